// kernels.h -- host-visible launch interface of the CUDA kernels (internal; the public boundary is include/mio_cuda.h)
#pragma once

#include "device_common.cuh"

namespace miocu
{

constexpr int INTEGRATE_THREADS = 256;
constexpr int EDGE_THREADS      = 256;

struct TrajectoryBuffers {
    int capacity;   // 0 = off
    int* npts;      // [R][N] points recorded so far (may exceed capacity; only `capacity` are stored)
    double* times;  // [R][N][capacity]
    double* values; // [R][N][capacity][n]
};

struct IntegrateArgs {
    Geometry geo;
    DeviceState st;
    const double* y_in;
    double* y_out;
    double* fac_out; // [R][N][G] mobility factor of the InfectedSymptoms compartment of the new snapshot (may be null)
    double t_end;
    int integrator;
    double abs_tol, rel_tol, dt_min, dt_max;
    TrajectoryBuffers traj;
    int num_runs;
};

// incidence lists of the graph: for every node the incident edges in global (reference) edge order
struct Incidence {
    const int* node_order; // [N] processing order of the edge kernels (heaviest nodes first)
    const int* inc_ptr;    // [N+1]
    const int* inc_edge;   // [2E] edge index
    const int* inc_other;  // [2E] the other node
    const int* inc_out;    // [2E] 1: this node is the edge's start (outgoing), 0: incoming
    const int* edge_start; // [E]
    const int* edge_end;   // [E]
    const double* coeff;   // [E][n]
};

struct EdgeArgs {
    Geometry geo;
    DeviceState st;
    Incidence inc;
    const double* snap;     // [R][N][n] m_last_state of the departure step
    const double* fac;      // [R][N][G] mobility factors of snap
    const double* y_in;     // [R][N][n] live state before this edge phase
    double* y_out;          // [R][N][n] live state after this edge phase
    const double* v0;       // return only: [R][N][n] state right after the departure exchange (pop of the Euler step)
    double* aux;            // [R][N][G][2] departure: written (lambda_i, critAdj_i of v0); return: read
    int* run_flags;         // [R] return only: set to 1 when an underflow fix-up is needed (exact pass required)
    long long* stats;       // [0] += number of (run, step) replays done by the exact kernel
    double t;               // time of this edge phase (m_t after the node phase)
    double t_dep;           // return only: departure time of the returning commuters
    double dt;              // dt of this graph step
    int num_runs;
    double* daily_out;      // optional [R][N][max_days][n]: record y_out as day `day_index`
    int day_index, max_days;
};

cudaError_t launch_integrate(const IntegrateArgs& a, cudaStream_t stream);
cudaError_t launch_depart(const EdgeArgs& a, cudaStream_t stream);
cudaError_t launch_return(const EdgeArgs& a, cudaStream_t stream);
// literal in-order replay of the return phase for runs with run_flags[r] != 0 (sequential edge semantics,
// metapopulation_mobility_instant.h:569-603 incl. the underflow fix-up)
cudaError_t launch_return_exact(const EdgeArgs& a, cudaStream_t stream);
cudaError_t launch_record_daily(const Geometry& geo, const double* y, double* daily, int day_index, int max_days,
                                int num_runs, cudaStream_t stream);

} // namespace miocu
