// kernels.h -- host-visible launch interface of the CUDA kernels (internal; the public boundary is include/mio_cuda.h)
#pragma once

#include "device_common.cuh"

namespace miocu
{

// launch shapes (tunable at build time; see tools/build_variants.py and profiles/ for the measurements behind them)
#ifndef MIO_INTEGRATE_THREADS
#define MIO_INTEGRATE_THREADS 256
#endif
#ifndef MIO_INTEGRATE_MINBLOCKS
#define MIO_INTEGRATE_MINBLOCKS 1
#endif
#ifndef MIO_EDGE_THREADS
#define MIO_EDGE_THREADS 256
#endif
#ifndef MIO_DEPART_MINBLOCKS
#define MIO_DEPART_MINBLOCKS 1
#endif
#ifndef MIO_RETURN_MINBLOCKS
#define MIO_RETURN_MINBLOCKS 1
#endif
#ifndef MIO_INTEGRATE_ITEM_PASSES
#define MIO_INTEGRATE_ITEM_PASSES 1024 /* whole runs unless there are too few runs to keep every CTA busy */
#endif
#ifndef MIO_EXACT_MINBLOCKS_FACT
#define MIO_EXACT_MINBLOCKS_FACT 16
#endif
#ifndef MIO_EXACT_MINBLOCKS
#define MIO_EXACT_MINBLOCKS 8
#endif
constexpr int INTEGRATE_THREADS = MIO_INTEGRATE_THREADS;
constexpr int EDGE_THREADS      = MIO_EDGE_THREADS;
// return_exact_blocks_kernel: T warps (in-order test / fix-up of a chunk of edges), as many R warps (returning
// populations of the next chunk), S warps (ordered sum into the source node)
#ifndef MIO_EXB_P1_WARPS
#define MIO_EXB_P1_WARPS 5
#endif
constexpr int EXB_P1_WARPS = MIO_EXB_P1_WARPS;
constexpr int EXB_P2_WARPS = 2;
constexpr int EXB_THREADS  = (2 * EXB_P1_WARPS + EXB_P2_WARPS) * 32;

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
    // written only by the node phase that precedes a departure (null otherwise):
    double* fac_out; // [R][N][W] get_mobility_factors of the new snapshot (InfectedSymptoms entry; the others are 1)
    double* u_out;   // [R][N][n] emigration base vector: snapshot * (1 - mobility damping) * factor * nondetection,
                     //           masked to the mobile compartments; mobile population of edge e = coeff_e * u_from
    const double* umask; // [n]
    double t_end;
    int integrator;
    double abs_tol, rel_tol, dt_min, dt_max;
    TrajectoryBuffers traj;
    int num_runs;
    int* work_counter; // [1] device scratch: the persistent node-phase kernel hands out its work items through it
    double* flows;      // [R][N][W][16] or null: the nodes are mio::FlowSimulation, cumulative flows (15 per lane + pad)
    double* traj_flows; // [R][N][capacity][G*15] or null: the flow series (with the trajectory buffers)
    double* aux_out;   // [R][N][W][4] or null: the launch that follows a departure writes the return coefficients of
                       // every node as a destination (by-product of its first right-hand side)
};

// Graph in the two forms the edge phase uses.
struct GraphCsr {
    // aggregated (SpMV) form
    const int* order_in;  // [N] nodes by descending in-degree  (departure kernel processing order)
    const int* order_out; // [N] nodes by descending out-degree (return kernel processing order)
    const int* in_ptr;    // [N+1] incoming edges sorted by (dst, src)
    const int* in_split;  // [N] incoming edges with src < dst: they precede dst's outgoing edges in reference order
    const int* in_src;    // [E]
    const double* w_in;   // factorized: [E][G] scalar per (edge, age); dense: [E][n]; incoming-CSR order
    const int* out_ptr;   // [N+1] outgoing edges = reference edge order
    const int* out_dst;   // [E]
    const double* w_out;  // [E][G][2] coefficient of the Susceptible and InfectedSevere compartment per age
    const double* s_out;  // [N][n] sum of the coefficient vectors of a node's outgoing edges (edge order)
    const double* umask;  // [n] 1/0 mobile compartment mask (all ones in dense mode)
    int factorized;       // coeff[e][i][c] == w[e][i] * umask[i][c] holds exactly for every edge
    int num_mob;          // number of mobile compartments (union over the age groups of the mask)
    int mob_comp[NC];     // their indices, padded by repeating a non-mobile compartment (whose u is zero)
    // literal form (exact replay of the sequential edge loop)
    const int* edge_start; // [E]
    const int* edge_end;   // [E]
    const double* coeff;   // [E][n]
    const double* w_edge;  // factorized graphs: [E][W] weight per (edge, lane), coeff[e][l][c] == w_edge[e][l] * umask[l][c]
};

struct EdgeArgs {
    Geometry geo;
    DeviceState st;
    GraphCsr csr;
    const double* snap;  // [R][N][n] m_last_state of the departure step
    const double* fac;   // [R][N][G] mobility factors of snap
    const double* u;     // [R][N][n] emigration base vectors of the departure step
    const double* y_in;  // [R][N][n] live state before this edge phase
    double* y_out;       // [R][N][n] live state after this edge phase
    double* vis;         // [R][N][n] visitors held by a node between departure and return (sum over its incoming edges)
    double* aux;         // [R][N][G][4] return coefficients of a node as destination (written by the node phase that
                         // follows the departure when aux_by_node_phase is set, else by the departure kernel)
    int aux_by_node_phase;
    int* run_flags;      // [R] return: set to 1 when an underflow fix-up is needed (exact replay required)
    long long* stats;    // [0] += number of (run, step) replays done by the exact kernel
    int* run_list;       // [R] scratch: the runs flagged for the exact replay of this step
    int* list_count;     // [1] scratch: length of run_list
    double t;            // time of this edge phase (m_t after the node phase)
    double t_dep;        // return only: departure time of the returning commuters
    double dt;           // dt of this graph step
    int num_runs;
    // schedule of the block-parallel literal replay (return_exact_blocks_kernel), run-invariant: chunk s = up to xepp
    // consecutive edges of ONE source node with distinct destinations; xslot[s][q] = {edge or -1, destination}
    const int* xslot;    // [xchunks][xepp][2] or null
    const int* xsrc;     // [xchunks][2] {source node, number of edges | hazard << 16} of a chunk; hazard: the source of
                         // the previous chunk is a destination of this one
    int xchunks, xepp;
    double* xdamp;       // [R][n] scratch: mobility damping of a flagged run at the departure time
    double* edge_m;      // [R][E][n] or null: mobile population of every edge (MobilityEdge::m_mobile_population's last
                         // value) written by the literal replay, for get_mobility_results
};

// MobilityEdge::add_mobility_result_time_point (metapopulation_mobility_instant.h:421-445) for all edges of all runs
struct MobilityResultArgs {
    int n, n_ref, E, K;        // device / reference state size, edges, index groups (saved_compartment_indices)
    const int* group_ptr;      // [K+1]
    const int* group_dev_idx;  // device index of every saved reference index, group after group
    const int* ref_to_dev;     // [n_ref] device index of reference element j (the total is summed in reference order)
    const unsigned char* edge_active; // [E] edge has a positive coefficient (edges without never carry commuters)
    const double* edge_m;      // [R][E][n]
    double* values;            // [R][E][capacity][K+1]
    int capacity, slot, num_runs;
};
// edge_m[r][e] = coeff_e * u[r][from(e)]: the mobile population of every edge right after a departure
cudaError_t launch_edge_mobile_depart(const EdgeArgs& a, cudaStream_t stream);
cudaError_t launch_mobility_results(const MobilityResultArgs& a, cudaStream_t stream);

cudaError_t launch_integrate(const IntegrateArgs& a, cudaStream_t stream);
cudaError_t launch_depart(const EdgeArgs& a, cudaStream_t stream);
cudaError_t launch_return(const EdgeArgs& a, cudaStream_t stream);
// literal in-order replay of the return phase for runs with run_flags[r] != 0 (sequential edge semantics,
// metapopulation_mobility_instant.h:569-603 incl. the underflow fix-up)
cudaError_t launch_return_exact(const EdgeArgs& a, cudaStream_t stream);
cudaError_t launch_record_daily(const Geometry& geo, const double* y, double* daily, int day_index, int max_days,
                                int num_runs, cudaStream_t stream);

// layered models (osecirvvs): device state order [LAY][G][NC] <-> reference order age*C + compartment
cudaError_t launch_permute_state(const Geometry& geo, const double* src, double* dst, size_t num_run_nodes,
                                 int to_device, cudaStream_t stream);

// ---- ensemble reduction (reduce.cu) -------------------------------------------------------------------------------
constexpr int SEL_MAX_P    = 8;
constexpr int SEL_MAX_SEG  = 8;
constexpr int SEL_BITS     = 11;
constexpr int SEL_BINS     = 1 << SEL_BITS;
#ifndef MIO_SELECT_THREADS
#define MIO_SELECT_THREADS 512
#endif
constexpr int SEL_THREADS  = MIO_SELECT_THREADS;

struct SelectArgs {
    const double* seg_ptr[SEL_MAX_SEG]; // segment s holds rows [col][seg_stride[s]], seg_count[s] valid runs each
    long long seg_stride[SEL_MAX_SEG];
    long long seg_count[SEL_MAX_SEG];
    int num_segments;
    int num_p;
    long long rank[SEL_MAX_P]; // zero based order statistic, size_t(num_runs * p)
    long long num_columns;
    long long total;           // sum of seg_count
    double* out_select;        // [num_p][num_columns] or null
    double* out_mean;          // [num_columns] or null
};


// y[R][N][n] of the current integer day -> columns of the ensemble matrix (per-node part and node-summed part)
// G, C: the reference's age groups and compartments per age group; lay: 1 (device order == reference order) or 3
cudaError_t launch_record_ensemble(const double* y, double* ens, int N, int G, int C, int lay, int num_runs, int day,
                                   int num_days, int age_sum, long long run_offset, long long stride,
                                   cudaStream_t stream);
cudaError_t launch_select(const SelectArgs& a, cudaStream_t stream);

} // namespace miocu
