// edges.cu -- edge phase of GraphSimulation::advance on B200: MobilityEdge::apply_mobility for all edges of all runs.
//
// Replaces cpp/memilio/mobility/metapopulation_mobility_instant.h:546-623 (apply_mobility), :460-471
// (calculate_mobility_returns: one explicit Euler step of the destination node's model with pop = destination state
// at departure), cpp/models/ode_secir/model.h:687-742 (get_mobility_factors, test_commuters) and the sequential
// in-place edge loop of cpp/memilio/mobility/graph_simulation.h:115-118.
//
// Formulation: the reference walks the edges in (start,end) order and updates both end nodes in place. Every
// contribution of an edge depends only on per-node snapshots (m_last_state of the source, post-departure state of the
// destination) and is linear in the source's snapshot, so the loop is re-expressed as two sparse matrix-vector
// gathers per node (visitors over the incoming-edge CSR at departure; destination coefficients over the outgoing-edge
// CSR at return) plus one Euler step per node, see depart_kernel / return_kernel. No atomics, deterministic; per node
// the blocks are combined in the reference's order (sources < k, own outgoing edges, sources > k). The traveler state
// of an edge is not stored: per destination node the summed visitor vectors are kept instead (2n doubles per node
// rather than n per edge). The one order dependent piece, the underflow fix-up of returns (:579-596), is detected
// (run_flags) and such runs are replayed for that step by return_exact_kernel with the literal sequential semantics.
#include "device_common.cuh"
#include "kernels.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace miocu
{

constexpr int AUX = 4; // per (run,node,lane): season*p_i*lambda_i (*ReducExposed), rISev_ICr, rISev_D, pad

template <int G, int LAY>
struct EdgeLane {
    Group<G, LAY> g;
    int run, node;
    bool valid;
    size_t rn;
};

__device__ __forceinline__ void load10(const double* p, double (&v)[NC])
{
    const double2* s = reinterpret_cast<const double2*>(p);
#pragma unroll
    for (int i = 0; i < NC / 2; ++i) {
        double2 x    = s[i];
        v[2 * i]     = x.x;
        v[2 * i + 1] = x.y;
    }
}
__device__ __forceinline__ void store10(double* p, const double (&v)[NC])
{
    double2* d = reinterpret_cast<double2*>(p);
#pragma unroll
    for (int i = 0; i < NC / 2; ++i) {
        d[i] = make_double2(v[2 * i], v[2 * i + 1]);
    }
}

// commuter_nondetection of a parameter set at time t (ode_secir/model.h:720-725)
__device__ __forceinline__ double nondetection_at(const double* pset, double t)
{
    return (t >= pset[hSTART_DETECTION] && t < pset[hEND_DETECTION]) ? pset[hNONDETECTION] : 1.0;
}

// mobile population of one edge for this lane's age group at departure time:
// m = m_last_state(from) * coeff(t) * mobility_factors(from), then test_commuters scaling (instant.h:608-615)
template <bool DAMPED>
__device__ __forceinline__ bool mobile_at_departure(const double* coeff_e, const double (&damp)[NC],
                                                    const double (&snap)[NC], double fac, double nd,
                                                    double (&m)[NC], double (&m_untested)[2])
{
    double c[NC];
    load10(coeff_e, c);
    bool pos = false;
#pragma unroll
    for (int i = 0; i < NC; ++i) {
        if (DAMPED) {
            c[i] = __dsub_rn(c[i], __dmul_rn(damp[i], c[i]));
        }
        pos  = pos || (c[i] > 0.0);
        m[i] = snap[i] * c[i];
    }
    m[cISy] *= fac;
    m_untested[0] = m[cINS];
    m_untested[1] = m[cISy];
    if (nd != 1.0) {
        m[cISy] *= nd;
        m[cINS] *= nd;
    }
    return pos;
}

// y1 = m + dt * f(pop = v0(dest), y = m): explicit Euler step of the destination's model on the traveler vector.
// slam = season*p_i*lambda_i(dest); the linear rates are the destination's.
template <int G, int LAY>
__device__ __forceinline__ void euler_return(const Group<G, LAY>& g, const LaneCoef& c, double slam, double rISev_ICr,
                                             double rISev_D, double dt, double (&m)[NC])
{
    double k[NC];
    flows_to_dydt(m[cS] * slam, c.rE_INS * m[cE], c.rINS_ISy * m[cINS], c.rINS_R * m[cINS], c.rINS_ISy * m[cINSC],
                  c.rINS_R * m[cINSC], c.rISy_ISev * m[cISy], c.rISy_R * m[cISy], c.rISy_ISev * m[cISyC],
                  c.rISy_R * m[cISyC], rISev_ICr * m[cISev], c.rISev_R * m[cISev], rISev_D * m[cISev],
                  c.rICr_D * m[cICr], c.rICr_R * m[cICr], k);
    route_recoveries<G, LAY>(g, k);
#pragma unroll
    for (int i = 0; i < NC; ++i) {
        m[i] = m[i] + dt * k[i];
    }
}

template <int G, int LAY>
__device__ __forceinline__ void edge_lane_setup(const Geometry& geo, const int* order, int num_runs,
                                                EdgeLane<G, LAY>& L, bool& cta_active)
{
    constexpr int W     = G * LAY;
    constexpr int NPW   = 32 / W;
    constexpr int WARPS = EDGE_THREADS / 32;
    const int chunks    = (geo.N + NPW * WARPS - 1) / (NPW * WARPS);
    L.run               = blockIdx.x / chunks;
    const int chunk     = blockIdx.x - L.run * chunks;
    cta_active          = L.run < num_runs;
    const int warp      = threadIdx.x >> 5;
    const int item      = (chunk * WARPS + warp) * NPW + L.g.lane / W;
    L.valid             = cta_active && !L.g.spare() && item < geo.N;
    L.node              = L.valid ? order[item] : 0;
    L.rn                = (size_t)(cta_active ? L.run : 0) * geo.N + L.node;
}

__device__ __forceinline__ void wait_mbar(uint64_t* mbar, uint32_t parity)
{
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n"
                     ".reg .pred p;\n"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
                     "selp.u32 %0, 1, 0, p;\n"
                     "}\n"
                     : "=r"(done)
                     : "r"(smem_u32(mbar)), "r"(parity)
                     : "memory");
    }
}

__device__ __forceinline__ int warp_max(int v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    }
    return v;
}

// V += coeff_e * u_src over the incoming edges [beg, beg+cnt) of this lane's node (SpMV row, edge order)
template <bool FACT, int W>
__device__ __forceinline__ void gather_visitors(const EdgeArgs& a, const Geometry& geo, int run, int gl, int beg,
                                                int cnt, int max_cnt, double (&V)[NC])
{
    if constexpr (FACT) {
        // batches of edges: the source indices and weights first, then the sources' records, so that the
        // index -> record dependency is paid once per batch and BATCH records are in flight per lane
        constexpr int BATCH = 4;
        for (int it0 = 0; it0 < max_cnt; it0 += BATCH) {
            int src[BATCH];
            double w[BATCH];
#pragma unroll
            for (int j = 0; j < BATCH; ++j) {
                const bool act = it0 + j < cnt;
                const int pos  = beg + (act ? it0 + j : 0);
                src[j]         = __ldg(a.csr.in_src + pos);
                w[j]           = act ? __ldg(a.csr.w_in + (size_t)pos * W + gl) : 0.0;
            }
            double us[BATCH][NC];
#pragma unroll
            for (int j = 0; j < BATCH; ++j) {
                load10(a.u + ((size_t)run * geo.N + src[j]) * geo.n + gl * NC, us[j]);
            }
#pragma unroll
            for (int j = 0; j < BATCH; ++j) {
#pragma unroll
                for (int i = 0; i < NC; ++i) {
                    V[i] += w[j] * us[j][i];
                }
            }
        }
    }
    else {
        for (int it = 0; it < max_cnt; ++it) {
            const bool act = it < cnt;
            const int pos  = beg + (act ? it : 0);
            const int src  = a.csr.in_src[pos];
            double us[NC];
            load10(a.u + ((size_t)run * geo.N + src) * geo.n + gl * NC, us);
            double w[NC];
            load10(a.csr.w_in + (size_t)pos * geo.n + gl * NC, w);
            if (act) {
#pragma unroll
                for (int i = 0; i < NC; ++i) {
                    V[i] += w[i] * us[i];
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Departure (all edges, one launch). With m_e = coeff_e * u_from (the mobile population of edge e) the in-place edge
// loop leaves node k at  y_k + sum_{e into k} m_e - sum_{e out of k} m_e. The visitors are gathered per destination
// (SpMV over the incoming-edge CSR; one vector V per node, see the note on the order at return_kernel) and kept for
// the return phase; the emigrants are s_out_k * u_k.
// Then the return coefficients of k as a destination (lambda_i, ICU adjusted rates) from v0 = y_out at time t.
// Everything of the departure after the visitor gather: test_commuters bookkeeping, the in-place node update, and the
// return coefficients of the node as a destination.
// NODE_AUX: the node phase that follows provides the return coefficients (a.aux_by_node_phase, known at compile time:
// the right-hand side below is then not part of the kernel at all, which halves its registers)
template <int G, int LAY, bool NODE_AUX = false>
__device__ __forceinline__ void depart_node_update(const EdgeArgs& a, const Geometry& geo, const EdgeLane<G, LAY>& L,
                                                   const double* smem, const double (&V)[NC],
                                                   const double* cm = nullptr)
{
    constexpr int W        = G * LAY;
    const Group<G, LAY>& g = L.g;
    const int ps_own       = a.st.node_pset ? a.st.node_pset[L.node] : 0;
    const double* pset     = smem + ps_own * geo.pset_stride;
    // a plain mio::Simulation<osecirvvs::Model> node has no test_commuters hook (instant.h:537-541)
    const bool hooks = !(LAY == VVS_LAY && geo.sim_kind != 1);
    double y[NC], uk[NC], so[NC];
    load10(a.y_in + L.rn * geo.n + g.gl * NC, y);
    load10(a.u + L.rn * geo.n + g.gl * NC, uk);
    load10(a.csr.s_out + (size_t)L.node * geo.n + g.gl * NC, so);
    // test_commuters (ode_secir/model.h:717-742, ode_secirvvs/model.h:887-940): detected commuters stay, in the
    // *Confirmed compartments (of their own immunity layer)
    const double nd = hooks ? nondetection_at(pset, a.t) : 1.0;
    if (nd != 1.0) {
        // u holds the tested (travelling) population; the untested one is rebuilt from the snapshot (= y here)
        const double* mob = geo.mob_max_pts > 0 ? smem + geo.mob_off : nullptr;
        double damp[NC];
        mobility_damping_eval((mob != nullptr && mob[0] > 0.0) ? mob : nullptr, geo.mob_max_pts, geo.n, g.gl, a.t,
                              damp);
        const double fk       = a.fac[L.rn * W + g.gl];
        const double lost_ins = so[cINS] * (y[cINS] * (1.0 - damp[cINS]) * a.csr.umask[g.gl * NC + cINS]);
        const double lost_isy = so[cISy] * (y[cISy] * (1.0 - damp[cISy]) * fk * a.csr.umask[g.gl * NC + cISy]);
        y[cINS] -= lost_ins * (1 - nd);
        y[cINSC] += lost_ins * (1 - nd);
        y[cISy] -= lost_isy * (1 - nd);
        y[cISyC] += lost_isy * (1 - nd);
    }
#pragma unroll
    for (int i = 0; i < NC; ++i) {
        y[i] = (y[i] + V[i]) - so[i] * uk[i];
    }
    if (L.valid) {
        store10(a.y_out + L.rn * geo.n + g.gl * NC, y);
        store10(a.vis + L.rn * geo.n + g.gl * NC, V);
    }
    // return coefficients of this node as destination: get_flows with pop = v0 (ode_secir/model.h:90-153,
    // ode_secirvvs/model.h:199-264). The node phase that follows evaluates exactly this right-hand side (same state,
    // same time) as its first stage and writes them then (aux_by_node_phase).
    if constexpr (!NODE_AUX) if (!a.aux_by_node_phase) {
        NodeRhs<G, LAY> rhs;
        rhs.age_rec = pset + PHDR + g.gl * AGE_REC;
        rhs.c.load(rhs.age_rec);
        rhs.p_eff         = rhs.c.rP;
        const int ci      = a.st.node_ctable ? a.st.node_ctable[L.node] : 0;
        rhs.ctable        = smem + geo.ct_off + ci * geo.ctable_stride;
        rhs.start_day     = pset[hSTART_DAY];
        rhs.seasonality   = pset[hSEASONALITY];
        rhs.icu_cap       = a.st.node_local[L.rn * 2 + 0];
        rhs.tnt_cap       = a.st.node_local[L.rn * 2 + 1];
        rhs.tnt_cap_hi    = rhs.tnt_cap * pset[hTNT_MAX_RISK];
        rhs.tnt_cap_hi_ns = rhs.tnt_cap * pset[hTNT_MAX_RISK_NOSYMP];
        rhs.icu_lo        = 0.90 * rhs.icu_cap;
        rhs.valid_lo      = DBL_MAX;
        rhs.valid_hi      = -DBL_MAX;
        rhs.ct_max_loc    = geo.ct_max_loc;
        rhs.ct_max_pts    = geo.ct_max_pts;
        if (cm != nullptr) {
            // C(t) of the run's contact table was evaluated once for the CTA (cm[G*G], shared memory)
#pragma unroll
            for (int q = 0; q < NodeRhs<G, LAY>::Q; ++q) {
                const int j = LAY == 1 ? q : q * LAY + g.layer;
                rhs.crow[q] = j < G ? cm[g.age * G + j] : 0.0;
            }
            rhs.valid_lo = a.t;
            rhs.valid_hi = a.t;
        }
        rhs.contact_row(g, a.t);
        const double season = rhs.season_at(a.t);
        double tnt, icu;
        rhs.node_loads(g, y, tnt, icu);
        const double lam = rhs.infection_pressure(g, y, tnt);
        double rISev_ICr, rISev_D;
        rhs.icu_rates(icu, rISev_ICr, rISev_D);
        const double slam = LAY == 1 ? (season * rhs.c.rP) * lam : rhs.c.rREDEXP * ((season * rhs.c.rP) * lam);
        if (L.valid) {
            double* ax = a.aux + (L.rn * W + g.gl) * AUX;
            reinterpret_cast<double2*>(ax)[0] = make_double2(slam, rISev_ICr);
            reinterpret_cast<double2*>(ax)[1] = make_double2(rISev_D, 0.0);
        }
    }
}

template <int G, int LAY, bool FACT>
__global__ void __launch_bounds__(EDGE_THREADS, MIO_DEPART_MINBLOCKS) depart_kernel(EdgeArgs a)
{
    extern __shared__ __align__(16) double smem[];
    __shared__ __align__(8) uint64_t mbar;
    constexpr int W     = G * LAY;
    const Geometry& geo = a.geo;
    EdgeLane<G, LAY> L;
    bool cta_active;
    edge_lane_setup<G, LAY>(geo, a.csr.order_in, a.num_runs, L, cta_active);
    if (!cta_active) {
        return;
    }
    stage_run_block(smem, a.st.run_block + (size_t)L.run * geo.block_stride, geo.block_stride * 8, &mbar);
    const Group<G, LAY>& g = L.g;
    const int beg          = L.valid ? a.csr.in_ptr[L.node] : 0;
    const int cnt     = L.valid ? a.csr.in_ptr[L.node + 1] - beg : 0;
    double V[NC];
#pragma unroll
    for (int i = 0; i < NC; ++i) {
        V[i] = 0.0;
    }
    gather_visitors<FACT, W>(a, geo, L.run, g.gl, beg, cnt, warp_max(cnt), V);
    depart_node_update<G, LAY>(a, geo, L, smem, V);
}

// ---------------------------------------------------------------------------------------------------------------
// Departure with the run's emigration base vectors staged in shared memory (factorized coefficients only).
// The gather of depart_kernel reads E*n doubles per run through L2 (48 GB per launch at R = 10^4) and is bound by L2
// bandwidth; here one CTA owns one run, copies the M mobile compartments of u[run] (N*G*M doubles, 96 KB for the
// 400-county graph) into shared memory once, and every incoming edge is then a shared-memory read.
#ifndef MIO_DEPART_SMEM_THREADS
#define MIO_DEPART_SMEM_THREADS 512
#endif
#ifndef MIO_DEPART_SMEM_MINBLOCKS
#define MIO_DEPART_SMEM_MINBLOCKS 1
#endif
#ifndef MIO_DEPART_BATCH
#define MIO_DEPART_BATCH 8
#endif
#ifndef MIO_DEPART_SMEM_MINBLOCKS_AUX /* CTAs per SM of the variant without the right-hand side (NODE_AUX) */
#define MIO_DEPART_SMEM_MINBLOCKS_AUX 1
#endif
constexpr int DEPART_SMEM_THREADS = MIO_DEPART_SMEM_THREADS;

// NR runs per CTA (1 or 2): the edge list and the weights are the same for every run, so a CTA that stages the
// emigration vectors of two consecutive runs fetches them once for both (the load/store unit, not HBM, limits this
// kernel; half of its wavefronts were these run-invariant global loads).
template <int G, int LAY, int M, int NR, bool NODE_AUX>
__global__ void __launch_bounds__(DEPART_SMEM_THREADS, NODE_AUX ? MIO_DEPART_SMEM_MINBLOCKS_AUX : MIO_DEPART_SMEM_MINBLOCKS)
    depart_smem_kernel(EdgeArgs a)
{
    extern __shared__ __align__(16) double smem[];
    __shared__ __align__(8) uint64_t mbar;
    const Geometry& geo = a.geo;
    constexpr int W     = G * LAY;
    constexpr int NPW   = 32 / W;
    constexpr int WARPS = DEPART_SMEM_THREADS / 32;
    const int run0      = blockIdx.x * NR;
    if (run0 >= a.num_runs) {
        return;
    }
    const int nr = min(NR, a.num_runs - run0); // runs of this CTA (the blocks and vectors of consecutive runs are
                                               // contiguous in global memory)
    stage_run_block(smem, a.st.run_block + (size_t)run0 * geo.block_stride, nr * geo.block_stride * 8, &mbar);
    double* us = smem + NR * geo.block_stride;
    const int run_recs = geo.N * W; // (node, lane) records of one run
    {
        // one (node, age) record per thread and iteration: only its M mobile compartments are read (M loads in
        // flight per thread, no register shuffling of the other compartments)
        const double* ug = a.u + (size_t)run0 * geo.N * geo.n;
        const int total  = nr * run_recs;
        int comp[M];
#pragma unroll
        for (int m = 0; m < M; ++m) {
            comp[m] = a.csr.mob_comp[m];
        }
#pragma unroll 2
        for (int rec = threadIdx.x; rec < total; rec += DEPART_SMEM_THREADS) {
            double x[M];
#pragma unroll
            for (int m = 0; m < M; ++m) {
                x[m] = __ldg(ug + (size_t)rec * NC + comp[m]);
            }
#pragma unroll
            for (int m = 0; m < M; ++m) {
                us[rec * M + m] = x[m];
            }
        }
    }
    // one contact table per run: C(t) is evaluated once for the CTA by its first warp (one entry per lane)
    __shared__ double cm_s[(G * G + 1) & ~1];
    __shared__ int cm_ok;
    if (threadIdx.x < 32) {
        double lo = 1.0, hi = 0.0;
        // only the right-hand side of the node update needs it (NR == 1: with NR > 1 the node phase provides the
        // return coefficients, see the launch)
        if (NR == 1 && !NODE_AUX && geo.num_ctables == 1 && !a.aux_by_node_phase) {
            warp_contact_matrix<G>(smem + geo.ct_off, geo.ct_max_loc, geo.ct_max_pts, a.t, cm_s, lo, hi);
        }
        if (threadIdx.x == 0) {
            cm_ok = lo <= hi ? 1 : 0;
        }
    }
    __syncthreads();
    const double* cm = cm_ok ? cm_s : nullptr;
    EdgeLane<G, LAY> L;
    const int warp   = threadIdx.x >> 5;
    const int passes = (geo.N + NPW * WARPS - 1) / (NPW * WARPS);
    for (int pass = 0; pass < passes; ++pass) {
        const int item   = (pass * WARPS + warp) * NPW + L.g.lane / W;
        const bool valid = !L.g.spare() && item < geo.N;
        L.node           = valid ? a.csr.order_in[item] : 0;
        const int gl     = L.g.gl;
        const int beg    = valid ? a.csr.in_ptr[L.node] : 0;
        const int cnt    = valid ? a.csr.in_ptr[L.node + 1] - beg : 0;
        double v[NR][M];
#pragma unroll
        for (int r = 0; r < NR; ++r) {
#pragma unroll
            for (int m = 0; m < M; ++m) {
                v[r][m] = 0.0;
            }
        }
        // the edge list and weights come from global memory (L2): the batch of BATCH edges after the current one is
        // fetched before the current batch's shared-memory reads and FMAs (software pipeline), so that the index ->
        // value dependency is not exposed once per batch
        constexpr int BATCH = MIO_DEPART_BATCH;
        const int mx        = warp_max(cnt);
        int src[BATCH], src_n[BATCH];
        double w[BATCH], w_n[BATCH];
        auto fetch = [&](int it0, int(&sr)[BATCH], double(&ww)[BATCH]) {
#pragma unroll
            for (int j = 0; j < BATCH; ++j) {
                const bool act = it0 + j < cnt;
                const int pos  = beg + (act ? it0 + j : 0);
                sr[j]          = __ldg(a.csr.in_src + pos);
                ww[j]          = act ? __ldg(a.csr.w_in + (size_t)pos * W + gl) : 0.0;
            }
        };
        if (mx > 0) {
            fetch(0, src_n, w_n);
        }
        for (int it0 = 0; it0 < mx; it0 += BATCH) {
#pragma unroll
            for (int j = 0; j < BATCH; ++j) {
                src[j] = src_n[j];
                w[j]   = w_n[j];
            }
            if (it0 + BATCH < mx) {
                fetch(it0 + BATCH, src_n, w_n);
            }
#pragma unroll
            for (int j = 0; j < BATCH; ++j) {
                // age groups that do not travel on this edge (weight 0: the sum would only gain +0.0) issue no
                // shared-memory reads
                if (w[j] != 0.0) {
                    const double* p = us + (src[j] * W + gl) * M;
#pragma unroll
                    for (int r = 0; r < NR; ++r) {
#pragma unroll
                        for (int m = 0; m < M; ++m) {
                            v[r][m] += w[j] * p[r * run_recs * M + m];
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            double V[NC];
#pragma unroll
            for (int i = 0; i < NC; ++i) {
                V[i] = 0.0;
            }
#pragma unroll
            for (int m = 0; m < M; ++m) {
                const int c = a.csr.mob_comp[m];
#pragma unroll
                for (int i = 0; i < NC; ++i) { // register arrays: select instead of a dynamic index
                    if (i == c) {
                        V[i] += v[r][m];
                    }
                }
            }
            L.run   = run0 + r;
            L.valid = valid && r < nr;
            L.rn    = (size_t)(r < nr ? run0 + r : run0) * geo.N + L.node;
            depart_node_update<G, LAY, NODE_AUX>(a, geo, L, smem + r * geo.block_stride, V, cm);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Return (all edges, one launch). calculate_mobility_returns is one explicit Euler step m + dt*A_dest*m of the
// destination's model, linear in the traveler vector m, so the sums over edges commute with it:
//   visitors leaving k :  (I + dt*A_k) V                                            (V from the departure gather)
//   own commuters back :  sum_e (I + dt*A_dst(e)) coeff_e*u_k = (I + dt*A0)(s_out*u_k) + dt*(destination dependent
//                         flows: S->E with sum_e coeff_e[S]*lambda_dst, ISev->ICr/D with the ICU adjusted rates)
// Order: the reference subtracts the visitors from sources < k, adds k's own commuters, then subtracts the visitors
// from sources > k. Without a fix-up the three blocks commute up to rounding, so all visitors of a node are kept as
// ONE vector V (half the traffic of two) and the node becomes (y - V') + own.
// The underflow fix-up (instant.h:579-596) triggers iff the node's running value goes negative at an incoming edge;
// the running value decreases along the incoming edges (returning visitor vectors are non-negative as long as
// dt*rate <= 1, checked) and the own commuters only add, so y - V' >= 0 implies that no fix-up happens anywhere in
// the sequence. The test y - V' < 0 is conservative (the own commuters might have covered the deficit before the
// later visitors left): such runs are merely replayed without need. Flagged runs are replayed literally by
// return_exact_kernel.
template <int G, int LAY, bool FACT>
__global__ void __launch_bounds__(EDGE_THREADS, MIO_RETURN_MINBLOCKS < 2 ? 2 : MIO_RETURN_MINBLOCKS)
    return_kernel(EdgeArgs a)
{
    extern __shared__ __align__(16) double smem[];
    __shared__ __align__(8) uint64_t mbar;
    constexpr int W     = G * LAY;
    const Geometry& geo = a.geo;
    EdgeLane<G, LAY> L;
    bool cta_active;
    edge_lane_setup<G, LAY>(geo, a.csr.order_out, a.num_runs, L, cta_active);
    if (!cta_active) {
        return;
    }
    stage_run_block(smem, a.st.run_block + (size_t)L.run * geo.block_stride, geo.block_stride * 8, &mbar);
    const Group<G, LAY>& g = L.g;
    const int ps_own       = a.st.node_pset ? a.st.node_pset[L.node] : 0;
    const double* pset     = smem + ps_own * geo.pset_stride;
    LaneCoef c_own;
    c_own.load(pset + PHDR + g.gl * AGE_REC);
    double slam_own, ricr_own, rd_own;
    {
        const double2* ax = reinterpret_cast<const double2*>(a.aux + (L.rn * W + g.gl) * AUX);
        double2 a0 = ax[0], a1 = ax[1];
        slam_own = a0.x;
        ricr_own = a0.y;
        rd_own   = a1.x;
    }
    double y[NC], V[NC];
    load10(a.y_in + L.rn * geo.n + g.gl * NC, y);
    bool underflow = false;
    // monotonicity guard: returning visitor vectors stay non-negative iff dt * (outflow rate) <= 1
    {
        const double rmax = fmax(fmax(fmax(slam_own, c_own.rE_INS), fmax(c_own.rINS_ISy + c_own.rINS_R,
                                                                         c_own.rISy_ISev + c_own.rISy_R)),
                                 fmax(ricr_own + c_own.rISev_R + rd_own, c_own.rICr_D + c_own.rICr_R));
        underflow = !(a.dt * rmax <= 1.0);
    }
    // the node's visitors return home (all of them: see the note on the order above)
    load10(a.vis + L.rn * geo.n + g.gl * NC, V);
    euler_return<G, LAY>(g, c_own, slam_own, ricr_own, rd_own, a.dt, V);
#pragma unroll
    for (int i = 0; i < NC; ++i) {
        y[i] -= V[i];
        underflow = underflow || (y[i] < 0.0);
    }
    // own commuters
    {
        double uk[NC], so[NC], q[NC];
        load10(a.u + L.rn * geo.n + g.gl * NC, uk);
        const int beg = L.valid ? a.csr.out_ptr[L.node] : 0;
        const int cnt = L.valid ? a.csr.out_ptr[L.node + 1] - beg : 0;
        const int mx  = warp_max(cnt);
        if (geo.num_psets == 1) {
            load10(a.csr.s_out + (size_t)L.node * geo.n + g.gl * NC, so);
            double a_slam = 0.0, a_icr = 0.0, a_d = 0.0;
            constexpr int BATCH = 4;
            for (int it0 = 0; it0 < mx; it0 += BATCH) {
                int dst[BATCH];
                double2 w[BATCH], a0[BATCH], a1[BATCH];
#pragma unroll
                for (int j = 0; j < BATCH; ++j) {
                    const bool act = it0 + j < cnt;
                    const int e    = beg + (act ? it0 + j : 0);
                    dst[j]         = __ldg(a.csr.out_dst + e);
                    w[j] = __ldg(reinterpret_cast<const double2*>(a.csr.w_out) + (size_t)e * W + g.gl);
                    if (!act) {
                        w[j] = make_double2(0.0, 0.0);
                    }
                }
#pragma unroll
                for (int j = 0; j < BATCH; ++j) {
                    const double2* ax =
                        reinterpret_cast<const double2*>(a.aux + (((size_t)L.run * geo.N + dst[j]) * W + g.gl) * AUX);
                    a0[j] = ax[0];
                    a1[j] = ax[1];
                }
#pragma unroll
                for (int j = 0; j < BATCH; ++j) {
                    a_slam += w[j].x * a0[j].x;
                    a_icr += w[j].y * a0[j].y;
                    a_d += w[j].y * a1[j].x;
                }
            }
#pragma unroll
            for (int i = 0; i < NC; ++i) {
                q[i] = so[i] * uk[i];
            }
            double k[NC];
            const LaneCoef& c = c_own;
            flows_to_dydt(a_slam * uk[cS], c.rE_INS * q[cE], c.rINS_ISy * q[cINS], c.rINS_R * q[cINS],
                          c.rINS_ISy * q[cINSC], c.rINS_R * q[cINSC], c.rISy_ISev * q[cISy], c.rISy_R * q[cISy],
                          c.rISy_ISev * q[cISyC], c.rISy_R * q[cISyC], a_icr * uk[cISev], c.rISev_R * q[cISev],
                          a_d * uk[cISev], c.rICr_D * q[cICr], c.rICr_R * q[cICr], k);
            route_recoveries<G, LAY>(g, k);
#pragma unroll
            for (int i = 0; i < NC; ++i) {
                y[i] += q[i] + a.dt * k[i];
            }
        }
        else {
            // parameter sets differ between nodes: Euler step per edge with the destination's own rates
            for (int it = 0; it < mx; ++it) {
                const bool act = it < cnt;
                const int e    = beg + (act ? it : 0);
                const int dst  = a.csr.out_dst[e];
                double m[NC];
                load10(a.csr.coeff + (size_t)e * geo.n + g.gl * NC, m);
#pragma unroll
                for (int i = 0; i < NC; ++i) {
                    m[i] *= uk[i];
                }
                LaneCoef c_o;
                c_o.load(smem + a.st.node_pset[dst] * geo.pset_stride + PHDR + g.gl * AGE_REC);
                const double2* ax =
                    reinterpret_cast<const double2*>(a.aux + (((size_t)L.run * geo.N + dst) * W + g.gl) * AUX);
                const double2 a0 = ax[0], a1 = ax[1];
                euler_return<G, LAY>(g, c_o, a0.x, a0.y, a1.x, a.dt, m);
                if (act) {
#pragma unroll
                    for (int i = 0; i < NC; ++i) {
                        y[i] += m[i];
                    }
                }
            }
        }
    }
    if (L.valid) {
        store10(a.y_out + L.rn * geo.n + g.gl * NC, y);
        if (underflow) {
            a.run_flags[L.run] = 1;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Return with the node records moved by TMA (one parameter set per run, i.e. the ensemble case). A lane's record is
// 80 contiguous bytes, so a warp-wide 16 byte load touches 30 different sectors and return_kernel is bound by the
// load/store unit, not by HBM. Here every warp owns a tile in shared memory: one elected lane issues a bulk copy per
// (node, array) -- 480 contiguous bytes each -- for the NPW nodes of the pass (state, both visitor vectors, emigration
// base vector); the lanes then read their records with conflict-free 16 byte shared-memory loads, and the new state
// leaves through bulk stores. Same arithmetic and order as return_kernel.
#ifndef MIO_RETURN_TMA_THREADS
#define MIO_RETURN_TMA_THREADS 256
#endif
#ifndef MIO_RETURN_TMA_MINBLOCKS
#define MIO_RETURN_TMA_MINBLOCKS 2
#endif
constexpr int RETURN_TMA_THREADS = MIO_RETURN_TMA_THREADS;
constexpr int RETURN_TMA_ARRAYS  = 3; // y, V, u

__device__ __forceinline__ void lds10(const double* p, double (&v)[NC])
{
    const double2* s = reinterpret_cast<const double2*>(p);
#pragma unroll
    for (int i = 0; i < NC / 2; ++i) {
        const double2 x = s[i];
        v[2 * i]        = x.x;
        v[2 * i + 1]    = x.y;
    }
}

template <int G, int LAY>
__global__ void __launch_bounds__(RETURN_TMA_THREADS, MIO_RETURN_TMA_MINBLOCKS) return_tma_kernel(EdgeArgs a)
{
    extern __shared__ __align__(16) double smem[];
    constexpr int WARPS = RETURN_TMA_THREADS / 32;
    __shared__ __align__(8) uint64_t mbar;         // parameter block
    __shared__ __align__(8) uint64_t wbar[WARPS];  // one per warp tile
    const Geometry& geo = a.geo;
    constexpr int W     = G * LAY;
    constexpr int NPW   = 32 / W;
    const int run       = blockIdx.x;
    if (run >= a.num_runs) {
        return;
    }
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        for (int w = 0; w < WARPS; ++w) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&wbar[w])));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int blk_bytes = geo.block_stride * 8;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(blk_bytes)
                     : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(smem)),
                     "l"(a.st.run_block + (size_t)run * geo.block_stride), "r"(blk_bytes), "r"(smem_u32(&mbar))
                     : "memory");
    }
    const int rec_bytes = geo.n * 8;                  // one node record
    const int tile_dbl  = NPW * geo.n;                // doubles of one array's tile
    double* tile        = smem + geo.block_stride + (size_t)warp * RETURN_TMA_ARRAYS * tile_dbl;
    const double* src[RETURN_TMA_ARRAYS] = {a.y_in, a.vis, a.u};
    const int passes = (geo.N + NPW * WARPS - 1) / (NPW * WARPS);
    EdgeLane<G, LAY> L;
    L.run                  = run;
    const Group<G, LAY>& g = L.g;
    bool underflow_any     = false;
    bool have_params       = false;
    LaneCoef c_own;
    for (int pass = 0; pass < passes; ++pass) {
        const int item0 = (pass * WARPS + warp) * NPW;
        if (item0 >= geo.N) { // uniform over the warp
            break;
        }
        const int cnt_nodes = min(NPW, geo.N - item0);
        if (lane == 0) {
            // the bulk stores of the previous pass have read the tile
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&wbar[warp])),
                         "r"(cnt_nodes * RETURN_TMA_ARRAYS * rec_bytes)
                         : "memory");
            for (int j = 0; j < cnt_nodes; ++j) {
                const size_t rn = (size_t)run * geo.N + a.csr.order_out[item0 + j];
#pragma unroll
                for (int q = 0; q < RETURN_TMA_ARRAYS; ++q) {
                    asm volatile(
                        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                            smem_u32(tile + q * tile_dbl + j * geo.n)),
                        "l"(src[q] + rn * geo.n), "r"(rec_bytes), "r"(smem_u32(&wbar[warp]))
                        : "memory");
                }
            }
        }
        const int item = item0 + g.lane / W;
        L.valid        = !g.spare() && item < geo.N;
        L.node         = L.valid ? a.csr.order_out[item] : 0;
        L.rn           = (size_t)run * geo.N + L.node;
        if (!have_params) {
            wait_mbar(&mbar, 0);
            c_own.load(smem + PHDR + g.gl * AGE_REC); // one parameter set per run
            have_params = true;
        }
        double slam_own, ricr_own, rd_own, so[NC];
        {
            const double2* ax = reinterpret_cast<const double2*>(a.aux + (L.rn * W + g.gl) * AUX);
            const double2 a0 = ax[0], a1 = ax[1];
            slam_own = a0.x;
            ricr_own = a0.y;
            rd_own   = a1.x;
        }
        load10(a.csr.s_out + (size_t)L.node * geo.n + g.gl * NC, so);
        // own commuters: destination dependent flows gathered over the outgoing edges (L2)
        double a_slam = 0.0, a_icr = 0.0, a_d = 0.0;
        {
            const int beg = L.valid ? a.csr.out_ptr[L.node] : 0;
            const int cnt = L.valid ? a.csr.out_ptr[L.node + 1] - beg : 0;
            const int mx  = warp_max(cnt);
            constexpr int BATCH = 4;
            for (int it0 = 0; it0 < mx; it0 += BATCH) {
                int dst[BATCH];
                double2 w[BATCH], a0[BATCH], a1[BATCH];
#pragma unroll
                for (int j = 0; j < BATCH; ++j) {
                    const bool act = it0 + j < cnt;
                    const int e    = beg + (act ? it0 + j : 0);
                    dst[j]         = __ldg(a.csr.out_dst + e);
                    w[j]           = __ldg(reinterpret_cast<const double2*>(a.csr.w_out) + (size_t)e * W + g.gl);
                    if (!act) {
                        w[j] = make_double2(0.0, 0.0);
                    }
                }
#pragma unroll
                for (int j = 0; j < BATCH; ++j) {
                    // lanes whose age group does not travel on this edge (both weights 0) issue no loads
                    a0[j] = make_double2(0.0, 0.0);
                    a1[j] = make_double2(0.0, 0.0);
                    if (w[j].x != 0.0 || w[j].y != 0.0) {
                        const double2* ax = reinterpret_cast<const double2*>(
                            a.aux + (((size_t)run * geo.N + dst[j]) * W + g.gl) * AUX);
                        a0[j] = ax[0];
                        a1[j] = ax[1];
                    }
                }
#pragma unroll
                for (int j = 0; j < BATCH; ++j) {
                    a_slam += w[j].x * a0[j].x;
                    a_icr += w[j].y * a0[j].y;
                    a_d += w[j].y * a1[j].x;
                }
            }
        }
        // the tile has arrived
        wait_mbar(&wbar[warp], pass & 1);
        const int slot_off = (g.spare() ? 0 : (g.lane / W)) * geo.n + g.gl * NC;
        double y[NC], V[NC], uk[NC];
        lds10(tile + 0 * tile_dbl + slot_off, y);
        lds10(tile + 1 * tile_dbl + slot_off, V);
        lds10(tile + 2 * tile_dbl + slot_off, uk);
        bool underflow = false;
        // monotonicity guard: returning visitor vectors stay non-negative iff dt * (outflow rate) <= 1
        {
            const double rmax = fmax(fmax(fmax(slam_own, c_own.rE_INS), fmax(c_own.rINS_ISy + c_own.rINS_R,
                                                                             c_own.rISy_ISev + c_own.rISy_R)),
                                     fmax(ricr_own + c_own.rISev_R + rd_own, c_own.rICr_D + c_own.rICr_R));
            underflow = !(a.dt * rmax <= 1.0);
        }
        // the node's visitors return home
        euler_return<G, LAY>(g, c_own, slam_own, ricr_own, rd_own, a.dt, V);
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            y[i] -= V[i];
            underflow = underflow || (y[i] < 0.0);
        }
        // own commuters
        {
            double q[NC], k[NC];
#pragma unroll
            for (int i = 0; i < NC; ++i) {
                q[i] = so[i] * uk[i];
            }
            const LaneCoef& c = c_own;
            flows_to_dydt(a_slam * uk[cS], c.rE_INS * q[cE], c.rINS_ISy * q[cINS], c.rINS_R * q[cINS],
                          c.rINS_ISy * q[cINSC], c.rINS_R * q[cINSC], c.rISy_ISev * q[cISy], c.rISy_R * q[cISy],
                          c.rISy_ISev * q[cISyC], c.rISy_R * q[cISyC], a_icr * uk[cISev], c.rISev_R * q[cISev],
                          a_d * uk[cISev], c.rICr_D * q[cICr], c.rICr_R * q[cICr], k);
            route_recoveries<G, LAY>(g, k);
#pragma unroll
            for (int i = 0; i < NC; ++i) {
                y[i] += q[i] + a.dt * k[i];
            }
        }
        if (L.valid) {
            store10(tile + slot_off, y); // shared memory: the state's tile is reused for the result
            underflow_any = underflow_any || underflow;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            for (int j = 0; j < cnt_nodes; ++j) {
                const size_t rn = (size_t)run * geo.N + a.csr.order_out[item0 + j];
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(a.y_out + rn * geo.n),
                             "r"(smem_u32(tile + j * geo.n)), "r"(rec_bytes)
                             : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        __syncwarp();
    }
    if (lane == 0) {
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    if (underflow_any) {
        a.run_flags[run] = 1;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Literal sequential replay of the return phase for flagged runs, two kernels:
//  collect_flagged_kernel  one CTA per run; a flagged run is appended to the run list, its flag cleared and its live
//                          state reset to the state before the return phase (y_out[run] = y_in[run]).
//  return_exact_kernel     one warp per NPW listed runs (group of W lanes = one run, lane = age group); the edges are
//                          walked in reference order and both end nodes are updated in place in y_out. The walk is a
//                          dependent chain per run, so it is software pipelined: edge indices are fetched two edges
//                          ahead, the edge's read-only operands (coefficients, destination coefficients) one edge
//                          ahead, and the source node's snapshot and running state stay in registers while consecutive
//                          edges share the source (edges are sorted by source).
template <int THREADS>
__global__ void __launch_bounds__(THREADS) collect_flagged_kernel(EdgeArgs a, int* __restrict__ run_list,
                                                                   int* __restrict__ list_count, int copy_state)
{
    const int run = blockIdx.x;
    if (run >= a.num_runs || a.run_flags[run] == 0) {
        return;
    }
    // (the block-parallel replay reads the pre-return state itself)
    const size_t total = copy_state ? (size_t)a.geo.N * a.geo.n / 2 : 0; // n is even
    const double2* src = reinterpret_cast<const double2*>(a.y_in + (size_t)run * a.geo.N * a.geo.n);
    double2* dst       = reinterpret_cast<double2*>(a.y_out + (size_t)run * a.geo.N * a.geo.n);
    for (size_t i = threadIdx.x; i < total; i += THREADS) {
        dst[i] = src[i];
    }
    if (!copy_state && a.xdamp != nullptr && a.geo.mob_max_pts > 0) {
        // mobility damping of the run at the departure time for the block-parallel replay (mobility_damping_eval)
        const double* mob = a.st.run_block + (size_t)run * a.geo.block_stride + a.geo.mob_off;
        const int K       = (int)mob[0];
        for (int j = threadIdx.x; j < a.geo.n && K > 0; j += THREADS) {
            const double* times = mob + 1;
            const double* D     = mob + 1 + a.geo.mob_max_pts + j;
            int k               = 0;
            while (k < K && !(a.t_dep < times[k])) {
                ++k;
            }
            const double* Dl = k > 0 ? D + (size_t)(k - 1) * a.geo.n : nullptr;
            const double* Dr = k < K ? D + (size_t)k * a.geo.n : Dl;
            const double xr  = k < K ? times[k] : DBL_MAX;
            a.xdamp[(size_t)run * a.geo.n + j] = smoother_cosine(a.t_dep, xr - 1.0, xr, Dl ? *Dl : 0.0, Dr ? *Dr : 0.0);
        }
    }
    __syncthreads(); // every thread has read the flag
    if (threadIdx.x == 0) {
        a.run_flags[run]                 = 0;
        run_list[atomicAdd(list_count, 1)] = run;
        atomicAdd((unsigned long long*)a.stats, 1ull);
    }
}

// FACT: the graph's coefficients factorize exactly (coeff[e][l][c] == w[e][l] * mask[l][c], mask 0/1): one weight per
// (edge, lane) is fetched instead of ten coefficients
template <int G, int LAY, bool FACT>
struct ExactEdge {
    int from, to;
    double c[FACT ? 1 : NC]; // coefficient vector of the edge for this lane (FACT: the weight)
    double2 a0, a1;          // return coefficients of the destination (aux)
};

template <int G, int LAY, bool FACT>
__device__ __forceinline__ void exact_edge_load(const EdgeArgs& a, const Geometry& geo, int run, int gl, int e,
                                                int from, int to, ExactEdge<G, LAY, FACT>& d)
{
    constexpr int W = G * LAY;
    d.from          = from;
    d.to            = to;
    if constexpr (FACT) {
        d.c[0] = __ldg(a.csr.w_edge + (size_t)e * W + gl);
    }
    else {
        const double2* cp = reinterpret_cast<const double2*>(a.csr.coeff + (size_t)e * geo.n + gl * NC);
#pragma unroll
        for (int i = 0; i < NC / 2; ++i) {
            const double2 v = __ldg(cp + i);
            d.c[2 * i]      = v.x;
            d.c[2 * i + 1]  = v.y;
        }
    }
    const double2* ax = reinterpret_cast<const double2*>(a.aux + (((size_t)run * geo.N + to) * W + gl) * AUX);
    d.a0              = ax[0];
    d.a1              = ax[1];
}

template <int G, int LAY, bool FACT>
__global__ void __launch_bounds__(32, FACT ? MIO_EXACT_MINBLOCKS_FACT : MIO_EXACT_MINBLOCKS) return_exact_kernel(EdgeArgs a, const int* __restrict__ run_list,
                                                          const int* __restrict__ list_count)
{
    constexpr int W     = G * LAY;
    constexpr int NPW   = 32 / W;
    const Geometry& geo = a.geo;
    // lane constants live in shared memory ([k][32], conflict free) rather than in registers: destination rate
    // coefficients (one parameter set per run), mobility damping, the held source node's snapshot
    __shared__ double ct_s[16 * 32];
    __shared__ double damp_s[NC * 32];
    __shared__ double sf_s[NC * 32];
    const int count     = *list_count;
    const int base      = blockIdx.x * NPW;
    if (base >= count) {
        return;
    }
    Group<G, LAY> g;
    const int lane   = threadIdx.x;
    const int slot   = base + lane / W;
    const bool act   = !g.spare() && slot < count;
    const int run    = run_list[act ? slot : base];
    const int gl     = g.gl;
    const bool hooks = !(LAY == VVS_LAY && geo.sim_kind != 1);
    // the run's parameter block is read from global memory (L2): mobility damping at the departure time, and the
    // rate coefficients of the destination's parameter set (kept in registers when the run has only one set)
    const double* blk = a.st.run_block + (size_t)run * geo.block_stride;
    const double* mob = geo.mob_max_pts > 0 ? blk + geo.mob_off : nullptr;
    const bool damped = mob != nullptr && mob[0] > 0.0;
    {
        double damp[NC];
        mobility_damping_eval(damped ? mob : nullptr, geo.mob_max_pts, geo.n, gl, a.t_dep, damp);
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            damp_s[i * 32 + lane] = damp[i];
        }
        const double* rec = blk + PHDR + gl * AGE_REC;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            ct_s[i * 32 + lane] = rec[i];
        }
    }
    const bool one_pset = geo.num_psets == 1;
    const size_t rbase  = (size_t)run * geo.N;
    unsigned maskbits   = 0u; // mobile compartments of this lane
    if constexpr (FACT) {
        for (int i = 0; i < NC; ++i) {
            maskbits |= (a.csr.umask[gl * NC + i] != 0.0 ? 1u : 0u) << i;
        }
    }

    const int E = geo.E;
    if (E == 0) {
        return;
    }
    // pipeline: indices of edge e+1, operands of edge e
    int f_next = E > 1 ? __ldg(a.csr.edge_start + 1) : 0;
    int t_next = E > 1 ? __ldg(a.csr.edge_end + 1) : 0;
    ExactEdge<G, LAY, FACT> cur;
    exact_edge_load<G, LAY, FACT>(a, geo, run, gl, 0, __ldg(a.csr.edge_start), __ldg(a.csr.edge_end), cur);
    // source node state held in registers while the source stays the same; it and the destination's running state
    // are fetched at the end of the previous edge's iteration (after that edge's stores, so that the loads see them)
    double yf[NC], ytn[NC], ff, nd;
    auto hold_source = [&](int node) {
        double sf[NC];
        load10(a.snap + (rbase + node) * geo.n + gl * NC, sf);
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            sf_s[i * 32 + lane] = sf[i];
        }
    };
    {
        const int from0 = cur.from;
        const int ps_f  = a.st.node_pset ? a.st.node_pset[from0] : 0;
        nd              = hooks ? nondetection_at(blk + ps_f * geo.pset_stride, a.t_dep) : 1.0;
        hold_source(from0);
        ff = a.fac[(rbase + from0) * W + gl];
        load10(a.y_out + (rbase + from0) * geo.n + gl * NC, yf);
        load10(a.y_out + (rbase + cur.to) * geo.n + gl * NC, ytn);
    }
    for (int e = 0; e < E; ++e) {
        // fetch ahead: indices of e+2, operands of e+1
        const int f_next2 = e + 2 < E ? __ldg(a.csr.edge_start + e + 2) : 0;
        const int t_next2 = e + 2 < E ? __ldg(a.csr.edge_end + e + 2) : 0;
        ExactEdge<G, LAY, FACT> nxt;
        if (e + 1 < E) {
            exact_edge_load<G, LAY, FACT>(a, geo, run, gl, e + 1, f_next, t_next, nxt);
        }
        const int from  = cur.from;
        const int to    = cur.to;
        const size_t rt = rbase + to;
        double yt[NC];
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            yt[i] = ytn[i];
        }
        LaneCoef c_t;
        if (!one_pset) {
            const int ps_t = a.st.node_pset ? a.st.node_pset[to] : 0;
            c_t.load(blk + ps_t * geo.pset_stride + PHDR + gl * AGE_REC);
        }
        else {
            double rec[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                rec[i] = ct_s[i * 32 + lane];
            }
            c_t.load(rec);
        }
        // mobile population of the edge at departure (instant.h:608-615)
        double m[NC];
        bool pos = false;
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            // FACT: w * mask with a 0/1 mask is the stored coefficient, bit for bit
            double c = FACT ? (((maskbits >> i) & 1u) ? cur.c[0] : 0.0) : cur.c[i];
            if (damped) {
                c = __dsub_rn(c, __dmul_rn(damp_s[i * 32 + lane], c));
            }
            pos  = pos || (c > 0.0);
            m[i] = sf_s[i * 32 + lane] * c;
        }
        m[cISy] *= ff;
        if (nd != 1.0) {
            m[cISy] *= nd;
            m[cINS] *= nd;
        }
        const bool any = g.any(pos); // the reference skips edges without a positive coefficient
        euler_return<G, LAY>(g, c_t, cur.a0.x, cur.a0.y, cur.a1.x, a.dt, m);
        // underflow fix-up, instant.h:579-596: per age group, the supernumerous returns are taken from the biggest
        // remaining compartment of that age group (first maximum in the reference's compartment order)
        double rem[NC];
        int fixes = 0;
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            rem[i] = yt[i] - m[i];
        }
        bool neg_any = false;
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            neg_any = neg_any || (rem[i] < 0);
        }
        if (__any_sync(0xffffffffu, neg_any)) { // uniform over the warp; no lane underflows on most edges
        int max_index = 0; // slot of the lane's own maximum, ties: lowest reference index
        if constexpr (LAY == 1) {
#pragma unroll
            for (int i = 1; i < NC; ++i) {
                if (rem[i] > rem[max_index]) {
                    max_index = i;
                }
            }
        }
        else {
#pragma unroll
            for (int i = 1; i < NC; ++i) {
                if (i != cR && (rem[i] > rem[max_index] ||
                                (rem[i] == rem[max_index] &&
                                 vvs_ref_comp(g.layer, i) < vvs_ref_comp(g.layer, max_index)))) {
                    max_index = i;
                }
            }
        }
        double neg = 0.0; // sum of the negative remainders of this lane
#pragma unroll
        for (int j = 0; j < NC; ++j) {
            if (rem[j] < 0) {
                if constexpr (LAY == 1) {
#pragma unroll
                    for (int q = 0; q < NC; ++q) { // register array: select instead of dynamic index
                        if (q == max_index) {
                            m[q] -= rem[j];
                        }
                    }
                }
                else {
                    neg += rem[j];
                }
                m[j] += rem[j];
                ++fixes;
            }
        }
        if constexpr (LAY > 1) {
            // the age group's maximum over the layers; the owning lane gives up the summed deficit of the age group
            double best = 0.0;
#pragma unroll
            for (int q = 0; q < NC; ++q) {
                if (q == max_index) {
                    best = rem[q];
                }
            }
            int best_ref = vvs_ref_comp(g.layer, max_index);
            bool owner   = true;
#pragma unroll
            for (int k = 1; k < LAY; ++k) {
                int l = gl + k * G;
                l     = l >= W ? l - W : l;
                const int src   = g.spare() ? lane : g.gbase + l;
                const double ov = __shfl_sync(0xffffffffu, best, src);
                const int oref  = __shfl_sync(0xffffffffu, best_ref, src);
                if (ov > best || (ov == best && oref < best_ref)) {
                    owner = false;
                }
            }
            const double deficit = g.sum_layers(neg);
            if (owner && deficit != 0.0) {
#pragma unroll
                for (int q = 0; q < NC; ++q) {
                    if (q == max_index) {
                        m[q] -= deficit;
                    }
                }
            }
        }
        }
        if (any) {
#pragma unroll
            for (int i = 0; i < NC; ++i) {
                yf[i] += m[i];
                yt[i] -= m[i];
            }
            if (a.edge_m != nullptr && act) { // the returning population of the edge (after the fix-up), :599
                store10(a.edge_m + ((size_t)run * E + e) * geo.n + gl * NC, m);
            }
            if (act) {
                store10(a.y_out + rt * geo.n + gl * NC, yt);
                if (fixes) {
                    atomicAdd(a.st.counters + rt * 4 + 3, fixes);
                }
            }
        }
        if (e + 1 < E) {
            if (nxt.from != from) { // uniform over the warp: the next edge starts a new source node
                if (act) {
                    store10(a.y_out + (rbase + from) * geo.n + gl * NC, yf);
                }
                const int ps_f = a.st.node_pset ? a.st.node_pset[nxt.from] : 0;
                nd             = hooks ? nondetection_at(blk + ps_f * geo.pset_stride, a.t_dep) : 1.0;
                hold_source(nxt.from);
                ff = a.fac[(rbase + nxt.from) * W + gl];
                load10(a.y_out + (rbase + nxt.from) * geo.n + gl * NC, yf);
            }
            // (an edge from a node to itself is not supported: its destination state would be the held source state)
            load10(a.y_out + (rbase + nxt.to) * geo.n + gl * NC, ytn);
            cur = nxt;
        }
        f_next = f_next2;
        t_next = t_next2;
    }
    if (act) {
        store10(a.y_out + (rbase + cur.from) * geo.n + gl * NC, yf);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// return_exact_blocks_kernel: the same literal replay, parallel over the edges of ONE source node.
// The reference's edge order is grouped by source; inside a source's block every destination occurs once, and the
// underflow test of an edge only reads the running state of its destination, so the edges of a block are independent
// of each other; only the blocks are sequential. One CTA owns a flagged run, its running state Y[N][n] lives in
// shared memory (N*n*8 bytes: 192 KB for 400 counties x 6 age groups); local memory must be avoided (the L1 left
// beside the state is tiny: spills and indexed local arrays cost an L2 round trip each). A chunk is <= EPP edges of one source; in step s three groups of warps work side by side
// on three chunks (ring of three chunk buffers B in shared memory), one barrier per step:
//   R warps (one W-lane group per edge): chunk s+1 -- the returning population of every edge by the explicit Euler step
//     of the destination (order independent; operands fetched two chunks ahead) -> B;
//   T warps (one W-lane group per edge): chunk s -- in order: test against Y[dest], fix-up (-> B), Y[dest] -= m;
//   S warps (one lane per state element): chunk s-1 -- Y[source] += m, edge after edge in edge order.
// When the source of chunk s-1 is a destination of chunk s (flag in the schedule) the T warps wait for the S warps.
// Every node's running state sees exactly the reference's sequence of operations (sources ascending, own block in edge
// order), so the result is bit-identical to return_exact_kernel and to the oracle. LAY == 1 (ODE-SECIR) only.
template <int G, bool FACT>
struct ExbOps {
    double c[FACT ? 1 : NC]; // coefficient vector of the edge for this lane (FACT: the weight)
    double2 a0, a1;          // return coefficients of the destination (aux)
    int e, to, from;         // e < 0: empty slot
};

template <int G, bool FACT>
__global__ void __launch_bounds__(EXB_THREADS, 1)
    return_exact_blocks_kernel(EdgeArgs a, const int* __restrict__ run_list, const int* __restrict__ list_count)
{
    constexpr int LAY = 1;
    constexpr int W   = G;
    constexpr int NPW = 32 / W;
    constexpr int TW  = EXB_P1_WARPS;
    constexpr int EPP = TW * NPW;
    extern __shared__ __align__(16) double smem[];
    const Geometry& geo = a.geo;
    const int n         = geo.n;
    double* const Y      = smem;                      // [N][n]
    double* const B      = Y + (size_t)geo.N * n;     // [3][EPP][n] returning populations of three chunks
    double* const damp_s = B + (size_t)3 * EPP * n;   // [W][NC]
    double* const rate_s = damp_s + W * NC;           // [W][17] (odd stride: the lanes of a group hit distinct banks)
    const int SS         = (n + W + 1) & ~1;          // (even: 16-byte loads)
    double* const snap_s = rate_s + W * 18;           // [2][SS] snapshot at departure and mobility factors of the
                                                      // source node of a chunk (staged by the S warps)
    int* const bany      = reinterpret_cast<int*>(snap_s + 2 * SS); // [3][EPP] the edge carries commuters
    const int count = *list_count;
    const int warp  = threadIdx.x >> 5;
    const bool is_t = warp < TW;
    const bool is_r = !is_t && warp < 2 * TW;
    const bool is_s = warp >= 2 * TW;
    Group<G, LAY> g;
    const int lane  = g.lane;
    const int gl    = g.gl;
    const int q     = (is_t ? warp : warp - TW) * NPW + lane / W; // T, R: the lane group's slot in a chunk
    const bool lane_on = !is_s && !g.spare();
    const int nchunks  = a.xchunks;
    const size_t total2 = (size_t)geo.N * n / 2;
    const size_t bstride = (size_t)EPP * n;
    unsigned maskbits   = 0u;
    if constexpr (FACT) {
        for (int i = 0; i < NC; ++i) {
            maskbits |= (a.csr.umask[gl * NC + i] != 0.0 ? 1u : 0u) << i;
        }
    }

    for (int item = blockIdx.x; item < count; item += gridDim.x) {
        const int run      = run_list[item];
        const size_t rbase = (size_t)run * geo.N;
        const double* blk  = a.st.run_block + (size_t)run * geo.block_stride;
        const double* mob  = geo.mob_max_pts > 0 ? blk + geo.mob_off : nullptr;
        const bool damped  = mob != nullptr && mob[0] > 0.0;
        const bool one_pset = geo.num_psets == 1;
        __syncthreads(); // the previous run's state has left shared memory
        {
            const double2* src = reinterpret_cast<const double2*>(a.y_in + rbase * n);
            double2* dst       = reinterpret_cast<double2*>(Y);
            for (size_t i = threadIdx.x; i < total2; i += EXB_THREADS) {
                dst[i] = src[i];
            }
            if (threadIdx.x < W * NC) {
                // mobility damping at the departure time, evaluated by collect_flagged_kernel (no call, and so no
                // stack frame, in this kernel)
                damp_s[threadIdx.x] = damped ? a.xdamp[(size_t)run * n + threadIdx.x] : 0.0;
            }
            if (threadIdx.x < W * 16) {
                const int gi = threadIdx.x / 16, ri = threadIdx.x - gi * 16;
                rate_s[gi * 17 + ri] = blk[PHDR + gi * AGE_REC + ri];
            }
        }
        const double nd_run = nondetection_at(blk, a.t_dep); // one parameter set per run

        // R: operands of (chunk, slot q) for this lane
        auto load_ops = [&](int e, int to, int from, ExbOps<G, FACT>& d) {
            d.e    = lane_on ? e : -1;
            d.to   = to;
            d.from = from;
            if (d.e >= 0) {
                if constexpr (FACT) {
                    d.c[0] = __ldg(a.csr.w_edge + (size_t)e * W + gl);
                }
                else {
                    const double2* cp = reinterpret_cast<const double2*>(a.csr.coeff + (size_t)e * n + gl * NC);
#pragma unroll
                    for (int i = 0; i < NC / 2; ++i) {
                        const double2 v = __ldg(cp + i);
                        d.c[2 * i]      = v.x;
                        d.c[2 * i + 1]  = v.y;
                    }
                }
                const double2* ax = reinterpret_cast<const double2*>(a.aux + ((rbase + to) * W + gl) * AUX);
                d.a0              = ax[0];
                d.a1              = ax[1];
            }
        };
        // S: snapshot and mobility factors of the source of chunk c -> shared memory (two chunks ahead of the R warps)
        auto stage_source = [&](int c) {
            if (c < nchunks) {
                const int from = __ldg(a.xsrc + 2 * c);
                double* dst    = snap_s + (c & 1) * SS;
                for (int j = threadIdx.x - 2 * TW * 32; j < n + W; j += EXB_P2_WARPS * 32) {
                    dst[j] = j < n ? a.snap[(rbase + from) * n + j] : a.fac[(rbase + from) * W + (j - n)];
                }
            }
        };
        // R: the returning population of the lane's edge (instant.h:608-615 at departure, :460-471 the Euler step)
        // -> chunk buffer
        auto returning = [&](const ExbOps<G, FACT>& d, int buf, int c) {
            bool pos = false;
            double m[NC];
            if (d.e >= 0) {
                const double* srec = snap_s + (c & 1) * SS;
                double sf[NC];
                lds10(srec + gl * NC, sf);
                const double ff = srec[n + gl];
                LaneCoef c_t;
                double nd = nd_run;
                if (one_pset) {
                    c_t.load(rate_s + gl * 17);
                }
                else {
                    const int ps_t = a.st.node_pset ? a.st.node_pset[d.to] : 0;
                    const int ps_f = a.st.node_pset ? a.st.node_pset[d.from] : 0;
                    c_t.load(blk + ps_t * geo.pset_stride + PHDR + gl * AGE_REC);
                    nd = nondetection_at(blk + ps_f * geo.pset_stride, a.t_dep);
                }
#pragma unroll
                for (int i = 0; i < NC; ++i) {
                    double c = FACT ? (((maskbits >> i) & 1u) ? d.c[0] : 0.0) : d.c[i];
                    if (damped) {
                        c = __dsub_rn(c, __dmul_rn(damp_s[gl * NC + i], c));
                    }
                    pos  = pos || (c > 0.0);
                    m[i] = sf[i] * c;
                }
                m[cISy] *= ff;
                if (nd != 1.0) {
                    m[cISy] *= nd;
                    m[cINS] *= nd;
                }
                euler_return<G, LAY>(g, c_t, d.a0.x, d.a0.y, d.a1.x, a.dt, m);
            }
            const bool any = g.any(pos); // the reference skips edges without a positive coefficient
            if (lane_on) {
                // a slot without commuters (empty, or no positive coefficient) holds -0.0: the S warps add every row
                // of the buffer, and x + -0.0 == x for every x
                if (!(d.e >= 0 && any)) {
#pragma unroll
                    for (int i = 0; i < NC; ++i) {
                        m[i] = -0.0;
                    }
                }
                store10(B + buf * bstride + (size_t)q * n + gl * NC, m);
                if (gl == 0) {
                    bany[buf * EPP + q] = (d.e >= 0 && any) ? 1 : 0;
                }
            }
        };
        auto slot_of = [&](int s, int& e, int& to, int& from) {
            e = -1, to = 0, from = 0;
            if (lane_on && s < nchunks) { // (spare lanes have no slot)
                const int2 v = __ldg(reinterpret_cast<const int2*>(a.xslot) + (size_t)s * EPP + q);
                e            = v.x;
                to           = v.y;
                from         = __ldg(a.xsrc + 2 * s);
            }
        };

        ExbOps<G, FACT> near, far;
        int e_n = -1, to_n = 0, from_n = 0; // R: slot of chunk s + 3; T: slot of chunk s + 1
        int e_cur = -1, to_cur = 0;        // T: slot of chunk s
        int from_prev = 0, from_next = 0;  // S: source of chunk s - 1 / s
        if (is_r) {
            int e0, t0, f0;
            slot_of(0, e0, t0, f0);
            load_ops(e0, t0, f0, far);
            slot_of(1, e0, t0, f0);
            load_ops(e0, t0, f0, near);
            slot_of(2, e_n, to_n, from_n);
        }
        else if (is_t) {
            int f0;
            slot_of(0, e_cur, to_cur, f0);
            if (!lane_on) {
                e_cur = -1;
            }
            slot_of(1, e_n, to_n, f0);
        }
        else {
            stage_source(0);
            stage_source(1);
            from_next = nchunks > 0 ? __ldg(a.xsrc) : 0;
        }
        __syncthreads(); // the state and the run constants are in shared memory
        if (is_r) {
            returning(far, 0, 0);
        }
        __syncthreads();
        for (int s = 0; s <= nchunks; ++s) {
            // chunk s - 1's source is a destination of chunk s: the T warps wait for the S warps
            const bool hazard = s > 0 && s < nchunks && (__ldg(a.xsrc + 2 * s + 1) >> 16) != 0;
            if (is_r) {
                if (s + 1 < nchunks) {
                    // fetch ahead: operands of chunk s + 3, slot of chunk s + 4
                    int e_f, to_f, from_f;
                    slot_of(s + 3, e_f, to_f, from_f);
                    load_ops(e_n, to_n, from_n, far);
                    e_n = e_f, to_n = to_f, from_n = from_f;
                    returning(near, (s + 1) % 3, s + 1);
                    near = far;
                }
            }
            else if (is_t) {
                if (hazard) {
                    asm volatile("bar.sync 1, %0;" ::"r"((TW + EXB_P2_WARPS) * 32) : "memory");
                }
                if (s < nchunks) {
                    // in order: underflow test against the destination's running state, fix-up (instant.h:579-596)
                    const int buf = s % 3;
                    int e_f, to_f, from_f;
                    slot_of(s + 2, e_f, to_f, from_f);
                    double* const mrec = B + buf * bstride + (size_t)q * n + gl * NC;
                    double* const yrec = Y + (size_t)to_cur * n + gl * NC;
                    const bool any_cur = e_cur >= 0 && bany[buf * EPP + q] != 0;
                    int fixes = 0;
                    double yt[NC], m[NC], rem[NC];
                    bool neg_any = false;
                    if (e_cur >= 0) {
                        lds10(yrec, yt);
                        lds10(mrec, m);
#pragma unroll
                        for (int i = 0; i < NC; ++i) {
                            rem[i]  = yt[i] - m[i];
                            neg_any = neg_any || (rem[i] < 0);
                        }
                    }
                    if (__any_sync(0xffffffffu, neg_any)) {
                        if (neg_any) {
                            // the deficits go to the age group's biggest remainder (first maximum); that element is
                            // kept in a scalar while the deficits are applied in compartment order
                            int max_index = 0;
                            double best = rem[0], mm = m[0];
#pragma unroll
                            for (int i = 1; i < NC; ++i) {
                                if (rem[i] > best) {
                                    best      = rem[i];
                                    max_index = i;
                                    mm        = m[i];
                                }
                            }
#pragma unroll
                            for (int j = 0; j < NC; ++j) {
                                if (rem[j] < 0) {
                                    mm -= rem[j];
                                    if (j == max_index) {
                                        mm += rem[j];
                                    }
                                    else {
                                        m[j] += rem[j];
                                    }
                                    ++fixes;
                                }
                            }
#pragma unroll
                            for (int k = 0; k < NC; ++k) {
                                if (k == max_index) {
                                    m[k] = mm;
                                }
                            }
                        }
                    }
                    if (any_cur) {
#pragma unroll
                        for (int i = 0; i < NC; ++i) {
                            yt[i] -= m[i];
                        }
                        store10(yrec, yt);
                        if (fixes) {
                            store10(mrec, m);
                            atomicAdd(a.st.counters + (rbase + to_cur) * 4 + 3, fixes);
                        }
                    }
                    e_cur  = lane_on ? e_n : -1;
                    to_cur = to_n;
                    e_n = e_f, to_n = to_f;
                }
            }
            else {
                if (s >= 1) {
                    // the commuters of chunk s - 1 come home: Y[source] += m, edge after edge (scalars, not arrays)
                    const int from = from_prev;
                    const double* bp = B + ((s - 1) % 3) * bstride;
                    constexpr int NN = W * NC; // == n
                    for (int comp = threadIdx.x - 2 * TW * 32; comp < NN; comp += EXB_P2_WARPS * 32) {
                        double acc       = Y[(size_t)from * NN + comp];
                        const double* mp = bp + comp;
#pragma unroll 5
                        for (int k = 0; k < EPP; ++k) {
                            acc += mp[k * NN];
                        }
                        Y[(size_t)from * NN + comp] = acc;
                    }
                }
                from_prev = from_next;
                from_next = s + 1 < nchunks ? __ldg(a.xsrc + 2 * (s + 1)) : 0;
                if (hazard) {
                    asm volatile("bar.sync 1, %0;" ::"r"((TW + EXB_P2_WARPS) * 32) : "memory");
                }
                stage_source(s + 2);
            }
            __syncthreads();
        }
        {
            const double2* src = reinterpret_cast<const double2*>(Y);
            double2* dst       = reinterpret_cast<double2*>(a.y_out + rbase * n);
            for (size_t i = threadIdx.x; i < total2; i += EXB_THREADS) {
                dst[i] = src[i];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// get_mobility_results support (single simulations that ask for it; not on the ensemble path).
// Departure: the mobile population of edge e is coeff_e * u_from (u = snapshot * (1 - damping) * mobility factor *
// commuter nondetection * mask, see integrate.cu), instant.h:608-615.
__global__ void edge_mobile_depart_kernel(int n, int N, int E, const int* __restrict__ edge_start,
                                          const double* __restrict__ coeff, const double* __restrict__ u,
                                          double* __restrict__ edge_m, size_t total)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) {
        const size_t re  = i / n;
        const int j      = (int)(i - re * n);
        const size_t run = re / E;
        const int e      = (int)(re - run * E);
        edge_m[i]        = coeff[(size_t)e * n + j] * u[(run * N + edge_start[e]) * n + j];
    }
}

cudaError_t launch_edge_mobile_depart(const EdgeArgs& a, cudaStream_t stream)
{
    const size_t total = (size_t)a.num_runs * a.geo.E * a.geo.n;
    if (total > 0) {
        edge_mobile_depart_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(
            a.geo.n, a.geo.N, a.geo.E, a.csr.edge_start, a.csr.coeff, a.u, a.edge_m, total);
    }
    return cudaGetLastError();
}

// add_mobility_result_time_point, instant.h:421-445: per edge the sums over the saved index groups (in the order the
// indices were given) and the total over all compartments (reference element order), one thread per (run, edge)
__global__ void mobility_results_kernel(MobilityResultArgs a)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)a.num_runs * a.E) {
        return;
    }
    const int e = (int)(i % a.E);
    if (!a.edge_active[e]) {
        return;
    }
    const double* m = a.edge_m + i * a.n;
    double* out     = a.values + (i * a.capacity + a.slot) * (a.K + 1);
    for (int g = 0; g < a.K; ++g) {
        double s = 0.0;
        for (int q = a.group_ptr[g]; q < a.group_ptr[g + 1]; ++q) {
            s += m[a.group_dev_idx[q]];
        }
        out[g] = s;
    }
    double tot = 0.0;
    for (int j = 0; j < a.n_ref; ++j) {
        tot += m[a.ref_to_dev[j]];
    }
    out[a.K] = tot;
}

cudaError_t launch_mobility_results(const MobilityResultArgs& a, cudaStream_t stream)
{
    const size_t total = (size_t)a.num_runs * a.E;
    if (total > 0) {
        mobility_results_kernel<<<(unsigned)((total + 127) / 128), 128, 0, stream>>>(a);
    }
    return cudaGetLastError();
}

__global__ void record_daily_kernel(int N, int n, const double* __restrict__ y, double* __restrict__ daily,
                                    int day_index, int max_days, size_t total)
{
    // y[R][N][n] -> daily[R][N][max_days][n] at slot day_index, 16-byte vectors
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t v = (size_t)n / 2;
    if (i < total) {
        const size_t rn = i / v;
        const size_t k  = i - rn * v;
        reinterpret_cast<double2*>(daily)[(rn * max_days + day_index) * v + k] =
            reinterpret_cast<const double2*>(y)[i];
    }
}

static int edge_sm_count()
{
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    return sms;
}
// MIO_EXACT_REPLAY=blocks selects return_exact_blocks_kernel. The default is return_exact_kernel: measured on the
// benchmark (10 000 runs x 400 counties, 38 124 replays per pass) the block-parallel kernel needs 165-220 ms per pass
// against 143 ms -- one run per SM (its state fills the shared memory) leaves two or three warps per scheduler, and the
// chain ordered-sum -> test of the next source's block (the source of a block is nearly always a destination of the
// next one in a county graph) costs about 3000 cycles per source node at that occupancy; see DESIGN.md section 9.
static bool exact_replay_blocks()
{
    static const bool blocks = [] {
        const char* v = std::getenv("MIO_EXACT_REPLAY");
        return v != nullptr && std::strcmp(v, "blocks") == 0;
    }();
    return blocks;
}

template <int G, int LAY>
static cudaError_t launch_edges_g(const EdgeArgs& a, int which, cudaStream_t stream)
{
    constexpr int W     = G * LAY;
    constexpr int NPW   = 32 / W;
    constexpr int WARPS = EDGE_THREADS / 32;
    const int chunks    = (a.geo.N + NPW * WARPS - 1) / (NPW * WARPS);
    const size_t grid   = (size_t)chunks * a.num_runs;
    const size_t shmem  = (size_t)a.geo.block_stride * 8;
#define MIO_EDGE_LAUNCH(KERN, GRID, THREADS)                                                                           \
    do {                                                                                                               \
        auto kern = KERN;                                                                                              \
        if (shmem > 48 * 1024) {                                                                                       \
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem);                       \
        }                                                                                                              \
        kern<<<(unsigned)(GRID), THREADS, shmem, stream>>>(a);                                                         \
    } while (0)
    if (which == 0) {
        const int M          = a.csr.num_mob <= 5 ? 5 : NC;
        const size_t shmem_u = shmem + sizeof(double) * (size_t)a.geo.N * W * M;
        if (a.csr.factorized && a.csr.num_mob > 0 && shmem_u <= 110 * 1024) {
            // one CTA of 512 threads per SM, one run per CTA. MIO_DEPART_RUNS=2: two runs per CTA (the run-invariant
            // edge list / weight loads are shared) -- measured slower (148 vs 131 ms per 30 days: 272 B of spills at
            // 128 registers and half as many CTAs to balance), kept for comparisons only
            static const bool one_run = [] {
                const char* v = std::getenv("MIO_DEPART_RUNS");
                return !(v != nullptr && std::atoi(v) == 2);
            }();
            const bool two = M == 5 && a.aux_by_node_phase && 2 * shmem_u <= 227 * 1024 && !one_run;
            if (two) {
                auto kern = depart_smem_kernel<G, LAY, 5, 2, true>;
                cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * shmem_u));
                kern<<<(unsigned)((a.num_runs + 1) / 2), DEPART_SMEM_THREADS, 2 * shmem_u, stream>>>(a);
            }
            else if (M == 5 && a.aux_by_node_phase) {
                auto kern = depart_smem_kernel<G, LAY, 5, 1, true>;
                cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem_u);
                kern<<<(unsigned)a.num_runs, DEPART_SMEM_THREADS, shmem_u, stream>>>(a);
            }
            else if (M == 5) {
                auto kern = depart_smem_kernel<G, LAY, 5, 1, false>;
                cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem_u);
                kern<<<(unsigned)a.num_runs, DEPART_SMEM_THREADS, shmem_u, stream>>>(a);
            }
            else {
                auto kern = depart_smem_kernel<G, LAY, NC, 1, false>;
                cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem_u);
                kern<<<(unsigned)a.num_runs, DEPART_SMEM_THREADS, shmem_u, stream>>>(a);
            }
        }
        else if (a.csr.factorized) {
            MIO_EDGE_LAUNCH((depart_kernel<G, LAY, true>), grid, EDGE_THREADS);
        }
        else {
            MIO_EDGE_LAUNCH((depart_kernel<G, LAY, false>), grid, EDGE_THREADS);
        }
    }
    else if (which == 1) {
        const size_t shmem_t =
            shmem + sizeof(double) * (size_t)(RETURN_TMA_THREADS / 32) * RETURN_TMA_ARRAYS * NPW * a.geo.n;
        if (a.csr.factorized && a.geo.num_psets == 1 && shmem_t <= 110 * 1024) {
            auto kern = return_tma_kernel<G, LAY>;
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem_t);
            kern<<<(unsigned)a.num_runs, RETURN_TMA_THREADS, shmem_t, stream>>>(a);
        }
        else if (a.csr.factorized) {
            MIO_EDGE_LAUNCH((return_kernel<G, LAY, true>), grid, EDGE_THREADS);
        }
        else {
            MIO_EDGE_LAUNCH((return_kernel<G, LAY, false>), grid, EDGE_THREADS);
        }
    }
    else {
        // a.run_list: [R] run indices, a.list_count: [1]
        cudaMemsetAsync(a.list_count, 0, sizeof(int), stream);
        // block-parallel replay when the run's state fits into shared memory (ODE-SECIR; not when the per-edge
        // populations are recorded)
        const size_t shmem_x = sizeof(double) * ((size_t)a.geo.N * a.geo.n + (size_t)3 * a.xepp * a.geo.n + W * NC + W * 18 + 2 * (a.geo.n + W + 1)) +
                               sizeof(int) * (size_t)3 * a.xepp;
        bool blocks = false;
        if constexpr (LAY == 1) {
            blocks = a.xslot != nullptr && a.edge_m == nullptr && shmem_x <= 227 * 1024 && exact_replay_blocks();
        }
        collect_flagged_kernel<256><<<(unsigned)a.num_runs, 256, 0, stream>>>(a, a.run_list, a.list_count, blocks ? 0 : 1);
        if constexpr (LAY == 1) {
            if (blocks) {
                const unsigned ctas = (unsigned)std::min(a.num_runs, edge_sm_count() * (shmem_x <= 110 * 1024 ? 2 : 1));
                if (a.csr.factorized && a.csr.w_edge != nullptr) {
                    auto kern = return_exact_blocks_kernel<G, true>;
                    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem_x);
                    kern<<<ctas, EXB_THREADS, shmem_x, stream>>>(a, a.run_list, a.list_count);
                }
                else {
                    auto kern = return_exact_blocks_kernel<G, false>;
                    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem_x);
                    kern<<<ctas, EXB_THREADS, shmem_x, stream>>>(a, a.run_list, a.list_count);
                }
                return cudaGetLastError();
            }
        }
        const unsigned warps = (unsigned)((a.num_runs + NPW - 1) / NPW);
        if (a.csr.factorized && a.csr.w_edge != nullptr) {
            return_exact_kernel<G, LAY, true><<<warps, 32, 0, stream>>>(a, a.run_list, a.list_count);
        }
        else {
            return_exact_kernel<G, LAY, false><<<warps, 32, 0, stream>>>(a, a.run_list, a.list_count);
        }
    }
#undef MIO_EDGE_LAUNCH
    return cudaGetLastError();
}

template <int LAY>
static cudaError_t launch_edges_lay(const EdgeArgs& a, int which, cudaStream_t stream)
{
    switch (a.geo.G) {
    case 1:
        return launch_edges_g<1, LAY>(a, which, stream);
    case 2:
        return launch_edges_g<2, LAY>(a, which, stream);
    case 3:
        return launch_edges_g<3, LAY>(a, which, stream);
    case 4:
        return launch_edges_g<4, LAY>(a, which, stream);
    case 5:
        return launch_edges_g<5, LAY>(a, which, stream);
    case 6:
        return launch_edges_g<6, LAY>(a, which, stream);
    case 7:
        return launch_edges_g<7, LAY>(a, which, stream);
    case 8:
        return launch_edges_g<8, LAY>(a, which, stream);
    default:
        return cudaErrorInvalidValue;
    }
}

static cudaError_t launch_edges(const EdgeArgs& a, int which, cudaStream_t stream)
{
    return a.geo.LAY == VVS_LAY ? launch_edges_lay<VVS_LAY>(a, which, stream) : launch_edges_lay<1>(a, which, stream);
}

cudaError_t launch_depart(const EdgeArgs& a, cudaStream_t stream)
{
    return launch_edges(a, 0, stream);
}
cudaError_t launch_return(const EdgeArgs& a, cudaStream_t stream)
{
    return launch_edges(a, 1, stream);
}
cudaError_t launch_return_exact(const EdgeArgs& a, cudaStream_t stream)
{
    return launch_edges(a, 2, stream);
}

// same copy for a layered model: device order [LAY][G][NC] -> reference order age*C + compartment
__global__ void record_daily_ref_kernel(int N, int G, int lay, int n_dev, int C_ref, const double* __restrict__ y,
                                        double* __restrict__ daily, int day_index, int max_days, size_t total)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int n_ref = G * C_ref;
    if (i < total) {
        const size_t rn = i / n_ref;
        const int j     = (int)(i - rn * n_ref);
        const int age   = j / C_ref;
        daily[(rn * max_days + day_index) * n_ref + j] = y[rn * n_dev + dev_index_of_ref(lay, G, age, j - age * C_ref)];
    }
}

cudaError_t launch_record_daily(const Geometry& geo, const double* y, double* daily, int day_index, int max_days,
                                int num_runs, cudaStream_t stream)
{
    const int threads = 256;
    if (geo.LAY == 1) {
        const size_t total = (size_t)num_runs * geo.N * (geo.n / 2);
        const size_t grid  = (total + threads - 1) / threads;
        record_daily_kernel<<<(unsigned)grid, threads, 0, stream>>>(geo.N, geo.n, y, daily, day_index, max_days,
                                                                    total);
    }
    else {
        const size_t total = (size_t)num_runs * geo.N * geo.n_ref;
        const size_t grid  = (total + threads - 1) / threads;
        record_daily_ref_kernel<<<(unsigned)grid, threads, 0, stream>>>(geo.N, geo.G, geo.LAY, geo.n, geo.C_ref, y,
                                                                        daily, day_index, max_days, total);
    }
    return cudaGetLastError();
}

// state buffers <-> reference element order (C-ABI side of a layered model); dst and src must not overlap
__global__ void permute_state_kernel(int G, int lay, int n_dev, int C_ref, const double* __restrict__ src,
                                     double* __restrict__ dst, size_t total, int to_device)
{
    const size_t i  = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int n_ref = G * C_ref;
    if (i < total) {
        if (to_device) { // i runs over device elements; the unused Recovered slots are zeroed
            const size_t rn = i / n_dev;
            const int d     = (int)(i - rn * n_dev);
            const int gl = d / NC, slot = d - gl * NC;
            const int layer = gl / G, age = gl - layer * G;
            const int comp  = vvs_ref_comp(layer, slot);
            dst[i]          = comp < 0 ? 0.0 : src[rn * n_ref + age * C_ref + comp];
        }
        else { // i runs over reference elements
            const size_t rn = i / n_ref;
            const int j     = (int)(i - rn * n_ref);
            const int age   = j / C_ref;
            dst[i]          = src[rn * n_dev + dev_index_of_ref(lay, G, age, j - age * C_ref)];
        }
    }
}

cudaError_t launch_permute_state(const Geometry& geo, const double* src, double* dst, size_t num_run_nodes,
                                 int to_device, cudaStream_t stream)
{
    const size_t total = num_run_nodes * (to_device ? geo.n : geo.n_ref);
    const int threads  = 256;
    const size_t grid  = (total + threads - 1) / threads;
    if (grid > 0) {
        permute_state_kernel<<<(unsigned)grid, threads, 0, stream>>>(geo.G, geo.LAY, geo.n, geo.C_ref, src, dst, total,
                                                                     to_device);
    }
    return cudaGetLastError();
}

} // namespace miocu
