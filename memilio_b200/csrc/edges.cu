// edges.cu -- edge phase of GraphSimulation::advance on B200: MobilityEdge::apply_mobility for all edges of all runs.
//
// Replaces cpp/memilio/mobility/metapopulation_mobility_instant.h:546-623 (apply_mobility), :460-471
// (calculate_mobility_returns: one explicit Euler step of the destination node's model with pop = destination state
// at departure), cpp/models/ode_secir/model.h:687-742 (get_mobility_factors, test_commuters) and the sequential
// in-place edge loop of cpp/memilio/mobility/graph_simulation.h:115-118.
//
// Formulation: the reference walks the edges in (start,end) order and updates both end nodes in place. Every
// contribution of an edge depends only on per-node snapshots (m_last_state of the source, post-departure state of the
// destination), so the loop is re-expressed as a *gather per node*: the lane group that owns node k walks k's incident
// edges in the same global order and applies the same additions/subtractions to k's state held in registers. No
// atomics, deterministic, and the floating point summation order per node equals the reference's. The traveler state
// of an edge is not stored between departure and return; it is recomputed from the source snapshot (bit-identical),
// which turns E*n doubles of HBM traffic per run and exchange into L2-resident snapshot reads.
// The one order dependent piece, the underflow fix-up of returns (:579-596), is detected by the gather (run_flags) and
// such runs are replayed for that step by return_exact_kernel with the literal sequential semantics.
#include "device_common.cuh"
#include "kernels.h"

namespace miocu
{

constexpr int AUX = 4; // per (run,node,age): season*p_i*lambda_i, rISev_ICr, rISev_D, pad

template <int G>
struct EdgeLane {
    Group<G> g;
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
__device__ __forceinline__ void euler_return(const LaneCoef& c, double slam, double rISev_ICr, double rISev_D,
                                             double dt, double (&m)[NC])
{
    double k[NC];
    flows_to_dydt(m[cS] * slam, c.rE_INS * m[cE], c.rINS_ISy * m[cINS], c.rINS_R * m[cINS], c.rINS_ISy * m[cINSC],
                  c.rINS_R * m[cINSC], c.rISy_ISev * m[cISy], c.rISy_R * m[cISy], c.rISy_ISev * m[cISyC],
                  c.rISy_R * m[cISyC], rISev_ICr * m[cISev], c.rISev_R * m[cISev], rISev_D * m[cISev],
                  c.rICr_D * m[cICr], c.rICr_R * m[cICr], k);
#pragma unroll
    for (int i = 0; i < NC; ++i) {
        m[i] = m[i] + dt * k[i];
    }
}

template <int G>
__device__ __forceinline__ void edge_lane_setup(const Geometry& geo, const Incidence& inc, int num_runs,
                                                EdgeLane<G>& L, bool& cta_active)
{
    constexpr int NPW   = 32 / G;
    constexpr int WARPS = EDGE_THREADS / 32;
    const int chunks    = (geo.N + NPW * WARPS - 1) / (NPW * WARPS);
    L.run               = blockIdx.x / chunks;
    const int chunk     = blockIdx.x - L.run * chunks;
    cta_active          = L.run < num_runs;
    const int warp      = threadIdx.x >> 5;
    const int item      = (chunk * WARPS + warp) * NPW + L.g.lane / G;
    L.valid             = cta_active && !L.g.spare() && item < geo.N;
    L.node              = L.valid ? inc.node_order[item] : 0;
    L.rn                = (size_t)(cta_active ? L.run : 0) * geo.N + L.node;
}

// ---------------------------------------------------------------------------------------------------------------
// Departure: y_out = y_in + sum(incoming mobile) - sum(outgoing mobile) in edge order; then the return coefficients of
// this node as a destination (lambda_i, ICU adjusted rates) from v0 = y_out at time t.
template <int G>
__global__ void __launch_bounds__(EDGE_THREADS) depart_kernel(EdgeArgs a)
{
    extern __shared__ __align__(16) double smem[];
    __shared__ __align__(8) uint64_t mbar;
    const Geometry& geo = a.geo;
    EdgeLane<G> L;
    bool cta_active;
    edge_lane_setup<G>(geo, a.inc, a.num_runs, L, cta_active);
    if (!cta_active) {
        return;
    }
    stage_run_block(smem, a.st.run_block + (size_t)L.run * geo.block_stride, geo.block_stride * 8, &mbar);
    const Group<G>& g  = L.g;
    const double* mob  = geo.mob_max_pts > 0 ? smem + geo.mob_off : nullptr;
    const bool damped  = mob != nullptr && mob[0] > 0.0;
    const int ps_own   = a.st.node_pset ? a.st.node_pset[L.node] : 0;
    const double* pset = smem + ps_own * geo.pset_stride;

    double y[NC], snapk[NC], damp[NC];
    load10(a.y_in + L.rn * geo.n + g.gl * NC, y);
    load10(a.snap + L.rn * geo.n + g.gl * NC, snapk);
    const double fack = a.fac[L.rn * G + g.gl];
    mobility_damping_eval(damped ? mob : nullptr, geo.mob_max_pts, geo.n, g.gl, a.t, damp);
    const double nd_own = nondetection_at(pset, a.t);

    const int beg = L.valid ? a.inc.inc_ptr[L.node] : 0;
    const int cnt = L.valid ? a.inc.inc_ptr[L.node + 1] - beg : 0;
    int max_cnt   = cnt;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        max_cnt = max(max_cnt, __shfl_xor_sync(0xffffffffu, max_cnt, o));
    }
    for (int it = 0; it < max_cnt; ++it) {
        const bool act = it < cnt;
        const int idx  = beg + (act ? it : 0);
        const int e    = act ? a.inc.inc_edge[idx] : 0;
        const int o    = act ? a.inc.inc_other[idx] : 0;
        const bool out = act ? (a.inc.inc_out[idx] != 0) : true;
        double m[NC], mu[2];
        bool pos;
        double nd = nd_own;
        if (out) {
            pos = damped ? mobile_at_departure<true>(a.inc.coeff + (size_t)e * geo.n + g.gl * NC, damp, snapk, fack,
                                                     nd, m, mu)
                         : mobile_at_departure<false>(a.inc.coeff + (size_t)e * geo.n + g.gl * NC, damp, snapk, fack,
                                                      nd, m, mu);
        }
        else {
            const size_t ro = (size_t)L.run * geo.N + o;
            double so[NC];
            load10(a.snap + ro * geo.n + g.gl * NC, so);
            const double fo = a.fac[ro * G + g.gl];
            if (geo.num_psets > 1) {
                nd = nondetection_at(smem + a.st.node_pset[o] * geo.pset_stride, a.t);
            }
            pos = damped ? mobile_at_departure<true>(a.inc.coeff + (size_t)e * geo.n + g.gl * NC, damp, so, fo, nd,
                                                     m, mu)
                         : mobile_at_departure<false>(a.inc.coeff + (size_t)e * geo.n + g.gl * NC, damp, so, fo, nd,
                                                      m, mu);
        }
        // (coeffs(t) > 0).any(), instant.h:605-606
        const bool any = g.any(act && pos);
        if (act && any) {
            if (out) {
                if (nd != 1.0) { // test_commuters acts on the source node, ode_secir/model.h:733-736
                    y[cINS] -= mu[0] * (1 - nd);
                    y[cINSC] += mu[0] * (1 - nd);
                    y[cISy] -= mu[1] * (1 - nd);
                    y[cISyC] += mu[1] * (1 - nd);
                }
#pragma unroll
                for (int i = 0; i < NC; ++i) {
                    y[i] -= m[i];
                }
            }
            else {
#pragma unroll
                for (int i = 0; i < NC; ++i) {
                    y[i] += m[i];
                }
            }
        }
    }
    if (L.valid) {
        store10(a.y_out + L.rn * geo.n + g.gl * NC, y);
    }
    // return coefficients of this node as destination: get_flows with pop = v0 (ode_secir/model.h:90-153)
    {
        LaneCoef c;
        const double* age_rec = pset + PHDR + g.gl * AGE_REC;
        c.load(age_rec);
        const int ci         = a.st.node_ctable ? a.st.node_ctable[L.node] : 0;
        const double icu_cap = a.st.node_local[L.rn * 2 + 0];
        const double tnt_cap = a.st.node_local[L.rn * 2 + 1];
        double crow[G], lo, hi;
        contact_row_eval<G>(smem + geo.ct_off + ci * geo.ctable_stride, geo.ct_max_loc, geo.ct_max_pts, g.gl, a.t,
                            crow, lo, hi);
        const double seasonality = pset[hSEASONALITY];
        const double season =
            seasonality == 0.0
                ? 1.0
                : 1 + seasonality *
                          sin(3.141592653589793238462643383279502884 * ((pset[hSTART_DAY] + a.t) / 182.5 + 0.5));
        const double tnt  = g.sum(c.rTNT * y[cINS]);
        const double icu  = g.sum(y[cICr]);
        const double Nj   = y[cS] + y[cE] + y[cINS] + y[cISy] + y[cISev] + y[cICr] + y[cR];
        const double divN = (Nj < 1e-12) ? 0.0 : 1.0 / Nj;
        const double risk = smoother_cosine(tnt, tnt_cap, tnt_cap * pset[hTNT_MAX_RISK], c.rRISK, c.rMAXRISK);
        const double w    = divN * (c.rREL * y[cINS] + risk * y[cISy]);
        double lam        = 0.0;
#pragma unroll
        for (int j = 0; j < G; ++j) {
            lam += crow[j] * g.get(w, j);
        }
        double rISev_ICr = c.rISev_ICr0, rISev_D = c.rISev_D0;
        const double icu_lo = 0.90 * icu_cap;
        if (icu > icu_lo) {
            const double crit = age_rec[rCRIT];
            const double adj  = smoother_cosine(icu, icu_lo, icu_cap, crit, 0.0);
            rISev_ICr         = adj / age_rec[rT_ISEV];
            rISev_D           = (age_rec[rDPS] + (crit - adj)) / age_rec[rT_ISEV];
        }
        if (L.valid) {
            double* ax = a.aux + (L.rn * G + g.gl) * AUX;
            reinterpret_cast<double2*>(ax)[0] = make_double2((season * c.rP) * lam, rISev_ICr);
            reinterpret_cast<double2*>(ax)[1] = make_double2(rISev_D, 0.0);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Return (optimistic gather): y_out = y_in + sum(returning own commuters) - sum(returning visitors), edge order.
template <int G>
__global__ void __launch_bounds__(EDGE_THREADS) return_kernel(EdgeArgs a)
{
    extern __shared__ __align__(16) double smem[];
    __shared__ __align__(8) uint64_t mbar;
    const Geometry& geo = a.geo;
    EdgeLane<G> L;
    bool cta_active;
    edge_lane_setup<G>(geo, a.inc, a.num_runs, L, cta_active);
    if (!cta_active) {
        return;
    }
    stage_run_block(smem, a.st.run_block + (size_t)L.run * geo.block_stride, geo.block_stride * 8, &mbar);
    const Group<G>& g  = L.g;
    const double* mob  = geo.mob_max_pts > 0 ? smem + geo.mob_off : nullptr;
    const bool damped  = mob != nullptr && mob[0] > 0.0;
    const int ps_own   = a.st.node_pset ? a.st.node_pset[L.node] : 0;
    const double* pset = smem + ps_own * geo.pset_stride;

    double y[NC], snapk[NC], damp[NC];
    load10(a.y_in + L.rn * geo.n + g.gl * NC, y);
    load10(a.snap + L.rn * geo.n + g.gl * NC, snapk);
    const double fack = a.fac[L.rn * G + g.gl];
    mobility_damping_eval(damped ? mob : nullptr, geo.mob_max_pts, geo.n, g.gl, a.t_dep, damp);
    const double nd_own = nondetection_at(pset, a.t_dep);
    LaneCoef c_own;
    c_own.load(pset + PHDR + g.gl * AGE_REC);
    double slam_own, ricr_own, rd_own;
    {
        const double2* ax = reinterpret_cast<const double2*>(a.aux + (L.rn * G + g.gl) * AUX);
        double2 a0 = ax[0], a1 = ax[1];
        slam_own = a0.x;
        ricr_own = a0.y;
        rd_own   = a1.x;
    }
    const int beg = L.valid ? a.inc.inc_ptr[L.node] : 0;
    const int cnt = L.valid ? a.inc.inc_ptr[L.node + 1] - beg : 0;
    int max_cnt   = cnt;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        max_cnt = max(max_cnt, __shfl_xor_sync(0xffffffffu, max_cnt, o));
    }
    bool underflow = false;
    for (int it = 0; it < max_cnt; ++it) {
        const bool act = it < cnt;
        const int idx  = beg + (act ? it : 0);
        const int e    = act ? a.inc.inc_edge[idx] : 0;
        const int o    = act ? a.inc.inc_other[idx] : 0;
        const bool out = act ? (a.inc.inc_out[idx] != 0) : true;
        const size_t ro = (size_t)L.run * geo.N + o;
        double m[NC], mu[2];
        bool pos;
        if (out) {
            pos = damped ? mobile_at_departure<true>(a.inc.coeff + (size_t)e * geo.n + g.gl * NC, damp, snapk, fack,
                                                     nd_own, m, mu)
                         : mobile_at_departure<false>(a.inc.coeff + (size_t)e * geo.n + g.gl * NC, damp, snapk, fack,
                                                      nd_own, m, mu);
            // destination = other node: its return coefficients and (if parameter sets differ) its rates
            const double2* ax = reinterpret_cast<const double2*>(a.aux + (ro * G + g.gl) * AUX);
            double2 a0 = ax[0], a1 = ax[1];
            if (geo.num_psets > 1) {
                LaneCoef c_o;
                c_o.load(smem + a.st.node_pset[o] * geo.pset_stride + PHDR + g.gl * AGE_REC);
                euler_return(c_o, a0.x, a0.y, a1.x, a.dt, m);
            }
            else {
                euler_return(c_own, a0.x, a0.y, a1.x, a.dt, m);
            }
        }
        else {
            double so[NC];
            load10(a.snap + ro * geo.n + g.gl * NC, so);
            const double fo = a.fac[ro * G + g.gl];
            double nd       = nd_own;
            if (geo.num_psets > 1) {
                nd = nondetection_at(smem + a.st.node_pset[o] * geo.pset_stride, a.t_dep);
            }
            pos = damped ? mobile_at_departure<true>(a.inc.coeff + (size_t)e * geo.n + g.gl * NC, damp, so, fo, nd,
                                                     m, mu)
                         : mobile_at_departure<false>(a.inc.coeff + (size_t)e * geo.n + g.gl * NC, damp, so, fo, nd,
                                                      m, mu);
            euler_return(c_own, slam_own, ricr_own, rd_own, a.dt, m);
        }
        const bool any = g.any(act && pos);
        if (act && any) {
            if (out) {
#pragma unroll
                for (int i = 0; i < NC; ++i) {
                    y[i] += m[i];
                }
            }
            else {
#pragma unroll
                for (int i = 0; i < NC; ++i) {
                    y[i] -= m[i];
                    underflow = underflow || (y[i] < 0.0);
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
// Literal sequential replay of the return phase for flagged runs: one warp per run, lane = age group, edges in
// reference order, both end nodes updated in place in y_out (which is first reset to y_in).
template <int G>
__global__ void __launch_bounds__(32) return_exact_kernel(EdgeArgs a)
{
    extern __shared__ __align__(16) double smem[];
    __shared__ __align__(8) uint64_t mbar;
    const Geometry& geo = a.geo;
    const int run       = blockIdx.x;
    if (run >= a.num_runs || a.run_flags[run] == 0) {
        return;
    }
    stage_run_block(smem, a.st.run_block + (size_t)run * geo.block_stride, geo.block_stride * 8, &mbar);
    const int lane     = threadIdx.x;
    const size_t base  = (size_t)run * geo.N * geo.n;
    const size_t total = (size_t)geo.N * geo.n;
    for (size_t i = lane; i < total; i += 32) {
        a.y_out[base + i] = a.y_in[base + i];
    }
    __syncwarp();
    if (lane == 0) {
        a.run_flags[run] = 0;
        atomicAdd((unsigned long long*)a.stats, 1ull);
    }
    if (lane >= G) {
        return;
    }
    const int gl      = lane;
    const double* mob = geo.mob_max_pts > 0 ? smem + geo.mob_off : nullptr;
    const bool damped = mob != nullptr && mob[0] > 0.0;
    double damp[NC];
    mobility_damping_eval(damped ? mob : nullptr, geo.mob_max_pts, geo.n, gl, a.t_dep, damp);
    const unsigned gmask = (1u << G) - 1u;
    for (int e = 0; e < geo.E; ++e) {
        const int from   = a.inc.edge_start[e];
        const int to     = a.inc.edge_end[e];
        const size_t rf  = (size_t)run * geo.N + from;
        const size_t rt  = (size_t)run * geo.N + to;
        const int ps_f   = a.st.node_pset ? a.st.node_pset[from] : 0;
        const int ps_t   = a.st.node_pset ? a.st.node_pset[to] : 0;
        const double nd  = nondetection_at(smem + ps_f * geo.pset_stride, a.t_dep);
        double sf[NC], m[NC], mu[2];
        load10(a.snap + rf * geo.n + gl * NC, sf);
        const double ff = a.fac[rf * G + gl];
        const bool pos =
            damped ? mobile_at_departure<true>(a.inc.coeff + (size_t)e * geo.n + gl * NC, damp, sf, ff, nd, m, mu)
                   : mobile_at_departure<false>(a.inc.coeff + (size_t)e * geo.n + gl * NC, damp, sf, ff, nd, m, mu);
        const bool any = (__ballot_sync(gmask, pos) & gmask) != 0u;
        if (!any) {
            continue;
        }
        LaneCoef c_t;
        c_t.load(smem + ps_t * geo.pset_stride + PHDR + gl * AGE_REC);
        const double2* ax = reinterpret_cast<const double2*>(a.aux + (rt * G + gl) * AUX);
        double2 a0 = ax[0], a1 = ax[1];
        euler_return(c_t, a0.x, a0.y, a1.x, a.dt, m);
        double yt[NC], yf[NC], rem[NC];
        load10(a.y_out + rt * geo.n + gl * NC, yt);
        load10(a.y_out + rf * geo.n + gl * NC, yf);
        // underflow fix-up, instant.h:579-596 (per age group: j runs over this lane's compartments in order)
        int fixes = 0;
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            rem[i] = yt[i] - m[i];
        }
        int max_index = 0;
#pragma unroll
        for (int i = 1; i < NC; ++i) {
            if (rem[i] > rem[max_index]) {
                max_index = i;
            }
        }
#pragma unroll
        for (int j = 0; j < NC; ++j) {
            if (rem[j] < 0) {
#pragma unroll
                for (int q = 0; q < NC; ++q) { // register array: select instead of dynamic index
                    if (q == max_index) {
                        m[q] -= rem[j];
                    }
                }
                m[j] += rem[j];
                ++fixes;
            }
        }
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            yf[i] += m[i];
            yt[i] -= m[i];
        }
        store10(a.y_out + rf * geo.n + gl * NC, yf);
        store10(a.y_out + rt * geo.n + gl * NC, yt);
        if (fixes) {
            atomicAdd(a.st.counters + rt * 4 + 3, fixes);
        }
    }
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

template <int G>
static cudaError_t launch_edges_g(const EdgeArgs& a, int which, cudaStream_t stream)
{
    constexpr int NPW   = 32 / G;
    constexpr int WARPS = EDGE_THREADS / 32;
    const int chunks    = (a.geo.N + NPW * WARPS - 1) / (NPW * WARPS);
    const size_t grid   = (size_t)chunks * a.num_runs;
    const size_t shmem  = (size_t)a.geo.block_stride * 8;
    if (which == 0) {
        if (shmem > 48 * 1024) {
            cudaFuncSetAttribute(depart_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem);
        }
        depart_kernel<G><<<(unsigned)grid, EDGE_THREADS, shmem, stream>>>(a);
    }
    else if (which == 1) {
        if (shmem > 48 * 1024) {
            cudaFuncSetAttribute(return_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem);
        }
        return_kernel<G><<<(unsigned)grid, EDGE_THREADS, shmem, stream>>>(a);
    }
    else {
        if (shmem > 48 * 1024) {
            cudaFuncSetAttribute(return_exact_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem);
        }
        return_exact_kernel<G><<<(unsigned)a.num_runs, 32, shmem, stream>>>(a);
    }
    return cudaGetLastError();
}

static cudaError_t launch_edges(const EdgeArgs& a, int which, cudaStream_t stream)
{
    switch (a.geo.G) {
    case 1:
        return launch_edges_g<1>(a, which, stream);
    case 2:
        return launch_edges_g<2>(a, which, stream);
    case 3:
        return launch_edges_g<3>(a, which, stream);
    case 4:
        return launch_edges_g<4>(a, which, stream);
    case 5:
        return launch_edges_g<5>(a, which, stream);
    case 6:
        return launch_edges_g<6>(a, which, stream);
    case 7:
        return launch_edges_g<7>(a, which, stream);
    case 8:
        return launch_edges_g<8>(a, which, stream);
    default:
        return cudaErrorInvalidValue;
    }
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

cudaError_t launch_record_daily(const Geometry& geo, const double* y, double* daily, int day_index, int max_days,
                                int num_runs, cudaStream_t stream)
{
    const size_t total = (size_t)num_runs * geo.N * (geo.n / 2);
    const int threads  = 256;
    const size_t grid  = (total + threads - 1) / threads;
    record_daily_kernel<<<(unsigned)grid, threads, 0, stream>>>(geo.N, geo.n, y, daily, day_index, max_days, total);
    return cudaGetLastError();
}

} // namespace miocu
