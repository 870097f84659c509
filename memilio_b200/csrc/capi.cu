// capi.cu -- C-ABI (include/mio_cuda.h) over the CUDA kernels: handle management, run upload, the
// GraphSimulation::advance loop (cpp/memilio/mobility/graph_simulation.h:101-120) and result download.
#include "../../include/mio_cuda.h"
#include "kernels.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <numeric>
#include <string>
#include <chrono>
#include <mutex>
#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace miocu;

namespace
{

thread_local std::string g_last_error;

int fail(int code, const std::string& msg)
{
    g_last_error = msg;
    return code;
}

#define CUDA_TRY(expr)                                                                                                 \
    do {                                                                                                               \
        cudaError_t err__ = (expr);                                                                                    \
        if (err__ != cudaSuccess) {                                                                                    \
            return fail(MIO_UNKNOWN_ERROR, std::string(#expr) + ": " + cudaGetErrorString(err__));                     \
        }                                                                                                              \
    } while (0)

template <class T>
cudaError_t dev_alloc(T** p, size_t count)
{
    *p = nullptr;
    if (count == 0) {
        return cudaSuccess;
    }
    return cudaMalloc((void**)p, count * sizeof(T));
}

} // namespace

// shared with the host-only translation units (study.cpp)
int mio_set_last_error(int code, const std::string& msg)
{
    return fail(code, msg);
}

struct mio_ensemble {
    int device = 0;
    int G = 0, C = 0, N = 0, num_days = 0, mode = 0, lay = 1;
    int first_day = 0; // day window [first_day, first_day + num_days) of the record
    long long capacity = 0, num_columns = 0;
    std::atomic<long long> num_runs{0}; // handles of several host threads record into one ensemble
    double* d_data = nullptr; // [num_columns][capacity]
};

struct mio_handle {
    mio_graph_desc desc;
    mio_ensemble* ens      = nullptr;
    long long ens_offset   = 0;
    Geometry geo;
    int device       = 0;
    cudaStream_t stream = nullptr;
    int max_runs = 0, active_runs = 0;
    // graph (device)
    int *d_node_pset = nullptr, *d_node_ctable = nullptr;
    int *d_order_in = nullptr, *d_order_out = nullptr, *d_in_ptr = nullptr, *d_in_split = nullptr, *d_in_src = nullptr,
        *d_out_ptr = nullptr, *d_out_dst = nullptr, *d_edge_start = nullptr, *d_edge_end = nullptr;
    double *d_coeff = nullptr, *d_w_in = nullptr, *d_w_out = nullptr, *d_s_out = nullptr, *d_umask = nullptr;
    double* d_w_edge = nullptr; // factorized graphs: [E][W] weight per (edge, lane) in reference edge order
    double* d_xdamp = nullptr;
    int *d_xslot = nullptr, *d_xsrc = nullptr; // schedule of the block-parallel literal replay (see EdgeArgs)
    int xchunks = 0, xepp = 0;
    int factorized = 0;
    int num_mob = 0;
    int mob_comp[NC] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    // per run (device)
    double *d_run_block = nullptr, *d_node_local = nullptr, *d_node_ctl = nullptr;
    int* d_counters     = nullptr;
    double* d_y[4]      = {nullptr, nullptr, nullptr, nullptr};
    double *d_fac = nullptr, *d_aux = nullptr, *d_u = nullptr, *d_vis = nullptr;
    double *d_vacc_partial = nullptr, *d_vacc_full = nullptr;
    int* d_run_flags = nullptr;
    double* d_daily  = nullptr;
    int* d_traj_npts = nullptr;
    double *d_traj_times = nullptr, *d_traj_values = nullptr;
    // mio::FlowSimulation nodes: cumulative flows [R][N][G][16] (15 per age group + padding), their time series
    double *d_flows = nullptr, *d_traj_flows = nullptr;
    // simulation state (GraphSimulationBase::m_t and the edges' common toggle / stored mobile population)
    double t            = 0.0;
    bool toggle         = false; // MobilityEdge::m_return_mobile_population
    bool have_mobile    = false; // a mobile population is stored (between departure and return)
    double t_dep        = 0.0, return_time = 0.0;
    int cur = 0, snap_idx = -1, v0_idx = -1;
    int day0 = 0, days_recorded = 0;
    // statistics
    long long launches = 0;
    long long* d_stats = nullptr; // [0] exact replays of return steps
    bool aux_pending = false; // a departure happened and the next node phase has to write the return coefficients
    int* d_work = nullptr;        // [0] work item counter of the persistent node-phase kernel, [1] run list length
    int* d_run_list = nullptr;    // [R] runs flagged for the exact replay of a return step
    // get_mobility_results (optional)
    int mob_K = 0, mob_cap = 0, mob_npts = 0;
    std::vector<int32_t> mob_gptr, mob_gidx;
    std::vector<double> mob_times;
    std::vector<unsigned char> edge_active;
    int *d_mob_gptr = nullptr, *d_mob_gdev = nullptr, *d_ref_to_dev = nullptr;
    unsigned char* d_edge_active = nullptr;
    double *d_edge_m = nullptr, *d_mob_values = nullptr;
    // optional per-kernel-class device timing (CUDA events on the launching stream)
    bool profiling = false;
    struct Span {
        int cls;
        cudaEvent_t a, b;
    };
    std::vector<Span> spans;
    std::vector<cudaEvent_t> event_pool;
    double prof_ms[MIO_NUM_KERNEL_CLASSES]        = {0, 0, 0, 0, 0, 0};
    long long prof_launches[MIO_NUM_KERNEL_CLASSES] = {0, 0, 0, 0, 0, 0};
};

namespace
{

int age_stride(int model)
{
    return model == MIO_MODEL_OSECIR ? MIO_OSECIR_AGE_STRIDE : (model == MIO_MODEL_OSECIRVVS ? MIO_OSECIRVVS_AGE_STRIDE : -1);
}
int64_t pset_size(int model, int G)
{
    return MIO_PSET_HEADER + int64_t(age_stride(model)) * G;
}
int64_t ctable_size(int G, int max_loc, int max_pts)
{
    return 1 + max_loc + int64_t(max_loc) * (max_pts + int64_t(2 + max_pts) * G * G);
}
int64_t mobtable_size(int n, int mob_max_pts)
{
    return 1 + mob_max_pts + int64_t(mob_max_pts) * n;
}

// raw parameter set (mio_cuda.h layout) -> derived device record (device_common.cuh). The coefficient expressions
// keep the operand order of cpp/models/ode_secir/model.h:152-211 so that coef*y is bit-identical to the reference.
void derive_pset_osecir(const double* raw, int G, double* out)
{
    for (int k = 0; k < PHDR; ++k) {
        out[k] = 0.0;
    }
    out[hSTART_DAY]       = raw[MIO_PH_START_DAY];
    out[hSEASONALITY]     = raw[MIO_PH_SEASONALITY];
    out[hTNT_MAX_RISK]    = raw[MIO_PH_TNT_MAX_RISK];
    out[hNONDETECTION]    = raw[MIO_PH_COMMUTER_NONDETECTION];
    out[hSTART_DETECTION] = raw[MIO_PH_START_COMMUTER_DETECTION];
    out[hEND_DETECTION]   = raw[MIO_PH_END_COMMUTER_DETECTION];
    for (int i = 0; i < G; ++i) {
        const double* a    = raw + MIO_PSET_HEADER + i * MIO_OSECIR_AGE_STRIDE;
        double* r          = out + PHDR + i * AGE_REC;
        const double T_E   = a[MIO_PA_TIME_EXPOSED];
        const double T_INS = a[MIO_PA_TIME_INS];
        const double T_ISy = a[MIO_PA_TIME_ISY];
        const double T_ISev = a[MIO_PA_TIME_ISEV];
        const double T_ICr = a[MIO_PA_TIME_ICR];
        const double rho   = a[MIO_PA_REC_PER_INS];
        const double sev   = a[MIO_PA_SEV_PER_ISY];
        const double crit  = a[MIO_PA_CRIT_PER_SEV];
        const double dps   = a[MIO_PA_DEATHS_PER_SEV];
        const double dpc   = a[MIO_PA_DEATHS_PER_CRIT];
        r[rE_INS]          = 1.0 / T_E;
        r[rINS_ISy]        = (1 - rho) * (1 / T_INS);
        r[rINS_R]          = rho * (1.0 / T_INS);
        r[rISy_ISev]       = sev / T_ISy;
        r[rISy_R]          = (1 - sev) / T_ISy;
        r[rISev_ICr0]      = crit / T_ISev;
        r[rISev_R]         = (1.0 - crit - dps) / T_ISev;
        r[rISev_D0]        = (dps + (crit - crit)) / T_ISev;
        r[rICr_D]          = dpc / T_ICr;
        r[rICr_R]          = (1.0 - dpc) / T_ICr;
        r[rTNT]            = (1 - rho) / T_INS;
        r[rP]              = a[MIO_PA_TRANS_PROB];
        r[rREL]            = a[MIO_PA_REL_TRANS_NOSYMP];
        r[rRISK]           = a[MIO_PA_RISK_SYMP];
        r[rMAXRISK]        = a[MIO_PA_MAX_RISK_SYMP];
        r[rCRIT]           = crit;
        r[rT_ISEV]         = T_ISev;
        r[rDPS]            = dps;
        r[rREDEXP]         = 1.0;
        r[rINFNV]          = 1.0;
    }
}

// osecirvvs: one record per (immunity layer, age group) lane, layer-major. Rate expressions in the operand order of
// cpp/models/ode_secirvvs/model.h:266-376 (the ICU-dependent ones at "ICU not short", adjusted on the device).
void derive_pset_osecirvvs(const double* raw, int G, double* out)
{
    for (int k = 0; k < PHDR; ++k) {
        out[k] = 0.0;
    }
    out[hSTART_DAY]             = raw[MIO_PH_START_DAY];
    out[hSEASONALITY]           = raw[MIO_PH_SEASONALITY];
    out[hTNT_MAX_RISK]          = raw[MIO_PH_TNT_MAX_RISK];
    out[hNONDETECTION]          = raw[MIO_PH_COMMUTER_NONDETECTION];
    out[hSTART_DETECTION]       = raw[MIO_PH_START_COMMUTER_DETECTION];
    out[hEND_DETECTION]         = raw[MIO_PH_END_COMMUTER_DETECTION];
    out[hTNT_MAX_RISK_NOSYMP]   = raw[MIO_PH_TNT_MAX_RISK_NOSYMP];
    out[hSTART_DAY_NEW_VARIANT] = raw[MIO_PH_START_DAY_NEW_VARIANT];
    for (int layer = 0; layer < VVS_LAY; ++layer) {
        for (int i = 0; i < G; ++i) {
            const double* a     = raw + MIO_PSET_HEADER + i * MIO_OSECIRVVS_AGE_STRIDE;
            double* r           = out + PHDR + (layer * G + i) * AGE_REC;
            const double T_E    = a[MIO_PA_TIME_EXPOSED];
            const double T_INS  = a[MIO_PA_TIME_INS];
            const double T_ISy  = a[MIO_PA_TIME_ISY];
            const double T_ISev = a[MIO_PA_TIME_ISEV];
            const double T_ICr  = a[MIO_PA_TIME_ICR];
            const double rho    = a[MIO_PA_REC_PER_INS];
            const double sev    = a[MIO_PA_SEV_PER_ISY];
            const double crit   = a[MIO_PA_CRIT_PER_SEV];
            const double dps    = a[MIO_PA_DEATHS_PER_SEV];
            const double dpc    = a[MIO_PA_DEATHS_PER_CRIT];
            const double rtm    = a[MIO_PV_REDUC_TIME_MILD];
            r[rE_INS]           = 1.0 / T_E;
            if (layer == 0) {
                r[rINS_ISy]   = (1 - rho) / T_INS;
                r[rINS_R]     = rho / T_INS;
                r[rISy_ISev]  = sev / T_ISy;
                r[rISy_R]     = (1 - sev) / T_ISy;
                r[rISev_ICr0] = crit / T_ISev;
                r[rISev_R]    = (1 - crit - dps) / T_ISev;
                r[rISev_D0]   = (dps + (crit - crit)) / T_ISev;
                r[rICr_D]     = dpc / T_ICr;
                r[rICr_R]     = (1 - dpc) / T_ICr;
                r[rREDEXP]    = 1.0;
            }
            else {
                const double red_exp  = a[layer == 1 ? MIO_PV_REDUC_EXPOSED_PI : MIO_PV_REDUC_EXPOSED_II];
                const double red_symp = a[layer == 1 ? MIO_PV_REDUC_INF_SYMP_PI : MIO_PV_REDUC_INF_SYMP_II];
                const double red_sev  = a[layer == 1 ? MIO_PV_REDUC_SEV_CRIT_DEAD_PI : MIO_PV_REDUC_SEV_CRIT_DEAD_II];
                r[rINS_ISy]   = (red_symp / red_exp) * (1 - rho) / (T_INS * rtm);
                r[rINS_R]     = (1 - (red_symp / red_exp) * (1 - rho)) / (T_INS * rtm);
                r[rISy_ISev]  = red_sev / red_symp * sev / (T_ISy * rtm);
                r[rISy_R]     = (1 - (red_sev / red_symp) * sev) / (T_ISy * rtm);
                r[rISev_ICr0] = red_sev / red_sev * crit / T_ISev;
                r[rISev_R]    = (1 - (red_sev / red_sev) * (crit + dps)) / T_ISev;
                r[rISev_D0]   = (red_sev / red_sev) * (dps + (crit - crit)) / T_ISev;
                r[rICr_D]     = (red_sev / red_sev) * dpc / T_ICr;
                r[rICr_R]     = (1 - (red_sev / red_sev) * dpc) / T_ICr;
                r[rREDEXP]    = red_exp;
            }
            r[rTNT]     = (1 - rho) / T_INS;
            r[rP]       = a[MIO_PA_TRANS_PROB];
            r[rREL]     = a[MIO_PA_REL_TRANS_NOSYMP];
            r[rRISK]    = a[MIO_PA_RISK_SYMP];
            r[rMAXRISK] = a[MIO_PA_MAX_RISK_SYMP];
            r[rCRIT]    = crit;
            r[rT_ISEV]  = T_ISev;
            r[rDPS]     = dps;
            r[rINFNV]   = a[MIO_PV_INFECTIOUSNESS_NEW_VARIANT];
        }
    }
}

void derive_pset(int model, const double* raw, int G, double* out)
{
    if (model == MIO_MODEL_OSECIRVVS) {
        derive_pset_osecirvvs(raw, G, out);
    }
    else {
        derive_pset_osecir(raw, G, out);
    }
}

// reference element order (age*C + compartment) -> device order [LAY][G][NC] of one node-sized vector; the unused
// slots get `pad`
void to_device_layout(int lay, int G, int C_ref, const double* ref, double* dev, double pad = 0.0)
{
    const int W = G * lay;
    std::fill(dev, dev + (size_t)W * NC, pad);
    for (int i = 0; i < G; ++i) {
        for (int c = 0; c < C_ref; ++c) {
            dev[dev_index_of_ref(lay, G, i, c)] = ref[i * C_ref + c];
        }
    }
}

void destroy(mio_handle* h)
{
    if (!h) {
        return;
    }
    cudaSetDevice(h->device);
    void* ptrs[] = {h->d_node_pset,  h->d_node_ctable, h->d_order_in,   h->d_order_out, h->d_in_ptr,
                    h->d_in_split,   h->d_in_src,      h->d_out_ptr,    h->d_out_dst,   h->d_w_in,
                    h->d_w_out,      h->d_s_out,       h->d_umask,      h->d_u,         h->d_vis,
                    h->d_edge_start, h->d_edge_end,    h->d_coeff,
                    h->d_run_block,  h->d_node_local,  h->d_node_ctl,   h->d_counters,
                    h->d_y[0],       h->d_y[1],        h->d_y[2],       h->d_y[3],      h->d_fac,
                    h->d_aux,        h->d_run_flags,   h->d_daily,      h->d_traj_npts, h->d_traj_times,
                    h->d_traj_values, h->d_stats,      h->d_vacc_partial, h->d_vacc_full, h->d_work, h->d_run_list, h->d_w_edge,
                    h->d_mob_gptr,   h->d_mob_gdev,    h->d_ref_to_dev, h->d_edge_active, h->d_edge_m, h->d_mob_values,
                    h->d_flows,      h->d_traj_flows,  h->d_xslot,      h->d_xsrc,        h->d_xdamp};
    for (void* p : ptrs) {
        if (p) {
            cudaFree(p);
        }
    }
    for (auto& sp : h->spans) {
        cudaEventDestroy(sp.a);
        cudaEventDestroy(sp.b);
    }
    for (auto e : h->event_pool) {
        cudaEventDestroy(e);
    }
    if (h->stream) {
        cudaStreamDestroy(h->stream);
    }
    delete h;
}

DeviceState device_state(const mio_handle* h)
{
    DeviceState st;
    st.run_block   = h->d_run_block;
    st.node_local  = h->d_node_local;
    st.node_pset   = h->d_node_pset;
    st.node_ctable = h->d_node_ctable;
    st.node_ctl    = h->d_node_ctl;
    st.counters    = h->d_counters;
    st.vacc_partial = h->d_vacc_partial;
    st.vacc_full    = h->d_vacc_full;
    return st;
}

cudaEvent_t get_event(mio_handle* h)
{
    if (!h->event_pool.empty()) {
        cudaEvent_t e = h->event_pool.back();
        h->event_pool.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}

struct ScopedSpan {
    mio_handle* h;
    int cls;
    cudaEvent_t a = nullptr;
    ScopedSpan(mio_handle* h_, int cls_)
        : h(h_)
        , cls(cls_)
    {
        ++h->prof_launches[cls];
        ++h->launches;
        if (h->profiling) {
            a = get_event(h);
            cudaEventRecord(a, h->stream);
        }
    }
    ~ScopedSpan()
    {
        if (h->profiling) {
            cudaEvent_t b = get_event(h);
            cudaEventRecord(b, h->stream);
            h->spans.push_back({cls, a, b});
        }
    }
};

void collect_spans(mio_handle* h)
{
    for (auto& s : h->spans) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, s.a, s.b) == cudaSuccess) {
            h->prof_ms[s.cls] += ms;
        }
        h->event_pool.push_back(s.a);
        h->event_pool.push_back(s.b);
    }
    h->spans.clear();
}

int free_buffer(const mio_handle* h, int a, int b, int c)
{
    for (int i = 0; i < 4; ++i) {
        if (i != a && i != b && i != c && i != h->snap_idx && i != h->v0_idx) {
            return i;
        }
    }
    return -1;
}

} // namespace

extern "C" {

const char* mio_cuda_last_error_string(void)
{
    return g_last_error.c_str();
}

int mio_cuda_version(void)
{
    return 100;
}

int mio_cuda_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        return 0;
    }
    return n;
}

int mio_num_compartments(int model)
{
    return model == MIO_MODEL_OSECIR ? 10 : (model == MIO_MODEL_OSECIRVVS ? 27 : -1);
}
int mio_pset_age_stride(int model)
{
    return age_stride(model);
}
int64_t mio_pset_size(int model, int num_groups)
{
    return pset_size(model, num_groups);
}
int64_t mio_ctable_size(int num_groups, int ct_max_loc, int ct_max_pts)
{
    return ctable_size(num_groups, ct_max_loc, ct_max_pts);
}
int64_t mio_mobtable_size(int model, int num_groups, int mob_max_pts)
{
    return mobtable_size(num_groups * mio_num_compartments(model), mob_max_pts);
}

int mio_cuda_graph_create(const mio_graph_desc* d, int device, mio_handle** out)
{
    if (!d || !out) {
        return fail(MIO_INVALID_VALUE, "null argument");
    }
    *out = nullptr;
    if (d->model != MIO_MODEL_OSECIR && d->model != MIO_MODEL_OSECIRVVS) {
        return fail(MIO_INVALID_VALUE, "unknown model kind");
    }
    if (d->model == MIO_MODEL_OSECIRVVS && d->sim_kind != MIO_SIM_BASE && d->sim_kind != MIO_SIM_MODEL) {
        return fail(MIO_INVALID_VALUE, "unknown node simulation kind");
    }
    if (d->model == MIO_MODEL_OSECIRVVS && d->vacc_days > 0 && (!d->vacc_partial || !d->vacc_full)) {
        return fail(MIO_INVALID_VALUE, "vaccination arrays missing");
    }
    if (d->num_groups < 1 || d->num_groups > MIO_MAX_GROUPS || d->num_nodes < 1 || d->num_edges < 0 ||
        d->max_runs < 1 || d->num_psets < 1 || d->num_ctables < 1 || d->ct_max_loc < 1 || d->ct_max_pts < 0 ||
        d->mob_max_pts < 0) {
        return fail(MIO_INVALID_VALUE, "invalid graph descriptor sizes");
    }
    if (!(d->dt_graph > 0.0)) {
        return fail(MIO_INVALID_VALUE, "dt_graph must be positive");
    }
    if (d->integrator.kind < 0 || d->integrator.kind > 2) {
        return fail(MIO_INVALID_VALUE, "unknown integrator kind");
    }
    if (d->num_edges > 0 && (!d->edge_start || !d->edge_end || !d->edge_coeff)) {
        return fail(MIO_INVALID_VALUE, "edge arrays missing");
    }
    if (d->flow_simulation && (d->model != MIO_MODEL_OSECIR || d->integrator.kind == MIO_INTEGRATOR_RKF45)) {
        return fail(MIO_INVALID_VALUE, "FlowSimulation nodes: ODE-SECIR with the Cash-Karp controller or Euler only");
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        return fail(MIO_UNKNOWN_ERROR, "no CUDA device available (this library has no CPU fallback)");
    }
    if (device < 0 || device >= ndev) {
        return fail(MIO_OUT_OF_RANGE, "device index out of range");
    }
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        return fail(MIO_UNKNOWN_ERROR, "device is not sm_100 (kernels are built for sm_100a only)");
    }
    const int G = d->num_groups, N = d->num_nodes, E = d->num_edges;
    const int LAY = d->model == MIO_MODEL_OSECIRVVS ? VVS_LAY : 1;
    const int W = G * LAY, n = W * NC;
    const int C_ref = mio_num_compartments(d->model), n_ref = G * C_ref;
    for (int e = 0; e < E; ++e) {
        if (d->edge_start[e] < 0 || d->edge_start[e] >= N || d->edge_end[e] < 0 || d->edge_end[e] >= N ||
            d->edge_start[e] == d->edge_end[e]) {
            return fail(MIO_OUT_OF_RANGE, "edge endpoint out of range");
        }
        if (e > 0 && !(d->edge_start[e - 1] < d->edge_start[e] ||
                       (d->edge_start[e - 1] == d->edge_start[e] && d->edge_end[e - 1] < d->edge_end[e]))) {
            return fail(MIO_INVALID_VALUE, "edges must be sorted by (start,end) like Graph::add_edge keeps them");
        }
    }
    for (int k = 0; k < N; ++k) {
        if ((d->node_pset && (d->node_pset[k] < 0 || d->node_pset[k] >= d->num_psets)) ||
            (d->node_ctable && (d->node_ctable[k] < 0 || d->node_ctable[k] >= d->num_ctables))) {
            return fail(MIO_OUT_OF_RANGE, "node parameter-set / contact-table index out of range");
        }
    }
    auto* h     = new mio_handle();
    h->desc     = *d;
    h->device   = device;
    h->max_runs = h->active_runs = d->max_runs;
    Geometry& geo    = h->geo;
    geo.G            = G;
    geo.N            = N;
    geo.E            = E;
    geo.n            = n;
    geo.LAY          = LAY;
    geo.W            = W;
    geo.C_ref        = C_ref;
    geo.n_ref        = n_ref;
    geo.sim_kind     = d->model == MIO_MODEL_OSECIRVVS ? d->sim_kind : MIO_SIM_MODEL;
    geo.vacc_days    = d->model == MIO_MODEL_OSECIRVVS ? d->vacc_days : 0;
    geo.num_psets    = d->num_psets;
    geo.num_ctables  = d->num_ctables;
    geo.ct_max_loc   = d->ct_max_loc;
    geo.ct_max_pts   = d->ct_max_pts;
    geo.mob_max_pts  = d->mob_max_pts;
    geo.pset_stride  = PHDR + AGE_REC * W;
    geo.ctable_stride = (int)ctable_size(G, d->ct_max_loc, d->ct_max_pts);
    geo.ct_off       = geo.num_psets * geo.pset_stride;
    geo.mob_off      = geo.ct_off + geo.num_ctables * geo.ctable_stride;
    int total        = geo.mob_off + (d->mob_max_pts > 0 ? (int)mobtable_size(n, d->mob_max_pts) : 0);
    geo.block_stride = (total + 1) & ~1; // cp.async.bulk wants multiples of 16 bytes
    if ((size_t)geo.block_stride * 8 > 200 * 1024) {
        delete h;
        return fail(MIO_OUT_OF_RANGE, "per-run parameter block exceeds the shared memory of one CTA");
    }
    h->t    = d->t0;
    h->day0 = (int)std::ceil(d->t0 - 1e-14);

#define H_TRY(expr)                                                                                                    \
    do {                                                                                                               \
        cudaError_t err__ = (expr);                                                                                    \
        if (err__ != cudaSuccess) {                                                                                    \
            std::string m = std::string(#expr) + ": " + cudaGetErrorString(err__);                                     \
            destroy(h);                                                                                                \
            return fail(MIO_UNKNOWN_ERROR, m);                                                                         \
        }                                                                                                              \
    } while (0)

    H_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    // graph
    if (d->node_pset) {
        H_TRY(dev_alloc(&h->d_node_pset, (size_t)N));
        H_TRY(cudaMemcpy(h->d_node_pset, d->node_pset, sizeof(int) * N, cudaMemcpyHostToDevice));
    }
    if (d->node_ctable) {
        H_TRY(dev_alloc(&h->d_node_ctable, (size_t)N));
        H_TRY(cudaMemcpy(h->d_node_ctable, d->node_ctable, sizeof(int) * N, cudaMemcpyHostToDevice));
    }
    if (geo.vacc_days > 0) {
        const size_t cnt = (size_t)N * G * geo.vacc_days;
        H_TRY(dev_alloc(&h->d_vacc_partial, cnt));
        H_TRY(dev_alloc(&h->d_vacc_full, cnt));
        H_TRY(cudaMemcpy(h->d_vacc_partial, d->vacc_partial, sizeof(double) * cnt, cudaMemcpyHostToDevice));
        H_TRY(cudaMemcpy(h->d_vacc_full, d->vacc_full, sizeof(double) * cnt, cudaMemcpyHostToDevice));
    }
    if (E > 0) {
        for (size_t i = 0; i < (size_t)E * n_ref; ++i) {
            if (!(d->edge_coeff[i] >= 0.0)) {
                destroy(h);
                return fail(MIO_INVALID_VALUE, "mobility coefficients must be non-negative");
            }
        }
        // coefficient vectors in the device's lane order (identity for osecir)
        std::vector<double> coeff_dev((size_t)E * n);
        h->edge_active.assign((size_t)E, 0);
        for (int e = 0; e < E; ++e) {
            to_device_layout(LAY, G, C_ref, d->edge_coeff + (size_t)e * n_ref, coeff_dev.data() + (size_t)e * n);
            for (int j = 0; j < n_ref; ++j) {
                if (d->edge_coeff[(size_t)e * n_ref + j] > 0.0) {
                    h->edge_active[(size_t)e] = 1;
                }
            }
        }
        const double* edge_coeff = coeff_dev.data();
        // outgoing CSR = reference edge order; incoming CSR sorted by (dst, src)
        std::vector<int> out_ptr(N + 1, 0), in_ptr(N + 1, 0), in_split(N, 0);
        for (int e = 0; e < E; ++e) {
            ++out_ptr[d->edge_start[e] + 1];
            ++in_ptr[d->edge_end[e] + 1];
            if (d->edge_start[e] < d->edge_end[e]) {
                ++in_split[d->edge_end[e]];
            }
        }
        for (int k = 0; k < N; ++k) {
            out_ptr[k + 1] += out_ptr[k];
            in_ptr[k + 1] += in_ptr[k];
        }
        std::vector<int> fill(in_ptr.begin(), in_ptr.end() - 1), in_src(E), in_edge(E);
        for (int e = 0; e < E; ++e) { // ascending (src, dst) => every destination's list is sorted by src
            const int t  = d->edge_end[e];
            in_src[fill[t]]  = d->edge_start[e];
            in_edge[fill[t]] = e;
            ++fill[t];
        }
        // factorized form: coeff[e][l][c] == w[e][l] * mask[l][c] with a 0/1 mask, exactly (l = lane of the node)
        std::vector<double> umask(n, 0.0), w_age((size_t)E * W, 0.0);
        for (int e = 0; e < E; ++e) {
            for (int j = 0; j < n; ++j) {
                if (edge_coeff[(size_t)e * n + j] != 0.0) {
                    umask[j] = 1.0;
                }
            }
        }
        bool fact = true;
        for (int e = 0; e < E && fact; ++e) {
            for (int i = 0; i < W && fact; ++i) {
                double w = 0.0;
                for (int c = 0; c < NC; ++c) {
                    if (umask[i * NC + c] != 0.0) {
                        w = edge_coeff[(size_t)e * n + i * NC + c];
                        break;
                    }
                }
                for (int c = 0; c < NC; ++c) {
                    if (edge_coeff[(size_t)e * n + i * NC + c] != w * umask[i * NC + c]) {
                        fact = false;
                    }
                }
                w_age[(size_t)e * W + i] = w;
            }
        }
        h->factorized = fact ? 1 : 0;
        if (!fact) {
            std::fill(umask.begin(), umask.end(), 1.0);
        }
        // mobile compartments = union of the mask over the lanes; the list is padded with a non-mobile
        // compartment (its emigration base entries are zero) or, if all ten are mobile, is simply complete
        {
            int pad = -1;
            for (int cc = 0; cc < NC; ++cc) {
                bool any = false;
                for (int i = 0; i < W; ++i) {
                    any = any || umask[i * NC + cc] != 0.0;
                }
                if (any) {
                    h->mob_comp[h->num_mob++] = cc;
                }
                else if (pad < 0) {
                    pad = cc;
                }
            }
            for (int m = h->num_mob; m < NC; ++m) {
                h->mob_comp[m] = pad < 0 ? 0 : pad;
            }
        }
        std::vector<double> w_in(fact ? (size_t)E * W : (size_t)E * n), w_out((size_t)E * W * 2),
            s_out((size_t)N * n, 0.0);
        for (int pos = 0; pos < E; ++pos) {
            const int e = in_edge[pos];
            if (fact) {
                std::copy(w_age.begin() + (size_t)e * W, w_age.begin() + (size_t)(e + 1) * W,
                          w_in.begin() + (size_t)pos * W);
            }
            else {
                std::copy(edge_coeff + (size_t)e * n, edge_coeff + (size_t)(e + 1) * n,
                          w_in.begin() + (size_t)pos * n);
            }
        }
        for (int e = 0; e < E; ++e) {
            for (int i = 0; i < W; ++i) {
                w_out[((size_t)e * W + i) * 2 + 0] = edge_coeff[(size_t)e * n + i * NC + cS];
                w_out[((size_t)e * W + i) * 2 + 1] = edge_coeff[(size_t)e * n + i * NC + cISev];
            }
            for (int j = 0; j < n; ++j) {
                s_out[(size_t)d->edge_start[e] * n + j] += edge_coeff[(size_t)e * n + j];
            }
        }
        std::vector<int> order_in(N), order_out(N);
        std::iota(order_in.begin(), order_in.end(), 0);
        std::iota(order_out.begin(), order_out.end(), 0);
        std::stable_sort(order_in.begin(), order_in.end(), [&](int a, int b) {
            return in_ptr[a + 1] - in_ptr[a] > in_ptr[b + 1] - in_ptr[b];
        });
        std::stable_sort(order_out.begin(), order_out.end(), [&](int a, int b) {
            return out_ptr[a + 1] - out_ptr[a] > out_ptr[b + 1] - out_ptr[b];
        });
        auto up_i = [&](int** dst, const std::vector<int>& v) {
            cudaError_t er = dev_alloc(dst, v.size());
            return er != cudaSuccess ? er : cudaMemcpy(*dst, v.data(), sizeof(int) * v.size(), cudaMemcpyHostToDevice);
        };
        auto up_d = [&](double** dst, const double* v, size_t cnt) {
            cudaError_t er = dev_alloc(dst, cnt);
            return er != cudaSuccess ? er : cudaMemcpy(*dst, v, sizeof(double) * cnt, cudaMemcpyHostToDevice);
        };
        std::vector<int> es(d->edge_start, d->edge_start + E), ee(d->edge_end, d->edge_end + E);
        H_TRY(up_i(&h->d_order_in, order_in));
        H_TRY(up_i(&h->d_order_out, order_out));
        H_TRY(up_i(&h->d_in_ptr, in_ptr));
        H_TRY(up_i(&h->d_in_split, in_split));
        H_TRY(up_i(&h->d_in_src, in_src));
        H_TRY(up_i(&h->d_out_ptr, out_ptr));
        H_TRY(up_i(&h->d_out_dst, ee));
        H_TRY(up_i(&h->d_edge_start, es));
        H_TRY(up_i(&h->d_edge_end, ee));
        H_TRY(up_d(&h->d_coeff, edge_coeff, (size_t)E * n));
        H_TRY(up_d(&h->d_w_in, w_in.data(), w_in.size()));
        H_TRY(up_d(&h->d_w_out, w_out.data(), w_out.size()));
        if (fact) {
            H_TRY(up_d(&h->d_w_edge, w_age.data(), w_age.size()));
        }
        H_TRY(up_d(&h->d_s_out, s_out.data(), s_out.size()));
        {
            // chunks of the block-parallel literal replay: consecutive edges of one source, distinct destinations
            const int epp = EXB_P1_WARPS * (32 / W);
            std::vector<int> xslot, xsrc;
            int e = 0;
            while (e < E) {
                const int src = d->edge_start[e];
                std::vector<int> seen;
                const size_t base = xslot.size();
                xslot.resize(base + (size_t)epp * 2, -1);
                int q = 0;
                while (e < E && q < epp && d->edge_start[e] == src &&
                       std::find(seen.begin(), seen.end(), d->edge_end[e]) == seen.end()) {
                    xslot[base + 2 * q]     = e;
                    xslot[base + 2 * q + 1] = d->edge_end[e];
                    seen.push_back(d->edge_end[e]);
                    ++q;
                    ++e;
                }
                for (; q < epp; ++q) {
                    xslot[base + 2 * q + 1] = 0;
                }
                const bool hazard = !xsrc.empty() && std::find(seen.begin(), seen.end(), xsrc[xsrc.size() - 2]) != seen.end();
                xsrc.push_back(src);
                xsrc.push_back((int)seen.size() | (hazard ? 1 << 16 : 0));
            }
            h->xchunks = (int)xsrc.size() / 2;
            h->xepp    = epp;
            H_TRY(up_i(&h->d_xslot, xslot));
            H_TRY(up_i(&h->d_xsrc, xsrc));
        }
        H_TRY(up_d(&h->d_umask, umask.data(), umask.size()));
    }
    // per-run state
    const size_t R = (size_t)d->max_runs;
    H_TRY(dev_alloc(&h->d_run_block, R * geo.block_stride));
    H_TRY(dev_alloc(&h->d_node_local, R * N * 2));
    H_TRY(dev_alloc(&h->d_node_ctl, R * N * 4));
    H_TRY(dev_alloc(&h->d_counters, R * N * 4));
    H_TRY(dev_alloc(&h->d_work, (size_t)4));
    const int nbuf = E > 0 ? 4 : 2;
    for (int b = 0; b < nbuf; ++b) {
        H_TRY(dev_alloc(&h->d_y[b], R * N * n));
    }
    if (E > 0) {
        H_TRY(dev_alloc(&h->d_fac, R * N * W));
        H_TRY(dev_alloc(&h->d_aux, R * N * W * 4));
        H_TRY(dev_alloc(&h->d_u, R * N * n));
        H_TRY(dev_alloc(&h->d_vis, R * N * n));
        H_TRY(dev_alloc(&h->d_run_flags, R));
        H_TRY(dev_alloc(&h->d_run_list, R));
        H_TRY(dev_alloc(&h->d_xdamp, R * n));
        H_TRY(cudaMemset(h->d_run_flags, 0, sizeof(int) * R));
        H_TRY(dev_alloc(&h->d_stats, (size_t)4));
        H_TRY(cudaMemset(h->d_stats, 0, sizeof(long long) * 4));
    }
    if (d->record_daily) {
        if (d->max_days < 1) {
            destroy(h);
            return fail(MIO_INVALID_VALUE, "record_daily needs max_days >= 1");
        }
        H_TRY(dev_alloc(&h->d_daily, R * N * (size_t)d->max_days * n_ref));
    }
    if (d->traj_capacity > 0) {
        H_TRY(dev_alloc(&h->d_traj_npts, R * N));
        H_TRY(dev_alloc(&h->d_traj_times, R * N * (size_t)d->traj_capacity));
        H_TRY(dev_alloc(&h->d_traj_values, R * N * (size_t)d->traj_capacity * n_ref));
    }
    if (d->flow_simulation) {
        H_TRY(dev_alloc(&h->d_flows, R * N * W * 16));
        H_TRY(cudaMemset(h->d_flows, 0, sizeof(double) * R * N * W * 16));
        if (d->traj_capacity > 0) {
            // time point 0 of the flow series is zero (FlowModel::get_initial_flows, flow_model.h:146-149)
            const size_t cnt = R * N * (size_t)d->traj_capacity * G * 15;
            H_TRY(dev_alloc(&h->d_traj_flows, cnt));
            H_TRY(cudaMemset(h->d_traj_flows, 0, sizeof(double) * cnt));
        }
    }
    H_TRY(cudaMemset(h->d_counters, 0, sizeof(int) * R * N * 4));
#undef H_TRY
    // the descriptor's host pointers are not valid after this call
    h->desc.edge_start = h->desc.edge_end = nullptr;
    h->desc.edge_coeff                    = nullptr;
    h->desc.node_pset = h->desc.node_ctable = nullptr;
    h->desc.vacc_partial = h->desc.vacc_full = nullptr;
    *out = h;
    return MIO_OK;
}

void mio_cuda_graph_destroy(mio_handle* h)
{
    destroy(h);
}

int mio_cuda_set_runs(mio_handle* h, int32_t slot0, int32_t count, const mio_run_data* data)
{
    if (!h || !data || !data->psets || !data->ctables || !data->node_local || !data->y0) {
        return fail(MIO_INVALID_VALUE, "null argument");
    }
    if (slot0 < 0 || count < 0 || slot0 + count > h->max_runs) {
        return fail(MIO_OUT_OF_RANGE, "run slots out of range");
    }
    if (h->geo.mob_max_pts > 0 && !data->mobtables) {
        return fail(MIO_INVALID_VALUE, "graph was created with mob_max_pts > 0 but no mobility tables were given");
    }
    if (count == 0) {
        return MIO_OK;
    }
    CUDA_TRY(cudaSetDevice(h->device));
    const Geometry& geo = h->geo;
    const int G = geo.G, N = geo.N, n = geo.n, n_ref = geo.n_ref;
    const int64_t psz = pset_size(h->desc.model, G);
    const int64_t csz = geo.ctable_stride;
    const int64_t msz     = geo.mob_max_pts > 0 ? mobtable_size(n, geo.mob_max_pts) : 0;     // device layout
    const int64_t msz_ref = geo.mob_max_pts > 0 ? mobtable_size(n_ref, geo.mob_max_pts) : 0; // caller's layout
    // build the device run blocks: derived parameter sets | contact tables | mobility table
    std::vector<double> block((size_t)count * geo.block_stride, 0.0);
    for (int r = 0; r < count; ++r) {
        double* b = block.data() + (size_t)r * geo.block_stride;
        for (int p = 0; p < geo.num_psets; ++p) {
            derive_pset(h->desc.model, data->psets + ((size_t)r * geo.num_psets + p) * psz, G,
                        b + p * geo.pset_stride);
        }
        std::memcpy(b + geo.ct_off, data->ctables + (size_t)r * geo.num_ctables * csz,
                    sizeof(double) * geo.num_ctables * csz);
        for (int c = 0; c < geo.num_ctables; ++c) {
            const double* ct = b + geo.ct_off + c * csz;
            const int L      = (int)ct[0];
            if (L < 1 || L > geo.ct_max_loc) {
                return fail(MIO_INVALID_VALUE, "contact table: number of locations out of range");
            }
            for (int l = 0; l < L; ++l) {
                if (ct[1 + l] < 0 || ct[1 + l] > geo.ct_max_pts) {
                    return fail(MIO_INVALID_VALUE, "contact table: number of change points out of range");
                }
            }
        }
        if (msz > 0) {
            const double* mt = data->mobtables + (size_t)r * msz_ref;
            if (mt[0] < 0 || mt[0] > geo.mob_max_pts) {
                return fail(MIO_INVALID_VALUE, "mobility table: number of change points out of range");
            }
            if (geo.LAY == 1) {
                std::memcpy(b + geo.mob_off, mt, sizeof(double) * msz);
            }
            else {
                double* md = b + geo.mob_off;
                std::memcpy(md, mt, sizeof(double) * (1 + geo.mob_max_pts));
                for (int k = 0; k < geo.mob_max_pts; ++k) {
                    to_device_layout(geo.LAY, G, geo.C_ref, mt + 1 + geo.mob_max_pts + (size_t)k * n_ref,
                                     md + 1 + geo.mob_max_pts + (size_t)k * n);
                }
            }
        }
    }
    const size_t s0 = (size_t)slot0;
    CUDA_TRY(cudaMemcpyAsync(h->d_run_block + s0 * geo.block_stride, block.data(), sizeof(double) * block.size(),
                             cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaMemcpyAsync(h->d_node_local + s0 * N * 2, data->node_local, sizeof(double) * (size_t)count * N * 2,
                             cudaMemcpyHostToDevice, h->stream));
    if (geo.LAY == 1) {
        CUDA_TRY(cudaMemcpyAsync(h->d_y[h->cur] + s0 * N * n, data->y0, sizeof(double) * (size_t)count * N * n,
                                 cudaMemcpyHostToDevice, h->stream));
    }
    else {
        // staged in a free state buffer (n >= n_ref), then permuted into the device's lane order
        const int stage = free_buffer(h, h->cur, -1, -1);
        if (stage < 0 || !h->d_y[stage]) {
            return fail(MIO_UNKNOWN_ERROR, "internal: no free state buffer");
        }
        CUDA_TRY(cudaMemcpyAsync(h->d_y[stage], data->y0, sizeof(double) * (size_t)count * N * n_ref,
                                 cudaMemcpyHostToDevice, h->stream));
        CUDA_TRY(launch_permute_state(geo, h->d_y[stage], h->d_y[h->cur] + s0 * N * n, (size_t)count * N, 1,
                                      h->stream));
        ++h->launches;
    }
    std::vector<double> tn((size_t)count * N, h->desc.t0), dtn((size_t)count * N, 0.1);
    if (data->dt0) {
        std::copy(data->dt0, data->dt0 + (size_t)count * N, dtn.begin());
    }
    for (double v : dtn) {
        if (!(v > 0.0)) {
            return fail(MIO_INVALID_VALUE, "initial integrator dt must be positive");
        }
    }
    // node clock records {t, dt, ICUCapacity, TestAndTraceCapacity}
    std::vector<double> nctl((size_t)count * N * 4);
    for (size_t i = 0; i < (size_t)count * N; ++i) {
        nctl[i * 4 + 0] = tn[i];
        nctl[i * 4 + 1] = dtn[i];
        nctl[i * 4 + 2] = data->node_local[i * 2 + 0];
        nctl[i * 4 + 3] = data->node_local[i * 2 + 1];
    }
    CUDA_TRY(cudaMemcpyAsync(h->d_node_ctl + s0 * N * 4, nctl.data(), sizeof(double) * nctl.size(),
                             cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaMemsetAsync(h->d_counters + s0 * N * 4, 0, sizeof(int) * (size_t)count * N * 4, h->stream));
    if (h->d_flows) {
        const size_t per = (size_t)N * geo.W * 16;
        CUDA_TRY(cudaMemsetAsync(h->d_flows + s0 * per, 0, sizeof(double) * per * count, h->stream));
    }
    if (h->d_daily) {
        // day 0 = initial value (interpolate_simulation_result, analyze_result.h:228-231)
        std::vector<double> tmp; // strided copy: slot day 0 of each (run,node)
        const size_t md = (size_t)h->desc.max_days;
        CUDA_TRY(cudaMemcpy2DAsync(h->d_daily + s0 * N * md * n_ref, sizeof(double) * md * n_ref, data->y0,
                                   sizeof(double) * n_ref, sizeof(double) * n_ref, (size_t)count * N,
                                   cudaMemcpyHostToDevice, h->stream));
        h->days_recorded = std::max(h->days_recorded, 1);
    }
    if (h->d_traj_npts) {
        const size_t cap = (size_t)h->desc.traj_capacity;
        std::vector<int> one((size_t)count * N, 1);
        CUDA_TRY(cudaMemcpyAsync(h->d_traj_npts + s0 * N, one.data(), sizeof(int) * one.size(),
                                 cudaMemcpyHostToDevice, h->stream));
        CUDA_TRY(cudaMemcpy2DAsync(h->d_traj_times + s0 * N * cap, sizeof(double) * cap, tn.data(), sizeof(double),
                                   sizeof(double), (size_t)count * N, cudaMemcpyHostToDevice, h->stream));
        CUDA_TRY(cudaMemcpy2DAsync(h->d_traj_values + s0 * N * cap * n_ref, sizeof(double) * cap * n_ref, data->y0,
                                   sizeof(double) * n_ref, sizeof(double) * n_ref, (size_t)count * N,
                                   cudaMemcpyHostToDevice, h->stream));
    }
    if (h->ens) {
        if (h->ens_offset + slot0 + count > h->ens->capacity) {
            return fail(MIO_OUT_OF_RANGE, "ensemble record: run capacity exceeded");
        }
        if (h->ens->first_day == 0) { // day 0 lies inside the record's window
            ++h->launches;
            ++h->prof_launches[MIO_KC_ENSEMBLE];
            CUDA_TRY(launch_record_ensemble(h->d_y[h->cur] + s0 * N * n, h->ens->d_data, N, G, geo.C_ref, geo.LAY,
                                            count, 0, h->ens->num_days, h->ens->mode, h->ens_offset + slot0,
                                            h->ens->capacity, h->stream));
        }
        {
            const long long upto = h->ens_offset + slot0 + count;
            long long seen       = h->ens->num_runs.load();
            while (seen < upto && !h->ens->num_runs.compare_exchange_weak(seen, upto)) {
            }
        }
    }
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return MIO_OK;
}

// rewind the graph simulation to t0 (run slots keep their data until set_runs overwrites them)
int mio_cuda_reset(mio_handle* h)
{
    if (!h) {
        return fail(MIO_INVALID_VALUE, "null handle");
    }
    CUDA_TRY(cudaSetDevice(h->device));
    h->t           = h->desc.t0;
    h->toggle      = false;
    h->have_mobile = false;
    h->snap_idx = h->v0_idx = -1;
    h->days_recorded        = 0;
    h->mob_npts             = 0;
    CUDA_TRY(cudaMemsetAsync(h->d_counters, 0, sizeof(int) * (size_t)h->max_runs * h->geo.N * 4, h->stream));
    if (h->d_run_flags) {
        CUDA_TRY(cudaMemsetAsync(h->d_run_flags, 0, sizeof(int) * (size_t)h->max_runs, h->stream));
    }
    return MIO_OK;
}

int mio_cuda_set_active_runs(mio_handle* h, int32_t count)
{
    if (!h || count < 0 || count > h->max_runs) {
        return fail(MIO_OUT_OF_RANGE, "active run count out of range");
    }
    h->active_runs = count;
    return MIO_OK;
}

// GraphSimulation::advance, graph_simulation.h:101-120
int mio_cuda_advance(mio_handle* h, double t_max)
{
    if (!h) {
        return fail(MIO_INVALID_VALUE, "null handle");
    }
    CUDA_TRY(cudaSetDevice(h->device));
    const Geometry& geo = h->geo;
    const int R         = h->active_runs;
    if (R == 0) {
        return MIO_OK;
    }
    if (h->ens && h->ens_offset + R > h->ens->capacity) {
        return fail(MIO_OUT_OF_RANGE, "ensemble record: run_offset + active runs exceeds the capacity");
    }
    double dt = h->desc.dt_graph;
    while (h->t < t_max) {
        if (h->t + dt > t_max) {
            dt = t_max - h->t;
        }
        const bool will_depart = geo.E > 0 && !h->toggle;
        // node phase
        const int out_idx = free_buffer(h, h->cur, -1, -1);
        if (out_idx < 0 || !h->d_y[out_idx]) {
            return fail(MIO_UNKNOWN_ERROR, "internal: no free state buffer");
        }
        IntegrateArgs ia;
        ia.geo        = geo;
        ia.st         = device_state(h);
        ia.y_in       = h->d_y[h->cur];
        ia.y_out      = h->d_y[out_idx];
        ia.fac_out    = will_depart ? h->d_fac : nullptr;
        ia.u_out      = will_depart ? h->d_u : nullptr;
        ia.umask      = h->d_umask;
        ia.t_end      = h->t + dt;
        ia.integrator = h->desc.integrator.kind;
        ia.abs_tol    = h->desc.integrator.abs_tol;
        ia.rel_tol    = h->desc.integrator.rel_tol;
        ia.dt_min     = h->desc.integrator.dt_min;
        ia.dt_max     = h->desc.integrator.dt_max;
        ia.traj       = TrajectoryBuffers{h->desc.traj_capacity, h->d_traj_npts, h->d_traj_times, h->d_traj_values};
        ia.num_runs   = R;
        ia.work_counter = h->d_work;
        ia.flows        = h->d_flows;
        ia.traj_flows   = h->d_traj_flows;
        // the node phase right after a departure also provides the return coefficients (see IntegrateArgs::aux_out)
        ia.aux_out      = (geo.E > 0 && h->have_mobile && h->aux_pending) ? h->d_aux : nullptr;
        h->aux_pending  = false;
        {
            ScopedSpan sp(h, MIO_KC_INTEGRATE);
            CUDA_TRY(launch_integrate(ia, h->stream));
        }
        h->cur = out_idx;
        h->t += dt;
        // edge phase, MobilityEdge::apply_mobility for all edges
        if (geo.E > 0) {
            EdgeArgs ea;
            ea.geo            = geo;
            ea.st             = device_state(h);
            ea.csr = GraphCsr{h->d_order_in, h->d_order_out, h->d_in_ptr,  h->d_in_split, h->d_in_src,
                              h->d_w_in,     h->d_out_ptr,   h->d_out_dst, h->d_w_out,    h->d_s_out,
                              h->d_umask,    h->factorized,  h->num_mob,   {0},           h->d_edge_start,
                              h->d_edge_end, h->d_coeff};
            std::copy(h->mob_comp, h->mob_comp + NC, ea.csr.mob_comp);
            ea.fac            = h->d_fac;
            ea.u              = h->d_u;
            ea.vis            = h->d_vis;
            ea.aux            = h->d_aux;
            ea.aux_by_node_phase = 1;
            ea.run_flags      = h->d_run_flags;
            ea.stats          = h->d_stats;
            ea.run_list       = h->d_run_list;
            ea.list_count     = h->d_work + 1;
            ea.csr.w_edge     = h->d_w_edge;
            ea.xslot          = h->d_xslot;
            ea.xsrc           = h->d_xsrc;
            ea.xchunks        = h->xchunks;
            ea.xepp           = h->xepp;
            ea.xdamp          = h->d_xdamp;
            ea.num_runs       = R;
            ea.edge_m         = h->d_edge_m;
            MobilityResultArgs mra{};
            bool mob_point = false; // this graph step moved commuters: one time point per edge
            if (h->d_edge_m) {
                mra = MobilityResultArgs{geo.n,          geo.n_ref,        geo.E,       h->mob_K,        h->d_mob_gptr,
                                         h->d_mob_gdev,  h->d_ref_to_dev,  h->d_edge_active, h->d_edge_m, h->d_mob_values,
                                         h->mob_cap,     h->mob_npts,      R};
            }
            ea.dt             = dt;
            ea.t              = h->t;
            if (h->have_mobile && h->return_time <= h->t) {
                const int ret_idx = free_buffer(h, h->cur, -1, -1);
                if (ret_idx < 0) {
                    return fail(MIO_UNKNOWN_ERROR, "internal: no free state buffer");
                }
                ea.snap  = h->d_y[h->snap_idx];
                ea.y_in  = h->d_y[h->cur];
                ea.y_out = h->d_y[ret_idx];
                ea.t_dep = h->t_dep;
                {
                    ScopedSpan sp(h, MIO_KC_RETURN);
                    CUDA_TRY(launch_return(ea, h->stream));
                }
                if (h->d_edge_m) {
                    // the per-edge returning populations come out of the literal replay: every run is replayed
                    CUDA_TRY(cudaMemsetAsync(h->d_run_flags, 1, sizeof(int) * (size_t)R, h->stream));
                }
                // runs that need the order dependent underflow fix-up are replayed literally (the kernel exits
                // at once for unflagged runs and clears the flag, so no host round trip is needed)
                {
                    ScopedSpan sp(h, MIO_KC_RETURN_EXACT);
                    CUDA_TRY(launch_return_exact(ea, h->stream));
                }
                if (h->d_edge_m && h->mob_npts < h->mob_cap) {
                    CUDA_TRY(launch_mobility_results(mra, h->stream));
                    ++h->launches;
                    mob_point = true;
                }
                h->cur         = ret_idx;
                h->have_mobile = false;
                h->snap_idx = h->v0_idx = -1;
            }
            if (!h->toggle) {
                if (h->have_mobile) {
                    return fail(MIO_INVALID_VALUE, "irregular stepping: a second departure before the previous "
                                                   "commuters returned is not supported");
                }
                const int dep_idx = free_buffer(h, h->cur, -1, -1);
                if (dep_idx < 0) {
                    return fail(MIO_UNKNOWN_ERROR, "internal: no free state buffer");
                }
                ea.snap  = h->d_y[h->cur];
                ea.y_in  = h->d_y[h->cur];
                ea.y_out = h->d_y[dep_idx];
                ea.t_dep = h->t;
                {
                    ScopedSpan sp(h, MIO_KC_DEPART);
                    CUDA_TRY(launch_depart(ea, h->stream));
                }
                if (h->d_edge_m && h->mob_npts < h->mob_cap) {
                    CUDA_TRY(launch_edge_mobile_depart(ea, h->stream));
                    CUDA_TRY(launch_mobility_results(mra, h->stream));
                    h->launches += 2;
                    mob_point = true;
                }
                h->snap_idx    = h->cur;
                h->v0_idx      = dep_idx;
                h->cur         = dep_idx;
                h->have_mobile = true;
                h->aux_pending = true;
                h->t_dep       = h->t;
                h->return_time = h->t + dt;
            }
            h->toggle = !h->toggle;
            if (mob_point) {
                h->mob_times[(size_t)h->mob_npts] = h->t;
                ++h->mob_npts;
            }
        }
        // daily record: the value the reference's interpolate_simulation_result yields at integer days
        if (h->d_daily || h->ens) {
            const double day = std::round(h->t);
            if (std::fabs(h->t - day) <= 1e-10 * std::max(1.0, std::fabs(day))) {
                const int slot = (int)day - h->day0;
                const int wslot = h->ens ? slot - h->ens->first_day : -1; // slot inside the record's day window
                if (h->ens && wslot >= 0 && wslot < h->ens->num_days) {
                    ScopedSpan sp(h, MIO_KC_ENSEMBLE);
                    CUDA_TRY(launch_record_ensemble(h->d_y[h->cur], h->ens->d_data, geo.N, geo.G, geo.C_ref, geo.LAY, R,
                                                    wslot, h->ens->num_days, h->ens->mode, h->ens_offset,
                                                    h->ens->capacity, h->stream));
                }
                if (h->d_daily && slot >= 0 && slot < h->desc.max_days) {
                    ScopedSpan sp(h, MIO_KC_RECORD);
                    CUDA_TRY(launch_record_daily(geo, h->d_y[h->cur], h->d_daily, slot, h->desc.max_days, R,
                                                 h->stream));
                    h->days_recorded = std::max(h->days_recorded, slot + 1);
                }
            }
        }
    }
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    collect_spans(h);
    return MIO_OK;
}

int mio_cuda_get_time(mio_handle* h, double* t)
{
    if (!h || !t) {
        return fail(MIO_INVALID_VALUE, "null argument");
    }
    *t = h->t;
    return MIO_OK;
}

static int check_slots(mio_handle* h, int32_t slot0, int32_t count, const void* out)
{
    if (!h || !out) {
        return fail(MIO_INVALID_VALUE, "null argument");
    }
    if (slot0 < 0 || count < 0 || slot0 + count > h->max_runs) {
        return fail(MIO_OUT_OF_RANGE, "run slots out of range");
    }
    return MIO_OK;
}

int mio_cuda_get_state(mio_handle* h, int32_t slot0, int32_t count, double* out)
{
    if (int rc = check_slots(h, slot0, count, out)) {
        return rc;
    }
    CUDA_TRY(cudaSetDevice(h->device));
    const size_t per = (size_t)h->geo.N * h->geo.n;
    if (h->geo.LAY == 1) {
        CUDA_TRY(cudaMemcpy(out, h->d_y[h->cur] + (size_t)slot0 * per, sizeof(double) * per * count,
                            cudaMemcpyDeviceToHost));
    }
    else {
        // permuted into the reference's element order in a free state buffer, then copied out
        const int stage = free_buffer(h, h->cur, -1, -1);
        if (stage < 0 || !h->d_y[stage]) {
            return fail(MIO_UNKNOWN_ERROR, "internal: no free state buffer");
        }
        CUDA_TRY(launch_permute_state(h->geo, h->d_y[h->cur] + (size_t)slot0 * per, h->d_y[stage],
                                      (size_t)count * h->geo.N, 0, h->stream));
        ++h->launches;
        CUDA_TRY(cudaMemcpyAsync(out, h->d_y[stage], sizeof(double) * (size_t)count * h->geo.N * h->geo.n_ref,
                                 cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
    }
    return MIO_OK;
}

int mio_cuda_get_node_dt(mio_handle* h, int32_t slot0, int32_t count, double* out)
{
    if (int rc = check_slots(h, slot0, count, out)) {
        return rc;
    }
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaMemcpy2D(out, sizeof(double), h->d_node_ctl + (size_t)slot0 * h->geo.N * 4 + 1, sizeof(double) * 4,
                          sizeof(double), (size_t)count * h->geo.N, cudaMemcpyDeviceToHost));
    return MIO_OK;
}

int mio_cuda_get_counters(mio_handle* h, int32_t slot0, int32_t count, int64_t* out)
{
    if (int rc = check_slots(h, slot0, count, out)) {
        return rc;
    }
    CUDA_TRY(cudaSetDevice(h->device));
    const size_t m = (size_t)count * h->geo.N * 4;
    std::vector<int> tmp(m);
    CUDA_TRY(cudaMemcpy(tmp.data(), h->d_counters + (size_t)slot0 * h->geo.N * 4, sizeof(int) * m,
                        cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < m; ++i) {
        out[i] = tmp[i];
    }
    return MIO_OK;
}

int mio_cuda_get_daily(mio_handle* h, int32_t slot0, int32_t count, int32_t* num_days, double* out,
                       int64_t out_capacity)
{
    if (int rc = check_slots(h, slot0, count, out)) {
        return rc;
    }
    if (!h->d_daily || !num_days) {
        return fail(MIO_INVALID_VALUE, "graph was created without record_daily");
    }
    CUDA_TRY(cudaSetDevice(h->device));
    const size_t nd = (size_t)h->days_recorded, md = (size_t)h->desc.max_days, n = (size_t)h->geo.n_ref;
    const size_t rows = (size_t)count * h->geo.N;
    if ((int64_t)(rows * nd * n) > out_capacity) {
        return fail(MIO_OUT_OF_RANGE, "output buffer too small for the daily record");
    }
    *num_days = (int32_t)nd;
    CUDA_TRY(cudaMemcpy2D(out, sizeof(double) * nd * n, h->d_daily + (size_t)slot0 * h->geo.N * md * n,
                          sizeof(double) * md * n, sizeof(double) * nd * n, rows, cudaMemcpyDeviceToHost));
    return MIO_OK;
}

int mio_cuda_get_trajectory(mio_handle* h, int32_t slot0, int32_t count, int32_t* npts, double* times,
                            double* values)
{
    if (int rc = check_slots(h, slot0, count, npts)) {
        return rc;
    }
    if (!h->d_traj_npts || !times || !values) {
        return fail(MIO_INVALID_VALUE, "graph was created without traj_capacity, or null output");
    }
    CUDA_TRY(cudaSetDevice(h->device));
    const size_t rows = (size_t)count * h->geo.N, cap = (size_t)h->desc.traj_capacity, n = (size_t)h->geo.n_ref;
    const size_t off = (size_t)slot0 * h->geo.N;
    CUDA_TRY(cudaMemcpy(npts, h->d_traj_npts + off, sizeof(int) * rows, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(times, h->d_traj_times + off * cap, sizeof(double) * rows * cap, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(values, h->d_traj_values + off * cap * n, sizeof(double) * rows * cap * n,
                        cudaMemcpyDeviceToHost));
    return MIO_OK;
}

int mio_cuda_get_flows(mio_handle* h, int32_t slot0, int32_t count, double* out)
{
    if (int rc = check_slots(h, slot0, count, out)) {
        return rc;
    }
    if (!h->d_flows) {
        return fail(MIO_INVALID_VALUE, "graph was created without flow_simulation");
    }
    CUDA_TRY(cudaSetDevice(h->device));
    const size_t lanes = (size_t)count * h->geo.N * h->geo.W;
    std::vector<double> tmp(lanes * 16);
    CUDA_TRY(cudaMemcpy(tmp.data(), h->d_flows + (size_t)slot0 * h->geo.N * h->geo.W * 16, sizeof(double) * tmp.size(),
                        cudaMemcpyDeviceToHost));
    for (size_t l = 0; l < lanes; ++l) {
        std::copy(tmp.begin() + l * 16, tmp.begin() + l * 16 + 15, out + l * 15);
    }
    return MIO_OK;
}

// Replace the contact tables of resident runs (dynamic NPIs change a node's contact patterns during the simulation,
// cpp/models/ode_secir/model.h:299-353): ctables[count][num_ctables][ctable_size], the layout of mio_cuda_set_runs.
int mio_cuda_set_contact_tables(mio_handle* h, int32_t slot0, int32_t count, const double* ctables)
{
    if (int rc = check_slots(h, slot0, count, ctables)) {
        return rc;
    }
    CUDA_TRY(cudaSetDevice(h->device));
    const Geometry& geo = h->geo;
    const size_t csz    = (size_t)geo.ctable_stride, nct = (size_t)geo.num_ctables;
    for (size_t r = 0; r < (size_t)count; ++r) {
        for (size_t c = 0; c < nct; ++c) {
            const double* ct = ctables + (r * nct + c) * csz;
            const int L      = (int)ct[0];
            if (L < 1 || L > geo.ct_max_loc) {
                return fail(MIO_INVALID_VALUE, "contact table: number of locations out of range");
            }
            for (int l = 0; l < L; ++l) {
                if (ct[1 + l] < 0 || ct[1 + l] > geo.ct_max_pts) {
                    return fail(MIO_INVALID_VALUE, "contact table: number of change points out of range");
                }
            }
        }
    }
    CUDA_TRY(cudaMemcpy2DAsync(h->d_run_block + (size_t)slot0 * geo.block_stride + geo.ct_off,
                               sizeof(double) * geo.block_stride, ctables, sizeof(double) * nct * csz,
                               sizeof(double) * nct * csz, (size_t)count, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return MIO_OK;
}

int mio_cuda_set_flows(mio_handle* h, int32_t slot0, int32_t count, const double* flows)
{
    if (int rc = check_slots(h, slot0, count, flows)) {
        return rc;
    }
    if (!h->d_flows) {
        return fail(MIO_INVALID_VALUE, "graph was created without flow_simulation");
    }
    CUDA_TRY(cudaSetDevice(h->device));
    const size_t lanes = (size_t)count * h->geo.N * h->geo.W;
    std::vector<double> tmp(lanes * 16, 0.0);
    for (size_t l = 0; l < lanes; ++l) {
        std::copy(flows + l * 15, flows + l * 15 + 15, tmp.begin() + l * 16);
    }
    CUDA_TRY(cudaMemcpy(h->d_flows + (size_t)slot0 * h->geo.N * h->geo.W * 16, tmp.data(), sizeof(double) * tmp.size(),
                        cudaMemcpyHostToDevice));
    if (h->d_traj_flows) { // time point 0 of the flow series
        const size_t cap = (size_t)h->desc.traj_capacity, nf = (size_t)h->geo.G * 15;
        CUDA_TRY(cudaMemcpy2D(h->d_traj_flows + (size_t)slot0 * h->geo.N * cap * nf, sizeof(double) * cap * nf, flows,
                              sizeof(double) * nf, sizeof(double) * nf, (size_t)count * h->geo.N,
                              cudaMemcpyHostToDevice));
    }
    return MIO_OK;
}

int mio_cuda_get_flow_trajectory(mio_handle* h, int32_t slot0, int32_t count, double* values)
{
    if (int rc = check_slots(h, slot0, count, values)) {
        return rc;
    }
    if (!h->d_traj_flows) {
        return fail(MIO_INVALID_VALUE, "graph was created without flow_simulation and traj_capacity");
    }
    CUDA_TRY(cudaSetDevice(h->device));
    const size_t per = (size_t)h->geo.N * (size_t)h->desc.traj_capacity * h->geo.G * 15;
    CUDA_TRY(cudaMemcpy(values, h->d_traj_flows + (size_t)slot0 * per, sizeof(double) * per * count,
                        cudaMemcpyDeviceToHost));
    return MIO_OK;
}

int mio_cuda_enable_mobility_results(mio_handle* h, int32_t num_groups, const int32_t* group_ptr,
                                     const int32_t* group_idx, int32_t capacity)
{
    if (!h || num_groups < 0 || capacity < 1 || (num_groups > 0 && (!group_ptr || !group_idx))) {
        return fail(MIO_INVALID_VALUE, "invalid arguments");
    }
    const Geometry& geo = h->geo;
    if (geo.E == 0) {
        return fail(MIO_INVALID_VALUE, "graph has no edges");
    }
    const size_t R = (size_t)h->max_runs;
    if (R * geo.E * geo.n * sizeof(double) > (size_t)8 << 30) {
        return fail(MIO_OUT_OF_RANGE, "mobility results keep the traveler state of every edge of every run: too many "
                                      "runs for this handle (use a handle with few runs)");
    }
    for (int g = 0; g < num_groups; ++g) {
        for (int q = group_ptr[g]; q < group_ptr[g + 1]; ++q) {
            if (group_idx[q] < 0 || group_idx[q] >= geo.n_ref) {
                return fail(MIO_OUT_OF_RANGE, "saved compartment index out of range");
            }
        }
    }
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    const int K1 = num_groups + 1;
    if (h->d_edge_m && h->mob_K != num_groups) {
        return fail(MIO_INVALID_VALUE, "mobility results are already enabled with other index groups");
    }
    if (!h->d_edge_m) {
        h->mob_K = num_groups;
        h->mob_gptr.assign(group_ptr, group_ptr + num_groups + 1);
        h->mob_gidx.assign(group_idx, group_idx + (num_groups > 0 ? group_ptr[num_groups] : 0));
        std::vector<int> gdev(h->mob_gidx.size()), r2d((size_t)geo.n_ref);
        for (int j = 0; j < geo.n_ref; ++j) {
            r2d[(size_t)j] = dev_index_of_ref(geo.LAY, geo.G, j / geo.C_ref, j % geo.C_ref);
        }
        for (size_t q = 0; q < gdev.size(); ++q) {
            gdev[q] = r2d[(size_t)h->mob_gidx[q]];
        }
        CUDA_TRY(dev_alloc(&h->d_mob_gptr, h->mob_gptr.size()));
        CUDA_TRY(cudaMemcpy(h->d_mob_gptr, h->mob_gptr.data(), sizeof(int) * h->mob_gptr.size(), cudaMemcpyHostToDevice));
        CUDA_TRY(dev_alloc(&h->d_mob_gdev, std::max<size_t>(gdev.size(), 1)));
        if (!gdev.empty()) {
            CUDA_TRY(cudaMemcpy(h->d_mob_gdev, gdev.data(), sizeof(int) * gdev.size(), cudaMemcpyHostToDevice));
        }
        CUDA_TRY(dev_alloc(&h->d_ref_to_dev, r2d.size()));
        CUDA_TRY(cudaMemcpy(h->d_ref_to_dev, r2d.data(), sizeof(int) * r2d.size(), cudaMemcpyHostToDevice));
        CUDA_TRY(dev_alloc(&h->d_edge_active, h->edge_active.size()));
        CUDA_TRY(cudaMemcpy(h->d_edge_active, h->edge_active.data(), h->edge_active.size(), cudaMemcpyHostToDevice));
        CUDA_TRY(dev_alloc(&h->d_edge_m, R * geo.E * geo.n));
        CUDA_TRY(cudaMemset(h->d_edge_m, 0, sizeof(double) * R * geo.E * geo.n));
    }
    if (capacity > h->mob_cap) {
        double* nv = nullptr;
        CUDA_TRY(dev_alloc(&nv, R * geo.E * (size_t)capacity * K1));
        CUDA_TRY(cudaMemset(nv, 0, sizeof(double) * R * geo.E * (size_t)capacity * K1));
        if (h->d_mob_values && h->mob_npts > 0) {
            CUDA_TRY(cudaMemcpy2D(nv, sizeof(double) * (size_t)capacity * K1, h->d_mob_values,
                                  sizeof(double) * (size_t)h->mob_cap * K1, sizeof(double) * (size_t)h->mob_npts * K1,
                                  R * geo.E, cudaMemcpyDeviceToDevice));
        }
        cudaFree(h->d_mob_values);
        h->d_mob_values = nv;
        h->mob_cap      = capacity;
        h->mob_times.resize((size_t)capacity, 0.0);
    }
    return MIO_OK;
}

int mio_cuda_get_mobility_results(mio_handle* h, int32_t slot0, int32_t count, int32_t* npts, double* times,
                                  double* values)
{
    if (int rc = check_slots(h, slot0, count, npts)) {
        return rc;
    }
    if (!h->d_mob_values || !times || !values) {
        return fail(MIO_INVALID_VALUE, "mobility results are not enabled, or null output");
    }
    CUDA_TRY(cudaSetDevice(h->device));
    const size_t E = (size_t)h->geo.E, K1 = (size_t)h->mob_K + 1, cap = (size_t)h->mob_cap;
    for (int r = 0; r < count; ++r) {
        for (size_t e = 0; e < E; ++e) {
            npts[(size_t)r * E + e] = h->edge_active[e] ? h->mob_npts : 0;
        }
    }
    std::copy(h->mob_times.begin(), h->mob_times.end(), times);
    CUDA_TRY(cudaMemcpy(values, h->d_mob_values + (size_t)slot0 * E * cap * K1, sizeof(double) * (size_t)count * E * cap * K1,
                        cudaMemcpyDeviceToHost));
    return MIO_OK;
}

int mio_cuda_get_stats(mio_handle* h, int64_t* launches, int64_t* exact_replays)
{
    if (!h) {
        return fail(MIO_INVALID_VALUE, "null handle");
    }
    if (launches) {
        *launches = h->launches;
    }
    if (exact_replays) {
        *exact_replays = 0;
        if (h->d_stats) {
            long long v = 0;
            CUDA_TRY(cudaSetDevice(h->device));
            CUDA_TRY(cudaMemcpy(&v, h->d_stats, sizeof(long long), cudaMemcpyDeviceToHost));
            *exact_replays = v;
        }
    }
    return MIO_OK;
}

int mio_cuda_get_stream(mio_handle* h, void** stream)
{
    if (!h || !stream) {
        return fail(MIO_INVALID_VALUE, "null argument");
    }
    *stream = (void*)h->stream;
    return MIO_OK;
}

int mio_cuda_set_profiling(mio_handle* h, int on)
{
    if (!h) {
        return fail(MIO_INVALID_VALUE, "null handle");
    }
    h->profiling = on != 0;
    for (int k = 0; k < MIO_NUM_KERNEL_CLASSES; ++k) {
        h->prof_ms[k]       = 0.0;
        h->prof_launches[k] = 0;
    }
    return MIO_OK;
}

int mio_cuda_get_profile(mio_handle* h, double* ms, int64_t* launches)
{
    if (!h || !ms || !launches) {
        return fail(MIO_INVALID_VALUE, "null argument");
    }
    for (int k = 0; k < MIO_NUM_KERNEL_CLASSES; ++k) {
        ms[k]       = h->prof_ms[k];
        launches[k] = h->prof_launches[k];
    }
    return MIO_OK;
}

int mio_cuda_device_state_ptr(mio_handle* h, void** y, int64_t* num_doubles)
{
    if (!h || !y || !num_doubles) {
        return fail(MIO_INVALID_VALUE, "null argument");
    }
    *y           = (void*)h->d_y[h->cur];
    *num_doubles = (int64_t)h->max_runs * h->geo.N * h->geo.n;
    return MIO_OK;
}

// ---- ensemble record and reduction ------------------------------------------------------------------------------
int mio_cuda_ensemble_create(int device, int32_t model, int32_t num_groups, int32_t num_nodes, int32_t num_days,
                             int32_t mode, int64_t capacity_runs, mio_ensemble** out)
{
    if (!out) {
        return fail(MIO_INVALID_VALUE, "null argument");
    }
    *out = nullptr;
    if ((model != MIO_MODEL_OSECIR && model != MIO_MODEL_OSECIRVVS) || num_groups < 1 || num_groups > MIO_MAX_GROUPS || num_nodes < 1 || num_days < 1 ||
        capacity_runs < 1 || (mode != MIO_ENS_FULL && mode != MIO_ENS_AGE_SUM)) {
        return fail(MIO_INVALID_VALUE, "invalid ensemble record sizes");
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        return fail(MIO_UNKNOWN_ERROR, "no CUDA device available (this library has no CPU fallback)");
    }
    if (device < 0 || device >= ndev) {
        return fail(MIO_OUT_OF_RANGE, "device index out of range");
    }
    CUDA_TRY(cudaSetDevice(device));
    auto* e        = new mio_ensemble();
    e->device      = device;
    e->G           = num_groups;
    e->C           = mio_num_compartments(model);
    e->lay         = model == MIO_MODEL_OSECIRVVS ? VVS_LAY : 1;
    e->N           = num_nodes;
    e->num_days    = num_days;
    e->mode        = mode;
    e->capacity    = capacity_runs;
    const int n    = num_groups * e->C;
    const int El   = mode == MIO_ENS_AGE_SUM ? e->C : n;
    e->num_columns = (long long)num_nodes * num_days * El + (long long)num_days * n;
    cudaError_t er = dev_alloc(&e->d_data, (size_t)e->num_columns * (size_t)capacity_runs);
    if (er != cudaSuccess) {
        delete e;
        return fail(MIO_UNKNOWN_ERROR, std::string("ensemble record allocation: ") + cudaGetErrorString(er));
    }
    *out = e;
    return MIO_OK;
}

void mio_cuda_ensemble_destroy(mio_ensemble* e)
{
    if (e) {
        cudaSetDevice(e->device);
        cudaFree(e->d_data);
        delete e;
    }
}

int mio_cuda_attach_ensemble(mio_handle* h, mio_ensemble* e, int64_t run_offset)
{
    if (!h) {
        return fail(MIO_INVALID_VALUE, "null handle");
    }
    if (e) {
        if (e->device != h->device || e->G != h->geo.G || e->N != h->geo.N || e->C != h->geo.C_ref) {
            return fail(MIO_INVALID_VALUE, "ensemble record does not match the graph (device, groups or nodes)");
        }
        if (run_offset < 0 || run_offset >= e->capacity) {
            return fail(MIO_OUT_OF_RANGE, "ensemble record: run_offset exceeds the capacity");
        }
        if (std::floor(h->desc.t0) != h->desc.t0) {
            return fail(MIO_INVALID_VALUE, "ensemble record needs an integer t0");
        }
    }
    h->ens        = e;
    h->ens_offset = run_offset;
    return MIO_OK;
}

int mio_cuda_ensemble_info(mio_ensemble* e, int64_t* num_columns, int64_t* stride, int64_t* num_runs,
                           void** device_ptr)
{
    if (!e) {
        return fail(MIO_INVALID_VALUE, "null argument");
    }
    if (num_columns) {
        *num_columns = e->num_columns;
    }
    if (stride) {
        *stride = e->capacity;
    }
    if (num_runs) {
        *num_runs = e->num_runs.load();
    }
    if (device_ptr) {
        *device_ptr = e->d_data;
    }
    return MIO_OK;
}

int mio_cuda_ensemble_set_window(mio_ensemble* e, int32_t first_day)
{
    if (!e || first_day < 0) {
        return fail(MIO_INVALID_VALUE, "invalid ensemble day window");
    }
    e->first_day = first_day;
    return MIO_OK;
}

int mio_cuda_select_percentiles(int device, int32_t num_segments, const void* const* seg_ptr,
                                const int64_t* seg_stride, const int64_t* seg_count, int64_t num_columns,
                                int32_t num_p, const double* p, double* percentiles, double* mean)
{
    if (num_segments < 1 || num_segments > SEL_MAX_SEG || !seg_ptr || !seg_stride || !seg_count || num_columns < 0 ||
        num_p < 0 || num_p > SEL_MAX_P || (num_p > 0 && (!p || !percentiles))) {
        return fail(MIO_INVALID_VALUE, "invalid arguments");
    }
    CUDA_TRY(cudaSetDevice(device));
    SelectArgs a{};
    a.num_segments = num_segments;
    a.total        = 0;
    for (int s = 0; s < num_segments; ++s) {
        if (seg_count[s] < 0 || seg_count[s] > seg_stride[s] || (!seg_ptr[s] && seg_count[s] > 0)) {
            return fail(MIO_INVALID_VALUE, "invalid segment");
        }
        a.seg_ptr[s]    = static_cast<const double*>(seg_ptr[s]);
        a.seg_stride[s] = seg_stride[s];
        a.seg_count[s]  = seg_count[s];
        a.total += seg_count[s];
    }
    if (a.total == 0) {
        return fail(MIO_INVALID_VALUE, "no runs recorded");
    }
    a.num_p = num_p;
    for (int q = 0; q < num_p; ++q) {
        // analyze_result.h:318,338: assert(p > 0 && p < 1); index static_cast<size_t>(num_runs * p)
        if (!(p[q] > 0.0 && p[q] < 1.0)) {
            return fail(MIO_INVALID_VALUE, "percentile must lie in (0, 1)");
        }
        a.rank[q] = (long long)(size_t)((double)a.total * p[q]);
    }
    a.num_columns = num_columns;
    // timing of the phases to stderr when MIO_CUDA_TRACE is set (allocation, wait for the record kernels, selection,
    // copy back)
    const bool trace = std::getenv("MIO_CUDA_TRACE") != nullptr;
    auto now         = [] { return std::chrono::steady_clock::now(); };
    auto ms          = [](std::chrono::steady_clock::time_point a0, std::chrono::steady_clock::time_point a1) {
        return std::chrono::duration<double, std::milli>(a1 - a0).count();
    };
    const auto t0 = now();
    // result buffers are kept per device between calls (grow only): cudaMalloc / cudaFree of a process that holds tens
    // of GB cost up to hundreds of milliseconds each
    static std::mutex scratch_mutex;
    static struct {
        double* p[2]  = {nullptr, nullptr};
        size_t cap[2] = {0, 0};
    } scratch[64];
    if (device < 0 || device >= 64) {
        return fail(MIO_INVALID_VALUE, "device index out of range");
    }
    std::lock_guard<std::mutex> guard(scratch_mutex);
    auto reserve = [&](int which, size_t count) -> double* {
        auto& sc = scratch[device];
        if (sc.cap[which] < count) {
            cudaFree(sc.p[which]);
            sc.p[which]   = nullptr;
            sc.cap[which] = 0;
            if (cudaMalloc(&sc.p[which], sizeof(double) * count) != cudaSuccess) {
                return nullptr;
            }
            sc.cap[which] = count;
        }
        return sc.p[which];
    };
    double *d_sel = nullptr, *d_mean = nullptr;
    if (num_p > 0) {
        d_sel = reserve(0, (size_t)num_p * num_columns);
        if (!d_sel) {
            return fail(MIO_UNKNOWN_ERROR, "select_percentiles: out of device memory");
        }
    }
    if (mean) {
        d_mean = reserve(1, (size_t)num_columns);
        if (!d_mean) {
            return fail(MIO_UNKNOWN_ERROR, "select_percentiles: out of device memory");
        }
    }
    a.out_select   = d_sel;
    a.out_mean     = d_mean;
    const auto t1  = now();
    cudaError_t er = cudaDeviceSynchronize(); // the record kernels run on the handles' streams
    const auto t2  = now();
    if (er == cudaSuccess) {
        er = launch_select(a, nullptr);
    }
    if (er == cudaSuccess && trace) {
        er = cudaDeviceSynchronize();
    }
    const auto t3 = now();
    if (er == cudaSuccess && d_sel) {
        er = cudaMemcpy(percentiles, d_sel, sizeof(double) * (size_t)num_p * num_columns, cudaMemcpyDeviceToHost);
    }
    if (er == cudaSuccess && d_mean) {
        er = cudaMemcpy(mean, d_mean, sizeof(double) * (size_t)num_columns, cudaMemcpyDeviceToHost);
    }
    const auto t4 = now();
    if (trace) {
        std::fprintf(stderr, "[mio_cuda] select: alloc %.2f ms, wait %.2f ms, kernel %.2f ms, copy back %.2f ms, free %.2f ms\n",
                     ms(t0, t1), ms(t1, t2), ms(t2, t3), ms(t3, t4), ms(t4, now()));
    }
    if (er != cudaSuccess) {
        return fail(MIO_UNKNOWN_ERROR, std::string("select_percentiles: ") + cudaGetErrorString(er));
    }
    return MIO_OK;
}

int mio_cuda_ensemble_reduce(mio_ensemble* e, int32_t num_p, const double* p, double* percentiles, double* mean)
{
    if (!e) {
        return fail(MIO_INVALID_VALUE, "null argument");
    }
    const void* ptr      = e->d_data;
    const int64_t stride = e->capacity, count = e->num_runs.load();
    return mio_cuda_select_percentiles(e->device, 1, &ptr, &stride, &count, e->num_columns, num_p, p, percentiles,
                                       mean);
}

int mio_cuda_reduce_host_columns(const double* columns, int64_t num_columns, int64_t num_runs, int32_t num_p,
                                 const double* p, double* percentiles, double* mean)
{
    if (!columns || num_columns < 0 || num_runs < 1) {
        return fail(MIO_INVALID_VALUE, "invalid arguments");
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        return fail(MIO_UNKNOWN_ERROR, "no CUDA device available (this library has no CPU fallback)");
    }
    int device = 0;
    CUDA_TRY(cudaGetDevice(&device));
    double* d = nullptr;
    CUDA_TRY(dev_alloc(&d, (size_t)num_columns * (size_t)num_runs));
    cudaError_t er = cudaMemcpy(d, columns, sizeof(double) * (size_t)num_columns * (size_t)num_runs,
                                cudaMemcpyHostToDevice);
    if (er != cudaSuccess) {
        cudaFree(d);
        return fail(MIO_UNKNOWN_ERROR, cudaGetErrorString(er));
    }
    const void* ptr = d;
    int rc = mio_cuda_select_percentiles(device, 1, &ptr, &num_runs, &num_runs, num_columns, num_p, p, percentiles,
                                         mean);
    cudaFree(d);
    return rc;
}

} // extern "C"
