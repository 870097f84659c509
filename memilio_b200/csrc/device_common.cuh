// device_common.cuh -- shared device helpers of the B200 Graph-ODE backend (sm_100a only).
//
// Lane mapping used by every kernel: a warp owns NPW = 32/W consecutive (run, node) pairs; the W = G*LAY lanes of a
// "group" own one (immunity layer, age group) each and keep that lane's NC = 10 compartment slots in registers.
// osecir: LAY = 1, the ten slots are the model's ten compartments. osecirvvs: LAY = 3 (naive, partial immunity,
// improved immunity), every layer has the same nine compartments S,E,INS,INSC,ISy,ISyC,ISev,ICr,D, stored in the
// osecir slot order with the Recovered slot unused (recoveries of all layers flow into S of the improved-immunity
// layer, cpp/models/ode_secirvvs/model.h:40-88); the device state of a node is [LAY][G][NC] and is permuted to and
// from the reference's age*27+compartment order at the C-ABI. Everything that couples lanes (force of infection,
// test-and-trace load, ICU occupancy, recoveries, error norm) is an exchange via warp shuffles.
#pragma once

#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>

namespace miocu
{

constexpr int NC      = 10; // compartment slots per lane = osecir compartments, ode_secir/infection_state.h:33-47
constexpr int VVS_NC  = 27; // osecirvvs compartments per age group, ode_secirvvs/infection_state.h:36-67
constexpr int VVS_LAY = 3;
constexpr int AGE_REC = 22; // derived per-age record (doubles): 20 slots + 2 of padding, so that the 16-byte reads of
                            // the six age groups of a node fall on different shared-memory banks
constexpr int PHDR    = 8;  // derived parameter-set header

// InfectionState order
enum : int
{
    cS = 0,
    cE,
    cINS,
    cINSC,
    cISy,
    cISyC,
    cISev,
    cICr,
    cR,
    cD
};

// derived age record slots (coefficients are precomputed with the reference's own expression order so that
// coef * y reproduces `a / b * y` of ode_secir/model.h:152-211 bit for bit)
enum : int
{
    rE_INS = 0,    // 1/TimeExposed
    rINS_ISy,      // (1-RecoveredPerInfectedNoSymptoms) * (1/TimeInfectedNoSymptoms)
    rINS_R,        // RecoveredPerInfectedNoSymptoms * (1/TimeInfectedNoSymptoms)
    rISy_ISev,     // SeverePerInfectedSymptoms / TimeInfectedSymptoms
    rISy_R,        // (1-SeverePerInfectedSymptoms) / TimeInfectedSymptoms
    rISev_ICr0,    // CriticalPerSevere / TimeInfectedSevere      (ICU not short)
    rISev_R,       // (1-CriticalPerSevere-DeathsPerSevere) / TimeInfectedSevere
    rISev_D0,      // DeathsPerSevere / TimeInfectedSevere        (ICU not short)
    rICr_D,        // DeathsPerCritical / TimeInfectedCritical
    rICr_R,        // (1-DeathsPerCritical) / TimeInfectedCritical
    rTNT,          // (1-RecoveredPerInfectedNoSymptoms) / TimeInfectedNoSymptoms
    rP,            // TransmissionProbabilityOnContact
    rREL,          // RelativeTransmissionNoSymptoms
    rRISK,         // RiskOfInfectionFromSymptomatic
    rMAXRISK,      // MaxRiskOfInfectionFromSymptomatic
    rCRIT,         // CriticalPerSevere
    rT_ISEV,       // TimeInfectedSevere
    rDPS,          // DeathsPerSevere
    rREDEXP,       // osecirvvs: ReducExposed{Partial,Improved}Immunity of the lane's layer (1 for naive / osecir)
    rINFNV         // osecirvvs: InfectiousnessNewVariant
};
// header slots
enum : int
{
    hSTART_DAY = 0,
    hSEASONALITY,
    hTNT_MAX_RISK,
    hNONDETECTION,
    hSTART_DETECTION,
    hEND_DETECTION,
    hTNT_MAX_RISK_NOSYMP, // osecirvvs: TestAndTraceCapacityMaxRiskNoSymptoms (slot 2 holds ...MaxRiskSymptoms)
    hSTART_DAY_NEW_VARIANT
};

// osecirvvs: reference compartment index of (layer, slot), -1 for the unused Recovered slot
// (ode_secirvvs/infection_state.h:36-67)
__host__ __device__ inline int vvs_ref_comp(int layer, int slot)
{
    constexpr int tab[VVS_LAY][NC] = {{0, 2, 5, 8, 11, 14, 17, 20, -1, 24},
                                      {1, 3, 6, 9, 12, 15, 18, 21, -1, 25},
                                      {23, 4, 7, 10, 13, 16, 19, 22, -1, 26}};
    return tab[layer][slot];
}
// device index (inside one node's [LAY][G][NC] block) of reference element (age, comp); LAY = 1: identity
__host__ __device__ inline int dev_index_of_ref(int lay, int G, int age, int comp)
{
    if (lay == 1) {
        return age * NC + comp;
    }
    constexpr int layer_of[VVS_NC] = {0, 1, 0, 1, 2, 0, 1, 2, 0, 1, 2, 0, 1, 2, 0, 1, 2, 0, 1, 2, 0, 1, 2, 2, 0, 1, 2};
    constexpr int slot_of[VVS_NC]  = {0, 0, 1, 1, 1, 2, 2, 2, 3, 3, 3, 4, 4, 4, 5, 5, 5, 6, 6, 6, 7, 7, 7, 0, 9, 9, 9};
    return (layer_of[comp] * G + age) * NC + slot_of[comp];
}

// problem geometry shared by all kernels (passed by value)
struct Geometry {
    int G, N, E, n;          // age groups, nodes, edges, n = W*NC doubles of one node's device state
    int LAY, W;              // immunity layers (1 osecir, 3 osecirvvs), lanes per node W = G*LAY
    int C_ref, n_ref;        // the reference's compartments per age group and state size G*C_ref
    int sim_kind;            // MIO_SIM_BASE / MIO_SIM_MODEL (osecirvvs: vaccination, variant, ADL mobility hooks)
    int vacc_days;           // osecirvvs: length of the cumulative vaccination arrays
    int num_psets, num_ctables;
    int ct_max_loc, ct_max_pts, mob_max_pts;
    int pset_stride;         // doubles per derived parameter set
    int ctable_stride;       // doubles per contact table
    int block_stride;        // doubles per run parameter block (psets | ctables | mobtable), multiple of 2
    int ct_off, mob_off;     // offsets of the contact tables / mobility table inside a run block
};

// per-handle device pointers (passed by value)
struct DeviceState {
    const double* run_block;   // [R][block_stride]
    const double* node_local;  // [R][N][2]
    const int* node_pset;      // [N]
    const int* node_ctable;    // [N]
    double* node_ctl;          // [R][N][4] node clock record: result.get_last_time(), SimulationBase::m_dt, and a copy
                               // of ICUCapacity, TestAndTraceCapacity (one 32-byte record per node: one bulk copy)
    int* counters;             // [R][N][4]
    const double* vacc_partial; // [N][G][vacc_days] cumulative DailyPartialVaccinations (osecirvvs), run-invariant
    const double* vacc_full;    // [N][G][vacc_days] cumulative DailyFullVaccinations
};

__device__ __forceinline__ double shfl(double v, int src)
{
    return __shfl_sync(0xffffffffu, v, src);
}

// W-lane group helpers, W = G*LAY; group lane gl = layer*G + age. Lanes >= NPW*W form no group and idle.
template <int G, int LAY = 1>
struct Group {
    static constexpr int W   = G * LAY;
    static constexpr int NPW = 32 / W;
    static_assert(W <= 32, "a node must fit into one warp");
    int lane, gl, gbase, age, layer;
    __device__ __forceinline__ Group()
    {
        lane  = threadIdx.x & 31;
        gl    = lane % W;
        gbase = lane - gl;
        if (lane >= NPW * W) { // spare lanes talk to themselves
            gl    = 0;
            gbase = lane;
        }
        layer = LAY == 1 ? 0 : gl / G;
        age   = LAY == 1 ? gl : gl - layer * G;
    }
    __device__ __forceinline__ bool spare() const
    {
        return lane >= NPW * W;
    }
    // value of group lane l (0 <= l < W)
    __device__ __forceinline__ double get_lane(double v, int l) const
    {
        return shfl(v, spare() ? lane : gbase + l);
    }
    // value of the lane of age group j in this lane's own layer
    __device__ __forceinline__ double get(double v, int j) const
    {
        return shfl(v, spare() ? lane : gbase + layer * G + j);
    }
    // all-reduce sum over the G age lanes of this lane's layer (pairwise tree where G allows)
    __device__ __forceinline__ double sum_ages(double v) const
    {
        const int abase = gbase + layer * G;
        if constexpr (G == 1) {
            return v;
        }
        else if constexpr (G == 2 || G == 4 || G == 8) {
#pragma unroll
            for (int o = 1; o < G; o <<= 1) {
                v += shfl(v, spare() ? lane : (abase + (age ^ o)));
            }
            return v;
        }
        else if constexpr (G == 6) {
            // pairs, then the two other pair sums
            v += shfl(v, spare() ? lane : (abase + (age ^ 1)));
            int s1   = age + 2 >= G ? age + 2 - G : age + 2;
            int s2   = age + 4 >= G ? age + 4 - G : age + 4;
            double a = shfl(v, spare() ? lane : abase + s1);
            double b = shfl(v, spare() ? lane : abase + s2);
            return v + a + b;
        }
        else {
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < G; ++j) {
                s += get(v, j);
            }
            return s;
        }
    }
    // all-reduce sum over the LAY layer lanes of this lane's age group
    __device__ __forceinline__ double sum_layers(double v) const
    {
        if constexpr (LAY == 1) {
            return v;
        }
        else {
            double s = v;
#pragma unroll
            for (int k = 1; k < LAY; ++k) {
                int l = gl + k * G;
                l     = l >= W ? l - W : l;
                s += shfl(v, spare() ? lane : gbase + l);
            }
            return s;
        }
    }
    // all-reduce sum over the whole group
    __device__ __forceinline__ double sum(double v) const
    {
        return sum_layers(sum_ages(v));
    }
    __device__ __forceinline__ double max(double v) const
    {
        if constexpr (W == 1) {
            return v;
        }
        else {
            double m = v;
            if constexpr (G > 1) {
                const int abase = gbase + layer * G;
#pragma unroll
                for (int j = 1; j < G; ++j) {
                    int s = age + j >= G ? age + j - G : age + j;
                    m     = fmax(m, shfl(v, spare() ? lane : abase + s));
                }
            }
            if constexpr (LAY > 1) {
                const double mm = m;
#pragma unroll
                for (int k = 1; k < LAY; ++k) {
                    int l = gl + k * G;
                    l     = l >= W ? l - W : l;
                    m     = fmax(m, shfl(mm, spare() ? lane : gbase + l));
                }
            }
            return m;
        }
    }
    __device__ __forceinline__ bool any(bool p) const
    {
        unsigned b    = __ballot_sync(0xffffffffu, p);
        unsigned mask = spare() ? (1u << lane) : (((1u << W) - 1u) << gbase);
        return (b & mask) != 0u;
    }
};

// cpp/memilio/math/smoother.h:32-55. The cosine ramp is taken only while a capacity threshold is being crossed; it is
// kept out of line so that the many inlined copies of the right-hand side stay small (instruction cache).
static __device__ __noinline__ double smoother_cosine_ramp(double x, double xleft, double xright, double yleft,
                                                           double yright)
{
    return 0.5 * (yleft - yright) * cos(3.141592653589793238462643383279502884 / (xright - xleft) * (x - xleft)) +
           0.5 * (yleft + yright);
}
__device__ __forceinline__ double smoother_cosine(double x, double xleft, double xright, double yleft, double yright)
{
    if (x <= xleft) {
        return yleft;
    }
    if (x >= xright) {
        return yright;
    }
    return smoother_cosine_ramp(x, xleft, xright, yleft, yright);
}

// ---------------------------------------------------------------------------------------------------------------
// Row i of ContactMatrixGroup::get_matrix_at(t) (contact_matrix.h:392-400, :190-194; damping.h:381-396) from the flat
// contact table. Also returns the closed interval [valid_lo, valid_hi] of times for which the row stays the same
// (empty, lo > hi, while any location is inside a cosine ramp). Written without FMA contraction so that the matrix
// is bit-identical to the reference's.
template <int G>
__device__ __noinline__ void contact_row_eval(const double* ct, int max_loc, int max_pts, int i, double t,
                                              double (&row)[G], double& valid_lo, double& valid_hi)
{
    const int L      = (int)ct[0];
    const int stride = max_pts + (2 + max_pts) * G * G;
    double lo = -DBL_MAX, hi = DBL_MAX;
#pragma unroll
    for (int j = 0; j < G; ++j) {
        row[j] = 0.0;
    }
    for (int l = 0; l < L; ++l) {
        const int K         = (int)ct[1 + l];
        const double* p     = ct + 1 + max_loc + l * stride;
        const double* times = p;
        const double* B     = p + max_pts + i * G;
        const double* M     = B + G * G;
        const double* D     = p + max_pts + 2 * G * G + i * G;
        int k               = 0;
        while (k < K && !(t < times[k])) {
            ++k;
        }
        // left = accumulated damping before the next change point, right = after
        const double* Dl = k > 0 ? D + (k - 1) * G * G : nullptr;
        const double* Dr = k < K ? D + k * G * G : Dl;
        const double xr  = k < K ? times[k] : DBL_MAX;
        const double xl  = xr - 1.0;
        const bool ramp  = (t > xl) && (t < xr);
        double c         = 0.0;
        if (ramp) {
            c = cos(3.141592653589793238462643383279502884 / (xr - xl) * (t - xl));
            lo = DBL_MAX; // not cacheable
            hi = -DBL_MAX;
        }
        else {
            if (k > 0) {
                lo = fmax(lo, times[k - 1]);
            }
            hi = fmin(hi, xl);
        }
#pragma unroll
        for (int j = 0; j < G; ++j) {
            const double yl = Dl ? Dl[j] : 0.0;
            const double yr = Dr ? Dr[j] : 0.0;
            double d;
            if (t <= xl) {
                d = yl;
            }
            else if (t >= xr) {
                d = yr;
            }
            else {
                d = __dadd_rn(__dmul_rn(__dmul_rn(0.5, __dsub_rn(yl, yr)), c), __dmul_rn(0.5, __dadd_rn(yl, yr)));
            }
            const double e = __dsub_rn(B[j], __dmul_rn(d, __dsub_rn(B[j], M[j])));
            row[j]         = __dadd_rn(row[j], e);
        }
    }
    valid_lo = lo;
    valid_hi = hi;
}

// C(t) of one contact table by the lanes of a warp, entry e = i*G + j per lane (contact_matrix.h:392-400, :190-194;
// damping.h:381-396): written to cm[G*G]; returns the closed interval of times for which it stays the same (empty,
// lo > hi, while any location is inside a cosine ramp: then cm is not written and the lanes evaluate their row at
// every stage time, see contact_row_slow). Same arithmetic as the reference, no FMA contraction.
template <int G>
__device__ __forceinline__ void warp_contact_matrix(const double* ct, int max_loc, int max_pts, double t, double* cm,
                                                    double& valid_lo, double& valid_hi)
{
    const int lane   = threadIdx.x & 31;
    const int L      = (int)ct[0];
    const int stride = max_pts + (2 + max_pts) * G * G;
    double lo = -DBL_MAX, hi = DBL_MAX;
    bool ramp = false;
    for (int l = 0; l < L; ++l) {
        const int K         = (int)ct[1 + l];
        const double* times = ct + 1 + max_loc + l * stride;
        int k               = 0;
        while (k < K && !(t < times[k])) {
            ++k;
        }
        const double xr = k < K ? times[k] : DBL_MAX;
        const double xl = xr - 1.0;
        if ((t > xl) && (t < xr)) {
            ramp = true;
        }
        else {
            if (k > 0) {
                lo = fmax(lo, times[k - 1]);
            }
            hi = fmin(hi, xl);
        }
    }
    if (ramp) {
        valid_lo = DBL_MAX;
        valid_hi = -DBL_MAX;
        return;
    }
    for (int e = lane; e < G * G; e += 32) {
        double acc = 0.0;
        for (int l = 0; l < L; ++l) {
            const int K         = (int)ct[1 + l];
            const double* p     = ct + 1 + max_loc + l * stride;
            const double* times = p;
            int k               = 0;
            while (k < K && !(t < times[k])) {
                ++k;
            }
            const double* D = p + max_pts + 2 * G * G;
            const double xr = k < K ? times[k] : DBL_MAX;
            const double xl = xr - 1.0;
            const double yl = k > 0 ? D[(k - 1) * G * G + e] : 0.0;
            const double yr = k < K ? D[k * G * G + e] : yl;
            const double d  = (t <= xl) ? yl : yr;
            const double B  = p[max_pts + e];
            const double M  = p[max_pts + G * G + e];
            acc             = __dadd_rn(acc, __dsub_rn(B, __dmul_rn(d, __dsub_rn(B, M))));
        }
        cm[e] = acc;
    }
    valid_lo = lo;
    valid_hi = hi;
}

// 1 - D(t) factor of the mobility coefficients for the C compartments of age group i:
// coeff_e(t) = B_e - D(t)*(B_e - 0), MobilityCoefficients = DampingMatrixExpression<VectorDampings>
// (metapopulation_mobility_instant.h:103-110). mob == nullptr or K == 0: D = 0.
static __device__ __noinline__ void mobility_damping_eval(const double* mob, int mob_max_pts, int n, int i, double t,
                                                   double (&d)[NC])
{
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        d[c] = 0.0;
    }
    if (mob == nullptr) {
        return;
    }
    const int K = (int)mob[0];
    if (K == 0) {
        return;
    }
    const double* times = mob + 1;
    const double* D     = mob + 1 + mob_max_pts + i * NC;
    int k               = 0;
    while (k < K && !(t < times[k])) {
        ++k;
    }
    const double* Dl = k > 0 ? D + (k - 1) * n : nullptr;
    const double* Dr = k < K ? D + k * n : Dl;
    const double xr  = k < K ? times[k] : DBL_MAX;
    const double xl  = xr - 1.0;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        d[c] = smoother_cosine(t, xl, xr, Dl ? Dl[c] : 0.0, Dr ? Dr[c] : 0.0);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// TMA bulk copy (cp.async.bulk, SASS UBLKCP) of one run's parameter block into shared memory.
__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void stage_run_block(double* smem_dst, const double* gmem_src, int bytes, uint64_t* mbar)
{
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar)), "r"(bytes)
                     : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(smem_dst)),
                     "l"(gmem_src), "r"(bytes), "r"(smem_u32(mbar))
                     : "memory");
    }
    // every thread waits for phase 0 of the barrier
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n"
                     ".reg .pred p;\n"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
                     "selp.u32 %0, 1, 0, p;\n"
                     "}\n"
                     : "=r"(done)
                     : "r"(smem_u32(mbar)), "r"(0)
                     : "memory");
    }
}

// osecir per-lane coefficients kept in registers
struct LaneCoef {
    double rE_INS, rINS_ISy, rINS_R, rISy_ISev, rISy_R, rISev_ICr0, rISev_R, rISev_D0, rICr_D, rICr_R, rTNT, rP,
        rREL, rRISK, rMAXRISK, rREDEXP;
    __device__ __forceinline__ void load(const double* a)
    {
        rE_INS     = a[miocu::rE_INS];
        rINS_ISy   = a[miocu::rINS_ISy];
        rINS_R     = a[miocu::rINS_R];
        rISy_ISev  = a[miocu::rISy_ISev];
        rISy_R     = a[miocu::rISy_R];
        rISev_ICr0 = a[miocu::rISev_ICr0];
        rISev_R    = a[miocu::rISev_R];
        rISev_D0   = a[miocu::rISev_D0];
        rICr_D     = a[miocu::rICr_D];
        rICr_R     = a[miocu::rICr_R];
        rTNT       = a[miocu::rTNT];
        rP         = a[miocu::rP];
        rREL       = a[miocu::rREL];
        rRISK      = a[miocu::rRISK];
        rMAXRISK   = a[miocu::rMAXRISK];
        rREDEXP    = a[miocu::rREDEXP];
    }
};

// flows -> dydt in the reference's accumulation order (flow_model.h:219-235 over the TypeList of model.h:45-59)
__device__ __forceinline__ void flows_to_dydt(double fSE, double fE_INS, double fINS_ISy, double fINS_R,
                                              double fINSC_ISyC, double fINSC_R, double fISy_ISev, double fISy_R,
                                              double fISyC_ISev, double fISyC_R, double fISev_ICr, double fISev_R,
                                              double fISev_D, double fICr_D, double fICr_R, double (&k)[NC])
{
    k[cS]    = 0.0 - fSE;
    k[cE]    = fSE - fE_INS;
    k[cINS]  = (fE_INS - fINS_ISy) - fINS_R;
    k[cINSC] = (0.0 - fINSC_ISyC) - fINSC_R;
    k[cISy]  = (fINS_ISy - fISy_ISev) - fISy_R;
    k[cISyC] = (fINSC_ISyC - fISyC_ISev) - fISyC_R;
    k[cISev] = (((fISy_ISev + fISyC_ISev) - fISev_ICr) - fISev_R) - fISev_D;
    k[cICr]  = (fISev_ICr - fICr_D) - fICr_R;
    k[cR]    = ((((fINS_R + fINSC_R) + fISy_R) + fISyC_R) + fISev_R) + fICr_R;
    k[cD]    = fISev_D + fICr_D;
}

// osecirvvs: the recoveries of every layer flow into SusceptibleImprovedImmunity (ode_secirvvs/model.h:40-88), i.e.
// into the S slot of the layer-2 lane of the same age group; the Recovered slot itself stays empty.
template <int G, int LAY>
__device__ __forceinline__ void route_recoveries(const Group<G, LAY>& g, double (&k)[NC])
{
    if constexpr (LAY > 1) {
        const double rec = g.sum_layers(k[cR]);
        k[cR]            = 0.0;
        if (g.layer == LAY - 1) {
            k[cS] += rec;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Right-hand side of one node as seen by one lane: cpp/models/ode_secir/model.h:82-214 (LAY = 1),
// cpp/models/ode_secirvvs/model.h:111-520 (LAY = 3) + cpp/memilio/compartments/flow_model.h:112-144.
template <int G, int LAY>
struct NodeRhs {
    // LAY == 1: the whole row i of C(t); LAY > 1: the layer lanes of an age group split the row, lane (layer, i) keeps
    // the entries j = q*LAY + layer
    static constexpr int Q = LAY == 1 ? G : (G + LAY - 1) / LAY;
    LaneCoef c;
    const double* age_rec; // shared memory, for the rarely used slots
    const double* ctable;  // shared memory
    double start_day, seasonality;
    double p_eff; // TransmissionProbabilityOnContact in effect (osecirvvs apply_variant changes it during advance)
    double tnt_cap, tnt_cap_hi, tnt_cap_hi_ns, icu_lo, icu_cap;
    double crow[Q];
    double valid_lo, valid_hi;
    int ct_max_loc, ct_max_pts;

    __device__ __forceinline__ double season_at(double ts) const
    {
        // model.h:128-130
        if (seasonality == 0.0) {
            return 1.0;
        }
        return 1 + seasonality * sin(3.141592653589793238462643383279502884 * ((start_day + ts) / 182.5 + 0.5));
    }

    __device__ __forceinline__ void contact_row(const Group<G, LAY>& g, double ts)
    {
        if (!(ts >= valid_lo && ts <= valid_hi)) {
            if constexpr (LAY == 1) {
                contact_row_eval<G>(ctable, ct_max_loc, ct_max_pts, g.age, ts, crow, valid_lo, valid_hi);
            }
            else {
                double row[G];
                contact_row_eval<G>(ctable, ct_max_loc, ct_max_pts, g.age, ts, row, valid_lo, valid_hi);
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    const int j = q * LAY + g.layer;
                    crow[q]     = j < G ? row[j] : 0.0;
                }
            }
        }
    }

    // sum_j C(t)(i,j) * divN_j * (infectious weight of age group j) with pop = y:
    // osecir  model.h:120-145 (risk and relative transmission are those of the infectious group j),
    // osecirvvs model.h:229-260 (both risk factors are those of the susceptible group i)
    __device__ __forceinline__ double infection_pressure(const Group<G, LAY>& g, const double (&y)[NC], double tnt) const
    {
        if constexpr (LAY == 1) {
            const double Nj   = y[cS] + y[cE] + y[cINS] + y[cISy] + y[cISev] + y[cICr] + y[cR];
            const double divN = (Nj < 1e-12) ? 0.0 : 1.0 / Nj;
            const double risk = smoother_cosine(tnt, tnt_cap, tnt_cap_hi, c.rRISK, c.rMAXRISK);
            const double w    = divN * (c.rREL * y[cINS] + risk * y[cISy]);
            double lam        = 0.0;
#pragma unroll
            for (int j = 0; j < G; ++j) {
                lam += crow[j] * g.get(w, j);
            }
            return lam;
        }
        else {
            // N_j counts every living compartment of the three layers, the *Confirmed ones included (model.h:244-248)
            const double npart =
                y[cS] + y[cE] + y[cINS] + y[cISy] + y[cISev] + y[cICr] + y[cINSC] + y[cISyC];
            const double Nj    = g.sum_layers(npart);
            const double INSj  = g.sum_layers(y[cINS]);
            const double ISyj  = g.sum_layers(y[cISy]);
            const double divN  = (Nj < 1e-12) ? 0.0 : 1.0 / Nj;
            const double w_ins = divN * INSj, w_isy = divN * ISyj;
            const double r_sy  = smoother_cosine(tnt, tnt_cap, tnt_cap_hi, c.rRISK, c.rMAXRISK);
            const double r_ns  = smoother_cosine(tnt, tnt_cap, tnt_cap_hi_ns, c.rREL, 1.0);
            double part        = 0.0;
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const int j  = q * LAY + g.layer;
                const int jj = j < G ? j : 0;
                part += crow[q] * (r_ns * g.get(w_ins, jj) + r_sy * g.get(w_isy, jj));
            }
            return g.sum_layers(part);
        }
    }

    // test_and_trace_required and icu_occupancy of pop = y (osecir model.h:90-97, osecirvvs model.h:199-206)
    __device__ __forceinline__ void node_loads(const Group<G, LAY>& g, const double (&y)[NC], double& tnt,
                                               double& icu) const
    {
        if constexpr (LAY == 1) {
            tnt = g.sum(c.rTNT * y[cINS]);
        }
        else {
            tnt = g.sum(c.rTNT * (y[cINS] + y[cINSC]));
        }
        icu = g.sum(y[cICr]);
    }

    // ICU capacity shortage (model.h:148-153): InfectedSevere -> InfectedCritical / Dead rates
    __device__ __forceinline__ void icu_rates(double icu, double& rISev_ICr, double& rISev_D) const
    {
        rISev_ICr = c.rISev_ICr0;
        rISev_D   = c.rISev_D0;
        if (icu > icu_lo) {
            const double crit = age_rec[rCRIT];
            const double adj  = smoother_cosine(icu, icu_lo, icu_cap, crit, 0.0);
            const double dadj = crit - adj;
            rISev_ICr         = adj / age_rec[rT_ISEV];
            rISev_D           = (age_rec[rDPS] + dadj) / age_rec[rT_ISEV];
        }
    }

    // dydt = f(y, ts) with pop == y; `season` is the seasonality factor at ts
    __device__ __forceinline__ void eval(const Group<G, LAY>& g, const double (&y)[NC], double ts, double season,
                                         double (&k)[NC])
    {
        contact_row(g, ts);
        double tnt, icu;
        node_loads(g, y, tnt, icu);
        const double lam = infection_pressure(g, y, tnt);
        double fSE;
        if constexpr (LAY == 1) {
            fSE = y[cS] * (season * p_eff) * lam;
        }
        else {
            fSE = (y[cS] * c.rREDEXP) * ((season * p_eff) * lam);
        }
        double rISev_ICr, rISev_D;
        icu_rates(icu, rISev_ICr, rISev_D);
        flows_to_dydt(fSE, c.rE_INS * y[cE], c.rINS_ISy * y[cINS], c.rINS_R * y[cINS], c.rINS_ISy * y[cINSC],
                      c.rINS_R * y[cINSC], c.rISy_ISev * y[cISy], c.rISy_R * y[cISy], c.rISy_ISev * y[cISyC],
                      c.rISy_R * y[cISyC], rISev_ICr * y[cISev], c.rISev_R * y[cISev], rISev_D * y[cISev],
                      c.rICr_D * y[cICr], c.rICr_R * y[cICr], k);
        route_recoveries(g, k);
    }
};

} // namespace miocu
