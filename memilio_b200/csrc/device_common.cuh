// device_common.cuh -- shared device helpers of the B200 Graph-ODE backend (sm_100a only).
//
// Lane mapping used by every kernel: a warp owns NPW = 32/G consecutive (run, node) pairs; the G lanes of a "group"
// own one age group each and keep that age group's C compartments in registers. Everything that couples age groups
// (force of infection, test-and-trace load, ICU occupancy, error norm) is a G-lane exchange via warp shuffles.
#pragma once

#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>

namespace miocu
{

constexpr int NC      = 10; // osecir compartments, cpp/models/ode_secir/infection_state.h:33-47
constexpr int AGE_REC = 20; // derived per-age record (doubles), see host_params.cpp
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
    rDPS           // DeathsPerSevere
};
// header slots
enum : int
{
    hSTART_DAY = 0,
    hSEASONALITY,
    hTNT_MAX_RISK,
    hNONDETECTION,
    hSTART_DETECTION,
    hEND_DETECTION
};

// problem geometry shared by all kernels (passed by value)
struct Geometry {
    int G, N, E, n;          // groups, nodes, edges, n = G*NC
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
    double* t_node;            // [R][N] node-local simulation time (result.get_last_time())
    double* dt_node;           // [R][N] SimulationBase::m_dt
    int* counters;             // [R][N][4]
};

__device__ __forceinline__ double shfl(double v, int src)
{
    return __shfl_sync(0xffffffffu, v, src);
}

// G-lane group helpers. Lanes >= NPW*G form no group and idle.
template <int G>
struct Group {
    static constexpr int NPW = 32 / G;
    int lane, gl, gbase;
    __device__ __forceinline__ Group()
    {
        lane  = threadIdx.x & 31;
        gl    = lane % G;
        gbase = lane - gl;
        if (lane >= NPW * G) { // spare lanes talk to themselves
            gl    = 0;
            gbase = lane;
        }
    }
    __device__ __forceinline__ bool spare() const
    {
        return lane >= NPW * G;
    }
    // value of group lane j
    __device__ __forceinline__ double get(double v, int j) const
    {
        int src = spare() ? lane : gbase + j;
        return shfl(v, src);
    }
    // all-reduce sum over the group (pairwise tree where G allows, otherwise sequential gather)
    __device__ __forceinline__ double sum(double v) const
    {
        if constexpr (G == 1) {
            return v;
        }
        else if constexpr (G == 2 || G == 4 || G == 8) {
#pragma unroll
            for (int o = 1; o < G; o <<= 1) {
                v += shfl(v, spare() ? lane : (gbase + (gl ^ o)));
            }
            return v;
        }
        else if constexpr (G == 6) {
            // pairs (gbase is even for G=6), then the two other pair sums
            v += shfl(v, spare() ? lane : (lane ^ 1));
            int s1   = gl + 2 >= G ? gl + 2 - G : gl + 2;
            int s2   = gl + 4 >= G ? gl + 4 - G : gl + 4;
            double a = shfl(v, spare() ? lane : gbase + s1);
            double b = shfl(v, spare() ? lane : gbase + s2);
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
    __device__ __forceinline__ double max(double v) const
    {
        if constexpr (G == 1) {
            return v;
        }
        else {
            double m = v;
#pragma unroll
            for (int j = 1; j < G; ++j) {
                int s = gl + j >= G ? gl + j - G : gl + j;
                m     = fmax(m, shfl(v, spare() ? lane : gbase + s));
            }
            return m;
        }
    }
    __device__ __forceinline__ bool any(bool p) const
    {
        unsigned b    = __ballot_sync(0xffffffffu, p);
        unsigned mask = spare() ? (1u << lane) : (((1u << G) - 1u) << gbase);
        return (b & mask) != 0u;
    }
};

// cpp/memilio/math/smoother.h:32-55
__device__ __forceinline__ double smoother_cosine(double x, double xleft, double xright, double yleft, double yright)
{
    if (x <= xleft) {
        return yleft;
    }
    if (x >= xright) {
        return yright;
    }
    return 0.5 * (yleft - yright) * cos(3.141592653589793238462643383279502884 / (xright - xleft) * (x - xleft)) +
           0.5 * (yleft + yright);
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
        rREL, rRISK, rMAXRISK;
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

} // namespace miocu
