// reduce.cu -- ensemble post-processing on the device: the interpolated (daily) results of all runs are kept as one
// column-major matrix [column][run] and reduced at the end of the study.
//
// Replaces cpp/memilio/data/analyze_result.h:186-265 (interpolate_simulation_result: with dt_graph = 0.5 every integer
// day is a time point of the node series, so the day value is the state after that step's edge phase), :267-287
// (sum_nodes), :289-314 (ensemble_mean), :316-343 (ensemble_percentile: sort over runs, index size_t(num_runs * p)).
//
// Columns: (node k, day d, element e) -> (k * D + d) * El + e for the per-node results (El = n, or C when the age
// groups are summed), followed by the node-summed full-resolution result (day d, element j) -> N*D*El + d * n + j.
// The order statistic is found by an exact radix select on the order-preserving 64 bit image of the doubles, all
// requested percentiles in the same passes; no approximation. HBM bound: 6 passes over the column.
#include "kernels.h"

#include <cstdint>

namespace miocu
{

// ---------------------------------------------------------------------------------------------------------------
// y[R][N][n] at an integer day -> ens[col][run_offset + r]
__global__ void __launch_bounds__(256) record_ensemble_kernel(const double* __restrict__ y, double* __restrict__ ens,
                                                              int N, int G, int C, int lay, int num_runs, int day,
                                                              int num_days, int age_sum, long long run_offset,
                                                              long long stride)
{
    extern __shared__ double tile[]; // [32][n + 1], reference element order
    const int n     = G * C;
    const int n_dev = G * lay * NC;
    const int node  = blockIdx.y;
    const int r0    = blockIdx.x * 32;
    const int rows  = min(32, num_runs - r0);
    for (int idx = threadIdx.x; idx < rows * n_dev; idx += blockDim.x) {
        const int rr = idx / n_dev, d = idx - rr * n_dev;
        int e        = d;
        if (lay != 1) {
            const int gl = d / NC, slot = d - gl * NC;
            const int layer = gl / G, age = gl - layer * G;
            const int comp  = vvs_ref_comp(layer, slot);
            e               = comp < 0 ? -1 : age * C + comp;
        }
        if (e >= 0) {
            tile[rr * (n + 1) + e] = y[((size_t)(r0 + rr) * N + node) * n_dev + d];
        }
    }
    __syncthreads();
    const int El   = age_sum ? C : n;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (lane < rows) {
        for (int el = w; el < El; el += nw) {
            double v;
            if (age_sum) {
                v = 0.0;
                for (int a = 0; a < G; ++a) {
                    v += tile[lane * (n + 1) + a * C + el];
                }
            }
            else {
                v = tile[lane * (n + 1) + el];
            }
            ens[((size_t)(node * num_days + day) * El + el) * stride + run_offset + r0 + lane] = v;
        }
    }
}

// sum_nodes: ens[base + day*n + j][run] = sum_k y[run][k][j], nodes added in index order
__global__ void __launch_bounds__(256) record_node_sum_kernel(const double* __restrict__ y, double* __restrict__ ens,
                                                              int N, int G, int C, int lay, int num_runs, int day,
                                                              long long base_col, long long run_offset,
                                                              long long stride)
{
    const int run   = blockIdx.x;
    const int j     = threadIdx.x;
    const int n     = G * C;
    const int n_dev = G * lay * NC;
    if (run >= num_runs || j >= n) {
        return;
    }
    const int age   = j / C;
    const double* p = y + (size_t)run * N * n_dev + dev_index_of_ref(lay, G, age, j - age * C);
    double s        = 0.0;
    for (int k = 0; k < N; ++k) {
        s += p[(size_t)k * n_dev];
    }
    ens[(size_t)(base_col + (long long)day * n + j) * stride + run_offset + run] = s;
}

// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long ordered_key(double v)
{
    unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_to_double(unsigned long long k)
{
    unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

// One CTA per column; exact radix select of all requested order statistics in the same passes over the digits, from
// the most significant. Percentiles whose prefixes agree so far share one histogram, so an element contributes to at
// most one histogram per pass, and equal (histogram, digit) pairs are aggregated inside the warp before the shared
// memory atomic (columns whose runs agree -- empty compartments -- would otherwise serialise on one counter).
// STAGED: the column's keys are first copied to shared memory (one pass over HBM, which also yields the mean) and the
// digit passes (8 bits each, stopping as soon as every order statistic is alone in its bin) read shared memory.
// Otherwise (columns too long for shared memory) the digit passes (11/10 bits) read the column from global memory.
template <bool STAGED>
__global__ void __launch_bounds__(SEL_THREADS) select_percentiles_kernel(SelectArgs a)
{
    constexpr int BITS = STAGED ? 8 : SEL_BITS;
    constexpr int BINS = 1 << BITS;
    extern __shared__ __align__(16) unsigned char sel_smem[];
    unsigned int(*hist)[BINS] = reinterpret_cast<unsigned int(*)[BINS]>(sel_smem);               // [num_p][BINS]
    unsigned long long* keys  = reinterpret_cast<unsigned long long*>(sel_smem + sizeof(unsigned int) * BINS * SEL_MAX_P);
    __shared__ unsigned long long prefix[SEL_MAX_P]; // of percentile q
    __shared__ long long remaining[SEL_MAX_P];
    __shared__ unsigned int bin_count[SEL_MAX_P];    // size of the bin percentile q fell into in the last pass
    __shared__ int hist_of[SEL_MAX_P];               // histogram used by percentile q (first q with the same prefix)
    __shared__ double red[SEL_THREADS];
    const long long col = blockIdx.x;
    const int tid       = threadIdx.x;
    const int lane      = tid & 31;
    const int total     = (int)a.total;
    // ensemble_mean: sum of value / num_runs (fixed summation tree => deterministic); staging pass
    if (a.out_mean != nullptr || STAGED) {
        double s           = 0.0;
        const double denom = (double)a.total;
        long long base     = 0;
        for (int sg = 0; sg < a.num_segments; ++sg) {
            const double* p = a.seg_ptr[sg] + (size_t)col * a.seg_stride[sg];
            // eight independent loads in flight per thread (the column streams in from HBM once)
            constexpr int UNROLL = 8;
            const long long cnt  = a.seg_count[sg];
            for (long long i0 = tid; i0 < cnt; i0 += (long long)UNROLL * SEL_THREADS) {
                double v[UNROLL];
#pragma unroll
                for (int j = 0; j < UNROLL; ++j) {
                    const long long i = i0 + (long long)j * SEL_THREADS;
                    v[j]              = i < cnt ? p[i] : 0.0;
                }
#pragma unroll
                for (int j = 0; j < UNROLL; ++j) {
                    const long long i = i0 + (long long)j * SEL_THREADS;
                    if (i < cnt) {
                        s += v[j] / denom;
                        if (STAGED) {
                            keys[base + i] = ordered_key(v[j]);
                        }
                    }
                }
            }
            base += cnt;
        }
        if (a.out_mean != nullptr) {
            red[tid] = s;
            __syncthreads();
            for (int o = SEL_THREADS / 2; o > 0; o >>= 1) {
                if (tid < o) {
                    red[tid] += red[tid + o];
                }
                __syncthreads();
            }
            if (tid == 0) {
                a.out_mean[col] = red[0];
            }
        }
    }
    if (a.out_select == nullptr || a.num_p == 0) {
        return;
    }
    if (tid < a.num_p) {
        prefix[tid]    = 0ull;
        remaining[tid] = a.rank[tid];
        bin_count[tid] = 0xffffffffu;
        hist_of[tid]   = 0;
    }
    __syncthreads();
    const int num_p = a.num_p;
    int shift       = 64;
    while (shift > 0) {
        int bits;
        if (STAGED) {
            bits = BITS;
        }
        else {
            bits = shift >= 20 + SEL_BITS ? SEL_BITS : (shift > 10 ? 10 : shift); // 11,11,11,11,10,10
        }
        shift -= bits;
        const unsigned int nb = 1u << bits;
        for (int i = tid; i < num_p * BINS; i += SEL_THREADS) {
            (&hist[0][0])[i] = 0u;
        }
        __syncthreads();
        // distinct prefixes of this pass (registers): histogram u belongs to the first percentile with that prefix
        unsigned long long upre[SEL_MAX_P];
        int nu = 0;
#pragma unroll
        for (int q = 0; q < SEL_MAX_P; ++q) {
            if (q < num_p && hist_of[q] == q) {
                upre[nu++] = prefix[q];
            }
        }
        const bool first = (shift + bits) == 64;
        const unsigned long long pre_lo = nu > 0 ? upre[0] : 0ull, pre_hi = nu > 0 ? upre[nu - 1] : 0ull;
        auto visit = [&](unsigned long long k) {
            const unsigned int digit = (unsigned int)(k >> shift) & (nb - 1u);
            int u = -1;
            if (first) {
                u = 0;
            }
            else {
                // the distinct prefixes ascend with the percentiles: most keys fall outside [first, last] after the
                // leading digits and are rejected with two comparisons
                const unsigned long long hi = k >> (shift + bits);
                if (hi >= pre_lo && hi <= pre_hi) {
#pragma unroll
                    for (int j = 0; j < SEL_MAX_P; ++j) {
                        if (j < nu && hi == upre[j]) {
                            u = j;
                        }
                    }
                }
            }
            // warp aggregation of equal (histogram, digit) pairs among the lanes that count
            if (u >= 0) {
                // columns whose runs agree (empty compartments) would serialise on one counter: when every counting
                // lane of the warp has the same (histogram, digit) one lane adds the count, else plain atomics
                const unsigned int code   = ((unsigned int)u << 16) | digit;
                const unsigned int act    = __activemask();
                const int leader          = __ffs(act) - 1;
                const unsigned int code0  = __shfl_sync(act, code, leader);
                if (__all_sync(act, code == code0)) {
                    if (lane == leader) {
                        atomicAdd(&hist[u][digit], (unsigned int)__popc(act));
                    }
                }
                else {
                    atomicAdd(&hist[u][digit], 1u);
                }
            }
        };
        if (STAGED) {
            for (int i = tid; i < total; i += SEL_THREADS) {
                visit(keys[i]);
            }
        }
        else {
            for (int sg = 0; sg < a.num_segments; ++sg) {
                const double* p = a.seg_ptr[sg] + (size_t)col * a.seg_stride[sg];
                // eight independent loads in flight per thread
                constexpr int UNROLL = 8;
                const long long cnt  = a.seg_count[sg];
                for (long long i0 = tid; i0 < cnt; i0 += (long long)UNROLL * SEL_THREADS) {
                    double v[UNROLL];
#pragma unroll
                    for (int j = 0; j < UNROLL; ++j) {
                        const long long i = i0 + (long long)j * SEL_THREADS;
                        v[j]              = i < cnt ? p[i] : 0.0;
                    }
#pragma unroll
                    for (int j = 0; j < UNROLL; ++j) {
                        if (i0 + (long long)j * SEL_THREADS < cnt) {
                            visit(ordered_key(v[j]));
                        }
                    }
                }
            }
        }
        __syncthreads();
        // warp q locates the digit of percentile q: per-lane partial sums, warp scan, then a short in-lane search.
        // histogram index of q: position of hist_of[q] among the distinct ones
        {
            const int wq = tid >> 5;
            if (wq < num_p) {
                int u = 0;
                for (int q = 0; q < hist_of[wq]; ++q) {
                    u += (hist_of[q] == q) ? 1 : 0;
                }
                const unsigned int* h = hist[u];
                const int per         = (int)(nb / 32u);
                unsigned int sum      = 0;
                for (int i = 0; i < per; ++i) {
                    sum += h[lane * per + i];
                }
                unsigned int inc = sum;
                for (int o = 1; o < 32; o <<= 1) {
                    const unsigned int t = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) {
                        inc += t;
                    }
                }
                const unsigned int exc = inc - sum;
                const long long rem    = remaining[wq];
                const bool mine        = rem >= (long long)exc && rem < (long long)inc;
                const unsigned int who = __ballot_sync(0xffffffffu, mine);
                if (who != 0u && lane == __ffs(who) - 1) {
                    long long r    = rem - exc;
                    unsigned int d = lane * per;
                    unsigned int c = 0;
                    for (;; ++d) {
                        c = h[d];
                        if (r < (long long)c) {
                            break;
                        }
                        r -= c;
                    }
                    remaining[wq] = r;
                    bin_count[wq] = c;
                    prefix[wq]    = (prefix[wq] << bits) | (unsigned long long)d;
                }
            }
        }
        __syncthreads();
        if (tid == 0) {
            for (int q = 0; q < num_p; ++q) {
                int hq = q;
                for (int j = 0; j < q; ++j) {
                    if (prefix[j] == prefix[q]) {
                        hq = j;
                        break;
                    }
                }
                hist_of[q] = hq;
            }
        }
        __syncthreads();
        if (STAGED && shift > 0) {
            // every order statistic alone in its bin: its key is the one element with that prefix
            bool alone = true;
            for (int q = 0; q < num_p; ++q) {
                alone = alone && bin_count[q] == 1u;
            }
            if (alone) { // uniform over the CTA
                for (int i = tid; i < total; i += SEL_THREADS) {
                    const unsigned long long k  = keys[i];
                    const unsigned long long hi = k >> shift;
                    for (int q = 0; q < num_p; ++q) {
                        if (hi == prefix[q]) {
                            a.out_select[(size_t)q * a.num_columns + col] = key_to_double(k);
                        }
                    }
                }
                return;
            }
        }
    }
    if (tid < a.num_p) {
        a.out_select[(size_t)tid * a.num_columns + col] = key_to_double(prefix[tid]);
    }
}

cudaError_t launch_record_ensemble(const double* y, double* ens, int N, int G, int C, int lay, int num_runs, int day,
                                   int num_days, int age_sum, long long run_offset, long long stride,
                                   cudaStream_t stream)
{
    const int n = G * C;
    dim3 grid((num_runs + 31) / 32, N);
    const size_t shmem = sizeof(double) * 32 * (n + 1);
    if (shmem > 48 * 1024) {
        cudaFuncSetAttribute(record_ensemble_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem);
    }
    record_ensemble_kernel<<<grid, 256, shmem, stream>>>(y, ens, N, G, C, lay, num_runs, day, num_days, age_sum,
                                                         run_offset, stride);
    const long long base = (long long)N * num_days * (age_sum ? C : n);
    record_node_sum_kernel<<<num_runs, ((n + 31) / 32) * 32, 0, stream>>>(y, ens, N, G, C, lay, num_runs, day, base,
                                                                         run_offset, stride);
    return cudaGetLastError();
}

cudaError_t launch_select(const SelectArgs& a, cudaStream_t stream)
{
    if (a.num_columns == 0) {
        return cudaSuccess;
    }
    // columns that fit into shared memory (two CTAs per SM) are staged there: one pass over HBM
    const size_t hist_staged = sizeof(unsigned int) * 256 * SEL_MAX_P;
    const size_t staged      = hist_staged + sizeof(unsigned long long) * (size_t)a.total;
    if (staged <= 200 * 1024) { // (two CTAs per SM up to 110 KB, one beyond)
        cudaFuncSetAttribute(select_percentiles_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)staged);
        select_percentiles_kernel<true><<<(unsigned)a.num_columns, SEL_THREADS, staged, stream>>>(a);
    }
    else {
        const int shmem = (int)sizeof(unsigned int) * SEL_BINS * SEL_MAX_P;
        cudaFuncSetAttribute(select_percentiles_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, shmem);
        select_percentiles_kernel<false><<<(unsigned)a.num_columns, SEL_THREADS, shmem, stream>>>(a);
    }
    return cudaGetLastError();
}

} // namespace miocu
