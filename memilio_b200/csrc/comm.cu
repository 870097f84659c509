// comm.cu -- the multi-GPU part of ParameterStudy::run behind the C-ABI: cpp/memilio/compartments/parameter_studies.h
// :163-213 shards the runs over MPI ranks (distribute_runs :354-367) and gathers the results on the root; here one
// process drives one GPU, the ensemble record of a rank holds its block of runs, and the post-processing
// (data/analyze_result.h:289-343) needs every column over the runs of ALL ranks. mio_cuda_ensemble_reduce_sharded
// re-shards the records from runs to columns with one grouped ncclSend/ncclRecv exchange over NVLink, selects the exact
// order statistics and the mean of the rank's column block (reduce.cu, multi-segment selection: the received blocks
// are used where they land, no concatenation), and gathers the reduced rows on rank 0.
// NCCL is bound at run time (dlopen of libnccl.so.2, the copy the process already has when torch is loaded), so the
// library itself has no link-time dependency on it.
#include "../../include/mio_cuda.h"
#include "kernels.h"

#include <dlfcn.h>

#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

int mio_set_last_error(int code, const std::string& msg);

namespace
{

// the slice of nccl.h this file needs (NCCL 2.x ABI)
typedef struct ncclComm* ncclComm_t;
typedef struct {
    char internal[128];
} ncclUniqueId;
typedef int ncclResult_t;
enum { ncclInt64 = 4, ncclFloat64 = 8 };

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*)                                               = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int)                         = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t)                                                  = nullptr;
    ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t)             = nullptr;
    ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t)                   = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t)      = nullptr;
    ncclResult_t (*GroupStart)()                                                             = nullptr;
    ncclResult_t (*GroupEnd)()                                                               = nullptr;
    const char* (*GetErrorString)(ncclResult_t)                                              = nullptr;
    std::string error;
};

NcclApi& nccl()
{
    static NcclApi api;
    if (api.lib || !api.error.empty()) {
        return api;
    }
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
        api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (api.lib) {
            break;
        }
    }
    if (!api.lib) {
        api.error = "NCCL library not found (libnccl.so.2)";
        return api;
    }
    auto sym = [&](const char* n) {
        void* p = dlsym(api.lib, n);
        if (!p && api.error.empty()) {
            api.error = std::string("NCCL symbol missing: ") + n;
        }
        return p;
    };
    api.GetUniqueId    = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank   = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy    = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.Send           = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
    api.Recv           = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
    api.AllGather      = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.GroupStart     = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd       = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    return api;
}

template <class T>
struct DevBuf { // grow-only device scratch
    T* p       = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t count)
    {
        if (count <= cap) {
            return cudaSuccess;
        }
        cudaFree(p);
        p   = nullptr;
        cap = 0;
        cudaError_t er = cudaMalloc((void**)&p, count * sizeof(T));
        if (er == cudaSuccess) {
            cap = count;
        }
        return er;
    }
    ~DevBuf()
    {
        cudaFree(p);
    }
};

} // namespace

struct mio_comm {
    int device = 0, rank = 0, world = 1;
    ncclComm_t comm     = nullptr;
    cudaStream_t stream = nullptr;
    DevBuf<double> recv, pack, rows, gathered;
    DevBuf<long long> counts;
};

// (capi.cu) the fields of the ensemble record this file needs
extern "C" int mio_cuda_ensemble_info(mio_ensemble* e, int64_t* num_columns, int64_t* stride, int64_t* num_runs,
                                      void** device_ptr);

#define NCCL_TRY(expr)                                                                                                 \
    do {                                                                                                               \
        ncclResult_t r__ = (expr);                                                                                     \
        if (r__ != 0) {                                                                                                \
            return mio_set_last_error(MIO_UNKNOWN_ERROR, std::string(#expr) + ": " + nccl().GetErrorString(r__));     \
        }                                                                                                              \
    } while (0)
#define CU_TRY(expr)                                                                                                   \
    do {                                                                                                               \
        cudaError_t e__ = (expr);                                                                                      \
        if (e__ != cudaSuccess) {                                                                                      \
            return mio_set_last_error(MIO_UNKNOWN_ERROR, std::string(#expr) + ": " + cudaGetErrorString(e__));        \
        }                                                                                                              \
    } while (0)

extern "C" {

int mio_cuda_comm_get_unique_id(char* id)
{
    if (!id) {
        return mio_set_last_error(MIO_INVALID_VALUE, "null argument");
    }
    NcclApi& api = nccl();
    if (!api.error.empty()) {
        return mio_set_last_error(MIO_UNKNOWN_ERROR, api.error);
    }
    ncclUniqueId uid;
    NCCL_TRY(api.GetUniqueId(&uid));
    std::memcpy(id, uid.internal, MIO_COMM_ID_BYTES);
    return MIO_OK;
}

int mio_cuda_comm_create(int device, int rank, int world, const char* id, mio_comm** out)
{
    if (!id || !out || world < 1 || world > miocu::SEL_MAX_SEG || rank < 0 || rank >= world) {
        return mio_set_last_error(MIO_INVALID_VALUE, "invalid communicator arguments (1..8 ranks)");
    }
    *out         = nullptr;
    NcclApi& api = nccl();
    if (!api.error.empty()) {
        return mio_set_last_error(MIO_UNKNOWN_ERROR, api.error);
    }
    CU_TRY(cudaSetDevice(device));
    auto* c   = new mio_comm();
    c->device = device;
    c->rank   = rank;
    c->world  = world;
    ncclUniqueId uid;
    std::memcpy(uid.internal, id, MIO_COMM_ID_BYTES);
    ncclResult_t r = api.CommInitRank(&c->comm, world, uid, rank);
    if (r != 0) {
        delete c;
        return mio_set_last_error(MIO_UNKNOWN_ERROR, std::string("ncclCommInitRank: ") + api.GetErrorString(r));
    }
    cudaError_t er = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (er != cudaSuccess) {
        api.CommDestroy(c->comm);
        delete c;
        return mio_set_last_error(MIO_UNKNOWN_ERROR, cudaGetErrorString(er));
    }
    *out = c;
    return MIO_OK;
}

void mio_cuda_comm_destroy(mio_comm* c)
{
    if (!c) {
        return;
    }
    cudaSetDevice(c->device);
    if (c->comm) {
        nccl().CommDestroy(c->comm);
    }
    if (c->stream) {
        cudaStreamDestroy(c->stream);
    }
    delete c;
}

int mio_cuda_ensemble_reduce_sharded(mio_ensemble* e, mio_comm* c, int64_t num_runs_local, int64_t num_columns,
                                     int32_t num_p, const double* p, double* percentiles, double* mean)
{
    using namespace miocu;
    if (!c || num_p < 0 || num_p > SEL_MAX_P || (num_p > 0 && !p) || num_runs_local < 0 || num_columns < 0) {
        return mio_set_last_error(MIO_INVALID_VALUE, "invalid arguments");
    }
    NcclApi& api = nccl();
    CU_TRY(cudaSetDevice(c->device));
    // this rank's record (a rank without runs passes e = NULL, num_runs_local = 0 and the column count)
    int64_t ncol = num_columns, stride = 0, recorded = 0;
    void* dptr = nullptr;
    if (e) {
        if (int rc = mio_cuda_ensemble_info(e, &ncol, &stride, &recorded, &dptr)) {
            return rc;
        }
        if (num_runs_local > stride) {
            return mio_set_last_error(MIO_OUT_OF_RANGE, "more runs than the record holds");
        }
    }
    else if (num_runs_local != 0) {
        return mio_set_last_error(MIO_INVALID_VALUE, "runs without a record");
    }
    const int W = c->world, me = c->rank;
    CU_TRY(cudaDeviceSynchronize()); // the record kernels run on the handles' streams
    // run counts of all ranks
    CU_TRY(c->counts.reserve((size_t)W + 1));
    long long mine = num_runs_local;
    CU_TRY(cudaMemcpyAsync(c->counts.p + W, &mine, sizeof(long long), cudaMemcpyHostToDevice, c->stream));
    NCCL_TRY(api.AllGather(c->counts.p + W, c->counts.p, 1, ncclInt64, c->comm, c->stream));
    std::vector<long long> counts((size_t)W);
    CU_TRY(cudaMemcpyAsync(counts.data(), c->counts.p, sizeof(long long) * W, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    long long total = 0;
    for (long long v : counts) {
        total += v;
    }
    if (total == 0) {
        return mio_set_last_error(MIO_INVALID_VALUE, "no runs recorded on any rank");
    }
    // column blocks (distribute_runs rule)
    std::vector<long long> c0((size_t)W), cn((size_t)W);
    for (int r = 0; r < W; ++r) {
        const long long base = ncol / W, rem = ncol % W;
        cn[(size_t)r] = base + (r < rem ? 1 : 0);
        c0[(size_t)r] = (long long)r * base + std::min<long long>(r, rem);
    }
    const long long my_c0 = c0[(size_t)me], my_cn = cn[(size_t)me];
    // receive blocks [my_cn][counts[s]] back to back
    std::vector<size_t> roff((size_t)W + 1, 0);
    for (int s = 0; s < W; ++s) {
        roff[(size_t)s + 1] = roff[(size_t)s] + (size_t)my_cn * (size_t)counts[(size_t)s];
    }
    CU_TRY(c->recv.reserve(std::max<size_t>(roff[(size_t)W], 1)));
    // rows of the record are `stride` apart: blocks are sent as they lie when the record is full (stride == runs),
    // else packed first
    const double* src = static_cast<const double*>(dptr);
    if (e && stride != num_runs_local && num_runs_local > 0) {
        CU_TRY(c->pack.reserve((size_t)ncol * (size_t)num_runs_local));
        CU_TRY(cudaMemcpy2DAsync(c->pack.p, sizeof(double) * num_runs_local, src, sizeof(double) * stride,
                                 sizeof(double) * num_runs_local, (size_t)ncol, cudaMemcpyDeviceToDevice, c->stream));
        src = c->pack.p;
    }
    NCCL_TRY(api.GroupStart());
    for (int peer = 0; peer < W; ++peer) {
        if (peer == me) {
            continue;
        }
        const size_t nsend = (size_t)cn[(size_t)peer] * (size_t)num_runs_local;
        if (nsend > 0) {
            NCCL_TRY(api.Send(src + (size_t)c0[(size_t)peer] * (size_t)num_runs_local, nsend, ncclFloat64, peer, c->comm,
                              c->stream));
        }
        const size_t nrecv = (size_t)my_cn * (size_t)counts[(size_t)peer];
        if (nrecv > 0) {
            NCCL_TRY(api.Recv(c->recv.p + roff[(size_t)peer], nrecv, ncclFloat64, peer, c->comm, c->stream));
        }
    }
    NCCL_TRY(api.GroupEnd());
    // selection over the run segments of all ranks (this rank's own block is read in place)
    SelectArgs a{};
    a.num_segments = W;
    a.total        = total;
    for (int s = 0; s < W; ++s) {
        if (s == me) {
            a.seg_ptr[s]    = num_runs_local > 0 ? src + (size_t)my_c0 * (size_t)num_runs_local : nullptr;
            a.seg_stride[s] = num_runs_local;
        }
        else {
            a.seg_ptr[s]    = c->recv.p + roff[(size_t)s];
            a.seg_stride[s] = counts[(size_t)s];
        }
        a.seg_count[s] = counts[(size_t)s];
    }
    a.num_p = num_p;
    for (int q = 0; q < num_p; ++q) {
        if (!(p[q] > 0.0 && p[q] < 1.0)) {
            return mio_set_last_error(MIO_INVALID_VALUE, "percentile must lie in (0, 1)");
        }
        a.rank[q] = (long long)(size_t)((double)total * p[q]);
    }
    const int K = num_p + 1; // rows: percentiles, then the mean
    a.num_columns = my_cn;
    CU_TRY(c->rows.reserve(std::max<size_t>((size_t)K * (size_t)my_cn, 1)));
    a.out_select = num_p > 0 ? c->rows.p : nullptr;
    a.out_mean   = c->rows.p + (size_t)num_p * (size_t)my_cn;
    CU_TRY(launch_select(a, c->stream));
    // gather the rows on rank 0 (parameter_studies.h:190-212)
    if (me == 0) {
        CU_TRY(c->gathered.reserve(std::max<size_t>((size_t)K * (size_t)ncol, 1)));
    }
    NCCL_TRY(api.GroupStart());
    if (me == 0) {
        for (int s = 1; s < W; ++s) {
            if (cn[(size_t)s] > 0) {
                NCCL_TRY(api.Recv(c->gathered.p + (size_t)K * (size_t)c0[(size_t)s], (size_t)K * (size_t)cn[(size_t)s],
                                  ncclFloat64, s, c->comm, c->stream));
            }
        }
    }
    else if (my_cn > 0) {
        NCCL_TRY(api.Send(c->rows.p, (size_t)K * (size_t)my_cn, ncclFloat64, 0, c->comm, c->stream));
    }
    NCCL_TRY(api.GroupEnd());
    if (me == 0) {
        // rank s delivered [K][cn_s] at offset K*c0_s: scatter the rows into [K][ncol] host arrays
        for (int s = 0; s < W; ++s) {
            const double* blk = s == 0 ? c->rows.p : c->gathered.p + (size_t)K * (size_t)c0[(size_t)s];
            for (int q = 0; q < K; ++q) {
                double* dst = q < num_p ? (percentiles ? percentiles + (size_t)q * (size_t)ncol : nullptr) : mean;
                if (dst && cn[(size_t)s] > 0) {
                    CU_TRY(cudaMemcpyAsync(dst + c0[(size_t)s], blk + (size_t)q * (size_t)cn[(size_t)s],
                                           sizeof(double) * (size_t)cn[(size_t)s], cudaMemcpyDeviceToHost, c->stream));
                }
            }
        }
    }
    CU_TRY(cudaStreamSynchronize(c->stream));
    return MIO_OK;
}

} // extern "C"
