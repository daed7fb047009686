// Multi-GPU exchange of the free segments: C-layers shard by orientation, every rank builds its own layers
// with no collective on the way, and ONE ncclAllGather of a fixed-size packed block per rank gives every rank the CSR
// offsets and the xL / xU / xM arrays (= the roadmap vertices) of all layers, which bridge-layer pairing and the host's
// connectMultiSlice need (HRM3D.cpp:158-224).  Meshes are never exchanged.
//
// The block has a CAPACITY chosen by the caller (the same on all ranks), so the exchange needs no size round trip and no
// host synchronisation: k_pack_block writes {lines, total segments, off[line_cap+1], xL[seg_cap], xU[seg_cap], xM[seg_cap]}
// on the side stream right behind the free-segment passes, ncclAllGather follows on the same stream, and the main stream
// meanwhile runs the fused kernel of the next batch.  A rank whose total exceeds seg_cap still sends its true total, so
// every receiver can tell that the capacity has to be raised.
//
// NCCL is bound at run time (dlopen): a single-GPU user needs no NCCL, and inside a PyTorch process the library that is
// already loaded (torch's bundled libnccl.so.2) is the one used, so a communicator created by the host framework and one
// created here (hrmgpu_comm_init) come from the same NCCL.  Only the few stable entry points below are declared.
#include "internal.cuh"

#include <dlfcn.h>

#include <cstring>
#include <mutex>

namespace hrmgpu {

// ---- minimal NCCL ABI (nccl.h, unchanged since NCCL 2.0) ----
typedef struct ncclComm* ncclComm_t;
typedef struct {
    char internal[128];
} ncclUniqueId;
typedef int ncclResult_t;  // ncclSuccess = 0
typedef int ncclDataType_t;
constexpr ncclDataType_t kNcclInt8 = 0;

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*CommCount)(const ncclComm_t, int*) = nullptr;
    ncclResult_t (*CommUserRank)(const ncclComm_t, int*) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
};

static NcclApi* nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        // the NCCL already in the process (e.g. PyTorch's) first; then an explicit path; then the system library
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (!h)
            if (const char* e = std::getenv("HRMGPU_NCCL_LIB")) h = dlopen(e, RTLD_NOW);
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW);
        if (!h) {
            api.error = std::string("NCCL not found (libnccl.so.2): ") + (dlerror() ? dlerror() : "");
            return;
        }
        api.handle = h;
        bool ok = true;
        auto sym = [&](const char* name) {
            void* p = dlsym(h, name);
            if (!p) ok = false, api.error = std::string("NCCL symbol missing: ") + name;
            return p;
        };
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.CommCount = reinterpret_cast<decltype(api.CommCount)>(sym("ncclCommCount"));
        api.CommUserRank = reinterpret_cast<decltype(api.CommUserRank)>(sym("ncclCommUserRank"));
        api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
        if (!ok) api.handle = nullptr;
    });
    return &api;
}

#define HRMGPU_NCCL(ctx, api, expr)                                                                     \
    do {                                                                                                \
        ncclResult_t _r = (expr);                                                                       \
        if (_r != 0) return fail(ctx, HRMGPU_E_CUDA, std::string(#expr) + ": " + (api)->GetErrorString(_r)); \
    } while (0)

// Block layout in 8-byte words: [0] lines (K*L of the sender), [1] total segments of the sender, [2 .. 2+line_cap] CSR offsets
// (entries past `lines` repeat the total), then xL[seg_cap], xU[seg_cap], xM[seg_cap].
__global__ void __launch_bounds__(256) k_pack_block(const long long* off, const double* seg, long long seg_stride, long long nl, long long line_cap,
                                                    long long seg_cap, long long* block) {
    const long long total = off[nl];
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    const long long t0 = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t0 == 0) block[0] = nl, block[1] = total;
    for (long long i = t0; i <= line_cap; i += stride) block[2 + i] = (i <= nl) ? off[i] : total;
    double* pay = reinterpret_cast<double*>(block + 2 + line_cap + 1);
    const long long cnt = total < seg_cap ? total : seg_cap;
    for (long long i = t0; i < 3 * cnt; i += stride) {
        const long long a = i / cnt, j = i - a * cnt;
        pay[a * seg_cap + j] = seg[a * seg_stride + j];
    }
}

void preload_exchange() {
    cudaFuncAttributes a;
    if (cudaFuncGetAttributes(&a, k_pack_block) != cudaSuccess) cudaGetLastError();
}

static long long block_bytes(long long line_cap, long long seg_cap) { return 8 * (2 + line_cap + 1) + 24 * seg_cap; }

}  // namespace hrmgpu

using namespace hrmgpu;

extern "C" {

int hrmgpu_comm_unique_id(void* id128) {
    if (!id128) return HRMGPU_E_ARG;
    NcclApi* api = nccl_api();
    if (!api->handle) return HRMGPU_E_STATE;
    ncclUniqueId id;
    if (api->GetUniqueId(&id) != 0) return HRMGPU_E_CUDA;
    std::memcpy(id128, id.internal, 128);
    return HRMGPU_OK;
}

int hrmgpu_comm_init(hrmgpu_ctx* c, int n_ranks, int rank, const void* id128) {
    if (!c) return HRMGPU_E_ARG;
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks || !id128) return fail(c, HRMGPU_E_ARG, "comm_init: bad arguments");
    NcclApi* api = nccl_api();
    if (!api->handle) return fail(c, HRMGPU_E_STATE, api->error);
    HRMGPU_CUDA(c, cudaSetDevice(c->device));
    if (c->comm) {
        api->CommDestroy(static_cast<ncclComm_t>(c->comm));
        c->comm = nullptr;
    }
    ncclUniqueId id;
    std::memcpy(id.internal, id128, 128);
    ncclComm_t comm = nullptr;
    HRMGPU_NCCL(c, api, api->CommInitRank(&comm, n_ranks, id, rank));
    c->comm = comm;
    return HRMGPU_OK;
}

int hrmgpu_comm_destroy(hrmgpu_ctx* c) {
    if (!c) return HRMGPU_E_ARG;
    if (c->comm) {
        NcclApi* api = nccl_api();
        cudaSetDevice(c->device);
        cudaStreamSynchronize(c->s2);
        if (api->handle) api->CommDestroy(static_cast<ncclComm_t>(c->comm));
        c->comm = nullptr;
    }
    return HRMGPU_OK;
}

int hrmgpu_allgather_layers(hrmgpu_ctx* c, void* nccl_comm, void* cuda_stream, long long line_cap, long long seg_cap) {
    if (!c) return HRMGPU_E_ARG;
    if (!c->built) return fail(c, HRMGPU_E_STATE, "allgather_layers: no layers built yet");
    const long long nl = static_cast<long long>(c->K) * c->L;
    if (line_cap < nl || seg_cap < 0) return fail(c, HRMGPU_E_ARG, "allgather_layers: line_cap must cover this rank's K*L lines");
    NcclApi* api = nccl_api();
    if (!api->handle) return fail(c, HRMGPU_E_STATE, api->error);
    ncclComm_t comm = nccl_comm ? static_cast<ncclComm_t>(nccl_comm) : static_cast<ncclComm_t>(c->comm);
    if (!comm) return fail(c, HRMGPU_E_STATE, "allgather_layers: no communicator (hrmgpu_comm_init or pass an ncclComm_t)");
    HRMGPU_CUDA(c, cudaSetDevice(c->device));
    int world = 0, rank = 0;
    HRMGPU_NCCL(c, api, api->CommCount(comm, &world));
    HRMGPU_NCCL(c, api, api->CommUserRank(comm, &rank));
    const long long bytes = block_bytes(line_cap, seg_cap);
    // Everything below is queued on the side stream (or the caller's), behind the free-segment passes of the last build.
    cudaStream_t st = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : c->s2;
    int rc;
    if ((rc = ensure_on(c, c->s2, c->xchg_send, static_cast<size_t>(bytes)))) return rc;
    if ((rc = ensure_on(c, c->s2, c->xchg_recv, static_cast<size_t>(bytes) * world))) return rc;
    if (st != c->s2) HRMGPU_CUDA(c, cudaStreamWaitEvent(st, c->ev_k3[c->ivl_cur], 0));
    if (nl > 0) {
        const long long work = (line_cap + 1) > 3 * seg_cap ? (line_cap + 1) : 3 * seg_cap;
        const int grid = static_cast<int>(std::min<long long>((work + 255) / 256, 4LL * c->num_sms));
        k_pack_block<<<grid > 0 ? grid : 1, 256, 0, st>>>(c->seg_off.p, c->segdata.p, c->seg_cap, nl, line_cap, seg_cap,
                                                           reinterpret_cast<long long*>(c->xchg_send.p));
        HRMGPU_CUDA(c, cudaGetLastError());
        c->launches += 1;
    } else {
        HRMGPU_CUDA(c, cudaMemsetAsync(c->xchg_send.p, 0, static_cast<size_t>(bytes), st));
    }
    HRMGPU_NCCL(c, api, api->AllGather(c->xchg_send.p, c->xchg_recv.p, static_cast<size_t>(bytes), kNcclInt8, comm, st));
    if (st != c->s2) {  // later side-stream work (the next passes reuse xchg_send) must follow the caller's stream
        HRMGPU_CUDA(c, cudaEventRecord(c->ev_join, st));
        HRMGPU_CUDA(c, cudaStreamWaitEvent(c->s2, c->ev_join, 0));
    }
    c->xchg_world = world, c->xchg_rank = rank, c->xchg_line_cap = line_cap, c->xchg_seg_cap = seg_cap;
    c->xchg_stream = st;
    return HRMGPU_OK;
}

int hrmgpu_gathered_info(hrmgpu_ctx* c, long long* out4, void** d_blocks) {
    if (!c || !out4) return HRMGPU_E_ARG;
    if (c->xchg_world <= 0) return fail(c, HRMGPU_E_STATE, "no exchange done yet");
    out4[0] = c->xchg_world, out4[1] = c->xchg_rank, out4[2] = c->xchg_line_cap, out4[3] = c->xchg_seg_cap;
    if (d_blocks) *d_blocks = c->xchg_recv.p;
    return HRMGPU_OK;
}

long long hrmgpu_get_gathered(hrmgpu_ctx* c, int rank, long long* lines_out, long long* off, double* xL, double* xU, double* xM, long long cap) {
    if (!c) return HRMGPU_E_ARG;
    if (c->xchg_world <= 0) return fail(c, HRMGPU_E_STATE, "no exchange done yet");
    if (rank < 0 || rank >= c->xchg_world) return fail(c, HRMGPU_E_ARG, "get_gathered: rank out of range");
    HRMGPU_CUDA(c, cudaSetDevice(c->device));
    cudaStream_t st = static_cast<cudaStream_t>(c->xchg_stream);
    const long long bytes = block_bytes(c->xchg_line_cap, c->xchg_seg_cap);
    const char* blk = c->xchg_recv.p + static_cast<size_t>(bytes) * rank;
    long long head[2] = {0, 0};
    HRMGPU_CUDA(c, cudaMemcpyAsync(head, blk, 16, cudaMemcpyDeviceToHost, st));
    HRMGPU_CUDA(c, cudaStreamSynchronize(st));
    const long long nl = head[0], total = head[1];
    if (lines_out) *lines_out = nl;
    if (total > c->xchg_seg_cap) return fail(c, HRMGPU_E_CAP, "get_gathered: the sender's segments exceeded the exchange capacity; repeat with a larger seg_cap");
    if (off) HRMGPU_CUDA(c, cudaMemcpyAsync(off, blk + 16, 8 * (nl + 1), cudaMemcpyDeviceToHost, st));
    if (xL) {
        if (!xU || !xM) return fail(c, HRMGPU_E_ARG, "get_gathered: xU/xM are NULL");
        if (cap < total) return fail(c, HRMGPU_E_CAP, "get_gathered: buffer too small");
        const char* pay = blk + 8 * (2 + c->xchg_line_cap + 1);
        if (total) {
            HRMGPU_CUDA(c, cudaMemcpyAsync(xL, pay, 8 * total, cudaMemcpyDeviceToHost, st));
            HRMGPU_CUDA(c, cudaMemcpyAsync(xU, pay + 8 * c->xchg_seg_cap, 8 * total, cudaMemcpyDeviceToHost, st));
            HRMGPU_CUDA(c, cudaMemcpyAsync(xM, pay + 16 * c->xchg_seg_cap, 8 * total, cudaMemcpyDeviceToHost, st));
        }
    }
    HRMGPU_CUDA(c, cudaStreamSynchronize(st));
    return total;
}

}  // extern "C"
