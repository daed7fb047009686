// C ABI of libhrmgpu.so (see include/hrmgpu.h): context, scene/sweep setup, build calls, getters.
// There is deliberately no CPU fallback: without a CUDA device hrmgpu_create() fails with
// HRMGPU_E_NODEVICE and nothing else can be called.
#include "internal.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <mutex>
#include <new>

using namespace hrmgpu;

namespace {

// sgn(f(theta)) * |f(theta)|^p, sgn(0) = 0 — the superquadric power function
// (reference src/util/ExponentialFunction.cpp:5-10), evaluated on the host in FP64 with libm because it
// is needed only 8*n times per object per scene (separable in eta and omega).
double power_fn(double theta, double p, bool sine) {
    const double v = sine ? std::sin(theta) : std::cos(theta);
    const double sg = (v < 0) ? -1.0 : ((v > 0) ? 1.0 : 0.0);
    return sg * std::pow(std::fabs(v), p);
}

// Parameter grid of SuperQuadrics::SuperQuadrics (reference src/geometry/SuperQuadrics.cpp:18-24):
// Eigen LinSpaced(n, lo, hi) = lo + i*step, last element exactly hi.
void lin_grid(int n, double lo, double hi, std::vector<double>& v) {
    v.resize(static_cast<size_t>(n));
    const double step = (hi - lo) / static_cast<double>(n - 1);
    const bool flip = std::fabs(hi) < std::fabs(lo);
    for (int i = 0; i < n; ++i) {
        if (flip)
            v[i] = (i == 0) ? lo : (hi - static_cast<double>(n - 1 - i) * step);
        else
            v[i] = (i == n - 1) ? hi : (lo + static_cast<double>(i) * step);
    }
}

// reference include/hrm/datastructure/DataType.h:30-32 (truncated on purpose)
const double kPI = 3.1415926;
const double kHALF_PI = kPI / 2.0;
const double kEPS = 1e-6;

template <class T>
int upload(Ctx* c, DevBuf<T>& buf, const T* h, size_t n) {
    int rc = ensure(c, buf, n);
    if (rc) return rc;
    if (n) HRMGPU_CUDA(c, cudaMemcpyAsync(buf.p, h, n * sizeof(T), cudaMemcpyHostToDevice, c->stream));
    return HRMGPU_OK;
}

template <class T>
void release(Ctx* c, DevBuf<T>& b) {
    if (b.p) cudaFreeAsync(b.p, c->stream);
    b.p = nullptr;
    b.n = 0;
}

// Flips to the other interval-table buffer and sizes it.  The free-segment passes that read this buffer two builds ago run on
// the side stream: the main stream waits for them (normally long finished) before the fused kernel overwrites the tables.
int next_interval_tables(Ctx* c, size_t cnt) {
    c->ivl_cur ^= 1;
    const int cur = c->ivl_cur;
    HRMGPU_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_k3[cur], 0));
    int rc;
    if ((rc = ensure(c, c->ivl_lo[cur], cnt))) return rc;
    return ensure(c, c->ivl_up[cur], cnt);
}

// planar Real[S][dim][M] of one layer -> double[S][M][dim] (the reference's column-major dim x M per surface)
template <class Real>
__global__ void __launch_bounds__(256) k_interleave_boundary(const Real* __restrict__ planar, double* __restrict__ out, int dim, int M, long long total) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long surf = i / (static_cast<long long>(dim) * M);
        const int r = static_cast<int>(i - surf * dim * M);
        const int pt = r / dim, d = r - pt * dim;
        out[i] = static_cast<double>(planar[surf * dim * M + static_cast<long long>(d) * M + pt]);
    }
}

int check_built(Ctx* c, int k) {
    if (!c->built) return fail(c, HRMGPU_E_STATE, "no layers built yet");
    if (k < 0 || k >= c->K) return fail(c, HRMGPU_E_ARG, "layer index out of range");
    return HRMGPU_OK;
}

}  // namespace

namespace hrmgpu {
cudaMemPool_t library_pool(int device) {
    static std::mutex mu;
    static std::vector<cudaMemPool_t> pools;
    std::lock_guard<std::mutex> lock(mu);
    if (device < 0) return nullptr;
    if (pools.size() <= static_cast<size_t>(device)) pools.resize(static_cast<size_t>(device) + 1, nullptr);
    if (!pools[device]) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        cudaMemPool_t pool = nullptr;
        if (cudaMemPoolCreate(&pool, &props) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        unsigned long long thr = ~0ULL;  // keep freed blocks; hrmgpu_destroy trims what exceeds kPoolKeepBytes
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        pools[device] = pool;
    }
    return pools[device];
}
}  // namespace hrmgpu

extern "C" {

// The library holds ~40 kernel variants; with CUDA's default lazy module loading their first uses cost tens of milliseconds
// spread over the first planner calls (measured: HRM3D::plan() 29 / 30 / 17 ms before settling at 12.8 ms; 15 / 13 / 13 ms when
// loaded up front).  The first hrmgpu_create of the process therefore loads the library's OWN kernels explicitly
// (cudaFuncGetAttributes per kernel, preload_*); nothing process-wide (CUDA_MODULE_LOADING, the default memory pool) is changed.
int hrmgpu_abi_version(void) { return HRMGPU_ABI_VERSION; }

int hrmgpu_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int hrmgpu_create(hrmgpu_ctx** out, int device, unsigned flags) {
    (void)flags;
    if (!out) return HRMGPU_E_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return HRMGPU_E_NODEVICE;
    }
    if (device < 0 || device >= n) return HRMGPU_E_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return HRMGPU_E_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return HRMGPU_E_CUDA;
    if (prop.major < 10) return HRMGPU_E_NODEVICE;  // kernels are built for sm_100a only
    hrmgpu_ctx* c = new (std::nothrow) hrmgpu_ctx();
    if (!c) return HRMGPU_E_NOMEM;
    c->device = device;
    c->num_sms = prop.multiProcessorCount;
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);  // (numerically lower = higher priority)
    // the side stream gets the higher priority: its short kernels are dispatched as soon as the fused kernel frees CTA slots
    bool ok = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithPriority(&c->s2, cudaStreamNonBlocking, prio_hi) == cudaSuccess && cudaEventCreate(&c->ev0) == cudaSuccess &&
              cudaEventCreate(&c->ev1) == cudaSuccess && cudaEventCreate(&c->ev2) == cudaSuccess && cudaEventCreate(&c->ev3) == cudaSuccess &&
              cudaEventCreateWithFlags(&c->ev_fused, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&c->ev_k3[0], cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&c->ev_k3[1], cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&c->ev_fetch, cudaEventDisableTiming) == cudaSuccess &&
              cudaMallocHost(reinterpret_cast<void**>(&c->h_info), 4 * sizeof(unsigned long long)) == cudaSuccess;
    if (!ok) {
        cudaGetLastError();
        hrmgpu_destroy(c);
        return HRMGPU_E_CUDA;
    }
    c->h_info[0] = c->h_info[1] = c->h_info[2] = c->h_info[3] = 0;
    c->stream = c->own_stream;
    c->pool = library_pool(device);
    if (!c->pool) {
        hrmgpu_destroy(c);
        return HRMGPU_E_CUDA;
    }
    {
        static std::once_flag once;
        std::call_once(once, [] { preload_mink3d(), preload_mink2d(), preload_segments(), preload_queries(), preload_exchange(), preload_connect(), preload_tfe(); });
    }
    // test hooks of the segment pass (tests/test_gpu_layers3d.py): list capacity of the 4-line kernel / one-line kernel only
    if (const char* e = std::getenv("HRMGPU_SEG_FAST_CAP")) c->seg_fast_cap = std::atoi(e);
    if (const char* e = std::getenv("HRMGPU_SEG_ONE_LINE")) c->force_seg_fallback = std::atoi(e) != 0;
    if (const char* e = std::getenv("HRMGPU_SEG_INIT_CAP")) c->seg_init_cap = std::atoll(e);
    if (const char* e = std::getenv("HRMGPU_BUCKET_CAP")) c->bucket_cap_hook = std::atoi(e);
    if (const char* e = std::getenv("HRMGPU_CLASSIC")) c->force_classic = std::atoi(e) != 0;
    if (const char* e = std::getenv("HRMGPU_WS_VARIANT")) c->ws_variant = std::atoi(e);
    if (const char* e = std::getenv("HRMGPU_WS_GENERIC")) c->ws_generic = std::atoi(e) != 0;
    if (const char* e = std::getenv("HRMGPU_NO_BANDS")) c->no_bands = std::atoi(e) != 0;
    if (const char* e = std::getenv("HRMGPU_NO_FACE_NORMALS")) c->no_face_normals = std::atoi(e) != 0;
    // test hook of the fused kernel: capacity of a warp's kept-cell list (small values force the multi-round path)
    if (const char* e = std::getenv("HRMGPU_CELL_SEG_CAP")) c->cell_seg_cap = std::atoi(e);
    if (ensure(c, c->counters, 4) != HRMGPU_OK) {
        hrmgpu_destroy(c);
        return HRMGPU_E_NOMEM;
    }
    cudaMemsetAsync(c->counters.p, 0, 4 * sizeof(unsigned long long), c->stream);
    *out = c;
    return HRMGPU_OK;
}

int hrmgpu_reset(hrmgpu_ctx* c) {
    if (!c) return HRMGPU_E_ARG;
    HRMGPU_CUDA(c, cudaSetDevice(c->device));
    HRMGPU_CUDA(c, cudaStreamSynchronize(c->stream));
    HRMGPU_CUDA(c, cudaStreamSynchronize(c->s2));
    c->stream = c->own_stream;
    c->built = false, c->bridge_built = false, c->sweep_set = false;
    c->dim = 0, c->K = 0, c->P = 0;
    c->h_off_valid = false, c->seg_pending = false, c->fetch_pending = false;
    c->seg_total = 0, c->xchg_world = 0;
    c->launches = 0, c->seg_redos = 0;
    c->err.clear();
    return HRMGPU_OK;
}

void hrmgpu_destroy(hrmgpu_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->s2) cudaStreamSynchronize(c->s2);
    hrmgpu_comm_destroy(c);
    c->stream = c->own_stream;  // (an external stream may not outlive this call: free on the context's own stream)
    if (!c->stream || !c->pool) {  // creation failed half-way: nothing was allocated from the pool yet
        if (c->own_stream) cudaStreamDestroy(c->own_stream);
        if (c->s2) cudaStreamDestroy(c->s2);
        if (c->h_info) cudaFreeHost(c->h_info);
        for (cudaEvent_t e : {c->ev0, c->ev1, c->ev2, c->ev3, c->ev_fused, c->ev_k3[0], c->ev_k3[1], c->ev_join, c->ev_fetch})
            if (e) cudaEventDestroy(e);
        delete c;
        return;
    }
    release(c, c->obj3);
    release(c, c->obj2);
    release(c, c->tab);
    release(c, c->txs);
    release(c, c->tys);
    release(c, c->body_semi);
    release(c, c->body_R);
    release(c, c->body_off);
    release(c, c->mesh);
    release(c, c->ivl_lo[0]);
    release(c, c->ivl_lo[1]);
    release(c, c->ivl_up[0]);
    release(c, c->ivl_up[1]);
    release(c, c->bcnt[0]);
    release(c, c->bcnt[1]);
    release(c, c->bucket[0]);
    release(c, c->bucket[1]);
    release(c, c->orig);
    release(c, c->orig_off);
    release(c, c->orig_cnt);
    release(c, c->seg_cnt);
    release(c, c->seg_off);
    release(c, c->segdata);
    release(c, c->seg_info);
    release(c, c->xchg_send);
    release(c, c->xchg_recv);
    release(c, c->bound_tmp);
    release(c, c->counters);
    release(c, c->bridge_mesh);
    release(c, c->dbg);
    release(c, c->q_a);
    release(c, c->q_b);
    release(c, c->q_out);
    release(c, c->bbox);
    release(c, c->q_flag);
    release(c, c->q_set);
    release(c, c->q_w);
    release(c, c->q_loff);
    release(c, c->q_loff2);
    release(c, c->pc_layers), release(c, c->pc_cnt), release(c, c->pc_m1), release(c, c->pc_match);
    release(c, c->pc_m0_first), release(c, c->pc_first), release(c, c->pc_info), release(c, c->pc_vtx);
    release(c, c->tfe_in), release(c, c->tfe_quat), release(c, c->pc_code);
    release(c, c->bridge_bbox);
    release(c, c->bridge_bands);
    release(c, c->face_n);
    release(c, c->scan_state_main), release(c, c->scan_state_side);
    release(c, c->bridge_semi);
    release(c, c->bridge_R);
    release(c, c->bridge_off);
    if (c->own_stream) cudaStreamSynchronize(c->own_stream);
    {
        cudaMemPool_t pool = c->pool;
        if (pool) {
            // give back what exceeds (memory still in use by other contexts + kPoolKeepBytes of free blocks); trimming below the
            // used size would release every free block and send the next context's allocations to the driver again
            unsigned long long reserved = 0, used = 0;
            if (cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved) == cudaSuccess &&
                cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used) == cudaSuccess && reserved > used + kPoolKeepBytes)
                cudaMemPoolTrimTo(pool, static_cast<size_t>(used + kPoolKeepBytes));
        }
    }
    for (cudaEvent_t e : {c->ev0, c->ev1, c->ev2, c->ev3, c->ev_fused, c->ev_k3[0], c->ev_k3[1], c->ev_join, c->ev_fetch})
        if (e) cudaEventDestroy(e);
    if (c->h_info) cudaFreeHost(c->h_info);
    if (c->h_off) cudaFreeHost(c->h_off);
    if (c->s2) cudaStreamDestroy(c->s2);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

const char* hrmgpu_last_error(const hrmgpu_ctx* c) { return c ? c->err.c_str() : "null context"; }

int hrmgpu_set_stream(hrmgpu_ctx* c, void* s) {
    if (!c) return HRMGPU_E_ARG;
    HRMGPU_CUDA(c, cudaStreamSynchronize(c->stream));
    HRMGPU_CUDA(c, cudaStreamSynchronize(c->s2));
    c->stream = s ? static_cast<cudaStream_t>(s) : c->own_stream;
    return HRMGPU_OK;
}

int hrmgpu_synchronize(hrmgpu_ctx* c) {
    if (!c) return HRMGPU_E_ARG;
    HRMGPU_CUDA(c, cudaStreamSynchronize(c->stream));
    HRMGPU_CUDA(c, cudaStreamSynchronize(c->s2));
    return HRMGPU_OK;
}

int hrmgpu_join(hrmgpu_ctx* c) {
    if (!c) return HRMGPU_E_ARG;
    HRMGPU_CUDA(c, cudaEventRecord(c->ev_join, c->s2));
    HRMGPU_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_join, 0));
    return HRMGPU_OK;
}

int hrmgpu_set_scene3d(hrmgpu_ctx* c, int n_arena, int n_obs, const double* sq, int n) {
    if (!c) return HRMGPU_E_ARG;
    if (n_arena < 0 || n_obs < 0 || n_arena + n_obs <= 0 || !sq) return fail(c, HRMGPU_E_ARG, "set_scene3d: bad object counts");
    if (n < 3 || n > kMaxGrid) return fail(c, HRMGPU_E_CAP, "set_scene3d: n_grid must be in [3, 200]");
    HRMGPU_CUDA(c, cudaSetDevice(c->device));
    const int S0 = n_arena + n_obs;
    std::vector<Obj3> objs(static_cast<size_t>(S0));
    std::vector<double> tab(static_cast<size_t>(S0) * 8 * n);
    std::vector<double> eta, omega;
    lin_grid(n, -kHALF_PI - kEPS, kHALF_PI + kEPS, eta);
    lin_grid(n, -kPI - kEPS, kPI + kEPS, omega);
    for (int i = 0; i < S0; ++i) {
        const double* r = sq + static_cast<size_t>(i) * 17;
        Obj3& o = objs[i];
        o.a = r[0], o.b = r[1], o.c = r[2];
        const double e1 = r[3], e2 = r[4];
        o.px = r[5], o.py = r[6], o.pz = r[7];
        for (int j = 0; j < 9; ++j) o.R[j] = r[8 + j];
        o.sign = (i < n_arena) ? -1.0 : 1.0;
        o.ia = 1.0 / o.a, o.ib = 1.0 / o.b, o.ic = 1.0 / o.c;
        double* t = tab.data() + static_cast<size_t>(i) * 8 * n;
        for (int j = 0; j < n; ++j) {
            t[0 * n + j] = power_fn(eta[j], e1, false);
            t[1 * n + j] = power_fn(eta[j], e1, true);
            t[2 * n + j] = power_fn(eta[j], 2 - e1, false);
            t[3 * n + j] = power_fn(eta[j], 2 - e1, true);
            t[4 * n + j] = power_fn(omega[j], e2, false);
            t[5 * n + j] = power_fn(omega[j], e2, true);
            t[6 * n + j] = power_fn(omega[j], 2 - e2, false);
            t[7 * n + j] = power_fn(omega[j], 2 - e2, true);
        }
    }
    int rc;
    if ((rc = upload(c, c->obj3, objs.data(), objs.size()))) return rc;
    if ((rc = upload(c, c->tab, tab.data(), tab.size()))) return rc;
    HRMGPU_CUDA(c, cudaStreamSynchronize(c->stream));  // host vectors go out of scope
    if (c->dim != 3) c->sweep_set = false;  // a 2D sweep grid (Lx forced to 1) is not a valid 3D grid
    c->dim = 3;
    c->n_arena = n_arena;
    c->n_obs = n_obs;
    c->n = n;
    c->built = false;
    c->bridge_built = false;
    return HRMGPU_OK;
}

// Sweep-line grids of HRM3D::sweepLineProcess / HRM2D::sweepLineProcess for the current scene (shared by set_sweep and resweep)
static int set_grid(hrmgpu_ctx* c, const double* lim, int Lx, int Ly, const char* who) {
    if (!lim) return fail(c, HRMGPU_E_ARG, "sweep grid: lim is NULL");
    if (c->dim == 0) return fail(c, HRMGPU_E_STATE, "sweep grid: set a scene first");
    HRMGPU_CUDA(c, cudaSetDevice(c->device));
    if (c->dim == 3) {
        if (Lx < 2 || Ly < 2) return fail(c, HRMGPU_E_ARG, "sweep grid: 3D needs Lx >= 2 and Ly >= 2");
    } else {
        Lx = 1;
        if (Ly < 1) return fail(c, HRMGPU_E_ARG, "sweep grid: Ly must be >= 1");
    }
    if (Lx > kMaxLinesAxis || Ly > kMaxLinesAxis || static_cast<long long>(Lx) * Ly > kMaxLines)
        return fail(c, HRMGPU_E_CAP, "sweep grid: at most 16383 lines per axis and 8192 lines per layer");
    (void)who;
    const int nl = (c->dim == 3) ? 6 : 4;
    for (int i = 0; i < nl; ++i) c->lim[i] = lim[i];
    if (c->dim == 2) c->lim[4] = c->lim[5] = 0.0;
    c->Lx = Lx;
    c->Ly = Ly;
    const double inf = std::numeric_limits<double>::infinity();
    std::vector<double> tx(static_cast<size_t>(Lx) + 2), ty(static_cast<size_t>(Ly) + 2);
    // reference src/planners/HRM3D.cpp:66-81 (3D) / HRM2D.cpp:56-61 (2D): t_i = lo + double(i) * d
    const double dx = (lim[1] - lim[0]) / static_cast<double>(Lx - 1);
    const double dy = (lim[3] - lim[2]) / (static_cast<double>(Ly) - 1);
    tx[0] = -inf;
    tx[Lx + 1] = inf;
    for (int i = 0; i < Lx; ++i) tx[i + 1] = (c->dim == 3) ? lim[0] + static_cast<double>(i) * dx : 0.0;
    ty[0] = -inf;
    ty[Ly + 1] = inf;
    for (int i = 0; i < Ly; ++i) ty[i + 1] = lim[2] + static_cast<double>(i) * dy;
    int rc;
    if ((rc = upload(c, c->txs, tx.data(), tx.size()))) return rc;
    if ((rc = upload(c, c->tys, ty.data(), ty.size()))) return rc;
    HRMGPU_CUDA(c, cudaStreamSynchronize(c->stream));
    c->sweep_set = true;
    return HRMGPU_OK;
}

int hrmgpu_set_sweep(hrmgpu_ctx* c, const double* lim, int Lx, int Ly) {
    if (!c) return HRMGPU_E_ARG;
    const int rc = set_grid(c, lim, Lx, Ly, "set_sweep");
    if (rc == HRMGPU_OK) c->built = false;
    return rc;
}

// Refinement path (HighwayRoadMap-inl.h:178-215, HRM3D.cpp:24-54 with isRefine_): the C-boundaries of the last build stay in
// HBM (the reference caches them in sliceBoundAll_ / sliceBoundMeshAll_) and only the sweep (K2) and the free segments (K3)
// are redone for a new grid.  Interval tables are sized by the new line count (the reference's are not: SURVEY finding 6).
int hrmgpu_resweep(hrmgpu_ctx* c, const double* lim, int Lx, int Ly) {
    if (!c) return HRMGPU_E_ARG;
    if (!c->built) return fail(c, HRMGPU_E_STATE, "resweep: no layers built");
    int rc = set_grid(c, lim, Lx, Ly, "resweep");
    if (rc) {
        return rc;
    }
    c->built = false;
    const int S0 = c->n_arena + c->n_obs;
    c->L = c->Lx * c->Ly;
    if ((rc = next_interval_tables(c, static_cast<size_t>(c->K) * c->S * c->L))) return rc;
    double *lo = c->ivl_lo[c->ivl_cur].p, *up = c->ivl_up[c->ivl_cur].p;
    HRMGPU_CUDA(c, cudaMemsetAsync(c->counters.p, 0, 4 * sizeof(unsigned long long), c->stream));
    HRMGPU_CUDA(c, cudaEventRecord(c->ev0, c->stream));
    if (c->dim == 3)
        rc = launch_minksweep3d(c, c->K, c->B, 0, S0, nullptr, nullptr, nullptr, c->prec, c->mesh.p, lo, up, true, true);
    else
        rc = launch_minksweep2d(c, c->K, c->B, 0, S0, c->body_semi.p, c->body_R.p, c->body_off.p, c->prec, c->mesh.p, lo, up, true, true);
    if (rc) return rc;
    HRMGPU_CUDA(c, cudaEventRecord(c->ev1, c->stream));
    if ((rc = launch_segments(c))) return rc;
    c->built = true;
    return HRMGPU_OK;
}

int hrmgpu_build_layers_dev(hrmgpu_ctx* c, int K, int B, const double* d_semi, const double* d_R, const double* d_off,
                            int precision) {
    if (!c) return HRMGPU_E_ARG;
    if (c->dim != 3) return fail(c, HRMGPU_E_STATE, "build_layers: no 3D scene set");
    if (!c->sweep_set || c->Lx < 2 || c->Ly < 2) return fail(c, HRMGPU_E_STATE, "build_layers: set_sweep first");
    if (K < 0 || B < 1 || !d_semi || (K > 0 && (!d_R || !d_off))) return fail(c, HRMGPU_E_ARG, "build_layers: bad arguments");
    if (precision < HRMGPU_F64_STRICT || precision > HRMGPU_F32) return fail(c, HRMGPU_E_ARG, "build_layers: unknown precision");
    HRMGPU_CUDA(c, cudaSetDevice(c->device));
    const int S0 = c->n_arena + c->n_obs;
    const int S = S0 * B, Sa = c->n_arena * B, L = c->Lx * c->Ly, M = c->n * c->n;
    const size_t real = (precision == HRMGPU_F32) ? 4 : 8;
    int rc;
    if ((rc = ensure(c, c->mesh, static_cast<size_t>(K) * S * 3 * M * real))) return rc;
    if ((rc = next_interval_tables(c, static_cast<size_t>(K) * S * L))) return rc;
    HRMGPU_CUDA(c, cudaMemsetAsync(c->counters.p, 0, 4 * sizeof(unsigned long long), c->stream));
    c->K = K, c->B = B, c->S = S, c->Sa = Sa, c->L = L, c->M = M, c->prec = precision;
    c->built = false;
    c->last_semi = d_semi, c->last_R = d_R, c->last_off = d_off;
    HRMGPU_CUDA(c, cudaEventRecord(c->ev0, c->stream));
    if ((rc = launch_minksweep3d(c, K, B, 0, S0, d_semi, d_R, d_off, precision, c->mesh.p, c->ivl_lo[c->ivl_cur].p, c->ivl_up[c->ivl_cur].p, true)))
        return rc;
    HRMGPU_CUDA(c, cudaEventRecord(c->ev1, c->stream));
    if ((rc = launch_segments(c))) return rc;
    c->built = true;
    return HRMGPU_OK;
}

int hrmgpu_build_layers(hrmgpu_ctx* c, int K, int B, const double* semi, const double* R, const double* off, int precision) {
    if (!c) return HRMGPU_E_ARG;
    if (K < 0 || B < 1 || !semi || (K > 0 && (!R || !off))) return fail(c, HRMGPU_E_ARG, "build_layers: bad arguments");
    HRMGPU_CUDA(c, cudaSetDevice(c->device));
    int rc;
    if ((rc = upload(c, c->body_semi, semi, static_cast<size_t>(B) * 3))) return rc;
    if ((rc = upload(c, c->body_R, R, static_cast<size_t>(K) * B * 9))) return rc;
    if ((rc = upload(c, c->body_off, off, static_cast<size_t>(K) * B * 3))) return rc;
    return hrmgpu_build_layers_dev(c, K, B, c->body_semi.p, c->body_R.p, c->body_off.p, precision);
}

int hrmgpu_set_scene2d(hrmgpu_ctx* c, int n_arena, int n_obs, const double* se, int num) {
    if (!c) return HRMGPU_E_ARG;
    if (n_arena < 0 || n_obs < 0 || n_arena + n_obs <= 0 || !se) return fail(c, HRMGPU_E_ARG, "set_scene2d: bad object counts");
    if (num < 3 || num > 8192) return fail(c, HRMGPU_E_CAP, "set_scene2d: num must be in [3, 8192]");
    HRMGPU_CUDA(c, cudaSetDevice(c->device));
    const int S0 = n_arena + n_obs;
    std::vector<Obj2> objs(static_cast<size_t>(S0));
    std::vector<double> tab(static_cast<size_t>(S0) * 4 * num);
    for (int i = 0; i < S0; ++i) {
        const double* r = se + static_cast<size_t>(i) * 6;
        Obj2& o = objs[i];
        o.a = r[0], o.b = r[1], o.eps = r[2], o.px = r[3], o.py = r[4];
        o.c1 = std::cos(r[5]);  // Eigen::Rotation2Dd(angle).matrix()
        o.s1 = std::sin(r[5]);
        o.sign = (i < n_arena) ? -1.0 : 1.0;
        double* t = tab.data() + static_cast<size_t>(i) * 4 * num;
        for (int j = 0; j < num; ++j) {
            const double th = 2 * j * kPI / (static_cast<double>(num) - 1);  // reference src/geometry/SuperEllipse.cpp:40,89
            t[0 * num + j] = power_fn(th, o.eps, false);
            t[1 * num + j] = power_fn(th, o.eps, true);
            t[2 * num + j] = power_fn(th, 2 - o.eps, false);
            t[3 * num + j] = power_fn(th, 2 - o.eps, true);
        }
    }
    int rc;
    if ((rc = upload(c, c->obj2, objs.data(), objs.size()))) return rc;
    if ((rc = upload(c, c->tab, tab.data(), tab.size()))) return rc;
    HRMGPU_CUDA(c, cudaStreamSynchronize(c->stream));
    c->dim = 2;
    c->n_arena = n_arena;
    c->n_obs = n_obs;
    c->n = num;
    c->built = false;
    c->bridge_built = false;
    c->sweep_set = false;
    return HRMGPU_OK;
}

int hrmgpu_build_layers2d(hrmgpu_ctx* c, int K, int B, const double* semi, const double* angle, const double* off, int precision) {
    if (!c) return HRMGPU_E_ARG;
    if (c->dim != 2) return fail(c, HRMGPU_E_STATE, "build_layers2d: no 2D scene set");
    if (!c->sweep_set) return fail(c, HRMGPU_E_STATE, "build_layers2d: set_sweep first");
    if (K < 0 || B < 1 || !semi || (K > 0 && (!angle || !off))) return fail(c, HRMGPU_E_ARG, "build_layers2d: bad arguments");
    if (precision < HRMGPU_F64_STRICT || precision > HRMGPU_F32) return fail(c, HRMGPU_E_ARG, "build_layers2d: unknown precision");
    if (precision == HRMGPU_F32) precision = HRMGPU_F64;  // the 2D path has no FP32 variant (it is tiny); FP64 fast math instead
    HRMGPU_CUDA(c, cudaSetDevice(c->device));
    std::vector<double> cs(static_cast<size_t>(K) * B * 2);
    for (size_t i = 0; i < static_cast<size_t>(K) * B; ++i) cs[2 * i] = std::cos(angle[i]), cs[2 * i + 1] = std::sin(angle[i]);
    int rc;
    if ((rc = upload(c, c->body_semi, semi, static_cast<size_t>(B) * 2))) return rc;
    if ((rc = upload(c, c->body_R, cs.data(), cs.size()))) return rc;
    if ((rc = upload(c, c->body_off, off, static_cast<size_t>(K) * B * 2))) return rc;
    const int S0 = c->n_arena + c->n_obs;
    const int S = S0 * B, Sa = c->n_arena * B, L = c->Ly, M = c->n;
    if ((rc = ensure(c, c->mesh, static_cast<size_t>(K) * S * 2 * M * 8))) return rc;
    if ((rc = next_interval_tables(c, static_cast<size_t>(K) * S * L))) return rc;
    HRMGPU_CUDA(c, cudaMemsetAsync(c->counters.p, 0, 4 * sizeof(unsigned long long), c->stream));
    c->K = K, c->B = B, c->S = S, c->Sa = Sa, c->L = L, c->M = M, c->prec = precision;
    c->built = false;
    HRMGPU_CUDA(c, cudaEventRecord(c->ev0, c->stream));
    if ((rc = launch_minksweep2d(c, K, B, 0, S0, c->body_semi.p, c->body_R.p, c->body_off.p, precision, c->mesh.p, c->ivl_lo[c->ivl_cur].p,
                                 c->ivl_up[c->ivl_cur].p, true)))
        return rc;
    HRMGPU_CUDA(c, cudaEventRecord(c->ev1, c->stream));
    HRMGPU_CUDA(c, cudaStreamSynchronize(c->stream));  // cs goes out of scope
    if ((rc = launch_segments(c))) return rc;
    c->built = true;
    return HRMGPU_OK;
}

int hrmgpu_test_edges(hrmgpu_ctx* c, int k, int E, const double* v1, const double* v2, unsigned char* free_out) {
    if (!c) return HRMGPU_E_ARG;
    int rc = check_built(c, k);
    if (rc) return rc;
    if (E < 0 || (E > 0 && (!v1 || !v2 || !free_out))) return fail(c, HRMGPU_E_ARG, "test_edges: bad arguments");
    if (E == 0) return HRMGPU_OK;
    HRMGPU_CUDA(c, cudaSetDevice(c->device));
    const size_t cnt = static_cast<size_t>(E) * c->dim;
    if ((rc = upload(c, c->q_a, v1, cnt))) return rc;
    if ((rc = upload(c, c->q_b, v2, cnt))) return rc;
    if ((rc = ensure(c, c->q_out, static_cast<size_t>(E)))) return rc;
    if ((rc = launch_edges(c, k, E, c->q_a.p, c->q_b.p, c->q_out.p))) return rc;
    HRMGPU_CUDA(c, cudaMemcpyAsync(free_out, c->q_out.p, static_cast<size_t>(E), cudaMemcpyDeviceToHost, c->stream));
    HRMGPU_CUDA(c, cudaStreamSynchronize(c->stream));
    return HRMGPU_OK;
}

int hrmgpu_test_edges_multi(hrmgpu_ctx* c, int E, const int* layer, const double* v1, const double* v2, unsigned char* free_out) {
    if (!c) return HRMGPU_E_ARG;
    int rc = check_built(c, 0);
    if (rc) return rc;
    if (E < 0 || (E > 0 && (!layer || !v1 || !v2 || !free_out))) return fail(c, HRMGPU_E_ARG, "test_edges_multi: bad arguments");
    if (E == 0) return HRMGPU_OK;
    for (int e = 0; e < E; ++e)
        if (layer[e] < 0 || layer[e] >= c->K) return fail(c, HRMGPU_E_ARG, "test_edges_multi: layer index out of range");
    HRMGPU_CUDA(c, cudaSetDevice(c->device));
    const size_t cnt = static_cast<size_t>(E) * c->dim;
    if ((rc = upload(c, c->q_a, v1, cnt))) return rc;
    if ((rc = upload(c, c->q_b, v2, cnt))) return rc;
    if ((rc = upload(c, c->q_set, layer, static_cast<size_t>(E)))) return rc;
    if ((rc = ensure(c, c->q_out, static_cast<size_t>(E)))) return rc;
    if ((rc = launch_edges(c, -1, E, c->q_a.p, c->q_b.p, c->q_out.p, c->q_set.p))) return rc;
    HRMGPU_CUDA(c, cudaMemcpyAsync(free_out, c->q_out.p, static_cast<size_t>(E), cudaMemcpyDeviceToHost, c->stream));
    HRMGPU_CUDA(c, cudaStreamSynchronize(c->stream));
    return HRMGPU_OK;
}

// The Minkowski sums of every obstacle with the P x B ellipsoids in c->bridge_semi / c->bridge_R (already on the device).
static int build_bridge_meshes(hrmgpu_ctx* c, int P, int B, int precision, int dim) {
    const int M = (dim == 3) ? c->n * c->n : c->n;
    const size_t real = (precision == HRMGPU_F32) ? 4 : 8;
    int rc;
    if ((rc = ensure(c, c->bridge_mesh, static_cast<size_t>(P) * c->n_obs * B * dim * M * real))) return rc;
    if (dim == 3) {
        // ONE launch for all pairs: layer = pair, with the pair's own body semi-axes (semi_per_layer)
        if ((rc = ensure(c, c->bridge_off, static_cast<size_t>(P > 0 ? P : 1) * B * 3))) return rc;
        HRMGPU_CUDA(c, cudaMemsetAsync(c->bridge_off.p, 0, sizeof(double) * static_cast<size_t>(P > 0 ? P : 1) * B * 3, c->stream));
        if (P > 0) {
            c->semi_per_layer = true;
            rc = launch_minksweep3d(c, P, B, c->n_arena, c->n_obs, c->bridge_semi.p, c->bridge_R.p, c->bridge_off.p, precision, c->bridge_mesh.p, nullptr, nullptr, false);
            c->semi_per_layer = false;
            if (rc) return rc;
        }
    } else {
        std::vector<double> zeros(static_cast<size_t>(B) * dim, 0.0);
        if ((rc = upload(c, c->bridge_off, zeros.data(), zeros.size()))) return rc;
        for (int pi = 0; pi < P; ++pi) {  // every pair has its own body shapes: one layer-batch per pair with that pair's semi-axes
            char* mesh = c->bridge_mesh.p + static_cast<size_t>(pi) * c->n_obs * B * dim * M * real;
            rc = launch_minksweep2d(c, 1, B, c->n_arena, c->n_obs, c->bridge_semi.p + static_cast<size_t>(pi) * B * 2,
                                    c->bridge_R.p + static_cast<size_t>(pi) * B * 2, c->bridge_off.p, precision, mesh, nullptr, nullptr, false);
            if (rc) return rc;
        }
    }
    c->P = P;
    c->bridge_B = B;
    c->bridge_prec = precision;
    if ((rc = launch_bridge_bbox(c))) return rc;
    HRMGPU_CUDA(c, cudaStreamSynchronize(c->stream));
    c->bridge_built = true;
    return HRMGPU_OK;
}

static int build_bridge_common(hrmgpu_ctx* c, int P, int B, const double* semi, const double* rot, int precision, int dim) {
    if (!c) return HRMGPU_E_ARG;
    if (c->dim != dim) return fail(c, HRMGPU_E_STATE, "build_bridge: scene dimension mismatch");
    if (P < 0 || B < 1 || (P > 0 && (!semi || !rot))) return fail(c, HRMGPU_E_ARG, "build_bridge: bad arguments");
    if (precision < HRMGPU_F64_STRICT || precision > HRMGPU_F32) return fail(c, HRMGPU_E_ARG, "build_bridge: unknown precision");
    if (dim == 2 && precision == HRMGPU_F32) precision = HRMGPU_F64;
    HRMGPU_CUDA(c, cudaSetDevice(c->device));
    int rc;
    if ((rc = upload(c, c->bridge_semi, semi, static_cast<size_t>(P) * B * dim))) return rc;
    std::vector<double> cs;
    if (dim == 3) {
        if ((rc = upload(c, c->bridge_R, rot, static_cast<size_t>(P) * B * 9))) return rc;
    } else {
        cs.resize(static_cast<size_t>(P) * B * 2);
        for (size_t i = 0; i < static_cast<size_t>(P) * B; ++i) cs[2 * i] = std::cos(rot[i]), cs[2 * i + 1] = std::sin(rot[i]);
        if ((rc = upload(c, c->bridge_R, cs.data(), cs.size()))) return rc;
    }
    return build_bridge_meshes(c, P, B, precision, dim);  // (synchronises: cs / the caller's arrays are pageable host memory)
}

int hrmgpu_build_bridge_tfe(hrmgpu_ctx* c, int P, int B, const double* body_semi, const double* quat_a, const double* quat_b, int num_step, int precision,
                            double* tfe_out) {
    if (!c) return HRMGPU_E_ARG;
    if (c->dim != 3) return fail(c, HRMGPU_E_STATE, "build_bridge_tfe: 3D scenes only");
    if (P < 0 || B < 1 || num_step < 1 || (P > 0 && (!body_semi || !quat_a || !quat_b))) return fail(c, HRMGPU_E_ARG, "build_bridge_tfe: bad arguments");
    if (precision < HRMGPU_F64_STRICT || precision > HRMGPU_F32) return fail(c, HRMGPU_E_ARG, "build_bridge_tfe: unknown precision");
    HRMGPU_CUDA(c, cudaSetDevice(c->device));
    int rc;
    const size_t PB = static_cast<size_t>(P) * B;
    std::vector<double> in(3 * static_cast<size_t>(B) + 8 * PB);
    std::copy(body_semi, body_semi + 3 * static_cast<size_t>(B), in.begin());
    if (PB) {
        std::copy(quat_a, quat_a + 4 * PB, in.begin() + 3 * static_cast<size_t>(B));
        std::copy(quat_b, quat_b + 4 * PB, in.begin() + 3 * static_cast<size_t>(B) + 4 * PB);
    }
    if ((rc = upload(c, c->tfe_in, in.data(), in.size()))) return rc;
    if ((rc = ensure(c, c->bridge_semi, 3 * PB))) return rc;
    if ((rc = ensure(c, c->bridge_R, 9 * PB))) return rc;
    if ((rc = ensure(c, c->tfe_quat, 4 * PB))) return rc;
    if ((rc = launch_tfe3d(c, P, B, c->tfe_in.p, c->tfe_in.p + 3 * static_cast<size_t>(B), c->tfe_in.p + 3 * static_cast<size_t>(B) + 4 * PB, num_step,
                           c->bridge_semi.p, c->bridge_R.p, c->tfe_quat.p)))
        return rc;
    if (tfe_out && PB) {  // [P][B][7]: semi-axes, quaternion (w, x, y, z)
        std::vector<double> s(3 * PB), q(4 * PB);
        HRMGPU_CUDA(c, cudaMemcpyAsync(s.data(), c->bridge_semi.p, sizeof(double) * 3 * PB, cudaMemcpyDeviceToHost, c->stream));
        HRMGPU_CUDA(c, cudaMemcpyAsync(q.data(), c->tfe_quat.p, sizeof(double) * 4 * PB, cudaMemcpyDeviceToHost, c->stream));
        HRMGPU_CUDA(c, cudaStreamSynchronize(c->stream));
        for (size_t i = 0; i < PB; ++i) {
            for (int d = 0; d < 3; ++d) tfe_out[7 * i + d] = s[3 * i + d];
            for (int d = 0; d < 4; ++d) tfe_out[7 * i + 3 + d] = q[4 * i + d];
        }
    }
    return build_bridge_meshes(c, P, B, precision, 3);  // (synchronises: `in` is pageable host memory)
}

int hrmgpu_build_bridge(hrmgpu_ctx* c, int P, int B, const double* tfe_semi, const double* tfe_R, int precision) {
    return build_bridge_common(c, P, B, tfe_semi, tfe_R, precision, 3);
}
int hrmgpu_build_bridge2d(hrmgpu_ctx* c, int P, int B, const double* tfe_semi, const double* tfe_angle, int precision) {
    return build_bridge_common(c, P, B, tfe_semi, tfe_angle, precision, 2);
}

int hrmgpu_test_bridge(hrmgpu_ctx* c, int p, int Q, const double* centres, unsigned char* free_out) {
    if (!c) return HRMGPU_E_ARG;
    if (!c->bridge_built) return fail(c, HRMGPU_E_STATE, "test_bridge: no bridge layers built");
    if (p < 0 || p >= c->P) return fail(c, HRMGPU_E_ARG, "test_bridge: pair index out of range");
    if (Q < 0 || (Q > 0 && (!centres || !free_out))) return fail(c, HRMGPU_E_ARG, "test_bridge: bad arguments");
    if (Q == 0) return HRMGPU_OK;
    HRMGPU_CUDA(c, cudaSetDevice(c->device));
    int rc;
    if ((rc = upload(c, c->q_a, centres, static_cast<size_t>(Q) * c->bridge_B * c->dim))) return rc;
    if ((rc = ensure(c, c->q_out, static_cast<size_t>(Q)))) return rc;
    if ((rc = launch_bridge_test(c, p, Q, c->q_a.p, c->q_out.p))) return rc;
    HRMGPU_CUDA(c, cudaMemcpyAsync(free_out, c->q_out.p, static_cast<size_t>(Q), cudaMemcpyDeviceToHost, c->stream));
    HRMGPU_CUDA(c, cudaStreamSynchronize(c->stream));
    return HRMGPU_OK;
}

int hrmgpu_test_bridge_multi(hrmgpu_ctx* c, int Q, const int* pair, const double* centres, unsigned char* free_out) {
    if (!c) return HRMGPU_E_ARG;
    if (!c->bridge_built) return fail(c, HRMGPU_E_STATE, "test_bridge_multi: no bridge layers built");
    if (Q < 0 || (Q > 0 && (!pair || !centres || !free_out))) return fail(c, HRMGPU_E_ARG, "test_bridge_multi: bad arguments");
    if (Q == 0) return HRMGPU_OK;
    for (int q = 0; q < Q; ++q)
        if (pair[q] < 0 || pair[q] >= c->P) return fail(c, HRMGPU_E_ARG, "test_bridge_multi: pair index out of range");
    HRMGPU_CUDA(c, cudaSetDevice(c->device));
    int rc;
    if ((rc = upload(c, c->q_a, centres, static_cast<size_t>(Q) * c->bridge_B * c->dim))) return rc;
    if ((rc = upload(c, c->q_set, pair, static_cast<size_t>(Q)))) return rc;
    if ((rc = ensure(c, c->q_out, static_cast<size_t>(Q)))) return rc;
    if ((rc = launch_bridge_test(c, -1, Q, c->q_a.p, c->q_out.p, c->q_set.p))) return rc;
    HRMGPU_CUDA(c, cudaMemcpyAsync(free_out, c->q_out.p, static_cast<size_t>(Q), cudaMemcpyDeviceToHost, c->stream));
    HRMGPU_CUDA(c, cudaStreamSynchronize(c->stream));
    return HRMGPU_OK;
}

static int test_transitions_common(hrmgpu_ctx* c, int C, const int* pair, const double* v0, const double* v1, int T, const double* w0,
                                   const double* w1, const double* link_off, const double* link_off2, unsigned char* free_out) {
    if (!c) return HRMGPU_E_ARG;
    if (!c->bridge_built) return fail(c, HRMGPU_E_STATE, "test_transitions: no bridge layers built");
    if (C < 0 || T < 1 || (C > 0 && (!pair || !v0 || !v1 || !w0 || !w1 || !link_off || !free_out)))
        return fail(c, HRMGPU_E_ARG, "test_transitions: bad arguments");
    if (link_off2 && c->dim != 3) return fail(c, HRMGPU_E_ARG, "test_transitions_articulated: 3D scenes only");
    if (C == 0) return HRMGPU_OK;
    for (int i = 0; i < C; ++i)
        if (pair[i] < 0 || pair[i] >= c->P) return fail(c, HRMGPU_E_ARG, "test_transitions: pair index out of range");
    HRMGPU_CUDA(c, cudaSetDevice(c->device));
    int rc;
    const size_t cnt = static_cast<size_t>(C) * c->dim;
    if ((rc = upload(c, c->q_a, v0, cnt))) return rc;
    if ((rc = upload(c, c->q_b, v1, cnt))) return rc;
    if ((rc = upload(c, c->q_set, pair, static_cast<size_t>(C)))) return rc;
    std::vector<double> w(2 * static_cast<size_t>(T));
    for (int s = 0; s < T; ++s) w[s] = w0[s], w[T + s] = w1[s];
    if ((rc = upload(c, c->q_w, w.data(), w.size()))) return rc;
    const size_t nl = static_cast<size_t>(c->P) * T * c->bridge_B * c->dim;
    if ((rc = upload(c, c->q_loff, link_off, nl))) return rc;
    if (link_off2 && (rc = upload(c, c->q_loff2, link_off2, nl))) return rc;
    if ((rc = ensure(c, c->q_out, static_cast<size_t>(C)))) return rc;
    if ((rc = launch_transitions(c, C, T, c->q_set.p, c->q_a.p, c->q_b.p, c->q_w.p, c->q_w.p + T, c->q_loff.p, c->q_out.p,
                                 link_off2 ? c->q_loff2.p : nullptr)))
        return rc;
    HRMGPU_CUDA(c, cudaMemcpyAsync(free_out, c->q_out.p, static_cast<size_t>(C), cudaMemcpyDeviceToHost, c->stream));
    HRMGPU_CUDA(c, cudaStreamSynchronize(c->stream));  // (w is pageable host memory: its copy has completed when the call returns)
    return HRMGPU_OK;
}

int hrmgpu_test_transitions(hrmgpu_ctx* c, int C, const int* pair, const double* v0, const double* v1, int T, const double* w0,
                            const double* w1, const double* link_off, unsigned char* free_out) {
    return test_transitions_common(c, C, pair, v0, v1, T, w0, w1, link_off, nullptr, free_out);
}

int hrmgpu_test_transitions_articulated(hrmgpu_ctx* c, int C, const int* pair, const double* v0, const double* v1, int T, const double* w0,
                                        const double* w1, const double* link_off, const double* chain_off, unsigned char* free_out) {
    if (c && !chain_off) return fail(c, HRMGPU_E_ARG, "test_transitions_articulated: bad arguments");
    return test_transitions_common(c, C, pair, v0, v1, T, w0, w1, link_off, chain_off, free_out);
}

long long hrmgpu_connect_slices3d(hrmgpu_ctx* c, unsigned char* code_out, long long cap, long long* layer_first) {
    if (!c) return HRMGPU_E_ARG;
    int rc = check_built(c, 0);
    if (rc) return rc;
    if (c->dim != 3) return fail(c, HRMGPU_E_STATE, "connect_slices3d: 3D layers only");
    if (cap < 0 || (cap > 0 && !code_out)) return fail(c, HRMGPU_E_ARG, "connect_slices3d: bad arguments");
    HRMGPU_CUDA(c, cudaSetDevice(c->device));
    if ((rc = finalize_segments(c))) return rc;
    // the free-segment arrays live on the side stream's timeline: the main stream must see them finished
    HRMGPU_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_k3[c->ivl_cur], 0));
    const long long items = 2LL * c->K * c->L;
    if ((rc = ensure(c, c->pc_cnt, static_cast<size_t>(items)))) return rc;
    if ((rc = ensure(c, c->pc_first, static_cast<size_t>(items) + 1))) return rc;
    if ((rc = ensure(c, c->pc_info, 2))) return rc;
    if ((rc = launch_slice_pairs(c, false, nullptr, c->pc_cnt.p, nullptr, nullptr, nullptr))) return rc;
    if ((rc = launch_scan_counts(c, c->pc_cnt.p, c->pc_first.p, items, c->pc_info.p, c->stream))) return rc;
    // the first pair of every layer (and the total) is all the host needs to size and split the answer
    std::vector<long long> first(static_cast<size_t>(c->K) + 1);
    HRMGPU_CUDA(c, cudaMemcpy2DAsync(first.data(), sizeof(long long), c->pc_first.p, sizeof(long long) * 2 * static_cast<size_t>(c->L), sizeof(long long),
                                     static_cast<size_t>(c->K) + 1, cudaMemcpyDeviceToHost, c->stream));
    HRMGPU_CUDA(c, cudaStreamSynchronize(c->stream));
    const long long total = first[c->K];
    if (layer_first) std::copy(first.begin(), first.end(), layer_first);
    if (total > cap) return fail(c, HRMGPU_E_CAP, "connect_slices3d: code_out too small");
    if (5 * total > 0x7FFFFFF0LL) return fail(c, HRMGPU_E_CAP, "connect_slices3d: too many candidate segments for one launch");
    if (total == 0) return 0;
    const size_t E = static_cast<size_t>(5 * total);
    if ((rc = ensure(c, c->q_a, 3 * E))) return rc;
    if ((rc = ensure(c, c->q_b, 3 * E))) return rc;
    if ((rc = ensure(c, c->q_set, E))) return rc;
    if ((rc = ensure(c, c->q_out, E))) return rc;
    if ((rc = ensure(c, c->pc_code, static_cast<size_t>(total)))) return rc;
    if ((rc = launch_slice_pairs(c, true, c->pc_first.p, nullptr, c->q_a.p, c->q_b.p, c->q_set.p))) return rc;
    if ((rc = launch_edges(c, -1, static_cast<int>(E), c->q_a.p, c->q_b.p, c->q_out.p, c->q_set.p))) return rc;
    if ((rc = launch_pair_codes(c, c->q_out.p, total, c->pc_code.p))) return rc;
    HRMGPU_CUDA(c, cudaMemcpyAsync(code_out, c->pc_code.p, static_cast<size_t>(total), cudaMemcpyDeviceToHost, c->stream));
    HRMGPU_CUDA(c, cudaStreamSynchronize(c->stream));
    return total;
}

long long hrmgpu_connect_pairs3d(hrmgpu_ctx* c, long long V, const double* vertex_xyz, int P, const int* range4, double thx, double thy, int T,
                                 const double* w0, const double* w1, const double* link_off, const double* chain_off, int* match_out,
                                 long long match_cap, long long* n_candidates_out) {
    if (!c) return HRMGPU_E_ARG;
    if (c->dim != 3 || !c->bridge_built) return fail(c, HRMGPU_E_STATE, "connect_pairs3d: no 3D bridge layers built");
    if (V < 0 || V > 0x7FFFFFF0LL || P < 0 || P != c->P || T < 1 || (P > 0 && (!vertex_xyz || !range4 || !w0 || !w1 || !link_off || !match_out)))
        return fail(c, HRMGPU_E_ARG, "connect_pairs3d: bad arguments (P must equal the number of bridge layers)");
    if (n_candidates_out) *n_candidates_out = 0;
    std::vector<long long> m0_first(static_cast<size_t>(P) + 1, 0);
    long long max_n0 = 0;
    for (int i = 0; i < P; ++i) {
        const int* r = range4 + 4 * static_cast<size_t>(i);
        if (r[0] < 0 || r[1] < r[0] || r[1] > V || r[2] < 0 || r[3] < r[2] || r[3] > V) return fail(c, HRMGPU_E_ARG, "connect_pairs3d: vertex range out of bounds");
        m0_first[i + 1] = m0_first[i] + (r[1] - r[0]);
        max_n0 = std::max<long long>(max_n0, r[1] - r[0]);
    }
    const long long M0 = m0_first[P];
    if (M0 > match_cap) return fail(c, HRMGPU_E_CAP, "connect_pairs3d: match_out too small");
    if (M0 == 0) return 0;
    if (M0 > 0x7FFFFFF0LL) return fail(c, HRMGPU_E_CAP, "connect_pairs3d: too many vertices");
    HRMGPU_CUDA(c, cudaSetDevice(c->device));
    int rc;
    if ((rc = upload(c, c->pc_vtx, vertex_xyz, 3 * static_cast<size_t>(V)))) return rc;
    if ((rc = upload(c, c->pc_layers, range4, 4 * static_cast<size_t>(P)))) return rc;
    if ((rc = upload(c, c->pc_m0_first, m0_first.data(), m0_first.size()))) return rc;
    if ((rc = ensure(c, c->pc_cnt, static_cast<size_t>(M0)))) return rc;
    if ((rc = ensure(c, c->pc_first, static_cast<size_t>(M0) + 1))) return rc;
    if ((rc = ensure(c, c->pc_info, 2))) return rc;
    PairParams p;
    p.vtx = c->pc_vtx.p;
    p.range = c->pc_layers.p;
    p.m0_first = c->pc_m0_first.p;
    p.cand_first = c->pc_first.p;
    p.cnt = c->pc_cnt.p;
    p.c_pair = nullptr, p.c_v0 = nullptr, p.c_v1 = nullptr, p.c_m1 = nullptr;
    p.P = P;
    p.thx = thx, p.thy = thy;
    if ((rc = launch_pair_candidates(c, p, max_n0, false))) return rc;
    if ((rc = launch_scan_counts(c, c->pc_cnt.p, c->pc_first.p, M0, c->pc_info.p, c->stream))) return rc;
    long long total = 0;
    HRMGPU_CUDA(c, cudaMemcpyAsync(&total, c->pc_first.p + M0, sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
    HRMGPU_CUDA(c, cudaStreamSynchronize(c->stream));  // (also: the uploads above come from pageable host memory)
    if (n_candidates_out) *n_candidates_out = total;
    if (total > 0x7FFFFFF0LL) return fail(c, HRMGPU_E_CAP, "connect_pairs3d: too many candidate pairs for one launch");
    const size_t C = static_cast<size_t>(total);
    if ((rc = ensure(c, c->pc_match, static_cast<size_t>(M0)))) return rc;
    if ((rc = ensure(c, c->pc_m1, C))) return rc;
    if ((rc = ensure(c, c->q_out, C))) return rc;
    if (C > 0) {
        if ((rc = ensure(c, c->q_set, C))) return rc;
        if ((rc = ensure(c, c->q_a, 3 * C))) return rc;
        if ((rc = ensure(c, c->q_b, 3 * C))) return rc;
        p.c_pair = c->q_set.p, p.c_v0 = c->q_a.p, p.c_v1 = c->q_b.p, p.c_m1 = c->pc_m1.p;
        if ((rc = launch_pair_candidates(c, p, max_n0, true))) return rc;
        std::vector<double> w(2 * static_cast<size_t>(T));
        for (int s = 0; s < T; ++s) w[s] = w0[s], w[T + s] = w1[s];
        if ((rc = upload(c, c->q_w, w.data(), w.size()))) return rc;
        const size_t nl = static_cast<size_t>(c->P) * T * c->bridge_B * 3;
        if ((rc = upload(c, c->q_loff, link_off, nl))) return rc;
        if (chain_off && (rc = upload(c, c->q_loff2, chain_off, nl))) return rc;
        if ((rc = launch_transitions(c, static_cast<int>(C), T, c->q_set.p, c->q_a.p, c->q_b.p, c->q_w.p, c->q_w.p + T, c->q_loff.p, c->q_out.p,
                                     chain_off ? c->q_loff2.p : nullptr)))
            return rc;
    }
    if ((rc = launch_pair_replay(c, P, c->pc_m0_first.p, c->pc_first.p, c->pc_m1.p, c->q_out.p, c->pc_match.p))) return rc;
    HRMGPU_CUDA(c, cudaMemcpyAsync(match_out, c->pc_match.p, sizeof(int) * static_cast<size_t>(M0), cudaMemcpyDeviceToHost, c->stream));
    HRMGPU_CUDA(c, cudaStreamSynchronize(c->stream));  // (w is pageable host memory: its copy has completed when the call returns)
    return M0;
}

int hrmgpu_get_dims(hrmgpu_ctx* c, long long* out) {
    if (!c || !out) return HRMGPU_E_ARG;
    if (!c->built) return fail(c, HRMGPU_E_STATE, "no layers built yet");
    if (int rc = finalize_segments(c)) return rc;
    out[0] = c->K, out[1] = c->B, out[2] = c->S, out[3] = c->Sa, out[4] = c->L, out[5] = c->M, out[6] = c->seg_total, out[7] = c->dim;
    return HRMGPU_OK;
}

int hrmgpu_get_boundary(hrmgpu_ctx* c, int k, int surface, double* xyz) {
    if (!c || !xyz) return HRMGPU_E_ARG;
    int rc = check_built(c, k);
    if (rc) return rc;
    if (surface < 0 || surface >= c->S) return fail(c, HRMGPU_E_ARG, "surface index out of range");
    HRMGPU_CUDA(c, cudaSetDevice(c->device));
    const int D = c->dim, M = c->M;
    const size_t cnt = static_cast<size_t>(D) * M;
    const size_t base = (static_cast<size_t>(k) * c->S + surface) * cnt;
    std::vector<double> planar(cnt);
    if (c->prec == HRMGPU_F32) {
        std::vector<float> pf(cnt);
        HRMGPU_CUDA(c, cudaMemcpyAsync(pf.data(), reinterpret_cast<float*>(c->mesh.p) + base, cnt * 4, cudaMemcpyDeviceToHost, c->stream));
        HRMGPU_CUDA(c, cudaStreamSynchronize(c->stream));
        for (size_t i = 0; i < cnt; ++i) planar[i] = pf[i];
    } else {
        HRMGPU_CUDA(c, cudaMemcpyAsync(planar.data(), reinterpret_cast<double*>(c->mesh.p) + base, cnt * 8, cudaMemcpyDeviceToHost, c->stream));
        HRMGPU_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    for (int i = 0; i < M; ++i)
        for (int d = 0; d < D; ++d) xyz[static_cast<size_t>(i) * D + d] = planar[static_cast<size_t>(d) * M + i];
    return HRMGPU_OK;
}

int hrmgpu_get_boundary_layer(hrmgpu_ctx* c, int k, double* xyz) {
    if (!c || !xyz) return HRMGPU_E_ARG;
    int rc = check_built(c, k);
    if (rc) return rc;
    HRMGPU_CUDA(c, cudaSetDevice(c->device));
    const int D = c->dim, M = c->M;
    const long long total = static_cast<long long>(c->S) * D * M;
    if ((rc = ensure(c, c->bound_tmp, static_cast<size_t>(total)))) return rc;
    const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, 8LL * c->num_sms));
    if (c->prec == HRMGPU_F32)
        k_interleave_boundary<float><<<grid, 256, 0, c->stream>>>(reinterpret_cast<const float*>(c->mesh.p) + static_cast<size_t>(k) * total, c->bound_tmp.p, D, M, total);
    else
        k_interleave_boundary<double><<<grid, 256, 0, c->stream>>>(reinterpret_cast<const double*>(c->mesh.p) + static_cast<size_t>(k) * total, c->bound_tmp.p, D, M, total);
    HRMGPU_CUDA(c, cudaGetLastError());
    c->launches += 1;
    HRMGPU_CUDA(c, cudaMemcpyAsync(xyz, c->bound_tmp.p, static_cast<size_t>(total) * 8, cudaMemcpyDeviceToHost, c->stream));
    HRMGPU_CUDA(c, cudaStreamSynchronize(c->stream));
    return HRMGPU_OK;
}

int hrmgpu_get_intervals(hrmgpu_ctx* c, int k, double* lo, double* up) {
    if (!c || !lo || !up) return HRMGPU_E_ARG;
    int rc = check_built(c, k);
    if (rc) return rc;
    HRMGPU_CUDA(c, cudaSetDevice(c->device));
    const size_t cnt = static_cast<size_t>(c->S) * c->L;
    std::vector<double> a(cnt), b(cnt);
    HRMGPU_CUDA(c, cudaMemcpyAsync(a.data(), c->ivl_lo[c->ivl_cur].p + static_cast<size_t>(k) * cnt, cnt * 8, cudaMemcpyDeviceToHost, c->stream));
    HRMGPU_CUDA(c, cudaMemcpyAsync(b.data(), c->ivl_up[c->ivl_cur].p + static_cast<size_t>(k) * cnt, cnt * 8, cudaMemcpyDeviceToHost, c->stream));
    HRMGPU_CUDA(c, cudaStreamSynchronize(c->stream));
    for (int s = 0; s < c->S; ++s)
        for (int l = 0; l < c->L; ++l) {
            lo[static_cast<size_t>(l) * c->S + s] = a[static_cast<size_t>(s) * c->L + l];
            up[static_cast<size_t>(l) * c->S + s] = b[static_cast<size_t>(s) * c->L + l];
        }
    return HRMGPU_OK;
}

// Host mirror of the CSR offsets (pinned), fetched the first time a getter needs it after a build.
static int fetch_offsets(hrmgpu_ctx* c) {
    int rc = finalize_segments(c);
    if (rc) return rc;
    if (c->h_off_valid) return HRMGPU_OK;
    const size_t nl = static_cast<size_t>(c->K) * c->L;
    if (c->h_off_cap < nl + 1) {
        if (c->h_off) cudaFreeHost(c->h_off);
        c->h_off = nullptr, c->h_off_cap = 0;
        HRMGPU_CUDA(c, cudaMallocHost(reinterpret_cast<void**>(&c->h_off), (nl + 1) * sizeof(long long)));
        c->h_off_cap = nl + 1;
    }
    if (nl == 0) {
        c->h_off[0] = 0;
    } else {
        HRMGPU_CUDA(c, cudaMemcpyAsync(c->h_off, c->seg_off.p, sizeof(long long) * (nl + 1), cudaMemcpyDeviceToHost, c->s2));
        HRMGPU_CUDA(c, cudaStreamSynchronize(c->s2));
    }
    c->h_off_valid = true;
    return HRMGPU_OK;
}

int hrmgpu_get_segments(hrmgpu_ctx* c, int k, int* off, double* xL, double* xU, double* xM, int cap) {
    if (!c) return HRMGPU_E_ARG;
    int rc = check_built(c, k);
    if (rc) return rc;
    HRMGPU_CUDA(c, cudaSetDevice(c->device));
    if ((rc = fetch_offsets(c))) return rc;
    const size_t l0 = static_cast<size_t>(k) * c->L;
    const long long b0 = c->h_off[l0], b1 = c->h_off[l0 + c->L];
    const long long cnt = b1 - b0;
    if (off)
        for (int l = 0; l <= c->L; ++l) off[l] = static_cast<int>(c->h_off[l0 + l] - b0);
    if (xL) {
        if (!xU || !xM) return fail(c, HRMGPU_E_ARG, "get_segments: xU/xM are NULL");
        if (cap < cnt) return fail(c, HRMGPU_E_CAP, "get_segments: buffer too small");
        if (cnt) {
            const double* d = c->segdata.p;
            HRMGPU_CUDA(c, cudaMemcpyAsync(xL, d + b0, cnt * 8, cudaMemcpyDeviceToHost, c->s2));
            HRMGPU_CUDA(c, cudaMemcpyAsync(xU, d + c->seg_cap + b0, cnt * 8, cudaMemcpyDeviceToHost, c->s2));
            HRMGPU_CUDA(c, cudaMemcpyAsync(xM, d + 2 * c->seg_cap + b0, cnt * 8, cudaMemcpyDeviceToHost, c->s2));
            HRMGPU_CUDA(c, cudaStreamSynchronize(c->s2));
        }
    }
    return static_cast<int>(cnt);
}

int hrmgpu_get_segments_all(hrmgpu_ctx* c, long long* off, double* xL, double* xU, double* xM, long long cap) {
    if (!c) return HRMGPU_E_ARG;
    if (!c->built) return fail(c, HRMGPU_E_STATE, "no layers built yet");
    HRMGPU_CUDA(c, cudaSetDevice(c->device));
    int rc = finalize_segments(c);
    if (rc) return rc;
    const size_t nl = static_cast<size_t>(c->K) * c->L;
    if (xL && (!xU || !xM)) return fail(c, HRMGPU_E_ARG, "get_segments_all: xU/xM are NULL");
    if (xL && cap < c->seg_total) return fail(c, HRMGPU_E_CAP, "get_segments_all: buffer too small");
    // offsets straight into the caller's buffer (no staging), payload behind them on the same stream
    if (off) {
        if (nl == 0)
            off[0] = 0;
        else
            HRMGPU_CUDA(c, cudaMemcpyAsync(off, c->seg_off.p, sizeof(long long) * (nl + 1), cudaMemcpyDeviceToHost, c->s2));
    }
    if (xL && c->seg_total) {
        const double* d = c->segdata.p;
        HRMGPU_CUDA(c, cudaMemcpyAsync(xL, d, c->seg_total * 8, cudaMemcpyDeviceToHost, c->s2));
        HRMGPU_CUDA(c, cudaMemcpyAsync(xU, d + c->seg_cap, c->seg_total * 8, cudaMemcpyDeviceToHost, c->s2));
        HRMGPU_CUDA(c, cudaMemcpyAsync(xM, d + 2 * c->seg_cap, c->seg_total * 8, cudaMemcpyDeviceToHost, c->s2));
    }
    HRMGPU_CUDA(c, cudaStreamSynchronize(c->s2));
    return HRMGPU_OK;
}

// Pipelined read-back: the copies are queued on the side stream right behind the free-segment passes of the last build and the
// call returns; the next hrmgpu_build_layers* may be issued at once (its fused kernel runs on the main stream meanwhile).
long long hrmgpu_fetch_segments_async(hrmgpu_ctx* c, long long* off, double* xL, double* xU, double* xM, long long cap) {
    if (!c) return HRMGPU_E_ARG;
    if (!c->built) return fail(c, HRMGPU_E_STATE, "no layers built yet");
    if (!off || cap < 0 || (cap > 0 && (!xL || !xU || !xM))) return fail(c, HRMGPU_E_ARG, "fetch_segments_async: bad arguments");
    HRMGPU_CUDA(c, cudaSetDevice(c->device));
    const size_t nl = static_cast<size_t>(c->K) * c->L;
    c->fetch_cap = cap;
    c->fetch_pending = true;
    if (nl == 0) {
        off[0] = 0;
        return 0;
    }
    HRMGPU_CUDA(c, cudaMemcpyAsync(off, c->seg_off.p, sizeof(long long) * (nl + 1), cudaMemcpyDeviceToHost, c->s2));
    const long long cnt = cap < c->seg_cap ? cap : c->seg_cap;
    if (cnt > 0) {
        const double* d = c->segdata.p;
        HRMGPU_CUDA(c, cudaMemcpyAsync(xL, d, cnt * 8, cudaMemcpyDeviceToHost, c->s2));
        HRMGPU_CUDA(c, cudaMemcpyAsync(xU, d + c->seg_cap, cnt * 8, cudaMemcpyDeviceToHost, c->s2));
        HRMGPU_CUDA(c, cudaMemcpyAsync(xM, d + 2 * c->seg_cap, cnt * 8, cudaMemcpyDeviceToHost, c->s2));
    }
    // the sizes of THIS build in their own pinned slot (h_info[0..1] belongs to the build queued last)
    HRMGPU_CUDA(c, cudaMemcpyAsync(c->h_info + 2, c->seg_info.p, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->s2));
    HRMGPU_CUDA(c, cudaEventRecord(c->ev_fetch, c->s2));
    return cnt;
}

long long hrmgpu_wait_segments(hrmgpu_ctx* c) {
    if (!c) return HRMGPU_E_ARG;
    if (!c->fetch_pending) return fail(c, HRMGPU_E_STATE, "wait_segments: no fetch in flight");
    HRMGPU_CUDA(c, cudaSetDevice(c->device));
    c->fetch_pending = false;
    if (static_cast<size_t>(c->K) * c->L == 0) return 0;
    HRMGPU_CUDA(c, cudaEventSynchronize(c->ev_fetch));
    const long long need_orig = static_cast<long long>(c->h_info[2]), need_seg = static_cast<long long>(c->h_info[3]);
    if (need_orig > c->orig_cap || need_seg > c->seg_cap || need_seg > c->fetch_cap)
        return fail(c, HRMGPU_E_CAP, "wait_segments: capacity exceeded (read the build with hrmgpu_get_segments_all, which grows the buffers)");
    return need_seg;
}

int hrmgpu_get_single_hit_count(hrmgpu_ctx* c, long long* out) {
    if (!c || !out) return HRMGPU_E_ARG;
    HRMGPU_CUDA(c, cudaSetDevice(c->device));
    unsigned long long v = 0;
    HRMGPU_CUDA(c, cudaMemcpyAsync(&v, c->counters.p, sizeof(v), cudaMemcpyDeviceToHost, c->stream));
    HRMGPU_CUDA(c, cudaStreamSynchronize(c->stream));
    *out = static_cast<long long>(v);
    return HRMGPU_OK;
}

long long hrmgpu_pack_segments_dev(hrmgpu_ctx* c, void* d_buf, long long cap_bytes) {
    if (!c) return HRMGPU_E_ARG;
    if (!c->built) return fail(c, HRMGPU_E_STATE, "no layers built yet");
    if (int rc = finalize_segments(c)) return rc;
    const long long nl = static_cast<long long>(c->K) * c->L;
    const long long need = (nl + 1) * 8 + 3 * c->seg_total * 8;
    if (!d_buf) return need;
    if (cap_bytes < need) return fail(c, HRMGPU_E_CAP, "pack_segments_dev: buffer too small");
    HRMGPU_CUDA(c, cudaSetDevice(c->device));
    char* dst = static_cast<char*>(d_buf);
    // on the context's MAIN stream (the one hrmgpu_set_stream installed), behind the free-segment passes
    HRMGPU_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_k3[c->ivl_cur], 0));
    HRMGPU_CUDA(c, cudaMemcpyAsync(dst, c->seg_off.p, (nl + 1) * 8, cudaMemcpyDeviceToDevice, c->stream));
    dst += (nl + 1) * 8;
    if (c->seg_total) {
        const double* d = c->segdata.p;
        HRMGPU_CUDA(c, cudaMemcpyAsync(dst, d, c->seg_total * 8, cudaMemcpyDeviceToDevice, c->stream));
        HRMGPU_CUDA(c, cudaMemcpyAsync(dst + c->seg_total * 8, d + c->seg_cap, c->seg_total * 8, cudaMemcpyDeviceToDevice, c->stream));
        HRMGPU_CUDA(c, cudaMemcpyAsync(dst + 2 * c->seg_total * 8, d + 2 * c->seg_cap, c->seg_total * 8, cudaMemcpyDeviceToDevice, c->stream));
    }
    // Stream semantics (see hrmgpu.h): with the context's own stream the call returns when the buffer is complete; with a
    // caller-installed stream the copies are ordered on THAT stream and the caller keeps using it.
    if (c->stream == c->own_stream) HRMGPU_CUDA(c, cudaStreamSynchronize(c->stream));
    return need;
}

// Profiling aid (not part of the drop-in surface): with enable != 0 the next 3D build records, per CTA of the
// fused kernel, clock64() durations of its phases; out (if non-NULL) receives [n_cta][8] values of the last build.
long long hrmgpu_debug_phase_times(hrmgpu_ctx* c, int enable, long long n_cta, long long* out) {
    if (!c) return HRMGPU_E_ARG;
    if (cudaSetDevice(c->device) != cudaSuccess) return HRMGPU_E_CUDA;
    if (!enable) {
        release(c, c->dbg);
        return 0;
    }
    if (out && c->dbg.p && c->dbg.n >= static_cast<size_t>(n_cta) * 8) {
        cudaStreamSynchronize(c->stream);
        cudaMemcpy(out, c->dbg.p, sizeof(long long) * n_cta * 8, cudaMemcpyDeviceToHost);
        return n_cta;
    }
    int rc = ensure(c, c->dbg, static_cast<size_t>(n_cta) * 8);
    return rc ? rc : 0;
}

// Profiling aid: re-runs K1 alone (boundary points of the last 3D build's poses, no sweep, meshes overwritten in place)
// and returns the kernel time in ms -- the ceiling of the fused kernel if the sweep were free.  The pose arrays of the last
// build must still be alive (they are the caller's for hrmgpu_build_layers_dev).
double hrmgpu_debug_k1_only_ms(hrmgpu_ctx* c) {
    if (!c || !c->built || c->dim != 3 || !c->last_semi || !c->last_R || !c->last_off) return -1.0;
    if (cudaSetDevice(c->device) != cudaSuccess) return -1.0;
    cudaEventRecord(c->ev0, c->stream);
    if (launch_minksweep3d(c, c->K, c->B, 0, c->n_arena + c->n_obs, c->last_semi, c->last_R, c->last_off, c->prec, c->mesh.p, nullptr, nullptr,
                           false))
        return -1.0;
    cudaEventRecord(c->ev1, c->stream);
    if (cudaEventSynchronize(c->ev1) != cudaSuccess) return -1.0;
    float ms = 0;
    if (cudaEventElapsedTime(&ms, c->ev0, c->ev1) != cudaSuccess) return -1.0;
    return static_cast<double>(ms);
}

long long hrmgpu_kernel_launches(hrmgpu_ctx* c, int reset) {
    if (!c) return HRMGPU_E_ARG;
    const long long v = c->launches;
    if (reset) c->launches = 0;
    return v;
}

double hrmgpu_last_segments_ms(hrmgpu_ctx* c) {
    if (!c || !c->built) return -1.0;
    if (cudaSetDevice(c->device) != cudaSuccess) return -1.0;
    if (cudaEventSynchronize(c->ev3) != cudaSuccess) return -1.0;
    float ms = 0;
    if (cudaEventElapsedTime(&ms, c->ev2, c->ev3) != cudaSuccess) return -1.0;
    return static_cast<double>(ms);
}

long long hrmgpu_segment_redos(hrmgpu_ctx* c, int reset) {
    if (!c) return HRMGPU_E_ARG;
    const long long v = c->seg_redos;
    if (reset) c->seg_redos = 0;
    return v;
}

double hrmgpu_last_minksweep_ms(hrmgpu_ctx* c) {
    if (!c || !c->built) return -1.0;
    if (cudaSetDevice(c->device) != cudaSuccess) return -1.0;
    if (cudaEventSynchronize(c->ev1) != cudaSuccess) return -1.0;
    float ms = 0;
    if (cudaEventElapsedTime(&ms, c->ev0, c->ev1) != cudaSuccess) return -1.0;
    return static_cast<double>(ms);
}

}  // extern "C"
