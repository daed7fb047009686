// Internal declarations shared by the translation units of libhrmgpu.so (sm_100a only).
#pragma once
#include "../../include/hrmgpu.h"

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

namespace hrmgpu {

constexpr int kThreads = 256;          // CTA size of the fused Minkowski+sweep kernel
constexpr int kCellCap = 2048;         // kept mesh cells listed per scan round in shared memory
constexpr int kMaxGrid = 200;          // n_grid limit (vertex line codes: 4*n^2 bytes of shared memory)
constexpr int kMaxLinesAxis = 16383;   // per-axis sweep-line limit (15-bit line codes)
constexpr int kMaxLines = 8192;        // Lx*Ly limit (two-slot hit tables in shared memory)
constexpr uint32_t kNoFace = 0xFFFFFFFFu;
constexpr size_t kPoolKeepBytes = size_t(1) << 30;  // freed device memory kept in the pool across contexts (hrmgpu_destroy trims the rest)

// Device-side description of one arena/obstacle superquadric (SoA would not help: 17 doubles read
// once per CTA through the read-only path).
struct Obj3 {
    double a, b, c;
    double px, py, pz;
    double R[9];
    double sign;  // -1 arena (Minkowski difference), +1 obstacle (sum): SuperQuadrics.cpp:84-85
    double ia, ib, ic;  // 1/a, 1/b, 1/c for the non-strict modes (the strict mode divides like the reference)
};

struct Obj2 {
    double a, b, eps;
    double px, py;
    double c1, s1;  // cos / sin of the object's angle, evaluated on the host with libm like the reference
    double sign;
};

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;  // capacity in elements
};

struct Ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaMemPool_t pool = nullptr;  // the library's private pool of this device (see ensure())
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;  // around the fused kernel of the last build (main stream)
    // Side stream of the free-segment passes (K3) and of the exchange: the passes of build i run behind `ev_fused` while the main
    // stream already runs the fused kernel of build i+1 into the other interval-table buffer.
    cudaStream_t s2 = nullptr;
    cudaEvent_t ev_fused = nullptr, ev2 = nullptr, ev3 = nullptr;  // fused kernel done; around the K3 passes (side stream)
    cudaEvent_t ev_k3[2] = {nullptr, nullptr};                     // K3 passes that read interval buffer [i] are done
    std::string err;
    long long launches = 0;

    // scene
    int dim = 0;  // 2 or 3 once a scene is set
    int n_arena = 0, n_obs = 0, n = 0;
    DevBuf<Obj3> obj3;
    DevBuf<Obj2> obj2;
    DevBuf<double> tab;    // 3D: [S0][8][n] power tables; 2D: [S0][4][num] (x0 factors ce,se; gradient factors)
    // sweep
    bool sweep_set = false;
    double lim[6] = {0, 0, 0, 0, 0, 0};
    int Lx = 0, Ly = 0;
    DevBuf<double> txs, tys;  // sentinel-padded line coordinates: [-inf, t_0..t_{L-1}, +inf]

    // last build
    int K = 0, B = 0, S = 0, Sa = 0, L = 0, M = 0, prec = 0;
    bool built = false;
    DevBuf<double> body_semi, body_R, body_off;  // staged copies of host inputs
    const double *last_semi = nullptr, *last_R = nullptr, *last_off = nullptr;  // device pose arrays of the last 3D build
    DevBuf<double> q_a, q_b;                     // staged query inputs (edge end points, body centres)
    DevBuf<unsigned char> q_out;                 // staged query results
    DevBuf<double> bbox;                         // [meshes][6] min/max per axis of the queried meshes
    DevBuf<unsigned int> q_flag;                 // per-query blocked flags
    DevBuf<int> q_set;                           // multi-layer queries: layer / pair index of every query
    DevBuf<char> mesh;                           // [K][S][dim][M] planar, double (or float in F32 mode)
    DevBuf<double> ivl_lo[2], ivl_up[2];         // [K][S][L], double-buffered (see s2)
    int ivl_cur = 0;                             // buffer of the last build
    // Per-line buckets of the obstacle intervals, filled by the warp-specialised fused kernel beside the dense tables (same double
    // buffering): bcnt [K][L] entries appended (may exceed bucket_cap: that line then takes the dense path), bucket [K][L][cap][2].
    DevBuf<int> bcnt[2];
    DevBuf<double> bucket[2];
    bool bucket_valid[2] = {false, false};
    int bucket_cap = 0;
    int bucket_cap_hook = 0;                     // test hook: bucket capacity (small values force the dense fallback per line)
    int ws_variant = 1;                          // experiment hook: warp / register split of the warp-specialised kernel (mink3d_ws.cu)
    bool semi_per_layer = false;                 // set around the bridge build: body semi-axes per layer (pair), see MinkParams
    bool force_classic = false;                  // test / profiling hook: never use the warp-specialised kernel
    bool ws_generic = false;                     // test hook (HRMGPU_WS_GENERIC): run-time grid size even where a fixed-n instantiation exists
    int num_sms = 148;                           // multiprocessors of the device
    int seg_fast_cap = 0;                        // test hook: list capacity of the 4-line segment pass (0 = default 128)
    long long seg_init_cap = 0;                  // test hook: first capacity of the segment pool / payload (forces the grow-and-repeat path)
    bool force_seg_fallback = false;             // test hook: always use the one-line-per-warp pass
    int cell_seg_cap = 0;                        // test hook: kept-cell list capacity per warp of the fused kernel (0 = default)
    DevBuf<double> orig;                         // pool of original (pre-enhancement) segments [orig_cap][2]
    DevBuf<long long> orig_off;                  // [K][L] first pool entry of a line
    DevBuf<int> orig_cnt;                        // [K][L]
    DevBuf<int> seg_cnt;                         // [K][L]
    DevBuf<long long> seg_off;                   // [K*L+1]
    DevBuf<double> segdata;                      // xL[seg_cap] xU[seg_cap] xM[seg_cap]
    DevBuf<unsigned long long> seg_info;         // [0] pool entries needed, [1] total segments
    long long orig_cap = 0, seg_cap = 0;         // capacities (entries) of orig / of each array in segdata
    long long orig_hint = 0, seg_hint = 0;       // what earlier builds needed (x1.5): first guess of the next build
    unsigned long long* h_info = nullptr;        // pinned host copy of seg_info
    long long* h_off = nullptr;                  // pinned host mirror of seg_off, fetched on demand
    size_t h_off_cap = 0;
    bool h_off_valid = false;
    bool seg_pending = false;                    // passes queued, sizes not yet checked by finalize_segments()
    long long seg_redos = 0;                     // passes repeated because a capacity was too small
    cudaEvent_t ev_join = nullptr;               // side-stream marker (hrmgpu_join, exchange on a caller's stream)
    cudaEvent_t ev_fetch = nullptr;              // end of the copies of hrmgpu_fetch_segments_async
    long long fetch_cap = 0;                     // capacity of the host arrays of the fetch in flight
    bool fetch_pending = false;
    DevBuf<char> xchg_send, xchg_recv;           // multi-GPU exchange (hrmgpu_allgather_layers): packed send / receive buffers
    DevBuf<double> bound_tmp;                    // hrmgpu_get_boundary_layer: transposed copy of one layer
    void* comm = nullptr;                        // ncclComm_t created by hrmgpu_comm_init (exchange.cu)
    int xchg_world = 0, xchg_rank = 0;           // last exchange
    long long xchg_line_cap = 0, xchg_seg_cap = 0;
    void* xchg_stream = nullptr;
    DevBuf<unsigned long long> counters;         // [0] single-hit count, [1] capacity overflow flag
    DevBuf<long long> dbg;                       // optional per-CTA phase timings (hrmgpu_debug_phase_times)
    long long seg_total = 0;                     // valid after finalize_segments()

    // bridge layers
    int P = 0, bridge_B = 0, bridge_prec = 0;
    bool bridge_built = false;
    DevBuf<char> bridge_mesh;  // [P][n_obs][B][dim][M]
    DevBuf<double> bridge_semi, bridge_R, bridge_off;
    DevBuf<double> bridge_bbox;                  // [P][n_obs * B][6] (3D)
    DevBuf<double> bridge_bands;                 // per bridge mesh: band boxes [2][n-1][4] + face boxes [2(n-1)^2][4] (3D, n - 1 <= 32): k_mesh_bands
    DevBuf<unsigned long long> scan_state_main, scan_state_side;  // k_scan_counts: ticket + per-tile inclusive totals, per stream
    DevBuf<double> face_n;                       // K4 3D: unit face normals of the queried layers' meshes (k_face_normals)
    bool no_face_normals = false;                // env HRMGPU_NO_FACE_NORMALS=1 (A/B switch)
    bool bridge_bands_ok = false, no_bands = false;  // no_bands: env HRMGPU_NO_BANDS=1 (A/B switch of the banded face walk)
    DevBuf<int> pc_layers, pc_cnt, pc_m1, pc_match;   // device-side vertex matching of connectMultiSlice (connect.cu)
    DevBuf<double> pc_vtx;
    DevBuf<unsigned char> pc_code;                    // connection code of every vertex pair of connectOneSlice
    DevBuf<double> tfe_in, tfe_quat;                  // device TFE: {body semi [B][3], quat a [P][B][4], quat b [P][B][4]}, result quaternions
    DevBuf<long long> pc_m0_first, pc_first;
    DevBuf<unsigned long long> pc_info;
    DevBuf<double> q_w, q_loff, q_loff2;         // transition tests: interpolation weights [2][T], link offsets [P][T][B][dim] (+ chain offsets)
};

}  // namespace hrmgpu
struct hrmgpu_ctx : public hrmgpu::Ctx {};
namespace hrmgpu {

#define HRMGPU_CUDA(ctx, expr)                                                                        \
    do {                                                                                              \
        cudaError_t _e = (expr);                                                                      \
        if (_e != cudaSuccess) {                                                                      \
            (ctx)->err = std::string(#expr) + ": " + cudaGetErrorString(_e);                          \
            return (_e == cudaErrorMemoryAllocation) ? HRMGPU_E_NOMEM : HRMGPU_E_CUDA;                \
        }                                                                                             \
    } while (0)

// Device buffers come from a stream-ordered memory pool PRIVATE to this library (one per device, created by the first
// hrmgpu_create on that device, cudaMallocFromPoolAsync on the context's stream): a planner object creates a context, grows
// ~30 buffers and destroys it again, and with plain cudaMalloc / cudaFree some of those calls take tens of milliseconds
// (measured: sporadic 20-60 ms inside K4 / K5a / buildLayers of HRM3D::plan()).  The pool keeps freed blocks (release
// threshold raised) so the next context reuses them; hrmgpu_destroy trims it to kPoolKeepBytes of free blocks.  The device's
// DEFAULT pool -- shared with every other cudaMallocAsync user of the process, e.g. PyTorch -- is never touched.
cudaMemPool_t library_pool(int device);  // ctx.cu; nullptr if the pool could not be created
template <class T>
inline int ensure(Ctx* c, DevBuf<T>& b, size_t n) {
    if (b.n >= n && b.p) return HRMGPU_OK;
    if (b.p) {
        cudaFreeAsync(b.p, c->stream);  // stream-ordered: work already queued on the stream may still use it
        b.p = nullptr;
        b.n = 0;
    }
    if (n == 0) n = 1;
    HRMGPU_CUDA(c, cudaMallocFromPoolAsync(reinterpret_cast<void**>(&b.p), n * sizeof(T), c->pool, c->stream));
    b.n = n;
    return HRMGPU_OK;
}

template <class T>
inline int ensure_on(Ctx* c, cudaStream_t st, DevBuf<T>& b, size_t n) {  // ensure() ordered on another stream of the context
    cudaStream_t keep = c->stream;
    c->stream = st;
    const int rc = ensure(c, b, n);
    c->stream = keep;
    return rc;
}

inline int fail(Ctx* c, int code, const std::string& msg) {
    c->err = msg;
    return code;
}

// Loads every kernel of a translation unit (cudaFuncGetAttributes) so that CUDA's lazy module loading does not spread
// 15-90 ms stalls over the first planner calls; called once per process by the first hrmgpu_create.
void preload_mink3d();
void preload_mink2d();
void preload_segments();
void preload_queries();
void preload_exchange();

// kernels / launchers implemented in the other translation units
int launch_minksweep3d(Ctx* c, int K, int B, int n_first_obj, int n_objs, const double* d_semi, const double* d_R,
                       const double* d_off, int precision, char* d_mesh, double* d_lo, double* d_up, bool do_sweep, bool reload = false);
int launch_minksweep2d(Ctx* c, int K, int B, int n_first_obj, int n_objs, const double* d_semi, const double* d_angle,
                       const double* d_off, int precision, char* d_mesh, double* d_lo, double* d_up, bool do_sweep, bool reload = false);
int launch_segments(Ctx* c);
int finalize_segments(Ctx* c);
int launch_edges(Ctx* c, int k, int E, const double* d_v1, const double* d_v2, unsigned char* d_out, const int* d_layer = nullptr);
int launch_bridge_test(Ctx* c, int p, int Q, const double* d_centres, unsigned char* d_out, const int* d_pair = nullptr);
int launch_bridge_bbox(Ctx* c);
// connect.cu: connectMultiSlice's vertex matching from the CSR free segments (HRM3D.cpp:158-224)
struct PairParams {
    const double* vtx;         // [V][3] vertex positions of the roadmap
    const int* range;          // [P][4] vertex id ranges {near begin, near end, cur begin, cur end}
    const long long* m0_first; // [P + 1] running count of m0 over the pairs
    long long* cand_first;     // [M0 + 1] offsets of every m0's candidates (fill pass)
    int* cnt;                  // [M0] (count pass)
    int* c_pair;               // [C]
    double* c_v0;              // [C][3]
    double* c_v1;              // [C][3]
    int* c_m1;                 // [C] vertex id of the candidate m1
    int P;
    double thx, thy;
};
int launch_pair_candidates(Ctx* c, const PairParams& p, long long max_n0, bool fill);
int launch_pair_replay(Ctx* c, int P, const long long* m0_first, const long long* cand_first, const int* c_m1, const unsigned char* free_flag, int* match);
int launch_slice_pairs(Ctx* c, bool fill, const long long* item_first, int* cnt, double* qa, double* qb, int* qlayer);
int launch_pair_codes(Ctx* c, const unsigned char* free5, long long n_pairs, unsigned char* code);
void preload_connect();
// tfe.cu: tightly-fitted ellipsoids of all (pair, body) combinations (TightFitEllipsoid.cpp:43-114)
int launch_tfe3d(Ctx* c, int P, int B, const double* d_body_semi, const double* d_quat_a, const double* d_quat_b, int num_step, double* d_semi,
                 double* d_rot, double* d_quat);
void preload_tfe();
// segments.cu: exclusive scan of n counts into n + 1 64-bit offsets (single CTA), info[1] = total
int launch_scan_counts(Ctx* c, const int* cnt, long long* off, long long n, unsigned long long* info, cudaStream_t st);
int launch_transitions(Ctx* c, int C, int T, const int* d_pair, const double* d_v0, const double* d_v1, const double* d_w0, const double* d_w1,
                       const double* d_link_off, unsigned char* d_out, const double* d_link_off2 = nullptr);

}  // namespace hrmgpu
