// K1+K2 fused, persistent and warp-specialised (the fast arithmetic modes at bench-like sizes).
//
// Same reference functions as mink3d.cu (SuperQuadrics.cpp:46-143, MultiBodyTree3D.cpp:91-105, MeshGenerator.cpp:105-131,
// FreeSpace3D.cpp:29-68, LineIntersection.cpp:56-171), same results; a different execution shape:
//
//   * 2 CTAs per SM stay resident for the whole launch and claim surfaces (layer, object, body) from a global counter,
//     arena surfaces first.  Per-launch work (sweep-line coordinates, shared-memory carve-up) is done once per CTA, not
//     once per surface.
//   * PRODUCER warps (8) run the boundary arithmetic of surface i+1 and stream its points to HBM (phase A of mink3d.cu,
//     same thread mapping: a group of lanes stores one whole eta-column per iteration) while the CONSUMER warps (4) sweep
//     surface i: SWAR scan of the vertex line codes, triangle tests of the kept cells, intervals.  The two halves meet
//     through two named barriers per buffer (FULL: codes + tables of a surface are in shared memory; EMPTY: the consumers
//     are done with them), so the HBM write stream never pauses for a sweep, whatever the CTAs of an SM are doing.
//   * The next surface's inputs (8 x n power table, object, body pose: ~6.7 KB) are fetched with cp.async one surface
//     ahead, so no warp ever waits for the L2/HBM round trip that opened every CTA of the one-surface kernel.
//   * The consumers never read a vertex back from global memory: a kept cell's corners are RE-EVALUATED from the
//     per-column / per-row factor tables the producers left in shared memory, with the same instruction sequence
//     (eval_point), hence bit-identical to the stored points.
//   * Every obstacle interval is also appended to a per-line bucket (one atomicAdd on the line's counter), so the
//     free-segment pass reads ~65 compact entries per line instead of striding over all 1001 surfaces.
//   * FP64 fast mode only (HRMGPU_F64; strict and FP32 builds run on mink3d.cu).  The squared norm of the gradient-like
//     vector u = ce2_c U_r + gz_c m (m = M1[:,2]) is evaluated in its expanded form a_c P_r + b_c Q_r + d_c with
//     P_r = |U_r|^2, Q_r = 2 U_r.m, a_c = ce2_c^2, b_c = ce2_c gz_c, d_c = gz_c^2 |m|^2: two FMAs instead of six, and U_r
//     never occupies registers.  Rounding differs from the direct form by a few ulps times the conditioning of M1
//     (<= 4 (max / min body semi-axis)^2); the points stay within the 1e-12 of HRMGPU_F64.
//   * The bench grid size n = 100 has its own instantiation (NF): strides and plane offsets become immediates.
#include "mink3d_common.cuh"

#include <cstdlib>

namespace hrmgpu {

constexpr int kPW = 8;                 // producer warps
constexpr int kPT = 32 * kPW;
#ifndef HRMGPU_WS_CW
#define HRMGPU_WS_CW 4
#endif
constexpr int kCW = HRMGPU_WS_CW;      // consumer warps (4; 5 = experiment: all threads at 72 registers, no setmaxnreg)
constexpr int kCT = 32 * kCW;
constexpr int kWsThreads = kPT + kCT;
constexpr int kDeferLines = 512;       // sweep lines whose bucket appends are deferred to the next surface (CT * lines per thread)
// Registers (2 CTAs per SM): 384 threads launched with 80 registers each; VAR 0 keeps 80 / 80, VAR 1 moves 8 per producer
// thread to the consumers (72 / 96, setmaxnreg).  (8 consumer warps at 48 / 56 registers were measured: 20-30 % slower.)
template <int VAR>
struct RegSplit {
    static constexpr int prod = (VAR == 1 && kCW == 4) ? 72 : 0;
    static constexpr int cons = (VAR == 1 && kCW == 4) ? 96 : 0;
};
constexpr int kWsCellCap = 1024;       // kept mesh cells listed per scan round (256 per consumer warp, as in the one-surface kernel)
constexpr int kZCap = 512;             // hit heights remembered per surface (phase C -> phase D)
constexpr uint32_t kZBits = 15, kZNone = (1u << kZBits) - 1u;  // hit key = face << 15 | pool index (kZNone: height not stored)
static_assert(kZCap < (1 << 15) && 2 * (kMaxGrid - 1) * (kMaxGrid - 1) < (1 << 17), "hit keys are 32-bit");
constexpr int kRowStride = 8;          // per-row factors V[3] W[3] P Q
constexpr int kWbStride = 24;          // per-warp scratch: 18 row-base coefficients, M1[:,2], M2[:,2]
constexpr int kRawExtra = 36;          // doubles behind the table in a raw input buffer: Obj3 (19, padded to 20), R[9], off[3], semi[3]
static_assert(sizeof(Obj3) == 19 * sizeof(double), "raw input layout");

enum : int { kBarProd = 1, kBarCons = 2, kBarFull = 3, kBarEmpty = 5 };  // named barriers (0 = __syncthreads)

struct WsParams {
    MinkParams m;
    int total;                     // work items = K * n_objs * B
    unsigned long long* work;      // claim counter (zeroed by the host before the launch)
    int* bcnt;                     // [K][L] entries appended to a line's bucket (may exceed bcap: the line then takes the dense path)
    double* bucket;                // [K][L][bcap][2]
    int bcap;
    int skip_a;                    // profiling builds only: producers skip phase A (what do the consumers cost on their own?)
    int o_raw[2], o_cc[2], o_rowf[2], o_vcode[2], o_hdr, o_items, o_zpool, o_stash, o_wtab;
    // (o_mats, o_txs, o_tys, o_queue, o_m0, o_m1, o_qcount of `m` are single copies: consumer- or setup-private)
};

__device__ __forceinline__ void bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ int bar_or(int id, int count, int pred) {
    int out;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\tsetp.ne.s32 q, %1, 0;\n\tbar.red.or.pred p, %2, %3, q;\n\tselp.s32 %0, 1, 0, p;\n\t}"
        : "=r"(out)
        : "r"(pred), "r"(id), "r"(count)
        : "memory");
    return out;
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
// Work-item claim whose issue point the compiler may not move (the answer is consumed a whole phase later).
__device__ __forceinline__ int claim_next(unsigned long long* counter) {
    unsigned long long old;
    asm volatile("atom.global.add.u64 %0, [%1], 1;" : "=l"(old) : "l"(counter) : "memory");
    return static_cast<int>(old);
}
// Bucket-slot request, likewise pinned to its place in the instruction stream.
__device__ __forceinline__ int bucket_slot(int* counter) {
    int old;
    asm volatile("atom.global.add.s32 %0, [%1], 1;" : "=r"(old) : "l"(counter) : "memory");
    return old;
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ double fm(double a, double b, double c) { return __fma_rn(a, b, c); }
__device__ __forceinline__ float fm(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ double mu(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float mu(float a, float b) { return __fmul_rn(a, b); }

// One boundary point from its column constants k = {ce1, ce2, C[3], WZ[3], a, b, d} and row factors f = {V[3], W[3], P, Q}
// (kColStride / kRowStride).  Every operation is an explicit fused multiply-add or multiply, so the producers (which store
// the point) and the consumers (which re-evaluate it) get the same bits.
__device__ __forceinline__ void eval_point(const double* k, const double* f, double* o) {
    const double nn = fm(k[8], f[6], fm(k[9], f[7], k[10]));
    const double inv = fast_rsqrt(nn);
#pragma unroll
    for (int d = 0; d < 3; ++d) o[d] = fm(fm(k[1], f[3 + d], k[5 + d]), inv, fm(k[0], f[d], k[2 + d]));
}

// The four corners of mesh cell (i, j) at once -- the same operations as four eval_point calls (bit-identical results), ordered
// for instruction-level parallelism: the four normalisations first (independent rsqrt chains), then the coordinates one axis
// at a time, so that a consumer thread waits for ONE dependent chain per stage instead of four whole evaluations in a row.
// cc_i = column constants of column i (the next column follows at +kColStride), rf_j = row factors of row j (next: +kRowStride).
__device__ __forceinline__ void eval_cell(const double* cc_i, const double* rf_j, D3 (&c)[2][2]) {  // c[col][row]
    double k01[2][2], inv[2][2], pq[2][2];
#pragma unroll
    for (int r = 0; r < 2; ++r) lds_pair(smem_u32(rf_j + r * kRowStride + 6), pq[r][0], pq[r][1]);
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        double k89[2];
        lds_pair(smem_u32(cc_i + a * kColStride), k01[a][0], k01[a][1]);
        lds_pair(smem_u32(cc_i + a * kColStride + 8), k89[0], k89[1]);
        const double k10 = cc_i[a * kColStride + 10];
#pragma unroll
        for (int r = 0; r < 2; ++r) inv[a][r] = fast_rsqrt(fm(k89[0], pq[r][0], fm(k89[1], pq[r][1], k10)));
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        double o[2][2];
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            const double kC = cc_i[a * kColStride + 2 + d], kW = cc_i[a * kColStride + 5 + d];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const double* f = rf_j + r * kRowStride;
                o[a][r] = fm(fm(k01[a][1], f[3 + d], kW), inv[a][r], fm(k01[a][0], f[d], kC));
            }
        }
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int r = 0; r < 2; ++r) (d == 0 ? c[a][r].x : (d == 1 ? c[a][r].y : c[a][r].z)) = o[a][r];
    }
}

// Height of the sweep line (ix, iy) on face f, evaluated from the shared tables (phase D's rare path: the hit pool was full).
__device__ __noinline__ double face_height(const double* cc, const double* rowf, const double* txs, const double* tys, int n, int Ly, uint32_t f, uint32_t l) {
    const uint32_t ix = l / static_cast<uint32_t>(Ly), iy = l - ix * static_cast<uint32_t>(Ly);
    const uint32_t nc1 = static_cast<uint32_t>(n - 1), ncell = nc1 * nc1;
    const uint32_t cell = (f < ncell) ? f : f - ncell;
    const uint32_t i = cell / nc1, j = cell - i * nc1;
    D3 cn[2][2];
    eval_cell(cc + i * kColStride, rowf + j * kRowStride, cn);
    double z = 0.0;
    vertical_hit_fast(txs[ix + 1], tys[iy + 1], cn[0][0], (f < ncell) ? cn[1][0] : cn[0][1], cn[1][1], z);  // the same test on the same vertices as in phase C
    return z;
}

// line_q (mink3d_common.cuh) without its range test, for surfaces whose vertices are PROVEN to lie within 2^15 line spacings
// of the grid origin (see range_ok below): ok = false only when the coordinate is within 4 * 2^-16 spacings of a grid line.
__device__ __forceinline__ uint32_t line_q_inrange(double x, double inv_d, double bias_q, bool& ok) {
    const uint32_t w = static_cast<uint32_t>(__double2loint(fma(x, inv_d, bias_q)));
    ok = ((w + 5u) & 0xFFF8u) != 0u;
    return w;
}

// NF: grid size n fixed at compile time (0: taken from the parameters).
template <int VAR, int NF, bool DBG>
__global__ void __launch_bounds__(kWsThreads, 2) k_minksweep3d_ws(const WsParams P) {
    constexpr int kDU = (kDeferLines + kCT - 1) / kCT;  // deferred lines per consumer thread
    using Regs = RegSplit<VAR>;
    using Code = uint16_t;  // 2 x 8-bit line codes (both axes <= 127 lines)
    using A = Fast;
    const MinkParams& p = P.m;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = NF ? NF : p.n, M = NF ? NF * NF : p.M;
    const int lpc = NF ? NF / 2 : p.lpc, groups = NF ? kPT / (NF / 2) : p.groups;  // (the host checks that pick_mapping agrees)
    const int tid = threadIdx.x;
    const int per_layer = p.n_objs * p.B;
    const int heavy = p.n_heavy * p.B;
    const int nk = P.total / per_layer;

    double* txs = reinterpret_cast<double*>(smem_raw + p.o_txs);
    double* tys = reinterpret_cast<double*>(smem_raw + p.o_tys);
    int* hdr = reinterpret_cast<int*>(smem_raw + P.o_hdr);      // [2][8]: item, slot, is_arena, k
    int* items = reinterpret_cast<int*>(smem_raw + P.o_items);  // [4][4] ring of claimed work items: item, object, pose (k * B + body), output slot
    for (int i = tid; i < p.Lx + 2; i += kWsThreads) txs[i] = __ldg(p.txs + i);
    for (int i = tid; i < p.Ly + 2; i += kWsThreads) tys[i] = __ldg(p.tys + i);
    // claim + decode by one thread (the integer divisions run once per surface, not once per thread)
    auto claim = [&](int ring, int item) {
        int k = 0, s = 0;
        if (item < P.total) {
            if (item < nk * heavy) {
                k = item / heavy;
                s = item - k * heavy;
            } else {
                const int rem = item - nk * heavy, light = per_layer - heavy;
                k = rem / light;
                s = heavy + (rem - k * light);
            }
        }
        const int oi = s / p.B;
        int* e = items + 4 * ring;
        e[0] = item, e[1] = p.first_obj + oi, e[2] = k * p.B + (s - oi * p.B), e[3] = k * per_layer + s;
    };
    if (DBG && (tid & 31) == 0) {  // hardware warp slot of every warp of the CTA (rows 6g.. producers, 7g.. consumers)
        unsigned wid;
        asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
        const int w = tid >> 5;
        p.dbg[(static_cast<size_t>(gridDim.x) * (w < kPW ? 6 : 7) + blockIdx.x) * 8 + (w < kPW ? w : w - kPW)] = wid;
    }
    if (tid == 0) {
        const int i0 = static_cast<int>(atomicAdd(P.work, 1ULL)), i1 = static_cast<int>(atomicAdd(P.work, 1ULL));
        claim(0, i0), claim(1, i1);
    }
    __syncthreads();

    if (tid < kPT) {
        // =============================== PRODUCERS: inputs -> tables -> boundary points ===============================
        if constexpr (Regs::prod != 0) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(Regs::prod));
        const int grp = tid / lpc;
        const int ln = tid - grp * lpc;
        const bool worker = grp < groups;
        const int lane = tid & 31;

        auto issue_load = [&](int ring, int rb) {
            const int obj_id = items[4 * ring + 1];
            const size_t pose = static_cast<size_t>(items[4 * ring + 2]);
            double* dst = reinterpret_cast<double*>(smem_raw + P.o_raw[rb]);
            const double* gtab = p.tab + static_cast<size_t>(obj_id) * 8 * n;
            for (int i = tid; i < 4 * n; i += kPT) cp_async16(dst + 2 * i, gtab + 2 * i);
            double* ex = dst + 8 * n;
            if (tid < 19) cp_async8(ex + tid, reinterpret_cast<const double*>(p.obj + obj_id) + tid);
            else if (tid >= 32 && tid < 41) cp_async8(ex + 20 + (tid - 32), p.R + pose * 9 + (tid - 32));
            else if (tid >= 64 && tid < 67) cp_async8(ex + 29 + (tid - 64), p.off + pose * 3 + (tid - 64));
            else if (tid >= 96 && tid < 99) cp_async8(ex + 32 + (tid - 96), p.semi + 3 * (p.semi_per_layer ? pose : pose % p.B) + (tid - 96));
        };

        // profiling builds (p.dbg): cycles of thread 0 in [1] waiting for inputs + the consumers, [2] tables, [3] phase A
        long long t_wait = 0, t_setup = 0, t_a = 0, t0 = 0, t_begin = 0;
        int done = 0;
        if (DBG) t_begin = clock64();
        if (items[0] < P.total) issue_load(0, 0);
        for (int it = 0;; ++it) {
            const int b = it & 1;
            const int item = items[4 * (it & 3)];
            const int slot = items[4 * (it & 3) + 3];
            // the item for iteration it + 2 is requested now (global atomic) and decoded behind phase A, when the answer is back:
            // nobody waits for the round trip
            int claimed = 0;
            if (tid == 0) claimed = claim_next(P.work);
            if (DBG) t0 = clock64();
            cp_async_wait_all();
            if (it >= 2) bar_sync(kBarEmpty + b, kWsThreads);  // the consumers are done with buffer b (surface it - 2)
            bar_sync(kBarProd, kPT);                           // raw[b] complete and visible to all producers
            if (DBG) {
                const long long t1 = clock64();
                t_wait += t1 - t0;
                t0 = t1;
            }
            if (item >= P.total) {
                if (tid == 0) hdr[8 * b] = -1;
                bar_arrive(kBarFull + b, kWsThreads);
                break;
            }
            const double* raw = reinterpret_cast<const double*>(smem_raw + P.o_raw[b]);
            const double* ro = raw + 8 * n;  // a b c px py pz R[9] sign ia ib ic | R2[9] @20 | off[3] @29 | semi[3] @32
            double* cc = reinterpret_cast<double*>(smem_raw + P.o_cc[b]);
            double* rowf = reinterpret_cast<double*>(smem_raw + P.o_rowf[b]);
            Code* vcode = reinterpret_cast<Code*>(smem_raw + P.o_vcode[b]);

            // ---- per-surface matrices: Tinv = (R2 diag) R2^T, M1 = Tinv R1, M2 = ((K Tinv) Tinv) R1  (SuperQuadrics.cpp:115,137-140).
            // Every warp evaluates them itself (entry e on lane e < 9, three dependent steps through shuffles, the reference's
            // association order), so no warp waits for another one here.
            double m1e, m2e;
            {
                const int me = lane < 9 ? lane : 0;
                const int i = me / 3, j = me - 3 * i;
                const double* r2 = ro + 20;
                const double md0 = ro[32], md1 = ro[33], md2 = ro[34], msign = ro[15];
                const double t = A::mad(A::mul(r2[3 * i + 2], md2), r2[3 * j + 2], A::mad(A::mul(r2[3 * i + 1], md1), r2[3 * j + 1], A::mul(A::mul(r2[3 * i], md0), r2[3 * j])));
                const double ti0 = __shfl_sync(0xFFFFFFFFu, t, 3 * i), ti1 = __shfl_sync(0xFFFFFFFFu, t, 3 * i + 1), ti2 = __shfl_sync(0xFFFFFFFFu, t, 3 * i + 2);
                const double tj0 = __shfl_sync(0xFFFFFFFFu, t, j), tj1 = __shfl_sync(0xFFFFFFFFu, t, 3 + j), tj2 = __shfl_sync(0xFFFFFFFFu, t, 6 + j);
                const double ktt = A::mad(A::mul(msign, ti2), tj2, A::mad(A::mul(msign, ti1), tj1, A::mul(A::mul(msign, ti0), tj0)));
                const double ki0 = __shfl_sync(0xFFFFFFFFu, ktt, 3 * i), ki1 = __shfl_sync(0xFFFFFFFFu, ktt, 3 * i + 1), ki2 = __shfl_sync(0xFFFFFFFFu, ktt, 3 * i + 2);
                const double r1a = ro[6 + j], r1b = ro[9 + j], r1c = ro[12 + j];
                m1e = A::mad(ti2, r1c, A::mad(ti1, r1b, A::mul(ti0, r1a)));
                m2e = A::mad(ki2, r1c, A::mad(ki1, r1b, A::mul(ki0, r1a)));
            }
            // The 18 row-base coefficients (R1[:,0] a, R1[:,1] b, M1[:,0] / a, M1[:,1] / b, M2[:,0] / a, M2[:,1] / b) and the third
            // columns of M1 / M2, one per lane, into the warp's own 24 doubles of shared memory: the table tasks below read them
            // with broadcast loads instead of every thread forming the products again.
            double* wb = reinterpret_cast<double*>(smem_raw + P.o_wtab) + (tid >> 5) * kWbStride;
            {
                const int d = lane % 3, kind = lane / 3;  // lanes 0..17: kind 0..5, component d; lanes 18..23: third columns
                const int src = lane < 18 ? 3 * d + (kind & 1) : 3 * d + 2;
                const double s1 = __shfl_sync(0xFFFFFFFFu, m1e, src), s2 = __shfl_sync(0xFFFFFFFFu, m2e, src);
                double val = 0.0;
                if (kind == 0) val = ro[6 + 3 * d] * ro[0];
                else if (kind == 1) val = ro[6 + 3 * d + 1] * ro[1];
                else if (kind == 2) val = s1 * ro[16];
                else if (kind == 3) val = s1 * ro[17];
                else if (kind == 4) val = s2 * ro[16];
                else if (kind == 5) val = s2 * ro[17];
                else if (kind == 6) val = s1;
                else val = s2;
                if (lane < 24) wb[lane] = val;
                __syncwarp();
            }
            // ---- factor tables: rows  V_r = R1[:,0] a cw1_r + R1[:,1] b sw1_r,  W_r = M2[:,0] cw2_r / a + M2[:,1] sw2_r / b,
            //      P_r = |U_r|^2, Q_r = 2 U_r . m  with U_r = M1[:,0] cw2_r / a + M1[:,1] sw2_r / b, m = M1[:,2];
            //      columns {ce1, ce2, C = R1[:,2] c se1 + (p - off), WZ = M2[:,2] gz, a = ce2^2, b = ce2 gz, d = gz^2 |m|^2}, gz = se2 / c
            for (int task = tid; task < 3 * n; task += kPT) {
                if (task < n) {
                    const int r = task;
                    const double cw1 = raw[4 * n + r], sw1 = raw[5 * n + r], cw2 = raw[6 * n + r], sw2 = raw[7 * n + r];
                    double* f = rowf + r * kRowStride;
                    double U[3];
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        f[d] = fm(wb[d], cw1, mu(wb[3 + d], sw1));
                        U[d] = fm(wb[6 + d], cw2, mu(wb[9 + d], sw2));
                        f[3 + d] = fm(wb[12 + d], cw2, mu(wb[15 + d], sw2));
                    }
                    f[6] = fm(U[2], U[2], fm(U[1], U[1], mu(U[0], U[0])));
                    f[7] = mu(2.0, fm(U[2], wb[20], fm(U[1], wb[19], mu(U[0], wb[18]))));
                } else {
                    const int c = (task - n) >> 1;
                    double* kc = cc + c * kColStride;
                    if ((task - n) & 1) {
                        const double ce1 = raw[c], Zc = mu(ro[2], raw[n + c]);  // c * se1
                        kc[0] = ce1;
#pragma unroll
                        for (int i = 0; i < 3; ++i) kc[2 + i] = fm(ro[6 + 3 * i + 2], Zc, ro[3 + i] - ro[29 + i]);
                    } else {
                        const double ce2 = raw[2 * n + c], gz = mu(raw[3 * n + c], ro[18]);  // se2 / c
                        const double mm = fm(wb[20], wb[20], fm(wb[19], wb[19], mu(wb[18], wb[18])));
                        kc[1] = ce2;
#pragma unroll
                        for (int i = 0; i < 3; ++i) kc[5 + i] = mu(wb[21 + i], gz);
                        kc[8] = mu(ce2, ce2), kc[9] = mu(ce2, gz), kc[10] = mu(mu(gz, gz), mm);
                        kc[11] = 0.0;
                    }
                }
            }
            if (tid == 0) {
                hdr[8 * b] = item;
                hdr[8 * b + 1] = slot;
                hdr[8 * b + 2] = ro[15] < 0.0 ? 1 : 0;  // arena surface (sign = -1)
                hdr[8 * b + 3] = slot / per_layer;      // layer
                // Every vertex lies within rad = a + b + c + max(body semi-axis) of p - off (|x0 - p| <= |(a, b, c)|, and the
                // offset term is Tinv applied to a unit vector): if that keeps the line coordinate inside +-32000 spacings, phase A
                // drops the per-vertex range test of line_q.  (NaN / inf inputs compare false: checked path.)
                const double rad = ro[0] + ro[1] + ro[2] + fmax(ro[32], fmax(ro[33], ro[34]));
                const double gxm = fma(fabs(ro[3] - ro[29]) + rad, p.inv_dx, fabs(p.bias_x));
                const double gym = fma(fabs(ro[4] - ro[30]) + rad, p.inv_dy, fabs(p.bias_y));
                hdr[8 * b + 4] = (gxm < 32000.0 && gym < 32000.0) ? 1 : 0;
            }
            bar_sync(kBarProd, kPT);
            // ---- inputs of the next surface: requested now, consumed after phase A
            if (items[4 * ((it + 1) & 3)] < P.total) issue_load((it + 1) & 3, b ^ 1);
            if (DBG) {
                const long long t1 = clock64();
                t_setup += t1 - t0;
                t0 = t1;
            }
            // ---- phase A: a thread owns the adjacent rows (r0, r0 + 1) and walks the columns of its group
            auto phase_a = [&](auto in_range_tag) {
                constexpr bool kInRange = decltype(in_range_tag)::value;
                using R2 = Real2<double>;
                double* mx = reinterpret_cast<double*>(p.mesh) + static_cast<size_t>(slot) * 3 * M;
                const int npass = n / (2 * lpc);
                const uint32_t cc_step = static_cast<uint32_t>(groups * kColStride * sizeof(double));
                const uint32_t v_step = static_cast<uint32_t>(groups * n * sizeof(Code));
                const size_t g_step = static_cast<size_t>(groups) * n;
                const uint32_t lx2 = static_cast<uint32_t>(p.Lx) * 0x00010001u, ly2 = static_cast<uint32_t>(p.Ly) * 0x00010001u;
                for (int h = 0; h < npass; ++h) {
                    const int r0 = 2 * (ln + h * lpc);
                    double F[2][8];  // V[3] W[3] P Q of the two rows
                    {
                        const uint32_t fa = smem_u32(rowf + r0 * kRowStride);
#pragma unroll
                        for (int h2 = 0; h2 < 2; ++h2)
#pragma unroll
                            for (int i = 0; i < 4; ++i) lds_pair(fa + (h2 * kRowStride + 2 * i) * sizeof(double), F[h2][2 * i], F[h2][2 * i + 1]);
                    }
                    uint32_t ca = smem_u32(cc) + static_cast<uint32_t>(grp * kColStride * sizeof(double));
                    uint32_t va = smem_u32(vcode) + static_cast<uint32_t>((grp * n + r0) * sizeof(Code));
                    double* gx = mx + grp * n + r0;
                    // Software pipeline over the columns (the loop is bound by the latency of its shared-memory loads, not by their
                    // number): the three norm constants {a, b, d} of a column are fetched one iteration ahead, so an iteration opens
                    // with its two rsqrt chains, and the other eight constants arrive underneath them.  Same operations on the same
                    // values as eval_point (the consumers re-evaluate corners with it).
                    double ka, kb, kd;
                    lds_pair(ca + 8 * sizeof(double), ka, kb);
                    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(kd) : "r"(ca + 10 * static_cast<uint32_t>(sizeof(double))));
                    for (int c = grp; c < n; c += groups, ca += cc_step, va += v_step, gx += g_step) {
                        const double nn0 = fm(ka, F[0][6], fm(kb, F[0][7], kd)), nn1 = fm(ka, F[1][6], fm(kb, F[1][7], kd));
                        double kk[8];
#pragma unroll
                        for (int i = 0; i < 4; ++i) lds_pair(ca + i * 2 * sizeof(double), kk[2 * i], kk[2 * i + 1]);
                        const double inv0 = fast_rsqrt(nn0), inv1 = fast_rsqrt(nn1);
                        // (the last iteration reads the 24 bytes behind the table: still inside this CTA's shared memory, never used)
                        lds_pair(ca + cc_step + 8 * sizeof(double), ka, kb);
                        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(kd) : "r"(ca + cc_step + 10 * static_cast<uint32_t>(sizeof(double))));
                        double o[2][3];
#pragma unroll
                        for (int d = 0; d < 3; ++d) {
                            o[0][d] = fm(fm(kk[1], F[0][3 + d], kk[5 + d]), inv0, fm(kk[0], F[0][d], kk[2 + d]));
                            o[1][d] = fm(fm(kk[1], F[1][3 + d], kk[5 + d]), inv1, fm(kk[0], F[1][d], kk[2 + d]));
                        }
                        // (evict-first stores, __stcs, were measured on this kernel: no difference, burst or sustained)
                        *reinterpret_cast<R2*>(gx) = R2{o[0][0], o[1][0]};
                        *reinterpret_cast<R2*>(gx + M) = R2{o[0][1], o[1][1]};
                        *reinterpret_cast<R2*>(gx + 2 * static_cast<size_t>(M)) = R2{o[0][2], o[1][2]};
                        bool k0, k1, k2, k3;
                        uint32_t wx0, wy0, wx1, wy1;
                        if constexpr (kInRange) {
                            wx0 = line_q_inrange(o[0][0], p.inv_dx, p.bias_xq, k0), wy0 = line_q_inrange(o[0][1], p.inv_dy, p.bias_yq, k1);
                            wx1 = line_q_inrange(o[1][0], p.inv_dx, p.bias_xq, k2), wy1 = line_q_inrange(o[1][1], p.inv_dy, p.bias_yq, k3);
                        } else {
                            wx0 = line_q(o[0][0], p.inv_dx, p.bias_xq, k0), wy0 = line_q(o[0][1], p.inv_dy, p.bias_yq, k1);
                            wx1 = line_q(o[1][0], p.inv_dx, p.bias_xq, k2), wy1 = line_q(o[1][1], p.inv_dy, p.bias_yq, k3);
                        }
                        if (k0 && k1 && k2 && k3) {
                            const uint32_t X = __vimin_s16x2_relu(__byte_perm(wx0, wx1, 0x7632), lx2);
                            const uint32_t Y = __vimin_s16x2_relu(__byte_perm(wy0, wy1, 0x7632), ly2);
                            asm volatile("st.shared.u32 [%0], %1;" ::"r"(va), "r"((X | (Y << 8)) << 1) : "memory");
                        } else {  // rare: a coordinate within rounding distance of a sweep line -> exact search
                            const uint32_t cx0 = line_code(o[0][0], txs, p.Lx, p.inv_dx, p.bias_x);
                            const uint32_t cy0 = line_code(o[0][1], tys, p.Ly, p.inv_dy, p.bias_y);
                            const uint32_t cx1 = line_code(o[1][0], txs, p.Lx, p.inv_dx, p.bias_x);
                            const uint32_t cy1 = line_code(o[1][1], tys, p.Ly, p.inv_dy, p.bias_y);
                            sts_codes(va, static_cast<uint16_t>(cx0 | (cy0 << 8)), static_cast<uint16_t>(cx1 | (cy1 << 8)));
                        }
                    }
                }
            };
            if (DBG && P.skip_a) {
                for (int i = tid; i < M + 8; i += kPT) vcode[i] = 0;  // no vertex near a sweep line: the scan keeps nothing
            } else if (worker) {
                if (hdr[8 * b + 4]) phase_a(std::true_type{});
                else phase_a(std::false_type{});
            }
            if (tid == 0) claim((it + 2) & 3, claimed);
            __syncwarp();
            bar_arrive(kBarFull + b, kWsThreads);  // codes and tables of this surface are in shared memory (bar.arrive orders them)
            if (DBG) t_a += clock64() - t0, ++done;
        }
        cp_async_wait_all();
        if (DBG && tid == 0) {
            long long* d = p.dbg + static_cast<size_t>(blockIdx.x) * 8;
            d[1] = t_wait, d[2] = t_setup, d[3] = t_a, d[7] = done;
        }
        return;
    }

    // ======================================= CONSUMERS: scan -> cell tests -> intervals =======================================
    if constexpr (Regs::cons != 0) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(Regs::cons));
    const int ctid = tid - kPT;
    const int lane = ctid & 31, warp = ctid >> 5;
    uint16_t* cellq = reinterpret_cast<uint16_t*>(smem_raw + p.o_queue);  // [kWsCellCap] kept cells (cell = i * (n-1) + j)
    uint16_t* heavyq = cellq + kWsCellCap;                                 // [kWsCellCap] of which: many candidates
    double* zpool = reinterpret_cast<double*>(smem_raw + P.o_zpool);       // [kZCap] heights of the hits found in phase C
    double2* stash = reinterpret_cast<double2*>(smem_raw + P.o_stash);     // [kCT][4] intervals whose bucket slot is still on its way
    // Bucket appends of the PREVIOUS surface: the slot of an interval comes from a global atomic, whose round trip (~2k cycles
    // under the write stream) nobody waits for -- the interval is parked in shared memory, the slot stays in a register, and the
    // append is completed when the next surface arrives.
    int ppos[kDU];
#pragma unroll
    for (int u = 0; u < kDU; ++u) ppos[u] = -1;
    size_t pbase = 0;
    auto flush_pending = [&]() {
#pragma unroll
        for (int u = 0; u < kDU; ++u) {
            if (ppos[u] >= 0 && ppos[u] < P.bcap) {
                const size_t gl = pbase + ctid + u * kCT;
                *reinterpret_cast<double2*>(P.bucket + (gl * P.bcap + ppos[u]) * 2) = stash[ctid * kDU + u];
            }
            ppos[u] = -1;
        }
    };
    uint32_t* m0 = reinterpret_cast<uint32_t*>(smem_raw + p.o_m0);         // [L] lowest hit face per line
    uint32_t* m1 = reinterpret_cast<uint32_t*>(smem_raw + p.o_m1);         // [L] second lowest
    int* qcount = reinterpret_cast<int*>(smem_raw + p.o_qcount);           // [2] heavy cells; [4..] cells per warp segment
    int* zcount = qcount;
    int* hcount = qcount + 2;
    int* segcnt = qcount + 4;
    int* scan_next = qcount + 12;  // next chunk of mesh rows / next batch of kept cells: the consumer warps claim their work, because the
    int* cell_next = qcount + 13;  // warp that shares its scheduler with the busiest producer warps runs at less than half the others' pace
    const int nc1 = n - 1;
    const int ncell = nc1 * nc1;
    const uint32_t nc1_magic = 0xFFFFFFFFu / static_cast<uint32_t>(nc1) + 1u;
    constexpr int kUPL = 4;
    constexpr int kSegStride = kWsCellCap / kCW;
    const int kSegCap = p.seg_cap < kSegStride ? p.seg_cap : kSegStride;
    uint16_t* seg = cellq + warp * kSegStride;
#ifndef HRMGPU_WS_KLIGHT
#define HRMGPU_WS_KLIGHT 16
#endif
    constexpr uint32_t kLight = HRMGPU_WS_KLIGHT;  // cells with more candidates than this get a whole warp

    for (int i = ctid; i < p.L; i += kCT) m0[i] = kNoFace, m1[i] = kNoFace;  // (afterwards phase D leaves every entry reset)
    if (ctid == 0) *scan_next = 0;
    bar_sync(kBarCons, kCT);
    long long c_wait = 0, c_sweep = 0, c_cells = 0, c_d = 0, c0 = 0;  // profiling builds: [4] waiting for the producers, [5] scan + cell tests, [6] phase D
    long long w_swar = 0, w_comp = 0, w_bar = 0, w_t = 0;            // per warp: SWAR tests, compaction, waiting for the other consumer warps
    long long w_first = 0, w_nchunk = 0;                             // ... of which the first chunk of every surface; chunks taken
    bool w_is_first = false;
    for (int it = 0;; ++it) {
        const int b = it & 1;
        if (DBG) c0 = clock64();
        bar_sync(kBarFull + b, kWsThreads);
        if (DBG) {
            const long long c1 = clock64();
            c_wait += c1 - c0;
            c0 = c1;
        }
        const int item = hdr[8 * b];
        if (item < 0) break;
        const int slot = hdr[8 * b + 1];
        const bool is_arena = hdr[8 * b + 2] != 0;
        const int k_layer = hdr[8 * b + 3];
        const double* cc = reinterpret_cast<const double*>(smem_raw + P.o_cc[b]);
        const double* rowf = reinterpret_cast<const double*>(smem_raw + P.o_rowf[b]);
        const Code* vcode = reinterpret_cast<const Code*>(smem_raw + P.o_vcode[b]);
        auto code_at = [&](int q) -> uint32_t { return __byte_perm(static_cast<uint32_t>(vcode[q]), 0u, 0x4140); };
        auto hit = [&](uint32_t ix, uint32_t iy, const D3& t0, const D3& t1, const D3& t2, double& z) -> bool {
            return vertical_hit_fast(txs[ix + 1], tys[iy + 1], t0, t1, t2, z);
        };

        // (no barrier here: m0 / m1 were reset by phase D of the previous surface, scan_next when its scan ended, and the barrier
        //  that ends the scan orders these two counters before the tests)
        if (ctid == 0) *hcount = 0, *zcount = 0;
        if (DBG) w_is_first = true;
        // A hit enters the line's ordered two-slot table keyed by its FACE INDEX (the reference keeps the first two hits in face
        // order, SURVEY finding 2); the low bits of the key say where its height is kept for phase D.
        // (Also measured: the producer warps scanning the first 64 mesh rows of the surface they have just finished, out of line,
        //  into their own list -- 1.7 % faster than the same build without it, but the added code costs that build 4.6 %.)
        // (Measured and rejected, kernel ms at 125 layers against 6.24 for this form: per-thread static pool slots instead of the
        //  shared counter 6.32; cell batches / scan chunks claimed one item ahead 6.94 / 6.36; ranks from ballots instead of the
        //  list-segment atomic 6.56 -- the consumer code sits at its register budget, and each of these moved spills into the loops.)
        auto record_hit = [&](uint32_t l, uint32_t face, double z) {
            const int zi = atomicAdd(zcount, 1);
            uint32_t key = (face << kZBits) | kZNone;
            if (zi < kZCap) zpool[zi] = z, key = (face << kZBits) | static_cast<uint32_t>(zi);
            insert_two_smallest(&m0[l], &m1[l], key);
        };

        // B1: cells whose inclusive xy bounding box (over the 4 corners) may contain a sweep line -- the SWAR test of mink3d.cu
        // (per axis the set of lines inside the box is empty <=> all four codes are equal and even).
        // Result bits of a unit: cell 0 -> bit 7, cell 2 -> 8, cell 1 -> 23, cell 3 -> 24 (where the SWAR form leaves them); four
        // units share one 32-bit word, unit uu shifted left by 2 * uu.
        // (a, bq: the two vertex rows; sa*, sb*: the same rows advanced by one vertex)
        auto swar4 = [](uint2 a, uint32_t sax, uint32_t say, uint2 bq, uint32_t sbx, uint32_t sby) -> uint32_t {
            const uint32_t dx = (a.x ^ bq.x) | (a.x ^ sax) | (a.x ^ sbx) | (a.x & 0x01010101u);
            const uint32_t dy = (a.y ^ bq.y) | (a.y ^ say) | (a.y ^ sby) | (a.y & 0x01010101u);
            const uint32_t nx = (((dx & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | dx), ny = (((dy & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | dy);
            const uint32_t fx = nx & (nx >> 8) & 0x00800080u, fy = ny & (ny >> 8) & 0x00800080u;
            return fx + (fy << 1);
        };
        auto cell_sets = [&](int q, uint32_t& loA, uint32_t& hiA, uint32_t& loB, uint32_t& hiB, uint32_t& nA, uint32_t& nB) {
            const uint32_t c00 = code_at(q), c11 = code_at(q + n + 1), c10 = code_at(q + n), c01 = code_at(q + 1);
            const uint32_t mnA = __vimin3_u16x2(c00, c11, c10), mxA = __vimax3_u16x2(c00, c11, c10);
            const uint32_t mnB = __vimin3_u16x2(c00, c11, c01), mxB = __vimax3_u16x2(c00, c11, c01);
            loA = (mnA >> 1) & 0x7FFF7FFFu, hiA = ((mxA + 0x00010001u) >> 1) & 0x7FFF7FFFu;
            loB = (mnB >> 1) & 0x7FFF7FFFu, hiB = ((mxB + 0x00010001u) >> 1) & 0x7FFF7FFFu;
            const uint32_t dA = __vsub2(hiA, loA), dB = __vsub2(hiB, loB);
            nA = (dA & 0xFFFFu) * (dA >> 16), nB = (dB & 0xFFFFu) * (dB >> 16);
        };
        auto test_lines = [&](uint32_t face, uint32_t lo2, uint32_t hi2, const D3& t0, const D3& t1, const D3& t2) {
            for (uint32_t ix = lo2 & 0xFFFFu; ix < (hi2 & 0xFFFFu); ++ix)
                for (uint32_t iy = lo2 >> 16; iy < (hi2 >> 16); ++iy) {
                    double z;
                    if (hit(ix, iy, t0, t1, t2, z)) record_hit(ix * p.Ly + iy, face, z);
                }
        };

        // Scan.  The code is straight-line on purpose (mask arithmetic instead of selects the compiler would turn into jumps):
        // the consumers are bound by instruction issue, so what counts is the number of instructions per cell and that the loads
        // and logic of independent cells interleave.  A lane owns the 4 cells j0 = 4 * lane .. j0 + 3 of every mesh row (n <= 128,
        // n % 4 == 0: one 8-byte load per vertex row), a warp a band of consecutive mesh rows; the lower vertex row of one step is
        // the upper row of the next, and the lane's cell mask is loop-invariant.  (1) all SWAR tests of a chunk (kWR words x 4
        // rows per lane) back to back; (2) ONE prefix sum over the lanes' kept-cell counts, then every lane writes its own cells.
        // What does not fit into the warp's list segment stays in w[] for the next round (after the tests of this round).
#ifndef HRMGPU_WS_KWR
#define HRMGPU_WS_KWR 3
#endif
        constexpr int kWR = HRMGPU_WS_KWR;  // words per lane and chunk (x 4 rows each): short on purpose, the unrolled body is ~30 instructions per row
        const int nchunk = (nc1 + kWR * kUPL - 1) / (kWR * kUPL);
        const int j0 = 4 * lane;
        const int jl = min(j0, n - 4);  // (lanes past the end of a row load its last unit and keep nothing)
        const int vcells = nc1 - j0;    // cells of this lane's unit inside a mesh row
        // cells 0 .. v-1 of a unit are bits 7, 23, 8, 24
        const uint32_t keep = (0x00000080u & (0u - static_cast<uint32_t>(vcells > 0))) | (0x00800000u & (0u - static_cast<uint32_t>(vcells > 1))) |
                              (0x00000100u & (0u - static_cast<uint32_t>(vcells > 2))) | (0x01000000u & (0u - static_cast<uint32_t>(vcells > 3)));
        auto chunk_row0 = [&](int chunk) { return chunk * (kWR * kUPL); };
        int chunk = 0;
        bool loaded = false;
        uint32_t w[kWR];
#pragma unroll
        for (int r = 0; r < kWR; ++r) w[r] = 0u;
        for (;;) {
            bool parked = false;
            if (lane == 0) segcnt[warp] = 0;
            __syncwarp();
            for (;;) {
                if (DBG) w_t = clock64();
                if (!loaded) {
                    int claimed_chunk = 0;
                    if (lane == 0) claimed_chunk = atomicAdd(scan_next, 1);
                    chunk = __shfl_sync(0xFFFFFFFFu, claimed_chunk, 0);
                    if (chunk >= nchunk) break;
                    const int row0 = chunk_row0(chunk);
                    const int band_end = nc1;  // (rows past the mesh in the last chunk keep nothing)
                    int row = min(row0, nc1 - 1);
                    uint2 ra = *reinterpret_cast<const uint2*>(vcode + row * n + jl);
                    uint32_t sax = __funnelshift_r(ra.x, ra.y, 16), say = __funnelshift_r(ra.y, static_cast<uint32_t>(vcode[row * n + jl + 4]), 16);
#pragma unroll
                    for (int r = 0; r < kWR; ++r) {
                        uint32_t acc = 0u;
#pragma unroll
                        for (int uu = 0; uu < kUPL; ++uu) {
                            const int i = row0 + r * kUPL + uu;  // mesh row (vertex rows i, i + 1)
                            const int ib = min(i + 1, n - 1);
                            const uint2 rb = *reinterpret_cast<const uint2*>(vcode + ib * n + jl);
                            const uint32_t sbx = __funnelshift_r(rb.x, rb.y, 16), sby = __funnelshift_r(rb.y, static_cast<uint32_t>(vcode[ib * n + jl + 4]), 16);
                            const uint32_t on = 0u - static_cast<uint32_t>(i < band_end);
                            acc += (swar4(ra, sax, say, rb, sbx, sby) & keep & on) << (2 * uu);
                            ra = rb, sax = sbx, say = sby;
                        }
                        w[r] = acc;
                    }
                    loaded = true;
                }
                if (DBG) {
                    const long long t = clock64();
                    w_swar += t - w_t;
                    if (w_is_first) w_first += t - w_t, w_is_first = false;
                    w_t = t, ++w_nchunk;
                }
                // Kept cells are sparse (~2 %): a lane that has some takes its slots in the warp's list segment with ONE shared-memory
                // atomic (the order inside the segment is irrelevant: hits are keyed by face index) -- no prefix sum over the lanes.
                int mine = 0;
#pragma unroll
                for (int r = 0; r < kWR; ++r) mine += __popc(w[r]);
                bool left = false;
                if (mine != 0) {
                    int pos = atomicAdd(&segcnt[warp], mine);
                    const int row0 = chunk_row0(chunk);
#pragma unroll
                    for (int r = 0; r < kWR; ++r) {
                        while (w[r] != 0u && pos < kSegCap) {
                            const int bit = __ffs(w[r]) - 1;
                            w[r] &= w[r] - 1u;
                            // bits 7.. hold cells 0 / 2, bits 23.. cells 1 / 3 of the word's four rows
                            const int hi = bit >= 23 ? 1 : 0, pb = bit - (hi ? 23 : 7);
                            seg[pos] = static_cast<uint16_t>((row0 + r * kUPL + (pb >> 1)) * nc1 + j0 + (((pb & 1) << 1) | hi));
                            ++pos;
                        }
                        left = left || (w[r] != 0u);
                    }
                }
                if (DBG) w_comp += clock64() - w_t;
                if (__any_sync(0xFFFFFFFFu, left)) {  // segment full: the tests come first, the rest of w[] follows in the next round
                    parked = true;
                    break;
                }
                loaded = false;
            }
            __syncwarp();
            if (lane == 0) segcnt[warp] = min(segcnt[warp], kSegCap);  // (a full segment: the counter ran past the capacity)
            if (ctid == 0) *cell_next = 0;
            __syncwarp();
            if (DBG) w_t = clock64();
            const int more = bar_or(kBarCons, kCT, parked ? 1 : 0);
            if (!more && ctid == 0) *scan_next = 0;  // every warp is done claiming chunks of this surface
            if (DBG) w_bar += clock64() - w_t;
            if (DBG) {
                const long long c1 = clock64();
                c_sweep += c1 - c0;
                c0 = c1;
            }
            // C: one kept cell per thread and pass
            int sc[kCW + 1];
            sc[0] = 0;
#pragma unroll
            for (int w = 0; w < kCW; ++w) sc[w + 1] = sc[w] + segcnt[w];
            const int ncq = sc[kCW];
            for (;;) {
                int base = 0;
                if (lane == 0) base = atomicAdd(cell_next, 32);
                base = __shfl_sync(0xFFFFFFFFu, base, 0);
                if (base >= ncq) break;
                const int ci = base + lane;
                if (ci < ncq) {
                    int w = 0, off_w = 0;
#pragma unroll
                    for (int ww = 1; ww < kCW; ++ww) w += (ci >= sc[ww]) ? 1 : 0, off_w = (ci >= sc[ww]) ? sc[ww] : off_w;
                    const uint32_t cell = cellq[w * kSegStride + (ci - off_w)];
                    const int i = static_cast<int>(__umulhi(cell, nc1_magic));
                    const int j = static_cast<int>(cell) - i * nc1;
                    const int q = i * n + j;
                    uint32_t loA, hiA, loB, hiB, nA, nB;
                    cell_sets(q, loA, hiA, loB, hiB, nA, nB);
                    if (nA + nB > kLight) {
                        heavyq[atomicAdd(hcount, 1)] = static_cast<uint16_t>(cell);
                    } else if (nA + nB != 0u) {
                        D3 cn[2][2];  // [column offset][row offset]: vertex (i + a, j + r) = point q + a * n + r
                        eval_cell(cc + i * kColStride, rowf + j * kRowStride, cn);
                        // faces (q, q+n, q+n+1) and (q, q+1, q+n+1): ONE copy of the line loop / triangle test in the code
                        // (the consumers' per-surface instruction footprint has to stay inside the 32 KB L1.5 instruction cache)
#pragma unroll 1
                        for (int fb = 0; fb < 2; ++fb) {
                            if ((fb ? nB : nA) == 0u) continue;
                            const D3 t1 = fb ? cn[0][1] : cn[1][0];
                            test_lines(cell + (fb ? static_cast<uint32_t>(ncell) : 0u), fb ? loB : loA, fb ? hiB : hiA, cn[0][0], t1, cn[1][1]);
                        }
                    }
                }
            }
            __syncwarp();
            bar_sync(kBarCons, kCT);
            const int nh = *hcount;
            if (nh != 0) {
                for (int h = warp; h < nh; h += kCW) {
                    const uint32_t cell = heavyq[h];
                    const int i = static_cast<int>(__umulhi(cell, nc1_magic));
                    const int j = static_cast<int>(cell) - i * nc1;
                    const int q = i * n + j;
                    uint32_t loA, hiA, loB, hiB, nA, nB;
                    cell_sets(q, loA, hiA, loB, hiB, nA, nB);
                    D3 cn[2][2];
                    eval_cell(cc + i * kColStride, rowf + j * kRowStride, cn);
                    const D3 c00 = cn[0][0], c10 = cn[1][0], c11 = cn[1][1], c01 = cn[0][1];
                    for (uint32_t idx = lane; idx < nA + nB; idx += 32) {
                        const bool second = idx >= nA;
                        const uint32_t k2 = second ? idx - nA : idx;
                        const uint32_t lo2 = second ? loB : loA, hi2 = second ? hiB : hiA;
                        const uint32_t wy = (hi2 >> 16) - (lo2 >> 16);
                        const uint32_t dxl = k2 / wy;
                        const uint32_t ix = (lo2 & 0xFFFFu) + dxl, iy = (lo2 >> 16) + (k2 - dxl * wy);
                        double z;
                        if (hit(ix, iy, c00, second ? c01 : c10, c11, z)) record_hit(ix * p.Ly + iy, cell + (second ? static_cast<uint32_t>(ncell) : 0u), z);
                    }
                }
                __syncwarp();
                bar_sync(kBarCons, kCT);
            }
            if (DBG) {
                const long long c1 = clock64();
                c_cells += c1 - c0;
                c0 = c1;
            }
            if (!more) break;
            if (ctid == 0) *hcount = 0;
            bar_sync(kBarCons, kCT);
        }

        if (DBG) {
            const long long c1 = clock64();
            c_sweep += c1 - c0;
            c0 = c1;
        }
        flush_pending();  // (the previous surface's bucket slots were requested a whole scan + test phase ago)
        // ---- phase D: intervals (FreeSpace3D.cpp:44-49 / 61-64) ----
        double* lo = p.lo + static_cast<size_t>(slot) * p.L;
        double* up = p.up + static_cast<size_t>(slot) * p.L;
        unsigned single = 0;
        auto face_z = [&](uint32_t key, int l) -> double {
            if ((key & kZNone) != kZNone) return zpool[key & kZNone];  // found in phase C
            return face_height(cc, rowf, txs, tys, n, p.Ly, key >> kZBits, static_cast<uint32_t>(l));  // (pool full: evaluate the face again)
        };
        // blocks of kDeferLines lines (one block when L <= 512): bucket appends deferred, see ppos.  Staged, so that the
        // shared-memory round trips of a thread's lines overlap instead of following one another: (1) the hit-table entries of all
        // its lines (reset for the next surface at once), (2) the hit heights, (3) intervals, stores and bucket-slot requests.
#pragma unroll 1
        for (int lb = 0; lb < p.L; lb += kDeferLines) {
            if (lb > 0) flush_pending();  // (more than 512 sweep lines: the slots of the previous block are awaited here)
            pbase = static_cast<size_t>(k_layer) * p.L + lb;
            uint32_t e0[kDU], e1[kDU];
#pragma unroll
            for (int u = 0; u < kDU; ++u) {
                const int l = lb + ctid + u * kCT;
                e0[u] = e1[u] = kNoFace;
                if (l < p.L && ctid + u * kCT < kDeferLines) {
                    e0[u] = m0[l], e1[u] = m1[l];
                    m0[l] = kNoFace, m1[l] = kNoFace;
                }
            }
            double z0[kDU], z1[kDU];
#pragma unroll
            for (int u = 0; u < kDU; ++u) {
                z0[u] = z1[u] = 0.0;
                if (e1[u] != kNoFace) z0[u] = face_z(e0[u], lb + ctid + u * kCT), z1[u] = face_z(e1[u], lb + ctid + u * kCT);
            }
#pragma unroll
            for (int u = 0; u < kDU; ++u) {
                const int l = lb + ctid + u * kCT;
                if (l < p.L && ctid + u * kCT < kDeferLines) {
                    double a = is_arena ? p.zlow : __longlong_as_double(0x7FF8000000000000LL);
                    double bb = is_arena ? p.zup : __longlong_as_double(0x7FF8000000000000LL);
                    if (e1[u] != kNoFace) {
                        const double zl = fmin(z0[u], z1[u]), zu = fmax(z0[u], z1[u]);
                        a = is_arena ? fmin(p.zlow, zl) : zl;
                        bb = is_arena ? fmax(p.zup, zu) : zu;
                        if (!is_arena && P.bcnt != nullptr && !isnan(zl) && !isnan(zu)) {  // an obstacle interval: also into the line's bucket
                            stash[ctid * kDU + u] = make_double2(a, bb);
                            ppos[u] = bucket_slot(P.bcnt + pbase + ctid + u * kCT);  // not read before flush_pending()
                        }
                    } else if (e0[u] != kNoFace) {
                        ++single;  // exactly one hit: reference reads out of bounds; defined here as "no interval"
                    }
                    lo[l] = a;
                    up[l] = bb;
                }
            }
        }
        if (single) atomicAdd(p.counters, static_cast<unsigned long long>(single));
        __syncwarp();
        bar_arrive(kBarEmpty + b, kWsThreads);  // (no fence: nothing the consumers wrote is read by the producers)
        if (DBG) c_d += clock64() - c0;
    }
    flush_pending();
    if (DBG && ctid == 0) {
        long long* d = p.dbg + static_cast<size_t>(blockIdx.x) * 8;
        d[4] = c_wait, d[5] = c_sweep, d[6] = c_d;
    }
    if (DBG && lane == 0) {  // second block of rows: [warp] SWAR, [4 + warp] compaction of every consumer warp ... and warp 0's barrier wait behind them
        long long* d = p.dbg + (static_cast<size_t>(gridDim.x) * 2 + blockIdx.x) * 8;
        d[warp] = w_swar, d[4 + warp] = w_comp;
        long long* d3 = p.dbg + (static_cast<size_t>(gridDim.x) * 3 + blockIdx.x) * 8;
        d3[warp] = w_first, d3[4 + warp] = w_nchunk;
        if (warp == 0) {
            p.dbg[(static_cast<size_t>(gridDim.x) * 4 + blockIdx.x) * 8] = w_bar;
            p.dbg[static_cast<size_t>(blockIdx.x) * 8] = c_cells;
        }
    }
}

// Shared-memory layout; returns the total.
static size_t ws_layout(WsParams& P, int n, int Lx, int Ly) {
    MinkParams& p = P.m;
    auto al = [](size_t v) { return (v + 15) & ~size_t(15); };
    size_t b = 0;
    for (int i = 0; i < 2; ++i) P.o_raw[i] = static_cast<int>(b), b += al(sizeof(double) * (8 * n + kRawExtra));
    for (int i = 0; i < 2; ++i) P.o_cc[i] = static_cast<int>(b), b += al(sizeof(double) * kColStride * n);
    for (int i = 0; i < 2; ++i) P.o_rowf[i] = static_cast<int>(b), b += al(sizeof(double) * kRowStride * n);
    for (int i = 0; i < 2; ++i) P.o_vcode[i] = static_cast<int>(b), b += al(2 * (static_cast<size_t>(n) * n + 8));
    p.o_mats = 0;
    P.o_wtab = static_cast<int>(b), b += sizeof(double) * kWbStride * kPW;
    p.o_txs = static_cast<int>(b), b += al(sizeof(double) * (Lx + 2));
    p.o_tys = static_cast<int>(b), b += al(sizeof(double) * (Ly + 2));
    p.o_queue = static_cast<int>(b), b += 2 * sizeof(uint16_t) * kWsCellCap;
    P.o_zpool = static_cast<int>(b), b += sizeof(double) * kZCap;
    P.o_stash = static_cast<int>(b), b += sizeof(double2) * kCT * ((kDeferLines + kCT - 1) / kCT);
    const size_t L = static_cast<size_t>(Lx) * Ly;
    p.o_m0 = static_cast<int>(b), b += al(sizeof(uint32_t) * L);
    p.o_m1 = static_cast<int>(b), b += al(sizeof(uint32_t) * L);
    p.o_qcount = static_cast<int>(b), b += 16 + 4 * 8 + 16;
    b = al(b);
    P.o_hdr = static_cast<int>(b), b += 2 * 8 * sizeof(int);
    P.o_items = static_cast<int>(b), b += 16 * sizeof(int);
    return al(b);
}

// Whether a build of this shape runs on the warp-specialised kernel (otherwise mink3d.cu's one-surface kernel).
bool ws_supported(const Ctx* c, int precision, bool do_sweep, bool reload, int lpc, int groups, int paired) {
    if (c->force_classic || !do_sweep || reload || precision != HRMGPU_F64) return false;
    if (!paired || c->n < 32 || c->n > 128 || (c->n & 3) != 0 || lpc * groups > kPT) return false;  // (scan: 8-byte rows, one unit per lane)
    if (c->Lx > 127 || c->Ly > 127) return false;  // 2 x 8-bit line codes
    return true;
}

template <int VAR, int NF, bool DBG>
static int launch_ws_t(Ctx* c, const WsParams& P, int grid, size_t smem) {
    auto kern = k_minksweep3d_ws<VAR, NF, DBG>;
    HRMGPU_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    kern<<<grid, kWsThreads, smem, c->stream>>>(P);
    HRMGPU_CUDA(c, cudaGetLastError());
    c->launches += 1;
    return HRMGPU_OK;
}

constexpr int kFixedN = 100;  // the grid size with its own instantiation (BASELINE.json configs[4]: 10^4 boundary points)

// Called by launch_minksweep3d (mink3d.cu) with the common parameters filled in.
int launch_minksweep3d_ws(Ctx* c, const MinkParams& base, int precision, long long items, int* d_bcnt, double* d_bucket, int bcap) {
    (void)precision;  // HRMGPU_F64 (ws_supported)
    WsParams P;
    P.m = base;
    P.total = static_cast<int>(items);
    P.work = c->counters.p + 2;
    P.bcnt = d_bcnt;
    P.bucket = d_bucket;
    P.bcap = bcap;
    P.skip_a = std::getenv("HRMGPU_WS_SKIPA") ? 1 : 0;
    const size_t smem = ws_layout(P, c->n, c->Lx, c->Ly);
    if (smem > 113 * 1024) return -1000;  // does not fit twice per SM: the caller falls back to the one-surface kernel
    HRMGPU_CUDA(c, cudaMemsetAsync(c->counters.p + 2, 0, sizeof(unsigned long long), c->stream));
    const int grid = static_cast<int>(std::min<long long>(items, 2LL * c->num_sms));
    const bool fixed = !c->ws_generic && c->n == kFixedN && base.lpc == kFixedN / 2 && base.groups == kPT / (kFixedN / 2);
    const int v = c->ws_variant;  // experiment hook (HRMGPU_WS_VARIANT): 0 = registers 80 / 80, 1 = 72 / 96 (producers / consumers)
    if (P.m.dbg) {  // profiling builds (tools/phase_times.py): per-role cycle counters
        if (fixed) return v == 0 ? launch_ws_t<0, kFixedN, true>(c, P, grid, smem) : launch_ws_t<1, kFixedN, true>(c, P, grid, smem);
        return v == 0 ? launch_ws_t<0, 0, true>(c, P, grid, smem) : launch_ws_t<1, 0, true>(c, P, grid, smem);
    }
    if (fixed) return v == 0 ? launch_ws_t<0, kFixedN, false>(c, P, grid, smem) : launch_ws_t<1, kFixedN, false>(c, P, grid, smem);
    return v == 0 ? launch_ws_t<0, 0, false>(c, P, grid, smem) : launch_ws_t<1, 0, false>(c, P, grid, smem);
}

template <class K>
static void touch_ws(K k) {
    cudaFuncAttributes a;
    if (cudaFuncGetAttributes(&a, k) != cudaSuccess) cudaGetLastError();
}
void preload_mink3d_ws() {
    touch_ws(k_minksweep3d_ws<0, 0, false>), touch_ws(k_minksweep3d_ws<1, 0, false>);
    touch_ws(k_minksweep3d_ws<0, kFixedN, false>), touch_ws(k_minksweep3d_ws<1, kFixedN, false>);
}

}  // namespace hrmgpu
