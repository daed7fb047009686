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
#include "mink3d_common.cuh"

namespace hrmgpu {

constexpr int kPW = 8;                 // producer warps
constexpr int kCW = 4;                 // consumer warps
constexpr int kPT = 32 * kPW, kCT = 32 * kCW, kWsThreads = kPT + kCT;
// Registers: the CTA is launched with 80 per thread (2 CTAs x 384 threads per SM); the 8 producer warps give 8 each to the
// 4 consumer warps (setmaxnreg), whose triangle tests + re-evaluation need more than the streaming phase A.
// Variant REBAL = false keeps 80 / 80.
constexpr int kProdRegs = 72, kConsRegs = 96;
constexpr int kRowStride = 10;         // per-row factors V[3] U[3] W[3] + pad
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
    int o_raw[2], o_cc[2], o_rowf[2], o_vcode[2], o_hdr, o_items;
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
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ double fm(double a, double b, double c) { return __fma_rn(a, b, c); }
__device__ __forceinline__ float fm(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ double mu(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float mu(float a, float b) { return __fmul_rn(a, b); }

// One boundary point from its column constants k = {ce1, ce2, C[3], UZ[3], WZ[3]} and row factors V, U, W (see kColStride in
// mink3d_common.cuh).  Every operation is an explicit fused multiply-add or multiply, so the producers (which store the
// point) and the consumers (which re-evaluate it) get the same bits.
template <class Real>
__device__ __forceinline__ void eval_point(const Real* k, const Real* V, const Real* U, const Real* W, Real* o) {
    const Real ux = fm(k[1], U[0], k[5]), uy = fm(k[1], U[1], k[6]), uz = fm(k[1], U[2], k[7]);
    const Real nn = fm(uz, uz, fm(uy, uy, mu(ux, ux)));
    Real inv;
    if constexpr (sizeof(Real) == 4)
        inv = rsqrtf(nn);
    else
        inv = fast_rsqrt(nn);
#pragma unroll
    for (int d = 0; d < 3; ++d) o[d] = fm(fm(k[1], W[d], k[8 + d]), inv, fm(k[0], V[d], k[2 + d]));
}

template <int MODE, bool REBAL>
__global__ void __launch_bounds__(kWsThreads, 2) k_minksweep3d_ws(const WsParams P) {
    using Real = typename RealOf<MODE>::type;
    using Code = uint16_t;  // 2 x 8-bit line codes (both axes <= 127 lines)
    using A = Fast;
    const MinkParams& p = P.m;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = p.n, M = p.M;
    const int tid = threadIdx.x;
    const int per_layer = p.n_objs * p.B;
    const int heavy = p.n_heavy * p.B;
    const int nk = P.total / per_layer;

    double* txs = reinterpret_cast<double*>(smem_raw + p.o_txs);
    double* tys = reinterpret_cast<double*>(smem_raw + p.o_tys);
    int* hdr = reinterpret_cast<int*>(smem_raw + P.o_hdr);      // [2][8]: item, slot, is_arena, k
    int* items = reinterpret_cast<int*>(smem_raw + P.o_items);  // [4] claimed work items, ring
    for (int i = tid; i < p.Lx + 2; i += kWsThreads) txs[i] = __ldg(p.txs + i);
    for (int i = tid; i < p.Ly + 2; i += kWsThreads) tys[i] = __ldg(p.tys + i);
    if (tid == 0) {
        items[0] = static_cast<int>(atomicAdd(P.work, 1ULL));
        items[1] = static_cast<int>(atomicAdd(P.work, 1ULL));
    }
    __syncthreads();

    auto decode = [&](int item, int& k, int& s) {
        if (item < nk * heavy) {
            k = item / heavy;
            s = item - k * heavy;
        } else {
            const int rem = item - nk * heavy, light = per_layer - heavy;
            k = rem / light;
            s = heavy + (rem - k * light);
        }
    };

    if (tid < kPT) {
        // =============================== PRODUCERS: inputs -> tables -> boundary points ===============================
        if constexpr (REBAL) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kProdRegs));
        const int grp = tid / p.lpc;
        const int ln = tid - grp * p.lpc;
        const bool worker = grp < p.groups;
        const int lane = tid & 31;
        double* mats = reinterpret_cast<double*>(smem_raw + p.o_mats);  // M1[9] M2[9] + scratch T[9] KTT[9]

        auto issue_load = [&](int item, int rb) {
            int k, s;
            decode(item, k, s);
            const int oi = s / p.B, b = s - oi * p.B, obj_id = p.first_obj + oi;
            double* dst = reinterpret_cast<double*>(smem_raw + P.o_raw[rb]);
            const double* gtab = p.tab + static_cast<size_t>(obj_id) * 8 * n;
            for (int i = tid; i < 4 * n; i += kPT) cp_async16(dst + 2 * i, gtab + 2 * i);
            double* ex = dst + 8 * n;
            const size_t pose = static_cast<size_t>(k) * p.B + b;
            if (tid < 19) cp_async8(ex + tid, reinterpret_cast<const double*>(p.obj + obj_id) + tid);
            else if (tid >= 32 && tid < 41) cp_async8(ex + 20 + (tid - 32), p.R + pose * 9 + (tid - 32));
            else if (tid >= 64 && tid < 67) cp_async8(ex + 29 + (tid - 64), p.off + pose * 3 + (tid - 64));
            else if (tid >= 96 && tid < 99) cp_async8(ex + 32 + (tid - 96), p.semi + 3 * b + (tid - 96));
        };

        if (items[0] < P.total) issue_load(items[0], 0);
        for (int it = 0;; ++it) {
            const int b = it & 1;
            const int item = items[it & 3];
            if (tid == 0) items[(it + 2) & 3] = static_cast<int>(atomicAdd(P.work, 1ULL));  // consumed two iterations from now
            cp_async_wait_all();
            if (it >= 2) bar_sync(kBarEmpty + b, kWsThreads);  // the consumers are done with buffer b (surface it - 2)
            bar_sync(kBarProd, kPT);                           // raw[b] complete and visible to all producers
            if (item >= P.total) {
                if (tid == 0) hdr[8 * b] = -1;
                __threadfence_block();
                bar_arrive(kBarFull + b, kWsThreads);
                break;
            }
            int k, s;
            decode(item, k, s);
            const int slot = k * per_layer + s;
            const double* raw = reinterpret_cast<const double*>(smem_raw + P.o_raw[b]);
            const double* ro = raw + 8 * n;  // a b c px py pz R[9] sign ia ib ic | R2[9] @20 | off[3] @29 | semi[3] @32
            Real* cc = reinterpret_cast<Real*>(smem_raw + P.o_cc[b]);
            Real* rowf = reinterpret_cast<Real*>(smem_raw + P.o_rowf[b]);
            Code* vcode = reinterpret_cast<Code*>(smem_raw + P.o_vcode[b]);

            // ---- per-surface matrices: Tinv = (R2 diag) R2^T, M1 = Tinv R1, M2 = ((K Tinv) Tinv) R1  (SuperQuadrics.cpp:115,137-140)
            if (tid < 32) {
                const int me = lane < 9 ? lane : 0;
                const int i = me / 3, j = me - 3 * i;
                const double* r2 = ro + 20;
                const double mi0 = r2[3 * i], mi1 = r2[3 * i + 1], mi2 = r2[3 * i + 2];
                const double mj0 = r2[3 * j], mj1 = r2[3 * j + 1], mj2 = r2[3 * j + 2];
                const double md0 = ro[32], md1 = ro[33], md2 = ro[34], msign = ro[15];
                const double r1a = ro[6 + j], r1b = ro[9 + j], r1c = ro[12 + j];
                double* Ts = mats + 18;
                const double t = A::mad(A::mul(mi2, md2), mj2, A::mad(A::mul(mi1, md1), mj1, A::mul(A::mul(mi0, md0), mj0)));
                if (lane < 9) Ts[me] = t;
                __syncwarp();
                const double ktt = A::mad(A::mul(msign, Ts[3 * i + 2]), Ts[6 + j],
                                          A::mad(A::mul(msign, Ts[3 * i + 1]), Ts[3 + j], A::mul(A::mul(msign, Ts[3 * i]), Ts[j])));
                if (lane < 9) Ts[9 + me] = ktt;
                __syncwarp();
                if (lane < 9) {
                    mats[me] = A::mad(Ts[3 * i + 2], r1c, A::mad(Ts[3 * i + 1], r1b, A::mul(Ts[3 * i], r1a)));
                    mats[9 + me] = A::mad(Ts[9 + 3 * i + 2], r1c, A::mad(Ts[9 + 3 * i + 1], r1b, A::mul(Ts[9 + 3 * i], r1a)));
                }
            }
            bar_sync(kBarProd, kPT);
            // ---- factor tables: rows  V_r = R1[:,0] a cw1_r + R1[:,1] b sw1_r,  U_r = M1[:,0] cw2_r / a + M1[:,1] sw2_r / b,  W_r likewise (M2);
            //      columns {ce1, ce2, C = R1[:,2] c se1 + (p - off), UZ = M1[:,2] se2 / c, WZ = M2[:,2] se2 / c}
            for (int task = tid; task < 3 * n; task += kPT) {
                if (task < n) {
                    const int r = task;
                    const Real cw1 = Real(raw[4 * n + r]), sw1 = Real(raw[5 * n + r]), cw2 = Real(raw[6 * n + r]), sw2 = Real(raw[7 * n + r]);
                    Real* f = rowf + r * kRowStride;
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        const Real ra = Real(ro[6 + 3 * d] * ro[0]), rb = Real(ro[6 + 3 * d + 1] * ro[1]);
                        const Real ua = Real(mats[3 * d] * ro[16]), ub = Real(mats[3 * d + 1] * ro[17]);
                        const Real wa = Real(mats[9 + 3 * d] * ro[16]), wb = Real(mats[9 + 3 * d + 1] * ro[17]);
                        f[d] = fm(ra, cw1, mu(rb, sw1));
                        f[3 + d] = fm(ua, cw2, mu(ub, sw2));
                        f[6 + d] = fm(wa, cw2, mu(wb, sw2));
                    }
                    f[9] = Real(0);
                } else {
                    const int c = (task - n) >> 1;
                    Real* kc = cc + c * kColStride;
                    if ((task - n) & 1) {
                        const Real ce1 = Real(raw[c]), Zc = mu(Real(ro[2]), Real(raw[n + c]));  // c * se1
                        kc[0] = ce1;
#pragma unroll
                        for (int i = 0; i < 3; ++i) kc[2 + i] = fm(Real(ro[6 + 3 * i + 2]), Zc, Real(ro[3 + i] - ro[29 + i]));
                    } else {
                        const Real ce2 = Real(raw[2 * n + c]), gz = mu(Real(raw[3 * n + c]), Real(ro[18]));  // se2 / c
                        kc[1] = ce2;
#pragma unroll
                        for (int i = 0; i < 3; ++i) kc[5 + i] = mu(Real(mats[3 * i + 2]), gz), kc[8 + i] = mu(Real(mats[9 + 3 * i + 2]), gz);
                        kc[11] = Real(0);
                    }
                }
            }
            if (tid == 0) {
                hdr[8 * b] = item;
                hdr[8 * b + 1] = slot;
                hdr[8 * b + 2] = ro[15] < 0.0 ? 1 : 0;  // arena surface (sign = -1)
                hdr[8 * b + 3] = k;
            }
            bar_sync(kBarProd, kPT);
            // ---- inputs of the next surface: requested now, consumed after phase A
            {
                const int next = items[(it + 1) & 3];
                if (next < P.total) issue_load(next, b ^ 1);
            }
            // ---- phase A: a thread owns the adjacent rows (r0, r0 + 1) and walks the columns of its group
            if (worker) {
                using R2 = Real2<Real>;
                Real* mx = reinterpret_cast<Real*>(p.mesh) + static_cast<size_t>(slot) * 3 * M;
                const int npass = n / (2 * p.lpc);
                const uint32_t cc_step = static_cast<uint32_t>(p.groups * kColStride * sizeof(Real));
                const uint32_t v_step = static_cast<uint32_t>(p.groups * n * sizeof(Code));
                const size_t g_step = static_cast<size_t>(p.groups) * n;
                const uint32_t lx2 = static_cast<uint32_t>(p.Lx) * 0x00010001u, ly2 = static_cast<uint32_t>(p.Ly) * 0x00010001u;
                for (int h = 0; h < npass; ++h) {
                    const int r0 = 2 * (ln + h * p.lpc);
                    Real F[2][9];  // V[3] U[3] W[3] of the two rows
                    {
                        const uint32_t fa = smem_u32(rowf + r0 * kRowStride);
#pragma unroll
                        for (int h2 = 0; h2 < 2; ++h2) {
#pragma unroll
                            for (int i = 0; i < 4; ++i) lds_pair(fa + (h2 * kRowStride + 2 * i) * sizeof(Real), F[h2][2 * i], F[h2][2 * i + 1]);
                            F[h2][8] = rowf[(r0 + h2) * kRowStride + 8];
                        }
                    }
                    uint32_t ca = smem_u32(cc) + static_cast<uint32_t>(grp * kColStride * sizeof(Real));
                    uint32_t va = smem_u32(vcode) + static_cast<uint32_t>((grp * n + r0) * sizeof(Code));
                    Real* gx = mx + grp * n + r0;
                    for (int c = grp; c < n; c += p.groups, ca += cc_step, va += v_step, gx += g_step) {
                        Real kk[12];
#pragma unroll
                        for (int i = 0; i < 6; ++i) lds_pair(ca + i * 2 * sizeof(Real), kk[2 * i], kk[2 * i + 1]);
                        Real o[2][3];
                        eval_point<Real>(kk, F[0], F[0] + 3, F[0] + 6, o[0]);
                        eval_point<Real>(kk, F[1], F[1] + 3, F[1] + 6, o[1]);
                        *reinterpret_cast<R2*>(gx) = R2{o[0][0], o[1][0]};
                        *reinterpret_cast<R2*>(gx + M) = R2{o[0][1], o[1][1]};
                        *reinterpret_cast<R2*>(gx + 2 * static_cast<size_t>(M)) = R2{o[0][2], o[1][2]};
                        bool k0, k1, k2, k3;
                        const uint32_t wx0 = line_q(static_cast<double>(o[0][0]), p.inv_dx, p.bias_xq, k0);
                        const uint32_t wy0 = line_q(static_cast<double>(o[0][1]), p.inv_dy, p.bias_yq, k1);
                        const uint32_t wx1 = line_q(static_cast<double>(o[1][0]), p.inv_dx, p.bias_xq, k2);
                        const uint32_t wy1 = line_q(static_cast<double>(o[1][1]), p.inv_dy, p.bias_yq, k3);
                        if (k0 && k1 && k2 && k3) {
                            const uint32_t X = __vimin_s16x2_relu(__byte_perm(wx0, wx1, 0x7632), lx2);
                            const uint32_t Y = __vimin_s16x2_relu(__byte_perm(wy0, wy1, 0x7632), ly2);
                            asm volatile("st.shared.u32 [%0], %1;" ::"r"(va), "r"((X | (Y << 8)) << 1) : "memory");
                        } else {  // rare: a coordinate within rounding distance of a sweep line -> exact search
                            const uint32_t cx0 = line_code(static_cast<double>(o[0][0]), txs, p.Lx, p.inv_dx, p.bias_x);
                            const uint32_t cy0 = line_code(static_cast<double>(o[0][1]), tys, p.Ly, p.inv_dy, p.bias_y);
                            const uint32_t cx1 = line_code(static_cast<double>(o[1][0]), txs, p.Lx, p.inv_dx, p.bias_x);
                            const uint32_t cy1 = line_code(static_cast<double>(o[1][1]), tys, p.Ly, p.inv_dy, p.bias_y);
                            sts_codes(va, static_cast<uint16_t>(cx0 | (cy0 << 8)), static_cast<uint16_t>(cx1 | (cy1 << 8)));
                        }
                    }
                }
            }
            __syncwarp();
            __threadfence_block();
            bar_arrive(kBarFull + b, kWsThreads);  // codes and tables of this surface are in shared memory
        }
        cp_async_wait_all();
        return;
    }

    // ======================================= CONSUMERS: scan -> cell tests -> intervals =======================================
    if constexpr (REBAL) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kConsRegs));
    const int ctid = tid - kPT;
    const int lane = ctid & 31, warp = ctid >> 5;
    uint16_t* cellq = reinterpret_cast<uint16_t*>(smem_raw + p.o_queue);  // [kCellCap] kept cells (cell = i * (n-1) + j)
    uint16_t* heavyq = cellq + kCellCap;                                   // [kCellCap] of which: many candidates
    uint32_t* m0 = reinterpret_cast<uint32_t*>(smem_raw + p.o_m0);         // [L] lowest hit face per line
    uint32_t* m1 = reinterpret_cast<uint32_t*>(smem_raw + p.o_m1);         // [L] second lowest
    int* qcount = reinterpret_cast<int*>(smem_raw + p.o_qcount);           // [2] heavy cells; [4..] cells per warp segment
    int* hcount = qcount + 2;
    int* segcnt = qcount + 4;
    const int nc1 = n - 1;
    const int ncell = nc1 * nc1;
    const int UPR = (nc1 + 3) >> 2;  // units (4 consecutive cells of a mesh row) per mesh row
    const int nunits = nc1 * UPR;
    const uint32_t upr_magic = (UPR > 1) ? 0xFFFFFFFFu / static_cast<uint32_t>(UPR) + 1u : 0u;
    const uint32_t nc1_magic = 0xFFFFFFFFu / static_cast<uint32_t>(nc1) + 1u;
    const bool vec_ok = (n & 3) == 0;
    constexpr int kUPL = 4;
    constexpr int kSegStride = kCellCap / kCW;
    const int kSegCap = p.seg_cap < kSegStride ? p.seg_cap : kSegStride;
    const unsigned lt_mask = (1u << lane) - 1u;
    uint16_t* seg = cellq + warp * kSegStride;
    constexpr uint32_t kLight = 16;  // cells with more candidates than this get a whole warp

    for (int it = 0;; ++it) {
        const int b = it & 1;
        bar_sync(kBarFull + b, kWsThreads);
        const int item = hdr[8 * b];
        if (item < 0) break;
        const int slot = hdr[8 * b + 1];
        const bool is_arena = hdr[8 * b + 2] != 0;
        const int k_layer = hdr[8 * b + 3];
        const Real* cc = reinterpret_cast<const Real*>(smem_raw + P.o_cc[b]);
        const Real* rowf = reinterpret_cast<const Real*>(smem_raw + P.o_rowf[b]);
        const Code* vcode = reinterpret_cast<const Code*>(smem_raw + P.o_vcode[b]);
        auto code_at = [&](int q) -> uint32_t { return __byte_perm(static_cast<uint32_t>(vcode[q]), 0u, 0x4140); };
        // vertex (i, j) of the parameter grid (point index i * n + j): the producers' arithmetic again
        auto vertex = [&](int i, int j) -> D3 {
            Real kk[12], F[9], o[3];
            const uint32_t ca = smem_u32(cc + i * kColStride), fa = smem_u32(rowf + j * kRowStride);
#pragma unroll
            for (int t = 0; t < 6; ++t) lds_pair(ca + t * 2 * sizeof(Real), kk[2 * t], kk[2 * t + 1]);
#pragma unroll
            for (int t = 0; t < 4; ++t) lds_pair(fa + t * 2 * sizeof(Real), F[2 * t], F[2 * t + 1]);
            F[8] = rowf[j * kRowStride + 8];
            eval_point<Real>(kk, F, F + 3, F + 6, o);
            return {static_cast<double>(o[0]), static_cast<double>(o[1]), static_cast<double>(o[2])};
        };
        auto hit = [&](uint32_t ix, uint32_t iy, const D3& t0, const D3& t1, const D3& t2, double& z) -> bool {
            return vertical_hit_fast(txs[ix + 1], tys[iy + 1], t0, t1, t2, z);
        };

        for (int i = ctid; i < p.L; i += kCT) m0[i] = kNoFace, m1[i] = kNoFace;
        if (ctid == 0) *hcount = 0;
        bar_sync(kBarCons, kCT);

        // B1: cells of unit u whose inclusive xy bounding box may contain a sweep line (see mink3d.cu for the SWAR form)
        auto unit_cells = [&](int unit_raw, bool fast) -> uint32_t {
            const int unit = min(unit_raw, nunits - 1);
            const int i = (UPR > 1) ? static_cast<int>(__umulhi(static_cast<uint32_t>(unit), upr_magic)) : unit;
            const int j0 = 4 * (unit - i * UPR);
            const int q = i * n + j0;
            uint32_t g;
            if (fast) {
                const uint2 a = *reinterpret_cast<const uint2*>(vcode + q);
                const uint2 bq = *reinterpret_cast<const uint2*>(vcode + q + n);
                const uint32_t a4 = static_cast<uint32_t>(vcode[q + 4]), b4 = static_cast<uint32_t>(vcode[q + n + 4]);
                const uint32_t sax = __funnelshift_r(a.x, a.y, 16), say = __funnelshift_r(a.y, a4, 16);
                const uint32_t sbx = __funnelshift_r(bq.x, bq.y, 16), sby = __funnelshift_r(bq.y, b4, 16);
                const uint32_t dx = (a.x ^ bq.x) | (a.x ^ sax) | (a.x ^ sbx) | (a.x & 0x01010101u);
                const uint32_t dy = (a.y ^ bq.y) | (a.y ^ say) | (a.y ^ sby) | (a.y & 0x01010101u);
                const uint32_t nx = (((dx & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | dx), ny = (((dy & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | dy);
                const uint32_t fx = nx & (nx >> 8) & 0x00800080u, fy = ny & (ny >> 8) & 0x00800080u;
                g = fx + (fy << 1);
            } else {
                uint32_t r0[5], r1[5];
#pragma unroll
                for (int e = 0; e < 5; ++e) r0[e] = code_at(q + e), r1[e] = code_at(q + n + e);
                g = 0;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const uint32_t d = (r0[e] ^ r1[e]) | (r0[e] ^ r0[e + 1]) | (r0[e] ^ r1[e + 1]) | (r0[e] & 0x00010001u);
                    if ((d & 0xFFFFu) != 0u && (d >> 16) != 0u) g |= 1u << ((e & 1 ? 23 : 7) + (e >> 1));
                }
            }
            const int v = nc1 - j0;
            const uint32_t keep = (v >= 4) ? 0x01800180u : ((v == 3) ? 0x00800180u : ((v == 2) ? 0x00800080u : 0x00000080u));
            return (unit_raw < nunits) ? (g & keep) : 0u;
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
                    if (hit(ix, iy, t0, t1, t2, z)) {
                        const uint32_t l = ix * p.Ly + iy;
                        insert_two_smallest(&m0[l], &m1[l], face);
                    }
                }
        };

        int bw = warp;
        bool batch_loaded = false;
        uint32_t w_bits = 0;
        for (;;) {
            int wcnt = 0;
            bool parked = false;
            while (bw * (32 * kUPL) < nunits) {
                if (!batch_loaded) {
                    w_bits = 0u;
                    if (vec_ok) {
#pragma unroll
                        for (int uu = 0; uu < kUPL; ++uu) w_bits += unit_cells((bw * kUPL + uu) * 32 + lane, true) << (2 * uu);
                    } else {
#pragma unroll
                        for (int uu = 0; uu < kUPL; ++uu) w_bits += unit_cells((bw * kUPL + uu) * 32 + lane, false) << (2 * uu);
                    }
                    batch_loaded = true;
                }
                unsigned any = __ballot_sync(0xFFFFFFFFu, w_bits != 0u);
                while (any != 0u) {
                    const int pos = wcnt + __popc(any & lt_mask);
                    if (w_bits != 0u && pos < kSegCap) {
                        const int bit = __ffs(w_bits) - 1;
                        w_bits &= w_bits - 1u;
                        const int hi = bit >= 23 ? 1 : 0, pb = bit - (hi ? 23 : 7);
                        const int unit = (bw * kUPL + (pb >> 1)) * 32 + lane;
                        const int i = (UPR > 1) ? static_cast<int>(__umulhi(static_cast<uint32_t>(unit), upr_magic)) : unit;
                        const int j = 4 * (unit - i * UPR) + (((pb & 1) << 1) | hi);
                        seg[pos] = static_cast<uint16_t>(i * nc1 + j);
                    }
                    wcnt += __popc(any);
                    if (wcnt >= kSegCap) {
                        wcnt = kSegCap;
                        break;
                    }
                    any = __ballot_sync(0xFFFFFFFFu, w_bits != 0u);
                }
                __syncwarp();
                if (__any_sync(0xFFFFFFFFu, w_bits != 0u)) {
                    parked = true;
                    break;
                }
                batch_loaded = false;
                bw += kCW;
            }
            if (lane == 0) segcnt[warp] = wcnt;
            __syncwarp();
            const int more = bar_or(kBarCons, kCT, parked ? 1 : 0);
            // C: one kept cell per thread and pass
            int sc[kCW + 1];
            sc[0] = 0;
#pragma unroll
            for (int w = 0; w < kCW; ++w) sc[w + 1] = sc[w] + segcnt[w];
            const int ncq = sc[kCW];
            for (int base = 0; base < ncq; base += kCT) {
                const int ci = base + ctid;
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
                        const D3 c00 = vertex(i, j), c11 = vertex(i + 1, j + 1);
                        if (nA != 0u) test_lines(cell, loA, hiA, c00, vertex(i + 1, j), c11);                                  // (q, q+n, q+n+1)
                        if (nB != 0u) test_lines(cell + static_cast<uint32_t>(ncell), loB, hiB, c00, vertex(i, j + 1), c11);  // (q, q+1, q+n+1)
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
                    const D3 c00 = vertex(i, j), c10 = vertex(i + 1, j), c11 = vertex(i + 1, j + 1), c01 = vertex(i, j + 1);
                    for (uint32_t idx = lane; idx < nA + nB; idx += 32) {
                        const bool second = idx >= nA;
                        const uint32_t k2 = second ? idx - nA : idx;
                        const uint32_t lo2 = second ? loB : loA, hi2 = second ? hiB : hiA;
                        const uint32_t wy = (hi2 >> 16) - (lo2 >> 16);
                        const uint32_t dxl = k2 / wy;
                        const uint32_t ix = (lo2 & 0xFFFFu) + dxl, iy = (lo2 >> 16) + (k2 - dxl * wy);
                        double z;
                        if (hit(ix, iy, c00, second ? c01 : c10, c11, z)) {
                            const uint32_t l = ix * p.Ly + iy;
                            insert_two_smallest(&m0[l], &m1[l], cell + (second ? static_cast<uint32_t>(ncell) : 0u));
                        }
                    }
                }
                __syncwarp();
                bar_sync(kBarCons, kCT);
            }
            if (!more) break;
            if (ctid == 0) *hcount = 0;
            bar_sync(kBarCons, kCT);
        }

        // ---- phase D: intervals (FreeSpace3D.cpp:44-49 / 61-64) ----
        double* lo = p.lo + static_cast<size_t>(slot) * p.L;
        double* up = p.up + static_cast<size_t>(slot) * p.L;
        unsigned single = 0;
        auto face_z = [&](uint32_t f, uint32_t ix, uint32_t iy) -> double {
            const uint32_t cell = (f < static_cast<uint32_t>(ncell)) ? f : f - static_cast<uint32_t>(ncell);
            const int i = static_cast<int>(__umulhi(cell, nc1_magic));
            const int j = static_cast<int>(cell) - i * nc1;
            const D3 t0 = vertex(i, j), t2 = vertex(i + 1, j + 1);
            const D3 t1 = (f < static_cast<uint32_t>(ncell)) ? vertex(i + 1, j) : vertex(i, j + 1);
            double z = 0.0;
            hit(ix, iy, t0, t1, t2, z);  // the same test on the same vertices as in phase C
            return z;
        };
        for (int l = ctid; l < p.L; l += kCT) {
            const uint32_t e0 = m0[l], e1 = m1[l];
            double a = is_arena ? p.zlow : __longlong_as_double(0x7FF8000000000000LL);
            double bb = is_arena ? p.zup : __longlong_as_double(0x7FF8000000000000LL);
            if (e1 != kNoFace) {
                const uint32_t ix = static_cast<uint32_t>(l) / static_cast<uint32_t>(p.Ly), iy = static_cast<uint32_t>(l) - ix * static_cast<uint32_t>(p.Ly);
                const double z0 = face_z(e0, ix, iy), z1 = face_z(e1, ix, iy);
                const double zl = fmin(z0, z1), zu = fmax(z0, z1);
                a = is_arena ? fmin(p.zlow, zl) : zl;
                bb = is_arena ? fmax(p.zup, zu) : zu;
                if (!is_arena && P.bcnt && !isnan(zl) && !isnan(zu)) {  // compact copy for the free-segment pass
                    const size_t gl = static_cast<size_t>(k_layer) * p.L + l;
                    const int pos = atomicAdd(P.bcnt + gl, 1);
                    if (pos < P.bcap) *reinterpret_cast<double2*>(P.bucket + (gl * P.bcap + pos) * 2) = make_double2(zl, zu);
                }
            } else if (e0 != kNoFace) {
                ++single;  // exactly one hit: reference reads out of bounds; defined here as "no interval"
            }
            lo[l] = a;
            up[l] = bb;
        }
        if (single) atomicAdd(p.counters, static_cast<unsigned long long>(single));
        __syncwarp();
        __threadfence_block();
        bar_arrive(kBarEmpty + b, kWsThreads);
    }
}

// Shared-memory layout; returns the total.
static size_t ws_layout(WsParams& P, int MODE, int n, int Lx, int Ly) {
    MinkParams& p = P.m;
    const size_t real = (MODE == HRMGPU_F32) ? 4 : 8;
    auto al = [](size_t v) { return (v + 15) & ~size_t(15); };
    size_t b = 0;
    for (int i = 0; i < 2; ++i) P.o_raw[i] = static_cast<int>(b), b += al(sizeof(double) * (8 * n + kRawExtra));
    for (int i = 0; i < 2; ++i) P.o_cc[i] = static_cast<int>(b), b += al(real * kColStride * n);
    for (int i = 0; i < 2; ++i) P.o_rowf[i] = static_cast<int>(b), b += al(real * kRowStride * n);
    for (int i = 0; i < 2; ++i) P.o_vcode[i] = static_cast<int>(b), b += al(2 * (static_cast<size_t>(n) * n + 8));
    p.o_mats = static_cast<int>(b), b += sizeof(double) * 36;
    p.o_txs = static_cast<int>(b), b += al(sizeof(double) * (Lx + 2));
    p.o_tys = static_cast<int>(b), b += al(sizeof(double) * (Ly + 2));
    p.o_queue = static_cast<int>(b), b += 2 * sizeof(uint16_t) * kCellCap;
    const size_t L = static_cast<size_t>(Lx) * Ly;
    p.o_m0 = static_cast<int>(b), b += al(sizeof(uint32_t) * L);
    p.o_m1 = static_cast<int>(b), b += al(sizeof(uint32_t) * L);
    p.o_qcount = static_cast<int>(b), b += 16 + 4 * kCW + 16;
    b = al(b);
    P.o_hdr = static_cast<int>(b), b += 2 * 8 * sizeof(int);
    P.o_items = static_cast<int>(b), b += 4 * sizeof(int);
    return al(b);
}

// Whether a build of this shape runs on the warp-specialised kernel (otherwise mink3d.cu's one-surface kernel).
bool ws_supported(const Ctx* c, int precision, bool do_sweep, bool reload, int lpc, int groups, int paired) {
    if (c->force_classic || !do_sweep || reload || precision == HRMGPU_F64_STRICT) return false;
    if (!paired || c->n < 32 || lpc * groups > kPT) return false;
    if (c->Lx > 127 || c->Ly > 127) return false;  // 2 x 8-bit line codes
    return true;
}

template <int MODE, bool REBAL>
static int launch_ws_t(Ctx* c, const WsParams& P, int grid, size_t smem) {
    auto kern = k_minksweep3d_ws<MODE, REBAL>;
    HRMGPU_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    kern<<<grid, kWsThreads, smem, c->stream>>>(P);
    HRMGPU_CUDA(c, cudaGetLastError());
    c->launches += 1;
    return HRMGPU_OK;
}

// Called by launch_minksweep3d (mink3d.cu) with the common parameters filled in.
int launch_minksweep3d_ws(Ctx* c, const MinkParams& base, int precision, long long items, int* d_bcnt, double* d_bucket, int bcap) {
    WsParams P;
    P.m = base;
    P.total = static_cast<int>(items);
    P.work = c->counters.p + 2;
    P.bcnt = d_bcnt;
    P.bucket = d_bucket;
    P.bcap = bcap;
    const size_t smem = ws_layout(P, precision, c->n, c->Lx, c->Ly);
    if (smem > 113 * 1024) return -1000;  // does not fit twice per SM: the caller falls back to the one-surface kernel
    HRMGPU_CUDA(c, cudaMemsetAsync(c->counters.p + 2, 0, sizeof(unsigned long long), c->stream));
    const int grid = static_cast<int>(std::min<long long>(items, 2LL * c->num_sms));
    if (precision == HRMGPU_F32) return launch_ws_t<HRMGPU_F32, false>(c, P, grid, smem);
    return c->ws_rebalance ? launch_ws_t<HRMGPU_F64, true>(c, P, grid, smem) : launch_ws_t<HRMGPU_F64, false>(c, P, grid, smem);
}

void preload_mink3d_ws() {
    cudaFuncAttributes a;
    if (cudaFuncGetAttributes(&a, k_minksweep3d_ws<HRMGPU_F64, false>) != cudaSuccess) cudaGetLastError();
    if (cudaFuncGetAttributes(&a, k_minksweep3d_ws<HRMGPU_F64, true>) != cudaSuccess) cudaGetLastError();
    if (cudaFuncGetAttributes(&a, k_minksweep3d_ws<HRMGPU_F32, false>) != cudaSuccess) cudaGetLastError();
}

}  // namespace hrmgpu
