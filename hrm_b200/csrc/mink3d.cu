// K1+K2 fused: closed-form Minkowski boundary of every (layer, object, body) surface and the vertical
// sweep-line intersection against its implicit triangle mesh, one CTA per surface.
//
// Replaces (reference, per surface and per sweep plane, all single-threaded FP64):
//   SuperQuadrics::getOriginShape / getMinkSum3D      src/geometry/SuperQuadrics.cpp:46-143
//   MultiBodyTree3D::minkSum (link-centre subtraction) src/datastructure/MultiBodyTree3D.cpp:91-105
//   getMeshFromParamSurface (implicit here)            src/geometry/MeshGenerator.cpp:105-131
//   FreeSpace3D::computeIntersectionInterval           src/datastructure/FreeSpace3D.cpp:29-68
//   intersectVerticalLineMesh3D / intersectLineTriangle3D  src/geometry/LineIntersection.cpp:56-171
//
// Thread mapping: the CTA's 256 threads form G groups of LPC lanes (LPC chosen on the host: a whole eta-column per group
// where possible, so that a group stores complete 32-byte sectors).  A group owns whole eta-columns c of the parameter
// grid; its lanes stride over the omega index r.  Point index k = r + c*n (omega fastest, SuperQuadrics.cpp:63-72), so
// a group's stores are contiguous.
//
// Phases inside one CTA:
//   A  boundary points.  All transcendental factors are separable (8 host-built tables per object); in the fast modes
//      every term is split into a row factor (kept in registers) and a column factor (shared table): 22 FP64 operations
//      per point.  Points are stored planar (x[M] y[M] z[M]); each vertex is classified against the sweep grid into a
//      packed "line code" (2*lo+eq per axis, lo = first grid line >= coordinate) kept in shared memory.
//   B  per mesh cell a SWAR test on the 4 corner codes keeps the cells whose inclusive xy bounding box can contain a sweep
//      line (LineIntersection.cpp:77-101); kept cells go to per-warp lists in shared memory.
//   C  a thread per kept cell: exact line sets of the two faces, triangle test per (face, line) candidate, hits recorded
//      with an ordered two-slot atomicMin on the FACE INDEX so that the two LOWEST face indices per line survive (the
//      reference's "first two hits in face order" rule, SURVEY.md finding 2).
//   D  per line the two surviving faces are evaluated again for their z and the [lo, up] interval is written.
#include "mink3d_common.cuh"

#include <cfloat>
#include <cmath>

namespace hrmgpu {

template <int MODE, bool SWEEP, bool CODE8>
// CTAs per SM: 3 (f64 and strict: 80 registers; strict keeps its matrices in shared memory for that, measured 8.6 % faster than 2 CTAs at
// 128 registers), 4 (f32: 64 registers without spills; measured +11 % over 3.  The f64 variant spills in phase A at 64 registers and loses 25 %).
__global__ void __launch_bounds__(kThreads, MODE == HRMGPU_F32 ? 4 : 3) k_minksweep3d(const MinkParams p) {
    using Code = typename std::conditional<CODE8, uint16_t, uint32_t>::type;  // packed line codes: 2 x 8 or 2 x 16 bits
    using A = typename std::conditional<MODE == HRMGPU_F64_STRICT, Strict, Fast>::type;
    using Real = typename RealOf<MODE>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];

    const int n = p.n, M = p.M;
    const int per_layer = p.n_objs * p.B;
    // Heavy surfaces first: the arena objects (eps ~ 0.1, a few huge faces, tens of thousands of sweep candidates) take ~20x
    // longer than an obstacle surface, so their CTAs are issued at the start of the grid instead of trailing it.
    int k, s;
    {
        const int heavy = p.n_heavy * p.B;  // surfaces [0, heavy) of every layer
        const int nk = gridDim.x / per_layer;
        const int bid = blockIdx.x;
        if (bid < nk * heavy) {
            k = bid / heavy;
            s = bid - k * heavy;
        } else {
            const int rem = bid - nk * heavy, light = per_layer - heavy;
            k = rem / light;
            s = heavy + (rem - k * light);
        }
    }
    const int slot = k * per_layer + s;  // position of this surface in the output arrays
    const int oi = s / p.B;
    const int b = s - oi * p.B;
    const int obj_id = p.first_obj + oi;
    const int tid = threadIdx.x;
    const int grp = tid / p.lpc;
    const int ln = tid - grp * p.lpc;
    const bool worker = grp < p.groups;

    long long t_start = 0, t_pro = 0, t_a = 0, t_scan = 0, t_drain = 0;
    __shared__ unsigned long long dbg_max[4];
    if (threadIdx.x < 4) dbg_max[threadIdx.x] = 0;
    int rounds = 0;
    if (p.dbg) t_start = clock64();
    // ---- shared memory carve-up (every region 16-byte aligned) ------------------------------
    Real* tab = reinterpret_cast<Real*>(smem_raw);                             // [8][n]
    double* mats = reinterpret_cast<double*>(smem_raw + p.o_mats);             // M1[9] M2[9] + scratch T[9] KTT[9]
    Real* rowb = reinterpret_cast<Real*>(mats + 36);                           // RA[3] RB[3] UA[3] UB[3] WA[3] WB[3]
    Real* cc = reinterpret_cast<Real*>(smem_raw + p.o_cc);                     // [n][kColStride] per-column constants (fast modes)
    double* txs = reinterpret_cast<double*>(smem_raw + p.o_txs);               // [Lx+2]
    double* tys = reinterpret_cast<double*>(smem_raw + p.o_tys);               // [Ly+2]
    uint16_t* cellq = reinterpret_cast<uint16_t*>(smem_raw + p.o_queue);       // [kCellCap] kept cells (cell = i * (n-1) + j)
    uint16_t* heavyq = cellq + kCellCap;                                       // [kCellCap] of which: many candidates
    Code* vcode = reinterpret_cast<Code*>(smem_raw + p.o_vcode);               // [M + 8]: the 5th code of the last unit
    auto pack_code = [](uint32_t cx, uint32_t cy) -> Code { return static_cast<Code>(CODE8 ? (cx | (cy << 8)) : (cx | (cy << 16))); };
    // expanded form used by the SIMD min/max tests: x code in bits 0..15, y code in bits 16..31
    auto code_at = [&](int q) -> uint32_t { return CODE8 ? __byte_perm(static_cast<uint32_t>(vcode[q]), 0u, 0x4140) : static_cast<uint32_t>(vcode[q]); };
    uint32_t* m0 = reinterpret_cast<uint32_t*>(smem_raw + p.o_m0);             // [L]
    uint32_t* m1 = reinterpret_cast<uint32_t*>(smem_raw + p.o_m1);             // [L]
    int* qcount = reinterpret_cast<int*>(smem_raw + p.o_qcount);               // [0] listed cells, [1] next scan batch, [2] heavy cells

    const Obj3& ob = p.obj[obj_id];
    const bool sign_neg = __ldg(&ob.sign) < 0.0;  // arena surface (requested here, consumed in phase D)
    Real* mesh = reinterpret_cast<Real*>(p.mesh) + static_cast<size_t>(slot) * 3 * M;
    Real* mx = mesh;
    Real* my = mesh + M;
    Real* mz = mesh + 2 * static_cast<size_t>(M);

    if (p.reload) {
        // ---- re-sweep (refinement path, HighwayRoadMap-inl.h:178-215): the cached boundary of this surface is classified
        // against the new sweep grid; nothing of K1 is recomputed ----
        if (SWEEP) {
            for (int i = tid; i < p.Lx + 2; i += kThreads) txs[i] = __ldg(p.txs + i);
            for (int i = tid; i < p.Ly + 2; i += kThreads) tys[i] = __ldg(p.tys + i);
            for (int i = tid; i < p.L; i += kThreads) m0[i] = kNoFace, m1[i] = kNoFace;
            if (tid == 0) qcount[0] = 0, qcount[1] = 0, qcount[2] = 0;
            __syncthreads();
            for (int i = tid; i < M; i += kThreads) {
                const uint32_t cx = line_code(static_cast<double>(mx[i]), txs, p.Lx, p.inv_dx, p.bias_x);
                const uint32_t cy = line_code(static_cast<double>(my[i]), tys, p.Ly, p.inv_dy, p.bias_y);
                vcode[i] = pack_code(cx, cy);
            }
        }
    } else {
    // ---- prologue: tables, line coordinates, per-surface matrices -------------------------
    // warp 0 requests the operands of the per-surface matrices first: their (HBM) latency overlaps the table loads below
    const int me = tid < 9 ? tid : 0;
    double mi0 = 0, mi1 = 0, mi2 = 0, mj0 = 0, mj1 = 0, mj2 = 0, md0 = 0, md1 = 0, md2 = 0, msign = 0, r1a = 0, r1b = 0, r1c = 0;
    if (tid < 32) {
        const int i = me / 3, j = me - 3 * i;
        const double* r2 = p.R + (static_cast<size_t>(k) * p.B + b) * 9;
        mi0 = __ldg(r2 + 3 * i), mi1 = __ldg(r2 + 3 * i + 1), mi2 = __ldg(r2 + 3 * i + 2);
        mj0 = __ldg(r2 + 3 * j), mj1 = __ldg(r2 + 3 * j + 1), mj2 = __ldg(r2 + 3 * j + 2);
        const double* sm = p.semi + 3 * (p.semi_per_layer ? static_cast<size_t>(k) * p.B + b : static_cast<size_t>(b));
        md0 = __ldg(sm), md1 = __ldg(sm + 1), md2 = __ldg(sm + 2);
        msign = __ldg(&ob.sign);
        r1a = __ldg(&ob.R[j]), r1b = __ldg(&ob.R[3 + j]), r1c = __ldg(&ob.R[6 + j]);
    }
    const double* gtab = p.tab + static_cast<size_t>(obj_id) * 8 * n;
    {
        // All global requests of the prologue are issued back to back (one L2/HBM round trip), the shared-memory
        // initialisation that needs no data runs while they are in flight, and only then are the values consumed.
        double v[4], vx = 0.0, vy = 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = (tid + j * kThreads < 8 * n) ? __ldg(gtab + tid + j * kThreads) : 0.0;
        if (SWEEP) {
            if (tid < p.Lx + 2) vx = __ldg(p.txs + tid);
            if (tid < p.Ly + 2) vy = __ldg(p.tys + tid);
            for (int i = tid; i < p.L; i += kThreads) m0[i] = kNoFace, m1[i] = kNoFace;
            if (tid == 0) qcount[0] = 0, qcount[1] = 0, qcount[2] = 0;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (tid + j * kThreads < 8 * n) tab[tid + j * kThreads] = static_cast<Real>(v[j]);
        for (int i = tid + 4 * kThreads; i < 8 * n; i += kThreads) tab[i] = static_cast<Real>(__ldg(gtab + i));
        if (SWEEP) {
            if (tid < p.Lx + 2) txs[tid] = vx;
            if (tid < p.Ly + 2) tys[tid] = vy;
            for (int i = tid + kThreads; i < p.Lx + 2; i += kThreads) txs[i] = __ldg(p.txs + i);
            for (int i = tid + kThreads; i < p.Ly + 2; i += kThreads) tys[i] = __ldg(p.tys + i);
        }
    }
    if (tid < 32) {
        // Tinv = (R2 * diag) * R2^T ;  M1 = Tinv * R1 ;  M2 = ((K * Tinv) * Tinv) * R1
        // (SuperQuadrics.cpp:115,137-140), each entry evaluated in the reference's association order by its own
        // lane (9 lanes, three dependent steps through shared memory instead of ~90 serial FP64 operations).
        double* Ts = mats + 18;   // scratch: T[9], KTT[9]
        const double t = A::mad(A::mul(mi2, md2), mj2, A::mad(A::mul(mi1, md1), mj1, A::mul(A::mul(mi0, md0), mj0)));
        if (tid < 9) Ts[me] = t;
        __syncwarp();
        const int i = me / 3, j = me - 3 * i;
        const double ktt = A::mad(A::mul(msign, Ts[3 * i + 2]), Ts[6 + j],
                                  A::mad(A::mul(msign, Ts[3 * i + 1]), Ts[3 + j], A::mul(A::mul(msign, Ts[3 * i]), Ts[j])));
        if (tid < 9) Ts[9 + me] = ktt;
        __syncwarp();
        if (tid < 9) {
            mats[me] = A::mad(Ts[3 * i + 2], r1c, A::mad(Ts[3 * i + 1], r1b, A::mul(Ts[3 * i], r1a)));
            mats[9 + me] = A::mad(Ts[9 + 3 * i + 2], r1c, A::mad(Ts[9 + 3 * i + 1], r1b, A::mul(Ts[9 + 3 * i], r1a)));
        }
    }
    const double offx = __ldg(p.off + (static_cast<size_t>(k) * p.B + b) * 3);
    const double offy = __ldg(p.off + (static_cast<size_t>(k) * p.B + b) * 3 + 1);
    const double offz = __ldg(p.off + (static_cast<size_t>(k) * p.B + b) * 3 + 2);

    // object constants for the column table, requested before the barrier so their latency overlaps the table loads
    double pre_a = 0, pre_b = 0, pre_c = 0, pre_ia = 0, pre_ib = 0, pre_ic = 0, pre_px = 0, pre_py = 0, pre_pz = 0, pre_R[9] = {};
    if constexpr (MODE != HRMGPU_F64_STRICT) {
        if (tid & 1) {
            pre_a = __ldg(&ob.a), pre_b = __ldg(&ob.b), pre_c = __ldg(&ob.c);
            pre_px = __ldg(&ob.px), pre_py = __ldg(&ob.py), pre_pz = __ldg(&ob.pz);
#pragma unroll
            for (int i = 0; i < 3; ++i) pre_R[i] = __ldg(&ob.R[3 * i]), pre_R[3 + i] = __ldg(&ob.R[3 * i + 1]), pre_R[6 + i] = __ldg(&ob.R[3 * i + 2]);
        } else {
            pre_ia = __ldg(&ob.ia), pre_ib = __ldg(&ob.ib), pre_ic = __ldg(&ob.ic);
        }
    }

    __syncthreads();
    if (p.dbg && tid == 0) dbg_max[2] = (unsigned long long)(clock64() - t_start);
    if constexpr (MODE == HRMGPU_F64_STRICT) {  // R1 beside M1, M2 (the scratch of the matrix products is free now)
        if (tid < 9) mats[18 + tid] = __ldg(&ob.R[tid]);
        __syncthreads();
    }
    if constexpr (MODE != HRMGPU_F64_STRICT) {
        // column table and row-base coefficients.  Two threads per column: the even one writes {ce2, UZ, WZ}, the odd one
        // {ce1, C}; the first two threads also write the row base.
        const int part = tid & 1;
        if (tid < 2) {  // (static indices only: pre_R must stay in registers)
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                if (part) {
                    rowb[d] = Real(pre_R[d] * pre_a), rowb[3 + d] = Real(pre_R[3 + d] * pre_b);
                } else {
                    rowb[6 + d] = Real(mats[3 * d] * pre_ia), rowb[9 + d] = Real(mats[3 * d + 1] * pre_ib);
                    rowb[12 + d] = Real(mats[9 + 3 * d] * pre_ia), rowb[15 + d] = Real(mats[9 + 3 * d + 1] * pre_ib);
                }
            }
        }
        for (int c = tid >> 1; c < n; c += kThreads >> 1) {
            Real* k = cc + c * kColStride;
            if (part == 0) {
                const Real ce2 = tab[2 * n + c], gz = tab[3 * n + c] * Real(pre_ic);  // se2 / c
                k[1] = ce2;
#pragma unroll
                for (int i = 0; i < 3; ++i) k[5 + i] = Real(mats[3 * i + 2]) * gz, k[8 + i] = Real(mats[9 + 3 * i + 2]) * gz;
            } else {
                const Real ce1 = tab[c], Zc = Real(pre_c) * tab[n + c];  // c * se1
                const Real tr[3] = {Real(pre_px - offx), Real(pre_py - offy), Real(pre_pz - offz)};
                k[0] = ce1;
#pragma unroll
                for (int i = 0; i < 3; ++i) k[2 + i] = Real(pre_R[6 + i]) * Zc + tr[i];
            }
        }
        __syncthreads();
    }
    if (p.dbg) t_pro = clock64();

    // ---- phase A: boundary points ---------------------------------------------------------
    if (worker) {
        const double oa = __ldg(&ob.a), obb = __ldg(&ob.b), oc = __ldg(&ob.c);
        const double px = __ldg(&ob.px), py = __ldg(&ob.py), pz = __ldg(&ob.pz);
        if constexpr (MODE == HRMGPU_F64_STRICT) {
            // IEEE replica of the reference's evaluation order; only exact products are hoisted per column.
            // The 27 matrix entries stay in shared memory (volatile: hoisted back into registers they take 54 of them and the kernel
            // drops to 2 CTAs per SM; measured 5.46 -> 4.99 ms per 32 layers).
            const volatile double* M1 = mats;
            const volatile double* M2 = mats + 9;
            const volatile double* R1 = mats + 18;
            for (int c = grp; c < n; c += p.groups) {
                const double ce1 = tab[c], se1 = tab[n + c], ce2 = tab[2 * n + c], se2 = tab[3 * n + c];
                const double Z = A::mul(oc, se1);   // c * (se1 * 1.0)
                const double gz = A::div(se2, oc);  // (se2 * 1.0) / c
                const double rz0 = A::mul(R1[2], Z), rz1 = A::mul(R1[5], Z), rz2 = A::mul(R1[8], Z);
                const double uz0 = A::mul(M1[2], gz), uz1 = A::mul(M1[5], gz), uz2 = A::mul(M1[8], gz);
                const double wz0 = A::mul(M2[2], gz), wz1 = A::mul(M2[5], gz), wz2 = A::mul(M2[8], gz);
                auto point = [&](int r) {
                    const double cw1 = tab[4 * n + r], sw1 = tab[5 * n + r], cw2 = tab[6 * n + r], sw2 = tab[7 * n + r];
                    const int kk = c * n + r;
                    // getOriginShape: canonical point, rotate, translate              SuperQuadrics.cpp:58-78
                    const double X = A::mul(oa, A::mul(ce1, cw1));
                    const double Y = A::mul(obb, A::mul(ce1, sw1));
                    const double x0 = A::add(A::add(A::mad(R1[1], Y, A::mul(R1[0], X)), rz0), px);
                    const double y0 = A::add(A::add(A::mad(R1[4], Y, A::mul(R1[3], X)), rz1), py);
                    const double z0 = A::add(A::add(A::mad(R1[7], Y, A::mul(R1[6], X)), rz2), pz);
                    // gradient                                                       SuperQuadrics.cpp:118-135
                    const double gx = A::div(A::mul(ce2, cw2), oa);
                    const double gy = A::div(A::mul(ce2, sw2), obb);
                    const double ux = A::add(A::mad(M1[1], gy, A::mul(M1[0], gx)), uz0);
                    const double uy = A::add(A::mad(M1[4], gy, A::mul(M1[3], gx)), uz1);
                    const double uz = A::add(A::mad(M1[7], gy, A::mul(M1[6], gx)), uz2);
                    const double wx = A::add(A::mad(M2[1], gy, A::mul(M2[0], gx)), wz0);
                    const double wy = A::add(A::mad(M2[4], gy, A::mul(M2[3], gx)), wz1);
                    const double wz = A::add(A::mad(M2[7], gy, A::mul(M2[6], gx)), wz2);
                    const double nrm = A::sqrt(A::mad(uz, uz, A::mad(uy, uy, A::mul(ux, ux))));  // :137-140
                    const double ox = A::sub(A::add(x0, A::div(wx, nrm)), offx);                   // MultiBodyTree3D.cpp:98-101
                    const double oy = A::sub(A::add(y0, A::div(wy, nrm)), offy);
                    const double oz = A::sub(A::add(z0, A::div(wz, nrm)), offz);
                    mx[kk] = ox;
                    my[kk] = oy;
                    mz[kk] = oz;
                    if (SWEEP) {
                        const uint32_t cx = line_code(ox, txs, p.Lx, p.inv_dx, p.bias_x);
                        const uint32_t cy = line_code(oy, tys, p.Ly, p.inv_dy, p.bias_y);
                        vcode[kk] = pack_code(cx, cy);
                    }
                };
                for (int r = ln; r < n; r += p.lpc) point(r);
            }
        } else {
            // Fast variants: x = P*cw1 + Q*sw1 + C,  u = U1*cw2 + U2*sw2 + U3,  w = W1*cw2 + W2*sw2 + W3 with the nine
            // 3-vectors depending on the column only (table `cc`, built below); out = x + w * rsqrt(u.u).
            // A thread keeps the omega-dependent factors of its rows in registers and streams the columns' constants from
            // shared memory (all lanes of a group read the same address: broadcast), which keeps the register count low
            // enough for 3 CTAs per SM.
            using R2 = Real2<Real>;
            constexpr int kRows = 2;
            const int nh = (n + p.lpc - 1) / p.lpc;
            if (p.paired) {
                // A thread owns the adjacent rows (r0, r0 + 1): the two points of a column are one 16-byte store per
                // coordinate plane and one packed store of their line codes; no tail handling (lpc divides n / 2).
                const int npass = n / (2 * p.lpc);
                const uint32_t cc_step = static_cast<uint32_t>(p.groups * kColStride * sizeof(Real));
                const uint32_t v_step = static_cast<uint32_t>(p.groups * n * sizeof(Code));
                const size_t g_step = static_cast<size_t>(p.groups) * n;
                const uint32_t lx2 = static_cast<uint32_t>(p.Lx) * 0x00010001u, ly2 = static_cast<uint32_t>(p.Ly) * 0x00010001u;
                for (int h = 0; h < npass; ++h) {
                    const int r0 = 2 * (ln + h * p.lpc);
                    Real V[2][3], U[2][3], W[2][3];  // row factors of the two rows (see kColStride)
                    {
                        Real cw1[2], sw1[2], cw2[2], sw2[2];
                        lds_pair(smem_u32(tab + 4 * n + r0), cw1[0], cw1[1]);
                        lds_pair(smem_u32(tab + 5 * n + r0), sw1[0], sw1[1]);
                        lds_pair(smem_u32(tab + 6 * n + r0), cw2[0], cw2[1]);
                        lds_pair(smem_u32(tab + 7 * n + r0), sw2[0], sw2[1]);
#pragma unroll
                        for (int h2 = 0; h2 < 2; ++h2)
#pragma unroll
                            for (int d = 0; d < 3; ++d) {
                                V[h2][d] = rowb[d] * cw1[h2] + rowb[3 + d] * sw1[h2];
                                U[h2][d] = rowb[6 + d] * cw2[h2] + rowb[9 + d] * sw2[h2];
                                W[h2][d] = rowb[12 + d] * cw2[h2] + rowb[15 + d] * sw2[h2];
                            }
                    }
                    uint32_t ca = smem_u32(cc) + static_cast<uint32_t>(grp * kColStride * sizeof(Real));
                    uint32_t va = smem_u32(vcode) + static_cast<uint32_t>((grp * n + r0) * sizeof(Code));
                    Real* gx = mx + grp * n + r0;
                    for (int c = grp; c < n; c += p.groups, ca += cc_step, va += v_step, gx += g_step) {
                        Real k[12];  // ce1, ce2, C[3], UZ[3], WZ[3], pad
#pragma unroll
                        for (int i = 0; i < 6; ++i) lds_pair(ca + i * 2 * sizeof(Real), k[2 * i], k[2 * i + 1]);
                        Real o[2][3];
#pragma unroll
                        for (int h2 = 0; h2 < 2; ++h2) {
                            const Real ux = k[1] * U[h2][0] + k[5], uy = k[1] * U[h2][1] + k[6], uz = k[1] * U[h2][2] + k[7];
                            Real inv;
                            if constexpr (MODE == HRMGPU_F32)
                                inv = rsqrtf(ux * ux + uy * uy + uz * uz);
                            else
                                inv = fast_rsqrt(ux * ux + uy * uy + uz * uz);
#pragma unroll
                            for (int d = 0; d < 3; ++d) o[h2][d] = (k[1] * W[h2][d] + k[8 + d]) * inv + (k[0] * V[h2][d] + k[2 + d]);
                        }
                        *reinterpret_cast<R2*>(gx) = R2{o[0][0], o[1][0]};
                        *reinterpret_cast<R2*>(gx + M) = R2{o[0][1], o[1][1]};
                        *reinterpret_cast<R2*>(gx + 2 * static_cast<size_t>(M)) = R2{o[0][2], o[1][2]};
                        if (SWEEP) {
                            bool k0, k1, k2, k3;
                            const uint32_t wx0 = line_q(static_cast<double>(o[0][0]), p.inv_dx, p.bias_xq, k0);
                            const uint32_t wy0 = line_q(static_cast<double>(o[0][1]), p.inv_dy, p.bias_yq, k1);
                            const uint32_t wx1 = line_q(static_cast<double>(o[1][0]), p.inv_dx, p.bias_xq, k2);
                            const uint32_t wy1 = line_q(static_cast<double>(o[1][1]), p.inv_dy, p.bias_yq, k3);
                            if (k0 && k1 && k2 && k3) {
                                // both points at once: [ceil(g1) : ceil(g0)] in 16-bit halves, clamped to [0, Lc]; code = 2 * index
                                const uint32_t X = __vimin_s16x2_relu(__byte_perm(wx0, wx1, 0x7632), lx2);
                                const uint32_t Y = __vimin_s16x2_relu(__byte_perm(wy0, wy1, 0x7632), ly2);
                                if constexpr (CODE8)  // index <= 127: no carry between the byte fields
                                    asm volatile("st.shared.u32 [%0], %1;" ::"r"(va), "r"((X | (Y << 8)) << 1) : "memory");
                                else
                                    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(va), "r"(__byte_perm(X, Y, 0x5410) << 1),
                                                 "r"(__byte_perm(X, Y, 0x7632) << 1)
                                                 : "memory");
                            } else {  // rare: a coordinate within rounding distance of a sweep line -> exact search
                                const uint32_t cx0 = line_code(static_cast<double>(o[0][0]), txs, p.Lx, p.inv_dx, p.bias_x);
                                const uint32_t cy0 = line_code(static_cast<double>(o[0][1]), tys, p.Ly, p.inv_dy, p.bias_y);
                                const uint32_t cx1 = line_code(static_cast<double>(o[1][0]), txs, p.Lx, p.inv_dx, p.bias_x);
                                const uint32_t cy1 = line_code(static_cast<double>(o[1][1]), tys, p.Ly, p.inv_dy, p.bias_y);
                                sts_codes(va, pack_code(cx0, cy0), pack_code(cx1, cy1));
                            }
                        }
                    }
                }
            } else
            for (int h0 = 0; h0 < nh; h0 += kRows) {
                int rr[kRows];
                bool rvalid[kRows];
                Real V[kRows][3], U[kRows][3], W[kRows][3];  // row factors (see kColStride)
#pragma unroll
                for (int h = 0; h < kRows; ++h) {
                    const int r = ln + (h0 + h) * p.lpc;
                    rvalid[h] = r < n;
                    rr[h] = rvalid[h] ? r : ln;  // rows past the end compute the thread's first row again and store nothing
                    const Real cw1 = tab[4 * n + rr[h]], sw1 = tab[5 * n + rr[h]], cw2 = tab[6 * n + rr[h]], sw2 = tab[7 * n + rr[h]];
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        V[h][d] = rowb[d] * cw1 + rowb[3 + d] * sw1;
                        U[h][d] = rowb[6 + d] * cw2 + rowb[9 + d] * sw2;
                        W[h][d] = rowb[12 + d] * cw2 + rowb[15 + d] * sw2;
                    }
                }
                int koff = grp * n;  // c * n
                const int kstep = p.groups * n;
                for (int c = grp; c < n; c += p.groups, koff += kstep) {
                    const R2* k2 = reinterpret_cast<const R2*>(cc + c * kColStride);
                    Real k[12];  // ce1, ce2, C[3], UZ[3], WZ[3], pad
#pragma unroll
                    for (int i = 0; i < 6; ++i) {
                        const R2 v = k2[i];
                        k[2 * i] = v.a, k[2 * i + 1] = v.b;
                    }
                    Real o[kRows][3];
#pragma unroll
                    for (int h = 0; h < kRows; ++h) {
                        const Real ux = k[1] * U[h][0] + k[5], uy = k[1] * U[h][1] + k[6], uz = k[1] * U[h][2] + k[7];
                        Real inv;
                        if constexpr (MODE == HRMGPU_F32)
                            inv = rsqrtf(ux * ux + uy * uy + uz * uz);
                        else
                            inv = fast_rsqrt(ux * ux + uy * uy + uz * uz);
#pragma unroll
                        for (int d = 0; d < 3; ++d) o[h][d] = (k[1] * W[h][d] + k[8 + d]) * inv + (k[0] * V[h][d] + k[2 + d]);
                    }
                    uint32_t cx[kRows], cy[kRows];
                    bool all_ok = true;
#pragma unroll
                    for (int h = 0; h < kRows; ++h) {
                        const int kk = koff + rr[h];
                        if (rvalid[h]) {
                            mx[kk] = o[h][0];
                            my[kk] = o[h][1];
                            mz[kk] = o[h][2];
                        }
                        if (SWEEP) {
                            bool okx, oky;
                            cx[h] = line_code_fast(static_cast<double>(o[h][0]), p.Lx, p.inv_dx, p.bias_x, okx);
                            cy[h] = line_code_fast(static_cast<double>(o[h][1]), p.Ly, p.inv_dy, p.bias_y, oky);
                            all_ok = all_ok && okx && oky;
                        }
                    }
                    if (SWEEP) {
                        if (!all_ok) {  // rare: some coordinate within rounding distance of a sweep line -> exact search
#pragma unroll
                            for (int h = 0; h < kRows; ++h) {
                                cx[h] = line_code(static_cast<double>(o[h][0]), txs, p.Lx, p.inv_dx, p.bias_x);
                                cy[h] = line_code(static_cast<double>(o[h][1]), tys, p.Ly, p.inv_dy, p.bias_y);
                            }
                        }
#pragma unroll
                        for (int h = 0; h < kRows; ++h)
                            if (rvalid[h]) vcode[koff + rr[h]] = pack_code(cx[h], cy[h]);
                    }
                }
            }
        }
    }
    }  // !p.reload
    if (!SWEEP) return;
    __syncwarp();  // lanes leave the column/row loops at different trips: reconverge before the block barrier
    __syncthreads();
    if (p.dbg) t_a = clock64();

    // Vertices of face f, and the triangle test of a loaded face against the sweep line (ix, iy).
    auto load_face = [&](uint32_t f, D3& t0, D3& t1, D3& t2) {
        int v0, v1, v2;
        face_vertices(static_cast<int>(f), n, v0, v1, v2);
        t0 = {static_cast<double>(mx[v0]), static_cast<double>(my[v0]), static_cast<double>(mz[v0])};
        t1 = {static_cast<double>(mx[v1]), static_cast<double>(my[v1]), static_cast<double>(mz[v1])};
        t2 = {static_cast<double>(mx[v2]), static_cast<double>(my[v2]), static_cast<double>(mz[v2])};
    };
    auto test_loaded = [&](uint32_t ix, uint32_t iy, const D3& t0, const D3& t1, const D3& t2, double& z) -> bool {
        if constexpr (MODE == HRMGPU_F64_STRICT) {
            const D3 o = {txs[ix + 1], tys[iy + 1], 0.0};
            const D3 d = {0.0, 0.0, 1.0};
            D3 pt;
            const bool hit = line_triangle<A>(o, d, t0, t1, t2, pt);
            z = pt.z;
            return hit;
        } else {
            return vertical_hit_fast(txs[ix + 1], tys[iy + 1], t0, t1, t2, z);
        }
    };

    // ---- phases B + C: scan cells -> list of kept cells -> per-cell triangle tests ----------------------------
    // B1 (warps, dynamic batches): SWAR test of every mesh cell, kept cells appended to a shared list.
    // C  (threads): a thread takes kept cells, forms the exact line sets of the cell's two faces from the vertex codes
    //    (LineIntersection.cpp:77-101), loads the four corners once and tests the few (face, line) candidates itself; a hit
    //    enters the line's ordered two-slot table by its FACE INDEX (the reference keeps the first two hits in face order,
    //    SURVEY finding 2).  Cells with many candidates (the eps = 0.1 arena: a few faces cover hundreds of lines) go to a
    //    second list whose lines are spread over the whole CTA.
    // D  re-evaluates the two surviving faces of a line for their z (their vertices were just read by this SM: L1 hits), so
    //    no per-hit storage, no commit rounds and no capacity limit on hits exist.
    const int nc1 = n - 1;
    const int ncell = nc1 * nc1;
    const int UPR = (nc1 + 3) >> 2;            // units (4 consecutive cells of a mesh row) per mesh row
    const int nunits = nc1 * UPR;
    const uint32_t upr_magic = (UPR > 1) ? 0xFFFFFFFFu / static_cast<uint32_t>(UPR) + 1u : 0u;  // u / UPR == umulhi(u, magic), u < 2^16
    const uint32_t nc1_magic = 0xFFFFFFFFu / static_cast<uint32_t>(nc1) + 1u;                   // cell / nc1 likewise (nc1 >= 2)
    const bool vec_ok = (n & 3) == 0;
    constexpr int kUPL = 4;  // units per lane and batch
    static_assert(kUPL <= 4, "unit_cells packs the cells of unit uu at bits 7 + 2 * uu, 8 + 2 * uu, 23 + 2 * uu, 24 + 2 * uu");
    const int lane = tid & 31;
    int* hcount = qcount + 2;
    // B1: cells of unit u whose inclusive xy bounding box (over the 4 corners) may contain a sweep line.  Per axis the set of
    // lines inside the box of a vertex set is empty  <=>  all codes (2*lo + eq) are equal and even, so a cell is kept iff on
    // BOTH axes some corner code differs from the first one or is odd: XOR / OR over packed codes, no min/max.
    // Result layout (per unit, before the batch shift): cell 0 -> bit 7, cell 2 -> bit 8, cell 1 -> bit 23, cell 3 -> bit 24
    // (where the SWAR form leaves them).
    auto unit_cells = [&](int unit_raw, bool fast) -> uint32_t {
        const int unit = min(unit_raw, nunits - 1);  // (units past the end are evaluated on the last one and dropped: no branch)
        const int i = (UPR > 1) ? static_cast<int>(__umulhi(static_cast<uint32_t>(unit), upr_magic)) : unit;
        const int j0 = 4 * (unit - i * UPR);
        const int q = i * n + j0;
        uint32_t g;
        if (fast) {
            // four vertices (x byte, y byte each) per 64-bit load; sa / sb = the same rows advanced by one vertex
            const uint2 a = *reinterpret_cast<const uint2*>(vcode + q);
            const uint2 b = *reinterpret_cast<const uint2*>(vcode + q + n);
            const uint32_t a4 = static_cast<uint32_t>(vcode[q + 4]), b4 = static_cast<uint32_t>(vcode[q + n + 4]);
            const uint32_t sax = __funnelshift_r(a.x, a.y, 16), say = __funnelshift_r(a.y, a4, 16);
            const uint32_t sbx = __funnelshift_r(b.x, b.y, 16), sby = __funnelshift_r(b.y, b4, 16);
            // (all four equal and odd is caught by the first corner's own low bits: unequal corners set a difference bit)
            const uint32_t dx = (a.x ^ b.x) | (a.x ^ sax) | (a.x ^ sbx) | (a.x & 0x01010101u);
            const uint32_t dy = (a.y ^ b.y) | (a.y ^ say) | (a.y ^ sby) | (a.y & 0x01010101u);
            // bit 7 of every non-zero byte, then both bytes of a vertex field
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
        // cells past the end of the mesh row (v = cells left in the row) and units past the end of the mesh
        // cells 0 .. v-1 of the unit are bits 7, 23, 8, 24 (v = 0: unit past the end).  Mask arithmetic, not a ladder of selects:
        // the compiler turns that into jumps (even a jump table), which serialises the independent units of a lane.
        const int v = (unit_raw < nunits) ? nc1 - j0 : 0;
        const uint32_t keep = (0x00000080u & (0u - static_cast<uint32_t>(v > 0))) | (0x00800000u & (0u - static_cast<uint32_t>(v > 1))) |
                              (0x00000100u & (0u - static_cast<uint32_t>(v > 2))) | (0x01000000u & (0u - static_cast<uint32_t>(v > 3)));
        return g & keep;
    };
    // exact line sets of the two faces of cell (i, j), q = i*n + j: per 16-bit half lo = min(2*lo+eq) >> 1,
    // hi+1 = (max(2*lo+eq) + 1) >> 1, lines [lo, hi); n* = number of (face, line) candidates
    auto cell_sets = [&](int q, uint32_t& loA, uint32_t& hiA, uint32_t& loB, uint32_t& hiB, uint32_t& nA, uint32_t& nB) {
        const uint32_t c00 = code_at(q), c11 = code_at(q + n + 1), c10 = code_at(q + n), c01 = code_at(q + 1);
        const uint32_t mnA = __vimin3_u16x2(c00, c11, c10), mxA = __vimax3_u16x2(c00, c11, c10);
        const uint32_t mnB = __vimin3_u16x2(c00, c11, c01), mxB = __vimax3_u16x2(c00, c11, c01);
        loA = (mnA >> 1) & 0x7FFF7FFFu, hiA = ((mxA + 0x00010001u) >> 1) & 0x7FFF7FFFu;
        loB = (mnB >> 1) & 0x7FFF7FFFu, hiB = ((mxB + 0x00010001u) >> 1) & 0x7FFF7FFFu;
        const uint32_t dA = __vsub2(hiA, loA), dB = __vsub2(hiB, loB);  // hi >= lo in both halves
        nA = (dA & 0xFFFFu) * (dA >> 16), nB = (dB & 0xFFFFu) * (dB >> 16);
    };
    auto vert = [&](int v) -> D3 { return {static_cast<double>(mx[v]), static_cast<double>(my[v]), static_cast<double>(mz[v])}; };
    auto test_lines = [&](uint32_t face, uint32_t lo2, uint32_t hi2, const D3& t0, const D3& t1, const D3& t2) {
        for (uint32_t ix = lo2 & 0xFFFFu; ix < (hi2 & 0xFFFFu); ++ix)
            for (uint32_t iy = lo2 >> 16; iy < (hi2 >> 16); ++iy) {
                double z;
                if (test_loaded(ix, iy, t0, t1, t2, z)) {
                    const uint32_t l = ix * p.Ly + iy;
                    insert_two_smallest(&m0[l], &m1[l], face);
                }
            }
    };
    constexpr uint32_t kLight = 16;  // cells with more candidates than this get a whole warp

    constexpr int kWarps = kThreads / 32;
    constexpr int kSegStride = kCellCap / kWarps;  // every warp lists its kept cells in its own segment: no atomics in the scan
    const int kSegCap = p.seg_cap;                 // (<= kSegStride)
    const int warp = tid >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    uint16_t* seg = cellq + warp * kSegStride;
    int* segcnt = qcount + 4;          // [kWarps]
    int bw = warp;                 // current batch (warp-uniform); batches are dealt round-robin (their cost is uniform)
    bool batch_loaded = false;
    uint32_t w_bits = 0;           // kept cells of the lane's units that are not listed yet
    for (;;) {
        long long tb1 = 0;
        if (p.dbg) tb1 = clock64();
        int wcnt = 0;
        bool parked = false;
        while (bw * (32 * kUPL) < nunits) {
            if (!batch_loaded) {
                // consecutive lanes read consecutive 8-byte words (no bank conflicts); the kUPL units of a lane are independent
                // straight-line code, so their loads and SWAR chains interleave
                w_bits = 0u;
                if (vec_ok && CODE8) {
#pragma unroll
                    for (int uu = 0; uu < kUPL; ++uu) w_bits += unit_cells((bw * kUPL + uu) * 32 + lane, true) << (2 * uu);
                } else {
#pragma unroll
                    for (int uu = 0; uu < kUPL; ++uu) w_bits += unit_cells((bw * kUPL + uu) * 32 + lane, false) << (2 * uu);
                }
                batch_loaded = true;
            }
            // append the kept cells: per round every lane that still has one writes its lowest; what does not fit into the
            // segment stays in w_bits for the next round
            unsigned any = __ballot_sync(0xFFFFFFFFu, w_bits != 0u);
            while (any != 0u) {
                const int pos = wcnt + __popc(any & lt_mask);
                if (w_bits != 0u && pos < kSegCap) {
                    const int bit = __ffs(w_bits) - 1;
                    w_bits &= w_bits - 1u;
                    // decode (see unit_cells): bits 7.. hold cells 0 / 2, bits 23.. cells 1 / 3 of the lane's kUPL units
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
            if (__any_sync(0xFFFFFFFFu, w_bits != 0u)) {  // segment full: the tests come first
                parked = true;
                break;
            }
            batch_loaded = false;
            bw += kWarps;
        }
        if (lane == 0) segcnt[warp] = wcnt;
        if (p.dbg && rounds == 0) atomicMax(&dbg_max[0], (unsigned long long)(clock64() - tb1));
        __syncwarp();
        const int more = __syncthreads_or(parked ? 1 : 0);  // no warp has cells left: this round is the last one
        long long tq0 = 0;
        if (p.dbg) tq0 = clock64();
        // C: one kept cell per thread and pass
        int sc[kWarps + 1];
        sc[0] = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) sc[w + 1] = sc[w] + segcnt[w];
        const int ncq = sc[kWarps];
        for (int base = 0; base < ncq; base += kThreads) {
            const int ci = base + tid;
            if (ci < ncq) {
                int w = 0;
#pragma unroll
                for (int ww = 1; ww < kWarps; ++ww) w += (ci >= sc[ww]) ? 1 : 0;
                int off_w = 0;
#pragma unroll
                for (int ww = 1; ww < kWarps; ++ww) off_w = (ci >= sc[ww]) ? sc[ww] : off_w;
                const uint32_t cell = cellq[w * kSegStride + (ci - off_w)];
                const uint32_t i = __umulhi(cell, nc1_magic);
                const int q = static_cast<int>(i) * n + static_cast<int>(cell - i * nc1);
                uint32_t loA, hiA, loB, hiB, nA, nB;
                cell_sets(q, loA, hiA, loB, hiB, nA, nB);
                if (nA + nB > kLight) {
                    heavyq[atomicAdd(hcount, 1)] = static_cast<uint16_t>(cell);
                } else if (nA + nB != 0u) {  // (the cell test is a superset of the two face tests)
                    const D3 c00 = vert(q), c11 = vert(q + n + 1);
                    if (nA != 0u) test_lines(cell, loA, hiA, c00, vert(q + n), c11);                                  // (q, q+n, q+n+1)
                    if (nB != 0u) test_lines(cell + static_cast<uint32_t>(ncell), loB, hiB, c00, vert(q + 1), c11);  // (q, q+1, q+n+1)
                }
            }
        }
        __syncwarp();
        __syncthreads();
        const int nh = *hcount;  // (at most one entry per listed cell)
        if (nh != 0) {
            // a warp per heavy cell, its (face, line) candidates over the lanes
            for (int h = warp; h < nh; h += kWarps) {
                const uint32_t cell = heavyq[h];
                const uint32_t i = __umulhi(cell, nc1_magic);
                const int q = static_cast<int>(i) * n + static_cast<int>(cell - i * nc1);
                uint32_t loA, hiA, loB, hiB, nA, nB;
                cell_sets(q, loA, hiA, loB, hiB, nA, nB);
                const D3 c00 = vert(q), c10 = vert(q + n), c11 = vert(q + n + 1), c01 = vert(q + 1);
                for (uint32_t idx = lane; idx < nA + nB; idx += 32) {
                    const bool second = idx >= nA;
                    const uint32_t k2 = second ? idx - nA : idx;
                    const uint32_t lo2 = second ? loB : loA, hi2 = second ? hiB : hiA;
                    const uint32_t wy = (hi2 >> 16) - (lo2 >> 16);
                    const uint32_t dxl = k2 / wy;
                    const uint32_t ix = (lo2 & 0xFFFFu) + dxl, iy = (lo2 >> 16) + (k2 - dxl * wy);
                    double z;
                    if (test_loaded(ix, iy, c00, second ? c01 : c10, c11, z)) {
                        const uint32_t l = ix * p.Ly + iy;
                        insert_two_smallest(&m0[l], &m1[l], cell + (second ? static_cast<uint32_t>(ncell) : 0u));
                    }
                }
            }
            __syncwarp();
            __syncthreads();
        }
        if (p.dbg) t_drain += clock64() - tq0;
        if (!more) break;
        if (tid == 0) *hcount = 0;
        ++rounds;
        __syncthreads();
    }

    if (p.dbg) t_scan = clock64();
    // ---- phase D: intervals ----------------------------------------------------------------
    const bool is_arena = sign_neg;
    double* lo = p.lo + static_cast<size_t>(slot) * p.L;
    double* up = p.up + static_cast<size_t>(slot) * p.L;
    unsigned single = 0;
    for (int l = tid; l < p.L; l += kThreads) {
        const uint32_t e0 = m0[l], e1 = m1[l];
        double a = is_arena ? p.zlow : __longlong_as_double(0x7FF8000000000000LL);
        double bb = is_arena ? p.zup : __longlong_as_double(0x7FF8000000000000LL);
        if (e1 != kNoFace) {
            // z of the two lowest hit faces: the same test on the same vertices as in phase C
            const uint32_t ix = static_cast<uint32_t>(l) / static_cast<uint32_t>(p.Ly), iy = static_cast<uint32_t>(l) - ix * static_cast<uint32_t>(p.Ly);
            D3 a0, a1, a2, b0, b1, b2;
            load_face(e0, a0, a1, a2);
            load_face(e1, b0, b1, b2);
            double z0 = 0.0, z1 = 0.0;
            test_loaded(ix, iy, a0, a1, a2, z0);
            test_loaded(ix, iy, b0, b1, b2, z1);
            const double zl = fmin(z0, z1), zu = fmax(z0, z1);
            a = is_arena ? fmin(p.zlow, zl) : zl;  // FreeSpace3D.cpp:44-49 / 61-64
            bb = is_arena ? fmax(p.zup, zu) : zu;
        } else if (e0 != kNoFace) {
            ++single;  // exactly one hit: reference reads out of bounds; defined here as "no interval"
        }
        lo[l] = a;
        up[l] = bb;
    }
    if (single) atomicAdd(p.counters, static_cast<unsigned long long>(single));
    if (p.dbg) {
        atomicMax(&dbg_max[1], (unsigned long long)(clock64() - t_scan));
        __syncthreads();
    }
    if (p.dbg && tid == 0) {
        long long* d = p.dbg + static_cast<size_t>(slot) * 8;
        d[3] = t_drain;              // of which phase C (cell tests)
        d[0] = t_pro - t_start;      // prologue
        d[1] = t_a - t_pro;          // phase A
        d[2] = t_scan - t_a;         // scan + tests
        d[4] = (long long)dbg_max[0];   // max over warps: B1 time in the first round
        d[6] = (long long)dbg_max[1];   // max over threads: phase D
        d[5] = (long long)dbg_max[2];   // prologue up to the first barrier (loads arrived, matrices built)
        d[7] = clock64() - t_start;
    }
}

// Shared-memory layout of k_minksweep3d (every region 16-byte aligned); fills the offsets of `p` and returns the total.
static size_t smem_layout(MinkParams& p, int MODE, bool sweep, bool code8, int n, int Lx, int Ly) {
    const size_t real = (MODE == HRMGPU_F32) ? 4 : 8;
    auto al = [](size_t v) { return (v + 15) & ~size_t(15); };
    size_t b = al(real * 8 * n);  // tab
    p.o_mats = static_cast<int>(b), b += sizeof(double) * (36 + 18);  // + 18 row-base coefficients (fast modes)
    p.o_txs = static_cast<int>(b), p.o_tys = static_cast<int>(b + (sweep ? sizeof(double) * (Lx + 2) : 0));
    if (sweep) b += al(sizeof(double) * (Lx + Ly + 4));
    // the column table is read in phase A only, the cell lists are used behind the barrier that ends phase A: one region
    p.o_cc = static_cast<int>(b);
    p.o_queue = static_cast<int>(b);
    {
        const size_t cc_bytes = (MODE != HRMGPU_F64_STRICT) ? al(real * kColStride * n) : 0;
        const size_t q_bytes = sweep ? 2 * sizeof(uint16_t) * kCellCap : 0;
        b += cc_bytes > q_bytes ? cc_bytes : q_bytes;
    }
    p.o_vcode = static_cast<int>(b), b += sweep ? al((code8 ? 2 : 4) * (static_cast<size_t>(n) * n + 8)) : 0;
    const size_t L = static_cast<size_t>(Lx) * Ly;
    p.o_m0 = static_cast<int>(b), p.o_m1 = static_cast<int>(b + (sweep ? sizeof(uint32_t) * L : 0)), b += sweep ? al(sizeof(uint32_t) * 2 * L) : 0;
    p.o_qcount = static_cast<int>(b);
    return b + 16 + 32;  // counters: listed cells, scan batch, heavy cells, pad, cells per warp segment [8]
}

// Lanes per column group: maximise n^2 / (threads * ceil(n/lpc) * ceil(n/groups)).
static void pick_mapping(int n, int& lpc, int& groups, int& paired) {
    double best = -1;
    lpc = 32;
    groups = kThreads / 32;
    paired = 0;
    if ((n & 1) == 0) {  // paired rows: lanes-per-column must divide n / 2
        // A group may be wider than a warp: with lpc = n / 2 a group stores a whole eta-column (n values, contiguous) per
        // iteration and the CTA `groups` adjacent columns -- whole 32-byte sectors only.  (A column split over two passes
        // leaves a half-written sector per column and plane whenever lpc * 16 bytes is not a multiple of 32: the L2 fills
        // it from DRAM, and those reads inside the write stream cost ~40 % of the store bandwidth; measured.)
        for (int l = (n / 2 < kThreads / 2 ? n / 2 : kThreads / 2); l >= 2; --l) {
            if ((n / 2) % l) continue;
            const int g = kThreads / l;
            const double steps = static_cast<double>(n / (2 * l)) * ((n + g - 1) / g);
            const double eff = static_cast<double>(n) * n / (2.0 * steps * kThreads);
            if (eff > best * 1.03) best = eff, lpc = l, groups = g, paired = 1;
        }
        if (paired) return;
        best = -1;
    }
    for (int l = 32; l >= 4; --l) {  // prefer wide groups (longer contiguous stores) unless a narrower one is >3% better
        const int g = kThreads / l;
        const double steps = static_cast<double>((n + l - 1) / l) * ((n + g - 1) / g);
        const double eff = static_cast<double>(n) * n / (steps * kThreads);
        if (eff > best * 1.03) best = eff, lpc = l, groups = g;
    }
}

template <int MODE, bool SWEEP, bool CODE8>
static int launch_t(Ctx* c, const MinkParams& p, int grid, size_t smem) {
    auto kern = k_minksweep3d<MODE, SWEEP, CODE8>;
    HRMGPU_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    kern<<<grid, kThreads, smem, c->stream>>>(p);
    HRMGPU_CUDA(c, cudaGetLastError());
    c->launches += 1;
    return HRMGPU_OK;
}

int launch_minksweep3d(Ctx* c, int K, int B, int n_first_obj, int n_objs, const double* d_semi, const double* d_R,
                       const double* d_off, int precision, char* d_mesh, double* d_lo, double* d_up, bool do_sweep, bool reload) {
    MinkParams p;
    p.obj = c->obj3.p;
    p.tab = c->tab.p;
    p.semi = d_semi;
    p.R = d_R;
    p.off = d_off;
    p.txs = c->txs.p;
    p.tys = c->tys.p;
    p.mesh = d_mesh;
    p.lo = d_lo;
    p.up = d_up;
    p.counters = c->counters.p;
    p.dbg = c->dbg.p && c->dbg.n >= static_cast<size_t>(K) * n_objs * B * 8 ? c->dbg.p : nullptr;
    p.n = c->n;
    p.M = c->n * c->n;
    p.B = B;
    p.first_obj = n_first_obj;
    p.n_objs = n_objs;
    p.n_heavy = (n_first_obj < c->n_arena) ? ((c->n_arena - n_first_obj < n_objs) ? c->n_arena - n_first_obj : n_objs) : 0;  // arena objects in this launch
    p.Lx = c->Lx;
    p.Ly = c->Ly;
    p.L = c->Lx * c->Ly;
    pick_mapping(c->n, p.lpc, p.groups, p.paired);
    p.reload = reload ? 1 : 0;
    p.semi_per_layer = c->semi_per_layer ? 1 : 0;
    p.seg_cap = (c->cell_seg_cap > 0 && c->cell_seg_cap < kCellCap / (kThreads / 32)) ? c->cell_seg_cap : kCellCap / (kThreads / 32);
    const double dx = (c->lim[1] - c->lim[0]) / static_cast<double>(c->Lx - 1);
    const double dy = (c->lim[3] - c->lim[2]) / static_cast<double>(c->Ly - 1);
    p.inv_dx = 1.0 / dx;
    p.inv_dy = 1.0 / dy;
    p.bias_x = -c->lim[0] * p.inv_dx;
    p.bias_y = -c->lim[2] * p.inv_dy;
    p.bias_xq = p.bias_x + (103079215104.0 + 0.9999847412109375);
    p.bias_yq = p.bias_y + (103079215104.0 + 0.9999847412109375);
    p.zlow = c->lim[4];
    p.zup = c->lim[5];
    const long long grid = static_cast<long long>(K) * n_objs * B;
    if (grid <= 0) return HRMGPU_OK;
    if (grid > 0x7FFFFFFFLL) return fail(c, HRMGPU_E_CAP, "K*S exceeds the CUDA grid limit");
    // 2 x 8-bit vertex line codes when both axes have at most 127 lines (code = 2*lo + eq <= 255), else 2 x 16-bit
    const bool code8 = do_sweep && c->Lx <= 127 && c->Ly <= 127;
    // Bench-like shapes in the fast modes run on the persistent warp-specialised kernel (mink3d_ws.cu), which also fills the
    // per-line interval buckets the free-segment pass prefers; everything else on the one-surface-per-CTA kernel below.
    c->bucket_valid[c->ivl_cur] = false;
    if (d_lo == c->ivl_lo[c->ivl_cur].p && ws_supported(c, precision, do_sweep, reload, p.lpc, p.groups, p.paired)) {
        const int So = c->n_obs * B;
        const int bcap = c->bucket_cap_hook > 0 ? c->bucket_cap_hook : (So < 128 ? (So > 0 ? So : 1) : 128);
        const size_t nl = static_cast<size_t>(K) * p.L;
        int rc;
        if ((rc = ensure(c, c->bcnt[c->ivl_cur], nl))) return rc;
        if ((rc = ensure(c, c->bucket[c->ivl_cur], nl * bcap * 2))) return rc;
        HRMGPU_CUDA(c, cudaMemsetAsync(c->bcnt[c->ivl_cur].p, 0, nl * sizeof(int), c->stream));
        rc = launch_minksweep3d_ws(c, p, precision, grid, c->bcnt[c->ivl_cur].p, c->bucket[c->ivl_cur].p, bcap);
        if (rc != -1000) {
            if (rc == HRMGPU_OK) c->bucket_valid[c->ivl_cur] = true, c->bucket_cap = bcap;
            return rc;
        }
    }
    const size_t smem = smem_layout(p, precision, do_sweep, code8, c->n, c->Lx, c->Ly);
    if (smem > 227 * 1024) return fail(c, HRMGPU_E_CAP, "n_grid / sweep-line count need more than 227 KB of shared memory");
    const int g = static_cast<int>(grid);
#define HRMGPU_LAUNCH3D(M_)                                                                      \
    do {                                                                                         \
        if (!do_sweep) return launch_t<M_, false, false>(c, p, g, smem);                         \
        return code8 ? launch_t<M_, true, true>(c, p, g, smem) : launch_t<M_, true, false>(c, p, g, smem); \
    } while (0)
    switch (precision) {
        case HRMGPU_F64_STRICT:
            HRMGPU_LAUNCH3D(HRMGPU_F64_STRICT);
        case HRMGPU_F64:
            HRMGPU_LAUNCH3D(HRMGPU_F64);
        case HRMGPU_F32:
            HRMGPU_LAUNCH3D(HRMGPU_F32);
        default:
            return fail(c, HRMGPU_E_ARG, "unknown precision");
    }
#undef HRMGPU_LAUNCH3D
}

// Loads this file's kernels up front (see preload_* in internal.cuh).
template <class K>
static void touch_kernel(K k) {
    cudaFuncAttributes a;
    if (cudaFuncGetAttributes(&a, k) != cudaSuccess) cudaGetLastError();
}
template <int MODE>
static void touch_mode() {
    touch_kernel(k_minksweep3d<MODE, false, false>), touch_kernel(k_minksweep3d<MODE, true, true>), touch_kernel(k_minksweep3d<MODE, true, false>);
}
void preload_mink3d() {
    touch_mode<HRMGPU_F64_STRICT>(), touch_mode<HRMGPU_F64>(), touch_mode<HRMGPU_F32>();
    preload_mink3d_ws();
}

}  // namespace hrmgpu
