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
// Phases inside one CTA (256 threads):
//   A  every thread evaluates boundary points k = r + c*n (omega index r fastest), stores them
//      planar (x[M] y[M] z[M]) with coalesced writes, and classifies each vertex against the sweep
//      grid into a packed 2x16-bit "line code" kept in shared memory.
//   B  every thread takes mesh cells, derives from the 4 corner codes (integer SIMD min/max) the
//      exact set of sweep lines inside each of the two faces' inclusive xy bounding boxes and
//      pushes the rare (face, line) candidates into a shared-memory queue.
//   C  the queue is drained densely: triangle test per candidate, hits recorded with an ordered
//      two-slot atomicMin so that the two LOWEST face indices per line survive (the reference's
//      "first two hits in face order" rule, SURVEY.md finding 2).
//   D  per line the two surviving faces are re-evaluated for z and the [lo, up] interval is written.
#include "devmath.cuh"
#include "internal.cuh"

#include <cfloat>
#include <cmath>

namespace hrmgpu {

struct MinkParams {
    const Obj3* obj;       // [S0]
    const double* tab;     // [S0][8][n]
    const double* semi;    // [B][3]
    const double* R;       // [K][B][9]
    const double* off;     // [K][B][3]
    const double* txs;     // [Lx+2] sentinel padded
    const double* tys;     // [Ly+2]
    void* mesh;            // [K][So][3][M] Real
    double* lo;            // [K][So][L]
    double* up;
    unsigned long long* counters;
    int n, M, B, first_obj, n_objs;
    int Lx, Ly, L;
    double x0, inv_dx, y0, inv_dy;  // line-index guess
    double zlow, zup;
};

// Smallest line index i in [0, Lc] with t_i >= x (Lc if none), and whether t_i == x; packed 2*i + eq.
// ts is sentinel padded: ts[0] = -inf, ts[1+i] = t_i, ts[Lc+1] = +inf.  Exact for any rounding of the guess.
__device__ __forceinline__ uint32_t line_code(double x, const double* ts, int Lc, double t0, double inv_d) {
    const double g = (x - t0) * inv_d;
    int i = (g > 0.0) ? ((g < static_cast<double>(Lc)) ? __double2int_ru(g) : Lc) : 0;
    while (i > 0 && ts[i] >= x) --i;
    while (ts[i + 1] < x) ++i;
    const uint32_t eq = (ts[i + 1] == x) ? 1u : 0u;
    return 2u * static_cast<uint32_t>(i) + eq;
}

template <int MODE>
struct RealOf {
    using type = double;
};
template <>
struct RealOf<HRMGPU_F32> {
    using type = float;
};

template <int MODE, bool SWEEP>
__global__ void __launch_bounds__(kThreads, 2) k_minksweep3d(const MinkParams p) {
    using A = typename std::conditional<MODE == HRMGPU_F64_STRICT, Strict, Fast>::type;
    using Real = typename RealOf<MODE>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];

    const int n = p.n, M = p.M;
    const int per_layer = p.n_objs * p.B;
    const int k = blockIdx.x / per_layer;
    const int s = blockIdx.x - k * per_layer;  // surface index inside the layer's output
    const int oi = s / p.B;
    const int b = s - oi * p.B;
    const int obj_id = p.first_obj + oi;
    const int tid = threadIdx.x;

    // ---- shared memory carve-up -----------------------------------------------------------
    Real* tab = reinterpret_cast<Real*>(smem_raw);                       // [8][n]
    size_t ofs = (sizeof(Real) * 8 * n + 15) & ~size_t(15);
    double* txs = reinterpret_cast<double*>(smem_raw + ofs);             // [Lx+2]
    double* tys = txs + (SWEEP ? p.Lx + 2 : 0);                           // [Ly+2]
    ofs += SWEEP ? sizeof(double) * (p.Lx + p.Ly + 4) : 0;
    uint32_t* vcode = reinterpret_cast<uint32_t*>(smem_raw + ofs);       // [M]
    ofs += SWEEP ? ((sizeof(uint32_t) * M + 15) & ~size_t(15)) : 0;
    uint32_t* m0 = reinterpret_cast<uint32_t*>(smem_raw + ofs);          // [L]
    uint32_t* m1 = m0 + (SWEEP ? p.L : 0);                                // [L]
    ofs += SWEEP ? ((sizeof(uint32_t) * 2 * p.L + 15) & ~size_t(15)) : 0;
    uint2* queue = reinterpret_cast<uint2*>(smem_raw + ofs);             // [kQueueCap]
    ofs += SWEEP ? sizeof(uint2) * kQueueCap : 0;
    int* qcount = reinterpret_cast<int*>(smem_raw + ofs);

    // ---- prologue: tables, line coordinates, per-surface matrices -------------------------
    const double* gtab = p.tab + static_cast<size_t>(obj_id) * 8 * n;
    for (int i = tid; i < 8 * n; i += kThreads) tab[i] = static_cast<Real>(__ldg(gtab + i));
    if (SWEEP) {
        for (int i = tid; i < p.Lx + 2; i += kThreads) txs[i] = __ldg(p.txs + i);
        for (int i = tid; i < p.Ly + 2; i += kThreads) tys[i] = __ldg(p.tys + i);
        for (int i = tid; i < p.L; i += kThreads) m0[i] = kNoFace, m1[i] = kNoFace;
        if (tid == 0) *qcount = 0;
    }

    const Obj3 ob = p.obj[obj_id];
    double R2[9], T[9], M1[9], M2[9];
    {
        const double* r2 = p.R + (static_cast<size_t>(k) * p.B + b) * 9;
#pragma unroll
        for (int i = 0; i < 9; ++i) R2[i] = __ldg(r2 + i);
        const double d0 = __ldg(p.semi + 3 * b), d1 = __ldg(p.semi + 3 * b + 1), d2 = __ldg(p.semi + 3 * b + 2);
        // Tinv = (R2 * diag) * R2^T                                  SuperQuadrics.cpp:115
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j)
                T[3 * i + j] = A::mad(A::mul(R2[3 * i + 2], d2), R2[3 * j + 2],
                                      A::mad(A::mul(R2[3 * i + 1], d1), R2[3 * j + 1], A::mul(A::mul(R2[3 * i], d0), R2[3 * j])));
        // M1 = Tinv * R1 ;  M2 = ((K * Tinv) * Tinv) * R1            SuperQuadrics.cpp:137-140
        double KT[9], KTT[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) KT[i] = A::mul(ob.sign, T[i]);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                M1[3 * i + j] = A::mad(T[3 * i + 2], ob.R[6 + j], A::mad(T[3 * i + 1], ob.R[3 + j], A::mul(T[3 * i], ob.R[j])));
                KTT[3 * i + j] = A::mad(KT[3 * i + 2], T[6 + j], A::mad(KT[3 * i + 1], T[3 + j], A::mul(KT[3 * i], T[j])));
            }
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j)
                M2[3 * i + j] = A::mad(KTT[3 * i + 2], ob.R[6 + j], A::mad(KTT[3 * i + 1], ob.R[3 + j], A::mul(KTT[3 * i], ob.R[j])));
    }
    const double offx = __ldg(p.off + (static_cast<size_t>(k) * p.B + b) * 3);
    const double offy = __ldg(p.off + (static_cast<size_t>(k) * p.B + b) * 3 + 1);
    const double offz = __ldg(p.off + (static_cast<size_t>(k) * p.B + b) * 3 + 2);

    Real* mesh = reinterpret_cast<Real*>(p.mesh) + static_cast<size_t>(blockIdx.x) * 3 * M;
    Real* mx = mesh;
    Real* my = mesh + M;
    Real* mz = mesh + 2 * static_cast<size_t>(M);
    __syncthreads();

    // ---- phase A: boundary points ---------------------------------------------------------
    if constexpr (MODE == HRMGPU_F32) {
        // FP32 throughput variant: same closed form, float arithmetic, fast reciprocal square root.
        float m1f[9], m2f[9], r1f[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) m1f[i] = static_cast<float>(M1[i]), m2f[i] = static_cast<float>(M2[i]), r1f[i] = static_cast<float>(ob.R[i]);
        const float fa = static_cast<float>(ob.a), fb = static_cast<float>(ob.b), fc = static_cast<float>(ob.c);
        const float ia = 1.0f / fa, ib = 1.0f / fb, ic = 1.0f / fc;
        const float fpx = static_cast<float>(ob.px - offx), fpy = static_cast<float>(ob.py - offy), fpz = static_cast<float>(ob.pz - offz);
        int c = tid / n, r = tid - c * n;
        const int dc = kThreads / n, dr = kThreads - dc * n;
        for (int kk = tid; kk < M; kk += kThreads) {
            const float ce1 = tab[c], se1 = tab[n + c], ce2 = tab[2 * n + c], se2 = tab[3 * n + c];
            const float cw1 = tab[4 * n + r], sw1 = tab[5 * n + r], cw2 = tab[6 * n + r], sw2 = tab[7 * n + r];
            const float X = fa * (ce1 * cw1), Y = fb * (ce1 * sw1), Z = fc * se1;
            const float gx = ce2 * cw2 * ia, gy = ce2 * sw2 * ib, gz = se2 * ic;
            const float ux = m1f[0] * gx + m1f[1] * gy + m1f[2] * gz;
            const float uy = m1f[3] * gx + m1f[4] * gy + m1f[5] * gz;
            const float uz = m1f[6] * gx + m1f[7] * gy + m1f[8] * gz;
            const float inv = rsqrtf(ux * ux + uy * uy + uz * uz);
            const float wx = m2f[0] * gx + m2f[1] * gy + m2f[2] * gz;
            const float wy = m2f[3] * gx + m2f[4] * gy + m2f[5] * gz;
            const float wz = m2f[6] * gx + m2f[7] * gy + m2f[8] * gz;
            const float ox = r1f[0] * X + r1f[1] * Y + r1f[2] * Z + fpx + wx * inv;
            const float oy = r1f[3] * X + r1f[4] * Y + r1f[5] * Z + fpy + wy * inv;
            const float oz = r1f[6] * X + r1f[7] * Y + r1f[8] * Z + fpz + wz * inv;
            reinterpret_cast<float*>(mx)[kk] = ox;
            reinterpret_cast<float*>(my)[kk] = oy;
            reinterpret_cast<float*>(mz)[kk] = oz;
            if (SWEEP) {
                const uint32_t cx = line_code(static_cast<double>(ox), txs, p.Lx, p.x0, p.inv_dx);
                const uint32_t cy = line_code(static_cast<double>(oy), tys, p.Ly, p.y0, p.inv_dy);
                vcode[kk] = cx | (cy << 16);
            }
            r += dr;
            c += dc;
            if (r >= n) r -= n, ++c;
        }
    } else {
        const double* dtab = reinterpret_cast<const double*>(tab);
        double ia = 0, ib = 0, ic = 0;
        if constexpr (MODE == HRMGPU_F64) ia = 1.0 / ob.a, ib = 1.0 / ob.b, ic = 1.0 / ob.c;
        int c = tid / n, r = tid - c * n;
        const int dc = kThreads / n, dr = kThreads - dc * n;
        for (int kk = tid; kk < M; kk += kThreads) {
            const double ce1 = dtab[c], se1 = dtab[n + c], ce2 = dtab[2 * n + c], se2 = dtab[3 * n + c];
            const double cw1 = dtab[4 * n + r], sw1 = dtab[5 * n + r], cw2 = dtab[6 * n + r], sw2 = dtab[7 * n + r];
            // getOriginShape: canonical point, rotate, translate                 SuperQuadrics.cpp:58-78
            const double X = A::mul(ob.a, A::mul(ce1, cw1));
            const double Y = A::mul(ob.b, A::mul(ce1, sw1));
            const double Z = A::mul(ob.c, se1);
            const double x0 = A::add(A::mad(ob.R[2], Z, A::mad(ob.R[1], Y, A::mul(ob.R[0], X))), ob.px);
            const double y0 = A::add(A::mad(ob.R[5], Z, A::mad(ob.R[4], Y, A::mul(ob.R[3], X))), ob.py);
            const double z0 = A::add(A::mad(ob.R[8], Z, A::mad(ob.R[7], Y, A::mul(ob.R[6], X))), ob.pz);
            // gradient                                                          SuperQuadrics.cpp:118-135
            double gx, gy, gz;
            if constexpr (MODE == HRMGPU_F64_STRICT) {
                gx = A::div(A::mul(ce2, cw2), ob.a);
                gy = A::div(A::mul(ce2, sw2), ob.b);
                gz = A::div(se2, ob.c);
            } else {
                gx = ce2 * cw2 * ia;
                gy = ce2 * sw2 * ib;
                gz = se2 * ic;
            }
            const double ux = A::mad(M1[2], gz, A::mad(M1[1], gy, A::mul(M1[0], gx)));
            const double uy = A::mad(M1[5], gz, A::mad(M1[4], gy, A::mul(M1[3], gx)));
            const double uz = A::mad(M1[8], gz, A::mad(M1[7], gy, A::mul(M1[6], gx)));
            const double wx = A::mad(M2[2], gz, A::mad(M2[1], gy, A::mul(M2[0], gx)));
            const double wy = A::mad(M2[5], gz, A::mad(M2[4], gy, A::mul(M2[3], gx)));
            const double wz = A::mad(M2[8], gz, A::mad(M2[7], gy, A::mul(M2[6], gx)));
            const double nn = A::mad(uz, uz, A::mad(uy, uy, A::mul(ux, ux)));
            double ox, oy, oz;
            if constexpr (MODE == HRMGPU_F64_STRICT) {
                const double nrm = A::sqrt(nn);                                  // SuperQuadrics.cpp:137-140
                ox = A::sub(A::add(x0, A::div(wx, nrm)), offx);                  // MultiBodyTree3D.cpp:98-101
                oy = A::sub(A::add(y0, A::div(wy, nrm)), offy);
                oz = A::sub(A::add(z0, A::div(wz, nrm)), offz);
            } else {
                const double inv = rsqrt(nn);
                ox = fma(wx, inv, x0) - offx;
                oy = fma(wy, inv, y0) - offy;
                oz = fma(wz, inv, z0) - offz;
            }
            reinterpret_cast<double*>(mx)[kk] = ox;
            reinterpret_cast<double*>(my)[kk] = oy;
            reinterpret_cast<double*>(mz)[kk] = oz;
            if (SWEEP) {
                const uint32_t cx = line_code(ox, txs, p.Lx, p.x0, p.inv_dx);
                const uint32_t cy = line_code(oy, tys, p.Ly, p.y0, p.inv_dy);
                vcode[kk] = cx | (cy << 16);
            }
            r += dr;
            c += dc;
            if (r >= n) r -= n, ++c;
        }
    }
    if (!SWEEP) return;
    __syncthreads();

    // One candidate: triangle test of face f against sweep line l, ordered insertion on a hit.
    auto test_face = [&](uint32_t f, uint32_t l, double& z) -> bool {
        int v0, v1, v2;
        face_vertices(static_cast<int>(f), n, v0, v1, v2);
        const D3 t0 = {static_cast<double>(mx[v0]), static_cast<double>(my[v0]), static_cast<double>(mz[v0])};
        const D3 t1 = {static_cast<double>(mx[v1]), static_cast<double>(my[v1]), static_cast<double>(mz[v1])};
        const D3 t2 = {static_cast<double>(mx[v2]), static_cast<double>(my[v2]), static_cast<double>(mz[v2])};
        const uint32_t ix = l / p.Ly, iy = l - ix * p.Ly;
        const D3 o = {txs[ix + 1], tys[iy + 1], 0.0};
        const D3 d = {0.0, 0.0, 1.0};
        D3 pt;
        const bool hit = line_triangle<A>(o, d, t0, t1, t2, pt);
        z = pt.z;
        return hit;
    };
    auto candidate = [&](uint32_t f, uint32_t l) {
        double z;
        if (test_face(f, l, z)) insert_two_smallest(&m0[l], &m1[l], f);
    };

    // ---- phase B: face bounding boxes vs the sweep grid, via the vertex line codes -----------
    const int nc1 = n - 1;
    const int ncell = nc1 * nc1;
    for (int cell = tid; cell < ncell; cell += kThreads) {
        const int i = cell / nc1, j = cell - i * nc1;
        const int q = i * n + j;
        const uint32_t c00 = vcode[q], c01 = vcode[q + 1], c10 = vcode[q + n], c11 = vcode[q + n + 1];
        const uint32_t dmin = __vminu2(c00, c11), dmax = __vmaxu2(c00, c11);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const uint32_t cm = half ? c01 : c10;
            const uint32_t mn = __vminu2(dmin, cm), mxx = __vmaxu2(dmax, cm);
            // lo = min(2*lo+eq) >> 1 ; hi+1 = (max(2*lo+eq) + 1) >> 1  (per 16-bit half)
            const uint32_t lo2 = (mn >> 1) & 0x7FFF7FFFu;
            const uint32_t hi2 = ((mxx + 0x00010001u) >> 1) & 0x7FFF7FFFu;
            const uint32_t lox = lo2 & 0xFFFFu, loy = lo2 >> 16, hix = hi2 & 0xFFFFu, hiy = hi2 >> 16;  // [lo, hi)
            if (lox < hix && loy < hiy) {
                const uint32_t f = half ? static_cast<uint32_t>(cell + ncell) : static_cast<uint32_t>(cell);
                for (uint32_t ix = lox; ix < hix; ++ix)
                    for (uint32_t iy = loy; iy < hiy; ++iy) {
                        const uint32_t l = ix * p.Ly + iy;
                        const int pos = atomicAdd(qcount, 1);
                        if (pos < kQueueCap)
                            queue[pos] = make_uint2(f, l);
                        else
                            candidate(f, l);  // queue full (huge faces / the arena): test in place
                    }
            }
        }
    }
    __syncthreads();

    // ---- phase C: dense drain of the candidate queue ------------------------------------------
    const int nq = min(*qcount, kQueueCap);
    for (int i = tid; i < nq; i += kThreads) candidate(queue[i].x, queue[i].y);
    __syncthreads();

    // ---- phase D: intervals ----------------------------------------------------------------
    const bool is_arena = ob.sign < 0.0;
    double* lo = p.lo + static_cast<size_t>(blockIdx.x) * p.L;
    double* up = p.up + static_cast<size_t>(blockIdx.x) * p.L;
    unsigned single = 0;
    for (int l = tid; l < p.L; l += kThreads) {
        const uint32_t f0 = m0[l], f1 = m1[l];
        double a = is_arena ? p.zlow : __longlong_as_double(0x7FF8000000000000LL);
        double bb = is_arena ? p.zup : __longlong_as_double(0x7FF8000000000000LL);
        if (f1 != kNoFace) {
            double z0, z1;
            test_face(f0, l, z0);
            test_face(f1, l, z1);
            const double zl = fmin(z0, z1), zu = fmax(z0, z1);
            a = is_arena ? fmin(p.zlow, zl) : zl;      // FreeSpace3D.cpp:44-49 / 61-64
            bb = is_arena ? fmax(p.zup, zu) : zu;
        } else if (f0 != kNoFace) {
            ++single;  // exactly one hit: reference reads out of bounds; defined here as "no interval"
        }
        lo[l] = a;
        up[l] = bb;
    }
    if (single) atomicAdd(p.counters, static_cast<unsigned long long>(single));
}

static size_t smem_bytes(int MODE, bool sweep, int n, int Lx, int Ly) {
    const size_t real = (MODE == HRMGPU_F32) ? 4 : 8;
    size_t b = (real * 8 * n + 15) & ~size_t(15);
    if (sweep) {
        b += sizeof(double) * (Lx + Ly + 4);
        b += (sizeof(uint32_t) * static_cast<size_t>(n) * n + 15) & ~size_t(15);
        b += (sizeof(uint32_t) * 2 * static_cast<size_t>(Lx) * Ly + 15) & ~size_t(15);
        b += sizeof(uint2) * kQueueCap;
    }
    return b + 16;
}

template <int MODE, bool SWEEP>
static int launch_t(Ctx* c, const MinkParams& p, int grid, size_t smem) {
    auto kern = k_minksweep3d<MODE, SWEEP>;
    HRMGPU_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    kern<<<grid, kThreads, smem, c->stream>>>(p);
    HRMGPU_CUDA(c, cudaGetLastError());
    c->launches += 1;
    return HRMGPU_OK;
}

int launch_minksweep3d(Ctx* c, int K, int B, int n_first_obj, int n_objs, const double* d_semi, const double* d_R,
                       const double* d_off, int precision, char* d_mesh, double* d_lo, double* d_up, bool do_sweep) {
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
    p.n = c->n;
    p.M = c->n * c->n;
    p.B = B;
    p.first_obj = n_first_obj;
    p.n_objs = n_objs;
    p.Lx = c->Lx;
    p.Ly = c->Ly;
    p.L = c->Lx * c->Ly;
    p.x0 = c->lim[0];
    p.y0 = c->lim[2];
    const double dx = (c->lim[1] - c->lim[0]) / static_cast<double>(c->Lx - 1);
    const double dy = (c->lim[3] - c->lim[2]) / static_cast<double>(c->Ly - 1);
    p.inv_dx = 1.0 / dx;
    p.inv_dy = 1.0 / dy;
    p.zlow = c->lim[4];
    p.zup = c->lim[5];
    const long long grid = static_cast<long long>(K) * n_objs * B;
    if (grid <= 0) return HRMGPU_OK;
    if (grid > 0x7FFFFFFFLL) return fail(c, HRMGPU_E_CAP, "K*S exceeds the CUDA grid limit");
    const size_t smem = smem_bytes(precision, do_sweep, c->n, c->Lx, c->Ly);
    if (smem > 227 * 1024) return fail(c, HRMGPU_E_CAP, "n_grid / sweep-line count need more than 227 KB of shared memory");
    const int g = static_cast<int>(grid);
    switch (precision) {
        case HRMGPU_F64_STRICT:
            return do_sweep ? launch_t<HRMGPU_F64_STRICT, true>(c, p, g, smem) : launch_t<HRMGPU_F64_STRICT, false>(c, p, g, smem);
        case HRMGPU_F64:
            return do_sweep ? launch_t<HRMGPU_F64, true>(c, p, g, smem) : launch_t<HRMGPU_F64, false>(c, p, g, smem);
        case HRMGPU_F32:
            return do_sweep ? launch_t<HRMGPU_F32, true>(c, p, g, smem) : launch_t<HRMGPU_F32, false>(c, p, g, smem);
        default:
            return fail(c, HRMGPU_E_ARG, "unknown precision");
    }
}

}  // namespace hrmgpu
