// K1' + K2' : 2D twin of the fused Minkowski + sweep kernel, one CTA per (layer, curve).
//
// Replaces (reference, single-threaded FP64):
//   SuperEllipse::getOriginShape / getMinkSum2D          src/geometry/SuperEllipse.cpp:32-101
//   MultiBodyTree2D::minkSum (link offset subtraction)    src/datastructure/MultiBodyTree2D.cpp:43-65
//   FreeSpace2D::computeIntersectionInterval              src/datastructure/FreeSpace2D.cpp:14-50
//   intersectHorizontalLinePolygon2D                      src/geometry/LineIntersection.cpp:215-248
//
// Phase A: every thread evaluates boundary points of the closed curve (4 host-built power tables per
// object) into shared memory and global memory (planar x[num] y[num]).  Phase B: one thread per sweep
// line walks the polygon edges IN ORDER (closing edge last) and keeps the first two crossings, which is
// what the reference's callers consume ([0] and [1] of the returned vector).
#include "devmath.cuh"
#include "internal.cuh"

namespace hrmgpu {

struct Mink2Params {
    const Obj2* obj;      // [S0]
    const double* tab;    // [S0][4][num]: ce(th,eps) se(th,eps) ce(th,2-eps) se(th,2-eps)
    const double* semi;   // [B][2]
    const double* cs;     // [K][B][2] cos/sin of the body angle (host libm)
    const double* off;    // [K][B][2]
    const double* tys;    // [Ly+2] sentinel padded
    double* mesh;         // [K][So][2][num]
    double* lo;           // [K][So][L]
    double* up;
    unsigned long long* counters;
    int num, B, first_obj, n_objs, Ly;
    int reload;  // re-sweep: the curve is read back from `mesh` instead of being computed
    double xlow, xup;
};

template <class A, bool SWEEP>
__global__ void __launch_bounds__(128) k_minksweep2d(const Mink2Params p) {
    extern __shared__ __align__(16) unsigned char smem2[];
    double* sx = reinterpret_cast<double*>(smem2);  // [num]
    double* sy = sx + p.num;                         // [num]
    const int per_layer = p.n_objs * p.B;
    const int k = blockIdx.x / per_layer;
    const int s = blockIdx.x - k * per_layer;
    const int oi = s / p.B;
    const int b = s - oi * p.B;
    const int obj_id = p.first_obj + oi;
    const int tid = threadIdx.x, num = p.num;
    const Obj2 ob = p.obj[obj_id];

    // ---- per-curve matrices, reference association order (SuperEllipse.cpp:79-98) ----
    const double a1 = ob.a, b1 = ob.b;
    const double c1 = ob.c1, s1 = ob.s1;  // cos/sin of the object's angle (host libm)
    const double a2 = __ldg(p.semi + 2 * b), b2 = __ldg(p.semi + 2 * b + 1);
    const double c2 = __ldg(p.cs + (static_cast<size_t>(k) * p.B + b) * 2), s2 = __ldg(p.cs + (static_cast<size_t>(k) * p.B + b) * 2 + 1);
    const double r = fmin(a1, b1);
    const double d0 = A::div(a2, r), d1 = A::div(b2, r);
    // Tinv = (R2 * diag) * R2^T
    const double RD[2][2] = {{A::mul(c2, d0), A::mul(-s2, d1)}, {A::mul(s2, d0), A::mul(c2, d1)}};
    const double R2T[2][2] = {{c2, s2}, {-s2, c2}};
    double T[2][2], KA[2][2], Bm[2][2], M1[2][2], M2[2][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) T[i][j] = A::mad(RD[i][1], R2T[1][j], A::mul(RD[i][0], R2T[0][j]));
    const double R1[2][2] = {{c1, -s1}, {s1, c1}};
    const double kr = A::mul(ob.sign, r);  // K * r
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) KA[i][j] = A::mul(kr, T[i][j]);
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) Bm[i][j] = A::mad(KA[i][1], T[1][j], A::mul(KA[i][0], T[0][j]));
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            M2[i][j] = A::mad(Bm[i][1], R1[1][j], A::mul(Bm[i][0], R1[0][j]));
            M1[i][j] = A::mad(T[i][1], R1[1][j], A::mul(T[i][0], R1[0][j]));
        }
    const double q2e = A::div(2.0, ob.eps);  // 2 / eps1
    const double offx = __ldg(p.off + (static_cast<size_t>(k) * p.B + b) * 2);
    const double offy = __ldg(p.off + (static_cast<size_t>(k) * p.B + b) * 2 + 1);
    const double* tab = p.tab + static_cast<size_t>(obj_id) * 4 * num;
    double* gx = p.mesh + static_cast<size_t>(blockIdx.x) * 2 * num;
    double* gy = gx + num;

    // ---- phase A ----
    for (int i = tid; i < num; i += blockDim.x) {
        if (p.reload) {  // refinement path: classify the cached curve against the new sweep lines
            if (SWEEP) sx[i] = gx[i], sy[i] = gy[i];
            continue;
        }
        const double ce = __ldg(tab + i), se = __ldg(tab + num + i), ce2 = __ldg(tab + 2 * num + i), se2 = __ldg(tab + 3 * num + i);
        const double X = A::mul(a1, ce), Y = A::mul(b1, se);                      // SuperEllipse.cpp:42-43
        const double x0 = A::add(A::mad(-s1, Y, A::mul(c1, X)), ob.px);            // :52
        const double y0 = A::add(A::mad(c1, Y, A::mul(s1, X)), ob.py);
        const double g0 = A::mul(q2e, ce2), g1 = A::mul(q2e, se2);                // :90-93 (no 1/a, 1/b: finding 5)
        const double u0 = A::mad(M1[0][1], g1, A::mul(M1[0][0], g0)), u1 = A::mad(M1[1][1], g1, A::mul(M1[1][0], g0));
        const double w0 = A::mad(M2[0][1], g1, A::mul(M2[0][0], g0)), w1 = A::mad(M2[1][1], g1, A::mul(M2[1][0], g0));
        const double nrm = A::sqrt(A::mad(u1, u1, A::mul(u0, u0)));
        const double ox = A::sub(A::add(x0, A::div(w0, nrm)), offx);              // :95-98, MultiBodyTree2D.cpp:62
        const double oy = A::sub(A::add(y0, A::div(w1, nrm)), offy);
        gx[i] = ox;
        gy[i] = oy;
        if (SWEEP) sx[i] = ox, sy[i] = oy;
    }
    if (!SWEEP) return;
    __syncwarp();
    __syncthreads();

    // ---- phase B: first two crossings per line in edge order ----
    const bool is_arena = ob.sign < 0.0;
    double* lo = p.lo + static_cast<size_t>(blockIdx.x) * p.Ly;
    double* up = p.up + static_cast<size_t>(blockIdx.x) * p.Ly;
    unsigned single = 0;
    for (int l = tid; l < p.Ly; l += blockDim.x) {
        const double ty = __ldg(p.tys + l + 1);
        double h[2];
        int cnt = 0;
        for (int i = 0; i < num && cnt < 2; ++i) {
            const double x1 = sx[i], y1 = sy[i];
            const int i2 = (i == num - 1) ? 0 : i + 1;
            const double x2 = sx[i2], y2 = sy[i2];
            if (ty >= fmin(y1, y2) && ty <= fmax(y1, y2)) {
                const double t = A::div(A::sub(ty, y2), A::sub(y1, y2));
                if (t >= 0.0 && t <= 1.0) h[cnt++] = A::add(A::mul(t, x1), A::mul(A::sub(1.0, t), x2));
            }
        }
        double a = is_arena ? p.xlow : __longlong_as_double(0x7FF8000000000000LL);
        double bb = is_arena ? p.xup : __longlong_as_double(0x7FF8000000000000LL);
        if (cnt == 2) {
            const double zl = fmin(h[0], h[1]), zu = fmax(h[0], h[1]);
            a = is_arena ? fmin(p.xlow, zl) : zl;   // FreeSpace2D.cpp:26-30 / 43-46
            bb = is_arena ? fmax(p.xup, zu) : zu;
        } else if (cnt == 1) {
            ++single;
        }
        lo[l] = a;
        up[l] = bb;
    }
    if (single) atomicAdd(p.counters, static_cast<unsigned long long>(single));
}

int launch_minksweep2d(Ctx* c, int K, int B, int n_first_obj, int n_objs, const double* d_semi, const double* d_cs,
                       const double* d_off, int precision, char* d_mesh, double* d_lo, double* d_up, bool do_sweep, bool reload) {
    Mink2Params p;
    p.obj = c->obj2.p;
    p.tab = c->tab.p;
    p.semi = d_semi;
    p.cs = d_cs;
    p.off = d_off;
    p.tys = c->tys.p;
    p.mesh = reinterpret_cast<double*>(d_mesh);
    p.lo = d_lo;
    p.up = d_up;
    p.counters = c->counters.p;
    p.num = c->n;
    p.B = B;
    p.first_obj = n_first_obj;
    p.n_objs = n_objs;
    p.Ly = c->Ly;
    p.reload = reload ? 1 : 0;
    p.xlow = c->lim[0];
    p.xup = c->lim[1];
    const long long grid = static_cast<long long>(K) * n_objs * B;
    if (grid <= 0) return HRMGPU_OK;
    if (grid > 0x7FFFFFFFLL) return fail(c, HRMGPU_E_CAP, "K*S exceeds the CUDA grid limit");
    const size_t smem = do_sweep ? sizeof(double) * 2 * c->n : 0;
    if (smem > 200 * 1024) return fail(c, HRMGPU_E_CAP, "num too large for the 2D sweep kernel");
    const int g = static_cast<int>(grid);
    const bool strict = precision == HRMGPU_F64_STRICT;
#define HRMGPU_LAUNCH2D(AR, SW)                                                                                              \
    do {                                                                                                                     \
        auto kern = k_minksweep2d<AR, SW>;                                                                                   \
        HRMGPU_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));    \
        kern<<<g, 128, smem, c->stream>>>(p);                                                                                \
    } while (0)
    if (strict && do_sweep) HRMGPU_LAUNCH2D(Strict, true);
    else if (strict) HRMGPU_LAUNCH2D(Strict, false);
    else if (do_sweep) HRMGPU_LAUNCH2D(Fast, true);
    else HRMGPU_LAUNCH2D(Fast, false);
#undef HRMGPU_LAUNCH2D
    HRMGPU_CUDA(c, cudaGetLastError());
    c->launches += 1;
    return HRMGPU_OK;
}

// Loads this file's kernels up front (see preload_* in internal.cuh).
template <class K>
static void touch_kernel(K k) {
    cudaFuncAttributes a;
    if (cudaFuncGetAttributes(&a, k) != cudaSuccess) cudaGetLastError();
}
void preload_mink2d() {
    touch_kernel(k_minksweep2d<Strict, true>), touch_kernel(k_minksweep2d<Strict, false>);
    touch_kernel(k_minksweep2d<Fast, true>), touch_kernel(k_minksweep2d<Fast, false>);
}

}  // namespace hrmgpu
