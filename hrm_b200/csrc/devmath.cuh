// Device arithmetic policies and the line/triangle test shared by the sweep, edge and bridge kernels.
//
// STRICT policy: every FP64 operation is issued through the *_rn intrinsics, which the compiler never
// contracts into FMAs, and divisions / square roots are the IEEE-rounded ones.  Fed the same inputs in
// the same order, this reproduces an x86-64 (SSE2, no FMA) evaluation of the reference bit for bit.
// FAST policy: plain operators (FMA contraction allowed), used by the throughput variants.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace hrmgpu {

struct Strict {
    __device__ __forceinline__ static double add(double a, double b) { return __dadd_rn(a, b); }
    __device__ __forceinline__ static double sub(double a, double b) { return __dsub_rn(a, b); }
    __device__ __forceinline__ static double mul(double a, double b) { return __dmul_rn(a, b); }
    __device__ __forceinline__ static double div(double a, double b) { return __ddiv_rn(a, b); }
    __device__ __forceinline__ static double sqrt(double a) { return __dsqrt_rn(a); }
    // a*b + c as two rounded operations
    __device__ __forceinline__ static double mad(double a, double b, double c) { return __dadd_rn(__dmul_rn(a, b), c); }
};

struct Fast {
    __device__ __forceinline__ static double add(double a, double b) { return a + b; }
    __device__ __forceinline__ static double sub(double a, double b) { return a - b; }
    __device__ __forceinline__ static double mul(double a, double b) { return a * b; }
    __device__ __forceinline__ static double div(double a, double b) { return a / b; }
    __device__ __forceinline__ static double sqrt(double a) { return ::sqrt(a); }
    __device__ __forceinline__ static double mad(double a, double b, double c) { return fma(a, b, c); }
};

// Branch-free 1/sqrt(x) for normal positive x (fast modes only): hardware seed (MUFU.RSQ64H, relative error ~2^-20)
// refined by ONE third-order (Halley) step: with e = 1 - x*y^2, 1/sqrt(x) = y * (1 - e)^(-1/2) = y * (1 + e/2 + 3e^2/8 +
// 5e^3/16 + ...); keeping the first three terms leaves 5|e|^3/16 < 2^-58 -- five FP64 operations instead of the eight
// of two Newton steps.  Unlike ::rsqrt() there is no special-case branch, so independent evaluations interleave.
// x = 0 / inf / NaN give inf-or-NaN / 0-or-NaN / NaN; the boundary arithmetic never feeds those (u is a unit-scale normal).
__device__ __forceinline__ double fast_rsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double t = x * y;
    const double e = fma(-t, y, 1.0);
    const double p = fma(0.375, e, 0.5);
    return fma(y * e, p, y);
}

struct D3 {
    double x, y, z;
};

template <class A>
__device__ __forceinline__ double dot3(const D3& a, const D3& b) {
    // (a.x*b.x + a.y*b.y) + a.z*b.z
    return A::mad(a.z, b.z, A::mad(a.y, b.y, A::mul(a.x, b.x)));
}
template <class A>
__device__ __forceinline__ D3 sub3(const D3& a, const D3& b) {
    return {A::sub(a.x, b.x), A::sub(a.y, b.y), A::sub(a.z, b.z)};
}

// Strict::mad(a,b,c) computes mul then add, i.e. (a*b)+c; dot3 above therefore evaluates
// ((ax*bx) + (ay*by)) + (az*bz) only if the additions are ordered that way:
//   mad(ay,by, ax*bx) = (ay*by) + (ax*bx)  -- IEEE addition is commutative, so this equals (ax*bx)+(ay*by).
//   mad(az,bz, t)     = (az*bz) + t        -- likewise equals t + (az*bz).

// intersectLineTriangle3D (reference src/geometry/LineIntersection.cpp:126-171).
// Line: origin o, direction d.  Triangle: t0, t0+u, t0+v.  Returns hit flag; pt = intersection point.
template <class A>
__device__ __forceinline__ bool line_triangle(const D3& o, const D3& d, const D3& t0, const D3& t1, const D3& t2, D3& pt) {
    const double tol = 1e-12;
    const D3 u = sub3<A>(t1, t0);
    const D3 v = sub3<A>(t2, t0);
    // n = u x v, normalised if non-zero (Eigen normalize())
    D3 n = {A::sub(A::mul(u.y, v.z), A::mul(u.z, v.y)), A::sub(A::mul(u.z, v.x), A::mul(u.x, v.z)),
            A::sub(A::mul(u.x, v.y), A::mul(u.y, v.x))};
    const double z = dot3<A>(n, n);
    if (z > 0.0) {
        const double s = A::sqrt(z);
        n.x = A::div(n.x, s);
        n.y = A::div(n.y, s);
        n.z = A::div(n.z, s);
    }
    const double a = -dot3<A>(n, sub3<A>(o, t0));
    const double b = dot3<A>(n, d);
    const double nn = A::sqrt(dot3<A>(n, n));
    if (!((fabs(b) > tol) && (nn > tol))) return false;
    const double pos = A::div(a, b);
    pt.x = A::add(o.x, A::mul(pos, d.x));
    pt.y = A::add(o.y, A::mul(pos, d.y));
    pt.z = A::add(o.z, A::mul(pos, d.z));
    const double uu = dot3<A>(u, u);
    const double uv = dot3<A>(u, v);
    const double vv = dot3<A>(v, v);
    const D3 w = sub3<A>(pt, t0);
    const double wu = dot3<A>(u, w);
    const double wv = dot3<A>(v, w);
    const double D = A::sub(A::mul(uv, uv), A::mul(uu, vv));
    const double s = A::div(A::sub(A::mul(uv, wv), A::mul(vv, wu)), D);
    if ((s < -tol) || (s > 1.0 + tol)) return false;
    const double t = A::div(A::sub(A::mul(uv, wu), A::mul(uu, wv)), D);
    return !((t < -tol) || (A::add(s, t) > 1.0 + tol));
}

// The test above split at the point where the line comes in.  face_unit_normal() is everything that depends on the triangle
// alone and is expensive (cross product, normalisation: one square root, three divisions, and the norm check with its second
// square root); line_triangle_n() continues from there with the very same operations, so for n = face_unit_normal(t0, t1, t2)
// it returns what line_triangle() returns, bit for bit.  A face the reference leaves at `nn > tol` gets n = 0: then b = 0 and
// the test bails out at `fabs(b) > tol` instead, with the same answer.  K4 tabulates n once per face and layer (k_face_normals).
template <class A>
__device__ __forceinline__ D3 face_unit_normal(const D3& t0, const D3& t1, const D3& t2) {
    const D3 u = sub3<A>(t1, t0);
    const D3 v = sub3<A>(t2, t0);
    D3 n = {A::sub(A::mul(u.y, v.z), A::mul(u.z, v.y)), A::sub(A::mul(u.z, v.x), A::mul(u.x, v.z)),
            A::sub(A::mul(u.x, v.y), A::mul(u.y, v.x))};
    const double z = dot3<A>(n, n);
    if (z > 0.0) {
        const double s = A::sqrt(z);
        n.x = A::div(n.x, s);
        n.y = A::div(n.y, s);
        n.z = A::div(n.z, s);
    }
    const double nn = A::sqrt(dot3<A>(n, n));
    if (!(nn > 1e-12)) n = {0.0, 0.0, 0.0};
    return n;
}
template <class A>
__device__ __forceinline__ bool line_triangle_n(const D3& o, const D3& d, const D3& t0, const D3& t1, const D3& t2, const D3& n, D3& pt) {
    const double tol = 1e-12;
    const D3 u = sub3<A>(t1, t0);
    const D3 v = sub3<A>(t2, t0);
    const double a = -dot3<A>(n, sub3<A>(o, t0));
    const double b = dot3<A>(n, d);
    if (!(fabs(b) > tol)) return false;
    const double pos = A::div(a, b);
    pt.x = A::add(o.x, A::mul(pos, d.x));
    pt.y = A::add(o.y, A::mul(pos, d.y));
    pt.z = A::add(o.z, A::mul(pos, d.z));
    const double uu = dot3<A>(u, u);
    const double uv = dot3<A>(u, v);
    const double vv = dot3<A>(v, v);
    const D3 w = sub3<A>(pt, t0);
    const double wu = dot3<A>(u, w);
    const double wv = dot3<A>(v, w);
    const double D = A::sub(A::mul(uv, uv), A::mul(uu, vv));
    const double num_s = A::sub(A::mul(uv, wv), A::mul(vv, wu));
    const double num_t = A::sub(A::mul(uv, wu), A::mul(uu, wv));
    // Certain rejections without the two divisions.  With q = num / D (real quotient) and correctly rounded division, which is
    // monotone: q < -3.9e-12 gives fl(q) < -tol, and q > 1 + 4e-12 gives fl(q) > 1 + 3.9e-12 -- for s that fails `s > 1 + tol`, for t
    // it fails `s + t > 1 + tol` once s >= -tol.  Either way the reference returns false, like here.  The thresholds are products
    // with |D| rounded once (relative 1.1e-16), far inside the factor-of-two margins; |D| outside [1e-280, 1e280], NaN operands and
    // anything undecided take the divisions below.  Measured on the maze scene: both divisions run in a fraction of the warps.
    const double aD = fabs(D);
    if (aD > 1e-280 && aD < 1e280) {
        const double ns = (D < 0.0) ? -num_s : num_s, nt = (D < 0.0) ? -num_t : num_t;
        const double lo = __dmul_rn(-4e-12, aD), hi = __dmul_rn(1.0 + 8e-12, aD);
        if ((ns < lo) || (ns > hi) || (nt < lo) || (nt > hi)) return false;
    }
    const double s = A::div(num_s, D);
    if ((s < -tol) || (s > 1.0 + tol)) return false;
    const double t = A::div(num_t, D);
    return !((t < -tol) || (A::add(s, t) > 1.0 + tol));
}

// Fast-mode variant of the test above for a VERTICAL line through (x, y): for a non-degenerate triangle the 3D
// barycentric coordinates of the plane intersection equal those of (x, y) in the projected triangle, so one division
// (by the projected area n_z) replaces the normalisation, the plane solve and the two barycentric divisions.  Same
// acceptance rule (tolerance 1e-12 on s, t, s+t; |n.d| / |n| > 1e-12); results differ from the reference only at
// rounding level, i.e. decisions can differ only for lines within ~1e-12 of a triangle edge.  z = hit height.
__device__ __forceinline__ bool vertical_hit_fast(double x, double y, const D3& t0, const D3& t1, const D3& t2, double& z) {
    const double tol = 1e-12;
    const double ux = t1.x - t0.x, uy = t1.y - t0.y, uz = t1.z - t0.z;
    const double vx = t2.x - t0.x, vy = t2.y - t0.y, vz = t2.z - t0.z;
    const double nz = ux * vy - uy * vx;
    const double nx = uy * vz - uz * vy, ny = uz * vx - ux * vz;
    const double nn = fma(nz, nz, fma(ny, ny, nx * nx));
    if (!(nz * nz > (tol * tol) * nn)) return false;  // |b| = |n_z| / |n| > tol and |n| > 0
    const double inv = 1.0 / nz;
    const double px = x - t0.x, py = y - t0.y;
    const double s = (px * vy - py * vx) * inv;
    if ((s < -tol) || (s > 1.0 + tol)) return false;
    const double t = (py * ux - px * uy) * inv;
    if ((t < -tol) || (s + t > 1.0 + tol)) return false;
    z = fma(t, vz, fma(s, uz, t0.z));
    return true;
}

// Implicit mesh connectivity of getMeshFromParamSurface (reference src/geometry/MeshGenerator.cpp:105-131):
// face f < F/2 : (q, q+n, q+n+1);  face f >= F/2 : (q, q+1, q+n+1), q = i*n + j, cell = i*(n-1)+j.
__device__ __forceinline__ void face_vertices(int f, int n, int& v0, int& v1, int& v2) {
    const int nc = (n - 1) * (n - 1);
    const int cell = (f < nc) ? f : f - nc;
    // cell / (n - 1) without the integer division: (cell + 0.5) / (n - 1) is at least 0.5 / (n - 1) away from every integer, the
    // float product is off by < 1e-6 of its value (< 2^15 here), exact for n - 1 <= 2048
    const int i = (n - 1 <= 2048) ? __float2int_rd((static_cast<float>(cell) + 0.5f) * (1.0f / static_cast<float>(n - 1))) : cell / (n - 1);
    const int j = cell - i * (n - 1);
    const int q = i * n + j;
    v0 = q;
    v1 = (f < nc) ? q + n : q + 1;
    v2 = q + n + 1;
}

// Ordered two-slot insertion: after all insertions for a line, (m0, m1) hold the two smallest face
// indices that were inserted — the faces the reference's face-order loop with `break` at two hits
// would have found (LineIntersection.cpp:46-50,116-120).  Deterministic for any arrival order.
__device__ __forceinline__ void insert_two_smallest(uint32_t* m0, uint32_t* m1, uint32_t f) {
    const uint32_t old = atomicMin(m0, f);
    const uint32_t v = (old > f) ? old : f;
    if (v != 0xFFFFFFFFu) atomicMin(m1, v);
}

}  // namespace hrmgpu
