// K4 (intra-layer edge tests) and K5 (bridge-layer point tests) against the meshes / curves a build left
// resident in HBM.
//
// Replaces (reference):
//   HRM3D::isSameSliceTransitionFree + intersectLineMesh3D       src/planners/HRM3D.cpp:319-365, LineIntersection.cpp:7-54
//   HRM2D::isSameSliceTransitionFree + isIntersectSegPolygon2D   src/planners/HRM2D.cpp:215-232, LineIntersection.cpp:173-213
//   HRM3D::isPtInCFree + intersectVerticalLineMesh3D             src/planners/HRM3D.cpp:394-425, LineIntersection.cpp:56-124
//   HRM2D::isPtInCFree + intersectHorizontalLinePolygon2D        src/planners/HRM2D.cpp:267-282, LineIntersection.cpp:215-248
//
// 3D: one WARP per (query, mesh).  Lanes test 32 consecutive faces, a ballot orders the hits, and the walk stops as
// soon as two hits exist in the scanned prefix -- exactly the reference's face-order loop with `break`.
// 2D: one thread per (query, curve) walks the polygon edges in order.
#include "devmath.cuh"
#include "internal.cuh"

namespace hrmgpu {

// min / max of every coordinate row of `count` meshes (planar [dim][M]): the reference's maxCoeff()/minCoeff().
// Block b covers mesh (b % per_set) of set (b / per_set); consecutive sets are `set_stride` values apart.
template <class Real>
__global__ void k_mesh_bbox(const Real* mesh, int dim, int M, double* bbox, int per_set, size_t set_stride) {
    const int set = blockIdx.x / per_set;
    const Real* m = mesh + static_cast<size_t>(set) * set_stride + static_cast<size_t>(blockIdx.x - set * per_set) * dim * M;
    __shared__ double smin[3][32], smax[3][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int d = 0; d < dim; ++d) {
        double mn = static_cast<double>(m[static_cast<size_t>(d) * M]), mx = mn;
        for (int i = threadIdx.x; i < M; i += blockDim.x) {
            const double v = static_cast<double>(m[static_cast<size_t>(d) * M + i]);
            mn = (v < mn) ? v : mn;
            mx = (v > mx) ? v : mx;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double a = __shfl_xor_sync(0xFFFFFFFFu, mn, o), b = __shfl_xor_sync(0xFFFFFFFFu, mx, o);
            mn = (a < mn) ? a : mn;
            mx = (b > mx) ? b : mx;
        }
        if (lane == 0) smin[d][warp] = mn, smax[d][warp] = mx;
    }
    __syncthreads();
    if (threadIdx.x < dim) {
        const int d = threadIdx.x;
        double mn = smin[d][0], mx = smax[d][0];
        for (int w = 1; w < static_cast<int>(blockDim.x >> 5); ++w) {
            mn = (smin[d][w] < mn) ? smin[d][w] : mn;
            mx = (smax[d][w] > mx) ? smax[d][w] : mx;
        }
        bbox[static_cast<size_t>(blockIdx.x) * 6 + 2 * d] = mn;
        bbox[static_cast<size_t>(blockIdx.x) * 6 + 2 * d + 1] = mx;
    }
}

// xy boxes of the bands and of the faces of every mesh (3D, n - 1 <= 32): band (0, i) = the cells of parameter column i (vertices
// [i*n, (i+2)*n)), band (1, j) = the cells of parameter row j (vertices i*n + j, i*n + j + 1 for every i).  A face's own xy box --
// the inclusive per-face test of intersectVerticalLineMesh3D (LineIntersection.cpp:77-101) -- lies inside the boxes of both of its
// bands, so a vertical line outside a band's box cannot pass any face of that band.  Per mesh: [2][n-1][4] band boxes, then
// [2 (n-1)^2][4] face boxes, each xmin xmax ymin ymax (min / max of the very doubles the face test compares).
__host__ __device__ inline size_t mesh_boxes_stride(int n) { return 4 * (2 * static_cast<size_t>(n - 1) + 2 * static_cast<size_t>(n - 1) * (n - 1)); }
template <class Real>
__global__ void __launch_bounds__(128) k_mesh_bands(const Real* mesh, int n, double* boxes, int n_mesh) {
    const int m = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int M = n * n;
    const Real* mx = mesh + static_cast<size_t>(m) * 3 * M;
    const Real* my = mx + M;
    double* out = boxes + static_cast<size_t>(m) * mesh_boxes_stride(n);
    for (int r = warp; r < 2 * (n - 1); r += 4) {  // one warp per band
        const int kind = r / (n - 1), b = r - kind * (n - 1);
        double x0 = 1.0 / 0.0, x1 = -x0, y0 = x0, y1 = -x0;
        for (int t = lane; t < 2 * n; t += 32) {
            const int q = kind == 0 ? b * n + t : (t >> 1) * n + b + (t & 1);
            const double x = static_cast<double>(mx[q]), y = static_cast<double>(my[q]);
            x0 = fmin(x0, x), x1 = fmax(x1, x), y0 = fmin(y0, y), y1 = fmax(y1, y);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            x0 = fmin(x0, __shfl_xor_sync(0xFFFFFFFFu, x0, o)), x1 = fmax(x1, __shfl_xor_sync(0xFFFFFFFFu, x1, o));
            y0 = fmin(y0, __shfl_xor_sync(0xFFFFFFFFu, y0, o)), y1 = fmax(y1, __shfl_xor_sync(0xFFFFFFFFu, y1, o));
        }
        if (lane == 0) out[4 * r] = x0, out[4 * r + 1] = x1, out[4 * r + 2] = y0, out[4 * r + 3] = y1;
    }
    double* fo = out + 8 * (n - 1);
    for (int f = threadIdx.x; f < 2 * (n - 1) * (n - 1); f += blockDim.x) {  // one thread per face
        int v0, v1, v2;
        face_vertices(f, n, v0, v1, v2);
        const double xa = static_cast<double>(mx[v0]), xb = static_cast<double>(mx[v1]), xc = static_cast<double>(mx[v2]);
        const double ya = static_cast<double>(my[v0]), yb = static_cast<double>(my[v1]), yc = static_cast<double>(my[v2]);
        fo[4 * f] = fmin(xa, fmin(xb, xc)), fo[4 * f + 1] = fmax(xa, fmax(xb, xc));
        fo[4 * f + 2] = fmin(ya, fmin(yb, yc)), fo[4 * f + 3] = fmax(ya, fmax(yb, yc));
    }
}

// First two hits of line (o, d) on the implicit mesh in face order.  VERTICAL adds the inclusive per-face xy
// bounding-box test of intersectVerticalLineMesh3D.  Returns the hit count (0..2); all lanes get h0, h1.  fn = the mesh's
// tabulated face normals [F][3] or NULL.
template <class A, class Real, bool VERTICAL, bool HASN = false>
__device__ __forceinline__ int warp_first_two_hits(const Real* mesh, int n, int M, const D3& o, const D3& d, D3& h0, D3& h1, const double* fn = nullptr) {
    const int lane = threadIdx.x & 31;
    const int F = 2 * (n - 1) * (n - 1);
    const Real* mx = mesh;
    const Real* my = mesh + M;
    const Real* mz = mesh + 2 * static_cast<size_t>(M);
    int cnt = 0;
    for (int base = 0; base < F && cnt < 2; base += 32) {
        const int f = base + lane;
        bool hit = false;
        D3 pt = {0, 0, 0};
        if (f < F) {
            int v0, v1, v2;
            face_vertices(f, n, v0, v1, v2);
            const D3 t0 = {static_cast<double>(mx[v0]), static_cast<double>(my[v0]), static_cast<double>(mz[v0])};
            const D3 t1 = {static_cast<double>(mx[v1]), static_cast<double>(my[v1]), static_cast<double>(mz[v1])};
            const D3 t2 = {static_cast<double>(mx[v2]), static_cast<double>(my[v2]), static_cast<double>(mz[v2])};
            bool skip = false;
            if (VERTICAL) {  // LineIntersection.cpp:77-101
                skip = (o.x < fmin(t0.x, fmin(t1.x, t2.x))) || (o.x > fmax(t0.x, fmax(t1.x, t2.x))) ||
                       (o.y < fmin(t0.y, fmin(t1.y, t2.y))) || (o.y > fmax(t0.y, fmax(t1.y, t2.y)));
            }
            if (!skip) {
                if (HASN) {  // tabulated unit normal of the face (k_face_normals)
                    const D3 nrm = {__ldg(fn + 3 * f), __ldg(fn + 3 * f + 1), __ldg(fn + 3 * f + 2)};
                    hit = line_triangle_n<A>(o, d, t0, t1, t2, nrm, pt);
                } else {
                    hit = line_triangle<A>(o, d, t0, t1, t2, pt);
                }
            }
        }
        unsigned bal = __ballot_sync(0xFFFFFFFFu, hit);
        while (bal && cnt < 2) {
            const int src = __ffs(bal) - 1;
            const double x = __shfl_sync(0xFFFFFFFFu, pt.x, src), y = __shfl_sync(0xFFFFFFFFu, pt.y, src),
                         z = __shfl_sync(0xFFFFFFFFu, pt.z, src);
            if (cnt == 0) h0 = {x, y, z}; else h1 = {x, y, z};
            ++cnt;
            bal &= bal - 1;
        }
    }
    return cnt;
}

// Unit normals of every face of `count` meshes (see face_unit_normal): [mesh][F][3].  Block b covers mesh (b % per_set) of set
// (b / per_set), like k_mesh_bbox.
template <class A, class Real>
__global__ void __launch_bounds__(128) k_face_normals(const Real* mesh, int n, int M, double* out, int per_set, size_t set_stride) {
    const int set = blockIdx.x / per_set;
    const Real* mx = mesh + static_cast<size_t>(set) * set_stride + static_cast<size_t>(blockIdx.x - set * per_set) * 3 * M;
    const Real* my = mx + M;
    const Real* mz = my + M;
    const int F = 2 * (n - 1) * (n - 1);
    double* o = out + static_cast<size_t>(blockIdx.x) * F * 3;
    for (int f = threadIdx.x; f < F; f += blockDim.x) {
        int v0, v1, v2;
        face_vertices(f, n, v0, v1, v2);
        const D3 t0 = {static_cast<double>(mx[v0]), static_cast<double>(my[v0]), static_cast<double>(mz[v0])};
        const D3 t1 = {static_cast<double>(mx[v1]), static_cast<double>(my[v1]), static_cast<double>(mz[v1])};
        const D3 t2 = {static_cast<double>(mx[v2]), static_cast<double>(my[v2]), static_cast<double>(mz[v2])};
        const D3 nrm = face_unit_normal<A>(t0, t1, t2);
        o[3 * f] = nrm.x, o[3 * f + 1] = nrm.y, o[3 * f + 2] = nrm.z;
    }
}

struct QueryParams {
    const void* mesh;      // first mesh of the queried set, planar [dim][M] each
    const double* bbox;    // [n_mesh][6]
    const double* face_n;  // K4 3D: [set][n_mesh][F][3] unit face normals (k_face_normals) or NULL
    const double* a;       // edges: v1 [E][dim]; bridge: centres [Q][B][dim]
    const double* b;       // edges: v2 [E][dim]
    unsigned int* flag;    // [n_query] set to 1 when blocked
    unsigned long long* counters;
    const int* qset;       // multi-set calls: [n_query] mesh set (layer / layer pair) of every query; nullptr = set 0
    size_t set_stride;     // values between the first meshes of consecutive sets
    int n, M, n_mesh, n_query, B, n_obs;
};

// ---- K4 3D ----
// Work item = (edge e, mesh m), m fastest.  A warp takes 32 consecutive items: every lane forms its edge's direction and
// applies the reference's mesh-level reject (LineIntersection.cpp:12-30: bounding box, only along (nearly) axis-parallel
// directions -- which is the common case, the candidate segments of connectOneSlice run between neighbouring sweep lines);
// the surviving items are then walked by the whole warp one after the other (faces in order, first two hits).
template <class A, class Real, bool HASN>
// (8 blocks per SM, 64 registers with 56 bytes of spills: 440 / 1407 us on the cluttered / maze map against 541 / 1754 at 4 blocks (124
// registers, no spills), 483 / 1548 at 6, 453 / 1441 at 10 and 656 / 2041 at 12 -- the kernel wants warps more than registers)
__global__ void __launch_bounds__(128, 8) k_edges3d(const QueryParams p) {
    const long long warp_global = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const long long item = warp_global * 32 + lane;
    const long long items = static_cast<long long>(p.n_query) * p.n_mesh;
    bool live = false;
    int e = 0, m = 0, set = 0;
    D3 t1 = {0, 0, 0}, t2 = {0, 0, 0}, d = {0, 0, 0};
    if (item < items) {
        e = static_cast<int>(item / p.n_mesh);
        m = static_cast<int>(item - static_cast<long long>(e) * p.n_mesh);
        t1 = {p.a[3 * static_cast<size_t>(e)], p.a[3 * static_cast<size_t>(e) + 1], p.a[3 * static_cast<size_t>(e) + 2]};
        t2 = {p.b[3 * static_cast<size_t>(e)], p.b[3 * static_cast<size_t>(e) + 1], p.b[3 * static_cast<size_t>(e) + 2]};
        d = sub3<A>(t2, t1);                           // HRM3D.cpp:322-325
        const double z = dot3<A>(d, d);
        if (z > 0.0) {
            const double s = A::sqrt(z);
            d.x = A::div(d.x, s), d.y = A::div(d.y, s), d.z = A::div(d.z, s);
        }
        set = p.qset ? __ldg(p.qset + e) : 0;
        const double* bb = p.bbox + (static_cast<size_t>(set) * p.n_mesh + m) * 6;
        const double oc[3] = {t1.x, t1.y, t1.z}, dc[3] = {d.x, d.y, d.z};
        live = true;
#pragma unroll
        for (int i = 0; i < 3; ++i)
            if (fabs(dc[i]) < 1e-6 && (oc[i] > bb[2 * i + 1] || oc[i] < bb[2 * i])) live = false;
    }
    unsigned todo = __ballot_sync(0xFFFFFFFFu, live);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const int es = __shfl_sync(0xFFFFFFFFu, e, src), ms = __shfl_sync(0xFFFFFFFFu, m, src), ss = __shfl_sync(0xFFFFFFFFu, set, src);
        const D3 a1 = {__shfl_sync(0xFFFFFFFFu, t1.x, src), __shfl_sync(0xFFFFFFFFu, t1.y, src), __shfl_sync(0xFFFFFFFFu, t1.z, src)};
        const D3 a2 = {__shfl_sync(0xFFFFFFFFu, t2.x, src), __shfl_sync(0xFFFFFFFFu, t2.y, src), __shfl_sync(0xFFFFFFFFu, t2.z, src)};
        const D3 dd = {__shfl_sync(0xFFFFFFFFu, d.x, src), __shfl_sync(0xFFFFFFFFu, d.y, src), __shfl_sync(0xFFFFFFFFu, d.z, src)};
        const Real* mesh = reinterpret_cast<const Real*>(p.mesh) + static_cast<size_t>(ss) * p.set_stride + static_cast<size_t>(ms) * 3 * p.M;
        D3 h0, h1;
        const double* fn = HASN ? p.face_n + (static_cast<size_t>(ss) * p.n_mesh + ms) * (6 * static_cast<size_t>(p.n - 1) * (p.n - 1)) : nullptr;
        const int cnt = warp_first_two_hits<A, Real, false, HASN>(mesh, p.n, p.M, a1, dd, h0, h1, fn);
        if (lane == 0) {
            if (cnt == 2) {
                const double s0 = dot3<A>(sub3<A>(h0, a1), sub3<A>(h0, a2));  // HRM3D.cpp:344-351
                const double s1 = dot3<A>(sub3<A>(h1, a1), sub3<A>(h1, a2));
                if ((s0 < 0) || (s1 < 0)) p.flag[es] = 1u;
            } else if (cnt == 1) {
                atomicAdd(p.counters + 2, 1ULL);  // single hit: reference reads out of bounds; defined as not blocking
            }
        }
    }
}

// Explicit 2x2 column-pivoted QR solve standing in for Eigen's colPivHouseholderQr().solve() (LineIntersection.cpp:203).
template <class A>
__device__ __forceinline__ void solve2x2(const double Am[2][2], const double b[2], double sol[2]) {
    const double n0 = A::mad(Am[1][0], Am[1][0], A::mul(Am[0][0], Am[0][0]));
    const double n1 = A::mad(Am[1][1], Am[1][1], A::mul(Am[0][1], Am[0][1]));
    const int p0 = (n1 > n0) ? 1 : 0, p1 = 1 - p0;
    const double a0 = Am[0][p0], a1 = Am[1][p0], c0 = Am[0][p1], c1 = Am[1][p1];
    const double nrm = A::sqrt(A::mad(a1, a1, A::mul(a0, a0)));
    sol[0] = sol[1] = 0.0;
    if (nrm == 0.0) return;
    const double q0x = A::div(a0, nrm), q0y = A::div(a1, nrm);
    const double r01 = A::mad(q0y, c1, A::mul(q0x, c0));
    const double ex = A::sub(c0, A::mul(r01, q0x)), ey = A::sub(c1, A::mul(r01, q0y));
    const double r11 = A::sqrt(A::mad(ey, ey, A::mul(ex, ex)));
    const double y0 = A::mad(q0y, b[1], A::mul(q0x, b[0]));
    const double thresh = A::mul(A::mul(2.220446049250313e-16, 2.0), nrm);
    if (r11 <= thresh) {
        sol[p0] = A::div(y0, nrm);
        sol[p1] = 0.0;
        return;
    }
    const double q1x = A::div(ex, r11), q1y = A::div(ey, r11);
    const double y1 = A::mad(q1y, b[1], A::mul(q1x, b[0]));
    const double x1 = A::div(y1, r11);
    const double x0 = A::div(A::sub(y0, A::mul(r01, x1)), nrm);
    sol[p0] = x0;
    sol[p1] = x1;
}

// ---- K4 2D ----
template <class A>
__global__ void __launch_bounds__(128) k_edges2d(const QueryParams p) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= p.n_query * p.n_mesh) return;
    const int e = t / p.n_mesh, m = t - e * p.n_mesh;
    const int set = p.qset ? __ldg(p.qset + e) : 0;
    const double* sx = reinterpret_cast<const double*>(p.mesh) + static_cast<size_t>(set) * p.set_stride + static_cast<size_t>(m) * 2 * p.M;
    const double* sy = sx + p.M;
    const double px1 = p.a[2 * e], py1 = p.a[2 * e + 1], px2 = p.b[2 * e], py2 = p.b[2 * e + 1];
    double Am[2][2], b[2], sol[2];
    Am[0][0] = A::sub(px2, px1);
    Am[1][0] = A::sub(py2, py1);
    for (int i = 0; i < p.M; ++i) {
        const int i2 = (i != p.M - 1) ? i + 1 : 0;
        const double x1 = sx[i], y1 = sy[i], x2 = sx[i2], y2 = sy[i2];
        Am[0][1] = A::sub(x1, x2);
        Am[1][1] = A::sub(y1, y2);
        b[0] = A::sub(x1, px1);
        b[1] = A::sub(y1, py1);
        solve2x2<A>(Am, b, sol);
        if ((sol[0] > 0 && sol[0] < 1) && (sol[1] > 0 && sol[1] < 1)) {
            p.flag[e] = 1u;
            return;
        }
    }
}

// ---- K5 3D: warp per (pose q, body b, obstacle o); mesh of (o, b) = index o*B + b of the pair's bridge set ----
template <class A, class Real>
__global__ void __launch_bounds__(128) k_bridge3d(const QueryParams p) {
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int per_q = p.B * p.n_obs;
    if (warp_global >= p.n_query * per_q) return;
    const int q = warp_global / per_q, rem = warp_global - q * per_q;
    const int b = rem / p.n_obs, o = rem - b * p.n_obs;
    const double* v = p.a + (static_cast<size_t>(q) * p.B + b) * 3;
    const D3 org = {v[0], v[1], v[2]};
    const D3 dir = {0.0, 0.0, 1.0};
    const int set = p.qset ? __ldg(p.qset + q) : 0;
    const Real* mesh = reinterpret_cast<const Real*>(p.mesh) + static_cast<size_t>(set) * p.set_stride + (static_cast<size_t>(o) * p.B + b) * 3 * p.M;
    D3 h0, h1;
    const int cnt = warp_first_two_hits<A, Real, true>(mesh, p.n, p.M, org, dir, h0, h1);
    if (lane == 0) {
        if (cnt == 2) {
            if (org.z > fmin(h0.z, h1.z) && org.z < fmax(h0.z, h1.z)) p.flag[q] = 1u;  // HRM3D.cpp:409-413
        } else if (cnt == 1) {
            atomicAdd(p.counters + 3, 1ULL);
        }
    }
}

// ---- K5 2D: thread per (pose q, body b, obstacle o) ----
template <class A>
__global__ void __launch_bounds__(128) k_bridge2d(const QueryParams p) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int per_q = p.B * p.n_obs;
    if (t >= p.n_query * per_q) return;
    const int q = t / per_q, rem = t - q * per_q;
    const int b = rem / p.n_obs, o = rem - b * p.n_obs;
    const double* v = p.a + (static_cast<size_t>(q) * p.B + b) * 2;
    const int set = p.qset ? __ldg(p.qset + q) : 0;
    const double* sx = reinterpret_cast<const double*>(p.mesh) + static_cast<size_t>(set) * p.set_stride + (static_cast<size_t>(o) * p.B + b) * 2 * p.M;
    const double* sy = sx + p.M;
    const double ty = v[1];
    double h[2];
    int cnt = 0;
    for (int i = 0; i < p.M && cnt < 2; ++i) {
        const int i2 = (i == p.M - 1) ? 0 : i + 1;
        const double x1 = sx[i], y1 = sy[i], x2 = sx[i2], y2 = sy[i2];
        if (ty >= fmin(y1, y2) && ty <= fmax(y1, y2)) {
            const double tt = A::div(A::sub(ty, y2), A::sub(y1, y2));
            if (tt >= 0.0 && tt <= 1.0) h[cnt++] = A::add(A::mul(tt, x1), A::mul(A::sub(1.0, tt), x2));
        }
    }
    if (cnt == 2) {
        if (v[0] > fmin(h[0], h[1]) && v[0] < fmax(h[0], h[1])) p.flag[q] = 1u;  // HRM2D.cpp:273-277
    } else if (cnt == 1) {
        atomicAdd(p.counters + 3, 1ULL);
    }
}

// ---- K5 over candidate vertex pairs: isMultiSliceTransitionFree (HRM3D.cpp:367-392, HRM2D.cpp:235-265) -------------------
// The robot moves from v0[c] to v1[c] in T interpolated poses.  Pose s has base centre t = w0[s]*v0 + w1[s]*v1 (two products,
// one sum, each rounded, like InterpolateSE3.cpp:5-29 on the host) and body b > 0 sits at link_off[pair][s][b] + t
// (MultiBodyTree3D.cpp:34-53: R_s * t_link + t; the interpolated rotation is the same for every candidate of a layer pair).
// Work item = (candidate c, pose s, body b, obstacle o), o fastest.
struct TransParams {
    const void* mesh;        // bridge meshes [P][n_obs][B][dim][M]
    const double* bbox;      // [P][n_obs * B][6] (3D)
    const double* bands;     // per mesh band + face boxes (3D, n - 1 <= 32) or NULL: k_mesh_bands
    const int* pair;         // [C]
    const double* v0;        // [C][dim]
    const double* v1;        // [C][dim]
    const double* w0;        // [T]
    const double* w1;        // [T]
    const double* link_off;  // [P][T][B][dim]
    const double* link_off2; // [P][T][B][dim] or NULL: articulated robots, centre = link_off + (link_off2 + t)
    unsigned int* flag;      // [C] set to 1 when some pose is blocked
    unsigned long long* counters;
    size_t set_stride;
    long long items;         // C * T * B * n_obs
    int n, M, C, T, B, n_obs;
};

// 3D: a warp takes 32 consecutive items.  Every lane forms its item's body centre and applies the mesh-level xy bounding-box
// reject of intersectVerticalLineMesh3D (LineIntersection.cpp:60-67), which removes most (centre, obstacle) combinations;
// the survivors are then walked by the whole warp one after the other (faces in order, first two hits).
#ifdef HRMGPU_K5_STATS
__device__ unsigned long long g_k5_stats[8];
#endif
// (6 blocks per SM: 80 registers; measured on the maze scene 1.74 ms against 1.97 ms unbounded (106), 1.76 ms at 7 and 1.88 ms at 8 blocks, which spill)
template <class A, class Real>
__global__ void __launch_bounds__(128, 6) k_transitions3d(const TransParams p) {
    const long long warp_global = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const long long item = warp_global * 32 + lane;
    bool live = false;
    int c = 0, set = 0, m = 0;
    unsigned mi = 0u, mj = 0u;
    double ox = 0, oy = 0, oz = 0;
    if (item < p.items) {
        const int per_pose = p.B * p.n_obs;
        const long long pose = item / per_pose;
        const int rem = static_cast<int>(item - pose * per_pose);
        const int b = rem / p.n_obs, o = rem - b * p.n_obs;
        c = static_cast<int>(pose / p.T);
        const int s = static_cast<int>(pose - static_cast<long long>(c) * p.T);
        if (*reinterpret_cast<volatile unsigned int*>(p.flag + c) == 0u) {  // an earlier pose already blocked this candidate: nothing to add
            set = __ldg(p.pair + c);
            const double w0 = __ldg(p.w0 + s), w1 = __ldg(p.w1 + s);
            const double* a0 = p.v0 + static_cast<size_t>(c) * 3;
            const double* a1 = p.v1 + static_cast<size_t>(c) * 3;
            ox = __dadd_rn(__dmul_rn(w0, __ldg(a0)), __dmul_rn(w1, __ldg(a1)));
            oy = __dadd_rn(__dmul_rn(w0, __ldg(a0 + 1)), __dmul_rn(w1, __ldg(a1 + 1)));
            oz = __dadd_rn(__dmul_rn(w0, __ldg(a0 + 2)), __dmul_rn(w1, __ldg(a1 + 2)));
            if (b > 0) {
                const size_t li = ((static_cast<size_t>(set) * p.T + s) * p.B + b) * 3;
                if (p.link_off2) {  // gBase * FK * tf_i: the chain's translation is added to the base position first (MultiBodyTree3D.cpp:63-89)
                    const double* l2 = p.link_off2 + li;
                    ox = __dadd_rn(__ldg(l2), ox), oy = __dadd_rn(__ldg(l2 + 1), oy), oz = __dadd_rn(__ldg(l2 + 2), oz);
                }
                const double* lo = p.link_off + li;
                ox = __dadd_rn(__ldg(lo), ox), oy = __dadd_rn(__ldg(lo + 1), oy), oz = __dadd_rn(__ldg(lo + 2), oz);
            }
            m = o * p.B + b;
            const double* bb = p.bbox + (static_cast<size_t>(set) * per_pose + m) * 6;
            live = !(ox > __ldg(bb + 1) || ox < __ldg(bb) || oy > __ldg(bb + 3) || oy < __ldg(bb + 2));
            if (live && p.bands) {  // which column / row bands of the mesh the vertical line can pass at all
                const double2* bd = reinterpret_cast<const double2*>(p.bands + (static_cast<size_t>(set) * per_pose + m) * mesh_boxes_stride(p.n));
                const int nm1 = p.n - 1;
#pragma unroll 3
                for (int t = 0; t < nm1; ++t) {
                    const double2 ux = __ldg(bd + 2 * t), uy = __ldg(bd + 2 * t + 1);
                    const double2 wx = __ldg(bd + 2 * (nm1 + t)), wy = __ldg(bd + 2 * (nm1 + t) + 1);
                    const bool in_i = !((ox < ux.x) | (ox > ux.y) | (oy < uy.x) | (oy > uy.y));
                    const bool in_j = !((ox < wx.x) | (ox > wx.y) | (oy < wy.x) | (oy > wy.y));
                    mi |= in_i ? 1u << t : 0u;
                    mj |= in_j ? 1u << t : 0u;
                }
                live = mi != 0u && mj != 0u;
            }
        }
    }
    unsigned todo = __ballot_sync(0xFFFFFFFFu, live);
#ifdef HRMGPU_K5_STATS
    {
        const bool skipped = item < p.items && *reinterpret_cast<volatile unsigned int*>(p.flag + c) != 0u;
        const int nf = live ? 2 * __popc(mi) * __popc(mj) : 0;
        atomicAdd(g_k5_stats + 0, 1ULL);
        atomicAdd(g_k5_stats + 1, live ? 1ULL : 0ULL);
        atomicAdd(g_k5_stats + 2, static_cast<unsigned long long>(nf));
        atomicAdd(g_k5_stats + 3, static_cast<unsigned long long>((nf + 31) / 32));
        atomicAdd(g_k5_stats + 4, skipped ? 1ULL : 0ULL);
        if (lane == 0) atomicAdd(g_k5_stats + 5, 1ULL);
        int tot = nf;
        for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xFFFFFFFFu, tot, o);
        if (lane == 0) atomicAdd(g_k5_stats + 6, static_cast<unsigned long long>((tot + 31) / 32));
    }
#endif
    const unsigned full = 0xFFFFFFFFu;
    if (p.bands) {
        // Two stages, both with the whole warp.  (1) Per surviving item, the faces of its bands take the per-face xy box test (two 16-byte
        // loads of the tabulated box, four comparisons); the (owner lane, face) pairs that pass are appended -- item by item, faces
        // ascending -- to a queue in shared memory.  (2) Whenever 32 pairs wait, every lane runs ONE line/triangle test (the expensive
        // part: ~250 FP64 instructions with 7 divisions / square roots in strict mode) and the hit heights go to their owner's slots in
        // queue order, which keeps the first two of every item in face order.  Testing item by item instead leaves 2-3 of 32 lanes busy
        // in the triangle test: 1.43 ms against 0.65 ms on the cluttered map, 3.98 against 1.60 ms on the dense maze.
        // Statistics of a launch (items, survivors, candidate faces, iterations): build with -DHRMGPU_K5_STATS (tools/build_variant.sh).
        __shared__ unsigned short s_queue[4][64];
        __shared__ unsigned char s_tab[4][64];
        unsigned short* queue = s_queue[(threadIdx.x >> 5) & 3];
        unsigned char* tab = s_tab[(threadIdx.x >> 5) & 3];
        const unsigned lt = (1u << lane) - 1u;
        const int nm1 = p.n - 1, nc = nm1 * nm1;
        __shared__ int s_cnt_all[4][32];
        __shared__ double s_hz_all[4][64];
        int* s_cnt = s_cnt_all[(threadIdx.x >> 5) & 3];  // hits of lane l's item so far (0..2) and the heights of the first two
        double* s_hz = s_hz_all[(threadIdx.x >> 5) & 3];
        s_cnt[lane] = 0;
        __syncwarp();
        int qn = 0;
        auto drain = [&](int take) {
            int owner = lane, f = 0;
            if (lane < take) {
                const unsigned e = queue[lane];
                owner = static_cast<int>(e >> 11), f = static_cast<int>(e & 2047u);
            }
            const D3 org = {__shfl_sync(full, ox, owner), __shfl_sync(full, oy, owner), __shfl_sync(full, oz, owner)};
            const int ss = __shfl_sync(full, set, owner), ms = __shfl_sync(full, m, owner);
            bool hit = false;
            D3 pt = {0, 0, 0};
            if (lane < take) {
                const Real* mx = reinterpret_cast<const Real*>(p.mesh) + static_cast<size_t>(ss) * p.set_stride + static_cast<size_t>(ms) * 3 * p.M;
                const Real* my = mx + p.M;
                const Real* mz = my + p.M;
                int v0, v1, v2;
                face_vertices(f, p.n, v0, v1, v2);
                const D3 t0 = {static_cast<double>(mx[v0]), static_cast<double>(my[v0]), static_cast<double>(mz[v0])};
                const D3 t1 = {static_cast<double>(mx[v1]), static_cast<double>(my[v1]), static_cast<double>(mz[v1])};
                const D3 t2 = {static_cast<double>(mx[v2]), static_cast<double>(my[v2]), static_cast<double>(mz[v2])};
                const D3 dir = {0.0, 0.0, 1.0};
                hit = line_triangle<A>(org, dir, t0, t1, t2, pt);
            }
            const unsigned bal = __ballot_sync(full, hit);
            if (bal) {
                // The hits of one item sit in consecutive lanes (the queue is item-major): a hit lane's rank among the hits of its
                // owner, added to the owner's count so far, is its place in face order; the first two are kept (the reference stops at
                // the second hit: LineIntersection.cpp:116-120).  No loop over the hits: they are most of the queued pairs.
                const unsigned group = __match_any_sync(full, owner);
                const unsigned mine = bal & group;
                const int before = hit ? s_cnt[owner] : 0;
                __syncwarp();
                if (hit) {
                    const int at = before + __popc(mine & lt);
                    if (at < 2) s_hz[2 * owner + at] = pt.z;
                    if ((mine & lt) == 0u) s_cnt[owner] = min(2, before + __popc(mine));
                }
                __syncwarp();
            }
            unsigned short rest = 0;
            if (take + lane < qn) rest = queue[take + lane];
            __syncwarp();
            if (take + lane < qn) queue[lane] = rest;
            __syncwarp();
            qn -= take;
        };
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const unsigned mis = __shfl_sync(full, mi, src), mjs = __shfl_sync(full, mj, src);
            const double x = __shfl_sync(full, ox, src), y = __shfl_sync(full, oy, src);
            const int ss = __shfl_sync(full, set, src), ms = __shfl_sync(full, m, src);
            const double2* fb = reinterpret_cast<const double2*>(p.bands + (static_cast<size_t>(ss) * (p.B * p.n_obs) + ms) * mesh_boxes_stride(p.n) + 8 * nm1);
            // rank -> band tables of this item: lane t owns bit t of either mask
            if ((mis >> lane) & 1u) tab[__popc(mis & lt)] = static_cast<unsigned char>(lane);
            if ((mjs >> lane) & 1u) tab[32 + __popc(mjs & lt)] = static_cast<unsigned char>(lane);
            __syncwarp();
            const int nbj = __popc(mjs), per_kind = __popc(mis) * nbj, total = 2 * per_kind;
            const float inv_nbj = 1.0f / static_cast<float>(nbj);
            for (int base = 0; base < total; base += 32) {
                const int cd = base + lane;
                bool pass = false;
                int f = 0;
                if (cd < total) {
                    const int kind = cd >= per_kind ? 1 : 0;
                    const int r = cd - kind * per_kind;
                    const int oi = __float2int_rd((static_cast<float>(r) + 0.5f) * inv_nbj);  // r / nbj: r < 1024, nbj <= 32, exact with the half offset
                    const int oj = r - oi * nbj;
                    f = kind * nc + tab[oi] * nm1 + tab[32 + oj];
                    const double2 bx = __ldg(fb + 2 * f), by = __ldg(fb + 2 * f + 1);
                    pass = !((x < bx.x) | (x > bx.y) | (y < by.x) | (y > by.y));  // LineIntersection.cpp:77-101
                }
                const unsigned bal = __ballot_sync(full, pass);
                if (pass) queue[qn + __popc(bal & lt)] = static_cast<unsigned short>((src << 11) | f);
                qn += __popc(bal);
                __syncwarp();
                if (qn >= 32) drain(32);
            }
            __syncwarp();  // (the tables are rewritten for the next item)
        }
        if (qn > 0) drain(qn);
        __syncwarp();
        if (live) {
            const int cnt = s_cnt[lane];
            const double hz0 = s_hz[2 * lane], hz1 = s_hz[2 * lane + 1];
            if (cnt == 2) {
                if (oz > fmin(hz0, hz1) && oz < fmax(hz0, hz1)) p.flag[c] = 1u;  // HRM3D.cpp:409-413
            } else if (cnt == 1) {
                atomicAdd(p.counters + 3, 1ULL);  // single hit: the reference reads out of bounds; defined as not blocking
            }
        }
        return;
    }
    while (todo) {  // meshes with more than 32 bands per direction: the plain walk
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const int cs = __shfl_sync(full, c, src), ss = __shfl_sync(full, set, src), ms = __shfl_sync(full, m, src);
        const D3 org = {__shfl_sync(full, ox, src), __shfl_sync(full, oy, src), __shfl_sync(full, oz, src)};
        const D3 dir = {0.0, 0.0, 1.0};
        const Real* mesh = reinterpret_cast<const Real*>(p.mesh) + static_cast<size_t>(ss) * p.set_stride + static_cast<size_t>(ms) * 3 * p.M;
        D3 h0, h1;
        const int cnt = warp_first_two_hits<A, Real, true>(mesh, p.n, p.M, org, dir, h0, h1);
        if (lane == 0) {
            if (cnt == 2) {
                if (org.z > fmin(h0.z, h1.z) && org.z < fmax(h0.z, h1.z)) p.flag[cs] = 1u;  // HRM3D.cpp:409-413
            } else if (cnt == 1) {
                atomicAdd(p.counters + 3, 1ULL);
            }
        }
    }
}

// 2D: thread per item
template <class A>
__global__ void __launch_bounds__(128) k_transitions2d(const TransParams p) {
    const long long item = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (item >= p.items) return;
    const int per_pose = p.B * p.n_obs;
    const long long pose = item / per_pose;
    const int rem = static_cast<int>(item - pose * per_pose);
    const int b = rem / p.n_obs, o = rem - b * p.n_obs;
    const int c = static_cast<int>(pose / p.T);
    const int s = static_cast<int>(pose - static_cast<long long>(c) * p.T);
    const int set = __ldg(p.pair + c);
    const double w0 = __ldg(p.w0 + s), w1 = __ldg(p.w1 + s);
    double vx = __dadd_rn(__dmul_rn(w0, __ldg(p.v0 + 2 * static_cast<size_t>(c))), __dmul_rn(w1, __ldg(p.v1 + 2 * static_cast<size_t>(c))));
    double vy = __dadd_rn(__dmul_rn(w0, __ldg(p.v0 + 2 * static_cast<size_t>(c) + 1)), __dmul_rn(w1, __ldg(p.v1 + 2 * static_cast<size_t>(c) + 1)));
    if (b > 0) {
        const double* lo = p.link_off + ((static_cast<size_t>(set) * p.T + s) * p.B + b) * 2;
        vx = __dadd_rn(__ldg(lo), vx), vy = __dadd_rn(__ldg(lo + 1), vy);
    }
    const double* sx = reinterpret_cast<const double*>(p.mesh) + static_cast<size_t>(set) * p.set_stride + (static_cast<size_t>(o) * p.B + b) * 2 * p.M;
    const double* sy = sx + p.M;
    double h[2];
    int cnt = 0;
    for (int i = 0; i < p.M && cnt < 2; ++i) {
        const int i2 = (i == p.M - 1) ? 0 : i + 1;
        const double x1 = sx[i], y1 = sy[i], x2 = sx[i2], y2 = sy[i2];
        if (vy >= fmin(y1, y2) && vy <= fmax(y1, y2)) {
            const double tt = A::div(A::sub(vy, y2), A::sub(y1, y2));
            if (tt >= 0.0 && tt <= 1.0) h[cnt++] = A::add(A::mul(tt, x1), A::mul(A::sub(1.0, tt), x2));
        }
    }
    if (cnt == 2) {
        if (vx > fmin(h[0], h[1]) && vx < fmax(h[0], h[1])) p.flag[c] = 1u;  // HRM2D.cpp:273-277
    } else if (cnt == 1) {
        atomicAdd(p.counters + 3, 1ULL);
    }
}

__global__ void k_flags_to_free(const unsigned int* flag, unsigned char* out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = flag[i] ? 0 : 1;
}

static int finish_flags(Ctx* c, int n, unsigned char* d_out) {
    k_flags_to_free<<<(n + 255) / 256, 256, 0, c->stream>>>(c->q_flag.p, d_out, n);
    HRMGPU_CUDA(c, cudaGetLastError());
    c->launches += 1;
    return HRMGPU_OK;
}

// k >= 0: all edges against layer k; k < 0: edge e against layer d_layer[e] (one launch for every layer of the build)
int launch_edges(Ctx* c, int k, int E, const double* d_v1, const double* d_v2, unsigned char* d_out, const int* d_layer) {
    const int So = c->S - c->Sa;
    const bool multi = k < 0;
    const int nset = multi ? c->K : 1;
    if (multi) k = 0;
    int rc;
    if ((rc = ensure(c, c->q_flag, static_cast<size_t>(E)))) return rc;
    HRMGPU_CUDA(c, cudaMemsetAsync(c->q_flag.p, 0, sizeof(unsigned int) * E, c->stream));
    if (So > 0) {
        QueryParams p{};
        p.a = d_v1;
        p.b = d_v2;
        p.flag = c->q_flag.p;
        p.counters = c->counters.p;
        p.n = c->n;
        p.M = c->M;
        p.n_mesh = So;
        p.n_query = E;
        p.B = c->B;
        p.n_obs = c->n_obs;
        p.qset = multi ? d_layer : nullptr;
        p.set_stride = static_cast<size_t>(c->S) * c->dim * c->M;
        const size_t real = (c->prec == HRMGPU_F32) ? 4 : 8;
        const char* mesh = c->mesh.p + (static_cast<size_t>(k) * c->S + c->Sa) * c->dim * c->M * real;
        p.mesh = mesh;
        const long long work = static_cast<long long>(E) * So;
        if ((work + 127) / 128 > 0x7FFFFFFFLL) return fail(c, HRMGPU_E_CAP, "test_edges: too many (edge, mesh) pairs for one launch");
        if (c->dim == 3) {
            if ((rc = ensure(c, c->bbox, static_cast<size_t>(nset) * So * 6))) return rc;
            p.bbox = c->bbox.p;
            if (c->prec == HRMGPU_F32)
                k_mesh_bbox<float><<<nset * So, 128, 0, c->stream>>>(reinterpret_cast<const float*>(mesh), 3, c->M, c->bbox.p, So, p.set_stride);
            else
                k_mesh_bbox<double><<<nset * So, 128, 0, c->stream>>>(reinterpret_cast<const double*>(mesh), 3, c->M, c->bbox.p, So, p.set_stride);
            HRMGPU_CUDA(c, cudaGetLastError());
            p.face_n = nullptr;
            const size_t fn_count = static_cast<size_t>(nset) * So * 6 * (c->n - 1) * (c->n - 1);
            if (fn_count * sizeof(double) <= (size_t(256) << 20) && !c->no_face_normals) {  // (beyond that the table would not stay in L2 anyway)
                if ((rc = ensure(c, c->face_n, fn_count))) return rc;
                if (c->prec == HRMGPU_F64_STRICT)
                    k_face_normals<Strict, double><<<nset * So, 128, 0, c->stream>>>(reinterpret_cast<const double*>(mesh), c->n, c->M, c->face_n.p, So, p.set_stride);
                else if (c->prec == HRMGPU_F64)
                    k_face_normals<Fast, double><<<nset * So, 128, 0, c->stream>>>(reinterpret_cast<const double*>(mesh), c->n, c->M, c->face_n.p, So, p.set_stride);
                else
                    k_face_normals<Fast, float><<<nset * So, 128, 0, c->stream>>>(reinterpret_cast<const float*>(mesh), c->n, c->M, c->face_n.p, So, p.set_stride);
                HRMGPU_CUDA(c, cudaGetLastError());
                c->launches += 1;
                p.face_n = c->face_n.p;
            }
            const unsigned grid = static_cast<unsigned>((work + 127) / 128);  // 32 items per warp, 4 warps per block
            if (c->prec == HRMGPU_F64_STRICT)
                p.face_n ? k_edges3d<Strict, double, true><<<grid, 128, 0, c->stream>>>(p) : k_edges3d<Strict, double, false><<<grid, 128, 0, c->stream>>>(p);
            else if (c->prec == HRMGPU_F64)
                p.face_n ? k_edges3d<Fast, double, true><<<grid, 128, 0, c->stream>>>(p) : k_edges3d<Fast, double, false><<<grid, 128, 0, c->stream>>>(p);
            else
                p.face_n ? k_edges3d<Fast, float, true><<<grid, 128, 0, c->stream>>>(p) : k_edges3d<Fast, float, false><<<grid, 128, 0, c->stream>>>(p);
            c->launches += 2;
        } else {
            p.bbox = nullptr;
            const unsigned grid = static_cast<unsigned>((work + 127) / 128);
            if (c->prec == HRMGPU_F64_STRICT)
                k_edges2d<Strict><<<grid, 128, 0, c->stream>>>(p);
            else
                k_edges2d<Fast><<<grid, 128, 0, c->stream>>>(p);
            c->launches += 1;
        }
        HRMGPU_CUDA(c, cudaGetLastError());
    }
    return finish_flags(c, E, d_out);
}

// pidx >= 0: all poses against the bridge layer of pair pidx; pidx < 0: pose q against pair d_pair[q]
int launch_bridge_test(Ctx* c, int pidx, int Q, const double* d_centres, unsigned char* d_out, const int* d_pair) {
    int rc;
    const bool multi = pidx < 0;
    if (multi) pidx = 0;
    if ((rc = ensure(c, c->q_flag, static_cast<size_t>(Q)))) return rc;
    HRMGPU_CUDA(c, cudaMemsetAsync(c->q_flag.p, 0, sizeof(unsigned int) * Q, c->stream));
    if (c->n_obs > 0) {
        QueryParams p{};
        p.a = d_centres;
        p.b = nullptr;
        p.bbox = nullptr;
        p.flag = c->q_flag.p;
        p.counters = c->counters.p;
        p.n = c->n;
        p.M = (c->dim == 3) ? c->n * c->n : c->n;
        p.n_mesh = c->n_obs * c->bridge_B;
        p.n_query = Q;
        p.B = c->bridge_B;
        p.n_obs = c->n_obs;
        p.qset = multi ? d_pair : nullptr;
        p.set_stride = static_cast<size_t>(c->n_obs) * c->bridge_B * c->dim * p.M;
        const size_t real = (c->bridge_prec == HRMGPU_F32) ? 4 : 8;
        p.mesh = c->bridge_mesh.p + static_cast<size_t>(pidx) * c->n_obs * c->bridge_B * c->dim * p.M * real;
        const long long work = static_cast<long long>(Q) * c->bridge_B * c->n_obs;
        if ((c->dim == 3 ? work * 32 + 127 : work + 127) / 128 > 0x7FFFFFFFLL) return fail(c, HRMGPU_E_CAP, "test_bridge: too many (pose, mesh) pairs for one launch");
        if (c->dim == 3) {
            const unsigned grid = static_cast<unsigned>((work * 32 + 127) / 128);
            if (c->bridge_prec == HRMGPU_F64_STRICT)
                k_bridge3d<Strict, double><<<grid, 128, 0, c->stream>>>(p);
            else if (c->bridge_prec == HRMGPU_F64)
                k_bridge3d<Fast, double><<<grid, 128, 0, c->stream>>>(p);
            else
                k_bridge3d<Fast, float><<<grid, 128, 0, c->stream>>>(p);
        } else {
            const unsigned grid = static_cast<unsigned>((work + 127) / 128);
            if (c->bridge_prec == HRMGPU_F64_STRICT)
                k_bridge2d<Strict><<<grid, 128, 0, c->stream>>>(p);
            else
                k_bridge2d<Fast><<<grid, 128, 0, c->stream>>>(p);
        }
        HRMGPU_CUDA(c, cudaGetLastError());
        c->launches += 1;
    }
    return finish_flags(c, Q, d_out);
}

// bounding boxes of all bridge meshes of the last build_bridge (3D): the mesh-level reject of the vertical-line test
int launch_bridge_bbox(Ctx* c) {
    const int total = c->P * c->n_obs * c->bridge_B;
    int rc;
    if ((rc = ensure(c, c->bridge_bbox, static_cast<size_t>(total > 0 ? total : 1) * 6))) return rc;
    if (total == 0 || c->dim != 3) return HRMGPU_OK;
    if (c->bridge_prec == HRMGPU_F32)
        k_mesh_bbox<float><<<total, 128, 0, c->stream>>>(reinterpret_cast<const float*>(c->bridge_mesh.p), 3, c->n * c->n, c->bridge_bbox.p, total, 0);
    else
        k_mesh_bbox<double><<<total, 128, 0, c->stream>>>(reinterpret_cast<const double*>(c->bridge_mesh.p), 3, c->n * c->n, c->bridge_bbox.p, total, 0);
    HRMGPU_CUDA(c, cudaGetLastError());
    c->launches += 1;
    c->bridge_bands_ok = false;
    if (c->n - 1 <= 32 && !c->no_bands) {
        if ((rc = ensure(c, c->bridge_bands, static_cast<size_t>(total) * mesh_boxes_stride(c->n)))) return rc;
        if (c->bridge_prec == HRMGPU_F32)
            k_mesh_bands<float><<<total, 128, 0, c->stream>>>(reinterpret_cast<const float*>(c->bridge_mesh.p), c->n, c->bridge_bands.p, total);
        else
            k_mesh_bands<double><<<total, 128, 0, c->stream>>>(reinterpret_cast<const double*>(c->bridge_mesh.p), c->n, c->bridge_bands.p, total);
        HRMGPU_CUDA(c, cudaGetLastError());
        c->launches += 1;
        c->bridge_bands_ok = true;
    }
    return HRMGPU_OK;
}

int launch_transitions(Ctx* c, int C, int T, const int* d_pair, const double* d_v0, const double* d_v1, const double* d_w0, const double* d_w1,
                       const double* d_link_off, unsigned char* d_out, const double* d_link_off2) {
    int rc;
    if ((rc = ensure(c, c->q_flag, static_cast<size_t>(C)))) return rc;
    HRMGPU_CUDA(c, cudaMemsetAsync(c->q_flag.p, 0, sizeof(unsigned int) * C, c->stream));
    if (c->n_obs > 0) {
        TransParams p;
        p.pair = d_pair, p.v0 = d_v0, p.v1 = d_v1, p.w0 = d_w0, p.w1 = d_w1, p.link_off = d_link_off, p.link_off2 = d_link_off2;
        p.flag = c->q_flag.p;
        p.counters = c->counters.p;
        p.bbox = c->bridge_bbox.p;
        p.bands = (c->dim == 3 && c->bridge_bands_ok) ? c->bridge_bands.p : nullptr;
        p.n = c->n;
        p.M = (c->dim == 3) ? c->n * c->n : c->n;
        p.C = C, p.T = T, p.B = c->bridge_B, p.n_obs = c->n_obs;
        p.set_stride = static_cast<size_t>(c->n_obs) * c->bridge_B * c->dim * p.M;
        p.mesh = c->bridge_mesh.p;
        p.items = static_cast<long long>(C) * T * c->bridge_B * c->n_obs;
        const long long blocks = (p.items + 127) / 128;  // 3D: 32 items per warp, 4 warps per block; 2D: one item per thread
        if (blocks > 0x7FFFFFFFLL) return fail(c, HRMGPU_E_CAP, "test_transitions: too many (pose, body, obstacle) items for one launch");
        const unsigned grid = static_cast<unsigned>(blocks);
        if (c->dim == 3) {
            if (c->bridge_prec == HRMGPU_F64_STRICT)
                k_transitions3d<Strict, double><<<grid, 128, 0, c->stream>>>(p);
            else if (c->bridge_prec == HRMGPU_F64)
                k_transitions3d<Fast, double><<<grid, 128, 0, c->stream>>>(p);
            else
                k_transitions3d<Fast, float><<<grid, 128, 0, c->stream>>>(p);
        } else {
            if (c->bridge_prec == HRMGPU_F64_STRICT)
                k_transitions2d<Strict><<<grid, 128, 0, c->stream>>>(p);
            else
                k_transitions2d<Fast><<<grid, 128, 0, c->stream>>>(p);
        }
        HRMGPU_CUDA(c, cudaGetLastError());
        c->launches += 1;
#ifdef HRMGPU_K5_STATS
        if (c->dim == 3) {
            unsigned long long h[8], z[8] = {0};
            cudaStreamSynchronize(c->stream);
            cudaMemcpyFromSymbol(h, g_k5_stats, sizeof(h));
            cudaMemcpyToSymbol(g_k5_stats, z, sizeof(z));
            fprintf(stderr, "[k5 stats] C=%d T=%d items=%llu live=%llu faces=%llu iters(per item)=%llu flagged-skips=%llu warps=%llu iters(flattened)=%llu\n", C, T, h[0], h[1],
                    h[2], h[3], h[4], h[5], h[6]);
        }
#endif
    }
    return finish_flags(c, C, d_out);
}

// Loads this file's kernels up front (see preload_* in internal.cuh).
template <class K>
static void touch_kernel(K k) {
    cudaFuncAttributes a;
    if (cudaFuncGetAttributes(&a, k) != cudaSuccess) cudaGetLastError();
}
void preload_queries() {
    touch_kernel(k_flags_to_free), touch_kernel(k_mesh_bbox<float>), touch_kernel(k_mesh_bbox<double>);
    touch_kernel(k_mesh_bands<float>), touch_kernel(k_mesh_bands<double>);
    touch_kernel(k_face_normals<Strict, double>), touch_kernel(k_face_normals<Fast, double>), touch_kernel(k_face_normals<Fast, float>);
    touch_kernel(k_edges3d<Strict, double, true>), touch_kernel(k_edges3d<Fast, double, true>), touch_kernel(k_edges3d<Fast, float, true>);
    touch_kernel(k_edges3d<Strict, double, false>), touch_kernel(k_edges3d<Fast, double, false>), touch_kernel(k_edges3d<Fast, float, false>);
    touch_kernel(k_edges2d<Strict>), touch_kernel(k_edges2d<Fast>);
    touch_kernel(k_bridge3d<Strict, double>), touch_kernel(k_bridge3d<Fast, double>), touch_kernel(k_bridge3d<Fast, float>);
    touch_kernel(k_bridge2d<Strict>), touch_kernel(k_bridge2d<Fast>);
    touch_kernel(k_transitions3d<Strict, double>), touch_kernel(k_transitions3d<Fast, double>), touch_kernel(k_transitions3d<Fast, float>);
    touch_kernel(k_transitions2d<Strict>), touch_kernel(k_transitions2d<Fast>);
}

}  // namespace hrmgpu
