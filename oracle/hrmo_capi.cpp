// ORACLE — TEST INFRASTRUCTURE ONLY. Nothing under hrm_b200/ may include, link or call this.
//
// Flat C entry points over the restatement so pytest (ctypes) and bench.py's cpu_baseline leg can
// call it.  Array conventions mirror include/hrmgpu.h so the parity tests feed both sides the same
// buffers: superquadric rows are [a,b,c,e1,e2,px,py,pz,qw,qx,qy,qz]; superellipse rows are
// [a,b,eps,px,py,angle]; surfaces are ordered j = object*B + body, arena objects first; sweep lines
// are ordered l = ix*Ly + iy; boundaries are Eigen-layout 3 x M (2 x N) column-major.
#include "hrmo_freespace.hpp"
#include "hrmo_articulated.hpp"

#include <chrono>
#include <cstring>
#include <thread>

using namespace hrmo;

static SuperQuadrics make_sq(const double* r, int n) {
    const double a[3] = {r[0], r[1], r[2]}, e[2] = {r[3], r[4]}, p[3] = {r[5], r[6], r[7]};
    return SuperQuadrics(a, e, p, Quat{r[8], r[9], r[10], r[11]}, n);
}
static SuperEllipse make_se(const double* r, int num) {
    const double a[2] = {r[0], r[1]}, p[2] = {r[3], r[4]};
    return SuperEllipse(a, r[2], p, r[5], num);
}

static MultiBodyTree3D make_robot3(int B, const double* semi, const double* quat, const double* off) {
    MultiBodyTree3D rb;
    const double e[2] = {1, 1};
    const double z[3] = {0, 0, 0};
    rb.base = SuperQuadrics(semi, e, z, Quat{quat[0], quat[1], quat[2], quat[3]}, 0);
    for (int b = 1; b < B; ++b) {
        SuperQuadrics l(semi + 3 * b, e, off + 3 * b, Quat{quat[4 * b], quat[4 * b + 1], quat[4 * b + 2], quat[4 * b + 3]}, 0);
        rb.link.push_back(l);
    }
    return rb;
}

extern "C" {

void hrmo_constants(double out[3]) {
    out[0] = PI;
    out[1] = HALF_PI;
    out[2] = EPSILON;
}

void hrmo_quat_to_rot(const double q[4], double R[9]) {
    const Mat3 m = quat_to_rot(Quat{q[0], q[1], q[2], q[3]});
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R[3 * i + j] = m.m[i][j];
}
void hrmo_rot_to_quat(const double R[9], double q[4]) {
    Mat3 m;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) m.m[i][j] = R[3 * i + j];
    const Quat r = rot_to_quat(m);
    q[0] = r.w, q[1] = r.x, q[2] = r.y, q[3] = r.z;
}
void hrmo_angle_axis_to_quat(double angle, const double axis[3], double q[4]) {
    Vec3 ax = {axis[0], axis[1], axis[2]};
    normalize(ax);
    const Quat r = angle_axis_to_quat(angle, ax);
    q[0] = r.w, q[1] = r.x, q[2] = r.y, q[3] = r.z;
}
void hrmo_slerp(const double a[4], double t, const double b[4], double out[4]) {
    const Quat r = slerp(Quat{a[0], a[1], a[2], a[3]}, t, Quat{b[0], b[1], b[2], b[3]});
    out[0] = r.w, out[1] = r.x, out[2] = r.y, out[3] = r.z;
}
double hrmo_angular_distance(const double a[4], const double b[4]) {
    return angular_distance(Quat{a[0], a[1], a[2], a[3]}, Quat{b[0], b[1], b[2], b[3]});
}
void hrmo_linspaced(int n, double lo, double hi, double* out) {
    const auto v = linspaced(n, lo, hi);
    std::memcpy(out, v.data(), sizeof(double) * n);
}
double hrmo_exponential_function(double theta, double power, int isSine) {
    return exponentialFunction(theta, power, isSine != 0);
}

// 8 power tables of one superquadric (a1/a2 hoisted form; the values the GPU scene setup consumes):
// out[0..n)=ce(eta,e1) [n..2n)=se(eta,e1) [2n..3n)=ce(eta,2-e1) [3n..4n)=se(eta,2-e1)
// [4n..5n)=ce(omega,e2) [5n..6n)=se(omega,e2) [6n..7n)=ce(omega,2-e2) [7n..8n)=se(omega,2-e2)
void hrmo_sq_tables(const double sq[12], int n, double* out) {
    const SuperQuadrics s = make_sq(sq, n);
    for (int i = 0; i < n; ++i) {
        out[0 * n + i] = exponentialFunction(s.etaLin[i], s.epsilon[0], false);
        out[1 * n + i] = exponentialFunction(s.etaLin[i], s.epsilon[0], true);
        out[2 * n + i] = exponentialFunction(s.etaLin[i], 2 - s.epsilon[0], false);
        out[3 * n + i] = exponentialFunction(s.etaLin[i], 2 - s.epsilon[0], true);
        out[4 * n + i] = exponentialFunction(s.omegaLin[i], s.epsilon[1], false);
        out[5 * n + i] = exponentialFunction(s.omegaLin[i], s.epsilon[1], true);
        out[6 * n + i] = exponentialFunction(s.omegaLin[i], 2 - s.epsilon[1], false);
        out[7 * n + i] = exponentialFunction(s.omegaLin[i], 2 - s.epsilon[1], true);
    }
}

void hrmo_origin_shape3d(const double sq[12], int n, double* out) {
    const Points p = make_sq(sq, n).getOriginShape();
    std::memcpy(out, p.p.data(), sizeof(double) * p.p.size());
}
void hrmo_minksum3d(const double sq[12], const double semiB[3], const double quatB[4], int n, int K, double* out) {
    const Points p = make_sq(sq, n).getMinkSum3D(semiB, Quat{quatB[0], quatB[1], quatB[2], quatB[3]}, K);
    std::memcpy(out, p.p.data(), sizeof(double) * p.p.size());
}
void hrmo_origin_shape2d(const double se[6], int num, double* out) {
    const Points p = make_se(se, num).getOriginShape();
    std::memcpy(out, p.p.data(), sizeof(double) * p.p.size());
}
void hrmo_minksum2d(const double se[6], const double semiB[2], double angleB, int num, int K, double* out) {
    const Points p = make_se(se, num).getMinkSum2D(semiB, angleB, K);
    std::memcpy(out, p.p.data(), sizeof(double) * p.p.size());
}

void hrmo_mvce2d(const double a[2], const double b[2], double thA, double thB, double out[3]) {
    const Ellipse2 e = getMVCE2D(a, b, thA, thB);
    out[0] = e.semi[0], out[1] = e.semi[1], out[2] = e.angle;
}
void hrmo_mvce3d(const double a[3], const double b[3], const double qa[4], const double qb[4], double out[7]) {
    const Ellipsoid3 e = getMVCE3D(a, b, Quat{qa[0], qa[1], qa[2], qa[3]}, Quat{qb[0], qb[1], qb[2], qb[3]});
    out[0] = e.semi[0], out[1] = e.semi[1], out[2] = e.semi[2];
    out[3] = e.quat.w, out[4] = e.quat.x, out[5] = e.quat.y, out[6] = e.quat.z;
}
void hrmo_tfe2d(const double a[2], double thA, double thB, int numStep, double out[3]) {
    const Ellipse2 e = getTFE2D(a, thA, thB, static_cast<size_t>(numStep));
    out[0] = e.semi[0], out[1] = e.semi[1], out[2] = e.angle;
}
void hrmo_tfe3d(const double a[3], const double qa[4], const double qb[4], int numStep, double out[7]) {
    const Ellipsoid3 e =
        getTFE3D(a, Quat{qa[0], qa[1], qa[2], qa[3]}, Quat{qb[0], qb[1], qb[2], qb[3]}, static_cast<size_t>(numStep));
    out[0] = e.semi[0], out[1] = e.semi[1], out[2] = e.semi[2];
    out[3] = e.quat.w, out[4] = e.quat.x, out[5] = e.quat.y, out[6] = e.quat.z;
}

// Vertical line (x, y) against the implicit mesh of a 3 x n^2 boundary: returns hit count (0..2) and z of each.
int hrmo_vertical_line_mesh(const double* bound, int n, double x, double y, double z0, double zhit[2]) {
    Points p;
    p.dim = 3, p.num = n * n;
    p.p.assign(bound, bound + 3 * n * n);
    const MeshMatrix m = getMeshFromParamSurface(p, n);
    const Line3D l = {{x, y, z0, 0, 0, 1}};
    const auto h = intersectVerticalLineMesh3D(l, m);
    for (size_t i = 0; i < h.size(); ++i) zhit[i] = h[i].z;
    return static_cast<int>(h.size());
}
// General line against the implicit mesh: returns hit count and the hit points.
int hrmo_line_mesh(const double* bound, int n, const double line[6], double hit[6]) {
    Points p;
    p.dim = 3, p.num = n * n;
    p.p.assign(bound, bound + 3 * n * n);
    const MeshMatrix m = getMeshFromParamSurface(p, n);
    Line3D l;
    std::memcpy(l.v, line, sizeof(l.v));
    const auto h = intersectLineMesh3D(l, m);
    for (size_t i = 0; i < h.size(); ++i) hit[3 * i] = h[i].x, hit[3 * i + 1] = h[i].y, hit[3 * i + 2] = h[i].z;
    return static_cast<int>(h.size());
}
int hrmo_horizontal_line_polygon(const double* bound, int num, double ty, double* out, int cap) {
    Points p;
    p.dim = 2, p.num = num;
    p.p.assign(bound, bound + 2 * num);
    const auto h = intersectHorizontalLinePolygon2D(ty, p);
    for (size_t i = 0; i < h.size() && static_cast<int>(i) < cap; ++i) out[i] = h[i];
    return static_cast<int>(h.size());
}
int hrmo_seg_polygon(const double* bound, int num, const double s1[2], const double s2[2]) {
    Points p;
    p.dim = 2, p.num = num;
    p.p.assign(bound, bound + 2 * num);
    return isIntersectSegPolygon2D(s1, s2, p) ? 1 : 0;
}

// Interval algebra of the restatement alone (Interval.cpp:10-100), same calling convention as oracle/ref_shim.cpp so that
// tests/test_ref_pin.py can run both on the same inputs.  op: 0 unions(a), 1 intersects(a), 2 complements(a = outer, b = inner).
int hrmo_interval_op(int op, const double* a, int na, const double* b, int nb, double* out, int cap) {
    auto make = [](const double* se, int n) {
        std::vector<Interval> v;
        for (int i = 0; i < n; ++i) v.push_back(Interval{se[2 * i], se[2 * i + 1]});
        return v;
    };
    std::vector<Interval> r;
    if (op == 0) r = unions(make(a, na));
    else if (op == 1) r = intersects(make(a, na));
    else r = complements(make(a, na), make(b, nb));
    const int n = static_cast<int>(r.size());
    for (int i = 0; i < n && i < cap; ++i) out[2 * i] = r[i].s, out[2 * i + 1] = r[i].e;
    return n;
}

// Free segments of one sweep plane from interval tables (Interval algebra + enhanceFreeSegment).
// lo/up are [Ly][Sa] arena and [Ly][So] obstacle tables. Output CSR over the Ly lines.
int hrmo_free_segments(int Ly, int Sa, int So, const double* alo, const double* aup, const double* olo, const double* oup,
                       int* off, double* xL, double* xU, double* xM, int cap) {
    std::vector<SuperQuadrics> none;
    MultiBodyTree3D rb;
    FreeSpace3D fs(rb, none, none);
    fs.intersect_.arenaLow.assign(Ly, std::vector<double>(Sa));
    fs.intersect_.arenaUpp.assign(Ly, std::vector<double>(Sa));
    fs.intersect_.obstacleLow.assign(Ly, std::vector<double>(So));
    fs.intersect_.obstacleUpp.assign(Ly, std::vector<double>(So));
    for (int i = 0; i < Ly; ++i) {
        for (int j = 0; j < Sa; ++j) fs.intersect_.arenaLow[i][j] = alo[i * Sa + j], fs.intersect_.arenaUpp[i][j] = aup[i * Sa + j];
        for (int j = 0; j < So; ++j)
            fs.intersect_.obstacleLow[i][j] = olo[i * So + j], fs.intersect_.obstacleUpp[i][j] = oup[i * So + j];
    }
    std::vector<double> ty(Ly, 0.0);
    fs.computeFreeSegment(ty);
    int tot = 0;
    for (int i = 0; i < Ly; ++i) {
        off[i] = tot;
        for (size_t j = 0; j < fs.segment_.xM[i].size(); ++j, ++tot)
            if (tot < cap) xL[tot] = fs.segment_.xL[i][j], xU[tot] = fs.segment_.xU[i][j], xM[tot] = fs.segment_.xM[i][j];
    }
    off[Ly] = tot;
    return tot;
}

// One full 3D C-layer: Minkowski boundaries -> meshes -> sweep -> free segments.
// Any output pointer may be NULL.  Returns the total segment count (or -1 on bad arguments).
int hrmo_layer3d(int n_arena, int n_obs, const double* sq, int n, int B, const double* body_semi, const double* body_quat,
                 const double* body_off, const double lim[6], int Lx, int Ly, double* bound, double* lo, double* up, int* seg_off,
                 double* xL, double* xU, double* xM, int seg_cap, long* single_hits) {
    if (n_arena < 0 || n_obs < 0 || B < 1 || n < 2 || Lx < 1 || Ly < 1) return -1;
    std::vector<SuperQuadrics> arena, obs;
    for (int i = 0; i < n_arena; ++i) arena.push_back(make_sq(sq + 12 * i, n));
    for (int i = 0; i < n_obs; ++i) obs.push_back(make_sq(sq + 12 * (n_arena + i), n));
    const MultiBodyTree3D rb = make_robot3(B, body_semi, body_quat, body_off);
    FreeSpace3D fs(rb, arena, obs);
    fs.setup(static_cast<unsigned>(Ly), lim[4], lim[5]);
    fs.computeCSpaceBoundary();
    fs.computeCSpaceBoundaryMesh(fs.cSpaceBoundary_);
    const int Sa = n_arena * B, So = n_obs * B, S = Sa + So, M = n * n;
    if (bound) {
        for (int j = 0; j < Sa; ++j) std::memcpy(bound + static_cast<size_t>(j) * 3 * M, fs.cSpaceBoundary_.arena[j].p.data(), sizeof(double) * 3 * M);
        for (int j = 0; j < So; ++j)
            std::memcpy(bound + static_cast<size_t>(Sa + j) * 3 * M, fs.cSpaceBoundary_.obstacle[j].p.data(), sizeof(double) * 3 * M);
    }
    // HRM3D.cpp:63-91 with the interval tables captured per plane
    std::vector<double> ty(Ly);
    const double dx = (lim[1] - lim[0]) / static_cast<double>(Lx - 1);
    const double dy = (lim[3] - lim[2]) / static_cast<double>(Ly - 1);
    for (int i = 0; i < Ly; ++i) ty[i] = lim[2] + static_cast<double>(i) * dy;
    int tot = 0;
    for (int ix = 0; ix < Lx; ++ix) {
        const double tx = lim[0] + static_cast<double>(ix) * dx;
        fs.computeIntersectionInterval(tx, ty);
        fs.computeFreeSegment(ty);
        for (int iy = 0; iy < Ly; ++iy) {
            const size_t l = static_cast<size_t>(ix) * Ly + iy;
            if (lo && up) {
                for (int j = 0; j < Sa; ++j) lo[l * S + j] = fs.intersect_.arenaLow[iy][j], up[l * S + j] = fs.intersect_.arenaUpp[iy][j];
                for (int j = 0; j < So; ++j)
                    lo[l * S + Sa + j] = fs.intersect_.obstacleLow[iy][j], up[l * S + Sa + j] = fs.intersect_.obstacleUpp[iy][j];
            }
            if (seg_off) seg_off[l] = tot;
            for (size_t j = 0; j < fs.segment_.xM[iy].size(); ++j, ++tot)
                if (xL && tot < seg_cap) xL[tot] = fs.segment_.xL[iy][j], xU[tot] = fs.segment_.xU[iy][j], xM[tot] = fs.segment_.xM[iy][j];
        }
    }
    if (seg_off) seg_off[static_cast<size_t>(Lx) * Ly] = tot;
    if (single_hits) *single_hits = fs.singleHitCount;
    return tot;
}

// One full 2D C-layer.  body_angle/body_off are the per-body poses (MultiBodyTree2D.cpp:43-65 applied by the caller,
// or by hrmo_robot2d_body_poses below).  lim = {xlo, xhi, ylo, yhi}.
int hrmo_layer2d(int n_arena, int n_obs, const double* se, int num, int B, const double* body_semi, const double* body_angle,
                 const double* body_off, const double lim[4], int Ly, double* bound, double* lo, double* up, int* seg_off, double* xL,
                 double* xU, double* xM, int seg_cap, long* single_hits) {
    if (n_arena < 0 || n_obs < 0 || B < 1 || num < 2 || Ly < 1) return -1;
    std::vector<SuperEllipse> arena, obs;
    for (int i = 0; i < n_arena; ++i) arena.push_back(make_se(se + 6 * i, num));
    for (int i = 0; i < n_obs; ++i) obs.push_back(make_se(se + 6 * (n_arena + i), num));
    MultiBodyTree2D rb;  // only used for sizing; boundaries are built per body below
    rb.link.resize(B - 1);
    FreeSpace2D fs(rb, arena, obs);
    fs.setup(static_cast<unsigned>(Ly), lim[0], lim[1]);
    auto mink = [&](const SuperEllipse& s, int k, std::vector<Points>& dst) {
        for (int b = 0; b < B; ++b) {
            Points p = s.getMinkSum2D(body_semi + 2 * b, body_angle[b], k);
            if (b > 0)
                for (int kk = 0; kk < p.num; ++kk) {
                    p.at(0, kk) = p.at(0, kk) - body_off[2 * b];
                    p.at(1, kk) = p.at(1, kk) - body_off[2 * b + 1];
                }
            dst.push_back(std::move(p));
        }
    };
    for (const auto& a : arena) mink(a, -1, fs.cSpaceBoundary_.arena);
    for (const auto& o : obs) mink(o, +1, fs.cSpaceBoundary_.obstacle);
    const int Sa = n_arena * B, So = n_obs * B, S = Sa + So;
    if (bound) {
        for (int j = 0; j < Sa; ++j) std::memcpy(bound + static_cast<size_t>(j) * 2 * num, fs.cSpaceBoundary_.arena[j].p.data(), sizeof(double) * 2 * num);
        for (int j = 0; j < So; ++j)
            std::memcpy(bound + static_cast<size_t>(Sa + j) * 2 * num, fs.cSpaceBoundary_.obstacle[j].p.data(), sizeof(double) * 2 * num);
    }
    const FreeSegment2D seg = sweepLineProcess2D(fs, lim, static_cast<size_t>(Ly));
    int tot = 0;
    for (int iy = 0; iy < Ly; ++iy) {
        if (lo && up) {
            for (int j = 0; j < Sa; ++j) lo[iy * S + j] = fs.intersect_.arenaLow[iy][j], up[iy * S + j] = fs.intersect_.arenaUpp[iy][j];
            for (int j = 0; j < So; ++j)
                lo[iy * S + Sa + j] = fs.intersect_.obstacleLow[iy][j], up[iy * S + Sa + j] = fs.intersect_.obstacleUpp[iy][j];
        }
        if (seg_off) seg_off[iy] = tot;
        for (size_t j = 0; j < seg.xM[iy].size(); ++j, ++tot)
            if (xL && tot < seg_cap) xL[tot] = seg.xL[iy][j], xU[tot] = seg.xU[iy][j], xM[tot] = seg.xM[iy][j];
    }
    if (seg_off) seg_off[Ly] = tot;
    if (single_hits) *single_hits = fs.singleHitCount;
    return tot;
}

// Per-body (angle, offset) of a 2D multi-body robot whose base has heading `theta`
// (MultiBodyTree2D.cpp:24-65).  robot rows: [a,b,eps,px,py,angle], first row = base, others relative.
void hrmo_robot2d_body_poses(int B, const double* robot, double theta, double* body_angle, double* body_off) {
    MultiBodyTree2D rb;
    rb.base = make_se(robot, 0);
    for (int b = 1; b < B; ++b) rb.addBody(make_se(robot + 6 * b, 0));
    Tf2 g;
    g.R[0][0] = std::cos(theta), g.R[0][1] = -std::sin(theta), g.R[1][0] = std::sin(theta), g.R[1][1] = std::cos(theta);
    g.t[0] = g.t[1] = 0.0;
    rb.robotTF(g);
    for (int b = 0; b < B; ++b) rb.bodyPose(static_cast<size_t>(b), body_angle[b], body_off + 2 * b);
}

// Per-body (quat, offset) of a 3D multi-body robot with base pose (quaternion q, position 0)
// (MultiBodyTree3D.cpp:21-53).  robot rows: [a,b,c,1,1,px,py,pz,qw,qx,qy,qz], links relative to base.
void hrmo_robot3d_body_poses(int B, const double* robot, const double q[4], double* body_quat, double* body_off) {
    MultiBodyTree3D rb;
    rb.base = make_sq(robot, 0);
    for (int b = 1; b < B; ++b) rb.addBody(make_sq(robot + 12 * b, 0));
    Tf3 g;
    g.R = quat_to_rot(Quat{q[0], q[1], q[2], q[3]});
    g.t[0] = g.t[1] = g.t[2] = 0.0;
    rb.robotTF(g);
    for (int b = 0; b < B; ++b) {
        const SuperQuadrics& s = (b == 0) ? rb.base : rb.link[b - 1];
        body_quat[4 * b] = s.quat.w, body_quat[4 * b + 1] = s.quat.x, body_quat[4 * b + 2] = s.quat.y, body_quat[4 * b + 3] = s.quat.z;
        for (int d = 0; d < 3; ++d) body_off[3 * b + d] = (b == 0) ? 0.0 : s.position[d];
    }
}

// CPU-baseline timer: builds `K` 3D C-layers (Minkowski + mesh + sweep + free segments; the region
// HighwayRoadMap-inl.h:67-69 times minus edges) with per-layer body quaternions [K][B][4] and offsets
// [K][B][3], on `threads` host threads (layers split contiguously).  Returns wall seconds; *checksum
// accumulates segment counts so the work cannot be elided.
double hrmo_time_layers3d(int n_arena, int n_obs, const double* sq, int n, int B, const double* body_semi, const double* body_quat,
                          const double* body_off, const double lim[6], int Lx, int Ly, int K, int threads, long* checksum) {
    if (threads < 1) threads = 1;
    std::vector<long> sums(static_cast<size_t>(threads), 0);
    const auto t0 = std::chrono::high_resolution_clock::now();
    auto work = [&](int t) {
        const int k0 = static_cast<int>(static_cast<long>(K) * t / threads), k1 = static_cast<int>(static_cast<long>(K) * (t + 1) / threads);
        for (int k = k0; k < k1; ++k)
            sums[t] += hrmo_layer3d(n_arena, n_obs, sq, n, B, body_semi, body_quat + static_cast<size_t>(k) * B * 4,
                                    body_off + static_cast<size_t>(k) * B * 3, lim, Lx, Ly, nullptr, nullptr, nullptr, nullptr, nullptr,
                                    nullptr, nullptr, 0, nullptr);
    };
    if (threads == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < threads; ++t) th.emplace_back(work, t);
        for (auto& x : th) x.join();
    }
    const double sec = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
    long s = 0;
    for (long v : sums) s += v;
    if (checksum) *checksum = s;
    return sec;
}

// ---- edge / bridge predicates ----------------------------------------------------------------------------------
static std::vector<MeshMatrix> meshes_from(const double* bound, int n, int count) {
    std::vector<MeshMatrix> out;
    const int M = n * n;
    for (int i = 0; i < count; ++i) {
        Points p;
        p.dim = 3, p.num = M;
        p.p.assign(bound + static_cast<size_t>(i) * 3 * M, bound + static_cast<size_t>(i + 1) * 3 * M);
        out.push_back(getMeshFromParamSurface(p, n));
    }
    return out;
}
static std::vector<Points> curves_from(const double* bound, int num, int count) {
    std::vector<Points> out;
    for (int i = 0; i < count; ++i) {
        Points p;
        p.dim = 2, p.num = num;
        p.p.assign(bound + static_cast<size_t>(i) * 2 * num, bound + static_cast<size_t>(i + 1) * 2 * num);
        out.push_back(std::move(p));
    }
    return out;
}

// isSameSliceTransitionFree for E segments against `count` C-obstacle boundaries ([count][M][3] Eigen layout).
long hrmo_edges3d(const double* bound, int n, int count, int E, const double* v1, const double* v2, unsigned char* out) {
    const auto meshes = meshes_from(bound, n, count);
    long single = 0;
    for (int e = 0; e < E; ++e) out[e] = isSameSliceTransitionFree3D(meshes, v1 + 3 * e, v2 + 3 * e, &single) ? 1 : 0;
    return single;
}
void hrmo_edges2d(const double* bound, int num, int count, int E, const double* v1, const double* v2, unsigned char* out) {
    const auto curves = curves_from(bound, num, count);
    for (int e = 0; e < E; ++e) out[e] = isSameSliceTransitionFree2D(curves, v1 + 2 * e, v2 + 2 * e) ? 1 : 0;
}
// isPtInCFree of Q points against `count` bridge C-obstacle boundaries of one body.
long hrmo_pt_in_cfree3d(const double* bound, int n, int count, int Q, const double* pts, unsigned char* out) {
    const auto meshes = meshes_from(bound, n, count);
    long single = 0;
    for (int q = 0; q < Q; ++q) out[q] = isPtInCFree3D(meshes, pts + 3 * q, &single) ? 1 : 0;
    return single;
}
long hrmo_pt_in_cfree2d(const double* bound, int num, int count, int Q, const double* pts, unsigned char* out) {
    const auto curves = curves_from(bound, num, count);
    long single = 0;
    for (int q = 0; q < Q; ++q) out[q] = isPtInCFree2D(curves, pts + 2 * q, &single) ? 1 : 0;
    return single;
}

// ---- planners ------------------------------------------------------------------------------------------------------
// HRM3D::plan on a scene given as rows (arena first).  robot rows: base first, links relative to base.
// quats [numSlice][4] (w,x,y,z).  lim[6], Lx, Ly, numPoint.  start/goal: 7 doubles (x,y,z,qw,qx,qy,qz).
// per_orientation = 0 reproduces the reference snapshot (frozen FreeSpace robot copy).
// out_info[8]: solved, n_vertex, n_edge, n_path, single_hits, n_edge_tests, -, -.   vertex [cap_v][7], edge [cap_e][2].
// Refinement rounds allowed in hrmo_plan3d / hrmo_plan2d (HighwayRoadMap::plan refines while unsolved and within the time
// limit; rounds are counted here to keep the runs deterministic).  out_info[7] = rounds done.
static int g_max_refine = 0;
void hrmo_set_max_refine(int n) { g_max_refine = n < 0 ? 0 : n; }
static int g_prob_refine_period = 60;  // ProbHRM3D.cpp:52-56
void hrmo_set_prob_refine_period(int n) { g_prob_refine_period = n < 0 ? 0 : n; }

int hrmo_plan3d(int n_arena, int n_obs, const double* sq, int n, int B, const double* robot, int numSlice, const double* quats,
                const double lim[6], int Lx, int Ly, int numPoint, const double* start, const double* goal, int per_orientation,
                long* out_info, double* vertex, long cap_v, long* edge, long cap_e, long* path, long cap_p) {
    std::vector<SuperQuadrics> arena, obs;
    for (int i = 0; i < n_arena; ++i) arena.push_back(make_sq(sq + 12 * i, n));
    for (int i = 0; i < n_obs; ++i) obs.push_back(make_sq(sq + 12 * (n_arena + i), n));
    MultiBodyTree3D rb;
    rb.base = make_sq(robot, n);
    for (int b = 1; b < B; ++b) rb.addBody(make_sq(robot + 12 * b, n));
    std::vector<Quat> qs;
    for (int i = 0; i < numSlice; ++i) qs.push_back(Quat{quats[4 * i], quats[4 * i + 1], quats[4 * i + 2], quats[4 * i + 3]});
    PlanningRequest req;
    req.parameters.boundaryLimits.assign(lim, lim + 6);
    req.parameters.numSlice = static_cast<size_t>(numSlice);
    req.parameters.numLineX = static_cast<size_t>(Lx);
    req.parameters.numLineY = static_cast<size_t>(Ly);
    req.parameters.numPoint = static_cast<size_t>(numPoint);
    req.start.assign(start, start + 7);
    req.goal.assign(goal, goal + 7);
    HRM3D hrm(rb, arena, obs, req, qs, per_orientation != 0);
    hrm.plan(static_cast<size_t>(g_max_refine));
    out_info[7] = static_cast<long>(hrm.numRefineDone);
    const auto& g = hrm.res_.graphStructure;
    out_info[0] = hrm.res_.solved ? 1 : 0;
    out_info[1] = static_cast<long>(g.vertex.size());
    out_info[2] = static_cast<long>(g.edge.size());
    out_info[3] = static_cast<long>(hrm.res_.PathId.size());
    out_info[4] = hrm.singleHitCount;
    out_info[5] = hrm.numEdgeTests;
    for (size_t i = 0; i < g.vertex.size() && static_cast<long>(i) < cap_v; ++i)
        for (int d = 0; d < 7; ++d) vertex[7 * i + d] = g.vertex[i][d];
    for (size_t i = 0; i < g.edge.size() && static_cast<long>(i) < cap_e; ++i) edge[2 * i] = static_cast<long>(g.edge[i].first), edge[2 * i + 1] = static_cast<long>(g.edge[i].second);
    for (size_t i = 0; i < hrm.res_.PathId.size() && static_cast<long>(i) < cap_p; ++i) path[i] = hrm.res_.PathId[i];
    return 0;
}

// HRM2D::plan.  se rows [a,b,eps,px,py,angle]; lim[4]; start/goal: (x, y, theta).  vertex [cap_v][3].
int hrmo_plan2d(int n_arena, int n_obs, const double* se, int num, int B, const double* robot, int numSlice, const double lim[4], int Ly,
                int numPoint, const double* start, const double* goal, int per_orientation, long* out_info, double* vertex, long cap_v,
                long* edge, long cap_e, long* path, long cap_p) {
    std::vector<SuperEllipse> arena, obs;
    for (int i = 0; i < n_arena; ++i) arena.push_back(make_se(se + 6 * i, num));
    for (int i = 0; i < n_obs; ++i) obs.push_back(make_se(se + 6 * (n_arena + i), num));
    MultiBodyTree2D rb;
    rb.base = make_se(robot, num);
    for (int b = 1; b < B; ++b) rb.addBody(make_se(robot + 6 * b, num));
    PlanningRequest req;
    req.parameters.boundaryLimits.assign(lim, lim + 4);
    req.parameters.numSlice = static_cast<size_t>(numSlice);
    req.parameters.numLineY = static_cast<size_t>(Ly);
    req.parameters.numPoint = static_cast<size_t>(numPoint);
    req.start.assign(start, start + 3);
    req.goal.assign(goal, goal + 3);
    HRM2D hrm(rb, arena, obs, req, per_orientation != 0);
    hrm.plan(static_cast<size_t>(g_max_refine));
    out_info[7] = static_cast<long>(hrm.numRefineDone);
    const auto& g = hrm.res_.graphStructure;
    out_info[0] = hrm.res_.solved ? 1 : 0;
    out_info[1] = static_cast<long>(g.vertex.size());
    out_info[2] = static_cast<long>(g.edge.size());
    out_info[3] = static_cast<long>(hrm.res_.PathId.size());
    out_info[4] = hrm.singleHitCount;
    out_info[5] = hrm.numEdgeTests;
    for (size_t i = 0; i < g.vertex.size() && static_cast<long>(i) < cap_v; ++i)
        for (int d = 0; d < 3; ++d) vertex[3 * i + d] = g.vertex[i][d];
    for (size_t i = 0; i < g.edge.size() && static_cast<long>(i) < cap_e; ++i) edge[2 * i] = static_cast<long>(g.edge[i].first), edge[2 * i + 1] = static_cast<long>(g.edge[i].second);
    for (size_t i = 0; i < hrm.res_.PathId.size() && static_cast<long>(i) < cap_p; ++i) path[i] = hrm.res_.PathId[i];
    return 0;
}


// ---- articulated path (hrmo_articulated.hpp) ---------------------------------------------------------------------------
// ParseURDF::getTransform(jointConfig, "body<k>") of the URDF's revolute tree: out16 = row-major 4x4.  Returns the number
// of joints, or -1 (file / link not found).
int hrmo_fk(const char* urdf, const double* joints, int nj, int body, double* out16) {
    try {
        KinematicTree kdl(urdf);
        const std::vector<double> q(joints, joints + nj);
        const Tf3 f = kdl.getTransform(q, body == 0 ? kdl.root : "body" + std::to_string(body));
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) out16[4 * i + j] = f.R.m[i][j];
            out16[4 * i + 3] = f.t[i];
        }
        out16[12] = out16[13] = out16[14] = 0.0, out16[15] = 1.0;
        return static_cast<int>(kdl.getNrOfJoints());
    } catch (const std::exception&) {
        return -1;
    }
}

// MultiBodyTree3D::robotTF(urdf, gBase, jointConfig): body poses [B][7] = (x,y,z,qw,qx,qy,qz), base first.
int hrmo_robot_tf_urdf(int n, int B, const double* robot, const char* urdf, const double* base7, const double* joints, int nj, double* out) {
    try {
        MultiBodyTree3D rb;
        rb.base = make_sq(robot, n);
        for (int b = 1; b < B; ++b) rb.addBody(make_sq(robot + 12 * b, n));
        KinematicTree kdl(urdf);
        Tf3 g;
        g.R = quat_to_rot(Quat{base7[3], base7[4], base7[5], base7[6]});
        for (int d = 0; d < 3; ++d) g.t[d] = base7[d];
        robotTF(rb, kdl, g, std::vector<double>(joints, joints + nj));
        for (int b = 0; b < B; ++b) {
            const SuperQuadrics& s = b == 0 ? rb.base : rb.link[static_cast<size_t>(b) - 1];
            for (int d = 0; d < 3; ++d) out[7 * b + d] = s.position[d];
            out[7 * b + 3] = s.quat.w, out[7 * b + 4] = s.quat.x, out[7 * b + 5] = s.quat.y, out[7 * b + 6] = s.quat.z;
        }
        return 0;
    } catch (const std::exception&) {
        return -1;
    }
}

// ProbHRM3D::plan with injected draws.  samples [n_samples][4 + nj] = (qw,qx,qy,qz, joints) for slices 2, 3, ...;
// start/goal: 7 + nj doubles.  vertex [cap_v][7 + nj].  out_info as hrmo_plan3d, [6] = number of C-slices built.
int hrmo_plan_prob3d(int n_arena, int n_obs, const double* sq, int n, int B, const double* robot, const char* urdf, int nj, int n_samples,
                     const double* samples, const double lim[6], int Lx, int Ly, int numPoint, const double* start, const double* goal,
                     int per_orientation, long* out_info, double* vertex, long cap_v, long* edge, long cap_e, long* path, long cap_p) {
    try {
        std::vector<SuperQuadrics> arena, obs;
        for (int i = 0; i < n_arena; ++i) arena.push_back(make_sq(sq + 12 * i, n));
        for (int i = 0; i < n_obs; ++i) obs.push_back(make_sq(sq + 12 * (n_arena + i), n));
        MultiBodyTree3D rb;
        rb.base = make_sq(robot, n);
        for (int b = 1; b < B; ++b) rb.addBody(make_sq(robot + 12 * b, n));
        PlanningRequest req;
        req.isRobotRigid = false;
        req.parameters.boundaryLimits.assign(lim, lim + 6);
        req.parameters.numLineX = static_cast<size_t>(Lx);
        req.parameters.numLineY = static_cast<size_t>(Ly);
        req.parameters.numPoint = static_cast<size_t>(numPoint);
        req.start.assign(start, start + 7 + nj);
        req.goal.assign(goal, goal + 7 + nj);
        std::vector<std::vector<double>> draws;
        for (int i = 0; i < n_samples; ++i) draws.emplace_back(samples + static_cast<size_t>(i) * (4 + nj), samples + static_cast<size_t>(i + 1) * (4 + nj));
        ProbHRM3D hrm(rb, urdf, arena, obs, req, draws, per_orientation != 0);
        if (static_cast<int>(hrm.kdl_.getNrOfJoints()) != nj) return -2;
        hrm.planProb(static_cast<size_t>(g_prob_refine_period));
        const auto& g = hrm.res_.graphStructure;
        const int W = 7 + nj;
        out_info[0] = hrm.res_.solved ? 1 : 0;
        out_info[1] = static_cast<long>(g.vertex.size());
        out_info[2] = static_cast<long>(g.edge.size());
        out_info[3] = static_cast<long>(hrm.res_.PathId.size());
        out_info[4] = hrm.singleHitCount;
        out_info[5] = hrm.numEdgeTests;
        out_info[6] = static_cast<long>(hrm.param_.numSlice);
        for (size_t i = 0; i < g.vertex.size() && static_cast<long>(i) < cap_v; ++i)
            for (int d = 0; d < W; ++d) vertex[static_cast<size_t>(W) * i + d] = g.vertex[i][d];
        for (size_t i = 0; i < g.edge.size() && static_cast<long>(i) < cap_e; ++i)
            edge[2 * i] = static_cast<long>(g.edge[i].first), edge[2 * i + 1] = static_cast<long>(g.edge[i].second);
        for (size_t i = 0; i < hrm.res_.PathId.size() && static_cast<long>(i) < cap_p; ++i) path[i] = hrm.res_.PathId[i];
        return 0;
    } catch (const std::exception&) {
        return -1;
    }
}
}  // extern "C"
