// ORACLE — TEST INFRASTRUCTURE ONLY. Nothing under hrm_b200/ may include, link or call this.
//
// CPU restatement (FP64, reference loop order, reference comparisons, reference constants) of
// the geometry layer of ChirikjianLab/hrm.  Every function cites the reference file:line it
// follows (paths relative to /root/reference).  The reference itself cannot be compiled in
// this image (Eigen/Boost/... absent), so this restatement IS the parity contract; it is
// pinned against the reference's own known-answer tests (test/TestGeometry.cpp) in
// tests/test_oracle_kat.py.  Parity for functions the reference tests do not pin is
// "unpinned by the reference" (see DESIGN.md §Oracle).
#pragma once
#include "hrmo_math.hpp"

#include <algorithm>
#include <cstdint>
#include <utility>
#include <vector>

namespace hrmo {

// src/util/ExponentialFunction.h:11 — sgn(0) = 0
inline double sgn(double v) { return (v < 0) ? -1.0 : ((v > 0) ? 1.0 : 0.0); }

// src/util/ExponentialFunction.cpp:5-10
inline double exponentialFunction(double theta, double power, bool isSine) {
    return isSine ? sgn(std::sin(theta)) * std::pow(std::fabs(std::sin(theta)), power)
                  : sgn(std::cos(theta)) * std::pow(std::fabs(std::cos(theta)), power);
}

// src/util/ExponentialFunction.cpp:12-25 — evaluated on every element of the n x n grid,
// sin/cos computed twice per element (sign and abs), exactly like the Eigen expression.
inline std::vector<double> exponentialFunctionMatrixForm(const std::vector<double>& thetaList, double power,
                                                         bool isSine) {
    std::vector<double> out(thetaList.size());
    if (isSine) {
        for (size_t i = 0; i < thetaList.size(); ++i)
            out[i] = sgn(std::sin(thetaList[i])) * std::pow(std::fabs(std::sin(thetaList[i])), power);
    } else {
        for (size_t i = 0; i < thetaList.size(); ++i)
            out[i] = sgn(std::cos(thetaList[i])) * std::pow(std::fabs(std::cos(thetaList[i])), power);
    }
    return out;
}

// 3 x M column-major point cloud (Eigen::MatrixXd BoundaryPoints): p[3*k + d].
struct Points {
    int dim = 3;
    int num = 0;
    std::vector<double> p;
    double& at(int d, int k) { return p[static_cast<size_t>(k) * dim + d]; }
    double at(int d, int k) const { return p[static_cast<size_t>(k) * dim + d]; }
};

// ---------------------------------------------------------------------------------------------
// SuperQuadrics (src/geometry/SuperQuadrics.cpp)
// ---------------------------------------------------------------------------------------------
struct SuperQuadrics {
    double semiAxis[3];
    double epsilon[2];
    double position[3];
    Quat quat;
    int num;  // points per parameter (num_); Num_ = num*num

    // SuperQuadrics.cpp:8-26 — eta_(r,c) = eta_c, omega_(r,c) = omega_r, both n x n col-major.
    std::vector<double> eta, omega;
    std::vector<double> etaLin, omegaLin;

    SuperQuadrics() : semiAxis{1, 1, 1}, epsilon{1, 1}, position{0, 0, 0}, quat{1, 0, 0, 0}, num(0) {}
    SuperQuadrics(const double a[3], const double eps[2], const double pos[3], const Quat& q, int n) : quat(q), num(n) {
        for (int i = 0; i < 3; ++i) semiAxis[i] = a[i], position[i] = pos[i];
        epsilon[0] = eps[0];
        epsilon[1] = eps[1];
        etaLin = linspaced(n, -HALF_PI - EPSILON, HALF_PI + EPSILON);
        omegaLin = linspaced(n, -PI - EPSILON, PI + EPSILON);
        eta.resize(static_cast<size_t>(n) * n);
        omega.resize(static_cast<size_t>(n) * n);
        for (int c = 0; c < n; ++c)
            for (int r = 0; r < n; ++r) {
                eta[static_cast<size_t>(c) * n + r] = etaLin[c];      // RowVector replicated down rows
                omega[static_cast<size_t>(c) * n + r] = omegaLin[r];  // Vector replicated across cols
            }
    }
    int getNum() const { return num * num; }
    int getNumParam() const { return num; }

    // The power function on the n x n parameter grid.  Faithful build: evaluated on all n^2 entries per call, like the
    // reference's exponentialFunctionMatrixForm on eta_ / omega_ (ExponentialFunction.cpp:12-25).  HRMO_FAST build ("hoisted",
    // SURVEY.md §8d): the grids are replications of n values, so n evaluations are replicated -- bit-identical values.
    std::vector<double> powerGrid(bool alongEta, double power, bool isSine) const {
#ifdef HRMO_FAST
        const std::vector<double> v = exponentialFunctionMatrixForm(alongEta ? etaLin : omegaLin, power, isSine);
        std::vector<double> out(static_cast<size_t>(num) * num);
        for (int c = 0; c < num; ++c)
            for (int r = 0; r < num; ++r) out[static_cast<size_t>(c) * num + r] = alongEta ? v[c] : v[r];
        return out;
#else
        return exponentialFunctionMatrixForm(alongEta ? eta : omega, power, isSine);
#endif
    }

    // SuperQuadrics.cpp:46-81
    Points getOriginShape() const {
        const int M = num * num;
        const std::vector<double> ce1 = powerGrid(true, epsilon[0], false);
        const std::vector<double> cw2 = powerGrid(false, epsilon[1], false);
        const std::vector<double> ce1b = powerGrid(true, epsilon[0], false);
        const std::vector<double> sw2 = powerGrid(false, epsilon[1], true);
        const std::vector<double> se1 = powerGrid(true, epsilon[0], true);
        const Mat3 R = quat_to_rot(quat);
        Points out;
        out.dim = 3;
        out.num = M;
        out.p.resize(static_cast<size_t>(3) * M);
        for (int k = 0; k < M; ++k) {
            const double x = semiAxis[0] * (ce1[k] * cw2[k]);
            const double y = semiAxis[1] * (ce1b[k] * sw2[k]);
            const double z = semiAxis[2] * (se1[k] * 1.0);
            out.at(0, k) = (R.m[0][0] * x + R.m[0][1] * y + R.m[0][2] * z) + position[0];
            out.at(1, k) = (R.m[1][0] * x + R.m[1][1] * y + R.m[1][2] * z) + position[1];
            out.at(2, k) = (R.m[2][0] * x + R.m[2][1] * y + R.m[2][2] * z) + position[2];
        }
        return out;
    }

    // SuperQuadrics.cpp:84-143.  shapeB = ellipsoid body (semi-axes a2,b2,c2, orientation quatB);
    // its position does not enter.  K = -1 (arena, Minkowski difference) / +1 (obstacle, sum).
    Points getMinkSum3D(const double semiB[3], const Quat& quatB, int K) const {
        const int M = num * num;
        const double a1 = semiAxis[0], b1 = semiAxis[1], c1 = semiAxis[2];
        const double eps1 = epsilon[0], eps2 = epsilon[1];
        const Mat3 R1 = quat_to_rot(quat);
        const Mat3 R2 = quat_to_rot(quatB);
        const Mat3 Tinv = mul_diag(R2, semiB) * transpose(R2);  // :115

        // :118-135 gradient, matrix-form power function evaluated per call like the reference
        const std::vector<double> ce = powerGrid(true, 2 - eps1, false);
        const std::vector<double> cw = powerGrid(false, 2 - eps2, false);
        const std::vector<double> ceb = powerGrid(true, 2 - eps1, false);
        const std::vector<double> sw = powerGrid(false, 2 - eps2, true);
        const std::vector<double> se = powerGrid(true, 2 - eps1, true);

        // :137-140  (K*Tinv*Tinv*R1*gradPhi) ./ ||Tinv*R1*gradPhi||, products associated left to right
        const Mat3 M2 = (scale(static_cast<double>(K), Tinv) * Tinv) * R1;
        const Mat3 M1 = Tinv * R1;
        Points out = getOriginShape();
        for (int k = 0; k < M; ++k) {
            const Vec3 g = {(ce[k] * cw[k]) / a1, (ceb[k] * sw[k]) / b1, (se[k] * 1.0) / c1};
            const Vec3 u = M1 * g;
            const Vec3 w = M2 * g;
            const double nrm = std::sqrt(u.x * u.x + u.y * u.y + u.z * u.z);
            out.at(0, k) = out.at(0, k) + w.x / nrm;
            out.at(1, k) = out.at(1, k) + w.y / nrm;
            out.at(2, k) = out.at(2, k) + w.z / nrm;
        }
        return out;
    }
};

// ---------------------------------------------------------------------------------------------
// SuperEllipse (src/geometry/SuperEllipse.cpp)
// ---------------------------------------------------------------------------------------------
struct SuperEllipse {
    double semiAxis[2];
    double epsilon;
    double position[2];
    double angle;
    int num;

    SuperEllipse() : semiAxis{1, 1}, epsilon(1), position{0, 0}, angle(0), num(0) {}
    SuperEllipse(const double a[2], double eps, const double pos[2], double ang, int n) : epsilon(eps), angle(ang), num(n) {
        semiAxis[0] = a[0], semiAxis[1] = a[1];
        position[0] = pos[0], position[1] = pos[1];
    }

    // SuperEllipse.cpp:32-55
    Points getOriginShape() const {
        Points out;
        out.dim = 2;
        out.num = num;
        out.p.resize(static_cast<size_t>(2) * num);
        const double c = std::cos(angle), s = std::sin(angle);  // Rotation2Dd(angle).matrix()
        for (int i = 0; i < num; ++i) {
            const double th = 2 * i * PI / (static_cast<double>(num) - 1);
            const double x = semiAxis[0] * exponentialFunction(th, epsilon, false);
            const double y = semiAxis[1] * exponentialFunction(th, epsilon, true);
            out.at(0, i) = (c * x + (-s) * y) + position[0];
            out.at(1, i) = (s * x + c * y) + position[1];
        }
        return out;
    }

    // SuperEllipse.cpp:58-101.  Note the gradient omits 1/a, 1/b (SURVEY.md finding 5).
    Points getMinkSum2D(const double semiB[2], double angleB, int K) const {
        const double a1 = semiAxis[0], b1 = semiAxis[1], th1 = angle, eps1 = epsilon;
        const double a2 = semiB[0], b2 = semiB[1], th2 = angleB;
        const double r = std::fmin(a1, b1);
        const double d0 = a2 / r, d1 = b2 / r;
        const double c1 = std::cos(th1), s1 = std::sin(th1);
        const double c2 = std::cos(th2), s2 = std::sin(th2);
        // Tinv = R2 * diag * R2^T
        const double RD[2][2] = {{c2 * d0, -s2 * d1}, {s2 * d0, c2 * d1}};
        const double R2T[2][2] = {{c2, s2}, {-s2, c2}};
        double T[2][2];
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 2; ++j) T[i][j] = RD[i][0] * R2T[0][j] + RD[i][1] * R2T[1][j];
        const double R1[2][2] = {{c1, -s1}, {s1, c1}};
        // M2 = (((K*r)*Tinv)*Tinv)*R1, M1 = Tinv*R1
        const double kr = static_cast<double>(K) * r;
        double A[2][2], Bm[2][2], M2[2][2], M1[2][2];
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 2; ++j) A[i][j] = kr * T[i][j];
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 2; ++j) Bm[i][j] = A[i][0] * T[0][j] + A[i][1] * T[1][j];
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 2; ++j) {
                M2[i][j] = Bm[i][0] * R1[0][j] + Bm[i][1] * R1[1][j];
                M1[i][j] = T[i][0] * R1[0][j] + T[i][1] * R1[1][j];
            }
        Points out = getOriginShape();
        for (int i = 0; i < num; ++i) {
            const double th = 2 * i * PI / (static_cast<double>(num) - 1);
            const double g0 = 2 / eps1 * exponentialFunction(th, 2 - eps1, false);
            const double g1 = 2 / eps1 * exponentialFunction(th, 2 - eps1, true);
            const double u0 = M1[0][0] * g0 + M1[0][1] * g1;
            const double u1 = M1[1][0] * g0 + M1[1][1] * g1;
            const double w0 = M2[0][0] * g0 + M2[0][1] * g1;
            const double w1 = M2[1][0] * g0 + M2[1][1] * g1;
            const double nrm = std::sqrt(u0 * u0 + u1 * u1);
            out.at(0, i) = out.at(0, i) + w0 / nrm;
            out.at(1, i) = out.at(1, i) + w1 / nrm;
        }
        return out;
    }
};

// ---------------------------------------------------------------------------------------------
// Mesh (src/geometry/MeshGenerator.cpp:105-131, include/hrm/geometry/MeshGenerator.h:23-29)
// ---------------------------------------------------------------------------------------------
struct MeshMatrix {
    Points vertices;            // 3 x M
    std::vector<double> faces;  // F x 3, stored as doubles like the reference, row-major here
    int numFaces = 0;
    int f(int i, int c) const { return static_cast<int>(faces[static_cast<size_t>(i) * 3 + c]); }
#ifdef HRMO_FAST
    mutable bool bboxKnown = false;  // HRMO_FAST: the per-call maxCoeff / minCoeff scans of LineIntersection.cpp:12-30,58-65 once per mesh
    mutable double bbox[3][2];
#endif
};

inline MeshMatrix getMeshFromParamSurface(const Points& surfBound, int n) {
    const int numSurfVtx = (n - 1) * (n - 1);
    std::vector<double> q(static_cast<size_t>(numSurfVtx));
    for (int i = 0; i < n - 1; ++i) {
        const std::vector<double> seg =
            linspaced(n - 1, static_cast<double>(i * n), static_cast<double>((i + 1) * n - 2));
        for (int j = 0; j < n - 1; ++j) q[static_cast<size_t>(i) * (n - 1) + j] = seg[j];
    }
    MeshMatrix Mm;
    Mm.vertices = surfBound;
    Mm.numFaces = 2 * numSurfVtx;
    Mm.faces.assign(static_cast<size_t>(Mm.numFaces) * 3, 0.0);
    for (int i = 0; i < numSurfVtx; ++i) {
        Mm.faces[static_cast<size_t>(i) * 3 + 0] = q[i];
        Mm.faces[static_cast<size_t>(i) * 3 + 1] = q[i] + n;
        Mm.faces[static_cast<size_t>(i) * 3 + 2] = q[i] + n + 1;
        Mm.faces[static_cast<size_t>(numSurfVtx + i) * 3 + 0] = q[i];
        Mm.faces[static_cast<size_t>(numSurfVtx + i) * 3 + 1] = q[i] + 1;
        Mm.faces[static_cast<size_t>(numSurfVtx + i) * 3 + 2] = q[i] + n + 1;
    }
    return Mm;
}

// ---------------------------------------------------------------------------------------------
// Line intersection (src/geometry/LineIntersection.cpp)
// ---------------------------------------------------------------------------------------------
struct Line3D {
    double v[6];  // origin(3), direction(3)
};

// LineIntersection.cpp:126-171
inline bool intersectLineTriangle3D(const Line3D& line, const Vec3& t0, const Vec3& u, const Vec3& v, Vec3& pt) {
    const double tol = 1e-12;
    Vec3 n = cross(u, v);
    normalize(n);
    const Vec3 o = {line.v[0], line.v[1], line.v[2]};
    const Vec3 d = {line.v[3], line.v[4], line.v[5]};
    const double a = -dot(n, o - t0);
    const double b = dot(n, d);
    if (!((std::fabs(b) > tol) && (norm(n) > tol))) return false;
    const double pos = a / b;
    pt = o + pos * d;
    const double uu = dot(u, u);
    const double uv = dot(u, v);
    const double vv = dot(v, v);
    const double wu = dot(u, pt - t0);
    const double wv = dot(v, pt - t0);
    const double D = std::pow(uv, 2) - uu * vv;
    const double s = (uv * wv - vv * wu) / D;
    if ((s < -tol) || (s > 1.0 + tol)) return false;
    const double t = (uv * wu - uu * wv) / D;
    return !((t < -tol) || (s + t > 1.0 + tol));
}

inline void rowMinMax(const Points& p, int row, double& mn, double& mx) {
    mn = p.at(row, 0);
    mx = p.at(row, 0);
    for (int k = 1; k < p.num; ++k) {
        const double v = p.at(row, k);
        if (v < mn) mn = v;
        if (v > mx) mx = v;
    }
}

// Bounding box of a mesh along one axis: scanned per call like the reference, or (HRMO_FAST) once per mesh -- same values.
inline void meshMinMax(const MeshMatrix& shape, int row, double& mn, double& mx) {
#ifdef HRMO_FAST
    if (!shape.bboxKnown) {
        for (int r = 0; r < 3; ++r) rowMinMax(shape.vertices, r, shape.bbox[r][0], shape.bbox[r][1]);
        shape.bboxKnown = true;
    }
    mn = shape.bbox[row][0], mx = shape.bbox[row][1];
#else
    rowMinMax(shape.vertices, row, mn, mx);
#endif
}

// LineIntersection.cpp:56-124 — first two hits in FACE-INDEX order (SURVEY.md finding 2).
inline std::vector<Vec3> intersectVerticalLineMesh3D(const Line3D& line, const MeshMatrix& shape) {
    std::vector<Vec3> points;
    double mn, mx;
    meshMinMax(shape, 0, mn, mx);
    if (line.v[0] > mx || line.v[0] < mn) return points;
    meshMinMax(shape, 1, mn, mx);
    if (line.v[1] > mx || line.v[1] < mn) return points;
    Vec3 pt;
    const Points& V = shape.vertices;
    for (int i = 0; i < shape.numFaces; ++i) {
        const int f0 = shape.f(i, 0), f1 = shape.f(i, 1), f2 = shape.f(i, 2);
        if (line.v[0] < std::fmin(V.at(0, f0), std::fmin(V.at(0, f1), V.at(0, f2))) ||
            line.v[0] > std::fmax(V.at(0, f0), std::fmax(V.at(0, f1), V.at(0, f2))))
            continue;
        if (line.v[1] < std::fmin(V.at(1, f0), std::fmin(V.at(1, f1), V.at(1, f2))) ||
            line.v[1] > std::fmax(V.at(1, f0), std::fmax(V.at(1, f1), V.at(1, f2))))
            continue;
        const Vec3 t0 = {V.at(0, f0), V.at(1, f0), V.at(2, f0)};
        const Vec3 u = Vec3{V.at(0, f1), V.at(1, f1), V.at(2, f1)} - t0;
        const Vec3 v = Vec3{V.at(0, f2), V.at(1, f2), V.at(2, f2)} - t0;
        if (intersectLineTriangle3D(line, t0, u, v, pt)) points.push_back(pt);
        if (points.size() == 2) break;
    }
    return points;
}

// LineIntersection.cpp:7-54 — general line; bbox reject only for near-axis-parallel components.
inline std::vector<Vec3> intersectLineMesh3D(const Line3D& line, const MeshMatrix& shape) {
    std::vector<Vec3> points;
    for (int i = 0; i < 3; ++i) {
        // direction.dot(basis_i) for unit basis vectors = d_i (+ exact zeros)
        const double di = line.v[3 + i];
        if (std::fabs(di) < 1e-6) {
            double mn, mx;
            meshMinMax(shape, i, mn, mx);
            if (line.v[i] > mx || line.v[i] < mn) return points;
        }
    }
    Vec3 pt;
    const Points& V = shape.vertices;
    for (int i = 0; i < shape.numFaces; ++i) {
        const int f0 = shape.f(i, 0), f1 = shape.f(i, 1), f2 = shape.f(i, 2);
        const Vec3 t0 = {V.at(0, f0), V.at(1, f0), V.at(2, f0)};
        const Vec3 u = Vec3{V.at(0, f1), V.at(1, f1), V.at(2, f1)} - t0;
        const Vec3 v = Vec3{V.at(0, f2), V.at(1, f2), V.at(2, f2)} - t0;
        if (intersectLineTriangle3D(line, t0, u, v, pt)) points.push_back(pt);
        if (points.size() == 2) break;
    }
    return points;
}

// Explicit 2x2 column-pivoted solve standing in for Eigen colPivHouseholderQr().solve()
// (LineIntersection.cpp:203).  For a regular system the solution is the exact one up to
// rounding; for a rank-deficient system Eigen returns the basic solution with the free
// component zero.  Only the open-interval tests (0,1)x(0,1) consume it.
inline void solve2x2ColPiv(const double A[2][2], const double b[2], double sol[2]) {
    const double n0 = A[0][0] * A[0][0] + A[1][0] * A[1][0];
    const double n1 = A[0][1] * A[0][1] + A[1][1] * A[1][1];
    const int p0 = (n1 > n0) ? 1 : 0, p1 = 1 - p0;
    const double a0 = A[0][p0], a1 = A[1][p0];
    const double c0 = A[0][p1], c1 = A[1][p1];
    const double nrm = std::sqrt(a0 * a0 + a1 * a1);
    sol[0] = sol[1] = 0.0;
    if (nrm == 0.0) return;
    // Householder-free equivalent: Gram-Schmidt QR of the pivoted matrix.
    const double q0x = a0 / nrm, q0y = a1 / nrm;
    const double r01 = q0x * c0 + q0y * c1;
    const double ex = c0 - r01 * q0x, ey = c1 - r01 * q0y;
    const double r11 = std::sqrt(ex * ex + ey * ey);
    const double y0 = q0x * b[0] + q0y * b[1];
    const double thresh = std::numeric_limits<double>::epsilon() * 2.0 * nrm;
    if (r11 <= thresh) {  // rank 1: free component zero
        sol[p0] = y0 / nrm;
        sol[p1] = 0.0;
        return;
    }
    const double q1x = ex / r11, q1y = ey / r11;
    const double y1 = q1x * b[0] + q1y * b[1];
    const double x1 = y1 / r11;
    const double x0 = (y0 - r01 * x1) / nrm;
    sol[p0] = x0;
    sol[p1] = x1;
}

// LineIntersection.cpp:173-213
inline bool isIntersectSegPolygon2D(const double s1[2], const double s2[2], const Points& shape) {
    const double px1 = s1[0], py1 = s1[1], px2 = s2[0], py2 = s2[1];
    double A[2][2];
    A[0][0] = px2 - px1;
    A[1][0] = py2 - py1;
    double b[2];
    const int N = shape.num;
    for (int i = 0; i < N; ++i) {
        const double x1 = shape.at(0, i), y1 = shape.at(1, i);
        double x2 = shape.at(0, 0), y2 = shape.at(1, 0);
        if (i != N - 1) {
            x2 = shape.at(0, i + 1);
            y2 = shape.at(1, i + 1);
        }
        A[0][1] = x1 - x2;
        A[1][1] = y1 - y2;
        b[0] = x1 - px1;
        b[1] = y1 - py1;
        double sol[2];
        solve2x2ColPiv(A, b, sol);
        if ((sol[0] > 0 && sol[0] < 1) && (sol[1] > 0 && sol[1] < 1)) return true;
    }
    return false;
}

// LineIntersection.cpp:215-248 — ALL crossings in edge order (closing edge last).
inline std::vector<double> intersectHorizontalLinePolygon2D(double ty, const Points& shape) {
    std::vector<double> points;
    double mn, mx;
    rowMinMax(shape, 1, mn, mx);
    if (ty > mx || ty < mn) return points;
    const int N = shape.num;
    for (int i = 0; i < N; ++i) {
        const double x1 = shape.at(0, i), y1 = shape.at(1, i);
        double x2, y2;
        if (i == N - 1) {
            x2 = shape.at(0, 0);
            y2 = shape.at(1, 0);
        } else {
            x2 = shape.at(0, i + 1);
            y2 = shape.at(1, i + 1);
        }
        if (ty >= std::fmin(y1, y2) && ty <= std::fmax(y1, y2)) {
            const double t = (ty - y2) / (y1 - y2);
            if (t >= 0.0 && t <= 1.0) points.push_back(t * x1 + (1.0 - t) * x2);
        }
    }
    return points;
}

// ---------------------------------------------------------------------------------------------
// Interpolation (src/util/InterpolateSE3.cpp)
// ---------------------------------------------------------------------------------------------
// :89-103
inline std::vector<Quat> interpolateSlerp(const Quat& qa, const Quat& qb, size_t numStep) {
    std::vector<Quat> out;
    const double dt = 1.0 / static_cast<double>(numStep);
    for (size_t i = 0; i <= numStep; ++i) out.push_back(slerp(qa, static_cast<double>(i) * dt, qb));
    return out;
}
// :104-122
inline std::vector<std::vector<double>> interpolateRn(const std::vector<double>& vStart, const std::vector<double>& vEnd,
                                                      size_t numStep) {
    std::vector<std::vector<double>> out;
    const double dt = 1.0 / static_cast<double>(numStep);
    for (size_t i = 0; i <= numStep; ++i) {
        std::vector<double> vStep;
        for (size_t j = 0; j < vStart.size(); ++j)
            vStep.push_back((1.0 - static_cast<double>(i) * dt) * vStart[j] + static_cast<double>(i) * dt * vEnd[j]);
        out.push_back(vStep);
    }
    return out;
}
// :5-29
inline std::vector<std::vector<double>> interpolateSE3(const std::vector<double>& vStart, const std::vector<double>& vEnd,
                                                       size_t numStep) {
    std::vector<std::vector<double>> out;
    const std::vector<Quat> qi =
        interpolateSlerp({vStart[3], vStart[4], vStart[5], vStart[6]}, {vEnd[3], vEnd[4], vEnd[5], vEnd[6]}, numStep);
    const std::vector<std::vector<double>> ti =
        interpolateRn({vStart[0], vStart[1], vStart[2]}, {vEnd[0], vEnd[1], vEnd[2]}, numStep);
    for (size_t i = 0; i <= numStep; ++i)
        out.push_back({ti[i][0], ti[i][1], ti[i][2], qi[i].w, qi[i].x, qi[i].y, qi[i].z});
    return out;
}
// :31-65
inline std::vector<std::vector<double>> interpolateCompoundSE3Rn(const std::vector<double>& vStart,
                                                                 const std::vector<double>& vEnd, size_t numStep) {
    std::vector<std::vector<double>> out;
    if (numStep == 0) out = {vStart, vEnd};
    if (vStart.size() == 7) {
        out = interpolateSE3(vStart, vEnd, numStep);
    } else {
        const std::vector<double> s3(vStart.begin(), vStart.begin() + 7), sn(vStart.begin() + 7, vStart.end());
        const std::vector<double> e3(vEnd.begin(), vEnd.begin() + 7), en(vEnd.begin() + 7, vEnd.end());
        const auto a = interpolateSE3(s3, e3, numStep);
        const auto b = interpolateRn(sn, en, numStep);
        for (size_t i = 0; i <= numStep; ++i) {
            std::vector<double> v = a.at(i);
            v.insert(v.end(), b.at(i).begin(), b.at(i).end());
            out.push_back(v);
        }
    }
    return out;
}

// ---------------------------------------------------------------------------------------------
// Tightly-fitted ellipsoids (src/geometry/TightFitEllipsoid.cpp)
// ---------------------------------------------------------------------------------------------
struct Ellipse2 {
    double semi[2];
    double angle;
};
struct Ellipsoid3 {
    double semi[3];
    Quat quat;
};

// Symmetric 2x2 "SVD": eigenvalues descending, U columns eigenvectors.
inline void sym_eig2(const double A[2][2], double ev[2], double U[2][2]) {
    const double a = A[0][0], b = 0.5 * (A[0][1] + A[1][0]), d = A[1][1];
    const double tr = a + d, df = a - d;
    const double rt = std::sqrt(df * df + 4.0 * b * b);
    ev[0] = 0.5 * (tr + rt);
    ev[1] = 0.5 * (tr - rt);
    double vx, vy;
    if (std::fabs(b) > 0.0) {
        vx = ev[0] - d;
        vy = b;
        if (std::fabs(vx) + std::fabs(vy) == 0.0) {
            vx = b;
            vy = ev[0] - a;
        }
    } else if (a >= d) {
        vx = 1.0;
        vy = 0.0;
    } else {
        vx = 0.0;
        vy = 1.0;
    }
    const double n = std::sqrt(vx * vx + vy * vy);
    vx /= n;
    vy /= n;
    U[0][0] = vx;
    U[1][0] = vy;
    U[0][1] = -vy;
    U[1][1] = vx;
}

// TightFitEllipsoid.cpp:6-41
inline Ellipse2 getMVCE2D(const double a[2], const double b[2], double thetaA, double thetaB) {
    const double Ra[2][2] = {{std::cos(thetaA), -std::sin(thetaA)}, {std::sin(thetaA), std::cos(thetaA)}};
    const double Rb[2][2] = {{std::cos(thetaB), -std::sin(thetaB)}, {std::sin(thetaB), std::cos(thetaB)}};
    const double r = std::fmin(b[0], b[1]);
    const double dg[2] = {r / b[0], r / b[1]};
    const double dA[2] = {std::pow(a[0], -2), std::pow(a[1], -2)};
    auto rdrt = [](const double R[2][2], const double d[2], double out[2][2]) {
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 2; ++j) out[i][j] = (R[i][0] * d[0]) * R[j][0] + (R[i][1] * d[1]) * R[j][1];
    };
    auto mul = [](const double X[2][2], const double Y[2][2], double out[2][2]) {
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 2; ++j) out[i][j] = X[i][0] * Y[0][j] + X[i][1] * Y[1][j];
    };
    double T[2][2], Am[2][2];
    rdrt(Rb, dg, T);
    rdrt(Ra, dA, Am);
    const double dt = T[0][0] * T[1][1] - T[0][1] * T[1][0];
    const double idt = 1.0 / dt;  // Eigen 2x2 inverse(): cofactors times 1/det
    const double Ti[2][2] = {{T[1][1] * idt, -T[0][1] * idt}, {-T[1][0] * idt, T[0][0] * idt}};
    double tmp[2][2], aP[2][2];
    mul(Ti, Am, tmp);
    mul(tmp, Ti, aP);
    double ev[2], U[2][2];
    sym_eig2(aP, ev, U);
    const double aPrime[2] = {std::pow(ev[0], -0.5), std::pow(ev[1], -0.5)};
    const double cPrime[2] = {std::max(aPrime[0], r), std::max(aPrime[1], r)};
    const double dC[2] = {std::pow(cPrime[0], -2), std::pow(cPrime[1], -2)};
    double C[2][2], TU[2][2], TUD[2][2];
    const double Ut[2][2] = {{U[0][0], U[1][0]}, {U[0][1], U[1][1]}};
    mul(T, U, TU);  // T*U*diagC*U^T*T, left to right
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j) TUD[i][j] = TU[i][j] * dC[j];
    mul(TUD, Ut, tmp);
    mul(tmp, T, C);
    sym_eig2(C, ev, U);
    Ellipse2 out;
    out.angle = std::acos(std::fmax(-1.0, std::fmin(1.0, U[0][0])));
    out.semi[0] = std::pow(ev[0], -0.5);
    out.semi[1] = std::pow(ev[1], -0.5);
    return out;
}

// TightFitEllipsoid.cpp:43-81
inline Ellipsoid3 getMVCE3D(const double a[3], const double b[3], const Quat& quatA, const Quat& quatB) {
    const Mat3 Ra = quat_to_rot(quatA);
    const Mat3 Rb = quat_to_rot(quatB);
    const double r = std::fmin(b[0], std::fmin(b[1], b[2]));
    const double dg[3] = {r / b[0], r / b[1], r / b[2]};
    const double dA[3] = {std::pow(a[0], -2.0), std::pow(a[1], -2.0), std::pow(a[2], -2.0)};
    const Mat3 T = mul_diag(Rb, dg) * transpose(Rb);
    const Mat3 Ti = inverse(T);
    const Mat3 aP = (Ti * (mul_diag(Ra, dA) * transpose(Ra))) * Ti;
    double ev[3];
    Mat3 U;
    sym_eig3(aP, ev, U);
    const double aPrime[3] = {std::pow(ev[0], -0.5), std::pow(ev[1], -0.5), std::pow(ev[2], -0.5)};
    const double cPrime[3] = {std::fmax(aPrime[0], r), std::fmax(aPrime[1], r), std::fmax(aPrime[2], r)};
    const double dC[3] = {std::pow(cPrime[0], -2.0), std::pow(cPrime[1], -2.0), std::pow(cPrime[2], -2.0)};
    const Mat3 C = (mul_diag(T * U, dC) * transpose(U)) * T;  // T*U*diagC*U^T*T, left to right
    sym_eig3(C, ev, U);
    Ellipsoid3 out;
    out.quat = rot_to_quat(U);
    out.semi[0] = std::pow(ev[0], -0.5);
    out.semi[1] = std::pow(ev[1], -0.5);
    out.semi[2] = std::pow(ev[2], -0.5);
    return out;
}

// TightFitEllipsoid.cpp:83-96
inline Ellipse2 getTFE2D(const double a[2], double thetaA, double thetaB, size_t numStep) {
    Ellipse2 e = getMVCE2D(a, a, thetaA, thetaB);
    const double dt = 1.0 / (static_cast<double>(numStep) - 1);
    for (size_t i = 0; i < numStep; ++i) {
        const double ci = static_cast<double>(i);
        const double thetaStep = (1 - ci * dt) * thetaA + ci * dt * thetaB;
        e = getMVCE2D(a, e.semi, thetaStep, e.angle);
    }
    return e;
}

// TightFitEllipsoid.cpp:98-114
inline Ellipsoid3 getTFE3D(const double a[3], const Quat& quatA, const Quat& quatB, size_t numStep) {
    const std::vector<Quat> qi = interpolateSlerp(quatA, quatB, numStep);
    Ellipsoid3 e = getMVCE3D(a, a, quatA, quatB);
    for (size_t i = 1; i < numStep; ++i) e = getMVCE3D(a, e.semi, qi.at(i), e.quat);
    return e;
}

}  // namespace hrmo
