// Host-side C++ layer of the B200 HRM path: small fixed-size linear algebra with the Eigen conventions the
// reference (ChirikjianLab/hrm) relies on.  Product code (not the test oracle): it feeds the C ABI of
// include/hrmgpu.h and runs the parts of the planner that stay on the host per BASELINE.json's north_star.
// Build with -ffp-contract=off (the reference is a plain x86-64 build without FMA contraction).
#pragma once
#include <array>
#include <cmath>
#include <limits>
#include <vector>

namespace hrm {

using Coordinate = double;
using Index = size_t;
using Indicator = int;

// reference include/hrm/datastructure/DataType.h:30-32 — truncated on purpose
static const double PI = 3.1415926;
static const double HALF_PI = PI / 2.0;
static const double EPSILON = 1e-6;

using Vec3 = std::array<double, 3>;
using Mat3 = std::array<std::array<double, 3>, 3>;
using Quaternion = std::array<double, 4>;  // (w, x, y, z), never normalised implicitly

inline Mat3 matMul(const Mat3& a, const Mat3& b) {
    Mat3 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r[i][j] = a[i][0] * b[0][j] + a[i][1] * b[1][j] + a[i][2] * b[2][j];
    return r;
}
inline Vec3 matVec(const Mat3& a, const Vec3& v) {
    return {a[0][0] * v[0] + a[0][1] * v[1] + a[0][2] * v[2], a[1][0] * v[0] + a[1][1] * v[1] + a[1][2] * v[2],
            a[2][0] * v[0] + a[2][1] * v[1] + a[2][2] * v[2]};
}
inline Mat3 transpose(const Mat3& a) {
    Mat3 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r[i][j] = a[j][i];
    return r;
}
inline Mat3 mulDiag(const Mat3& a, const Vec3& d) {  // a * diag(d)
    Mat3 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r[i][j] = a[i][j] * d[j];
    return r;
}
inline double det(const Mat3& a) {
    return a[0][0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1]) - a[0][1] * (a[1][0] * a[2][2] - a[1][2] * a[2][0]) +
           a[0][2] * (a[1][0] * a[2][1] - a[1][1] * a[2][0]);
}
inline Mat3 inverse(const Mat3& a) {  // cofactors times 1/det, like Eigen's fixed-size inverse()
    Mat3 c;
    c[0][0] = a[1][1] * a[2][2] - a[1][2] * a[2][1];
    c[0][1] = a[0][2] * a[2][1] - a[0][1] * a[2][2];
    c[0][2] = a[0][1] * a[1][2] - a[0][2] * a[1][1];
    c[1][0] = a[1][2] * a[2][0] - a[1][0] * a[2][2];
    c[1][1] = a[0][0] * a[2][2] - a[0][2] * a[2][0];
    c[1][2] = a[0][2] * a[1][0] - a[0][0] * a[1][2];
    c[2][0] = a[1][0] * a[2][1] - a[1][1] * a[2][0];
    c[2][1] = a[0][1] * a[2][0] - a[0][0] * a[2][1];
    c[2][2] = a[0][0] * a[1][1] - a[0][1] * a[1][0];
    const double d = a[0][0] * c[0][0] + a[0][1] * c[1][0] + a[0][2] * c[2][0];
    const double id = 1.0 / d;
    for (auto& row : c)
        for (auto& v : row) v = id * v;
    return c;
}

// Eigen QuaternionBase::toRotationMatrix() on the raw coefficients (reference SuperQuadrics.cpp:78,109-110)
inline Mat3 toRotationMatrix(const Quaternion& q) {
    const double w = q[0], x = q[1], y = q[2], z = q[3];
    const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w;
    const double txx = tx * x, txy = ty * x, txz = tz * x;
    const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
    Mat3 r;
    r[0] = {1.0 - (tyy + tzz), txy - twz, txz + twy};
    r[1] = {txy + twz, 1.0 - (txx + tzz), tyz - twx};
    r[2] = {txz - twy, tyz + twx, 1.0 - (txx + tyy)};
    return r;
}

// Eigen Quaterniond(Matrix3d) (reference MultiBodyTree3D.cpp:38-50)
inline Quaternion toQuaternion(const Mat3& a) {
    double q[3], w;
    double t = a[0][0] + a[1][1] + a[2][2];
    if (t > 0.0) {
        t = std::sqrt(t + 1.0);
        w = 0.5 * t;
        t = 0.5 / t;
        q[0] = (a[2][1] - a[1][2]) * t;
        q[1] = (a[0][2] - a[2][0]) * t;
        q[2] = (a[1][0] - a[0][1]) * t;
    } else {
        int i = 0;
        if (a[1][1] > a[0][0]) i = 1;
        if (a[2][2] > a[i][i]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = std::sqrt(a[i][i] - a[j][j] - a[k][k] + 1.0);
        q[i] = 0.5 * t;
        t = 0.5 / t;
        w = (a[k][j] - a[j][k]) * t;
        q[j] = (a[j][i] + a[i][j]) * t;
        q[k] = (a[k][i] + a[i][k]) * t;
    }
    return {w, q[0], q[1], q[2]};
}

inline Quaternion quatMul(const Quaternion& a, const Quaternion& b) {
    return {a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
            a[0] * b[2] + a[2] * b[0] + a[3] * b[1] - a[1] * b[3], a[0] * b[3] + a[3] * b[0] + a[1] * b[2] - a[2] * b[1]};
}

// Eigen 3.3 angularDistance (reference HRM3D.cpp:168,456-468)
inline double angularDistance(const Quaternion& a, const Quaternion& b) {
    const Quaternion d = quatMul(a, {b[0], -b[1], -b[2], -b[3]});
    return 2.0 * std::atan2(std::sqrt(d[1] * d[1] + d[2] * d[2] + d[3] * d[3]), std::fabs(d[0]));
}

// Eigen 3.3 slerp, not re-normalised (reference InterpolateSE3.cpp:98)
inline Quaternion slerp(const Quaternion& a, double t, const Quaternion& b) {
    const double one = 1.0 - std::numeric_limits<double>::epsilon();
    const double d = a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3];
    const double ad = std::fabs(d);
    double s0, s1;
    if (ad >= one) {
        s0 = 1.0 - t;
        s1 = t;
    } else {
        const double th = std::acos(ad), st = std::sin(th);
        s0 = std::sin((1.0 - t) * th) / st;
        s1 = std::sin(t * th) / st;
    }
    if (d < 0.0) s1 = -s1;
    return {s0 * a[0] + s1 * b[0], s0 * a[1] + s1 * b[1], s0 * a[2] + s1 * b[2], s0 * a[3] + s1 * b[3]};
}

// reference src/util/DistanceMetric.cpp:5-13
inline double vectorEuclidean(const std::vector<Coordinate>& v1, const std::vector<Coordinate>& v2) {
    double acc = 0.0;
    for (size_t i = 0; i < v1.size(); ++i) {
        const double d = v1[i] - v2[i];
        acc = acc + d * d;
    }
    return std::sqrt(acc);
}

// Symmetric 3x3 eigen-decomposition (cyclic Jacobi), eigenvalues descending, U canonicalised to det +1: stands in for
// Eigen::JacobiSVD on the SPD matrices of TightFitEllipsoid.cpp.
inline void symEig3(const Mat3& A, Vec3& ev, Mat3& U) {
    double a[3][3], v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) a[i][j] = 0.5 * (A[i][j] + A[j][i]);
    for (int sweep = 0; sweep < 64; ++sweep) {
        const double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
        const double diag = a[0][0] * a[0][0] + a[1][1] * a[1][1] + a[2][2] * a[2][2];
        if (off <= 1e-32 * diag || off == 0.0) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (a[p][q] == 0.0) continue;
                const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; ++k) {
                    const double akp = a[k][p], akq = a[k][q];
                    a[k][p] = c * akp - s * akq;
                    a[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < 3; ++k) {
                    const double apk = a[p][k], aqk = a[q][k];
                    a[p][k] = c * apk - s * aqk;
                    a[q][k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 3; ++k) {
                    const double vkp = v[k][p], vkq = v[k][q];
                    v[k][p] = c * vkp - s * vkq;
                    v[k][q] = s * vkp + c * vkq;
                }
            }
    }
    int idx[3] = {0, 1, 2};
    for (int i = 0; i < 2; ++i)
        for (int j = i + 1; j < 3; ++j)
            if (a[idx[j]][idx[j]] > a[idx[i]][idx[i]]) std::swap(idx[i], idx[j]);
    for (int c = 0; c < 3; ++c) {
        ev[c] = a[idx[c]][idx[c]];
        for (int r = 0; r < 3; ++r) U[r][c] = v[r][idx[c]];
    }
    if (det(U) < 0.0)
        for (int r = 0; r < 3; ++r) U[r][2] = -U[r][2];
}

}  // namespace hrm
