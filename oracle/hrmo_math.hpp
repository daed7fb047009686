// ORACLE — TEST INFRASTRUCTURE ONLY. Nothing under hrm_b200/ may include, link or call this.
//
// Dependency-free FP64 restatement of the Eigen 3.3.x primitives the reference's hot path
// uses (Eigen is absent from /root/reference and from this image).
// Formulas are stated from the published Eigen 3.3.7 algorithm (Ubuntu 20.04 libeigen3-dev,
// the version the reference's Docker image pins: /root/reference/Docker/Dockerfile:1,14).
// Build with -O3 -ffp-contract=off and no -march (the reference is plain x86-64: no FMA).
#pragma once
#include <cmath>
#include <cstddef>
#include <limits>
#include <vector>

namespace hrmo {

// include/hrm/datastructure/DataType.h:30-32 — truncated constants, used verbatim.
static const double PI = 3.1415926;
static const double HALF_PI = PI / 2.0;
static const double EPSILON = 1e-6;

struct Vec3 {
    double x, y, z;
};
inline Vec3 operator+(const Vec3& a, const Vec3& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec3 operator-(const Vec3& a, const Vec3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec3 operator*(double s, const Vec3& a) { return {s * a.x, s * a.y, s * a.z}; }
inline double dot(const Vec3& a, const Vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Vec3 cross(const Vec3& a, const Vec3& b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
// Eigen MatrixBase::normalize(): z = squaredNorm(); if (z > 0) v /= sqrt(z)
inline void normalize(Vec3& v) {
    const double z = dot(v, v);
    if (z > 0.0) {
        const double s = std::sqrt(z);
        v.x /= s;
        v.y /= s;
        v.z /= s;
    }
}
inline double norm(const Vec3& v) { return std::sqrt(dot(v, v)); }

struct Mat3 {
    double m[3][3];
    double& operator()(int i, int j) { return m[i][j]; }
    double operator()(int i, int j) const { return m[i][j]; }
};
inline Mat3 mat3_identity() { return {{{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}}; }
inline Mat3 operator*(const Mat3& a, const Mat3& b) {
    Mat3 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j];
    return r;
}
inline Vec3 operator*(const Mat3& a, const Vec3& v) {
    return {a.m[0][0] * v.x + a.m[0][1] * v.y + a.m[0][2] * v.z, a.m[1][0] * v.x + a.m[1][1] * v.y + a.m[1][2] * v.z,
            a.m[2][0] * v.x + a.m[2][1] * v.y + a.m[2][2] * v.z};
}
inline Mat3 transpose(const Mat3& a) {
    Mat3 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[j][i];
    return r;
}
inline Mat3 scale(double s, const Mat3& a) {
    Mat3 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r.m[i][j] = s * a.m[i][j];
    return r;
}
// A * diag(d): scales columns.
inline Mat3 mul_diag(const Mat3& a, const double d[3]) {
    Mat3 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][j] * d[j];
    return r;
}
inline double det(const Mat3& a) {
    return a.m[0][0] * (a.m[1][1] * a.m[2][2] - a.m[1][2] * a.m[2][1]) -
           a.m[0][1] * (a.m[1][0] * a.m[2][2] - a.m[1][2] * a.m[2][0]) +
           a.m[0][2] * (a.m[1][0] * a.m[2][1] - a.m[1][1] * a.m[2][0]);
}
// Cofactor inverse (Eigen uses the cofactor formula for fixed 3x3 inverse()).
inline Mat3 inverse(const Mat3& a) {
    Mat3 c;
    c.m[0][0] = a.m[1][1] * a.m[2][2] - a.m[1][2] * a.m[2][1];
    c.m[0][1] = a.m[0][2] * a.m[2][1] - a.m[0][1] * a.m[2][2];
    c.m[0][2] = a.m[0][1] * a.m[1][2] - a.m[0][2] * a.m[1][1];
    c.m[1][0] = a.m[1][2] * a.m[2][0] - a.m[1][0] * a.m[2][2];
    c.m[1][1] = a.m[0][0] * a.m[2][2] - a.m[0][2] * a.m[2][0];
    c.m[1][2] = a.m[0][2] * a.m[1][0] - a.m[0][0] * a.m[1][2];
    c.m[2][0] = a.m[1][0] * a.m[2][1] - a.m[1][1] * a.m[2][0];
    c.m[2][1] = a.m[0][1] * a.m[2][0] - a.m[0][0] * a.m[2][1];
    c.m[2][2] = a.m[0][0] * a.m[1][1] - a.m[0][1] * a.m[1][0];
    const double d = a.m[0][0] * c.m[0][0] + a.m[0][1] * c.m[1][0] + a.m[0][2] * c.m[2][0];
    const double id = 1.0 / d;
    return scale(id, c);
}

struct Quat {
    double w, x, y, z;
};

// Eigen QuaternionBase::toRotationMatrix() — no normalisation (SURVEY.md finding 4).
inline Mat3 quat_to_rot(const Quat& q) {
    const double tx = 2.0 * q.x, ty = 2.0 * q.y, tz = 2.0 * q.z;
    const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
    const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
    const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    Mat3 r;
    r.m[0][0] = 1.0 - (tyy + tzz);
    r.m[0][1] = txy - twz;
    r.m[0][2] = txz + twy;
    r.m[1][0] = txy + twz;
    r.m[1][1] = 1.0 - (txx + tzz);
    r.m[1][2] = tyz - twx;
    r.m[2][0] = txz - twy;
    r.m[2][1] = tyz + twx;
    r.m[2][2] = 1.0 - (txx + tyy);
    return r;
}

// Eigen quaternionbase_assign_impl<Matrix3,3,3>::run — Quaterniond(Matrix3d).
inline Quat rot_to_quat(const Mat3& a) {
    double q[3];
    double w;
    double t = a.m[0][0] + a.m[1][1] + a.m[2][2];
    if (t > 0.0) {
        t = std::sqrt(t + 1.0);
        w = 0.5 * t;
        t = 0.5 / t;
        q[0] = (a.m[2][1] - a.m[1][2]) * t;
        q[1] = (a.m[0][2] - a.m[2][0]) * t;
        q[2] = (a.m[1][0] - a.m[0][1]) * t;
    } else {
        int i = 0;
        if (a.m[1][1] > a.m[0][0]) i = 1;
        if (a.m[2][2] > a.m[i][i]) i = 2;
        const int j = (i + 1) % 3;
        const int k = (j + 1) % 3;
        t = std::sqrt(a.m[i][i] - a.m[j][j] - a.m[k][k] + 1.0);
        q[i] = 0.5 * t;
        t = 0.5 / t;
        w = (a.m[k][j] - a.m[j][k]) * t;
        q[j] = (a.m[j][i] + a.m[i][j]) * t;
        q[k] = (a.m[k][i] + a.m[i][k]) * t;
    }
    return {w, q[0], q[1], q[2]};
}

// Quaterniond(AngleAxisd(angle, axis)): w = cos(angle/2), vec = sin(angle/2) * axis.
inline Quat angle_axis_to_quat(double angle, const Vec3& axis) {
    const double ha = 0.5 * angle;
    const double s = std::sin(ha);
    return {std::cos(ha), s * axis.x, s * axis.y, s * axis.z};
}

inline Quat quat_mul(const Quat& a, const Quat& b) {
    return {a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
            a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z, a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x};
}
inline Quat quat_conj(const Quat& q) { return {q.w, -q.x, -q.y, -q.z}; }
inline double quat_dot(const Quat& a, const Quat& b) { return a.w * b.w + a.x * b.x + a.y * b.y + a.z * b.z; }

// Eigen 3.3 QuaternionBase::angularDistance: d = a * conj(b); 2 * atan2(|d.vec|, |d.w|).
inline double angular_distance(const Quat& a, const Quat& b) {
    const Quat d = quat_mul(a, quat_conj(b));
    return 2.0 * std::atan2(std::sqrt(d.x * d.x + d.y * d.y + d.z * d.z), std::fabs(d.w));
}

// Eigen 3.3 QuaternionBase::slerp (result not re-normalised).
inline Quat slerp(const Quat& a, double t, const Quat& b) {
    const double one = 1.0 - std::numeric_limits<double>::epsilon();
    const double d = quat_dot(a, b);
    const double absD = std::fabs(d);
    double scale0, scale1;
    if (absD >= one) {
        scale0 = 1.0 - t;
        scale1 = t;
    } else {
        const double theta = std::acos(absD);
        const double sinTheta = std::sin(theta);
        scale0 = std::sin((1.0 - t) * theta) / sinTheta;
        scale1 = std::sin(t * theta) / sinTheta;
    }
    if (d < 0.0) scale1 = -scale1;
    return {scale0 * a.w + scale1 * b.w, scale0 * a.x + scale1 * b.x, scale0 * a.y + scale1 * b.y,
            scale0 * a.z + scale1 * b.z};
}

// Eigen 3.3.7 DenseBase::LinSpaced(n, low, high) for floating scalars (linspaced_op_impl).
inline std::vector<double> linspaced(int n, double low, double high) {
    std::vector<double> v(static_cast<size_t>(n));
    if (n == 1) {
        v[0] = high;
        return v;
    }
    const int size1 = n - 1;
    const double step = (high - low) / static_cast<double>(n - 1);
    const bool flip = std::fabs(high) < std::fabs(low);
    for (int i = 0; i < n; ++i) {
        if (flip)
            v[i] = (i == 0) ? low : (high - static_cast<double>(size1 - i) * step);
        else
            v[i] = (i == size1) ? high : (low + static_cast<double>(i) * step);
    }
    return v;
}

// Symmetric 3x3 eigen-decomposition by cyclic Jacobi rotations, eigenvalues sorted descending,
// eigenvectors in the columns of U, det(U) canonicalised to +1.  Stands in for
// Eigen::JacobiSVD on the symmetric positive-definite matrices of TightFitEllipsoid.cpp:65-78
// (singular values == eigenvalues, U == V).  Column signs are unspecified in Eigen; see
// SURVEY.md App. C13 for why only U*diag*U^T matters downstream.
inline void sym_eig3(const Mat3& A, double eval[3], Mat3& U) {
    double a[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) a[i][j] = 0.5 * (A.m[i][j] + A.m[j][i]);
    double v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int sweep = 0; sweep < 64; ++sweep) {
        const double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
        const double diag = a[0][0] * a[0][0] + a[1][1] * a[1][1] + a[2][2] * a[2][2];
        if (off <= 1e-32 * diag || off == 0.0) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (a[p][q] == 0.0) continue;
                const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double c = 1.0 / std::sqrt(t * t + 1.0);
                const double s = t * c;
                for (int k = 0; k < 3; ++k) {  // A <- A * J
                    const double akp = a[k][p], akq = a[k][q];
                    a[k][p] = c * akp - s * akq;
                    a[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < 3; ++k) {  // A <- J^T * A
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
            if (a[idx[j]][idx[j]] > a[idx[i]][idx[i]]) {
                const int t = idx[i];
                idx[i] = idx[j];
                idx[j] = t;
            }
    for (int c = 0; c < 3; ++c) {
        eval[c] = a[idx[c]][idx[c]];
        for (int r = 0; r < 3; ++r) U.m[r][c] = v[r][idx[c]];
    }
    if (det(U) < 0.0)
        for (int r = 0; r < 3; ++r) U.m[r][2] = -U.m[r][2];
}

}  // namespace hrmo
