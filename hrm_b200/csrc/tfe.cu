// Tightly-fitted ellipsoids of all (layer pair, body) combinations in one launch (reference
// src/geometry/TightFitEllipsoid.cpp:43-114, src/util/InterpolateSE3.cpp:89-103): one thread per (pair, body) runs the
// reference's chain -- slerp between the two orientations, then numStep successive minimum-volume enclosing ellipsoids, each
// two 3x3 symmetric eigen-decompositions -- and leaves semi-axes and rotation where hrmgpu_build_bridge expects them, so the
// bridge layers are built without a host round trip.  The arithmetic is the host layer's (hrm_host.hpp getMVCE3D / getTFE3D,
// hrm_math.hpp symEig3: cyclic Jacobi standing in for Eigen::JacobiSVD on these SPD matrices), operation for operation; this
// file is compiled with -fmad=false, so +, -, *, /, sqrt round like the host's.  pow / acos / sin are CUDA's (<= 2 ulp), not
// glibc's: the ellipsoids can differ from the host's in the last bits, the bridge meshes then by ~1e-16 relative.
#include "internal.cuh"

namespace hrmgpu {

struct M3 {
    double a[3][3];
};

__device__ static M3 mat_mul(const M3& x, const M3& y) {
    M3 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r.a[i][j] = x.a[i][0] * y.a[0][j] + x.a[i][1] * y.a[1][j] + x.a[i][2] * y.a[2][j];
    return r;
}
__device__ static M3 transpose(const M3& x) {
    M3 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r.a[i][j] = x.a[j][i];
    return r;
}
__device__ static M3 mul_diag(const M3& x, const double d[3]) {
    M3 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r.a[i][j] = x.a[i][j] * d[j];
    return r;
}
__device__ static double det3(const M3& m) {
    const double(*a)[3] = m.a;
    return a[0][0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1]) - a[0][1] * (a[1][0] * a[2][2] - a[1][2] * a[2][0]) +
           a[0][2] * (a[1][0] * a[2][1] - a[1][1] * a[2][0]);
}
__device__ static M3 inverse3(const M3& m) {  // cofactors times 1 / det
    const double(*a)[3] = m.a;
    M3 c;
    c.a[0][0] = a[1][1] * a[2][2] - a[1][2] * a[2][1];
    c.a[0][1] = a[0][2] * a[2][1] - a[0][1] * a[2][2];
    c.a[0][2] = a[0][1] * a[1][2] - a[0][2] * a[1][1];
    c.a[1][0] = a[1][2] * a[2][0] - a[1][0] * a[2][2];
    c.a[1][1] = a[0][0] * a[2][2] - a[0][2] * a[2][0];
    c.a[1][2] = a[0][2] * a[1][0] - a[0][0] * a[1][2];
    c.a[2][0] = a[1][0] * a[2][1] - a[1][1] * a[2][0];
    c.a[2][1] = a[0][1] * a[2][0] - a[0][0] * a[2][1];
    c.a[2][2] = a[0][0] * a[1][1] - a[0][1] * a[1][0];
    const double d = a[0][0] * c.a[0][0] + a[0][1] * c.a[1][0] + a[0][2] * c.a[2][0];
    const double id = 1.0 / d;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) c.a[i][j] = id * c.a[i][j];
    return c;
}
__device__ static M3 quat_to_rot(const double q[4]) {  // Eigen toRotationMatrix() on the raw coefficients
    const double w = q[0], x = q[1], y = q[2], z = q[3];
    const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w;
    const double txx = tx * x, txy = ty * x, txz = tz * x;
    const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
    M3 r;
    r.a[0][0] = 1.0 - (tyy + tzz), r.a[0][1] = txy - twz, r.a[0][2] = txz + twy;
    r.a[1][0] = txy + twz, r.a[1][1] = 1.0 - (txx + tzz), r.a[1][2] = tyz - twx;
    r.a[2][0] = txz - twy, r.a[2][1] = tyz + twx, r.a[2][2] = 1.0 - (txx + tyy);
    return r;
}
__device__ static void rot_to_quat(const M3& m, double out[4]) {  // Eigen Quaterniond(Matrix3d)
    const double(*a)[3] = m.a;
    double q[3], w;
    double t = a[0][0] + a[1][1] + a[2][2];
    if (t > 0.0) {
        t = sqrt(t + 1.0);
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
        t = sqrt(a[i][i] - a[j][j] - a[k][k] + 1.0);
        q[i] = 0.5 * t;
        t = 0.5 / t;
        w = (a[k][j] - a[j][k]) * t;
        q[j] = (a[j][i] + a[i][j]) * t;
        q[k] = (a[k][i] + a[i][k]) * t;
    }
    out[0] = w, out[1] = q[0], out[2] = q[1], out[3] = q[2];
}
__device__ static void slerp(const double a[4], double t, const double b[4], double out[4]) {  // Eigen 3.3 slerp, not re-normalised
    const double one = 1.0 - 2.220446049250313e-16;
    const double d = a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3];
    const double ad = fabs(d);
    double s0, s1;
    if (ad >= one) {
        s0 = 1.0 - t;
        s1 = t;
    } else {
        const double th = acos(ad), st = sin(th);
        s0 = sin((1.0 - t) * th) / st;
        s1 = sin(t * th) / st;
    }
    if (d < 0.0) s1 = -s1;
    for (int i = 0; i < 4; ++i) out[i] = s0 * a[i] + s1 * b[i];
}
// cyclic Jacobi, eigenvalues descending, U canonicalised to det +1 (hrm_math.hpp symEig3)
__device__ static void sym_eig3(const M3& A, double ev[3], M3& U) {
    double a[3][3], v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) a[i][j] = 0.5 * (A.a[i][j] + A.a[j][i]);
    for (int sweep = 0; sweep < 64; ++sweep) {
        const double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
        const double diag = a[0][0] * a[0][0] + a[1][1] * a[1][1] + a[2][2] * a[2][2];
        if (off <= 1e-32 * diag || off == 0.0) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (a[p][q] == 0.0) continue;
                const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
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
            if (a[idx[j]][idx[j]] > a[idx[i]][idx[i]]) {
                const int t = idx[i];
                idx[i] = idx[j], idx[j] = t;
            }
    for (int c = 0; c < 3; ++c) {
        ev[c] = a[idx[c]][idx[c]];
        for (int r = 0; r < 3; ++r) U.a[r][c] = v[r][idx[c]];
    }
    if (det3(U) < 0.0)
        for (int r = 0; r < 3; ++r) U.a[r][2] = -U.a[r][2];
}
// TightFitEllipsoid.cpp:43-81
__device__ static void mvce3d(const double a[3], const double b[3], const double quatA[4], const double quatB[4], double semi[3], double quat[4]) {
    const M3 Ra = quat_to_rot(quatA), Rb = quat_to_rot(quatB);
    const double r = fmin(b[0], fmin(b[1], b[2]));
    const double dg[3] = {r / b[0], r / b[1], r / b[2]};
    const double dA[3] = {pow(a[0], -2.0), pow(a[1], -2.0), pow(a[2], -2.0)};
    const M3 T = mat_mul(mul_diag(Rb, dg), transpose(Rb));
    const M3 Ti = inverse3(T);
    const M3 aP = mat_mul(mat_mul(Ti, mat_mul(mul_diag(Ra, dA), transpose(Ra))), Ti);
    double ev[3];
    M3 U;
    sym_eig3(aP, ev, U);
    double dC[3];
    for (int i = 0; i < 3; ++i) dC[i] = pow(fmax(pow(ev[i], -0.5), r), -2.0);
    const M3 C = mat_mul(mat_mul(mul_diag(mat_mul(T, U), dC), transpose(U)), T);
    sym_eig3(C, ev, U);
    rot_to_quat(U, quat);
    for (int i = 0; i < 3; ++i) semi[i] = pow(ev[i], -0.5);
}

// body_semi [B][3], quat_a / quat_b [P][B][4]  ->  semi_out [P][B][3], rot_out [P][B][9] (row-major), quat_out [P][B][4]
__global__ void __launch_bounds__(64) k_tfe3d(int PB, int B, const double* body_semi, const double* quat_a, const double* quat_b, int num_step, double* semi_out,
                                              double* rot_out, double* quat_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= PB) return;
    const int b = i % B;
    double a[3], qa[4], qb[4];
    for (int d = 0; d < 3; ++d) a[d] = body_semi[3 * b + d];
    for (int d = 0; d < 4; ++d) qa[d] = quat_a[4 * static_cast<size_t>(i) + d], qb[d] = quat_b[4 * static_cast<size_t>(i) + d];
    double semi[3], quat[4];
    mvce3d(a, a, qa, qb, semi, quat);  // getTFE3D, :98-114
    const double dt = 1.0 / static_cast<double>(num_step);
    for (int s = 1; s < num_step; ++s) {
        double qi[4], ns[3], nq[4];
        slerp(qa, static_cast<double>(s) * dt, qb, qi);
        mvce3d(a, semi, qi, quat, ns, nq);
        for (int d = 0; d < 3; ++d) semi[d] = ns[d];
        for (int d = 0; d < 4; ++d) quat[d] = nq[d];
    }
    const M3 R = quat_to_rot(quat);  // (build_bridge takes the rotation matrix of the ellipsoid's quaternion)
    for (int d = 0; d < 3; ++d) semi_out[3 * static_cast<size_t>(i) + d] = semi[d];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) rot_out[9 * static_cast<size_t>(i) + 3 * r + c] = R.a[r][c];
    for (int d = 0; d < 4; ++d) quat_out[4 * static_cast<size_t>(i) + d] = quat[d];
}

int launch_tfe3d(Ctx* c, int P, int B, const double* d_body_semi, const double* d_quat_a, const double* d_quat_b, int num_step, double* d_semi,
                 double* d_rot, double* d_quat) {
    const int PB = P * B;
    if (PB <= 0) return HRMGPU_OK;
    k_tfe3d<<<(PB + 63) / 64, 64, 0, c->stream>>>(PB, B, d_body_semi, d_quat_a, d_quat_b, num_step, d_semi, d_rot, d_quat);
    HRMGPU_CUDA(c, cudaGetLastError());
    c->launches += 1;
    return HRMGPU_OK;
}

void preload_tfe() {
    cudaFuncAttributes a;
    if (cudaFuncGetAttributes(&a, k_tfe3d) != cudaSuccess) cudaGetLastError();
}

}  // namespace hrmgpu
