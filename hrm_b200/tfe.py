"""Host-side (Python binding) tightly-fitted ellipsoids and pose interpolation — the inputs of the bridge-layer
kernels.  Pure Python floats (IEEE double, no FMA), stated in the same operation order everywhere so results are
reproducible bit for bit.

Mirrors the reference's src/geometry/TightFitEllipsoid.cpp (getMVCE2D/3D :6-81, getTFE2D/3D :83-114) and
src/util/InterpolateSE3.cpp (slerp / R^n / compound interpolation :5-122).  Eigen::JacobiSVD on the symmetric
positive-definite matrices involved is replaced by a cyclic Jacobi eigen-solver (singular values == eigenvalues,
descending, U = V); U is canonicalised to det +1.
"""
import math

from . import hostmath as hm

_EPS = 2.220446049250313e-16


def quat_dot(a, b):
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3]


def quat_mul(a, b):
    return [a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
            a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
            a[0] * b[2] + a[2] * b[0] + a[3] * b[1] - a[1] * b[3],
            a[0] * b[3] + a[3] * b[0] + a[1] * b[2] - a[2] * b[1]]


def angular_distance(a, b):
    """Eigen 3.3 QuaternionBase::angularDistance (used by HRM3D.cpp:168,456-468)."""
    d = quat_mul(a, [b[0], -b[1], -b[2], -b[3]])
    return 2.0 * math.atan2(math.sqrt(d[1] * d[1] + d[2] * d[2] + d[3] * d[3]), math.fabs(d[0]))


def slerp(a, t, b):
    """Eigen 3.3 QuaternionBase::slerp; result not re-normalised."""
    one = 1.0 - _EPS
    d = quat_dot(a, b)
    ad = math.fabs(d)
    if ad >= one:
        s0, s1 = 1.0 - t, t
    else:
        th = math.acos(ad)
        st = math.sin(th)
        s0 = math.sin((1.0 - t) * th) / st
        s1 = math.sin(t * th) / st
    if d < 0.0:
        s1 = -s1
    return [s0 * a[0] + s1 * b[0], s0 * a[1] + s1 * b[1], s0 * a[2] + s1 * b[2], s0 * a[3] + s1 * b[3]]


def interpolate_slerp(qa, qb, num_step):  # InterpolateSE3.cpp:89-103
    dt = 1.0 / float(num_step)
    return [slerp(qa, float(i) * dt, qb) for i in range(num_step + 1)]


def interpolate_rn(v0, v1, num_step):  # :104-122
    dt = 1.0 / float(num_step)
    return [[(1.0 - float(i) * dt) * v0[j] + float(i) * dt * v1[j] for j in range(len(v0))] for i in range(num_step + 1)]


def interpolate_se3(v0, v1, num_step):  # :5-29, vertices (x,y,z,qw,qx,qy,qz)
    q = interpolate_slerp(v0[3:7], v1[3:7], num_step)
    t = interpolate_rn(v0[0:3], v1[0:3], num_step)
    return [t[i] + q[i] for i in range(num_step + 1)]


# ---- small dense helpers with a fixed accumulation order ((a0*b0 + a1*b1) + a2*b2) -------------------------------
def _mm(a, b):
    return [[a[i][0] * b[0][j] + a[i][1] * b[1][j] + a[i][2] * b[2][j] for j in range(3)] for i in range(3)]


def _t(a):
    return [[a[j][i] for j in range(3)] for i in range(3)]


def _mul_diag(a, d):
    return [[a[i][j] * d[j] for j in range(3)] for i in range(3)]


def _det(a):
    return (a[0][0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1]) - a[0][1] * (a[1][0] * a[2][2] - a[1][2] * a[2][0]) +
            a[0][2] * (a[1][0] * a[2][1] - a[1][1] * a[2][0]))


def _inv(a):
    c = [[0.0] * 3 for _ in range(3)]
    c[0][0] = a[1][1] * a[2][2] - a[1][2] * a[2][1]
    c[0][1] = a[0][2] * a[2][1] - a[0][1] * a[2][2]
    c[0][2] = a[0][1] * a[1][2] - a[0][2] * a[1][1]
    c[1][0] = a[1][2] * a[2][0] - a[1][0] * a[2][2]
    c[1][1] = a[0][0] * a[2][2] - a[0][2] * a[2][0]
    c[1][2] = a[0][2] * a[1][0] - a[0][0] * a[1][2]
    c[2][0] = a[1][0] * a[2][1] - a[1][1] * a[2][0]
    c[2][1] = a[0][1] * a[2][0] - a[0][0] * a[2][1]
    c[2][2] = a[0][0] * a[1][1] - a[0][1] * a[1][0]
    d = a[0][0] * c[0][0] + a[0][1] * c[1][0] + a[0][2] * c[2][0]
    idet = 1.0 / d
    return [[idet * c[i][j] for j in range(3)] for i in range(3)]


def sym_eig3(A):
    """Eigenvalues (descending) and eigenvectors (columns of U, det +1) of a symmetric 3x3 by cyclic Jacobi."""
    a = [[0.5 * (A[i][j] + A[j][i]) for j in range(3)] for i in range(3)]
    v = [[1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]]
    for _ in range(64):
        off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2]
        diag = a[0][0] * a[0][0] + a[1][1] * a[1][1] + a[2][2] * a[2][2]
        if off <= 1e-32 * diag or off == 0.0:
            break
        for p in range(2):
            for q in range(p + 1, 3):
                if a[p][q] == 0.0:
                    continue
                theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q])
                t = (1.0 if theta >= 0.0 else -1.0) / (math.fabs(theta) + math.sqrt(theta * theta + 1.0))
                c = 1.0 / math.sqrt(t * t + 1.0)
                s = t * c
                for k in range(3):
                    akp, akq = a[k][p], a[k][q]
                    a[k][p] = c * akp - s * akq
                    a[k][q] = s * akp + c * akq
                for k in range(3):
                    apk, aqk = a[p][k], a[q][k]
                    a[p][k] = c * apk - s * aqk
                    a[q][k] = s * apk + c * aqk
                for k in range(3):
                    vkp, vkq = v[k][p], v[k][q]
                    v[k][p] = c * vkp - s * vkq
                    v[k][q] = s * vkp + c * vkq
    idx = [0, 1, 2]
    for i in range(2):
        for j in range(i + 1, 3):
            if a[idx[j]][idx[j]] > a[idx[i]][idx[i]]:
                idx[i], idx[j] = idx[j], idx[i]
    ev = [a[idx[c]][idx[c]] for c in range(3)]
    U = [[v[r][idx[c]] for c in range(3)] for r in range(3)]
    if _det(U) < 0.0:
        for r in range(3):
            U[r][2] = -U[r][2]
    return ev, U


def get_mvce3d(a, b, quat_a, quat_b):
    """TightFitEllipsoid.cpp:43-81 -> (semi[3], quat[4])."""
    Ra, Rb = hm.quat_to_rot(quat_a), hm.quat_to_rot(quat_b)
    r = math.fmin(b[0], math.fmin(b[1], b[2])) if hasattr(math, "fmin") else min(b[0], min(b[1], b[2]))
    dg = [r / b[0], r / b[1], r / b[2]]
    dA = [math.pow(a[0], -2.0), math.pow(a[1], -2.0), math.pow(a[2], -2.0)]
    T = _mm(_mul_diag(Rb, dg), _t(Rb))
    Ti = _inv(T)
    aP = _mm(_mm(Ti, _mm(_mul_diag(Ra, dA), _t(Ra))), Ti)
    ev, U = sym_eig3(aP)
    a_prime = [math.pow(e, -0.5) for e in ev]
    c_prime = [max(x, r) for x in a_prime]
    dC = [math.pow(x, -2.0) for x in c_prime]
    C = _mm(_mm(_mul_diag(_mm(T, U), dC), _t(U)), T)
    ev, U = sym_eig3(C)
    return [math.pow(e, -0.5) for e in ev], hm.rot_to_quat(U)


def get_tfe3d(a, quat_a, quat_b, num_step):
    """TightFitEllipsoid.cpp:98-114."""
    qi = interpolate_slerp(quat_a, quat_b, num_step)
    semi, quat = get_mvce3d(a, a, quat_a, quat_b)
    for i in range(1, num_step):
        semi, quat = get_mvce3d(a, semi, qi[i], quat)
    return semi, quat


# ---- 2D ---------------------------------------------------------------------------------------------------------
def _sym_eig2(A):
    a, b, d = A[0][0], 0.5 * (A[0][1] + A[1][0]), A[1][1]
    tr, df = a + d, a - d
    rt = math.sqrt(df * df + 4.0 * b * b)
    ev = [0.5 * (tr + rt), 0.5 * (tr - rt)]
    if math.fabs(b) > 0.0:
        vx, vy = ev[0] - d, b
        if math.fabs(vx) + math.fabs(vy) == 0.0:
            vx, vy = b, ev[0] - a
    elif a >= d:
        vx, vy = 1.0, 0.0
    else:
        vx, vy = 0.0, 1.0
    n = math.sqrt(vx * vx + vy * vy)
    vx /= n
    vy /= n
    return ev, [[vx, -vy], [vy, vx]]


def get_mvce2d(a, b, theta_a, theta_b):
    """TightFitEllipsoid.cpp:6-41 -> (semi[2], angle)."""
    Ra = [[math.cos(theta_a), -math.sin(theta_a)], [math.sin(theta_a), math.cos(theta_a)]]
    Rb = [[math.cos(theta_b), -math.sin(theta_b)], [math.sin(theta_b), math.cos(theta_b)]]
    r = min(b[0], b[1])
    dg = [r / b[0], r / b[1]]
    dA = [math.pow(a[0], -2), math.pow(a[1], -2)]

    def rdrt(R, d):
        return [[(R[i][0] * d[0]) * R[j][0] + (R[i][1] * d[1]) * R[j][1] for j in range(2)] for i in range(2)]

    def mul(X, Y):
        return [[X[i][0] * Y[0][j] + X[i][1] * Y[1][j] for j in range(2)] for i in range(2)]

    T = rdrt(Rb, dg)
    Am = rdrt(Ra, dA)
    dt = T[0][0] * T[1][1] - T[0][1] * T[1][0]
    idt = 1.0 / dt
    Ti = [[T[1][1] * idt, -T[0][1] * idt], [-T[1][0] * idt, T[0][0] * idt]]
    aP = mul(mul(Ti, Am), Ti)
    ev, U = _sym_eig2(aP)
    a_prime = [math.pow(ev[0], -0.5), math.pow(ev[1], -0.5)]
    c_prime = [max(a_prime[0], r), max(a_prime[1], r)]
    dC = [math.pow(c_prime[0], -2), math.pow(c_prime[1], -2)]
    Ut = [[U[0][0], U[1][0]], [U[0][1], U[1][1]]]
    TU = mul(T, U)
    TUD = [[TU[i][j] * dC[j] for j in range(2)] for i in range(2)]
    C = mul(mul(TUD, Ut), T)
    ev, U = _sym_eig2(C)
    angle = math.acos(max(-1.0, min(1.0, U[0][0])))
    return [math.pow(ev[0], -0.5), math.pow(ev[1], -0.5)], angle


def get_tfe2d(a, theta_a, theta_b, num_step):
    """TightFitEllipsoid.cpp:83-96."""
    semi, angle = get_mvce2d(a, a, theta_a, theta_b)
    dt = 1.0 / (float(num_step) - 1)
    for i in range(num_step):
        ci = float(i)
        theta_step = (1 - ci * dt) * theta_a + ci * dt * theta_b
        semi, angle = get_mvce2d(a, semi, theta_step, angle)
    return semi, angle
