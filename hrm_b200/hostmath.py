"""Host-side rotation helpers of the Python binding (pure Python floats = IEEE double, no FMA).

They restate the Eigen conventions the reference relies on so that the matrices handed to the C ABI are
exactly the ones the reference would hold: Quaterniond::toRotationMatrix() on un-normalised quaternions
(reference src/geometry/SuperQuadrics.cpp:78,109-110), Quaterniond(Matrix3d)
(src/datastructure/MultiBodyTree3D.cpp:38-50) and Quaterniond(AngleAxisd)
(src/test/util/ParsePlanningSettings.cpp:197-199).
"""
import math

PI = 3.1415926  # reference include/hrm/datastructure/DataType.h:30-32 (truncated on purpose)
HALF_PI = PI / 2.0
EPSILON = 1e-6


def quat_to_rot(q):
    """q = (w, x, y, z) -> 3x3 row-major nested list; no normalisation."""
    w, x, y, z = (float(v) for v in q)
    tx, ty, tz = 2.0 * x, 2.0 * y, 2.0 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    return [
        [1.0 - (tyy + tzz), txy - twz, txz + twy],
        [txy + twz, 1.0 - (txx + tzz), tyz - twx],
        [txz - twy, tyz + twx, 1.0 - (txx + tyy)],
    ]


def rot_to_quat(m):
    t = m[0][0] + m[1][1] + m[2][2]
    q = [0.0, 0.0, 0.0]
    if t > 0.0:
        t = math.sqrt(t + 1.0)
        w = 0.5 * t
        t = 0.5 / t
        q[0] = (m[2][1] - m[1][2]) * t
        q[1] = (m[0][2] - m[2][0]) * t
        q[2] = (m[1][0] - m[0][1]) * t
    else:
        i = 0
        if m[1][1] > m[0][0]:
            i = 1
        if m[2][2] > m[i][i]:
            i = 2
        j = (i + 1) % 3
        k = (j + 1) % 3
        t = math.sqrt(m[i][i] - m[j][j] - m[k][k] + 1.0)
        q[i] = 0.5 * t
        t = 0.5 / t
        w = (m[k][j] - m[j][k]) * t
        q[j] = (m[j][i] + m[i][j]) * t
        q[k] = (m[k][i] + m[i][k]) * t
    return [w, q[0], q[1], q[2]]


def mat_mul(a, b):
    return [[a[i][0] * b[0][j] + a[i][1] * b[1][j] + a[i][2] * b[2][j] for j in range(3)] for i in range(3)]


def mat_vec(a, v):
    return [a[i][0] * v[0] + a[i][1] * v[1] + a[i][2] * v[2] for i in range(3)]


def angle_axis_to_quat(angle, axis):
    ax = [float(v) for v in axis]
    z = ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]
    if z > 0.0:
        s = math.sqrt(z)
        ax = [v / s for v in ax]
    ha = 0.5 * angle
    s = math.sin(ha)
    return [math.cos(ha), s * ax[0], s * ax[1], s * ax[2]]


def flat9(m):
    return [m[0][0], m[0][1], m[0][2], m[1][0], m[1][1], m[1][2], m[2][0], m[2][1], m[2][2]]
