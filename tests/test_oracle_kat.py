"""Pins the CPU oracle against every known-answer value the reference's own tests hold for the hot path
(/root/reference/test/TestGeometry.cpp) and against analytic identities for the parts the reference leaves unpinned."""
import numpy as np

from tests import util


def test_constants(oracle):
    c = oracle.constants()  # include/hrm/datastructure/DataType.h:30-32
    assert c["PI"] == 3.1415926 and c["HALF_PI"] == 3.1415926 / 2.0 and c["EPSILON"] == 1e-6


def test_superellipse_boundary_sampling(oracle):
    """TestGeometry.cpp:11-20 — ASSERT_DOUBLE_EQ (4 ULP) on point 0."""
    b = oracle.origin_shape2d([5.0, 3.0, 1.25, -2.6, 3.2, 0.0], 50)
    assert abs(b[0, 0] - 2.4) <= 4 * np.spacing(2.4)
    assert abs(b[1, 0] - 3.2) <= 4 * np.spacing(3.2)


def test_superellipse_minkowski_boundary_sampling(oracle):
    """TestGeometry.cpp:22-32 — the only Minkowski KAT of the reference."""
    m = oracle.minksum2d([5.0, 3.0, 1.25, 3.2, -2.5, 0.0], [2.5, 1.5], 0.0, 50, +1)
    assert abs(m[0, 0] - 10.7) <= 4 * np.spacing(10.7)
    assert abs(m[1, 0] + 2.5) <= 4 * np.spacing(2.5)


def test_superquadrics_boundary_sampling(oracle):
    """TestGeometry.cpp:35-50 — samples within 1e-6 of the axis have z = 6 or 2 (the 2x20 pole samples)."""
    o = oracle.origin_shape3d([5.0, 3.0, 2.0, 1.25, 0.3, 2.32, -1.5, 4.0, 0.0, 1.0, 0.0, 0.0], 20)
    sel = (np.abs(o[0] - 2.32) < 1e-6) & (np.abs(o[1] + 1.5) < 1e-6)
    assert sel.sum() == 40
    assert np.all((np.abs(o[2, sel] - 6.0) < 1e-6) | (np.abs(o[2, sel] - 2.0) < 1e-6))


def test_superquadrics_minkowski_sampling_is_vacuous_but_holds(oracle):
    """TestGeometry.cpp:52-71 — the reference's condition never fires for n=20 (SURVEY.md finding 10); whatever samples
    do come within 1e-6 of the axis must satisfy its assertion."""
    m = oracle.minksum3d([5.0, 3.0, 2.0, 1.25, 0.3, -3.2, 3.24, -2.13, 0.0, 1.0, 0.0, 0.0], [2.5, 1.5, 1.0], [0.0, 1.0, 0.0, 0.0], 20, +1)
    sel = (np.abs(m[0] + 3.2) < 1e-6) & (np.abs(m[1] - 3.24) < 1e-6)
    assert np.all((np.abs(m[2, sel] - (3.0 - 2.13)) < 1e-6) | (np.abs(m[2, sel] - (-3.0 - 2.13)) < 1e-6))
    # analytic anchor: the surface must reach the poles z = pz +/- (c1 + c2) up to the grid resolution
    assert abs(m[2].max() - (3.0 - 2.13)) < 1e-3 and abs(m[2].min() - (-3.0 - 2.13)) < 1e-3


def test_mvce2d(oracle):
    """TestGeometry.cpp:74-82"""
    s = oracle.mvce2d([5.0, 3.0], [5.0, 3.0], 0.0, oracle.constants()["HALF_PI"])
    assert abs(s[0] - 5.0) < 1e-6 and abs(s[1] - 5.0) < 1e-6


def test_mvce3d(oracle):
    """TestGeometry.cpp:84-96"""
    s = oracle.mvce3d([5.0, 3.0, 2.0], [2.0, 3.0, 5.0], [0.0, 1.0, 0.0, 0.0], [0.0, 1.0, 0.0, 0.0])
    assert abs(s[0] - 3.0) < 1e-6 and abs(s[1] - 5.0) < 1e-6 and abs(s[2] - 5.0) < 1e-6


def test_sphere_plus_sphere_is_a_sphere(oracle):
    """eps = (1,1): sphere (+) sphere of radii 3 and 2 is a sphere of radius 5 about the centre; (-) gives radius 1."""
    c = np.array([[1.0], [2.0], [3.0]])
    for K, r in ((+1, 5.0), (-1, 1.0)):
        m = oracle.minksum3d([3, 3, 3, 1, 1, 1, 2, 3, 1, 0, 0, 0], [2, 2, 2], [0.5, -0.5, 0.5, 0.5], 15, K)
        assert np.max(np.abs(np.linalg.norm(m - c, axis=0) - r)) < 1e-12


def test_minkowski_contains_origin_shape_direction(oracle):
    """x = x0 + K * T^2 R1 g / |T R1 g|: the offset has the sign of K along the outward normal."""
    sq = [6, 4, 3, 0.8, 1.3, 0, 0, 0, 1, 0, 0, 0]
    o = oracle.origin_shape3d(sq, 12)
    mp = oracle.minksum3d(sq, [2, 1.5, 1], [1, 0, 0, 0], 12, +1)
    mm = oracle.minksum3d(sq, [2, 1.5, 1], [1, 0, 0, 0], 12, -1)
    assert np.allclose(mp - o, -(mm - o), rtol=0, atol=1e-12)
    assert np.all(np.linalg.norm(mp, axis=0) >= np.linalg.norm(o, axis=0) - 1e-9)


def test_eigen_conventions(oracle):
    # toRotationMatrix of a unit quaternion is orthonormal; rot_to_quat inverts it (up to sign)
    q = np.array([0.5, -0.5, 0.5, 0.5])
    R = oracle.quat_to_rot(q)
    assert np.allclose(R @ R.T, np.eye(3), atol=1e-15)
    q2 = oracle.rot_to_quat(R)
    assert np.allclose(q2, q) or np.allclose(q2, -q)
    # slerp end points, angular distance of a quaternion with itself / its negative
    a, b = np.array([1.0, 0, 0, 0]), np.array([0.0, 1, 0, 0])
    assert np.allclose(oracle.slerp(a, 0.0, b), a) and np.allclose(oracle.slerp(a, 1.0, b), b)
    assert abs(oracle.angular_distance(a, a)) < 1e-15 and abs(oracle.angular_distance(a, -a)) < 1e-15
    assert abs(oracle.angular_distance(a, b) - np.pi) < 1e-12
    # LinSpaced: last element is exactly `high`
    v = oracle.linspaced(10, -1.5707963 - 1e-6, 1.5707963 + 1e-6)
    assert v[-1] == 1.5707963 + 1e-6 and v[0] == -1.5707963 - 1e-6
    # the un-normalised quaternion is used as is (finding 4)
    assert not np.allclose(oracle.quat_to_rot([0.9, 0, 0, 0.5]) @ oracle.quat_to_rot([0.9, 0, 0, 0.5]).T, np.eye(3))


def test_power_tables_are_the_matrix_form_values(oracle):
    """The separable 8-table form used for the GPU reproduces the per-element matrix form bit for bit."""
    sq = [5.0, 3.0, 2.0, 0.3, 1.7, 0.1, -0.2, 0.3, 0.9, 0.1, -0.3, 0.2]
    n = 9
    t = oracle.sq_tables(sq, n)
    o = oracle.origin_shape3d(sq, n)
    R = oracle.quat_to_rot(sq[8:])
    for c in range(n):
        for r in range(n):
            X = np.array([sq[0] * (t[0, c] * t[4, r]), sq[1] * (t[0, c] * t[5, r]), sq[2] * (t[1, c] * 1.0)])
            want = [(R[i, 0] * X[0] + R[i, 1] * X[1] + R[i, 2] * X[2]) + sq[5 + i] for i in range(3)]
            assert np.array_equal(o[:, r + c * n], want)


def test_interval_algebra_and_enhancement(oracle):
    """Interval.cpp / FreeSpace-inl.h:101-174 on a hand-checked case: touching obstacle intervals merge, zero-length
    segments are kept, enhancement appends the neighbour's end points and sorts xL/xU/xM independently."""
    nan = np.nan
    alo, aup = np.array([[-10.0], [-10.0], [-10.0]]), np.array([[10.0], [10.0], [10.0]])
    olo = np.array([[-2.0, 1.0, nan], [8.0, nan, nan], [nan, -3.0, -3.0]])
    oup = np.array([[1.0, 4.0, nan], [12.0, nan, nan], [nan, -3.0, 2.0]])
    off, xL, xU, xM = oracle.free_segments(alo, aup, olo, oup)
    # line 0: obstacles [-2,1],[1,4] touch -> merge to [-2,4]; free [-10,-2], [4,10] (no neighbour end point falls inside)
    # line 1: free [-10,8] ((12,inf) clipped to the arena is empty); pair (0,1) appends -2 and 4, pair (1,2) appends -3 and 2,
    #         then xL, xU, xM are sorted independently
    # line 2 (last, never sorted): originals [-10,-3],[2,10] then nothing appended
    assert off.tolist() == [0, 2, 7, 9]
    assert xL[:2].tolist() == [-10.0, 4.0] and xU[:2].tolist() == [-2.0, 10.0] and xM[:2].tolist() == [-6.0, 7.0]
    assert xL[2:7].tolist() == [-10.0, -3.0, -2.0, 2.0, 4.0]
    assert xU[2:7].tolist() == [-3.0, -2.0, 2.0, 4.0, 8.0]
    assert xM[2:7].tolist() == [-3.0, -2.0, -1.0, 2.0, 4.0]
    assert xL[7:].tolist() == [-10.0, 2.0] and xU[7:].tolist() == [-3.0, 10.0] and xM[7:].tolist() == [-6.5, 6.0]


def test_first_two_hits_in_face_order(oracle):
    """SURVEY.md finding 2: on the bundled eps=0.1 arena the mesh folds over itself; the oracle returns the hits of the
    two LOWEST face indices, which is not the min/max-z pair on a fair share of the sweep lines."""
    sc = util.scenes()["3D_superquadrics_cluttered_rabbit"]
    semi, R, off, quat = util.frozen_pose3d(sc["robot"])
    n = 10
    arena = oracle.minksum3d(sc["arena"][0], semi[0], quat[0][0], n, -1)
    lim, Lx, Ly = sc["lim"], sc["Lx"], sc["Ly"]
    two = 0
    for ix in range(Lx):
        for iy in range(Ly):
            x = lim[0] + ix * (lim[1] - lim[0]) / (Lx - 1)
            y = lim[2] + iy * (lim[3] - lim[2]) / (Ly - 1)
            c, z = oracle.vertical_line_mesh(arena, n, x, y)
            assert c in (0, 2)
            two += c == 2
    assert two == Lx * Ly


def test_planner_solves_reference_test_scenes(oracle):
    """test/TestHRM3D.cpp (sparse/rabbit, n=10, q_icosahedron_60) and test/TestHRM2D.cpp::MultiBody (sparse/rabbit, 50
    points, 10 slices, numPoint 5, auto numLineY = 3) assert only `solved`."""
    S = util.scenes()
    sc = S["3D_superquadrics_sparse_rabbit"]
    r = oracle.plan3d(1, len(sc["obstacle"]), np.array(sc["arena"] + sc["obstacle"]), 10, np.array(sc["robot"]), np.array(sc["quats"]),
                      sc["lim"], sc["Lx"], sc["Ly"], 5, sc["end_points"][0], sc["end_points"][1])
    assert r["solved"] and len(r["vertex"]) > 100 and len(r["path"]) >= 2
    sc = S["2D_superellipse_sparse_rabbit"]
    r = oracle.plan2d(1, len(sc["obstacle"]), np.array(sc["arena"] + sc["obstacle"]), 50, np.array(sc["robot"]), 10, sc["lim"], sc["Ly"],
                      5, sc["end_points"][0], sc["end_points"][1])
    assert sc["Ly"] == 3 and r["solved"]
