"""Pins the CPU oracle against every known-answer value the reference's own tests hold for the hot path
(/root/reference/test/TestGeometry.cpp) and against analytic identities for the parts the reference leaves unpinned."""
import os

import numpy as np
import pytest

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


# ---- articulated path: URDF revolute tree, forward kinematics, ProbHRM3D (oracle/hrmo_articulated.hpp) -------------------
# KDL / urdfdom are third-party dependencies absent from the reference tree: the FK restatement is pinned by analytic
# identities, the planner by the assertion of the reference's own test (test/TestProbHRM3D.cpp:8-62: `solved`).
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_fk_zero_configuration_is_a_translation_chain(oracle):
    urdf = os.path.join(GOLDEN, "snake.urdf")
    for body, x in ((0, 0.0), (1, 7.0), (2, 15.2), (3, 23.4)):  # joint origins 7.0, +8.2, +8.2 along x
        T, nj = oracle.fk(urdf, [0.0, 0.0, 0.0], body)
        assert nj == 3
        assert np.allclose(T[:3, :3], np.eye(3), atol=1e-15) and np.allclose(T[:3, 3], [x, 0, 0], atol=1e-14)


def test_fk_known_rotations(oracle):
    urdf = os.path.join(GOLDEN, "snake.urdf")
    h = np.pi / 2
    # joint 1 about +y by 90 deg: the chain's x axis turns to -z; body2's origin sits 8.2 along the rotated x
    T2, _ = oracle.fk(urdf, [h, 0.0, 0.0], 2)
    assert np.allclose(T2[:3, 3], [7.0, 0.0, -8.2], atol=1e-14)
    assert np.allclose(T2[:3, :3], [[0, 0, 1], [0, 1, 0], [-1, 0, 0]], atol=1e-15)
    # joint 2 about +z by 90 deg (joint 1 at rest): body3's origin moves from x to y
    T3, _ = oracle.fk(urdf, [0.0, h, 0.0], 3)
    assert np.allclose(T3[:3, 3], [15.2, 8.2, 0.0], atol=1e-14)
    # composition: FK(q) of body3 == T01(q0) T12(q1) T23(q2) built by hand
    q = [0.3, -0.7, 1.1]

    def rot(axis, a):
        c, s = np.cos(a), np.sin(a)
        return {"y": np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]]), "z": np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])}[axis]

    def hom(R, t):
        M = np.eye(4)
        M[:3, :3], M[:3, 3] = R, t
        return M

    want = hom(rot("y", q[0]), [7.0, 0, 0]) @ hom(rot("z", q[1]), [8.2, 0, 0]) @ hom(rot("y", q[2]), [8.2, 0, 0])
    got, _ = oracle.fk(urdf, q, 3)
    assert np.allclose(got, want, atol=1e-14)


def test_fk_tree_joint_numbering(oracle):
    """tree.urdf: three chains off the base; KDL numbers joints depth first in name order, so joint k drives body k+1."""
    urdf = os.path.join(GOLDEN, "tree.urdf")
    T0, nj = oracle.fk(urdf, np.zeros(9), 9)
    assert nj == 9 and np.allclose(T0[:3, 3], [0, 0, 2.1 + 8.2 + 8.2], atol=1e-14)
    for k in range(9):
        q = np.zeros(9)
        q[k] = 0.5
        for body in range(1, 10):
            T, _ = oracle.fk(urdf, q, body)
            base, _ = oracle.fk(urdf, np.zeros(9), body)
            chain = (body - 1) // 3
            moved = not np.allclose(T, base, atol=1e-15)
            assert moved == (chain == k // 3 and body - 1 >= k), (k, body)


def test_robot_tf_urdf_matches_rigid_tf_at_zero_joints(oracle):
    """With all joints at rest the articulated transform must place the links where the rigid-body tree does, given that the
    robot file's link poses are expressed relative to the URDF frames of body<k> (resources/3D/robot_snake_3D.csv)."""
    sc = util.scenes()["3D_superquadrics_sparse_snake"]
    urdf = os.path.join(GOLDEN, "snake.urdf")
    base = [1.0, -2.0, 0.5, 0.5, 0.5, 0.5, 0.5]
    poses = oracle.robot_tf_urdf(10, np.array(sc["robot"]), urdf, base, [0.0, 0.0, 0.0])
    R = np.array([[0, 0, 1], [1, 0, 0], [0, 1, 0]], dtype=float)  # rotation of the quaternion (.5,.5,.5,.5)
    assert np.allclose(poses[0, :3], base[:3]) and np.allclose(poses[0, 3:], base[3:], atol=1e-15)
    for k, x in enumerate((7.0 + 4.1, 15.2 + 4.1, 23.4 + 4.1)):
        assert np.allclose(poses[k + 1, :3], np.array(base[:3]) + R @ np.array([x, 0, 0]), atol=1e-13)
        assert np.allclose(np.abs(poses[k + 1, 3:]), 0.5, atol=1e-15)


def _halton(i, base):
    f, r = 1.0, 0.0
    while i > 0:
        f /= base
        r += f * (i % base)
        i //= base
    return r


def prob_samples(count, nj, max_joint=np.pi / 2):
    """Deterministic stand-in for Quaterniond::UnitRandom() + ompl::RNG::uniformReal: Halton points mapped to unit quaternions
    (Shoemake) and joint angles in [-max_joint, max_joint]."""
    out = []
    primes = [2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37]
    for i in range(1, count + 1):
        u1, u2, u3 = _halton(i, 2), _halton(i, 3), _halton(i, 5)
        q = [np.sqrt(u1) * np.cos(2 * np.pi * u3), np.sqrt(1 - u1) * np.sin(2 * np.pi * u2), np.sqrt(1 - u1) * np.cos(2 * np.pi * u2),
             np.sqrt(u1) * np.sin(2 * np.pi * u3)]
        joints = [(2 * _halton(i, primes[3 + j]) - 1) * max_joint for j in range(nj)]
        out.append(q + joints)
    return np.array(out)


@pytest.mark.parametrize("per_orientation", [False, True])
def test_prob_hrm3d_sparse_snake_solves(oracle, per_orientation):
    """test/TestProbHRM3D.cpp: superquadrics/sparse/snake, n = 10, default sweep lines -> solved.  (With the reference's
    search rule -- start and goal are matched to their nearest vertices of the slices built so far -- the sparse map is
    already solved inside the first slice.)"""
    sc = util.scenes()["3D_superquadrics_sparse_snake"]
    rows = np.array(sc["arena"] + sc["obstacle"])
    res = oracle.plan_prob3d(len(sc["arena"]), len(sc["obstacle"]), rows, 10, np.array(sc["robot"]), os.path.join(GOLDEN, "snake.urdf"),
                             prob_samples(40, 3), sc["lim"], sc["Lx"], sc["Ly"], 5, sc["end_points"][0], sc["end_points"][1],
                             per_orientation=per_orientation)
    assert res["solved"] and res["slices"] >= 1
    assert res["vertex"].shape[1] == 10  # x, y, z, quaternion, 3 joint angles
    assert np.array_equal(res["vertex"][0, 7:], np.array(sc["end_points"][0][7:10]))  # slice 0 carries the start joint angles
    assert len(res["path"]) >= 2


def test_prob_hrm3d_home_snake_needs_several_slices(oracle):
    """BASELINE.json configs[3] scene (home map, snake): with per-orientation layers the start slice alone does not connect
    start and goal, so the random slices, the nearest-slice bridge (TFE over the interpolated articulated motion) and the
    joint-carrying vertices are all exercised."""
    sc = util.scenes()["3D_superquadrics_home_snake"]
    rows = np.array(sc["arena"] + sc["obstacle"])
    res = oracle.plan_prob3d(1, len(sc["obstacle"]), rows, 10, np.array(sc["robot"]), os.path.join(GOLDEN, "snake.urdf"), prob_samples(30, 3),
                             sc["lim"], sc["Lx"], sc["Ly"], 5, sc["end_points"][0], sc["end_points"][1], per_orientation=True)
    assert res["solved"] and res["slices"] > 2
    joints = np.unique(res["vertex"][:, 7:], axis=0)
    assert len(joints) == res["slices"]  # one joint configuration per slice
    assert np.all(np.abs(joints) <= np.pi / 2 + 1e-6)


def test_refinement_solves_when_first_roadmap_does_not(oracle):
    """HighwayRoadMap::plan refines (doubled sweep lines on the cached C-boundaries, connectExistSlice) while unsolved.  The
    reference itself throws here (interval tables keep their initial size, SURVEY.md finding 6); the restatement sizes them by
    the current line count: 5 x 3 lines do not connect start and goal on the narrow map, the first refinement round does."""
    sc = util.scenes()["3D_superquadrics_narrow_rabbit"]
    rows = np.array(sc["arena"] + sc["obstacle"])
    args = (1, len(sc["obstacle"]), rows, 10, np.array(sc["robot"]), np.array(sc["quats"]), sc["lim"], 5, 3, 5, sc["end_points"][0],
            sc["end_points"][1])
    try:
        oracle.set_max_refine(0)
        first = oracle.plan3d(*args)
        oracle.set_max_refine(2)
        refined = oracle.plan3d(*args)
    finally:
        oracle.set_max_refine(0)
    assert not first["solved"] and first["refine_rounds"] == 0
    assert refined["solved"] and refined["refine_rounds"] == 1
    nv = len(first["vertex"])
    assert np.array_equal(refined["vertex"][:nv], first["vertex"])  # the first roadmap is kept, refinement only appends
    assert np.array_equal(refined["edge"][: len(first["edge"])], first["edge"])
    assert len(refined["vertex"]) > nv


def test_hoisted_oracle_build_is_bit_identical(oracle):
    """oracle/libhrm_oracle_fast.so (-DHRMO_FAST: power function on n values instead of the n x n grids, mesh bounding boxes once
    per mesh -- the `fast` CPU baseline of SURVEY.md §8d) against the faithful build: every point, interval and segment equal."""
    from tests import util

    rows = util.synthetic_scene3d(9, 12)
    rng = np.random.default_rng(9)
    robot = [[10, 5, 4, 1, 1, 0, 0, 0, 1, 0, 0, 0], [6, 3, 2, 1, 1, -7, 0, 4, 0.70710599, 0, 0.70710802, 0]]
    semi, R, off, quat = util.body_poses3d(robot, util.random_unit_quats(rng, 2))
    lim = [-58.0, 58.0, -28.0, 28.0, -18.0, 18.0]
    for k in range(2):
        a = oracle.layer3d(1, 12, rows, 18, semi, quat[k], off[k], lim, 12, 6)
        b = oracle.layer3d(1, 12, rows, 18, semi, quat[k], off[k], lim, 12, 6, fast=True)
        assert a["single_hits"] == b["single_hits"] and a["off"][-1] > 0
        for key in ("bound", "lo", "up", "off", "xL", "xU", "xM"):
            assert np.array_equal(a[key], b[key], equal_nan=True), key
