"""Planner-level parity on the reference's bundled maps: the GPU-backed HRM3D / HRM2D (hrm_b200/planner.py: K1-K5 on the
GPU, host replay) must produce the SAME roadmap as the CPU oracle's restatement of the reference planners — identical
vertex list (bit for bit in STRICT mode), identical edge list, identical planning success."""
import numpy as np
import pytest

from hrm_b200 import capi
from hrm_b200.planner import HRM2D, HRM3D, PlanningRequest
from tests import util

pytestmark = pytest.mark.gpu


def request3d(sc, num_point=5):
    req = PlanningRequest()
    req.start, req.goal = list(sc["end_points"][0][:7]), list(sc["end_points"][1][:7])
    req.parameters.boundaryLimits = list(sc["lim"])
    req.parameters.numLineX, req.parameters.numLineY = sc["Lx"], sc["Ly"]
    req.parameters.numPoint = num_point
    return req


@pytest.mark.parametrize("name", ["3D_superquadrics_sparse_rabbit", "3D_superquadrics_cluttered_rabbit", "3D_superquadrics_maze_rabbit",
                                  "3D_superquadrics_narrow_rabbit"])
@pytest.mark.parametrize("per_orientation", [False, True])
def test_hrm3d_roadmap_identical(oracle, gpu_ctx, name, per_orientation):
    sc = util.scenes()[name]
    n = 10
    rows = np.array(sc["arena"] + sc["obstacle"])
    ref = oracle.plan3d(len(sc["arena"]), len(sc["obstacle"]), rows, n, np.array(sc["robot"]), np.array(sc["quats"]), sc["lim"], sc["Lx"],
                        sc["Ly"], 5, sc["end_points"][0], sc["end_points"][1], per_orientation=per_orientation)
    hrm = HRM3D(util.robot3d(sc["robot"], n), util.sq_objects(sc["arena"], n), util.sq_objects(sc["obstacle"], n), request3d(sc),
                sc["quats"], ctx=gpu_ctx, precision=capi.F64_STRICT, per_orientation=per_orientation)
    hrm.plan()
    res = hrm.getPlanningResult()
    assert len(res.vertex) == len(ref["vertex"])
    assert np.array_equal(np.array(res.vertex), ref["vertex"])
    assert np.array_equal(np.array(res.edge, dtype=np.int64).reshape(-1, 2), ref["edge"])
    assert res.solved == ref["solved"]
    if ref["solved"]:
        assert res.PathId[0] == ref["path"][0] and res.PathId[-1] == ref["path"][-1]
    assert hrm.ctx.kernel_launches() > 0


def test_hrm3d_home_map_first_layers(oracle, gpu_ctx):
    """The 21-object home map (config 4 scene) restricted to 8 orientations to keep the oracle fast."""
    sc = util.scenes()["3D_superquadrics_home_rabbit"]
    n = 10
    quats = sc["quats"][:8]
    rows = np.array(sc["arena"] + sc["obstacle"])
    ref = oracle.plan3d(1, len(sc["obstacle"]), rows, n, np.array(sc["robot"]), np.array(quats), sc["lim"], sc["Lx"], sc["Ly"], 5,
                        sc["end_points"][0], sc["end_points"][1], per_orientation=True)
    hrm = HRM3D(util.robot3d(sc["robot"], n), util.sq_objects(sc["arena"], n), util.sq_objects(sc["obstacle"], n), request3d(sc), quats,
                ctx=gpu_ctx, precision=capi.F64_STRICT, per_orientation=True)
    hrm.plan()
    res = hrm.getPlanningResult()
    assert np.array_equal(np.array(res.vertex), ref["vertex"])
    assert np.array_equal(np.array(res.edge, dtype=np.int64).reshape(-1, 2), ref["edge"])
    assert res.solved == ref["solved"]


@pytest.mark.parametrize("name,Ly", [("2D_superellipse_sparse_rabbit", None), ("2D_superellipse_sparse_rabbit", 30),
                                     ("2D_superellipse_cluttered_rabbit", 30), ("2D_superellipse_maze2_rabbit", 30),
                                     ("2D_ellipse_sparse_rabbit", 20)])
@pytest.mark.parametrize("per_orientation", [False, True])
def test_hrm2d_roadmap_identical(oracle, gpu_ctx, name, Ly, per_orientation):
    sc = util.scenes()[name]
    Ly = Ly or sc["Ly"]
    num, num_slice = 50, 10 if Ly < 10 else 20  # TestHRM2D.cpp: 50 / 10; README bench example: 20 slices, 30 lines
    rows = np.array(sc["arena"] + sc["obstacle"])
    ref = oracle.plan2d(len(sc["arena"]), len(sc["obstacle"]), rows, num, np.array(sc["robot"]), num_slice, sc["lim"], Ly, 5,
                        sc["end_points"][0], sc["end_points"][1], per_orientation=per_orientation)
    req = PlanningRequest()
    req.start, req.goal = list(sc["end_points"][0]), list(sc["end_points"][1])
    req.parameters.boundaryLimits = list(sc["lim"])
    req.parameters.numLineY, req.parameters.numSlice, req.parameters.numPoint = Ly, num_slice, 5
    hrm = HRM2D(util.robot2d(sc["robot"], num), util.se_objects(sc["arena"], num), util.se_objects(sc["obstacle"], num), req, ctx=gpu_ctx,
                precision=capi.F64_STRICT, per_orientation=per_orientation)
    hrm.plan()
    res = hrm.getPlanningResult()
    assert np.array_equal(np.array(res.vertex), ref["vertex"])
    assert np.array_equal(np.array(res.edge, dtype=np.int64).reshape(-1, 2), ref["edge"])
    assert res.solved == ref["solved"]


def test_hrm3d_fast_mode_same_connectivity(oracle, gpu_ctx):
    """The bench's arithmetic mode (FP64 with FMA) on the reference's TestHRM3D scene: same vertex count, same edges,
    same success; vertices within 1e-6."""
    sc = util.scenes()["3D_superquadrics_sparse_rabbit"]
    n = 10
    rows = np.array(sc["arena"] + sc["obstacle"])
    ref = oracle.plan3d(1, len(sc["obstacle"]), rows, n, np.array(sc["robot"]), np.array(sc["quats"]), sc["lim"], sc["Lx"], sc["Ly"], 5,
                        sc["end_points"][0], sc["end_points"][1], per_orientation=True)
    hrm = HRM3D(util.robot3d(sc["robot"], n), util.sq_objects(sc["arena"], n), util.sq_objects(sc["obstacle"], n), request3d(sc),
                sc["quats"], ctx=gpu_ctx, precision=capi.F64, per_orientation=True)
    hrm.plan()
    res = hrm.getPlanningResult()
    assert len(res.vertex) == len(ref["vertex"])
    assert np.max(np.abs(np.array(res.vertex) - ref["vertex"])) <= 1e-6
    assert np.array_equal(np.array(res.edge, dtype=np.int64).reshape(-1, 2), ref["edge"])
    assert res.solved == ref["solved"] is True


# ---- the same parity through the C++ host layer (libhrmhost.so: hrm::planners::HRM3D / HRM2D over the C ABI) ----------
@pytest.mark.parametrize("name", ["3D_superquadrics_sparse_rabbit", "3D_superquadrics_cluttered_rabbit", "3D_superquadrics_maze_rabbit",
                                  "3D_superquadrics_narrow_rabbit"])
@pytest.mark.parametrize("per_orientation", [False, True])
def test_cpp_host_hrm3d_roadmap_identical(oracle, name, per_orientation):
    from hrm_b200 import hostlib

    sc = util.scenes()[name]
    n = 10
    rows = np.array(sc["arena"] + sc["obstacle"])
    args = (len(sc["arena"]), len(sc["obstacle"]), rows, n, np.array(sc["robot"]), np.array(sc["quats"]), sc["lim"], sc["Lx"], sc["Ly"], 5,
            sc["end_points"][0], sc["end_points"][1])
    ref = oracle.plan3d(*args, per_orientation=per_orientation)
    got = hostlib.plan3d(*args, per_orientation=per_orientation, precision=capi.F64_STRICT)
    assert np.array_equal(got["vertex"], ref["vertex"])
    assert np.array_equal(got["edge"], ref["edge"])
    assert got["solved"] == ref["solved"]
    assert got["edge_tests"] > 0 and got["gpu_launches"] > 0
    if ref["solved"]:
        assert got["path"][0] == ref["path"][0] and got["path"][-1] == ref["path"][-1]


@pytest.mark.parametrize("name,Ly", [("2D_superellipse_sparse_rabbit", 3), ("2D_superellipse_cluttered_rabbit", 30),
                                     ("2D_superellipse_maze2_rabbit", 30)])
@pytest.mark.parametrize("per_orientation", [False, True])
def test_cpp_host_hrm2d_roadmap_identical(oracle, name, Ly, per_orientation):
    from hrm_b200 import hostlib

    sc = util.scenes()[name]
    num, num_slice = 50, 10 if Ly < 10 else 20
    rows = np.array(sc["arena"] + sc["obstacle"])
    args = (len(sc["arena"]), len(sc["obstacle"]), rows, num, np.array(sc["robot"]), num_slice, sc["lim"], Ly, 5, sc["end_points"][0],
            sc["end_points"][1])
    ref = oracle.plan2d(*args, per_orientation=per_orientation)
    got = hostlib.plan2d(*args, per_orientation=per_orientation, precision=capi.F64_STRICT)
    assert np.array_equal(got["vertex"], ref["vertex"])
    assert np.array_equal(got["edge"], ref["edge"])
    assert got["solved"] == ref["solved"]


# ---- ProbHRM3D: articulated robot, batched C-layers, injected draws (hrm_b200/host/hrm_articulated.hpp) ---------------------
def _prob_case(oracle, name, per_orientation, n_samples, batch):
    import os

    from hrm_b200 import hostlib
    from tests.test_oracle_kat import GOLDEN, prob_samples

    sc = util.scenes()[name]
    rows = np.array(sc["arena"] + sc["obstacle"])
    tree = name.endswith("_tree")  # tree.urdf: three chains off the base, 9 joints; snake.urdf: one chain, 3 joints
    urdf = os.path.join(GOLDEN, "tree.urdf" if tree else "snake.urdf")
    draws = prob_samples(n_samples, 9 if tree else 3)
    nj = draws.shape[1] - 4
    ep = [list(e) + [0.0] * (7 + nj - len(e)) for e in sc["end_points"]]  # (tree: the bundled end points name 6 of the 9 joints)
    args = (len(sc["arena"]), len(sc["obstacle"]), rows, 10, np.array(sc["robot"]), urdf, draws, sc["lim"], sc["Lx"], sc["Ly"], 5, ep[0], ep[1])
    ref = oracle.plan_prob3d(*args, per_orientation=per_orientation)
    got = hostlib.plan_prob3d(*args, per_orientation=per_orientation, precision=capi.F64_STRICT, batch=batch)
    return ref, got


@pytest.mark.parametrize("name,per_orientation,batch", [("3D_superquadrics_sparse_snake", False, 8), ("3D_superquadrics_cluttered_snake", True, 4),
                                                        ("3D_superquadrics_home_snake", True, 8), ("3D_superquadrics_home_snake", True, 3),
                                                        ("3D_superquadrics_sparse_tree", True, 4), ("3D_superquadrics_sparse_tree", False, 8)])
def test_prob_hrm3d_roadmap_identical(oracle, name, per_orientation, batch):
    """Same draws -> the GPU-backed ProbHRM3D (layers built `batch` slices ahead) must reproduce the oracle's sequential
    planner: slice count, vertex list (with joint angles), edge list, solved, path."""
    ref, got = _prob_case(oracle, name, per_orientation, 30, batch)
    assert got["slices"] == ref["slices"]
    assert got["vertex"].shape == ref["vertex"].shape
    assert np.array_equal(got["vertex"], ref["vertex"])
    assert np.array_equal(got["edge"], ref["edge"])
    assert got["solved"] == ref["solved"]
    assert np.array_equal(got["path"], ref["path"])
    assert got["gpu_launches"] > 0
    if name.endswith("home_snake"):
        assert ref["slices"] > 2  # several random slices and nearest-slice bridges were needed


def test_prob_hrm3d_slice_by_slice_path_identical(oracle, monkeypatch):
    """HRM_PROB_PER_SLICE=1: one bridge build and one matching call per slice instead of the batch staged ahead of its turn
    (stageBatch) -- the same roadmap as the oracle, hence as the staged path the other tests run."""
    monkeypatch.setenv("HRM_PROB_PER_SLICE", "1")
    ref, got = _prob_case(oracle, "3D_superquadrics_home_snake", True, 30, 8)
    assert np.array_equal(got["vertex"], ref["vertex"]) and np.array_equal(got["edge"], ref["edge"])
    assert got["solved"] == ref["solved"] and np.array_equal(got["path"], ref["path"])
    per_slice_launches = got["gpu_launches"]
    monkeypatch.delenv("HRM_PROB_PER_SLICE")
    _, got2 = _prob_case(oracle, "3D_superquadrics_home_snake", True, 30, 8)
    staged_launches = got2["gpu_launches"]
    assert np.array_equal(got2["vertex"], got["vertex"]) and np.array_equal(got2["edge"], got["edge"])
    assert staged_launches < per_slice_launches  # fewer, larger launches


def test_prob_hrm3d_stops_when_draws_run_out(oracle):
    """Too few draws to solve the per-orientation home map: both planners stop unsolved after the same number of slices."""
    ref, got = _prob_case(oracle, "3D_superquadrics_home_snake", True, 3, 8)
    assert not ref["solved"] and not got["solved"]
    assert got["slices"] == ref["slices"] == 5
    assert np.array_equal(got["vertex"], ref["vertex"]) and np.array_equal(got["edge"], ref["edge"])


# ---- refinement path: hrmgpu_resweep + connectExistSlice ----------------------------------------------------
@pytest.fixture
def refine_rounds(oracle):
    from hrm_b200 import hostlib

    def set_rounds(n):
        oracle.set_max_refine(n)
        hostlib.set_max_refine(n)

    yield set_rounds
    set_rounds(0)


@pytest.mark.parametrize("name,nq,Lx,Ly,per_orientation,rounds", [("3D_superquadrics_narrow_rabbit", 60, 5, 3, False, 2),
                                                                  ("3D_superquadrics_maze_rabbit", 12, 4, 2, True, 2),
                                                                  ("3D_superquadrics_cluttered_rabbit", 8, 3, 2, True, 1)])
def test_hrm3d_refinement_identical(oracle, refine_rounds, name, nq, Lx, Ly, per_orientation, rounds):
    """Too few sweep lines for the first roadmap -> refineExistRoadmap: the C++ host planner re-sweeps all cached layers in one
    hrmgpu_resweep call per round and must reproduce the oracle's slice-by-slice refinement (vertices, edges, early stop)."""
    from hrm_b200 import hostlib

    sc = util.scenes()[name]
    rows = np.array(sc["arena"] + sc["obstacle"])
    args = (1, len(sc["obstacle"]), rows, 10, np.array(sc["robot"]), np.array(sc["quats"][:nq]), sc["lim"], Lx, Ly, 5, sc["end_points"][0],
            sc["end_points"][1])
    refine_rounds(rounds)
    ref = oracle.plan3d(*args, per_orientation=per_orientation)
    got = hostlib.plan3d(*args, per_orientation=per_orientation, precision=capi.F64_STRICT)
    if name.endswith("cluttered_rabbit"):
        assert ref["refine_rounds"] == 0  # solved at once: the refinement loop must not run
    else:
        assert ref["refine_rounds"] >= 1
    assert got["refine_rounds"] == ref["refine_rounds"]
    assert np.array_equal(got["vertex"], ref["vertex"])
    assert np.array_equal(got["edge"], ref["edge"])
    assert got["solved"] == ref["solved"]
    assert np.array_equal(got["path"], ref["path"])
    if name.endswith("narrow_rabbit"):
        assert ref["solved"] and ref["refine_rounds"] == 1  # found during the first refinement round


@pytest.mark.parametrize("name,Ly,per_orientation", [("2D_superellipse_maze2_rabbit", 6, False), ("2D_superellipse_cluttered_rabbit", 5, True)])
def test_hrm2d_refinement_identical(oracle, refine_rounds, name, Ly, per_orientation):
    from hrm_b200 import hostlib

    sc = util.scenes()[name]
    rows = np.array(sc["arena"] + sc["obstacle"])
    args = (len(sc["arena"]), len(sc["obstacle"]), rows, 50, np.array(sc["robot"]), 10, sc["lim"], Ly, 5, sc["end_points"][0], sc["end_points"][1])
    refine_rounds(3)
    ref = oracle.plan2d(*args, per_orientation=per_orientation)
    got = hostlib.plan2d(*args, per_orientation=per_orientation, precision=capi.F64_STRICT)
    assert ref["refine_rounds"] == 3 and got["refine_rounds"] == 3
    assert np.array_equal(got["vertex"], ref["vertex"])
    assert np.array_equal(got["edge"], ref["edge"])
    assert got["solved"] == ref["solved"]


def test_free_space_layer_view_follows_the_reference_call_order(oracle):
    """hrm::FreeSpaceLayer3D (host/hrm_host.hpp) presents one C-layer of a batched GPU build behind the reference's
    FreeSpaceComputator method set; driven in the order of HRM3D::constructOneSlice / sweepLineProcess (HRM3D.cpp:24-54,63-91) --
    per x-plane computeIntersectionInterval({tx, ty}), computeFreeSegment(ty), getFreeSegment() -- it must hand out exactly the
    oracle's boundary points, interval tables and free segments of every layer."""
    from hrm_b200 import hostlib

    sc = util.scenes()["3D_superquadrics_cluttered_rabbit"]
    rows = np.array(sc["arena"] + sc["obstacle"])
    quats = np.array(sc["quats"][:5])
    got = hostlib.layer_views3d(1, len(sc["obstacle"]), rows, 10, np.array(sc["robot"]), quats, sc["lim"], sc["Lx"], sc["Ly"], precision=0)
    semi, R, off, bq = util.body_poses3d(sc["robot"], quats)
    for k in range(5):
        ref = oracle.layer3d(1, len(sc["obstacle"]), rows, 10, semi, bq[k], off[k], sc["lim"], sc["Lx"], sc["Ly"])
        assert np.array_equal(got[k]["lo"], ref["lo"], equal_nan=True) and np.array_equal(got[k]["up"], ref["up"], equal_nan=True)
        assert np.array_equal(got[k]["off"], ref["off"])
        for key in ("xL", "xU", "xM"):
            assert np.array_equal(got[k][key], ref[key]), key
        if k == 0:
            assert np.array_equal(got[0]["bound"], ref["bound"])
    assert sum(len(g["xM"]) for g in got) > 100


def test_prob_hrm3d_tree_robot_bridges_identical(oracle):
    """tree.urdf (three chains off the base) on the home map: 12 draws do not solve it, so all 14 slices are built and every one
    is bridged to its nearest earlier slice through the articulated TFE / transition tests -- the GPU-backed planner (look-ahead
    batches, device matching) must give the oracle's roadmap vertex for vertex and edge for edge."""
    import os

    from hrm_b200 import hostlib
    from tests.test_oracle_kat import GOLDEN, prob_samples

    S = util.scenes()
    sc, tree = S["3D_superquadrics_home_snake"], S["3D_superquadrics_sparse_tree"]
    rows = np.array(sc["arena"] + sc["obstacle"])
    ep0 = list(sc["end_points"][0][:7]) + list(tree["end_points"][0][7:])
    ep1 = list(sc["end_points"][1][:7]) + list(tree["end_points"][1][7:])
    ep0, ep1 = ep0 + [0.0] * (16 - len(ep0)), ep1 + [0.0] * (16 - len(ep1))  # (the bundled end points name 6 of the 9 joints)
    args = (len(sc["arena"]), len(sc["obstacle"]), rows, 10, np.array(tree["robot"]), os.path.join(GOLDEN, "tree.urdf"), prob_samples(12, 9), sc["lim"],
            sc["Lx"], sc["Ly"], 5, ep0, ep1)
    ref = oracle.plan_prob3d(*args, per_orientation=True)
    got = hostlib.plan_prob3d(*args, per_orientation=True, precision=capi.F64_STRICT, batch=5)
    assert ref["slices"] == got["slices"] == 14 and not ref["solved"] and not got["solved"]
    assert np.array_equal(got["vertex"], ref["vertex"])
    assert np.array_equal(got["edge"], ref["edge"]) and len(ref["edge"]) > 10000


@pytest.mark.parametrize("Lx,Ly,draws,period,batch", [(6, 6, 6, 4, 8), (6, 5, 7, 4, 3)])
def test_prob_hrm3d_periodic_refinement_identical(oracle, Lx, Ly, draws, period, batch):
    """ProbHRM3D doubles the sweep lines of all slices every 60 slices (ProbHRM3D.cpp:52-56); with the period set to 4 the home
    map sees two refinements within 8-9 slices (6x6 -> 24x24 lines).  The GPU-backed planner rebuilds all slices in one call per
    refinement and must reproduce the oracle's roadmap (vertex ids, bridge vertices, connectExistSlice edges) exactly -- once
    unsolved after 8 slices, once solved at slice 9."""
    import os

    from hrm_b200 import hostlib
    from tests.test_oracle_kat import GOLDEN, prob_samples

    sc = util.scenes()["3D_superquadrics_home_snake"]
    rows = np.array(sc["arena"] + sc["obstacle"])
    args = (len(sc["arena"]), len(sc["obstacle"]), rows, 10, np.array(sc["robot"]), os.path.join(GOLDEN, "snake.urdf"), prob_samples(draws, 3), sc["lim"],
            Lx, Ly, 5, sc["end_points"][0], sc["end_points"][1])
    oracle.set_prob_refine_period(period)
    hostlib.set_prob_refine_period(period)
    try:
        ref = oracle.plan_prob3d(*args, per_orientation=True)
        got = hostlib.plan_prob3d(*args, per_orientation=True, precision=capi.F64_STRICT, batch=batch)
    finally:
        oracle.set_prob_refine_period(60)
        hostlib.set_prob_refine_period(60)
    assert got["slices"] == ref["slices"] and got["solved"] == ref["solved"]
    assert len(ref["vertex"]) > 7000  # (two refinements happened: ~1500 vertices without them)
    assert np.array_equal(got["vertex"], ref["vertex"])
    assert np.array_equal(got["edge"], ref["edge"])
    assert np.array_equal(got["path"], ref["path"])
