"""End-to-end run of hrm_b200/lib/hrm_run (one table-driven driver for HRM3D / HRM2D / ProbHRM3D): resources tree ->
parsePlanningConfig -> loadEnvironment -> GPU-backed planner -> result/benchmark/time_*.csv (the column schema the reference's
plotting scripts read) + vertex/edge/path/segment dumps.
The resources tree is rebuilt from tests/golden/scenes.json (the reference's own files are not on the GPU box); the oracle
plans on exactly the numbers the CLI parsed (hostlib.load_scene on the same config directory)."""
import math
import os
import subprocess

import numpy as np
import pytest

from hrm_b200 import hostlib
from tests import util

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "hrm_b200", "lib", "hrm_run")


def _angle_axis_row(row):
    """[a,b,c,e1,e2,x,y,z,qw,qx,qy,qz] -> the resource format [..., ax,ay,az,angle]."""
    w, x, y, z = row[8:12]
    nv = math.sqrt(x * x + y * y + z * z)
    axis = [x / nv, y / nv, z / nv] if nv > 0 else [0.0, 0.0, 1.0]
    return list(row[:8]) + axis + [2.0 * math.atan2(nv, w)]


def _write(path, rows):
    with open(path, "w") as f:
        f.write("# generated from tests/golden/scenes.json\n")
        for r in rows:
            f.write(",".join("%.9g" % v for v in r) + "\n")


def make_resources(sc, root, env="t", robot="r"):
    os.makedirs(root / "3D" / "urdf"), os.makedirs(root / "SO3_sequence")
    _write(root / "3D" / f"env_superquadrics_{env}_3D_arena.csv", [_angle_axis_row(r) for r in sc["arena"]])
    _write(root / "3D" / f"env_superquadrics_{env}_3D_obstacle.csv", [_angle_axis_row(r) for r in sc["obstacle"]])
    _write(root / "3D" / f"robot_{robot}_3D.csv", [_angle_axis_row(r) for r in sc["robot"]])
    ends = []
    for e in sc["end_points"]:
        aa = _angle_axis_row([0] * 8 + list(e[3:7]))
        ends.append(list(e[:3]) + aa[8:12] + list(e[7:]))  # joint angles of articulated robots follow the base pose
    _write(root / "3D" / f"setting_superquadrics_{env}_3D.csv", ends)
    _write(root / "SO3_sequence" / f"q_icosahedron_{len(sc['quats'])}.csv", [[q[1], q[2], q[3], q[0]] for q in sc["quats"]])


@pytest.mark.parametrize("name", ["3D_superquadrics_sparse_rabbit", "3D_superquadrics_cluttered_rabbit"])
def test_bench_hrm3d_cli(oracle, name, tmp_path):
    sc = util.scenes()[name]
    res, cfg, out = tmp_path / "resources", tmp_path / "config", tmp_path / "result"
    make_resources(sc, res)
    os.makedirs(cfg), os.makedirs(out / "details"), os.makedirs(out / "benchmark")
    env = dict(os.environ, HRM_RESOURCES_PATH=str(res), HRM_CONFIG_PATH=str(cfg), HRM_RESULT_PATH=str(out))
    ns = len(sc["quats"])
    p = subprocess.run([EXE, "hrm3d", "map=t", "robot=r", "trials=2", "time=60", f"slices={ns}", "so3=icosahedron", "details=1"], env=env,
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    assert f"{ns} C-slices" in p.stdout and "trial 2/2: solved=" in p.stdout

    # what the CLI planned on, and the oracle's roadmap for exactly those numbers
    got = hostlib.load_scene("3D", str(cfg) + "/", str(res / "SO3_sequence" / f"q_icosahedron_{ns}.csv"))
    assert got["numSlice"] == ns and np.array_equal(got["quats"], np.array(sc["quats"]))
    assert np.allclose(got["obstacle"], np.array(sc["obstacle"]), rtol=2e-5, atol=2e-6)  # %g round trip of the angle-axis detour
    rows = np.vstack([got["arena"], got["obstacle"]])
    ref = oracle.plan3d(len(got["arena"]), len(got["obstacle"]), rows, 10, got["robot"], got["quats"], got["lim"], got["Lx"], got["Ly"], 5,
                        got["end_points"][0][:7], got["end_points"][1][:7], per_orientation=False)

    stats = (out / "benchmark" / "time_high_3D.csv").read_text().strip().split("\n")
    assert stats[0] == "SUCCESS,BUILD_TIME,SEARCH_TIME,PLAN_TIME,N_LAYERS,N_X,N_Y,GRAPH_NODE,GRAPH_EDGE,PATH_NODE"
    assert len(stats) == 3
    for line in stats[1:]:
        c = line.split(",")
        assert int(c[0]) == int(ref["solved"]) and int(c[4]) == ns and (int(c[5]), int(c[6])) == (got["Lx"], got["Ly"])
        assert int(c[7]) == len(ref["vertex"]) and int(c[8]) == len(ref["edge"]) and int(c[9]) == len(ref["path"])

    V = np.array([[float(v) for v in l.rstrip(",").split(",")] for l in (out / "details" / "vertex_hrm_3D.csv").read_text().strip().split("\n")])
    assert V.shape == ref["vertex"].shape and np.allclose(V, ref["vertex"], rtol=1e-5, atol=1e-5)  # file holds 6 significant digits
    E = np.array([[int(v) for v in l.split(",")] for l in (out / "details" / "edge_hrm_3D.csv").read_text().strip().split("\n")])
    assert np.array_equal(E, ref["edge"])
    ids = [int(v) for v in (out / "details" / "path_id_hrm_3D.csv").read_text().rstrip(",").split(",") if v]
    assert ids == [int(v) for v in ref["path"]]
    if ref["solved"]:  # HRM3D::getInterpolatedSolutionPath: start, path vertices, goal, densified
        sol = (out / "details" / "solution_path_hrm_3D.csv").read_text().strip().split("\n")
        ip = (out / "details" / "interpolated_path_hrm_3D.csv").read_text().strip().split("\n")
        assert len(sol) == len(ids) + 2 and len(ip) >= len(sol)
        assert all(len(l.rstrip(",").split(",")) == 7 for l in ip)
        assert ip[0] == sol[0] and ip[-1] == sol[-1]
    seg = (out / "details" / "segment_hrm_3D.csv").read_text().strip().split("\n")
    assert len(seg) > 0 and all(len(l.split(",")) == 5 for l in seg)
    bound = (out / "details" / "mink_bound_hrm_3D.csv").read_text().rstrip("\n").split("\n")
    assert len(bound) == 3 * (len(got["arena"]) + len(got["obstacle"])) * len(got["robot"])
    assert all(len(l.split()) == 100 for l in bound)


def test_bench_prob_hrm3d_cli(tmp_path):
    """hrm_run prob3d on the home map with the snake robot: resources -> config (robot type "snake" keeps 3 joint angles in
    the end points) -> URDF FK -> batched C-layers -> solved, with the reference's CSV columns."""
    import shutil

    sc = util.scenes()["3D_superquadrics_home_snake"]
    res, cfg, out = tmp_path / "resources", tmp_path / "config", tmp_path / "result"
    make_resources(sc, res, env="home", robot="snake")
    shutil.copy(os.path.join(ROOT, "tests", "golden", "snake.urdf"), res / "3D" / "urdf" / "snake.urdf")
    os.makedirs(cfg), os.makedirs(out / "details"), os.makedirs(out / "benchmark")
    env = dict(os.environ, HRM_RESOURCES_PATH=str(res), HRM_CONFIG_PATH=str(cfg), HRM_RESULT_PATH=str(out))
    p = subprocess.run([EXE, "prob3d", "map=home", "robot=snake", "trials=2", "time=120", "per_orientation=1", "seed=5"], env=env,
                       capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "trial 2/2: solved=1" in p.stdout
    ends = (cfg / "end_points_3D.csv").read_text().strip().split("\n")
    assert all(len(l.split(",")) == 10 for l in ends)  # x,y,z, quaternion, 3 joints (parseStartGoalConfig, robotType == "snake")
    stats = (out / "benchmark" / "time_prob_high_3D.csv").read_text().strip().split("\n")
    assert stats[0] == "SUCCESS,PLAN_TIME,N_LAYERS,N_X,N_Y,GRAPH_NODE,GRAPH_EDGE,PATH_NODE"
    assert len(stats) == 3
    for line in stats[1:]:
        c = line.split(",")
        assert int(c[0]) == 1 and int(c[2]) >= 1 and int(c[5]) > 0 and int(c[6]) > 0 and int(c[7]) >= 2


def test_bench_hrm2d_cli(oracle, tmp_path):
    """hrm_run hrm2d (the reference README's example: sparse rabbit, 20 slices, 30 sweep lines): same roadmap size as the oracle planning on the
    numbers the CLI parsed; 2D resource rows are already in their final form (a,b,eps,x,y,angle)."""
    from hrm_b200 import hostlib

    sc = util.scenes()["2D_superellipse_sparse_rabbit"]
    res, cfg, out = tmp_path / "resources", tmp_path / "config", tmp_path / "result"
    os.makedirs(res / "2D"), os.makedirs(cfg), os.makedirs(out / "details"), os.makedirs(out / "benchmark")
    _write(res / "2D" / "env_superellipse_t_2D_arena.csv", sc["arena"])
    _write(res / "2D" / "env_superellipse_t_2D_obstacle.csv", sc["obstacle"])
    _write(res / "2D" / "robot_r_2D.csv", sc["robot"])
    _write(res / "2D" / "setting_superellipse_t_2D.csv", sc["end_points"])
    env = dict(os.environ, HRM_RESOURCES_PATH=str(res), HRM_CONFIG_PATH=str(cfg), HRM_RESULT_PATH=str(out))
    p = subprocess.run([EXE, "hrm2d", "map=t", "robot=r", "trials=2", "slices=20", "ly=30"], env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    got = hostlib.load_scene("2D", str(cfg) + "/", num_param=50, num_line_y=30)
    rows = np.vstack([got["arena"], got["obstacle"]])
    ref = oracle.plan2d(len(got["arena"]), len(got["obstacle"]), rows, 50, got["robot"], 20, got["lim"], 30, 5, got["end_points"][0][:3],
                        got["end_points"][1][:3], per_orientation=False)
    stats = (out / "benchmark" / "time_hrm_2D.csv").read_text().strip().split("\n")
    assert stats[0] == "BUILD_TIME,SEARCH_TIME,PLAN_TIME,GRAPH_NODE,GRAPH_EDGE,PATH_NODE" and len(stats) == 3
    for line in stats[1:]:
        c = line.split(",")
        assert int(c[3]) == len(ref["vertex"]) and int(c[4]) == len(ref["edge"]) and int(c[5]) == len(ref["path"])
    assert ref["solved"] and "path_nodes=%d" % len(ref["path"]) in p.stdout
