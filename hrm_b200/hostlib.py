"""ctypes binding of libhrmhost.so — the C++ host layer (hrm_b200/host/hrm_host.hpp: hrm::planners::HRM3D / HRM2D,
FreeSpaceGPU3D/2D, MultiBodyTree*, TFE) that runs the reference's planner API on top of the CUDA C ABI."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libhrmhost.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libhrmhost.so not built: run `python -m hrm_b200.build`")
        _lib = C.CDLL(LIB_PATH)
        _lib.hrmhost_last_error.restype = C.c_char_p
    return _lib


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _out(info, vertex, edge, path):
    nv, ne, npth = int(info[1]), int(info[2]), int(info[3])
    return {"solved": bool(info[0]), "vertex": vertex[:nv].copy(), "edge": edge[:ne].copy(), "path": path[:npth].copy(),
            "edge_tests": int(info[5]), "gpu_launches": int(info[6]), "total_time_s": info[7] * 1e-6, "refine_rounds": int(info[4])}


def set_max_refine(rounds):
    """Refinement rounds (doubled sweep lines on the cached C-boundaries) plan3d / plan2d may run while unsolved; default 0."""
    load().hrmhost_set_max_refine(C.c_int(rounds))


def set_prob_refine_period(n):
    """ProbHRM3D doubles the sweep lines of all slices every `n` slices (60 in the reference, ProbHRM3D.cpp:52-56; 0 = never)."""
    load().hrmhost_set_prob_refine_period(C.c_int(n))


def plan3d(n_arena, n_obs, sq12, n, robot_rows, quats, lim, Lx, Ly, num_point, start, goal, per_orientation=False, precision=0, device=0,
           cap_v=1 << 18, cap_e=1 << 21):
    """hrm::planners::HRM3D(robot, arena, obstacle, request).plan() on the GPU-backed free space."""
    sq12, robot_rows, quats, lim = _d(sq12), _d(robot_rows), _d(quats), _d(lim)
    start, goal = _d(start)[:7].copy(), _d(goal)[:7].copy()
    info = np.zeros(8, dtype=np.int64)
    vertex, edge, path = np.zeros((cap_v, 7)), np.zeros((cap_e, 2), dtype=np.int64), np.zeros(cap_v, dtype=np.int64)
    rc = load().hrmhost_plan3d(C.c_int(n_arena), C.c_int(n_obs), _p(sq12), C.c_int(n), C.c_int(robot_rows.shape[0]), _p(robot_rows),
                               C.c_int(quats.shape[0]), _p(quats), _p(lim), C.c_int(Lx), C.c_int(Ly), C.c_int(num_point), _p(start), _p(goal),
                               C.c_int(int(per_orientation)), C.c_int(precision), C.c_int(device), _p(info), _p(vertex), C.c_long(cap_v),
                               _p(edge), C.c_long(cap_e), _p(path), C.c_long(cap_v))
    if rc != 0:
        raise RuntimeError(load().hrmhost_last_error().decode())
    return _out(info, vertex, edge, path)


def layer_views3d(n_arena, n_obs, sq12, n, robot_rows, quats, lim, Lx, Ly, precision=0, device=0, cap=1 << 16):
    """Builds one C-layer per quaternion with FreeSpaceGPU3D::buildLayers, then walks every layer through a FreeSpaceLayer3D view in
    the reference's per-layer call order (computeIntersectionInterval({tx, ty}) -> computeFreeSegment(ty) -> getFreeSegment() per
    x-plane).  Returns one dict per layer: lo, up [L][S], off [L+1], xL, xU, xM (and bound [S][3][M] for layer 0)."""
    sq12, robot_rows, quats, lim = _d(sq12), _d(robot_rows), _d(quats).reshape(-1, 4), _d(lim)
    K, B = quats.shape[0], robot_rows.shape[0]
    S, L, M = (n_arena + n_obs) * B, Lx * Ly, n * n
    lo, up = np.zeros((K, L, S)), np.zeros((K, L, S))
    off = np.zeros((K, L + 1), dtype=np.int32)
    xL, xU, xM = np.zeros((K, cap)), np.zeros((K, cap)), np.zeros((K, cap))
    cnt = np.zeros(K, dtype=np.int64)
    bound0 = np.zeros((S, M, 3))
    rc = load().hrmhost_layer_views3d(C.c_int(n_arena), C.c_int(n_obs), _p(sq12), C.c_int(n), C.c_int(B), _p(robot_rows), C.c_int(K), _p(quats),
                                      _p(lim), C.c_int(Lx), C.c_int(Ly), C.c_int(precision), C.c_int(device), _p(lo), _p(up), _p(off), _p(xL),
                                      _p(xU), _p(xM), C.c_long(cap), _p(cnt), _p(bound0))
    if rc != 0:
        raise _err()
    assert (cnt <= cap).all()
    return [{"lo": lo[k], "up": up[k], "off": off[k], "xL": xL[k, :cnt[k]].copy(), "xU": xU[k, :cnt[k]].copy(), "xM": xM[k, :cnt[k]].copy(),
             "bound": np.ascontiguousarray(bound0.transpose(0, 2, 1)) if k == 0 else None} for k in range(K)]


def plan2d(n_arena, n_obs, se6, num, robot_rows, num_slice, lim, Ly, num_point, start, goal, per_orientation=False, precision=0, device=0,
           cap_v=1 << 16, cap_e=1 << 20):
    """hrm::planners::HRM2D(robot, arena, obstacle, request).plan()."""
    se6, robot_rows, lim = _d(se6), _d(robot_rows), _d(lim)
    start, goal = _d(start)[:3].copy(), _d(goal)[:3].copy()
    info = np.zeros(8, dtype=np.int64)
    vertex, edge, path = np.zeros((cap_v, 3)), np.zeros((cap_e, 2), dtype=np.int64), np.zeros(cap_v, dtype=np.int64)
    rc = load().hrmhost_plan2d(C.c_int(n_arena), C.c_int(n_obs), _p(se6), C.c_int(num), C.c_int(robot_rows.shape[0]), _p(robot_rows),
                               C.c_int(num_slice), _p(lim), C.c_int(Ly), C.c_int(num_point), _p(start), _p(goal), C.c_int(int(per_orientation)),
                               C.c_int(precision), C.c_int(device), _p(info), _p(vertex), C.c_long(cap_v), _p(edge), C.c_long(cap_e), _p(path),
                               C.c_long(cap_v))
    if rc != 0:
        raise RuntimeError(load().hrmhost_last_error().decode())
    return _out(info, vertex, edge, path)


# ---- scene / config I/O (hrm_b200/host/hrm_io.hpp) --------------------------------------------------------------------
def _err():
    return RuntimeError(load().hrmhost_last_error().decode())


def parse_csv(filename, stride=32, cap_rows=1 << 16):
    """hrm::parse2DCsvFile: list of rows (std::stof cells, '#' comments skipped)."""
    lib = load()
    lib.hrmhost_parse_csv.restype = C.c_long
    out = np.zeros((cap_rows, stride))
    row_len = np.zeros(cap_rows, dtype=np.int32)
    n = lib.hrmhost_parse_csv(filename.encode(), _p(out), C.c_int(stride), _p(row_len), C.c_long(cap_rows))
    if n < 0:
        raise _err()
    return [out[i, :row_len[i]].tolist() for i in range(min(n, cap_rows))]


def parse_planning_config(object_type, env_type, robot_type, dim, resources, config):
    """hrm::parsePlanningConfig: resources/<dim>/*.csv (angle-axis) -> config/*_config_<dim>.csv (quaternion)."""
    rc = load().hrmhost_parse_planning_config(object_type.encode(), env_type.encode(), robot_type.encode(), dim.encode(), resources.encode(),
                                              config.encode())
    if rc != 0:
        raise _err()


def load_scene(dim, config_prefix, quat_file="0", num_param=10, num_line_x=0, num_line_y=0, cap_rows=4096):
    """PlannerSetting::loadEnvironment + loadRobotMultiBody{2,3}D + defineParameters on a config directory."""
    w = 12 if dim == "3D" else 6
    counts = np.zeros(8, dtype=np.int64)
    arena, obstacle, robot = np.zeros((cap_rows, w)), np.zeros((cap_rows, w)), np.zeros((cap_rows, w))
    ends, quats, lim = np.zeros((cap_rows, 16)), np.zeros((cap_rows, 4)), np.zeros(6)
    rc = load().hrmhost_load_scene(dim.encode(), config_prefix.encode(), quat_file.encode(), C.c_int(num_param), C.c_int(num_line_x),
                                   C.c_int(num_line_y), _p(counts), _p(arena), _p(obstacle), _p(robot), _p(ends), _p(quats), _p(lim),
                                   C.c_long(cap_rows))
    if rc != 0:
        raise _err()
    na, no, nr, ne, nq, lx, ly, ns = (int(v) for v in counts)
    return {"arena": arena[:na].copy(), "obstacle": obstacle[:no].copy(), "robot": robot[:nr].copy(), "end_points": ends[:ne].copy(),
            "quats": quats[:nq].copy(), "lim": lim[: 6 if dim == "3D" else 4].copy(), "Lx": lx, "Ly": ly, "numSlice": ns}


def store_result(details_dir, suffix, vertex, edge, path):
    """hrm::storeGraphInfo + hrm::storePathInfo (vertex_/edge_/path_id_/solution_path_/interpolated_path_<suffix>.csv)."""
    vertex = _d(vertex)
    edge = np.ascontiguousarray(edge, dtype=np.int64).reshape(-1, 2)
    path = np.ascontiguousarray(path, dtype=np.int64)
    rc = load().hrmhost_store_result(details_dir.encode(), suffix.encode(), C.c_int(vertex.shape[1]), C.c_long(vertex.shape[0]), _p(vertex),
                                     C.c_long(edge.shape[0]), _p(edge), C.c_long(path.shape[0]), _p(path))
    if rc != 0:
        raise _err()


# ---- articulated robots (hrm_b200/host/hrm_articulated.hpp) -------------------------------------------------------------
def fk(urdf, joints, body):
    """hrm::ParseURDF::getTransform(jointConfig, "body<k>") -> (4x4, number of joints)."""
    joints = _d(joints)
    out = np.zeros(16)
    nj = load().hrmhost_fk(str(urdf).encode(), _p(joints), C.c_int(joints.shape[0]), C.c_int(body), _p(out))
    if nj < 0:
        raise _err()
    return out.reshape(4, 4), nj


def plan_prob3d(n_arena, n_obs, sq12, n, robot_rows, urdf, samples, lim, Lx, Ly, num_point, start, goal, per_orientation=False, precision=0,
                device=0, batch=8, seed=20220726, time_limit=0.0, nj=None, cap_v=1 << 18, cap_e=1 << 21):
    """hrm::planners::ProbHRM3D(robot, urdf, arena, obstacle, request).plan(); `samples` [N][4 + nj] replace the random draws
    (None: seeded RNG, then `nj` is required)."""
    sq12, robot_rows, lim = _d(sq12), _d(robot_rows), _d(lim)
    if samples is not None:
        samples = _d(samples)
        nj = samples.shape[1] - 4
        ns = samples.shape[0]
    else:
        samples, ns = np.zeros((1, 4 + nj)), -1
    start, goal = _d(start)[: 7 + nj].copy(), _d(goal)[: 7 + nj].copy()
    info = np.zeros(8, dtype=np.int64)
    vertex, edge, path = np.zeros((cap_v, 7 + nj)), np.zeros((cap_e, 2), dtype=np.int64), np.zeros(cap_v, dtype=np.int64)
    rc = load().hrmhost_plan_prob3d(C.c_int(n_arena), C.c_int(n_obs), _p(sq12), C.c_int(n), C.c_int(robot_rows.shape[0]), _p(robot_rows),
                                    str(urdf).encode(), C.c_int(nj), C.c_int(ns), _p(samples), C.c_ulonglong(seed), C.c_int(batch), _p(lim),
                                    C.c_int(Lx), C.c_int(Ly), C.c_int(num_point), _p(start), _p(goal), C.c_int(int(per_orientation)),
                                    C.c_int(precision), C.c_int(device), C.c_double(time_limit), _p(info), _p(vertex), C.c_long(cap_v),
                                    _p(edge), C.c_long(cap_e), _p(path), C.c_long(cap_v))
    if rc != 0:
        raise _err()
    out = _out(info, vertex, edge, path)
    out["slices"] = int(info[4])
    return out
