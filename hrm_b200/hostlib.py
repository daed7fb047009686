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
            "edge_tests": int(info[5]), "gpu_launches": int(info[6]), "total_time_s": info[7] * 1e-6}


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
