"""ctypes loader for the CPU oracle (oracle/libhrm_oracle.so).

ORACLE — TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py; never from hrm_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build(force=False):
    so = os.path.join(_HERE, "libhrm_oracle.so")
    fast = os.path.join(_HERE, "libhrm_oracle_fast.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".hpp", ".cpp")) and not f.startswith("ref_")]
    if force or not os.path.exists(so) or not os.path.exists(fast) or any(os.path.getmtime(s) > min(os.path.getmtime(so), os.path.getmtime(fast)) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


_FAST = None


def fast_lib():
    """The hoisted build (libhrm_oracle_fast.so, -DHRMO_FAST): same results, loop-invariant work done once."""
    global _FAST
    if _FAST is None:
        build()
        _FAST = C.CDLL(os.path.join(_HERE, "libhrm_oracle_fast.so"))
        _FAST.hrmo_time_layers3d.restype = C.c_double
    return _FAST


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.hrmo_exponential_function.restype = C.c_double
        _LIB.hrmo_exponential_function.argtypes = [C.c_double, C.c_double, C.c_int]
        _LIB.hrmo_angular_distance.restype = C.c_double
        _LIB.hrmo_time_layers3d.restype = C.c_double
    return _LIB


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def constants():
    out = np.zeros(3)
    lib().hrmo_constants(_p(out))
    return {"PI": out[0], "HALF_PI": out[1], "EPSILON": out[2]}


def quat_to_rot(q):
    q = _d(q)
    R = np.zeros(9)
    lib().hrmo_quat_to_rot(_p(q), _p(R))
    return R.reshape(3, 3)


def rot_to_quat(R):
    R = _d(R).reshape(9)
    q = np.zeros(4)
    lib().hrmo_rot_to_quat(_p(R), _p(q))
    return q


def angle_axis_to_quat(angle, axis):
    axis = _d(axis)
    q = np.zeros(4)
    lib().hrmo_angle_axis_to_quat(C.c_double(angle), _p(axis), _p(q))
    return q


def slerp(a, t, b):
    a, b = _d(a), _d(b)
    out = np.zeros(4)
    lib().hrmo_slerp(_p(a), C.c_double(t), _p(b), _p(out))
    return out


def angular_distance(a, b):
    a, b = _d(a), _d(b)
    return lib().hrmo_angular_distance(_p(a), _p(b))


def linspaced(n, lo, hi):
    out = np.zeros(n)
    lib().hrmo_linspaced(C.c_int(n), C.c_double(lo), C.c_double(hi), _p(out))
    return out


def exponential_function(theta, power, is_sine):
    return lib().hrmo_exponential_function(theta, power, int(bool(is_sine)))


def sq_tables(sq, n):
    sq = _d(sq)
    out = np.zeros(8 * n)
    lib().hrmo_sq_tables(_p(sq), C.c_int(n), _p(out))
    return out.reshape(8, n)


def origin_shape3d(sq, n):
    sq = _d(sq)
    out = np.zeros(3 * n * n)
    lib().hrmo_origin_shape3d(_p(sq), C.c_int(n), _p(out))
    return out.reshape(n * n, 3).T.copy()  # 3 x M like Eigen


def minksum3d(sq, semiB, quatB, n, K):
    sq, semiB, quatB = _d(sq), _d(semiB), _d(quatB)
    out = np.zeros(3 * n * n)
    lib().hrmo_minksum3d(_p(sq), _p(semiB), _p(quatB), C.c_int(n), C.c_int(K), _p(out))
    return out.reshape(n * n, 3).T.copy()


def origin_shape2d(se, num):
    se = _d(se)
    out = np.zeros(2 * num)
    lib().hrmo_origin_shape2d(_p(se), C.c_int(num), _p(out))
    return out.reshape(num, 2).T.copy()


def minksum2d(se, semiB, angleB, num, K):
    se, semiB = _d(se), _d(semiB)
    out = np.zeros(2 * num)
    lib().hrmo_minksum2d(_p(se), _p(semiB), C.c_double(angleB), C.c_int(num), C.c_int(K), _p(out))
    return out.reshape(num, 2).T.copy()


def mvce2d(a, b, thA, thB):
    a, b = _d(a), _d(b)
    out = np.zeros(3)
    lib().hrmo_mvce2d(_p(a), _p(b), C.c_double(thA), C.c_double(thB), _p(out))
    return out


def mvce3d(a, b, qa, qb):
    a, b, qa, qb = _d(a), _d(b), _d(qa), _d(qb)
    out = np.zeros(7)
    lib().hrmo_mvce3d(_p(a), _p(b), _p(qa), _p(qb), _p(out))
    return out


def tfe2d(a, thA, thB, num_step):
    a = _d(a)
    out = np.zeros(3)
    lib().hrmo_tfe2d(_p(a), C.c_double(thA), C.c_double(thB), C.c_int(num_step), _p(out))
    return out


def tfe3d(a, qa, qb, num_step):
    a, qa, qb = _d(a), _d(qa), _d(qb)
    out = np.zeros(7)
    lib().hrmo_tfe3d(_p(a), _p(qa), _p(qb), C.c_int(num_step), _p(out))
    return out


def vertical_line_mesh(bound3xM, n, x, y, z0=0.0):
    b = _d(np.asarray(bound3xM).T)  # -> M x 3 == col-major 3 x M
    z = np.zeros(2)
    c = lib().hrmo_vertical_line_mesh(_p(b), C.c_int(n), C.c_double(x), C.c_double(y), C.c_double(z0), _p(z))
    return c, z[:c].copy()


def line_mesh(bound3xM, n, line6):
    b = _d(np.asarray(bound3xM).T)
    line6 = _d(line6)
    h = np.zeros(6)
    c = lib().hrmo_line_mesh(_p(b), C.c_int(n), _p(line6), _p(h))
    return c, h.reshape(2, 3)[:c].copy()


def horizontal_line_polygon(bound2xN, ty, cap=64):
    b = _d(np.asarray(bound2xN).T)
    out = np.zeros(cap)
    c = lib().hrmo_horizontal_line_polygon(_p(b), C.c_int(b.shape[0]), C.c_double(ty), _p(out), C.c_int(cap))
    return c, out[: min(c, cap)].copy()


def seg_polygon(bound2xN, s1, s2):
    b = _d(np.asarray(bound2xN).T)
    s1, s2 = _d(s1), _d(s2)
    return bool(lib().hrmo_seg_polygon(_p(b), C.c_int(b.shape[0]), _p(s1), _p(s2)))


def free_segments(alo, aup, olo, oup, cap=1 << 16):
    """Interval tables [Ly][Sa] / [Ly][So] -> (off[Ly+1], xL, xU, xM) incl. enhanceFreeSegment."""
    alo, aup, olo, oup = _d(alo), _d(aup), _d(olo), _d(oup)
    Ly, Sa = alo.shape
    So = olo.shape[1]
    off = np.zeros(Ly + 1, dtype=np.int32)
    xL, xU, xM = np.zeros(cap), np.zeros(cap), np.zeros(cap)
    tot = lib().hrmo_free_segments(C.c_int(Ly), C.c_int(Sa), C.c_int(So), _p(alo), _p(aup), _p(olo), _p(oup), _p(off),
                                   _p(xL), _p(xU), _p(xM), C.c_int(cap))
    assert tot <= cap
    return off, xL[:tot].copy(), xU[:tot].copy(), xM[:tot].copy()


def interval_op(op, a, b=None, impl="oracle"):
    """Interval::unions ("unions"), intersects, complements on lists of (start, end) pairs.  impl = "oracle" runs the
    restatement, impl = "reference" the reference's own object code (oracle/_ref/libhrm_ref.so, see oracle/ref_build.mk)."""
    code = {"unions": 0, "intersects": 1, "complements": 2}[op]
    a = _d(a).reshape(-1, 2)
    b = _d(b if b is not None else np.zeros((0, 2))).reshape(-1, 2)
    cap = 2 * (len(a) + len(b)) + 4
    out = np.zeros((cap, 2))
    fn = lib().hrmo_interval_op if impl == "oracle" else ref_lib().hrmref_interval_op
    n = fn(C.c_int(code), _p(a), C.c_int(len(a)), _p(b), C.c_int(len(b)), _p(out), C.c_int(cap))
    assert 0 <= n <= cap
    return out[:n].copy()


_REF = None


def ref_path():
    return os.path.join(_HERE, "_ref", "libhrm_ref.so")


def build_ref(reference="/root/reference"):
    """Compiles the reference's own Interval.cpp / Parse2dCsvFile.cpp into oracle/_ref (only where the reference tree exists:
    in the build container; the GPU box receives the prebuilt library with the repo snapshot)."""
    if os.path.isdir(reference):
        subprocess.check_call(["make", "-s", "-f", os.path.join("oracle", "ref_build.mk"), "REF=" + reference], cwd=os.path.dirname(_HERE))
    return ref_path() if os.path.exists(ref_path()) else None


def ref_lib():
    global _REF
    if _REF is None:
        if not os.path.exists(ref_path()):
            raise RuntimeError("oracle/_ref/libhrm_ref.so not built (make -f oracle/ref_build.mk needs /root/reference)")
        _REF = C.CDLL(ref_path())
        _REF.hrmref_parse_csv.restype = C.c_long
    return _REF


def ref_parse_csv(filename, stride=64, cap_rows=1 << 16):
    """hrm::parse2DCsvFile of the reference's own object code: list of rows."""
    out = np.zeros((cap_rows, stride))
    row_len = np.zeros(cap_rows, dtype=np.int32)
    n = ref_lib().hrmref_parse_csv(filename.encode(), _p(out), C.c_int(stride), _p(row_len), C.c_long(cap_rows))
    assert n <= cap_rows
    return [out[r, : row_len[r]].copy() for r in range(n)]


def robot3d_body_poses(robot_rows, q):
    robot_rows = _d(robot_rows)
    B = robot_rows.shape[0]
    q = _d(q)
    bq, bo = np.zeros((B, 4)), np.zeros((B, 3))
    lib().hrmo_robot3d_body_poses(C.c_int(B), _p(robot_rows), _p(q), _p(bq), _p(bo))
    return bq, bo


def robot2d_body_poses(robot_rows, theta):
    robot_rows = _d(robot_rows)
    B = robot_rows.shape[0]
    ba, bo = np.zeros(B), np.zeros((B, 2))
    lib().hrmo_robot2d_body_poses(C.c_int(B), _p(robot_rows), C.c_double(theta), _p(ba), _p(bo))
    return ba, bo


def layer3d(n_arena, n_obs, sq, n, body_semi, body_quat, body_off, lim, Lx, Ly, want_bound=True, seg_cap=1 << 18, fast=False):
    """One 3D C-layer. Returns dict(bound[S,3,M], lo[L,S], up[L,S], off[L+1], xL, xU, xM, single_hits)."""
    sq, body_semi, body_quat, body_off, lim = _d(sq), _d(body_semi), _d(body_quat), _d(body_off), _d(lim)
    B = body_semi.shape[0]
    S = (n_arena + n_obs) * B
    M = n * n
    L = Lx * Ly
    bound = np.zeros((S, M, 3)) if want_bound else None
    lo, up = np.zeros((L, S)), np.zeros((L, S))
    off = np.zeros(L + 1, dtype=np.int32)
    xL, xU, xM = np.zeros(seg_cap), np.zeros(seg_cap), np.zeros(seg_cap)
    sh = C.c_long(0)
    tot = (fast_lib() if fast else lib()).hrmo_layer3d(C.c_int(n_arena), C.c_int(n_obs), _p(sq), C.c_int(n), C.c_int(B), _p(body_semi), _p(body_quat),
                             _p(body_off), _p(lim), C.c_int(Lx), C.c_int(Ly), _p(bound), _p(lo), _p(up), _p(off), _p(xL), _p(xU),
                             _p(xM), C.c_int(seg_cap), C.byref(sh))
    assert 0 <= tot <= seg_cap, tot
    return {
        "bound": None if bound is None else np.ascontiguousarray(bound.transpose(0, 2, 1)),
        "lo": lo, "up": up, "off": off, "xL": xL[:tot].copy(), "xU": xU[:tot].copy(), "xM": xM[:tot].copy(),
        "single_hits": sh.value,
    }


def layer3d_parallel(n_arena, n_obs, sq, n, body_semi, body_quat, body_off, lim, Lx, Ly, threads=None, want_bound=False):
    """layer3d() for layers with ~10^3 surfaces: the surfaces of a layer are independent (FreeSpace-inl.h:42-60,
    FreeSpace3D.cpp:29-68), so chunks of objects go through hrmo_layer3d on separate threads (ctypes releases the GIL) and
    the free segments are then formed from the merged interval tables, plane by plane, with hrmo_free_segments -- the same
    code path layer3d() runs on the whole tables.  Same return value as layer3d()."""
    import os
    from concurrent.futures import ThreadPoolExecutor

    sq = _d(sq).reshape(n_arena + n_obs, 12)
    body_semi = _d(body_semi)
    B = body_semi.shape[0]
    threads = threads or len(os.sched_getaffinity(0))
    S0 = n_arena + n_obs
    cuts = np.linspace(0, S0, min(4 * threads, S0) + 1).astype(int)
    chunks = [(a, b) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]

    def run(ab):
        a, b = ab
        na = max(0, min(b, n_arena) - a)  # arena objects in this chunk (they come first)
        return layer3d(na, (b - a) - na, sq[a:b], n, body_semi, body_quat, body_off, lim, Lx, Ly, want_bound=want_bound)

    with ThreadPoolExecutor(max_workers=threads) as ex:
        parts = list(ex.map(run, chunks))
    lo = np.concatenate([p_["lo"] for p_ in parts], axis=1)
    up = np.concatenate([p_["up"] for p_ in parts], axis=1)
    Sa = n_arena * B
    off = [0]
    xs = ([], [], [])
    for ix in range(Lx):
        r = slice(ix * Ly, (ix + 1) * Ly)
        o, xL, xU, xM = free_segments(lo[r, :Sa], up[r, :Sa], lo[r, Sa:], up[r, Sa:])
        off.extend((o[1:] + off[ix * Ly]).tolist())
        for dst, src in zip(xs, (xL, xU, xM)):
            dst.append(src)
    return {
        "bound": np.concatenate([p_["bound"] for p_ in parts]) if want_bound else None,
        "lo": lo, "up": up, "off": np.array(off, dtype=np.int32), "xL": np.concatenate(xs[0]), "xU": np.concatenate(xs[1]),
        "xM": np.concatenate(xs[2]), "single_hits": sum(p_["single_hits"] for p_ in parts),
    }


def layer2d(n_arena, n_obs, se, num, body_semi, body_angle, body_off, lim, Ly, seg_cap=1 << 16):
    se, body_semi, body_angle, body_off, lim = _d(se), _d(body_semi), _d(body_angle), _d(body_off), _d(lim)
    B = body_semi.shape[0]
    S = (n_arena + n_obs) * B
    bound = np.zeros((S, num, 2))
    lo, up = np.zeros((Ly, S)), np.zeros((Ly, S))
    off = np.zeros(Ly + 1, dtype=np.int32)
    xL, xU, xM = np.zeros(seg_cap), np.zeros(seg_cap), np.zeros(seg_cap)
    sh = C.c_long(0)
    tot = lib().hrmo_layer2d(C.c_int(n_arena), C.c_int(n_obs), _p(se), C.c_int(num), C.c_int(B), _p(body_semi), _p(body_angle),
                             _p(body_off), _p(lim), C.c_int(Ly), _p(bound), _p(lo), _p(up), _p(off), _p(xL), _p(xU), _p(xM),
                             C.c_int(seg_cap), C.byref(sh))
    assert 0 <= tot <= seg_cap, tot
    return {
        "bound": np.ascontiguousarray(bound.transpose(0, 2, 1)), "lo": lo, "up": up, "off": off,
        "xL": xL[:tot].copy(), "xU": xU[:tot].copy(), "xM": xM[:tot].copy(), "single_hits": sh.value,
    }


def time_layers3d(n_arena, n_obs, sq, n, body_semi, body_quat, body_off, lim, Lx, Ly, threads=1, fast=False):
    """CPU baseline: seconds to build K layers (body_quat [K,B,4], body_off [K,B,3]); fast = the hoisted build."""
    sq, body_semi, body_quat, body_off, lim = _d(sq), _d(body_semi), _d(body_quat), _d(body_off), _d(lim)
    K, B = body_quat.shape[0], body_quat.shape[1]
    cs = C.c_long(0)
    sec = (fast_lib() if fast else lib()).hrmo_time_layers3d(C.c_int(n_arena), C.c_int(n_obs), _p(sq), C.c_int(n), C.c_int(B), _p(body_semi), _p(body_quat),
                                   _p(body_off), _p(lim), C.c_int(Lx), C.c_int(Ly), C.c_int(K), C.c_int(threads), C.byref(cs))
    return sec, cs.value


def _bounds_flat(bounds):
    """list/array of dim x M boundaries -> contiguous [count][M][dim] (Eigen column-major per boundary)."""
    arr = np.stack([np.asarray(b).T for b in bounds])
    return np.ascontiguousarray(arr, dtype=np.float64)


def edges3d(bounds, n, v1, v2):
    """isSameSliceTransitionFree (HRM3D.cpp:319-365) of E segments against C-obstacle boundaries (each 3 x n^2)."""
    b = _bounds_flat(bounds)
    v1, v2 = _d(v1).reshape(-1, 3), _d(v2).reshape(-1, 3)
    out = np.zeros(v1.shape[0], dtype=np.uint8)
    lib().hrmo_edges3d.restype = C.c_long
    single = lib().hrmo_edges3d(_p(b), C.c_int(n), C.c_int(b.shape[0]), C.c_int(v1.shape[0]), _p(v1), _p(v2), _p(out))
    return out.astype(bool), single


def edges2d(bounds, v1, v2):
    b = _bounds_flat(bounds)
    v1, v2 = _d(v1).reshape(-1, 2), _d(v2).reshape(-1, 2)
    out = np.zeros(v1.shape[0], dtype=np.uint8)
    lib().hrmo_edges2d(_p(b), C.c_int(b.shape[1]), C.c_int(b.shape[0]), C.c_int(v1.shape[0]), _p(v1), _p(v2), _p(out))
    return out.astype(bool)


def pt_in_cfree3d(bounds, n, pts):
    b = _bounds_flat(bounds)
    pts = _d(pts).reshape(-1, 3)
    out = np.zeros(pts.shape[0], dtype=np.uint8)
    lib().hrmo_pt_in_cfree3d.restype = C.c_long
    single = lib().hrmo_pt_in_cfree3d(_p(b), C.c_int(n), C.c_int(b.shape[0]), C.c_int(pts.shape[0]), _p(pts), _p(out))
    return out.astype(bool), single


def pt_in_cfree2d(bounds, pts):
    b = _bounds_flat(bounds)
    pts = _d(pts).reshape(-1, 2)
    out = np.zeros(pts.shape[0], dtype=np.uint8)
    lib().hrmo_pt_in_cfree2d.restype = C.c_long
    single = lib().hrmo_pt_in_cfree2d(_p(b), C.c_int(b.shape[1]), C.c_int(b.shape[0]), C.c_int(pts.shape[0]), _p(pts), _p(out))
    return out.astype(bool), single


def _plan_out(info, vertex, edge, path):
    nv, ne, npth = int(info[1]), int(info[2]), int(info[3])
    return {"solved": bool(info[0]), "vertex": vertex[:nv].copy(), "edge": edge[:ne].copy(), "path": path[:npth].copy(),
            "single_hits": int(info[4]), "edge_tests": int(info[5]), "refine_rounds": int(info[7])}


def set_max_refine(n):
    """Refinement rounds plan3d / plan2d may run when the first roadmap does not connect start and goal (default 0)."""
    lib().hrmo_set_max_refine(C.c_int(n))


def set_prob_refine_period(n):
    """ProbHRM3D doubles the sweep lines of all slices every `n` slices (60 in the reference, ProbHRM3D.cpp:52-56; 0 = never)."""
    lib().hrmo_set_prob_refine_period(C.c_int(n))


def plan3d(n_arena, n_obs, sq, n, robot_rows, quats, lim, Lx, Ly, num_point, start, goal, per_orientation=False,
           cap_v=1 << 18, cap_e=1 << 21):
    sq, robot_rows, quats, lim, start, goal = _d(sq), _d(robot_rows), _d(quats), _d(lim), _d(start)[:7].copy(), _d(goal)[:7].copy()
    info = np.zeros(8, dtype=np.int64)
    vertex = np.zeros((cap_v, 7))
    edge = np.zeros((cap_e, 2), dtype=np.int64)
    path = np.zeros(cap_v, dtype=np.int64)
    lib().hrmo_plan3d(C.c_int(n_arena), C.c_int(n_obs), _p(sq), C.c_int(n), C.c_int(robot_rows.shape[0]), _p(robot_rows),
                      C.c_int(quats.shape[0]), _p(quats), _p(lim), C.c_int(Lx), C.c_int(Ly), C.c_int(num_point), _p(start), _p(goal),
                      C.c_int(int(per_orientation)), _p(info), _p(vertex), C.c_long(cap_v), _p(edge), C.c_long(cap_e), _p(path),
                      C.c_long(cap_v))
    assert info[1] <= cap_v and info[2] <= cap_e
    return _plan_out(info, vertex, edge, path)


def plan2d(n_arena, n_obs, se, num, robot_rows, num_slice, lim, Ly, num_point, start, goal, per_orientation=False,
           cap_v=1 << 16, cap_e=1 << 20):
    se, robot_rows, lim, start, goal = _d(se), _d(robot_rows), _d(lim), _d(start)[:3].copy(), _d(goal)[:3].copy()
    info = np.zeros(8, dtype=np.int64)
    vertex = np.zeros((cap_v, 3))
    edge = np.zeros((cap_e, 2), dtype=np.int64)
    path = np.zeros(cap_v, dtype=np.int64)
    lib().hrmo_plan2d(C.c_int(n_arena), C.c_int(n_obs), _p(se), C.c_int(num), C.c_int(robot_rows.shape[0]), _p(robot_rows),
                      C.c_int(num_slice), _p(lim), C.c_int(Ly), C.c_int(num_point), _p(start), _p(goal), C.c_int(int(per_orientation)),
                      _p(info), _p(vertex), C.c_long(cap_v), _p(edge), C.c_long(cap_e), _p(path), C.c_long(cap_v))
    assert info[1] <= cap_v and info[2] <= cap_e
    return _plan_out(info, vertex, edge, path)


# ---- articulated path (hrmo_articulated.hpp) ----------------------------------------------------------------------------
def fk(urdf, joints, body):
    """ParseURDF::getTransform(jointConfig, "body<k>") (body 0 = root link) -> (4x4 matrix, number of joints)."""
    joints = _d(joints)
    out = np.zeros(16)
    nj = lib().hrmo_fk(str(urdf).encode(), _p(joints), C.c_int(joints.shape[0]), C.c_int(body), _p(out))
    if nj < 0:
        raise RuntimeError("hrmo_fk failed for %s body %d" % (urdf, body))
    return out.reshape(4, 4), nj


def robot_tf_urdf(n, robot_rows, urdf, base7, joints):
    """MultiBodyTree3D::robotTF(urdf, gBase, jointConfig) -> body poses [B][7] (x,y,z,qw,qx,qy,qz), base first."""
    robot_rows, base7, joints = _d(robot_rows), _d(base7), _d(joints)
    out = np.zeros((robot_rows.shape[0], 7))
    rc = lib().hrmo_robot_tf_urdf(C.c_int(n), C.c_int(robot_rows.shape[0]), _p(robot_rows), str(urdf).encode(), _p(base7), _p(joints),
                                  C.c_int(joints.shape[0]), _p(out))
    if rc != 0:
        raise RuntimeError("hrmo_robot_tf_urdf failed")
    return out


def plan_prob3d(n_arena, n_obs, sq, n, robot_rows, urdf, samples, lim, Lx, Ly, num_point, start, goal, per_orientation=False,
                cap_v=1 << 18, cap_e=1 << 21):
    """ProbHRM3D::plan with the random draws injected: samples [N][4 + nj] for the slices after start and goal."""
    sq, robot_rows, samples, lim = _d(sq), _d(robot_rows), _d(samples), _d(lim)
    nj = samples.shape[1] - 4
    start, goal = _d(start)[: 7 + nj].copy(), _d(goal)[: 7 + nj].copy()
    info = np.zeros(8, dtype=np.int64)
    vertex = np.zeros((cap_v, 7 + nj))
    edge = np.zeros((cap_e, 2), dtype=np.int64)
    path = np.zeros(cap_v, dtype=np.int64)
    rc = lib().hrmo_plan_prob3d(C.c_int(n_arena), C.c_int(n_obs), _p(sq), C.c_int(n), C.c_int(robot_rows.shape[0]), _p(robot_rows),
                                str(urdf).encode(), C.c_int(nj), C.c_int(samples.shape[0]), _p(samples), _p(lim), C.c_int(Lx), C.c_int(Ly),
                                C.c_int(num_point), _p(start), _p(goal), C.c_int(int(per_orientation)), _p(info), _p(vertex),
                                C.c_long(cap_v), _p(edge), C.c_long(cap_e), _p(path), C.c_long(cap_v))
    if rc != 0:
        raise RuntimeError("hrmo_plan_prob3d failed (%d)" % rc)
    assert info[1] <= cap_v and info[2] <= cap_e
    out = _plan_out(info, vertex, edge, path)
    out["slices"] = int(info[6])
    return out
