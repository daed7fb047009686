"""Shared helpers of the parity tests: bundled scenes (tests/golden/scenes.json), seeded synthetic scenes,
and conversion into the flat arrays both the oracle and the C ABI take."""
import json
import os

import numpy as np

from hrm_b200 import hostmath as hm
from hrm_b200.model import MultiBodyTree2D, MultiBodyTree3D, SuperEllipse, SuperQuadrics

HERE = os.path.dirname(os.path.abspath(__file__))
_SCENES = None


def scenes():
    global _SCENES
    if _SCENES is None:
        with open(os.path.join(HERE, "golden", "scenes.json")) as f:
            _SCENES = json.load(f)
    return _SCENES


def sq_objects(rows12, n):
    return [SuperQuadrics(r[0:3], r[3:5], r[5:8], r[8:12], n) for r in rows12]


def se_objects(rows6, num):
    return [SuperEllipse(r[0:2], r[2], r[3:5], r[5], num) for r in rows6]


def robot3d(rows12, n=0):
    parts = sq_objects(rows12, n)
    rb = MultiBodyTree3D(parts[0])
    for p in parts[1:]:
        rb.addBody(p)
    return rb


def robot2d(rows6, num=0):
    parts = se_objects(rows6, num)
    rb = MultiBodyTree2D(parts[0])
    for p in parts[1:]:
        rb.addBody(p)
    return rb


def sq17(rows12, n):
    return np.array([o.row17() for o in sq_objects(rows12, n)], dtype=np.float64)


def body_poses3d(robot_rows, quats):
    """Per-orientation body poses: base rotation R(q_k) at the origin (HRM3D.cpp:27-31,510-518).
    Returns semi [B,3], R [K,B,9], off [K,B,3], quat [K,B,4] (the latter for the oracle)."""
    rb = robot3d(robot_rows)
    Rs, offs, qs = [], [], []
    semi = None
    for q in quats:
        rb.robotTF(hm.quat_to_rot(q), [0.0, 0.0, 0.0])
        semi, R, off = rb.body_arrays()
        Rs.append(R)
        offs.append(off)
        qs.append([rb.base.quat] + [l.quat for l in rb.link])
    return np.array(semi), np.array(Rs), np.array(offs), np.array(qs)


def frozen_pose3d(robot_rows):
    """The as-loaded robot pose the reference's FreeSpace copy keeps for every layer (SURVEY.md finding 1)."""
    rb = robot3d(robot_rows)
    semi, R, off = rb.body_arrays()
    q = [rb.base.quat] + [l.quat for l in rb.link]
    return np.array(semi), np.array([R]), np.array([off]), np.array([q])


def body_poses2d(robot_rows, thetas):
    rb = robot2d(robot_rows)
    angs, offs = [], []
    semi = None
    for th in thetas:
        rb.robotTF(th)
        semi, a, o = rb.body_arrays()
        angs.append(a)
        offs.append(o)
    return np.array(semi), np.array(angs), np.array(offs)


def f32round(a):
    return np.asarray(a, dtype=np.float32).astype(np.float64)


def random_unit_quats(rng, k):
    q = rng.normal(size=(k, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    return f32round(q)  # (w,x,y,z), float-rounded like the CSV path => not exactly unit norm


def synthetic_scene3d(seed, n_obs, arena=(70.0, 40.0, 30.0), box=(60.0, 32.0, 22.0), semi_rng=(2.0, 10.0)):
    """Seeded synthetic scene of SURVEY.md §8(d): one arena (eps 0.1, identity pose) + n_obs random
    superquadric obstacles; every value float-rounded like the CSV path.  Rows are 12-column
    [a,b,c,e1,e2,px,py,pz,qw,qx,qy,qz]."""
    rng = np.random.default_rng(seed)
    rows = [[arena[0], arena[1], arena[2], 0.1, 0.1, 0, 0, 0, 1, 0, 0, 0]]
    semi = rng.uniform(semi_rng[0], semi_rng[1], size=(n_obs, 3))
    eps = rng.uniform(0.1, 1.8, size=(n_obs, 2))
    pos = rng.uniform(-1, 1, size=(n_obs, 3)) * np.array(box)
    q = random_unit_quats(rng, n_obs)
    for i in range(n_obs):
        rows.append(list(semi[i]) + list(eps[i]) + list(pos[i]) + list(q[i]))
    return f32round(np.array(rows, dtype=np.float64))


def csr_lists(off, *arrays):
    return [[a[off[i]:off[i + 1]] for i in range(len(off) - 1)] for a in arrays]
