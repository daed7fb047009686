"""CPU-side tests: the C-ABI library loads and exports every declared symbol, fails loudly without a GPU, the host-side
mirror classes produce the arrays the reference would hold, scene fixtures are what defineParameters gives."""
import os
import re

import numpy as np
import pytest

from hrm_b200 import capi
from hrm_b200 import hostmath as hm
from hrm_b200 import workload as wl
from tests import util


def test_library_exports_every_declared_symbol():
    lib = capi.load()
    names = capi.declared_symbols()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), n
    assert lib.hrmgpu_abi_version() == 2


def test_ctypes_signatures_match_the_header():
    """Every prototype in include/hrmgpu.h has a ctypes signature in capi.load() with the same number of parameters (a drifted
    binding would pass garbage through the C ABI without any error)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    txt = open(os.path.join(root, "include", "hrmgpu.h")).read()
    txt = re.sub(r"/\*.*?\*/", " ", txt, flags=re.S)  # comments out
    protos = dict(re.findall(r"\b(hrmgpu_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", txt, flags=re.S))
    lib = capi.load()
    assert set(protos) == set(capi.declared_symbols())
    checked = 0
    for name, args in protos.items():
        fn = getattr(lib, name)
        if fn.argtypes is None:
            continue  # (debug hooks bound ad hoc by the tools)
        n_params = 0 if args.strip() in ("", "void") else len(args.split(","))
        assert len(fn.argtypes) == n_params, (name, n_params, len(fn.argtypes))
        checked += 1
    assert checked >= 40


def test_cpp_host_library_loads_and_exports_planners():
    from hrm_b200 import hostlib

    lib = hostlib.load()
    for n in ("hrmhost_plan3d", "hrmhost_plan2d", "hrmhost_last_error"):
        assert hasattr(lib, n), n


def test_cpp_host_fails_loudly_without_device():
    import torch

    from hrm_b200 import hostlib
    from tests import util as u

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    sc = u.scenes()["3D_superquadrics_sparse_rabbit"]
    rows = np.array(sc["arena"] + sc["obstacle"])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        hostlib.plan3d(1, len(sc["obstacle"]), rows, 10, np.array(sc["robot"]), np.array(sc["quats"][:2]), sc["lim"], sc["Lx"], sc["Ly"], 5,
                       sc["end_points"][0], sc["end_points"][1])


def test_no_cpu_fallback_without_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert capi.load().hrmgpu_device_count() == 0
    with pytest.raises(capi.HrmGpuError) as e:
        capi.Context(0)
    assert e.value.code == capi.E_NODEVICE


def test_product_never_imports_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dirpath, _, files in os.walk(os.path.join(root, "hrm_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                for pat in (r"^\s*(from|import)\s+oracle", r"#include\s*[\"<][^\n]*(oracle|hrmo_)", r"hrm_oracle", r"\bhrmo\b", r"libhrm_oracle"):
                    assert not re.search(pat, txt, flags=re.M), (pat, os.path.join(dirpath, f))


def test_hostmath_matches_oracle_bitwise(oracle):
    rng = np.random.default_rng(0)
    for _ in range(50):
        q = util.f32round(rng.normal(size=4))
        assert np.array_equal(np.array(hm.quat_to_rot(q)), oracle.quat_to_rot(q))
        R = oracle.quat_to_rot(q / np.linalg.norm(q))
        assert np.array_equal(np.array(hm.rot_to_quat(R.tolist())), oracle.rot_to_quat(R))
    assert np.array_equal(np.array(hm.angle_axis_to_quat(0.7, [0.2, 0.4, -1.3])), oracle.angle_axis_to_quat(0.7, [0.2, 0.4, -1.3]))


def test_multibody_poses_match_oracle(oracle):
    sc = util.scenes()["3D_superquadrics_cluttered_rabbit"]
    for q in sc["quats"][:10]:
        semi, R, off, quat = util.body_poses3d(sc["robot"], [q])
        bq, bo = oracle.robot3d_body_poses(np.array(sc["robot"]), q)
        assert np.array_equal(quat[0], bq) and np.array_equal(off[0], bo)
    sc = util.scenes()["2D_superellipse_sparse_rabbit"]
    for th in (-3.1415926, -1.0, 0.0, 2.2):
        semi, ang, off = util.body_poses2d(sc["robot"], [th])
        ba, bo = oracle.robot2d_body_poses(np.array(sc["robot"]), th)
        assert np.array_equal(ang[0], ba) and np.array_equal(off[0], bo)


def test_scene_fixtures_follow_define_parameters():
    S = util.scenes()
    c = S["3D_superquadrics_cluttered_rabbit"]
    assert c["lim"] == [-58.0, 58.0, -28.0, 28.0, -18.0, 18.0] and (c["Lx"], c["Ly"]) == (11, 5) and len(c["quats"]) == 60
    assert len(c["arena"]) == 1 and len(c["obstacle"]) == 7 and len(c["robot"]) == 3
    s = S["2D_superellipse_sparse_rabbit"]
    assert s["lim"] == [-62.5, 62.5, -32.5, 32.5] and s["Ly"] == 3
    # values went through stof: they are exactly representable as float32
    a = np.array(c["obstacle"])
    assert np.array_equal(a, a.astype(np.float32).astype(np.float64))


def test_workload_is_seeded_and_float_rounded():
    r1, r2 = wl.scene_rows12(50), wl.scene_rows12(50)
    assert np.array_equal(r1, r2) and r1.shape == (51, 12)
    assert np.array_equal(r1, r1.astype(np.float32).astype(np.float64))
    assert wl.algorithmic_bytes_per_layer(1001, 10000, 512, 8) == 3 * 1001 * 10000 * 8 + 2 * 512 * 1001 * 8
    r17 = wl.scene_rows17(r1)
    assert r17.shape == (51, 17) and np.array_equal(r17[0, 8:], np.eye(3).ravel())
