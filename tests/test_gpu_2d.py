"""GPU parity of the 2D C-layer construction (K1' + K2' + K3) against the CPU oracle through the C ABI."""
import numpy as np
import pytest

from hrm_b200 import capi
from hrm_b200 import hostmath as hm
from tests import util

pytestmark = pytest.mark.gpu


def headings(num_slice):
    dr = 2 * hm.PI / (float(num_slice) - 1)  # reference src/planners/HRM2D.cpp:46-51
    return [-hm.PI + dr * float(i) for i in range(num_slice)]


def run_both(oracle, ctx, sc, num, Ly, thetas, precision):
    rows = sc["arena"] + sc["obstacle"]
    n_arena, n_obs = len(sc["arena"]), len(sc["obstacle"])
    semi, ang, off = util.body_poses2d(sc["robot"], thetas)
    ctx.set_scene2d(n_arena, n_obs, np.array(rows), num)
    ctx.set_sweep(sc["lim"], 1, Ly)
    ctx.build_layers2d(semi, ang, off, precision)
    out = []
    for k in range(len(thetas)):
        ref = oracle.layer2d(n_arena, n_obs, np.array(rows), num, semi, ang[k], off[k], sc["lim"], Ly)
        S = ref["bound"].shape[0]
        got = {"bound": np.stack([ctx.boundary(k, s) for s in range(S)])}
        got["lo"], got["up"] = ctx.intervals(k)
        got["off"], got["xL"], got["xU"], got["xM"] = ctx.segments(k)
        out.append((ref, got))
    return out


@pytest.mark.parametrize("name", ["2D_superellipse_sparse_rabbit", "2D_superellipse_cluttered_rabbit", "2D_superellipse_maze2_rabbit",
                                  "2D_superellipse_NAO_lab_rabbit", "2D_ellipse_sparse_rabbit"])
@pytest.mark.parametrize("Ly", [3, 30])
def test_bundled_2d_strict_bit_exact(oracle, gpu_ctx, name, Ly):
    sc = util.scenes()[name]
    res = run_both(oracle, gpu_ctx, sc, 50, Ly, headings(10), capi.F64_STRICT)
    nseg = 0
    for ref, got in res:
        assert np.array_equal(ref["bound"], got["bound"])
        assert np.array_equal(ref["lo"], got["lo"], equal_nan=True) and np.array_equal(ref["up"], got["up"], equal_nan=True)
        assert np.array_equal(ref["off"], got["off"])
        for key in ("xL", "xU", "xM"):
            assert np.array_equal(ref[key], got[key]), key
        nseg += len(ref["xM"])
    assert nseg > 0


def test_reference_kat_minkowski_2d(gpu_ctx):
    """The reference's own known-answer test (test/TestGeometry.cpp:22-32): S({5,3},1.25,{3.2,-2.5},0,50) (+) E({2.5,1.5})
    -> point 0 == (10.7, -2.5)."""
    gpu_ctx.set_scene2d(0, 1, np.array([[5.0, 3.0, 1.25, 3.2, -2.5, 0.0]]), 50)
    gpu_ctx.set_sweep([-1.0, 1.0, -1.0, 1.0], 1, 2)
    gpu_ctx.build_layers2d(np.array([[2.5, 1.5]]), np.array([[0.0]]), np.zeros((1, 1, 2)), capi.F64_STRICT)
    b = gpu_ctx.boundary(0, 0)
    assert b[0, 0] == pytest.approx(10.7, rel=0, abs=4 * np.spacing(10.7))
    assert b[1, 0] == pytest.approx(-2.5, rel=0, abs=4 * np.spacing(2.5))


def test_fast_f64_within_tolerance_2d(oracle, gpu_ctx):
    sc = util.scenes()["2D_superellipse_cluttered_rabbit"]
    for ref, got in run_both(oracle, gpu_ctx, sc, 50, 20, headings(6), capi.F64):
        assert np.max(np.abs(ref["bound"] - got["bound"])) / 100.0 <= 1e-12
        assert np.array_equal(ref["off"], got["off"])
        for key in ("xL", "xU", "xM"):
            assert np.max(np.abs(ref[key] - got[key]), initial=0.0) <= 1e-6


def test_resweep2d_equals_fresh_build(gpu_ctx):
    """2D refinement path: hrmgpu_resweep on the cached curves == a fresh build with the doubled line count."""
    sc = util.scenes()["2D_superellipse_cluttered_rabbit"]
    rows = sc["arena"] + sc["obstacle"]
    num = 50
    semi, ang, off = util.body_poses2d(sc["robot"], np.linspace(-3.0, 3.0, 5))
    gpu_ctx.set_scene2d(len(sc["arena"]), len(sc["obstacle"]), np.array(rows), num)
    gpu_ctx.set_sweep(sc["lim"], 1, 20)
    gpu_ctx.build_layers2d(semi, ang, off, capi.F64_STRICT)
    gpu_ctx.resweep(sc["lim"], 1, 40)
    re = [(gpu_ctx.intervals(k), gpu_ctx.segments(k)) for k in range(5)]
    fresh = capi.Context(0)
    fresh.set_scene2d(len(sc["arena"]), len(sc["obstacle"]), np.array(rows), num)
    fresh.set_sweep(sc["lim"], 1, 40)
    fresh.build_layers2d(semi, ang, off, capi.F64_STRICT)
    for k in range(5):
        (lo, up), seg = re[k]
        flo, fup = fresh.intervals(k)
        assert lo.shape[0] == 40
        assert np.array_equal(lo, flo, equal_nan=True) and np.array_equal(up, fup, equal_nan=True)
        for a, b in zip(seg, fresh.segments(k)):
            assert np.array_equal(a, b)
    fresh.close()
