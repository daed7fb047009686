"""Full-size parity of the headline path: ONE complete C-layer of the bench workloads (1001 surfaces x 10^4 boundary points x
512 sweep lines, BASELINE.json configs[4]) against the CPU oracle itself -- not against another GPU mode.

  * HRMGPU_F64_STRICT: every interval, every free segment and a sample of the boundary points bit-identical;
  * HRMGPU_F64 (the mode bench.py times): the same hit / no-hit pattern on all 512 x 1001 (line, surface) pairs, interval
    end points and free-segment end points within 1e-6 (north_star), boundary points within 1e-12 of the surface scale.

Two scenes: "scaling" (the C-obstacles cover nearly everything, a layer keeps a handful of free segments) and
"freespace" (same generator, obstacle semi-axes U[0.5,2]: hundreds of free segments per layer, so the complement / enhance /
CSR passes are compared on real data).  The oracle runs its surfaces on all host threads (oracle.hrmo.layer3d_parallel)."""
import numpy as np
import pytest

from hrm_b200 import capi
from hrm_b200 import workload as wl

pytestmark = pytest.mark.gpu

SEG_TOL = 1e-6   # north_star: free-segment end points within 1e-6
REL64 = 1e-12    # north_star: FP64 kernel variant matches to 1e-12
SAMPLE = list(range(0, 1001, 53))  # surfaces whose boundary points are compared (arena + 18 obstacles)


@pytest.fixture(scope="module", params=["scaling", "freespace"])
def full_layer(request, oracle):
    name = request.param
    rows = wl.scene_rows12(wl.N_OBS, workload=name)
    quat = wl.orientations(1, seed=wl.SEED + 31).reshape(1, 1, 4)
    semi = np.array(wl.BODY_SEMI)
    off = np.zeros((1, 1, 3))
    ref = oracle.layer3d_parallel(1, wl.N_OBS, rows, wl.N_GRID, semi, quat[0], off[0], wl.LIM, wl.LX, wl.LY)
    bounds = {}
    for s in SAMPLE:  # boundary of single surfaces: one-object scenes give the same points (surfaces are independent)
        r = oracle.layer3d(1 if s == 0 else 0, 0 if s == 0 else 1, rows[s:s + 1], wl.N_GRID, semi, quat[0], off[0], wl.LIM, wl.LX, wl.LY)
        bounds[s] = r["bound"][0]
    return name, rows, semi, wl.body_rotations(quat[:, 0]), off, ref, bounds


def build(ctx, rows, semi, R, off, precision):
    ctx.set_scene3d(1, wl.N_OBS, wl.scene_rows17(rows), wl.N_GRID)
    ctx.set_sweep(wl.LIM, wl.LX, wl.LY)
    ctx.build_layers(semi, R, off, precision)
    lo, up = ctx.intervals(0)
    return lo, up, ctx.segments(0)


def test_full_layer_strict_bit_exact(gpu_ctx, full_layer):
    name, rows, semi, R, off, ref, bounds = full_layer
    lo, up, (o, xL, xU, xM) = build(gpu_ctx, rows, semi, R, off, capi.F64_STRICT)
    assert np.array_equal(lo, ref["lo"], equal_nan=True) and np.array_equal(up, ref["up"], equal_nan=True)
    assert np.array_equal(o, ref["off"])
    assert np.array_equal(xL, ref["xL"]) and np.array_equal(xU, ref["xU"]) and np.array_equal(xM, ref["xM"])
    assert gpu_ctx.single_hit_count() == ref["single_hits"]
    for s in SAMPLE:
        assert np.array_equal(gpu_ctx.boundary(0, s), bounds[s]), s
    if name == "freespace":
        assert (np.diff(ref["off"]) > 0).mean() >= 0.3 and ref["off"][-1] >= 200  # the scene really has free space


def test_full_layer_bench_mode_against_oracle(gpu_ctx, full_layer):
    name, rows, semi, R, off, ref, bounds = full_layer
    lo, up, (o, xL, xU, xM) = build(gpu_ctx, rows, semi, R, off, capi.F64)
    flips = int((np.isnan(lo) != np.isnan(ref["lo"])).sum())
    assert flips == 0, "%d of %d hit decisions differ from the oracle" % (flips, lo.size)
    assert np.nanmax(np.abs(lo - ref["lo"])) <= SEG_TOL and np.nanmax(np.abs(up - ref["up"])) <= SEG_TOL
    assert np.array_equal(o, ref["off"])
    for a, b in ((xL, ref["xL"]), (xU, ref["xU"]), (xM, ref["xM"])):
        assert np.max(np.abs(a - b), initial=0.0) <= SEG_TOL
    for s in SAMPLE:
        scale = max(rows[s][:3]) + max(semi[0])
        assert np.max(np.abs(gpu_ctx.boundary(0, s) - bounds[s])) / scale <= REL64, s
