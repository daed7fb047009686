"""GPU parity of the 3D C-layer construction (K1 Minkowski + K2 sweep + K3 free segments) against the
CPU oracle, through the C ABI.  STRICT mode must be bit-identical; the fast FP64 and FP32 variants must
meet the tolerances BASELINE.json's north_star states (1e-12 / 1e-5 relative on boundary points,
1e-6 on free-segment end points)."""
import numpy as np
import pytest

from hrm_b200 import capi
from tests import util

pytestmark = pytest.mark.gpu

REL64 = 1e-12   # north_star: FP64 kernel variant matches to 1e-12 (relative to the surface scale)
REL32 = 1e-5    # north_star: Minkowski boundary points within 1e-5 relative
SEG_TOL = 1e-6  # north_star: free-segment endpoints within 1e-6


def run_both(oracle, ctx, rows, n_arena, n, semi, R, off, quat, lim, Lx, Ly, precision, layers=None):
    n_obs = len(rows) - n_arena
    ctx.set_scene3d(n_arena, n_obs, util.sq17(rows, n), n)
    ctx.set_sweep(lim, Lx, Ly)
    ctx.build_layers(semi, R, off, precision)
    K = R.shape[0]
    out = []
    for k in (layers if layers is not None else range(K)):
        ref = oracle.layer3d(n_arena, n_obs, np.array(rows), n, semi, quat[k], off[k], lim, Lx, Ly)
        S = ref["bound"].shape[0]
        got = {"bound": np.stack([ctx.boundary(k, s) for s in range(S)])}
        got["lo"], got["up"] = ctx.intervals(k)
        got["off"], got["xL"], got["xU"], got["xM"] = ctx.segments(k)
        out.append((ref, got))
    return out


def assert_bit_exact(ref, got):
    assert np.array_equal(ref["bound"], got["bound"]), "boundary points differ"
    assert np.array_equal(ref["lo"], got["lo"], equal_nan=True)
    assert np.array_equal(ref["up"], got["up"], equal_nan=True)
    assert np.array_equal(ref["off"], got["off"])
    for key in ("xL", "xU", "xM"):
        assert np.array_equal(ref[key], got[key]), key


@pytest.mark.parametrize("name", ["3D_superquadrics_cluttered_rabbit", "3D_superquadrics_sparse_rabbit",
                                  "3D_superquadrics_maze_rabbit", "3D_superquadrics_narrow_rabbit"])
def test_bundled_frozen_pose_strict_bit_exact(oracle, gpu_ctx, name):
    sc = util.scenes()[name]
    rows = sc["arena"] + sc["obstacle"]
    semi, R, off, quat = util.frozen_pose3d(sc["robot"])
    for n in (10, 20):
        (ref, got), = run_both(oracle, gpu_ctx, rows, len(sc["arena"]), n, semi, R, off, quat, sc["lim"], sc["Lx"], sc["Ly"],
                               capi.F64_STRICT)
        assert_bit_exact(ref, got)
        assert gpu_ctx.single_hit_count() == ref["single_hits"]
        assert ref["off"][-1] > 0


def test_bundled_per_orientation_strict_bit_exact(oracle, gpu_ctx):
    sc = util.scenes()["3D_superquadrics_cluttered_rabbit"]
    rows = sc["arena"] + sc["obstacle"]
    semi, R, off, quat = util.body_poses3d(sc["robot"], sc["quats"])
    res = run_both(oracle, gpu_ctx, rows, 1, 10, semi, R, off, quat, sc["lim"], sc["Lx"], sc["Ly"], capi.F64_STRICT)
    assert len(res) == 60
    for ref, got in res:
        assert_bit_exact(ref, got)


def test_home_snake_strict_bit_exact(oracle, gpu_ctx):
    sc = util.scenes()["3D_superquadrics_home_snake"]
    rows = sc["arena"] + sc["obstacle"]
    semi, R, off, quat = util.body_poses3d(sc["robot"], sc["quats"][:6])
    res = run_both(oracle, gpu_ctx, rows, 1, 10, semi, R, off, quat, sc["lim"], 20, 20, capi.F64_STRICT)
    for ref, got in res:
        assert_bit_exact(ref, got)


@pytest.mark.parametrize("seed,n_obs,n,Lx,Ly", [(1, 12, 16, 8, 6), (2, 40, 25, 16, 8), (3, 5, 7, 3, 2), (4, 30, 33, 32, 16)])
def test_synthetic_strict_bit_exact(oracle, gpu_ctx, seed, n_obs, n, Lx, Ly):
    rows = util.synthetic_scene3d(seed, n_obs)
    rng = np.random.default_rng(100 + seed)
    robot = [[10, 5, 4, 1, 1, 0, 0, 0, 1, 0, 0, 0], [6, 3, 2, 1, 1, -7, 0, 4, 0.70710599, 0, 0.70710802, 0]]
    semi, R, off, quat = util.body_poses3d(robot, util.random_unit_quats(rng, 3))
    lim = [-58.0, 58.0, -28.0, 28.0, -18.0, 18.0]
    for ref, got in run_both(oracle, gpu_ctx, rows.tolist(), 1, n, semi, R, off, quat, lim, Lx, Ly, capi.F64_STRICT):
        assert_bit_exact(ref, got)


def scale_of(rows, semi):
    return np.max(np.array(rows)[:, :3]) + np.max(semi)


@pytest.mark.parametrize("precision,tol", [(capi.F64, REL64), (capi.F32, REL32)])
def test_fast_variants_within_tolerance(oracle, gpu_ctx, precision, tol):
    sc = util.scenes()["3D_superquadrics_cluttered_rabbit"]
    rows = sc["arena"] + sc["obstacle"]
    semi, R, off, quat = util.body_poses3d(sc["robot"], sc["quats"][:8])
    res = run_both(oracle, gpu_ctx, rows, 1, 20, semi, R, off, quat, sc["lim"], sc["Lx"], sc["Ly"], precision)
    for ref, got in res:
        S = ref["bound"].shape[0]
        for s in range(S):
            obj = s // semi.shape[0]
            scale = max(rows[obj][:3]) + np.max(semi)
            err = np.max(np.abs(ref["bound"][s] - got["bound"][s])) / scale
            assert err <= tol, (s, err)
        if precision == capi.F64:
            # same hit decisions, same segment structure, end points within 1e-6
            assert np.array_equal(np.isnan(ref["lo"]), np.isnan(got["lo"]))
            assert np.array_equal(ref["off"], got["off"])
            for key in ("xL", "xU", "xM"):
                assert np.max(np.abs(ref[key] - got[key]), initial=0.0) <= SEG_TOL


def test_roundtrip_properties_full_size(gpu_ctx):
    """Size-independent properties at the BASELINE synthetic sizes (1001 surfaces, n=100, 32x16 lines),
    where the oracle would take minutes: every obstacle interval is ordered, free segments lie inside the
    arena interval, are ordered and disjoint from every obstacle interval of their line."""
    rows = util.synthetic_scene3d(20220726, 1000)
    rng = np.random.default_rng(7)
    semi = np.array([[10.0, 5.0, 4.0]])
    quats = util.random_unit_quats(rng, 2)
    from hrm_b200 import hostmath as hm
    R = np.array([[hm.flat9(hm.quat_to_rot(q))] for q in quats])
    off = np.zeros((2, 1, 3))
    lim = [-58.0, 58.0, -28.0, 28.0, -18.0, 18.0]
    gpu_ctx.set_scene3d(1, 1000, util.sq17(rows.tolist(), 100), 100)
    gpu_ctx.set_sweep(lim, 32, 16)
    gpu_ctx.build_layers(semi, R, off, capi.F64_STRICT)
    for k in range(2):
        lo, up = gpu_ctx.intervals(k)
        ok = ~np.isnan(lo[:, 1:])
        assert np.all(lo[:, 1:][ok] <= up[:, 1:][ok])
        assert ok.sum() > 1000
        offs, xL, xU, xM = gpu_ctx.segments(k)
        for l in range(512):
            a, b = offs[l], offs[l + 1]
            for j in range(a, b):
                assert lo[l, 0] <= xL[j] <= xU[j] <= up[l, 0]
                inside = ok[l] & (lo[l, 1:] < xM[j]) & (xM[j] < up[l, 1:])
                assert not inside.any()
        b = gpu_ctx.boundary(k, 5)
        assert b.shape == (3, 10000) and np.isfinite(b).all()


def test_bench_sizes_strict_bit_exact_on_a_sample(oracle, gpu_ctx):
    """The bench's own grid (n = 100, 32 x 16 lines, rabbit base body) on the arena + the first 40 obstacles of the bench
    scene: every point, interval and segment of one layer bit-identical to the oracle (the full 1001 surfaces take the oracle minutes)."""
    rows = util.synthetic_scene3d(20220726, 1000)[:41]
    rng = np.random.default_rng(11)
    robot = [[10, 5, 4, 1, 1, 0, 0, 0, 1, 0, 0, 0]]
    semi, R, off, quat = util.body_poses3d(robot, util.random_unit_quats(rng, 1))
    lim = [-58.0, 58.0, -28.0, 28.0, -18.0, 18.0]
    (ref, got), = run_both(oracle, gpu_ctx, rows.tolist(), 1, 100, semi, R, off, quat, lim, 32, 16, capi.F64_STRICT)
    assert_bit_exact(ref, got)
    assert got["off"][-1] > 100


def test_bench_mode_against_strict_full_size(gpu_ctx):
    """The bench's headline mode (F64) against the bit-exact STRICT mode at the full BASELINE sizes (1001 surfaces x 10^4 points
    x 512 lines, 2 layers): points within 1e-12 of the surface scale, the same hit decisions, the same segment structure,
    end points within 1e-6; the F32 variant within 1e-5 on the points."""
    rows = util.synthetic_scene3d(20220726, 1000)
    rng = np.random.default_rng(13)
    robot = [[10, 5, 4, 1, 1, 0, 0, 0, 1, 0, 0, 0]]
    semi, R, off, quat = util.body_poses3d(robot, util.random_unit_quats(rng, 2))
    lim = [-58.0, 58.0, -28.0, 28.0, -18.0, 18.0]
    gpu_ctx.set_scene3d(1, 1000, util.sq17(rows.tolist(), 100), 100)
    gpu_ctx.set_sweep(lim, 32, 16)
    res = {}
    sample = list(range(0, 1001, 37))
    for prec in (capi.F64_STRICT, capi.F64, capi.F32):
        gpu_ctx.build_layers(semi, R, off, prec)
        res[prec] = [(gpu_ctx.intervals(k), gpu_ctx.segments(k), [gpu_ctx.boundary(k, s) for s in sample]) for k in range(2)]
    for k in range(2):
        (lo0, up0), (o0, xL0, xU0, xM0), b0 = res[capi.F64_STRICT][k]
        (lo1, up1), (o1, xL1, xU1, xM1), b1 = res[capi.F64][k]
        assert np.array_equal(np.isnan(lo0), np.isnan(lo1))
        assert np.nanmax(np.abs(lo0 - lo1)) <= SEG_TOL and np.nanmax(np.abs(up0 - up1)) <= SEG_TOL
        assert np.array_equal(o0, o1)
        assert max(np.max(np.abs(xL0 - xL1)), np.max(np.abs(xU0 - xU1)), np.max(np.abs(xM0 - xM1))) <= SEG_TOL
        b2 = res[capi.F32][k][2]
        for i, s in enumerate(sample):
            scale = max(rows[s][:3]) + 10.0
            assert np.max(np.abs(b0[i] - b1[i])) / scale <= REL64, s
            assert np.max(np.abs(b0[i] - b2[i])) / scale <= REL32, s
        (lo2, up2), _, _ = res[capi.F32][k]
        assert (np.isnan(lo0) == np.isnan(lo2)).mean() > 0.9995  # FP32 points may flip a hit next to a sweep line


@pytest.mark.parametrize("env", [{"HRMGPU_SEG_FAST_CAP": "2"}, {"HRMGPU_SEG_ONE_LINE": "1"}])
def test_segment_pass_variants_bit_exact(oracle, env, monkeypatch):
    """K3 pass 1 has a 4-lines-per-warp fast path with short interval lists and a one-line-per-warp pass with full-size
    lists for lines that overflow them.  A list capacity of 2 sends every busy line through the fallback; the second
    variant disables the fast path.  Both must still match the oracle bit for bit (L = 8 x 12 is divisible by 4)."""
    for k_, v in env.items():
        monkeypatch.setenv(k_, v)
    ctx = capi.Context(0)  # the hooks are read when the context is created
    sc = util.scenes()["3D_superquadrics_cluttered_rabbit"]
    rows = sc["arena"] + sc["obstacle"]
    semi, R, off, quat = util.body_poses3d(sc["robot"], sc["quats"][:3])
    res = run_both(oracle, ctx, rows, 1, 10, semi, R, off, quat, sc["lim"], 8, 12, capi.F64_STRICT)
    assert any((np.diff(ref["off"]) > 2).any() for ref, _ in res)
    for ref, got in res:
        assert_bit_exact(ref, got)


@pytest.mark.parametrize("cap", ["1", "5"])
def test_cell_list_overflow_rounds_bit_exact(oracle, cap, monkeypatch):
    """The fused kernel lists the kept mesh cells of a surface in per-warp segments (256 cells each) and tests them behind a
    barrier; when a segment is full the scan parks and another round follows.  No realistic surface fills a segment, so a
    capacity of 1 / 5 cells forces tens of rounds (incl. the heavy-cell list of the arena): results stay bit-identical."""
    monkeypatch.setenv("HRMGPU_CELL_SEG_CAP", cap)
    ctx = capi.Context(0)  # the hook is read when the context is created
    rows = util.synthetic_scene3d(21, 10)
    rng = np.random.default_rng(121)
    robot = [[10, 5, 4, 1, 1, 0, 0, 0, 1, 0, 0, 0], [6, 3, 2, 1, 1, -7, 0, 4, 0.70710599, 0, 0.70710802, 0]]
    semi, R, off, quat = util.body_poses3d(robot, util.random_unit_quats(rng, 2))
    lim = [-58.0, 58.0, -28.0, 28.0, -18.0, 18.0]
    for n, Lx, Ly in ((24, 16, 8), (13, 5, 140)):
        for ref, got in run_both(oracle, ctx, rows.tolist(), 1, n, semi, R, off, quat, lim, Lx, Ly, capi.F64_STRICT):
            assert_bit_exact(ref, got)
            assert got["off"][-1] > 10
    ctx.close()


@pytest.mark.parametrize("precision", [capi.F64_STRICT, capi.F64, capi.F32])
def test_resweep_equals_fresh_build(oracle, gpu_ctx, precision):
    """Refinement path (hrmgpu_resweep): doubling the sweep lines on the cached boundaries must give exactly the intervals
    and free segments of a fresh build at the doubled grid -- and, in strict mode, the oracle's."""
    sc = util.scenes()["3D_superquadrics_cluttered_rabbit"]
    rows = sc["arena"] + sc["obstacle"]
    n = 12
    semi, R, off, quat = util.body_poses3d(sc["robot"], sc["quats"][:3])
    Lx, Ly = sc["Lx"], sc["Ly"]
    gpu_ctx.set_scene3d(1, len(sc["obstacle"]), util.sq17(rows, n), n)
    gpu_ctx.set_sweep(sc["lim"], Lx, Ly)
    gpu_ctx.build_layers(semi, R, off, precision)
    bound0 = gpu_ctx.boundary(1, 2).copy()
    gpu_ctx.resweep(sc["lim"], 2 * Lx, 2 * Ly)
    re = [(gpu_ctx.intervals(k), gpu_ctx.segments(k)) for k in range(3)]
    assert np.array_equal(gpu_ctx.boundary(1, 2), bound0)  # boundaries untouched
    fresh = capi.Context(0)
    fresh.set_scene3d(1, len(sc["obstacle"]), util.sq17(rows, n), n)
    fresh.set_sweep(sc["lim"], 2 * Lx, 2 * Ly)
    fresh.build_layers(semi, R, off, precision)
    for k in range(3):
        (lo, up), seg = re[k]
        flo, fup = fresh.intervals(k)
        assert np.array_equal(lo, flo, equal_nan=True) and np.array_equal(up, fup, equal_nan=True)
        for a, b in zip(seg, fresh.segments(k)):
            assert np.array_equal(a, b)
        if precision == capi.F64_STRICT:
            ref = oracle.layer3d(1, len(sc["obstacle"]), np.array(rows), n, semi, quat[k], off[k], sc["lim"], 2 * Lx, 2 * Ly)
            assert np.array_equal(lo, ref["lo"], equal_nan=True) and np.array_equal(seg[3], ref["xM"])
    fresh.close()
    # edge tests still run on the cached meshes
    assert gpu_ctx.test_edges(0, np.zeros((1, 3)), np.ones((1, 3))).shape == (1,)
