"""Edge cases of the C ABI on the GPU: empty inputs, arena-only scenes, the 2 x 16-bit line-code path (more than 127 sweep
lines on an axis), the n_grid limits, and the error behaviour (status codes, never a crash)."""
import numpy as np
import pytest

from hrm_b200 import capi
from tests import util
from tests.test_gpu_layers3d import assert_bit_exact, run_both

pytestmark = pytest.mark.gpu

LIM = [-58.0, 58.0, -28.0, 28.0, -18.0, 18.0]
ROBOT = [[10, 5, 4, 1, 1, 0, 0, 0, 1, 0, 0, 0], [6, 3, 2, 1, 1, -7, 0, 4, 0.70710599, 0, 0.70710802, 0]]


def poses(seed, K):
    rng = np.random.default_rng(seed)
    return util.body_poses3d(ROBOT, util.random_unit_quats(rng, K))


def test_arena_only_scene_strict_bit_exact(oracle, gpu_ctx):
    """No obstacle at all: every sweep line that hits the arena yields exactly one free segment (its arena interval)."""
    rows = util.synthetic_scene3d(5, 0)
    assert len(rows) == 1
    semi, R, off, quat = poses(50, 2)
    for ref, got in run_both(oracle, gpu_ctx, rows.tolist(), 1, 12, semi, R, off, quat, LIM, 6, 4, capi.F64_STRICT):
        assert_bit_exact(ref, got)
        assert np.all(np.diff(got["off"]) <= 1)
    # K4 / K5 with no obstacle mesh: everything is free
    v = np.zeros((3, 3))
    assert gpu_ctx.test_edges(0, v, v + 1.0).all()
    assert gpu_ctx.test_edges_multi(np.array([0, 1, 0]), v, v + 1.0).all()


@pytest.mark.parametrize("Lx,Ly", [(140, 3), (3, 130), (129, 60)])
def test_wide_line_codes_strict_bit_exact(oracle, gpu_ctx, Lx, Ly):
    """More than 127 sweep lines on an axis: the vertex line codes switch from 2 x 8 to 2 x 16 bits."""
    rows = util.synthetic_scene3d(6, 6)
    semi, R, off, quat = poses(60, 1)
    for ref, got in run_both(oracle, gpu_ctx, rows.tolist(), 1, 12, semi, R, off, quat, LIM, Lx, Ly, capi.F64_STRICT):
        assert_bit_exact(ref, got)
        assert got["off"][-1] > 0


@pytest.mark.parametrize("precision", [capi.F64, capi.F32])
def test_wide_line_codes_fast_modes_same_decisions(oracle, gpu_ctx, precision):
    rows = util.synthetic_scene3d(6, 6)
    semi, R, off, quat = poses(61, 1)
    (ref, got), = run_both(oracle, gpu_ctx, rows.tolist(), 1, 16, semi, R, off, quat, LIM, 150, 4, precision)
    same = np.isnan(ref["lo"]) == np.isnan(got["lo"])
    assert same.mean() > 0.999  # a boundary point within rounding distance of a sweep line may flip a hit in the fast modes
    if precision == capi.F64:
        assert same.all()
        assert np.array_equal(ref["off"], got["off"])
        assert np.max(np.abs(ref["xM"] - got["xM"]), initial=0.0) <= 1e-6


@pytest.mark.parametrize("n", [3, 4, 5, 200])
def test_grid_size_limits_strict_bit_exact(oracle, gpu_ctx, n):
    """Smallest meshes (n = 3: 8 faces) and the largest supported parameter grid (n = 200: 40 000 points per surface)."""
    rows = util.synthetic_scene3d(7, 2)
    semi, R, off, quat = poses(70, 1)
    (ref, got), = run_both(oracle, gpu_ctx, rows.tolist(), 1, n, semi[:1], R[:, :1], off[:, :1], quat[:, :1], LIM, 5, 4, capi.F64_STRICT)
    assert_bit_exact(ref, got)


def test_empty_batches(gpu_ctx):
    rows = util.synthetic_scene3d(8, 3)
    semi, R, off, quat = poses(80, 1)
    gpu_ctx.set_scene3d(1, 3, util.sq17(rows.tolist(), 10), 10)
    gpu_ctx.set_sweep(LIM, 4, 3)
    gpu_ctx.build_layers(semi, R, off, capi.F64_STRICT)
    e = np.zeros((0, 3))
    assert gpu_ctx.test_edges(0, e, e).shape == (0,)
    assert gpu_ctx.test_edges_multi(np.zeros(0, dtype=np.int32), e, e).shape == (0,)
    B = semi.shape[0]
    gpu_ctx.build_bridge(np.ones((1, B, 3)), np.tile(np.eye(3).reshape(1, 1, 9), (1, B, 1)), capi.F64_STRICT)
    assert gpu_ctx.test_bridge(0, np.zeros((0, B, 3))).shape == (0,)
    assert gpu_ctx.test_bridge_multi(np.zeros(0, dtype=np.int32), np.zeros((0, B, 3))).shape == (0,)
    w = np.array([1.0, 0.0])
    assert gpu_ctx.test_transitions(np.zeros(0, dtype=np.int32), e, e, w, w[::-1].copy(), np.zeros((1, 2, B, 3))).shape == (0,)


def test_errors_are_status_codes_not_crashes():
    ctx = capi.Context(0)
    try:
        rows = util.synthetic_scene3d(9, 3)
        semi, R, off, quat = poses(90, 1)
        with pytest.raises(RuntimeError):  # no scene yet
            ctx.build_layers(semi, R, off, capi.F64_STRICT)
        ctx.set_scene3d(1, 3, util.sq17(rows.tolist(), 10), 10)
        with pytest.raises(RuntimeError):  # no sweep grid yet
            ctx.build_layers(semi, R, off, capi.F64_STRICT)
        with pytest.raises(RuntimeError):  # the reference divides by numLine - 1
            ctx.set_sweep(LIM, 1, 4)
        with pytest.raises(RuntimeError):  # more than 8192 sweep lines per layer
            ctx.set_sweep(LIM, 129, 128)
        with pytest.raises(RuntimeError):  # n_grid beyond the shared-memory limit
            ctx.set_scene3d(1, 3, util.sq17(rows.tolist(), 201), 201)
        ctx.set_scene3d(1, 3, util.sq17(rows.tolist(), 10), 10)
        ctx.set_sweep(LIM, 4, 3)
        with pytest.raises(RuntimeError):  # nothing built yet
            ctx.segments(0)
        with pytest.raises(RuntimeError):
            ctx.test_edges(0, np.zeros((1, 3)), np.ones((1, 3)))
        with pytest.raises(RuntimeError):
            ctx.test_bridge(0, np.zeros((1, 2, 3)))
        ctx.build_layers(semi, R, off, capi.F64_STRICT)
        with pytest.raises(RuntimeError):  # layer index out of range
            ctx.segments(1)
        with pytest.raises(RuntimeError):
            ctx.build_layers(semi, R, off, 7)  # unknown precision
        assert len(ctx.last_error()) > 0
        o, xL, xU, xM = ctx.segments(0)  # the context is still usable after the failures
        assert o[-1] == len(xM)
        # hrmgpu_reset: back to "no scene, nothing built" with the buffers kept; then usable again with the same results
        assert ctx.lib.hrmgpu_reset(ctx.h) == 0
        with pytest.raises(RuntimeError):
            ctx.segments(0)
        with pytest.raises(RuntimeError):
            ctx.build_layers(semi, R, off, capi.F64_STRICT)
        ctx.set_scene3d(1, 3, util.sq17(rows.tolist(), 10), 10)
        ctx.set_sweep(LIM, 4, 3)
        ctx.build_layers(semi, R, off, capi.F64_STRICT)
        o2, xL2, xU2, xM2 = ctx.segments(0)
        assert np.array_equal(o, o2) and np.array_equal(xM, xM2)
    finally:
        ctx.close()


@pytest.mark.parametrize("K,Lx,Ly", [(16, 32, 16), (17, 32, 16), (3, 91, 90), (5, 7, 3)])
def test_segment_offsets_across_scan_tiles(gpu_ctx, K, Lx, Ly):
    """The CSR offsets of a batch come from a chained multi-CTA scan over K * Lx * Ly line counts (tiles of 8192): exactly one
    tile (16 x 512), one line more than a tile boundary allows (17 x 512 = 2 tiles), a ragged count (3 x 8190) and a single small
    tile must all give, layer by layer, the segments of that layer built on its own (a one-tile scan)."""
    rows = util.synthetic_scene3d(7, 12, semi_rng=(3.0, 9.0))
    rng = np.random.default_rng(5)
    robot = [[6, 4, 3, 1, 1, 0, 0, 0, 1, 0, 0, 0]]
    semi, R, off, quat = util.body_poses3d(robot, util.random_unit_quats(rng, K))
    lim = [-58.0, 58.0, -28.0, 28.0, -18.0, 18.0]
    gpu_ctx.set_scene3d(1, 12, util.sq17(rows.tolist(), 20), 20)
    gpu_ctx.set_sweep(lim, Lx, Ly)
    gpu_ctx.build_layers(semi, R, off, capi.F64_STRICT)
    batch = [gpu_ctx.segments(k) for k in range(K)]
    total = 0
    for k in sorted({0, 1, K // 2, K - 1}):
        gpu_ctx.build_layers(semi, R[k:k + 1], off[k:k + 1], capi.F64_STRICT)
        alone = gpu_ctx.segments(0)
        for a, b in zip(batch[k], alone):
            assert np.array_equal(a, b)
        total += len(alone[1])
    assert total > 0
