"""The persistent warp-specialised fused kernel (hrm_b200/csrc/mink3d_ws.cu) takes the fast arithmetic modes at n >= 32: its
producers store the boundary points, its consumers re-evaluate the corners of the kept cells from shared tables, and obstacle
intervals are also appended to per-line buckets that the free-segment pass reads.  Everything it produces is checked against
the CPU oracle (tolerances of BASELINE.json's north_star) and against the one-surface-per-CTA kernel (HRMGPU_CLASSIC=1)."""
import numpy as np
import pytest

from hrm_b200 import capi
from tests import util

pytestmark = pytest.mark.gpu

LIM = [-58.0, 58.0, -28.0, 28.0, -18.0, 18.0]
ROBOT = [[10, 5, 4, 1, 1, 0, 0, 0, 1, 0, 0, 0], [6, 3, 2, 1, 1, -7, 0, 4, 0.70710599, 0, 0.70710802, 0]]


def scene(seed, n_obs, layers, semi_rng=(2.0, 10.0)):
    rows = util.synthetic_scene3d(seed, n_obs, semi_rng=semi_rng)
    rng = np.random.default_rng(1000 + seed)
    semi, R, off, quat = util.body_poses3d(ROBOT, util.random_unit_quats(rng, layers))
    return rows, semi, R, off, quat


def build(ctx, rows, n, Lx, Ly, semi, R, off, prec):
    ctx.set_scene3d(1, len(rows) - 1, util.sq17(rows.tolist(), n), n)
    ctx.set_sweep(LIM, Lx, Ly)
    ctx.build_layers(semi, R, off, prec)
    K = R.shape[0]
    return [(ctx.intervals(k), ctx.segments(k)) for k in range(K)]


def check_against_oracle(oracle, ctx, rows, n, Lx, Ly, semi, R, off, quat, prec, layers):
    got = build(ctx, rows, n, Lx, Ly, semi, R, off, prec)
    nseg = 0
    for k in layers:
        ref = oracle.layer3d(1, len(rows) - 1, rows, n, semi, quat[k], off[k], LIM, Lx, Ly)
        (lo, up), (o, xL, xU, xM) = got[k]
        bound = ctx.boundary_layer(k)
        tol = 1e-12 if prec == capi.F64 else 1e-5
        for s in range(bound.shape[0]):
            scale = max(rows[s // semi.shape[0]][:3]) + np.max(semi)
            assert np.max(np.abs(bound[s] - ref["bound"][s])) / scale <= tol, (k, s)
        if prec == capi.F64:
            assert np.array_equal(np.isnan(lo), np.isnan(ref["lo"]))
            assert np.nanmax(np.abs(lo - ref["lo"])) <= 1e-6 and np.nanmax(np.abs(up - ref["up"])) <= 1e-6
            assert np.array_equal(o, ref["off"])
            for a, b in ((xL, ref["xL"]), (xU, ref["xU"]), (xM, ref["xM"])):
                assert np.max(np.abs(a - b), initial=0.0) <= 1e-6
            nseg += len(xM)
    return nseg


@pytest.mark.parametrize("n,Lx,Ly,n_obs,layers,generic", [(32, 16, 8, 40, 20, False), (64, 20, 12, 12, 3, False), (100, 32, 16, 5, 1, False),
                                                          (100, 32, 16, 5, 1, True)])
def test_ws_f64_against_oracle(oracle, monkeypatch, n, Lx, Ly, n_obs, layers, generic):
    """(32, ...): 820 surfaces, more than the 296 resident CTAs, so every CTA walks several surfaces of different layers;
    (100, ...): fewer surfaces than CTAs; n = 100 runs on the fixed-size instantiation, and once more (generic) on the
    run-time-size one."""
    if generic:
        monkeypatch.setenv("HRMGPU_WS_GENERIC", "1")
    rows, semi, R, off, quat = scene(n, n_obs, layers)
    gpu_ctx = capi.Context(0)
    try:
        _ws_oracle_case(oracle, gpu_ctx, rows, n, Lx, Ly, semi, R, off, quat, layers)
    finally:
        gpu_ctx.close()


def _ws_oracle_case(oracle, gpu_ctx, rows, n, Lx, Ly, semi, R, off, quat, layers):
    nseg = check_against_oracle(oracle, gpu_ctx, rows, n, Lx, Ly, semi, R, off, quat, capi.F64, range(0, layers, max(1, layers // 3)))
    assert nseg > 0


def test_ws_f32_points_within_tolerance(oracle, gpu_ctx):
    rows, semi, R, off, quat = scene(7, 20, 4)
    check_against_oracle(oracle, gpu_ctx, rows, 48, 16, 8, semi, R, off, quat, capi.F32, [0, 3])


@pytest.mark.parametrize("env", [{}, {"HRMGPU_CELL_SEG_CAP": "3"}, {"HRMGPU_BUCKET_CAP": "2"}, {"HRMGPU_WS_VARIANT": "0"}])
def test_ws_equals_one_surface_kernel(env, monkeypatch):
    """Same intervals (bit for bit: both kernels evaluate the same FMA sequence on the same tables? no -- the tables are built
    slightly differently, so: same hit decisions, <= 1e-9 on end points) and the same free segments as the classic kernel; with
    a tiny kept-cell list (multi-round sweep), with tiny buckets (every busy line falls back to the dense tables) and with the
    72 / 96 register split."""
    rows, semi, R, off, quat = scene(11, 30, 6, semi_rng=(0.5, 4.0))
    monkeypatch.setenv("HRMGPU_CLASSIC", "1")
    classic = capi.Context(0)
    ref = build(classic, rows, 40, 24, 12, semi, R, off, capi.F64)
    classic.close()
    monkeypatch.delenv("HRMGPU_CLASSIC")
    for k_, v in env.items():
        monkeypatch.setenv(k_, v)
    ws = capi.Context(0)
    got = build(ws, rows, 40, 24, 12, semi, R, off, capi.F64)
    ws.close()
    total = 0
    for ((lo0, up0), s0), ((lo1, up1), s1) in zip(ref, got):
        assert np.array_equal(np.isnan(lo0), np.isnan(lo1))
        assert np.nanmax(np.abs(lo0 - lo1)) <= 1e-9 and np.nanmax(np.abs(up0 - up1)) <= 1e-9
        assert np.array_equal(s0[0], s1[0])
        for a, b in zip(s0[1:], s1[1:]):
            assert np.max(np.abs(a - b), initial=0.0) <= 1e-9
        total += len(s0[1])
    assert total > 500


def test_ws_far_away_surface_takes_the_range_checked_path(monkeypatch):
    """Phase A drops the per-vertex range test of the fixed-point line index for surfaces it proves to be within 2^15 line
    spacings of the grid; an obstacle 10^6 units away (and one 10^5 units long) fails the proof and must give the same
    intervals as the one-surface kernel, as must all the other surfaces of the layer."""
    rows, semi, R, off, quat = scene(13, 12, 3, semi_rng=(0.5, 4.0))
    rows = np.array(rows, dtype=float)
    rows[3, 5] = 1.0e6      # centre x of obstacle 2
    rows[5, 0] = 1.0e5      # first semi-axis of obstacle 4: a needle through the arena
    monkeypatch.setenv("HRMGPU_CLASSIC", "1")
    classic = capi.Context(0)
    ref = build(classic, rows, 40, 24, 12, semi, R, off, capi.F64)
    classic.close()
    monkeypatch.delenv("HRMGPU_CLASSIC")
    ws = capi.Context(0)
    got = build(ws, rows, 40, 24, 12, semi, R, off, capi.F64)
    ws.close()
    for ((lo0, up0), s0), ((lo1, up1), s1) in zip(ref, got):
        assert np.array_equal(np.isnan(lo0), np.isnan(lo1))
        assert np.nanmax(np.abs(lo0 - lo1)) <= 1e-6 and np.nanmax(np.abs(up0 - up1)) <= 1e-6
        assert np.array_equal(s0[0], s1[0])


def test_ws_resweep_and_queries_use_the_stored_points(gpu_ctx):
    """The points the producers stored are what resweep / edge tests read later: a resweep at the same grid (classic kernel in
    reload mode, vertices read back from HBM) must reproduce the warp-specialised kernel's intervals exactly -- i.e. the
    consumers' re-evaluated corners are bit-identical to the stored ones."""
    rows, semi, R, off, quat = scene(5, 25, 3)
    first = build(gpu_ctx, rows, 36, 20, 10, semi, R, off, capi.F64)
    gpu_ctx.resweep(LIM, 20, 10)
    for k in range(3):
        lo, up = gpu_ctx.intervals(k)
        assert np.array_equal(lo, first[k][0][0], equal_nan=True) and np.array_equal(up, first[k][0][1], equal_nan=True)
        for a, b in zip(gpu_ctx.segments(k), first[k][1]):
            assert np.array_equal(a, b)
