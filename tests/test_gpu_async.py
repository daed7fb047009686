"""The build calls only queue work: the fused kernel on the main stream, the free-segment passes on the side stream behind it,
with capacity-bounded buffers instead of a size read-back in the middle of a build.  These tests pin the host-visible
behaviour of that design through the C ABI: results identical to the oracle whatever the capacities, pipelined read-back,
layer-wide boundary getter, stream semantics of hrmgpu_pack_segments_dev."""
import numpy as np
import pytest

from hrm_b200 import capi
from tests import util

pytestmark = pytest.mark.gpu


def cluttered(n_layers, n=10):
    sc = util.scenes()["3D_superquadrics_cluttered_rabbit"]
    rows = sc["arena"] + sc["obstacle"]
    semi, R, off, quat = util.body_poses3d(sc["robot"], sc["quats"][:n_layers])
    return sc, rows, semi, R, off, quat


def setup(ctx, sc, rows, n=10, Lx=None, Ly=None):
    ctx.set_scene3d(1, len(sc["obstacle"]), util.sq17(rows, n), n)
    ctx.set_sweep(sc["lim"], Lx or sc["Lx"], Ly or sc["Ly"])


@pytest.mark.parametrize("init_cap", ["1", "7", "200"])
def test_capacity_overflow_repeats_passes_bit_exact(oracle, init_cap, monkeypatch):
    """Segment pool and payload start far too small (test hook): the passes report the sizes they need, the first getter grows
    the buffers and repeats them, and the result is still the oracle's bit for bit."""
    monkeypatch.setenv("HRMGPU_SEG_INIT_CAP", init_cap)
    ctx = capi.Context(0)
    sc, rows, semi, R, off, quat = cluttered(4)
    setup(ctx, sc, rows, Lx=8, Ly=12)
    ctx.build_layers(semi, R, off, capi.F64_STRICT)
    d = ctx.dims()
    assert ctx.segment_redos() >= 1 and d["segments"] > 200
    for k in range(4):
        ref = oracle.layer3d(1, len(sc["obstacle"]), np.array(rows), 10, semi, quat[k], off[k], sc["lim"], 8, 12)
        o, xL, xU, xM = ctx.segments(k)
        assert np.array_equal(o, ref["off"]) and np.array_equal(xL, ref["xL"]) and np.array_equal(xU, ref["xU"]) and np.array_equal(xM, ref["xM"])
    # the next build of the same size fits at once
    ctx.segment_redos(reset=True)
    ctx.build_layers(semi, R, off, capi.F64_STRICT)
    assert ctx.dims()["segments"] == d["segments"] and ctx.segment_redos() == 0
    ctx.close()


def test_pipelined_fetch_matches_synchronous_getters(gpu_ctx):
    """fetch_segments_async(build A) ... build B ... wait: A's arrays are A's although B was queued in between; B's getters
    return B."""
    import torch

    sc, rows, semi, R, off, quat = cluttered(12)
    setup(gpu_ctx, sc, rows)
    gpu_ctx.build_layers(semi, R[:6], off[:6], capi.F64_STRICT)
    refA = gpu_ctx.segments_all()
    gpu_ctx.build_layers(semi, R[6:], off[6:], capi.F64_STRICT)
    refB = gpu_ctx.segments_all()
    assert not np.array_equal(refA[3], refB[3])
    nl = 6 * sc["Lx"] * sc["Ly"]
    cap = 2 * max(len(refA[1]), len(refB[1]))
    pin = [torch.zeros(nl + 1, dtype=torch.int64).pin_memory()] + [torch.zeros(cap, dtype=torch.float64).pin_memory() for _ in range(3)]
    arr = [t.numpy() for t in pin]
    gpu_ctx.build_layers(semi, R[:6], off[:6], capi.F64_STRICT)
    gpu_ctx.fetch_segments_async(*arr)
    gpu_ctx.build_layers(semi, R[6:], off[6:], capi.F64_STRICT)  # queued while A's passes / copies may still run
    tot = gpu_ctx.wait_segments()
    assert tot == len(refA[1]) and np.array_equal(arr[0], refA[0])
    for a, r in zip(arr[1:], refA[1:]):
        assert np.array_equal(a[:tot], r)
    gotB = gpu_ctx.segments_all()
    for a, r in zip(gotB, refB):
        assert np.array_equal(a, r)
    # a fetch whose host arrays are too small reports it
    small = [arr[0]] + [a[:3] for a in arr[1:]]
    gpu_ctx.fetch_segments_async(*small)
    with pytest.raises(capi.HrmGpuError) as e:
        gpu_ctx.wait_segments()
    assert e.value.code == capi.E_CAP


def test_boundary_layer_equals_per_surface_getter(gpu_ctx):
    sc, rows, semi, R, off, quat = cluttered(3)
    setup(gpu_ctx, sc, rows)
    for prec in (capi.F64_STRICT, capi.F32):
        gpu_ctx.build_layers(semi, R, off, prec)
        d = gpu_ctx.dims()
        all_ = gpu_ctx.boundary_layer(2)
        assert all_.shape == (d["S"], 3, d["M"])
        for s in range(d["S"]):
            assert np.array_equal(all_[s], gpu_ctx.boundary(2, s))


def test_join_orders_main_stream_after_segment_passes(gpu_ctx):
    """hrmgpu_join: a main-stream event recorded after join covers the side-stream passes; K3 timer is readable."""
    sc, rows, semi, R, off, quat = cluttered(8)
    setup(gpu_ctx, sc, rows)
    gpu_ctx.build_layers(semi, R, off, capi.F64)
    gpu_ctx.join()
    gpu_ctx.synchronize()
    assert gpu_ctx.last_segments_ms() > 0 and gpu_ctx.last_minksweep_ms() > 0


def test_pack_segments_dev_on_own_stream_is_complete_on_return():
    """a context on its OWN stream packs into a torch tensor; the bytes must be complete when the call returns
    (torch reads the buffer on a different stream right away)."""
    import torch

    from hrm_b200 import dist as hdist

    ctx = capi.Context(0)  # no set_stream: the context's own non-blocking stream
    sc, rows, semi, R, off, quat = cluttered(30)
    setup(ctx, sc, rows, Lx=22, Ly=10)
    for _ in range(3):
        ctx.build_layers(semi, R, off, capi.F64)
        need = ctx.pack_segments_dev()
        buf = torch.empty(need, dtype=torch.uint8, device="cuda:0")
        ctx.pack_segments_dev(buf.data_ptr(), need)
        got = hdist.unpack_segments(buf.cpu().numpy(), 30 * 220)
        for a, r in zip(got, ctx.segments_all()):
            assert np.array_equal(a, r)
    ctx.close()
