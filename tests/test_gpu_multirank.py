"""Multi-GPU exchange on real GPUs (skips on a 1-GPU box; run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multirank.py -m gpu`):
every rank builds its shard of the C-layers, ONE hrmgpu_allgather_layers exchanges the packed free segments over NCCL, and
every rank's merged CSR must be bit-identical to a single-GPU build of all layers."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, K):
    import torch
    import torch.distributed as dist

    from hrm_b200 import capi
    from hrm_b200 import dist as hdist
    from tests import util

    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    sc = util.scenes()["3D_superquadrics_cluttered_rabbit"]
    rows = sc["arena"] + sc["obstacle"]
    semi, R, off, _ = util.body_poses3d(sc["robot"], sc["quats"][:K])
    Lx, Ly = 12, 8
    ctx = capi.Context(rank)
    try:
        ctx.set_scene3d(1, len(sc["obstacle"]), util.sq17(rows, 10), 10)
        ctx.set_sweep(sc["lim"], Lx, Ly)
        hdist.init_exchange(ctx, dist)
        k0, k1 = hdist.shard_range(K, rank, world)
        line_cap = ((K + world - 1) // world) * Lx * Ly
        for prec in (capi.F64_STRICT, capi.F64):
            ctx.build_layers(semi, R[k0:k1], off[k0:k1], prec)
            ctx.allgather_layers(line_cap, 1 << 16)
            merged = hdist.gathered_csr(ctx)
            info = ctx.gathered_info()
            assert info["world"] == world and info["rank"] == rank
            # the same layers on ONE GPU
            ctx.build_layers(semi, R, off, prec)
            single = ctx.segments_all()
            assert single[0][-1] > 100
            for a, b in zip(merged, single):
                assert np.array_equal(a, b)
        # exchange capacity too small: every receiver can tell
        ctx.build_layers(semi, R[k0:k1], off[k0:k1], capi.F64)
        ctx.allgather_layers(line_cap, 4)
        with pytest.raises(capi.HrmGpuError) as e:
            ctx.gathered(0)
        assert e.value.code == capi.E_CAP
        # pipelined: exchange of batch A overlaps the build of batch B (no host sync in between)
        ctx.build_layers(semi, R[k0:k1], off[k0:k1], capi.F64)
        ctx.allgather_layers(line_cap, 1 << 16)
        ctx.build_layers(semi, R[k0:k1][::-1].copy(), off[k0:k1][::-1].copy(), capi.F64)
        ctx.synchronize()
        again = hdist.gathered_csr(ctx)
        for a, b in zip(again, merged):
            assert np.array_equal(a, b)
    finally:
        ctx.close()
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.parametrize("K", [10, 9])
def test_allgather_layers_equals_single_gpu_build(K):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    world = min(torch.cuda.device_count(), 4)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    mp.spawn(_worker, args=(world, _free_port(), K), nprocs=world, join=True)


def _bridge_worker(rank, world, port, K, n_gpus):
    """Layers sharded by orientation, segments exchanged, bridge pairs re-sharded by pair index: the matches of
    all pairs assembled from the ranks must equal the matches one context computes for all pairs."""
    import torch
    import torch.distributed as dist

    from hrm_b200 import capi
    from hrm_b200 import dist as hdist
    from tests import util

    dev = rank if n_gpus >= world else 0  # (a 1-GPU box: both ranks share the device, the segments go through the host)
    torch.cuda.set_device(dev)
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    sc = util.scenes()["3D_superquadrics_cluttered_rabbit"]
    rows = sc["arena"] + sc["obstacle"]
    quats = sc["quats"][:K]
    semi, R, off, _ = util.body_poses3d(sc["robot"], quats)
    rb = util.robot3d(sc["robot"])
    body_semi = [rb.base.semiAxis] + [l.semiAxis for l in rb.link]
    Lx, Ly, num_point = 10, 6, 5
    ctx = capi.Context(dev)
    try:
        ctx.set_scene3d(1, len(sc["obstacle"]), util.sq17(rows, 10), 10)
        ctx.set_sweep(sc["lim"], Lx, Ly)
        k0, k1 = hdist.shard_range(K, rank, world)
        ctx.build_layers(semi, R[k0:k1], off[k0:k1], capi.F64_STRICT)
        if n_gpus >= world:
            hdist.init_exchange(ctx, dist)
            ctx.allgather_layers(((K + world - 1) // world) * Lx * Ly, 1 << 16)
            csr = hdist.gathered_csr(ctx)
        else:
            shards = [None] * world
            dist.all_gather_object(shards, tuple(ctx.segments_all()))
            csr = hdist.merge_shards(shards)
        pairs, got = hdist.bridge_matches_sharded(ctx, dist, csr, K, body_semi, rb.tf, quats, sc["lim"], Lx, Ly, num_point, capi.F64_STRICT)
        assert len(pairs) == K - 1 and all(m is not None for m in got)
        # one context, all layers, all pairs
        ctx.build_layers(semi, R, off, capi.F64_STRICT)
        single = ctx.segments_all()
        for a, b in zip(csr, single):
            assert np.array_equal(a, b)
        xyz, first = hdist.layer_vertices(single, sc["lim"], Lx, Ly, K)
        want = hdist.bridge_matches(ctx, xyz, first, pairs, body_semi, rb.tf, quats, sc["lim"], Lx, Ly, num_point, capi.F64_STRICT)
        for p in range(len(pairs)):
            assert np.array_equal(got[p], want[p]), pairs[p]
        linked = sum(int((m >= 0).sum()) for m in want)
        assert linked > 20 and any((m < 0).any() for m in want)
        for (near, cur), m in zip(pairs, want):  # matches point into the `cur` layer of their pair
            assert ((m < 0) | ((m >= first[cur]) & (m < first[cur + 1]))).all()
    finally:
        ctx.close()
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.parametrize("K", [9])
def test_bridge_pairs_sharded_by_pair_index(K):
    """Runs on any box: with >= 2 GPUs one rank per GPU and hrmgpu_allgather_layers over NCCL, with one GPU two ranks on the same
    device and the segments exchanged through the host."""
    import torch
    import torch.multiprocessing as mp

    n_gpus = torch.cuda.device_count()
    world = 2
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    mp.spawn(_bridge_worker, args=(world, _free_port(), K, n_gpus), nprocs=world, join=True)
