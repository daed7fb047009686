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
