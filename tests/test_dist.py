"""Multi-GPU host logic on CPU: orientation sharding and the all-gather of packed free-segment arrays, with a
world_size-2 gloo process group (the GPU path uses the same code over NCCL)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hrm_b200 import dist as hdist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_covers_everything():
    for K in (1, 7, 60, 10000):
        for world in (1, 2, 4, 8):
            spans = [hdist.shard_range(K, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == K
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


def test_pack_unpack_roundtrip():
    rng = np.random.default_rng(0)
    cnt = rng.integers(0, 4, size=12)
    off = np.concatenate([[0], np.cumsum(cnt)])
    xL, xU, xM = rng.normal(size=off[-1]), rng.normal(size=off[-1]), rng.normal(size=off[-1])
    buf = hdist.pack_segments_host(off, xL, xU, xM)
    o2, a, b, c = hdist.unpack_segments(buf, 12)
    assert np.array_equal(o2, off) and np.array_equal(a, xL) and np.array_equal(b, xU) and np.array_equal(c, xM)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import hrmo
    from tests import util

    sc = util.scenes()["3D_superquadrics_cluttered_rabbit"]
    rows = np.array(sc["arena"] + sc["obstacle"])
    quats = sc["quats"][:6]
    k0, k1 = hdist.shard_range(len(quats), rank, world)
    semi, R, off, quat = util.body_poses3d(sc["robot"], quats[k0:k1])
    L = sc["Lx"] * sc["Ly"]
    offs, Ls, Us, Ms = [np.zeros(1, dtype=np.int64)], [], [], []
    base = 0
    for k in range(k1 - k0):  # the CPU oracle stands in for the GPU build in this host-logic test
        r = hrmo.layer3d(1, 7, rows, 10, semi, quat[k], off[k], sc["lim"], sc["Lx"], sc["Ly"], want_bound=False)
        offs.append(r["off"][1:].astype(np.int64) + base)
        base += int(r["off"][-1])
        Ls.append(r["xL"]), Us.append(r["xU"]), Ms.append(r["xM"])
    local = hdist.pack_segments_host(np.concatenate(offs), np.concatenate(Ls), np.concatenate(Us), np.concatenate(Ms))
    parts = hdist.allgather_bytes(torch.from_numpy(local.copy()), dist, torch.device("cpu"))
    shards = []
    for r in range(world):
        a, b = hdist.shard_range(len(quats), r, world)
        shards.append(hdist.unpack_segments(parts[r].numpy(), (b - a) * L))
    merged = hdist.merge_shards(shards)
    q.put((rank, merged[0].tolist(), merged[3].tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_allgather_segments_world2_gloo(oracle):
    from tests import util

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # both ranks hold the same merged CSR, equal to the single-process result
    assert res[0][1] == res[1][1] and res[0][2] == res[1][2]
    sc = util.scenes()["3D_superquadrics_cluttered_rabbit"]
    rows = np.array(sc["arena"] + sc["obstacle"])
    semi, R, off, quat = util.body_poses3d(sc["robot"], sc["quats"][:6])
    xm = []
    for k in range(6):
        xm += oracle.layer3d(1, 7, rows, 10, semi, quat[k], off[k], sc["lim"], sc["Lx"], sc["Ly"], want_bound=False)["xM"].tolist()
    assert res[0][2] == xm


def test_bridge_pair_sharding_helpers():
    """Pure host logic of the pair-sharded bridge stage: every pair has exactly one owner, the pairing is the
    planner's nearest EARLIER layer, vertex ids follow (layer, line, segment) order."""
    from hrm_b200 import dist as hd
    from hrm_b200 import tfe
    from tests import util

    for n in (0, 1, 7, 59):
        for w in (1, 2, 3, 8):
            assert sorted(sum([hd.pair_shard(n, r, w) for r in range(w)], [])) == list(range(n))
    quats = util.scenes()["3D_superquadrics_cluttered_rabbit"]["quats"][:12]
    pairs = hd.nearest_earlier_layers(quats)
    assert [c for _, c in pairs] == list(range(1, 12))
    for near, cur in pairs:
        assert near < cur
        d = [tfe.angular_distance(quats[cur], quats[j]) for j in range(cur)]
        assert near == int(np.argmin(d))  # (first minimum: the reference keeps the earlier layer on ties)
    # two layers, 2 x 3 lines
    off = np.array([0, 1, 1, 3, 3, 3, 4, 4, 6, 6, 6, 6, 7])
    xM = np.arange(7, dtype=np.float64)
    xyz, first = hd.layer_vertices((off, xM, xM, xM), [0.0, 1.0, 0.0, 2.0, -1.0, 1.0], 2, 3, 2)
    assert first.tolist() == [0, 4, 7]
    assert xyz[:, 2].tolist() == xM.tolist()
    assert xyz[:4, :2].tolist() == [[0.0, 0.0], [0.0, 2.0], [0.0, 2.0], [1.0, 2.0]]  # lines 0, 2, 2, 5 of layer 0
    assert xyz[4:, :2].tolist() == [[0.0, 1.0], [0.0, 1.0], [1.0, 2.0]]              # lines 1, 1, 5 of layer 1
