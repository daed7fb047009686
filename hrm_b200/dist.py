"""Multi-GPU plumbing: C-layers shard by orientation (no data-path collective while building them);
the one exchange step is an all-gather of every rank's free-segment arrays (CSR offsets + xL/xU/xM) so
that each rank holds the vertices of all layers for bridge-layer pairing.
torch.distributed (NCCL on GPUs, gloo in the CPU tests) is only the transport."""
import numpy as np


def shard_range(K, rank, world):
    """Contiguous block of orientations [k0, k1) owned by `rank` (keeps neighbouring orientations together)."""
    k0 = (K * rank) // world
    k1 = (K * (rank + 1)) // world
    return k0, k1


def pack_segments_host(off, xL, xU, xM):
    """Same byte layout hrmgpu_pack_segments_dev produces: int64 off[KL+1], then xL, xU, xM (float64)."""
    off = np.ascontiguousarray(off, dtype=np.int64)
    return np.concatenate([off.view(np.uint8)] + [np.ascontiguousarray(a, dtype=np.float64).view(np.uint8) for a in (xL, xU, xM)])


def unpack_segments(buf, n_lines):
    """Inverse of the packing for one rank's byte buffer (numpy uint8 or anything exposing .view)."""
    b = np.ascontiguousarray(buf, dtype=np.uint8)
    off = b[: 8 * (n_lines + 1)].view(np.int64).copy()
    tot = int(off[-1])
    base = 8 * (n_lines + 1)
    xL = b[base: base + 8 * tot].view(np.float64).copy()
    xU = b[base + 8 * tot: base + 16 * tot].view(np.float64).copy()
    xM = b[base + 16 * tot: base + 24 * tot].view(np.float64).copy()
    return off, xL, xU, xM


def merge_shards(shards):
    """Concatenate per-rank (off, xL, xU, xM) in rank order into one global CSR."""
    offs, Ls, Us, Ms = [np.zeros(1, dtype=np.int64)], [], [], []
    base = 0
    for off, xL, xU, xM in shards:
        offs.append(off[1:] + base)
        base += int(off[-1])
        Ls.append(xL)
        Us.append(xU)
        Ms.append(xM)
    return np.concatenate(offs), np.concatenate(Ls), np.concatenate(Us), np.concatenate(Ms)


def init_exchange(ctx, dist, device=None):
    """Creates the context's NCCL communicator (hrmgpu_comm_init): rank 0 draws the NCCL unique id and torch.distributed
    broadcasts its 128 bytes -- torch is only the bootstrap, the exchange itself is hrmgpu_allgather_layers."""
    import torch

    rank, world = dist.get_rank(), dist.get_world_size()
    use_dev = device is not None and dist.get_backend() == "nccl"
    ident = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        ident = torch.from_numpy(ctx.comm_unique_id().copy())
    if use_dev:
        ident = ident.to(device)
    dist.broadcast(ident, src=0)
    ctx.comm_init(world, rank, ident.cpu().numpy())


def gathered_csr(ctx):
    """Global CSR (off, xL, xU, xM) over the layers of all ranks, in rank order, from the last hrmgpu_allgather_layers."""
    info = ctx.gathered_info()
    return merge_shards([ctx.gathered(r) for r in range(info["world"])])


def allgather_bytes(local_u8, dist, device):
    """All-gather of variable-length byte tensors through torch.distributed: sizes first, then payload padded to the max.
    (Exact-size variant for host frameworks that want torch tensors; the hot path uses hrmgpu_allgather_layers.)"""
    import torch

    world = dist.get_world_size()
    n = torch.tensor([local_u8.numel()], dtype=torch.int64, device=device)
    sizes = torch.zeros(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(sizes, n)
    sizes_h = [int(v) for v in sizes.cpu()]
    maxb = max(sizes_h)
    pad = torch.zeros(maxb, dtype=torch.uint8, device=device)
    pad[: local_u8.numel()] = local_u8
    out = torch.empty(world * maxb, dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(out, pad)
    return [out[r * maxb: r * maxb + sizes_h[r]] for r in range(world)]


def allgather_segments(ctx, dist, device):
    """Packs the context's last build on the device and exchanges it with torch.distributed; returns per-rank uint8 device
    tensors.  hrmgpu_pack_segments_dev copies on the context's main stream: unless that IS torch's current stream (installed
    with ctx.set_stream), the context is synchronised before torch touches the buffer."""
    import torch

    need = ctx.pack_segments_dev()
    buf = torch.empty(need, dtype=torch.uint8, device=device)
    ctx.pack_segments_dev(buf.data_ptr(), need)
    ctx.synchronize()
    return allgather_bytes(buf, dist, device)


# ---- bridge stage over the gathered layers: pairs re-sharded by pair index -------------------------------
def nearest_earlier_layers(quats):
    """connectMultiSlice's layer pairing (reference src/planners/HRM3D.cpp:160-183): layer i >= 1 is bridged to the EARLIER layer
    of the smallest angular distance; returns [(near, cur)] for cur = 1 .. K-1."""
    from hrm_b200 import tfe

    pairs = []
    for i in range(1, len(quats)):
        best, near = float("inf"), 0
        for j in range(i):
            d = tfe.angular_distance(quats[i], quats[j])
            if d < best:
                best, near = d, j
        pairs.append((near, i))
    return pairs


def pair_shard(n_pairs, rank, world):
    """Pair indices owned by `rank`: round robin, so that consecutive (similarly sized) pairs spread over the ranks."""
    return list(range(rank, n_pairs, world))


def layer_vertices(csr, lim, Lx, Ly, K):
    """Vertex positions (sweep-line x, y and segment mid point) of all layers of a (gathered) CSR in the reference's order
    (HRM3D.cpp:93-113) and the first vertex id of every layer."""
    off, _, _, xM = csr
    L = Lx * Ly
    dx, dy = (lim[1] - lim[0]) / float(Lx - 1), (lim[3] - lim[2]) / float(Ly - 1)
    tx = np.array([lim[0] + float(i) * dx for i in range(Lx)])
    ty = np.array([lim[2] + float(i) * dy for i in range(Ly)])
    cnt = np.diff(np.asarray(off, dtype=np.int64))
    line = np.repeat(np.arange(K * L) % L, cnt)
    xyz = np.stack([tx[line // Ly], ty[line % Ly], np.asarray(xM, dtype=np.float64)], axis=1)
    first = np.concatenate([[0], np.cumsum(cnt.reshape(K, L).sum(axis=1))]).astype(np.int64)
    return xyz, first


def bridge_matches(ctx, xyz, first, pairs, body_semi, link_tf, quats, lim, Lx, Ly, num_point, precision):
    """K5 for the layer pairs `pairs` [(near, cur)] on this context: tightly-fitted ellipsoids + bridge layers in one
    hrmgpu_build_bridge_tfe call, candidate enumeration + transition tests + the replay of the reference's moving lower bound in one
    hrmgpu_connect_pairs3d call.  link_tf = [(R 3x3, t 3)] of the bodies after the base.  Returns one int array per pair:
    match[m0 - first[near]] = vertex id in `cur` that vertex m0 of `near` connects to, or -1."""
    from hrm_b200 import hostmath as hm
    from hrm_b200 import tfe

    if not pairs:
        return []
    B = 1 + len(link_tf)
    T = int(num_point) + 1
    qa, qb = np.zeros((len(pairs), B, 4)), np.zeros((len(pairs), B, 4))
    link_off = np.zeros((len(pairs), T, B, 3))
    for p, (near, cur) in enumerate(pairs):
        for b in range(B):  # computeTFE (HRM3D.cpp:521-539): body b turns with the base
            if b == 0:
                qa[p, b], qb[p, b] = quats[cur], quats[near]
            else:
                Rl = link_tf[b - 1][0]
                qa[p, b] = hm.rot_to_quat(hm.mat_mul(hm.quat_to_rot(quats[cur]), Rl))
                qb[p, b] = hm.rot_to_quat(hm.mat_mul(hm.quat_to_rot(quats[near]), Rl))
        # rotated link translations of the interpolated poses (the same for every candidate of the pair)
        steps = tfe.interpolate_se3([0.0, 0.0, 0.0] + list(quats[near]), [0.0, 0.0, 0.0] + list(quats[cur]), int(num_point))
        for s, v in enumerate(steps):
            Rs = hm.quat_to_rot(v[3:7])
            for b in range(1, B):
                link_off[p, s, b] = hm.mat_vec(Rs, link_tf[b - 1][1])
    ctx.build_bridge_tfe(body_semi, qa, qb, int(num_point), precision)
    w1 = np.array([float(s) * (1.0 / float(num_point)) for s in range(T)])
    w0 = np.array([1.0 - float(s) * (1.0 / float(num_point)) for s in range(T)])
    thx, thy = 2.0 * (lim[1] - lim[0]) / float(Lx), 2.0 * (lim[3] - lim[2]) / float(Ly)
    range4 = [[first[near], first[near + 1], first[cur], first[cur + 1]] for near, cur in pairs]
    match, _ = ctx.connect_pairs3d(xyz, range4, thx, thy, w0, w1, link_off)
    out, at = [], 0
    for near, _cur in pairs:
        n = int(first[near + 1] - first[near])
        out.append(np.array(match[at:at + n], dtype=np.int64))
        at += n
    return out


def bridge_matches_sharded(ctx, dist, csr, K, body_semi, link_tf, quats, lim, Lx, Ly, num_point, precision):
    """The bridge stage of all K layers with the pairs re-sharded by pair index: every rank holds all vertices (gathered CSR),
    tests only its pairs -- any rank can build any bridge layer, it needs the obstacles and the two orientations only -- and the
    matches (a few KB) are exchanged as objects.  Returns (pairs, matches) for ALL pairs, identical on every rank."""
    rank, world = dist.get_rank(), dist.get_world_size()
    xyz, first = layer_vertices(csr, lim, Lx, Ly, K)
    pairs = nearest_earlier_layers(quats)
    mine = pair_shard(len(pairs), rank, world)
    local = bridge_matches(ctx, xyz, first, [pairs[p] for p in mine], body_semi, link_tf, quats, lim, Lx, Ly, num_point, precision)
    gathered = [None] * world
    dist.all_gather_object(gathered, (mine, local))
    matches = [None] * len(pairs)
    for ids, ms in gathered:
        for p, m in zip(ids, ms):
            matches[p] = m
    return pairs, matches
