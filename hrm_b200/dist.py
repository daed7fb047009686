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
