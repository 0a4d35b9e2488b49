"""Multi-GPU decomposition of the path: one process per GPU.

Independent stereo pairs are split into contiguous ranges, one per rank, with no
data-path collective (the reference is single-GPU; pairs never interact). A single very
large frame is split into row bands whose boundary rows of disparity are exchanged once
per propagation iteration. Only the host-side bookkeeping lives here; it is covered on CPU
with a world-size-2 gloo group (tests/test_sharding.py)."""


def shard_range(n, rank, world):
    """Contiguous split pairs[g*n/G : (g+1)*n/G] (512 -> 512/256/128/64 per GPU)."""
    if world < 1 or not (0 <= rank < world) or n < 0:
        raise ValueError("bad shard request n=%d rank=%d world=%d" % (n, rank, world))
    return (rank * n) // world, ((rank + 1) * n) // world


def band_rows(height, rank, world, halo):
    """Row band of one rank for a frame split across `world` GPUs.

    Returns (own_lo, own_hi, load_lo, load_hi): rows this rank owns, and the rows it loads
    (its band plus `halo` rows of its neighbours, clipped to the image)."""
    lo, hi = shard_range(height, rank, world)
    return lo, hi, max(lo - halo, 0), min(hi + halo, height)


def halo_exchanges(rank, world):
    """Neighbour ranks a band swaps boundary rows with: [(peer, 'up'|'down'), ...]."""
    out = []
    if rank > 0:
        out.append((rank - 1, "up"))
    if rank < world - 1:
        out.append((rank + 1, "down"))
    return out


def run_sharded(match_fn, left, right, group=None, gather=True):
    """Runs match_fn(left[lo:hi], right[lo:hi], first_pair_index=lo) on this rank's range.

    `group` is a torch.distributed process group (None = single process). With gather=True
    rank 0 receives every rank's (disp_l, disp_r) and returns the full arrays; other ranks
    return their own slice. The gather is bookkeeping for tests/tools, not part of the path."""
    import numpy as np
    n = left.shape[0]
    if group is None:
        return match_fn(left, right, 0)
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_range(n, rank, world)
    dl, dr = match_fn(left[lo:hi], right[lo:hi], lo)
    if not gather:
        return dl, dr
    parts = [None] * world
    dist.gather_object((lo, hi, dl, dr), parts if rank == 0 else None, dst=0, group=group)
    if rank != 0:
        return dl, dr
    full_l = np.empty((n,) + dl.shape[1:], dl.dtype)
    full_r = np.empty((n,) + dr.shape[1:], dr.dtype)
    for plo, phi, pl, pr in parts:
        full_l[plo:phi], full_r[plo:phi] = pl, pr
    return full_l, full_r
