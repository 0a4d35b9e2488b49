"""Multi-GPU decomposition of the path: one process per GPU.

Independent stereo pairs are split into contiguous ranges, one per rank, with no
data-path collective (the reference is single-GPU; pairs never interact). A single very
large frame is split into row bands of whole column-sweep chunks whose overlap rows of
{disparity, cost} are swapped with the neighbour bands after EVERY column sweep, i.e. twice
per propagation iteration (pm_band_plan / pm_band_exchange_rows in include/pm_b200.h; bands.py
drives the exchange). Only host-side bookkeeping lives here; it is covered on CPU with a
world-size-2 gloo group (tests/test_sharding.py, tests/test_bands.py)."""


def shard_range(n, rank, world):
    """Contiguous split pairs[g*n/G : (g+1)*n/G] (512 -> 512/256/128/64 per GPU)."""
    if world < 1 or not (0 <= rank < world) or n < 0:
        raise ValueError("bad shard request n=%d rank=%d world=%d" % (n, rank, world))
    return (rank * n) // world, ((rank + 1) * n) // world


def band_rows(params, height, rank, world):
    """Row band of one rank: a thin wrapper over pm_band_plan (the plan the engine runs).

    Returns (own_lo, own_hi, load_lo, load_hi): the frame rows this rank produces, and the rows it
    must be given (its band plus overlap + 4 halo rows, clipped to the frame)."""
    from .engine import band_plan
    lay = band_plan(params, height, rank, world)
    return lay.own_lo, lay.own_hi, lay.load_lo, lay.load_hi


def halo_exchanges(params, height, rank, world, direction):
    """The frame-row intervals a band swaps after a column sweep of `direction` (+1 / -1):
    [(peer, 'send'|'recv', (lo, hi)), ...] from pm_band_exchange_rows; empty intervals dropped."""
    from .engine import band_exchange_rows
    rows = band_exchange_rows(params, height, rank, world, direction)
    out = []
    for name, peer in (("send_prev", rank - 1), ("recv_prev", rank - 1), ("send_next", rank + 1),
                       ("recv_next", rank + 1)):
        lo, hi = rows[name]
        if hi > lo:
            out.append((peer, name.split("_")[0], (lo, hi)))
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
