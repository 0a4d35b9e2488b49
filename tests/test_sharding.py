"""Host-side multi-GPU bookkeeping on CPU: contiguous pair shards, row bands, and a
world-size-2 gloo run of the sharded driver (the data path itself has no collective)."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _sh():
    sys.path.insert(0, ROOT)
    return importlib.import_module("ocean-perception_b200.sharding")


def test_shard_range_partitions_exactly():
    sh = _sh()
    for n in (0, 1, 7, 64, 512, 513):
        for world in (1, 2, 3, 4, 8):
            cover = []
            for r in range(world):
                lo, hi = sh.shard_range(n, r, world)
                assert 0 <= lo <= hi <= n
                cover += list(range(lo, hi))
            assert cover == list(range(n))
    assert [sh.shard_range(512, r, 8) for r in range(8)][3] == (192, 256)   # 64 pairs per GPU
    with pytest.raises(ValueError):
        sh.shard_range(4, 2, 2)


def test_band_rows_and_neighbours(built_lib):
    """sharding.band_rows / halo_exchanges are the engine's own plan (pm_band_plan,
    pm_band_exchange_rows): config C5, 2160 rows in bands of 1080 / 540 / 270 = whole column-sweep
    chunks of 135 rows, halo = overlap + 4 rows, two exchanges per iteration."""
    sh = _sh()
    P = importlib.import_module("ocean-perception_b200").PatchmatchGpu.Params()
    P.init_mode = "random"
    ov = P.sweep_overlap
    for world, rows in ((2, 1080), (4, 540), (8, 270)):
        owned = 0
        for r in range(world):
            lo, hi, llo, lhi = sh.band_rows(P, 2160, r, world)
            assert hi - lo == rows and lo % 135 == 0
            assert llo == max(lo - (ov + 4), 0) and lhi == min(hi + ov + 4, 2160)
            owned += hi - lo
            for d in (+1, -1):
                ex = sh.halo_exchanges(P, 2160, r, world, d)
                peers = {p for p, _, _ in ex}
                assert ((r - 1) in peers) == (r > 0) and ((r + 1) in peers) == (r < world - 1)
                for peer, kind, (a, b) in ex:
                    assert 0 < b - a <= 2 * ov + 3
                    # what I send is what the peer receives
                    back = sh.halo_exchanges(P, 2160, peer, world, d)
                    assert (r, "recv" if kind == "send" else "send", (a, b)) in back
        assert owned == 2160
    with pytest.raises(Exception):
        sh.band_rows(P, 2160, 0, 3)     # 3 does not divide sweep_chunks = 16


def _fake_match(L, R, first):
    # stands in for the engine: the result depends on the data and on the global pair index
    idx = np.arange(first, first + L.shape[0], dtype=np.float32)[:, None, None]
    return L.astype(np.float32) + idx, R.astype(np.float32) - idx


def _worker(rank, world, port, n, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sh = importlib.import_module("ocean-perception_b200.sharding")
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    L = rng.integers(0, 255, (n, 6, 8)).astype(np.uint8)
    R = rng.integers(0, 255, (n, 6, 8)).astype(np.uint8)
    dl, dr = sh.run_sharded(_fake_match, L, R, group=dist.group.WORLD)
    if rank == 0:
        wl, wr = _fake_match(L, R, 0)
        q.put(bool(np.array_equal(dl, wl) and np.array_equal(dr, wr)))
    else:
        lo, hi = sh.shard_range(n, rank, world)
        q.put(dl.shape[0] == hi - lo)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [5, 8])
def test_sharded_run_world2_gloo(n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + n
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(res)
