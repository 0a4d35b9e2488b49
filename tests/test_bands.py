"""Row-band bookkeeping of pm_band_plan / pm_band_exchange_rows (host-only, no GPU) and a
world-size-2 gloo run of the exchange: what one rank sends is exactly what its neighbour
expects, row for row."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _eng():
    sys.path.insert(0, ROOT)
    build = importlib.import_module("ocean-perception_b200.build")
    build.build()
    return importlib.import_module("ocean-perception_b200.engine")


def _params(E, **kw):
    P = E.PatchmatchGpu.Params()
    for k, v in kw.items():
        setattr(P, k, v)
    return P


@pytest.mark.parametrize("frame_h", [2160, 720, 250, 199])
def test_band_plan_partitions_the_frame(frame_h):
    E = _eng()
    P = _params(E)
    cs = frame_h // P.sweep_chunks
    for world in (1, 2, 4, 8, 16):
        cover = []
        for r in range(world):
            lay = E.band_plan(P, frame_h, r, world)
            assert lay.nk == 16 // world and lay.k_lo == r * lay.nk
            assert lay.own_lo == lay.k_lo * cs
            assert lay.load_lo == max(lay.own_lo - 9, 0) and lay.load_hi == min(lay.own_hi + 9, frame_h)
            cover += list(range(lay.own_lo, lay.own_hi))
        assert cover == list(range(frame_h))     # remainder rows go to the last band


def test_band_plan_refusals():
    E = _eng()
    with pytest.raises(E.PmError) as ei:
        E.band_plan(_params(E), 2160, 0, 3)      # 3 does not divide 16 chunks
    assert ei.value.code == -2
    with pytest.raises(E.PmError):
        E.band_plan(_params(E, pyramid_levels=2), 2160, 0, 2)
    with pytest.raises(E.PmError):
        E.band_plan(_params(E), 100, 0, 2)       # chunks of 6 rows < 2*overlap+2
    with pytest.raises(E.PmError):
        E.band_plan(_params(E), 2160, 2, 2)


@pytest.mark.parametrize("direction", [1, -1])
@pytest.mark.parametrize("overlap", [5, 3, 0])
def test_exchange_rows_agree_between_neighbours(direction, overlap):
    E = _eng()
    P = _params(E, sweep_overlap=overlap)
    for frame_h in (2160, 250, 720):
        for world in (2, 4, 8, 16):
            lays = [E.band_plan(P, frame_h, r, world) for r in range(world)]
            rows = [E.band_exchange_rows(P, frame_h, r, world, direction) for r in range(world)]
            assert rows[0]["send_prev"] == rows[0]["recv_prev"] == (0, 0)
            assert rows[-1]["send_next"] == rows[-1]["recv_next"] == (0, 0)
            for r in range(world - 1):
                a, b = rows[r], rows[r + 1]
                assert a["send_next"] == b["recv_prev"] and b["send_prev"] == a["recv_next"]
                # a rank only sends rows it holds and receives rows inside its planes
                for k in ("send_next", "recv_next"):
                    assert lays[r].load_lo <= a[k][0] < a[k][1] <= lays[r].load_hi
                for k in ("send_prev", "recv_prev"):
                    assert lays[r + 1].load_lo <= b[k][0] < b[k][1] <= lays[r + 1].load_hi
                # buffers hold 2*overlap+3 rows
                assert max(a[k][1] - a[k][0] for k in a) <= 2 * overlap + 3
            # after the exchange every rank is current on [own_lo-ov-2, own_hi+ov+2): what it
            # computed itself plus what it received is contiguous
            for r in range(world):
                got = rows[r]
                lo = got["recv_prev"][0] if r > 0 else 0
                hi = got["recv_next"][1] if r < world - 1 else frame_h
                if r > 0:
                    assert lo == lays[r].own_lo - overlap - 2
                if r < world - 1:
                    assert hi == lays[r].own_hi + overlap + 2


def _worker(rank, world, port, frame_h, q):
    import torch
    import torch.distributed as dist
    E = _eng()
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    P = _params(E)
    ok = True
    for direction in (1, -1):
        rows = E.band_exchange_rows(P, frame_h, rank, world, direction)
        ops, recv = [], {}
        for name, peer in (("send_prev", rank - 1), ("recv_prev", rank - 1),
                           ("send_next", rank + 1), ("recv_next", rank + 1)):
            lo, hi = rows[name]
            if hi <= lo:
                continue
            if name.startswith("send"):   # a "row" is its own frame index, stamped with the sender
                t = torch.arange(lo, hi, dtype=torch.int64) * 100 + rank
                ops.append(dist.P2POp(dist.isend, t, peer))
            else:
                t = torch.empty(hi - lo, dtype=torch.int64)
                recv[name] = (t, lo, hi, peer)
                ops.append(dist.P2POp(dist.irecv, t, peer))
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        for name, (t, lo, hi, peer) in recv.items():
            ok &= bool((t == torch.arange(lo, hi, dtype=torch.int64) * 100 + peer).all())
    q.put(ok)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("frame_h", [2160, 250])
def test_exchange_world2_gloo(frame_h):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000) + frame_h % 7
    procs = [ctx.Process(target=_worker, args=(r, 2, port, frame_h, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(res)
