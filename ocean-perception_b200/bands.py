"""One very large frame split into row bands, one band per GPU.

Host side of include/pm_b200.h's pm_band_* calls: one process per GPU, the engine packs the
rows a column sweep leaves for its neighbours, and this module moves them with
torch.distributed point-to-point ops (NCCL send/recv over NVLink). PyTorch is plumbing
here: device buffers, the stream everything is ordered on, and the process group.

A band is a whole number of the reference's column-sweep chunks (patchmatch_gpu.cu:196-202
in the reference), so the banded result is bit-identical to the single-GPU one.
"""
import numpy as np

from .engine import PatchmatchGpu, band_plan


class _DevBuf:
    """A raw device pointer as a __cuda_array_interface__ object (torch.as_tensor takes it)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1",
                                         "data": (int(ptr), False), "version": 3}


def _as_tensor(ptr, nbytes, device):
    import torch
    return torch.as_tensor(_DevBuf(ptr, nbytes), device=device)


def _xfer_list(x):
    """[(kind, peer_offset, ptr, bytes)] of a CBandXfer; peer_offset is -1 (prev) / +1 (next)."""
    return [("send", -1, x.send_prev, x.send_prev_bytes), ("recv", -1, x.recv_prev, x.recv_prev_bytes),
            ("send", +1, x.send_next, x.send_next_bytes), ("recv", +1, x.recv_next, x.recv_next_bytes)]


class BandedMatcher:
    """PatchmatchGpu::Match for one frame spread over the ranks of a process group.

    Every rank calls Match with the WHOLE frame (or only its rows, see rows=) and gets the
    disparity rows it owns: (disp_l, disp_r, (own_lo, own_hi))."""

    def __init__(self, params, group=None, device=0, p2p=True):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.group = group
        self.rank = dist.get_rank(group) if group is not None or dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if group is not None or dist.is_initialized() else 1
        self.device = torch.device("cuda", device)
        self.params = params
        self.eng = PatchmatchGpu(params, device=device)
        # everything (kernels, packing, NCCL) is ordered on ONE non-default stream: handle 0
        # (torch's default stream) means "the engine's own stream" in the C ABI
        self.stream = torch.cuda.Stream(self.device)
        self.exchanges = 0
        self.exchange_bytes = 0
        self.overlap = "none: each exchange runs on the kernels' stream between two sweeps"
        self._last_x = None
        self._per_frame = 0
        self.want_p2p = bool(p2p)
        self.p2p = False          # the halo rows go over peer memory (pm_band_p2p_*), not NCCL
        self._p2p_width = None

    def close(self):
        self.eng.close()

    def layout(self, frame_h):
        return band_plan(self.params, frame_h, self.rank, self.world)

    def setup_p2p(self, width):
        """Peer-memory exchange: every rank exports its receive region as a CUDA IPC handle, the handles
        are all-gathered over the process group (host side), every rank maps its neighbours' regions.
        Collective over the group. Falls back to the NCCL transport when IPC is not available."""
        dist = self.dist
        if self.world == 1 or not self.want_p2p or self._p2p_width == width:
            return self.p2p
        ok, handle = 1, b""
        try:
            handle, _ = self.eng.band_p2p_export(width)
        except Exception:
            ok = 0
        handles = [None] * self.world
        dist.all_gather_object(handles, (ok, handle), group=self.group)
        if all(o for o, _ in handles):
            try:
                self.eng.band_p2p_connect(handles[self.rank - 1][1] if self.rank > 0 else None,
                                          handles[self.rank + 1][1] if self.rank < self.world - 1 else None)
            except Exception:
                ok = 0
        else:
            ok = 0
        flags = [None] * self.world
        dist.all_gather_object(flags, ok, group=self.group)
        self.p2p = all(flags)
        if not self.p2p:
            self.eng.band_p2p_disable()
        else:
            self.overlap = ("peer memory: the pack kernel stores the halo rows into the neighbour's buffer over "
                            "NVLink and publishes a sequence flag; no NCCL call, no staging copy")
        self._p2p_width = width
        return self.p2p

    def upload(self, iml, imr, seed_l=None, seed_r=None):
        """Copies this rank's rows of the frame to the device; returns the resident band."""
        torch = self.torch
        h, w = iml.shape
        self.setup_p2p(w)
        lay = self.layout(h)
        sl = slice(lay.load_lo, lay.load_hi)
        dev = {"h": h, "w": w, "lay": lay,
               "L": torch.from_numpy(np.ascontiguousarray(iml[sl])).to(self.device),
               "R": torch.from_numpy(np.ascontiguousarray(imr[sl])).to(self.device),
               "SL": None, "SR": None}
        if seed_l is not None:
            dev["SL"] = torch.from_numpy(np.ascontiguousarray(seed_l[sl], np.float32)).to(self.device)
            dev["SR"] = torch.from_numpy(np.ascontiguousarray(seed_r[sl], np.float32)).to(self.device)
        own = lay.own_hi - lay.own_lo
        dev["OL"] = torch.empty((own, w), dtype=torch.float32, device=self.device)
        dev["OR"] = torch.empty((own, w), dtype=torch.float32, device=self.device)
        return dev

    def _exchange(self, x):
        dist = self.dist
        ops, keep = [], []
        for kind, off, ptr, nbytes in _xfer_list(x):
            if not nbytes:
                continue
            t = _as_tensor(ptr, nbytes, self.device)
            keep.append(t)
            peer = self.rank + off
            if self.group is not None:
                peer = dist.get_global_rank(self.group, peer)
            ops.append(dist.P2POp(dist.isend if kind == "send" else dist.irecv, t, peer, self.group))
            self.exchange_bytes += nbytes if kind == "send" else 0
        if ops:
            for wk in dist.batch_isend_irecv(ops):
                wk.wait()  # stream-ordered for NCCL: the current stream waits, the host does not
        self.exchanges += 1
        self._last_x = x

    def measure_exchange_ms(self, dev, reps=5):
        """Device time of one frame's halo exchanges ALONE (no kernels between them): the buffers of
        the last exchange moved exchanges-per-frame times, CUDA events on the band stream, averaged
        over `reps` frames. Every rank must call it (the transfers are collective between neighbours)."""
        torch = self.torch
        if self.p2p:
            return None   # inside the engine: see the band_exchange stage time
        if self.world == 1 or self._last_x is None or not self._per_frame:
            return 0.0
        saved = (self.exchanges, self.exchange_bytes)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(self.stream):
            self._exchange(self._last_x)          # warm-up
            ev0.record(self.stream)
            for _ in range(reps * self._per_frame):
                self._exchange(self._last_x)
            ev1.record(self.stream)
        self.stream.synchronize()
        self.exchanges, self.exchange_bytes = saved
        return ev0.elapsed_time(ev1) / reps

    def run(self, dev, pair_index=0):
        """The device part, asynchronous: ordered after torch's current stream on entry, and
        the current stream waits for it on exit."""
        torch = self.torch
        cur = torch.cuda.current_stream(self.device)
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            out = self._run(dev, pair_index)
        cur.wait_stream(self.stream)
        return out

    def _run(self, dev, pair_index):
        st = self.stream.cuda_stream
        lay, w = dev["lay"], dev["w"]
        self.eng.band_begin(dev["L"].data_ptr(), dev["R"].data_ptr(), w, w, dev["h"], self.rank,
                            self.world, dev["SL"].data_ptr() if dev["SL"] is not None else None,
                            dev["SR"].data_ptr() if dev["SR"] is not None else None, w * 4,
                            pair_index, st)
        n = 0
        while True:
            x = self.eng.band_step()
            if x is None:
                break
            self._exchange(x)
            n += 1
        if self.p2p and self.world > 1:   # the engine ran the exchanges itself (two per iteration)
            n = 2 * self.params.patchmatch_iters
            self.exchanges += n
        self._per_frame = n
        self.eng.band_finish(dev["OL"].data_ptr(), dev["OR"].data_ptr(), w * 4)
        return dev["OL"], dev["OR"]

    def Match(self, iml, imr, seed_l=None, seed_r=None, pair_index=0):
        dev = self.upload(iml, imr, seed_l, seed_r)
        ol, orr = self.run(dev, pair_index)
        lay = dev["lay"]
        return ol.cpu().numpy(), orr.cpu().numpy(), (lay.own_lo, lay.own_hi)


def match_bands_one_device_p2p(params, iml, imr, world, device=0, pair_index=0):
    """All `world` bands of a frame on ONE GPU with the PEER-MEMORY exchange (pm_band_p2p_*): one
    engine and one stream per band, the neighbours' receive regions connected by pointer. The bands
    run concurrently; each waits on the flags its neighbours publish. Random init only."""
    import torch
    dev = torch.device("cuda", device)
    h, w = iml.shape
    engs = [PatchmatchGpu(params, device=device) for _ in range(world)]
    streams = [torch.cuda.Stream(dev) for _ in range(world)]
    try:
        regions = [e.band_p2p_export(w)[1] for e in engs]
        for r, e in enumerate(engs):
            e.band_p2p_connect(region_prev=regions[r - 1] if r > 0 else None,
                               region_next=regions[r + 1] if r < world - 1 else None)
        res = []
        for r, e in enumerate(engs):
            lay = band_plan(params, h, r, world)
            sl = slice(lay.load_lo, lay.load_hi)
            d = {"lay": lay, "L": torch.from_numpy(np.ascontiguousarray(iml[sl])).to(dev),
                 "R": torch.from_numpy(np.ascontiguousarray(imr[sl])).to(dev)}
            own = lay.own_hi - lay.own_lo
            d["OL"] = torch.empty((own, w), dtype=torch.float32, device=dev)
            d["OR"] = torch.empty((own, w), dtype=torch.float32, device=dev)
            res.append(d)
        torch.cuda.synchronize(dev)
        # every allocation first (it synchronises the device), then every band's whole schedule
        for r, (e, d) in enumerate(zip(engs, res)):
            e.band_begin(d["L"].data_ptr(), d["R"].data_ptr(), w, w, h, r, world, None, None, w * 4,
                         pair_index, streams[r].cuda_stream)
        for e in engs:
            assert e.band_step() is None
        for e, d in zip(engs, res):
            e.band_finish(d["OL"].data_ptr(), d["OR"].data_ptr(), w * 4)
        for r, e in enumerate(engs):
            e.synchronize(streams[r].cuda_stream)
        dl = np.concatenate([d["OL"].cpu().numpy() for d in res])
        dr = np.concatenate([d["OR"].cpu().numpy() for d in res])
        return dl, dr
    finally:
        for e in engs:
            e.close()


def match_bands_one_device(params, iml, imr, world, seed_l=None, seed_r=None, device=0, pair_index=0):
    """All `world` bands of a frame on ONE GPU, in lock step, in one process: the ranks'
    exchange buffers are copied device-to-device where NCCL would move them. This is how
    the multi-rank path is exercised when fewer GPUs than ranks are at hand."""
    import torch
    dev = torch.device("cuda", device)
    h, w = iml.shape
    engs = [PatchmatchGpu(params, device=device) for _ in range(world)]
    stream = torch.cuda.Stream(dev)   # one non-default stream orders kernels and copies
    stream.wait_stream(torch.cuda.current_stream(dev))
    try:
        with torch.cuda.stream(stream):
            return _bands_one_device(params, iml, imr, world, seed_l, seed_r, dev, pair_index, engs,
                                     stream.cuda_stream)
    finally:
        for e in engs:
            e.close()


def _bands_one_device(params, iml, imr, world, seed_l, seed_r, dev, pair_index, engs, st):
    import torch
    h, w = iml.shape
    res = []
    for r, e in enumerate(engs):
        lay = band_plan(params, h, r, world)
        sl = slice(lay.load_lo, lay.load_hi)
        d = {"lay": lay,
             "L": torch.from_numpy(np.ascontiguousarray(iml[sl])).to(dev),
             "R": torch.from_numpy(np.ascontiguousarray(imr[sl])).to(dev)}
        if seed_l is not None:
            d["SL"] = torch.from_numpy(np.ascontiguousarray(seed_l[sl], np.float32)).to(dev)
            d["SR"] = torch.from_numpy(np.ascontiguousarray(seed_r[sl], np.float32)).to(dev)
        own = lay.own_hi - lay.own_lo
        d["OL"] = torch.empty((own, w), dtype=torch.float32, device=dev)
        d["OR"] = torch.empty((own, w), dtype=torch.float32, device=dev)
        res.append(d)
        e.band_begin(d["L"].data_ptr(), d["R"].data_ptr(), w, w, h, r, world,
                     d["SL"].data_ptr() if "SL" in d else None,
                     d["SR"].data_ptr() if "SR" in d else None, w * 4, pair_index, st)
    while True:
        xs = [e.band_step() for e in engs]
        if all(x is None for x in xs):
            break
        assert all(x is not None for x in xs), "bands left the schedule at different points"
        for r in range(world - 1):
            a, b = xs[r], xs[r + 1]
            assert a.send_next_bytes == b.recv_prev_bytes and b.send_prev_bytes == a.recv_next_bytes
            _as_tensor(b.recv_prev, b.recv_prev_bytes, dev).copy_(
                _as_tensor(a.send_next, a.send_next_bytes, dev))
            _as_tensor(a.recv_next, a.recv_next_bytes, dev).copy_(
                _as_tensor(b.send_prev, b.send_prev_bytes, dev))
    for e, d in zip(engs, res):
        e.band_finish(d["OL"].data_ptr(), d["OR"].data_ptr(), w * 4)
    torch.cuda.synchronize(dev)
    dl = np.concatenate([d["OL"].cpu().numpy() for d in res])
    dr = np.concatenate([d["OR"].cpu().numpy() for d in res])
    return dl, dr
