#!/usr/bin/env python
"""Config C5: one synthetic 3840x2160 frame, 256-disparity range, split into row bands
across the ranks of a torchrun job (one process per GPU, NCCL send/recv of the halo rows).

    python tools/band_bench.py                                  # 1 GPU, whole frame as one band
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/band_bench.py --check

Prints one JSON line on rank 0: frames/s (device time, max over ranks), exchange volume,
and with --check whether the gathered bands are bit-identical to the whole-frame pass that
rank 0 runs on its own GPU afterwards."""
import argparse
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--max-disp", type=int, default=256)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--check", action="store_true")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pkg = importlib.import_module("ocean-perception_b200")
    bands = importlib.import_module("ocean-perception_b200.bands")

    W, H, D = a.width, a.height, a.max_disp
    L, R, T = pkg.synth.make_pair(0, W, H, D)
    P = pkg.PatchmatchGpu.Params()
    P.init_mode, P.max_disp, P.patchmatch_iters, P.clamp_disp = "random", D, a.iters, 1
    bm = bands.BandedMatcher(P, device=local_rank)
    band = bm.upload(L, R)
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(a.warmup, 3)):
        bm.run(band)
    barrier()
    bm.exchanges = bm.exchange_bytes = 0
    bm.eng.launch_count(reset=True)
    bm.eng.set_profiling(True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(a.steps):
        bm.run(band)
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    stage = bm.eng.stage_ms()
    bm.eng.set_profiling(False)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())

    ol, orr = band["OL"], band["OR"]
    lay = band["lay"]
    ok = None
    if a.check:
        full_l = torch.zeros((H, W), dtype=torch.float32, device=dev)
        full_r = torch.zeros((H, W), dtype=torch.float32, device=dev)
        full_l[lay.own_lo:lay.own_hi] = ol
        full_r[lay.own_lo:lay.own_hi] = orr
        if world > 1:  # bands are disjoint: a sum gathers them (test bookkeeping, not the path)
            dist.all_reduce(full_l)
            dist.all_reduce(full_r)
        if rank == 0:
            eng = pkg.PatchmatchGpu(P, device=local_rank)
            wl, wr = eng.Match(L, R)
            eng.close()
            ok = bool(np.array_equal(full_l.cpu().numpy(), wl) and np.array_equal(full_r.cpu().numpy(), wr))
    if rank == 0:
        dl = ol.cpu().numpy()
        Tb = T[lay.own_lo:lay.own_hi]
        found = (dl > 0) & (Tb > 0)
        print(json.dumps({
            "metric": "frames_per_s_%dx%d_d%d_row_bands" % (W, H, D), "value": a.steps / (ms_max * 1e-3),
            "unit": "frames/s", "n_gpus": world, "steps": a.steps, "ms_per_frame": ms_max / a.steps,
            "scaling": "strong", "band_rows": lay.own_hi - lay.own_lo,
            "exchanges_per_frame": bm.exchanges / a.steps,
            "exchange_bytes_sent_per_frame_rank0": bm.exchange_bytes / a.steps,
            "gpu_launches_per_frame": bm.eng.launch_count() / a.steps,
            "stage_ms_per_frame_rank0": {k: v[0] / a.steps for k, v in stage.items()},
            "bit_identical_to_whole_frame": ok,
            "within_1px_of_truth_rank0": float((np.abs(dl - Tb)[found] <= 1).mean()) if found.any() else None,
        }))
    if world > 1:
        dist.destroy_process_group()
    bm.close()


if __name__ == "__main__":
    main()
