#!/usr/bin/env python
"""Streaming use of the path: a sequence of frames where every frame is seeded
with the previous frame's disparity instead of SparseInit -- the motivation the reference's
author states for the GPU library (src/vehicle/patchmatch_gpu/README.md:7) and the shape of
PatchmatchGpuTest.Sequence (test/stereo_matching/patchmatch_gpu_test.cpp:95-138), which plays a
EuRoC-layout dataset through Match() frame by frame.

    python tools/sequence_bench.py [--frames 30] [--width 1280 --height 720]

Synthetic sequence: one scene whose right image drifts by a sub-pixel amount per frame
(synth.make_pair with a per-frame disparity offset). Frame 0 runs the reference's SparseInit on
the device; frames k > 0 pass the maps of frame k-1 as seed_l / seed_r (device pointers, no host
round trip). Prints one JSON line: per-frame latency (device events), accuracy per frame."""
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
    ap.add_argument("--frames", type=int, default=30)
    ap.add_argument("--width", type=int, default=1280)
    ap.add_argument("--height", type=int, default=720)
    ap.add_argument("--max-disp", type=int, default=128)
    ap.add_argument("--iters", type=int, default=3)
    a = ap.parse_args()
    import torch
    pkg = importlib.import_module("ocean-perception_b200")
    W, H, D = a.width, a.height, a.max_disp
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    frames = [pkg.synth.make_pair(7, W, H, D, shift=0.25 * k) for k in range(a.frames)]
    P = pkg.PatchmatchGpu.Params()
    P.init_mode, P.patchmatch_iters = "sparse", a.iters
    P.matcher_params.max_disp = D
    eng = pkg.PatchmatchGpu(P, device=0)
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    dL = [torch.from_numpy(f[0]).to(dev) for f in frames]
    dR = [torch.from_numpy(f[1]).to(dev) for f in frames]
    out = [[torch.empty((H, W), dtype=torch.float32, device=dev) for _ in range(2)] for _ in range(2)]

    def run(k, seeded):
        cur, prv = out[k & 1], out[(k - 1) & 1]
        eng.match_batch_device(1, dL[k].data_ptr(), dR[k].data_ptr(), W, H, W, cur[0].data_ptr(),
                               cur[1].data_ptr(), W * 4,
                               d_seed_l=prv[0].data_ptr() if seeded else None,
                               d_seed_r=prv[1].data_ptr() if seeded else None,
                               stream=stream.cuda_stream)

    for k in range(min(3, a.frames)):   # warm-up
        run(k, k > 0)
    torch.cuda.synchronize(dev)
    ms, acc, valid = [], [], []
    for k in range(a.frames):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        run(k, k > 0)
        e1.record(stream)
        torch.cuda.synchronize(dev)
        ms.append(e0.elapsed_time(e1))
        d = out[k & 1][0].cpu().numpy()
        t = frames[k][2]
        found = (d > 0) & (t > 0)
        valid.append(float(found.mean()))
        acc.append(float((np.abs(d - t)[found] <= 1.0).mean()) if found.any() else 0.0)
    eng.close()
    print(json.dumps({
        "workload": "%d-frame synthetic sequence %dx%d, %d iterations, frame 0 SparseInit, "
                    "frames k>0 seeded with frame k-1's maps (device-resident)" % (a.frames, W, H, a.iters),
        "frame0_ms": ms[0], "seeded_ms_median": float(np.median(ms[1:])) if len(ms) > 1 else None,
        "frames_per_s_seeded": 1e3 / float(np.median(ms[1:])) if len(ms) > 1 else None,
        "valid_frac": [round(v, 4) for v in valid], "within_1px": [round(v, 4) for v in acc]}))


if __name__ == "__main__":
    main()
