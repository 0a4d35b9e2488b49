#!/usr/bin/env python
"""Config C1 of BASELINE.json: the reference's own CPU-runnable case -- stereo::Patchmatch on
the fixture pair fsl1/fsr1 at 376x240 (test/stereo_matching/patchmatch_test.cpp:116-188),
single frame. Times, on this box:
  * the CPU algorithm (oracle port of patchmatch.cpp + the test's functor and schedule), one core;
  * the same stage library on the GPU (pm_cpu_* : Initialize + 4 x (AddNoise, Propagate) +
    RemoveBackground), result compared bit for bit with the cv2-literal golden;
  * the GPU library's Match (reference defaults, device SparseInit) on the same pair.
Prints one JSON line."""
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    import pmo
    pkg = importlib.import_module("ocean-perception_b200")
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "c1_inputs.npz")))
    want = dict(np.load(os.path.join(ROOT, "tests", "golden", "c1_cpu.npz")))["final"]
    il, ir = g["il"], g["ir"]
    t0 = time.perf_counter()
    seed = pmo.c_initialize(il, ir, 1)
    cpu = pmo.c_estimate_disparity(il, ir, seed)
    t_cpu = time.perf_counter() - t0
    eng = pkg.PatchmatchGpu(device=0)
    pmc = pkg.Patchmatch(eng)
    pmc.EstimateDisparity(il, ir)   # warm-up
    n = 5
    t0 = time.perf_counter()
    for _ in range(n):
        got = pmc.EstimateDisparity(il, ir)
    t_gpu_c = (time.perf_counter() - t0) / n
    eng.Match(il, ir)
    t0 = time.perf_counter()
    for _ in range(20):
        dl, dr = eng.Match(il, ir)
    t_gpu_g = (time.perf_counter() - t0) / 20
    eng.close()
    print(json.dumps({
        "config": "C1: fsl1/fsr1 at 376x240, single frame",
        "cpu_stage_library_1core_s": t_cpu, "cpu_equals_golden": bool(np.array_equal(cpu, want)),
        "gpu_stage_library_s": t_gpu_c, "gpu_equals_golden": bool(np.array_equal(got, want)),
        "gpu_library_match_s": t_gpu_g, "gpu_library_valid_frac": float((dl > 0).mean()),
        "note": "a raster pass of stereo::Patchmatch is one dependent chain per image line: 240-376 chains "
                "per pass, one warp each (lanes = patch pixels); Match is the chunked-sweep GPU library path"}))


if __name__ == "__main__":
    main()
