#!/usr/bin/env python
"""Second quality baseline: the reference's SGBM wrapper
stereo::EstimateDisparity (src/vehicle/stereo_matching/stereo_matching.cpp:11-41 =
cv::StereoSGBM::create(0, num_disp, block_size), MODE_SGBM, result / 16) next to the
PatchMatch engine on the same synthetic pairs with ground truth.

    python tools/sgbm_compare.py [--pairs 4] [--width 1280 --height 720 --max-disp 128]

Needs a GPU for the engine and cv2 for SGBM (library code: not part of the product path).
Prints one JSON line: valid fraction, fraction within 1 px of the truth, mean end-point
error on valid pixels, and wall-clock per pair (SGBM on the host's cores, engine on the GPU
through the host-buffer call)."""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def estimate_disparity_sgbm(il, ir, num_disp, block_size):
    """stereo::EstimateDisparity, stereo_matching.cpp:11-41."""
    import cv2
    sgbm = cv2.StereoSGBM_create(0, num_disp, block_size)
    sgbm.setMinDisparity(0)
    sgbm.setNumDisparities(num_disp)
    sgbm.setMode(cv2.StereoSGBM_MODE_SGBM)
    disp = sgbm.compute(il, ir)                      # int16, fixed point x16 (:34-38)
    return disp.astype(np.float32) * np.float32(1.0 / 16.0)


def score(d, t):
    found = (d > 0) & (t > 0)
    err = np.abs(d - t)[found]
    return {"valid_frac": float(found.mean()),
            "within_1px": float((err <= 1.0).mean()) if err.size else 0.0,
            "mean_epe": float(err.mean()) if err.size else None}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=4)
    ap.add_argument("--width", type=int, default=1280)
    ap.add_argument("--height", type=int, default=720)
    ap.add_argument("--max-disp", type=int, default=128)
    ap.add_argument("--block-size", type=int, default=5)
    ap.add_argument("--init", default="sparse", choices=["sparse", "random"])
    a = ap.parse_args()
    pkg = importlib.import_module("ocean-perception_b200")
    L, R, T = pkg.synth.make_batch(0, a.pairs, a.width, a.height, a.max_disp, unique=a.pairs)
    P = pkg.PatchmatchGpu.Params()
    P.init_mode, P.max_disp = a.init, a.max_disp
    P.matcher_params.max_disp = a.max_disp
    eng = pkg.PatchmatchGpu(P, device=0)
    eng.MatchBatch(L[:1], R[:1])                     # warm-up (workspace, noise image)
    t0 = time.perf_counter()
    dl, _ = eng.MatchBatch(L, R)
    t_pm = (time.perf_counter() - t0) / a.pairs
    eng.close()
    t0 = time.perf_counter()
    ds = np.stack([estimate_disparity_sgbm(L[i], R[i], a.max_disp, a.block_size) for i in range(a.pairs)])
    t_sg = (time.perf_counter() - t0) / a.pairs
    out = {"workload": "%d synthetic %dx%d pairs, %d disparities" % (a.pairs, a.width, a.height, a.max_disp),
           "patchmatch_b200": dict(score(dl, T), s_per_pair=t_pm, init=a.init),
           "sgbm_cv2": dict(score(ds, T), s_per_pair=t_sg, block_size=a.block_size,
                            cores=os.cpu_count())}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
