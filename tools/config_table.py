#!/usr/bin/env python
"""The results table of BASELINE.md section 7: every config of BASELINE.json on one B200, with the CPU
legs timed on the box's host cores in the same run, and the reference's OWN kernels (oracle/_ref,
patchmatch_gpu.cu:18-295 recompiled for sm_100a) timed on the same GPU as a second baseline.

    python tools/config_table.py > profiles/r2_config_table.json      (needs a GPU)

Test/bench harness: uses the oracle and oracle/_ref as the things measured against, never the product."""
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def gpu_pairs_per_s(pkg, torch, P, L, R, n, reps=5):
    """Device-resident batch of n copies of (L, R) through pm_match_batch_device, CUDA events."""
    h, w = L.shape
    eng = pkg.PatchmatchGpu(P, device=0)
    dL = torch.from_numpy(L).cuda().repeat(n, 1, 1).contiguous()
    dR = torch.from_numpy(R).cuda().repeat(n, 1, 1).contiguous()
    oL = torch.empty((n, h, w), dtype=torch.float32, device="cuda")
    oR = torch.empty((n, h, w), dtype=torch.float32, device="cuda")
    st = torch.cuda.Stream()
    torch.cuda.synchronize()

    def run():
        eng.match_batch_device(n, dL.data_ptr(), dR.data_ptr(), w, h, w, oL.data_ptr(), oR.data_ptr(), w * 4,
                               stream=st.cuda_stream)
    for _ in range(3):
        run()
    eng.synchronize(st.cuda_stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps):
        run()
    e1.record(st)
    eng.synchronize(st.cuda_stream)
    ms = e0.elapsed_time(e1) / reps
    out = oL[0].cpu().numpy(), oR[0].cpu().numpy()
    eng.close()
    return n / (ms * 1e-3), ms, out


def main():
    import torch
    import pmo
    import pmref
    pkg = importlib.import_module("ocean-perception_b200")
    rows = []
    cores = os.cpu_count()

    def P_(**kw):
        P = pkg.PatchmatchGpu.Params()
        for k, v in kw.items():
            setattr(P, k, v)
        return P

    # C1: the reference's CPU-runnable case (stereo::Patchmatch on the fixture)
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "c1_inputs.npz")))
    want = dict(np.load(os.path.join(ROOT, "tests", "golden", "c1_cpu.npz")))["final"]
    t0 = time.perf_counter()
    cpu = pmo.c_estimate_disparity(g["il"], g["ir"], pmo.c_initialize(g["il"], g["ir"], 1))
    c1_cpu = time.perf_counter() - t0
    pmc = pkg.Patchmatch(device=0)
    pmc.EstimateDisparity(g["il"], g["ir"])
    t0 = time.perf_counter()
    for _ in range(5):
        got = pmc.EstimateDisparity(g["il"], g["ir"])
    c1_gpu = (time.perf_counter() - t0) / 5
    pmc.close()
    pps, ms, _ = gpu_pairs_per_s(pkg, torch, P_(), g["il"], g["ir"], 1, reps=20)
    rows.append({"config": "C1 fsl1/fsr1 376x240, stereo::Patchmatch stage library (CPU semantics)",
                 "cpu_1core_s_per_frame": c1_cpu, "gpu_s_per_frame": c1_gpu,
                 "cpu_equals_golden": bool(np.array_equal(cpu, want)), "gpu_equals_golden": bool(np.array_equal(got, want)),
                 "gpu_library_match_ms_per_pair": ms})

    for name, w, h, D, levels, init in (("C2 752x480 D64 1 level", 752, 480, 64, 1, "random"),
                                        ("C3 1280x720 D128 2 levels", 1280, 720, 128, 2, "random"),
                                        ("reference defaults 1280x720 (device SparseInit, 1 level)", 1280, 720, 128, 1, "sparse")):
        L, R, T = pkg.synth.make_pair(1, w, h, D)
        P = P_(init_mode=init, max_disp=D, pyramid_levels=levels)
        one, ms1, (dl, dr) = gpu_pairs_per_s(pkg, torch, P, L, R, 1, reps=20)
        many, ms64, _ = gpu_pairs_per_s(pkg, torch, P, L, R, 64, reps=3)
        po = pmo.default_params(init_mode=1 if init == "random" else 0, max_disp=D, pyramid_levels=levels)
        t0 = time.perf_counter()
        if init == "random":
            wl, wr = pmo.g_match(po, L, R, pair_index=0)
        else:
            sl, sr = pmo.s_match_seeds(L, R, 4)
            wl, wr = pmo.g_match(po, L, R, sl, sr)
        cpu_s = time.perf_counter() - t0
        found = (dl > 0) & (T > 0)
        row = {"config": name, "gpu_single_pair_ms": ms1, "gpu_single_pair_pairs_per_s": one,
               "gpu_batch64_pairs_per_s": many, "cpu_port_1core_pairs_per_s": 1.0 / cpu_s,
               "bit_exact_vs_oracle": bool(np.array_equal(dl, wl) and np.array_equal(dr, wr)),
               "valid_frac": float(found.mean()),
               "within_1px_of_truth": float((np.abs(dl - T)[found] <= 1).mean())}
        if init == "sparse":
            # the reference's own kernels on this GPU, same planes, same seeds, stock launches
            noise = pmo.rng_uniform(123, -1, 1, w * h).reshape(h, w)
            tot = 0.0
            for view in (0, 1):
                planes = pmo.g_planes(L, R, view)
                seed = sl if view == 0 else np.ascontiguousarray(sr[:, ::-1])
                tot += pmref.match_view_timed(*planes, noise, seed, reps=5)
            row["reference_kernels_b200_ms_per_pair_device_loop_only"] = tot
            row["reference_kernels_b200_pairs_per_s_device_loop_only"] = 1e3 / tot
            row["note"] = ("reference kernels = patchmatch_gpu.cu:18-295 compiled verbatim for sm_100a, both views, "
                           "device loop only (:394-410; no SparseInit, no upload/download, no GradientMagnitude)")
        rows.append(row)
    print(json.dumps({"host_cores": cores, "rows": rows}, indent=1))


if __name__ == "__main__":
    main()
