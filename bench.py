#!/usr/bin/env python
"""Headline benchmark: stereo pairs/s at 1280x720, 128 disparities (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU algorithm

Workload (config C4 of BASELINE.md): a batch of synthetic 1280x720 pairs, 128-disparity
range, 3 PatchMatch iterations, 2 pyramid levels, the reference's stage list (noise, four
chunked sweeps, background mask, left-right check) and per-pixel random init. One process
per GPU; every rank owns `--pairs-per-gpu` independent pairs (no data-path collective:
scaling "weak"). A step is one pass of the hot path over the rank's batch.

Prints ONE JSON line on rank 0. `value` = pairs all ranks processed / max-over-ranks device
time with inputs and outputs resident in HBM; `e2e` = the same through the reference-facing
C-ABI call with pinned HOST buffers (host<->device copies inside the timed region).
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "stereo_pairs_per_s_1280x720_d128"
UNIT = "pairs/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs-per-gpu", type=int, default=64)
    ap.add_argument("--unique-pairs", type=int, default=8,
                    help="distinct synthetic pairs generated per rank (repeated cyclically)")
    ap.add_argument("--width", type=int, default=1280)
    ap.add_argument("--height", type=int, default=720)
    ap.add_argument("--max-disp", type=int, default=128)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--levels", type=int, default=2)
    ap.add_argument("--init", default="random", choices=["random", "sparse"],
                    help="random: the north-star workload (per-pixel Philox init); sparse: the "
                         "reference's SparseInit (GFTT + template matching + dilate) on the device")
    ap.add_argument("--cpu-sample-pairs", type=int, default=8,
                    help="pairs the cpu_baseline leg times on one host core")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--max-batch", type=int, default=0,
                    help="pairs per device pass (0 = the engine's choice)")
    return ap.parse_args()


def workload_config(a):
    return {
        "workload": "C4: batch of synthetic %dx%d pairs, %d-disparity range, %d iterations, "
                    "%d-level pyramid, %s init, reference stage list" %
                    (a.width, a.height, a.max_disp, a.iters, a.levels, a.init),
        "pairs_per_gpu": a.pairs_per_gpu, "width": a.width, "height": a.height,
        "max_disp": a.max_disp, "iters": a.iters, "pyramid_levels": a.levels,
        "sweep_chunks": 16, "sweep_overlap": 5, "cost": "l1grad_x5 (5 taps)",
        "parallelism": "independent pairs sharded across ranks, no collective",
        "l2": "per-step working set (>= 1 GB of planes per device pass) exceeds the 126 MB L2",
    }


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md):
    one `nvidia-smi -lms 50` process started before the region and stopped after it."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def summary(self):
        rows = []
        if self.proc is not None:
            self.proc.terminate()
            try:
                out, _ = self.proc.communicate(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
                out, _ = self.proc.communicate()
            rows = [[c.strip() for c in l.split(",")] for l in out.splitlines() if l.strip()]

        def num(v):
            try:
                return float(v)
            except ValueError:
                return None
        sm = [num(r[0]) for r in rows if num(r[0]) is not None]
        mx = [num(r[1]) for r in rows if len(r) > 1 and num(r[1]) is not None]
        pw = [num(r[2]) for r in rows if len(r) > 2 and num(r[2]) is not None]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows if len(r) >= 7 for i in range(4)
                          if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "power_w_max": max(pw) if pw else None,
                "reasons": reasons, "samples": len(rows)}


def bind_to_gpu_numa_node(gpu_index):
    """Pin this rank's host threads to the CPUs NVML reports as local to its GPU, before any pinned
    buffer is allocated (first touch then lands on the GPU's NUMA node). With 8 ranks on one box
    the host<->device copies of the e2e leg otherwise cross sockets. Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = {64 * i + b for i, wd in enumerate(words) for b in range(64) if (wd >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def measured_sweep_traffic(a):
    """Mean DRAM bytes per sweep launch from the committed ncu capture of this workload
    (profiles/r1d_sweep_dram_bytes.json, tools/profile_round.sh); None for other workloads."""
    path = os.path.join(ROOT, "profiles", "r1d_sweep_dram_bytes.json")
    try:
        with open(path) as f:
            d = json.load(f)
    except OSError:
        return None, None
    wl = d.get("workload", {})
    mine = {"pairs_per_gpu": a.pairs_per_gpu, "width": a.width, "height": a.height,
            "pyramid_levels": a.levels, "iters": a.iters}
    if wl != mine:
        return None, None
    return d["mean_dram_bytes_per_launch"], "profiles/r1d_sweep_dram_bytes.json (ncu, per launch)"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback (B200_PROFILING.md)"


# Algorithmic work of one pair (DESIGN.md "Work accounting").
def sweep_bytes_per_pair(n_px):
    # one sweep pass, gradient stored: 2 views x (2 u8 images + 2 f32 gradients + disparity
    # read + disparity write) = 36 bytes per pixel
    return 36.0 * n_px


def evals_per_pair(n_px, iters, levels, chunks=16, overlap=5, width=1280, height=720):
    """Hypothesis evaluations the params specify (cost of the current disparity cached):
    per view and iteration: 1 (noise) + 4 sweeps; + 1 for the background mask."""
    total = 0.0
    for l in range(levels):
        total += 2 * (iters * 5.0) * (n_px / 4.0 ** l)
    total += 2 * n_px  # MaskBackground: cost(0)
    return total


OPS_PER_EVAL = 5 * 21  # 5 taps x (2 lerps = 6, 2 |diff| = 4, weighted sum = 3, frac/floor = 8)


def run_reference(a, rank, world, out_line):
    """The reference's CPU algorithm (oracle port: the reference cannot be compiled here)
    on all host threads, rank 0 only."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pmo
    from concurrent.futures import ThreadPoolExecutor
    pkg = importlib.import_module("ocean-perception_b200")
    cores = os.cpu_count() or 1
    n = max(cores, 1)  # one pair per thread per step: a bounded sample of the workload
    L, R, _ = pkg.synth.make_batch(0, n, a.width, a.height, a.max_disp, unique=min(n, a.unique_pairs))
    p = pmo.default_params(init_mode=1 if a.init == "random" else 0, max_disp=a.max_disp,
                           pyramid_levels=a.levels, patchmatch_iters=a.iters)
    pmo.lib()

    def one(i):
        if a.init == "random":
            pmo.g_match(p, L[i], R[i], pair_index=i)
        else:
            sl, sr = pmo.s_match_seeds(L[i], R[i], 4)
            pmo.g_match(p, L[i], R[i], sl, sr, pair_index=i)

    def step():
        with ThreadPoolExecutor(cores) as ex:
            list(ex.map(one, range(n)))

    for _ in range(min(a.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step()
    dt = time.perf_counter() - t0
    value = n * a.steps / dt
    cfg = workload_config(a)
    sample = "%d pairs per step (one per host thread) of the same synthetic workload" % n
    out_line.emit({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus,
        "steps": a.steps, "warmup": min(a.warmup, 1), "ms_per_step": 1e3 * dt / a.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


class OneLineStdout:
    """Everything libraries print (NCCL banners, warnings) goes to stderr; stdout carries
    exactly the one JSON line the driver parses."""

    def __init__(self):
        sys.stdout.flush()
        self.fd = os.dup(1)
        os.dup2(2, 1)

    def emit(self, obj):
        sys.stdout.flush()
        os.write(self.fd, (json.dumps(obj) + "\n").encode())


def main():
    a = parse_args()
    out_line = OneLineStdout()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if a.impl == "reference":
        run_reference(a, rank, world, out_line)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU path "
                         "(use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    numa_cpus = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    pkg = importlib.import_module("ocean-perception_b200")
    importlib.import_module("ocean-perception_b200.build").build()

    W, H, B = a.width, a.height, a.pairs_per_gpu
    n_px = W * H
    first = rank * B
    Lh, Rh, Th = pkg.synth.make_batch(first, B, W, H, a.max_disp, unique=a.unique_pairs)

    P = pkg.PatchmatchGpu.Params()
    P.init_mode, P.max_disp, P.pyramid_levels, P.patchmatch_iters = a.init, a.max_disp, a.levels, a.iters
    P.max_batch = a.max_batch
    eng = pkg.PatchmatchGpu(P, device=local_rank)

    dev = torch.device("cuda", local_rank)
    dL = torch.from_numpy(Lh).to(dev)
    dR = torch.from_numpy(Rh).to(dev)
    oL = torch.empty((B, H, W), dtype=torch.float32, device=dev)
    oR = torch.empty((B, H, W), dtype=torch.float32, device=dev)
    stream = torch.cuda.Stream(dev)  # a real (non-default) stream: the engine launches on it
    torch.cuda.set_stream(stream)

    def step_device():
        eng.match_batch_device(B, dL.data_ptr(), dR.data_ptr(), W, H, W, oL.data_ptr(),
                               oR.data_ptr(), W * 4, first_pair_index=first,
                               stream=stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(a.warmup, 3)):
        step_device()
    barrier()
    eng.launch_count(reset=True)
    eng.set_profiling(True)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)  # let the sampler come up; the GPU idles meanwhile
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for _ in range(a.steps):
        step_device()
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = eng.launch_count()
    stage = eng.stage_ms()
    eng.set_profiling(False)
    clocks = sampler.summary() if sampler else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * B * a.steps / (ms_max * 1e-3)

    # accuracy against the synthetic truth (information only)
    dl = oL[: min(B, a.unique_pairs)].cpu().numpy()
    T = Th[: dl.shape[0]]
    found = (dl > 0) & (T > 0)
    quality = {"valid_frac": float(found.mean()),
               "within_1px_of_truth": float((np.abs(dl - T)[found] <= 1.0).mean()) if found.any() else 0.0}

    # ---- e2e: the C-ABI host call, pinned host buffers, copies inside the timed region
    e2e = None
    if not a.no_e2e:
        pL = torch.from_numpy(Lh).pin_memory()
        pR = torch.from_numpy(Rh).pin_memory()
        hL = torch.empty((B, H, W), dtype=torch.float32).pin_memory()
        hR = torch.empty((B, H, W), dtype=torch.float32).pin_memory()
        import ctypes as C

        def step_host():
            rc = eng._lib.pm_match_batch_host(eng._h, B, C.c_void_p(pL.data_ptr()),
                                              C.c_void_p(pR.data_ptr()), W, H, W, None, None, first,
                                              C.c_void_p(hL.data_ptr()), C.c_void_p(hR.data_ptr()), W * 4)
            if rc != 0:
                raise RuntimeError(eng._lib.pm_last_error(eng._h).decode())

        step_host()
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            step_host()
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        td = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(td, op=dist.ReduceOp.MAX)
        e2e = {"value": world * B * a.steps / float(td.item()), "unit": UNIT,
               "h2d_bytes_per_step": 2 * B * n_px, "d2h_bytes_per_step": 2 * B * n_px * 4,
               "checksum": float(hL.sum().item())}

    if rank == 0:
        peaks, peak_src = measured_peaks()
        # dominant kernel = the sweep stages (row + column): algorithmic bytes per launch
        # (36 B/px/pair/sweep) over its CUDA-event time inside the timed region
        sw_ms = stage["sweep_row"][0] + stage["sweep_col"][0]
        sw_n = stage["sweep_row"][1] + stage["sweep_col"][1]
        total_stage_ms = sum(v[0] for v in stage.values())
        # bytes of all sweep launches in the timed region: every level, 4 sweeps x iters
        sweep_bytes = sum(sweep_bytes_per_pair(n_px / 4.0 ** l) * 4 * a.iters for l in range(a.levels))
        sweep_bytes *= B * a.steps
        achieved = sweep_bytes / (sw_ms * 1e-3) / 1e9 if sw_ms > 0 else 0.0
        traffic, traffic_src = measured_sweep_traffic(a)
        roofline = {"bound": "hbm", "kernel": "k_sweep (row + column sweeps)",
                    "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": achieved / peaks["hbm_gbs"], "traffic": traffic,
                    "traffic_source": traffic_src,
                    "algorithmic_bytes_per_launch": sweep_bytes / max(sw_n, 1),
                    # the same launches counted with the bytes ncu saw move (planes are float2
                    # {I,G} and {d,cost}: 32 B per pixel and view instead of the canonical 18)
                    "dram_gbs": (traffic * sw_n / (sw_ms * 1e-3) / 1e9) if traffic and sw_ms > 0 else None,
                    "peak_source": peak_src, "launches": sw_n,
                    "avg_launch_ms": sw_ms / max(sw_n, 1),
                    "share_of_step": sw_ms / total_stage_ms if total_stage_ms else None}
        fp32_peak = 148 * 128 * 2 * peaks.get("sm_max_mhz", 1965.0) * 1e6 / 1e12
        evals = evals_per_pair(n_px, a.iters, a.levels) * B * a.steps * world
        alu = {"evals_per_pair": evals_per_pair(n_px, a.iters, a.levels),
               "ops_per_eval": OPS_PER_EVAL,
               "achieved_tflops": evals * OPS_PER_EVAL / (ms_max * 1e-3) / 1e12 / world,
               "peak_tflops": fp32_peak, "peak_source": "148 SM x 128 lanes x 2 x sm_max_mhz"}
        alu["frac"] = alu["achieved_tflops"] / fp32_peak
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": max(a.warmup, 3), "ms_per_step": ms_max / a.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(a), "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
            "host_cpus_bound_to_gpu_node": numa_cpus,
            "roofline": roofline, "alu": alu,
            "stage_ms_per_step": {k: v[0] / a.steps for k, v in stage.items()},
            "quality": quality,
        }
        if not a.no_cpu_baseline and world == 1:   # the CPU baseline leg runs at N = 1 only
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import pmo
            n = max(1, a.cpu_sample_pairs)
            p = pmo.default_params(init_mode=1 if a.init == "random" else 0, max_disp=a.max_disp,
                                   pyramid_levels=a.levels, patchmatch_iters=a.iters)
            t0 = time.perf_counter()
            agree = []
            for i in range(n):
                if a.init == "random":
                    wl, wr = pmo.g_match(p, Lh[i], Rh[i], pair_index=first + i)
                else:
                    sl, sr = pmo.s_match_seeds(Lh[i], Rh[i], 4)
                    wl, wr = pmo.g_match(p, Lh[i], Rh[i], sl, sr, pair_index=first + i)
                if i < dl.shape[0]:
                    agree.append(bool(np.array_equal(wl, dl[i])))
            dt = time.perf_counter() - t0
            out["cpu_baseline"] = {"value": n / dt, "unit": UNIT, "cores": 1, "kind": "port",
                                   "sample": "%d pairs of the same workload, one host thread "
                                             "(oracle/pm_oracle.c, -O3 -march=native)" % n,
                                   "bit_exact_vs_gpu": agree}
        out_line.emit(out)
    if world > 1:
        dist.destroy_process_group()
    eng.close()


if __name__ == "__main__":
    main()
