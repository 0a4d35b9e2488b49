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
    ap.add_argument("--pairs-total", type=int, default=512,
                    help="config C4: the batch of pairs one step processes, split over the ranks "
                         "(512 / 256 / 128 / 64 per GPU at 1 / 2 / 4 / 8 GPUs)")
    ap.add_argument("--pairs-per-gpu", type=int, default=0,
                    help="> 0: fixed pairs per rank instead (weak scaling; tools and profiling)")
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
    ap.add_argument("--no-c5", action="store_true", help="skip the row-band (C5) leg at N >= 2")
    ap.add_argument("--c5-frames", type=int, default=10)
    ap.add_argument("--c5-nccl", action="store_true",
                    help="C5 leg: move the halo rows with NCCL send/recv instead of peer-memory stores")
    ap.add_argument("--max-batch", type=int, default=0,
                    help="pairs per device pass (0 = the engine's choice)")
    return ap.parse_args()


def pairs_of_rank(a, rank, world):
    """(first pair, pairs) of one rank: the contiguous split of SURVEY.md 8e, or a fixed count."""
    if a.pairs_per_gpu > 0:
        return rank * a.pairs_per_gpu, a.pairs_per_gpu
    lo, hi = (rank * a.pairs_total) // world, ((rank + 1) * a.pairs_total) // world
    return lo, hi - lo


def workload_config(a, world=1):
    strong = a.pairs_per_gpu <= 0
    first, per = pairs_of_rank(a, 0, world)
    return {
        "workload": "C4: batch of %s synthetic %dx%d pairs, %d-disparity range, %d iterations, "
                    "%d-level pyramid, %s init, reference stage list" %
                    ("%d" % a.pairs_total if strong else "%d per GPU" % a.pairs_per_gpu,
                     a.width, a.height, a.max_disp, a.iters, a.levels, a.init),
        "pairs_total": a.pairs_total if strong else per * world, "pairs_per_gpu": per,
        "width": a.width, "height": a.height,
        "max_disp": a.max_disp, "iters": a.iters, "pyramid_levels": a.levels,
        "sweep_chunks": 16, "sweep_overlap": 5, "cost": "l1grad_x5 (5 taps)",
        "parallelism": "independent pairs sharded across ranks (contiguous ranges), no collective",
        "unique_pairs": "%d distinct synthetic pairs per rank, repeated cyclically" % a.unique_pairs,
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


def measured_sweep_traffic(a, pairs_per_launch):
    """Mean DRAM bytes per sweep launch from the committed ncu capture of this workload
    (profiles/*_sweep_dram_bytes.json, tools/profile_r2.sh); None for other workloads. The capture
    ran device passes of `pairs_per_gpu` pairs; the bytes of a launch are proportional to the pairs
    it covers (every plane is streamed once per sweep), so they are scaled to this run's pass size."""
    for name in ("r5_sweep_dram_bytes.json", "r4_sweep_dram_bytes.json", "r3_sweep_dram_bytes.json", "r2_sweep_dram_bytes.json",
                 "r1d_sweep_dram_bytes.json"):
        path = os.path.join(ROOT, "profiles", name)
        try:
            with open(path) as f:
                d = json.load(f)
        except OSError:
            continue
        wl = dict(d.get("workload", {}))
        captured = wl.pop("pairs_per_gpu", None)
        mine = {"width": a.width, "height": a.height, "pyramid_levels": a.levels, "iters": a.iters}
        if wl == mine and captured:
            scale = pairs_per_launch / float(captured)
            return d["mean_dram_bytes_per_launch"] * scale, \
                "profiles/%s (ncu, per launch of %d pairs, scaled to %.0f pairs per launch)" % (
                    name, captured, pairs_per_launch)
    return None, None


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


def evals_per_pair(n_px, iters, levels):
    """Hypothesis evaluations the params specify (cost of the current disparity cached):
    per view and iteration: 1 (noise) + 4 sweeps; + 1 for the background mask."""
    total = 0.0
    for l in range(levels):
        total += 2 * (iters * 5.0) * (n_px / 4.0 ** l)
    total += 2 * n_px  # MaskBackground: cost(0)
    return total


OPS_PER_EVAL = 5 * 21  # 5 taps x (2 lerps = 6, 2 |diff| = 4, weighted sum = 3, frac/floor = 8)


def oracle():
    """The CPU oracle: only the cpu_baseline legs and the reference arm come here."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pmo
    pmo.lib()
    return pmo


def cpu_c1_leg(pmo):
    """Config C1, the reference's own CPU-runnable case: stereo::Patchmatch (the oracle's port of
    patchmatch.cpp with the functor and schedule of patchmatch_test.cpp:116-188) on the fixture pair
    fsl1/fsr1 at 376x240, one frame, one core."""
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "c1_inputs.npz")))
    want = dict(np.load(os.path.join(ROOT, "tests", "golden", "c1_cpu.npz")))["final"]
    t0 = time.perf_counter()
    seed = pmo.c_initialize(g["il"], g["ir"], 1)
    cpu = pmo.c_estimate_disparity(g["il"], g["ir"], seed)
    dt = time.perf_counter() - t0
    return g, want, cpu, dt


def run_reference(a, rank, world, out_line):
    """The reference's CPU algorithm (oracle port: the reference cannot be compiled here)
    on all host threads, rank 0 only."""
    if rank != 0:
        return
    pmo = oracle()
    from concurrent.futures import ThreadPoolExecutor
    pkg = importlib.import_module("ocean-perception_b200")
    cores = os.cpu_count() or 1
    n = max(cores, 1)  # one pair per thread per step: a bounded sample of the workload
    L, R, _ = pkg.synth.make_batch(0, n, a.width, a.height, a.max_disp, unique=min(n, a.unique_pairs))
    p = pmo.default_params(init_mode=1 if a.init == "random" else 0, max_disp=a.max_disp,
                           pyramid_levels=a.levels, patchmatch_iters=a.iters)

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
    _, want, cpu, c1_dt = cpu_c1_leg(pmo)
    cfg = workload_config(a, world)
    sample = ("%d pairs per step (one per host thread) of the same synthetic workload; the algorithm is "
              "the GPU library's (patchmatch_gpu.cu) restated for the CPU, which is FASTER than the "
              "reference's own CPU class stereo::Patchmatch (see c1_stereo_patchmatch)" % n)
    out_line.emit({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus,
        "steps": a.steps, "warmup": min(a.warmup, 1), "ms_per_step": 1e3 * dt / a.steps,
        "higher_is_better": True, "scaling": "strong" if a.pairs_per_gpu <= 0 else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": sample},
        "c1_stereo_patchmatch": {"seconds_per_frame_1core": c1_dt, "frame": "fsl1/fsr1 at 376x240",
                                 "equals_cv2_golden": bool(np.array_equal(cpu, want))},
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


def copy_ceiling(torch, dist, dev, world, pL, pR, hL, hR, dL, dR, oL, oR, reps=3):
    """Host<->device copies of one step's buffers with NO kernel, both directions at once on two
    streams, all ranks together: the pairs/s the box's pinned-copy path allows (max time over ranks)."""
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    best = None
    for _ in range(reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        with torch.cuda.stream(s_in):
            dL.copy_(pL, non_blocking=True)
            dR.copy_(pR, non_blocking=True)
        with torch.cuda.stream(s_out):
            hL.copy_(oL, non_blocking=True)
            hR.copy_(oR, non_blocking=True)
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        td = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(td, op=dist.ReduceOp.MAX)
        dt = float(td.item())
        best = dt if best is None else min(best, dt)
    return best


def c5_band_leg(a, pkg, torch, dist, rank, local_rank, world):
    """Config C5 under the driver's eyes: one synthetic 3840x2160 frame, 256-disparity range, split
    into row bands over the ranks (NCCL send/recv of the halo rows, bands.py); returns the dict of
    the `c5_band` key on rank 0."""
    bands = importlib.import_module("ocean-perception_b200.bands")
    W, H, D = 3840, 2160, 256
    dev = torch.device("cuda", local_rank)
    L, R, T = pkg.synth.make_pair(0, W, H, D)
    P = pkg.PatchmatchGpu.Params()
    P.init_mode, P.max_disp, P.patchmatch_iters, P.clamp_disp = "random", D, a.iters, 1
    bm = bands.BandedMatcher(P, device=local_rank, p2p=not a.c5_nccl)
    band = bm.upload(L, R)

    def barrier():
        dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(3):
        bm.run(band)
    barrier()
    bm.exchanges = bm.exchange_bytes = 0
    bm.eng.set_profiling(True)
    stream = torch.cuda.current_stream(dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(a.c5_frames):
        bm.run(band)
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    stage = bm.eng.stage_ms()
    bm.eng.set_profiling(False)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    # NCCL transport: the exchanges alone, back to back, no kernels; peer memory: the engine's stage time
    xms = bm.measure_exchange_ms(band, reps=5)
    if xms is None:
        xms = stage["band_exchange"][0] / a.c5_frames
    lay = band["lay"]
    full_l = torch.zeros((H, W), dtype=torch.float32, device=dev)
    full_r = torch.zeros((H, W), dtype=torch.float32, device=dev)
    full_l[lay.own_lo:lay.own_hi] = band["OL"]
    full_r[lay.own_lo:lay.own_hi] = band["OR"]
    dist.all_reduce(full_l)   # bands are disjoint: a sum gathers them (a check, not the path)
    dist.all_reduce(full_r)
    out = None
    if rank == 0:
        eng = pkg.PatchmatchGpu(P, device=local_rank)
        wl, wr = eng.Match(L, R)
        eng.close()
        ok = bool(np.array_equal(full_l.cpu().numpy(), wl) and np.array_equal(full_r.cpu().numpy(), wr))
        found = (wl > 0) & (T > 0)
        out = {"workload": "C5: one synthetic %dx%d frame, %d-disparity range, %d iterations, row bands "
                           "of whole column-sweep chunks over %d GPUs, halo rows %s"
                           % (W, H, D, a.iters, world,
                              "stored into the neighbour's buffer over NVLink (peer memory)" if bm.p2p
                              else "by NCCL send/recv"),
               "frames_per_s": a.c5_frames / (ms_max * 1e-3), "ms_per_frame": ms_max / a.c5_frames,
               "frames": a.c5_frames, "band_rows": lay.own_hi - lay.own_lo,
               "exchanges_per_frame": bm.exchanges / a.c5_frames,
               "exchange_bytes_sent_per_frame_rank0": bm.exchange_bytes / a.c5_frames,
               "exchange_ms": xms, "exchange_transport": "peer memory (pm_band_p2p)" if bm.p2p else "NCCL send/recv",
               "exchange_overlap": bm.overlap,
               "bit_identical": ok,
               "stage_ms": {k: v[0] / a.c5_frames for k, v in stage.items()},
               "within_1px_of_truth": float((np.abs(wl - T)[found] <= 1).mean()) if found.any() else None}
    bm.close()
    return out


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

    W, H = a.width, a.height
    n_px = W * H
    first, B = pairs_of_rank(a, rank, world)
    total_pairs = a.pairs_total if a.pairs_per_gpu <= 0 else B * world
    U = max(1, min(a.unique_pairs, B))
    Lu, Ru, Tu = pkg.synth.make_batch(first, U, W, H, a.max_disp)

    P = pkg.PatchmatchGpu.Params()
    P.init_mode, P.max_disp, P.pyramid_levels, P.patchmatch_iters = a.init, a.max_disp, a.levels, a.iters
    P.max_batch = a.max_batch
    eng = pkg.PatchmatchGpu(P, device=local_rank)

    dev = torch.device("cuda", local_rank)
    reps = (B + U - 1) // U
    dL = torch.from_numpy(Lu).to(dev).repeat(reps, 1, 1)[:B].contiguous()
    dR = torch.from_numpy(Ru).to(dev).repeat(reps, 1, 1)[:B].contiguous()
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

    warm = max(a.warmup, 3)
    for _ in range(warm):
        step_device()
    barrier()
    fp32_peak = eng.measure_fp32_peak() if rank == 0 else None
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
    value = total_pairs * a.steps / (ms_max * 1e-3)

    # accuracy against the synthetic truth (information only)
    dl = oL[:U].cpu().numpy()
    found = (dl > 0) & (Tu > 0)
    quality = {"valid_frac": float(found.mean()),
               "within_1px_of_truth": float((np.abs(dl - Tu)[found] <= 1.0).mean()) if found.any() else 0.0}

    # ---- e2e: the C-ABI host calls, pinned host buffers, copies inside the timed region
    e2e = None
    if not a.no_e2e:
        pL = torch.from_numpy(Lu).repeat(reps, 1, 1)[:B].contiguous().pin_memory()
        pR = torch.from_numpy(Ru).repeat(reps, 1, 1)[:B].contiguous().pin_memory()
        hL = torch.empty((B, H, W), dtype=torch.float32).pin_memory()
        hR = torch.empty((B, H, W), dtype=torch.float32).pin_memory()

        def timed(fn):
            barrier()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize(dev)
            dt = time.perf_counter() - t0
            td = torch.tensor([dt], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(td, op=dist.ReduceOp.MAX)
            return float(td.item())

        def run_async():       # a stream of batches: every step enqueued, one wait at the end
            for _ in range(a.steps):
                eng.match_batch_host_async(B, pL.data_ptr(), pR.data_ptr(), W, H, W, hL.data_ptr(),
                                           hR.data_ptr(), W * 4, first_pair_index=first)
            eng.wait()

        def run_sync():        # one blocking call per step
            for _ in range(a.steps):
                eng.match_batch_host(B, pL.data_ptr(), pR.data_ptr(), W, H, W, hL.data_ptr(),
                                     hR.data_ptr(), W * 4, first_pair_index=first)

        run_sync()             # warm-up: allocates the host-path device buffers
        dt_async = timed(run_async)
        exact = bool(np.array_equal(hL[:U].numpy(), dl))
        dt_sync = timed(run_sync)
        ceil_s = copy_ceiling(torch, dist, dev, world, pL, pR, hL, hR, dL, dR, oL, oR)
        e2e = {"value": total_pairs * a.steps / dt_async, "unit": UNIT,
               "h2d_bytes_per_step": 2 * B * n_px, "d2h_bytes_per_step": 2 * B * n_px * 4,
               "api": "pm_match_batch_host_async x steps + pm_wait (pinned host buffers; uploads, kernels "
                      "and downloads of consecutive steps overlap)",
               "sync_call_value": total_pairs * a.steps / dt_sync,
               "sync_call_api": "pm_match_batch_host, one blocking call per step",
               "copy_ceiling_pairs_per_s": total_pairs / ceil_s,
               "copy_ceiling_note": "one step's H2D + D2H with no kernels, both directions at once, all "
                                    "ranks together (max over ranks, best of 3)",
               "equals_device_resident_result": exact,
               "checksum": float(hL[:U].sum().item())}

    # ---- config C5 under the same launch: one 3840x2160 frame in row bands over the ranks
    c5 = None
    if world > 1 and not a.no_c5 and 16 % world == 0:
        torch.cuda.set_stream(torch.cuda.default_stream(dev))
        del dL, dR, oL, oR
        torch.cuda.empty_cache()
        c5 = c5_band_leg(a, pkg, torch, dist, rank, local_rank, world)

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
        # pairs one sweep launch covers (the engine picks the pass size): every pass runs
        # 4 sweeps x iters x levels launches
        pairs_per_launch = B * a.steps * 4.0 * a.iters * a.levels / max(sw_n, 1)
        traffic, traffic_src = measured_sweep_traffic(a, pairs_per_launch)
        evals = evals_per_pair(n_px, a.iters, a.levels) * total_pairs * a.steps
        alu_tflops = evals * OPS_PER_EVAL / (ms_max * 1e-3) / 1e12 / world
        roofline = {
            # what ncu shows binding (profiles/r4_ncu_k_sweep_*.md): neither DRAM (28-52 %) nor the
            # FP32 pipes (19-23 %): one in-order warp per dependent chain, 2 warps per scheduler in
            # the row kernel (issue 43-53 %, shared-memory wavefronts 59-71 %), L1 misses of the
            # matched row on the chain in the column kernel (long scoreboard 8.9 cycles / instruction)
            "bound": "hbm",
            "binding_resource": "latency of dependent chains, neither HBM nor FP32 (ncu, profiles/r4_ncu_k_sweep_*.md): "
                                "row sweeps 8 warps/SM, issue-active 43-53 %, LSU/shared wavefront pipe 59-71 %, "
                                "DRAM 28-36 %; column sweeps issue-active 53 %, L1 wavefront pipe 65 %, long-scoreboard "
                                "8.9 cycles per instruction, DRAM 52 %; see DESIGN.md 4.3",
            "kernel": "k_sweep (row + column sweeps)",
            "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": achieved / peaks["hbm_gbs"], "traffic": traffic,
            "traffic_source": traffic_src,
            "algorithmic_bytes_per_launch": sweep_bytes / max(sw_n, 1),
            # the same launches counted with the bytes ncu saw move (planes are float2
            # {I,G} and {d,cost}: 32 B per pixel and view instead of the canonical 18)
            "dram_gbs": (traffic * sw_n / (sw_ms * 1e-3) / 1e9) if traffic and sw_ms > 0 else None,
            "dram_frac": (traffic * sw_n / (sw_ms * 1e-3) / 1e9 / peaks["hbm_gbs"]) if traffic and sw_ms > 0 else None,
            "fp32_frac": alu_tflops / fp32_peak if fp32_peak else None,
            "peak_source": peak_src, "launches": sw_n, "pairs_per_launch": pairs_per_launch,
            "avg_launch_ms": sw_ms / max(sw_n, 1),
            "share_of_step": sw_ms / total_stage_ms if total_stage_ms else None}
        alu = {"evals_per_pair": evals_per_pair(n_px, a.iters, a.levels),
               "ops_per_eval": OPS_PER_EVAL, "achieved_tflops": alu_tflops,
               "peak_tflops": fp32_peak,
               "peak_source": "measured: dependent-free FFMA kernel on this GPU (pm_measure_fp32_peak)",
               "nominal_tflops": 148 * 128 * 2 * peaks.get("sm_max_mhz", 1965.0) * 1e6 / 1e12,
               "frac": alu_tflops / fp32_peak if fp32_peak else None}
        # the streaming stages against the same measured copy bandwidth: canonical algorithmic
        # bytes (u8 images, f32 gradient / disparity / cost planes, each logical plane
        # once) and the bytes the engine's layouts really move (float2 {I,G} and {d,cost} planes,
        # profiles/r4_streaming_kernels.md), over the CUDA-event time of the stage in this run
        lv_px = sum(n_px / 4.0 ** l for l in range(a.levels))
        stage_bytes = {
            # u8 pair in; gradient planes out (canonical 2 + 8); engine: refT + matI of both views,
            # left-to-right and flipped = 4 float2 planes per view
            "preprocess": (10.0 * lv_px, (2.0 + 64.0) * lv_px + (2.5 * n_px if a.levels > 1 else 0.0)),
            # MaskBackground: one sweep-sized pass (36 B/px); engine, per view: {I,G} of the reference
            # and the matched view + {d,cost} read once by the rolling pass (24 B), d written (4 B)
            "mask_bg": (36.0 * n_px, 2.0 * (24.0 + 4.0) * n_px),
            # MaskOcclusions + flip back: both disparity planes in, both out
            "finalize": (16.0 * n_px, 16.0 * n_px),
        }
        roofline_stages = {}
        for k, (canon, moved) in stage_bytes.items():
            ms = stage[k][0]
            if ms > 0:
                roofline_stages[k] = {
                    "ms_per_step": ms / a.steps,
                    "hbm_frac_algorithmic": canon * B * a.steps / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                    "hbm_frac_bytes_moved": moved * B * a.steps / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"]}
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": warm, "ms_per_step": ms_max / a.steps, "higher_is_better": True,
            "scaling": "strong" if a.pairs_per_gpu <= 0 else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(a, world), "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
            "host_cpus_bound_to_gpu_node": numa_cpus,
            "roofline": roofline, "alu": alu, "roofline_streaming_stages": roofline_stages,
            "stage_ms_per_step": {k: v[0] / a.steps for k, v in stage.items()},
            "quality": quality,
        }
        if c5 is not None:
            out["c5_band"] = c5
        if not a.no_cpu_baseline and world == 1:   # the CPU baseline legs run at N = 1 only
            pmo = oracle()
            n = max(1, min(a.cpu_sample_pairs, U))
            p = pmo.default_params(init_mode=1 if a.init == "random" else 0, max_disp=a.max_disp,
                                   pyramid_levels=a.levels, patchmatch_iters=a.iters)
            t0 = time.perf_counter()
            agree = []
            for i in range(n):
                if a.init == "random":
                    wl, wr = pmo.g_match(p, Lu[i], Ru[i], pair_index=first + i)
                else:
                    sl, sr = pmo.s_match_seeds(Lu[i], Ru[i], 4)
                    wl, wr = pmo.g_match(p, Lu[i], Ru[i], sl, sr, pair_index=first + i)
                agree.append(bool(np.array_equal(wl, dl[i])))
            dt = time.perf_counter() - t0
            out["cpu_baseline"] = {"value": n / dt, "unit": UNIT, "cores": 1, "kind": "port",
                                   "sample": "%d pairs of the same workload, one host thread "
                                             "(oracle/pm_oracle.c, -O3 -march=native): the GPU library's "
                                             "algorithm (patchmatch_gpu.cu) restated for the CPU" % n,
                                   "bit_exact_vs_gpu": agree}
            # config C1: the reference's ACTUAL CPU path, stereo::Patchmatch, on its own fixture,
            # beside the same stage library on the GPU (pm_cpu_estimate_disparity)
            g, want, cpu, c1_dt = cpu_c1_leg(pmo)
            pmc = pkg.Patchmatch(device=local_rank)   # its own engine: the reference's default params
            pmc.EstimateDisparity(g["il"], g["ir"])
            t0 = time.perf_counter()
            for _ in range(3):
                got = pmc.EstimateDisparity(g["il"], g["ir"])
            c1_gpu = (time.perf_counter() - t0) / 3
            pmc.close()
            out["cpu_baseline_c1"] = {
                "workload": "C1: stereo::Patchmatch (Initialize + 4 x (AddNoise, Propagate) + RemoveBackground, "
                            "patchmatch_test.cpp:116-188) on fsl1/fsr1 at 376x240, one frame",
                "value": 1.0 / c1_dt, "unit": "frames/s", "cores": 1, "kind": "port",
                "seconds_per_frame": c1_dt, "cpu_equals_cv2_golden": bool(np.array_equal(cpu, want)),
                "gpu_seconds_per_frame": c1_gpu, "gpu_equals_cv2_golden": bool(np.array_equal(got, want))}
        out_line.emit(out)
    if world > 1:
        dist.destroy_process_group()
    eng.close()


if __name__ == "__main__":
    main()
