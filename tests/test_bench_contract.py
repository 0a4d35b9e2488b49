"""bench.py's contract on a box without a GPU: the reference arm (the reference's CPU algorithm
through the oracle port, all host threads) prints exactly one JSON line with the keys the driver
reads, for both init modes; the CUDA arm refuses to run without a device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SMALL = ["--width", "320", "--height", "208", "--max-disp", "32", "--levels", "1", "--steps", "1",
         "--warmup", "0", "--unique-pairs", "2"]


@pytest.mark.parametrize("init", ["random", "sparse"])
def test_reference_arm_prints_one_json_line(init):
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--init", init]
                       + SMALL, capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "pairs/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1 and d["scaling"] == "strong"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and init in d["config"]["workload"]
    assert d["config"]["pairs_total"] == 512 and "512" in d["config"]["workload"]
    # config C1 sits in the line: the reference's own CPU class on its fixture, equal to the cv2 golden
    assert d["c1_stereo_patchmatch"]["equals_cv2_golden"] is True
    assert d["c1_stereo_patchmatch"]["seconds_per_frame_1core"] > 0


def test_cuda_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + SMALL, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
    assert not [l for l in r.stdout.splitlines() if l.strip().startswith("{")]
