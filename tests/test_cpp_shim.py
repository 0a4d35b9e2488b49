"""The C++ host class (include/patchmatch_gpu.h): builds with g++ against the C ABI library,
parses the reference-shaped YAML, fails loudly without a GPU, and on a GPU matches the oracle."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

YAML = """%YAML:1.0
PatchmatchGpu:
  init_mode: random
  max_disp: 48
  pyramid_levels: 2
  FeatureDetector:
    max_features_per_frame: 200
    min_distance_btw_tracked_and_detected_features: 20
    gftt_quality_level: 0.01
    gftt_block_size: 5
    gftt_use_harris_corner_detector: 0
  StereoMatcher:
    templ_cols: 31
    templ_rows: 11
    max_disp: 128
    max_matching_cost: 0.15
    bidirectional: 1
    subpixel_refinement: 0
"""


@pytest.fixture(scope="module")
def shim(built_lib, tmp_path_factory):
    d = tmp_path_factory.mktemp("shim")
    exe = str(d / "shim_test")
    libdir = os.path.dirname(built_lib)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "shim_test.cpp"), "-o", exe,
                           "-L", libdir, "-lpm_b200", "-Wl,-rpath," + libdir])
    y = d / "pm.yaml"
    y.write_text(YAML)
    return exe, str(y), d


def test_cpp_params_and_no_gpu_error(shim):
    import torch
    exe, yaml, _ = shim
    out = subprocess.run([exe, "--no-gpu", yaml], capture_output=True, text=True, check=True).stdout
    assert "templ=31x11" in out and "max_disp=128" in out and "levels=2" in out
    assert "yaml error: cannot open /nonexistent.yaml" in out
    if torch.cuda.is_available():
        assert "engine created" in out
    else:
        assert "create failed" in out and "no CPU path" in out


@pytest.mark.gpu
def test_cpp_match_equals_oracle(shim, pmo, pkg):
    exe, yaml, d = shim
    w, h = 640, 400
    L, R, _ = pkg.synth.make_pair(4, w, h, 48)
    L.tofile(d / "l.raw"); R.tofile(d / "r.raw")
    subprocess.run([exe, yaml, str(w), str(h), str(d / "l.raw"), str(d / "r.raw"),
                    str(d / "dl.raw"), str(d / "dr.raw")], check=True)
    dl = np.fromfile(d / "dl.raw", np.float32).reshape(h, w)
    dr = np.fromfile(d / "dr.raw", np.float32).reshape(h, w)
    wl, wr = pmo.g_match(pmo.default_params(init_mode=1, max_disp=48, pyramid_levels=2), L, R)
    assert np.array_equal(dl, wl) and np.array_equal(dr, wr)


@pytest.mark.gpu
def test_cpp_match_band_single_band(shim, pkg):
    exe, yaml, d = shim
    w, h = 416, 266
    L, R, _ = pkg.synth.make_pair(2, w, h, 48)
    L.tofile(d / "bl.raw"); R.tofile(d / "br.raw")
    out = subprocess.run([exe, "--band", yaml, str(w), str(h), str(d / "bl.raw"), str(d / "br.raw")],
                         capture_output=True, text=True, check=True).stdout
    assert "band ok" in out and "does not divide sweep_chunks" in out


@pytest.mark.gpu
def test_cpp_reference_drivers_with_default_params(shim, pmo, c1, c1_cpu):
    """The reference's two drivers through the C++ classes with default params: Match seeds itself
    (SparseInit on the device), stereo::Patchmatch::EstimateDisparity(il, ir) runs Initialize + the
    test schedule. Against the cv2-literal goldens of the reference's fixture."""
    exe, _, d = shim
    il, ir = c1["il"], c1["ir"]
    h, w = il.shape
    il.tofile(d / "fl.raw"); ir.tofile(d / "fr.raw")
    outs = [str(d / n) for n in ("sd.raw", "sdr.raw", "seed.raw", "cpu.raw")]
    subprocess.run([exe, "--sparse", str(w), str(h), str(d / "fl.raw"), str(d / "fr.raw")] + outs, check=True)
    disp, dispr, seed, cpu = [np.fromfile(o, np.float32).reshape(h, w) for o in outs]
    assert np.array_equal(seed, c1["seed_gpu_l"])                       # cv2 SparseInit
    wl, wr = pmo.g_match(pmo.default_params(), il, ir, c1["seed_gpu_l"], c1["seed_gpu_r"])
    assert np.array_equal(disp, wl) and np.array_equal(dispr, wr)
    assert np.array_equal(cpu, c1_cpu["final"])                         # cv2-literal CPU pipeline
