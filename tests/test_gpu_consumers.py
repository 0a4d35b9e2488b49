"""The callers and data formats either side of the path, on the GPU (-m gpu):
ForegroundTextureMask against its cv2 golden, the mesher-facing vertex adapter against the oracle's
DispToDepth/Backproject, the PatchmatchGpuTest.Sequence driver over an EuRoC-layout folder, and the
reference's SGBM wrapper as the second quality baseline."""
import importlib
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def test_foreground_texture_mask_equals_cv2_golden(pmo, engine_factory, c1):
    """patchmatch.cpp:19-49 on the device == cv2 (golden made by oracle/gen_goldens_texture_mask.py)."""
    g = dict(np.load(os.path.join(GOLDEN, "texture_mask.npz")))
    e = engine_factory()
    for i, (k, mg, d) in enumerate(g["cases"]):
        got = e.ForegroundTextureMask(c1["il"], int(k), float(mg), int(d))
        assert np.array_equal(got, g["mask%d" % i]), (k, mg, d, int((got != g["mask%d" % i]).sum()))
    # a larger, synthetic image against the oracle
    pkg = importlib.import_module("ocean-perception_b200")
    L, _, _ = pkg.synth.make_pair(1, 1280, 720, 128)
    assert np.array_equal(e.ForegroundTextureMask(L), pmo.c_foreground_texture_mask(L))
    for bad in ((2, 35.0, 2), (7, 35.0, 9)):        # the reference CHECK-fails on these
        with pytest.raises(pkg.PmError) as ei:
            e.ForegroundTextureMask(c1["il"], *bad)
        assert ei.value.code == -1


def test_mesh_vertices_from_dense_disparity(pmo, pkg, engine_factory):
    """ObjectMesher::BuildTriangleMesh's per-vertex arithmetic (object_mesher.cpp:139-150) on keypoints
    sampled from the dense map == the oracle's dense DispToDepth/Backproject at those pixels."""
    L, R, _ = pkg.synth.make_pair(3, 672, 376, 64)          # ZED Mini VGA size (config/shared/ZEDMini.yaml)
    e = engine_factory(init_mode="random", max_disp=64)
    dl, _ = e.Match(L, R, pair_index=3)
    rig = pkg.StereoCamera(fx=338.0, fy=338.0, cx=336.0, cy=188.0, baseline=0.063, height=752, width=1344)
    scale = 376.0 / 752.0
    rng = np.random.default_rng(0)
    kp = np.stack([rng.integers(0, 672, 300), rng.integers(0, 376, 300)], 1).astype(np.float32)
    mask = e.ForegroundTextureMask(L)
    vd, xyz = e.MeshVertices(dl, kp, rig, mask=mask)
    depth, dense = pmo.x_disp_to_depth(dl, rig.fx, rig.fy, rig.cx, rig.cy, rig.baseline, scale)
    xs, ys = kp[:, 0].astype(int), kp[:, 1].astype(int)
    gate = mask[ys, xs] > 0
    want_d = np.where(gate, dl[ys, xs], 0.0).astype(np.float32)
    assert np.array_equal(vd, want_d)
    want_xyz = np.where((gate & (want_d > 0))[:, None], dense[ys, xs], 0.0).astype(np.float32)
    assert np.array_equal(xyz, want_xyz)
    assert (vd > 0).sum() > 50


def test_sequence_over_euroc_layout(pmo, pkg, engine_factory, tmp_path):
    """PatchmatchGpuTest.Sequence (patchmatch_gpu_test.cpp:95-138): EurocDataset playback, gray, half
    size, Match - here over a synthetic EuRoC tree, every frame compared with the oracle."""
    ds = pkg.dataset
    D = 32
    frames = []
    for i in range(3):
        L, R, _ = pkg.synth.make_pair(20 + i, 640, 416, 2 * D)
        frames.append((np.stack([L, L, L], -1), np.stack([R, R, R], -1)))   # colour PNGs: exercises BGR2GRAY
    ds.write_euroc_sequence(str(tmp_path / "zed_dataset"), frames, dt_ns=10_000_000)
    dataset = ds.EurocDataset(str(tmp_path / "zed_dataset"))
    e = engine_factory()                                   # the test's params are the defaults but max_disp
    got = []

    def stereo_cb(ts, left, right):
        iml = ds.resize_half(ds.maybe_convert_to_gray(left))
        imr = ds.resize_half(ds.maybe_convert_to_gray(right))
        got.append((iml, imr) + e.Match(iml, imr))

    dataset.RegisterStereoCallback(stereo_cb)
    assert dataset.Playback(20.0) == 3
    for iml, imr, dl, dr in got:
        assert iml.shape == (208, 320)
        sl, sr = pmo.s_match_seeds(iml, imr, 4)
        wl, wr = pmo.g_match(pmo.default_params(), iml, imr, sl, sr)
        assert np.array_equal(dl, wl) and np.array_equal(dr, wr)


def test_sgbm_comparator(pkg, engine_factory):
    """The reference's second stereo path, stereo::EstimateDisparity = cv::StereoSGBM (stereo_matching.cpp:
    11-41), as a quality yardstick on a synthetic pair with ground truth (tools/sgbm_compare.py)."""
    try:
        import cv2  # noqa: F401
    except ImportError:
        pytest.skip("cv2 not importable")
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    sg = importlib.import_module("sgbm_compare")
    L, R, T = pkg.synth.make_pair(0, 640, 400, 64)
    ds_ = sg.estimate_disparity_sgbm(L, R, 64, 5)
    e = engine_factory(init_mode="random", max_disp=64)
    dl, _ = e.Match(L, R)
    s_pm, s_sg = sg.score(dl, T), sg.score(ds_, T)
    print("PatchMatch %s | SGBM %s" % (s_pm, s_sg))
    assert s_pm["within_1px"] > 0.95 and s_sg["within_1px"] > 0.9
    assert s_pm["valid_frac"] > 0.6
