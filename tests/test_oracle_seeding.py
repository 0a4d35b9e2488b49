"""The oracle's restatement of the reference's sparse seeding step (oracle/pm_oracle_seed.c:
FeatureDetector::Detect, StereoMatcher::MatchRectified, PatchmatchGpu::SparseInit,
Patchmatch::Initialize) against cv2-literal goldens made on the reference's own stereo
fixtures (oracle/gen_goldens_seeding.py, oracle/gen_goldens.py). CPU only."""
import os

import numpy as np
import pytest

from conftest import GOLDEN


@pytest.fixture(scope="module")
def sg():
    return dict(np.load(os.path.join(GOLDEN, "seeding.npz")))


PAIRS = ["fs1", "farm", "caddy", "vk"]


@pytest.mark.parametrize("name", PAIRS)
def test_keypoints_match_cv2(pmo, sg, name):
    # same keypoints in the same (selection) order as cv::GFTTDetector, on both views'
    # reference images (patchmatch_gpu.cu:335 and :362-365)
    il, ir = sg[name + "_il"], sg[name + "_ir"]
    k, ncand = pmo.s_good_features(il)
    assert ncand >= len(k)
    assert np.array_equal(k, sg[name + "_kps"].astype(np.int64))
    kr, _ = pmo.s_good_features(np.ascontiguousarray(ir[:, ::-1]))
    assert np.array_equal(kr, sg[name + "_kps_r"].astype(np.int64))


@pytest.mark.parametrize("name", PAIRS)
def test_match_rectified_matches_cv2(pmo, sg, name):
    il, ir = sg[name + "_il"], sg[name + "_ir"]
    d = pmo.s_match_rectified(il, ir, sg[name + "_kps"].astype(np.int64))
    assert np.array_equal(d, sg[name + "_disps"])
    ilf, irf = np.ascontiguousarray(il[:, ::-1]), np.ascontiguousarray(ir[:, ::-1])
    dr = pmo.s_match_rectified(irf, ilf, sg[name + "_kps_r"].astype(np.int64))
    assert np.array_equal(dr, sg[name + "_disps_r"])
    assert (d >= 0).sum() > 0.6 * d.size  # the fixtures do match


@pytest.mark.parametrize("name", PAIRS)
def test_sparse_init_matches_cv2(pmo, sg, name):
    il, ir = sg[name + "_il"], sg[name + "_ir"]
    sl, sr = pmo.s_match_seeds(il, ir, 4)
    assert np.array_equal(sl, sg[name + "_seed_l"])
    assert np.array_equal(sr, sg[name + "_seed_r"])
    assert np.array_equal(pmo.s_sparse_init(il, ir, 4), sg[name + "_seed_l"])


def test_c1_fixture_seeds(pmo, c1):
    # the goldens of oracle/gen_goldens.py (config C1), incl. Patchmatch::Initialize
    il, ir = c1["il"], c1["ir"]
    k, _ = pmo.s_good_features(il)
    assert np.array_equal(k, c1["kps"].astype(np.int64))
    assert np.array_equal(pmo.s_match_rectified(il, ir, k), c1["kp_disps"])
    sl, sr = pmo.s_match_seeds(il, ir, 4)
    assert np.array_equal(sl, c1["seed_gpu_l"]) and np.array_equal(sr, c1["seed_gpu_r"])
    assert np.array_equal(pmo.c_initialize(il, ir, 1), c1["seed_cpu"])


def test_detector_param_variants(pmo, c1, sg):
    il = c1["il"]
    v1 = pmo.seed_params(max_features=50, quality_level=0.05, min_distance=10, block_size=3)
    assert np.array_equal(pmo.s_good_features(il, v1)[0], sg["v1_kps"].astype(np.int64))
    v2 = pmo.seed_params(use_harris=1, harris_k=0.04)
    assert np.array_equal(pmo.s_good_features(il, v2)[0], sg["v2_kps"].astype(np.int64))
    v3 = pmo.seed_params(max_features=400, min_distance=7, block_size=7)
    assert np.array_equal(pmo.s_good_features(il, v3)[0], sg["v3_kps"].astype(np.int64))
    assert len(sg["v1_kps"]) == 50 and len(sg["v3_kps"]) > 200


def test_matcher_param_variants(pmo, c1, sg):
    il, ir = c1["il"], c1["ir"]
    k = c1["kps"].astype(np.int64)
    v4 = pmo.seed_params(templ_cols=21, templ_rows=7, max_disp=64, max_matching_cost=0.1)
    assert np.array_equal(pmo.s_match_rectified(il, ir, k, v4), sg["v4_disps"])
    v5 = pmo.seed_params(templ_cols=41, templ_rows=15, max_disp=200, max_matching_cost=0.3)
    assert np.array_equal(pmo.s_match_rectified(il, ir, k, v5), sg["v5_disps"])


def test_initialize_and_dilate_variants(pmo, c1, sg):
    il, ir = c1["il"], c1["ir"]
    for f in (1, 2, 4):   # Patchmatch::Initialize(downsample_factor), patchmatch.cpp:75-81
        assert np.array_equal(pmo.c_initialize(il, ir, f), sg["init_f%d" % f]), f
    assert np.array_equal(pmo.s_sparse_init(il, ir, 2), sg["sparse_f2"])


def test_response_map_close_to_cv2(pmo, c1, sg):
    # OpenCV evaluates the response in float32 with SIMD-dependent rounding; the oracle is
    # the exact value rounded once. Tolerance: 1e-6 of the map's maximum.
    il = c1["il"]
    rows = sg["resp_rows"]
    e = pmo.s_corner_response(il, 5, False)
    assert np.abs(e[rows] - sg["resp_eig"]).max() <= 1e-6 * sg["resp_eig_max"]
    hmap = pmo.s_corner_response(il, 5, True, 0.04)
    assert np.abs(hmap[rows] - sg["resp_harris"]).max() <= 1e-5 * sg["resp_harris_max"]


def test_seeding_edge_cases(pmo):
    # flat image: response 0 everywhere, no keypoints, all-zero seeds
    flat = np.full((64, 160), 90, np.uint8)
    k, nc = pmo.s_good_features(flat)
    assert len(k) == 0 and nc == 0
    assert not pmo.s_sparse_init(flat, flat, 4).any()
    # keypoints near the top/bottom edge return -1 (stereo_matcher.cpp:35-37, 65-67); a
    # keypoint near the left edge shifts its template inward (:42-45)
    rng = np.random.default_rng(5)
    im = rng.integers(0, 256, (80, 200), dtype=np.uint8)
    d = pmo.s_match_rectified(im, im, [(100, 3), (100, 76), (100, 40), (5, 40), (196, 40)])
    assert d[0] == -1 and d[1] == -1
    assert d[2] == 0 and d[3] == 0 and d[4] == 0   # identical images: zero disparity
