"""The reference's sparse seeding step (PatchmatchGpu::SparseInit, patchmatch_gpu.cu:414-442:
FeatureDetector::Detect, StereoMatcher::MatchRectified, dilate) on the GPU through the C ABI,
against the cv2-literal goldens made on the reference's own fixtures and, bit for bit, against
the CPU oracle (run on the B200 box: -m gpu)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

PAIRS = ["fs1", "farm", "caddy", "vk"]


@pytest.fixture(scope="module")
def sg():
    return dict(np.load(os.path.join(GOLDEN, "seeding.npz")))


def _flip(a):
    return np.ascontiguousarray(a[:, ::-1])


def _texture(w, h, seed):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 256, (h, w)).astype(np.int32)
    a = (a + np.roll(a, 1, 1) + np.roll(a, 1, 0) + np.roll(a, -1, 1) + 2) // 4
    return a.astype(np.uint8)


def _set(P, **kw):
    for k, v in kw.items():
        tgt = P.detector_params if hasattr(P.detector_params, k) else (
            P.matcher_params if hasattr(P.matcher_params, k) else P)
        assert hasattr(tgt, k), k
        setattr(tgt, k, v)
    return P


@pytest.fixture()
def eng(pkg, built_lib):
    made = []

    def factory(**kw):
        e = pkg.PatchmatchGpu(_set(pkg.PatchmatchGpu.Params(), **kw), device=0)
        made.append(e)
        return e

    yield factory
    for e in made:
        e.close()


# ------------------------------------------------------------------ stages

@pytest.mark.parametrize("cfg", [dict(), dict(gftt_block_size=3), dict(gftt_block_size=7),
                                 dict(gftt_block_size=4), dict(gftt_block_size=9),
                                 dict(gftt_use_harris_corner_detector=True),
                                 dict(gftt_use_harris_corner_detector=True, gftt_block_size=6, gftt_k=0.06)])
def test_corner_response_bit_exact(pmo, eng, c1, cfg):
    e = eng(**cfg)
    d = e.params.detector_params
    for im in (c1["il"], c1["ir"][:77, :131]):
        got = e.stage_corner_response(im)
        want = pmo.s_corner_response(im, d.gftt_block_size, d.gftt_use_harris_corner_detector, d.gftt_k)
        assert np.array_equal(got, want), cfg


@pytest.mark.parametrize("name", PAIRS)
def test_detect_matches_cv2_and_oracle(pmo, eng, sg, name):
    e = eng()
    for img, key in ((sg[name + "_il"], "_kps"), (_flip(sg[name + "_ir"]), "_kps_r")):
        k, nc = e.stage_detect(img)
        wk, wnc = pmo.s_good_features(img)
        assert np.array_equal(k, wk) and nc == wnc
        assert np.array_equal(k, sg[name + key].astype(np.int64))   # cv::GFTTDetector, same order


def test_detect_param_variants(pmo, eng, c1, sg):
    il = c1["il"]
    k, _ = eng(max_features_per_frame=50, gftt_quality_level=0.05,
               min_distance_btw_tracked_and_detected_features=10, gftt_block_size=3).stage_detect(il)
    assert np.array_equal(k, sg["v1_kps"].astype(np.int64))
    k, _ = eng(gftt_use_harris_corner_detector=True).stage_detect(il)
    assert np.array_equal(k, sg["v2_kps"].astype(np.int64))
    k, _ = eng(max_features_per_frame=400, min_distance_btw_tracked_and_detected_features=7,
               gftt_block_size=7).stage_detect(il)
    assert np.array_equal(k, sg["v3_kps"].astype(np.int64))
    # no minimum distance: the strongest local maxima in order
    e = eng(max_features_per_frame=1000, min_distance_btw_tracked_and_detected_features=0)
    k, nc = e.stage_detect(il)
    wk, wnc = pmo.s_good_features(il, pmo.seed_params(max_features=1000, min_distance=0))
    assert np.array_equal(k, wk) and nc == wnc and len(k) == 1000


def test_detect_many_candidates_global_sort(pmo, eng):
    """A 1280x720 texture has far more local maxima than the shared-memory sort holds."""
    img = _texture(1280, 720, 3)
    e = eng()
    k, nc = e.stage_detect(img)
    wk, wnc = pmo.s_good_features(img)
    assert nc == wnc and nc > 8192
    assert np.array_equal(k, wk) and len(k) == 200
    flat = np.full((100, 300), 77, np.uint8)
    k, nc = e.stage_detect(flat)
    assert len(k) == 0 and nc == 0
    # a minimum distance so large that the strongest ~2000 candidates cannot fill max_features:
    # the selection falls back to sorting the whole list in global memory
    e2 = eng(min_distance_btw_tracked_and_detected_features=300)
    k, nc = e2.stage_detect(img)
    wk, wnc = pmo.s_good_features(img, pmo.seed_params(min_distance=300))
    assert nc == wnc and np.array_equal(k, wk) and 3 <= len(k) < 40


@pytest.mark.parametrize("name", PAIRS)
def test_match_rectified_matches_cv2(pmo, eng, sg, name):
    e = eng()
    il, ir = sg[name + "_il"], sg[name + "_ir"]
    d = e.stage_match_rectified(il, ir, sg[name + "_kps"].astype(np.int32))
    assert np.array_equal(d, sg[name + "_disps"])
    d = e.stage_match_rectified(_flip(ir), _flip(il), sg[name + "_kps_r"].astype(np.int32))
    assert np.array_equal(d, sg[name + "_disps_r"])


def test_match_rectified_variants_and_borders(pmo, eng, c1, sg):
    il, ir = c1["il"], c1["ir"]
    k = c1["kps"].astype(np.int32)
    d = eng(templ_cols=21, templ_rows=7, max_disp=64, max_matching_cost=0.1).stage_match_rectified(il, ir, k)
    assert np.array_equal(d, sg["v4_disps"])
    d = eng(templ_cols=41, templ_rows=15, max_disp=200, max_matching_cost=0.3).stage_match_rectified(il, ir, k)
    assert np.array_equal(d, sg["v5_disps"])
    # every pixel of a coarse grid incl. all four borders, more keypoints than max_features
    h, w = il.shape
    grid = np.array([(x, y) for y in list(range(0, 16)) + list(range(h - 16, h)) + [h // 2]
                     for x in list(range(0, w, 7)) + [w - 1]], np.int32)
    assert len(grid) > 200
    assert np.array_equal(eng().stage_match_rectified(il, ir, grid), pmo.s_match_rectified(il, ir, grid))


@pytest.mark.parametrize("name", PAIRS)
def test_sparse_init_matches_cv2(pmo, eng, sg, name):
    e = eng()
    il, ir = sg[name + "_il"], sg[name + "_ir"]
    assert np.array_equal(e.SparseInit(il, ir, 4), sg[name + "_seed_l"])
    # the right view as the reference calls it (patchmatch_gpu.cu:362-365)
    assert np.array_equal(_flip(e.SparseInit(_flip(ir), _flip(il), 4)), sg[name + "_seed_r"])


def test_initialize_matches_cv2(pkg, eng, c1, sg):
    pmc = pkg.Patchmatch(eng())
    for f in (1, 2, 4):
        assert np.array_equal(pmc.Initialize(c1["il"], c1["ir"], f), sg["init_f%d" % f]), f
    assert np.array_equal(pmc.Initialize(c1["il"], c1["ir"], 1), c1["seed_cpu"])
    assert np.array_equal(eng().SparseInit(c1["il"], c1["ir"], 2), sg["sparse_f2"])


# ------------------------------------------------------------ whole pipeline

def test_match_with_device_seeding_on_the_fixture(pmo, eng, c1):
    """PatchmatchGpu::Match(iml, imr, disp, dispr) with the reference's defaults: SparseInit for
    both views on the device, then the iterations. Equal to the oracle fed with the cv2 seeds."""
    e = eng()
    dl, dr = e.Match(c1["il"], c1["ir"])
    wl, wr = pmo.g_match(pmo.default_params(), c1["il"], c1["ir"], c1["seed_gpu_l"], c1["seed_gpu_r"])
    assert np.array_equal(dl, wl) and np.array_equal(dr, wr)
    assert (dl > 0).sum() > 5000
    # and equal to passing the same seeds explicitly
    dl2, dr2 = e.Match(c1["il"], c1["ir"], c1["seed_gpu_l"], c1["seed_gpu_r"])
    assert np.array_equal(dl, dl2) and np.array_equal(dr, dr2)


@pytest.mark.parametrize("levels", [1, 2])
def test_batch_with_device_seeding(pkg, pmo, eng, levels):
    w, h, D, n = 640, 400, 64, 5
    pairs = [pkg.synth.make_pair(i, w, h, D) for i in range(n)]
    L = np.stack([p[0] for p in pairs]); R = np.stack([p[1] for p in pairs])
    e = eng(pyramid_levels=levels, max_batch=2)   # 3 device passes: 2 + 2 + 1
    outl, outr = e.MatchBatch(L, R)
    p = pmo.default_params(pyramid_levels=levels)
    hits = 0
    for i in range(n):
        sl, sr = pmo.s_match_seeds(L[i], R[i], 4)
        wl, wr = pmo.g_match(p, L[i], R[i], sl, sr)
        assert np.array_equal(outl[i], wl) and np.array_equal(outr[i], wr), i
        hits += int((sl > 0).sum())
    assert hits > 0


def test_estimate_disparity_without_a_seed(pkg, eng, c1, c1_cpu):
    """stereo::Patchmatch::EstimateDisparity(iml, imr) end to end (Initialize + schedule) ==
    the cv2-literal T0 on the reference's fixture."""
    pmc = pkg.Patchmatch(eng())
    assert np.array_equal(pmc.EstimateDisparity(c1["il"], c1["ir"]), c1_cpu["final"])


def test_seeding_error_paths(pkg, eng, c1):
    il, ir = c1["il"], c1["ir"]
    with pytest.raises(pkg.PmError) as ei:
        eng(subpixel_refinement=True).SparseInit(il, ir, 4)
    assert ei.value.code == -2 and "subpixel_refinement" in str(ei.value)
    with pytest.raises(pkg.PmError) as ei:
        eng(max_disp=512).SparseInit(il, ir, 4)          # stripe wider than the image
    assert ei.value.code == -2
    with pytest.raises(pkg.PmError) as ei:
        eng(max_features_per_frame=5000).SparseInit(il, ir, 4)
    assert ei.value.code == -2
    with pytest.raises(pkg.PmError) as ei:
        eng(templ_cols=200).SparseInit(il, ir, 4)        # template wider than the stripe
    assert ei.value.code == -1
    # an engine that failed a seeding call still works
    e = eng(subpixel_refinement=True, init_mode="random", max_disp=32)
    with pytest.raises(pkg.PmError):
        e.SparseInit(il, ir, 4)
    dl, _ = e.Match(il, ir)
    assert dl.shape == il.shape


def test_disp_to_depth_and_points(pkg, pmo, eng, c1):
    """StereoCamera::DispToDepth / PinholeCamera::Backproject per pixel with the ZED
    Mini rig of config/shared/ZEDMini.yaml:39-60 and ObjectMesher's resolution scaling."""
    e = eng()
    rig = pkg.StereoCamera(336.135986, 336.135986, 317.032654, 178.710770, 0.062939, height=376, width=672)
    disp = c1["seed_gpu_l"]                       # integers with large zero areas
    rng = np.random.default_rng(2)
    disp = (disp + rng.uniform(0, 0.9, disp.shape).astype(np.float32) * (disp > 0)).astype(np.float32)
    depth, xyz = e.DispToDepth(disp, rig, want_points=True)
    scale = disp.shape[0] / 376.0
    wd, wx = pmo.x_disp_to_depth(disp, rig.fx, rig.fy, rig.cx, rig.cy, rig.baseline, scale)
    assert np.array_equal(depth, wd) and np.array_equal(xyz, wx)
    assert np.all(depth[disp <= 0] == 0) and np.all(depth[disp > 0] > 0)
    # against the scalar reference formulas in double (closed-form K^-1: 1e-6 relative in float32)
    ys, xs = np.nonzero(disp > 0)
    for y, x in list(zip(ys, xs))[::997]:
        z = rig.DispToDepth(float(disp[y, x]) / scale)
        assert abs(depth[y, x] - z) <= 1e-6 * z
        assert abs(xyz[y, x, 0] - z * (x / scale - rig.cx) / rig.fx) <= 1e-5 * z
        assert abs(xyz[y, x, 1] - z * (y / scale - rig.cy) / rig.fy) <= 1e-5 * z
    with pytest.raises(ValueError):
        rig.DispToDepth(0.0)                      # the reference CHECK-fails (stereo_camera.cpp:51)
    with pytest.raises(pkg.PmError):
        e.DispToDepth(disp, pkg.StereoCamera(0.0, 1.0, 0.0, 0.0, 0.1))


@pytest.mark.parametrize("w,h", [(211, 97), (333, 205), (258, 130), (640, 199), (1001, 64), (136, 260)])
def test_ragged_sizes_against_the_oracle(pkg, pmo, eng, w, h):
    """Odd widths and heights (tile edges of every seeding kernel, the scalar preprocess path, images
    barely wider than the search stripe): keypoints, matches, seed maps and -- where the sweep
    schedule allows the size -- the whole Match equal the oracle."""
    L, R, _ = pkg.synth.make_pair(w * 7 + h, w, h, 48)
    e = eng()
    k, nc = e.stage_detect(L)
    wk, wnc = pmo.s_good_features(L)
    assert np.array_equal(k, wk) and nc == wnc
    assert np.array_equal(e.stage_corner_response(R), pmo.s_corner_response(R, 5, False))
    assert np.array_equal(e.stage_match_rectified(L, R, wk), pmo.s_match_rectified(L, R, wk))
    sl, sr = pmo.s_match_seeds(L, R, 4)
    assert np.array_equal(e.SparseInit(L, R, 4), sl)
    assert np.array_equal(_flip(e.SparseInit(_flip(R), _flip(L), 4)), sr)
    if w // 16 >= 12 and h // 16 >= 12:
        dl, dr = e.Match(L, R)
        wl, wr = pmo.g_match(pmo.default_params(), L, R, sl, sr)
        assert np.array_equal(dl, wl) and np.array_equal(dr, wr)


def test_disp_to_depth_device_batch(pkg, pmo, eng):
    """The device-pointer variant over a batch of maps with a row stride, on a caller's stream."""
    import torch
    e = eng()
    rig = pkg.StereoCamera(336.135986, 336.135986, 317.032654, 178.710770, 0.062939, height=376, width=672)
    n, h, w, stride = 3, 120, 200, 208
    rng = np.random.default_rng(4)
    disp = (rng.uniform(0, 60, (n, h, stride)) * (rng.uniform(0, 1, (n, h, stride)) > 0.3)).astype(np.float32)
    dev = torch.device("cuda", 0)
    d_disp = torch.from_numpy(disp).to(dev)
    d_depth = torch.full((n, h, stride), -1.0, dtype=torch.float32, device=dev)
    d_xyz = torch.zeros((n, h, w, 3), dtype=torch.float32, device=dev)
    st = torch.cuda.Stream(dev)
    with torch.cuda.stream(st):
        e.disp_to_depth_device(n, d_disp.data_ptr(), w, h, stride * 4, rig, h / 376.0, d_depth.data_ptr(),
                               stride * 4, d_xyz.data_ptr(), stream=st.cuda_stream)
    st.synchronize()
    depth, xyz = d_depth.cpu().numpy(), d_xyz.cpu().numpy()
    for i in range(n):
        wd, wx = pmo.x_disp_to_depth(disp[i, :, :w], rig.fx, rig.fy, rig.cx, rig.cy, rig.baseline, h / 376.0)
        assert np.array_equal(depth[i, :, :w], wd) and np.array_equal(xyz[i], wx)
    assert np.all(depth[:, :, w:] == -1.0)   # padding untouched
