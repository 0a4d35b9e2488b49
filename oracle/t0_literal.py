"""T0: literal Python/cv2 transliteration of the reference's CPU PatchMatch path.

TEST INFRASTRUCTURE ONLY. Each function calls the same OpenCV entry point the
reference calls (cv2 4.13 here, OpenCV 3.4.0 there) in the same order, so its
outputs are the closest thing to "the reference run here" this container can
produce. It is slow (one cv2.getRectSubPix per patch) and only runs in
oracle/gen_goldens.py; the committed fixtures under tests/golden/ pin the C
oracle (T1) to it.

Citations are relative to /root/reference.
"""
import cv2
import numpy as np

F32 = np.float32

# Run OpenCV's own C++ kernels, not Intel IPP's: with IPP on, cv2.getRectSubPix goes
# through ippiCopySubpixIntersect, whose rounding differs from OpenCV's published
# fixed-point sampler (measured here: 2066 of 20000 random u8 patches differ by 1 LSB,
# 6819 of 20000 f32 patches in the last bits). The restatement follows the published
# algorithm (imgproc/samplers.cpp), which is what cv2 executes with IPP off.
cv2.ipp.setUseIPP(False)


# ---------------------------------------------------------------- primitives

def rng_uniform(shape, lo, hi, seed=123):
    """cv::RNG rng(seed); rng.fill(m, UNIFORM, lo, hi) -- patchmatch.cpp:146-147."""
    cv2.setRNGSeed(seed)
    m = np.zeros(shape, F32)
    cv2.randu(m, float(lo), float(hi))
    return m


def compute_gradient(im):
    """ComputeGradient, test/stereo_matching/patchmatch_test.cpp:48-64."""
    dx = cv2.Sobel(im, cv2.CV_32F, 1, 0, ksize=3)
    dy = cv2.Sobel(im, cv2.CV_32F, 0, 1, ksize=3)
    dx = cv2.pow(dx, 2)
    dy = cv2.pow(dy, 2)
    return cv2.sqrt(dx + dy)


def get_patch_subpix(im, x, y, pw, ph):
    """GetPatchSubpix, patchmatch.cpp:98-111."""
    return cv2.getRectSubPix(im, (pw, ph), (float(x), float(y)))


def l1_cost(pl, pr):
    """L1CostFunction<ImageT>, patchmatch_test.cpp:20-26."""
    return F32(cv2.mean(cv2.absdiff(pl, pr))[0])


def to_image1b(g):
    """Implicit cv::Mat_<float> -> cv::Mat_<uchar> at the functor boundary
    (patchmatch_test.cpp:30-33 takes Image1b, patchmatch.hpp:18 passes Image1f)."""
    return np.clip(np.rint(g), 0, 255).astype(np.uint8)  # cvRound (half-even) + saturate


def l1_gradient_cost(pl, pr, gl, gr):
    """L1GradientCostFunction, patchmatch_test.cpp:30-45."""
    alpha = F32(0.7)
    tau_color = F32(50.0)
    tau_grad = F32(20.0)
    gl8 = to_image1b(gl)
    gr8 = to_image1b(gr)
    error_color = min(l1_cost(pl, pr), tau_color)
    # L1CostFunction<Image1f>(gl, gr): the u8 patches are converted back to f32
    error_grad = min(l1_cost(gl8.astype(F32), gr8.astype(F32)), tau_grad)
    return F32(alpha * error_color + (F32(1) - alpha) * error_grad)


# ------------------------------------------------------- stereo::Patchmatch

def add_noise(disp, amount):
    """Patchmatch::AddNoise with mask = disp > 0, patchmatch.cpp:143-155."""
    noise = rng_uniform(disp.shape, -amount, amount, 123)
    mask = (disp > 0).astype(np.uint8) * 255
    out = disp.copy()
    cv2.add(disp, noise, out, mask=mask)
    return cv2.max(out, 0)


def propagate_neighbors(iml, imr, Gl, Gr, x, y, disp, pw, ph, x_off, y_off):
    """PropagateNeighbors (1-neighbour overload), patchmatch.cpp:158-196."""
    ref = get_patch_subpix(iml, x, y, pw, ph)
    gref = get_patch_subpix(Gl, x, y, pw, ph)
    d0 = disp[y, x]
    d0 = F32(min(max(d0, F32(0)), F32(x) - F32(pw // 2)))
    dl = disp[y + y_off, x + x_off]
    p0 = get_patch_subpix(imr, F32(x) - d0, y, pw, ph)
    g0 = get_patch_subpix(Gr, F32(x) - d0, y, pw, ph)
    costs = [l1_gradient_cost(ref, p0, gref, g0)]
    cands = [d0]
    if (F32(x) - dl) >= (pw // 2):
        pl = get_patch_subpix(imr, F32(x) - dl, y, pw, ph)
        gl = get_patch_subpix(Gr, F32(x) - dl, y, pw, ph)
        costs.append(l1_gradient_cost(ref, pl, gref, gl))
        cands.append(dl)
    best = 0
    for i in range(1, len(costs)):  # Argmin, patchmatch.cpp:115-126
        if costs[i] < costs[best]:
            best = i
    disp[y, x] = cands[best]


def _border(x, y, w, h, pw, ph):
    return y < ph // 2 or x < pw // 2 or y > h - ph // 2 - 1 or x > w - pw // 2 - 1


def propagate_pass(iml, imr, Gl, Gr, disp, ph, pw, which):
    """One of the four raster passes of Patchmatch::Propagate, patchmatch.cpp:264-310."""
    h, w = iml.shape
    if which in (0, 1):
        xo, yo = (-1, 0) if which == 0 else (0, -1)
        for y in range(1, h):
            for x in range(1, w):
                if _border(x, y, w, h, pw, ph):
                    continue
                propagate_neighbors(iml, imr, Gl, Gr, x, y, disp, pw, ph, xo, yo)
    else:
        xo, yo = (1, 0) if which == 2 else (0, 1)
        for y in range(h - 2, -1, -1):
            for x in range(w - 2, -1, -1):
                if _border(x, y, w, h, pw, ph):
                    continue
                propagate_neighbors(iml, imr, Gl, Gr, x, y, disp, pw, ph, xo, yo)


def propagate(iml, imr, Gl, Gr, disp, ph, pw):
    for which in range(4):
        propagate_pass(iml, imr, Gl, Gr, disp, ph, pw, which)


def remove_background(iml, imr, Gl, Gr, disp, ph, pw, win_by_factor):
    """Patchmatch::RemoveBackground, patchmatch.cpp:314-360."""
    h, w = iml.shape
    for y in range(1, h):
        for x in range(1, w):
            if _border(x, y, w, h, pw, ph):
                continue
            ref = get_patch_subpix(iml, x, y, pw, ph)
            gref = get_patch_subpix(Gl, x, y, pw, ph)
            d0 = disp[y, x]
            d0 = F32(min(max(d0, F32(0)), F32(x) - F32(pw // 2)))
            p0 = get_patch_subpix(imr, F32(x) - d0, y, pw, ph)
            g0 = get_patch_subpix(Gr, F32(x) - d0, y, pw, ph)
            cost_cur = l1_gradient_cost(ref, p0, gref, g0)
            pb = get_patch_subpix(imr, x, y, pw, ph)
            gb = get_patch_subpix(Gr, x, y, pw, ph)
            cost_zero = l1_gradient_cost(ref, pb, gref, gb)
            if cost_cur > F32(cost_zero / F32(win_by_factor)):
                disp[y, x] = 0


# ------------------------------------------------------------------ seeding

def detect_gftt(img, max_features=200, quality=0.01, min_dist=20, block=5, harris=False, k=0.04):
    """FeatureDetector::Detect with no tracked keypoints, feature_detector.cpp:89-122.
    ANMS early-returns because #kp <= max_features (feature_detector.cpp:67-69)."""
    det = cv2.GFTTDetector_create(max_features, quality, min_dist, block, harris, k)
    mask = np.full(img.shape, 255, np.uint8)
    kps = det.detect(img, mask)
    return [kp.pt for kp in kps]


def match_rectified(left, right, kp, templ_cols=31, templ_rows=11, max_disp=128, max_cost=0.15):
    """StereoMatcher::MatchRectified, feature_tracking/stereo_matcher.cpp:22-116
    (subpixel_refinement = false)."""
    stripe_rows = templ_rows + 2
    # C round(): half away from zero
    rx = int(np.floor(abs(kp[0]) + 0.5) * np.sign(kp[0])) if kp[0] != 0 else 0
    ry = int(np.floor(abs(kp[1]) + 0.5) * np.sign(kp[1])) if kp[1] != 0 else 0
    ty = ry - (templ_rows - 1) // 2
    if ty < 0 or (ty + templ_rows) >= left.shape[0]:
        return -1.0
    offset_x = 0
    tx = rx - (templ_cols - 1) // 2
    if tx < 0:
        offset_x = tx
        tx = 0
    if (tx + templ_cols) >= left.shape[1]:
        assert offset_x == 0
        offset_x = (tx + templ_cols) - (left.shape[1] - 1)
        tx -= offset_x
    patch = left[ty:ty + templ_rows, tx:tx + templ_cols]
    sy = ry - (stripe_rows - 1) // 2
    if sy < 0 or (sy + stripe_rows) >= right.shape[0]:
        return -1.0
    sx = rx + (templ_cols - 1) // 2 - max_disp
    if sx + max_disp > right.shape[1] - 1:
        sx -= (sx + max_disp) - (right.shape[1] - 1)
    if sx < 0:
        sx = 0
    stripe = right[sy:sy + stripe_rows, sx:sx + max_disp]
    result = cv2.matchTemplate(stripe, patch, cv2.TM_SQDIFF_NORMED)
    min_val, _, min_loc, _ = cv2.minMaxLoc(result)
    mx = min_loc[0] + sx + (templ_cols - 1) // 2 + offset_x
    if min_val < max_cost and kp[0] >= float(mx):
        return float(F32(kp[0]) - F32(mx))
    return -1.0


def _scatter_seeds(shape, kps, disps):
    seeds = np.zeros(shape, F32)
    for kp, d in zip(kps, disps):
        if d >= 0:  # patchmatch_gpu.cu:431, patchmatch.cpp:69
            yy = int(np.floor(kp[1] + 0.5))
            xx = int(np.floor(kp[0] + 0.5))
            seeds[yy, xx] = F32(d)
    return seeds


def sparse_init_gpu(iml, imr, dilate_factor=4, **kw):
    """PatchmatchGpu::SparseInit, patchmatch_gpu.cu:414-442."""
    kps = detect_gftt(iml)
    disps = [match_rectified(iml, imr, kp, **kw) for kp in kps]
    seeds = _scatter_seeds(iml.shape, kps, disps)
    ds = int(2 ** dilate_factor) + 1
    el = cv2.getStructuringElement(cv2.MORPH_RECT, (2 * ds + 1, 2 * ds + 1), (ds, ds))
    return cv2.dilate(seeds, el), kps, disps


def initialize_cpu(iml, imr, downsample_factor=1, **kw):
    """Patchmatch::Initialize, patchmatch.cpp:52-87."""
    kps = detect_gftt(iml)
    disps = [match_rectified(iml, imr, kp, **kw) for kp in kps]
    seeds = _scatter_seeds(iml.shape, kps, disps)
    ds = int(2 ** (downsample_factor - 1)) + 1
    el = cv2.getStructuringElement(cv2.MORPH_RECT, (2 * ds + 1, 2 * ds + 1), (ds, ds))
    seeds = cv2.dilate(seeds, el)
    h, w = seeds.shape
    seeds = cv2.resize(seeds, (w // downsample_factor, h // downsample_factor),
                       interpolation=cv2.INTER_NEAREST)
    seeds = (seeds / F32(2 ** downsample_factor)).astype(F32)
    return seeds, kps, disps


def load_fixture_pair(left_path, right_path):
    """imread GRAYSCALE + cv::resize(size/2), patchmatch_test.cpp:121-133."""
    il = cv2.imread(left_path, cv2.IMREAD_GRAYSCALE)
    ir = cv2.imread(right_path, cv2.IMREAD_GRAYSCALE)
    il = cv2.resize(il, (il.shape[1] // 2, il.shape[0] // 2))
    ir = cv2.resize(ir, (ir.shape[1] // 2, ir.shape[0] // 2))
    return il, ir
