"""Generate tests/golden/seeding.npz: cv2-literal outputs of the reference's sparse
seeding step (oracle/t0_literal.py) on more of the reference's own stereo fixtures.

Run in the authoring container only (needs /root/reference and cv2):
    python oracle/gen_goldens_seeding.py
The GPU box never runs this; tests read the committed .npz.

Per pair `name` (images are stored so the tests do not need /root/reference):
  {name}_il, {name}_ir          grayscale images as the reference's tests load them
  {name}_kps, {name}_kps_r      FeatureDetector::Detect keypoints of the left image and of
                                the flipped right image (patchmatch_gpu.cu:362-365)
  {name}_disps, {name}_disps_r  StereoMatcher::MatchRectified of those keypoints
  {name}_seed_l, {name}_seed_r  PatchmatchGpu::SparseInit maps (right one flipped back)
Variants on the C1 fixture (fsl1/fsr1 at half size): non-default detector / matcher
params, the Harris response, Patchmatch::Initialize with downsample_factor 1, 2 and 4,
and rows of cv2.cornerMinEigenVal / cornerHarris for the tolerance check of the response.
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import t0_literal as t0  # noqa: E402

RES = "/root/reference/test/resources"
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "seeding.npz")
F32 = np.float32


def load(lp, rp, half):
    il = cv2.imread(os.path.join(RES, lp), cv2.IMREAD_GRAYSCALE)
    ir = cv2.imread(os.path.join(RES, rp), cv2.IMREAD_GRAYSCALE)
    if half:
        il = cv2.resize(il, (il.shape[1] // 2, il.shape[0] // 2))
        ir = cv2.resize(ir, (ir.shape[1] // 2, ir.shape[0] // 2))
    return il, ir


def detect(img, **kw):
    return np.array(t0.detect_gftt(img, **kw), F32).reshape(-1, 2)


def main():
    out = {}
    pairs = [
        ("fs1", "images/fsl1.png", "images/fsr1.png", False),       # 752x480: GFTT hits max_features
        ("farm", "farmsim_01_left.png", "farmsim_01_right.png", True),
        ("caddy", "caddy_32_left.jpg", "caddy_32_right.jpg", True),
        ("vk", "images/vkl.jpg", "images/vkr.jpg", True),
    ]
    names = []
    for name, lp, rp, half in pairs:
        il, ir = load(lp, rp, half)
        names.append(name)
        out[name + "_il"], out[name + "_ir"] = il, ir
        sl, kps, disps = t0.sparse_init_gpu(il, ir, 4)
        ilf = np.ascontiguousarray(il[:, ::-1])
        irf = np.ascontiguousarray(ir[:, ::-1])
        srf, kps_r, disps_r = t0.sparse_init_gpu(irf, ilf, 4)
        out[name + "_kps"] = np.array(kps, F32).reshape(-1, 2)
        out[name + "_disps"] = np.array(disps, np.float64)
        out[name + "_kps_r"] = np.array(kps_r, F32).reshape(-1, 2)
        out[name + "_disps_r"] = np.array(disps_r, np.float64)
        out[name + "_seed_l"] = sl
        out[name + "_seed_r"] = np.ascontiguousarray(srf[:, ::-1])
        print(name, il.shape, len(kps), "kps,", sum(d >= 0 for d in disps), "matched;",
              len(kps_r), "/", sum(d >= 0 for d in disps_r), "on the flipped right image")
    out["names"] = np.array(names)

    # ---- variants on the C1 fixture (images are in c1_inputs.npz)
    il, ir = load("images/fsl1.png", "images/fsr1.png", True)
    out["v1_kps"] = detect(il, max_features=50, quality=0.05, min_dist=10, block=3)
    out["v2_kps"] = detect(il, harris=True, k=0.04)
    out["v3_kps"] = detect(il, max_features=400, min_dist=7, block=7)
    kps = t0.detect_gftt(il)
    out["v4_disps"] = np.array([t0.match_rectified(il, ir, kp, templ_cols=21, templ_rows=7, max_disp=64,
                                                   max_cost=0.1) for kp in kps], np.float64)
    out["v5_disps"] = np.array([t0.match_rectified(il, ir, kp, templ_cols=41, templ_rows=15, max_disp=200,
                                                   max_cost=0.3) for kp in kps], np.float64)
    out["init_f1"] = t0.initialize_cpu(il, ir, 1)[0]
    out["init_f2"] = t0.initialize_cpu(il, ir, 2)[0]
    out["init_f4"] = t0.initialize_cpu(il, ir, 4)[0]
    out["sparse_f2"] = t0.sparse_init_gpu(il, ir, 2)[0]
    rows = np.array([0, 1, 2, 57, 120, 237, 238, 239])
    out["resp_rows"] = rows
    out["resp_eig"] = cv2.cornerMinEigenVal(il, 5, ksize=3)[rows]
    out["resp_eig_max"] = F32(cv2.cornerMinEigenVal(il, 5, ksize=3).max())
    out["resp_harris"] = cv2.cornerHarris(il, 5, 3, 0.04)[rows]
    out["resp_harris_max"] = F32(cv2.cornerHarris(il, 5, 3, 0.04).max())
    np.savez_compressed(OUT, **out)
    print(OUT, os.path.getsize(OUT))


if __name__ == "__main__":
    main()
