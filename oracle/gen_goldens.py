"""Generate tests/golden/*.npz from the reference's own fixture with the literal
cv2 transliteration (oracle/t0_literal.py).

Run in the authoring container only (needs /root/reference and cv2):
    python oracle/gen_goldens.py
The GPU box never runs this; tests read the committed .npz files.

What is frozen (config C1 of BASELINE.md: fsl1/fsr1 at 376x240):
  c1_inputs.npz   the two grayscale half-size images, GFTT keypoints, matched
                  disparities, Patchmatch::Initialize seeds, PatchmatchGpu::SparseInit
                  seeds for both views
  c1_cpu.npz      stereo::Patchmatch driven by the test schedule
                  (patchmatch_test.cpp:173-183): disparity after the first
                  AddNoise+Propagate and after RemoveBackground
  kat.npz         known-answer vectors for cv::RNG, getRectSubPix (u8/f32), the
                  test's cost functor, Sobel magnitude, resize/2, dilate
"""
import os
import sys
import time

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import t0_literal as t0  # noqa: E402

REF = "/root/reference/test/resources/images"
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
F32 = np.float32


def main():
    os.makedirs(OUT, exist_ok=True)
    il, ir = t0.load_fixture_pair(os.path.join(REF, "fsl1.png"), os.path.join(REF, "fsr1.png"))
    h, w = il.shape
    print("fixture", il.shape)

    # ---- seeds
    seed_cpu, kps, disps = t0.initialize_cpu(il, ir, 1)
    seed_gl, _, _ = t0.sparse_init_gpu(il, ir, 4)
    ilf = np.ascontiguousarray(il[:, ::-1])
    irf = np.ascontiguousarray(ir[:, ::-1])
    seed_gr_flipped, kps_r, disps_r = t0.sparse_init_gpu(irf, ilf, 4)  # patchmatch_gpu.cu:362-365
    seed_gr = np.ascontiguousarray(seed_gr_flipped[:, ::-1])            # right-image coordinates
    np.savez_compressed(
        os.path.join(OUT, "c1_inputs.npz"), il=il, ir=ir,
        kps=np.array(kps, F32), kp_disps=np.array(disps, np.float64),
        kps_r=np.array(kps_r, F32), kp_disps_r=np.array(disps_r, np.float64),
        seed_cpu=seed_cpu, seed_gpu_l=seed_gl, seed_gpu_r=seed_gr)

    # ---- KATs
    rng = np.random.default_rng(20261018)
    Gl = t0.compute_gradient(il)
    Gr = t0.compute_gradient(ir)
    full = cv2.imread(os.path.join(REF, "fsl1.png"), cv2.IMREAD_GRAYSCALE)
    n_sp = 400
    sp_pw = rng.choice([3, 5], n_sp).astype(np.int32)
    sp_cx = np.empty(n_sp, F32)
    sp_cy = np.empty(n_sp, F32)
    for i in range(n_sp):
        if i % 4 == 0:   # anywhere, incl. outside the image
            sp_cx[i] = rng.uniform(-4, w + 4); sp_cy[i] = rng.uniform(-4, h + 4)
        elif i % 4 == 1:  # integral row, fractional column (the PatchMatch case)
            sp_cx[i] = rng.uniform(1, w - 1); sp_cy[i] = rng.integers(0, h)
        elif i % 4 == 2:  # exact ties
            sp_cx[i] = rng.integers(0, w) + rng.choice([0.0, 0.5]); sp_cy[i] = rng.integers(0, h)
        else:             # last row / column (replicate-border branch)
            sp_cx[i] = w - 1 - rng.uniform(0, 3); sp_cy[i] = h - 1 - rng.integers(0, 3)
    sp_u8 = np.zeros((n_sp, 25), np.uint8)
    sp_f32 = np.zeros((n_sp, 25), F32)
    for i in range(n_sp):
        pw = int(sp_pw[i])
        sp_u8[i, :pw * pw] = cv2.getRectSubPix(il, (pw, pw), (float(sp_cx[i]), float(sp_cy[i]))).ravel()
        sp_f32[i, :pw * pw] = cv2.getRectSubPix(Gl, (pw, pw), (float(sp_cx[i]), float(sp_cy[i]))).ravel()
    n_c = 2000
    c_pw = rng.choice([3, 5], n_c).astype(np.int32)
    c_x = np.empty(n_c, np.int32); c_y = np.empty(n_c, np.int32)
    c_d = np.empty(n_c, F32); c_cost = np.empty(n_c, F32)
    for i in range(n_c):
        pw = int(c_pw[i])
        c_x[i] = rng.integers(pw // 2, w - pw // 2)
        c_y[i] = rng.integers(pw // 2, h - pw // 2)
        hi = c_x[i] - pw // 2
        c_d[i] = F32(rng.uniform(0, hi)) if hi > 0 else F32(0)
        if i % 5 == 0:
            c_d[i] = F32(np.floor(c_d[i]))
        x, y, d = int(c_x[i]), int(c_y[i]), c_d[i]
        c_cost[i] = t0.l1_gradient_cost(
            t0.get_patch_subpix(il, x, y, pw, pw), t0.get_patch_subpix(ir, F32(x) - d, y, pw, pw),
            t0.get_patch_subpix(Gl, x, y, pw, pw), t0.get_patch_subpix(Gr, F32(x) - d, y, pw, pw))
    dil_in = np.zeros((64, 96), F32)
    dil_in[rng.integers(0, 64, 12), rng.integers(0, 96, 12)] = rng.uniform(1, 98, 12).astype(F32)
    el = cv2.getStructuringElement(cv2.MORPH_RECT, (35, 35), (17, 17))
    np.savez_compressed(
        os.path.join(OUT, "kat.npz"),
        rng123_unit=t0.rng_uniform((4, 64), -1, 1, 123).ravel(),
        rng123_32=t0.rng_uniform((4, 64), -32, 32, 123).ravel(),
        rng123_half=t0.rng_uniform((4, 64), -0.5, 0.5, 123).ravel(),
        rng7_unit=t0.rng_uniform((4, 64), -1, 1, 7).ravel(),
        grad_l_rows=Gl[[0, 1, 117, h - 2, h - 1]], grad_rows_idx=np.array([0, 1, 117, h - 2, h - 1]),
        grad_l_sum=np.float64(Gl.astype(np.float64).sum()),
        full_rows=full[200:204], half_rows=il[100:102],
        sp_pw=sp_pw, sp_cx=sp_cx, sp_cy=sp_cy, sp_u8=sp_u8, sp_f32=sp_f32,
        c_pw=c_pw, c_x=c_x, c_y=c_y, c_d=c_d, c_cost=c_cost,
        dil_in=dil_in, dil_out=cv2.dilate(dil_in, el))

    # ---- CPU pipeline, test schedule (patchmatch_test.cpp:173-183)
    disp = seed_cpu.copy()
    stages = {}
    t_all = time.time()
    for s, (amount, patch) in enumerate([(32.0, 5), (8.0, 5), (2.0, 3), (0.5, 3)]):
        t = time.time()
        disp = t0.add_noise(disp, amount)
        if s == 0:
            stages["noise0"] = disp.copy()
        t0.propagate(il, ir, Gl, Gr, disp, patch, patch)
        print("stage", s, "%.1fs" % (time.time() - t))
        if s == 0:
            stages["prop0"] = disp.copy()
    stages["prop3"] = disp.copy()
    t0.remove_background(il, ir, Gl, Gr, disp, 3, 3, 1.5)
    stages["final"] = disp.copy()
    print("cpu pipeline %.1fs" % (time.time() - t_all))
    np.savez_compressed(os.path.join(OUT, "c1_cpu.npz"), **stages)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
