"""Golden vectors of ForegroundTextureMask (reference: src/vehicle/stereo_matching/patchmatch.cpp:19-49)
made with the same OpenCV calls in the same order (cv2 4.13, IPP off) on the reference's fixture
fsl1.png at 376x240 (tests/golden/c1_inputs.npz 'il'). Run in the authoring container:

    python oracle/gen_goldens_texture_mask.py     ->  tests/golden/texture_mask.npz
"""
import os

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = ((7, 35.0, 2), (7, 35.0, 1), (9, 20.0, 2), (5, 50.0, 2), (4, 10.0, 1))   # (ksize, min_grad, downsize)


def foreground_texture_mask(gray, ksize=7, min_grad=35.0, downsize=2):
    """Literal transliteration of patchmatch.cpp:19-49."""
    assert 1 <= downsize <= 8
    sk = ksize // downsize
    assert sk > 1
    kw = 2 * sk + 1
    kernel = cv2.getStructuringElement(cv2.MORPH_RECT, (kw, kw), (sk, sk))
    h, w = gray.shape
    if downsize > 1:
        small = cv2.resize(gray, (w // downsize, h // downsize), interpolation=cv2.INTER_LINEAR)
        grad = cv2.morphologyEx(small, cv2.MORPH_GRADIENT, kernel, anchor=(-1, -1), iterations=1)
        m = (grad > min_grad).astype(np.uint8) * 255           # cv::Mat comparison: 255 / 0
        return cv2.resize(m, (w, h), interpolation=cv2.INTER_LINEAR)
    grad = cv2.morphologyEx(gray, cv2.MORPH_GRADIENT, kernel, anchor=(-1, -1), iterations=1)
    return (grad > min_grad).astype(np.uint8) * 255


def main():
    try:
        cv2.ipp.setUseIPP(False)
    except Exception:
        pass
    il = dict(np.load(os.path.join(ROOT, "tests", "golden", "c1_inputs.npz")))["il"]
    out = {"cases": np.array(CASES, np.float64)}
    for i, (k, g, d) in enumerate(CASES):
        out["mask%d" % i] = foreground_texture_mask(il, k, g, d)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "texture_mask.npz"), **out)
    print({k: (v.shape, int((v > 0).sum())) for k, v in out.items()})


if __name__ == "__main__":
    main()
