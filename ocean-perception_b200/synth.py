"""Deterministic synthetic stereo pairs.

left  = clamp(128 + 40*t), t = per-pixel splitmix64 noise in U(-1,1), box-blurred 3x3
        twice, plus a low-frequency ramp (mean ~128, sigma ~30 like the fsl1 fixture);
truth = background plane D/8, 12 axis-aligned rectangles with d in [D/8, 7D/8) in
        0.25-px steps painted near-over-far, one flat-intensity rectangle with d = 0;
right = every layer's texture shifted by its disparity (bilinear), composited
        near-over-far, dis-occlusions filled with fresh noise.
The generator runs on the host (numpy) and is part of the bench/test harness only.
"""
import numpy as np

_GOLD = 0x9E3779B97F4A7C15
_M64 = (1 << 64) - 1


def _splitmix64(x):
    """Vectorised splitmix64 finaliser over a uint64 array."""
    with np.errstate(over="ignore"):
        z = x + np.uint64(_GOLD)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def _noise(seed, salt, h, w):
    idx = np.arange(h * w, dtype=np.uint64).reshape(h, w)
    with np.errstate(over="ignore"):
        key = np.uint64((seed + salt * 0xD1B54A32D192ED03) & _M64)
        r = _splitmix64(idx * np.uint64(0x2545F4914F6CDD1D) + key)
    return (r >> np.uint64(40)).astype(np.float64) * (2.0 / (1 << 24)) - 1.0


def _box3(a):
    p = np.pad(a, 1, mode="edge")
    return (p[:-2, :-2] + p[:-2, 1:-1] + p[:-2, 2:] + p[1:-1, :-2] + p[1:-1, 1:-1] + p[1:-1, 2:] +
            p[2:, :-2] + p[2:, 1:-1] + p[2:, 2:]) / 9.0


def _texture(seed, salt, h, w):
    t = _box3(_box3(_noise(seed, salt, h, w))) * 5.5  # two box blurs shrink sigma by ~5x
    yy, xx = np.mgrid[0:h, 0:w]
    ramp = 0.35 * np.sin(xx * (2 * np.pi / max(w, 1)) * 1.5 + 0.7) * np.cos(yy * (2 * np.pi / max(h, 1)))
    return t + ramp


class _Rng:
    """Tiny scalar splitmix64 stream for the scene layout."""

    def __init__(self, seed):
        self.s = seed & _M64

    def u64(self):
        self.s = (self.s + _GOLD) & _M64
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
        return z ^ (z >> 31)

    def uniform(self, lo, hi):
        return lo + (hi - lo) * ((self.u64() >> 11) / float(1 << 53))

    def randint(self, lo, hi):
        return lo + int(self.u64() % max(1, hi - lo))


def make_pair(pair_index, w, h, max_disp, shift=0.0):
    """Returns (left u8, right u8, truth f32): truth is the left-view disparity, 0 where
    the scene is at infinity (the flat rectangle). `shift` adds a constant to the disparity of
    every layer but the one at infinity (frames of a sequence: tools/sequence_bench.py)."""
    seed = (_GOLD * (pair_index + 1)) & _M64
    rng = _Rng(seed ^ 0xA5A5A5A5)
    D = float(max_disp)
    tex = _texture(seed, 1, h, w)
    left_f = 128.0 + 40.0 * tex
    truth = np.full((h, w), np.float32(D / 8.0 + shift), np.float32)
    layers = [(D / 8.0 + shift, 0, h, 0, w, False)]  # (d, y0, y1, x0, x1, flat)
    rects = []
    for _ in range(12):
        rw = rng.randint(w // 10, w // 3)
        rh = rng.randint(h // 10, h // 3)
        x0 = rng.randint(0, w - rw)
        y0 = rng.randint(0, h - rh)
        d = np.floor(rng.uniform(D / 8.0, 7.0 * D / 8.0) * 4.0) / 4.0 + shift
        rects.append((d, y0, y0 + rh, x0, x0 + rw, False))
    rects.sort(key=lambda r: r[0])  # far first, near painted last
    fw, fh = w // 6, h // 6
    fx0, fy0 = rng.randint(0, w - fw), rng.randint(0, h - fh)
    flat = (0.0, fy0, fy0 + fh, fx0, fx0 + fw, True)
    for r in rects:
        truth[r[1]:r[2], r[3]:r[4]] = np.float32(r[0])
    truth[flat[1]:flat[2], flat[3]:flat[4]] = 0.0
    left_f[flat[1]:flat[2], flat[3]:flat[4]] = 128.0 + 40.0 * 0.2
    layers += rects
    # the flat rectangle sits at infinity: paint it first in the right view (farthest)
    right_f = 128.0 + 40.0 * _texture(seed, 2, h, w)  # fresh noise for dis-occlusions
    order = [flat] + layers
    # ownership map: a layer only contributes the left pixels it actually owns
    owner = np.zeros((h, w), np.int32)
    for i, r in enumerate(layers[1:], start=1):
        owner[r[1]:r[2], r[3]:r[4]] = i
    owner[flat[1]:flat[2], flat[3]:flat[4]] = -1
    xs = np.arange(w, dtype=np.float64)
    for li, (d, y0, y1, x0, x1, is_flat) in enumerate(order):
        lid = -1 if is_flat else li - 1
        src = xs + d                      # right pixel xr shows left pixel xr + d
        i0 = np.floor(src).astype(np.int64)
        t = src - i0
        ok = (i0 >= 0) & (i0 + 1 <= w - 1)
        i0c = np.clip(i0, 0, w - 2)
        rows = left_f[y0:y1]
        own = owner[y0:y1]
        val = rows[:, i0c] * (1 - t) + rows[:, i0c + 1] * t
        valid = ok[None, :] & (own[:, i0c] == lid) & (own[:, np.clip(i0c + (t > 0), 0, w - 1)] == lid)
        sub = right_f[y0:y1]
        sub[valid] = val[valid]
    left = np.clip(np.rint(left_f), 0, 255).astype(np.uint8)
    right = np.clip(np.rint(right_f), 0, 255).astype(np.uint8)
    return left, right, truth


def make_batch(first_index, n, w, h, max_disp, unique=None):
    """[n,h,w] uint8 left/right and float32 truth; with `unique` < n the first `unique`
    pairs are generated and repeated cyclically (same pixels, same amount of work)."""
    u = n if unique is None else min(unique, n)
    L = np.empty((n, h, w), np.uint8)
    R = np.empty((n, h, w), np.uint8)
    T = np.empty((n, h, w), np.float32)
    for i in range(u):
        L[i], R[i], T[i] = make_pair(first_index + i, w, h, max_disp)
    for i in range(u, n):
        L[i], R[i], T[i] = L[i % u], R[i % u], T[i % u]
    return L, R, T
