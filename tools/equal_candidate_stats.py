"""How often is a sweep's candidate bit-equal to the pixel's current disparity? (lever (a))

If prev == cur.d the cached cost IS the candidate's cost, the strict `<` fails and the step's result
is known without an evaluation. Measured with the oracle's sweep on the bench workload (random init,
2 levels, 1280x720, D=128): per sweep, the fraction of visited pixels that are equal ("lane"), and the
fraction of groups of adjacent lines (the lanes of a warp) in which every lane is equal ("warp32",
"half16"): only those save issue slots; per-lane equality saves gathers (predicated loads).
"""
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pmo  # noqa: E402

pkg = importlib.import_module("ocean-perception_b200")


def stats(d_in, d_out, along_x, direction):
    # candidate at position p = the chain's value at p - dir = d_out there (chunk starts aside)
    axis = 1 if along_x else 0
    cand = np.roll(d_out, direction, axis=axis)
    eq = cand == d_in
    h, w = eq.shape
    inner = eq[1:-1, 1:-1]
    if along_x:   # lanes = adjacent rows
        g32 = eq[: h // 32 * 32].reshape(h // 32, 32, w).all(1)
        g16 = eq[: h // 16 * 16].reshape(h // 16, 16, w).all(1)
    else:         # lanes = adjacent columns
        g32 = eq[:, : w // 32 * 32].reshape(h, w // 32, 32).all(2)
        g16 = eq[:, : w // 16 * 16].reshape(h, w // 16, 16).all(2)
    return {"lane": float(inner.mean()), "warp32": float(g32.mean()), "half16": float(g16.mean())}


def main():
    W, H, D = 1280, 720, 128
    levels, iters = 2, 3
    L, R, _ = pkg.synth.make_pair(0, W, H, D)
    p = pmo.default_params(init_mode=1, max_disp=D, pyramid_levels=levels)
    imgs = [(L, R)]
    for l in range(1, levels):
        imgs.append((pmo.resize_half(imgs[-1][0]), pmo.resize_half(imgs[-1][1])))
    out = []
    prev = None
    for l in range(levels - 1, -1, -1):
        Ll, Rl = imgs[l]
        h, w = Ll.shape
        planes = pmo.g_planes(Ll, Rl, 0)
        noise = pmo.rng_uniform(123, -1, 1, w * h).reshape(h, w)
        if l == levels - 1:
            disp = pmo.x_random_init(p, w, h, 0, 0, l, D / float(1 << l))
        else:
            disp = pmo.x_upsample2(prev, w, h)
        iter0 = (levels - 1 - l) * iters
        for it in range(iters):
            scale = 32.0 / (1 << l) / 2.0 ** (iter0 + it)
            disp = pmo.g_add_noise(disp, noise, scale)
            for s, (ax, dr) in enumerate(((1, 1), (0, 1), (1, -1), (0, -1))):
                cost = pmo.g_cost_map(*planes, disp, 0.9)
                d2, _ = pmo.g_sweep_chains(*planes, disp, cost, ax, dr)
                st = stats(disp, d2, ax, dr)
                st.update(level=l, iter=it, sweep="%s%+d" % ("row" if ax else "col", dr),
                          changed=float((d2 != disp).mean()))
                out.append(st)
                print(st)
                disp = d2
        prev = disp
    with open(os.path.join(ROOT, "profiles", "r2_equal_candidate_stats.json"), "w") as f:
        json.dump({"workload": "1280x720, D=128, random init, 2 levels, 3 iters, left view, pair 0",
                   "sweeps": out}, f, indent=1)


if __name__ == "__main__":
    main()
