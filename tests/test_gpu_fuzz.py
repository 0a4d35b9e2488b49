"""Randomised configurations (fixed seed): whole pipelines on the GPU against the oracle, bit for bit.
Sizes, sweep chunking, iterations, pyramid levels, both init modes, both cost modes and the
extension switches are drawn together, so that kernel-selection boundaries (shared-memory row kernel,
block column kernel, transposed row sweeps, generic kernel, fused noise) are crossed in combination."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _configs(n, seed):
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < n:
        chunks = int(rng.choice([4, 8, 16]))
        ov = int(rng.integers(2, 6))
        levels = int(rng.choice([1, 1, 2]))
        need = (2 * ov + 2) * chunks << (levels - 1)
        w = int(rng.integers(max(need, 140), max(need, 140) + 420))
        h = int(rng.integers(need, need + 260))
        if rng.uniform() < 0.15:
            w = int(rng.integers(1340, 1500))   # wider than the shared-memory row kernel stages
            h = int(rng.integers(need, need + 40))
        cfg = dict(w=w, h=h, sweep_chunks=chunks, sweep_overlap=ov, pyramid_levels=levels,
                   patchmatch_iters=int(rng.integers(0, 3)), max_disp=int(rng.choice([24, 48, 64])),
                   init_mode=str(rng.choice(["random", "sparse"])),
                   cost_mode=str(rng.choice(["l1grad_x5", "l1grad_x5", "l1grad_full"])),
                   lr_mode=str(rng.choice(["ratio", "abs1px"])), subpixel=int(rng.integers(0, 2)),
                   median_ksize=int(rng.choice([0, 0, 3, 5])),
                   noise_accept=str(rng.choice(["always", "improve"])), clamp_disp=int(rng.integers(0, 2)),
                   cost_alpha=float(rng.choice([0.9, 0.5, 1.0])), seed=int(rng.integers(1, 1000)))
        if cfg["init_mode"] == "sparse" and cfg["w"] <= 128:
            continue
        out.append(cfg)
    return out


CONFIGS = _configs(48, 20261018)


@pytest.mark.parametrize("i", range(len(CONFIGS)))
def test_random_configuration(pkg, pmo, engine_factory, i):
    cfg = dict(CONFIGS[i])
    w, h = cfg.pop("w"), cfg.pop("h")
    L, R, _ = pkg.synth.make_pair(100 + i, w, h, cfg["max_disp"])
    e = engine_factory(**cfg)
    enum = {"init_mode": {"sparse": 0, "random": 1}, "noise_accept": {"always": 0, "improve": 1},
            "lr_mode": {"ratio": 0, "abs1px": 1}, "cost_mode": {"l1grad_x5": 0, "l1grad_full": 1}}
    p = pmo.default_params(**{k: (enum[k][v] if k in enum else v) for k, v in cfg.items()})
    dl, dr = e.Match(L, R, pair_index=i)
    if cfg["init_mode"] == "sparse":
        sl, sr = pmo.s_match_seeds(L, R, 4)
        wl, wr = pmo.g_match(p, L, R, sl, sr, pair_index=i)
    else:
        wl, wr = pmo.g_match(p, L, R, pair_index=i)
    assert np.array_equal(dl, wl) and np.array_equal(dr, wr), (cfg, w, h, int((dl != wl).sum()))
