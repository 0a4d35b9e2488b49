"""The stages the north-star names and the reference lacks or leaves as dead code, against the oracle
(-m gpu): the generic L1GradientCost with a 5x5 patch (patchmatch_gpu.cu:45-69; the reference's CPU
driver runs 5x5 patches, patchmatch_test.cpp:173-176), the kernels' patch_radius = 2, a census/Hamming
cost, and K-candidate random-search refinement. Bit-exact: same float operations in the same order."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rand_disp(w, h, seed, hi=40.0, zero_frac=0.3):
    rng = np.random.default_rng(seed)
    return (rng.uniform(0, hi, (h, w)) * (rng.uniform(0, 1, (h, w)) > zero_frac)).astype(np.float32)


@pytest.fixture()
def oracle_mode(pmo):
    """Sets the oracle's cost mode / patch size for stage calls and restores the defaults."""
    def set_(mode, patch):
        pmo.set_cost_mode(mode)
        pmo.set_patch_size(patch)
    yield set_
    pmo.set_cost_mode(0)
    pmo.set_patch_size(3)


@pytest.mark.parametrize("mode,name,patch", [(1, "l1grad_full", 5), (2, "census", 5), (2, "census", 3),
                                             (0, "l1grad_x5", 5)])
def test_stages_with_patch_radius_and_cost_modes(pmo, pmref_or_none, engine_factory, c1, oracle_mode,
                                                 mode, name, patch):
    il, ir = c1["il"], c1["ir"]
    h, w = il.shape
    r = patch // 2
    e = engine_factory(cost_mode=name, patch_size=patch)
    e.stage_load_pair(il, ir)
    oracle_mode(mode, patch)
    for view in (0, 1):
        planes = pmo.g_planes(il, ir, view)
        d0 = _rand_disp(w, h, 21 + view)
        e.stage_set_disp(view, d0)
        _, cost = e.stage_get_disp(view, want_cost=True)
        want = pmo.g_cost_map(*planes, d0, 0.9)
        assert np.array_equal(cost[r:h - r, r:w - r], want[r:h - r, r:w - r]), (name, patch, view)
        if mode == 1 and pmref_or_none is not None:
            # the reference's own generic L1GradientCost, ph = pw = 5
            ref = pmref_or_none.cost_map(*planes, d0, 0.9, ph=patch, pw=patch)
            assert np.array_equal(ref[r:h - r, r:w - r], want[r:h - r, r:w - r])
        for along_x, direction in ((1, 1), (0, 1), (1, -1), (0, -1)):
            e.stage_set_disp(view, d0)
            e.stage_propagate(view, along_x, direction)
            got = e.stage_get_disp(view)
            assert np.array_equal(got, pmo.g_propagate(*planes, d0, along_x, direction)), (name, patch, view, along_x, direction)
        e.stage_set_disp(view, d0)
        e.stage_mask_background(view)
        assert np.array_equal(e.stage_get_disp(view), pmo.g_mask_background(*planes, d0))
        # the border the kernels must not touch: `patch_radius` rows and columns
        e.stage_set_disp(view, d0)
        e.stage_propagate(view, 1, 1)
        got = e.stage_get_disp(view)
        assert np.array_equal(got[:r], d0[:r]) and np.array_equal(got[:, :r], d0[:, :r])


@pytest.fixture()
def pmref_or_none():
    try:
        import pmref
        pmref.lib()
        return pmref if pmref.lib().pmref_device_count() > 0 else None
    except Exception:
        return None


@pytest.mark.parametrize("kw", [
    dict(cost_mode="l1grad_full", patch_size=5, init_mode="random", max_disp=48),
    dict(cost_mode="census", patch_size=5, init_mode="random", max_disp=48, pyramid_levels=2),
    dict(random_search_k=2, init_mode="random", max_disp=48),
    dict(random_search_k=3, init_mode="random", max_disp=48, pyramid_levels=2, clamp_disp=1,
         noise_accept="improve", subpixel=1, median_ksize=3, lr_mode="abs1px"),
    dict(random_search_k=2, cost_mode="census", patch_size=5, init_mode="random", max_disp=48),
])
def test_pipelines_with_extensions_equal_the_oracle(pmo, pkg, engine_factory, kw):
    L, R, T = pkg.synth.make_pair(2, 640, 400, 48)
    e = engine_factory(**kw)
    dl, dr = e.Match(L, R, pair_index=7)
    enum = {"init_mode": {"random": 1}, "noise_accept": {"improve": 1}, "lr_mode": {"abs1px": 1},
            "cost_mode": {"l1grad_full": 1, "census": 2}}
    okw = {k: (enum[k][v] if k in enum else v) for k, v in kw.items()}
    wl, wr = pmo.g_match(pmo.default_params(**okw), L, R, pair_index=7)
    assert np.array_equal(dl, wl), (kw, int((dl != wl).sum()))
    assert np.array_equal(dr, wr)
    found = (dl > 0) & (T > 0)
    assert found.mean() > 0.3


def test_random_search_improves_the_result(pkg, engine_factory):
    """The refinement is improve-only: with it the synthetic-truth accuracy must not drop, and on a
    random-init run with few iterations it visibly helps."""
    L, R, T = pkg.synth.make_pair(5, 640, 400, 64)
    res = {}
    for k in (0, 4):
        e = engine_factory(init_mode="random", max_disp=64, patchmatch_iters=2, random_search_k=k)
        dl, _ = e.Match(L, R, pair_index=5)
        found = (dl > 0) & (T > 0)
        res[k] = (found.mean(), float((np.abs(dl - T)[found] <= 0.5).mean()))
    print("random search K=0 / K=4: valid %.3f / %.3f, within 0.5 px %.3f / %.3f" %
          (res[0][0], res[4][0], res[0][1], res[4][1]))
    assert res[4][1] >= res[0][1] - 0.01


def test_census_cost_is_robust_to_gain_and_offset(pmo, pkg, engine_factory):
    """Census compares orderings: a gain/offset change of the right image leaves the census result
    unchanged where nothing saturates, while the L1 cost degrades."""
    L, R, T = pkg.synth.make_pair(6, 512, 320, 48)
    R2 = np.clip(np.rint(R.astype(np.float32) * 0.8 + 20.0), 0, 255).astype(np.uint8)
    e = engine_factory(cost_mode="census", patch_size=5, init_mode="random", max_disp=48)
    d1, _ = e.Match(L, R, pair_index=1)
    d2, _ = e.Match(L, R2, pair_index=1)
    ok1 = (np.abs(d1 - T)[(d1 > 0) & (T > 0)] <= 1).mean()
    ok2 = (np.abs(d2 - T)[(d2 > 0) & (T > 0)] <= 1).mean()
    print("census within 1 px: %.3f, with gain 0.8 / offset 20 on the right image: %.3f" % (ok1, ok2))
    assert ok2 > ok1 - 0.05
