"""stereo::Patchmatch (the reference's CPU stage library) on the GPU against the goldens
frozen from the literal cv2 transliteration on the reference's own fixture (config C1),
and against the C oracle on other inputs. Bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture()
def pmc(pkg, engine_factory):
    return pkg.Patchmatch(engine_factory())


def test_cost_functor_kat(pmc, c1, kat):
    # L1GradientCostFunction on getRectSubPix patches: 2000 samples made with cv2
    pmc.load(c1["il"], c1["ir"])
    got = pmc.cost(kat["c_x"], kat["c_y"], kat["c_d"], kat["c_pw"])
    assert np.array_equal(got, kat["c_cost"])


def test_cost_functor_at_borders(pmc, pmo, c1):
    il, ir = c1["il"], c1["ir"]
    gl, gr = pmo.gradient_mag(il), pmo.gradient_mag(ir)
    pmc.load(il, ir)
    h, w = il.shape
    xs, ys, ds, pws = [], [], [], []
    rng = np.random.default_rng(4)
    for pw in (3, 5):
        for x in (pw // 2, pw // 2 + 1, w - pw // 2 - 2, w - pw // 2 - 1):
            for y in (pw // 2, h - pw // 2 - 1, h // 2):
                for d in (0.0, 0.5, float(x - pw // 2), float(rng.uniform(0, max(x - pw // 2, 0)))):
                    xs.append(x); ys.append(y); ds.append(d); pws.append(pw)
    got = pmc.cost(xs, ys, ds, pws)
    want = np.array([pmo.c_cost(il, ir, gl, gr, x, y, d, p, p) for x, y, d, p in zip(xs, ys, ds, pws)],
                    np.float32)
    assert np.array_equal(got, want)


def test_add_noise_and_first_propagate_match_t0(pmc, c1, c1_cpu):
    pmc.load(c1["il"], c1["ir"])
    pmc.set_disp(c1["seed_cpu"])
    pmc.AddNoise(32.0)
    assert np.array_equal(pmc.get_disp(), c1_cpu["noise0"])
    pmc.Propagate(5, 5)
    assert np.array_equal(pmc.get_disp(), c1_cpu["prop0"])


def test_single_passes_match_oracle(pmc, pmo, c1):
    il, ir = c1["il"][:200, :360].copy(), c1["ir"][:200, :360].copy()
    gl, gr = pmo.gradient_mag(il), pmo.gradient_mag(ir)
    rng = np.random.default_rng(8)
    disp = (rng.uniform(0, 40, il.shape) * (rng.uniform(0, 1, il.shape) > 0.3)).astype(np.float32)
    pmc.load(il, ir)
    for ps in range(4):
        for patch in (3, 5):
            pmc.set_disp(disp)
            pmc.Propagate(patch, patch, single_pass=ps)
            want = pmo.c_propagate(il, ir, gl, gr, disp, patch, patch, passes=[ps])
            assert np.array_equal(pmc.get_disp(), want), (ps, patch)


def test_remove_background_matches_t0(pmc, c1, c1_cpu):
    pmc.load(c1["il"], c1["ir"])
    pmc.set_disp(c1_cpu["prop3"])
    pmc.RemoveBackground(3, 3, 1.5)
    assert np.array_equal(pmc.get_disp(), c1_cpu["final"])


def test_estimate_disparity_matches_t0(pmc, c1, c1_cpu):
    """The whole CPU pipeline of the reference's test (patchmatch_test.cpp:156-183) on the
    reference's fixture: GPU == cv2-literal T0, bit for bit."""
    out = pmc.EstimateDisparity(c1["il"], c1["ir"], c1["seed_cpu"])
    assert np.array_equal(out, c1_cpu["final"])
    assert (out > 0).sum() > 10000


def test_non_power_of_two_noise_amount(pmc, pmo, c1):
    pmc.load(c1["il"], c1["ir"])
    pmc.set_disp(c1["seed_cpu"])
    pmc.AddNoise(3.3)
    assert np.array_equal(pmc.get_disp(), pmo.c_add_noise(c1["seed_cpu"], 3.3))
