"""One frame in row bands (pm_band_*) on the GPU: all bands of a
frame are run in lock step on ONE device, their exchange buffers copied where NCCL would
move them; the result must be bit-identical to the whole-frame pass (and so to the oracle).
The real NCCL transport is exercised by `tools/band_bench.py --check` under torchrun and by the
c5_band leg of bench.py at N >= 2."""
import importlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _frame(pkg, w, h, D, idx=3):
    return pkg.synth.make_pair(idx, w, h, D)[:2]


def _params(pkg, **kw):
    P = pkg.PatchmatchGpu.Params()
    for k, v in kw.items():
        setattr(P, k, v)
    return P


@pytest.mark.parametrize("world", [2, 4, 8, 16])
def test_bands_equal_whole_frame_random_init(pkg, built_lib, pmo, world):
    bands = importlib.import_module("ocean-perception_b200.bands")
    w, h, D = 416, 266, 48          # 266 rows: chunks of 16, 10 remainder rows in the last band
    L, R = _frame(pkg, w, h, D)
    P = _params(pkg, init_mode="random", max_disp=D)
    eng = pkg.PatchmatchGpu(P, device=0)
    wl, wr = eng.Match(L, R, pair_index=5)
    eng.close()
    dl, dr = bands.match_bands_one_device(P, L, R, world, pair_index=5)
    assert np.array_equal(dl, wl) and np.array_equal(dr, wr)
    if world == 2:                  # and the whole-frame pass is the oracle's
        ol, orr = pmo.g_match(pmo.default_params(init_mode=1, max_disp=D), L, R, pair_index=5)
        assert np.array_equal(wl, ol) and np.array_equal(wr, orr)


@pytest.mark.parametrize("world", [2, 8])
def test_bands_equal_whole_frame_seeds_and_extensions(pkg, built_lib, c1, world):
    bands = importlib.import_module("ocean-perception_b200.bands")
    P = _params(pkg, lr_mode="abs1px", subpixel=1, median_ksize=3, clamp_disp=1, max_disp=64,
                noise_accept="improve")
    eng = pkg.PatchmatchGpu(P, device=0)
    wl, wr = eng.Match(c1["il"], c1["ir"], c1["seed_gpu_l"], c1["seed_gpu_r"])
    eng.close()
    dl, dr = bands.match_bands_one_device(P, c1["il"], c1["ir"], world, c1["seed_gpu_l"], c1["seed_gpu_r"])
    assert np.array_equal(dl, wl) and np.array_equal(dr, wr)


def test_single_band_is_the_whole_frame(pkg, built_lib):
    bands = importlib.import_module("ocean-perception_b200.bands")
    L, R = _frame(pkg, 320, 200, 32)
    P = _params(pkg, init_mode="random", max_disp=32, patchmatch_iters=2)
    eng = pkg.PatchmatchGpu(P, device=0)
    wl, wr = eng.Match(L, R)
    eng.close()
    dl, dr = bands.match_bands_one_device(P, L, R, 1)
    assert np.array_equal(dl, wl) and np.array_equal(dr, wr)


def test_wide_frame_bands(pkg, built_lib):
    """A frame wider than the shared-memory row kernel takes (w > 1334)."""
    bands = importlib.import_module("ocean-perception_b200.bands")
    w, h, D = 1920, 240, 96
    L, R = _frame(pkg, w, h, D)
    P = _params(pkg, init_mode="random", max_disp=D, patchmatch_iters=2, clamp_disp=1)
    eng = pkg.PatchmatchGpu(P, device=0)
    wl, wr = eng.Match(L, R)
    eng.close()
    dl, dr = bands.match_bands_one_device(P, L, R, 4)
    assert np.array_equal(dl, wl) and np.array_equal(dr, wr)


def test_band_state_errors(pkg, built_lib):
    eng = pkg.PatchmatchGpu(_params(pkg, init_mode="random"), device=0)
    with pytest.raises(pkg.PmError) as ei:
        eng.band_step()
    assert ei.value.code == -6
    with pytest.raises(pkg.PmError):
        eng.band_finish(0, 0, 0)
    eng.close()


def test_c5_full_size_whole_frame_and_two_bands_equal_the_oracle(pmo):
    """BASELINE config C5 at its real size: 3840x2160, 256-disparity range, 3 iterations, random init.
    The whole-frame pass (wide-frame row sweeps = the column kernel on transposed planes) and the
    2-band split (both bands in lock step on one device) against the ORACLE, not against each other."""
    pkg = importlib.import_module("ocean-perception_b200")
    bands = importlib.import_module("ocean-perception_b200.bands")
    W, H, D = 3840, 2160, 256
    L, R, T = pkg.synth.make_pair(0, W, H, D)
    P = pkg.PatchmatchGpu.Params()
    P.init_mode, P.max_disp, P.clamp_disp = "random", D, 1
    wl, wr = pmo.g_match(pmo.default_params(init_mode=1, max_disp=D, clamp_disp=1), L, R)
    eng = pkg.PatchmatchGpu(P, device=0)
    dl, dr = eng.Match(L, R)
    eng.close()
    assert np.array_equal(dl, wl) and np.array_equal(dr, wr), (int((dl != wl).sum()), int((dr != wr).sum()))
    bl, br = bands.match_bands_one_device(P, L, R, 2)
    assert np.array_equal(bl, wl) and np.array_equal(br, wr)
    found = (dl > 0) & (T > 0)
    assert found.mean() > 0.7 and (np.abs(dl - T)[found] <= 1.0).mean() > 0.97


@pytest.mark.parametrize("world", [2, 4])
def test_peer_memory_exchange_equals_the_whole_frame(pmo, world):
    """pm_band_p2p_*: the bands push their halo rows into each other's buffers and wait on sequence
    flags (no host in the loop). All bands run concurrently on ONE device here, one stream each, the
    receive regions connected by pointer; the result must equal the whole frame and the oracle."""
    pkg = importlib.import_module("ocean-perception_b200")
    bands = importlib.import_module("ocean-perception_b200.bands")
    W, H, D = 1920, 480, 64
    L, R, _ = pkg.synth.make_pair(4, W, H, D)
    P = pkg.PatchmatchGpu.Params()
    P.init_mode, P.max_disp, P.clamp_disp = "random", D, 1
    wl, wr = pmo.g_match(pmo.default_params(init_mode=1, max_disp=D, clamp_disp=1), L, R, pair_index=2)
    bl, br = bands.match_bands_one_device_p2p(P, L, R, world, pair_index=2)
    assert np.array_equal(bl, wl), int((bl != wl).sum())
    assert np.array_equal(br, wr)
    bl2, br2 = bands.match_bands_one_device_p2p(P, L, R, world, pair_index=2)   # second frame: flags keep counting
    assert np.array_equal(bl2, wl) and np.array_equal(br2, wr)
