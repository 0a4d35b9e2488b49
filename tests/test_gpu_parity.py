"""CUDA path vs the CPU oracle through the C ABI (run on the B200 box: -m gpu).

Bit-exact (np.array_equal) for every stage and for the whole pipeline: the kernels
pin the same float operations as the oracle (pm_device.cuh / pm_oracle.c)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _blur(a):
    a = a.astype(np.int32)
    return ((a + np.roll(a, 1, 1) + np.roll(a, 1, 0) + np.roll(a, -1, 1) + 2) // 4).astype(np.uint8)


def _pair(w, h, shift=6, seed=0):
    rng = np.random.default_rng(seed + w * 7 + h)
    L = _blur(rng.integers(0, 256, (h, w)).astype(np.uint8))
    R = np.roll(L, -shift, axis=1)
    R[:, -shift:] = rng.integers(0, 256, (h, shift))
    return L, R


def _rand_disp(w, h, seed, hi=40.0, zero_frac=0.3):
    rng = np.random.default_rng(seed)
    return (rng.uniform(0, hi, (h, w)) * (rng.uniform(0, 1, (h, w)) > zero_frac)).astype(np.float32)


# ------------------------------------------------------------------ stages

def test_preprocess_planes(pmo, engine_factory, c1):
    e = engine_factory()
    e.stage_load_pair(c1["il"], c1["ir"])
    for view in (0, 1):
        i_ref, g_ref, i_mat, g_mat = e.stage_get_planes(view)
        w_iref, w_imat, w_gref, w_gmat = pmo.g_planes(c1["il"], c1["ir"], view)
        assert np.array_equal(i_ref, w_iref) and np.array_equal(g_ref, w_gref)
        assert np.array_equal(i_mat, w_imat) and np.array_equal(g_mat, w_gmat)


def test_noise_image(pmo, engine_factory):
    e = engine_factory()
    for (w, h) in ((376, 240), (1280, 720), (333, 77)):
        got = e.stage_noise_image(w, h)
        assert np.array_equal(got.ravel(), pmo.rng_uniform(123, -1, 1, w * h))
    e2 = engine_factory(seed=7)
    assert np.array_equal(e2.stage_noise_image(64, 4).ravel(), pmo.rng_uniform(7, -1, 1, 256))


def test_downscale2(pmo, engine_factory, c1):
    e = engine_factory()
    for im in (c1["il"], c1["ir"][:239, :375]):
        assert np.array_equal(e.stage_downscale2(im), pmo.resize_half(im))


@pytest.mark.parametrize("view", [0, 1])
def test_eval_cost(pmo, engine_factory, c1, view):
    e = engine_factory()
    e.stage_load_pair(c1["il"], c1["ir"])
    h, w = c1["il"].shape
    planes = pmo.g_planes(c1["il"], c1["ir"], view)
    for seed, hi in ((1, 40.0), (2, 400.0)):   # second case exercises the x-1 / xr >= 1 clamp
        disp = _rand_disp(w, h, seed, hi)
        if seed == 2:
            disp = np.floor(disp)               # integral disparities: t == 0 path
        e.stage_set_disp(view, disp)
        d, c = e.stage_get_disp(view, want_cost=True)
        assert np.array_equal(d, disp)
        want = pmo.g_cost_map(*planes, disp, 0.9)
        assert np.array_equal(c[1:-1, 1:-1], want[1:-1, 1:-1])


def test_add_noise(pmo, engine_factory, c1):
    e = engine_factory()
    e.stage_load_pair(c1["il"], c1["ir"])
    h, w = c1["il"].shape
    noise = pmo.rng_uniform(123, -1, 1, w * h).reshape(h, w)
    disp = _rand_disp(w, h, 5)
    for scale in (32.0, 16.0, 8.0):
        e.stage_set_disp(0, disp)
        e.stage_add_noise(0, scale)
        got, cost = e.stage_get_disp(0, want_cost=True)
        want = pmo.g_add_noise(disp, noise, scale)
        assert np.array_equal(got, want)
        wc = pmo.g_cost_map(*pmo.g_planes(c1["il"], c1["ir"], 0), want, 0.9)
        assert np.array_equal(cost[1:-1, 1:-1], wc[1:-1, 1:-1])


# (672, 200) is wide enough for the second-generation row kernel (deep load ring), the others
# run the first-generation block kernels or the generic one
@pytest.mark.parametrize("size", [(376, 240), (257, 211), (200, 193), (672, 200)])
def test_propagate_all_sweeps(pmo, engine_factory, c1, size):
    w, h = size
    if size == (376, 240):
        L, R = c1["il"], c1["ir"]
    else:
        L, R = _pair(w, h)
    e = engine_factory()
    e.stage_load_pair(L, R)
    for view in (0, 1):
        planes = pmo.g_planes(L, R, view)
        disp = _rand_disp(w, h, 11 + view)
        for along_x in (1, 0):
            for direction in (1, -1):
                e.stage_set_disp(view, disp)
                e.stage_propagate(view, along_x, direction)
                got, cost = e.stage_get_disp(view, want_cost=True)
                want = pmo.g_propagate(*planes, disp, along_x, direction)
                assert np.array_equal(got, want), (view, along_x, direction)
                assert (want != disp).sum() > 100
                wc = pmo.g_cost_map(*planes, want, 0.9)
                assert np.array_equal(cost[1:-1, 1:-1], wc[1:-1, 1:-1])


def test_propagate_other_chunkings(pmo, engine_factory):
    L, R = _pair(320, 200)
    for chunks, ov in ((1, 0), (4, 2), (8, 5), (16, 0)):
        e = engine_factory(sweep_chunks=chunks, sweep_overlap=ov)
        e.stage_load_pair(L, R)
        planes = pmo.g_planes(L, R, 0)
        disp = _rand_disp(320, 200, chunks)
        for along_x, direction in ((1, 1), (0, -1)):
            e.stage_set_disp(0, disp)
            e.stage_propagate(0, along_x, direction)
            want = pmo.g_propagate(*planes, disp, along_x, direction, chunks=chunks, overlap=ov)
            assert np.array_equal(e.stage_get_disp(0), want), (chunks, ov, along_x)


def test_mask_background(pmo, engine_factory, c1):
    e = engine_factory()
    e.stage_load_pair(c1["il"], c1["ir"])
    h, w = c1["il"].shape
    for view in (0, 1):
        disp = _rand_disp(w, h, 21 + view)
        e.stage_set_disp(view, disp)
        e.stage_mask_background(view)
        want = pmo.g_mask_background(*pmo.g_planes(c1["il"], c1["ir"], view), disp)
        assert np.array_equal(e.stage_get_disp(view), want)
        assert 0 < (want > 0).sum() < (disp > 0).sum()


def test_mask_occlusions(pmo, engine_factory):
    rng = np.random.default_rng(9)
    dl = _rand_disp(300, 100, 1, 60.0)
    dr = (dl * rng.uniform(0.5, 1.6, dl.shape)).astype(np.float32)
    for mode, name in ((0, "ratio"), (1, "abs1px")):
        e = engine_factory(lr_mode=name)
        got = e.stage_mask_occlusions(dl, dr)
        assert np.array_equal(got, pmo.g_mask_occlusions(dl, dr, mode))


def test_extension_stages(pmo, engine_factory, c1):
    e = engine_factory()
    e.stage_load_pair(c1["il"], c1["ir"])
    h, w = c1["il"].shape
    p = pmo.default_params()
    for view in (0, 1):
        e.stage_random_init(view, 5, 1, 64.0)
        assert np.array_equal(e.stage_get_disp(view), pmo.x_random_init(p, w, h, 5, view, 1, 64.0))
    disp = np.floor(_rand_disp(w, h, 31, 30.0)) + np.float32(0.25)
    disp[disp < 1] = 0
    e.stage_set_disp(0, disp)
    e.stage_subpixel(0)
    want = pmo.x_subpixel(*pmo.g_planes(c1["il"], c1["ir"], 0), disp)
    assert np.array_equal(e.stage_get_disp(0), want)
    for k in (3, 5):
        assert np.array_equal(e.stage_median(disp, k), pmo.x_median(disp, k))


# ---------------------------------------------------------------- pipeline

def test_c1_fixture_pipeline(pmo, engine_factory, c1):
    """PatchmatchGpu::Match on the reference's own fixture with the reference's defaults
    and the cv2-literal SparseInit seeds (patchmatch_gpu_test.cpp:47-92)."""
    e = engine_factory()
    dl, dr = e.Match(c1["il"], c1["ir"], c1["seed_gpu_l"], c1["seed_gpu_r"])
    wl, wr = pmo.g_match(pmo.default_params(), c1["il"], c1["ir"], c1["seed_gpu_l"], c1["seed_gpu_r"])
    assert np.array_equal(dl, wl) and np.array_equal(dr, wr)
    assert (dl > 0).sum() > 10000
    assert e.launch_count() >= 15


@pytest.mark.parametrize("kw", [
    dict(init_mode="random", max_disp=48),
    dict(init_mode="random", max_disp=48, pyramid_levels=2),
    dict(init_mode="random", max_disp=48, pyramid_levels=2, noise_accept="improve", clamp_disp=1),
    dict(init_mode="random", max_disp=48, lr_mode="abs1px", subpixel=1, median_ksize=3),
    dict(init_mode="random", max_disp=48, patchmatch_iters=1, sweep_chunks=8, sweep_overlap=3,
         cost_alpha=0.7, cost_improve_factor=0.9, seed=99, median_ksize=5),
])
def test_synthetic_pipeline_variants(pmo, pkg, engine_factory, kw):
    L, R, T = pkg.synth.make_pair(2, 640, 400, 48)
    e = engine_factory(**kw)
    dl, dr = e.Match(L, R, pair_index=2)
    enum = {"init_mode": {"random": 1}, "noise_accept": {"improve": 1}, "lr_mode": {"abs1px": 1}}
    okw = {k: (enum[k][v] if k in enum else v) for k, v in kw.items()}
    wl, wr = pmo.g_match(pmo.default_params(**okw), L, R, pair_index=2)
    assert np.array_equal(dl, wl) and np.array_equal(dr, wr)
    found = (dl > 0) & (T > 0)
    assert found.mean() > 0.4


def test_batch_equals_single_and_is_deterministic(pkg, engine_factory):
    L, R, T = pkg.synth.make_batch(0, 5, 320, 208, 32)
    e = engine_factory(init_mode="random", max_disp=32, max_batch=2)   # 3 device passes: 2+2+1
    bl, br = e.MatchBatch(L, R, first_pair_index=0)
    e1 = engine_factory(init_mode="random", max_disp=32)
    for i in range(5):
        dl, dr = e1.Match(L[i], R[i], pair_index=i)
        assert np.array_equal(bl[i], dl) and np.array_equal(br[i], dr)
    bl2, br2 = e.MatchBatch(L, R, first_pair_index=0)
    assert np.array_equal(bl, bl2) and np.array_equal(br, br2)


def test_tapered_blocking_batch_equals_async_and_single(pkg, engine_factory):
    """pm_match_batch_host cuts a large batch into tapering device passes (here 72 pairs: 8, 18,
    18, 14, 8, 6); the asynchronous call uses uniform passes of 18. Both must give what a pair gives
    on its own with the same pair index, whatever the partition."""
    import torch
    n, w, h, D = 72, 256, 208, 32
    L, R, T = pkg.synth.make_batch(3, 6, w, h, D)
    L = np.ascontiguousarray(np.tile(L, (n // 6, 1, 1)))
    R = np.ascontiguousarray(np.tile(R, (n // 6, 1, 1)))
    e = engine_factory(init_mode="random", max_disp=D)
    bl, br = e.MatchBatch(L, R, first_pair_index=5)            # blocking, tapered
    pL, pR = torch.from_numpy(L).pin_memory(), torch.from_numpy(R).pin_memory()
    al = torch.empty((n, h, w), dtype=torch.float32).pin_memory()
    ar = torch.empty((n, h, w), dtype=torch.float32).pin_memory()
    e.match_batch_host_async(n, pL.data_ptr(), pR.data_ptr(), w, h, w, al.data_ptr(), ar.data_ptr(),
                             w * 4, first_pair_index=5)
    e.wait()
    assert np.array_equal(bl, al.numpy()) and np.array_equal(br, ar.numpy())
    e1 = engine_factory(init_mode="random", max_disp=D)
    for i in (0, 7, 8, 25, 26, 43, 44, 57, 58, 65, 66, 71):   # both sides of every pass boundary
        dl, dr = e1.Match(L[i], R[i], pair_index=5 + i)
        assert np.array_equal(bl[i], dl) and np.array_equal(br[i], dr), i


def test_strided_buffers_and_resize(pmo, pkg, engine_factory):
    """Row strides larger than the width, and an engine reused across resolutions
    (the reference never re-sizes its noise image: SURVEY.md A.4-5)."""
    import ctypes as C
    e = engine_factory(init_mode="random", max_disp=32)
    for (w, h) in ((300, 200), (256, 192)):
        L, R, _ = pkg.synth.make_pair(1, w, h, 32)
        stride = w + 20
        Lp = np.zeros((h, stride), np.uint8); Lp[:, :w] = L
        Rp = np.zeros((h, stride), np.uint8); Rp[:, :w] = R
        ostride = w + 12
        ol = np.full((h, ostride), -1, np.float32); orr = np.full((h, ostride), -1, np.float32)
        rc = e._lib.pm_match_host(e._h, C.c_void_p(Lp.ctypes.data), C.c_void_p(Rp.ctypes.data), w, h,
                                  stride, None, None, 1, C.c_void_p(ol.ctypes.data),
                                  C.c_void_p(orr.ctypes.data), ostride * 4)
        assert rc == 0
        wl, wr = pmo.g_match(pmo.default_params(init_mode=1, max_disp=32), L, R, pair_index=1)
        assert np.array_equal(ol[:, :w], wl) and np.array_equal(orr[:, :w], wr)
        assert np.all(ol[:, w:] == -1)  # padding untouched


def test_error_paths(pkg, engine_factory):
    e = engine_factory()  # init_mode = sparse: seed maps are optional, but come in pairs
    L = np.zeros((64, 64), np.uint8)
    import ctypes as C
    out = np.zeros((64, 64), np.float32)
    rc = e._lib.pm_match_host(e._h, C.c_void_p(L.ctypes.data), C.c_void_p(L.ctypes.data), 64, 64, 64,
                              C.c_void_p(out.ctypes.data), None, 0, C.c_void_p(out.ctypes.data),
                              C.c_void_p(out.ctypes.data), 256)
    assert rc == -1
    e2 = engine_factory(init_mode="random")
    with pytest.raises(pkg.PmError) as ei:
        e2.Match(L, L)  # 64/16 = 4-pixel chunks < 2*overlap+2
    assert ei.value.code == -2
    with pytest.raises(pkg.PmError):
        e2.stage_propagate(0, 1, 1)  # no pair loaded


def test_empty_scene_stays_background(engine_factory):
    """All-zero seeds: noise is masked by d > 0 and nothing can propagate."""
    rng = np.random.default_rng(0)
    L = _blur(rng.integers(0, 256, (200, 320)).astype(np.uint8))
    e = engine_factory()
    z = np.zeros((200, 320), np.float32)
    dl, dr = e.Match(L, L, z, z)
    assert not dl.any() and not dr.any()


@pytest.mark.parametrize("cfg", [
    ("C2", 752, 480, 64, 1),     # BASELINE config C2: 752x480, 64-disparity range, 3 iterations
    ("C3", 1280, 720, 128, 2),   # BASELINE config C3: 1280x720, 128-disparity range, 2 levels
])
def test_baseline_configs_equal_oracle(pmo, pkg, engine_factory, cfg):
    name, w, h, D, levels = cfg
    L, R, T = pkg.synth.make_pair(1, w, h, D)
    e = engine_factory(init_mode="random", max_disp=D, pyramid_levels=levels)
    dl, dr = e.Match(L, R, pair_index=1)
    wl, wr = pmo.g_match(pmo.default_params(init_mode=1, max_disp=D, pyramid_levels=levels), L, R,
                         pair_index=1)
    assert np.array_equal(dl, wl) and np.array_equal(dr, wr), name
    found = (dl > 0) & (T > 0)
    assert (np.abs(dl - T)[found] <= 1.0).mean() > 0.97


# --------------------------------------------------- full-size properties

def test_full_size_properties(pkg, engine_factory):
    """BASELINE config C3 (1280x720, D = 128, 2 levels): size-independent properties."""
    w, h, D = 1280, 720, 128
    L, R, T = pkg.synth.make_batch(0, 2, w, h, D)
    e = engine_factory(init_mode="random", max_disp=D, pyramid_levels=2)
    dl, dr = e.MatchBatch(L, R)
    xs = np.arange(w, dtype=np.float32)[None, None, :]
    assert dl.min() >= 0 and dr.min() >= 0 and np.isfinite(dl).all() and np.isfinite(dr).all()
    assert dl.max() < 2 * D and dr.max() < 2 * D   # init range + noise, never a runaway value
    # left-right consistency of what survives the occlusion mask (ratio test, :292)
    for i in range(2):
        ys, xs_ = np.nonzero(dl[i])
        d = dl[i][ys, xs_]
        xr = np.maximum(xs_.astype(np.float32) - d, 0).astype(np.int64)
        r = dr[i][ys, xr]
        assert np.all((r.astype(np.float64) <= 1.4 * d.astype(np.float64)) &
                      (r.astype(np.float64) >= 0.7 * d.astype(np.float64)))
    # accuracy against the synthetic ground truth
    found = (dl > 0) & (T > 0)
    assert found.mean() > 0.7
    assert (np.abs(dl - T)[found] <= 1.0).mean() > 0.97
    # idempotence of the masks: re-applying MaskOcclusions changes nothing
    again = e.stage_mask_occlusions(dl[0], dr[0])
    assert np.array_equal(again, dl[0])
    # determinism and batch-order independence
    dl2, dr2 = e.MatchBatch(L[::-1].copy(), R[::-1].copy(), first_pair_index=0)
    e.params  # (pair index keys the random init: reversed order with the same indices differs)
    dl3, dr3 = e.MatchBatch(L, R)
    assert np.array_equal(dl, dl3) and np.array_equal(dr, dr3)


def test_swap_and_mirror_symmetry(pmo, engine_factory, c1):
    """Matching the mirrored, swapped pair gives the mirrored right map (the reference
    obtains its right view exactly this way, patchmatch_gpu.cu:357-368)."""
    e = engine_factory()
    il, ir = c1["il"], c1["ir"]
    sl, sr = c1["seed_gpu_l"], c1["seed_gpu_r"]
    dl, dr = e.Match(il, ir, sl, sr)
    f = lambda a: np.ascontiguousarray(a[:, ::-1])
    dl_m, dr_m = e.Match(f(ir), f(il), f(sr), f(sl))
    # dr_m (right map of the mirrored problem) is the mirrored, un-occlusion-masked left map
    unmasked = f(dr_m)
    assert np.array_equal(unmasked[dl > 0], dl[dl > 0])


def test_wide_frame_transposed_row_sweeps(pkg, pmo, engine_factory):
    """Images wider than the shared-memory row kernel can stage (w > 1330) run their row sweeps as
    the column kernel on transposed planes: stage by stage and end to end equal to the oracle."""
    w, h, D = 1424, 208, 48
    L, R, _ = pkg.synth.make_pair(3, w, h, D)
    e = engine_factory(init_mode="random", max_disp=D, patchmatch_iters=2)
    e.stage_load_pair(L, R)
    for view in (0, 1):
        Il, Ir, Gl, Gr = pmo.g_planes(L, R, view)
        d0 = _rand_disp(w, h, 11 + view, hi=40.0)
        for direction in (+1, -1):
            e.stage_set_disp(view, d0)
            e.stage_propagate(view, 1, direction)
            want = pmo.g_propagate(Il, Ir, Gl, Gr, d0, 1, direction)
            assert np.array_equal(e.stage_get_disp(view), want), (view, direction)
    dl, dr = e.Match(L, R, pair_index=2)
    wl, wr = pmo.g_match(pmo.default_params(init_mode=1, max_disp=D, patchmatch_iters=2), L, R, pair_index=2)
    assert np.array_equal(dl, wl) and np.array_equal(dr, wr)
    # a batch, so that several views share one launch
    Ls = np.stack([L, L[::-1].copy()]); Rs = np.stack([R, R[::-1].copy()])
    ol, orr = e.MatchBatch(Ls, Rs, first_pair_index=2)
    assert np.array_equal(ol[0], wl) and np.array_equal(orr[0], wr)
    w2l, w2r = pmo.g_match(pmo.default_params(init_mode=1, max_disp=D, patchmatch_iters=2), Ls[1], Rs[1], pair_index=3)
    assert np.array_equal(ol[1], w2l) and np.array_equal(orr[1], w2r)


def test_full_patch_cost_mode(pkg, pmo, engine_factory, c1):
    """cost_mode l1grad_full (the reference's L1GradientCost, patchmatch_gpu.cu:45-69): stages and
    whole pipelines equal the oracle; runs on the one-thread-per-chain kernels."""
    il, ir = c1["il"], c1["ir"]
    h, w = il.shape
    e = engine_factory(cost_mode="l1grad_full")
    e.stage_load_pair(il, ir)
    try:
        pmo.set_cost_mode(1)
        noise = pmo.rng_uniform(123, -1, 1, w * h).reshape(h, w)
        for view in (0, 1):
            Il, Ir, Gl, Gr = pmo.g_planes(il, ir, view)
            d0 = _rand_disp(w, h, 21 + view, hi=40.0)
            e.stage_set_disp(view, d0)
            _, cost = e.stage_get_disp(view, want_cost=True)
            assert np.array_equal(cost, pmo.g_cost_map(Il, Ir, Gl, Gr, d0, 0.9))
            for along_x, direction in ((1, 1), (0, 1), (1, -1), (0, -1)):
                e.stage_set_disp(view, d0)
                e.stage_propagate(view, along_x, direction)
                assert np.array_equal(e.stage_get_disp(view),
                                      pmo.g_propagate(Il, Ir, Gl, Gr, d0, along_x, direction)), (view, along_x, direction)
            e.stage_set_disp(view, d0)
            e.stage_mask_background(view)
            assert np.array_equal(e.stage_get_disp(view), pmo.g_mask_background(Il, Ir, Gl, Gr, d0))
    finally:
        pmo.set_cost_mode(0)
    dl, dr = e.Match(il, ir, c1["seed_gpu_l"], c1["seed_gpu_r"])
    wl, wr = pmo.g_match(pmo.default_params(cost_mode=1), il, ir, c1["seed_gpu_l"], c1["seed_gpu_r"])
    assert np.array_equal(dl, wl) and np.array_equal(dr, wr)
    e2 = engine_factory(cost_mode="l1grad_full", init_mode="random", max_disp=48, pyramid_levels=2,
                        subpixel=1, noise_accept="improve")
    L, R, _ = pkg.synth.make_pair(1, 512, 400, 48)
    dl, dr = e2.Match(L, R, pair_index=4)
    p = pmo.default_params(cost_mode=1, init_mode=1, max_disp=48, pyramid_levels=2, subpixel=1, noise_accept=1)
    wl, wr = pmo.g_match(p, L, R, pair_index=4)
    assert np.array_equal(dl, wl) and np.array_equal(dr, wr)


def test_two_devices_in_one_process(pkg, pmo):
    """One engine per GPU inside one process (the C++ host's model): kernel attributes are per
    device, both engines must run the shared-memory kernels and agree with the oracle."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    w, h, D = 640, 400, 48
    L, R, _ = pkg.synth.make_pair(9, w, h, D)
    outs = []
    for dev in (0, 1):
        P = pkg.PatchmatchGpu.Params()
        e = pkg.PatchmatchGpu(P, device=dev)        # reference defaults: device SparseInit + sweeps
        outs.append(e.Match(L, R))
        e.close()
    sl, sr = pmo.s_match_seeds(L, R, 4)
    wl, wr = pmo.g_match(pmo.default_params(), L, R, sl, sr)
    for dl, dr in outs:
        assert np.array_equal(dl, wl) and np.array_equal(dr, wr)
