"""The device-pointer entry points of the C ABI against the oracle (-m gpu).

pm_match_batch_device is what bench.py's `value` is measured through: strided device buffers, a
caller-owned stream, more pairs than one device pass holds, with and without seed maps, and
calls on different streams sharing the engine's single workspace.
pm_match_planes_device is the reference's float-plane overload
Match(GpuMat iml, imr, Gl, Gr, GpuMat& disp) (patchmatch_gpu.h:104-108, patchmatch_gpu.cu:379-411).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _dev(a, torch):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _strided(arr, stride, fill, torch):
    """[n, h, w] -> device tensor [n, h, stride] with the payload in [..., :w]"""
    n, h, w = arr.shape
    t = torch.full((n, h, stride), fill, dtype=torch.from_numpy(arr).dtype, device="cuda")
    t[:, :, :w] = torch.from_numpy(arr).cuda()
    return t


@pytest.mark.parametrize("init", ["random", "seeds", "sparse"])
def test_match_batch_device_strided_stream_vs_oracle(pmo, pkg, engine_factory, init):
    import torch
    w, h, D, n = 640, 400, 64, 3
    L, R, _ = pkg.synth.make_batch(5, n, w, h, D)
    e = engine_factory(init_mode="random" if init == "random" else "sparse", max_disp=D,
                       max_batch=2)                       # two device passes: 2 + 1 pairs
    istride, ostride = w + 48, w + 16                     # elements
    dL, dR = _strided(L, istride, 7, torch), _strided(R, istride, 9, torch)
    oL = torch.full((n, h, ostride), -1.0, dtype=torch.float32, device="cuda")
    oR = torch.full((n, h, ostride), -1.0, dtype=torch.float32, device="cuda")
    seeds = None
    if init == "seeds":   # caller-supplied SparseInit maps, same stride as the outputs
        sl = np.zeros((n, h, w), np.float32); sr = np.zeros((n, h, w), np.float32)
        for i in range(n):
            sl[i], sr[i] = pmo.s_match_seeds(L[i], R[i], 4)
        seeds = (_strided(sl, ostride, 0.0, torch), _strided(sr, ostride, 0.0, torch))
    stream = torch.cuda.Stream()
    torch.cuda.synchronize()                              # the fills above ran on the default stream
    with torch.cuda.stream(stream):
        e.match_batch_device(n, dL.data_ptr(), dR.data_ptr(), w, h, istride, oL.data_ptr(),
                             oR.data_ptr(), ostride * 4,
                             d_seed_l=seeds[0].data_ptr() if seeds else None,
                             d_seed_r=seeds[1].data_ptr() if seeds else None,
                             first_pair_index=5, stream=stream.cuda_stream)
    e.synchronize(stream.cuda_stream)
    gl, gr = oL.cpu().numpy(), oR.cpu().numpy()
    assert np.all(gl[:, :, w:] == -1) and np.all(gr[:, :, w:] == -1)      # padding untouched
    for i in range(n):
        if init == "random":
            wl, wr = pmo.g_match(pmo.default_params(init_mode=1, max_disp=D), L[i], R[i], pair_index=5 + i)
        else:
            s_l, s_r = pmo.s_match_seeds(L[i], R[i], 4)
            wl, wr = pmo.g_match(pmo.default_params(), L[i], R[i], s_l, s_r)
        assert np.array_equal(gl[i, :, :w], wl), (init, i, int((gl[i, :, :w] != wl).sum()))
        assert np.array_equal(gr[i, :, :w], wr), (init, i)


def test_device_calls_on_two_streams_share_the_workspace(pmo, pkg, engine_factory):
    """Back-to-back asynchronous calls on DIFFERENT streams, then a host call: the engine orders them
    with events on its single workspace, results equal the oracle's."""
    import torch
    w, h, D = 640, 400, 64
    L, R, _ = pkg.synth.make_batch(1, 2, w, h, D)
    e = engine_factory(init_mode="random", max_disp=D)
    dL, dR = _dev(L, torch), _dev(R, torch)
    outs = [(torch.empty((1, h, w), dtype=torch.float32, device="cuda"),
             torch.empty((1, h, w), dtype=torch.float32, device="cuda")) for _ in range(2)]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    torch.cuda.synchronize()
    for rep in range(3):
        for i in range(2):
            e.match_batch_device(1, dL[i].data_ptr(), dR[i].data_ptr(), w, h, w, outs[i][0].data_ptr(),
                                 outs[i][1].data_ptr(), w * 4, first_pair_index=1 + i,
                                 stream=streams[i].cuda_stream)
    hl, hr = e.Match(L[0], R[0], pair_index=1)            # host call on the engine's own stream
    for s in streams:
        e.synchronize(s.cuda_stream)
    for i in range(2):
        wl, wr = pmo.g_match(pmo.default_params(init_mode=1, max_disp=D), L[i], R[i], pair_index=1 + i)
        assert np.array_equal(outs[i][0][0].cpu().numpy(), wl)
        assert np.array_equal(outs[i][1][0].cpu().numpy(), wr)
        if i == 0:
            assert np.array_equal(hl, wl) and np.array_equal(hr, wr)


def test_device_call_on_the_engines_own_stream(pmo, pkg, engine_factory):
    """stream == NULL: asynchronous on the engine's stream; pm_synchronize(e, NULL) waits for it."""
    import torch
    w, h, D = 512, 320, 48
    L, R, _ = pkg.synth.make_batch(3, 1, w, h, D)
    e = engine_factory(init_mode="random", max_disp=D)
    dL, dR = _dev(L, torch), _dev(R, torch)
    oL = torch.empty((1, h, w), dtype=torch.float32, device="cuda")
    oR = torch.empty((1, h, w), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    e.match_batch_device(1, dL.data_ptr(), dR.data_ptr(), w, h, w, oL.data_ptr(), oR.data_ptr(), w * 4,
                         first_pair_index=3)
    e.synchronize()
    wl, wr = pmo.g_match(pmo.default_params(init_mode=1, max_disp=D), L[0], R[0], pair_index=3)
    assert np.array_equal(oL[0].cpu().numpy(), wl) and np.array_equal(oR[0].cpu().numpy(), wr)


@pytest.mark.parametrize("size", [(376, 240), (1280, 720)])
def test_match_planes_device_vs_oracle_and_reference_kernels(pmo, pkg, engine_factory, c1, size):
    """The reference's float-plane overload: caller-owned Il, Ir, Gl, Gr, disp seed -> result in place.
    Compared with the oracle's pmo_g_match_view and, statement by statement, with the reference's own
    kernels (oracle/_ref, stock 16x16 launch: equal to the lock-step schedule on B200, see
    test_gpu_reference_kernels.py)."""
    import torch
    import pmref
    w, h = size
    if size == (376, 240):
        L, R = c1["il"], c1["ir"]
    else:
        L, R, _ = pkg.synth.make_pair(4, w, h, 128)
    sl, sr = pmo.s_match_seeds(L, R, 4)
    noise = pmo.rng_uniform(123, -1, 1, w * h).reshape(h, w)
    e = engine_factory()
    stride = w + 24                                              # elements: pitched planes
    for view in (0, 1):
        planes = pmo.g_planes(L, R, view)                        # Il, Ir, Gl, Gr as the reference holds them
        seed = sl if view == 0 else np.ascontiguousarray(sr[:, ::-1])
        want = pmo.g_match_view(pmo.default_params(), *planes, noise, seed)
        dev = []
        for pl in planes:
            t = torch.zeros((h, stride), dtype=torch.float32, device="cuda")
            t[:, :w] = torch.from_numpy(pl).cuda()
            dev.append(t)
        disp = torch.full((h, stride), -7.0, dtype=torch.float32, device="cuda")
        disp[:, :w] = torch.from_numpy(seed).cuda()
        stream = torch.cuda.Stream()
        torch.cuda.synchronize()
        e.match_planes_device(dev[0].data_ptr(), dev[1].data_ptr(), dev[2].data_ptr(), dev[3].data_ptr(),
                              w, h, stride * 4, disp.data_ptr(), stride * 4, stream=stream.cuda_stream)
        e.synchronize(stream.cuda_stream)
        got = disp.cpu().numpy()
        assert np.all(got[:, w:] == -7.0)
        assert np.array_equal(got[:, :w], want), (view, int((got[:, :w] != want).sum()))
        ref = pmref.match_view(*planes, noise, seed)             # the reference's kernels, stock launch
        agree = float((ref == got[:, :w]).mean())
        print("planes Match view %d %dx%d: %.4f%% equal to the reference's own kernels" % (view, w, h, 100 * agree))
        assert agree >= 0.999
    # a second size on the same engine, then back: the workspace follows
    e2 = engine_factory(pyramid_levels=2)
    with pytest.raises(pkg.PmError):
        e2.match_planes_device(dev[0].data_ptr(), dev[1].data_ptr(), dev[2].data_ptr(), dev[3].data_ptr(),
                               w, h, stride * 4, disp.data_ptr(), stride * 4)
