"""ctypes binding of the CPU oracle (oracle/pm_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py. The product package never
imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libpm_oracle.so")


def build(force=False):
    """gcc build of the C restatement (oracle/Makefile)."""
    deps = [os.path.join(_HERE, f) for f in ("pm_oracle.c", "pm_oracle_seed.c", "pm_oracle.h")]
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(d) for d in deps)):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _LIB_PATH


class Params(C.Structure):
    _fields_ = [
        ("cost_alpha", C.c_float),
        ("patchmatch_iters", C.c_int),
        ("cost_improve_factor", C.c_float),
        ("sweep_chunks", C.c_int),
        ("sweep_overlap", C.c_int),
        ("noise_scale0", C.c_float),
        ("noise_accept", C.c_int),
        ("seed", C.c_uint64),
        ("init_mode", C.c_int),
        ("max_disp", C.c_int),
        ("clamp_disp", C.c_int),
        ("pyramid_levels", C.c_int),
        ("lr_mode", C.c_int),
        ("subpixel", C.c_int),
        ("median_ksize", C.c_int),
        ("cost_mode", C.c_int),
        ("patch_size", C.c_int),
        ("random_search_k", C.c_int),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.pmo_g_cost5.restype = C.c_float
        _lib.pmo_c_cost.restype = C.c_float
        _lib.pmo_philox_u01.restype = C.c_float
        _lib.pmo_g_match.restype = C.c_int
        _lib.pmo_s_good_features.restype = C.c_int
        _lib.pmo_s_match_rectified.restype = C.c_double
    return _lib


def default_params(**kw):
    p = Params()
    lib().pmo_params_default(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a, a.ctypes.data_as(C.POINTER(C.c_uint8))


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(C.POINTER(C.c_float))


def set_cost_mode(mode):
    """0 = 5-tap L1GradientCost3x3 (reference), 1 = full 3x3 L1GradientCost; affects every (G) stage."""
    lib().pmo_set_cost_mode(int(mode))


def set_patch_size(patch_size):
    """3 (reference) or 5: patch radius of every (G) stage and window of cost modes 1 and 2."""
    lib().pmo_set_patch_size(int(patch_size))


def x_random_search(params, Il, Ir, Gl, Gr, disp, pair_index, view, level, iter_global, scale,
                    dmax=float("inf")):
    Il, a = _f32(Il); Ir, b = _f32(Ir); Gl, c = _f32(Gl); Gr, d = _f32(Gr)
    h, w = Il.shape
    disp = np.array(disp, np.float32, copy=True, order="C")
    lib().pmo_x_random_search(C.byref(params), a, b, c, d, w, h, C.c_uint32(pair_index),
                              C.c_uint32(view), C.c_uint32(level), C.c_uint32(iter_global),
                              C.c_float(scale), C.c_float(dmax),
                              disp.ctypes.data_as(C.POINTER(C.c_float)))
    return disp


def rng_uniform(seed, lo, hi, n):
    out = np.empty(n, np.float32)
    lib().pmo_rng_uniform_f32(C.c_uint64(seed), C.c_float(lo), C.c_float(hi),
                              out.ctypes.data_as(C.POINTER(C.c_float)), C.c_size_t(n))
    return out


def resize_half(im):
    im, p = _u8(im)
    h, w = im.shape
    out = np.empty((h // 2, w // 2), np.uint8)
    lib().pmo_resize_half_u8(p, w, h, out.ctypes.data_as(C.POINTER(C.c_uint8)))
    return out


def gradient_mag(im):
    im, p = _u8(im)
    h, w = im.shape
    out = np.empty((h, w), np.float32)
    lib().pmo_gradient_mag_u8(p, w, h, out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def get_rect_subpix_u8(im, pw, ph, cx, cy):
    im, p = _u8(im)
    h, w = im.shape
    out = np.empty((ph, pw), np.uint8)
    lib().pmo_get_rect_subpix_u8(p, w, h, pw, ph, C.c_float(cx), C.c_float(cy),
                                 out.ctypes.data_as(C.POINTER(C.c_uint8)))
    return out


def get_rect_subpix_f32(im, pw, ph, cx, cy):
    im, p = _f32(im)
    h, w = im.shape
    out = np.empty((ph, pw), np.float32)
    lib().pmo_get_rect_subpix_f32(p, w, h, pw, ph, C.c_float(cx), C.c_float(cy),
                                  out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def dilate_rect(im, r):
    im, p = _f32(im)
    h, w = im.shape
    out = np.empty((h, w), np.float32)
    lib().pmo_dilate_rect_f32(p, w, h, r, out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def philox_u01(seed, c0, c1, c2, c3):
    return float(lib().pmo_philox_u01(C.c_uint64(seed), C.c_uint32(c0), C.c_uint32(c1),
                                      C.c_uint32(c2), C.c_uint32(c3)))


# ------------------------------------------------------------ (G) semantics

def g_planes(L, R, view=0):
    """f32 image + gradient planes of one view (view 1 = flipped, swapped)."""
    L = np.ascontiguousarray(L, np.uint8)
    R = np.ascontiguousarray(R, np.uint8)
    if view == 0:
        ref, mat = L, R
        return (ref.astype(np.float32), mat.astype(np.float32), gradient_mag(ref), gradient_mag(mat))
    ref, mat = R, L
    f = lambda a: np.ascontiguousarray(a[:, ::-1])
    return (f(ref.astype(np.float32)), f(mat.astype(np.float32)),
            f(gradient_mag(ref)), f(gradient_mag(mat)))


def g_cost5(Il, Ir, Gl, Gr, y, x, xr, alpha):
    Il, a = _f32(Il); Ir, b = _f32(Ir); Gl, c = _f32(Gl); Gr, d = _f32(Gr)
    h, w = Il.shape
    return float(lib().pmo_g_cost5(a, b, c, d, w, h, int(y), int(x), C.c_float(xr), C.c_float(alpha)))


def g_cost_map(Il, Ir, Gl, Gr, disp, alpha):
    """cost(d) at every interior pixel with xr = max(x - d, 1); border = 0."""
    h, w = Il.shape
    out = np.zeros((h, w), np.float32)
    Il, a = _f32(Il); Ir, b = _f32(Ir); Gl, c = _f32(Gl); Gr, d = _f32(Gr); disp, e = _f32(disp)
    lib().pmo_g_cost_map(a, b, c, d, w, h, e, C.c_float(alpha), out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def g_add_noise(disp, unit_noise, scale):
    disp = np.array(disp, np.float32, copy=True, order="C")
    un, pn = _f32(unit_noise)
    lib().pmo_g_add_noise(disp.ctypes.data_as(C.POINTER(C.c_float)), pn, C.c_size_t(disp.size),
                          C.c_float(scale))
    return disp


def g_propagate(Il, Ir, Gl, Gr, disp, along_x, direction, alpha=0.9, chunks=16, overlap=5):
    Il, a = _f32(Il); Ir, b = _f32(Ir); Gl, c = _f32(Gl); Gr, d = _f32(Gr)
    h, w = Il.shape
    disp = np.array(disp, np.float32, copy=True, order="C")
    fn = lib().pmo_g_propagate_row if along_x else lib().pmo_g_propagate_col
    fn(a, b, c, d, w, h, disp.ctypes.data_as(C.POINTER(C.c_float)), int(direction),
       C.c_float(alpha), int(chunks), int(overlap))
    return disp


def g_sweep_chains(Il, Ir, Gl, Gr, disp, cost, along_x, direction, alpha=0.9, chunks=16, overlap=5):
    Il, a = _f32(Il); Ir, b = _f32(Ir); Gl, c = _f32(Gl); Gr, d = _f32(Gr)
    h, w = Il.shape
    di, pdi = _f32(disp); ci, pci = _f32(cost)
    do = np.empty((h, w), np.float32); co = np.empty((h, w), np.float32)
    lib().pmo_g_sweep_chains(a, b, c, d, w, h, pdi, pci, do.ctypes.data_as(C.POINTER(C.c_float)),
                             co.ctypes.data_as(C.POINTER(C.c_float)), int(along_x), int(direction),
                             C.c_float(alpha), int(chunks), int(overlap))
    return do, co


def g_mask_background(Il, Ir, Gl, Gr, disp, alpha=0.9, improve=0.8):
    Il, a = _f32(Il); Ir, b = _f32(Ir); Gl, c = _f32(Gl); Gr, d = _f32(Gr)
    h, w = Il.shape
    disp = np.array(disp, np.float32, copy=True, order="C")
    lib().pmo_g_mask_background(a, b, c, d, w, h, disp.ctypes.data_as(C.POINTER(C.c_float)),
                                C.c_float(alpha), C.c_float(improve))
    return disp


def g_mask_occlusions(displ, dispr, lr_mode=0):
    displ = np.array(displ, np.float32, copy=True, order="C")
    dispr, pr = _f32(dispr)
    h, w = displ.shape
    lib().pmo_g_mask_occlusions(displ.ctypes.data_as(C.POINTER(C.c_float)), pr, w, h, int(lr_mode))
    return displ


def g_match_view(params, Il, Ir, Gl, Gr, unit_noise, disp, level_scale=1.0, iter0=0, do_mask=1):
    Il, a = _f32(Il); Ir, b = _f32(Ir); Gl, c = _f32(Gl); Gr, d = _f32(Gr)
    un, pn = _f32(unit_noise)
    h, w = Il.shape
    disp = np.array(disp, np.float32, copy=True, order="C")
    lib().pmo_g_match_view(C.byref(params), a, b, c, d, w, h, pn, C.c_float(level_scale),
                           int(iter0), int(do_mask), disp.ctypes.data_as(C.POINTER(C.c_float)))
    return disp


def g_match(params, L, R, seed_l=None, seed_r=None, pair_index=0):
    L, pl = _u8(L); R, pr = _u8(R)
    h, w = L.shape
    dl = np.zeros((h, w), np.float32)
    dr = np.zeros((h, w), np.float32)
    sl = sr = None
    psl = psr = None
    if seed_l is not None:
        sl, psl = _f32(seed_l)
        sr, psr = _f32(seed_r)
    rc = lib().pmo_g_match(C.byref(params), pl, pr, w, h, psl, psr, C.c_uint32(pair_index),
                           dl.ctypes.data_as(C.POINTER(C.c_float)),
                           dr.ctypes.data_as(C.POINTER(C.c_float)))
    if rc != 0:
        raise RuntimeError("pmo_g_match failed: %d" % rc)
    return dl, dr


def x_random_init(params, w, h, pair_index, view, level, rng):
    out = np.empty((h, w), np.float32)
    lib().pmo_x_random_init(C.byref(params), w, h, C.c_uint32(pair_index), C.c_uint32(view),
                            C.c_uint32(level), C.c_float(rng),
                            out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def x_median(disp, k):
    disp, p = _f32(disp)
    h, w = disp.shape
    out = np.empty((h, w), np.float32)
    lib().pmo_x_median(p, w, h, int(k), out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def x_subpixel(Il, Ir, Gl, Gr, disp, alpha=0.9):
    Il, a = _f32(Il); Ir, b = _f32(Ir); Gl, c = _f32(Gl); Gr, d = _f32(Gr)
    h, w = Il.shape
    disp = np.array(disp, np.float32, copy=True, order="C")
    lib().pmo_x_subpixel(a, b, c, d, w, h, C.c_float(alpha),
                         disp.ctypes.data_as(C.POINTER(C.c_float)))
    return disp


def x_upsample2(src, w, h):
    src, p = _f32(src)
    sh, sw = src.shape
    out = np.empty((h, w), np.float32)
    lib().pmo_x_upsample2(p, sw, sh, w, h, out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


# ------------------------------------------------------------ (C) semantics

def c_cost(Il, Ir, Gl, Gr, x, y, d, pw, ph):
    Il, a = _u8(Il); Ir, b = _u8(Ir); Gl, c = _f32(Gl); Gr, e = _f32(Gr)
    h, w = Il.shape
    return float(lib().pmo_c_cost(a, b, c, e, w, h, int(x), int(y), C.c_float(d), int(pw), int(ph)))


def c_add_noise(disp, amount):
    disp = np.array(disp, np.float32, copy=True, order="C")
    h, w = disp.shape
    lib().pmo_c_add_noise(disp.ctypes.data_as(C.POINTER(C.c_float)), w, h, C.c_float(amount))
    return disp


def c_propagate(Il, Ir, Gl, Gr, disp, ph, pw, passes=None):
    Il, a = _u8(Il); Ir, b = _u8(Ir); Gl, c = _f32(Gl); Gr, e = _f32(Gr)
    h, w = Il.shape
    disp = np.array(disp, np.float32, copy=True, order="C")
    pd = disp.ctypes.data_as(C.POINTER(C.c_float))
    if passes is None:
        lib().pmo_c_propagate(a, b, c, e, w, h, pd, int(ph), int(pw))
    else:
        for ps in passes:
            lib().pmo_c_propagate_pass(a, b, c, e, w, h, pd, int(ph), int(pw), int(ps))
    return disp


def c_remove_background(Il, Ir, Gl, Gr, disp, ph, pw, win_by_factor=2.0):
    Il, a = _u8(Il); Ir, b = _u8(Ir); Gl, c = _f32(Gl); Gr, e = _f32(Gr)
    h, w = Il.shape
    disp = np.array(disp, np.float32, copy=True, order="C")
    lib().pmo_c_remove_background(a, b, c, e, w, h, disp.ctypes.data_as(C.POINTER(C.c_float)),
                                  int(ph), int(pw), C.c_float(win_by_factor))
    return disp


def c_estimate_disparity(Il, Ir, seed):
    Il, a = _u8(Il); Ir, b = _u8(Ir)
    h, w = Il.shape
    disp = np.array(seed, np.float32, copy=True, order="C")
    lib().pmo_c_estimate_disparity(a, b, w, h, disp.ctypes.data_as(C.POINTER(C.c_float)))
    return disp


# ------------------------------------------------------------ (S) seeding

class SeedParams(C.Structure):
    _fields_ = [
        ("max_features", C.c_int), ("min_distance", C.c_int), ("quality_level", C.c_double),
        ("block_size", C.c_int), ("use_harris", C.c_int), ("harris_k", C.c_double),
        ("templ_cols", C.c_int), ("templ_rows", C.c_int), ("max_disp", C.c_int),
        ("max_matching_cost", C.c_double),
    ]


def seed_params(**kw):
    p = SeedParams()
    lib().pmo_seed_params_default(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def s_corner_response(im, block=5, harris=False, k=0.04):
    im, p = _u8(im)
    h, w = im.shape
    out = np.empty((h, w), np.float32)
    lib().pmo_s_corner_response(p, w, h, int(block), int(harris), C.c_double(k),
                                out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def s_good_features(im, sp=None):
    """[(x, y)] in selection order, number of local-maximum candidates."""
    sp = sp or seed_params()
    im, p = _u8(im)
    h, w = im.shape
    kx = np.zeros(sp.max_features, np.int32)
    ky = np.zeros(sp.max_features, np.int32)
    nc = C.c_int()
    n = lib().pmo_s_good_features(p, w, h, C.byref(sp), kx.ctypes.data_as(C.POINTER(C.c_int)),
                                  ky.ctypes.data_as(C.POINTER(C.c_int)), C.byref(nc))
    return np.stack([kx[:n], ky[:n]], 1), int(nc.value)


def s_match_rectified(L, R, kps, sp=None):
    sp = sp or seed_params()
    L, pl = _u8(L); R, pr = _u8(R)
    h, w = L.shape
    return np.array([lib().pmo_s_match_rectified(pl, pr, w, h, C.byref(sp), int(x), int(y))
                     for x, y in kps], np.float64)


def s_sparse_init(L, R, dilate_factor=4, sp=None):
    sp = sp or seed_params()
    L, pl = _u8(L); R, pr = _u8(R)
    h, w = L.shape
    out = np.empty((h, w), np.float32)
    lib().pmo_s_sparse_init(pl, pr, w, h, C.byref(sp), int(dilate_factor),
                            out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def c_initialize(L, R, downsample_factor=1, sp=None):
    sp = sp or seed_params()
    L, pl = _u8(L); R, pr = _u8(R)
    h, w = L.shape
    f = int(downsample_factor)
    out = np.empty((h // f, w // f), np.float32)
    lib().pmo_c_initialize(pl, pr, w, h, C.byref(sp), f, out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def s_match_seeds(L, R, dilate_factor=4, sp=None):
    sp = sp or seed_params()
    L, pl = _u8(L); R, pr = _u8(R)
    h, w = L.shape
    sl = np.empty((h, w), np.float32)
    sr = np.empty((h, w), np.float32)
    lib().pmo_s_match_seeds(pl, pr, w, h, C.byref(sp), int(dilate_factor),
                            sl.ctypes.data_as(C.POINTER(C.c_float)),
                            sr.ctypes.data_as(C.POINTER(C.c_float)))
    return sl, sr


def x_disp_to_depth(disp, fx, fy, cx, cy, baseline, scale=1.0):
    """(depth [h,w], xyz [h,w,3]) of StereoCamera::DispToDepth / PinholeCamera::Backproject."""
    disp, p = _f32(disp)
    h, w = disp.shape
    depth = np.empty((h, w), np.float32)
    xyz = np.empty((h, w, 3), np.float32)
    lib().pmo_x_disp_to_depth(p, w, h, C.c_double(fx), C.c_double(fy), C.c_double(cx), C.c_double(cy),
                              C.c_double(baseline), C.c_double(scale),
                              depth.ctypes.data_as(C.POINTER(C.c_float)),
                              xyz.ctypes.data_as(C.POINTER(C.c_float)))
    return depth, xyz


def c_foreground_texture_mask(gray, ksize=7, min_grad=35.0, downsize=2):
    gray, p = _u8(gray)
    h, w = gray.shape
    out = np.empty((h, w), np.uint8)
    rc = lib().pmo_c_foreground_texture_mask(p, w, h, int(ksize), C.c_double(min_grad), int(downsize),
                                             out.ctypes.data_as(C.POINTER(C.c_uint8)))
    if rc != 0:
        raise ValueError("ForegroundTextureMask: unsupported arguments (%d)" % rc)
    return out
