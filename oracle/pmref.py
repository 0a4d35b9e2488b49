"""ctypes binding of oracle/_ref/libpm_ref_kernels.so: the reference's own CUDA kernels
(/root/reference/src/vehicle/patchmatch_gpu/patchmatch_gpu.cu:18-295, compiled verbatim behind a
PtrStepSz shim by oracle/ref/build_ref.py).

TEST INFRASTRUCTURE ONLY: loaded by tests/ (-m gpu) to pin the oracle and the CUDA path to the
reference's executable semantics. The product package never imports this module.
"""
import ctypes as C
import importlib.util
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libpm_ref_kernels.so")
SYMBOLS = ("pmref_device_count", "pmref_get_subpixel", "pmref_cost_map", "pmref_propagate",
           "pmref_mask_background", "pmref_mask_occlusions", "pmref_match_view", "pmref_match_view_timed")


def _builder():
    spec = importlib.util.spec_from_file_location("pm_build_ref", os.path.join(_HERE, "ref", "build_ref.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def build(force=False):
    """Builds the library where /root/reference exists; elsewhere returns the prebuilt path or None."""
    return _builder().build(force=force)


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = build()
        if path is None or not os.path.exists(path):
            raise RuntimeError("oracle/_ref/libpm_ref_kernels.so is missing and /root/reference is not "
                               "here to build it from (python oracle/ref/build_ref.py)")
        _lib = C.CDLL(path)
    return _lib


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(C.POINTER(C.c_float))


def _chk(rc, what):
    if rc != 0:
        raise RuntimeError("%s failed (%d)" % (what, rc))


def get_subpixel(im, rows, cols):
    im, p = _f32(im)
    rows, pr = _f32(rows)
    cols, pc = _f32(cols)
    h, w = im.shape
    out = np.empty(rows.size, np.float32)
    _chk(lib().pmref_get_subpixel(p, w, h, pr, pc, rows.size, out.ctypes.data_as(C.POINTER(C.c_float))),
         "pmref_get_subpixel")
    return out


def cost_map(Il, Ir, Gl, Gr, disp, alpha=0.9, ph=0, pw=0):
    """L1GradientCost3x3 (ph = pw = 0) or the generic L1GradientCost at xr = fmaxf(x - d, pw/2) for
    every interior pixel; border = 0."""
    Il, a = _f32(Il); Ir, b = _f32(Ir); Gl, c = _f32(Gl); Gr, d = _f32(Gr); disp, e = _f32(disp)
    h, w = Il.shape
    out = np.empty((h, w), np.float32)
    _chk(lib().pmref_cost_map(a, b, c, d, w, h, e, int(ph), int(pw), C.c_float(alpha),
                              out.ctypes.data_as(C.POINTER(C.c_float))), "pmref_cost_map")
    return out


def propagate(Il, Ir, Gl, Gr, disp, along_x, direction, alpha=0.9, stripes=16, lines=16):
    """PropagateRow (along_x) / PropagateCol with the reference's launch shape; stripes = 1 gives
    one thread per line (no concurrent writers)."""
    Il, a = _f32(Il); Ir, b = _f32(Ir); Gl, c = _f32(Gl); Gr, d = _f32(Gr)
    h, w = Il.shape
    disp = np.array(disp, np.float32, copy=True, order="C")
    _chk(lib().pmref_propagate(a, b, c, d, w, h, disp.ctypes.data_as(C.POINTER(C.c_float)),
                               int(bool(along_x)), int(direction), int(stripes), int(lines),
                               C.c_float(alpha)), "pmref_propagate")
    return disp


def mask_background(Il, Ir, Gl, Gr, disp, alpha=0.9, improve=0.8):
    Il, a = _f32(Il); Ir, b = _f32(Ir); Gl, c = _f32(Gl); Gr, d = _f32(Gr)
    h, w = Il.shape
    disp = np.array(disp, np.float32, copy=True, order="C")
    _chk(lib().pmref_mask_background(a, b, c, d, w, h, disp.ctypes.data_as(C.POINTER(C.c_float)),
                                     C.c_float(alpha), C.c_float(improve)), "pmref_mask_background")
    return disp


def mask_occlusions(displ, dispr):
    displ = np.array(displ, np.float32, copy=True, order="C")
    dispr, pr = _f32(dispr)
    h, w = displ.shape
    _chk(lib().pmref_mask_occlusions(displ.ctypes.data_as(C.POINTER(C.c_float)), pr, w, h),
         "pmref_mask_occlusions")
    return displ


def match_view(Il, Ir, Gl, Gr, unit_noise, disp, iters=3, alpha=0.9, improve=0.8, stripes=16, lines=16,
               do_mask=True):
    """PatchmatchGpu::Match(GpuMat...) (patchmatch_gpu.cu:379-411) on one view."""
    Il, a = _f32(Il); Ir, b = _f32(Ir); Gl, c = _f32(Gl); Gr, d = _f32(Gr); un, n = _f32(unit_noise)
    h, w = Il.shape
    disp = np.array(disp, np.float32, copy=True, order="C")
    _chk(lib().pmref_match_view(a, b, c, d, w, h, n, disp.ctypes.data_as(C.POINTER(C.c_float)),
                                int(iters), C.c_float(alpha), C.c_float(improve), int(stripes),
                                int(lines), int(bool(do_mask))), "pmref_match_view")
    return disp


def match_view_timed(Il, Ir, Gl, Gr, unit_noise, seed, iters=3, alpha=0.9, improve=0.8, reps=5):
    """Milliseconds per view of the reference's own device Match loop (stock 16x16 launches, a device
    sync after each kernel, patchmatch_gpu.cu:394-410) with the planes resident on the GPU."""
    Il, a = _f32(Il); Ir, b = _f32(Ir); Gl, c = _f32(Gl); Gr, d = _f32(Gr); un, n = _f32(unit_noise)
    seed, s = _f32(seed)
    h, w = Il.shape
    ms = C.c_float()
    _chk(lib().pmref_match_view_timed(a, b, c, d, w, h, n, s, int(iters), C.c_float(alpha),
                                      C.c_float(improve), 16, 16, int(reps), C.byref(ms)),
         "pmref_match_view_timed")
    return float(ms.value)
