"""ctypes binding of libpm_b200.so and the PatchmatchGpu class on top of it.

Mirrors the reference's C++ interface (citations relative to /root/reference):
  PatchmatchGpu::Params          src/vehicle/patchmatch_gpu/patchmatch_gpu.h:79-92
  PatchmatchGpu(const Params&)   patchmatch_gpu.h:96
  Match(iml, imr, disp, dispr)   patchmatch_gpu.h:99-102, patchmatch_gpu.cu:331-376
numpy arrays stand in for cv::Mat: uint8 HxW images in, float32 HxW disparities out.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "lib", os.environ.get("PM_B200_LIB", "libpm_b200.so"))

PM_N_STAGES = 10


class PmError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("pm_b200 error %d: %s" % (code, message))
        self.code = code


class CParams(C.Structure):
    """struct pm_params (include/pm_b200.h)."""
    _fields_ = [
        ("cost_alpha", C.c_float), ("patchmatch_iters", C.c_int),
        ("init_dilate_factor", C.c_int), ("cost_improve_factor", C.c_float),
        ("sm_templ_cols", C.c_int), ("sm_templ_rows", C.c_int), ("sm_max_disp", C.c_int),
        ("sm_max_matching_cost", C.c_double), ("sm_bidirectional", C.c_int),
        ("sm_subpixel_refinement", C.c_int),
        ("fd_max_features_per_frame", C.c_int), ("fd_min_distance", C.c_int),
        ("fd_gftt_quality_level", C.c_double), ("fd_gftt_block_size", C.c_int),
        ("fd_gftt_use_harris", C.c_int), ("fd_gftt_k", C.c_double),
        ("patch_size", C.c_int), ("sweep_chunks", C.c_int), ("sweep_overlap", C.c_int),
        ("noise_scale0", C.c_float), ("seed", C.c_uint64),
        ("init_mode", C.c_int), ("max_disp", C.c_int), ("clamp_disp", C.c_int),
        ("pyramid_levels", C.c_int), ("cost_mode", C.c_int), ("lr_mode", C.c_int),
        ("noise_accept", C.c_int), ("subpixel", C.c_int), ("median_ksize", C.c_int),
        ("max_batch", C.c_int), ("random_search_k", C.c_int),
    ]


class CStereoRig(C.Structure):
    """struct pm_stereo_rig (include/pm_b200.h)."""
    _fields_ = [("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
                ("baseline", C.c_double)]


class StereoCamera:
    """core::StereoCamera (vision_core/stereo_camera.hpp:10-45) reduced to what the dense path
    needs: the left camera's intrinsics, the nominal image height and the baseline
    (config/shared/ZEDMini.yaml:39-60)."""

    def __init__(self, fx, fy, cx, cy, baseline, height=None, width=None):
        self.fx, self.fy, self.cx, self.cy, self.baseline = fx, fy, cx, cy, baseline
        self.height, self.width = height, width

    def DispToDepth(self, disp):       # stereo_camera.cpp:49-53
        if not disp > 0:
            raise ValueError("Cannot convert zero disparity to depth (inf)!")
        return self.fx * self.baseline / disp

    def DepthToDisp(self, depth):      # stereo_camera.cpp:56-60
        if not depth > 0:
            raise ValueError("depth must be positive")
        return self.fx * self.baseline / depth

    def to_c(self):
        return CStereoRig(self.fx, self.fy, self.cx, self.cy, self.baseline)


class CBandLayout(C.Structure):
    """struct pm_band_layout (include/pm_b200.h)."""
    _fields_ = [("own_lo", C.c_int), ("own_hi", C.c_int), ("load_lo", C.c_int), ("load_hi", C.c_int),
                ("k_lo", C.c_int), ("nk", C.c_int)]


class CBandXfer(C.Structure):
    """struct pm_band_xfer (include/pm_b200.h)."""
    _fields_ = [("send_prev", C.c_void_p), ("send_prev_bytes", C.c_size_t),
                ("recv_prev", C.c_void_p), ("recv_prev_bytes", C.c_size_t),
                ("send_next", C.c_void_p), ("send_next_bytes", C.c_size_t),
                ("recv_next", C.c_void_p), ("recv_next_bytes", C.c_size_t)]


_lib = None


def lib_path():
    return _LIB


def load_library():
    """Loads lib/libpm_b200.so. Raises if it has not been built: there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB):
        raise PmError(-3, "%s is missing: build it with `python ocean-perception_b200/build.py` "
                          "(nvcc, sm_100a); this package has no CPU or PyTorch fallback" % _LIB)
    lib = C.CDLL(_LIB)
    lib.pm_last_error.restype = C.c_char_p
    lib.pm_last_error.argtypes = [C.c_void_p]
    lib.pm_stage_name.restype = C.c_char_p
    vp, u8p, f32p = C.c_void_p, C.c_void_p, C.c_void_p
    lib.pm_create.argtypes = [C.POINTER(CParams), C.c_int, C.POINTER(C.c_void_p)]
    lib.pm_destroy.argtypes = [vp]
    lib.pm_get_params.argtypes = [vp, C.POINTER(CParams)]
    lib.pm_params_default.argtypes = [C.POINTER(CParams)]
    lib.pm_params_load_yaml.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(CParams), C.c_char_p, C.c_size_t]
    lib.pm_match_host.argtypes = [vp, u8p, u8p, C.c_int, C.c_int, C.c_size_t, f32p, f32p,
                                  C.c_uint32, f32p, f32p, C.c_size_t]
    lib.pm_match_batch_host.argtypes = [vp, C.c_int, u8p, u8p, C.c_int, C.c_int, C.c_size_t, f32p,
                                        f32p, C.c_uint32, f32p, f32p, C.c_size_t]
    lib.pm_match_batch_host_async.argtypes = lib.pm_match_batch_host.argtypes
    lib.pm_wait.argtypes = [vp]
    lib.pm_measure_fp32_peak.argtypes = [vp, C.POINTER(C.c_double)]
    lib.pm_mesh_vertices_host.argtypes = [vp, f32p, C.c_int, C.c_int, C.c_size_t, u8p, C.c_size_t, f32p,
                                          C.c_int, C.POINTER(CStereoRig), C.c_double, f32p, f32p]
    lib.pm_foreground_texture_mask_host.argtypes = [vp, u8p, C.c_int, C.c_int, C.c_size_t, C.c_int,
                                                    C.c_double, C.c_int, u8p, C.c_size_t]
    lib.pm_match_batch_device.argtypes = [vp, C.c_int, u8p, u8p, C.c_int, C.c_int, C.c_size_t, f32p,
                                          f32p, C.c_uint32, f32p, f32p, C.c_size_t, vp]
    lib.pm_synchronize.argtypes = [vp, vp]
    lib.pm_match_planes_device.argtypes = [vp, f32p, f32p, f32p, f32p, C.c_int, C.c_int, C.c_size_t,
                                           f32p, C.c_size_t, vp]
    lib.pm_host_alloc.argtypes = [C.c_size_t, C.POINTER(C.c_void_p)]
    lib.pm_host_free.argtypes = [vp]
    lib.pm_launch_count.argtypes = [vp, C.POINTER(C.c_uint64)]
    lib.pm_launch_count_reset.argtypes = [vp]
    lib.pm_set_profiling.argtypes = [vp, C.c_int]
    lib.pm_last_stage_ms.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_uint32)]
    lib.pm_stage_load_pair.argtypes = [vp, u8p, u8p, C.c_int, C.c_int, C.c_size_t]
    lib.pm_stage_get_planes.argtypes = [vp, C.c_int, f32p, f32p, f32p, f32p]
    lib.pm_stage_noise_image.argtypes = [vp, C.c_int, C.c_int, f32p]
    lib.pm_stage_set_disp.argtypes = [vp, C.c_int, f32p]
    lib.pm_stage_get_disp.argtypes = [vp, C.c_int, f32p, f32p]
    lib.pm_stage_add_noise.argtypes = [vp, C.c_int, C.c_float]
    lib.pm_stage_propagate.argtypes = [vp, C.c_int, C.c_int, C.c_int]
    lib.pm_stage_mask_background.argtypes = [vp, C.c_int]
    lib.pm_stage_mask_occlusions.argtypes = [vp, f32p, f32p, C.c_int, C.c_int]
    lib.pm_stage_downscale2.argtypes = [vp, u8p, C.c_int, C.c_int, u8p]
    lib.pm_stage_random_init.argtypes = [vp, C.c_int, C.c_uint32, C.c_uint32, C.c_float]
    lib.pm_stage_subpixel.argtypes = [vp, C.c_int]
    lib.pm_stage_median.argtypes = [vp, f32p, C.c_int, C.c_int, C.c_int, f32p]
    lib.pm_sparse_init_host.argtypes = [vp, u8p, u8p, C.c_int, C.c_int, C.c_size_t, C.c_int, f32p,
                                        C.c_size_t]
    lib.pm_cpu_initialize.argtypes = [vp, u8p, u8p, C.c_int, C.c_int, C.c_size_t, C.c_int, f32p,
                                      C.c_size_t]
    lib.pm_stage_corner_response.argtypes = [vp, u8p, C.c_int, C.c_int, C.c_size_t, f32p]
    lib.pm_stage_detect.argtypes = [vp, u8p, C.c_int, C.c_int, C.c_size_t, C.c_int, vp,
                                    C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.pm_stage_match_rectified.argtypes = [vp, u8p, u8p, C.c_int, C.c_int, C.c_size_t, vp, C.c_int,
                                             vp]
    lib.pm_disp_to_depth_host.argtypes = [vp, f32p, C.c_int, C.c_int, C.c_size_t, C.POINTER(CStereoRig),
                                          C.c_double, f32p, f32p]
    lib.pm_disp_to_depth_device.argtypes = [vp, C.c_int, f32p, C.c_int, C.c_int, C.c_size_t,
                                            C.POINTER(CStereoRig), C.c_double, f32p, C.c_size_t, f32p, vp]
    lib.pm_cpu_set_disp.argtypes = [vp, f32p]
    lib.pm_cpu_get_disp.argtypes = [vp, f32p]
    lib.pm_cpu_add_noise.argtypes = [vp, C.c_float]
    lib.pm_cpu_propagate.argtypes = [vp, C.c_int, C.c_int, C.c_int]
    lib.pm_cpu_remove_background.argtypes = [vp, C.c_int, C.c_int, C.c_float]
    lib.pm_cpu_estimate_disparity.argtypes = [vp, u8p, u8p, C.c_int, C.c_int, C.c_size_t, f32p, f32p,
                                              C.c_size_t]
    lib.pm_cpu_cost.argtypes = [vp, C.c_int, vp, vp, vp, vp, f32p]
    lib.pm_band_plan.argtypes = [C.POINTER(CParams), C.c_int, C.c_int, C.c_int, C.POINTER(CBandLayout)]
    lib.pm_band_exchange_rows.argtypes = [C.POINTER(CParams), C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.POINTER(C.c_int * 8)]
    lib.pm_band_begin.argtypes = [vp, u8p, u8p, C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_int, f32p,
                                  f32p, C.c_size_t, C.c_uint32, vp]
    lib.pm_band_step.argtypes = [vp, C.POINTER(CBandXfer)]
    lib.pm_band_finish.argtypes = [vp, f32p, f32p, C.c_size_t]
    lib.pm_band_p2p_export.argtypes = [vp, C.c_int, vp, C.POINTER(C.c_void_p)]
    lib.pm_band_p2p_connect.argtypes = [vp, vp, vp, vp, vp]
    lib.pm_band_p2p_disable.argtypes = [vp]
    if lib.pm_abi_version() != 2:
        raise PmError(-1, "ABI version mismatch")
    _lib = lib
    return lib


class StereoMatcherParams:
    """ft::StereoMatcher::Params, feature_tracking/stereo_matcher.hpp:18-30."""

    def __init__(self):
        self.templ_cols = 31
        self.templ_rows = 11
        self.max_disp = 128
        self.max_matching_cost = 0.15
        self.bidirectional = False
        self.subpixel_refinement = False


class FeatureDetectorParams:
    """ft::FeatureDetector::Params, feature_tracking/feature_detector.hpp:26-51."""

    def __init__(self):
        self.max_features_per_frame = 200
        self.min_distance_btw_tracked_and_detected_features = 20
        self.gftt_quality_level = 0.01
        self.gftt_block_size = 5
        self.gftt_use_harris_corner_detector = False
        self.gftt_k = 0.04


_INIT = {"sparse": 0, "seeds": 0, "random": 1}
_LR = {"ratio": 0, "abs1px": 1}
_NOISE = {"always": 0, "improve": 1}
_COST = {"l1grad_x5": 0, "l1grad_full": 1, "census": 2}


def _enum(v, table):
    return table[v] if isinstance(v, str) else int(v)


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class PatchmatchGpu:
    """bm::pm::PatchmatchGpu (patchmatch_gpu.h:77-124)."""

    class Params:
        """PatchmatchGpu::Params (patchmatch_gpu.h:79-92) plus the extension keys of
        include/pm_b200.h; defaults reproduce the reference."""

        def __init__(self, yaml_path=None, subtree=""):
            self.detector_params = FeatureDetectorParams()
            self.matcher_params = StereoMatcherParams()
            self.cost_alpha = 0.9
            self.patchmatch_iters = 3
            self.init_dilate_factor = 4
            self.cost_improve_factor = 0.8
            self.patch_size = 3
            self.sweep_chunks = 16
            self.sweep_overlap = 5
            self.noise_scale0 = 32.0
            self.seed = 123
            self.init_mode = "sparse"
            self.max_disp = 128
            self.clamp_disp = 0
            self.pyramid_levels = 1
            self.cost_mode = "l1grad_x5"
            self.lr_mode = "ratio"
            self.noise_accept = "always"
            self.subpixel = 0
            self.median_ksize = 0
            self.max_batch = 0
            self.random_search_k = 0
            if yaml_path is not None:
                self._load_yaml(yaml_path, subtree)

        def _load_yaml(self, path, subtree):
            c = CParams()
            err = C.create_string_buffer(512)
            rc = load_library().pm_params_load_yaml(os.fsencode(path), subtree.encode(), C.byref(c),
                                                    err, 512)
            if rc != 0:
                raise PmError(rc, err.value.decode())
            self._from_c(c)

        def _from_c(self, c):
            d, m = self.detector_params, self.matcher_params
            m.templ_cols, m.templ_rows, m.max_disp = c.sm_templ_cols, c.sm_templ_rows, c.sm_max_disp
            m.max_matching_cost = c.sm_max_matching_cost
            m.bidirectional = bool(c.sm_bidirectional)
            m.subpixel_refinement = bool(c.sm_subpixel_refinement)
            d.max_features_per_frame = c.fd_max_features_per_frame
            d.min_distance_btw_tracked_and_detected_features = c.fd_min_distance
            d.gftt_quality_level = c.fd_gftt_quality_level
            d.gftt_block_size = c.fd_gftt_block_size
            d.gftt_use_harris_corner_detector = bool(c.fd_gftt_use_harris)
            d.gftt_k = c.fd_gftt_k
            for k in ("cost_alpha", "patchmatch_iters", "init_dilate_factor", "cost_improve_factor",
                      "patch_size", "sweep_chunks", "sweep_overlap", "noise_scale0", "seed",
                      "max_disp", "clamp_disp", "pyramid_levels", "subpixel", "median_ksize",
                      "max_batch", "random_search_k"):
                setattr(self, k, getattr(c, k))
            self.init_mode = ["sparse", "random"][c.init_mode]
            self.lr_mode = ["ratio", "abs1px"][c.lr_mode]
            self.noise_accept = ["always", "improve"][c.noise_accept]
            self.cost_mode = ["l1grad_x5", "l1grad_full", "census"][c.cost_mode]

        def to_c(self):
            c = CParams()
            d, m = self.detector_params, self.matcher_params
            c.cost_alpha = self.cost_alpha
            c.patchmatch_iters = self.patchmatch_iters
            c.init_dilate_factor = self.init_dilate_factor
            c.cost_improve_factor = self.cost_improve_factor
            c.sm_templ_cols, c.sm_templ_rows, c.sm_max_disp = m.templ_cols, m.templ_rows, int(m.max_disp)
            c.sm_max_matching_cost = m.max_matching_cost
            c.sm_bidirectional = int(m.bidirectional)
            c.sm_subpixel_refinement = int(m.subpixel_refinement)
            c.fd_max_features_per_frame = d.max_features_per_frame
            c.fd_min_distance = d.min_distance_btw_tracked_and_detected_features
            c.fd_gftt_quality_level = d.gftt_quality_level
            c.fd_gftt_block_size = d.gftt_block_size
            c.fd_gftt_use_harris = int(d.gftt_use_harris_corner_detector)
            c.fd_gftt_k = d.gftt_k
            c.patch_size = self.patch_size
            c.sweep_chunks = self.sweep_chunks
            c.sweep_overlap = self.sweep_overlap
            c.noise_scale0 = self.noise_scale0
            c.seed = self.seed
            c.init_mode = _enum(self.init_mode, _INIT)
            c.max_disp = self.max_disp
            c.clamp_disp = int(self.clamp_disp)
            c.pyramid_levels = self.pyramid_levels
            c.cost_mode = _enum(self.cost_mode, _COST)
            c.lr_mode = _enum(self.lr_mode, _LR)
            c.noise_accept = _enum(self.noise_accept, _NOISE)
            c.subpixel = int(self.subpixel)
            c.median_ksize = self.median_ksize
            c.max_batch = self.max_batch
            c.random_search_k = int(self.random_search_k)
            return c

    def __init__(self, params=None, device=0):
        self._lib = load_library()
        self._h = C.c_void_p()
        self.params = params if params is not None else PatchmatchGpu.Params()
        c = self.params.to_c()
        rc = self._lib.pm_create(C.byref(c), int(device), C.byref(self._h))
        if rc != 0:
            raise PmError(rc, self._lib.pm_last_error(None).decode())
        self.device = device

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.pm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise PmError(rc, self._lib.pm_last_error(self._h).decode())

    # ---- SparseInit, patchmatch_gpu.h:110-112, patchmatch_gpu.cu:414-442
    def SparseInit(self, iml, imr, dilate_factor=None):
        """GFTT keypoints of iml matched along the rows of imr, scattered and dilated."""
        iml = np.ascontiguousarray(iml, np.uint8)
        imr = np.ascontiguousarray(imr, np.uint8)
        h, w = iml.shape
        f = self.params.init_dilate_factor if dilate_factor is None else int(dilate_factor)
        out = np.empty((h, w), np.float32)
        self._check(self._lib.pm_sparse_init_host(self._h, _ptr(iml), _ptr(imr), w, h, w, f,
                                                  _ptr(out), w * 4))
        return out

    # ---- ForegroundTextureMask, stereo_matching/patchmatch.hpp:22-26, patchmatch.cpp:19-49
    def ForegroundTextureMask(self, gray, ksize=7, min_grad=35.0, downsize=2):
        """uint8 mask: non-zero = textured foreground (morphological gradient above min_grad)."""
        gray = np.ascontiguousarray(gray, np.uint8)
        h, w = gray.shape
        mask = np.empty((h, w), np.uint8)
        self._check(self._lib.pm_foreground_texture_mask_host(self._h, _ptr(gray), w, h, w, int(ksize),
                                                              float(min_grad), int(downsize), _ptr(mask), w))
        return mask

    def MeshVertices(self, disp, keypoints, rig, mask=None):
        """Mesh vertices the way ObjectMesher::BuildTriangleMesh makes them (object_mesher.cpp:139-150),
        with the disparities read from the dense map at the keypoints: (vertex_disps [n], xyz [n, 3])."""
        disp = np.ascontiguousarray(disp, np.float32)
        kp = np.ascontiguousarray(keypoints, np.float32).reshape(-1, 2)
        h, w = disp.shape
        scale = float(h) / float(rig.height) if rig.height else 1.0
        if mask is not None:
            mask = np.ascontiguousarray(mask, np.uint8)
        vd = np.empty(len(kp), np.float32)
        xyz = np.empty((len(kp), 3), np.float32)
        c = rig.to_c()
        self._check(self._lib.pm_mesh_vertices_host(self._h, _ptr(disp), w, h, w * 4, _ptr(mask), w, _ptr(kp),
                                                    len(kp), C.byref(c), scale, _ptr(vd), _ptr(xyz)))
        return vd, xyz

    # ---- the consumer of the maps: metric depth and points in the left camera's RDF frame
    def DispToDepth(self, disp, rig, want_points=False):
        """Per pixel StereoCamera::DispToDepth (and PinholeCamera::Backproject) with ObjectMesher's
        scale handling (mesher/object_mesher.cpp:147-150): scale = map height / rig height."""
        disp = np.ascontiguousarray(disp, np.float32)
        h, w = disp.shape
        scale = float(h) / float(rig.height) if rig.height else 1.0
        depth = np.empty((h, w), np.float32)
        xyz = np.empty((h, w, 3), np.float32) if want_points else None
        c = rig.to_c()
        self._check(self._lib.pm_disp_to_depth_host(self._h, _ptr(disp), w, h, w * 4, C.byref(c), scale,
                                                    _ptr(depth), _ptr(xyz)))
        return (depth, xyz) if want_points else depth

    def disp_to_depth_device(self, n, d_disp, w, h, disp_stride, rig, scale, d_depth=None,
                             depth_stride=0, d_xyz=None, stream=None):
        """pm_disp_to_depth_device: raw device pointers (ints), asynchronous on `stream`."""
        c = rig.to_c()
        self._check(self._lib.pm_disp_to_depth_device(
            self._h, n, C.c_void_p(d_disp), w, h, disp_stride, C.byref(c), float(scale),
            C.c_void_p(d_depth) if d_depth else None, depth_stride,
            C.c_void_p(d_xyz) if d_xyz else None, C.c_void_p(stream) if stream else None))

    # ---- Match (host images), patchmatch_gpu.cu:331-376
    def Match(self, iml, imr, seed_l=None, seed_r=None, pair_index=0):
        """Returns (disp, dispr): float32 left (occlusion-masked) and right disparity.
        init_mode "sparse" without seed maps runs SparseInit for both views on the device."""
        iml = np.ascontiguousarray(iml, np.uint8)
        imr = np.ascontiguousarray(imr, np.uint8)
        if iml.ndim != 2 or iml.shape != imr.shape:
            raise ValueError("Match expects two HxW uint8 images of equal size")
        h, w = iml.shape
        disp = np.empty((h, w), np.float32)
        dispr = np.empty((h, w), np.float32)
        if seed_l is not None:
            seed_l = np.ascontiguousarray(seed_l, np.float32)
            seed_r = np.ascontiguousarray(seed_r, np.float32)
        self._check(self._lib.pm_match_host(self._h, _ptr(iml), _ptr(imr), w, h, w, _ptr(seed_l),
                                            _ptr(seed_r), pair_index, _ptr(disp), _ptr(dispr), w * 4))
        return disp, dispr

    def MatchBatch(self, left, right, seed_l=None, seed_r=None, first_pair_index=0, out=None):
        """n pairs: uint8 [n,H,W] arrays in, float32 [n,H,W] out (host buffers)."""
        left = np.ascontiguousarray(left, np.uint8)
        right = np.ascontiguousarray(right, np.uint8)
        n, h, w = left.shape
        if out is None:
            out = (np.empty((n, h, w), np.float32), np.empty((n, h, w), np.float32))
        if seed_l is not None:
            seed_l = np.ascontiguousarray(seed_l, np.float32)
            seed_r = np.ascontiguousarray(seed_r, np.float32)
        self._check(self._lib.pm_match_batch_host(self._h, n, _ptr(left), _ptr(right), w, h, w,
                                                  _ptr(seed_l), _ptr(seed_r), first_pair_index,
                                                  _ptr(out[0]), _ptr(out[1]), w * 4))
        return out

    def match_batch_host_async(self, n, p_left, p_right, w, h, stride, p_disp_l, p_disp_r, disp_stride,
                               first_pair_index=0):
        """pm_match_batch_host_async on raw (pinned) host pointers given as ints; call wait() before
        reading the outputs."""
        self._check(self._lib.pm_match_batch_host_async(
            self._h, n, C.c_void_p(p_left), C.c_void_p(p_right), w, h, stride, None, None,
            first_pair_index, C.c_void_p(p_disp_l), C.c_void_p(p_disp_r), disp_stride))

    def match_batch_host(self, n, p_left, p_right, w, h, stride, p_disp_l, p_disp_r, disp_stride,
                         first_pair_index=0):
        """pm_match_batch_host (blocking) on raw host pointers given as ints."""
        self._check(self._lib.pm_match_batch_host(
            self._h, n, C.c_void_p(p_left), C.c_void_p(p_right), w, h, stride, None, None,
            first_pair_index, C.c_void_p(p_disp_l), C.c_void_p(p_disp_r), disp_stride))

    def measure_fp32_peak(self):
        """TFLOP/s of a dependent-free FFMA kernel on this engine's device."""
        v = C.c_double()
        self._check(self._lib.pm_measure_fp32_peak(self._h, C.byref(v)))
        return float(v.value)

    def wait(self):
        self._check(self._lib.pm_wait(self._h))

    def match_batch_device(self, n, d_left, d_right, w, h, stride, d_disp_l, d_disp_r, disp_stride,
                           d_seed_l=None, d_seed_r=None, first_pair_index=0, stream=None):
        """Raw device pointers (ints), asynchronous on `stream` (int cudaStream_t or None)."""
        self._check(self._lib.pm_match_batch_device(
            self._h, n, C.c_void_p(d_left), C.c_void_p(d_right), w, h, stride,
            C.c_void_p(d_seed_l) if d_seed_l else None, C.c_void_p(d_seed_r) if d_seed_r else None,
            first_pair_index, C.c_void_p(d_disp_l), C.c_void_p(d_disp_r), disp_stride,
            C.c_void_p(stream) if stream else None))

    def match_planes_device(self, d_il, d_ir, d_gl, d_gr, w, h, plane_stride_bytes, d_disp,
                            disp_stride_bytes, stream=None):
        """Match(GpuMat iml, imr, Gl, Gr, GpuMat& disp) (patchmatch_gpu.h:104-108): one view on
        float32 device planes (raw pointers as ints), disp seed -> result in place."""
        self._check(self._lib.pm_match_planes_device(
            self._h, C.c_void_p(d_il), C.c_void_p(d_ir), C.c_void_p(d_gl), C.c_void_p(d_gr), w, h,
            plane_stride_bytes, C.c_void_p(d_disp), disp_stride_bytes,
            C.c_void_p(stream) if stream else None))

    def synchronize(self, stream=None):
        """pm_synchronize: waits for `stream` (None = the engine's) and raises the deferred status."""
        self._check(self._lib.pm_synchronize(self._h, C.c_void_p(stream) if stream else None))

    # ---- one frame in row bands (include/pm_b200.h, "row bands"); device pointers are ints
    def band_begin(self, d_left, d_right, w, stride, frame_h, rank, world, d_seed_l=None,
                   d_seed_r=None, seed_stride=0, pair_index=0, stream=None):
        self._check(self._lib.pm_band_begin(
            self._h, C.c_void_p(d_left), C.c_void_p(d_right), w, stride, frame_h, rank, world,
            C.c_void_p(d_seed_l) if d_seed_l else None, C.c_void_p(d_seed_r) if d_seed_r else None,
            seed_stride, pair_index, C.c_void_p(stream) if stream else None))

    def band_step(self):
        """None when the iterations are done, else the CBandXfer of the pending exchange."""
        x = CBandXfer()
        rc = self._lib.pm_band_step(self._h, C.byref(x))
        if rc < 0:
            self._check(rc)
        return x if rc == 1 else None

    def band_finish(self, d_disp_l, d_disp_r, disp_stride):
        self._check(self._lib.pm_band_finish(self._h, C.c_void_p(d_disp_l), C.c_void_p(d_disp_r),
                                             disp_stride))

    # ---- halo exchange over peer memory (include/pm_b200.h, pm_band_p2p_*)
    def band_p2p_export(self, width):
        """(64-byte CUDA IPC handle, device pointer) of this engine's receive region."""
        h = C.create_string_buffer(64)
        reg = C.c_void_p()
        self._check(self._lib.pm_band_p2p_export(self._h, int(width), h, C.byref(reg)))
        return h.raw, int(reg.value)

    def band_p2p_connect(self, handle_prev=None, handle_next=None, region_prev=None, region_next=None):
        hp = C.create_string_buffer(handle_prev, 64) if handle_prev else None
        hn = C.create_string_buffer(handle_next, 64) if handle_next else None
        self._check(self._lib.pm_band_p2p_connect(
            self._h, hp, hn, C.c_void_p(region_prev) if region_prev else None,
            C.c_void_p(region_next) if region_next else None))

    def band_p2p_disable(self):
        self._check(self._lib.pm_band_p2p_disable(self._h))

    # ---- bookkeeping
    def launch_count(self, reset=False):
        v = C.c_uint64()
        self._check(self._lib.pm_launch_count(self._h, C.byref(v)))
        if reset:
            self._lib.pm_launch_count_reset(self._h)
        return int(v.value)

    def set_profiling(self, on):
        self._check(self._lib.pm_set_profiling(self._h, int(on)))

    def stage_ms(self):
        """{stage: (milliseconds, spans)} accumulated since set_profiling(True)."""
        ms = (C.c_float * PM_N_STAGES)()
        n = (C.c_uint32 * PM_N_STAGES)()
        self._check(self._lib.pm_last_stage_ms(self._h, ms, n))
        return {self._lib.pm_stage_name(i).decode(): (float(ms[i]), int(n[i])) for i in range(PM_N_STAGES)}

    # ---- per-stage entry points (parity tests)
    def stage_load_pair(self, iml, imr):
        iml = np.ascontiguousarray(iml, np.uint8)
        imr = np.ascontiguousarray(imr, np.uint8)
        h, w = iml.shape
        self._shape = (h, w)
        self._check(self._lib.pm_stage_load_pair(self._h, _ptr(iml), _ptr(imr), w, h, w))

    def stage_get_planes(self, view):
        outs = [np.empty(self._shape, np.float32) for _ in range(4)]
        self._check(self._lib.pm_stage_get_planes(self._h, view, *[_ptr(o) for o in outs]))
        return outs

    def stage_noise_image(self, w, h):
        out = np.empty((h, w), np.float32)
        self._check(self._lib.pm_stage_noise_image(self._h, w, h, _ptr(out)))
        return out

    def stage_set_disp(self, view, disp):
        disp = np.ascontiguousarray(disp, np.float32)
        assert disp.shape == self._shape
        self._check(self._lib.pm_stage_set_disp(self._h, view, _ptr(disp)))

    def stage_get_disp(self, view, want_cost=False):
        d = np.empty(self._shape, np.float32)
        c = np.empty(self._shape, np.float32) if want_cost else None
        self._check(self._lib.pm_stage_get_disp(self._h, view, _ptr(d), _ptr(c)))
        return (d, c) if want_cost else d

    def stage_add_noise(self, view, scale):
        self._check(self._lib.pm_stage_add_noise(self._h, view, scale))

    def stage_propagate(self, view, along_x, direction):
        self._check(self._lib.pm_stage_propagate(self._h, view, int(along_x), int(direction)))

    def stage_mask_background(self, view):
        self._check(self._lib.pm_stage_mask_background(self._h, view))

    def stage_mask_occlusions(self, disp_l, disp_r):
        dl = np.array(disp_l, np.float32, copy=True, order="C")
        dr = np.ascontiguousarray(disp_r, np.float32)
        h, w = dl.shape
        self._check(self._lib.pm_stage_mask_occlusions(self._h, _ptr(dl), _ptr(dr), w, h))
        return dl

    def stage_downscale2(self, im):
        im = np.ascontiguousarray(im, np.uint8)
        h, w = im.shape
        out = np.empty((h // 2, w // 2), np.uint8)
        self._check(self._lib.pm_stage_downscale2(self._h, _ptr(im), w, h, _ptr(out)))
        return out

    def stage_corner_response(self, img):
        """cv::cornerMinEigenVal / cornerHarris as goodFeaturesToTrack evaluates it."""
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        out = np.empty((h, w), np.float32)
        self._check(self._lib.pm_stage_corner_response(self._h, _ptr(img), w, h, w, _ptr(out)))
        return out

    def stage_detect(self, img):
        """FeatureDetector::Detect(img, {}, kp): (n, 2) int32 (x, y) in selection order and the
        number of local-maximum candidates."""
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        cap = self.params.detector_params.max_features_per_frame
        xy = np.zeros((max(cap, 1), 2), np.int32)
        n, nc = C.c_int(), C.c_int()
        self._check(self._lib.pm_stage_detect(self._h, _ptr(img), w, h, w, cap, _ptr(xy), C.byref(n),
                                              C.byref(nc)))
        return xy[:n.value].copy(), int(nc.value)

    def stage_match_rectified(self, iml, imr, kps):
        """StereoMatcher::MatchRectified(iml, imr, kps): float64 disparities, -1 = no match."""
        iml = np.ascontiguousarray(iml, np.uint8)
        imr = np.ascontiguousarray(imr, np.uint8)
        h, w = iml.shape
        xy = np.ascontiguousarray(kps, np.int32).reshape(-1, 2)
        out = np.empty(len(xy), np.float64)
        self._check(self._lib.pm_stage_match_rectified(self._h, _ptr(iml), _ptr(imr), w, h, w, _ptr(xy),
                                                       len(xy), _ptr(out)))
        return out

    def stage_random_init(self, view, pair_index, level, rng):
        self._check(self._lib.pm_stage_random_init(self._h, view, pair_index, level, rng))

    def stage_subpixel(self, view):
        self._check(self._lib.pm_stage_subpixel(self._h, view))

    def stage_median(self, disp, k):
        disp = np.ascontiguousarray(disp, np.float32)
        h, w = disp.shape
        out = np.empty((h, w), np.float32)
        self._check(self._lib.pm_stage_median(self._h, _ptr(disp), w, h, k, _ptr(out)))
        return out


def band_plan(params, frame_h, rank, world):
    """pm_band_plan: the rows rank `rank` of `world` owns and must be given. Host-only."""
    lay = CBandLayout()
    c = params.to_c()
    lib = load_library()
    rc = lib.pm_band_plan(C.byref(c), frame_h, rank, world, C.byref(lay))
    if rc != 0:
        raise PmError(rc, lib.pm_last_error(None).decode())
    return lay


def band_exchange_rows(params, frame_h, rank, world, direction):
    """pm_band_exchange_rows: {name: (lo, hi)} frame rows swapped after a column sweep."""
    rows = (C.c_int * 8)()
    c = params.to_c()
    lib = load_library()
    rc = lib.pm_band_exchange_rows(C.byref(c), frame_h, rank, world, direction, C.byref(rows))
    if rc != 0:
        raise PmError(rc, lib.pm_last_error(None).decode())
    names = ("send_prev", "recv_prev", "send_next", "recv_next")
    return {n: (rows[2 * i], rows[2 * i + 1]) for i, n in enumerate(names)}


class Patchmatch:
    """bm::stereo::Patchmatch (src/vehicle/stereo_matching/patchmatch.hpp:29-81 in the
    reference): the CPU stage library, run on the GPU with the cost functor of the
    reference's only driver (test/stereo_matching/patchmatch_test.cpp:30-45). Works on an
    existing PatchmatchGpu engine; numpy arrays stand in for cv::Mat."""

    def __init__(self, engine=None, device=0):
        self._own = engine is None
        self.eng = engine if engine is not None else PatchmatchGpu(device=device)
        self._lib, self._h = self.eng._lib, self.eng._h

    def close(self):
        if self._own:
            self.eng.close()

    def _check(self, rc):
        self.eng._check(rc)

    def load(self, iml, imr):
        self.eng.stage_load_pair(iml, imr)
        self._shape = self.eng._shape

    def set_disp(self, disp):
        disp = np.ascontiguousarray(disp, np.float32)
        self._check(self._lib.pm_cpu_set_disp(self._h, _ptr(disp)))

    def get_disp(self):
        out = np.empty(self._shape, np.float32)
        self._check(self._lib.pm_cpu_get_disp(self._h, _ptr(out)))
        return out

    # Patchmatch::Initialize, patchmatch.cpp:52-87
    def Initialize(self, iml, imr, downsample_factor=1):
        iml = np.ascontiguousarray(iml, np.uint8)
        imr = np.ascontiguousarray(imr, np.uint8)
        h, w = iml.shape
        f = int(downsample_factor)
        out = np.empty((h // f, w // f), np.float32)
        self._check(self._lib.pm_cpu_initialize(self._h, _ptr(iml), _ptr(imr), w, h, w, f, _ptr(out),
                                                (w // f) * 4))
        return out

    # Patchmatch::AddNoise(disp, amount, disp > 0), patchmatch.cpp:143-155
    def AddNoise(self, amount):
        self._check(self._lib.pm_cpu_add_noise(self._h, amount))

    # Patchmatch::Propagate, patchmatch.cpp:248-311
    def Propagate(self, patch_height, patch_width, single_pass=-1):
        self._check(self._lib.pm_cpu_propagate(self._h, patch_height, patch_width, single_pass))

    # Patchmatch::RemoveBackground, patchmatch.cpp:314-360
    def RemoveBackground(self, patch_height, patch_width, win_by_factor=2.0):
        self._check(self._lib.pm_cpu_remove_background(self._h, patch_height, patch_width, win_by_factor))

    # Patchmatch::EstimateDisparity, patchmatch.hpp:48 (schedule of patchmatch_test.cpp:173-183)
    def EstimateDisparity(self, iml, imr, seed=None):
        iml = np.ascontiguousarray(iml, np.uint8)
        imr = np.ascontiguousarray(imr, np.uint8)
        if seed is None:  # Initialize(iml, imr, 1), patchmatch_test.cpp:142
            seed = self.Initialize(iml, imr, 1)
        seed = np.ascontiguousarray(seed, np.float32)
        h, w = iml.shape
        out = np.empty((h, w), np.float32)
        self._check(self._lib.pm_cpu_estimate_disparity(self._h, _ptr(iml), _ptr(imr), w, h, w,
                                                        _ptr(seed), _ptr(out), w * 4))
        self.eng._shape = self._shape = (h, w)
        return out

    def cost(self, xs, ys, ds, patch):
        xs = np.ascontiguousarray(xs, np.int32); ys = np.ascontiguousarray(ys, np.int32)
        ds = np.ascontiguousarray(ds, np.float32); patch = np.ascontiguousarray(patch, np.int32)
        out = np.empty(xs.size, np.float32)
        self._check(self._lib.pm_cpu_cost(self._h, xs.size, _ptr(xs), _ptr(ys), _ptr(ds), _ptr(patch),
                                          _ptr(out)))
        return out
