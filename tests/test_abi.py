"""The C-ABI shared library on a box without a GPU: it loads, exports every symbol
include/pm_b200.h declares, parses params, and refuses to create an engine (no CPU path)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "pm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pm_[a-z0-9_]+)\s*\(", text)))


def test_exports_every_declared_symbol(built_lib):
    lib = C.CDLL(built_lib)
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "libpm_b200.so does not export " + n


def test_defaults_are_the_reference_defaults(pkg, built_lib):
    lib = pkg.load_library()
    from importlib import import_module
    CParams = import_module("ocean-perception_b200.engine").CParams
    p = CParams()
    assert lib.pm_params_default(C.byref(p)) == 0
    # patchmatch_gpu.h:85-88
    assert abs(p.cost_alpha - 0.9) < 1e-7 and p.patchmatch_iters == 3
    assert p.init_dilate_factor == 4 and abs(p.cost_improve_factor - 0.8) < 1e-7
    # stereo_matcher.hpp:20-25, feature_detector.hpp:31-40
    assert (p.sm_templ_cols, p.sm_templ_rows, p.sm_max_disp) == (31, 11, 128)
    assert p.sm_max_matching_cost == 0.15 and p.fd_max_features_per_frame == 200
    # literals of the launch sites, patchmatch_gpu.cu:385-408, 341
    assert (p.patch_size, p.sweep_chunks, p.sweep_overlap, p.seed) == (3, 16, 5, 123)
    assert p.noise_scale0 == 32.0 and p.pyramid_levels == 1 and p.init_mode == 0


YAML_OK = """%YAML:1.0
# shape of config/auv/lcm_nodes/ObjectMesherLcm.yaml:37-59
PatchmatchGpu:
  cost_alpha: 0.85
  patchmatch_iters: 4
  pyramid_levels: 2
  init_mode: random     # extension key, by name
  lr_mode: 1
  seed: 77
  FeatureDetector:
    max_features_per_frame: 150
    subpixel_corners: 0 # bool
    min_distance_btw_tracked_and_detected_features: 20
    gftt_quality_level: 0.01
    gftt_block_size: 9
    gftt_use_harris_corner_detector: 0 # bool
    gftt_k: 0.04

  StereoMatcher:
    templ_cols: 31
    templ_rows: 31
    max_disp: 96
    # max_matching_cost: 0.15
    max_matching_cost: 0.10
    bidirectional: 1 # bool
    subpixel_refinement: 0 # bool
"""


def test_yaml_params(pkg, built_lib, tmp_path):
    f = tmp_path / "pm.yaml"
    f.write_text(YAML_OK)
    P = pkg.PatchmatchGpu.Params(str(f), "PatchmatchGpu")
    assert abs(P.cost_alpha - 0.85) < 1e-6 and P.patchmatch_iters == 4
    assert P.pyramid_levels == 2 and P.init_mode == "random" and P.lr_mode == "abs1px" and P.seed == 77
    assert P.detector_params.max_features_per_frame == 150 and P.detector_params.gftt_block_size == 9
    assert P.matcher_params.templ_rows == 31 and P.matcher_params.max_disp == 96
    assert P.matcher_params.max_matching_cost == 0.10 and P.matcher_params.bidirectional is True
    assert P.max_disp == 96  # defaults to StereoMatcher/max_disp
    assert P.cost_improve_factor == pytest.approx(0.8)  # untouched default


def test_yaml_missing_required_key_is_an_error_not_an_abort(pkg, built_lib, tmp_path):
    # the reference CHECK-aborts (yaml_parser.cpp:82); the C ABI returns PM_ERR_YAML
    f = tmp_path / "bad.yaml"
    f.write_text(YAML_OK.replace("    templ_rows: 31\n", ""))
    with pytest.raises(pkg.PmError) as ei:
        pkg.PatchmatchGpu.Params(str(f), "PatchmatchGpu")
    assert ei.value.code == -5 and "templ_rows" in str(ei.value)
    with pytest.raises(pkg.PmError) as ei:
        pkg.PatchmatchGpu.Params(str(tmp_path / "nope.yaml"))
    assert ei.value.code == -5


def test_invalid_params_rejected(pkg, built_lib):
    P = pkg.PatchmatchGpu.Params()
    P.patch_size = 7                      # 3 (reference) and 5 are supported
    with pytest.raises(pkg.PmError) as ei:
        pkg.PatchmatchGpu(P)
    assert ei.value.code == -2 and "patch_size" in str(ei.value)
    P = pkg.PatchmatchGpu.Params()
    P.random_search_k = 99
    with pytest.raises(pkg.PmError) as ei:
        pkg.PatchmatchGpu(P)
    assert ei.value.code == -1
    P = pkg.PatchmatchGpu.Params()
    P.sweep_overlap = 40
    with pytest.raises(pkg.PmError) as ei:
        pkg.PatchmatchGpu(P)
    assert ei.value.code == -1


def test_no_cpu_fallback(pkg, built_lib):
    """Without a CUDA device pm_create fails loudly; with one it succeeds."""
    import torch
    if torch.cuda.is_available():
        e = pkg.PatchmatchGpu()
        e.close()
    else:
        with pytest.raises(pkg.PmError) as ei:
            pkg.PatchmatchGpu()
        assert ei.value.code == -3 and "no CPU path" in str(ei.value)


def test_product_never_touches_the_oracle():
    """Nothing under the package may import, include or link oracle/."""
    pkg_dir = os.path.join(ROOT, "ocean-perception_b200")
    for base, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(base, f), errors="ignore").read()
                assert "pm_oracle" not in text and "import pmo" not in text and "pmo_" not in text, f


def test_reference_kernel_library_builds_and_exports():
    """oracle/_ref/libpm_ref_kernels.so: the reference's kernels compiled verbatim (oracle/ref/).
    Built here when /root/reference is present; the GPU box loads the prebuilt file."""
    import pmref
    path = pmref.build()
    if path is None:
        pytest.skip("no /root/reference and no prebuilt oracle/_ref library")
    lib = pmref.lib()
    for name in pmref.SYMBOLS:
        assert hasattr(lib, name), name
    # nothing of the reference's source is stored in the repository
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    assert not [f for f in os.listdir(ref_dir) if f.endswith((".cu", ".cuh", ".h", ".cpp"))]
