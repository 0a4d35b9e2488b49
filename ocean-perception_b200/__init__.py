"""B200-native PatchMatch stereo engine: Python host-side mirror of the reference's
bm::pm::PatchmatchGpu interface (src/vehicle/patchmatch_gpu/patchmatch_gpu.h:77-124 in
the reference) over the C ABI of lib/libpm_b200.so (include/pm_b200.h).

There is no CPU path: importing works anywhere, but creating an engine without the
compiled CUDA library or without a CUDA device raises.
"""
from .engine import (FeatureDetectorParams, Patchmatch, PatchmatchGpu, PmError,  # noqa: F401
                     StereoCamera, StereoMatcherParams, lib_path, load_library)
from . import dataset, synth  # noqa: F401

__all__ = ["PatchmatchGpu", "Patchmatch", "PmError", "FeatureDetectorParams", "StereoMatcherParams",
           "StereoCamera",
           "load_library", "lib_path", "synth", "dataset"]
