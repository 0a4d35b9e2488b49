"""Builds oracle/_ref/libpm_ref_kernels.so: the reference's own PatchMatch kernels for sm_100a.

TEST INFRASTRUCTURE. Run where /root/reference exists (the authoring container; __graft_entry__.build()
calls it). The GPU box only loads the prebuilt library, which travels with the gpurun snapshot
(oracle/_ref/ is git-ignored, not gpurun-ignored).

Recipe: cut lines 18-295 of /root/reference/src/vehicle/patchmatch_gpu/patchmatch_gpu.cu (the three
__device__ functions and four __global__ kernels; the rest of the file is host code on OpenCV-CUDA
GpuMats) into a TEMPORARY directory, and compile oracle/ref/ref_launchers.cu, which #includes that
extract behind oracle/ref/cv_cuda_shim.h, with nvcc's default flags (-fmad=true as in the reference's
build). Nothing of the reference is written into the repository; the .so holds machine code only.
"""
import hashlib
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(os.path.dirname(HERE), "_ref")
OUT = os.path.join(OUT_DIR, "libpm_ref_kernels.so")
REF_CU = "/root/reference/src/vehicle/patchmatch_gpu/patchmatch_gpu.cu"
FIRST, LAST = 18, 295   # 1-based, inclusive


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    return "nvcc"


def available():
    return os.path.exists(REF_CU)


def build(force=False, verbose=False):
    """Returns the path of the library, or None when the reference tree is absent and no prebuilt
    library exists."""
    deps = [os.path.join(HERE, f) for f in ("ref_launchers.cu", "cv_cuda_shim.h", "build_ref.py")]
    if not available():
        return OUT if os.path.exists(OUT) else None
    if (not force and os.path.exists(OUT)
            and os.path.getmtime(OUT) >= max(os.path.getmtime(d) for d in deps + [REF_CU])):
        return OUT
    with open(REF_CU) as f:
        lines = f.readlines()
    cut = lines[FIRST - 1:LAST]
    # sanity: the cut must start at GetSubpixel's template line and end at MaskOcclusions' brace
    if not cut[0].startswith("template <typename T>") or cut[-1].strip() != "}":
        raise RuntimeError("reference source moved: lines %d-%d are not the kernel block" % (FIRST, LAST))
    os.makedirs(OUT_DIR, exist_ok=True)
    with tempfile.TemporaryDirectory(prefix="pmref_") as tmp:
        ext = os.path.join(tmp, "ref_kernels_extract.cuh")
        with open(ext, "w") as f:
            f.writelines(cut)
        cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++14",
               "-shared", "-Xcompiler", "-fPIC", "-I", HERE, '-DPM_REF_EXTRACT="%s"' % ext,
               os.path.join(HERE, "ref_launchers.cu"), "-o", OUT]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed on the reference kernels")
    with open(os.path.join(OUT_DIR, "BUILD_INFO.txt"), "w") as f:
        f.write("source: %s lines %d-%d, sha256 of the cut %s\nflags: nvcc default (-fmad=true), sm_100a\n"
                % (REF_CU, FIRST, LAST, hashlib.sha256("".join(cut).encode()).hexdigest()))
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
