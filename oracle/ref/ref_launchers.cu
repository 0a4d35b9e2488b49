// ref_launchers.cu -- the reference's own PatchMatch kernels, compiled as they are.
//
// TEST INFRASTRUCTURE (oracle/ref/README.md): only tests/ load the library this builds.
//
// PM_REF_EXTRACT names a file holding lines 18-295 of
// /root/reference/src/vehicle/patchmatch_gpu/patchmatch_gpu.cu VERBATIM (GetSubpixel, L1GradientCost,
// L1GradientCost3x3, PropagateRow, PropagateCol, MaskBackground, MaskOcclusions). oracle/ref/build_ref.py
// cuts it out of the read-only reference tree into a temporary directory at build time; no reference
// source is stored in this repository. It is compiled with nvcc's defaults (-fmad=true, like the
// reference's CMake, src/vehicle/patchmatch_gpu/CMakeLists.txt:3-4), so the float contraction is
// whatever nvcc chooses for the reference's own expressions.
//
// Below the include: thin host launchers with the reference's launch shapes
// (patchmatch_gpu.cu:385-392, 406-408, 370-372) over HOST buffers (dense, stride = width), and
// two wrapper kernels that expose the __device__ functions per pixel.
#include <cstdint>
#include <cstdio>
#include <cmath>

#include "cv_cuda_shim.h"

namespace bm {
namespace pm {
namespace cu = cv::cuda;

#ifndef PM_REF_EXTRACT
#error "PM_REF_EXTRACT must name the extracted reference lines (oracle/ref/build_ref.py)"
#endif
#include PM_REF_EXTRACT

// ---- wrappers around the reference's __device__ functions (not reference code) ----

// cost of the disparity map at every interior pixel, sampled where the sweeps sample it:
// L1GradientCost3x3(..., tRow, col, y, fmaxf(x - d, patch_radius), alpha), patchmatch_gpu.cu:161-162
__global__ void WrapCost3x3(const cu::PtrStepSz<float> iml, const cu::PtrStepSz<float> imr,
                            const cu::PtrStepSz<float> Gl, const cu::PtrStepSz<float> Gr,
                            const cu::PtrStepSz<float> disp, cu::PtrStepSz<float> cost, float alpha)
{
  const int tCol = blockIdx.x * blockDim.x + threadIdx.x;
  const int tRow = blockIdx.y * blockDim.y + threadIdx.y;
  if (tRow < 1 || tRow > iml.rows - 2 || tCol < 1 || tCol > iml.cols - 2) return;
  const float y = __int2float_rd(tRow);
  const float x = __int2float_rd(tCol);
  cost(tRow, tCol) = L1GradientCost3x3(iml, imr, Gl, Gr, tRow, tCol, y, fmaxf(x - disp(tRow, tCol), 1), alpha);
}

// the generic ph x pw L1GradientCost (patchmatch_gpu.cu:45-69), same sampling rule with radius pw/2
__global__ void WrapCostGeneric(const cu::PtrStepSz<float> iml, const cu::PtrStepSz<float> imr,
                                const cu::PtrStepSz<float> Gl, const cu::PtrStepSz<float> Gr,
                                const cu::PtrStepSz<float> disp, cu::PtrStepSz<float> cost,
                                int ph, int pw, float alpha)
{
  const int tCol = blockIdx.x * blockDim.x + threadIdx.x;
  const int tRow = blockIdx.y * blockDim.y + threadIdx.y;
  const int ry = ph / 2, rx = pw / 2;
  if (tRow < ry || tRow > iml.rows - ry - 1 || tCol < rx || tCol > iml.cols - rx - 1) return;
  const float y = __int2float_rd(tRow);
  const float x = __int2float_rd(tCol);
  cost(tRow, tCol) = L1GradientCost(iml, imr, Gl, Gr, tRow, tCol, y, fmaxf(x - disp(tRow, tCol), rx), ph, pw, alpha);
}

__global__ void WrapGetSubpixel(const cu::PtrStepSz<float> im, const float* rows, const float* cols,
                                int n, float* out)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = GetSubpixel<float>(im, rows[i], cols[i]);
}

// AddForegroundNoise (patchmatch_gpu.cu:298-304) is four OpenCV-CUDA library calls, not reference
// source: threshold(disp > 0 -> 1), scaleAdd (= addWeighted(noise, scale, disp, 1, 0):
// a*alpha + b*beta + gamma in float), multiply, max(.., 0). Restated one call per statement. The scale
// is 32 / 2^iter (:395), a power of two, so noise*scale is exact and every contraction of
// a*alpha + b*beta gives the same float.
__global__ void LibThreshold(const cu::PtrStepSz<float> disp, cu::PtrStepSz<float> mask)
{
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x < disp.cols && y < disp.rows) mask(y, x) = disp(y, x) > 0.0f ? 1.0f : 0.0f;
}
__global__ void LibScaleAdd(const cu::PtrStepSz<float> noise, float scale, cu::PtrStepSz<float> disp)
{
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x < disp.cols && y < disp.rows) disp(y, x) = noise(y, x) * scale + disp(y, x) * 1.0f + 0.0f;
}
__global__ void LibMultiply(cu::PtrStepSz<float> disp, const cu::PtrStepSz<float> mask)
{
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x < disp.cols && y < disp.rows) disp(y, x) = disp(y, x) * mask(y, x);
}
__global__ void LibMax0(cu::PtrStepSz<float> disp)
{
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x < disp.cols && y < disp.rows) disp(y, x) = fmaxf(disp(y, x), 0.0f);
}

}  // namespace pm
}  // namespace bm

namespace {

namespace cu = cv::cuda;
using cu::device::divUp;

struct Planes {
  float* d[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  size_t pitch = 0;
  int w = 0, h = 0, n = 0;
  ~Planes() { for (int i = 0; i < n; ++i) if (d[i]) cudaFree(d[i]); }
  // cv::cuda::GpuMat allocates with cudaMallocPitch (opencv core/src/cuda/gpu_mat.cu)
  bool alloc(int n_, int w_, int h_) {
    n = n_; w = w_; h = h_;
    for (int i = 0; i < n; ++i)
      if (cudaMallocPitch(&d[i], &pitch, (size_t)w * sizeof(float), h) != cudaSuccess) return false;
    return true;
  }
  bool up(int i, const float* src) {
    return cudaMemcpy2D(d[i], pitch, src, (size_t)w * 4, (size_t)w * 4, h, cudaMemcpyHostToDevice) == cudaSuccess;
  }
  bool down(int i, float* dst) {
    return cudaMemcpy2D(dst, (size_t)w * 4, d[i], pitch, (size_t)w * 4, h, cudaMemcpyDeviceToHost) == cudaSuccess;
  }
  cu::PtrStepSz<float> m(int i) const { return cu::PtrStepSz<float>(h, w, d[i], pitch); }
};

int done() {
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) { fprintf(stderr, "pmref: %s\n", cudaGetErrorString(e)); return -1; }
  return 0;
}

// The launch shapes of PatchmatchGpu::Match (device), patchmatch_gpu.cu:385-403, with the two
// literals of the reference as parameters: stripes (16 = column_stripes = row_stripes) and lines
// (16 = the other block dimension). stripes = 1 gives one thread per line: no concurrent writers.
void launch_row(const Planes& P, int disp_i, int direction, int stripes, int lines, float alpha) {
  const dim3 row_block(stripes, lines);
  const dim3 row_grid(divUp(stripes, row_block.x), divUp(P.h, row_block.y));
  bm::pm::PropagateRow<<<row_grid, row_block>>>(P.m(0), P.m(1), P.m(2), P.m(3), P.m(disp_i), direction, 3, alpha);
}
void launch_col(const Planes& P, int disp_i, int direction, int stripes, int lines, float alpha) {
  const dim3 col_block(lines, stripes);
  const dim3 col_grid(divUp(P.w, col_block.x), divUp(stripes, col_block.y));
  bm::pm::PropagateCol<<<col_grid, col_block>>>(P.m(0), P.m(1), P.m(2), P.m(3), P.m(disp_i), direction, 3, alpha);
}

}  // namespace

extern "C" {

int pmref_device_count(void) {
  int n = 0;
  return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
}

int pmref_get_subpixel(const float* im, int w, int h, const float* rows, const float* cols, int n, float* out) {
  Planes P;
  if (!P.alloc(1, w, h) || !P.up(0, im)) return -1;
  float *dr = 0, *dc = 0, *dout = 0;
  cudaMalloc(&dr, n * 4); cudaMalloc(&dc, n * 4); cudaMalloc(&dout, n * 4);
  cudaMemcpy(dr, rows, n * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dc, cols, n * 4, cudaMemcpyHostToDevice);
  bm::pm::WrapGetSubpixel<<<(n + 127) / 128, 128>>>(P.m(0), dr, dc, n, dout);
  int rc = done();
  if (rc == 0) cudaMemcpy(out, dout, n * 4, cudaMemcpyDeviceToHost);
  cudaFree(dr); cudaFree(dc); cudaFree(dout);
  return rc;
}

// ph == pw == 0: L1GradientCost3x3 (the cost the library launches); else the generic L1GradientCost.
int pmref_cost_map(const float* Il, const float* Ir, const float* Gl, const float* Gr, int w, int h,
                   const float* disp, int ph, int pw, float alpha, float* cost) {
  Planes P;
  if (!P.alloc(6, w, h)) return -1;
  const float* src[5] = {Il, Ir, Gl, Gr, disp};
  for (int i = 0; i < 5; ++i) if (!P.up(i, src[i])) return -1;
  cudaMemset2D(P.d[5], P.pitch, 0, (size_t)w * 4, h);
  const dim3 block(16, 16), grid(divUp(w, 16), divUp(h, 16));
  if (ph == 0 && pw == 0)
    bm::pm::WrapCost3x3<<<grid, block>>>(P.m(0), P.m(1), P.m(2), P.m(3), P.m(4), P.m(5), alpha);
  else
    bm::pm::WrapCostGeneric<<<grid, block>>>(P.m(0), P.m(1), P.m(2), P.m(3), P.m(4), P.m(5), ph, pw, alpha);
  if (done()) return -1;
  return P.down(5, cost) ? 0 : -1;
}

int pmref_propagate(const float* Il, const float* Ir, const float* Gl, const float* Gr, int w, int h,
                    float* disp, int along_x, int direction, int stripes, int lines, float alpha) {
  Planes P;
  if (!P.alloc(5, w, h)) return -1;
  const float* src[5] = {Il, Ir, Gl, Gr, disp};
  for (int i = 0; i < 5; ++i) if (!P.up(i, src[i])) return -1;
  if (along_x) launch_row(P, 4, direction, stripes, lines, alpha);
  else launch_col(P, 4, direction, stripes, lines, alpha);
  if (done()) return -1;
  return P.down(4, disp) ? 0 : -1;
}

int pmref_mask_background(const float* Il, const float* Ir, const float* Gl, const float* Gr, int w,
                          int h, float* disp, float alpha, float improve) {
  Planes P;
  if (!P.alloc(5, w, h)) return -1;
  const float* src[5] = {Il, Ir, Gl, Gr, disp};
  for (int i = 0; i < 5; ++i) if (!P.up(i, src[i])) return -1;
  const dim3 block(16, 16), grid(divUp(w, block.x), divUp(h, block.y));   // :406-407
  bm::pm::MaskBackground<<<grid, block>>>(P.m(0), P.m(1), P.m(2), P.m(3), P.m(4), 3, alpha, improve);
  if (done()) return -1;
  return P.down(4, disp) ? 0 : -1;
}

int pmref_mask_occlusions(float* displ, const float* dispr, int w, int h) {
  Planes P;
  if (!P.alloc(2, w, h) || !P.up(0, displ) || !P.up(1, dispr)) return -1;
  const dim3 block(16, 16), grid(divUp(w, block.x), divUp(h, block.y));   // :370-371
  bm::pm::MaskOcclusions<<<grid, block>>>(P.m(0), P.m(1));
  if (done()) return -1;
  return P.down(0, displ) ? 0 : -1;
}

// PatchmatchGpu::Match(GpuMat iml, imr, Gl, Gr, GpuMat& disp), patchmatch_gpu.cu:379-411, statement
// by statement: the reference's kernels with the reference's shapes and a device-wide sync after
// each; AddForegroundNoise through the restated library calls above.
int pmref_match_view(const float* Il, const float* Ir, const float* Gl, const float* Gr, int w, int h,
                     const float* unit_noise, float* disp, int iters, float alpha, float improve,
                     int stripes, int lines, int do_mask) {
  Planes P;
  if (!P.alloc(7, w, h)) return -1;
  const float* src[6] = {Il, Ir, Gl, Gr, disp, unit_noise};
  for (int i = 0; i < 6; ++i) if (!P.up(i, src[i])) return -1;
  const dim3 block(16, 16), grid(divUp(w, block.x), divUp(h, block.y));
  for (int iter = 0; iter < iters; ++iter) {
    const float scale = (float)(32.0 / std::pow(2.0, (float)iter));   // :395
    bm::pm::LibThreshold<<<grid, block>>>(P.m(4), P.m(6));
    bm::pm::LibScaleAdd<<<grid, block>>>(P.m(5), scale, P.m(4));
    bm::pm::LibMultiply<<<grid, block>>>(P.m(4), P.m(6));
    bm::pm::LibMax0<<<grid, block>>>(P.m(4));
    cudaDeviceSynchronize();
    launch_row(P, 4, 1, stripes, lines, alpha);
    cudaDeviceSynchronize();
    launch_col(P, 4, 1, stripes, lines, alpha);
    cudaDeviceSynchronize();
    launch_row(P, 4, -1, stripes, lines, alpha);
    cudaDeviceSynchronize();
    launch_col(P, 4, -1, stripes, lines, alpha);
  }
  if (do_mask)
    bm::pm::MaskBackground<<<grid, block>>>(P.m(0), P.m(1), P.m(2), P.m(3), P.m(4), 3, alpha, improve);
  if (done()) return -1;
  return P.down(4, disp) ? 0 : -1;
}

// The same sequence timed with CUDA events, planes resident in device memory: `reps` runs from the
// same seed, the average milliseconds per view in *ms. This is the reference's GPU hot loop
// (patchmatch_gpu.cu:394-410: kernels + a device-wide sync after each) recompiled for sm_100a -
// the baseline tools/config_table.py puts beside the engine, never a product path.
int pmref_match_view_timed(const float* Il, const float* Ir, const float* Gl, const float* Gr, int w, int h,
                           const float* unit_noise, const float* seed, int iters, float alpha,
                           float improve, int stripes, int lines, int reps, float* ms) {
  Planes P;
  if (!P.alloc(8, w, h)) return -1;
  const float* src[6] = {Il, Ir, Gl, Gr, seed, unit_noise};
  for (int i = 0; i < 6; ++i) if (!P.up(i, src[i])) return -1;
  if (!P.up(7, seed)) return -1;
  const dim3 block(16, 16), grid(divUp(w, block.x), divUp(h, block.y));
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  float total = 0.0f;
  for (int rep = 0; rep < reps + 1; ++rep) {
    cudaMemcpy2D(P.d[4], P.pitch, P.d[7], P.pitch, (size_t)w * 4, h, cudaMemcpyDeviceToDevice);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int iter = 0; iter < iters; ++iter) {
      const float scale = (float)(32.0 / std::pow(2.0, (float)iter));
      bm::pm::LibThreshold<<<grid, block>>>(P.m(4), P.m(6));
      bm::pm::LibScaleAdd<<<grid, block>>>(P.m(5), scale, P.m(4));
      bm::pm::LibMultiply<<<grid, block>>>(P.m(4), P.m(6));
      bm::pm::LibMax0<<<grid, block>>>(P.m(4));
      cudaDeviceSynchronize();
      launch_row(P, 4, 1, stripes, lines, alpha);
      cudaDeviceSynchronize();
      launch_col(P, 4, 1, stripes, lines, alpha);
      cudaDeviceSynchronize();
      launch_row(P, 4, -1, stripes, lines, alpha);
      cudaDeviceSynchronize();
      launch_col(P, 4, -1, stripes, lines, alpha);
    }
    bm::pm::MaskBackground<<<grid, block>>>(P.m(0), P.m(1), P.m(2), P.m(3), P.m(4), 3, alpha, improve);
    cudaDeviceSynchronize();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float t = 0.0f;
    cudaEventElapsedTime(&t, a, b);
    if (rep > 0) total += t;   // the first run warms up
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  if (done()) return -1;
  *ms = total / reps;
  return 0;
}

}  // extern "C"
