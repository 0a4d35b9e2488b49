// pm_cpu_semantics.cu -- the reference's CPU stage library, stereo::Patchmatch
// (src/vehicle/stereo_matching/patchmatch.cpp), on the GPU.
//
// Patchmatch::Propagate (patchmatch.cpp:248-311) is four strict raster passes, each
// using one neighbour: left, top, right, bottom. A pass that only looks along its own
// row (or column) leaves the rows (columns) independent, so one thread per line walks
// it serially and the result is the raster scan's, bit for bit. The cost is the functor
// the reference's only driver supplies (test/stereo_matching/patchmatch_test.cpp:30-45)
// on patches fetched with cv::getRectSubPix (patchmatch.cpp:98-111): u8 patches through
// 16-bit fixed-point bilinear weights, f32 gradient patches converted to u8 at the
// functor boundary, truncated mean-L1. All of it is integer or single-rounded float
// arithmetic restated from OpenCV's published algorithm (imgproc/samplers.cpp).
//
// These kernels serve parity with the reference's CPU path (config C1); the throughput
// path is pm_sweep.cu. Citations are relative to /root/reference.
#include <cstdlib>

#include "pm_kernels.h"

namespace pm {

namespace {

constexpr int kMaxPatch = 5;

struct Adjust {
  long off;
  int rx, ry, rw, rh;
};

// adjustRect (OpenCV imgproc/samplers.cpp)
__device__ __forceinline__ Adjust adjust_rect(int sw, int sh, int ww, int wh, int ipx, int ipy) {
  Adjust r;
  long off = 0;
  if (ipx >= 0) { off += ipx; r.rx = 0; }
  else { r.rx = -ipx; if (r.rx > ww) r.rx = ww; }
  if (ipx < sw - ww) r.rw = ww;
  else { r.rw = sw - ipx - 1; if (r.rw < 0) { off += r.rw; r.rw = 0; } }
  if (ipy >= 0) { off += (long)ipy * sw; r.ry = 0; }
  else r.ry = -ipy;
  if (ipy < sh - wh) r.rh = wh;
  else { r.rh = sh - ipy - 1; if (r.rh < 0) { off += (long)r.rh * sw; r.rh = 0; } }
  r.off = off - r.rx;
  return r;
}

// Pixel (x, y) of component C of an interleaved {I, G} plane (pitch in float2 elements).
template <int C>
__device__ __forceinline__ float px(const float2* __restrict__ im, int pitch, long idx, int sw) {
  // idx is a linear index into the dense sw-wide image the algorithm is written for
  const long y = idx / sw, x = idx - y * sw;
  const float2 v = im[y * pitch + x];
  return C == 0 ? v.x : v.y;
}

// cv::getRectSubPix (getRectSubPix_Cn_), u8 -> u8 when FIX, f32 -> f32 otherwise.
// FIX: weights cvRound(w * 65536), value (sum + 32768) >> 16.
template <int C, bool FIX>
__device__ void rect_subpix(const float2* __restrict__ im, int pitch, int sw, int sh, int pw, int ph,
                            float cx, float cy, float* __restrict__ dst) {
  cx = __fsub_rn(cx, __fmul_rn(__int2float_rn(pw - 1), 0.5f));
  cy = __fsub_rn(cy, __fmul_rn(__int2float_rn(ph - 1), 0.5f));
  const int ipx = __float2int_rd(cx), ipy = __float2int_rd(cy);
  const float a = __fsub_rn(cx, __int2float_rn(ipx)), b = __fsub_rn(cy, __int2float_rn(ipy));
  const float oma = __fsub_rn(1.f, a), omb = __fsub_rn(1.f, b);
  const float fa11 = __fmul_rn(oma, omb), fa12 = __fmul_rn(a, omb);
  const float fa21 = __fmul_rn(oma, b), fa22 = __fmul_rn(a, b);
  int ia11 = 0, ia12 = 0, ia21 = 0, ia22 = 0, ib1 = 0, ib2 = 0;
  if (FIX) {
    ia11 = __float2int_rn(__fmul_rn(fa11, 65536.f)); ia12 = __float2int_rn(__fmul_rn(fa12, 65536.f));
    ia21 = __float2int_rn(__fmul_rn(fa21, 65536.f)); ia22 = __float2int_rn(__fmul_rn(fa22, 65536.f));
    ib1 = __float2int_rn(__fmul_rn(omb, 65536.f)); ib2 = __float2int_rn(__fmul_rn(b, 65536.f));
  }
  auto mix4 = [&](float p00, float p01, float p10, float p11) -> float {
    if (FIX) {
      const int s = (int)p00 * ia11 + (int)p01 * ia12 + (int)p10 * ia21 + (int)p11 * ia22;
      return (float)((s + (1 << 15)) >> 16);
    }
    return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(p00, fa11), __fmul_rn(p01, fa12)),
                               __fmul_rn(p10, fa21)), __fmul_rn(p11, fa22));
  };
  auto mix2 = [&](float p0, float p1) -> float {
    if (FIX) {
      const int s = (int)p0 * ib1 + (int)p1 * ib2;
      return (float)((s + (1 << 15)) >> 16);
    }
    return __fadd_rn(__fmul_rn(p0, omb), __fmul_rn(p1, b));
  };
  if (0 <= ipx && ipx < sw - pw && 0 <= ipy && ipy < sh - ph) {
    for (int i = 0; i < ph; ++i)
      for (int j = 0; j < pw; ++j) {
        const long s = (long)(ipy + i) * sw + ipx + j;
        dst[i * pw + j] = mix4(px<C>(im, pitch, s, sw), px<C>(im, pitch, s + 1, sw),
                               px<C>(im, pitch, s + sw, sw), px<C>(im, pitch, s + sw + 1, sw));
      }
  } else {
    const Adjust r = adjust_rect(sw, sh, pw, ph, ipx, ipy);
    long src = r.off;
    for (int i = 0; i < ph; ++i) {
      long src2 = src + sw;
      if (i < r.ry || i >= r.rh) src2 -= sw;
      float s0 = mix2(px<C>(im, pitch, src + r.rx, sw), px<C>(im, pitch, src2 + r.rx, sw));
      for (int j = 0; j < r.rx; ++j) dst[i * pw + j] = s0;
      s0 = mix2(px<C>(im, pitch, src + r.rw, sw), px<C>(im, pitch, src2 + r.rw, sw));
      for (int j = r.rw; j < pw; ++j) dst[i * pw + j] = s0;
      for (int j = r.rx; j < r.rw; ++j)
        dst[i * pw + j] = mix4(px<C>(im, pitch, src + j, sw), px<C>(im, pitch, src + j + 1, sw),
                               px<C>(im, pitch, src2 + j, sw), px<C>(im, pitch, src2 + j + 1, sw));
      if (i < r.rh) src = src2;
    }
  }
}

// One pixel (row i, column j of the patch) of cv::getRectSubPix: the arithmetic of rect_subpix for
// that pixel alone, so that the lanes of a warp can each take one pixel of a patch.
template <int C, bool FIX>
__device__ float rect_subpix_px(const float2* __restrict__ im, int pitch, int sw, int sh, int pw,
                                int ph, float cx, float cy, int i, int j) {
  cx = __fsub_rn(cx, __fmul_rn(__int2float_rn(pw - 1), 0.5f));
  cy = __fsub_rn(cy, __fmul_rn(__int2float_rn(ph - 1), 0.5f));
  const int ipx = __float2int_rd(cx), ipy = __float2int_rd(cy);
  const float a = __fsub_rn(cx, __int2float_rn(ipx)), b = __fsub_rn(cy, __int2float_rn(ipy));
  const float oma = __fsub_rn(1.f, a), omb = __fsub_rn(1.f, b);
  auto at = [&](int y, int x) -> float {
    const float2 v = im[(size_t)y * pitch + x];
    return C == 0 ? v.x : v.y;
  };
  auto at_lin = [&](long idx) -> float {   // linear index into the dense sw-wide image
    const int y = (int)idx / sw, x = (int)idx - y * sw;
    return at(y, x);
  };
  auto mix4 = [&](float p00, float p01, float p10, float p11) -> float {
    const float fa11 = __fmul_rn(oma, omb), fa12 = __fmul_rn(a, omb);
    const float fa21 = __fmul_rn(oma, b), fa22 = __fmul_rn(a, b);
    if (FIX) {
      const int ia11 = __float2int_rn(__fmul_rn(fa11, 65536.f)), ia12 = __float2int_rn(__fmul_rn(fa12, 65536.f));
      const int ia21 = __float2int_rn(__fmul_rn(fa21, 65536.f)), ia22 = __float2int_rn(__fmul_rn(fa22, 65536.f));
      const int s = (int)p00 * ia11 + (int)p01 * ia12 + (int)p10 * ia21 + (int)p11 * ia22;
      return (float)((s + (1 << 15)) >> 16);
    }
    return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(p00, fa11), __fmul_rn(p01, fa12)),
                               __fmul_rn(p10, fa21)), __fmul_rn(p11, fa22));
  };
  auto mix2 = [&](float p0, float p1) -> float {
    if (FIX) {
      const int ib1 = __float2int_rn(__fmul_rn(omb, 65536.f)), ib2 = __float2int_rn(__fmul_rn(b, 65536.f));
      const int s = (int)p0 * ib1 + (int)p1 * ib2;
      return (float)((s + (1 << 15)) >> 16);
    }
    return __fadd_rn(__fmul_rn(p0, omb), __fmul_rn(p1, b));
  };
  if (0 <= ipx && ipx < sw - pw && 0 <= ipy && ipy < sh - ph) {
    const int y = ipy + i, x = ipx + j;
    return mix4(at(y, x), at(y, x + 1), at(y + 1, x), at(y + 1, x + 1));
  }
  const Adjust r = adjust_rect(sw, sh, pw, ph, ipx, ipy);
  long src = r.off, src2 = 0;
  for (int ii = 0;; ++ii) {       // the row loop of rect_subpix up to row i
    src2 = src + sw;
    if (ii < r.ry || ii >= r.rh) src2 -= sw;
    if (ii == i) break;
    if (ii < r.rh) src = src2;
  }
  if (j >= r.rw) return mix2(at_lin(src + r.rw), at_lin(src2 + r.rw));
  if (j < r.rx) return mix2(at_lin(src + r.rx), at_lin(src2 + r.rx));
  return mix4(at_lin(src + j), at_lin(src + j + 1), at_lin(src2 + j), at_lin(src2 + j + 1));
}

// cv::saturate_cast<uchar>(float): round half to even, clamp
__device__ __forceinline__ int sat_u8(float v) {
  const int r = __float2int_rn(v);
  return r < 0 ? 0 : (r > 255 ? 255 : r);
}

// L1GradientCostFunction on GetPatchSubpix patches (patchmatch_test.cpp:30-45,
// patchmatch.cpp:171-180) for reference pixel (x, y) and disparity d.
__device__ float c_cost(const float2* __restrict__ ref, const float2* __restrict__ mat, int pitch,
                        int w, int h, int x, int y, float d, int pw, int ph) {
  float r8[kMaxPatch * kMaxPatch], c8[kMaxPatch * kMaxPatch];
  float rg[kMaxPatch * kMaxPatch], cg[kMaxPatch * kMaxPatch];
  const float xf = __int2float_rn(x), yf = __int2float_rn(y);
  const float xc = __fsub_rn(xf, d);
  rect_subpix<0, true>(ref, pitch, w, h, pw, ph, xf, yf, r8);
  rect_subpix<1, false>(ref, pitch, w, h, pw, ph, xf, yf, rg);
  rect_subpix<0, true>(mat, pitch, w, h, pw, ph, xc, yf, c8);
  rect_subpix<1, false>(mat, pitch, w, h, pw, ph, xc, yf, cg);
  const int n = pw * ph;
  int sc = 0, sg = 0;
  for (int i = 0; i < n; ++i) {
    sc += abs((int)r8[i] - (int)c8[i]);
    sg += abs(sat_u8(rg[i]) - sat_u8(cg[i]));  // f32 -> u8 at the functor's Image1b parameters
  }
  // cv::mean = sum * (1.0 / N) in double, then float
  const double inv = 1.0 / (double)n;
  const float ec = fminf((float)((double)sc * inv), 50.0f);
  const float eg = fminf((float)((double)sg * inv), 20.0f);
  const float alpha = 0.7f;
  return __fadd_rn(__fmul_rn(alpha, ec), __fmul_rn(__fsub_rn(1.0f, alpha), eg));
}

__device__ __forceinline__ bool c_border(int x, int y, int w, int h, int pw, int ph) {
  return y < ph / 2 || x < pw / 2 || y > h - ph / 2 - 1 || x > w - pw / 2 - 1;
}

// One raster pass of Patchmatch::Propagate: pass 0 left, 1 top, 2 right, 3 bottom neighbour
// (patchmatch.cpp:264-310) with PropagateNeighbors' update rule (:158-196): clamp d0 and
// write it back, admit the neighbour iff x - dl >= pw/2, first minimum wins.
__global__ void k_c_propagate_pass(const float2* __restrict__ ref, const float2* __restrict__ mat,
                                   float* __restrict__ disp, int w, int h, int pitch, int dpitch,
                                   int ph, int pw, int pass) {
  const int line = blockIdx.x * blockDim.x + threadIdx.x;
  const bool along_x = (pass == 0 || pass == 2);
  const int nlines = along_x ? h : w, len = along_x ? w : h;
  if (line >= nlines) return;
  const int fwd = pass < 2;
  // forward passes visit 1 .. len-1, backward passes len-2 .. 0 (patchmatch.cpp:264,288)
  for (int s = 0; s < len - 1; ++s) {
    const int pos = fwd ? 1 + s : len - 2 - s;
    const int x = along_x ? pos : line, y = along_x ? line : pos;
    if (fwd ? (y < 1 || x < 1) : (y > h - 2 || x > w - 2)) continue;  // loop bounds of the other axis
    if (c_border(x, y, w, h, pw, ph)) continue;
    const int nx = x + (pass == 0 ? -1 : (pass == 2 ? 1 : 0));
    const int ny = y + (pass == 1 ? -1 : (pass == 3 ? 1 : 0));
    float d0 = disp[(size_t)y * dpitch + x];
    d0 = fminf(fmaxf(d0, 0.0f), __fsub_rn(__int2float_rn(x), __int2float_rn(pw / 2)));
    const float dl = disp[(size_t)ny * dpitch + nx];
    const float cost_cur = c_cost(ref, mat, pitch, w, h, x, y, d0, pw, ph);
    float best = d0;
    if (__fsub_rn(__int2float_rn(x), dl) >= __int2float_rn(pw / 2)) {
      const float cost_n = c_cost(ref, mat, pitch, w, h, x, y, dl, pw, ph);
      if (cost_n < cost_cur) best = dl;
    }
    disp[(size_t)y * dpitch + x] = best;
  }
}

// The same pass with ONE WARP per line: lane k < pw*ph owns pixel k of the patches (its bilinear
// fetches from the reference image, the reference gradient, and the two candidates), the two
// integer sums of absolute differences are reduced with warp shuffles (integer addition: any
// order gives the same sum), every lane forms the same cost, lane 0 stores. ~100x the speed of
// the one-thread-per-line form, whose patches live in local memory.
__device__ __forceinline__ float c_cost_warp(const float2* __restrict__ mat, int pitch, int w, int h,
                                             float xc, float yf, int pw, int ph, bool on, int pi,
                                             int pj, int r8, int rg8) {
  int sc = 0, sg = 0;
  if (on) {
    const int c8 = (int)rect_subpix_px<0, true>(mat, pitch, w, h, pw, ph, xc, yf, pi, pj);
    const int cg8 = sat_u8(rect_subpix_px<1, false>(mat, pitch, w, h, pw, ph, xc, yf, pi, pj));
    sc = abs(r8 - c8);
    sg = abs(rg8 - cg8);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    sc += __shfl_xor_sync(0xffffffffu, sc, d);
    sg += __shfl_xor_sync(0xffffffffu, sg, d);
  }
  const double inv = 1.0 / (double)(pw * ph);
  const float ec = fminf((float)((double)sc * inv), 50.0f);
  const float eg = fminf((float)((double)sg * inv), 20.0f);
  const float alpha = 0.7f;
  return __fadd_rn(__fmul_rn(alpha, ec), __fmul_rn(__fsub_rn(1.0f, alpha), eg));
}

__global__ void __launch_bounds__(128)
k_c_propagate_pass_warp(const float2* __restrict__ ref, const float2* __restrict__ mat,
                        float* disp, int w, int h, int pitch, int dpitch, int ph, int pw, int pass) {
  const int lane = threadIdx.x & 31;
  const int line = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const bool along_x = (pass == 0 || pass == 2);
  const int nlines = along_x ? h : w, len = along_x ? w : h;
  if (line >= nlines) return;  // whole warps leave together
  const bool on = lane < pw * ph;
  const int pi = on ? lane / pw : 0, pj = on ? lane - pi * pw : 0;
  const int fwd = pass < 2;
  for (int s = 0; s < len - 1; ++s) {
    const int pos = fwd ? 1 + s : len - 2 - s;
    const int x = along_x ? pos : line, y = along_x ? line : pos;
    if (fwd ? (y < 1 || x < 1) : (y > h - 2 || x > w - 2)) continue;
    if (c_border(x, y, w, h, pw, ph)) continue;
    const int nx = x + (pass == 0 ? -1 : (pass == 2 ? 1 : 0));
    const int ny = y + (pass == 1 ? -1 : (pass == 3 ? 1 : 0));
    float d0 = disp[(size_t)y * dpitch + x];
    d0 = fminf(fmaxf(d0, 0.0f), __fsub_rn(__int2float_rn(x), __int2float_rn(pw / 2)));
    const float dl = disp[(size_t)ny * dpitch + nx];
    const float xf = __int2float_rn(x), yf = __int2float_rn(y);
    int r8 = 0, rg8 = 0;
    if (on) {
      r8 = (int)rect_subpix_px<0, true>(ref, pitch, w, h, pw, ph, xf, yf, pi, pj);
      rg8 = sat_u8(rect_subpix_px<1, false>(ref, pitch, w, h, pw, ph, xf, yf, pi, pj));
    }
    const float cost_cur = c_cost_warp(mat, pitch, w, h, __fsub_rn(xf, d0), yf, pw, ph, on, pi, pj, r8, rg8);
    float best = d0;
    if (__fsub_rn(xf, dl) >= __int2float_rn(pw / 2)) {   // uniform across the warp
      const float cost_n = c_cost_warp(mat, pitch, w, h, __fsub_rn(xf, dl), yf, pw, ph, on, pi, pj, r8, rg8);
      if (cost_n < cost_cur) best = dl;
    }
    if (lane == 0) disp[(size_t)y * dpitch + x] = best;
    __syncwarp();   // the next step of this line reads what lane 0 just stored
  }
}

// Patchmatch::RemoveBackground (patchmatch.cpp:314-360)
__global__ void k_c_remove_background(const float2* __restrict__ ref, const float2* __restrict__ mat,
                                      float* __restrict__ disp, int w, int h, int pitch, int dpitch,
                                      int ph, int pw, float win_by_factor) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x < 1 || x >= w || y < 1 || y >= h) return;
  if (c_border(x, y, w, h, pw, ph)) return;
  float d0 = disp[(size_t)y * dpitch + x];
  d0 = fminf(fmaxf(d0, 0.0f), __fsub_rn(__int2float_rn(x), __int2float_rn(pw / 2)));
  const float cost_cur = c_cost(ref, mat, pitch, w, h, x, y, d0, pw, ph);
  const float cost_zero = c_cost(ref, mat, pitch, w, h, x, y, 0.0f, pw, ph);
  if (cost_cur > __fdiv_rn(cost_zero, win_by_factor)) disp[(size_t)y * dpitch + x] = 0.0f;
}

// Patchmatch::AddNoise with mask = disp > 0 (patchmatch.cpp:143-155): cv::add under the
// mask, then max(disp, 0).
__global__ void k_c_add_noise(float* __restrict__ disp, const float* __restrict__ noise, int w,
                              int h, int dpitch, int npitch) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= w) return;
  float d = disp[(size_t)y * dpitch + x];
  if (d > 0.0f) d = __fadd_rn(d, noise[(size_t)y * npitch + x]);
  disp[(size_t)y * dpitch + x] = d > 0.0f ? d : 0.0f;
}

// cost of one pixel list (parity tests of the functor)
__global__ void k_c_cost_list(const float2* __restrict__ ref, const float2* __restrict__ mat, int w,
                              int h, int pitch, const int* __restrict__ xs, const int* __restrict__ ys,
                              const float* __restrict__ ds, const int* __restrict__ pws, int n,
                              float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = c_cost(ref, mat, pitch, w, h, xs[i], ys[i], ds[i], pws[i], pws[i]);
}

}  // namespace

static inline unsigned cdivu(long a, long b) { return (unsigned)((a + b - 1) / b); }

int launch_c_propagate_pass(const float2* ref, const float2* mat, float* disp, int w, int h,
                            int pitch, int dpitch, int ph, int pw, int pass, cudaStream_t st) {
  if (pw > kMaxPatch || ph > kMaxPatch || pw < 1 || ph < 1 || !(pw & 1) || !(ph & 1)) return -1;
  const int nlines = (pass == 0 || pass == 2) ? h : w;
  static const bool thread_per_line = [] { const char* e = getenv("PM_CPU_THREAD_PER_LINE"); return e && e[0] == '1'; }();
  if (thread_per_line)
    k_c_propagate_pass<<<cdivu(nlines, 32), 32, 0, st>>>(ref, mat, disp, w, h, pitch, dpitch, ph, pw, pass);
  else
    k_c_propagate_pass_warp<<<cdivu(nlines, 4), 128, 0, st>>>(ref, mat, disp, w, h, pitch, dpitch, ph, pw, pass);
  return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_c_remove_background(const float2* ref, const float2* mat, float* disp, int w, int h,
                               int pitch, int dpitch, int ph, int pw, float win_by_factor,
                               cudaStream_t st) {
  if (pw > kMaxPatch || ph > kMaxPatch || pw < 1 || ph < 1 || !(pw & 1) || !(ph & 1)) return -1;
  dim3 grid(cdivu(w, 64), h);
  k_c_remove_background<<<grid, 64, 0, st>>>(ref, mat, disp, w, h, pitch, dpitch, ph, pw, win_by_factor);
  return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_c_add_noise(float* disp, const float* noise, int w, int h, int dpitch, int npitch,
                       cudaStream_t st) {
  dim3 grid(cdivu(w, 128), h);
  k_c_add_noise<<<grid, 128, 0, st>>>(disp, noise, w, h, dpitch, npitch);
  return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_c_cost_list(const float2* ref, const float2* mat, int w, int h, int pitch, const int* xs,
                       const int* ys, const float* ds, const int* pws, int n, float* out,
                       cudaStream_t st) {
  k_c_cost_list<<<cdivu(n, 64), 64, 0, st>>>(ref, mat, w, h, pitch, xs, ys, ds, pws, n, out);
  return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace pm
