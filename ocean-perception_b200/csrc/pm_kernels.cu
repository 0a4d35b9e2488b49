// pm_kernels.cu -- streaming stages of the PatchMatch pipeline (sm_100a).
// Sweep kernels live in pm_sweep.cu. Citations are relative to /root/reference.
#include "pm_kernels.h"

namespace pm {

#define PM_LAUNCH_CHECK(n) (cudaGetLastError() == cudaSuccess ? (n) : -1)

static inline unsigned cdiv(long a, long b) { return (unsigned)((a + b - 1) / b); }

// ------------------------------------------------------------------ downscale

__global__ void k_downscale2(const uint8_t* __restrict__ src, size_t spitch, size_t splane,
                             uint8_t* __restrict__ dst, int dw, int dh, size_t dpitch,
                             size_t dplane) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= dw || y >= dh) return;
  const uint8_t* r0 = src + blockIdx.z * splane + (size_t)(2 * y) * spitch + 2 * x;
  const uint8_t* r1 = r0 + spitch;
  dst[blockIdx.z * dplane + (size_t)y * dpitch + x] =
      (uint8_t)((r0[0] + r0[1] + r1[0] + r1[1] + 2) >> 2);
}

int launch_downscale2(const uint8_t* src, int sw, int sh, size_t spitch, size_t splane,
                      uint8_t* dst, size_t dpitch, size_t dplane, int n, cudaStream_t st) {
  const int dw = sw / 2, dh = sh / 2;
  dim3 grid(cdiv(dw, 128), dh, n);
  k_downscale2<<<grid, 128, 0, st>>>(src, spitch, splane, dst, dw, dh, dpitch, dplane);
  return PM_LAUNCH_CHECK(1);
}

// ----------------------------------------------------------------- preprocess

__device__ __forceinline__ int reflect101(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * n - 2 - i;
  return i;
}

// Sobel 3x3 (scale 1, BORDER_REFLECT_101) magnitude of a u8 image at (x,y):
// integers under a correctly rounded sqrt (patchmatch_gpu.cu:307-319).
__device__ __forceinline__ float2 ig_at(const uint8_t* __restrict__ im, size_t pitch, int w, int h,
                                        int x, int y, int y_off, int full_h) {
  // the border reflects at the FRAME's edge; at a band's edge the neighbour row is simply
  // missing (clamped): such rows lie in the band's halo and their gradient is never used
  const int ym = min(max(reflect101(y + y_off - 1, full_h) - y_off, 0), h - 1);
  const int yp = min(max(reflect101(y + y_off + 1, full_h) - y_off, 0), h - 1);
  const uint8_t* r0 = im + (size_t)ym * pitch;
  const uint8_t* r1 = im + (size_t)y * pitch;
  const uint8_t* r2 = im + (size_t)yp * pitch;
  const int xm = reflect101(x - 1, w), xp = reflect101(x + 1, w);
  const int a = r0[xm], b = r0[x], c = r0[xp];
  const int d = r1[xm], e = r1[x], f = r1[xp];
  const int g = r2[xm], hh = r2[x], i = r2[xp];
  const int gx = (c + 2 * f + i) - (a + 2 * d + g);
  const int gy = (g + 2 * hh + i) - (a + 2 * b + c);
  return make_float2(__int2float_rn(e), __fsqrt_rn(__int2float_rn(gx * gx + gy * gy)));
}

__global__ void k_preprocess(const uint8_t* __restrict__ L, const uint8_t* __restrict__ R,
                             size_t ipitch, size_t iplane, float2* __restrict__ ref,
                             float2* __restrict__ mat, ViewGeom g) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const int p = blockIdx.z;
  if (x >= g.w) return;
  const float2 l = ig_at(L + p * iplane, ipitch, g.w, g.h, x, y, g.y_off, g.full_h);
  const float2 r = ig_at(R + p * iplane, ipitch, g.w, g.h, x, y, g.y_off, g.full_h);
  const size_t v0 = (size_t)(2 * p) * g.plane, v1 = v0 + g.plane;
  const size_t o = (size_t)y * g.pitch + x, of = (size_t)y * g.pitch + (g.w - 1 - x);
  ref[v0 + o] = l;
  mat[v0 + o] = r;
  ref[v1 + of] = r;  // right view: flipped and swapped planes (patchmatch_gpu.cu:357-367)
  mat[v1 + of] = l;
}

// Four pixels per thread (w, pitches and pointers multiples of 4): one 32-bit load per image row
// plus the two neighbours, 16-byte stores; same integers under the same sqrt as ig_at.
__device__ __forceinline__ void ig4_at(const uint8_t* __restrict__ im, size_t pitch, int w, int h,
                                       int x4, int y, int y_off, int full_h, float2 out[4]) {
  const int ym = min(max(reflect101(y + y_off - 1, full_h) - y_off, 0), h - 1);
  const int yp = min(max(reflect101(y + y_off + 1, full_h) - y_off, 0), h - 1);
  const int rows[3] = {ym, y, yp};
  int px[3][6];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const uint8_t* row = im + (size_t)rows[r] * pitch;
    const uchar4 c = *reinterpret_cast<const uchar4*>(row + x4);
    px[r][0] = row[x4 > 0 ? x4 - 1 : 1];          // BORDER_REFLECT_101
    px[r][1] = c.x; px[r][2] = c.y; px[r][3] = c.z; px[r][4] = c.w;
    px[r][5] = row[x4 + 4 < w ? x4 + 4 : w - 2];
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int gx = (px[0][k + 2] + 2 * px[1][k + 2] + px[2][k + 2]) - (px[0][k] + 2 * px[1][k] + px[2][k]);
    const int gy = (px[2][k] + 2 * px[2][k + 1] + px[2][k + 2]) - (px[0][k] + 2 * px[0][k + 1] + px[0][k + 2]);
    out[k] = make_float2(__int2float_rn(px[1][k + 1]), __fsqrt_rn(__int2float_rn(gx * gx + gy * gy)));
  }
}

__global__ void __launch_bounds__(128)
k_preprocess4(const uint8_t* __restrict__ L, const uint8_t* __restrict__ R, size_t ipitch,
              size_t iplane, float2* __restrict__ ref, float2* __restrict__ mat, ViewGeom g) {
  const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int y = blockIdx.y;
  const int p = blockIdx.z;
  if (x4 >= g.w) return;
  float2 l[4], r[4];
  ig4_at(L + p * iplane, ipitch, g.w, g.h, x4, y, g.y_off, g.full_h, l);
  ig4_at(R + p * iplane, ipitch, g.w, g.h, x4, y, g.y_off, g.full_h, r);
  const size_t v0 = (size_t)(2 * p) * g.plane, v1 = v0 + g.plane;
  const size_t o = (size_t)y * g.pitch + x4, of = (size_t)y * g.pitch + (g.w - 4 - x4);
  float4* r0 = reinterpret_cast<float4*>(ref + v0 + o);
  float4* m0 = reinterpret_cast<float4*>(mat + v0 + o);
  r0[0] = make_float4(l[0].x, l[0].y, l[1].x, l[1].y); r0[1] = make_float4(l[2].x, l[2].y, l[3].x, l[3].y);
  m0[0] = make_float4(r[0].x, r[0].y, r[1].x, r[1].y); m0[1] = make_float4(r[2].x, r[2].y, r[3].x, r[3].y);
  // right view: flipped and swapped planes (patchmatch_gpu.cu:357-367)
  float4* r1 = reinterpret_cast<float4*>(ref + v1 + of);
  float4* m1 = reinterpret_cast<float4*>(mat + v1 + of);
  r1[0] = make_float4(r[3].x, r[3].y, r[2].x, r[2].y); r1[1] = make_float4(r[1].x, r[1].y, r[0].x, r[0].y);
  m1[0] = make_float4(l[3].x, l[3].y, l[2].x, l[2].y); m1[1] = make_float4(l[1].x, l[1].y, l[0].x, l[0].y);
}

int launch_preprocess(const uint8_t* L, const uint8_t* R, size_t ipitch, size_t iplane,
                      float2* ref, float2* mat, ViewGeom g, int npairs, cudaStream_t st) {
  const bool vec4 = g.w % 4 == 0 && g.w >= 8 && ipitch % 4 == 0 && iplane % 4 == 0 &&
                    (reinterpret_cast<uintptr_t>(L) | reinterpret_cast<uintptr_t>(R)) % 4 == 0;
  if (vec4) {
    dim3 grid(cdiv(g.w / 4, 128), g.h, npairs);
    k_preprocess4<<<grid, 128, 0, st>>>(L, R, ipitch, iplane, ref, mat, g);
  } else {
    dim3 grid(cdiv(g.w, 128), g.h, npairs);
    k_preprocess<<<grid, 128, 0, st>>>(L, R, ipitch, iplane, ref, mat, g);
  }
  return PM_LAUNCH_CHECK(1);
}

// k_preprocess4 that also leaves the two derived layouts the shared-memory row sweeps read: the
// transposed reference plane refT ([x][pitchT], rows contiguous) and the slot-interleaved matched
// plane matI ([row group of 16][column][16 rows]) of both views. A block owns one row group x 64
// columns; the values pass through a shared tile and leave as 128-byte segments (16 rows x 8 bytes),
// so k_transpose2 and k_interleave16 do not have to read the row-major planes back.
constexpr int kFuseCols = 64;

__global__ void __launch_bounds__(256)
k_preprocess_fused(const uint8_t* __restrict__ L, const uint8_t* __restrict__ R, size_t ipitch,
                   size_t iplane, float2* __restrict__ ref, float2* __restrict__ mat,
                   float2* __restrict__ refT, int pitchT, size_t planeT, float2* __restrict__ matI,
                   int cols, size_t planeI, ViewGeom g) {
  __shared__ float2 tl[16][kFuseCols + 1], tr[16][kFuseCols + 1];
  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * kFuseCols, grp = blockIdx.y, y0 = grp * 16;
  const int p = blockIdx.z;
  const size_t v0 = (size_t)(2 * p), v1 = v0 + 1;
  {
    const int cx = tid & 15, ry = tid >> 4;
    const int x4 = x0 + 4 * cx, y = y0 + ry;
    float2 l[4], r[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) l[k] = r[k] = make_float2(0.0f, 0.0f);
    if (x4 < g.w && y < g.h) {
      ig4_at(L + p * iplane, ipitch, g.w, g.h, x4, y, g.y_off, g.full_h, l);
      ig4_at(R + p * iplane, ipitch, g.w, g.h, x4, y, g.y_off, g.full_h, r);
      const size_t o = (size_t)y * g.pitch + x4, of = (size_t)y * g.pitch + (g.w - 4 - x4);
      float4* r0 = reinterpret_cast<float4*>(ref + v0 * g.plane + o);
      float4* m0 = reinterpret_cast<float4*>(mat + v0 * g.plane + o);
      r0[0] = make_float4(l[0].x, l[0].y, l[1].x, l[1].y); r0[1] = make_float4(l[2].x, l[2].y, l[3].x, l[3].y);
      m0[0] = make_float4(r[0].x, r[0].y, r[1].x, r[1].y); m0[1] = make_float4(r[2].x, r[2].y, r[3].x, r[3].y);
      // right view: flipped and swapped planes (patchmatch_gpu.cu:357-367)
      float4* r1 = reinterpret_cast<float4*>(ref + v1 * g.plane + of);
      float4* m1 = reinterpret_cast<float4*>(mat + v1 * g.plane + of);
      r1[0] = make_float4(r[3].x, r[3].y, r[2].x, r[2].y); r1[1] = make_float4(r[1].x, r[1].y, r[0].x, r[0].y);
      m1[0] = make_float4(l[3].x, l[3].y, l[2].x, l[2].y); m1[1] = make_float4(l[1].x, l[1].y, l[0].x, l[0].y);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      tl[ry][4 * cx + k] = l[k];
      tr[ry][4 * cx + k] = r[k];
    }
  }
  __syncthreads();
  const int rr = tid & 15, cl = tid >> 4, y = y0 + rr;
  float2* refT0 = refT + v0 * planeT;
  float2* refT1 = refT + v1 * planeT;
  float2* mI0 = matI + v0 * planeI + (size_t)grp * cols * 16;
  float2* mI1 = matI + v1 * planeI + (size_t)grp * cols * 16;
#pragma unroll
  for (int k = 0; k < kFuseCols / 16; ++k) {
    const int c = cl + 16 * k, x = x0 + c;
    if (x >= g.w) continue;
    const float2 lv = tl[rr][c], rv = tr[rr][c];   // zero on the rows past the image
    const int xf = g.w - 1 - x;
    if (y < g.h) {
      refT0[(size_t)x * pitchT + y] = lv;
      refT1[(size_t)xf * pitchT + y] = rv;
    }
    mI0[(size_t)x * 16 + rr] = rv;
    mI1[(size_t)xf * 16 + rr] = lv;
  }
  // the pad columns [w, cols) of the interleaved planes are zero, like the pad of the matched plane
  if (blockIdx.x == 0) {
    const int c = g.w + cl;
    if (c < cols) mI0[(size_t)c * 16 + rr] = mI1[(size_t)c * 16 + rr] = make_float2(0.0f, 0.0f);
  }
}

bool preprocess_fused_supported(const uint8_t* L, const uint8_t* R, size_t ipitch, size_t iplane,
                                ViewGeom g, int cols) {
  return g.w % 4 == 0 && g.w >= 8 && ipitch % 4 == 0 && iplane % 4 == 0 && cols - g.w <= 16 &&
         (reinterpret_cast<uintptr_t>(L) | reinterpret_cast<uintptr_t>(R)) % 4 == 0;
}

int launch_preprocess_fused(const uint8_t* L, const uint8_t* R, size_t ipitch, size_t iplane,
                            float2* ref, float2* mat, float2* refT, int pitchT, size_t planeT,
                            float2* matI, int cols, size_t planeI, ViewGeom g, int npairs,
                            cudaStream_t st) {
  dim3 grid(cdiv(g.w, kFuseCols), cdiv(g.h, 16), npairs);
  k_preprocess_fused<<<grid, 256, 0, st>>>(L, R, ipitch, iplane, ref, mat, refT, pitchT, planeT, matI,
                                           cols, planeI, g);
  return PM_LAUNCH_CHECK(1);
}

// ---------------------------------------------------------------- noise image

// cv::RNG is the multiply-with-carry generator s' = A*lo(s) + hi(s), A = 4164903690,
// which is the Lehmer generator s' = A*s mod M, M = A*2^32 - 1. Each thread jumps to
// the start of its run with A^n mod M (square-and-multiply over a host-built table of
// A^(2^j)) and then steps sequentially.
__device__ __forceinline__ uint64_t addmod(uint64_t a, uint64_t b, uint64_t m) {
  const uint64_t s = a + b;
  return (s < a || s >= m) ? s - m : s;
}

__device__ uint64_t mulmod(uint64_t a, uint64_t b, uint64_t m) {
  uint64_t r = 0;
  while (b) {
    if (b & 1) r = addmod(r, a, m);
    a = addmod(a, a, m);
    b >>= 1;
  }
  return r;
}

struct MwcTable {
  uint64_t pow2[40];  // A^(2^j) mod M
};

constexpr int kNoiseRun = 128;

__global__ void k_noise_image(float* __restrict__ noise, int w, int h, int pitch, uint64_t seed,
                              MwcTable tab, float p0, float p1, long first) {
  // element i of this buffer is element first + i of the generator's stream
  const uint64_t A = 4164903690ull, M = (A << 32) - 1;
  const long run = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long n = (long)w * h;
  long i = run * kNoiseRun;
  if (i >= n) return;
  // state before stream element first + i = A^(first+i) * seed mod M
  uint64_t s = seed % M;
  uint64_t e = (uint64_t)(first + i);
  for (int j = 0; e; ++j, e >>= 1)
    if (e & 1) s = mulmod(s, tab.pow2[j], M);
  if (first + i == 0) s = seed;  // the first step runs on the raw seed
  const long end = min(i + (long)kNoiseRun, n);
  for (; i < end; ++i) {
    s = (uint64_t)(uint32_t)s * A + (uint32_t)(s >> 32);
    const int t = (int)(uint32_t)s;
    const int y = (int)(i / w), x = (int)(i - (long)y * w);
    noise[(size_t)y * pitch + x] = __fadd_rn(__fmul_rn(__int2float_rn(t), p0), p1);
  }
}

int launch_rng_uniform(float* out, int w, int h, int pitch, uint64_t seed, float lo, float hi,
                       long first, cudaStream_t st) {
  const unsigned __int128 M = ((unsigned __int128)4164903690ull << 32) - 1;
  MwcTable tab;
  unsigned __int128 a = 4164903690ull;
  for (int j = 0; j < 40; ++j) {
    tab.pow2[j] = (uint64_t)a;
    a = (a * a) % M;
  }
  if (seed == 0) seed = 0xffffffffull;  // cv::RNG(0) (OpenCV core operations.hpp)
  // p0 = (float)((hi-lo) * 2^-32), p1 = (float)((hi+lo)/2)  (OpenCV core rand.cpp)
  const double da = lo < hi ? lo : hi, db = lo < hi ? hi : lo;
  const float p0 = (float)((db - da) * 2.3283064365386963e-10), p1 = (float)((da + db) * 0.5);
  const long runs = ((long)w * h + kNoiseRun - 1) / kNoiseRun;
  k_noise_image<<<cdiv(runs, 64), 64, 0, st>>>(out, w, h, pitch, seed, tab, p0, p1, first);
  return PM_LAUNCH_CHECK(1);
}

int launch_noise_image(float* noise, int w, int h, int pitch, uint64_t seed, long first,
                       cudaStream_t st) {
  return launch_rng_uniform(noise, w, h, pitch, seed, -1.0f, 1.0f, first, st);
}

// ----------------------------------------------------------------------- init

__global__ void k_init_random(float2* __restrict__ dc, ViewGeom g, uint64_t seed,
                              uint32_t first_pair, uint32_t level, float range) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y, v = blockIdx.z;
  if (x >= g.w) return;
  const uint32_t idx = (uint32_t)((y + g.y_off) * g.w + x);
  const float u = philox_u01(seed, idx, first_pair + (uint32_t)(v >> 1),
                             ((uint32_t)(v & 1) << 8) | level, 0x50524d49u);
  dc[(size_t)v * g.plane + (size_t)y * g.pitch + x] = make_float2(__fmul_rn(u, range), 0.0f);
}

int launch_init_random(float2* dc, ViewGeom g, int nviews, uint64_t seed, uint32_t first_pair,
                       uint32_t level, float range, cudaStream_t st) {
  dim3 grid(cdiv(g.w, 128), g.h, nviews);
  k_init_random<<<grid, 128, 0, st>>>(dc, g, seed, first_pair, level, range);
  return PM_LAUNCH_CHECK(1);
}

__global__ void k_init_seeds(float2* __restrict__ dc, ViewGeom g, const float* __restrict__ seed_l,
                             const float* __restrict__ seed_r, size_t spitch, size_t splane,
                             int level, float level_scale) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y, v = blockIdx.z;
  if (x >= g.w) return;
  const float* s = ((v & 1) ? seed_r : seed_l) + (size_t)(v >> 1) * splane;
  const int sx = (v & 1) ? (g.w - 1 - x) : x;  // view coordinates -> image coordinates
  const float d = s[(size_t)(y << level) * spitch + ((size_t)sx << level)];
  dc[(size_t)v * g.plane + (size_t)y * g.pitch + x] = make_float2(__fmul_rn(d, level_scale), 0.0f);
}

int launch_init_seeds(float2* dc, ViewGeom g, int npairs, const float* seed_l, const float* seed_r,
                      size_t spitch, size_t splane, int level, cudaStream_t st) {
  dim3 grid(cdiv(g.w, 128), g.h, 2 * npairs);
  k_init_seeds<<<grid, 128, 0, st>>>(dc, g, seed_l, seed_r, spitch, splane, level,
                                     1.0f / (float)(1 << level));
  return PM_LAUNCH_CHECK(1);
}

// One thread per source element: its 2x2 block of the finer level (and, on the last source
// column / row, whatever the odd size leaves over) gets {2 d, 0}; aligned pairs leave as one
// 16-byte store.
// pstride: 1 = a plane of disparities, 2 = the coarser level's {d, cost} plane itself (its .x).
__global__ void k_upsample2(float2* __restrict__ dc, ViewGeom g, const float* __restrict__ prev,
                            int pw, int ph, int ppitch, size_t pplane, int pstride) {
  const int sx = blockIdx.x * blockDim.x + threadIdx.x;
  const int sy = blockIdx.y, v = blockIdx.z;
  if (sx >= pw) return;
  const float d = __fmul_rn(2.0f, prev[((size_t)v * pplane + (size_t)sy * ppitch + sx) * pstride]);
  const int x0 = 2 * sx, x1 = sx == pw - 1 ? g.w : x0 + 2;   // min(x >> 1, pw - 1) == sx
  const int y0 = 2 * sy, y1 = sy == ph - 1 ? g.h : y0 + 2;
  for (int y = y0; y < y1; ++y) {
    float2* row = dc + (size_t)v * g.plane + (size_t)y * g.pitch;
    *reinterpret_cast<float4*>(row + x0) = make_float4(d, 0.0f, d, 0.0f);  // x0 even, pitch even
    for (int x = x0 + 2; x < x1; ++x) row[x] = make_float2(d, 0.0f);
  }
}

int launch_upsample2(float2* dc, ViewGeom g, int nviews, const float* prev, int pw, int ph,
                     int ppitch, size_t pplane, cudaStream_t st, int pstride) {
  dim3 grid(cdiv(pw, 128), ph, nviews);
  k_upsample2<<<grid, 128, 0, st>>>(dc, g, prev, pw, ph, ppitch, pplane, pstride);
  return PM_LAUNCH_CHECK(1);
}

__global__ void k_extract_disp(const float2* __restrict__ dc, ViewGeom g, float* __restrict__ out,
                               int opitch, size_t oplane, int want_cost) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y, v = blockIdx.z;
  if (x >= g.w) return;
  const float2 e = dc[(size_t)v * g.plane + (size_t)y * g.pitch + x];
  out[(size_t)v * oplane + (size_t)y * opitch + x] = want_cost ? e.y : e.x;
}

int launch_extract_disp(const float2* dc, ViewGeom g, int nviews, float* out, int opitch,
                        size_t oplane, cudaStream_t st) {
  dim3 grid(cdiv(g.w, 128), g.h, nviews);
  k_extract_disp<<<grid, 128, 0, st>>>(dc, g, out, opitch, oplane, 0);
  return PM_LAUNCH_CHECK(1);
}

int launch_extract_cost(const float2* dc, ViewGeom g, int nviews, float* out, int opitch,
                        size_t oplane, cudaStream_t st) {
  dim3 grid(cdiv(g.w, 128), g.h, nviews);
  k_extract_disp<<<grid, 128, 0, st>>>(dc, g, out, opitch, oplane, 1);
  return PM_LAUNCH_CHECK(1);
}

__global__ void k_set_disp(float2* __restrict__ dc, ViewGeom g, const float* __restrict__ in,
                           int ipitch, size_t iplane) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y, v = blockIdx.z;
  if (x >= g.w) return;
  dc[(size_t)v * g.plane + (size_t)y * g.pitch + x] =
      make_float2(in[(size_t)v * iplane + (size_t)y * ipitch + x], 0.0f);
}

int launch_set_disp(float2* dc, ViewGeom g, int nviews, const float* in, int ipitch,
                    size_t iplane, cudaStream_t st) {
  dim3 grid(cdiv(g.w, 128), g.h, nviews);
  k_set_disp<<<grid, 128, 0, st>>>(dc, g, in, ipitch, iplane);
  return PM_LAUNCH_CHECK(1);
}

// Caller-owned float planes (the reference's Il/Gl or Ir/Gr GpuMats, patchmatch_gpu.h:104-108) ->
// one interleaved {I, G} plane of the workspace. Pad elements stay as they are (zero).
__global__ void k_interleave_ig(const float* __restrict__ I, const float* __restrict__ G,
                                size_t ipitch, float2* __restrict__ out, ViewGeom g) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= g.w) return;
  out[(size_t)y * g.pitch + x] = make_float2(I[(size_t)y * ipitch + x], G[(size_t)y * ipitch + x]);
}

int launch_interleave_ig(const float* I, const float* G, size_t ipitch, float2* out, ViewGeom g,
                         cudaStream_t st) {
  dim3 grid(cdiv(g.w, 128), g.h);
  k_interleave_ig<<<grid, 128, 0, st>>>(I, G, ipitch, out, g);
  return PM_LAUNCH_CHECK(1);
}

// ------------------------------------------------------------------ transpose

// float2 plane [h][pitch] -> [w][pitchT] (rows contiguous), 32x32 tiles through shared
// memory; both sides move 256-byte row segments.
__global__ void k_transpose2(const float2* __restrict__ src, int w, int h, int pitch, size_t plane,
                             float2* __restrict__ dst, int pitchT, size_t planeT) {
  __shared__ float2 tile[32][33];
  const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
  src += blockIdx.z * plane;
  dst += blockIdx.z * planeT;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const int x = x0 + tx, y = y0 + ty + j;
    if (x < w && y < h) tile[ty + j][tx] = src[(size_t)y * pitch + x];
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const int y = y0 + tx, x = x0 + ty + j;
    if (x < w && y < h) dst[(size_t)x * pitchT + y] = tile[tx][ty + j];
  }
}

int launch_transpose2(const float2* src, int w, int h, int pitch, size_t plane, float2* dst,
                      int pitchT, size_t planeT, int n, cudaStream_t st) {
  dim3 grid(cdiv(w, 32), cdiv(h, 32), n);
  k_transpose2<<<grid, dim3(32, 8), 0, st>>>(src, w, h, pitch, plane, dst, pitchT, planeT);
  return PM_LAUNCH_CHECK(1);
}

__global__ void k_transpose1(const float* __restrict__ src, int w, int h, int pitch,
                             float* __restrict__ dst, int pitchT) {
  __shared__ float tile[32][33];
  const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const int x = x0 + tx, y = y0 + ty + j;
    if (x < w && y < h) tile[ty + j][tx] = src[(size_t)y * pitch + x];
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const int y = y0 + tx, x = x0 + ty + j;
    if (x < w && y < h) dst[(size_t)x * pitchT + y] = tile[tx][ty + j];
  }
}

int launch_transpose1(const float* src, int w, int h, int pitch, float* dst, int pitchT,
                      cudaStream_t st) {
  dim3 grid(cdiv(w, 32), cdiv(h, 32));
  k_transpose1<<<grid, dim3(32, 8), 0, st>>>(src, w, h, pitch, dst, pitchT);
  return PM_LAUNCH_CHECK(1);
}

// ---------------------------------------------------------------- noise + cost

// AddForegroundNoise: mask = d > 0; d = max((noise*scale + d) * mask, 0)
// (patchmatch_gpu.cu:300-303; scaleAdd contracts to fma). The cost of the new d is
// evaluated in the same pass so the sweeps never recompute cost(d0).
template <int MODE>
__global__ void __launch_bounds__(128)
k_noise_cost(const float2* __restrict__ ref, const float2* __restrict__ mat,
             float2* __restrict__ dc, ViewGeom g, const float* __restrict__ noise, int npitch,
             float scale, float dmax, int improve, float alpha, float w1) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y, v = blockIdx.z;
  if (x >= g.w) return;
  const size_t vo = (size_t)v * g.plane;
  const size_t o = vo + (size_t)y * g.pitch + x;
  const float d = dc[o].x;
  const bool interior = row_interior(g, y) && col_interior(g, x);
  const int rad = MODE == 0 ? 1 : g.radius;
  float dn = 0.0f;
  if (d > 0.0f) {
    if (scale != 0.0f) {
      const float t = __fmaf_rn(scale, noise[(size_t)y * npitch + x], d);
      dn = fminf(t > 0.0f ? t : 0.0f, dmax);
    } else {
      dn = d;
    }
  }
  float c = 0.0f;
  if (interior) {
    c = cost_at<MODE>(g, ref + vo, mat + vo, y, x, xr_of_r(x, dn, rad), alpha, w1);
    if (improve && d > 0.0f) {
      const float c_old = cost_at<MODE>(g, ref + vo, mat + vo, y, x, xr_of_r(x, d, rad), alpha, w1);
      if (!(c < c_old)) { dn = d; c = c_old; }
    }
  } else if (improve && d > 0.0f) {
    dn = d;  // border pixels have no cost: keep d
  }
  dc[o] = make_float2(dn, c);
}

int launch_noise_cost(const float2* ref, const float2* mat, float2* dc, ViewGeom g, int nviews,
                      const float* noise, int npitch, float scale, float dmax, int improve,
                      float alpha, cudaStream_t st) {
  dim3 grid(cdiv(g.w, 128), g.h, nviews);
  if (g.cost_mode != 0 || g.radius != 1)
    k_noise_cost<1><<<grid, 128, 0, st>>>(ref, mat, dc, g, noise, npitch, scale, dmax, improve, alpha,
                                          1 - alpha);
  else
    k_noise_cost<0><<<grid, 128, 0, st>>>(ref, mat, dc, g, noise, npitch, scale, dmax, improve, alpha,
                                          1 - alpha);
  return PM_LAUNCH_CHECK(1);
}

// ------------------------------------------------------------ mask background

template <int MODE>
__global__ void __launch_bounds__(128)
k_mask_background(const float2* __restrict__ ref, const float2* __restrict__ mat,
                  const float2* __restrict__ dc, ViewGeom g, float alpha, float w1, float improve,
                  int do_mask, float* __restrict__ out, int opitch, size_t oplane) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y, v = blockIdx.z;
  if (x >= g.w) return;
  const size_t vo = (size_t)v * g.plane;
  const float2 e = dc[vo + (size_t)y * g.pitch + x];
  float d = e.x;
  if (do_mask && row_interior(g, y) && col_interior(g, x)) {
    const float cost0 = cost_at<MODE>(g, ref + vo, mat + vo, y, x, __int2float_rn(x), alpha, w1);
    if (!(e.y < __fmul_rn(improve, cost0))) d = 0.0f;  // patchmatch_gpu.cu:267-269
  }
  out[(size_t)v * oplane + (size_t)y * opitch + x] = d;
}

// MaskBackground for the 5-tap cost, as a streaming pass. The hypothesis is d = 0: the sample
// column is the pixel's own (integral) column, every lerp weight is exactly 0 and returns its first
// operand, so the five matched samples are plain elements and
//   cost(0) at (x, y) = T(x-1,y-1) + T(x+1,y-1) + T(x,y) + T(x-1,y+1) + T(x+1,y+1)   (in that order),
// with T(p) = alpha |Iref(p) - Imat(p)| + (1 - alpha) |Gref(p) - Gmat(p)| a function of the position
// alone. A warp walks kMaskRows rows of a 30-column strip: each lane computes T of one column once per
// row (one reference and one matched element), its neighbours come by shuffle and three rows of T roll
// in registers - 24 bytes read and 4 written per pixel instead of 15 taps. Bit-identical to
// k_mask_background<0> (same operations on the same operands in the same order).
constexpr int kMaskRows = 24;

__global__ void __launch_bounds__(256)
k_mask_background_d0(const float2* __restrict__ ref, const float2* __restrict__ mat,
                     const float2* __restrict__ dc, ViewGeom g, float alpha, float w1, float improve,
                     float* __restrict__ out, int opitch, size_t oplane) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int strip = blockIdx.x * 8 + warp;
  const int xo = strip * 30 + lane - 1;                 // lanes 1..30 own a column
  if (strip * 30 >= g.w) return;
  const int xl = min(max(xo, 0), g.w - 1);              // the column this lane loads
  const int ya = blockIdx.y * kMaskRows, yb = min(ya + kMaskRows, g.h);
  const size_t vo = (size_t)blockIdx.z * g.plane;
  ref += vo + xl;
  mat += vo + xl;
  dc += vo + xl;
  out += (size_t)blockIdx.z * oplane + xl;
  const bool owner = lane >= 1 && lane <= 30 && xo < g.w;
  const bool xin = owner && col_interior(g, xo);
  auto T_row = [&](int y, float& l, float& c, float& r) {
    const size_t o = (size_t)min(max(y, 0), g.h - 1) * g.pitch;
    c = tap_term(__ldg(ref + o), __ldg(mat + o), alpha, w1);
    l = __shfl_up_sync(0xffffffffu, c, 1);
    r = __shfl_down_sync(0xffffffffu, c, 1);
  };
  float ml, mc, mr, cl, cc, cr, pl, pc, pr;
  T_row(ya - 1, ml, mc, mr);
  T_row(ya, cl, cc, cr);
#pragma unroll 4
  for (int y = ya; y < yb; ++y) {
    T_row(y + 1, pl, pc, pr);
    const float2 e = __ldg(dc + (size_t)y * g.pitch);
    float cost0 = __fadd_rn(ml, mr);
    cost0 = __fadd_rn(cost0, cc);
    cost0 = __fadd_rn(cost0, pl);
    cost0 = __fadd_rn(cost0, pr);
    float d = e.x;
    if (xin && row_interior(g, y) && !(e.y < __fmul_rn(improve, cost0))) d = 0.0f;  // patchmatch_gpu.cu:267-269
    if (owner) out[(size_t)y * opitch] = d;
    ml = cl; mc = cc; mr = cr;
    cl = pl; cc = pc; cr = pr;
  }
  (void)mc;
}

int launch_mask_background(const float2* ref, const float2* mat, const float2* dc, ViewGeom g,
                           int nviews, float alpha, float improve, int do_mask, float* out,
                           int opitch, size_t oplane, cudaStream_t st) {
  static const bool d0_on = [] { const char* v = getenv("PM_MASK_D0"); return !(v && v[0] == '0'); }();
  if (d0_on && do_mask && g.cost_mode == 0 && g.radius == 1 && g.w >= 3 && g.h >= 3) {
    dim3 grid0(cdiv(cdiv(g.w, 30), 8), cdiv(g.h, kMaskRows), nviews);
    k_mask_background_d0<<<grid0, 256, 0, st>>>(ref, mat, dc, g, alpha, 1 - alpha, improve, out, opitch,
                                                oplane);
    return PM_LAUNCH_CHECK(1);
  }
  dim3 grid(cdiv(g.w, 128), g.h, nviews);
  if (g.cost_mode != 0 || g.radius != 1)
    k_mask_background<1><<<grid, 128, 0, st>>>(ref, mat, dc, g, alpha, 1 - alpha, improve, do_mask,
                                               out, opitch, oplane);
  else
    k_mask_background<0><<<grid, 128, 0, st>>>(ref, mat, dc, g, alpha, 1 - alpha, improve, do_mask,
                                               out, opitch, oplane);
  return PM_LAUNCH_CHECK(1);
}

// ------------------------------------------------------------- random search
// Extension (north-star "random-search refinement"): K perturbed
// disparities per foreground pixel, d + (2u-1) * scale / 2^(k+1) with u = Philox keyed by (pixel,
// pair, view, level, iteration, k), clamped to [0, dmax], kept only where the cost drops. Pixel-local.
template <int MODE>
__global__ void __launch_bounds__(128)
k_random_search(const float2* __restrict__ ref, const float2* __restrict__ mat,
                float2* __restrict__ dc, ViewGeom g, uint64_t seed, uint32_t first_pair,
                uint32_t level, uint32_t iter_global, int K, float scale, float dmax, float alpha,
                float w1) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y, v = blockIdx.z;
  if (x >= g.w || !row_interior(g, y) || !col_interior(g, x)) return;
  const size_t vo = (size_t)v * g.plane;
  const size_t o = vo + (size_t)y * g.pitch + x;
  float2 cur = dc[o];
  if (!(cur.x > 0.0f)) return;
  const int rad = MODE == 0 ? 1 : g.radius;
  const uint32_t idx = (uint32_t)((y + g.y_off) * g.w + x);
  const uint32_t pair = first_pair + (uint32_t)(v >> 1), view = (uint32_t)(v & 1);
  for (int k = 0; k < K; ++k) {
    const float u = philox_u01(seed, idx, pair, (view << 8) | level,
                               0x52530000u | ((iter_global & 0xffu) << 8) | (uint32_t)k);
    const float radius = __fmul_rn(scale, 1.0f / (float)(1u << (k + 1)));
    const float t = __fmaf_rn(__fsub_rn(__fmul_rn(2.0f, u), 1.0f), radius, cur.x);
    const float dn = fminf(t > 0.0f ? t : 0.0f, dmax);
    const float cn = cost_at<MODE>(g, ref + vo, mat + vo, y, x, xr_of_r(x, dn, rad), alpha, w1);
    if (cn < cur.y) cur = make_float2(dn, cn);
  }
  dc[o] = cur;
}

int launch_random_search(const float2* ref, const float2* mat, float2* dc, ViewGeom g, int nviews,
                         uint64_t seed, uint32_t first_pair, uint32_t level, uint32_t iter_global,
                         int K, float scale, float dmax, float alpha, cudaStream_t st) {
  dim3 grid(cdiv(g.w, 128), g.h, nviews);
  if (g.cost_mode != 0 || g.radius != 1)
    k_random_search<1><<<grid, 128, 0, st>>>(ref, mat, dc, g, seed, first_pair, level, iter_global, K,
                                             scale, dmax, alpha, 1 - alpha);
  else
    k_random_search<0><<<grid, 128, 0, st>>>(ref, mat, dc, g, seed, first_pair, level, iter_global, K,
                                             scale, dmax, alpha, 1 - alpha);
  return PM_LAUNCH_CHECK(1);
}

// ------------------------------------------------------------------- subpixel

template <int MODE>
__global__ void __launch_bounds__(128)
k_subpixel(const float2* __restrict__ ref, const float2* __restrict__ mat, ViewGeom g, float alpha,
           float w1, float* __restrict__ disp, int dpitch, size_t dplane) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y, v = blockIdx.z;
  if (!col_interior(g, x) || !row_interior(g, y)) return;
  float* p = disp + (size_t)v * dplane + (size_t)y * dpitch + x;
  const float d = *p;
  const float xf = __int2float_rn(x);
  const float dp1 = __fadd_rn(d, 1.0f), dm1 = __fsub_rn(d, 1.0f);
  if (!(d >= 1.0f) || !(__fsub_rn(xf, dp1) >= __int2float_rn(g.radius))) return;
  const size_t vo = (size_t)v * g.plane;
  const float c0 = cost_at<MODE>(g, ref + vo, mat + vo, y, x, __fsub_rn(xf, d), alpha, w1);
  const float cm = cost_at<MODE>(g, ref + vo, mat + vo, y, x, __fsub_rn(xf, dm1), alpha, w1);
  const float cp = cost_at<MODE>(g, ref + vo, mat + vo, y, x, __fsub_rn(xf, dp1), alpha, w1);
  const float den = __fsub_rn(__fadd_rn(cm, cp), __fmul_rn(2.0f, c0));
  if (den > 0.0f && c0 <= cm && c0 <= cp)
    *p = __fadd_rn(d, __fdiv_rn(__fmul_rn(0.5f, __fsub_rn(cm, cp)), den));
}

int launch_subpixel(const float2* ref, const float2* mat, ViewGeom g, int nviews, float alpha,
                    float* disp, int dpitch, size_t dplane, cudaStream_t st) {
  dim3 grid(cdiv(g.w, 128), g.h, nviews);
  if (g.cost_mode != 0 || g.radius != 1)
    k_subpixel<1><<<grid, 128, 0, st>>>(ref, mat, g, alpha, 1 - alpha, disp, dpitch, dplane);
  else
    k_subpixel<0><<<grid, 128, 0, st>>>(ref, mat, g, alpha, 1 - alpha, disp, dpitch, dplane);
  return PM_LAUNCH_CHECK(1);
}

// ------------------------------------------------------------------- finalize

__device__ __forceinline__ bool occluded(float dl, float dr, int lr_mode) {
  if (lr_mode == 0)  // double-precision literals, patchmatch_gpu.cu:292
    return ((double)dr > 1.4 * (double)dl) || ((double)dr < 0.7 * (double)dl);
  return fabsf(__fsub_rn(dl, dr)) > 1.0f;
}

// view 1 holds the right result in flipped coordinates: cu::flip (:368), then
// MaskOcclusions (:273-295) on the left map; both maps are written to the caller.
__global__ void k_finalize(const float* __restrict__ dispv, int vpitch, size_t vplane, int w, int h,
                           int lr_mode, float* __restrict__ out_l, float* __restrict__ out_r,
                           size_t opitch, size_t oplane) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y, p = blockIdx.z;
  if (x >= w) return;
  const float* v0 = dispv + (size_t)(2 * p) * vplane + (size_t)y * vpitch;
  const float* v1 = v0 + vplane;
  const float dl = v0[x];
  const int xr = (int)fmaxf(__fsub_rn(__int2float_rn(x), dl), 0.0f);
  const float dr = v1[w - 1 - xr];
  float* ol = (float*)((char*)out_l + p * oplane + (size_t)y * opitch);
  float* orr = (float*)((char*)out_r + p * oplane + (size_t)y * opitch);
  ol[x] = occluded(dl, dr, lr_mode) ? 0.0f : dl;
  orr[x] = v1[w - 1 - x];
}

// Four pixels per thread: 16-byte loads of both views' rows (the right one read backwards) and
// 16-byte stores; the gather of MaskOcclusions stays per pixel.
__global__ void __launch_bounds__(128)
k_finalize4(const float* __restrict__ dispv, int vpitch, size_t vplane, int w, int h, int lr_mode,
            float* __restrict__ out_l, float* __restrict__ out_r, size_t opitch, size_t oplane) {
  const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int y = blockIdx.y, p = blockIdx.z;
  if (x >= w) return;
  const float* v0 = dispv + (size_t)(2 * p) * vplane + (size_t)y * vpitch;
  const float* v1 = v0 + vplane;
  const float4 dl = *reinterpret_cast<const float4*>(v0 + x);
  const float4 rv = *reinterpret_cast<const float4*>(v1 + (w - 4 - x));   // columns w-4-x .. w-1-x
  const float d[4] = {dl.x, dl.y, dl.z, dl.w};
  float o[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int xr = (int)fmaxf(__fsub_rn(__int2float_rn(x + k), d[k]), 0.0f);
    const float dr = __ldg(v1 + (w - 1 - xr));
    o[k] = occluded(d[k], dr, lr_mode) ? 0.0f : d[k];
  }
  float* ol = (float*)((char*)out_l + p * oplane + (size_t)y * opitch);
  float* orr = (float*)((char*)out_r + p * oplane + (size_t)y * opitch);
  *reinterpret_cast<float4*>(ol + x) = make_float4(o[0], o[1], o[2], o[3]);
  *reinterpret_cast<float4*>(orr + x) = make_float4(rv.w, rv.z, rv.y, rv.x);
}

int launch_finalize(const float* dispv, int vpitch, size_t vplane, int w, int h, int npairs,
                    int lr_mode, float* out_l, float* out_r, size_t opitch_bytes,
                    size_t oplane_bytes, cudaStream_t st) {
  const bool vec4 = w % 4 == 0 && vpitch % 4 == 0 && vplane % 4 == 0 && opitch_bytes % 16 == 0 &&
                    oplane_bytes % 16 == 0 &&
                    (reinterpret_cast<uintptr_t>(dispv) | reinterpret_cast<uintptr_t>(out_l) |
                     reinterpret_cast<uintptr_t>(out_r)) % 16 == 0;
  if (vec4) {
    dim3 grid4(cdiv(w / 4, 128), h, npairs);
    k_finalize4<<<grid4, 128, 0, st>>>(dispv, vpitch, vplane, w, h, lr_mode, out_l, out_r,
                                       opitch_bytes, oplane_bytes);
    return PM_LAUNCH_CHECK(1);
  }
  dim3 grid(cdiv(w, 128), h, npairs);
  k_finalize<<<grid, 128, 0, st>>>(dispv, vpitch, vplane, w, h, lr_mode, out_l, out_r,
                                   opitch_bytes, oplane_bytes);
  return PM_LAUNCH_CHECK(1);
}

__global__ void k_mask_occlusions(float* __restrict__ disp_l, const float* __restrict__ disp_r,
                                  int w, int h, int lr_mode) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= w) return;
  const float dl = disp_l[(size_t)y * w + x];
  const int xr = (int)fmaxf(__fsub_rn(__int2float_rn(x), dl), 0.0f);
  const float dr = disp_r[(size_t)y * w + xr];
  if (occluded(dl, dr, lr_mode)) disp_l[(size_t)y * w + x] = 0.0f;
}

int launch_mask_occlusions(float* disp_l, const float* disp_r, int w, int h, int lr_mode,
                           cudaStream_t st) {
  dim3 grid(cdiv(w, 128), h, 1);
  k_mask_occlusions<<<grid, 128, 0, st>>>(disp_l, disp_r, w, h, lr_mode);
  return PM_LAUNCH_CHECK(1);
}

// --------------------------------------------------------------------- median

template <int K>
__global__ void k_median(const float* __restrict__ src, float* __restrict__ dst, int w, int h,
                         size_t pitch, size_t plane) {
  constexpr int R = K / 2, N = K * K;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= w) return;
  const char* sb = (const char*)src + blockIdx.z * plane;
  char* db = (char*)dst + blockIdx.z * plane;
  float* o = (float*)(db + (size_t)y * pitch) + x;
  if (x < R || x >= w - R || y < R || y >= h - R) {
    *o = ((const float*)(sb + (size_t)y * pitch))[x];
    return;
  }
  float v[N];
#pragma unroll
  for (int j = 0; j < K; ++j) {
    const float* r = (const float*)(sb + (size_t)(y + j - R) * pitch) + x - R;
#pragma unroll
    for (int i = 0; i < K; ++i) v[j * K + i] = r[i];
  }
  // partial selection sort up to the median: N/2+1 minima
#pragma unroll
  for (int a = 0; a <= N / 2; ++a) {
#pragma unroll
    for (int b = a + 1; b < N; ++b) {
      const float lo = fminf(v[a], v[b]), hi = fmaxf(v[a], v[b]);
      v[a] = lo;
      v[b] = hi;
    }
  }
  *o = v[N / 2];
}

int launch_median(const float* src, float* dst, int w, int h, size_t pitch_bytes,
                  size_t plane_bytes, int n, int k, cudaStream_t st) {
  dim3 grid(cdiv(w, 128), h, n);
  if (k == 3)
    k_median<3><<<grid, 128, 0, st>>>(src, dst, w, h, pitch_bytes, plane_bytes);
  else if (k == 5)
    k_median<5><<<grid, 128, 0, st>>>(src, dst, w, h, pitch_bytes, plane_bytes);
  else
    return -1;
  return PM_LAUNCH_CHECK(1);
}

// --------------------------------------------------- disparity -> depth / points
// StereoCamera::DispToDepth (vision_core/stereo_camera.cpp:49-53): fx * baseline / disp, and
// PinholeCamera::Backproject (vision_core/pinhole_camera.cpp:41-45): depth * K^-1 * (x, y, 1), as
// ObjectMesher applies them to a map computed at another resolution (mesher/object_mesher.cpp:
// 147-150: pixel and disparity divided by scale_factor). Double precision like the reference;
// disparity <= 0 (invalid/background, where the reference CHECK-fails) gives 0.
__global__ void k_disp_to_depth(const float* __restrict__ disp, int w, int h, size_t dpitch,
                                size_t dplane, double fxb, double scale, float* __restrict__ depth,
                                size_t opitch, size_t oplane, float* __restrict__ xyz, double ifx,
                                double ify, double cx, double cy) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= w) return;
  const float d = disp[blockIdx.z * dplane + (size_t)y * dpitch + x];
  double z = 0.0;
  if (d > 0.0f) z = fxb / ((double)d / scale);
  if (depth) depth[blockIdx.z * oplane + (size_t)y * opitch + x] = (float)z;
  if (xyz) {
    float* p = xyz + ((size_t)blockIdx.z * h * w + (size_t)y * w + x) * 3;
    const double u = (double)x / scale, v = (double)y / scale;
    p[0] = (float)(z * ((u - cx) * ifx));
    p[1] = (float)(z * ((v - cy) * ify));
    p[2] = (float)z;
  }
}

int launch_disp_to_depth(const float* disp, int w, int h, size_t dpitch, size_t dplane, int n,
                         double fx, double fy, double cx, double cy, double baseline, double scale,
                         float* depth, size_t opitch, size_t oplane, float* xyz, cudaStream_t st) {
  dim3 grid(cdiv(w, 128), h, n);
  k_disp_to_depth<<<grid, 128, 0, st>>>(disp, w, h, dpitch, dplane, fx * baseline, scale, depth,
                                        opitch, oplane, xyz, 1.0 / fx, 1.0 / fy, cx, cy);
  return PM_LAUNCH_CHECK(1);
}

}  // namespace pm

// ------------------------------------------------------------ FP32 peak probe
// Dependent-free FFMA loop: 16 independent accumulators per thread, 8 warps per scheduler. The
// bench uses the measured rate as the FP32 ceiling of its ALU roofline instead of a figure
// derived from the clock.
namespace pm {

__global__ void __launch_bounds__(1024)
k_fma_peak(float* out, int iters, float a, float b) {
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = (float)(threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = __fmaf_rn(acc[i], a, b);
  }
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  if (s == 123.456f) out[0] = s;   // keeps the loop alive, practically never true
}

int launch_fma_peak(float* scratch, int blocks, int iters, cudaStream_t st) {
  k_fma_peak<<<blocks, 1024, 0, st>>>(scratch, iters, 1.0000001f, 1e-7f);
  return PM_LAUNCH_CHECK(1);
}

}  // namespace pm

// ------------------------------------------------------- ForegroundTextureMask
// stereo_matching/patchmatch.cpp:19-49 (declared patchmatch.hpp:22-26; the consumer the dense depth
// was meant to feed): morphological gradient with a (2k+1)^2 rectangle (pixels outside
// the image never win), threshold at min_grad, and - for downsize 2 - OpenCV's INTER_LINEAR resize
// back to the full size in its 11-bit fixed point.
namespace pm {

// one block = a 32 x 8 tile; the (32+2k) x (8+2k) neighbourhood staged in shared memory,
// separable running max / min are not worth it for k <= 7
__global__ void __launch_bounds__(256)
k_morph_gradient_mask(const uint8_t* __restrict__ src, int w, int h, size_t pitch, int k,
                      float min_grad, uint8_t* __restrict__ dst, size_t dpitch) {
  extern __shared__ uint8_t tile[];
  const int tw = 32 + 2 * k, th = 8 + 2 * k;
  const int x0 = blockIdx.x * 32 - k, y0 = blockIdx.y * 8 - k;
  for (int i = threadIdx.x; i < tw * th; i += 256) {
    const int x = x0 + i % tw, y = y0 + i / tw;
    const bool in = x >= 0 && x < w && y >= 0 && y < h;
    // outside pixels never win: carry them as "no value" via two planes (max plane 0, min plane 255)
    tile[i] = in ? src[(size_t)y * pitch + x] : 0;
    tile[tw * th + i] = in ? src[(size_t)y * pitch + x] : 255;
  }
  __syncthreads();
  const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
  const int x = blockIdx.x * 32 + lx, y = blockIdx.y * 8 + ly;
  if (x >= w || y >= h) return;
  int mx = 0, mn = 255;
  for (int j = 0; j <= 2 * k; ++j)
    for (int i = 0; i <= 2 * k; ++i) {
      const int o = (ly + j) * tw + lx + i;
      mx = max(mx, (int)tile[o]);
      mn = min(mn, (int)tile[tw * th + o]);
    }
  dst[(size_t)y * dpitch + x] = (float)(mx - mn) > min_grad ? 255 : 0;
}

// cv::resize(u8, INTER_LINEAR) to exactly twice the size (HResizeLinear, 11-bit coefficients;
// VResizeLinear<uchar>: ((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2)
__global__ void k_resize_up2_u8(const uint8_t* __restrict__ src, int sw, int sh, size_t spitch,
                                uint8_t* __restrict__ dst, size_t dpitch) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= 2 * sw) return;
  int sy = (y + 1) / 2 - 1, b1 = (y & 1) ? 512 : 1536;
  if (sy < 0) { sy = 0; b1 = 0; }
  if (sy >= sh - 1) { sy = sh - 1; b1 = 0; }
  const int b0 = 2048 - b1, sy1 = min(sy + 1, sh - 1);
  int sx = (x + 1) / 2 - 1, a1 = (x & 1) ? 512 : 1536;
  if (sx < 0) { sx = 0; a1 = 0; }
  if (sx >= sw - 1) { sx = sw - 1; a1 = 0; }
  const int a0 = 2048 - a1, sx1 = min(sx + 1, sw - 1);
  const int S0 = src[(size_t)sy * spitch + sx] * a0 + src[(size_t)sy * spitch + sx1] * a1;
  const int S1 = src[(size_t)sy1 * spitch + sx] * a0 + src[(size_t)sy1 * spitch + sx1] * a1;
  dst[(size_t)y * dpitch + x] = (uint8_t)((((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2);
}

int launch_morph_gradient_mask(const uint8_t* src, int w, int h, size_t pitch, int k, float min_grad,
                               uint8_t* dst, size_t dpitch, cudaStream_t st) {
  dim3 grid(cdiv(w, 32), cdiv(h, 8));
  const size_t smem = 2 * (size_t)(32 + 2 * k) * (8 + 2 * k);
  k_morph_gradient_mask<<<grid, 256, smem, st>>>(src, w, h, pitch, k, min_grad, dst, dpitch);
  return PM_LAUNCH_CHECK(1);
}

int launch_resize_up2_u8(const uint8_t* src, int sw, int sh, size_t spitch, uint8_t* dst, size_t dpitch,
                         cudaStream_t st) {
  dim3 grid(cdiv(2 * sw, 128), 2 * sh);
  k_resize_up2_u8<<<grid, 128, 0, st>>>(src, sw, sh, spitch, dst, dpitch);
  return PM_LAUNCH_CHECK(1);
}

}  // namespace pm

// ------------------------------------------------ mesher-facing vertex adapter
// ObjectMesher::BuildTriangleMesh (mesher/object_mesher.cpp:139-150) turns a keypoint and its disparity
// into a mesh vertex: Backproject(pixel / scale_factor, DispToDepth(disp / scale_factor)). Here the
// disparity comes from the dense map at the keypoint's (rounded) pixel, optionally gated by a
// foreground mask (EstimateForegroundMask / ForegroundTextureMask, object_mesher.cpp:198-199).
namespace pm {

__global__ void k_mesh_vertices(const float* __restrict__ disp, int w, int h, size_t dpitch,
                                const uint8_t* __restrict__ mask, size_t mpitch,
                                const float2* __restrict__ kps, int n, double fxb, double scale,
                                double ifx, double ify, double cx, double cy,
                                float* __restrict__ out_disp, float* __restrict__ out_xyz) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float2 k = kps[i];
  const int x = __float2int_rn(k.x), y = __float2int_rn(k.y);
  float d = 0.0f;
  if (x >= 0 && x < w && y >= 0 && y < h && (!mask || mask[(size_t)y * mpitch + x] > 0))
    d = disp[(size_t)y * dpitch + x];
  double z = 0.0;
  if (d > 0.0f) z = fxb / ((double)d / scale);
  out_disp[i] = d;
  const double u = (double)k.x / scale, v = (double)k.y / scale;   // the keypoint itself, not the rounded pixel
  out_xyz[3 * i + 0] = (float)(z * ((u - cx) * ifx));
  out_xyz[3 * i + 1] = (float)(z * ((v - cy) * ify));
  out_xyz[3 * i + 2] = (float)z;
}

int launch_mesh_vertices(const float* disp, int w, int h, size_t dpitch, const uint8_t* mask,
                         size_t mpitch, const float2* kps, int n, double fx, double fy, double cx,
                         double cy, double baseline, double scale, float* out_disp, float* out_xyz,
                         cudaStream_t st) {
  k_mesh_vertices<<<cdiv(n, 128), 128, 0, st>>>(disp, w, h, dpitch, mask, mpitch, kps, n, fx * baseline,
                                               scale, 1.0 / fx, 1.0 / fy, cx, cy, out_disp, out_xyz);
  return PM_LAUNCH_CHECK(1);
}

}  // namespace pm

// ------------------------------------------- row bands: halo push over peer memory
// The rows a column sweep leaves for a neighbour band are written STRAIGHT into that neighbour's
// receive buffer (a peer-mapped allocation: NVLink stores), followed by a sequence-numbered flag;
// the receiver spins on its own flag before unpacking. No NCCL call, no staging copy on the sender.
namespace pm {

// rows [row0, row0 + nrows) of both views of a {d, cost} plane -> dst[view][row][pitch] (remote)
__global__ void __launch_bounds__(256)
k_band_push(const float2* __restrict__ dc, size_t plane, int pitch, int row0, int nrows,
            float4* __restrict__ dst) {
  const size_t per_view = (size_t)nrows * pitch / 2;          // float4 = two elements; pitch is even
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * per_view) return;
  const int v = i >= per_view;
  const size_t o = i - (size_t)v * per_view;
  const float4* src = reinterpret_cast<const float4*>(dc + (size_t)v * plane + (size_t)row0 * pitch);
  dst[i] = src[o];
}

// after the pushes of this exchange (stream order): publish its sequence number at the peer(s)
__global__ void k_band_signal(unsigned long long* flag_a, unsigned long long* flag_b,
                              unsigned long long seq) {
  if (threadIdx.x == 0) {
    __threadfence_system();
    if (flag_a) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag_a), "l"(seq) : "memory");
    if (flag_b) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag_b), "l"(seq) : "memory");
  }
}

// wait until both neighbours have published `seq` (bounded: ~timeout_ns, then *err = 1)
__global__ void k_band_wait(const unsigned long long* flag_a, const unsigned long long* flag_b,
                            unsigned long long seq, unsigned long long timeout_ns, int* err) {
  if (threadIdx.x != 0) return;
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    unsigned long long a = seq, b = seq;
    if (flag_a) asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(a) : "l"(flag_a) : "memory");
    if (flag_b) asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(b) : "l"(flag_b) : "memory");
    if (a >= seq && b >= seq) return;
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (t - t0 > timeout_ns) { *err = 1; return; }
    __nanosleep(200);
  }
}

// The same exchange in TWO launches instead of six. (1) Both pushes and the signal: every thread
// stores its 16 bytes into the neighbour's buffer and fences; the block that draws the last ticket
// publishes the sequence number at both neighbours.
__global__ void __launch_bounds__(256)
k_band_push_signal(const float2* __restrict__ dc, size_t plane, int pitch, int row_p, int n_p,
                   float4* __restrict__ dst_p, int row_n, int n_n, float4* __restrict__ dst_n,
                   unsigned long long* flag_p, unsigned long long* flag_n, unsigned long long seq,
                   unsigned* ticket) {
  const size_t pv_p = (size_t)n_p * pitch / 2, pv_n = (size_t)n_n * pitch / 2;   // float4 per view
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 2 * pv_p) {
    const int v = i >= pv_p;
    const float4* src = reinterpret_cast<const float4*>(dc + (size_t)v * plane + (size_t)row_p * pitch);
    dst_p[i] = src[i - (size_t)v * pv_p];
  } else if ((i -= 2 * pv_p) < 2 * pv_n) {
    const int v = i >= pv_n;
    const float4* src = reinterpret_cast<const float4*>(dc + (size_t)v * plane + (size_t)row_n * pitch);
    dst_n[i] = src[i - (size_t)v * pv_n];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicAdd(ticket, 1u) == gridDim.x - 1) {   // every block's stores are fenced by now
      *ticket = 0;
      __threadfence_system();
      if (flag_p) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag_p), "l"(seq) : "memory");
      if (flag_n) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag_n), "l"(seq) : "memory");
    }
  }
}

// (2) The wait and both unpacks: thread 0 of every block spins on the local flags (bounded), then
// the block copies its share of the two receive buffers into the plane (reads past L1: the rows
// were written by the neighbours' GPUs).
__global__ void __launch_bounds__(256)
k_band_wait_unpack(const unsigned long long* flag_p, const unsigned long long* flag_n,
                   unsigned long long seq, unsigned long long timeout_ns, int* err, float2* dc,
                   size_t plane, int pitch, int row_p, int n_p, const float4* __restrict__ src_p,
                   int row_n, int n_n, const float4* __restrict__ src_n) {
  if (threadIdx.x == 0) {
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
      unsigned long long a = seq, b = seq;
      if (flag_p) asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(a) : "l"(flag_p) : "memory");
      if (flag_n) asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(b) : "l"(flag_n) : "memory");
      if (a >= seq && b >= seq) break;
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (t - t0 > timeout_ns) { *err = 1; break; }
      __nanosleep(200);
    }
  }
  __syncthreads();
  // few blocks, grid-stride: the spinning blocks must never crowd out the kernels they wait for
  // (all bands share one GPU in the one-device emulation)
  const size_t pv_p = (size_t)n_p * pitch / 2, pv_n = (size_t)n_n * pitch / 2;
  for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < 2 * (pv_p + pv_n);
       j += (size_t)gridDim.x * blockDim.x) {
    if (j < 2 * pv_p) {
      const int v = j >= pv_p;
      float4* dst = reinterpret_cast<float4*>(dc + (size_t)v * plane + (size_t)row_p * pitch);
      dst[j - (size_t)v * pv_p] = __ldcg(src_p + j);
    } else {
      const size_t i = j - 2 * pv_p;
      const int v = i >= pv_n;
      float4* dst = reinterpret_cast<float4*>(dc + (size_t)v * plane + (size_t)row_n * pitch);
      dst[i - (size_t)v * pv_n] = __ldcg(src_n + i);
    }
  }
}

int launch_band_push_signal(const float2* dc, size_t plane, int pitch, int row_p, int n_p, void* dst_p,
                            int row_n, int n_n, void* dst_n, unsigned long long* flag_p,
                            unsigned long long* flag_n, unsigned long long seq, unsigned* ticket,
                            cudaStream_t st) {
  const size_t n = (size_t)(max(n_p, 0) + max(n_n, 0)) * pitch;   // float4 count over both views
  k_band_push_signal<<<max(cdiv((long)n, 256), 1), 256, 0, st>>>(
      dc, plane, pitch, row_p, max(n_p, 0), (float4*)dst_p, row_n, max(n_n, 0), (float4*)dst_n,
      flag_p, flag_n, seq, ticket);
  return PM_LAUNCH_CHECK(1);
}

int launch_band_wait_unpack(const unsigned long long* flag_p, const unsigned long long* flag_n,
                            unsigned long long seq, unsigned long long timeout_ns, int* err, float2* dc,
                            size_t plane, int pitch, int row_p, int n_p, const void* src_p, int row_n,
                            int n_n, const void* src_n, cudaStream_t st) {
  const size_t n = (size_t)(max(n_p, 0) + max(n_n, 0)) * pitch;
  k_band_wait_unpack<<<min(max(cdiv((long)n, 256), 1), 32), 256, 0, st>>>(
      flag_p, flag_n, seq, timeout_ns, err, dc, plane, pitch, row_p, max(n_p, 0), (const float4*)src_p,
      row_n, max(n_n, 0), (const float4*)src_n);
  return PM_LAUNCH_CHECK(1);
}

int launch_band_push(const float2* dc, size_t plane, int pitch, int row0, int nrows, void* dst,
                     cudaStream_t st) {
  if (nrows <= 0) return 0;
  const size_t n = (size_t)nrows * pitch;   // float4 count over both views
  k_band_push<<<cdiv((long)n, 256), 256, 0, st>>>(dc, plane, pitch, row0, nrows, (float4*)dst);
  return PM_LAUNCH_CHECK(1);
}

int launch_band_signal(unsigned long long* flag_a, unsigned long long* flag_b, unsigned long long seq,
                       cudaStream_t st) {
  k_band_signal<<<1, 32, 0, st>>>(flag_a, flag_b, seq);
  return PM_LAUNCH_CHECK(1);
}

int launch_band_wait(const unsigned long long* flag_a, const unsigned long long* flag_b,
                     unsigned long long seq, unsigned long long timeout_ns, int* err, cudaStream_t st) {
  k_band_wait<<<1, 32, 0, st>>>(flag_a, flag_b, seq, timeout_ns, err);
  return PM_LAUNCH_CHECK(1);
}

}  // namespace pm
