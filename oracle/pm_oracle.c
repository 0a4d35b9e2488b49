/*
 * pm_oracle.c -- CPU restatement of the reference PatchMatch stereo path.
 * TEST INFRASTRUCTURE ONLY (see pm_oracle.h).  Plain C99, single threaded.
 * Build: gcc -O3 -march=native -ffp-contract=off -fPIC -shared (oracle/Makefile).
 * -ffp-contract=off is REQUIRED: every fused multiply-add below is an explicit
 * fmaf() that pins the contraction nvcc applies to the reference's expression
 * shapes; anything else is individually rounded.
 *
 * All citations are relative to /root/reference.
 */
#include "pm_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define PMO_MIN(a, b) ((a) < (b) ? (a) : (b))
#define PMO_MAX(a, b) ((a) > (b) ? (a) : (b))

void pmo_params_default(pmo_params* p) {
  memset(p, 0, sizeof(*p));
  p->cost_alpha = 0.9f;          /* patchmatch_gpu.h:85 */
  p->patchmatch_iters = 3;       /* patchmatch_gpu.h:86 */
  p->cost_improve_factor = 0.8f; /* patchmatch_gpu.h:88 */
  p->sweep_chunks = 16;          /* patchmatch_gpu.cu:385-386 */
  p->sweep_overlap = 5;          /* patchmatch_gpu.cu:143-144 */
  p->noise_scale0 = 32.0f;       /* patchmatch_gpu.cu:395 */
  p->noise_accept = 0;
  p->seed = 123;                 /* patchmatch_gpu.cu:341 */
  p->init_mode = 0;
  p->max_disp = 128;             /* stereo_matcher.hpp:23 */
  p->clamp_disp = 0;
  p->pyramid_levels = 1;
  p->lr_mode = 0;
  p->subpixel = 0;
  p->median_ksize = 0;
  p->cost_mode = 0;
  p->patch_size = 3;             /* patchmatch_gpu.cu:397-408 */
  p->random_search_k = 0;
}

/* ===================================================== OpenCV primitives */

/* cv::RNG: state = (uint32)state * A + (state >> 32); value = (float)(int)state * p0 + p1
 * with p0 = (float)((hi-lo) * 2^-32), p1 = (float)((hi+lo)/2) (OpenCV core rand.cpp,
 * randf_32f).  KAT: seed 123, U(-1,1) -> 0.550433, 0.532670, 0.872047, -0.466683 ... */
void pmo_rng_uniform_f32(uint64_t seed, float lo, float hi, float* out, size_t n) {
  uint64_t state = seed ? seed : 0xffffffffULL;
  const double a = fmin((double)lo, (double)hi), b = fmax((double)lo, (double)hi);
  const float p0 = (float)(fmin(DBL_MAX, b - a) * 2.3283064365386963e-10);
  const float p1 = (float)((a + b) * 0.5);
  for (size_t i = 0; i < n; ++i) {
    state = (uint64_t)(uint32_t)state * 4164903690U + (uint32_t)(state >> 32);
    const int t = (int)(uint32_t)state;
    out[i] = (float)t * p0 + p1;
  }
}

void pmo_resize_half_u8(const uint8_t* src, int w, int h, uint8_t* dst) {
  const int dw = w / 2, dh = h / 2;
  for (int y = 0; y < dh; ++y) {
    const uint8_t* r0 = src + (size_t)(2 * y) * w;
    const uint8_t* r1 = r0 + w;
    for (int x = 0; x < dw; ++x) {
      dst[(size_t)y * dw + x] =
          (uint8_t)((r0[2 * x] + r0[2 * x + 1] + r1[2 * x] + r1[2 * x + 1] + 2) >> 2);
    }
  }
}

static inline int reflect101(int i, int n) {
  if (n == 1) return 0;
  if (i < 0) i = -i;
  if (i >= n) i = 2 * n - 2 - i;
  return i;
}

void pmo_gradient_mag_u8(const uint8_t* im, int w, int h, float* g) {
  for (int y = 0; y < h; ++y) {
    const uint8_t* r0 = im + (size_t)reflect101(y - 1, h) * w;
    const uint8_t* r1 = im + (size_t)y * w;
    const uint8_t* r2 = im + (size_t)reflect101(y + 1, h) * w;
    for (int x = 0; x < w; ++x) {
      const int xm = reflect101(x - 1, w), xp = reflect101(x + 1, w);
      const int gx = (r0[xp] + 2 * r1[xp] + r2[xp]) - (r0[xm] + 2 * r1[xm] + r2[xm]);
      const int gy = (r2[xm] + 2 * r2[x] + r2[xp]) - (r0[xm] + 2 * r0[x] + r0[xp]);
      g[(size_t)y * w + x] = sqrtf((float)(gx * gx + gy * gy));
    }
  }
}

void pmo_flip_h_u8(const uint8_t* src, int w, int h, uint8_t* dst) {
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) dst[(size_t)y * w + x] = src[(size_t)y * w + (w - 1 - x)];
}

void pmo_flip_h_f32(const float* src, int w, int h, float* dst) {
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) dst[(size_t)y * w + x] = src[(size_t)y * w + (w - 1 - x)];
}

void pmo_u8_to_f32(const uint8_t* src, size_t n, float* dst) {
  for (size_t i = 0; i < n; ++i) dst[i] = (float)src[i];
}

/* ---- cv::getRectSubPix (OpenCV imgproc samplers.cpp, getRectSubPix_Cn_ + adjustRect).
 * The window's top-left is centre - (size-1)/2; bilinear weights from the
 * fractional part; a window touching the last row/column or leaving the image
 * takes the replicate-border branch.  The same body serves u8 (16-bit fixed
 * point weights, (s + 2^15) >> 16) and f32 (float weights). */

typedef struct {
  ptrdiff_t off; /* element offset of the adjusted source origin */
  int rx, ry, rw, rh;
} pmo_adjust;

static pmo_adjust adjust_rect(int sw, int sh, int ww, int wh, int ipx, int ipy) {
  pmo_adjust r;
  ptrdiff_t off = 0;
  if (ipx >= 0) {
    off += ipx;
    r.rx = 0;
  } else {
    r.rx = -ipx;
    if (r.rx > ww) r.rx = ww;
  }
  if (ipx < sw - ww) {
    r.rw = ww;
  } else {
    r.rw = sw - ipx - 1;
    if (r.rw < 0) {
      off += r.rw;
      r.rw = 0;
    }
  }
  if (ipy >= 0) {
    off += (ptrdiff_t)ipy * sw;
    r.ry = 0;
  } else {
    r.ry = -ipy;
  }
  if (ipy < sh - wh) {
    r.rh = wh;
  } else {
    r.rh = sh - ipy - 1;
    if (r.rh < 0) {
      off += (ptrdiff_t)r.rh * sw;
      r.rh = 0;
    }
  }
  r.off = off - r.rx;
  return r;
}

#define PMO_DEFINE_SUBPIX(NAME, T, WT, SCALE, CAST)                                          \
  void NAME(const T* src0, int sw, int sh, int pw, int ph, float cx, float cy, T* dst) {     \
    cx -= (float)(pw - 1) * 0.5f;                                                            \
    cy -= (float)(ph - 1) * 0.5f;                                                            \
    const int ipx = (int)floorf(cx), ipy = (int)floorf(cy);                                  \
    const float a = cx - (float)ipx, b = cy - (float)ipy;                                    \
    const WT a11 = SCALE((1.f - a) * (1.f - b)), a12 = SCALE(a * (1.f - b));                 \
    const WT a21 = SCALE((1.f - a) * b), a22 = SCALE(a * b);                                 \
    const WT b1 = SCALE(1.f - b), b2 = SCALE(b);                                             \
    if (0 <= ipx && ipx < sw - pw && 0 <= ipy && ipy < sh - ph) {                            \
      const T* src = src0 + (ptrdiff_t)ipy * sw + ipx;                                       \
      for (int i = 0; i < ph; ++i, src += sw, dst += pw)                                     \
        for (int j = 0; j < pw; ++j) {                                                       \
          const WT s0 = src[j] * a11 + src[j + 1] * a12 + src[j + sw] * a21 +                \
                        src[j + sw + 1] * a22;                                               \
          dst[j] = CAST(s0);                                                                 \
        }                                                                                    \
    } else {                                                                                 \
      const pmo_adjust r = adjust_rect(sw, sh, pw, ph, ipx, ipy);                            \
      const T* src = src0 + r.off;                                                           \
      for (int i = 0; i < ph; ++i, dst += pw) {                                              \
        const T* src2 = src + sw;                                                            \
        if (i < r.ry || i >= r.rh) src2 -= sw;                                               \
        WT s0 = src[r.rx] * b1 + src2[r.rx] * b2;                                            \
        for (int j = 0; j < r.rx; ++j) dst[j] = CAST(s0);                                    \
        s0 = src[r.rw] * b1 + src2[r.rw] * b2;                                               \
        for (int j = r.rw; j < pw; ++j) dst[j] = CAST(s0);                                   \
        for (int j = r.rx; j < r.rw; ++j) {                                                  \
          s0 = src[j] * a11 + src[j + 1] * a12 + src2[j] * a21 + src2[j + 1] * a22;          \
          dst[j] = CAST(s0);                                                                 \
        }                                                                                    \
        if (i < r.rh) src = src2;                                                            \
      }                                                                                      \
    }                                                                                        \
  }

static inline int scale_fixpt(float x) { return (int)lrintf(x * 65536.f); }
static inline uint8_t cast_8u(int s) { return (uint8_t)((s + (1 << 15)) >> 16); }
static inline float scale_nop(float x) { return x; }
static inline float cast_nop(float x) { return x; }

PMO_DEFINE_SUBPIX(pmo_get_rect_subpix_u8, uint8_t, int, scale_fixpt, cast_8u)
PMO_DEFINE_SUBPIX(pmo_get_rect_subpix_f32, float, float, scale_nop, cast_nop)

void pmo_dilate_rect_f32(const float* src, int w, int h, int r, float* dst) {
  /* separable max; cells outside the image do not take part (cv::dilate's
   * default border value is -DBL_MAX). */
  float* tmp = (float*)malloc((size_t)w * h * sizeof(float));
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      float m = -FLT_MAX;
      const int x0 = PMO_MAX(x - r, 0), x1 = PMO_MIN(x + r, w - 1);
      for (int k = x0; k <= x1; ++k) m = fmaxf(m, src[(size_t)y * w + k]);
      tmp[(size_t)y * w + x] = m;
    }
  for (int y = 0; y < h; ++y) {
    const int y0 = PMO_MAX(y - r, 0), y1 = PMO_MIN(y + r, h - 1);
    for (int x = 0; x < w; ++x) {
      float m = -FLT_MAX;
      for (int k = y0; k <= y1; ++k) m = fmaxf(m, tmp[(size_t)k * w + x]);
      dst[(size_t)y * w + x] = m;
    }
  }
  free(tmp);
}

/* ---- Philox-4x32-10 (Salmon et al., SC'11), key = 64-bit seed. */
float pmo_philox_u01(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return (float)(c0 >> 8) * 5.9604644775390625e-8f; /* 2^-24 */
}

/* ============================================ (G) GPU-library semantics */

/* GetSubpixel (patchmatch_gpu.cu:18-42) at an integral row: row0 == row1 and
 * trow == 0, so c0 = c00 and c1 = c01 exactly; the column interpolation is
 * (1-t)*c0 + t*c1, which nvcc contracts to fma(1-t, c0, t*c1). */
static inline float g_sample(const float* row, float col) {
  const int c0i = (int)floorf(col), c1i = (int)ceilf(col);
  const float t = col - (float)c0i;
  return fmaf(1.0f - t, row[c0i], t * row[c1i]);
}

/* cost_mode 1: L1GradientCost (patchmatch_gpu.cu:45-69), the full ph x pw patch the 5-tap
 * version was cut down from (dead code in the reference library): 3 x 3 taps in raster order,
 * sample column xr - float(pw/2) + float(col) evaluated left to right, same term. */
static int g_cost_mode = 0;
static int g_radius = 1;   /* patch_size / 2: 1 (the reference's launch sites, :397-408) or 2 */
void pmo_set_cost_mode(int mode) { g_cost_mode = mode; }
void pmo_set_patch_size(int patch_size) { g_radius = patch_size / 2; }

/* cost_mode 1: L1GradientCost (patchmatch_gpu.cu:45-69), the full ph x pw patch the 5-tap version
 * was cut down from (dead code in the reference library), ph = pw = 2r+1: taps in raster order,
 * sample column xr - float(pw/2) + float(col) evaluated left to right, same term. */
static float g_cost_full(const float* Il, const float* Ir, const float* Gl, const float* Gr,
                         int w, int yl, int xl, float xr, float alpha, int r) {
  if (xr > (float)(w - 1 - r)) xr = (float)(w - 1 - r);
  const float w1 = 1 - alpha;
  float cost = 0;
  for (int row = 0; row < 2 * r + 1; ++row)
    for (int col = 0; col < 2 * r + 1; ++col) {
      const size_t lo = (size_t)(yl - r + row) * w + (xl - r + col);
      const float* irow = Ir + (size_t)(yl - r + row) * w;
      const float* grow = Gr + (size_t)(yl - r + row) * w;
      const float xri = (xr - (float)r) + (float)col;
      const float di = fabsf(Il[lo] - g_sample(irow, xri));
      const float dg = fabsf(Gl[lo] - g_sample(grow, xri));
      cost = cost + fmaf(di, alpha, w1 * dg);
    }
  return cost;
}

/* cost_mode 2 (extension; the reference has no census cost): census transform of
 * the (2r+1)^2 window on the intensity, Hamming distance. Reference bit: Il(tap) < Il(centre);
 * matched bit: S(tap) < S(centre) with S the reference's GetSubpixel sampling at columns
 * xr - r + col. The cost is the number of differing bits as a float (0 .. (2r+1)^2 - 1). */
static float g_cost_census(const float* Il, const float* Ir, int w, int yl, int xl, float xr, int r) {
  if (xr > (float)(w - 1 - r)) xr = (float)(w - 1 - r);
  const float cl = Il[(size_t)yl * w + xl];
  const float cr = g_sample(Ir + (size_t)yl * w, (xr - (float)r) + (float)r);
  int ham = 0;
  for (int row = 0; row < 2 * r + 1; ++row)
    for (int col = 0; col < 2 * r + 1; ++col) {
      if (row == r && col == r) continue;
      const float a = Il[(size_t)(yl - r + row) * w + (xl - r + col)];
      const float b = g_sample(Ir + (size_t)(yl - r + row) * w, (xr - (float)r) + (float)col);
      ham += (a < cl) != (b < cr);
    }
  return (float)ham;
}

float pmo_g_cost5(const float* Il, const float* Ir, const float* Gl, const float* Gr,
                  int w, int h, int yl, int xl, float xr, float alpha) {
  if (g_cost_mode == 1) return g_cost_full(Il, Ir, Gl, Gr, w, yl, xl, xr, alpha, g_radius);
  if (g_cost_mode == 2) return g_cost_census(Il, Ir, w, yl, xl, xr, g_radius);
  /* taps TL, TR, C, BL, BR in source order (patchmatch_gpu.cu:84-111);
   * each term alpha*|dI| + (1-alpha)*|dG| contracts to fma(|dI|, alpha, (1-alpha)*|dG|). */
  static const int dy[5] = {-1, -1, 0, 1, 1};
  static const int dx[5] = {-1, 1, 0, -1, 1};
  (void)h;
  /* safety clamp: no effect for d >= 0 (the reference's domain) */
  if (xr > (float)(w - 2)) xr = (float)(w - 2);
  const float w1 = 1 - alpha;
  float cost = 0;
  for (int k = 0; k < 5; ++k) {
    const size_t lo = (size_t)(yl + dy[k]) * w + (xl + dx[k]);
    const float* irow = Ir + (size_t)(yl + dy[k]) * w;
    const float* grow = Gr + (size_t)(yl + dy[k]) * w;
    const float col = xr + (float)dx[k];
    const float di = fabsf(Il[lo] - g_sample(irow, col));
    const float dg = fabsf(Gl[lo] - g_sample(grow, col));
    cost = cost + fmaf(di, alpha, w1 * dg);
  }
  return cost;
}

/* fmaxf(x - d, patch_radius), patchmatch_gpu.cu:162 */
static inline float g_xr(int x, float d) { return fmaxf((float)x - d, (float)g_radius); }

/* cost(d) of a whole map at the sample column the sweeps use, xr = fmaxf(x - d, 1)
 * (patchmatch_gpu.cu:161-162); border pixels get 0. */
void pmo_g_cost_map(const float* Il, const float* Ir, const float* Gl, const float* Gr, int w, int h,
                    const float* disp, float alpha, float* cost) {
  memset(cost, 0, (size_t)w * h * sizeof(float));
  const int r = g_radius;
  for (int y = r; y <= h - 1 - r; ++y)
    for (int x = r; x <= w - 1 - r; ++x)
      cost[(size_t)y * w + x] = pmo_g_cost5(Il, Ir, Gl, Gr, w, h, y, x, g_xr(x, disp[(size_t)y * w + x]), alpha);
}

void pmo_g_add_noise(float* disp, const float* unit_noise, size_t n, float scale) {
  /* threshold(>0) -> mask; scaleAdd(noise, scale, disp); multiply(mask); max(0)
   * (patchmatch_gpu.cu:300-303). scaleAdd is scale*a + b -> fma. Zero results are
   * canonicalised to +0. */
  for (size_t i = 0; i < n; ++i) {
    const float d = disp[i];
    if (d > 0) {
      const float t = fmaf(scale, unit_noise[i], d);
      disp[i] = t > 0 ? t : 0.0f;
    } else {
      disp[i] = 0.0f;
    }
  }
}

/* One lock-step sweep over `nlines` independent lines (rows or columns).
 * line l, position c maps to pixel (y,x) via the strides; chunk k of a line walks
 * positions [max(k*cs-ov, 1), min((k+1)*cs+ov, len-2)) in direction +1, or
 * (min.., max..] downwards in direction -1 (patchmatch_gpu.cu:141-156). All
 * chunks of a line take step i together: loads first, stores after. */
static void g_sweep(const float* Il, const float* Ir, const float* Gl, const float* Gr,
                    int w, int h, float* disp, int dir, float alpha, int chunks, int ov,
                    int along_x) {
  const int len = along_x ? w : h;      /* length of a line */
  const int nlines = along_x ? h : w;   /* number of lines */
  const int cs = len / chunks;          /* iml.cols / blockDim.x */
  int* start = (int*)malloc(sizeof(int) * chunks * 2);
  int* stop = start + chunks;
  int maxsteps = 0;
  for (int k = 0; k < chunks; ++k) {
    const int mn = PMO_MAX(k * cs - ov, g_radius);
    const int mx = PMO_MIN((k + 1) * cs + ov, len - g_radius - 1);
    if (mn >= len) { start[k] = 0; stop[k] = 0; continue; }
    start[k] = dir > 0 ? mn : mx;
    stop[k] = dir > 0 ? mx : mn;
    const int steps = dir > 0 ? (stop[k] - start[k]) : (start[k] - stop[k]);
    if (steps > maxsteps) maxsteps = steps;
  }
  float* newd = (float*)malloc(sizeof(float) * chunks);
  int* newp = (int*)malloc(sizeof(int) * chunks);
  for (int l = g_radius; l <= nlines - 1 - g_radius; ++l) {
    for (int i = 0; i < maxsteps; ++i) {
      int nw = 0;
      for (int k = 0; k < chunks; ++k) {
        const int c = start[k] + dir * i;
        if (!(dir > 0 ? c < stop[k] : c > stop[k])) continue;
        const int y = along_x ? l : c, x = along_x ? c : l;
        const int py = along_x ? l : c - dir, px = along_x ? c - dir : l;
        const float d0 = disp[(size_t)y * w + x];
        const float d1 = disp[(size_t)py * w + px];
        const float cost0 = pmo_g_cost5(Il, Ir, Gl, Gr, w, h, y, x, g_xr(x, d0), alpha);
        const float cost1 = pmo_g_cost5(Il, Ir, Gl, Gr, w, h, y, x, g_xr(x, d1), alpha);
        if (cost1 < cost0) {
          newp[nw] = y * w + x;
          newd[nw] = fminf(d1, (float)x - (float)g_radius);
          ++nw;
        }
      }
      for (int j = 0; j < nw; ++j) disp[newp[j]] = newd[j];
    }
  }
  free(newp);
  free(newd);
  free(start);
}

void pmo_g_propagate_row(const float* Il, const float* Ir, const float* Gl, const float* Gr,
                         int w, int h, float* disp, int dir, float alpha, int chunks, int ov) {
  g_sweep(Il, Ir, Gl, Gr, w, h, disp, dir, alpha, chunks, ov, 1);
}

void pmo_g_propagate_col(const float* Il, const float* Ir, const float* Gl, const float* Gr,
                         int w, int h, float* disp, int dir, float alpha, int chunks, int ov) {
  g_sweep(Il, Ir, Gl, Gr, w, h, disp, dir, alpha, chunks, ov, 0);
}

/* The same sweep in the form the CUDA kernels compute it (pm_sweep.cu): every
 * chunk is an independent chain over the PRE-sweep {d, cost} planes that first
 * replays the head of the next chunk, then walks its own range, and writes only
 * the positions no earlier chunk overwrites. tests/ proves it equal to the
 * lock-step schedule above; it is not used by any other oracle function. */
static void chunk_range(int k, int cs, int ov, int len, int dir, int* start, int* stop) {
  const int mn = PMO_MAX(k * cs - ov, g_radius);
  const int mx = PMO_MIN((k + 1) * cs + ov, len - g_radius - 1);
  *start = dir > 0 ? mn : mx;
  *stop = dir > 0 ? mx : mn;
}

void pmo_g_sweep_chains(const float* Il, const float* Ir, const float* Gl, const float* Gr,
                        int w, int h, const float* d_in, const float* c_in, float* d_out,
                        float* c_out, int along_x, int dir, float alpha, int chunks, int ov) {
  const int len = along_x ? w : h, nlines = along_x ? h : w, cs = len / chunks;
  memcpy(d_out, d_in, (size_t)w * h * sizeof(float));
  memcpy(c_out, c_in, (size_t)w * h * sizeof(float));
#define IDX(l, c) (along_x ? (size_t)(l) * w + (c) : (size_t)(c) * w + (l))
  for (int l = g_radius; l <= nlines - 1 - g_radius; ++l)
    for (int k = 0; k < chunks; ++k) {
      int start, stop;
      chunk_range(k, cs, ov, len, dir, &start, &stop);
      const int nsteps = dir > 0 ? stop - start : start - stop;
      if (nsteps <= 0) continue;
      int n_ov = 0, start_n = 0, n_head = 0;
      const int kn = k + dir, kp = k - dir;
      if (kn >= 0 && kn < chunks) {
        int sn, en;
        chunk_range(kn, cs, ov, len, dir, &sn, &en);
        const int nn = dir > 0 ? en - sn : sn - en;
        if (nn > 0) {
          start_n = sn;
          n_ov = dir > 0 ? stop - sn : sn - stop;
          n_ov = PMO_MAX(0, PMO_MIN(n_ov, PMO_MIN(nn, 16)));
        }
      }
      if (kp >= 0 && kp < chunks) {
        int sp, ep;
        chunk_range(kp, cs, ov, len, dir, &sp, &ep);
        const int np = dir > 0 ? ep - sp : sp - ep;
        if (np > 0) n_head = PMO_MAX(0, dir > 0 ? ep - start : start - ep);
      }
      float hd[16], hc[16];
      if (n_ov > 0) {
        float prev = d_in[IDX(l, start_n - dir)];
        for (int j = 0; j < n_ov; ++j) {
          const int pos = start_n + dir * j;
          const int y = along_x ? l : pos, x = along_x ? pos : l;
          float d = d_in[IDX(l, pos)], c = c_in[IDX(l, pos)];
          const float c1 = pmo_g_cost5(Il, Ir, Gl, Gr, w, h, y, x, g_xr(x, prev), alpha);
          if (c1 < c) { d = fminf(prev, (float)x - (float)g_radius); c = c1; }
          hd[j] = d; hc[j] = c; prev = d;
        }
      }
      float prev = d_in[IDX(l, start - dir)];
      const int first_ov = nsteps - n_ov;
      for (int i = 0; i < nsteps; ++i) {
        const int pos = start + dir * i;
        const int y = along_x ? l : pos, x = along_x ? pos : l;
        float d, c;
        if (i >= first_ov) { d = hd[i - first_ov]; c = hc[i - first_ov]; }
        else { d = d_in[IDX(l, pos)]; c = c_in[IDX(l, pos)]; }
        const float c1 = pmo_g_cost5(Il, Ir, Gl, Gr, w, h, y, x, g_xr(x, prev), alpha);
        if (c1 < c) { d = fminf(prev, (float)x - (float)g_radius); c = c1; }
        prev = d;
        if (i >= n_head) { d_out[IDX(l, pos)] = d; c_out[IDX(l, pos)] = c; }
      }
    }
#undef IDX
}

void pmo_g_mask_background(const float* Il, const float* Ir, const float* Gl, const float* Gr,
                           int w, int h, float* disp, float alpha, float improve) {
  for (int y = g_radius; y <= h - 1 - g_radius; ++y)
    for (int x = g_radius; x <= w - 1 - g_radius; ++x) {
      const float d1 = disp[(size_t)y * w + x];
      const float cost0 = pmo_g_cost5(Il, Ir, Gl, Gr, w, h, y, x, (float)x, alpha);
      const float cost1 = pmo_g_cost5(Il, Ir, Gl, Gr, w, h, y, x, g_xr(x, d1), alpha);
      if (!(cost1 < improve * cost0)) disp[(size_t)y * w + x] = 0.0f;
    }
}

void pmo_g_mask_occlusions(float* displ, const float* dispr, int w, int h, int lr_mode) {
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      const float dl = displ[(size_t)y * w + x];
      const int xr = (int)fmaxf((float)x - dl, 0.0f); /* float index truncated, :289 */
      const float dr = dispr[(size_t)y * w + xr];
      int occluded;
      if (lr_mode == 0) {
        occluded = ((double)dr > 1.4 * (double)dl) || ((double)dr < 0.7 * (double)dl);
      } else {
        occluded = fabsf(dl - dr) > 1.0f;
      }
      if (occluded) displ[(size_t)y * w + x] = 0.0f;
    }
}

/* noise with the improve-only rule (extension, noise_accept = 1): the perturbed,
 * clamped candidate replaces d only where it strictly lowers the cost. Border
 * pixels have no cost and keep d. */
static void g_noise_improve(const float* Il, const float* Ir, const float* Gl, const float* Gr,
                            int w, int h, float* disp, const float* unit_noise, float scale,
                            float alpha, float dmax) {
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      const size_t i = (size_t)y * w + x;
      const float d = disp[i];
      if (!(d > 0)) { disp[i] = 0.0f; continue; }
      if (y < g_radius || y > h - 1 - g_radius || x < g_radius || x > w - 1 - g_radius) continue;
      const float t = fmaf(scale, unit_noise[i], d);
      const float dn = fminf(t > 0 ? t : 0.0f, dmax);
      const float c_old = pmo_g_cost5(Il, Ir, Gl, Gr, w, h, y, x, g_xr(x, d), alpha);
      const float c_new = pmo_g_cost5(Il, Ir, Gl, Gr, w, h, y, x, g_xr(x, dn), alpha);
      if (c_new < c_old) disp[i] = dn;
    }
}

/* extension (north-star "random-search refinement"; the reference has none):
 * after the four sweeps of an iteration every foreground pixel tests K perturbed disparities
 * d + (2u-1) * scale / 2^(k+1), u = Philox(seed; pixel, pair, view<<8|level, 'RS'<<16|iter<<8|k),
 * clamped to [0, dmax], and keeps one only where it strictly lowers the cost. Pixel-local. */
void pmo_x_random_search(const pmo_params* p, const float* Il, const float* Ir, const float* Gl,
                         const float* Gr, int w, int h, uint32_t pair_index, uint32_t view,
                         uint32_t level, uint32_t iter_global, float scale, float dmax, float* disp) {
  const int r = g_radius;
  for (int y = r; y <= h - 1 - r; ++y)
    for (int x = r; x <= w - 1 - r; ++x) {
      const size_t i = (size_t)y * w + x;
      float d = disp[i];
      if (!(d > 0)) continue;
      float c = pmo_g_cost5(Il, Ir, Gl, Gr, w, h, y, x, g_xr(x, d), p->cost_alpha);
      for (int k = 0; k < p->random_search_k; ++k) {
        const float u = pmo_philox_u01(p->seed, (uint32_t)i, pair_index, (view << 8) | level,
                                       0x52530000u | ((iter_global & 0xffu) << 8) | (uint32_t)k);
        const float radius = scale * (1.0f / (float)(1u << (k + 1)));
        const float t = fmaf(2.0f * u - 1.0f, radius, d);
        const float dn = fminf(t > 0 ? t : 0.0f, dmax);
        const float cn = pmo_g_cost5(Il, Ir, Gl, Gr, w, h, y, x, g_xr(x, dn), p->cost_alpha);
        if (cn < c) { d = dn; c = cn; }
      }
      disp[i] = d;
    }
}

void pmo_g_match_view(const pmo_params* p, const float* Il, const float* Ir,
                      const float* Gl, const float* Gr, int w, int h,
                      const float* unit_noise, float level_scale, int iter0,
                      int do_mask, float* disp) {
  pmo_g_match_view_ex(p, Il, Ir, Gl, Gr, w, h, unit_noise, level_scale, iter0, do_mask, 0, 0, 0, disp);
}

void pmo_g_match_view_ex(const pmo_params* p, const float* Il, const float* Ir,
                         const float* Gl, const float* Gr, int w, int h,
                         const float* unit_noise, float level_scale, int iter0,
                         int do_mask, uint32_t pair_index, uint32_t view, uint32_t level,
                         float* disp) {
  const size_t n = (size_t)w * h;
  const float a = p->cost_alpha;
  for (int it = 0; it < p->patchmatch_iters; ++it) {
    /* 32.0 / pow(2.0, iter), patchmatch_gpu.cu:395 */
    const float scale = (float)((double)p->noise_scale0 * (double)level_scale /
                                pow(2.0, (double)(iter0 + it)));
    const float dmax = p->clamp_disp ? (float)p->max_disp * level_scale : INFINITY;
    if (p->noise_accept == 0) {
      pmo_g_add_noise(disp, unit_noise, n, scale);
      if (p->clamp_disp)
        for (size_t i = 0; i < n; ++i) disp[i] = fminf(disp[i], dmax);
    } else {
      g_noise_improve(Il, Ir, Gl, Gr, w, h, disp, unit_noise, scale, a, dmax);
    }
    pmo_g_propagate_row(Il, Ir, Gl, Gr, w, h, disp, +1, a, p->sweep_chunks, p->sweep_overlap);
    pmo_g_propagate_col(Il, Ir, Gl, Gr, w, h, disp, +1, a, p->sweep_chunks, p->sweep_overlap);
    pmo_g_propagate_row(Il, Ir, Gl, Gr, w, h, disp, -1, a, p->sweep_chunks, p->sweep_overlap);
    pmo_g_propagate_col(Il, Ir, Gl, Gr, w, h, disp, -1, a, p->sweep_chunks, p->sweep_overlap);
    if (p->random_search_k > 0)
      pmo_x_random_search(p, Il, Ir, Gl, Gr, w, h, pair_index, view, level, (uint32_t)(iter0 + it),
                          scale, dmax, disp);
  }
  if (do_mask)
    pmo_g_mask_background(Il, Ir, Gl, Gr, w, h, disp, a, p->cost_improve_factor);
}

void pmo_x_random_init(const pmo_params* p, int w, int h, uint32_t pair_index,
                       uint32_t view, uint32_t level, float range, float* disp) {
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      const uint32_t idx = (uint32_t)(y * w + x);
      const float u = pmo_philox_u01(p->seed, idx, pair_index, (view << 8) | level, 0x50524d49u);
      disp[(size_t)y * w + x] = u * range;
    }
}

void pmo_x_upsample2(const float* src, int sw, int sh, int w, int h, float* dst) {
  for (int y = 0; y < h; ++y) {
    const int sy = PMO_MIN(y >> 1, sh - 1);
    for (int x = 0; x < w; ++x) {
      const int sx = PMO_MIN(x >> 1, sw - 1);
      dst[(size_t)y * w + x] = 2.0f * src[(size_t)sy * sw + sx];
    }
  }
}

static int cmp_float(const void* a, const void* b) {
  const float fa = *(const float*)a, fb = *(const float*)b;
  return (fa > fb) - (fa < fb);
}

void pmo_x_median(const float* src, int w, int h, int k, float* dst) {
  const int r = k / 2;
  float win[25];
  memcpy(dst, src, (size_t)w * h * sizeof(float));
  if (k != 3 && k != 5) return;
  for (int y = r; y < h - r; ++y)
    for (int x = r; x < w - r; ++x) {
      int n = 0;
      for (int j = -r; j <= r; ++j)
        for (int i = -r; i <= r; ++i) win[n++] = src[(size_t)(y + j) * w + (x + i)];
      qsort(win, (size_t)n, sizeof(float), cmp_float);
      dst[(size_t)y * w + x] = win[n / 2];
    }
}

void pmo_x_subpixel(const float* Il, const float* Ir, const float* Gl, const float* Gr,
                    int w, int h, float alpha, float* disp) {
  /* parabola through cost(d-1), cost(d), cost(d+1); only where all three samples
   * stay inside the unclamped range and d is a discrete minimum. */
  for (int y = g_radius; y <= h - 1 - g_radius; ++y)
    for (int x = g_radius; x <= w - 1 - g_radius; ++x) {
      const size_t i = (size_t)y * w + x;
      const float d = disp[i];
      if (!(d >= 1.0f) || !((float)x - (d + 1.0f) >= (float)g_radius)) continue;
      const float c0 = pmo_g_cost5(Il, Ir, Gl, Gr, w, h, y, x, (float)x - d, alpha);
      const float cm = pmo_g_cost5(Il, Ir, Gl, Gr, w, h, y, x, (float)x - (d - 1.0f), alpha);
      const float cp = pmo_g_cost5(Il, Ir, Gl, Gr, w, h, y, x, (float)x - (d + 1.0f), alpha);
      const float den = (cm + cp) - 2.0f * c0;
      if (den > 0 && c0 <= cm && c0 <= cp) disp[i] = d + (0.5f * (cm - cp)) / den;
    }
}

int pmo_g_match(const pmo_params* p, const uint8_t* L, const uint8_t* R, int w, int h,
                const float* seed_l, const float* seed_r, uint32_t pair_index,
                float* disp_l, float* disp_r) {
  const int levels = p->pyramid_levels < 1 ? 1 : p->pyramid_levels;
  if (levels > 8) return -1;
  if (p->init_mode == 0 && (!seed_l || !seed_r)) return -2;
  const int saved_cost_mode = g_cost_mode;  /* restored on exit: the stage functions share the switches */
  const int saved_radius = g_radius;
  g_cost_mode = p->cost_mode;
  g_radius = (p->patch_size > 0 ? p->patch_size : 3) / 2;
  int lw[8], lh[8];
  uint8_t* Lp[8];
  uint8_t* Rp[8];
  lw[0] = w; lh[0] = h;
  Lp[0] = (uint8_t*)L; Rp[0] = (uint8_t*)R;
  for (int l = 1; l < levels; ++l) {
    lw[l] = lw[l - 1] / 2; lh[l] = lh[l - 1] / 2;
    Lp[l] = (uint8_t*)malloc((size_t)lw[l] * lh[l]);
    Rp[l] = (uint8_t*)malloc((size_t)lw[l] * lh[l]);
    pmo_resize_half_u8(Lp[l - 1], lw[l - 1], lh[l - 1], Lp[l]);
    pmo_resize_half_u8(Rp[l - 1], lw[l - 1], lh[l - 1], Rp[l]);
  }
  const size_t n0 = (size_t)w * h;
  float* out[2] = {disp_l, (float*)malloc(n0 * sizeof(float))};
  float* prev = (float*)malloc(n0 * sizeof(float));
  for (int view = 0; view < 2; ++view) {
    int pw_ = 0, ph_ = 0;
    for (int l = levels - 1; l >= 0; --l) {
      const int cw = lw[l], ch = lh[l];
      const size_t n = (size_t)cw * ch;
      float* Iref = (float*)malloc(n * sizeof(float));
      float* Imat = (float*)malloc(n * sizeof(float));
      float* Gref = (float*)malloc(n * sizeof(float));
      float* Gmat = (float*)malloc(n * sizeof(float));
      float* tmp = (float*)malloc(n * sizeof(float));
      float* noise = (float*)malloc(n * sizeof(float));
      /* upload + convertTo(CV_32FC1) + GradientMagnitude (:346-352); the right view
       * uses horizontally flipped, swapped planes (:357-367). */
      if (view == 0) {
        pmo_u8_to_f32(Lp[l], n, Iref);
        pmo_u8_to_f32(Rp[l], n, Imat);
        pmo_gradient_mag_u8(Lp[l], cw, ch, Gref);
        pmo_gradient_mag_u8(Rp[l], cw, ch, Gmat);
      } else {
        pmo_u8_to_f32(Rp[l], n, tmp); pmo_flip_h_f32(tmp, cw, ch, Iref);
        pmo_u8_to_f32(Lp[l], n, tmp); pmo_flip_h_f32(tmp, cw, ch, Imat);
        pmo_gradient_mag_u8(Rp[l], cw, ch, tmp); pmo_flip_h_f32(tmp, cw, ch, Gref);
        pmo_gradient_mag_u8(Lp[l], cw, ch, tmp); pmo_flip_h_f32(tmp, cw, ch, Gmat);
      }
      pmo_rng_uniform_f32(p->seed, -1.0f, 1.0f, noise, n); /* :339-344 */
      const float level_scale = 1.0f / (float)(1 << l);
      float* disp = out[view];
      if (l == levels - 1) {
        if (p->init_mode == 1) {
          pmo_x_random_init(p, cw, ch, pair_index, (uint32_t)view, (uint32_t)l,
                            (float)p->max_disp * level_scale, disp);
        } else {
          const float* seed = view == 0 ? seed_l : seed_r;
          for (int y = 0; y < ch; ++y)
            for (int x = 0; x < cw; ++x) {
              const int sx = view == 0 ? x : (cw - 1 - x); /* view coords -> image coords */
              disp[(size_t)y * cw + x] =
                  seed[(size_t)(y << l) * w + ((size_t)sx << l)] * level_scale;
            }
        }
      } else {
        memcpy(prev, disp, (size_t)pw_ * ph_ * sizeof(float));
        pmo_x_upsample2(prev, pw_, ph_, cw, ch, disp);
      }
      pmo_g_match_view_ex(p, Iref, Imat, Gref, Gmat, cw, ch, noise, level_scale,
                          (levels - 1 - l) * p->patchmatch_iters, l == 0, pair_index, (uint32_t)view,
                          (uint32_t)l, disp);
      if (l == 0 && p->subpixel)
        pmo_x_subpixel(Iref, Imat, Gref, Gmat, cw, ch, p->cost_alpha, disp);
      pw_ = cw; ph_ = ch;
      free(Iref); free(Imat); free(Gref); free(Gmat); free(tmp); free(noise);
    }
  }
  /* cu::flip(disp_gpu_r_) then MaskOcclusions (:368-372) */
  pmo_flip_h_f32(out[1], w, h, disp_r);
  pmo_g_mask_occlusions(disp_l, disp_r, w, h, p->lr_mode);
  if (p->median_ksize == 3 || p->median_ksize == 5) {
    memcpy(prev, disp_l, n0 * sizeof(float));
    pmo_x_median(prev, w, h, p->median_ksize, disp_l);
    memcpy(prev, disp_r, n0 * sizeof(float));
    pmo_x_median(prev, w, h, p->median_ksize, disp_r);
  }
  free(prev);
  free(out[1]);
  for (int l = 1; l < levels; ++l) { free(Lp[l]); free(Rp[l]); }
  g_cost_mode = saved_cost_mode;
  g_radius = saved_radius;
  return 0;
}

/* ========================================= (C) CPU stage-library semantics */

static inline uint8_t sat_u8_from_f32(float v) {
  /* cv::saturate_cast<uchar>(float): cvRound (half to even) then clamp */
  const long r = lrintf(v);
  return (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
}

float pmo_c_cost(const uint8_t* Il, const uint8_t* Ir, const float* Gl, const float* Gr,
                 int w, int h, int x, int y, float d, int pw, int ph) {
  uint8_t ref[81], cand[81];
  float gref[81], gcand[81];
  const int n = pw * ph;
  /* GetPatchSubpix (patchmatch.cpp:98-111) */
  pmo_get_rect_subpix_u8(Il, w, h, pw, ph, (float)x, (float)y, ref);
  pmo_get_rect_subpix_f32(Gl, w, h, pw, ph, (float)x, (float)y, gref);
  pmo_get_rect_subpix_u8(Ir, w, h, pw, ph, (float)x - d, (float)y, cand);
  pmo_get_rect_subpix_f32(Gr, w, h, pw, ph, (float)x - d, (float)y, gcand);
  /* L1GradientCostFunction (patchmatch_test.cpp:30-45). The gradient patches
   * reach the functor through `const Image1b&` parameters, i.e. they are
   * converted f32 -> u8 first. cv::mean = sum * (1/N) in double. */
  int sc = 0, sg = 0;
  for (int i = 0; i < n; ++i) {
    sc += abs((int)ref[i] - (int)cand[i]);
    sg += abs((int)sat_u8_from_f32(gref[i]) - (int)sat_u8_from_f32(gcand[i]));
  }
  const double inv = 1.0 / (double)n;
  const float alpha = 0.7f, tau_color = 50.0f, tau_grad = 20.0f;
  const float error_color = fminf((float)((double)sc * inv), tau_color);
  const float error_grad = fminf((float)((double)sg * inv), tau_grad);
  return alpha * error_color + (1 - alpha) * error_grad;
}

void pmo_c_add_noise(float* disp, int w, int h, float amount) {
  /* fresh cv::RNG(123) on every call (patchmatch.cpp:146); cv::add under
   * mask = disp > 0, then max(disp, 0). */
  const size_t n = (size_t)w * h;
  float* noise = (float*)malloc(n * sizeof(float));
  pmo_rng_uniform_f32(123, -amount, amount, noise, n);
  for (size_t i = 0; i < n; ++i) {
    float d = disp[i];
    if (d > 0) d = d + noise[i];
    disp[i] = d > 0 ? d : 0.0f;
  }
  free(noise);
}

static inline int c_is_border(int x, int y, int w, int h, int pw, int ph) {
  return y < ph / 2 || x < pw / 2 || y > h - ph / 2 - 1 || x > w - pw / 2 - 1;
}

/* PropagateNeighbors, single-neighbour overload (patchmatch.cpp:158-196) */
static void c_propagate_px(const uint8_t* Il, const uint8_t* Ir, const float* Gl, const float* Gr,
                           int w, int h, int x, int y, float* disp, int ph, int pw,
                           int xo, int yo) {
  float d0 = disp[(size_t)y * w + x];
  d0 = fminf(fmaxf(d0, 0.0f), (float)x - (float)(pw / 2));
  const float dl = disp[(size_t)(y + yo) * w + (x + xo)];
  const float cost_cur = pmo_c_cost(Il, Ir, Gl, Gr, w, h, x, y, d0, pw, ph);
  float best = d0;
  if (((float)x - dl) >= (float)(pw / 2)) {
    const float cost_n = pmo_c_cost(Il, Ir, Gl, Gr, w, h, x, y, dl, pw, ph);
    if (cost_n < cost_cur) best = dl; /* Argmin keeps the first minimum, :115-126 */
  }
  disp[(size_t)y * w + x] = best;
}

void pmo_c_propagate_pass(const uint8_t* Il, const uint8_t* Ir, const float* Gl, const float* Gr,
                          int w, int h, float* disp, int ph, int pw, int pass) {
  if (pass == 0 || pass == 1) { /* patchmatch.cpp:264-285 */
    for (int y = 1; y < h; ++y)
      for (int x = 1; x < w; ++x) {
        if (c_is_border(x, y, w, h, pw, ph)) continue;
        c_propagate_px(Il, Ir, Gl, Gr, w, h, x, y, disp, ph, pw, pass == 0 ? -1 : 0,
                       pass == 0 ? 0 : -1);
      }
  } else { /* patchmatch.cpp:288-310 */
    for (int y = h - 2; y >= 0; --y)
      for (int x = w - 2; x >= 0; --x) {
        if (c_is_border(x, y, w, h, pw, ph)) continue;
        c_propagate_px(Il, Ir, Gl, Gr, w, h, x, y, disp, ph, pw, pass == 2 ? 1 : 0,
                       pass == 2 ? 0 : 1);
      }
  }
}

void pmo_c_propagate(const uint8_t* Il, const uint8_t* Ir, const float* Gl, const float* Gr,
                     int w, int h, float* disp, int ph, int pw) {
  for (int pass = 0; pass < 4; ++pass)
    pmo_c_propagate_pass(Il, Ir, Gl, Gr, w, h, disp, ph, pw, pass);
}

void pmo_c_remove_background(const uint8_t* Il, const uint8_t* Ir, const float* Gl, const float* Gr,
                             int w, int h, float* disp, int ph, int pw, float win_by_factor) {
  for (int y = 1; y < h; ++y)
    for (int x = 1; x < w; ++x) {
      if (c_is_border(x, y, w, h, pw, ph)) continue;
      float d0 = disp[(size_t)y * w + x];
      d0 = fminf(fmaxf(d0, 0.0f), (float)x - (float)(pw / 2));
      const float cost_cur = pmo_c_cost(Il, Ir, Gl, Gr, w, h, x, y, d0, pw, ph);
      const float cost_zero = pmo_c_cost(Il, Ir, Gl, Gr, w, h, x, y, 0.0f, pw, ph);
      if (cost_cur > (cost_zero / win_by_factor)) disp[(size_t)y * w + x] = 0.0f;
    }
}

void pmo_c_estimate_disparity(const uint8_t* Il, const uint8_t* Ir, int w, int h, float* disp) {
  const size_t n = (size_t)w * h;
  float* Gl = (float*)malloc(n * sizeof(float));
  float* Gr = (float*)malloc(n * sizeof(float));
  pmo_gradient_mag_u8(Il, w, h, Gl); /* ComputeGradient, patchmatch_test.cpp:48-64 */
  pmo_gradient_mag_u8(Ir, w, h, Gr);
  static const float amount[4] = {32.0f, 8.0f, 2.0f, 0.5f}; /* patchmatch_test.cpp:173-180 */
  static const int patch[4] = {5, 5, 3, 3};
  for (int s = 0; s < 4; ++s) {
    pmo_c_add_noise(disp, w, h, amount[s]);
    pmo_c_propagate(Il, Ir, Gl, Gr, w, h, disp, patch[s], patch[s]);
  }
  pmo_c_remove_background(Il, Ir, Gl, Gr, w, h, disp, 3, 3, 1.5f); /* :183 */
  free(Gl);
  free(Gr);
}

/* ============================== the consumer: disparity -> depth / points */

/* StereoCamera::DispToDepth (vision_core/stereo_camera.cpp:49-53) and
 * PinholeCamera::Backproject (vision_core/pinhole_camera.cpp:41-45) per pixel, with the
 * resolution handling of mesher/object_mesher.cpp:147-150. K^-1 in closed form
 * (1/fx, 1/fy, -cx/fx, -cy/fy); the reference inverts K with Eigen, so its last bits may
 * differ: compare at 1e-12 relative. disparity <= 0 -> 0. */
void pmo_x_disp_to_depth(const float* disp, int w, int h, double fx, double fy, double cx,
                         double cy, double baseline, double scale, float* depth, float* xyz) {
  const double fxb = fx * baseline, ifx = 1.0 / fx, ify = 1.0 / fy;
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      const float d = disp[(size_t)y * w + x];
      double z = 0.0;
      if (d > 0.0f) z = fxb / ((double)d / scale);
      if (depth) depth[(size_t)y * w + x] = (float)z;
      if (xyz) {
        float* p = xyz + ((size_t)y * w + x) * 3;
        const double u = (double)x / scale, v = (double)y / scale;
        p[0] = (float)(z * ((u - cx) * ifx));
        p[1] = (float)(z * ((v - cy) * ify));
        p[2] = (float)z;
      }
    }
}

/* ================================= ForegroundTextureMask (patchmatch.cpp:19-49) */

/* cv::morphologyEx(MORPH_GRADIENT) with a (2k+1)^2 rectangle anchored at its centre and the default
 * border (morphologyDefaultBorderValue: pixels outside the image never win the max / min):
 * dilate - erode on u8. */
static void morph_gradient_u8(const uint8_t* src, int w, int h, int k, uint8_t* dst) {
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      int mx = 0, mn = 255;
      for (int j = PMO_MAX(y - k, 0); j <= PMO_MIN(y + k, h - 1); ++j)
        for (int i = PMO_MAX(x - k, 0); i <= PMO_MIN(x + k, w - 1); ++i) {
          const int v = src[(size_t)j * w + i];
          if (v > mx) mx = v;
          if (v < mn) mn = v;
        }
      dst[(size_t)y * w + x] = (uint8_t)(mx - mn);
    }
}

/* cv::resize(u8, INTER_LINEAR) to exactly twice the size: the fixed-point path of OpenCV's resize
 * (HResizeLinear with 11-bit coefficients, then VResizeLinear<uchar>:
 * ((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2), source indices clamped at the border. */
static void resize_up2_u8(const uint8_t* src, int sw, int sh, uint8_t* dst) {
  const int w = 2 * sw, h = 2 * sh;
  for (int y = 0; y < h; ++y) {
    int sy = (y + 1) / 2 - 1;                 /* floor((y + 0.5) / 2 - 0.5) */
    int b1 = (y & 1) ? 512 : 1536, b0;        /* fraction 0.25 (odd y) or 0.75 (even y), x 2048 */
    if (sy < 0) { sy = 0; b1 = 0; }
    if (sy >= sh - 1) { sy = sh - 1; b1 = 0; }
    b0 = 2048 - b1;
    const int sy1 = PMO_MIN(sy + 1, sh - 1);
    for (int x = 0; x < w; ++x) {
      int sx = (x + 1) / 2 - 1;
      int a1 = (x & 1) ? 512 : 1536, a0;
      if (sx < 0) { sx = 0; a1 = 0; }
      if (sx >= sw - 1) { sx = sw - 1; a1 = 0; }
      a0 = 2048 - a1;
      const int sx1 = PMO_MIN(sx + 1, sw - 1);
      const int S0 = src[(size_t)sy * sw + sx] * a0 + src[(size_t)sy * sw + sx1] * a1;
      const int S1 = src[(size_t)sy1 * sw + sx] * a0 + src[(size_t)sy1 * sw + sx1] * a1;
      dst[(size_t)y * w + x] = (uint8_t)((((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2);
    }
  }
}

int pmo_c_foreground_texture_mask(const uint8_t* gray, int w, int h, int ksize, double min_grad,
                                  int downsize, uint8_t* mask) {
  if (downsize < 1 || downsize > 2) return -1;      /* the reference allows 1..8; 1 and 2 are restated */
  const int sk = ksize / downsize;
  if (sk <= 1) return -2;                           /* CHECK_GT(scaled_ksize, 1) */
  if (downsize == 1) {
    uint8_t* g = (uint8_t*)malloc((size_t)w * h);
    morph_gradient_u8(gray, w, h, sk, g);
    for (size_t i = 0; i < (size_t)w * h; ++i) mask[i] = (double)g[i] > min_grad ? 255 : 0;
    free(g);
    return 0;
  }
  if ((w | h) & 1) return -3;                       /* exact halving only */
  const int sw = w / 2, sh = h / 2;
  uint8_t* small = (uint8_t*)malloc((size_t)sw * sh * 3);
  uint8_t* grad = small + (size_t)sw * sh;
  uint8_t* m = grad + (size_t)sw * sh;
  pmo_resize_half_u8(gray, w, h, small);            /* INTER_LINEAR at scale 2 == 2x2 area average */
  morph_gradient_u8(small, sw, sh, sk, grad);
  for (size_t i = 0; i < (size_t)sw * sh; ++i) m[i] = (double)grad[i] > min_grad ? 255 : 0;
  resize_up2_u8(m, sw, sh, mask);
  free(small);
  return 0;
}
