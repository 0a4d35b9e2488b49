/*
 * pm_oracle_seed.c -- CPU restatement of the reference's sparse seeding step.
 * TEST INFRASTRUCTURE ONLY (see pm_oracle.h).  Plain C99, single threaded.
 *
 * Restates (citations relative to /root/reference):
 *   PatchmatchGpu::SparseInit            src/vehicle/patchmatch_gpu/patchmatch_gpu.cu:414-442
 *   Patchmatch::Initialize               src/vehicle/stereo_matching/patchmatch.cpp:52-87
 *   FeatureDetector::Detect              src/vehicle/feature_tracking/feature_detector.cpp:89-122
 *   StereoMatcher::MatchRectified        src/vehicle/feature_tracking/stereo_matcher.cpp:22-116
 * and the OpenCV 3.4 primitives they call, which are not vendored in the reference
 * (CMakeLists.txt:28 pins OpenCV 3.4.0): cv::GFTTDetector -> cv::goodFeaturesToTrack
 * (imgproc/featureselect.cpp: cornerMinEigenVal / cornerHarris, threshold at
 * quality * max, 3x3 dilate local maxima, descending sort, greedy minimum distance) and
 * cv::matchTemplate(TM_SQDIFF_NORMED) + cv::minMaxLoc (imgproc/templmatch.cpp).
 *
 * Arithmetic.  OpenCV evaluates both primitives in float32 (scaled Sobel, sliding box
 * sums, DFT cross-correlation), so its last bits depend on SIMD width and FMA use
 * (measured here: cv2's own Sobel rounds differently in its vector body and scalar
 * tail).  Every quantity involved is an integer function of the u8 images, so this
 * restatement evaluates them EXACTLY (int64) and rounds once at the end:
 *   min-eigenvalue response = s^2/2 * ((A+C) - sqrt((A-C)^2 + 4 B^2)),  s = 1/(4*block*255),
 *   A,B,C = box sums of dx^2, dx*dy, dy^2 with dx,dy the unscaled 3x3 Sobel responses;
 *   SQDIFF_NORMED = sum (T-I)^2 / (sqrt(sum I^2) * sqrt(sum T^2)).
 * It is the infinitely precise version of what OpenCV approximates (relative deviation
 * measured at 2e-7 of the response maximum) and picks the same keypoints, in the same
 * order, with the same disparities as cv2 4.13 on the reference's fixtures
 * (tests/golden/seeding.npz, tests/test_oracle_seeding.py).
 */
#include "pm_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define S_MIN(a, b) ((a) < (b) ? (a) : (b))
#define S_MAX(a, b) ((a) > (b) ? (a) : (b))

void pmo_seed_params_default(pmo_seed_params* p) {
  p->max_features = 200;      /* feature_detector.hpp:35 */
  p->min_distance = 20;       /* :38 */
  p->quality_level = 0.01;    /* :39 */
  p->block_size = 5;          /* :40 */
  p->use_harris = 0;          /* :41 */
  p->harris_k = 0.04;         /* :42 */
  p->templ_cols = 31;         /* stereo_matcher.hpp:21 */
  p->templ_rows = 11;         /* :22 */
  p->max_disp = 128;          /* :23 */
  p->max_matching_cost = 0.15; /* :24 */
}

/* BORDER_REFLECT_101 for any offset (cv::borderInterpolate). */
static int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) {
    if (i < 0) i = -i;
    else i = 2 * n - 2 - i;
  }
  return i;
}

/* cv::cornerMinEigenVal / cv::cornerHarris (imgproc/corner.cpp, cornerEigenValsVecs with
 * aperture 3): response of every pixel, float32. */
void pmo_s_corner_response(const uint8_t* im, int w, int h, int block, int harris, double k,
                           float* out) {
  const size_t n = (size_t)w * h;
  int32_t* dx = (int32_t*)malloc(n * sizeof(int32_t));
  int32_t* dy = (int32_t*)malloc(n * sizeof(int32_t));
  for (int y = 0; y < h; ++y) {
    const uint8_t* r0 = im + (size_t)reflect101(y - 1, h) * w;
    const uint8_t* r1 = im + (size_t)y * w;
    const uint8_t* r2 = im + (size_t)reflect101(y + 1, h) * w;
    for (int x = 0; x < w; ++x) {
      const int xm = reflect101(x - 1, w), xp = reflect101(x + 1, w);
      dx[(size_t)y * w + x] = (r0[xp] - r0[xm]) + 2 * (r1[xp] - r1[xm]) + (r2[xp] - r2[xm]);
      dy[(size_t)y * w + x] = (r2[xm] - r0[xm]) + 2 * (r2[x] - r0[x]) + (r2[xp] - r0[xp]);
    }
  }
  /* boxFilter(cov, block x block, anchor = block/2, normalize = false, BORDER_REFLECT_101) */
  const int a0 = block / 2;
  const double s = 1.0 / (4.0 * (double)block * 255.0);
  const double k_eig = 0.5 * s * s, k_har = (s * s) * (s * s);
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      int64_t A = 0, B = 0, C = 0;
      for (int j = 0; j < block; ++j) {
        const int yy = reflect101(y - a0 + j, h);
        for (int i = 0; i < block; ++i) {
          const int xx = reflect101(x - a0 + i, w);
          const int64_t gx = dx[(size_t)yy * w + xx], gy = dy[(size_t)yy * w + xx];
          A += gx * gx; B += gx * gy; C += gy * gy;
        }
      }
      double r;
      if (!harris) {
        const int64_t dif = A - C;
        const double disc = sqrt((double)(dif * dif + 4 * B * B));
        r = k_eig * ((double)(A + C) - disc);
      } else {
        const double det = (double)(A * C - B * B);
        const double tr = (double)(A + C);
        r = k_har * (det - (k * tr) * tr);
      }
      out[(size_t)y * w + x] = (float)r;
    }
  free(dx);
  free(dy);
}

typedef struct { float v; int ofs; } s_cand;

/* featureselect.cpp greaterThanPtr: value descending, ties by address descending. */
static int cand_cmp(const void* a, const void* b) {
  const s_cand* p = (const s_cand*)a;
  const s_cand* q = (const s_cand*)b;
  if (p->v > q->v) return -1;
  if (p->v < q->v) return 1;
  return p->ofs > q->ofs ? -1 : (p->ofs < q->ofs ? 1 : 0);
}

/* cv::goodFeaturesToTrack with an all-pass mask, as FeatureDetector::Detect runs it with
 * no tracked keypoints (feature_detector.cpp:96-102); the ANMS step that follows
 * (:106-108) returns its input because GFTT already caps the count at
 * max_features_per_frame (:67-69). Returns the number of keypoints written (<= max_features). */
int pmo_s_good_features(const uint8_t* im, int w, int h, const pmo_seed_params* sp, int* kx,
                        int* ky, int* n_candidates) {
  const size_t n = (size_t)w * h;
  float* eig = (float*)malloc(n * sizeof(float));
  pmo_s_corner_response(im, w, h, sp->block_size, sp->use_harris, sp->harris_k, eig);
  float mx = eig[0];
  for (size_t i = 1; i < n; ++i) mx = eig[i] > mx ? eig[i] : mx;
  /* threshold(eig, eig, maxVal * qualityLevel, 0, THRESH_TOZERO): the threshold is cast to float */
  const float thr = (float)((double)mx * sp->quality_level);
  for (size_t i = 0; i < n; ++i) eig[i] = eig[i] > thr ? eig[i] : 0.0f;
  s_cand* c = (s_cand*)malloc(n * sizeof(s_cand));
  size_t nc = 0;
  for (int y = 1; y < h - 1; ++y)
    for (int x = 1; x < w - 1; ++x) {
      const float v = eig[(size_t)y * w + x];
      if (v == 0.0f) continue;
      float m = -FLT_MAX;  /* dilate(eig, tmp, Mat()): 3x3, pixels outside do not take part */
      for (int j = -1; j <= 1; ++j)
        for (int i = -1; i <= 1; ++i) m = fmaxf(m, eig[(size_t)(y + j) * w + (x + i)]);
      if (v == m) { c[nc].v = v; c[nc].ofs = y * w + x; ++nc; }
    }
  qsort(c, nc, sizeof(s_cand), cand_cmp);
  if (n_candidates) *n_candidates = (int)nc;
  int na = 0;
  const int md2 = sp->min_distance * sp->min_distance;
  for (size_t i = 0; i < nc && na < sp->max_features; ++i) {
    const int y = c[i].ofs / w, x = c[i].ofs - y * w;
    int good = 1;
    if (sp->min_distance >= 1)
      /* the cell grid of featureselect.cpp only accelerates this test (cell size =
       * minDistance, 3x3 cells searched) */
      for (int j = 0; j < na; ++j) {
        const int ddx = x - kx[j], ddy = y - ky[j];
        if (ddx * ddx + ddy * ddy < md2) { good = 0; break; }
      }
    if (good) { kx[na] = x; ky[na] = y; ++na; }
  }
  free(c);
  free(eig);
  return na;
}

/* StereoMatcher::MatchRectified for one keypoint (stereo_matcher.cpp:22-116),
 * subpixel_refinement = false (as in the PatchMatch drivers, patchmatch_gpu_test.cpp:76).
 * Keypoints are integral (GFTT without cornerSubPix), so round() is the identity. */
double pmo_s_match_rectified(const uint8_t* L, const uint8_t* R, int w, int h,
                             const pmo_seed_params* sp, int kpx, int kpy) {
  const int tc = sp->templ_cols, tr = sp->templ_rows, md = sp->max_disp;
  const int stripe_rows = tr + 2;
  const int ty = kpy - (tr - 1) / 2;
  if (ty < 0 || ty + tr >= h) return -1.0;
  int offset_x = 0;
  int tx = kpx - (tc - 1) / 2;
  if (tx < 0) { offset_x = tx; tx = 0; }
  if (tx + tc >= w) {
    offset_x = (tx + tc) - (w - 1);
    tx -= offset_x;
  }
  const int sy = kpy - (stripe_rows - 1) / 2;
  if (sy < 0 || sy + stripe_rows >= h) return -1.0;
  int sx = kpx + (tc - 1) / 2 - md;
  if (sx + md > w - 1) sx -= (sx + md) - (w - 1);
  if (sx < 0) sx = 0;
  /* matchTemplate(stripe, patch, TM_SQDIFF_NORMED) -> float32 result; minMaxLoc takes the
   * first minimum in raster order */
  int64_t t2 = 0;
  for (int j = 0; j < tr; ++j)
    for (int i = 0; i < tc; ++i) {
      const int64_t t = L[(size_t)(ty + j) * w + tx + i];
      t2 += t * t;
    }
  const double tnorm = sqrt((double)t2);
  float best = FLT_MAX;
  int best_i = 0;
  for (int pj = 0; pj <= stripe_rows - tr; ++pj)
    for (int pi = 0; pi <= md - tc; ++pi) {
      int64_t ssd = 0, w2 = 0;
      for (int j = 0; j < tr; ++j) {
        const uint8_t* a = L + (size_t)(ty + j) * w + tx;
        const uint8_t* b = R + (size_t)(sy + pj + j) * w + sx + pi;
        for (int i = 0; i < tc; ++i) {
          const int64_t d = (int64_t)a[i] - b[i];
          ssd += d * d;
          w2 += (int64_t)b[i] * b[i];
        }
      }
      double num = (double)ssd;
      const double t = sqrt((double)w2) * tnorm;
      if (fabs(num) < t) num /= t;
      else if (fabs(num) < t * 1.125) num = num > 0 ? 1.0 : -1.0;
      else num = 1.0;
      const float r = (float)num;
      if (r < best) { best = r; best_i = pi; }
    }
  const int mx = best_i + sx + (tc - 1) / 2 + offset_x;
  if ((double)best < sp->max_matching_cost && kpx >= mx) return (double)((float)kpx - (float)mx);
  return -1.0;
}

/* zero map + keypoint disparities (patchmatch_gpu.cu:423-433, patchmatch.cpp:62-72) */
static int scatter_seeds(const uint8_t* L, const uint8_t* R, int w, int h,
                         const pmo_seed_params* sp, float* seeds) {
  int* kx = (int*)malloc(sizeof(int) * (size_t)sp->max_features * 2);
  int* ky = kx + sp->max_features;
  const int nk = pmo_s_good_features(L, w, h, sp, kx, ky, NULL);
  memset(seeds, 0, (size_t)w * h * sizeof(float));
  for (int i = 0; i < nk; ++i) {
    const double d = pmo_s_match_rectified(L, R, w, h, sp, kx[i], ky[i]);
    if ((float)d >= 0) seeds[(size_t)ky[i] * w + kx[i]] = (float)d;
  }
  free(kx);
  return nk;
}

/* PatchmatchGpu::SparseInit (patchmatch_gpu.cu:414-442). */
void pmo_s_sparse_init(const uint8_t* L, const uint8_t* R, int w, int h,
                       const pmo_seed_params* sp, int dilate_factor, float* seeds) {
  float* tmp = (float*)malloc((size_t)w * h * sizeof(float));
  scatter_seeds(L, R, w, h, sp, tmp);
  const int r = (1 << dilate_factor) + 1; /* (int)pow(2, f) + 1; element (2r+1)^2 anchored at r */
  pmo_dilate_rect_f32(tmp, w, h, r, seeds);
  free(tmp);
}

/* Patchmatch::Initialize (patchmatch.cpp:52-87); seeds has (w/f) x (h/f) elements. */
void pmo_c_initialize(const uint8_t* L, const uint8_t* R, int w, int h,
                      const pmo_seed_params* sp, int f, float* seeds) {
  const size_t n = (size_t)w * h;
  float* tmp = (float*)malloc(2 * n * sizeof(float));
  float* dil = tmp + n;
  scatter_seeds(L, R, w, h, sp, tmp);
  const int r = (1 << (f - 1)) + 1; /* (int)pow(2, f - 1) + 1 */
  pmo_dilate_rect_f32(tmp, w, h, r, dil);
  /* cv::resize(size / f, INTER_NEAREST): src = min(floor(dst * (src_size / dst_size)), src_size - 1)
   * then disps /= pow(2, f)  (the quirk of patchmatch.cpp:81: not f) */
  const int ow = w / f, oh = h / f;
  const double fx = (double)w / ow, fy = (double)h / oh;
  const float div = (float)(1 << f);
  for (int y = 0; y < oh; ++y) {
    const int sy = S_MIN((int)floor(y * fy), h - 1);
    for (int x = 0; x < ow; ++x) {
      const int sx = S_MIN((int)floor(x * fx), w - 1);
      seeds[(size_t)y * ow + x] = dil[(size_t)sy * w + sx] / div;
    }
  }
  free(tmp);
}

/* The seeds PatchmatchGpu::Match (host overload) computes for its two views
 * (patchmatch_gpu.cu:335, 357-365): left = SparseInit(L, R); right =
 * SparseInit(flip(R), flip(L)), returned here flipped back to right-image coordinates
 * (the convention of pmo_g_match's seed_r). */
void pmo_s_match_seeds(const uint8_t* L, const uint8_t* R, int w, int h,
                       const pmo_seed_params* sp, int dilate_factor, float* seed_l,
                       float* seed_r) {
  const size_t n = (size_t)w * h;
  pmo_s_sparse_init(L, R, w, h, sp, dilate_factor, seed_l);
  uint8_t* Lf = (uint8_t*)malloc(2 * n);
  uint8_t* Rf = Lf + n;
  float* tmp = (float*)malloc(n * sizeof(float));
  pmo_flip_h_u8(L, w, h, Lf);
  pmo_flip_h_u8(R, w, h, Rf);
  pmo_s_sparse_init(Rf, Lf, w, h, sp, dilate_factor, tmp);
  pmo_flip_h_f32(tmp, w, h, seed_r);
  free(tmp);
  free(Lf);
}
