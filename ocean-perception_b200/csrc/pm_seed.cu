// pm_seed.cu -- the reference's sparse seeding step on the GPU (sm_100a).
//
// PatchmatchGpu::SparseInit (patchmatch_gpu.cu:414-442) runs on the host in the reference:
// FeatureDetector::Detect (feature_detector.cpp:89-122; cv::GFTTDetector ==
// cv::goodFeaturesToTrack), StereoMatcher::MatchRectified (stereo_matcher.cpp:22-116;
// cv::matchTemplate TM_SQDIFF_NORMED + cv::minMaxLoc) and a rectangular cv::dilate, twice per
// pair. Here every seeding problem (view 2p: left reference; view 2p+1: the flipped right
// image against the flipped left one, patchmatch_gpu.cu:362-365) of a device pass runs through
// the same kernels:
//   k_seed_response    3x3 Sobel (dx, dy) of a u8 tile in shared memory, box sums of dx^2, dx*dy,
//                      dy^2 (rolling window) -> min-eigenvalue / Harris response, per-view maximum
//   k_seed_candidates  threshold at quality*max, 3x3 local maxima -> 64-bit sort keys
//   k_seed_select      one block per view: bitonic sort (value desc, address desc) and the
//                      greedy minimum-distance selection of goodFeaturesToTrack
//   k_seed_match       one block per keypoint: SQDIFF_NORMED over the search stripe, first
//                      minimum, acceptance test -> keypoint disparity
//   k_seed_paint       scatter + rectangular dilate (+ nearest resize and division for
//                      Patchmatch::Initialize, patchmatch.cpp:75-81) -> seed maps
// Arithmetic: every quantity is an integer function of the u8 images, so sums are
// exact integers and each result is rounded once, in double, then to float.
// Citations are relative to /root/reference.
#include <climits>
#include <cmath>

#include "pm_kernels.h"

namespace pm {

#define PM_LAUNCH_CHECK(n) (cudaGetLastError() == cudaSuccess ? (n) : -1)

static inline unsigned cdiv(long a, long b) { return (unsigned)((a + b - 1) / b); }

namespace {

// BORDER_REFLECT_101 for any offset (cv::borderInterpolate)
__device__ __forceinline__ int reflect_any(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
  return i;
}

// reference image of problem v: L for even v, R read right-to-left for odd v
__device__ __forceinline__ const uint8_t* ref_image(const SeedImages& im, int v) {
  return ((v & 1) ? im.R : im.L) + (size_t)(v >> 1) * im.iplane;
}
__device__ __forceinline__ const uint8_t* mat_image(const SeedImages& im, int v) {
  return ((v & 1) ? im.L : im.R) + (size_t)(v >> 1) * im.iplane;
}
__device__ __forceinline__ int px(const uint8_t* img, const SeedImages& im, int v, int x, int y) {
  return img[(size_t)y * im.ipitch + ((v & 1) ? im.w - 1 - x : x)];
}

// floats ordered as unsigned integers (any sign)
__device__ __forceinline__ unsigned ord_of(float f) {
  const unsigned b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float float_of(unsigned o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

}  // namespace

// ---------------------------------------------------------------- response

constexpr int kRespTW = 128, kRespTH = 32;

__device__ __forceinline__ float response_of(long long A, long long B, long long C, int harris,
                                             double k, double k_eig, double k_har) {
  double r;
  if (!harris) {
    const long long dif = A - C;
    const double disc = sqrt((double)(dif * dif + 4 * B * B));
    r = k_eig * ((double)(A + C) - disc);
  } else {
    const double det = (double)(A * C - B * B);
    const double tr = (double)(A + C);
    r = k_har * (det - (k * tr) * tr);
  }
  return (float)r;
}

// The same value from 32-bit sums (box sizes up to 7: |A|, |B|, |C| < 2^26). Every intermediate is an
// integer that a double holds exactly (squares and products below 2^53; 4*B*B is an exact scaling),
// and the one sum that can exceed 2^53 rounds exactly as the conversion of the exact integer would.
__device__ __forceinline__ float response_of_i32(int A, int B, int C, int harris, double k,
                                                 double k_eig, double k_har) {
  const double a = (double)A, b = (double)B, c = (double)C;
  double r;
  if (!harris) {
    const double dif = a - c;
    const double disc = sqrt(dif * dif + 4.0 * (b * b));
    r = k_eig * ((a + c) - disc);
  } else {
    const double det = a * c - b * b;
    const double tr = a + c;
    r = k_har * (det - (k * tr) * tr);
  }
  return (float)r;
}

// BS > 0: compile-time box size with a rolling window of row sums; BS == 0: any size, direct.
// Shared memory: the u8 tile the gradients need (border-reflected pixels, BORDER_REFLECT_101 of
// cv::Sobel), then the packed (dx, dy) of rows [ty0-a0, ty0-a0+GH) x cols [tx0-a0, tx0-a0+GW).
template <int BS>
__global__ void __launch_bounds__(kRespTW)
k_seed_response(SeedImages im, int bsz, int harris, double k, float* __restrict__ resp, int rpitch,
                size_t rplane, unsigned* __restrict__ vmax) {
  extern __shared__ int s_tile[];
  const int b = BS > 0 ? BS : bsz;
  const int a0 = b / 2;
  const int GW = kRespTW + b - 1, GH = kRespTH + b - 1;
  const int UW = GW + 2, UH = GH + 2, UP = (UW + 3) & ~3;
  unsigned char* s_u8 = reinterpret_cast<unsigned char*>(s_tile + GW * GH);
  const int w = im.w, h = im.h;
  const int v = blockIdx.z;
  const int tx0 = blockIdx.x * kRespTW, ty0 = blockIdx.y * kRespTH;
  const int ux0 = tx0 - a0 - 1, uy0 = ty0 - a0 - 1;  // image coordinates of u8 slot (0, 0)
  const uint8_t* img = ref_image(im, v);
  // rows outer (row index and reflection uniform per iteration), columns per thread (source
  // column resolved once): the per-element work is one byte load and one byte store
  {
    int xsrc[2], cidx[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      cidx[q] = threadIdx.x + q * kRespTW;
      const int xr = reflect_any(ux0 + cidx[q], w);
      xsrc[q] = (v & 1) ? w - 1 - xr : xr;
    }
    for (int r = 0; r < UH; ++r) {
      const uint8_t* row = img + (size_t)reflect_any(uy0 + r, h) * im.ipitch;
#pragma unroll
      for (int q = 0; q < 2; ++q)
        if (cidx[q] < UW) s_u8[r * UP + cidx[q]] = row[xsrc[q]];
    }
  }
  __syncthreads();
  {
    // boxFilter's border: the covariance terms of the reflected PIXEL, i.e. the gradient taken at
    // the reflected position (its own 3x3 neighbourhood lies in the tile for every output that
    // is stored; clamped for the others)
    int sc[2], cidx[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      cidx[q] = threadIdx.x + q * kRespTW;
      sc[q] = min(max(reflect_any(tx0 - a0 + cidx[q], w) - ux0, 1), UW - 2);
    }
    for (int r = 0; r < GH; ++r) {
      const int sr = min(max(reflect_any(ty0 - a0 + r, h) - uy0, 1), UH - 2);
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        if (cidx[q] >= GW) continue;
        const unsigned char* u = s_u8 + sr * UP + sc[q];
        const int p00 = u[-UP - 1], p01 = u[-UP], p02 = u[-UP + 1];
        const int p10 = u[-1], p12 = u[1];
        const int p20 = u[UP - 1], p21 = u[UP], p22 = u[UP + 1];
        const int dx = (p02 - p00) + 2 * (p12 - p10) + (p22 - p20);
        const int dy = (p20 - p00) + 2 * (p21 - p01) + (p22 - p02);
        s_tile[r * GW + cidx[q]] = (dx & 0xffff) | (dy << 16);
      }
    }
  }
  __syncthreads();
  const double s = 1.0 / (4.0 * (double)b * 255.0);
  const double k_eig = 0.5 * s * s, k_har = (s * s) * (s * s);
  const int x = tx0 + threadIdx.x;
  float best = -INFINITY;
  if (BS > 0) {
    int rA[BS > 0 ? BS : 1], rB[BS > 0 ? BS : 1], rC[BS > 0 ? BS : 1];
#pragma unroll
    for (int j = 0; j < BS; ++j) rA[j] = rB[j] = rC[j] = 0;
    int sA = 0, sB = 0, sC = 0;  // 5x5 of 1020^2 = 2.6e7: fits 32 bits up to BS = 45
    for (int r0 = 0; r0 < GH; r0 += BS) {
#pragma unroll
      for (int j = 0; j < BS; ++j) {
        const int r = r0 + j;
        if (r < GH) {
          int hA = 0, hB = 0, hC = 0;
#pragma unroll
          for (int i = 0; i < BS; ++i) {
            const int p = s_tile[r * GW + threadIdx.x + i];
            const int dx = (short)(p & 0xffff), dy = p >> 16;
            hA += dx * dx; hB += dx * dy; hC += dy * dy;
          }
          sA += hA - rA[j]; sB += hB - rB[j]; sC += hC - rC[j];
          rA[j] = hA; rB[j] = hB; rC[j] = hC;
          const int y = ty0 + r - (BS - 1);
          if (r >= BS - 1 && y < h && x < w) {
            const float e = response_of_i32(sA, sB, sC, harris, k, k_eig, k_har);
            resp[(size_t)v * rplane + (size_t)y * rpitch + x] = e;
            best = fmaxf(best, e);
          }
        }
      }
    }
  } else {
    for (int ry = 0; ry < kRespTH; ++ry) {
      const int y = ty0 + ry;
      if (y >= h || x >= w) break;
      long long A = 0, B = 0, C = 0;
      for (int j = 0; j < b; ++j)
        for (int i = 0; i < b; ++i) {
          const int p = s_tile[(ry + j) * GW + threadIdx.x + i];
          const long long dx = (short)(p & 0xffff), dy = p >> 16;
          A += dx * dx; B += dx * dy; C += dy * dy;
        }
      const float e = response_of(A, B, C, harris, k, k_eig, k_har);
      resp[(size_t)v * rplane + (size_t)y * rpitch + x] = e;
      best = fmaxf(best, e);
    }
  }
  // per-view maximum (minMaxLoc over the whole map, featureselect.cpp)
  unsigned o = best == -INFINITY ? 0u : ord_of(best);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) o = max(o, __shfl_xor_sync(0xffffffffu, o, d));
  if ((threadIdx.x & 31) == 0 && o) atomicMax(vmax + v, o);
}

// -------------------------------------------------------------- candidates

// threshold(eig, max*quality, THRESH_TOZERO); dilate 3x3; corners where val != 0 && val == dilated,
// rows 1..h-2, cols 1..w-2 (featureselect.cpp). key = ordered(value) << 32 | (y*w + x).
// A thread walks kCandRows rows of one column with a rolling 3-row window; a block appends its
// candidates with ONE global atomic (a counter per view would otherwise serialise ~30k atomics).
constexpr int kCandRows = 16;

__global__ void __launch_bounds__(128)
k_seed_candidates(const float* __restrict__ resp, int rpitch, size_t rplane, int w, int h,
                  const unsigned* __restrict__ vmax, double quality,
                  unsigned long long* __restrict__ keys, size_t kplane, int cap,
                  int* __restrict__ count) {
  __shared__ unsigned long long s_k[128 * 4];
  __shared__ int s_n, s_base;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y0 = blockIdx.y * kCandRows, v = blockIdx.z;
  const unsigned om = vmax[v];
  if (!om) return;
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  const float thr = (float)((double)float_of(om) * quality);
  const float* r = resp + (size_t)v * rplane;
  const bool col_ok = x >= 1 && x <= w - 2;
  float m0 = 0.0f, m1 = 0.0f, c1 = 0.0f;  // row maxima of rows y-2, y-1 and the centre of row y-1
  const int y_end = min(y0 + kCandRows, h - 1);  // last row examined is y_end - 1 <= h-2
  if (col_ok)
    for (int y = max(y0, 1) - 1; y <= y_end; ++y) {
      const float* row = r + (size_t)y * rpitch + x;
      float a = row[-1], b = row[0], c = row[1];
      a = a > thr ? a : 0.0f; b = b > thr ? b : 0.0f; c = c > thr ? c : 0.0f;
      const float m2 = fmaxf(fmaxf(a, b), c);
      const int yc = y - 1;  // the row whose 3x3 window is now complete
      if (yc >= max(y0, 1) && c1 != 0.0f && c1 == fmaxf(fmaxf(m0, m1), m2)) {
        const int slot = atomicAdd(&s_n, 1);
        const unsigned long long key = ((unsigned long long)ord_of(c1) << 32) | (unsigned)(yc * w + x);
        if (slot < 128 * 4) s_k[slot] = key;
        else {  // more than a third of the tile are maxima: plateaus; append one by one
          const int g = atomicAdd(count + v, 1);
          if (g < cap) keys[(size_t)v * kplane + g] = key;
        }
      }
      m0 = m1; m1 = m2; c1 = b;
    }
  __syncthreads();
  const int n = min(s_n, 128 * 4);
  if (n == 0) return;
  if (threadIdx.x == 0) s_base = atomicAdd(count + v, n);
  __syncthreads();
  const int base = s_base;
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    if (base + i < cap) keys[(size_t)v * kplane + base + i] = s_k[i];
}

// ------------------------------------------------------------------ select

constexpr int kSelThreads = 1024;
constexpr int kSelSmemKeys = 8192;

__device__ __forceinline__ void bitonic_desc(unsigned long long* a, int n) {
  for (int k = 2; k <= n; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int l = i ^ j;
        if (l > i) {
          const unsigned long long p = a[i], q = a[l];
          const bool first_half = (i & k) == 0;
          if ((p < q) == first_half) { a[i] = q; a[l] = p; }
        }
      }
      __syncthreads();
    }
}

// The greedy pass of goodFeaturesToTrack over candidates a[0..n) sorted in selection order:
// a candidate is kept unless a kept one lies closer than min_distance; stops at max_features.
// Kept corners are appended to s_acc / *s_nacc (shared).
__device__ __forceinline__ void greedy_select(const unsigned long long* a, int n, int w,
                                              int max_features, int min_distance, int2* s_acc,
                                              int* s_nacc, int* s_first) {
  const int md2 = min_distance * min_distance;
  for (int base = 0; base < n; base += blockDim.x) {
    if (*s_nacc >= max_features) break;
    const int i = base + threadIdx.x;
    bool alive = i < n;
    int x = 0, y = 0;
    if (alive) {
      const unsigned ofs = (unsigned)(a[i] & 0xffffffffu);
      y = ofs / w; x = ofs - y * w;
      if (min_distance >= 1) {
        const int na = *s_nacc;
        for (int j = 0; j < na; ++j) {
          const int dx = x - s_acc[j].x, dy = y - s_acc[j].y;
          if (dx * dx + dy * dy < md2) { alive = false; break; }
        }
      }
    }
    // the survivors of this batch, in order
    while (true) {
      __syncthreads();
      if (threadIdx.x == 0) *s_first = INT_MAX;
      __syncthreads();
      if (alive) atomicMin(s_first, (int)threadIdx.x);
      __syncthreads();
      const int f = *s_first;
      if (f == INT_MAX) break;
      if ((int)threadIdx.x == f) {
        s_acc[*s_nacc] = make_int2(x, y);
        *s_nacc = *s_nacc + 1;
        alive = false;
      }
      __syncthreads();
      const int na = *s_nacc;
      if (na >= max_features) break;
      if (alive && min_distance >= 1) {
        const int dx = x - s_acc[na - 1].x, dy = y - s_acc[na - 1].y;
        if (dx * dx + dy * dy < md2) alive = false;
      }
    }
    __syncthreads();
  }
  __syncthreads();
}

constexpr int kSelBins = 4096;       // histogram over the top 12 bits of the ordered value
constexpr int kSelHeadTarget = 2048; // candidates wanted in the head (strongest) group

// One block per view. Sorts the view's candidates (descending value; equal values by
// descending address: greaterThanPtr of featureselect.cpp) and runs the greedy selection.
// Views with more candidates than the shared-memory sort holds first try the strongest ones
// alone: a 4096-bin histogram of the values picks the bins that hold about kSelHeadTarget
// candidates; every candidate in them precedes every other one in selection order, so when the
// greedy pass fills max_features from that head the rest never matters. Otherwise the whole
// list is sorted in global memory and the pass restarts.
__global__ void __launch_bounds__(kSelThreads)
k_seed_select(unsigned long long* __restrict__ keys, size_t kplane, int cap,
              const int* __restrict__ count, int w, int max_features, int min_distance,
              int2* __restrict__ kps, int* __restrict__ nkp, int* __restrict__ status) {
  extern __shared__ unsigned long long s_keys[];
  __shared__ int s_nacc, s_first, s_cut, s_head;
  __shared__ int2 s_acc[kMaxSeedFeatures];
  const int v = blockIdx.x;
  int n = count[v];
  if (n > cap) {  // cannot happen without large plateaus of equal responses
    n = cap;
    if (threadIdx.x == 0) atomicOr(status, 1);
  }
  unsigned long long* gk = keys + (size_t)v * kplane;
  if (threadIdx.x == 0) s_nacc = 0;
  __syncthreads();
  bool done = false;
  if (n > kSelSmemKeys) {
    int* hist = reinterpret_cast<int*>(s_keys + kSelSmemKeys) - kSelBins;  // tail of the key buffer
    for (int i = threadIdx.x; i < kSelBins; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) atomicAdd(&hist[(unsigned)(gk[i] >> 52)], 1);
    __syncthreads();
    if (threadIdx.x == 0) {  // highest bins first until the head holds enough candidates
      int acc = 0, b = kSelBins - 1;
      for (; b >= 0; --b) {
        acc += hist[b];
        if (acc >= kSelHeadTarget) break;
      }
      s_cut = max(b, 0);
      s_head = 0;
    }
    __syncthreads();
    const unsigned cut = (unsigned)s_cut;
    const int room = kSelSmemKeys - kSelBins * (int)sizeof(int) / (int)sizeof(unsigned long long);
    __syncthreads();  // hist is dead from here: the head is gathered into the same buffer
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const unsigned long long k = gk[i];
      if ((unsigned)(k >> 52) >= cut) {
        const int slot = atomicAdd(&s_head, 1);
        if (slot < room) s_keys[slot] = k;
      }
    }
    __syncthreads();
    const int nh = s_head;
    if (nh <= room && nh < n) {
      int npad = 1;
      while (npad < nh) npad <<= 1;
      for (int i = nh + threadIdx.x; i < npad; i += blockDim.x) s_keys[i] = 0ull;
      __syncthreads();
      bitonic_desc(s_keys, npad);
      greedy_select(s_keys, nh, w, max_features, min_distance, s_acc, &s_nacc, &s_first);
      done = s_nacc >= max_features;
      __syncthreads();
      if (!done && threadIdx.x == 0) s_nacc = 0;  // restart over the whole list
      __syncthreads();
    }
  }
  if (!done) {
    int npad = 1;
    while (npad < n) npad <<= 1;
    unsigned long long* a;
    if (npad <= kSelSmemKeys) {
      a = s_keys;
      for (int i = threadIdx.x; i < npad; i += blockDim.x) a[i] = i < n ? gk[i] : 0ull;
    } else {
      a = gk;
      for (int i = n + threadIdx.x; i < npad; i += blockDim.x) a[i] = 0ull;
    }
    __syncthreads();
    bitonic_desc(a, npad);
    greedy_select(a, n, w, max_features, min_distance, s_acc, &s_nacc, &s_first);
  }
  const int na = s_nacc;
  for (int i = threadIdx.x; i < na; i += blockDim.x) kps[(size_t)v * max_features + i] = s_acc[i];
  if (threadIdx.x == 0) nkp[v] = na;
}

// ------------------------------------------------------------------- match

constexpr int kMatchThreads = 160;  // 294 positions of the default stripe: two rounds

// StereoMatcher::MatchRectified for keypoint blockIdx.x of view blockIdx.y. Template and stripe
// are staged as packed bytes; sum (T-I)^2 = sum T^2 - 2 sum T*I + sum I^2 with the two window
// sums as byte dot products (DP4A) over words assembled from the stripe at any byte offset.
__global__ void __launch_bounds__(kMatchThreads)
k_seed_match(SeedImages im, const int2* __restrict__ kps, const int* __restrict__ nkp,
             int max_features, int tc, int tr, int md, double max_cost, float* __restrict__ kpd) {
  extern __shared__ unsigned s_w[];  // template [tr][tcw] words (zero padded), stripe [tr+2][mdw] words
  __shared__ unsigned long long s_red[kMatchThreads / 32];
  __shared__ unsigned s_t2[kMatchThreads / 32];
  const int k = blockIdx.x, v = blockIdx.y;
  if (k >= nkp[v]) return;
  const int2 kp = kps[(size_t)v * max_features + k];
  float* out = kpd + (size_t)v * max_features + k;
  const int w = im.w, h = im.h;
  const int stripe_rows = tr + 2;
  const int ty = kp.y - (tr - 1) / 2;
  const int sy = kp.y - (stripe_rows - 1) / 2;
  if (ty < 0 || ty + tr >= h || sy < 0 || sy + stripe_rows >= h) {  // stereo_matcher.cpp:35-37, 65-67
    if (threadIdx.x == 0) *out = -1.0f;
    return;
  }
  int offset_x = 0;
  int tx = kp.x - (tc - 1) / 2;
  if (tx < 0) { offset_x = tx; tx = 0; }
  if (tx + tc >= w) { offset_x = (tx + tc) - (w - 1); tx -= offset_x; }
  int sx = kp.x + (tc - 1) / 2 - md;
  if (sx + md > w - 1) sx -= (sx + md) - (w - 1);
  if (sx < 0) sx = 0;
  const int tcw = (tc + 3) >> 2, mdw = ((md + 3) >> 2) + 1;  // one spare word for the funnel shift
  unsigned* Tw = s_w;
  unsigned* Sw = s_w + tr * tcw;
  unsigned char* T = reinterpret_cast<unsigned char*>(Tw);
  unsigned char* S = reinterpret_cast<unsigned char*>(Sw);
  const uint8_t* ref = ref_image(im, v);
  const uint8_t* mat = mat_image(im, v);
  for (int i = threadIdx.x; i < tr * tcw + stripe_rows * mdw; i += blockDim.x) s_w[i] = 0u;
  __syncthreads();
  unsigned t2 = 0;
  for (int i = threadIdx.x; i < tr * tc; i += blockDim.x) {
    const int r = i / tc, c = i - r * tc;
    const unsigned p = px(ref, im, v, tx + c, ty + r);
    T[r * tcw * 4 + c] = (unsigned char)p;
    t2 += p * p;
  }
  for (int i = threadIdx.x; i < stripe_rows * md; i += blockDim.x) {
    const int r = i / md, c = i - r * md;
    S[r * mdw * 4 + c] = (unsigned char)px(mat, im, v, sx + c, sy + r);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) t2 += __shfl_xor_sync(0xffffffffu, t2, d);
  if ((threadIdx.x & 31) == 0) s_t2[threadIdx.x >> 5] = t2;
  __syncthreads();
  t2 = 0;
  for (int i = 0; i < kMatchThreads / 32; ++i) t2 += s_t2[i];
  const double tnorm = sqrt((double)t2);
  const int nx = md - tc + 1, npos = nx * (stripe_rows - tr + 1);
  const unsigned last_mask = (tc & 3) ? ((1u << (8 * (tc & 3))) - 1u) : 0xffffffffu;
  unsigned long long best = ~0ull;
  for (int p = threadIdx.x; p < npos; p += blockDim.x) {
    const int pj = p / nx, pi = p - pj * nx;
    const int sh = 8 * (pi & 3);
    unsigned ab = 0, w2 = 0;
    for (int j = 0; j < tr; ++j) {
      const unsigned* a = Tw + j * tcw;
      const unsigned* bw = Sw + (pj + j) * mdw + (pi >> 2);
      unsigned lo = bw[0];
      for (int i = 0; i < tcw; ++i) {
        const unsigned hi = bw[i + 1];
        unsigned bb = __funnelshift_r(lo, hi, sh);  // stripe bytes pi+4i .. pi+4i+3
        if (i == tcw - 1) bb &= last_mask;
        ab = __dp4a(a[i], bb, ab);
        w2 = __dp4a(bb, bb, w2);
        lo = hi;
      }
    }
    const unsigned long long ssd = (unsigned long long)t2 + w2 - 2ull * ab;
    // common_matchTemplate (templmatch.cpp), TM_SQDIFF_NORMED
    double num = (double)ssd;
    const double t = sqrt((double)w2) * tnorm;
    if (fabs(num) < t) num = num / t;
    else if (fabs(num) < t * 1.125) num = num > 0 ? 1.0 : -1.0;
    else num = 1.0;
    const float r = (float)num;  // >= 0: its bits order like unsigned integers
    const unsigned long long key = ((unsigned long long)__float_as_uint(r) << 32) | (unsigned)p;
    best = key < best ? key : best;  // smallest value, then first in raster order (minMaxLoc)
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    const unsigned long long o = __shfl_xor_sync(0xffffffffu, best, d);
    best = o < best ? o : best;
  }
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < kMatchThreads / 32; ++i) best = s_red[i] < best ? s_red[i] : best;
    const float minval = __uint_as_float((unsigned)(best >> 32));
    const int p = (int)(best & 0xffffffffu);
    const int pi = p % nx;
    const int mx = pi + sx + (tc - 1) / 2 + offset_x;
    // has_good_matching_score && match_is_to_the_left (stereo_matcher.cpp:107-115)
    *out = ((double)minval < max_cost && kp.x >= mx) ? __fsub_rn((float)kp.x, (float)mx) : -1.0f;
  }
}

// ------------------------------------------------------------------- paint

constexpr int kPaintTW = 128, kPaintTH = 32;

// Seed map of every view: zero map, keypoint disparities >= 0 written at the keypoints, dilated
// with a (2r+1)^2 rectangle (patchmatch_gpu.cu:423-439). With ow x oh != w x h the dilated map
// is sampled like cv::resize(INTER_NEAREST) and divided by `div` (Patchmatch::Initialize,
// patchmatch.cpp:75-81). View 2p goes to out_l[p], view 2p+1 to out_r[p] flipped back to
// right-image coordinates.
__global__ void __launch_bounds__(256)
k_seed_paint(const int2* __restrict__ kps, const float* __restrict__ kpd,
             const int* __restrict__ nkp, int max_features, int w, int h, int r, int ow, int oh,
             double fx, double fy, float div, float* __restrict__ out_l,
             float* __restrict__ out_r, size_t opitch, size_t oplane) {
  __shared__ int s_n;
  __shared__ int s_x[kMaxSeedFeatures], s_y[kMaxSeedFeatures];
  __shared__ float s_d[kMaxSeedFeatures];
  const int v = blockIdx.z;
  const int x0 = blockIdx.x * kPaintTW, y0 = blockIdx.y * kPaintTH;
  const int x1 = min(x0 + kPaintTW, ow) - 1, y1 = min(y0 + kPaintTH, oh) - 1;
  // source range of this tile of the output
  const int sx0 = min((int)floor(x0 * fx), w - 1), sx1 = min((int)floor(x1 * fx), w - 1);
  const int sy0 = min((int)floor(y0 * fy), h - 1), sy1 = min((int)floor(y1 * fy), h - 1);
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  const int n = nkp[v];
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int2 kp = kps[(size_t)v * max_features + i];
    const float d = kpd[(size_t)v * max_features + i];
    if (d >= 0.0f && kp.x >= sx0 - r && kp.x <= sx1 + r && kp.y >= sy0 - r && kp.y <= sy1 + r) {
      const int s = atomicAdd(&s_n, 1);
      s_x[s] = kp.x; s_y[s] = kp.y; s_d[s] = d;
    }
  }
  __syncthreads();
  const int m = s_n;
  const bool same = ow == w && oh == h;
  float* out = ((v & 1) ? out_r : out_l) + (size_t)(v >> 1) * oplane;
  if (same && ow % 4 == 0 && opitch % 4 == 0) {
    // SparseInit: four pixels per thread, one 16-byte store (reversed for the flipped view)
    for (int i = threadIdx.x; i < kPaintTW * kPaintTH / 4; i += blockDim.x) {
      const int x = x0 + 4 * (i % (kPaintTW / 4)), y = y0 + i / (kPaintTW / 4);
      if (x >= ow || y >= oh) continue;
      float4 best = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int j = 0; j < m; ++j) {
        if (abs(y - s_y[j]) > r) continue;
        const int dx = x - s_x[j];
        const float d = s_d[j];
        if (abs(dx) <= r) best.x = fmaxf(best.x, d);
        if (abs(dx + 1) <= r) best.y = fmaxf(best.y, d);
        if (abs(dx + 2) <= r) best.z = fmaxf(best.z, d);
        if (abs(dx + 3) <= r) best.w = fmaxf(best.w, d);
      }
      if (div != 1.0f) {
        best.x = __fdiv_rn(best.x, div); best.y = __fdiv_rn(best.y, div);
        best.z = __fdiv_rn(best.z, div); best.w = __fdiv_rn(best.w, div);
      }
      if (v & 1)
        *reinterpret_cast<float4*>(out + (size_t)y * opitch + (ow - 4 - x)) =
            make_float4(best.w, best.z, best.y, best.x);
      else
        *reinterpret_cast<float4*>(out + (size_t)y * opitch + x) = best;
    }
    return;
  }
  for (int i = threadIdx.x; i < kPaintTW * kPaintTH; i += blockDim.x) {
    const int x = x0 + (i % kPaintTW), y = y0 + (i / kPaintTW);
    if (x >= ow || y >= oh) continue;
    // no resize (SparseInit): the source pixel is the output pixel
    const int sx = same ? x : min((int)floor(x * fx), w - 1);
    const int sy = same ? y : min((int)floor(y * fy), h - 1);
    float best = 0.0f;
    for (int j = 0; j < m; ++j)
      if (abs(sx - s_x[j]) <= r && abs(sy - s_y[j]) <= r) best = fmaxf(best, s_d[j]);
    out[(size_t)y * opitch + ((v & 1) ? ow - 1 - x : x)] = __fdiv_rn(best, div);
  }
}

// --------------------------------------------------------------- launchers

int launch_seed_detect(const SeedImages& im, int nviews, const SeedDetect& sp, float* resp, int pitch, size_t plane, unsigned long long* keys,
                       size_t kplane, int cap, SeedState s, cudaStream_t st) {
  if (cudaMemsetAsync(s.vmax, 0, sizeof(unsigned) * nviews, st) != cudaSuccess) return -1;
  if (cudaMemsetAsync(s.ncand, 0, sizeof(int) * nviews, st) != cudaSuccess) return -1;
  const int b = sp.block_size;
  dim3 g2(cdiv(im.w, kRespTW), cdiv(im.h, kRespTH), nviews);
  const size_t smem = (size_t)(kRespTW + b - 1) * (kRespTH + b - 1) * sizeof(int) +
                      (size_t)((kRespTW + b + 1 + 3) & ~3) * (kRespTH + b + 1);
#define PM_RESP(BS)                                                                         \
  k_seed_response<BS><<<g2, kRespTW, smem, st>>>(im, b, sp.use_harris, sp.harris_k, resp, pitch, \
                                                 plane, s.vmax)
  if (b == 3) PM_RESP(3);
  else if (b == 5) PM_RESP(5);
  else if (b == 7) PM_RESP(7);
  else PM_RESP(0);
#undef PM_RESP
  dim3 g3(cdiv(im.w, 128), cdiv(im.h, kCandRows), nviews);
  k_seed_candidates<<<g3, 128, 0, st>>>(resp, pitch, plane, im.w, im.h, s.vmax, sp.quality_level,
                                        keys, kplane, cap, s.ncand);
  static bool attr_set[64] = {false};  // per device of this process
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev & 63]) {
    if (cudaFuncSetAttribute(k_seed_select, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             kSelSmemKeys * (int)sizeof(unsigned long long)) != cudaSuccess) return -1;
    attr_set[dev & 63] = true;
  }
  k_seed_select<<<nviews, kSelThreads, kSelSmemKeys * sizeof(unsigned long long), st>>>(
      keys, kplane, cap, s.ncand, im.w, sp.max_features, sp.min_distance, s.kps, s.nkp, s.status);
  return PM_LAUNCH_CHECK(3);
}

int launch_seed_match(const SeedImages& im, int nviews, const SeedMatch& mp, int max_features,
                      SeedState s, cudaStream_t st) {
  dim3 grid(max_features, nviews);
  const size_t smem = 4 * ((size_t)mp.templ_rows * ((mp.templ_cols + 3) / 4) +
                           (size_t)(mp.templ_rows + 2) * ((mp.max_disp + 3) / 4 + 1));
  k_seed_match<<<grid, kMatchThreads, smem, st>>>(im, s.kps, s.nkp, max_features, mp.templ_cols,
                                                  mp.templ_rows, mp.max_disp, mp.max_matching_cost,
                                                  s.kpd);
  return PM_LAUNCH_CHECK(1);
}

int launch_seed_paint(int nviews, int w, int h, int radius, int ow, int oh, float div,
                      int max_features, SeedState s, float* out_l, float* out_r, size_t opitch,
                      size_t oplane, cudaStream_t st) {
  dim3 grid(cdiv(ow, kPaintTW), cdiv(oh, kPaintTH), nviews);
  k_seed_paint<<<grid, 256, 0, st>>>(s.kps, s.kpd, s.nkp, max_features, w, h, radius, ow, oh,
                                     (double)w / ow, (double)h / oh, div, out_l, out_r, opitch, oplane);
  return PM_LAUNCH_CHECK(1);
}

}  // namespace pm
