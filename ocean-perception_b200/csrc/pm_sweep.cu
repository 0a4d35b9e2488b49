// pm_sweep.cu -- spatial propagation sweeps (PropagateRow / PropagateCol,
// patchmatch_gpu.cu:116-230) for sm_100a.
//
// Schedule. The reference cuts every line (row or column) into `chunks` chunks
// that overlap their neighbours by `overlap` pixels on both sides and lets one
// thread walk each chunk while all of them read and write the same disparity
// plane. On the hardware it was written for the chunks of a line advance in
// lock step (one warp), which fixes the outcome of that race: when chunk k
// reaches the head of chunk k+1 (in walking order) it finds the values chunk k+1
// left there during its own first steps, and what chunk k writes there is final.
// Everything else a chunk reads is still the pre-sweep plane. The kernels below
// compute exactly that, with every chain independent of every other thread:
//   1. replay the first n_ov steps of the next chunk on the pre-sweep plane,
//   2. walk the own chunk, taking {d, cost} from (1) inside the overlap,
//   3. write only the positions no earlier chunk will overwrite.
// dc_in is never written, dc_out receives the final values (the caller copies the
// plane first so untouched pixels carry over).
//
// The cost of the current disparity is read from the {d, cost} plane instead of
// being recomputed at every step (patchmatch_gpu.cu:161-162 recomputes it): the
// cost is a pure function of (pixel, d), so the comparison is identical.
#include "pm_kernels.h"

namespace pm {

constexpr int kMaxOverlap2 = 16;  // 2 * overlap upper bound

template <bool ALONG_X>
struct Walk {
  int line, pitch;
  __device__ __forceinline__ size_t idx(int pos) const {
    return ALONG_X ? (size_t)line * pitch + pos : (size_t)pos * pitch + line;
  }
  __device__ __forceinline__ int x(int pos) const { return ALONG_X ? pos : line; }
  __device__ __forceinline__ int y(int pos) const { return ALONG_X ? line : pos; }
};

// One propagation step at position pos with candidate `cand` coming from the
// previous position (patchmatch_gpu.cu:158-170).
template <bool ALONG_X>
__device__ __forceinline__ float2 step(const Walk<ALONG_X>& wk, const float2* __restrict__ ref,
                                       const float2* __restrict__ mat, int pos, float2 cur,
                                       float cand, float alpha, float w1) {
  const int x = wk.x(pos), y = wk.y(pos);
  const RefTaps L = load_ref_taps(ref, wk.pitch, y, x);
  const float c1 = cost5(L, mat, wk.pitch, y, xr_of(x, cand), alpha, w1);
  if (c1 < cur.y) {
    cur.x = fminf(cand, __int2float_rn(x - 1));
    cur.y = c1;
  }
  return cur;
}

template <bool ALONG_X>
__global__ void __launch_bounds__(128)
k_sweep_generic(const float2* __restrict__ ref, const float2* __restrict__ mat,
                const float2* __restrict__ dc_in, float2* __restrict__ dc_out, ViewGeom g,
                int nviews, int dir, int chunks, int ov, float alpha, float w1) {
  const int nlines = ALONG_X ? g.h : g.w, len = ALONG_X ? g.w : g.h;
  const int nl = nlines - 2;
  const long tid = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= (long)nl * chunks * nviews) return;
  const int line = 1 + (int)(tid % nl);
  const int k = (int)((tid / nl) % chunks);
  const int v = (int)(tid / ((long)nl * chunks));
  const size_t vo = (size_t)v * g.plane;
  ref += vo; mat += vo; dc_in += vo; dc_out += vo;
  const int cs = len / chunks;
  Walk<ALONG_X> wk{line, g.pitch};

  int start, stop;
  chunk_range(k, cs, ov, len, dir, start, stop);
  const int nsteps = dir > 0 ? stop - start : start - stop;
  if (nsteps <= 0) return;

  // next chunk in walking order: its head overlaps my tail
  int n_ov = 0, start_n = 0;
  const int kn = k + dir, kp = k - dir;
  if (kn >= 0 && kn < chunks) {
    int sn, en;
    chunk_range(kn, cs, ov, len, dir, sn, en);
    const int nn = dir > 0 ? en - sn : sn - en;
    if (nn > 0) {
      start_n = sn;
      n_ov = dir > 0 ? stop - sn : sn - stop;
      n_ov = max(0, min(n_ov, min(nn, kMaxOverlap2)));
    }
  }
  // previous chunk in walking order: its tail overwrites my head
  int n_head = 0;
  if (kp >= 0 && kp < chunks) {
    int sp, ep;
    chunk_range(kp, cs, ov, len, dir, sp, ep);
    const int np = dir > 0 ? ep - sp : sp - ep;
    if (np > 0) n_head = max(0, dir > 0 ? ep - start : start - ep);
  }

  float hd[kMaxOverlap2], hc[kMaxOverlap2];
  if (n_ov > 0) {
    float prev = dc_in[wk.idx(start_n - dir)].x;
    for (int j = 0; j < n_ov; ++j) {
      const int pos = start_n + dir * j;
      const float2 o = step(wk, ref, mat, pos, dc_in[wk.idx(pos)], prev, alpha, w1);
      hd[j] = o.x;
      hc[j] = o.y;
      prev = o.x;
    }
  }

  float prev = dc_in[wk.idx(start - dir)].x;
  const int first_ov = nsteps - n_ov;
  for (int i = 0; i < nsteps; ++i) {
    const int pos = start + dir * i;
    float2 cur;
    if (i >= first_ov) {
      cur.x = hd[i - first_ov];
      cur.y = hc[i - first_ov];
    } else {
      cur = dc_in[wk.idx(pos)];
    }
    cur = step(wk, ref, mat, pos, cur, prev, alpha, w1);
    prev = cur.x;
    if (i >= n_head) dc_out[wk.idx(pos)] = cur;
  }
}


// ---------------------------------------------------------------------------
// Row sweep with the matched image staged in shared memory.
//
// A row sweep walks along x, the axis the disparity shifts along, so the 32
// lanes of a warp (32 different rows or chunks) gather from 32 different lines
// of the matched image: through L1 that costs one wavefront per lane per load.
// Shared memory serves such gathers at bank granularity instead. One block owns
// kRows consecutive rows and all chunks of those rows:
//   * the kRows+2 matched-image rows are copied into shared memory once (full
//     width, so any disparity up to x-1 stays inside, as the reference allows);
//   * a half-warp holds the 16 rows of one chunk, so the pre-sweep {d, cost} plane
//     and the reference taps are read from TRANSPOSED planes (rows contiguous)
//     with 128-byte coalesced loads;
//   * results leave through a 16x16 shared tile per half-warp and are written
//     row-major with 128-byte stores. Every pixel of the block's rows is written
//     exactly once (final writer, or a copy where no chunk visits), so the output
//     plane needs no pre-copy.
// The arithmetic and the schedule are those of k_sweep_generic.

constexpr int kRows = 16;
constexpr int kTilePitch = 17;

struct RowEmitter {
  float2* tile;        // [16][kTilePitch] of this half-warp
  float2* out;         // row-major output plane of the view
  int pitch, y0, h, r; // r = row within the block = lane within the half-warp
  int own_lo, own_hi;  // positions this chunk is the final writer (or copier) of
  int edge;            // 15 walking up, 0 walking down
  unsigned mask;

  __device__ __forceinline__ void flush(int p) {
    __syncwarp(mask);
    const int pos = (p & ~15) + r;
    if (pos >= own_lo && pos < own_hi) {
#pragma unroll
      for (int rr = 0; rr < 16; ++rr)
        if (y0 + rr < h) out[(size_t)(y0 + rr) * pitch + pos] = tile[rr * kTilePitch + r];
    }
    __syncwarp(mask);
  }
  // all lanes of the half-warp call emit with the same p
  __device__ __forceinline__ void emit(int p, float2 v, bool last) {
    tile[r * kTilePitch + (p & 15)] = v;
    if ((p & 15) == edge || last) flush(p);
  }
};

__global__ void __launch_bounds__(512)
k_sweep_row_smem(const float2* __restrict__ refT, const float2* __restrict__ mat,
                 const float2* __restrict__ dcT_in, float2* __restrict__ dc_out, ViewGeom g,
                 int pitchT, size_t planeT, int dir, int chunks, int ov, float alpha, float w1) {
  extern __shared__ float2 smem[];
  const int w = g.w, h = g.h;
  const int spitch = w + 1;  // odd pitch in 8-byte words: rows spread over all banks
  float2* smat = smem;
  float2* tiles = smem + (size_t)(kRows + 2) * spitch;
  const int t = threadIdx.x, r = t & 15, k = t >> 4;
  const int y0 = blockIdx.x * kRows, v = blockIdx.y;
  refT += (size_t)v * planeT;
  dcT_in += (size_t)v * planeT;
  mat += (size_t)v * g.plane;
  dc_out += (size_t)v * g.plane;

  // stage the matched rows y0-1 .. y0+kRows (incl. the finite pad element at column w)
  for (int row = 0; row < kRows + 2; ++row) {
    const int gy = min(max(y0 - 1 + row, 0), h - 1);
    const float2* src = mat + (size_t)gy * g.pitch;
    float2* dst = smat + row * spitch;
    for (int c = t; c < spitch; c += blockDim.x) dst[c] = src[c];
  }
  __syncthreads();

  const int y = y0 + r;
  const bool valid = y < h;
  const int yc = valid ? y : h - 1;              // idle lanes mirror the last row (no stores)
  const bool active = valid && y >= 1 && y <= h - 2;
  const int len = w, cs = len / chunks;

  int start, stop;
  chunk_range(k, cs, ov, len, dir, start, stop);
  const int nsteps = dir > 0 ? stop - start : start - stop;
  const int kn = k + dir, kp = k - dir;
  const bool has_n = kn >= 0 && kn < chunks, has_p = kp >= 0 && kp < chunks;
  int n_ov = 0, start_n = 0, n_head = 0, stop_p = 0;
  if (has_n) {
    int sn, en;
    chunk_range(kn, cs, ov, len, dir, sn, en);
    start_n = sn;
    n_ov = max(0, min(dir > 0 ? stop - sn : sn - stop, kMaxOverlap2));
  }
  if (has_p) {
    int sp, ep;
    chunk_range(kp, cs, ov, len, dir, sp, ep);
    stop_p = ep;
    n_head = max(0, dir > 0 ? ep - start : start - ep);
  }

  RowEmitter em;
  em.tile = tiles + (size_t)k * 16 * kTilePitch;
  em.out = dc_out; em.pitch = g.pitch; em.y0 = y0; em.h = h; em.r = r;
  em.edge = dir > 0 ? 15 : 0;
  em.mask = 0xFFFFu << (threadIdx.x & 16);
  if (dir > 0) {
    em.own_lo = has_p ? stop_p : 0;
    em.own_hi = has_n ? stop : len;
  } else {
    em.own_lo = has_n ? stop + 1 : 0;
    em.own_hi = has_p ? stop_p + 1 : len;
  }
  // walking order over everything this chunk emits: [first, last]
  const int first = dir > 0 ? em.own_lo : em.own_hi - 1;
  const int last = dir > 0 ? em.own_hi - 1 : em.own_lo;

  const float2* in_col = dcT_in + yc;  // + p * pitchT
  const float2* srow = smat + (size_t)(r + 1) * spitch;  // matched row y

  // 1. positions before the first visited one (first chunk in walking order only)
  const int first_visit = start + dir * n_head;  // first position this chunk finally writes
  for (int p = first; p != first_visit && (dir > 0 ? p < first_visit : p > first_visit); p += dir)
    em.emit(p, in_col[(size_t)p * pitchT], p == last);

  // 2. replay the head of the next chunk on the pre-sweep plane
  float hd[kMaxOverlap2], hc[kMaxOverlap2];
  if (active && n_ov > 0) {
    float prev = in_col[(size_t)(start_n - dir) * pitchT].x;
    for (int j = 0; j < n_ov; ++j) {
      const int p = start_n + dir * j;
      float2 cur = in_col[(size_t)p * pitchT];
      const float2* rt = refT + (size_t)p * pitchT + y;
      RefTaps L;
      L.tl = rt[-pitchT - 1]; L.bl = rt[-pitchT + 1];
      L.c = rt[0];
      L.tr = rt[pitchT - 1]; L.br = rt[pitchT + 1];
      const float c1 = cost5(L, srow, spitch, 0, xr_of(p, prev), alpha, w1);
      if (c1 < cur.y) { cur.x = fminf(prev, __int2float_rn(p - 1)); cur.y = c1; }
      hd[j] = cur.x; hc[j] = cur.y; prev = cur.x;
    }
  }

  // 3. the chunk itself
  {
    float prev = in_col[(size_t)(start - dir) * pitchT].x;
    const int first_ov = nsteps - n_ov;
    for (int i = 0; i < nsteps; ++i) {
      const int p = start + dir * i;
      float2 cur;
      if (i >= first_ov && active) { cur.x = hd[i - first_ov]; cur.y = hc[i - first_ov]; }
      else cur = in_col[(size_t)p * pitchT];
      if (active) {
        const float2* rt = refT + (size_t)p * pitchT + y;
        RefTaps L;
        L.tl = rt[-pitchT - 1]; L.bl = rt[-pitchT + 1];
        L.c = rt[0];
        L.tr = rt[pitchT - 1]; L.br = rt[pitchT + 1];
        const float c1 = cost5(L, srow, spitch, 0, xr_of(p, prev), alpha, w1);
        if (c1 < cur.y) { cur.x = fminf(prev, __int2float_rn(p - 1)); cur.y = c1; }
        prev = cur.x;
      }
      if (i >= n_head) em.emit(p, cur, p == last);
    }
  }

  // 4. positions after the last visited one (last chunk in walking order only)
  for (int p = stop; dir > 0 ? p <= last : p >= last; p += dir)
    em.emit(p, in_col[(size_t)p * pitchT], p == last);
}

size_t sweep_row_smem_bytes(int w, int chunks) {
  return ((size_t)(kRows + 2) * (w + 1) + (size_t)chunks * 16 * kTilePitch) * sizeof(float2);
}

int launch_sweep_row_smem(const float2* refT, const float2* mat, const float2* dcT_in,
                          float2* dc_out, ViewGeom g, int pitchT, size_t planeT, int nviews,
                          int dir, SweepParams sp, cudaStream_t st) {
  const size_t bytes = sweep_row_smem_bytes(g.w, sp.chunks);
  static size_t configured = 0;
  if (bytes > configured) {
    if (cudaFuncSetAttribute(k_sweep_row_smem, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)bytes) != cudaSuccess) return -1;
    configured = bytes;
  }
  dim3 grid((g.h + kRows - 1) / kRows, nviews);
  k_sweep_row_smem<<<grid, 16 * sp.chunks, bytes, st>>>(refT, mat, dcT_in, dc_out, g, pitchT, planeT,
                                                      dir, sp.chunks, sp.overlap, sp.alpha,
                                                      1 - sp.alpha);
  return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_sweep(const float2* ref, const float2* mat, const float2* dc_in, float2* dc_out,
                 ViewGeom g, int nviews, int along_x, int dir, SweepParams sp, cudaStream_t st) {
  const int nlines = along_x ? g.h : g.w;
  const long chains = (long)(nlines - 2) * sp.chunks * nviews;
  if (chains <= 0) return 0;
  const unsigned blocks = (unsigned)((chains + 127) / 128);
  if (along_x)
    k_sweep_generic<true><<<blocks, 128, 0, st>>>(ref, mat, dc_in, dc_out, g, nviews, dir,
                                                  sp.chunks, sp.overlap, sp.alpha, 1 - sp.alpha);
  else
    k_sweep_generic<false><<<blocks, 128, 0, st>>>(ref, mat, dc_in, dc_out, g, nviews, dir,
                                                   sp.chunks, sp.overlap, sp.alpha, 1 - sp.alpha);
  return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace pm
