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
