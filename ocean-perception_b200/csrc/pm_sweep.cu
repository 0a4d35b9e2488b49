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
// Everything else a chunk reads is still the pre-sweep plane.
//
// Two realisations of exactly that schedule live here:
//   * k_sweep_generic: every chain is an independent thread that first replays the
//     head of the next chunk on the pre-sweep plane (any size, any chunking);
//   * k_sweep_row / k_sweep_col: all chunks of a line sit in one block, write every
//     position they walk to the output plane, and after one block barrier a chunk
//     simply reads the head of its successor back from that plane. No replay, no
//     pre-copied output plane, every output pixel written by the block that owns it.
//
// The cost of the current disparity is read from the {d, cost} plane instead of
// being recomputed at every step (patchmatch_gpu.cu:161-162 recomputes it): the
// cost is a pure function of (pixel, d), so the comparison is identical.
#include <climits>
#include <cstdlib>

#include "pm_kernels.h"

namespace pm {

constexpr int kMaxOverlap2 = 16;  // 2 * overlap upper bound
#ifndef PM_COL_L2_PREFETCH
#define PM_COL_L2_PREFETCH 0
#endif
#ifndef PM_COL_CUR_PREFETCH
#define PM_COL_CUR_PREFETCH 2
#endif
#ifndef PM_COL_REF_PREFETCH
#define PM_COL_REF_PREFETCH 0
#endif
constexpr bool kColRefPrefetch = PM_COL_REF_PREFETCH != 0;
constexpr int kColCurPrefetch = PM_COL_CUR_PREFETCH;  // {d, cost} lines pulled into L1 beyond the ring
constexpr int kColL2Prefetch = PM_COL_L2_PREFETCH;  // column kernel: steps ahead prefetched into L2
constexpr int kRowTPrefetch = 12;      // transposed row sweeps: sample columns prefetched ahead
constexpr int kGenericPrefetch = 12;   // generic kernel: walk positions prefetched ahead

static int skip_eq_flag() {
  static const int v = [] { const char* e = getenv("PM_SKIP_EQ"); return e && e[0] == '0' ? 0 : 1; }();
  return v;
}

static bool use_v1() {
  static const int v = [] { const char* e = getenv("PM_SWEEP_V1"); return e && e[0] == '1' ? 1 : 0; }();
  return v != 0;
}


// ============================================================ generic kernel

template <bool ALONG_X>
struct Walk {
  int line, pitch, cost_mode, radius;
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
  const float xr = xr_of_r(x, cand, wk.radius);
  float c1;
  if (wk.cost_mode == 1) {
    c1 = cost_full(ref, mat, wk.pitch, y, x, xr, alpha, w1, wk.radius);
  } else if (wk.cost_mode == 2) {
    c1 = cost_census(ref, mat, wk.pitch, y, x, xr, wk.radius);
  } else {
    const RefTaps L = load_ref_taps(ref, wk.pitch, y, x);
    c1 = cost5(L, mat, wk.pitch, y, xr, alpha, w1);
  }
  if (c1 < cur.y) {
    cur.x = fminf(cand, __int2float_rn(x - wk.radius));
    cur.y = c1;
  }
  return cur;
}

template <bool ALONG_X>
__global__ void __launch_bounds__(128)
k_sweep_generic(const float2* __restrict__ ref, const float2* __restrict__ mat,
                const float2* __restrict__ dc_in, float2* __restrict__ dc_out, ViewGeom g,
                int nviews, int dir, int chunks, int ov, float alpha, float w1, int k_lo, int nk) {
  // Positions along a column are FRAME rows: a band (g.y_off, g.full_h) runs chunks
  // [k_lo, k_lo + nk) of the frame's chunking on its local planes. Rows are never split.
  const int nlines = ALONG_X ? g.h : g.w, len = ALONG_X ? g.w : g.full_h;
  const int rad = g.radius;
  const int nl = nlines - 2 * rad;
  const long tid = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= (long)nl * nk * nviews) return;
  const int line = rad + (int)(tid % nl);
  const int k = k_lo + (int)((tid / nl) % nk);
  const int v = (int)(tid / ((long)nl * nk));
  if (ALONG_X && !row_interior(g, line)) return;
  const size_t vo = (size_t)v * g.plane;
  ref += vo; mat += vo; dc_in += vo; dc_out += vo;
  if (!ALONG_X) {  // frame row -> local row
    const ptrdiff_t shift = -(ptrdiff_t)g.y_off * g.pitch;
    ref += shift; mat += shift; dc_in += shift; dc_out += shift;
  }
  const int cs = len / chunks;
  Walk<ALONG_X> wk{line, g.pitch, g.cost_mode, rad};

  int start, stop;
  chunk_range(k, cs, ov, len, dir, start, stop, rad);
  const int nsteps = dir > 0 ? stop - start : start - stop;
  if (nsteps <= 0) return;

  // next chunk in walking order: its head overlaps my tail
  int n_ov = 0, start_n = 0;
  const int kn = k + dir, kp = k - dir;
  if (kn >= 0 && kn < chunks) {
    int sn, en;
    chunk_range(kn, cs, ov, len, dir, sn, en, rad);
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
    chunk_range(kp, cs, ov, len, dir, sp, ep, rad);
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
    // Chains here are few and long (row bands, odd shapes): each step would otherwise wait for
    // DRAM three times on the dependent chain. Pull the lines of the position kGenericPrefetch
    // steps ahead into L1: {d, cost}, the reference taps and the matched row where the current
    // disparity would sample it.
    if (i + kGenericPrefetch < nsteps) {
      const int pp = pos + dir * kGenericPrefetch;
      const size_t o = wk.idx(pp);
      asm volatile("prefetch.global.L1 [%0];" ::"l"(dc_in + o));
      asm volatile("prefetch.global.L1 [%0];" ::"l"(ref + o));
      const int xs = min(max(__float2int_rd(__fsub_rn(__int2float_rn(wk.x(pp)), prev)), 0), g.w);
      asm volatile("prefetch.global.L1 [%0];" ::"l"(mat + (size_t)wk.y(pp) * g.pitch + xs));
    }
  }
}

int launch_sweep(const float2* ref, const float2* mat, const float2* dc_in, float2* dc_out,
                 ViewGeom g, int nviews, int along_x, int dir, SweepParams sp, cudaStream_t st,
                 int k_lo, int nk) {
  if (nk <= 0 || along_x) { k_lo = 0; nk = sp.chunks; }
  const int nlines = along_x ? g.h : g.w;
  const long chains = (long)(nlines - 2 * g.radius) * nk * nviews;
  if (chains <= 0) return 0;
  const unsigned blocks = (unsigned)((chains + 127) / 128);
  if (along_x)
    k_sweep_generic<true><<<blocks, 128, 0, st>>>(ref, mat, dc_in, dc_out, g, nviews, dir, sp.chunks,
                                                  sp.overlap, sp.alpha, 1 - sp.alpha, k_lo, nk);
  else
    k_sweep_generic<false><<<blocks, 128, 0, st>>>(ref, mat, dc_in, dc_out, g, nviews, dir, sp.chunks,
                                                   sp.overlap, sp.alpha, 1 - sp.alpha, k_lo, nk);
  return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// ==================================================== block-per-line kernels
//
// Walk of one chunk, in walking order (index j, position q = walk_first + dir*j):
//   [0, vis_lo)        positions before the first visited one; only the first chunk of
//                      a line has any (the border position): copied through
//   [vis_lo, vis_hi)   the positions the reference's thread visits: evaluated
//   [tail_lo, vis_hi)  the part of those that the next chunk visited first: {d, cost}
//                      comes from the output plane (written by that chunk before the
//                      block barrier at j == kBarrierStep) instead of the input plane
//   [vis_hi, nwalk)    positions after the last visited one; only the last chunk has
//                      any: copied through
// Every walked position is stored; where two chunks store the same position the later
// store (the predecessor's tail) comes after the barrier and is the final value.

constexpr int kPF = 4;             // software prefetch distance (steps), row kernel
#ifndef PM_PF_COL
#define PM_PF_COL 2
#endif
constexpr int kPFCol = PM_PF_COL;  // column kernel (coalesced, mostly L1/L2 hits)
constexpr int kColPrefetchDisp = 144;  // column kernel: L1 prefetch reach to the left of a warp
constexpr int kRowBarrierStep = 15;  // row kernel: barrier after the first tile flush

struct ChainGeom {
  int walk_first, nwalk, vis_lo, vis_hi, tail_lo, start;
};

__host__ __device__ inline void chunk_range_hd(int k, int cs, int ov, int len, int dir, int& start,
                                               int& stop) {
  const int a = k * cs - ov, b = (k + 1) * cs + ov;
  const int mn = a > 1 ? a : 1;
  const int mx = b < len - 2 ? b : len - 2;
  start = dir > 0 ? mn : mx;
  stop = dir > 0 ? mx : mn;
}

__host__ __device__ inline ChainGeom chain_geom(int k, int chunks, int cs, int ov, int len, int dir) {
  ChainGeom c;
  int start, stop;
  chunk_range_hd(k, cs, ov, len, dir, start, stop);
  const int nsteps = dir > 0 ? stop - start : start - stop;
  const int kn = k + dir, kp = k - dir;
  const bool has_n = kn >= 0 && kn < chunks, has_p = kp >= 0 && kp < chunks;
  c.start = start;
  c.walk_first = has_p ? start : (dir > 0 ? 0 : len - 1);
  c.vis_lo = dir > 0 ? start - c.walk_first : c.walk_first - start;
  c.vis_hi = c.vis_lo + nsteps;
  c.nwalk = has_n ? c.vis_hi : c.vis_hi + (dir > 0 ? len - stop : stop + 1);
  c.tail_lo = INT_MAX;
  if (has_n) {
    int sn, en;
    chunk_range_hd(kn, cs, ov, len, dir, sn, en);
    const int n_ov = dir > 0 ? stop - sn : sn - stop;
    c.tail_lo = c.vis_hi - (n_ov > 0 ? n_ov : 0);
  }
  return c;
}

// Host-side check that the barrier scheme is valid for this line length, and the
// block-uniform trip count.
bool sweep_block_plan(int len, int chunks, int ov, int bar_step, int pf, int max_chunks,
                      int* max_walk) {
  if (chunks < 2 || chunks > max_chunks || ov > 8) return false;
  const int cs = len / chunks;
  int mw = 0;
  for (int dir = -1; dir <= 1; dir += 2)
    for (int k = 0; k < chunks; ++k) {
      const ChainGeom c = chain_geom(k, chunks, cs, ov, len, dir);
      if (c.vis_hi <= c.vis_lo) return false;
      // the first handover read (prefetched kPF steps early) must come after the barrier,
      // and every head (at most 2*ov steps) must be stored before it
      if (c.tail_lo != INT_MAX && c.tail_lo - pf <= bar_step) return false;
      if (c.nwalk <= bar_step + 1 || 2 * ov > bar_step + 1) return false;
      if (c.nwalk > mw) mw = c.nwalk;
    }
  *max_walk = mw;
  return true;
}

struct TapPair { float2 l, r; };

struct Slot {      // what one walk step needs from memory, prefetched kPF steps ahead
  float2 cur;      // {d, cost} at the position
  RefTaps taps;    // reference taps around the position
};

// ------------------------------------------------------------------ row sweep
//
// A row sweep walks along x, the axis the disparity shifts along, so the lanes of a
// warp (different rows) gather from different lines of the matched image: through L1
// that is one wavefront per lane per load. Shared memory serves such gathers at bank
// granularity instead. One block owns kRows consecutive rows and all chunks of those rows:
//   * the kRows+2 matched-image rows are staged in shared memory once, full width, so
//     any disparity up to x-1 stays inside (the reference has no other bound);
//   * a half-warp holds the 16 rows of one chunk: the pre-sweep {d, cost} plane and the
//     reference taps come from TRANSPOSED planes (rows contiguous), 128-byte loads;
//   * results leave through a 16x16 shared tile per half-warp, written row-major with
//     128-byte stores.

constexpr int kRows = 16;
constexpr int kTilePitch = 17;

__host__ __device__ inline int row_spitch(int w) { return w + 1 + ((3 - (w + 1)) & 15); }

// Second-generation kernel: rows arrive by TMA bulk copies, which need 16-byte aligned shared
// rows, i.e. an even pitch: 2 (mod 16) float2. Bank pair of (row r, column c) is then
// (2r + c) mod 16: rows 8 apart share a bank when their columns agree, a two-way conflict
// that measurements show is not what bounds the kernel (it is latency-bound at 8 warps/SM).
__host__ __device__ inline int row_spitch2(int w) { return w + 2 + ((2 - (w + 2)) & 15); }
__host__ __device__ inline int row_copy_elems(int w) { return (w + 2) & ~1; }  // w+1 rounded to even

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
// TMA 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void tma_load_1d(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__global__ void __launch_bounds__(512)
k_sweep_row(const float2* __restrict__ refT, const float2* __restrict__ mat,
            const float2* __restrict__ dcT_in, float2* dc_out, ViewGeom g, int pitchT,
            size_t planeT, int dir, int chunks, int ov, int max_walk, float alpha, float w1) {
  extern __shared__ float2 smem[];
  const int w = g.w, h = g.h;
  // shared pitch = 3 (mod 16) float2: bank pair of (row r, column c) is (3r + c) mod 16, so
  // the 16 rows of a half-warp are conflict-free when their columns agree AND when
  // neighbouring rows differ by one column (slanted surfaces)
  const int spitch = row_spitch(w);
  float2* smat = smem;
  float2* tiles = smem + (size_t)(kRows + 2) * spitch;
  const int t = threadIdx.x, r = t & 15, k = t >> 4;
  const int y0 = blockIdx.x * kRows, v = blockIdx.y;
  refT += (size_t)v * planeT;
  dcT_in += (size_t)v * planeT;
  mat += (size_t)v * g.plane;
  dc_out += (size_t)v * g.plane;

  // stage the matched rows y0-1 .. y0+kRows (incl. the finite pad element at column w):
  // 8-byte cp.async, all in flight at once; the odd shared pitch rules out 16-byte copies
  {
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(smat);
    for (int row = 0; row < kRows + 2; ++row) {
      const int gy = min(max(y0 - 1 + row, 0), h - 1);
      const float2* src = mat + (size_t)gy * g.pitch;
      const unsigned dst = sbase + (unsigned)(row * spitch) * 8u;
      for (int c = t; c <= w; c += blockDim.x)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + 8u * c), "l"(src + c));
    }
    asm volatile("cp.async.commit_group;");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
  __syncthreads();

  const int y = y0 + r;
  const int yc = min(y, h - 1);                 // rows past the image mirror the last one
  const bool active = y < h && row_interior(g, y);  // rows the reference sweeps (:134)
  const ChainGeom cg = chain_geom(k, chunks, w / chunks, ov, w, dir);

  const float2* m1 = smat + (size_t)(r + 1) * spitch;  // matched row y
  const float2* m0 = m1 - spitch;
  const float2* m2 = m1 + spitch;
  float2* tile = tiles + (size_t)k * 16 * kTilePitch;

  const ptrdiff_t in_step = (ptrdiff_t)dir * pitchT;
  const float2* in_p = dcT_in + (size_t)cg.walk_first * pitchT + yc;    // input at walk index jj
  const float2* ref_p = refT + (size_t)cg.walk_first * pitchT + yc;
  const float2* ho_p = dc_out + (size_t)yc * g.pitch + cg.walk_first;   // handover (row-major)

  auto fetch = [&](Slot& s, int jj) {
    // in_p / ref_p / ho_p point at walk index jj
    const bool inside = jj < cg.nwalk;
    const bool vis = inside && active && jj >= cg.vis_lo && jj < cg.vis_hi;
    if (inside) s.cur = (vis && jj >= cg.tail_lo) ? __ldcg(ho_p) : *in_p;
    if (vis) {
      s.taps.tl = ref_p[-pitchT - 1];
      s.taps.bl = ref_p[-pitchT + 1];
      s.taps.c = ref_p[0];
      s.taps.tr = ref_p[pitchT - 1];
      s.taps.br = ref_p[pitchT + 1];
    }
    in_p += in_step;
    ref_p += in_step;
    ho_p += dir;
  };

  Slot ring[kPF];
#pragma unroll
  for (int u = 0; u < kPF; ++u) fetch(ring[u], u);

  // candidate for the first visited position: the pre-sweep disparity just before it
  float prev = dcT_in[(size_t)(cg.start - dir) * pitchT + yc].x;
  float xq = __int2float_rn(cg.walk_first);  // position as float, stepped exactly
  const float fdir = (float)dir;

  for (int j0 = 0; j0 < max_walk; j0 += kPF) {
#pragma unroll
    for (int u = 0; u < kPF; ++u) {
      const int j = j0 + u;
      Slot& s = ring[u];
      float2 cur = s.cur;
      if (active && j >= cg.vis_lo && j < cg.vis_hi) {
        const float xr = fmaxf(__fsub_rn(xq, prev), 1.0f);
        const float c1 = cost5_rows(s.taps, m0, m1, m2, xr, alpha, w1);
        if (c1 < cur.y) {
          cur.x = fminf(prev, __fsub_rn(xq, 1.0f));
          cur.y = c1;
        }
        prev = cur.x;
      }
      tile[r * kTilePitch + (j & 15)] = cur;
      fetch(s, j + kPF);
      xq = __fadd_rn(xq, fdir);
      if ((j & 15) == 15 || j == max_walk - 1) {
        // flush walk indices [j & ~15, j]: lane r stores tile column r of all 16 rows
        __syncwarp();
        const int jc = (j & ~15) + r;
        if (jc <= j && jc < cg.nwalk) {
          float2* o = dc_out + (size_t)y0 * g.pitch + (cg.walk_first + dir * jc);
#pragma unroll
          for (int rr = 0; rr < 16; ++rr)
            if (y0 + rr < h) o[(size_t)rr * g.pitch] = tile[rr * kTilePitch + r];
        }
        __syncwarp();
        if (j == kRowBarrierStep) __syncthreads();  // heads are stored: successors may read them
      }
    }
  }
}

// ---------------------------------------------------- row sweep, second generation
//
// Same block shape, shared-memory layout and barrier handover as k_sweep_row; the step itself
// is rebuilt for issue slots (the first version needed ~190 per step at 8 warps per SM):
// direction as a template parameter, one running pointer per plane, FADD.RM floor, packed
// f32x2 arithmetic, rolling reference taps (the column ahead of step i is the column behind
// step i+2: 3 loads per step instead of 5) and a register ring kRowP steps deep for the
// loads that come from HBM.

#ifndef PM_ROW_P
#define PM_ROW_P 13
#endif
constexpr int kRowP = PM_ROW_P;  // P + 3 ring slots; 16 (the tile period) must be a multiple
#ifndef PM_ROW_WINDOW
#define PM_ROW_WINDOW 1
#endif
constexpr bool kRowWindow = PM_ROW_WINDOW != 0;  // candidate evaluations through MatWin (pm_device.cuh)

// NOISE: AddForegroundNoise (patchmatch_gpu.cu:298-304) and the cost refresh run inside the
// sweep: dcT_in is the plane BEFORE the noise, every walked position first becomes
// {d', cost(d')} exactly as k_noise_cost would have left it (its gathers come from the rows
// already staged here instead of a second trip to HBM), then the sweep step follows.
struct RowNoise {
  const float* noiseT;   // the U(-1,1) image, transposed like dcT ([x][pitchT])
  float scale, dmax;
  int skip_eq;           // plain sweeps: skip warp-steps whose candidates all equal the own disparity
  int walk_end;          // longest walk of any chunk: the last tile period stops there (block-uniform)
};

__device__ __forceinline__ float noised(float d, float nz, float scale, float dmax) {
  const float t = __fmaf_rn(scale, nz, d);
  return d > 0.0f ? fminf(t > 0.0f ? t : 0.0f, dmax) : 0.0f;
}

// RMIN: the pre-sweep {d, cost} plane is read ROW-MAJOR (dc_rm) through the block's 16x16 tiles
// instead of from a transposed copy: at the start of a 16-step period a half-warp loads the NEXT
// period's 16 rows x 16 positions with 128-byte row segments into registers, and after the
// period's results have left the tile those registers refill it; a step then finds its {d, cost}
// in the tile slot it will overwrite with its result. No k_transpose2 pass before the sweep.
// Handed-over positions (j >= tail_lo) are read from dc_out, which is row-major anyway; their
// loads are issued after the block barrier, so tail_lo >= 32 is required (sweep_row_rm_supported).
//
// SPEC: one-step speculation. A step's candidate is the disparity the previous position ended up
// with, which is one of two values known BEFORE that position's comparison: the candidate it tested
// (clamped), if it accepted, or its own disparity, if it kept it. Both costs of the next position
// are evaluated while the current decision is still pending, and the decision then only selects:
// the dependent chain per position shrinks from a whole evaluation to a compare and two selects,
// at the price of one more (chain-independent) evaluation per step. Same operands, same
// operations, same order of decisions: bit-identical to the serial form.
//
// IL: the block's 16 matched rows are staged slot-interleaved ([column][16 rows], pm_device.cuh
// RowsIL) from matI, a copy of the matched plane in that order made once per level
// (k_interleave16): bank-conflict-free gathers whatever the columns. The halo rows come row-major.
struct RowIL {
  const float2* matI;   // [view][row group][column][16 rows]
  int cols;             // columns per group (= row_copy_elems(w))
  size_t plane;         // elements per view
};

template <int DIR, bool NOISE, bool RMIN, bool SPEC, bool IL>
__global__ void __launch_bounds__(256)
k_sweep_row2(const float2* __restrict__ refT, const float2* __restrict__ mat,
             const float2* __restrict__ dcT_in, float2* dc_out, ViewGeom g, int pitchT,
             size_t planeT, int chunks, int ov, int max_walk, float alpha, float w1,
             RowNoise nz, const float2* __restrict__ dc_rm, RowIL il) {
  constexpr int P = kRowP, NA = P + 3;
  extern __shared__ __align__(16) float2 smem2[];
  const int w = g.w, h = g.h;
  const int spitch = row_spitch2(w);
  float2* smat = smem2;
  float2* tiles = smem2 + (size_t)(kRows + 2) * spitch;
  __shared__ __align__(8) unsigned long long stage_bar;
  const int t = threadIdx.x, r = t & 15, k = t >> 4;
  const int y0 = blockIdx.x * kRows, v = blockIdx.y;
  const int W1 = row_copy_elems(w);   // IL: columns staged per row
  refT += (size_t)v * planeT;
  if (!RMIN) dcT_in += (size_t)v * planeT;
  if (RMIN) dc_rm += (size_t)v * g.plane;
  mat += (size_t)v * g.plane;
  dc_out += (size_t)v * g.plane;

  // stage the matched rows y0-1 .. y0+kRows (incl. the finite pad element at column w) with
  // one TMA bulk copy per row: a single thread issues them, the bytes land asynchronously
  // and are counted on an mbarrier
  {
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&stage_bar);
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(smat);
    const unsigned row_bytes = (unsigned)row_copy_elems(w) * 8u;
    if (t == 0) {
      mbar_init(bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (t == 0) {
      mbar_expect_tx(bar, row_bytes * (kRows + 2));
      if (IL) {
        // [column][16 rows] of this row group in four bulk copies, then the two halo rows
        const float2* grp = il.matI + (size_t)v * il.plane + (size_t)blockIdx.x * il.cols * 16;
        const unsigned total = row_bytes * kRows;
        const unsigned part = (total / 4u) & ~15u;
        for (int i = 0; i < 4; ++i) {
          const unsigned off = part * i, len = i == 3 ? total - off : part;
          tma_load_1d(sbase + off, reinterpret_cast<const char*>(grp) + off, len, bar);
        }
        const int gt = max(y0 - 1, 0), gb = min(y0 + kRows, h - 1);
        tma_load_1d(sbase + total, mat + (size_t)gt * g.pitch, row_bytes, bar);
        tma_load_1d(sbase + total + row_bytes, mat + (size_t)gb * g.pitch, row_bytes, bar);
      } else {
        for (int row = 0; row < kRows + 2; ++row) {
          const int gy = min(max(y0 - 1 + row, 0), h - 1);
          tma_load_1d(sbase + (unsigned)(row * spitch) * 8u, mat + (size_t)gy * g.pitch, row_bytes, bar);
        }
      }
    }
  }

  const int y = y0 + r;
  const int yc = min(y, h - 1);                     // rows past the image mirror the last one
  const bool active = y < h && row_interior(g, y);  // rows the reference sweeps (:134)
  const ChainGeom cg = chain_geom(k, chunks, w / chunks, ov, w, DIR);

  const float2* m1 = smat + (size_t)(r + 1) * spitch;  // matched row y
  const float2* m0 = m1 - spitch;
  const float2* m2 = m1 + spitch;
  float2* tile = tiles + (size_t)k * 16 * kTilePitch;

  const ptrdiff_t se = (ptrdiff_t)DIR * pitchT;
  // pointers at walk index j + P (the loads run P steps ahead of the evaluation)
  const float2* in_p = dcT_in + (size_t)cg.walk_first * pitchT + yc;
  const float2* rf_p = refT + (size_t)cg.walk_first * pitchT + yc;
  const float2* ho_p = dc_out + (size_t)yc * g.pitch + cg.walk_first;   // handover (row-major)

  // ring: A[(j+2) % NA] = reference column ahead of walk index j (rows y-1, y+1); the column
  // behind index j is the one that was ahead of index j-2. Loads only where evaluated.
  TapPair A[NA];
  float2 C[NA], CUR[NA];
  float NZ[NA];
  const float* nz_p = NOISE ? nz.noiseT + (size_t)cg.walk_first * pitchT + yc : nullptr;
  auto visible = [&](int jj) { return active && jj >= cg.vis_lo && jj < cg.vis_hi; };
  // positions whose cost exists (k_noise_cost's `interior`): with NOISE their taps are needed
  // whether the sweep visits them or not
  auto costed = [&](int jj) {
    const int x = cg.walk_first + DIR * jj;
    return active && jj < cg.nwalk && x >= 1 && x <= w - 2;
  };
  auto needs_taps = [&](int jj) { return NOISE ? costed(jj) : visible(jj); };
  auto fetch = [&](int slot, int jj) {   // in_p / rf_p / ho_p point at walk index jj
    if (jj < cg.nwalk) {
      if (!RMIN) {
        const bool vis = visible(jj);
        CUR[slot] = (vis && jj >= cg.tail_lo) ? __ldcg(ho_p) : *in_p;
      }
      if (NOISE) NZ[slot] = *nz_p;
      if (needs_taps(jj)) C[slot] = rf_p[0];
      if (needs_taps(jj) || needs_taps(jj + 2)) {      // ahead of jj == behind jj+2
        A[(slot + 2) % NA].l = rf_p[se - 1];
        A[(slot + 2) % NA].r = rf_p[se + 1];
      }
    }
    in_p += se;
    rf_p += se;
    ho_p += DIR;
    if (NOISE) nz_p += se;
  };
  // the columns behind walk indices 0 and 1 ("ahead" of the indices -2 and -1)
  if (needs_taps(0)) { A[0].l = rf_p[-se - 1]; A[0].r = rf_p[-se + 1]; }
  if (needs_taps(1)) { A[1].l = rf_p[-1];      A[1].r = rf_p[1]; }
#pragma unroll
  for (int u = 0; u < P; ++u) fetch(u, u);

  // RMIN: IN[rr] = {d, cost} of row y0+rr at walk index 16*period + r (tile column r)
  float2 IN[16];
  const unsigned act16 = __ballot_sync(0xffffffffu, active) & 0xffffu;  // rows the reference sweeps
  auto load_period = [&](int period) {
    const int jc = 16 * period + r;
    if (jc < cg.nwalk) {
      const bool tail = jc >= cg.vis_lo && jc < cg.vis_hi && jc >= cg.tail_lo;
      const int xp = cg.walk_first + DIR * jc;
#pragma unroll
      for (int rr = 0; rr < 16; ++rr) {
        const size_t o = (size_t)min(y0 + rr, h - 1) * g.pitch + xp;
        IN[rr] = (tail && ((act16 >> rr) & 1u)) ? __ldcg(dc_out + o) : dc_rm[o];
      }
    }
  };
  auto store_period = [&]() {
#pragma unroll
    for (int rr = 0; rr < 16; ++rr) tile[rr * kTilePitch + r] = IN[rr];
  };
  if (RMIN) {
    load_period(0);
    store_period();
    __syncwarp();
    load_period(1);
  }

  // candidate for the first visited position: the (noised) pre-sweep disparity before it
  mbar_wait((unsigned)__cvta_generic_to_shared(&stage_bar), 0);   // the rows have landed

  MatWin win;
  win.cc = INT_MIN / 2;
  win.t = 0.0f;
#pragma unroll
  for (int i = 0; i < 4; ++i) win.a[i] = win.b[i] = make_float2(0.0f, 0.0f);
  win.c[0] = win.c[1] = make_float2(0.0f, 0.0f);
  const unsigned s1a = (unsigned)__cvta_generic_to_shared(m1);
  const unsigned s0a = s1a - (unsigned)spitch * 8u, s2a = s1a + (unsigned)spitch * 8u;
  RowsIL RI;
  {
    const char* mainp = reinterpret_cast<const char*>(smat);
    const char* topp = mainp + (size_t)W1 * 128;
    const char* botp = topp + (size_t)W1 * 8;
    RI.p1 = mainp + r * 8;
    RI.p0 = r == 0 ? topp : mainp + (r - 1) * 8;
    RI.p2 = r == kRows - 1 ? botp : mainp + (r + 1) * 8;
    RI.s0 = r == 0 ? 8u : 128u;
    RI.s2 = r == kRows - 1 ? 8u : 128u;
    RI.b0 = (unsigned)__cvta_generic_to_shared(RI.p0);
    RI.b1 = (unsigned)__cvta_generic_to_shared(RI.p1);
    RI.b2 = (unsigned)__cvta_generic_to_shared(RI.p2);
  }
  auto cost_full = [&](const RefTaps& L, float xr) {
    return IL ? cost5_packed_il(L, RI, xr, alpha, w1) : cost5_packed<false>(L, m0, m1, m2, xr, alpha, w1);
  };
  auto cost_win = [&](const RefTaps& L, float xr) {
    return IL ? cost5_window_il<DIR>(L, win, RI, xr, alpha, w1)
              : cost5_window<DIR>(L, win, s0a, s1a, s2a, xr, alpha, w1);
  };

  float prev = RMIN ? dc_rm[(size_t)yc * g.pitch + (cg.start - DIR)].x
                    : dcT_in[(size_t)(cg.start - DIR) * pitchT + yc].x;
  if (NOISE) prev = noised(prev, nz.noiseT[(size_t)(cg.start - DIR) * pitchT + yc], nz.scale, nz.dmax);
  float xq = __int2float_rn(cg.walk_first);  // position as float, stepped exactly
  const float fdir = (float)DIR, wf = __int2float_rn(w - 2);

  // SPEC state: candidate and cost at the CURRENT position if the previous one accepted (a, ea)
  // or kept its own disparity (b, eb); before the first visited position: the pre-sweep neighbour
  float sp_a = 0.0f, sp_ea = 0.0f, sp_b = prev, sp_eb = 0.0f;
  bool sp_acc = false;
  if (SPEC) {
    RefTaps L0;
    L0.c = C[0];
    if (DIR > 0) { L0.tl = A[0].l; L0.bl = A[0].r; L0.tr = A[2].l; L0.br = A[2].r; }
    else         { L0.tr = A[0].l; L0.br = A[0].r; L0.tl = A[2].l; L0.bl = A[2].r; }
    sp_eb = cost_full(L0, fminf(fmaxf(__fsub_rn(xq, prev), 1.0f), wf));
  }

  static_assert(16 % NA == 0, "the tile period must be a whole number of ring turns");
  for (int j0 = 0; j0 < max_walk; j0 += 16) {   // max_walk is a multiple of 16
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int j = j0 + u;
      float2 cur = RMIN ? tile[r * kTilePitch + u] : CUR[u % NA];
      if (SPEC) {
        const bool vis = visible(j);
        const TapPair bh = A[u % NA], ah = A[(u + 2) % NA];
        if (NOISE && !(vis && j >= cg.tail_lo)) {   // handed-over positions are already done
          RefTaps L;
          L.c = C[u % NA];
          if (DIR > 0) { L.tl = bh.l; L.bl = bh.r; L.tr = ah.l; L.br = ah.r; }
          else         { L.tr = bh.l; L.br = bh.r; L.tl = ah.l; L.bl = ah.r; }
          const float dn = noised(cur.x, NZ[u % NA], nz.scale, nz.dmax);
          const float xn = fminf(fmaxf(__fsub_rn(xq, dn), 1.0f), wf);
          const float cn = cost_full(L, xn);
          cur.x = dn;
          cur.y = costed(j) ? cn : 0.0f;
        }
        // the candidate of this position and its cost, both evaluated one step ago
        const float cd = sp_acc ? sp_a : sp_b;
        const float c1 = sp_acc ? sp_ea : sp_eb;
        // what this pixel holds if it accepts (patchmatch_gpu.cu:169) = the next candidate then
        const float an = fminf(cd, __fsub_rn(xq, 1.0f));
        // ... and if it keeps its own disparity (positions before the first visited one pass the
        // pre-sweep neighbour on)
        const float bn = vis ? cur.x : sp_b;
        // reference taps of the NEXT position
        const TapPair nbh = A[(u + 1) % NA], nah = A[(u + 3) % NA];
        RefTaps Ln;
        Ln.c = C[(u + 1) % NA];
        if (DIR > 0) { Ln.tl = nbh.l; Ln.bl = nbh.r; Ln.tr = nah.l; Ln.br = nah.r; }
        else         { Ln.tr = nbh.l; Ln.br = nbh.r; Ln.tl = nah.l; Ln.bl = nah.r; }
        const float xnext = __fadd_rn(xq, fdir);
        const float xra = fminf(fmaxf(__fsub_rn(xnext, an), 1.0f), wf);
        const float xrb = fminf(fmaxf(__fsub_rn(xnext, bn), 1.0f), wf);
        sp_ea = cost_win(Ln, xra);
        sp_eb = cost_full(Ln, xrb);
        const bool acc = vis && c1 < cur.y;
        if (acc) {
          cur.x = an;
          cur.y = c1;
        }
        sp_a = an;
        sp_b = bn;
        sp_acc = acc;
      } else {

        // evaluated on every lane (no branch: one basic block per step); lanes that the
        // reference does not visit hold stale taps, their result is dropped by `vis`.
        // The clamp to w is a no-op where visited (xr <= x) and keeps stale lanes in range.
        const bool vis = visible(j);
        const TapPair bh = A[u % NA], ah = A[(u + 2) % NA];
        RefTaps L;
        L.c = C[u % NA];
        if (DIR > 0) { L.tl = bh.l; L.bl = bh.r; L.tr = ah.l; L.br = ah.r; }
        else         { L.tr = bh.l; L.br = bh.r; L.tl = ah.l; L.bl = ah.r; }
        if (NOISE && !(vis && j >= cg.tail_lo)) {   // handed-over positions are already done
          const float dn = noised(cur.x, NZ[u % NA], nz.scale, nz.dmax);
          const float xn = fminf(fmaxf(__fsub_rn(xq, dn), 1.0f), wf);
          const float cn = cost_full(L, xn);
          cur.x = dn;
          cur.y = costed(j) ? cn : 0.0f;
        }
        // Plain sweeps: when no visited lane of the warp has a candidate that differs from its own
        // disparity, every cached cost equals the candidate's and the strict `<` fails on all of
        // them: skip the evaluation (warp-uniform; never the case right after the noise).
        if (NOISE || !nz.skip_eq || __any_sync(0xffffffffu, vis && prev != cur.x)) {
          const float xr = fminf(fmaxf(__fsub_rn(xq, prev), 1.0f), wf);
          const float c1 = kRowWindow ? cost_win(L, xr) : cost_full(L, xr);
          if (vis && c1 < cur.y) {
            cur.x = fminf(prev, __fsub_rn(xq, 1.0f));
            cur.y = c1;
          }
        }
        if (vis) prev = cur.x;
      }
      tile[r * kTilePitch + u] = cur;
      fetch((u + P) % NA, j + P);
      xq = __fadd_rn(xq, fdir);
    }
    // flush walk indices [j0, j0+16): lane r stores tile column r of all 16 rows
    __syncwarp();
    const int jc = j0 + r;
    if (jc < cg.nwalk) {
      float2* o = dc_out + (size_t)y0 * g.pitch + (cg.walk_first + DIR * jc);
#pragma unroll
      for (int rr = 0; rr < 16; ++rr)
        if (y0 + rr < h) o[(size_t)rr * g.pitch] = tile[rr * kTilePitch + r];
    }
    __syncwarp();
    if (RMIN) {   // the next period's inputs take the slots the results just left
      store_period();
      __syncwarp();
    }
    if (j0 == 0) __syncthreads();  // heads (< 16 steps) are stored: successors may read them
    if (RMIN) load_period(j0 / 16 + 2);
  }
}

// ---------------------------------------------------- row sweep, third generation
#ifndef PM_ROW_AHEAD
#define PM_ROW_AHEAD 4
#endif
#ifndef PM_ROW_AHEAD_IN
#define PM_ROW_AHEAD_IN 1
#endif
//
// k_sweep_row2 with the slot-interleaved staging (RowIL) and NO data-dependent branch inside the
// sixteen unrolled steps of a tile period: conditional prefetches are predicated loads, the noise
// refresh and the decision are selects, the right-hand taps are always split on their own
// (cost5_full_bf / cost5_win_bf). The period is one basic block, so the scheduler overlaps the
// loads, address arithmetic and window moves of step j+1 with the dependent chain of step j.
// SPEC: one-step speculation as described at k_sweep_row2.
template <int DIR, bool NOISE, bool SPEC, bool RMIN>
__global__ void __launch_bounds__(256)
k_sweep_row3(const float2* __restrict__ refT, const float2* __restrict__ mat,
             const float2* __restrict__ dcT_in, float2* dc_out, ViewGeom g, int pitchT,
             size_t planeT, int chunks, int ov, int max_walk, float alpha, float w1,
             RowNoise nz, RowIL il, const float2* __restrict__ dc_rm) {
  constexpr int P = kRowP, NA = P + 3;
  extern __shared__ __align__(16) float2 smem2[];
  const int w = g.w, h = g.h;
  const int spitch = row_spitch2(w);
  float2* smat = smem2;
  float2* tiles = smem2 + (size_t)(kRows + 2) * spitch;
  __shared__ __align__(8) unsigned long long stage_bar;
  const int t = threadIdx.x, r = t & 15, k = t >> 4;
  const int y0 = blockIdx.x * kRows, v = blockIdx.y;
  const int W1 = row_copy_elems(w);
  refT += (size_t)v * planeT;
  if (!RMIN) dcT_in += (size_t)v * planeT;
  if (RMIN) dc_rm += (size_t)v * g.plane;
  mat += (size_t)v * g.plane;
  dc_out += (size_t)v * g.plane;

  {
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&stage_bar);
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(smat);
    const unsigned row_bytes = (unsigned)W1 * 8u;
    if (t == 0) {
      mbar_init(bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (t == 0) {
      mbar_expect_tx(bar, row_bytes * (kRows + 2));
      const float2* grp = il.matI + (size_t)v * il.plane + (size_t)blockIdx.x * il.cols * 16;
      const unsigned total = row_bytes * kRows;
      const unsigned part = (total / 4u) & ~15u;
      for (int i = 0; i < 4; ++i) {
        const unsigned off = part * i, len = i == 3 ? total - off : part;
        tma_load_1d(sbase + off, reinterpret_cast<const char*>(grp) + off, len, bar);
      }
      const int gt = max(y0 - 1, 0), gb = min(y0 + kRows, h - 1);
      tma_load_1d(sbase + total, mat + (size_t)gt * g.pitch, row_bytes, bar);
      tma_load_1d(sbase + total + row_bytes, mat + (size_t)gb * g.pitch, row_bytes, bar);
    }
  }

  const int y = y0 + r;
  const int yc = min(y, h - 1);
  const bool active = y < h && row_interior(g, y);
  const ChainGeom cg = chain_geom(k, chunks, w / chunks, ov, w, DIR);
  float2* tile = tiles + (size_t)k * 16 * kTilePitch;

  const ptrdiff_t se = (ptrdiff_t)DIR * pitchT;
  const float2* in_p = dcT_in + (size_t)cg.walk_first * pitchT + yc;
  const float2* rf_p = refT + (size_t)cg.walk_first * pitchT + yc;
  const float2* ho_p = dc_out + (size_t)yc * g.pitch + cg.walk_first;
  const float* nz_p = NOISE ? nz.noiseT + (size_t)cg.walk_first * pitchT + yc : nullptr;
  const int pf_dy = r == 0 ? (yc > 0 ? -1 : 0) : (r == kRows - 1 && yc < h - 1 ? 1 : 0);

  TapPair A[NA];
  float2 C[NA], CUR[NA];
  float NZ[NA];
#pragma unroll
  for (int i = 0; i < NA; ++i) {
    A[i].l = A[i].r = C[i] = CUR[i] = make_float2(0.0f, 0.0f);
    NZ[i] = 0.0f;
  }
  auto visible = [&](int jj) { return active && jj >= cg.vis_lo && jj < cg.vis_hi; };
  auto costed = [&](int jj) {
    const int x = cg.walk_first + DIR * jj;
    return active && jj < cg.nwalk && x >= 1 && x <= w - 2;
  };
  auto needs_taps = [&](int jj) { return NOISE ? costed(jj) : visible(jj); };
  auto fetch = [&](int slot, int jj) {   // pointers are at walk index jj; every load is predicated
    const bool in = jj < cg.nwalk;
    if (!RMIN) {
      const bool tail = visible(jj) && jj >= cg.tail_lo;
      ldg_cg_f2_if(CUR[slot], tail ? ho_p : in_p, in);
    }
    if (NOISE) ldg_nc_f32_if(NZ[slot], nz_p, in);
    const bool nt = needs_taps(jj);
    ldg_nc_f2_if(C[slot], rf_p, in && nt);
    const bool na = in && (nt || needs_taps(jj + 2));      // ahead of jj == behind jj+2
    ldg_nc_f2_if(A[(slot + 2) % NA].l, rf_p + se - 1, na);
    ldg_nc_f2_if(A[(slot + 2) % NA].r, rf_p + se + 1, na);
#if PM_ROW_AHEAD > 0
    // ptxas puts the ring loads on a scoreboard that other instructions of the period wait on, so
    // the kernel ends up waiting for the youngest ring load a few times per period: make those
    // loads L1 hits by pulling their lines in a few steps earlier (prefetches carry no scoreboard).
    // Lanes 0 and 15 of a half-warp take the rows just outside the block.
    if (jj + PM_ROW_AHEAD + 1 < cg.nwalk) {
      asm volatile("prefetch.global.L1 [%0];" ::"l"(rf_p + (PM_ROW_AHEAD + 1) * se + pf_dy));
      if (NOISE) asm volatile("prefetch.global.L1 [%0];" ::"l"(nz_p + PM_ROW_AHEAD * se));
    }
#endif
    in_p += se;
    rf_p += se;
    ho_p += DIR;
    if (NOISE) nz_p += se;
  };
  ldg_nc_f2_if(A[0].l, rf_p - se - 1, needs_taps(0));
  ldg_nc_f2_if(A[0].r, rf_p - se + 1, needs_taps(0));
  ldg_nc_f2_if(A[1].l, rf_p - 1, needs_taps(1));
  ldg_nc_f2_if(A[1].r, rf_p + 1, needs_taps(1));
#pragma unroll
  for (int u = 0; u < P; ++u) fetch(u, u);

  // RMIN: the pre-sweep {d, cost} plane is read ROW-MAJOR through the block's tiles (see k_sweep_row2):
  // IN[rr] = row y0+rr at walk index 16*period + r, loaded one period ahead
  float2 IN[16];
  const unsigned act16 = __ballot_sync(0xffffffffu, active) & 0xffffu;
  auto load_period = [&](int period) {
    const int jc = 16 * period + r;
    const bool in = jc < cg.nwalk;
    const bool tail = jc >= cg.vis_lo && jc < cg.vis_hi && jc >= cg.tail_lo;
    const int xp = cg.walk_first + DIR * jc;
#pragma unroll
    for (int rr = 0; rr < 16; ++rr) {
      const size_t o = (size_t)min(y0 + rr, h - 1) * g.pitch + xp;
      const bool t = tail && ((act16 >> rr) & 1u);
      ldg_cg_f2_if(IN[rr], (t ? (const float2*)dc_out : dc_rm) + o, in);
#if PM_ROW_AHEAD_IN > 0
      // the segment of the period after goes to L2 now (never a handed-over one: those are in L2)
      if (jc + 16 * PM_ROW_AHEAD_IN < cg.nwalk)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(dc_rm + o + DIR * 16 * PM_ROW_AHEAD_IN));
#endif
    }
  };
  auto store_period = [&]() {
#pragma unroll
    for (int rr = 0; rr < 16; ++rr) tile[rr * kTilePitch + r] = IN[rr];
  };
  if (RMIN) {
#pragma unroll
    for (int rr = 0; rr < 16; ++rr) IN[rr] = make_float2(0.0f, 0.0f);
    load_period(0);
    store_period();
    __syncwarp();
    load_period(1);
  }

  mbar_wait((unsigned)__cvta_generic_to_shared(&stage_bar), 0);   // the rows have landed

  MatWin win;
  win.cc = INT_MIN / 2;
  win.t = 0.0f;
#pragma unroll
  for (int i = 0; i < 4; ++i) win.a[i] = win.b[i] = make_float2(0.0f, 0.0f);
  win.c[0] = win.c[1] = make_float2(0.0f, 0.0f);
  RowsIL RI;
  {
    const char* mainp = reinterpret_cast<const char*>(smat);
    const char* topp = mainp + (size_t)W1 * 128;
    const char* botp = topp + (size_t)W1 * 8;
    RI.p1 = mainp + r * 8;
    RI.p0 = r == 0 ? topp : mainp + (r - 1) * 8;
    RI.p2 = r == kRows - 1 ? botp : mainp + (r + 1) * 8;
    RI.s0 = r == 0 ? 8u : 128u;
    RI.s2 = r == kRows - 1 ? 8u : 128u;
    RI.b0 = (unsigned)__cvta_generic_to_shared(RI.p0);
    RI.b1 = (unsigned)__cvta_generic_to_shared(RI.p1);
    RI.b2 = (unsigned)__cvta_generic_to_shared(RI.p2);
  }

  float prev = RMIN ? dc_rm[(size_t)yc * g.pitch + (cg.start - DIR)].x
                    : dcT_in[(size_t)(cg.start - DIR) * pitchT + yc].x;
  if (NOISE) prev = noised(prev, nz.noiseT[(size_t)(cg.start - DIR) * pitchT + yc], nz.scale, nz.dmax);
  float xq = __int2float_rn(cg.walk_first);
  const float fdir = (float)DIR, wf = __int2float_rn(w - 2);
  auto clampx = [&](float x) { return fminf(fmaxf(x, 1.0f), wf); };
  auto taps_of = [&](int s0, int s2, int sc) {   // ring slots: behind, ahead, centre
    RefTaps L;
    L.c = C[sc];
    if (DIR > 0) { L.tl = A[s0].l; L.bl = A[s0].r; L.tr = A[s2].l; L.br = A[s2].r; }
    else         { L.tr = A[s0].l; L.br = A[s0].r; L.tl = A[s2].l; L.bl = A[s2].r; }
    return L;
  };

  float sp_a = 0.0f, sp_ea = 0.0f, sp_b = prev, sp_eb = 0.0f;
  bool sp_acc = false;
  if (SPEC) sp_eb = cost5_full_bf(taps_of(0, 2, 0), RI, clampx(__fsub_rn(xq, prev)), alpha, w1);

  static_assert(16 % NA == 0, "the tile period must be a whole number of ring turns");
  for (int j0 = 0; j0 < max_walk; j0 += 16) {   // max_walk is a multiple of 16
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int j = j0 + u;
      // the walk of the longest chunk ends inside the last tile period (90 of 96 steps at w = 1280,
      // 50 of 64 at the 640-wide pyramid level): stop there, block-uniformly
      if (j >= nz.walk_end) break;
      float2 cur = RMIN ? tile[r * kTilePitch + u] : CUR[u % NA];
      const bool vis = visible(j);
      if (NOISE) {
        // AddForegroundNoise + cost refresh on every position but the handed-over ones
        const bool refresh = !(vis && j >= cg.tail_lo);
        const float dn = noised(cur.x, NZ[u % NA], nz.scale, nz.dmax);
        const float cn = cost5_full_bf(taps_of(u % NA, (u + 2) % NA, u % NA), RI,
                                       clampx(__fsub_rn(xq, dn)), alpha, w1);
        cur.x = refresh ? dn : cur.x;
        cur.y = refresh ? (costed(j) ? cn : 0.0f) : cur.y;
      }
      if (SPEC) {
        const float cd = sp_acc ? sp_a : sp_b;       // this position's candidate, its cost
        const float c1 = sp_acc ? sp_ea : sp_eb;
        const float an = fminf(cd, __fsub_rn(xq, 1.0f));   // what an accepting pixel holds (:169)
        const float bn = vis ? cur.x : sp_b;               // what a declining one holds
        const RefTaps Ln = taps_of((u + 1) % NA, (u + 3) % NA, (u + 1) % NA);
        const float xnext = __fadd_rn(xq, fdir);
#ifdef PM_ROW3_NOWIN
        sp_ea = cost5_full_bf(Ln, RI, clampx(__fsub_rn(xnext, an)), alpha, w1);
#else
        sp_ea = cost5_win_bf<DIR>(Ln, win, RI, clampx(__fsub_rn(xnext, an)), alpha, w1);
#endif
        sp_eb = cost5_full_bf(Ln, RI, clampx(__fsub_rn(xnext, bn)), alpha, w1);
        const bool acc = vis && c1 < cur.y;
        cur.x = acc ? an : cur.x;
        cur.y = acc ? c1 : cur.y;
        sp_a = an;
        sp_b = bn;
        sp_acc = acc;
      } else {
#ifdef PM_ROW3_NOWIN
        const float c1 = cost5_full_bf(taps_of(u % NA, (u + 2) % NA, u % NA), RI,
                                       clampx(__fsub_rn(xq, prev)), alpha, w1);
#else
        const float c1 = cost5_win_bf<DIR>(taps_of(u % NA, (u + 2) % NA, u % NA), win, RI,
                                           clampx(__fsub_rn(xq, prev)), alpha, w1);
#endif
        const bool acc = vis && c1 < cur.y;
        cur.x = acc ? fminf(prev, __fsub_rn(xq, 1.0f)) : cur.x;
        cur.y = acc ? c1 : cur.y;
        prev = vis ? cur.x : prev;
      }
      tile[r * kTilePitch + u] = cur;
      fetch((u + P) % NA, j + P);
      xq = __fadd_rn(xq, fdir);
    }
    // flush walk indices [j0, j0+16): lane r stores tile column r of all 16 rows
    __syncwarp();
    const int jc = j0 + r;
    if (jc < cg.nwalk) {
      float2* o = dc_out + (size_t)y0 * g.pitch + (cg.walk_first + DIR * jc);
#pragma unroll
      for (int rr = 0; rr < 16; ++rr)
        if (y0 + rr < h) o[(size_t)rr * g.pitch] = tile[rr * kTilePitch + r];
    }
    __syncwarp();
    if (RMIN) {   // the next period's inputs take the slots the results just left
      store_period();
      __syncwarp();
    }
    if (j0 == 0) __syncthreads();  // heads (< 16 steps) are stored: successors may read them
    if (RMIN) load_period(j0 / 16 + 2);
  }
}

static size_t sweep_row2_smem_bytes(int w, int chunks) {
  return ((size_t)(kRows + 2) * row_spitch2(w) + (size_t)chunks * 16 * kTilePitch) * sizeof(float2);
}

size_t sweep_row_smem_bytes(int w, int chunks) {
  return ((size_t)(kRows + 2) * row_spitch(w) + (size_t)chunks * 16 * kTilePitch) * sizeof(float2);
}

bool sweep_row_fuses_noise(int w, int chunks, int ov) {
  int mw;
  return !use_v1() && sweep_row_supported(w, chunks, ov) &&
         sweep_row2_smem_bytes(w, chunks) + 64 <= (size_t)227 * 1024 &&
         sweep_block_plan(w, chunks, ov, kRowBarrierStep, kRowP, 16, &mw);
}

// every handed-over position must be loaded after the block barrier that follows the first tile
// period: loads run up to two periods (32 steps) ahead
static bool row_rm_plan(int w, int chunks, int ov) {
  const int cs = w / chunks;
  for (int dir = -1; dir <= 1; dir += 2)
    for (int k = 0; k < chunks; ++k) {
      const ChainGeom c = chain_geom(k, chunks, cs, ov, w, dir);
      if (c.tail_lo != INT_MAX && c.tail_lo < 32) return false;
    }
  return true;
}

// Row-major input (no k_transpose2 before a row sweep). Measured on B200, 64 pairs of 1280x720: in the
// second-generation kernel it cost 3.0 ms against the 2.07 ms of transposes it saves; in the third
// generation (per-step exits, predicated loads) 1.2 ms: a net gain of 0.5-0.9 ms per device pass, so it
// is on by default there (PM_ROW_RMIN=0 switches it off).
static bool use_rm() {
  static const int v = [] { const char* e = getenv("PM_ROW_RMIN"); return e && e[0] == '0' ? 0 : 1; }();
  return v != 0;
}

bool sweep_row_reads_rowmajor(int w, int chunks, int ov) {
  return use_rm() && sweep_row_fuses_noise(w, chunks, ov) && row_rm_plan(w, chunks, ov);
}
bool sweep_row_interleaved(int w, int chunks, int ov);

struct Row2Args {
  const float2 *refT, *mat, *dcT_in;
  float2* dc_out;
  ViewGeom g;
  int pitchT;
  size_t planeT;
  int chunks, ov, max_walk;
  float alpha, w1;
  RowNoise nz;
  const float2* dc_rm;
  RowIL il;
};

// one instantiation: raises its dynamic shared-memory limit once per device, then launches
template <int DIR, bool NOISE, bool RMIN, bool SPEC, bool IL>
static bool row2_launch(dim3 grid, int threads, size_t bytes, cudaStream_t st, const Row2Args& a) {
  static size_t configured[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (bytes > configured[dev & 63]) {
    if (cudaFuncSetAttribute(k_sweep_row2<DIR, NOISE, RMIN, SPEC, IL>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess)
      return false;
    configured[dev & 63] = bytes;
  }
  k_sweep_row2<DIR, NOISE, RMIN, SPEC, IL><<<grid, threads, bytes, st>>>(
      a.refT, a.mat, a.dcT_in, a.dc_out, a.g, a.pitchT, a.planeT, a.chunks, a.ov, a.max_walk, a.alpha,
      a.w1, a.nz, a.dc_rm, a.il);
  return true;
}

template <bool NOISE, bool RMIN, bool SPEC, bool IL>
static bool row2_dir(int dir, dim3 grid, int threads, size_t bytes, cudaStream_t st, const Row2Args& a) {
  return dir > 0 ? row2_launch<1, NOISE, RMIN, SPEC, IL>(grid, threads, bytes, st, a)
                 : row2_launch<-1, NOISE, RMIN, SPEC, IL>(grid, threads, bytes, st, a);
}

template <int DIR, bool NOISE, bool SPEC, bool RMIN = false>
static bool row3_launch(dim3 grid, int threads, size_t bytes, cudaStream_t st, const Row2Args& a) {
  static size_t configured[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (bytes > configured[dev & 63]) {
    if (cudaFuncSetAttribute(k_sweep_row3<DIR, NOISE, SPEC, RMIN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)bytes) != cudaSuccess)
      return false;
    configured[dev & 63] = bytes;
  }
  k_sweep_row3<DIR, NOISE, SPEC, RMIN><<<grid, threads, bytes, st>>>(
      a.refT, a.mat, a.dcT_in, a.dc_out, a.g, a.pitchT, a.planeT, a.chunks, a.ov, a.max_walk, a.alpha,
      a.w1, a.nz, a.il, a.dc_rm);
  return true;
}

template <bool NOISE, bool SPEC, bool RMIN = false>
static bool row3_dir(int dir, dim3 grid, int threads, size_t bytes, cudaStream_t st, const Row2Args& a) {
  return dir > 0 ? row3_launch<1, NOISE, SPEC, RMIN>(grid, threads, bytes, st, a)
                 : row3_launch<-1, NOISE, SPEC, RMIN>(grid, threads, bytes, st, a);
}

static bool env_on(const char* name, bool dflt) {
  const char* e = getenv(name);
  if (!e || !e[0]) return dflt;
  return e[0] != '0';
}

bool sweep_row_interleaved(int w, int chunks, int ov) {
  static const bool on = env_on("PM_ROW_IL", true);
  return on && sweep_row_fuses_noise(w, chunks, ov);
}

int launch_sweep_row(const float2* refT, const float2* mat, const float2* dcT_in, float2* dc_out,
                     ViewGeom g, int pitchT, size_t planeT, int nviews, int dir, SweepParams sp,
                     cudaStream_t st, const float* noiseT, float noise_scale, float noise_dmax,
                     const float2* dc_rm, const float2* matI, size_t planeI) {
  int max_walk = 0;
  if (!sweep_block_plan(g.w, sp.chunks, sp.overlap, kRowBarrierStep, kPF, 32, &max_walk)) return -1;
  max_walk = (max_walk + kPF - 1) / kPF * kPF;
  const size_t bytes = sweep_row_smem_bytes(g.w, sp.chunks);
  // cudaFuncSetAttribute is per device: one high-water mark per device of this process
  static size_t configured_dev[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  size_t& configured = configured_dev[dev & 63];
  if (bytes > configured) {
    if (cudaFuncSetAttribute(k_sweep_row, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)bytes) != cudaSuccess) return -1;
    configured = bytes;
  }
  dim3 grid((g.h + kRows - 1) / kRows, nviews);
  int mw2 = 0;
  const size_t bytes2 = sweep_row2_smem_bytes(g.w, sp.chunks);
  if (!use_v1() && bytes2 + 64 <= (size_t)227 * 1024 &&
      sweep_block_plan(g.w, sp.chunks, sp.overlap, kRowBarrierStep, kRowP, 16, &mw2)) {
    const int mw_exact = mw2;
    mw2 = (mw2 + 15) / 16 * 16;
    const bool rm = dc_rm != nullptr, ilv = matI != nullptr;
    if (rm && !sweep_row_reads_rowmajor(g.w, sp.chunks, sp.overlap)) return -1;
    static const bool spec_on = env_on("PM_ROW_SPEC", false);
    const bool spec = spec_on && !rm;
    const Row2Args a{refT, mat, dcT_in, dc_out, g, pitchT, planeT, sp.chunks, sp.overlap, mw2, sp.alpha,
                     1 - sp.alpha, RowNoise{noiseT, noise_scale, noise_dmax, skip_eq_flag(), env_on("PM_ROW_EARLY_END", true) ? mw_exact : mw2}, dc_rm,
                     RowIL{matI, row_copy_elems(g.w), planeI}};
    const int th = 16 * sp.chunks;
    bool ok;
    static const bool gen3 = env_on("PM_ROW_GEN3", true);
    if (ilv && gen3 && rm) {
      ok = noiseT ? row3_dir<true, false, true>(dir, grid, th, bytes2, st, a)
                  : row3_dir<false, false, true>(dir, grid, th, bytes2, st, a);
      return ok && cudaGetLastError() == cudaSuccess ? 1 : -1;
    }
    if (ilv && gen3) {
      ok = noiseT ? (spec ? row3_dir<true, true>(dir, grid, th, bytes2, st, a)
                          : row3_dir<true, false>(dir, grid, th, bytes2, st, a))
                  : (spec ? row3_dir<false, true>(dir, grid, th, bytes2, st, a)
                          : row3_dir<false, false>(dir, grid, th, bytes2, st, a));
      return ok && cudaGetLastError() == cudaSuccess ? 1 : -1;
    }
    if (noiseT) {
      if (rm) ok = row2_dir<true, true, false, false>(dir, grid, th, bytes2, st, a);
      else if (ilv) ok = spec ? row2_dir<true, false, true, true>(dir, grid, th, bytes2, st, a)
                              : row2_dir<true, false, false, true>(dir, grid, th, bytes2, st, a);
      else ok = spec ? row2_dir<true, false, true, false>(dir, grid, th, bytes2, st, a)
                     : row2_dir<true, false, false, false>(dir, grid, th, bytes2, st, a);
    } else {
      if (rm) ok = row2_dir<false, true, false, false>(dir, grid, th, bytes2, st, a);
      else if (ilv) ok = spec ? row2_dir<false, false, true, true>(dir, grid, th, bytes2, st, a)
                              : row2_dir<false, false, false, true>(dir, grid, th, bytes2, st, a);
      else ok = spec ? row2_dir<false, false, true, false>(dir, grid, th, bytes2, st, a)
                     : row2_dir<false, false, false, false>(dir, grid, th, bytes2, st, a);
    }
    return ok && cudaGetLastError() == cudaSuccess ? 1 : -1;
  }
  if (noiseT || dc_rm) return -1;   // only the second-generation kernel fuses the noise / reads row-major
  k_sweep_row<<<grid, 16 * sp.chunks, bytes, st>>>(refT, mat, dcT_in, dc_out, g, pitchT, planeT, dir,
                                                   sp.chunks, sp.overlap, max_walk, sp.alpha,
                                                   1 - sp.alpha);
  return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// matched plane -> [view][row group of 16][column][16 rows] (RowIL); rows past the image stay zero
__global__ void __launch_bounds__(256)
k_interleave16(const float2* __restrict__ mat, ViewGeom g, float2* __restrict__ matI, int cols,
               size_t planeI) {
  __shared__ float2 tile[16][33];
  const int c0 = blockIdx.x * 32, grp = blockIdx.y, v = blockIdx.z;
  mat += (size_t)v * g.plane;
  matI += (size_t)v * planeI + (size_t)grp * cols * 16;
  const int tid = threadIdx.x;
  {
    const int cx = tid & 31, ry = tid >> 5;   // 32 columns x 8 rows, twice
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int rr = ry + 8 * k, y = grp * 16 + rr, c = c0 + cx;
      tile[rr][cx] = (y < g.h && c < cols) ? mat[(size_t)y * g.pitch + c] : make_float2(0.0f, 0.0f);
    }
  }
  __syncthreads();
  {
    const int rr = tid & 15, cl = tid >> 4;   // 16 rows x 16 columns, twice
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int c = c0 + cl + 16 * k;
      if (c < cols) matI[(size_t)c * 16 + rr] = tile[rr][cl + 16 * k];
    }
  }
}

size_t sweep_row_interleaved_plane(int w, int h) {
  return (size_t)((h + kRows - 1) / kRows) * row_copy_elems(w) * 16;
}

int sweep_row_interleaved_cols(int w) { return row_copy_elems(w); }

int launch_interleave16(const float2* mat, ViewGeom g, int nviews, float2* matI, cudaStream_t st) {
  const int cols = row_copy_elems(g.w);
  dim3 grid((cols + 31) / 32, (g.h + kRows - 1) / kRows, nviews);
  k_interleave16<<<grid, 256, 0, st>>>(mat, g, matI, cols, sweep_row_interleaved_plane(g.w, g.h));
  return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

bool sweep_row_supported(int w, int chunks, int ov) {
  int mw;
  return sweep_block_plan(w, chunks, ov, kRowBarrierStep, kPF, 32, &mw) &&
         sweep_row_smem_bytes(w, chunks) <= (size_t)227 * 1024;
}

// --------------------------------------------------------------- column sweep
//
// A column sweep walks along y; the lanes of a warp are 32 adjacent columns, so every
// plane is read and written row-major with coalesced accesses and the matched-image
// gathers of a warp land within a few neighbouring lines (L1). One block owns 32
// columns and all chunks of those columns (one warp per chunk).

// ROWT: the same kernel runs ROW sweeps on TRANSPOSED planes (refT, matT, dcT: [x][pitchT], rows
// contiguous): the lanes of a warp are 32 adjacent rows, the walk goes along x, every plane access
// is coalesced and a gather touches one 256-byte segment per distinct sample column in the warp
// (one or two once the disparity is locally smooth). It serves the widths the shared-memory row
// kernel cannot stage (w > 1330, e.g. 3840-wide frames and their row bands); `pitch`/`plane` are
// then pitchT/planeT and the result is the transposed {d, cost} plane.
template <bool ROWT>
__device__ __forceinline__ float cost5_lines(const RefTaps& L, const float2* __restrict__ mp, int pitch,
                                             float xr, float alpha, float w1) {
  if (!ROWT) return cost5_rows(L, mp - pitch, mp, mp + pitch, xr, alpha, w1);
  // mp = matT + y: element (row y + j, column c) is mp[c * pitch + j]
  int cc;
  float t, om;
  col_split(xr, cc, t, om);
  const float colp = __fadd_rn(xr, 1.0f);
  int cp = cc + 1;
  float tp = t, op = om;
  if (__fsub_rn(colp, 1.0f) != xr) col_split(colp, cp, tp, op);  // rare: xr+1 was rounded
  const float2* a = mp + (size_t)cc * pitch;
  const float2* b = mp + (size_t)cp * pitch;
  float cost = tap_term(L.tl, lerp2(__ldg(a - pitch - 1), __ldg(a - 1), t, om), alpha, w1);
  cost = __fadd_rn(cost, tap_term(L.tr, lerp2(__ldg(b - 1), __ldg(b + pitch - 1), tp, op), alpha, w1));
  cost = __fadd_rn(cost, tap_term(L.c, lerp2(__ldg(a), __ldg(a + pitch), t, om), alpha, w1));
  cost = __fadd_rn(cost, tap_term(L.bl, lerp2(__ldg(a - pitch + 1), __ldg(a + 1), t, om), alpha, w1));
  cost = __fadd_rn(cost, tap_term(L.br, lerp2(__ldg(b + 1), __ldg(b + pitch + 1), tp, op), alpha, w1));
  return cost;
}

#ifndef PM_COL_MINBLOCKS
#define PM_COL_MINBLOCKS 2
#endif
template <bool ROWT>
__global__ void __launch_bounds__(512, PM_COL_MINBLOCKS)
k_sweep_col(const float2* __restrict__ ref, const float2* __restrict__ mat,
            const float2* __restrict__ dc_in, float2* dc_out, ViewGeom g, int pitch, size_t plane,
            int dir, int chunks, int ov, int max_walk, int bar_step, float alpha, float w1,
            int skip_eq) {
  // nl lines of length len: columns walked along y, or (ROWT) rows walked along x
  const int nl = ROWT ? g.h : g.w, len = ROWT ? g.w : g.h;
  const int w = nl;  // name kept from the column form: the lane axis
  const int lane = threadIdx.x & 31, k = threadIdx.x >> 5;
  const int xs = blockIdx.x * 32 + lane, v = blockIdx.y;
  const bool valid = xs < w;           // lines past the image mirror the last one, no stores
  const int x = valid ? xs : w - 1;
  const size_t vo = (size_t)v * plane;
  ref += vo; mat += vo; dc_in += vo; dc_out += vo;
  // lines the reference sweeps (:134, :192)
  const bool active = valid && (ROWT ? row_interior(g, x) : (x >= 1 && x <= w - 2));
  const ChainGeom cg = chain_geom(k, chunks, len / chunks, ov, len, dir);

  // One 32-bit element offset per walk (a view's plane has far fewer than 2^31 elements) instead of
  // one 64-bit pointer per plane: the view's base pointers are block-uniform, the kernel lives at
  // its 64-register cap. fo: the position the ring fetches; oo: the position being evaluated.
  const int step_e = dir * pitch;
  int fo = cg.walk_first * pitch + x;
  int oo = fo;

  auto fetch = [&](Slot& s, int jj) {
    const bool inside = jj < cg.nwalk;
    const bool vis = inside && active && jj >= cg.vis_lo && jj < cg.vis_hi;
    if (inside) s.cur = (vis && jj >= cg.tail_lo) ? __ldcg(dc_out + fo) : dc_in[fo];
    if (vis) {
      const float2* ref_p = ref + fo;
      s.taps.tl = ref_p[-pitch - 1];
      if (ROWT) { s.taps.bl = ref_p[-pitch + 1]; s.taps.tr = ref_p[pitch - 1]; }
      else      { s.taps.tr = ref_p[-pitch + 1]; s.taps.bl = ref_p[pitch - 1]; }
      s.taps.c = ref_p[0];
      s.taps.br = ref_p[pitch + 1];
    }
    fo += step_e;
  };

  Slot ring[kPFCol];
#pragma unroll
  for (int u = 0; u < kPFCol; ++u) fetch(ring[u], u);

  float prev = dc_in[(size_t)(cg.start - dir) * pitch + x].x;
  // image column of the evaluated pixel: the lane (column sweep) or the walk position (ROWT)
  float xf = __int2float_rn(ROWT ? cg.walk_first : x);
  const float fdir = (float)dir;
  // ROWT: matT + y, the gathers index it by sample column; else the matched row of the step
  const float2* const matT_p = mat + x;

  for (int j0 = 0; j0 < max_walk; j0 += kPFCol) {
#pragma unroll
    for (int u = 0; u < kPFCol; ++u) {
      const int j = j0 + u;
      Slot& s = ring[u];
      float2 cur = s.cur;
      if (active && j >= cg.vis_lo && j < cg.vis_hi) {
        // A candidate bit-equal to the pixel's own disparity has the cached cost: the strict
        // `<` (patchmatch_gpu.cu:168) fails whatever it is, so the lane skips its ten gathers;
        // a warp whose lanes all agree skips the evaluation altogether.
        if (!skip_eq || prev != cur.x) {
          const float xr = fmaxf(__fsub_rn(xf, prev), 1.0f);
          const float2* mat_p = ROWT ? matT_p : mat + (oo - x);
          const float c1 = cost5_lines<ROWT>(s.taps, mat_p, pitch, xr, alpha, w1);
          if (c1 < cur.y) {
            cur.x = fminf(prev, __fsub_rn(xf, 1.0f));
            cur.y = c1;
          }
        }
        prev = cur.x;
      }
      if (valid && j < cg.nwalk) dc_out[oo] = cur;
      fetch(s, j + kPFCol);
      // the matched row two steps ahead is first touched on the dependent chain: pull the
      // lines this warp can reach (its 32 columns and kColPrefetchDisp px to their left)
      // into L1 now; one instruction per warp and step
      if (!ROWT) {
        const int pc = blockIdx.x * 32 - kColPrefetchDisp + 16 * lane;
        if (lane < (kColPrefetchDisp + 48) / 16 && pc >= 0 && pc < w && j + 2 < cg.nwalk)
          asm volatile("prefetch.global.L1 [%0];" ::"l"(mat + (oo - x) + 2 * step_e + pc));
        if (kColCurPrefetch > 0 && j + kPFCol + kColCurPrefetch < cg.nwalk)
        {
          asm volatile("prefetch.global.L1 [%0];" ::"l"(dc_in + fo + kColCurPrefetch * step_e));
          if (kColRefPrefetch) asm volatile("prefetch.global.L1 [%0];" ::"l"(ref + fo + (kColCurPrefetch + 1) * step_e));
        }
        // further ahead, into L2 only (L1 cannot hold more rows of 32 warps): the lines the register
        // ring and the L1 prefetch above will ask for kColL2Prefetch steps from now
        if (kColL2Prefetch > 0 && j + kColL2Prefetch + 1 < cg.nwalk) {
          if (lane < (kColPrefetchDisp + 48) / 16 && pc >= 0 && pc < w)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(mat + (oo - x) + (kColL2Prefetch + 1) * step_e + pc));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(dc_in + fo + (kColL2Prefetch - kPFCol) * step_e));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(ref + fo + (kColL2Prefetch - kPFCol + 1) * step_e));
        }
      } else {
        // the sample column kRowTPrefetch steps ahead, if the disparity stays what it is: its rows
        // (this lane's; lanes 0 and 31 take the rows just outside the warp) go to L1 now, so that
        // the gather that first touches them does not wait for DRAM on the dependent chain
        if (active && j + kRowTPrefetch + 1 < cg.nwalk) {  // +1: the reference row one step further
          int pc = __float2int_rd(__fsub_rn(xf, prev)) + dir * kRowTPrefetch;
          pc = min(max(pc, 0), len);
          const int dy = lane == 0 ? -1 : (lane == 31 ? 1 : 0);
          asm volatile("prefetch.global.L1 [%0];" ::"l"(matT_p + (size_t)pc * pitch + dy));
          // the {d, cost} and reference rows of that step too: the register ring runs only
          // kPFCol steps ahead, which hides DRAM latency with many warps per SM, not with few
          asm volatile("prefetch.global.L1 [%0];" ::"l"(dc_in + fo + (kRowTPrefetch - kPFCol) * step_e));
          asm volatile("prefetch.global.L1 [%0];" ::"l"(ref + fo + (kRowTPrefetch - kPFCol + 1) * step_e + dy));
        }
        xf = __fadd_rn(xf, fdir);
      }
      oo += step_e;
      if (j == bar_step) __syncthreads();  // heads are stored: successors may read them
    }
  }
}

// ------------------------------------------------- column sweep, third generation
//
// The same schedule, block shape and arithmetic as k_sweep_col with the per-step overhead taken
// out (the first kernel spends ~120 of its ~250 issue slots per step on 64-bit address arithmetic
// re-materialised under the 64-register cap):
//   * one running 64-bit pointer per stream ({d, cost} in, out, reference row, matched row), each
//     advanced with one IMAD.WIDE per step; every load and store is pointer + immediate;
//   * the reference taps roll: the row ahead of step j is the row behind step j+2, and its centre
//     element is the centre tap of step j+1, so a step loads three adjacent elements of ONE row
//     (x-1, x, x+1) instead of five taps of three rows;
//   * DIR is a template parameter, the trip count is the warp's own walk length, the packed
//     f32x2 pipe lerps and differences {I, G} pairs, floor() is one FADD.RM;
//   * WPC warps per chunk: a block owns 32*WPC adjacent columns, so the matched-image lines a
//     block pulls into L1 are shared by more lanes (reach D + 32*WPC columns per row).
template <int DIR>
__device__ __forceinline__ float cost5_ptr(float2 tl, float2 tr, float2 c, float2 bl, float2 br,
                                           const char* mrow, long long pitchB, float xr, float alpha,
                                           float w1) {
  int cc;
  float t, om;
  col_split_rd(xr, cc, t, om);
#ifdef PM_C3_FAKE   // diagnosis only: every gather lands in the same few lines
  cc = (cc & 15) + 1;
#endif
  const float colp = __fadd_rn(xr, 1.0f);
  const char* a1 = mrow + (long long)cc * 8;
  const char* a0 = a1 - pitchB;
  const char* a2 = a1 + pitchB;
  auto ld = [](const char* p, int e) { return __ldg((const float2*)p + e); };
  float2 mtr, mbr;
  if (__fsub_rn(colp, 1.0f) != xr) {  // rare: xr+1 was rounded, split it like the reference does
    int cp;
    float tp, op;
    col_split_rd(colp, cp, tp, op);
    const int e = cp - cc;
    mtr = lerp2p(ld(a0, e), ld(a0, e + 1), tp, op);
    mbr = lerp2p(ld(a2, e), ld(a2, e + 1), tp, op);
  } else {
    mtr = lerp2p(ld(a0, 1), ld(a0, 2), t, om);
    mbr = lerp2p(ld(a2, 1), ld(a2, 2), t, om);
  }
  float cost = tap_term_p(tl, lerp2p(ld(a0, -1), ld(a0, 0), t, om), alpha, w1);
  cost = __fadd_rn(cost, tap_term_p(tr, mtr, alpha, w1));
  cost = __fadd_rn(cost, tap_term_p(c, lerp2p(ld(a1, 0), ld(a1, 1), t, om), alpha, w1));
  cost = __fadd_rn(cost, tap_term_p(bl, lerp2p(ld(a2, -1), ld(a2, 0), t, om), alpha, w1));
  cost = __fadd_rn(cost, tap_term_p(br, mbr, alpha, w1));
  return cost;
}

// The same evaluation on a TRANSPOSED matched plane (row sweeps of wide frames): `mlane` points at
// (column 0, this lane's row y); element (row y + j, column c) is mlane[c * pitch + j].
__device__ __forceinline__ float cost5_ptrT(float2 tl, float2 tr, float2 c, float2 bl, float2 br,
                                            const char* mlane, int pitchB, float xr, float alpha,
                                            float w1) {
  int cc;
  float t, om;
  col_split_rd(xr, cc, t, om);
  const float colp = __fadd_rn(xr, 1.0f);
  const char* a = mlane + (long long)cc * pitchB;   // column cc
  const char* am = a - pitchB;                      // column cc - 1
  auto ld = [](const char* p, int e) { return __ldg((const float2*)p + e); };
  float2 mtr, mbr;
  if (__fsub_rn(colp, 1.0f) != xr) {  // rare: xr+1 was rounded, split it like the reference does
    int cp;
    float tp, op;
    col_split_rd(colp, cp, tp, op);
    const char* b = mlane + (long long)cp * pitchB;
    const char* bp = b + pitchB;
    mtr = lerp2p(ld(b, -1), ld(bp, -1), tp, op);
    mbr = lerp2p(ld(b, 1), ld(bp, 1), tp, op);
  } else {
    const char* b = a + pitchB;
    const char* bp = b + pitchB;
    mtr = lerp2p(ld(b, -1), ld(bp, -1), t, om);
    mbr = lerp2p(ld(b, 1), ld(bp, 1), t, om);
  }
  float cost = tap_term_p(tl, lerp2p(ld(am, -1), ld(a, -1), t, om), alpha, w1);
  cost = __fadd_rn(cost, tap_term_p(tr, mtr, alpha, w1));
  cost = __fadd_rn(cost, tap_term_p(c, lerp2p(ld(a, 0), ld(a + pitchB, 0), t, om), alpha, w1));
  cost = __fadd_rn(cost, tap_term_p(bl, lerp2p(ld(am, 1), ld(a, 1), t, om), alpha, w1));
  cost = __fadd_rn(cost, tap_term_p(br, mbr, alpha, w1));
  return cost;
}

struct LeadTaps { float2 l, r; };

constexpr int kCol3Unroll = 4;
#ifndef PM_C3_CEN
#define PM_C3_CEN 1
#endif
#ifndef PM_C3_PF
#define PM_C3_PF 1
#endif
#ifndef PM_C3_PFD
#define PM_C3_PFD 2
#endif
#ifndef PM_C3_AHEAD
#define PM_C3_AHEAD 4
#endif
#ifndef PM_C3_PFSECT
#define PM_C3_PFSECT 1
#endif
#ifndef PM_C3_PFSELF
#define PM_C3_PFSELF 0
#endif

// Geometry of one chain for the in-place kernel: only the evaluated steps are walked (the rows and
// columns the reference skips simply keep their values in place).
struct Chain3 { int start, nsteps, tail_lo; };
__host__ __device__ inline Chain3 chain3(int k, int chunks, int cs, int ov, int len, int dir) {
  const ChainGeom c = chain_geom(k, chunks, cs, ov, len, dir);
  Chain3 r;
  r.start = c.start;
  r.nsteps = c.vis_hi - c.vis_lo;
  r.tail_lo = c.tail_lo == INT_MAX ? INT_MAX : c.tail_lo - c.vis_lo;
  return r;
}

// steps before the block barrier: every head (2*ov steps) is stored by then
__host__ __device__ inline int col3_bar_steps(int ov) {
  return (2 * ov + kCol3Unroll - 1) / kCol3Unroll * kCol3Unroll;
}

// IN PLACE on the {d, cost} plane `dc`: under the lock-step schedule a chunk reads every position
// before it writes it, reads its successor's head after the barrier (which is the handover) and
// nothing else of the plane changes under it, so source and destination can be one plane; {d, cost}
// loads go to L2 (ld.cg), where the stores of the other warps of the block are visible after the
// barrier.
// ROWT: the same kernel runs ROW sweeps in place on TRANSPOSED planes (refT, matT, dcT: [x][pitchT],
// rows contiguous; `pitch`/`plane` are then pitchT/planeT): the lanes of a warp are 32 adjacent rows, the
// walk goes along x, every plane access is coalesced and a gather touches one 256-byte segment per
// distinct sample column in the warp. It serves the widths the shared-memory row kernel cannot
// stage (w > 1330: 3840-wide frames and their row bands).
template <int DIR, int WPC, bool ROWT>
__global__ void __launch_bounds__(512 * WPC, ROWT ? 1 : 2 / WPC)   // ROWT: few blocks, no register cap
k_sweep_col3(const float2* __restrict__ ref, const float2* __restrict__ mat, float2* dc, ViewGeom g,
             int pitch, size_t plane, int chunks, int ov, float alpha, float w1) {
  // nl lines of length len: columns walked along y, or (ROWT) rows walked along x
  const int w = ROWT ? g.h : g.w, len = ROWT ? g.w : g.h;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int k = warp / WPC, sub = warp % WPC;
  const int x0 = blockIdx.x * (32 * WPC);
  const int xs = x0 + sub * 32 + lane;
  // lines the reference sweeps (:134, :192); the others mirror one, without stores
  const bool active = ROWT ? (xs < w && row_interior(g, xs)) : (xs >= 1 && xs <= w - 2);
  const int x = min(max(xs, 1), w - 2);
  const size_t vo = (size_t)blockIdx.y * plane;
  const Chain3 cg = chain3(k, chunks, len / chunks, ov, len, DIR);
  const long long pitchB = (long long)pitch * 8;
  const long long stepB = DIR * pitchB;

  // fetch jj loads the reference row one ahead of walk position jj (start + DIR*(jj+1)) and the
  // {d, cost} element of position jj; both run two steps ahead of the evaluation
  const size_t e0 = vo + (size_t)cg.start * pitch + x;
  const char* p_ref = (const char*)(ref + e0) - stepB;   // fetch -2
  const char* p_cur = (const char*)(dc + e0);
  char* p_out = (char*)(dc + e0);
  // column sweep: the matched row of the step; ROWT: this lane's row of the transposed plane
  const char* p_mat = ROWT ? (const char*)(mat + vo + x)
                           : (const char*)(mat + vo + (size_t)cg.start * pitch);
  // matched lines this block can reach two steps from now go to L1 (one instruction per step and
  // chunk): lane i takes the 128-byte line holding column x0 - kColPrefetchDisp + 16 i
  // one warp per chunk: lane i takes the 128-byte line holding column x0 - kColPrefetchDisp + 16 i;
  // two warps per chunk: one prefetch per 32-byte SECTOR (4 elements) spread over the 64 lanes
  // (measured: column sweeps 8.16 -> 8.02 ms per 64 pairs)
  const bool sect = PM_C3_PFSECT && WPC == 2;
  const int pc = x0 - kColPrefetchDisp + (sect ? 4 * (lane + 32 * sub) : 16 * lane);
  const bool pf_lane = !ROWT && pc >= 0 && pc < w &&
      (sect ? pc < x0 + 32 * WPC + 4 : (sub == 0 && lane < (kColPrefetchDisp + 32 * WPC + 16) / 16));
  const char* p_pf = p_mat + PM_C3_PFD * stepB + (long long)(pf_lane ? pc : 0) * 8;
  // ROWT: lanes 0 and 31 prefetch the rows just outside the warp
  const int pf_dyB = ROWT ? (lane == 0 ? -8 : (lane == 31 ? 8 : 0)) : 0;

  LeadTaps L[4];
#if PM_C3_CEN
  float2 Cn[2];   // centre taps of steps j, j+1 (mod 2), loaded from the row behind the lead row
#else
  float2 Cn[4];
#endif
  float2 curv[2];

  auto fetch_ref = [&](int slot, bool on) {
    if (on) {
      const float2* q = (const float2*)p_ref;
      L[slot].l = __ldg(q - 1);
#if PM_C3_CEN
      Cn[slot & 1] = __ldg((const float2*)(p_ref - stepB));   // row of walk position `slot`
#else
      Cn[slot] = __ldg(q);
#endif
      L[slot].r = __ldg(q + 1);
    }
    p_ref += stepB;
  };
  auto fetch_cur = [&](int slot, bool on) {
    if (on) curv[slot] = __ldcg((const float2*)p_cur);
    p_cur += stepB;
  };
  int rem = cg.nsteps;   // steps left, counted down
#if PM_C3_CEN
  {
    const float2* q = (const float2*)p_ref;
    L[2].l = __ldg(q - 1);
    L[2].r = __ldg(q + 1);
    p_ref += stepB;
    q = (const float2*)p_ref;
    L[3].l = __ldg(q - 1);
    L[3].r = __ldg(q + 1);
    p_ref += stepB;
  }
#else
  fetch_ref(2, true);
  fetch_ref(3, true);
#endif
  fetch_ref(0, true);
  fetch_ref(1, rem > 1);
  fetch_cur(0, true);
  fetch_cur(1, rem > 1);
  float prev = __ldcg(dc + vo + (size_t)(cg.start - DIR) * pitch + x).x;
  // image column of the evaluated pixel: the lane (column sweep) or the walk position (ROWT)
  float xf = __int2float_rn(ROWT ? cg.start : x);
  const int rem_bar = cg.nsteps - col3_bar_steps(ov);

  while (rem > 0) {
#pragma unroll
    for (int u = 0; u < kCol3Unroll; ++u) {
      float2 cur = curv[u & 1];
      // a candidate bit-equal to the pixel's own disparity has the cached cost: the strict `<`
      // (patchmatch_gpu.cu:168) fails whatever it is
      if (rem > u && active && prev != cur.x) {
        const float xr = fmaxf(__fsub_rn(xf, prev), 1.0f);
        const LeadTaps& lead = L[u];
        const LeadTaps& trail = L[(u + 2) & 3];
#if PM_C3_CEN
        const float2 cen = Cn[u & 1];
#else
        const float2 cen = Cn[(u + 3) & 3];
#endif
        // lead/trail: the line ahead of / behind the walk; .l/.r: its elements before / after the lane
        float c1;
        if (ROWT)
          c1 = DIR > 0
              ? cost5_ptrT(trail.l, lead.l, cen, trail.r, lead.r, p_mat, (int)pitchB, xr, alpha, w1)
              : cost5_ptrT(lead.l, trail.l, cen, lead.r, trail.r, p_mat, (int)pitchB, xr, alpha, w1);
        else
          c1 = DIR > 0
              ? cost5_ptr<DIR>(trail.l, trail.r, cen, lead.l, lead.r, p_mat, pitchB, xr, alpha, w1)
              : cost5_ptr<DIR>(lead.l, lead.r, cen, trail.l, trail.r, p_mat, pitchB, xr, alpha, w1);
        if (c1 < cur.y) {
          cur.x = fminf(prev, __fsub_rn(xf, 1.0f));
          cur.y = c1;
          *(float2*)p_out = cur;
        }
      }
      if (rem > u) prev = cur.x;
      const bool more = rem > u + 2;
#if PM_C3_PFSELF
      // every lane pulls the sector its own next candidate would sample (odd lanes: the upper end)
      if (!ROWT && active && rem > u + PM_C3_PFD) {
        const int pcol = max(__float2int_rd(__fsub_rn(xf, prev)), 1) + ((lane & 1) ? 2 : -1);
        asm volatile("prefetch.global.L1 [%0];" ::"l"(p_mat + PM_C3_PFD * stepB + (long long)pcol * 8));
      }
#elif PM_C3_PF
      if (pf_lane && rem > u + PM_C3_PFD) asm volatile("prefetch.global.L1 [%0];" ::"l"(p_pf));
#endif
#if PM_C3_AHEAD > 0
      // The register ring is two steps deep, but ptxas shares its scoreboard with a gather, so a
      // step ends up waiting for the ring loads of the step before: pull their lines close first
      // (reference row -> L1, {d, cost} -> L2, where ld.cg reads it), then the ring loads are hits.
      if (rem > u + 2 + PM_C3_AHEAD) {
        asm volatile("prefetch.global.L1 [%0];" ::"l"(p_ref + PM_C3_AHEAD * stepB));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p_cur + PM_C3_AHEAD * stepB));
      }
#endif
      fetch_ref((u + 2) & 3, more);
      fetch_cur(u & 1, more);
      p_out += stepB;
      if (ROWT) {
        // the sample column kRowTPrefetch steps ahead, if the disparity stays what it is: its rows go
        // to L1 now, so that the gather that first touches them does not wait for DRAM on the chain
        if (active && rem > u + kRowTPrefetch + 1) {
          int pcol = __float2int_rd(__fsub_rn(xf, prev)) + DIR * kRowTPrefetch;
          pcol = min(max(pcol, 0), len);
          asm volatile("prefetch.global.L1 [%0];" ::"l"(p_mat + (long long)pcol * (int)pitchB + pf_dyB));
        }
        xf = __fadd_rn(xf, (float)DIR);
      } else {
        p_mat += stepB;
        p_pf += stepB;
      }
    }
    rem -= kCol3Unroll;
    if (rem == rem_bar) __syncthreads();  // heads are stored: successors may read them
  }
}

// ---- fourth generation: the column kernel with the image rows in shared memory -------------------
// Same schedule, block shape (64 adjacent columns x all chunks, two warps per chunk), arithmetic and
// in-place {d, cost} plane as k_sweep_col3<DIR, 2, false>. What changes is where the image elements
// of an evaluation come from: every chunk keeps a ring of kC4Ring rows in shared memory - the matched
// plane's columns [x0 - kC4Reach, x0 + 72) and the reference plane's columns [x0 - 2, x0 + 66). One
// lane per chunk fills it with TMA bulk copies (cp.async.bulk, completion counted on an mbarrier per
// slot) three steps ahead of the walk; the gathers are ld.shared (two wavefronts for 32 adjacent
// columns whatever their alignment, a fixed latency on the chain) instead of L1 hits-or-misses
// behind a 65 %-busy L1 wavefront pipe, and the L1 prefetches and the register ring of reference
// taps disappear. A slot is handed back through a second mbarrier (the chunk's other warp arrives
// when it has read the row, the filling lane waits for it). The ring length equals the unroll
// factor, so every slot address is an immediate. A candidate whose sample column lies left of the
// staged window (d > kC4Reach - 2, possible because the reference does not bound d) takes the
// global-memory evaluation of the third generation.
constexpr int kC4Reach = 144;                 // matched columns staged to the left of the block (even)
constexpr int kC4W = kC4Reach + 72;           // staged matched row, elements
constexpr int kC4RefW = 68;                   // staged reference row, elements
constexpr int kC4Ring = 6;                    // rows per chunk: 3 live + 3 in flight; = unroll factor
constexpr unsigned kC4MatB = kC4W * 8u, kC4RefB = kC4RefW * 8u;
constexpr unsigned kC4SlotB = kC4MatB + kC4RefB;
constexpr unsigned kC4BarB = 2048;            // 16 chunks x 8 x {full, empty} x 8 bytes
#ifndef PM_C4_L2PF
#define PM_C4_L2PF 0
#endif

__host__ __device__ inline size_t col4_smem_bytes(int chunks) {
  return kC4BarB + (size_t)chunks * kC4Ring * kC4SlotB;
}
__host__ __device__ inline int col4_bar_steps(int ov) {
  return (2 * ov + kC4Ring - 1) / kC4Ring * kC4Ring;
}

__device__ __forceinline__ void mbar_arrive(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

template <int DIR>
__device__ __noinline__ float cost5_glob(float2 tl, float2 tr, float2 c, float2 bl, float2 br,
                                         const char* mrow, long long pitchB, float xr, float alpha,
                                         float w1) {
  return cost5_ptr<DIR>(tl, tr, c, bl, br, mrow, pitchB, xr, alpha, w1);
}

// s0, s1, s2: shared addresses of (image) column 0 of the matched rows y-1, y, y+1
__device__ __forceinline__ float cost5_lds(float2 tl, float2 tr, float2 c, float2 bl, float2 br,
                                           unsigned s0, unsigned s1, unsigned s2, int cc, float t,
                                           float om, float xr, float alpha, float w1) {
  const float colp = __fadd_rn(xr, 1.0f);
  const unsigned o = (unsigned)cc * 8u;
  float2 a0, a1, a2, a3, c0, c1, b0, b1, b2, b3;
  lds_f2(a0, s0 + o - 8u); lds_f2(a1, s0 + o);
  lds_f2(c0, s1 + o);      lds_f2(c1, s1 + o + 8u);
  lds_f2(b0, s2 + o - 8u); lds_f2(b1, s2 + o);
  float2 mtr, mbr;
  if (__fsub_rn(colp, 1.0f) != xr) {  // rare: xr+1 was rounded, split it like the reference does
    int cp;
    float tp, op;
    col_split_rd(colp, cp, tp, op);
    const unsigned q = (unsigned)cp * 8u;
    lds_f2(a2, s0 + q); lds_f2(a3, s0 + q + 8u);
    lds_f2(b2, s2 + q); lds_f2(b3, s2 + q + 8u);
    mtr = lerp2p(a2, a3, tp, op);
    mbr = lerp2p(b2, b3, tp, op);
  } else {
    lds_f2(a2, s0 + o + 8u); lds_f2(a3, s0 + o + 16u);
    lds_f2(b2, s2 + o + 8u); lds_f2(b3, s2 + o + 16u);
    mtr = lerp2p(a2, a3, t, om);
    mbr = lerp2p(b2, b3, t, om);
  }
  float cost = tap_term_p(tl, lerp2p(a0, a1, t, om), alpha, w1);
  cost = __fadd_rn(cost, tap_term_p(tr, mtr, alpha, w1));
  cost = __fadd_rn(cost, tap_term_p(c, lerp2p(c0, c1, t, om), alpha, w1));
  cost = __fadd_rn(cost, tap_term_p(bl, lerp2p(b0, b1, t, om), alpha, w1));
  cost = __fadd_rn(cost, tap_term_p(br, mbr, alpha, w1));
  return cost;
}

template <int DIR>
__global__ void __launch_bounds__(1024, 1)
k_sweep_col4(const float2* __restrict__ ref, const float2* __restrict__ mat, float2* dc, ViewGeom g,
             int pitch, size_t plane, int chunks, int ov, float alpha, float w1) {
  extern __shared__ __align__(128) unsigned char c4_smem[];
  const int w = g.w, len = g.h;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int k = warp >> 1, sub = warp & 1;
  const int x0 = blockIdx.x * 64;
  const int xs = x0 + sub * 32 + lane;
  const bool active = xs >= 1 && xs <= w - 2;
  const int x = min(max(xs, 1), w - 2);
  const size_t vo = (size_t)blockIdx.y * plane;
  const Chain3 cg = chain3(k, chunks, len / chunks, ov, len, DIR);
  const long long pitchB = (long long)pitch * 8;
  const long long stepB = DIR * pitchB;

  // ring index i holds row cg.start + DIR * (i - 1): step j reads indices j (the row behind the
  // walk), j + 1 and j + 2 (the row ahead); index i lives in slot i % 6 with mbarrier parity (i / 6) & 1
  const int lo = max(0, x0 - kC4Reach), rlo = max(0, x0 - 2);
  const unsigned mbytes = (unsigned)min(kC4W, pitch - lo) * 8u;
  const unsigned rbytes = (unsigned)min(kC4RefW, pitch - rlo) * 8u;
  const unsigned sb = (unsigned)__cvta_generic_to_shared(c4_smem);
  const unsigned full0 = sb + (unsigned)k * 64u, empty0 = sb + 1024u + (unsigned)k * 64u;
  const unsigned rows0 = sb + kC4BarB + (unsigned)k * (kC4Ring * kC4SlotB);
  const bool producer = sub == 0 && lane == 0;
  if (producer) {
    for (int s = 0; s < kC4Ring; ++s) { mbar_init(full0 + 8u * s, 1); mbar_init(empty0 + 8u * s, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // running offset (bytes) of the next row to stage, relative to the planes' bases
  long long fill_off = (long long)(vo + (size_t)(cg.start - DIR) * pitch) * 8;
  const char* mat_b = (const char*)mat + (long long)lo * 8;
  const char* ref_b = (const char*)ref + (long long)rlo * 8;
  auto stage = [&](unsigned sl) {
    mbar_expect_tx(full0 + 8u * sl, mbytes + rbytes);
    tma_load_1d(rows0 + kC4SlotB * sl, mat_b + fill_off, mbytes, full0 + 8u * sl);
    tma_load_1d(rows0 + kC4SlotB * sl + kC4MatB, ref_b + fill_off, rbytes, full0 + 8u * sl);
  };
  if (producer) {
    for (int i = 0; i < kC4Ring; ++i) {   // every chain has at least 6 rows (>= 12 steps)
      stage(i);
      fill_off += stepB;
    }
  } else {
    fill_off += kC4Ring * stepB;
  }

  const size_t e0 = vo + (size_t)cg.start * pitch + x;
  const char* p_cur = (const char*)(dc + e0);
  char* p_out = (char*)(dc + e0);
  float2 curv[2];
  auto fetch_cur = [&](int slot, bool on) {
    if (on) curv[slot] = __ldcg((const float2*)p_cur);
    p_cur += stepB;
  };
  int rem = cg.nsteps;
  fetch_cur(0, true);
  fetch_cur(1, rem > 1);
  float prev = __ldcg(dc + vo + (size_t)(cg.start - DIR) * pitch + x).x;
  const float xf = __int2float_rn(x);
  const int rem_bar = cg.nsteps - col4_bar_steps(ov);
  // shared addresses, slot 0: matched column 0, and this lane's reference column
  const unsigned lbase = rows0 - (unsigned)lo * 8u;
  const unsigned rbase = rows0 + kC4MatB + (unsigned)(x - rlo) * 8u;

  mbar_wait(full0, 0);
  mbar_wait(full0 + 8u, 0);
  unsigned ph = 0;   // parity of the ring indices jj .. jj + 5 (jj = steps done, a multiple of 6)
  while (rem > 0) {
#pragma unroll
    for (int u = 0; u < kC4Ring; ++u) {
      constexpr int R = kC4Ring;
      const unsigned slT = u % R, slC = (u + 1) % R, slL = (u + 2) % R;
      float2 cur = curv[u & 1];
      if (rem > u) mbar_wait(full0 + 8u * slL, u + 2 < R ? ph : ph ^ 1u);   // the row ahead has landed
      if (rem > u && active && prev != cur.x) {
        const float xr = fmaxf(__fsub_rn(xf, prev), 1.0f);
        // reference taps: rows behind / at / ahead of the walk, columns x-1, x, x+1
        float2 tl_, tr_, cen, ll_, lr_;
        lds_f2(tl_, rbase + kC4SlotB * slT - 8u); lds_f2(tr_, rbase + kC4SlotB * slT + 8u);
        lds_f2(cen, rbase + kC4SlotB * slC);
        lds_f2(ll_, rbase + kC4SlotB * slL - 8u); lds_f2(lr_, rbase + kC4SlotB * slL + 8u);
        int cc;
        float t, om;
        col_split_rd(xr, cc, t, om);
        float c1;
        if (cc > lo) {
          const unsigned sT = lbase + kC4SlotB * slT, sC = lbase + kC4SlotB * slC,
                         sL = lbase + kC4SlotB * slL;
          c1 = DIR > 0
              ? cost5_lds(tl_, tr_, cen, ll_, lr_, sT, sC, sL, cc, t, om, xr, alpha, w1)
              : cost5_lds(ll_, lr_, cen, tl_, tr_, sL, sC, sT, cc, t, om, xr, alpha, w1);
        } else {   // left of the staged window
          const int y = cg.start + DIR * (cg.nsteps - rem + u);
          const char* mrow = (const char*)(mat + vo + (size_t)y * pitch);
          c1 = DIR > 0
              ? cost5_glob<DIR>(tl_, tr_, cen, ll_, lr_, mrow, pitchB, xr, alpha, w1)
              : cost5_glob<DIR>(ll_, lr_, cen, tl_, tr_, mrow, pitchB, xr, alpha, w1);
        }
        if (c1 < cur.y) {
          cur.x = fminf(prev, __fsub_rn(xf, 1.0f));
          cur.y = c1;
          *(float2*)p_out = cur;
        }
      }
      if (rem > u) {
        prev = cur.x;
        // ring index jj + u is done with: hand its slot back and refill it with index jj + u + 6
        __syncwarp();
        if (lane == 0) {
          if (sub) {
            mbar_arrive(empty0 + 8u * slT);
          } else if (rem > u + R - 2) {
            mbar_wait(empty0 + 8u * slT, ph);
            stage(slT);
#if PM_C4_L2PF > 0
            // the rows PM_C4_L2PF steps further on: into L2 now, so that their TMA copy is an L2 hit
            if (rem > u + R - 2 + PM_C4_L2PF) {
              asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(mat_b + fill_off + PM_C4_L2PF * stepB), "r"(mbytes) : "memory");
              asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ref_b + fill_off + PM_C4_L2PF * stepB), "r"(rbytes) : "memory");
            }
#endif
          }
        }
      }
      fill_off += stepB;
      const bool more = rem > u + 2;
      if (rem > u + 2 + 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(p_cur + 4 * stepB));
      fetch_cur(u & 1, more);
      p_out += stepB;
    }
    rem -= kC4Ring;
    ph ^= 1u;
    if (rem == rem_bar) __syncthreads();  // heads are stored: successors may read them
  }
}

// every head (at most 2*ov steps) is stored directly, so the barrier can come right after
static int col_bar_step(int ov) { return 2 * ov > 0 ? 2 * ov - 1 : 0; }

int launch_sweep_col(const float2* ref, const float2* mat, const float2* dc_in, float2* dc_out,
                     ViewGeom g, int nviews, int dir, SweepParams sp, cudaStream_t st) {
  int max_walk = 0;
  const int bar_step = col_bar_step(sp.overlap);
  if (!sweep_block_plan(g.h, sp.chunks, sp.overlap, bar_step, kPFCol, 16, &max_walk)) return -1;
  max_walk = (max_walk + kPFCol - 1) / kPFCol * kPFCol;
  dim3 grid((g.w + 31) / 32, nviews);
  k_sweep_col<false><<<grid, 32 * sp.chunks, 0, st>>>(ref, mat, dc_in, dc_out, g, g.pitch, g.plane,
                                                      dir, sp.chunks, sp.overlap, max_walk, bar_step,
                                                      sp.alpha, 1 - sp.alpha, skip_eq_flag());
  return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

static int col3_mode() {   // 0: off, 1: 32 columns per block, 2: 64 columns per block, 3: by grid size
  static const int v = [] { const char* e = getenv("PM_COL_V3"); return e ? atoi(e) : 3; }();
  return v;
}

static bool col4_on() {   // fourth-generation column kernel (shared-memory matched rows)
  static const bool v = [] { const char* e = getenv("PM_COL_V4"); return e && e[0] == '1'; }();
  return v;
}

static bool col4_supported(int len, int chunks, int ov) {
  if (chunks < 2 || chunks > 16 || ov > 8) return false;
  const int cs = len / chunks, nb = col4_bar_steps(ov);
  for (int dir = -1; dir <= 1; dir += 2)
    for (int k = 0; k < chunks; ++k) {
      const Chain3 c = chain3(k, chunks, cs, ov, len, dir);
      if (c.nsteps < nb || c.nsteps < kC4Ring) return false;
      if (c.tail_lo != INT_MAX && c.tail_lo - 2 < nb) return false;
    }
  return true;
}

// nl lines of length len, each cut into `chunks` chunks
bool sweep_col_inplace_supported(int nl, int len, int chunks, int ov) {
  if (col3_mode() <= 0 || nl < 3 || chunks < 2 || chunks > 16 || ov > 8) return false;
  const int cs = len / chunks, nb = col3_bar_steps(ov);
  for (int dir = -1; dir <= 1; dir += 2)
    for (int k = 0; k < chunks; ++k) {
      const Chain3 c = chain3(k, chunks, cs, ov, len, dir);
      // every chunk reaches the barrier, and its first handover read (fetched two steps early)
      // comes after it
      if (c.nsteps < nb || c.nsteps < 2) return false;
      if (c.tail_lo != INT_MAX && c.tail_lo - 2 < nb) return false;
    }
  return true;
}

template <bool ROWT>
static int launch_col3(const float2* ref, const float2* mat, float2* dc, ViewGeom g, int pitch,
                       size_t plane, int nviews, int dir, SweepParams sp, cudaStream_t st) {
  const int nl = ROWT ? g.h : g.w, len = ROWT ? g.w : g.h;
  if (g.cost_mode != 0 || g.radius != 1 || !sweep_col_inplace_supported(nl, len, sp.chunks, sp.overlap))
    return -1;
  // 64 lines per block (one block of 1024 threads per SM) share more of the matched lines in L1;
  // launches that would not fill the GPU twice over keep 32 lines per block
  int wpc = col3_mode() == 1 ? 1 : 2;
  if (col3_mode() != 2 && (long)((nl + 63) / 64) * nviews < 2 * 148) wpc = 1;
  if (ROWT) wpc = 1;   // wide single frames: few blocks, one block of 512 threads with ~90 registers
  dim3 grid((nl + 32 * wpc - 1) / (32 * wpc), nviews);
  const int th = 32 * wpc * sp.chunks;
  const float a = sp.alpha, b = 1 - sp.alpha;
  if (!ROWT && wpc == 2 && col4_on() && col4_supported(len, sp.chunks, sp.overlap) && pitch % 2 == 0 &&
      plane % 2 == 0 && (reinterpret_cast<uintptr_t>(mat) | reinterpret_cast<uintptr_t>(ref)) % 16 == 0) {
    // fourth generation: matched rows staged in shared memory by TMA (needs 16-byte aligned rows)
    static const bool attr = [] {
      const int bytes = (int)col4_smem_bytes(16);
      return cudaFuncSetAttribute(k_sweep_col4<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) == cudaSuccess &&
             cudaFuncSetAttribute(k_sweep_col4<-1>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) == cudaSuccess;
    }();
    if (attr) {
      const size_t bytes = col4_smem_bytes(sp.chunks);
      if (dir > 0) k_sweep_col4<1><<<grid, th, bytes, st>>>(ref, mat, dc, g, pitch, plane, sp.chunks, sp.overlap, a, b);
      else k_sweep_col4<-1><<<grid, th, bytes, st>>>(ref, mat, dc, g, pitch, plane, sp.chunks, sp.overlap, a, b);
      return cudaGetLastError() == cudaSuccess ? 1 : -1;
    }
    (void)cudaGetLastError();
  }
  if (ROWT || wpc == 1) {
    if (dir > 0) k_sweep_col3<1, 1, ROWT><<<grid, th, 0, st>>>(ref, mat, dc, g, pitch, plane, sp.chunks, sp.overlap, a, b);
    else k_sweep_col3<-1, 1, ROWT><<<grid, th, 0, st>>>(ref, mat, dc, g, pitch, plane, sp.chunks, sp.overlap, a, b);
  } else {
    if (dir > 0) k_sweep_col3<1, 2, false><<<grid, th, 0, st>>>(ref, mat, dc, g, pitch, plane, sp.chunks, sp.overlap, a, b);
    else k_sweep_col3<-1, 2, false><<<grid, th, 0, st>>>(ref, mat, dc, g, pitch, plane, sp.chunks, sp.overlap, a, b);
  }
  return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_sweep_col_inplace(const float2* ref, const float2* mat, float2* dc, ViewGeom g, int nviews,
                             int dir, SweepParams sp, cudaStream_t st) {
  return launch_col3<false>(ref, mat, dc, g, g.pitch, g.plane, nviews, dir, sp, st);
}

int launch_sweep_rowT_inplace(const float2* refT, const float2* matT, float2* dcT, ViewGeom g, int pitchT,
                              size_t planeT, int nviews, int dir, SweepParams sp, cudaStream_t st) {
  return launch_col3<true>(refT, matT, dcT, g, pitchT, planeT, nviews, dir, sp, st);
}

int launch_sweep_rowT(const float2* refT, const float2* matT, const float2* dcT_in, float2* dcT_out,
                      ViewGeom g, int pitchT, size_t planeT, int nviews, int dir, SweepParams sp,
                      cudaStream_t st) {
  int max_walk = 0;
  const int bar_step = col_bar_step(sp.overlap);
  if (!sweep_block_plan(g.w, sp.chunks, sp.overlap, bar_step, kPFCol, 16, &max_walk)) return -1;
  max_walk = (max_walk + kPFCol - 1) / kPFCol * kPFCol;
  dim3 grid((g.h + 31) / 32, nviews);
  k_sweep_col<true><<<grid, 32 * sp.chunks, 0, st>>>(refT, matT, dcT_in, dcT_out, g, pitchT, planeT,
                                                     dir, sp.chunks, sp.overlap, max_walk, bar_step,
                                                     sp.alpha, 1 - sp.alpha, skip_eq_flag());
  return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

bool sweep_rowT_supported(int w, int chunks, int ov) {
  int mw;
  return sweep_block_plan(w, chunks, ov, col_bar_step(ov), kPFCol, 16, &mw);
}

bool sweep_col_supported(int h, int chunks, int ov) {
  int mw;
  return sweep_block_plan(h, chunks, ov, col_bar_step(ov), kPFCol, 16, &mw);
}

}  // namespace pm
