// pm_device.cuh -- device-side building blocks shared by every kernel.
//
// All float arithmetic that decides a result is written with explicit
// __f*_rn intrinsics so that nvcc can neither contract nor reorder it: the
// forms below are the contractions nvcc applies to the reference's own
// expressions (profiles/contraction_evidence.txt) and the CPU oracle pins the
// same ones.  Citations are relative to /root/reference.
#pragma once

#include <climits>
#include <cuda_runtime.h>
#include <stdint.h>

namespace pm {

// One view problem = a reference image and the image it is matched against,
// each stored as interleaved {intensity, gradient magnitude} float2 planes
// (the reference's Il/Gl and Ir/Gr GpuMats, patchmatch_gpu.cu:346-352), plus the
// running {disparity, cost(disparity)} plane.  Rows are `pitch` elements apart
// and every row has at least one zeroed pad element after column w-1.
struct ViewGeom {
  int w, h, pitch;          // pitch in float2 elements
  size_t plane;             // elements between consecutive views
  // Row-band mode (one frame split across GPUs): the planes hold rows
  // [y_off, y_off + h) of a frame of full_h rows. Whole frames: y_off = 0, full_h = h.
  int y_off, full_h;
  // 0 = L1GradientCost3x3, the 5 taps the reference evaluates; 1 = the full (2r+1)^2 L1GradientCost
  // (patchmatch_gpu.cu:45-69); 2 = census + Hamming over the same window (extension). Only the
  // one-thread-per-pixel / per-chain kernels look at it; the block sweep kernels are built for
  // mode 0 with radius 1 and are not selected otherwise.
  int cost_mode;
  // patch_size / 2, the patch_radius of the reference's kernels (patchmatch_gpu.cu:129, 188, 244):
  // rows and columns closer than this to the border are skipped, xr = fmaxf(x - d, radius), an
  // accepted disparity is clamped to x - radius.
  int radius;
};

// Rows whose cost the reference evaluates (1 .. rows-2 of the FRAME, patchmatch_gpu.cu:134)
// and whose neighbour rows are present in these planes.
__host__ __device__ __forceinline__ bool row_interior(const ViewGeom& g, int y) {
  const int yg = y + g.y_off, r = g.radius;
  return yg >= r && yg <= g.full_h - 1 - r && y >= r && y <= g.h - 1 - r;
}
__host__ __device__ __forceinline__ bool col_interior(const ViewGeom& g, int x) {
  return x >= g.radius && x <= g.w - 1 - g.radius;
}

// GetSubpixel (patchmatch_gpu.cu:18-42) at an integral row: row0 == row1 and
// trow == 0, hence c0 = c00, c1 = c01 exactly and only the column lerp
// (1-t)*c0 + t*c1 -> fma(1-t, c0, t*c1) remains. ceil(col) == floor(col) when
// t == 0; reading floor+1 instead is identical because its weight is then 0 and
// the pad element is finite.
//
// floor() without the conversion unit: for 0 <= col < 2^22, col + 2^23 rounds to the
// nearest integer, whose low mantissa bits are that integer; one compare turns
// round-to-nearest into floor. F2I/I2F run on the quarter-rate XU pipe, these are
// plain FADD/IADD. The values are the ones __float2int_rd / __int2float_rn give.
__device__ __forceinline__ void col_split(float col, int& c0, float& t, float& omt) {
  const float big = 8388608.0f;
  const float r = __fadd_rn(col, big);
  float fl = __fsub_rn(r, big);
  int i = __float_as_int(r) - 0x4B000000;
  if (fl > col) { fl = __fsub_rn(fl, 1.0f); i -= 1; }
  c0 = i;
  t = __fsub_rn(col, fl);
  omt = __fsub_rn(1.0f, t);
}

__device__ __forceinline__ float2 lerp_ig(const float2* __restrict__ row, int c0, float t, float omt) {
  const float2 a = row[c0];
  const float2 b = row[c0 + 1];
  float2 r;
  r.x = __fmaf_rn(omt, a.x, __fmul_rn(t, b.x));
  r.y = __fmaf_rn(omt, a.y, __fmul_rn(t, b.y));
  return r;
}

__device__ __forceinline__ float2 lerp2(float2 a, float2 b, float t, float omt) {
  float2 r;
  r.x = __fmaf_rn(omt, a.x, __fmul_rn(t, b.x));
  r.y = __fmaf_rn(omt, a.y, __fmul_rn(t, b.y));
  return r;
}

// alpha*|dI| + (1-alpha)*|dG| -> fma(|dI|, alpha, (1-alpha)*|dG|)
__device__ __forceinline__ float tap_term(float2 l, float2 r, float alpha, float w1) {
  const float di = fabsf(__fsub_rn(l.x, r.x));
  const float dg = fabsf(__fsub_rn(l.y, r.y));
  return __fmaf_rn(di, alpha, __fmul_rn(w1, dg));
}

// The five reference-image taps of L1GradientCost3x3 (TL, TR, C, BL, BR,
// patchmatch_gpu.cu:84-111); independent of the hypothesis.
struct RefTaps {
  float2 tl, tr, c, bl, br;
};

__device__ __forceinline__ RefTaps load_ref_taps(const float2* __restrict__ ref, int pitch, int y, int x) {
  const float2* r0 = ref + (size_t)(y - 1) * pitch + x;
  const float2* r1 = r0 + pitch;
  const float2* r2 = r1 + pitch;
  RefTaps t;
  t.tl = r0[-1];
  t.tr = r0[1];
  t.c = r1[0];
  t.bl = r2[-1];
  t.br = r2[1];
  return t;
}

// L1GradientCost3x3 (patchmatch_gpu.cu:72-114) for the hypothesis whose centre lands at
// column xr (>= 1) of the matched rows m0 (y-1), m1 (y), m2 (y+1).
//
// The three sample columns are xr-1, xr, xr+1. xr-1 is always exact in float (xr >= 1),
// so its floor is floor(xr)-1 and its fraction is that of xr. xr+1 is exact unless it
// crosses into a binade with a coarser ulp; then (and only then) it is split on its own,
// as the reference's per-tap GetSubpixel would.
__device__ __forceinline__ float cost5_rows(const RefTaps& L, const float2* __restrict__ m0,
                                            const float2* __restrict__ m1,
                                            const float2* __restrict__ m2, float xr, float alpha,
                                            float w1) {
  int cc;
  float t, om;
  col_split(xr, cc, t, om);
  const float colp = __fadd_rn(xr, 1.0f);
  int cp = cc + 1;
  float tp = t, op = om;
  if (__fsub_rn(colp, 1.0f) != xr) col_split(colp, cp, tp, op);  // rare: xr+1 was rounded
  float cost = tap_term(L.tl, lerp2(m0[cc - 1], m0[cc], t, om), alpha, w1);
  cost = __fadd_rn(cost, tap_term(L.tr, lerp2(m0[cp], m0[cp + 1], tp, op), alpha, w1));
  cost = __fadd_rn(cost, tap_term(L.c, lerp2(m1[cc], m1[cc + 1], t, om), alpha, w1));
  cost = __fadd_rn(cost, tap_term(L.bl, lerp2(m2[cc - 1], m2[cc], t, om), alpha, w1));
  cost = __fadd_rn(cost, tap_term(L.br, lerp2(m2[cp], m2[cp + 1], tp, op), alpha, w1));
  return cost;
}

__device__ __forceinline__ float cost5(const RefTaps& L, const float2* __restrict__ mat, int pitch,
                                       int y, float xr, float alpha, float w1) {
  const float2* m1 = mat + (size_t)y * pitch;
  return cost5_rows(L, m1 - pitch, m1, m1 + pitch, xr, alpha, w1);
}

// ---- the same evaluation with Blackwell's packed f32x2 pipe (FMUL2 / FFMA2 / FADD2) -------
// Every lane of a packed instruction is an IEEE round-to-nearest operation, so the
// results are bit-identical to the scalar forms above; what changes is the number of
// issue slots: {I, G} pairs are lerped and differenced with one instruction instead of two.

// floor() with one FADD.RM: for 0 <= col < 2^22, col + 2^23 rounded towards -inf is
// floor(col) + 2^23 exactly, and its low mantissa bits are that integer.
__device__ __forceinline__ void col_split_rd(float col, int& c0, float& t, float& omt) {
  const float big = 8388608.0f;
  const float r = __fadd_rd(col, big);
  c0 = __float_as_int(r) - 0x4B000000;
  t = __fsub_rn(col, __fsub_rn(r, big));
  omt = __fsub_rn(1.0f, t);
}

__device__ __forceinline__ float2 lerp2p(float2 a, float2 b, float t, float omt) {
  return __ffma2_rn(a, make_float2(omt, omt), __fmul2_rn(b, make_float2(t, t)));
}

__device__ __forceinline__ float tap_term_p(float2 l, float2 r, float alpha, float w1) {
  const float2 d = __fadd2_rn(l, make_float2(-r.x, -r.y));
  return __fmaf_rn(fabsf(d.x), alpha, __fmul_rn(w1, fabsf(d.y)));
}

// cost5_rows with packed arithmetic. GLOBAL: the matched rows live in global memory and are
// read through the read-only path (ld.global.nc, L1-resident); a plain ld.global of data the
// compiler cannot prove read-only took ~7x longer per step on B200. Otherwise: shared memory.
template <bool GLOBAL>
__device__ __forceinline__ float2 ld_mat(const float2* p) {
  if (GLOBAL) return __ldg(p);
  return *p;
}

template <bool GLOBAL>
__device__ __forceinline__ float cost5_packed(const RefTaps& L, const float2* m0, const float2* m1,
                                              const float2* m2, float xr, float alpha, float w1) {
  int cc;
  float t, om;
  col_split_rd(xr, cc, t, om);
  const float colp = __fadd_rn(xr, 1.0f);
  const float2* a0 = m0 + cc;
  const float2* a1 = m1 + cc;
  const float2* a2 = m2 + cc;
  float2 tr, br;
  if (__fsub_rn(colp, 1.0f) != xr) {  // rare: xr+1 was rounded, split it like the reference does
    int cp;
    float tp, op;
    col_split_rd(colp, cp, tp, op);
    tr = lerp2p(ld_mat<GLOBAL>(m0 + cp), ld_mat<GLOBAL>(m0 + cp + 1), tp, op);
    br = lerp2p(ld_mat<GLOBAL>(m2 + cp), ld_mat<GLOBAL>(m2 + cp + 1), tp, op);
  } else {
    tr = lerp2p(ld_mat<GLOBAL>(a0 + 1), ld_mat<GLOBAL>(a0 + 2), t, om);
    br = lerp2p(ld_mat<GLOBAL>(a2 + 1), ld_mat<GLOBAL>(a2 + 2), t, om);
  }
  float cost = tap_term_p(L.tl, lerp2p(ld_mat<GLOBAL>(a0 - 1), ld_mat<GLOBAL>(a0), t, om), alpha, w1);
  cost = __fadd_rn(cost, tap_term_p(L.tr, tr, alpha, w1));
  cost = __fadd_rn(cost, tap_term_p(L.c, lerp2p(ld_mat<GLOBAL>(a1), ld_mat<GLOBAL>(a1 + 1), t, om), alpha, w1));
  cost = __fadd_rn(cost, tap_term_p(L.bl, lerp2p(ld_mat<GLOBAL>(a2 - 1), ld_mat<GLOBAL>(a2), t, om), alpha, w1));
  cost = __fadd_rn(cost, tap_term_p(L.br, br, alpha, w1));
  return cost;
}

// ---- rolling window over the matched rows (row sweeps, shared memory) ------------------------
// Along a row sweep about four candidate evaluations out of five have the disparity of the step
// before (the sweep just copied it), so their sample column moved by exactly one pixel with
// the same fraction: of the ten matched-image elements an evaluation needs (columns cc-1..cc+2
// of rows y-1 and y+1, cc..cc+1 of row y) seven are the previous step's. A lane keeps them in
// registers and loads only the three new ones; the other seven loads are predicated off, which
// is what counts on the shared-memory data pipe (wavefronts are per active lane). Lanes whose
// column or fraction changed reload everything. Same operands, same operations: bit-identical.
struct MatWin {
  float2 a[4], c[2], b[4];  // rows y-1, y, y+1
  int cc;                   // column of a[1] / c[0] / b[1]; INT_MIN/2 = nothing usable
  float t;                  // fraction the window was loaded for
};

__device__ __forceinline__ void lds_f2(float2& v, unsigned addr) {
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
}
__device__ __forceinline__ void lds_f2_if(float2& v, unsigned addr, bool on) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.s32 p, %3, 0;\n"
      "@p ld.shared.v2.f32 {%0, %1}, [%2];\n"
      "}" : "+f"(v.x), "+f"(v.y) : "r"(addr), "r"((int)on));
}

// cost5_packed<false> through the window; s0/s1/s2 = shared addresses of rows y-1, y, y+1.
template <int DIR>
__device__ __forceinline__ float cost5_window(const RefTaps& L, MatWin& W, unsigned s0, unsigned s1,
                                              unsigned s2, float xr, float alpha, float w1) {
  int cc;
  float t, om;
  col_split_rd(xr, cc, t, om);
  const float colp = __fadd_rn(xr, 1.0f);
  const bool exactp = __fsub_rn(colp, 1.0f) == xr;
  const bool roll = exactp && cc == W.cc + DIR && t == W.t;
  const bool full = !roll;
  const unsigned o = (unsigned)cc * 8u;
  if (DIR > 0) {
    if (roll) {
      W.a[0] = W.a[1]; W.a[1] = W.a[2]; W.a[2] = W.a[3];
      W.c[0] = W.c[1];
      W.b[0] = W.b[1]; W.b[1] = W.b[2]; W.b[2] = W.b[3];
    }
    lds_f2_if(W.a[0], s0 + o - 8u, full); lds_f2_if(W.a[1], s0 + o, full);
    lds_f2_if(W.a[2], s0 + o + 8u, full); lds_f2(W.a[3], s0 + o + 16u);
    lds_f2_if(W.c[0], s1 + o, full);      lds_f2(W.c[1], s1 + o + 8u);
    lds_f2_if(W.b[0], s2 + o - 8u, full); lds_f2_if(W.b[1], s2 + o, full);
    lds_f2_if(W.b[2], s2 + o + 8u, full); lds_f2(W.b[3], s2 + o + 16u);
  } else {
    if (roll) {
      W.a[3] = W.a[2]; W.a[2] = W.a[1]; W.a[1] = W.a[0];
      W.c[1] = W.c[0];
      W.b[3] = W.b[2]; W.b[2] = W.b[1]; W.b[1] = W.b[0];
    }
    lds_f2(W.a[0], s0 + o - 8u);          lds_f2_if(W.a[1], s0 + o, full);
    lds_f2_if(W.a[2], s0 + o + 8u, full); lds_f2_if(W.a[3], s0 + o + 16u, full);
    lds_f2(W.c[0], s1 + o);               lds_f2_if(W.c[1], s1 + o + 8u, full);
    lds_f2(W.b[0], s2 + o - 8u);          lds_f2_if(W.b[1], s2 + o, full);
    lds_f2_if(W.b[2], s2 + o + 8u, full); lds_f2_if(W.b[3], s2 + o + 16u, full);
  }
  W.cc = exactp ? cc : INT_MIN / 2;
  W.t = t;
  float2 tr, br;
  if (!exactp) {  // rare: xr+1 was rounded, split it like the reference does
    int cp;
    float tp, op;
    col_split_rd(colp, cp, tp, op);
    float2 p0, p1, q0, q1;
    lds_f2(p0, s0 + (unsigned)cp * 8u); lds_f2(p1, s0 + (unsigned)cp * 8u + 8u);
    lds_f2(q0, s2 + (unsigned)cp * 8u); lds_f2(q1, s2 + (unsigned)cp * 8u + 8u);
    tr = lerp2p(p0, p1, tp, op);
    br = lerp2p(q0, q1, tp, op);
  } else {
    tr = lerp2p(W.a[2], W.a[3], t, om);
    br = lerp2p(W.b[2], W.b[3], t, om);
  }
  float cost = tap_term_p(L.tl, lerp2p(W.a[0], W.a[1], t, om), alpha, w1);
  cost = __fadd_rn(cost, tap_term_p(L.tr, tr, alpha, w1));
  cost = __fadd_rn(cost, tap_term_p(L.c, lerp2p(W.c[0], W.c[1], t, om), alpha, w1));
  cost = __fadd_rn(cost, tap_term_p(L.bl, lerp2p(W.b[0], W.b[1], t, om), alpha, w1));
  cost = __fadd_rn(cost, tap_term_p(L.br, br, alpha, w1));
  return cost;
}

// ---- the same two evaluations over the slot-interleaved staging of the row kernel -------------
// The 16 matched rows of a block are staged as [column][16 rows] (128 bytes per column): lane r
// always reads slot r +- 1, so the 16 lanes of a half-warp hit 16 different bank pairs WHATEVER
// their columns are - gathers at random columns (the noise step, fresh candidates) cost the two
// wavefronts a 64-bit warp load needs at least, instead of 4-6 with row-major rows. The two halo
// rows (above the block's first row, below its last) stay row-major and are read by one lane each.
// A lane therefore addresses its three rows as base + column * stride with its own stride.
struct RowsIL {
  unsigned b0, b1, b2;   // shared addresses of column 0 in rows y-1, y, y+1
  unsigned s0, s2;       // bytes per column in rows y-1 and y+1 (128, or 8 for a halo row); row y: 128
  const char *p0, *p1, *p2;  // the same three bases as pointers (plain loads, free to be scheduled)
};

__device__ __forceinline__ float2 ld_il(const char* base, unsigned off) {
  return *reinterpret_cast<const float2*>(base + off);
}

__device__ __forceinline__ float cost5_packed_il(const RefTaps& L, const RowsIL& R, float xr, float alpha,
                                                 float w1) {
  int cc;
  float t, om;
  col_split_rd(xr, cc, t, om);
  const float colp = __fadd_rn(xr, 1.0f);
  const unsigned o0 = (unsigned)cc * R.s0, o1 = (unsigned)cc * 128u, o2 = (unsigned)cc * R.s2;
  float2 tr, br;
#ifdef PM_ASSUME_EXACT
  if (false) {
#else
  if (__fsub_rn(colp, 1.0f) != xr) {  // rare: xr+1 was rounded, split it like the reference does
#endif
    int cp;
    float tp, op;
    col_split_rd(colp, cp, tp, op);
    tr = lerp2p(ld_il(R.p0, (unsigned)cp * R.s0), ld_il(R.p0, (unsigned)(cp + 1) * R.s0), tp, op);
    br = lerp2p(ld_il(R.p2, (unsigned)cp * R.s2), ld_il(R.p2, (unsigned)(cp + 1) * R.s2), tp, op);
  } else {
    tr = lerp2p(ld_il(R.p0, o0 + R.s0), ld_il(R.p0, o0 + 2u * R.s0), t, om);
    br = lerp2p(ld_il(R.p2, o2 + R.s2), ld_il(R.p2, o2 + 2u * R.s2), t, om);
  }
  float cost = tap_term_p(L.tl, lerp2p(ld_il(R.p0, o0 - R.s0), ld_il(R.p0, o0), t, om), alpha, w1);
  cost = __fadd_rn(cost, tap_term_p(L.tr, tr, alpha, w1));
  cost = __fadd_rn(cost, tap_term_p(L.c, lerp2p(ld_il(R.p1, o1), ld_il(R.p1, o1 + 128u), t, om), alpha, w1));
  cost = __fadd_rn(cost, tap_term_p(L.bl, lerp2p(ld_il(R.p2, o2 - R.s2), ld_il(R.p2, o2), t, om), alpha, w1));
  cost = __fadd_rn(cost, tap_term_p(L.br, br, alpha, w1));
  return cost;
}

template <int DIR>
__device__ __forceinline__ float cost5_window_il(const RefTaps& L, MatWin& W, const RowsIL& R, float xr,
                                                 float alpha, float w1) {
  int cc;
  float t, om;
  col_split_rd(xr, cc, t, om);
  const float colp = __fadd_rn(xr, 1.0f);
#ifdef PM_ASSUME_EXACT
  const bool exactp = true;
#else
  const bool exactp = __fsub_rn(colp, 1.0f) == xr;
#endif
  const bool roll = exactp && cc == W.cc + DIR && t == W.t;
  const bool full = !roll;
  const unsigned a0 = R.b0 + (unsigned)cc * R.s0, a1 = R.b1 + (unsigned)cc * 128u,
                 a2 = R.b2 + (unsigned)cc * R.s2;
  if (DIR > 0) {
    if (roll) {
      W.a[0] = W.a[1]; W.a[1] = W.a[2]; W.a[2] = W.a[3];
      W.c[0] = W.c[1];
      W.b[0] = W.b[1]; W.b[1] = W.b[2]; W.b[2] = W.b[3];
    }
    lds_f2_if(W.a[0], a0 - R.s0, full); lds_f2_if(W.a[1], a0, full);
    lds_f2_if(W.a[2], a0 + R.s0, full); lds_f2(W.a[3], a0 + 2u * R.s0);
    lds_f2_if(W.c[0], a1, full);        lds_f2(W.c[1], a1 + 128u);
    lds_f2_if(W.b[0], a2 - R.s2, full); lds_f2_if(W.b[1], a2, full);
    lds_f2_if(W.b[2], a2 + R.s2, full); lds_f2(W.b[3], a2 + 2u * R.s2);
  } else {
    if (roll) {
      W.a[3] = W.a[2]; W.a[2] = W.a[1]; W.a[1] = W.a[0];
      W.c[1] = W.c[0];
      W.b[3] = W.b[2]; W.b[2] = W.b[1]; W.b[1] = W.b[0];
    }
    lds_f2(W.a[0], a0 - R.s0);          lds_f2_if(W.a[1], a0, full);
    lds_f2_if(W.a[2], a0 + R.s0, full); lds_f2_if(W.a[3], a0 + 2u * R.s0, full);
    lds_f2(W.c[0], a1);                 lds_f2_if(W.c[1], a1 + 128u, full);
    lds_f2(W.b[0], a2 - R.s2);          lds_f2_if(W.b[1], a2, full);
    lds_f2_if(W.b[2], a2 + R.s2, full); lds_f2_if(W.b[3], a2 + 2u * R.s2, full);
  }
  W.cc = exactp ? cc : INT_MIN / 2;
  W.t = t;
  float2 tr, br;
  if (!exactp) {  // rare: xr+1 was rounded, split it like the reference does
    int cp;
    float tp, op;
    col_split_rd(colp, cp, tp, op);
    float2 p0, p1, q0, q1;
    lds_f2(p0, R.b0 + (unsigned)cp * R.s0); lds_f2(p1, R.b0 + (unsigned)(cp + 1) * R.s0);
    lds_f2(q0, R.b2 + (unsigned)cp * R.s2); lds_f2(q1, R.b2 + (unsigned)(cp + 1) * R.s2);
    tr = lerp2p(p0, p1, tp, op);
    br = lerp2p(q0, q1, tp, op);
  } else {
    tr = lerp2p(W.a[2], W.a[3], t, om);
    br = lerp2p(W.b[2], W.b[3], t, om);
  }
  float cost = tap_term_p(L.tl, lerp2p(W.a[0], W.a[1], t, om), alpha, w1);
  cost = __fadd_rn(cost, tap_term_p(L.tr, tr, alpha, w1));
  cost = __fadd_rn(cost, tap_term_p(L.c, lerp2p(W.c[0], W.c[1], t, om), alpha, w1));
  cost = __fadd_rn(cost, tap_term_p(L.bl, lerp2p(W.b[0], W.b[1], t, om), alpha, w1));
  cost = __fadd_rn(cost, tap_term_p(L.br, br, alpha, w1));
  return cost;
}

// ---- branch-free forms (third-generation row kernel) -------------------------------------------
// The column xr+1 of the right-hand taps is ALWAYS split on its own, as the reference's per-tap
// GetSubpixel does (when xr+1 is exact the split is bit-equal to {floor(xr)+1, t, 1-t}, so nothing
// changes; when it was rounded this IS the reference's arithmetic). No data-dependent branch is left
// in an evaluation, so the sixteen unrolled steps of the sweep form one basic block and the
// compiler overlaps a step's independent work with the previous step's dependent chain.
__device__ __forceinline__ float cost5_full_bf(const RefTaps& L, const RowsIL& R, float xr, float alpha,
                                               float w1) {
  int cc, cp;
  float t, om, tp, op;
  col_split_rd(xr, cc, t, om);
  col_split_rd(__fadd_rn(xr, 1.0f), cp, tp, op);
  const unsigned o0 = (unsigned)cc * R.s0, o1 = (unsigned)cc * 128u, o2 = (unsigned)cc * R.s2;
  const unsigned q0 = (unsigned)cp * R.s0, q2 = (unsigned)cp * R.s2;
  float cost = tap_term_p(L.tl, lerp2p(ld_il(R.p0, o0 - R.s0), ld_il(R.p0, o0), t, om), alpha, w1);
  cost = __fadd_rn(cost, tap_term_p(L.tr, lerp2p(ld_il(R.p0, q0), ld_il(R.p0, q0 + R.s0), tp, op), alpha, w1));
  cost = __fadd_rn(cost, tap_term_p(L.c, lerp2p(ld_il(R.p1, o1), ld_il(R.p1, o1 + 128u), t, om), alpha, w1));
  cost = __fadd_rn(cost, tap_term_p(L.bl, lerp2p(ld_il(R.p2, o2 - R.s2), ld_il(R.p2, o2), t, om), alpha, w1));
  cost = __fadd_rn(cost, tap_term_p(L.br, lerp2p(ld_il(R.p2, q2), ld_il(R.p2, q2 + R.s2), tp, op), alpha, w1));
  return cost;
}

// Rolling window, branch-free: the elements are keyed by COLUMN only (the weights are recomputed at
// every step), so a step whose floor column moved by exactly one pixel keeps the elements it can
// and loads the rest; the right-hand pair is always loaded at its own split's column. W.cc is the
// column the kept elements belong to, or invalid when the right-hand split fell outside cc + 1.
template <int DIR>
__device__ __forceinline__ float cost5_win_bf(const RefTaps& L, MatWin& W, const RowsIL& R, float xr,
                                              float alpha, float w1) {
  int cc, cp;
  float t, om, tp, op;
  col_split_rd(xr, cc, t, om);
  col_split_rd(__fadd_rn(xr, 1.0f), cp, tp, op);
  const bool roll = cc == W.cc + DIR;
  const bool full = !roll;
  const unsigned a0 = R.b0 + (unsigned)cc * R.s0, a1 = R.b1 + (unsigned)cc * 128u,
                 a2 = R.b2 + (unsigned)cc * R.s2;
  const unsigned r0 = R.b0 + (unsigned)cp * R.s0, r2 = R.b2 + (unsigned)cp * R.s2;
  if (DIR > 0) {
    // kept: columns cc-1, cc (were cc, cc+1) and cc of the middle row (was cc+1)
    if (roll) { W.a[0] = W.a[1]; W.a[1] = W.a[2]; W.c[0] = W.c[1]; W.b[0] = W.b[1]; W.b[1] = W.b[2]; }
    lds_f2_if(W.a[0], a0 - R.s0, full); lds_f2_if(W.a[1], a0, full);
    lds_f2(W.a[2], r0);                 lds_f2(W.a[3], r0 + R.s0);
    lds_f2_if(W.c[0], a1, full);        lds_f2(W.c[1], a1 + 128u);
    lds_f2_if(W.b[0], a2 - R.s2, full); lds_f2_if(W.b[1], a2, full);
    lds_f2(W.b[2], r2);                 lds_f2(W.b[3], r2 + R.s2);
    W.cc = cp == cc + 1 ? cc : INT_MIN / 2;   // a[2] must be column cc+1 to be kept next time
  } else {
    // kept: column cc (was cc-1... i.e. the old a[0]) and cc+1 of the middle row (the old c[0])
    if (roll) { W.a[1] = W.a[0]; W.c[1] = W.c[0]; W.b[1] = W.b[0]; }
    lds_f2(W.a[0], a0 - R.s0);          lds_f2_if(W.a[1], a0, full);
    lds_f2(W.a[2], r0);                 lds_f2(W.a[3], r0 + R.s0);
    lds_f2(W.c[0], a1);                 lds_f2_if(W.c[1], a1 + 128u, full);
    lds_f2(W.b[0], a2 - R.s2);          lds_f2_if(W.b[1], a2, full);
    lds_f2(W.b[2], r2);                 lds_f2(W.b[3], r2 + R.s2);
    W.cc = cc;
  }
  float cost = tap_term_p(L.tl, lerp2p(W.a[0], W.a[1], t, om), alpha, w1);
  cost = __fadd_rn(cost, tap_term_p(L.tr, lerp2p(W.a[2], W.a[3], tp, op), alpha, w1));
  cost = __fadd_rn(cost, tap_term_p(L.c, lerp2p(W.c[0], W.c[1], t, om), alpha, w1));
  cost = __fadd_rn(cost, tap_term_p(L.bl, lerp2p(W.b[0], W.b[1], t, om), alpha, w1));
  cost = __fadd_rn(cost, tap_term_p(L.br, lerp2p(W.b[2], W.b[3], tp, op), alpha, w1));
  return cost;
}

// predicated global loads (no branch around a conditional prefetch)
__device__ __forceinline__ void ldg_nc_f2_if(float2& v, const float2* p, bool on) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.s32 p, %3, 0;\n"
      "@p ld.global.nc.v2.f32 {%0, %1}, [%2];\n"
      "}" : "+f"(v.x), "+f"(v.y) : "l"(p), "r"((int)on));
}
__device__ __forceinline__ void ldg_cg_f2_if(float2& v, const float2* p, bool on) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.s32 p, %3, 0;\n"
      "@p ld.global.cg.v2.f32 {%0, %1}, [%2];\n"
      "}" : "+f"(v.x), "+f"(v.y) : "l"(p), "r"((int)on));
}
__device__ __forceinline__ void ldg_nc_f32_if(float& v, const float* p, bool on) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.s32 p, %2, 0;\n"
      "@p ld.global.nc.f32 %0, [%1];\n"
      "}" : "+f"(v) : "l"(p), "r"((int)on));
}

// L1GradientCost (patchmatch_gpu.cu:45-69) with ph = pw = 2r+1: taps in raster order, sample
// column xr - float(pw/2) + float(col) evaluated left to right, each split on its own.
__device__ __forceinline__ float cost_full(const float2* __restrict__ ref,
                                           const float2* __restrict__ mat, int pitch, int y, int x,
                                           float xr, float alpha, float w1, int r) {
  float cost = 0.0f;
  const float xb = __fsub_rn(xr, __int2float_rn(r));
  for (int row = 0; row < 2 * r + 1; ++row) {
    const float2* rr = ref + (size_t)(y - r + row) * pitch + (x - r);
    const float2* mr = mat + (size_t)(y - r + row) * pitch;
    for (int col = 0; col < 2 * r + 1; ++col) {
      int c0;
      float t, om;
      col_split(__fadd_rn(xb, __int2float_rn(col)), c0, t, om);
      cost = __fadd_rn(cost, tap_term(rr[col], lerp_ig(mr, c0, t, om), alpha, w1));
    }
  }
  return cost;
}

// Census + Hamming over the (2r+1)^2 window of the intensity (extension):
// reference bit Il(tap) < Il(centre), matched bit S(tap) < S(centre)
// with S the reference's GetSubpixel sampling at columns xr - r + col; the number of differing bits.
__device__ __forceinline__ float cost_census(const float2* __restrict__ ref,
                                             const float2* __restrict__ mat, int pitch, int y, int x,
                                             float xr, int r) {
  const float xb = __fsub_rn(xr, __int2float_rn(r));
  const float cl = ref[(size_t)y * pitch + x].x;
  int c0;
  float t, om;
  col_split(__fadd_rn(xb, __int2float_rn(r)), c0, t, om);
  const float cr = lerp_ig(mat + (size_t)y * pitch, c0, t, om).x;
  int ham = 0;
  for (int row = 0; row < 2 * r + 1; ++row) {
    const float2* rr = ref + (size_t)(y - r + row) * pitch + (x - r);
    const float2* mr = mat + (size_t)(y - r + row) * pitch;
    for (int col = 0; col < 2 * r + 1; ++col) {
      if (row == r && col == r) continue;
      col_split(__fadd_rn(xb, __int2float_rn(col)), c0, t, om);
      const float b = lerp_ig(mr, c0, t, om).x;
      ham += (rr[col].x < cl) != (b < cr);
    }
  }
  return __int2float_rn(ham);
}

// the cost of hypothesis column xr at reference pixel (y, x): MODE 0 = the reference's 5 taps
// (compile-time, the fast kernels), MODE 1 = whatever the view's cost_mode / radius say
template <int MODE>
__device__ __forceinline__ float cost_at(const ViewGeom& g, const float2* __restrict__ ref,
                                         const float2* __restrict__ mat, int y, int x, float xr,
                                         float alpha, float w1) {
  if (MODE == 1) {
    if (g.cost_mode == 1) return cost_full(ref, mat, g.pitch, y, x, xr, alpha, w1, g.radius);
    if (g.cost_mode == 2) return cost_census(ref, mat, g.pitch, y, x, xr, g.radius);
  }
  const RefTaps L = load_ref_taps(ref, g.pitch, y, x);
  return cost5(L, mat, g.pitch, y, xr, alpha, w1);
}

// fmaxf(x - d, patch_radius), patchmatch_gpu.cu:162
__device__ __forceinline__ float xr_of(int x, float d) {
  return fmaxf(__fsub_rn(__int2float_rn(x), d), 1.0f);
}
__device__ __forceinline__ float xr_of_r(int x, float d, int r) {
  return fmaxf(__fsub_rn(__int2float_rn(x), d), __int2float_rn(r));
}

// Philox-4x32-10, key = 64-bit seed; returns U[0,1) with 24 bits.
__device__ __forceinline__ float philox_u01(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2,
                                            uint32_t c3) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return __fmul_rn(__uint2float_rn(c0 >> 8), 5.9604644775390625e-8f);
}

// Chunk k of a sweep line of length len: positions [mn, mx) walked upwards
// (dir > 0) or (mn, mx] walked downwards (patchmatch_gpu.cu:141-156).
__device__ __forceinline__ void chunk_range(int k, int cs, int ov, int len, int dir, int& start,
                                            int& stop, int r = 1) {
  const int mn = max(k * cs - ov, r);
  const int mx = min((k + 1) * cs + ov, len - r - 1);
  start = dir > 0 ? mn : mx;
  stop = dir > 0 ? mx : mn;
}

}  // namespace pm
