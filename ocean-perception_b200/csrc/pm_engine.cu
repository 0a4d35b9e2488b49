// pm_engine.cu -- engine state and the C ABI (include/pm_b200.h).
//
// Host-side orchestration of PatchmatchGpu::Match (patchmatch_gpu.cu:331-411):
// one pass over a sub-batch of stereo pairs runs every view problem (left and
// right reference of every pair) through the same kernels. Citations are relative
// to /root/reference.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/pm_b200.h"
#include "pm_kernels.h"

namespace pm {
int launch_extract_cost(const float2* dc, ViewGeom g, int nviews, float* out, int opitch,
                        size_t oplane, cudaStream_t st);
}

using namespace pm;

namespace {

thread_local std::string g_create_error;

constexpr int kMaxLevels = 6;

enum Stage { ST_PRE = 0, ST_INIT, ST_NOISE, ST_SWEEP_ROW, ST_SWEEP_COL, ST_MASK, ST_FINAL, ST_COPY,
             ST_SEED, ST_XCHG };
const char* kStageNames[PM_N_STAGES] = {"preprocess", "init",      "noise_cost", "sweep_row",
                                        "sweep_col",  "mask_bg",   "finalize",   "plane_copy",
                                        "sparse_init", "band_exchange"};

struct Level {
  int w = 0, h = 0, pitch = 0, pitch8 = 0, npitch = 0, pitchT = 0;
  size_t plane = 0, plane8 = 0, planeT = 0;
  bool row_smem = false;  // row sweeps run the shared-memory block kernel at this level
  bool row_T = false;     // row sweeps run the column kernel on transposed planes (wide images)
  bool col_block = false; // column sweeps run the block kernel at this level
  bool col_inplace = false; // ... its third generation, which sweeps the {d, cost} plane in place
  uint8_t* L8 = nullptr;  // levels >= 1 only (level 0 reads the caller's images)
  uint8_t* R8 = nullptr;
  float* noise = nullptr;
  float* noiseT = nullptr;  // transposed copy ([w][pitchT]) for the fused noise + row sweep
  bool fuse_noise = false;  // the row(+1) sweep applies the iteration's noise itself
  bool row_rm = false;      // the shared-memory row kernel reads the {d, cost} plane row-major
  bool row_il = false;      // ... and stages the matched rows from the slot-interleaved copy (matI)
  size_t planeI = 0;        // elements per view of matI at this level
};

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

}  // namespace

struct pm_engine {
  pm_params p;
  int device = 0;
  std::string err;
  cudaStream_t stream = nullptr, s_in = nullptr, s_out = nullptr;
  // workspace
  int w = 0, h = 0, nb = 0, levels = 1;
  Level lv[kMaxLevels];
  float2 *ref = nullptr, *mat = nullptr, *dcA = nullptr, *dcB = nullptr;
  float2 *refT = nullptr, *dcT = nullptr;  // transposed planes (rows contiguous) for row sweeps
  float2 *matT = nullptr, *dcT2 = nullptr; // levels with row_T: transposed matched plane, sweep output
  float2 *matI = nullptr;                  // levels with row_il: slot-interleaved matched plane
  float* dispv = nullptr;  // [2*nb][h][pitch] plain disparity planes
  float* dprev = nullptr;  // previous pyramid level's disparity
  // host path: double-buffered device input/output
  uint8_t* d_in[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // [slot][L/R]
  float* d_seed[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
  float* d_out[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
  bool seed_alloc = false;
  // sparse seeding (pm_seed.cu): per-view keypoints and the seed maps of a device pass
  SeedState seed = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  float* seed_map[2] = {nullptr, nullptr};  // [nb][h][npitch], left / right-image coordinates
  int seed_views = 0, seed_maxf = 0;
  cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr},
              ev_out[2] = {nullptr, nullptr};
  uint64_t launches = 0;
  // One workspace, possibly several streams (pm_match_batch_device runs on the caller's): the last
  // stream that used the workspace records ev_ws when its call has been enqueued, the next user of
  // another stream waits for it before touching the planes.
  cudaStream_t last_stream = nullptr;
  cudaEvent_t ev_ws = nullptr;
  bool ws_used = false;
  // row bands over peer memory (pm_band_p2p_*): this rank's receive region (IPC-exported) and the
  // neighbours' regions mapped into this process; see include/pm_b200.h
  struct P2P {
    bool on = false;
    char* region = nullptr;
    size_t buf_bytes = 0;            // one receive buffer; the region holds 4 + the flags
    int width = 0;
    char *peer_prev = nullptr, *peer_next = nullptr;
    bool ipc_prev = false, ipc_next = false;   // opened with cudaIpcOpenMemHandle (to be closed)
    unsigned long long seq = 0;      // exchanges issued so far (both neighbours count alike)
  } p2p;
  // host path: device passes issued so far (slot = pass & 1; the slot events persist across
  // calls, so consecutive asynchronous calls keep the copy/compute pipeline full) and whether a
  // call is waiting for pm_wait
  uint64_t host_pass = 0;
  bool host_pending = false, host_pending_seedcheck = false;
  // profiling
  bool profiling = false;
  struct ProfSpan { int stage; cudaEvent_t a, b; };
  std::vector<ProfSpan> spans;       // recorded, not yet resolved
  std::vector<cudaEvent_t> ev_pool;  // recycled timing events
  float stage_ms[PM_N_STAGES] = {0};
  uint32_t stage_launches[PM_N_STAGES] = {0};
  // stage API state
  bool stage_loaded = false;
  // row-band mode (pm_band_*): this engine holds rows [load_lo, load_hi) of a frame
  struct Band {
    bool ws = false;        // the workspace was built for a band
    bool running = false;   // between pm_band_begin and pm_band_finish
    int rank = 0, world = 1, frame_h = 0;
    int own_lo = 0, own_hi = 0, load_lo = 0, load_hi = 0, k_lo = 0, nk = 0;
    int op = 0, nops = 0;   // next operation of the schedule
    int pending_dir = 0;    // column sweep whose received rows are not unpacked yet
    size_t xrows = 0;       // capacity of each exchange buffer, in rows (both views)
    float2 *send_prev = nullptr, *recv_prev = nullptr, *send_next = nullptr, *recv_next = nullptr;
    float* out[2] = {nullptr, nullptr};  // whole-band result planes
    const uint8_t *dL = nullptr, *dR = nullptr;
    size_t ipitch = 0;
    const float *seedL = nullptr, *seedR = nullptr;
    size_t spitch = 0;
    uint32_t pair_index = 0;
    cudaStream_t st = nullptr;
  } band;
};

namespace {

int fail(pm_engine* e, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (e) e->err = buf; else g_create_error = buf;
  return code;
}

#define PM_CUDA(e, call)                                                                   \
  do {                                                                                     \
    cudaError_t _st = (call);                                                              \
    if (_st != cudaSuccess)                                                                \
      return fail(e, _st == cudaErrorMemoryAllocation ? PM_ERR_OOM : PM_ERR_CUDA,          \
                  "%s failed: %s", #call, cudaGetErrorString(_st));                        \
  } while (0)

#define PM_LAUNCH(e, call)                                                                 \
  do {                                                                                     \
    int _n = (call);                                                                       \
    if (_n < 0)                                                                            \
      return fail(e, PM_ERR_CUDA, "%s failed: %s", #call,                                  \
                  cudaGetErrorString(cudaGetLastError()));                                 \
    (e)->launches += (uint64_t)_n;                                                         \
  } while (0)

constexpr size_t kP2PFlagBytes = 128;   // 4 sequence flags (u64) + the error word, after the 4 buffers
inline unsigned long long* p2p_flag(char* region, size_t buf, int idx) {
  return reinterpret_cast<unsigned long long*>(region + 4 * buf) + idx;
}
inline int* p2p_err(char* region, size_t buf) { return reinterpret_cast<int*>(region + 4 * buf + 64); }
inline unsigned* p2p_ticket(char* region, size_t buf) { return reinterpret_cast<unsigned*>(region + 4 * buf + 96); }

void p2p_close(pm_engine* e) {
  pm_engine::P2P& P = e->p2p;
  if (P.ipc_prev && P.peer_prev) cudaIpcCloseMemHandle(P.peer_prev);
  if (P.ipc_next && P.peer_next) cudaIpcCloseMemHandle(P.peer_next);
  P.peer_prev = P.peer_next = nullptr;
  P.ipc_prev = P.ipc_next = false;
  P.on = false;
}

// Stream ordering of the shared workspace (see pm_engine::ev_ws).
int ws_acquire(pm_engine* e, cudaStream_t st) {
  if (e->ws_used && e->last_stream != st) PM_CUDA(e, cudaStreamWaitEvent(st, e->ev_ws, 0));
  return PM_OK;
}
int ws_release(pm_engine* e, cudaStream_t st) {
  PM_CUDA(e, cudaEventRecord(e->ev_ws, st));
  e->last_stream = st;
  e->ws_used = true;
  return PM_OK;
}

ViewGeom geom(const pm_engine* e, const Level& l) {
  ViewGeom g;
  g.w = l.w; g.h = l.h; g.pitch = l.pitch; g.plane = l.plane;
  g.y_off = e->band.ws ? e->band.load_lo : 0;
  g.full_h = e->band.ws ? e->band.frame_h : l.h;
  g.cost_mode = e->p.cost_mode;
  g.radius = e->p.patch_size / 2;
  return g;
}

void free_workspace(pm_engine* e) {
  auto F = [](void* p) { if (p) cudaFree(p); };
  for (int l = 0; l < kMaxLevels; ++l) {
    F(e->lv[l].L8); F(e->lv[l].R8); F(e->lv[l].noise); F(e->lv[l].noiseT);
    e->lv[l] = Level();
  }
  F(e->ref); F(e->mat); F(e->dcA); F(e->dcB); F(e->dispv); F(e->dprev); F(e->refT); F(e->dcT);
  F(e->matT); F(e->dcT2); F(e->matI);
  e->ref = e->mat = e->dcA = e->dcB = e->refT = e->dcT = e->matT = e->dcT2 = e->matI = nullptr;
  e->dispv = e->dprev = nullptr;
  for (int s = 0; s < 2; ++s)
    for (int k = 0; k < 2; ++k) {
      F(e->d_in[s][k]); F(e->d_seed[s][k]); F(e->d_out[s][k]);
      e->d_in[s][k] = nullptr; e->d_seed[s][k] = nullptr; e->d_out[s][k] = nullptr;
    }
  e->seed_alloc = false;
  F(e->seed.kps); F(e->seed.kpd); F(e->seed.nkp); F(e->seed.ncand); F(e->seed.vmax); F(e->seed.status);
  F(e->seed_map[0]); F(e->seed_map[1]);
  e->seed = SeedState{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  e->seed_map[0] = e->seed_map[1] = nullptr;
  e->seed_views = e->seed_maxf = 0;
  e->w = e->h = e->nb = 0;
  e->stage_loaded = false;
  F(e->band.send_prev); F(e->band.recv_prev); F(e->band.send_next); F(e->band.recv_next);
  F(e->band.out[0]); F(e->band.out[1]);
  e->band = pm_engine::Band();
}

// h is the height the column sweeps are chunked over: the frame's, also for a row band.
int validate_size(pm_engine* e, int w, int h) {
  const pm_params& p = e->p;
  if (w < 8 || h < 8) return fail(e, PM_ERR_INVALID_ARG, "image %dx%d is too small", w, h);
  const int lw = w >> (p.pyramid_levels - 1), lh = h >> (p.pyramid_levels - 1);
  const int need = 2 * p.sweep_overlap + 2;
  if (lw / p.sweep_chunks < need || lh / p.sweep_chunks < need)
    return fail(e, PM_ERR_UNSUPPORTED,
                "coarsest level %dx%d gives sweep chunks shorter than 2*overlap+2 = %d pixels "
                "(sweep_chunks = %d): the lock-step schedule is undefined there",
                lw, lh, need, p.sweep_chunks);
  return PM_OK;
}

// (Re)allocates the workspace for nb pairs of w x h. Buffers are zeroed once: the
// pad element after each row must stay finite (pm_device.cuh, lerp_ig).
// band_frame_h > 0: the planes hold rows [band_load_lo, band_load_lo + h) of a frame of
// band_frame_h rows (one pair, one pyramid level).
// sweeps = false: the caller only runs the seeding kernels, which have no chunking constraint.
int ensure_workspace(pm_engine* e, int w, int h, int nb, bool host_path, bool need_seed,
                     int band_frame_h = 0, int band_load_lo = 0, bool sweeps = true) {
  const bool band = band_frame_h > 0;
  if (sweeps) {
    if (int rc = validate_size(e, w, band ? band_frame_h : h)) return rc;
  } else if (w < 8 || h < 8) {
    return fail(e, PM_ERR_INVALID_ARG, "image %dx%d is too small", w, h);
  }
  if (e->band.running && !band)
    return fail(e, PM_ERR_STATE, "a row-band pass is in flight: call pm_band_finish first");
  const bool same = (w == e->w && h == e->h && nb <= e->nb && band == e->band.ws &&
                     (!band || (band_frame_h == e->band.frame_h && band_load_lo == e->band.load_lo)));
  if (same && (!host_path || e->d_in[0][0]) && (!need_seed || !host_path || e->seed_alloc))
    return PM_OK;
  if (!same) {
    PM_CUDA(e, cudaDeviceSynchronize());
    free_workspace(e);
    e->w = w; e->h = h; e->nb = nb;
    e->levels = band ? 1 : e->p.pyramid_levels;
    e->band.ws = band;
    e->band.frame_h = band_frame_h;
    e->band.load_lo = band_load_lo;
    const size_t V = 2 * (size_t)nb;
    for (int l = 0; l < e->levels; ++l) {
      Level& L = e->lv[l];
      L.w = w >> l; L.h = h >> l;
      L.pitch = round_up(L.w + 1, 16);
      L.plane = (size_t)L.pitch * L.h;
      L.pitch8 = round_up(L.w, 16);
      L.plane8 = (size_t)L.pitch8 * L.h;
      L.npitch = round_up(L.w, 32);
      L.pitchT = round_up(L.h, 16);
      L.planeT = (size_t)L.pitchT * (L.w + 1);  // + the pad column of the matched plane
      // the block sweep kernels evaluate the reference's 5-tap cost; other cost modes run the
      // one-thread-per-chain kernel
      const bool x5 = e->p.cost_mode == PM_COST_L1GRAD_X5 && e->p.patch_size == 3;
      static const bool force_rowT = [] { const char* v = getenv("PM_FORCE_ROWT"); return v && v[0] == '1'; }();
      L.row_smem = x5 && !force_rowT && sweep_row_supported(L.w, e->p.sweep_chunks, e->p.sweep_overlap);
      L.row_T = x5 && !L.row_smem && sweep_rowT_supported(L.w, e->p.sweep_chunks, e->p.sweep_overlap);
      // the block column kernel owns all chunks of a column: whole frames only
      L.col_block = x5 && !band && sweep_col_supported(L.h, e->p.sweep_chunks, e->p.sweep_overlap);
      L.col_inplace = L.col_block && sweep_col_inplace_supported(L.w, L.h, e->p.sweep_chunks, e->p.sweep_overlap);
      if (l > 0) {
        PM_CUDA(e, cudaMalloc(&L.L8, L.plane8 * nb));
        PM_CUDA(e, cudaMalloc(&L.R8, L.plane8 * nb));
      }
      PM_CUDA(e, cudaMalloc(&L.noise, (size_t)L.npitch * L.h * sizeof(float)));
      PM_CUDA(e, cudaMemsetAsync(L.noise, 0, (size_t)L.npitch * L.h * sizeof(float), e->stream));
      PM_LAUNCH(e, launch_noise_image(L.noise, L.w, L.h, L.npitch, e->p.seed,
                                      (long)band_load_lo * L.w, e->stream));
      L.fuse_noise = L.row_smem && e->p.noise_accept == PM_NOISE_ALWAYS &&
                     sweep_row_fuses_noise(L.w, e->p.sweep_chunks, e->p.sweep_overlap);
      L.row_rm = L.row_smem && sweep_row_reads_rowmajor(L.w, e->p.sweep_chunks, e->p.sweep_overlap);
      L.row_il = L.row_smem && sweep_row_interleaved(L.w, e->p.sweep_chunks, e->p.sweep_overlap);
      L.planeI = L.row_il ? sweep_row_interleaved_plane(L.w, L.h) : 0;
      if (L.fuse_noise) {
        PM_CUDA(e, cudaMalloc(&L.noiseT, (size_t)L.pitchT * L.w * sizeof(float)));
        PM_CUDA(e, cudaMemsetAsync(L.noiseT, 0, (size_t)L.pitchT * L.w * sizeof(float), e->stream));
        PM_LAUNCH(e, launch_transpose1(L.noise, L.w, L.h, L.npitch, L.noiseT, L.pitchT, e->stream));
      }
    }
    const size_t plane0 = e->lv[0].plane;
    const size_t bytes2 = plane0 * V * sizeof(float2) + 256;
    PM_CUDA(e, cudaMalloc(&e->ref, bytes2));
    PM_CUDA(e, cudaMalloc(&e->mat, bytes2));
    PM_CUDA(e, cudaMalloc(&e->dcA, bytes2));
    PM_CUDA(e, cudaMalloc(&e->dcB, bytes2));
    const size_t bytesT = e->lv[0].planeT * V * sizeof(float2) + 256;
    PM_CUDA(e, cudaMalloc(&e->refT, bytesT));
    PM_CUDA(e, cudaMalloc(&e->dcT, bytesT));
    PM_CUDA(e, cudaMemsetAsync(e->refT, 0, bytesT, e->stream));
    PM_CUDA(e, cudaMemsetAsync(e->dcT, 0, bytesT, e->stream));
    bool any_T = false;
    for (int l = 0; l < e->levels; ++l) any_T = any_T || e->lv[l].row_T;
    if (any_T) {
      PM_CUDA(e, cudaMalloc(&e->matT, bytesT));
      PM_CUDA(e, cudaMalloc(&e->dcT2, bytesT));
      PM_CUDA(e, cudaMemsetAsync(e->matT, 0, bytesT, e->stream));
      PM_CUDA(e, cudaMemsetAsync(e->dcT2, 0, bytesT, e->stream));
    }
    size_t bytesI = 0;
    for (int l = 0; l < e->levels; ++l) bytesI = std::max(bytesI, e->lv[l].planeI * V * sizeof(float2));
    if (bytesI) {
      PM_CUDA(e, cudaMalloc(&e->matI, bytesI + 256));
      PM_CUDA(e, cudaMemsetAsync(e->matI, 0, bytesI + 256, e->stream));
    }
    PM_CUDA(e, cudaMalloc(&e->dispv, plane0 * V * sizeof(float)));
    PM_CUDA(e, cudaMalloc(&e->dprev, plane0 * V * sizeof(float)));
    PM_CUDA(e, cudaMemsetAsync(e->ref, 0, bytes2, e->stream));
    PM_CUDA(e, cudaMemsetAsync(e->mat, 0, bytes2, e->stream));
    PM_CUDA(e, cudaMemsetAsync(e->dcA, 0, bytes2, e->stream));
    PM_CUDA(e, cudaMemsetAsync(e->dcB, 0, bytes2, e->stream));
    PM_CUDA(e, cudaMemsetAsync(e->dispv, 0, plane0 * V * sizeof(float), e->stream));
    PM_CUDA(e, cudaMemsetAsync(e->dprev, 0, plane0 * V * sizeof(float), e->stream));
  }
  if (host_path && !e->d_in[0][0]) {
    const Level& L0 = e->lv[0];
    for (int s = 0; s < 2; ++s)
      for (int k = 0; k < 2; ++k) {
        PM_CUDA(e, cudaMalloc(&e->d_in[s][k], L0.plane8 * e->nb));
        PM_CUDA(e, cudaMalloc(&e->d_out[s][k], (size_t)L0.npitch * L0.h * e->nb * sizeof(float)));
      }
  }
  if (host_path && need_seed && !e->seed_alloc) {
    const Level& L0 = e->lv[0];
    for (int s = 0; s < 2; ++s)
      for (int k = 0; k < 2; ++k)
        PM_CUDA(e, cudaMalloc(&e->d_seed[s][k], (size_t)L0.npitch * L0.h * e->nb * sizeof(float)));
    e->seed_alloc = true;
  }
  PM_CUDA(e, cudaStreamSynchronize(e->stream));
  return PM_OK;
}

// CUDA-event span around one stage on the launching stream. Nothing synchronises
// here: spans are resolved by pm_last_stage_ms after the stream has drained.
struct StageTimer {
  pm_engine* e;
  cudaStream_t st;
  pm_engine::ProfSpan span;
  static cudaEvent_t get(pm_engine* e) {
    if (!e->ev_pool.empty()) { cudaEvent_t ev = e->ev_pool.back(); e->ev_pool.pop_back(); return ev; }
    cudaEvent_t ev = nullptr;
    cudaEventCreate(&ev);
    return ev;
  }
  StageTimer(pm_engine* e_, cudaStream_t st_, int stage) : e(e_), st(st_) {
    span.stage = stage; span.a = span.b = nullptr;
    if (e->profiling) { span.a = get(e); span.b = get(e); cudaEventRecord(span.a, st); }
  }
  ~StageTimer() {
    if (span.a) { cudaEventRecord(span.b, st); e->spans.push_back(span); }
  }
};

void resolve_spans(pm_engine* e) {
  for (auto& sp : e->spans) {
    float ms = 0;
    if (cudaEventSynchronize(sp.b) == cudaSuccess && cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) {
      e->stage_ms[sp.stage] += ms;
      e->stage_launches[sp.stage] += 1;
    }
    e->ev_pool.push_back(sp.a);
    e->ev_pool.push_back(sp.b);
  }
  e->spans.clear();
}

float noise_scale(const pm_params& p, int level, int git) {
  // 32.0 / pow(2.0, iter) (patchmatch_gpu.cu:395), continued across pyramid levels
  return (float)((double)p.noise_scale0 * (1.0 / (double)(1 << level)) / std::pow(2.0, (double)git));
}

// One sweep over `nviews` views starting at view `v0`: dcA -> dcB, then the two are swapped
// (whole-plane pointers, so callers always sweep all resident views or copy back).
// noise_scale > 0 (row sweeps of levels with fuse_noise only): the sweep first applies
// AddForegroundNoise + the cost refresh, i.e. src is the plane before the noise.
// in_place != nullptr: the caller accepts a sweep that leaves its result in `src` (the in-place
// column kernel); *in_place says whether that happened. Otherwise the result is always in `dst`.
int sweep_views(pm_engine* e, const Level& L, int nviews, size_t v0, int along_x, int dir,
                float2* src, float2* dst, cudaStream_t st, float noise_scale = 0.0f,
                float noise_dmax = 0.0f, bool* in_place = nullptr) {
  if (in_place) *in_place = false;
  const pm_params& p = e->p;
  const ViewGeom g = geom(e, L);
  const SweepParams sp{p.sweep_chunks, p.sweep_overlap, p.cost_alpha};
  const size_t vo = v0 * L.plane, voT = v0 * L.planeT;
  if (along_x && L.row_smem) {
    if (!L.row_rm) {
      StageTimer t(e, st, ST_COPY);
      PM_LAUNCH(e, launch_transpose2(src + vo, L.w, L.h, L.pitch, L.plane, e->dcT + voT, L.pitchT,
                                     L.planeT, nviews, st));
    }
    StageTimer t(e, st, ST_SWEEP_ROW);
    PM_LAUNCH(e, launch_sweep_row(e->refT + voT, e->mat + vo, e->dcT + voT, dst + vo, g, L.pitchT,
                                  L.planeT, nviews, dir, sp, st,
                                  noise_scale > 0.0f ? L.noiseT : nullptr, noise_scale, noise_dmax,
                                  L.row_rm ? src + vo : nullptr,
                                  L.row_il ? e->matI + v0 * L.planeI : nullptr, L.planeI));
    return PM_OK;
  }
  if (along_x && L.row_T) {
    // wide images: the row sweep is the column kernel on transposed planes; the result comes
    // back transposed and is turned row-major again
    {
      StageTimer t(e, st, ST_COPY);
      PM_LAUNCH(e, launch_transpose2(src + vo, L.w, L.h, L.pitch, L.plane, e->dcT + voT, L.pitchT,
                                     L.planeT, nviews, st));
    }
    const bool inpl = sweep_col_inplace_supported(L.h, L.w, sp.chunks, sp.overlap);
    {
      StageTimer t(e, st, ST_SWEEP_ROW);
      if (inpl)
        PM_LAUNCH(e, launch_sweep_rowT_inplace(e->refT + voT, e->matT + voT, e->dcT + voT, g, L.pitchT,
                                               L.planeT, nviews, dir, sp, st));
      else
        PM_LAUNCH(e, launch_sweep_rowT(e->refT + voT, e->matT + voT, e->dcT + voT, e->dcT2 + voT, g,
                                       L.pitchT, L.planeT, nviews, dir, sp, st));
    }
    StageTimer t(e, st, ST_COPY);
    PM_LAUNCH(e, launch_transpose2((inpl ? e->dcT : e->dcT2) + voT, L.h, L.w, L.pitchT, L.planeT,
                                   dst + vo, L.pitch, L.plane, nviews, st));
    return PM_OK;
  }
  if (!along_x && L.col_inplace && in_place) {
    StageTimer t(e, st, ST_SWEEP_COL);
    PM_LAUNCH(e, launch_sweep_col_inplace(e->ref + vo, e->mat + vo, src + vo, g, nviews, dir, sp, st));
    *in_place = true;
    return PM_OK;
  }
  if (!along_x && L.col_block) {
    StageTimer t(e, st, ST_SWEEP_COL);
    PM_LAUNCH(e, launch_sweep_col(e->ref + vo, e->mat + vo, src + vo, dst + vo, g, nviews, dir, sp, st));
    return PM_OK;
  }
  {
    StageTimer t(e, st, ST_COPY);
    PM_CUDA(e, cudaMemcpyAsync(dst + vo, src + vo, L.plane * nviews * sizeof(float2),
                               cudaMemcpyDeviceToDevice, st));
  }
  StageTimer t(e, st, along_x ? ST_SWEEP_ROW : ST_SWEEP_COL);
  PM_LAUNCH(e, launch_sweep(e->ref + vo, e->mat + vo, src + vo, dst + vo, g, nviews, along_x, dir,
                            sp, st, e->band.ws ? e->band.k_lo : 0, e->band.ws ? e->band.nk : 0));
  return PM_OK;
}

int run_sweep(pm_engine* e, const Level& L, int nviews, size_t v0, int along_x, int dir,
              cudaStream_t st, float noise_scale = 0.0f, float noise_dmax = 0.0f) {
  bool in_place = false;
  if (int rc = sweep_views(e, L, nviews, v0, along_x, dir, e->dcA, e->dcB, st, noise_scale, noise_dmax,
                           &in_place))
    return rc;
  if (!in_place) std::swap(e->dcA, e->dcB);
  return PM_OK;
}

// AddForegroundNoise of global iteration `it` at level l (patchmatch_gpu.cu:395) + cost refresh.
int run_noise(pm_engine* e, int l, int nviews, int it, cudaStream_t st) {
  const pm_params& p = e->p;
  const Level& L = e->lv[l];
  const float dmax = p.clamp_disp ? (float)p.max_disp / (float)(1 << l) : INFINITY;
  const int iter0 = (e->levels - 1 - l) * p.patchmatch_iters;
  StageTimer t(e, st, ST_NOISE);
  if (it < 0)  // no iterations: only evaluate the cost of the initial disparity
    PM_LAUNCH(e, launch_noise_cost(e->ref, e->mat, e->dcA, geom(e, L), nviews, L.noise, L.npitch,
                                   0.0f, INFINITY, 0, p.cost_alpha, st));
  else
    PM_LAUNCH(e, launch_noise_cost(e->ref, e->mat, e->dcA, geom(e, L), nviews, L.noise, L.npitch,
                                   noise_scale(p, l, iter0 + it), dmax, p.noise_accept,
                                   p.cost_alpha, st));
  return PM_OK;
}

// The iterations of PatchmatchGpu::Match (device overload, patchmatch_gpu.cu:394-404)
// on the views currently held in dcA at pyramid level l.
int run_iterations(pm_engine* e, int l, int nviews, cudaStream_t st, uint32_t first_pair = 0) {
  const pm_params& p = e->p;
  const Level& L = e->lv[l];
  if (p.patchmatch_iters == 0) return run_noise(e, l, nviews, -1, st);
  const float dmax = p.clamp_disp ? (float)p.max_disp / (float)(1 << l) : INFINITY;
  const int iter0 = (e->levels - 1 - l) * p.patchmatch_iters;
  for (int it = 0; it < p.patchmatch_iters; ++it) {
    // the first sweep of an iteration (row, +1) can apply the noise itself
    const float fused = L.fuse_noise ? noise_scale(p, l, iter0 + it) : 0.0f;
    if (!(fused > 0.0f))
      if (int rc = run_noise(e, l, nviews, it, st)) return rc;
    for (int s = 0; s < 4; ++s) {
      const int along_x = (s % 2 == 0), dir = s < 2 ? +1 : -1;
      if (int rc = run_sweep(e, L, nviews, 0, along_x, dir, st, s == 0 ? fused : 0.0f, dmax)) return rc;
    }
    if (p.random_search_k > 0) {   // extension: K random-search candidates per pixel
      StageTimer t(e, st, ST_NOISE);
      PM_LAUNCH(e, launch_random_search(e->ref, e->mat, e->dcA, geom(e, L), nviews, p.seed, first_pair,
                                        (uint32_t)l, (uint32_t)(iter0 + it), p.random_search_k,
                                        noise_scale(p, l, iter0 + it), dmax, p.cost_alpha, st));
    }
  }
  return PM_OK;
}

// MaskBackground (+ subpixel), flip + MaskOcclusions (+ median) on the level-0 views held in
// dcA, written to the caller's maps (patchmatch_gpu.cu:406-410, 368-375).
int finish_level0(pm_engine* e, int nb, float* dOutL, float* dOutR, size_t opitch_bytes,
                  size_t oplane_bytes, cudaStream_t st) {
  const pm_params& p = e->p;
  const Level& L = e->lv[0];
  const ViewGeom g = geom(e, L);
  const int V = 2 * nb;
  {
    StageTimer t(e, st, ST_MASK);
    PM_LAUNCH(e, launch_mask_background(e->ref, e->mat, e->dcA, g, V, p.cost_alpha,
                                        p.cost_improve_factor, 1, e->dispv, L.pitch, L.plane, st));
    if (p.subpixel)
      PM_LAUNCH(e, launch_subpixel(e->ref, e->mat, g, V, p.cost_alpha, e->dispv, L.pitch, L.plane, st));
  }
  StageTimer t(e, st, ST_FINAL);
  if (p.median_ksize == 3 || p.median_ksize == 5) {
    // finalize into dprev-backed dense maps, then median into the caller's buffers
    float* tl = e->dprev;
    float* tr = e->dprev + (size_t)nb * L.plane;
    const size_t tp = (size_t)L.pitch * sizeof(float), tpl = L.plane * sizeof(float);
    PM_LAUNCH(e, launch_finalize(e->dispv, L.pitch, L.plane, L.w, L.h, nb, p.lr_mode, tl, tr, tp,
                                 tpl, st));
    // median needs equal pitches on both sides: go through dispv as a second temp
    float* ml = e->dispv;
    float* mr = e->dispv + (size_t)nb * L.plane;
    PM_LAUNCH(e, launch_median(tl, ml, L.w, L.h, tp, tpl, nb, p.median_ksize, st));
    PM_LAUNCH(e, launch_median(tr, mr, L.w, L.h, tp, tpl, nb, p.median_ksize, st));
    PM_CUDA(e, cudaMemcpy2DAsync(dOutL, opitch_bytes, ml, tp, L.w * sizeof(float),
                                 (size_t)L.h * nb, cudaMemcpyDeviceToDevice, st));
    PM_CUDA(e, cudaMemcpy2DAsync(dOutR, opitch_bytes, mr, tp, L.w * sizeof(float),
                                 (size_t)L.h * nb, cudaMemcpyDeviceToDevice, st));
  } else {
    PM_LAUNCH(e, launch_finalize(e->dispv, L.pitch, L.plane, L.w, L.h, nb, p.lr_mode, dOutL, dOutR,
                                 opitch_bytes, oplane_bytes, st));
  }
  return PM_OK;
}

// ---- sparse seeding: PatchmatchGpu::SparseInit (patchmatch_gpu.cu:414-442) on the device

// The detector / matcher parameters the seeding kernels support, checked when seeding is used.
int check_seed_params(pm_engine* e, int w, int h) {
  const pm_params& p = e->p;
  if (p.fd_max_features_per_frame < 1 || p.fd_max_features_per_frame > kMaxSeedFeatures)
    return fail(e, PM_ERR_UNSUPPORTED, "max_features_per_frame %d outside [1, %d]",
                p.fd_max_features_per_frame, kMaxSeedFeatures);
  if (p.fd_min_distance < 0) return fail(e, PM_ERR_INVALID_ARG, "min_distance %d", p.fd_min_distance);
  if (p.fd_gftt_block_size < 1 || p.fd_gftt_block_size > 15)
    return fail(e, PM_ERR_UNSUPPORTED, "gftt_block_size %d outside [1, 15]", p.fd_gftt_block_size);
  if (!(p.fd_gftt_quality_level > 0.0))
    return fail(e, PM_ERR_INVALID_ARG, "gftt_quality_level %g", p.fd_gftt_quality_level);
  if (p.sm_subpixel_refinement)
    return fail(e, PM_ERR_UNSUPPORTED, "StereoMatcher subpixel_refinement (cv::cornerSubPix, "
                "stereo_matcher.cpp:96-104) is not implemented; the PatchMatch drivers run with it off "
                "(patchmatch_gpu_test.cpp:76)");
  const int tc = p.sm_templ_cols, tr = p.sm_templ_rows, md = p.sm_max_disp;
  if (tc < 1 || tr < 1 || md < tc)
    return fail(e, PM_ERR_INVALID_ARG, "template %dx%d against a %d-pixel stripe", tc, tr, md);
  if ((size_t)tc * tr > 32768 ||
      4 * ((size_t)tr * ((tc + 3) / 4) + (size_t)(tr + 2) * ((md + 3) / 4 + 1)) > 48 * 1024)
    return fail(e, PM_ERR_UNSUPPORTED, "template %dx%d / max_disp %d exceed the matcher's tile", tc, tr, md);
  if (w < 2 * p.fd_gftt_block_size || h < 2 * p.fd_gftt_block_size)
    return fail(e, PM_ERR_UNSUPPORTED, "image %dx%d is smaller than two corner windows", w, h);
  if (w <= md || h <= tr + 2)
    return fail(e, PM_ERR_UNSUPPORTED, "image %dx%d is not larger than the search stripe %dx%d "
                "(cv::Mat ROI assertion in the reference, stereo_matcher.cpp:81)", w, h, md, tr + 2);
  return PM_OK;
}

int ensure_seed_ws(pm_engine* e, int nviews, bool maps) {
  const int maxf = e->p.fd_max_features_per_frame;
  if (e->seed_views < nviews || e->seed_maxf != maxf) {
    auto F = [](void* p) { if (p) cudaFree(p); };
    PM_CUDA(e, cudaDeviceSynchronize());  // the buffers may be in use on a caller's stream
    F(e->seed.kps); F(e->seed.kpd); F(e->seed.nkp); F(e->seed.ncand); F(e->seed.vmax); F(e->seed.status);
    e->seed = SeedState{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    e->seed_views = 0;
    PM_CUDA(e, cudaMalloc(&e->seed.kps, sizeof(int2) * (size_t)nviews * maxf));
    PM_CUDA(e, cudaMalloc(&e->seed.kpd, sizeof(float) * (size_t)nviews * maxf));
    PM_CUDA(e, cudaMalloc(&e->seed.nkp, sizeof(int) * nviews));
    PM_CUDA(e, cudaMalloc(&e->seed.ncand, sizeof(int) * nviews));
    PM_CUDA(e, cudaMalloc(&e->seed.vmax, sizeof(unsigned) * nviews));
    PM_CUDA(e, cudaMalloc(&e->seed.status, sizeof(int)));
    PM_CUDA(e, cudaMemset(e->seed.status, 0, sizeof(int)));  // synchronous: any stream may run the seeding
    e->seed_views = nviews;
    e->seed_maxf = maxf;
  }
  if (maps && !e->seed_map[0]) {
    const Level& L0 = e->lv[0];
    const size_t bytes = (size_t)L0.npitch * L0.h * e->nb * sizeof(float);
    PM_CUDA(e, cudaMalloc(&e->seed_map[0], bytes));
    PM_CUDA(e, cudaMalloc(&e->seed_map[1], bytes));
  }
  return PM_OK;
}

// FeatureDetector::Detect + StereoMatcher::MatchRectified for `nviews` seeding problems over
// the level-0 images (even problems: L against R; odd: both read right-to-left). dprev and
// dispv are free before the first pyramid level is set up and serve as scratch.
int seed_keypoints(pm_engine* e, int nviews, const uint8_t* dL, const uint8_t* dR, size_t ipitch,
                   size_t iplane, bool match, cudaStream_t st) {
  const pm_params& p = e->p;
  const Level& L0 = e->lv[0];
  if (int rc = check_seed_params(e, L0.w, L0.h)) return rc;
  if (int rc = ensure_seed_ws(e, nviews, false)) return rc;
  const SeedImages im{dL, dR, ipitch, iplane, L0.w, L0.h};
  const SeedDetect det{p.fd_max_features_per_frame, p.fd_min_distance, p.fd_gftt_block_size,
                       p.fd_gftt_use_harris, p.fd_gftt_quality_level, p.fd_gftt_k};
  const size_t kplane = L0.plane / 2;  // 64-bit keys in a view's float plane
  int cap = 1;
  while ((size_t)cap * 2 <= kplane) cap *= 2;
  PM_LAUNCH(e, launch_seed_detect(im, nviews, det, e->dispv, L0.pitch, L0.plane,
                                  reinterpret_cast<unsigned long long*>(e->dprev), kplane, cap,
                                  e->seed, st));
  if (match) {
    const SeedMatch mp{p.sm_templ_cols, p.sm_templ_rows, p.sm_max_disp, p.sm_max_matching_cost};
    PM_LAUNCH(e, launch_seed_match(im, nviews, mp, p.fd_max_features_per_frame, e->seed, st));
  }
  return PM_OK;
}

// Both seed maps of every pair of a device pass -> e->seed_map (patchmatch_gpu.cu:335, 362-365).
int run_seeding(pm_engine* e, int nb, const uint8_t* dL, const uint8_t* dR, size_t ipitch,
                size_t iplane, cudaStream_t st) {
  const Level& L0 = e->lv[0];
  StageTimer t(e, st, ST_SEED);
  if (int rc = ensure_seed_ws(e, 2 * nb, true)) return rc;
  if (int rc = seed_keypoints(e, 2 * nb, dL, dR, ipitch, iplane, true, st)) return rc;
  const int radius = (1 << e->p.init_dilate_factor) + 1;  // (int)pow(2, f) + 1, :436
  PM_LAUNCH(e, launch_seed_paint(2 * nb, L0.w, L0.h, radius, L0.w, L0.h, 1.0f,
                                 e->p.fd_max_features_per_frame, e->seed, e->seed_map[0],
                                 e->seed_map[1], L0.npitch, (size_t)L0.npitch * L0.h, st));
  return PM_OK;
}

// upload/convertTo/GradientMagnitude/flip (patchmatch_gpu.cu:346-360) and the initial
// disparity of pyramid level l.
int setup_level(pm_engine* e, int l, int nb, const uint8_t* dL, const uint8_t* dR, size_t ipitch,
                size_t iplane, const float* dSeedL, const float* dSeedR, size_t spitch,
                size_t splane, uint32_t first_pair, cudaStream_t st) {
  const pm_params& p = e->p;
  const Level& L = e->lv[l];
  const ViewGeom g = geom(e, L);
  const int V = 2 * nb;
  {
    StageTimer t(e, st, ST_PRE);
    const uint8_t* sl = l == 0 ? dL : L.L8;
    const uint8_t* sr = l == 0 ? dR : L.R8;
    const size_t sp8 = l == 0 ? ipitch : (size_t)L.pitch8, spl8 = l == 0 ? iplane : L.plane8;
    // PM_PRE_FUSED=0: the separate transpose / interleave passes (diagnosis)
    static const bool fuse_on = [] { const char* v = getenv("PM_PRE_FUSED"); return !(v && v[0] == '0'); }();
    const int icols = sweep_row_interleaved_cols(L.w);
    if (fuse_on && L.row_smem && L.row_il && !L.row_T &&
        preprocess_fused_supported(sl, sr, sp8, spl8, g, icols)) {
      // one pass writes the row-major planes, refT and matI of both views
      PM_LAUNCH(e, launch_preprocess_fused(sl, sr, sp8, spl8, e->ref, e->mat, e->refT, L.pitchT,
                                           L.planeT, e->matI, icols, L.planeI, g, nb, st));
    } else {
      PM_LAUNCH(e, launch_preprocess(sl, sr, sp8, spl8, e->ref, e->mat, g, nb, st));
      if (L.row_smem || L.row_T)
        PM_LAUNCH(e, launch_transpose2(e->ref, L.w, L.h, L.pitch, L.plane, e->refT, L.pitchT,
                                       L.planeT, V, st));
      if (L.row_T)  // with the zeroed pad column, which becomes the last row of matT
        PM_LAUNCH(e, launch_transpose2(e->mat, L.w + 1, L.h, L.pitch, L.plane, e->matT, L.pitchT,
                                       L.planeT, V, st));
      if (L.row_il) PM_LAUNCH(e, launch_interleave16(e->mat, g, V, e->matI, st));
    }
  }
  StageTimer t(e, st, ST_INIT);
  if (l == e->levels - 1) {
    if (p.init_mode == PM_INIT_RANDOM) {
      PM_LAUNCH(e, launch_init_random(e->dcA, g, V, p.seed, first_pair, (uint32_t)l,
                                      (float)p.max_disp / (float)(1 << l), st));
    } else {
      PM_LAUNCH(e, launch_init_seeds(e->dcA, g, nb, dSeedL, dSeedR, spitch, splane, l, st));
    }
  } else {
    // the coarser level left its {d, cost} plane in dcA (coarse geometry): the finer level's initial
    // plane is written into dcB straight from it, and the two swap
    const Level& P = e->lv[l + 1];
    PM_LAUNCH(e, launch_upsample2(e->dcB, g, V, reinterpret_cast<const float*>(e->dcA), P.w, P.h,
                                  P.pitch, P.plane, st, 2));
    std::swap(e->dcA, e->dcB);
  }
  return PM_OK;
}

// One device pass over nb pairs (patchmatch_gpu.cu:331-376 without the host round trips).
int run_device(pm_engine* e, int nb, const uint8_t* dL, const uint8_t* dR, size_t ipitch,
               size_t iplane, const float* dSeedL, const float* dSeedR, size_t spitch,
               size_t splane, uint32_t first_pair, float* dOutL, float* dOutR,
               size_t opitch_bytes, size_t oplane_bytes, cudaStream_t st) {
  const pm_params& p = e->p;
  const int V = 2 * nb;
  if (p.init_mode == PM_INIT_SPARSE && !dSeedL) {  // SparseInit on the device
    if (int rc = run_seeding(e, nb, dL, dR, ipitch, iplane, st)) return rc;
    dSeedL = e->seed_map[0];
    dSeedR = e->seed_map[1];
    spitch = e->lv[0].npitch;
    splane = (size_t)e->lv[0].npitch * e->lv[0].h;
  }
  // pyramid (extension): level l = level l-1 resized by 1/2
  {
    StageTimer t(e, st, ST_PRE);
    for (int l = 1; l < e->levels; ++l) {
      const Level& S = e->lv[l - 1];
      const Level& D = e->lv[l];
      const uint8_t* sl = l == 1 ? dL : S.L8;
      const uint8_t* sr = l == 1 ? dR : S.R8;
      const size_t sp_ = l == 1 ? ipitch : (size_t)S.pitch8, spl = l == 1 ? iplane : S.plane8;
      PM_LAUNCH(e, launch_downscale2(sl, S.w, S.h, sp_, spl, D.L8, D.pitch8, D.plane8, nb, st));
      PM_LAUNCH(e, launch_downscale2(sr, S.w, S.h, sp_, spl, D.R8, D.pitch8, D.plane8, nb, st));
    }
  }
  for (int l = e->levels - 1; l >= 0; --l) {
    if (int rc = setup_level(e, l, nb, dL, dR, ipitch, iplane, dSeedL, dSeedR, spitch, splane,
                             first_pair, st)) return rc;
    if (int rc = run_iterations(e, l, V, st, first_pair)) return rc;
    if (l == 0)
      if (int rc = finish_level0(e, nb, dOutL, dOutR, opitch_bytes, oplane_bytes, st)) return rc;
  }
  return PM_OK;
}


int check_params(const pm_params* p, std::string* why) {
  char b[256];
#define BAD(...) do { snprintf(b, sizeof(b), __VA_ARGS__); *why = b; return PM_ERR_INVALID_ARG; } while (0)
  if (!(p->cost_alpha >= 0.f && p->cost_alpha <= 1.f)) BAD("cost_alpha %g outside [0,1]", p->cost_alpha);
  if (p->patchmatch_iters < 0 || p->patchmatch_iters > 64) BAD("patchmatch_iters %d", p->patchmatch_iters);
  if (p->patch_size != 3 && p->patch_size != 5) {
    snprintf(b, sizeof(b), "patch_size %d: 3 (the reference's launch sites, patchmatch_gpu.cu:397-408) "
             "and 5 are supported", p->patch_size);
    *why = b;
    return PM_ERR_UNSUPPORTED;
  }
  if (p->random_search_k < 0 || p->random_search_k > 16) BAD("random_search_k %d outside [0,16]", p->random_search_k);
  if (p->sweep_chunks < 1 || p->sweep_chunks > 64) BAD("sweep_chunks %d", p->sweep_chunks);
  if (p->sweep_overlap < 0 || p->sweep_overlap > 8) BAD("sweep_overlap %d outside [0,8]", p->sweep_overlap);
  if (p->pyramid_levels < 1 || p->pyramid_levels > kMaxLevels) BAD("pyramid_levels %d", p->pyramid_levels);
  if (p->init_mode != PM_INIT_SPARSE && p->init_mode != PM_INIT_RANDOM) BAD("init_mode %d", p->init_mode);
  if (p->init_dilate_factor < 0 || p->init_dilate_factor > 12) BAD("init_dilate_factor %d", p->init_dilate_factor);
  if (p->cost_mode != PM_COST_L1GRAD_X5 && p->cost_mode != PM_COST_L1GRAD_FULL &&
      p->cost_mode != PM_COST_CENSUS) BAD("cost_mode %d", p->cost_mode);
  if (p->lr_mode != PM_LR_RATIO && p->lr_mode != PM_LR_ABS1PX) BAD("lr_mode %d", p->lr_mode);
  if (p->noise_accept != PM_NOISE_ALWAYS && p->noise_accept != PM_NOISE_IMPROVE) BAD("noise_accept %d", p->noise_accept);
  if (p->median_ksize != 0 && p->median_ksize != 3 && p->median_ksize != 5) BAD("median_ksize %d", p->median_ksize);
  if (p->max_disp < 1) BAD("max_disp %d", p->max_disp);
  if (p->max_batch < 0) BAD("max_batch %d", p->max_batch);
#undef BAD
  return PM_OK;
}

int auto_batch(const pm_engine* e, int w, int h, int n, bool host_path) {
  if (e->p.max_batch > 0) return std::min(n, e->p.max_batch);
  // Device-resident batches: as many pairs per pass as a third of the GPU's memory holds (~120 MB of
  // planes per 1280x720 pair; at most 512 pairs): every sweep launch ends with a partly filled wave of
  // blocks, and the longer the launch the less that tail weighs. Measured at 1280x720, 512 pairs:
  // 16 pairs per pass 2040, 64: 2725, 148: 2778, 256: 2781, 512: 2807 pairs/s. 180 GB of HBM3e is
  // what makes the large pass affordable.
  const double px = (double)w * h;
  const double per_pair = 132.0 * px * (e->p.pyramid_levels > 1 ? 1.05 : 1.0);
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { (void)cudaGetLastError(); total_b = free_b = 0; }
  double budget = std::min(64.0e9, (double)total_b / 3.0);
  const bool have = e->w == w && e->h == h && e->nb > 0 && !e->band.ws;
  if (have) budget = std::max(budget, per_pair * e->nb);          // what is allocated already is paid for
  else budget = std::min(budget, 0.6 * (double)free_b);           // do not crowd out another tenant
  int nb = (int)std::max(1.0, std::min(512.0, budget / per_pair));
  if (total_b == 0) nb = (int)std::max(1.0, std::min(64.0, 64.0e6 / px));
  // Host batches are cut into at least four passes so that the upload of pass k+1 and the
  // download of pass k-1 overlap the kernels of pass k.
  if (host_path) nb = std::min(nb, std::max(1, (n + 3) / 4));
  return std::min(n, nb);
}

}  // namespace

// ------------------------------------------------------------------- C ABI

extern "C" {

int pm_abi_version(void) { return PM_B200_ABI_VERSION; }

static int seed_status(pm_engine* e);
static int check_rig(pm_engine* e, const pm_stereo_rig* rig, double scale);

int pm_params_default(pm_params* p) {
  if (!p) return PM_ERR_INVALID_ARG;
  std::memset(p, 0, sizeof(*p));
  p->cost_alpha = 0.9f;
  p->patchmatch_iters = 3;
  p->init_dilate_factor = 4;
  p->cost_improve_factor = 0.8f;
  p->sm_templ_cols = 31;
  p->sm_templ_rows = 11;
  p->sm_max_disp = 128;
  p->sm_max_matching_cost = 0.15;
  p->sm_bidirectional = 0;
  p->sm_subpixel_refinement = 0;
  p->fd_max_features_per_frame = 200;
  p->fd_min_distance = 20;
  p->fd_gftt_quality_level = 0.01;
  p->fd_gftt_block_size = 5;
  p->fd_gftt_use_harris = 0;
  p->fd_gftt_k = 0.04;
  p->patch_size = 3;
  p->sweep_chunks = 16;
  p->sweep_overlap = 5;
  p->noise_scale0 = 32.0f;
  p->seed = 123;
  p->init_mode = PM_INIT_SPARSE;
  p->max_disp = 128;
  p->clamp_disp = 0;
  p->pyramid_levels = 1;
  p->cost_mode = PM_COST_L1GRAD_X5;
  p->lr_mode = PM_LR_RATIO;
  p->noise_accept = PM_NOISE_ALWAYS;
  p->subpixel = 0;
  p->median_ksize = 0;
  p->max_batch = 0;
  p->random_search_k = 0;
  return PM_OK;
}

int pm_create(const pm_params* params, int device, pm_engine** out) {
  if (!params || !out) return fail(nullptr, PM_ERR_INVALID_ARG, "pm_create: null argument");
  *out = nullptr;
  std::string why;
  if (int rc = check_params(params, &why)) return fail(nullptr, rc, "pm_create: %s", why.c_str());
  int ndev = 0;
  cudaError_t st = cudaGetDeviceCount(&ndev);
  if (st != cudaSuccess || ndev == 0)
    return fail(nullptr, PM_ERR_CUDA, "pm_create: no CUDA device (%s); this engine has no CPU path",
                cudaGetErrorString(st));
  if (device < 0 || device >= ndev)
    return fail(nullptr, PM_ERR_INVALID_ARG, "pm_create: device %d of %d", device, ndev);
  pm_engine* e = new (std::nothrow) pm_engine();
  if (!e) return fail(nullptr, PM_ERR_OOM, "pm_create: out of host memory");
  e->p = *params;
  e->device = device;
  auto bail = [&](const char* what, cudaError_t s) {
    std::string m = std::string(what) + ": " + cudaGetErrorString(s);
    pm_destroy(e);
    return fail(nullptr, PM_ERR_CUDA, "pm_create: %s", m.c_str());
  };
  if ((st = cudaSetDevice(device)) != cudaSuccess) return bail("cudaSetDevice", st);
  if ((st = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", st);
  if ((st = cudaStreamCreateWithFlags(&e->s_in, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", st);
  if ((st = cudaStreamCreateWithFlags(&e->s_out, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", st);
  if ((st = cudaEventCreateWithFlags(&e->ev_ws, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", st);
  for (int s = 0; s < 2; ++s) {
    if ((st = cudaEventCreateWithFlags(&e->ev_in[s], cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", st);
    if ((st = cudaEventCreateWithFlags(&e->ev_done[s], cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", st);
    if ((st = cudaEventCreateWithFlags(&e->ev_out[s], cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", st);
  }
  *out = e;
  return PM_OK;
}

int pm_destroy(pm_engine* e) {
  if (!e) return PM_OK;
  cudaSetDevice(e->device);
  cudaDeviceSynchronize();
  free_workspace(e);
  p2p_close(e);
  if (e->p2p.region) cudaFree(e->p2p.region);
  for (int s = 0; s < 2; ++s) {
    if (e->ev_in[s]) cudaEventDestroy(e->ev_in[s]);
    if (e->ev_done[s]) cudaEventDestroy(e->ev_done[s]);
    if (e->ev_out[s]) cudaEventDestroy(e->ev_out[s]);
  }
  if (e->ev_ws) cudaEventDestroy(e->ev_ws);
  resolve_spans(e);
  for (cudaEvent_t ev : e->ev_pool) cudaEventDestroy(ev);
  if (e->stream) cudaStreamDestroy(e->stream);
  if (e->s_in) cudaStreamDestroy(e->s_in);
  if (e->s_out) cudaStreamDestroy(e->s_out);
  delete e;
  return PM_OK;
}

const char* pm_last_error(const pm_engine* e) { return e ? e->err.c_str() : g_create_error.c_str(); }

int pm_get_params(const pm_engine* e, pm_params* out) {
  if (!e || !out) return PM_ERR_INVALID_ARG;
  *out = e->p;
  return PM_OK;
}

int pm_host_alloc(size_t bytes, void** out) {
  if (!out) return PM_ERR_INVALID_ARG;
  return cudaHostAlloc(out, bytes, cudaHostAllocDefault) == cudaSuccess ? PM_OK : PM_ERR_OOM;
}

int pm_host_free(void* p) { return cudaFreeHost(p) == cudaSuccess ? PM_OK : PM_ERR_CUDA; }

int pm_launch_count(const pm_engine* e, uint64_t* out) {
  if (!e || !out) return PM_ERR_INVALID_ARG;
  *out = e->launches;
  return PM_OK;
}

int pm_launch_count_reset(pm_engine* e) {
  if (!e) return PM_ERR_INVALID_ARG;
  e->launches = 0;
  return PM_OK;
}

int pm_set_profiling(pm_engine* e, int on) {
  if (!e) return PM_ERR_INVALID_ARG;
  resolve_spans(e);
  e->profiling = on != 0;
  std::memset(e->stage_ms, 0, sizeof(e->stage_ms));
  std::memset(e->stage_launches, 0, sizeof(e->stage_launches));
  return PM_OK;
}

int pm_last_stage_ms(pm_engine* e, float* ms, uint32_t* spans) {
  if (!e || !ms) return PM_ERR_INVALID_ARG;
  resolve_spans(e);
  std::memcpy(ms, e->stage_ms, sizeof(e->stage_ms));
  if (spans) std::memcpy(spans, e->stage_launches, sizeof(e->stage_launches));
  return PM_OK;
}

const char* pm_stage_name(int i) { return (i >= 0 && i < PM_N_STAGES) ? kStageNames[i] : ""; }

static int check_io(pm_engine* e, int n, const void* l, const void* r, int w, int h, size_t stride,
                    const void* sl, const void* sr, const void* ol, const void* orr, size_t ostride) {
  if (!e) return PM_ERR_INVALID_ARG;
  if (n < 1 || !l || !r || !ol || !orr) return fail(e, PM_ERR_INVALID_ARG, "null image/output pointer or n < 1");
  if (w < 1 || h < 1 || stride < (size_t)w || ostride < (size_t)w * sizeof(float) || ostride % sizeof(float))
    return fail(e, PM_ERR_INVALID_ARG, "bad size/stride: %dx%d stride %zu out stride %zu", w, h, stride, ostride);
  if ((sl == nullptr) != (sr == nullptr))
    return fail(e, PM_ERR_INVALID_ARG, "seed_l and seed_r must be given together (or both NULL: "
                "SparseInit then runs on the device)");
  return PM_OK;
}

int pm_match_batch_device(pm_engine* e, int n, const uint8_t* d_left, const uint8_t* d_right,
                          int width, int height, size_t stride_bytes, const float* d_seed_l,
                          const float* d_seed_r, uint32_t first_pair_index, float* d_disp_l,
                          float* d_disp_r, size_t disp_stride_bytes, void* stream) {
  if (int rc = check_io(e, n, d_left, d_right, width, height, stride_bytes, d_seed_l, d_seed_r,
                        d_disp_l, d_disp_r, disp_stride_bytes)) return rc;
  PM_CUDA(e, cudaSetDevice(e->device));
  const int nb = auto_batch(e, width, height, n, false);
  if (int rc = ensure_workspace(e, width, height, nb, false, false)) return rc;
  cudaStream_t st = stream ? (cudaStream_t)stream : e->stream;
  if (int rc = ws_acquire(e, st)) return rc;
  const size_t iplane = stride_bytes * height, oplane = disp_stride_bytes * height;
  const size_t spitch = disp_stride_bytes / sizeof(float), splane = spitch * height;
  for (int i = 0; i < n; i += nb) {
    const int m = std::min(nb, n - i);
    if (int rc = run_device(e, m, d_left + i * iplane, d_right + i * iplane, stride_bytes, iplane,
                            d_seed_l ? d_seed_l + i * splane : nullptr,
                            d_seed_r ? d_seed_r + i * splane : nullptr, spitch, splane,
                            first_pair_index + (uint32_t)i,
                            (float*)((char*)d_disp_l + i * oplane),
                            (float*)((char*)d_disp_r + i * oplane), disp_stride_bytes, oplane, st))
      return rc;
  }
  return ws_release(e, st);
}

int pm_match_planes_device(pm_engine* e, const float* d_il, const float* d_ir, const float* d_gl,
                           const float* d_gr, int width, int height, size_t plane_stride_bytes,
                           float* d_disp, size_t disp_stride_bytes, void* stream) {
  if (!e) return PM_ERR_INVALID_ARG;
  if (!d_il || !d_ir || !d_gl || !d_gr || !d_disp)
    return fail(e, PM_ERR_INVALID_ARG, "pm_match_planes_device: null plane pointer");
  if (width < 1 || height < 1 || plane_stride_bytes < (size_t)width * sizeof(float) ||
      plane_stride_bytes % sizeof(float) || disp_stride_bytes < (size_t)width * sizeof(float) ||
      disp_stride_bytes % sizeof(float))
    return fail(e, PM_ERR_INVALID_ARG, "bad size/stride: %dx%d plane stride %zu disparity stride %zu",
                width, height, plane_stride_bytes, disp_stride_bytes);
  if (e->p.pyramid_levels != 1)
    return fail(e, PM_ERR_UNSUPPORTED, "the float-plane Match runs the planes it is given: pyramid_levels "
                "must be 1 (the reference has no pyramid, patchmatch_gpu.cu:379-411)");
  PM_CUDA(e, cudaSetDevice(e->device));
  if (int rc = ensure_workspace(e, width, height, std::max(1, e->nb * (e->w == width && e->h == height)),
                                false, false)) return rc;
  cudaStream_t st = stream ? (cudaStream_t)stream : e->stream;
  if (int rc = ws_acquire(e, st)) return rc;
  e->stage_loaded = false;
  const Level& L = e->lv[0];
  const ViewGeom g = geom(e, L);
  const size_t ip = plane_stride_bytes / sizeof(float), dp = disp_stride_bytes / sizeof(float);
  {
    StageTimer t(e, st, ST_PRE);
    PM_LAUNCH(e, launch_interleave_ig(d_il, d_gl, ip, e->ref, g, st));
    PM_LAUNCH(e, launch_interleave_ig(d_ir, d_gr, ip, e->mat, g, st));
    if (L.row_smem || L.row_T)
      PM_LAUNCH(e, launch_transpose2(e->ref, L.w, L.h, L.pitch, L.plane, e->refT, L.pitchT, L.planeT, 1, st));
    if (L.row_T)
      PM_LAUNCH(e, launch_transpose2(e->mat, L.w + 1, L.h, L.pitch, L.plane, e->matT, L.pitchT, L.planeT, 1, st));
    if (L.row_il) PM_LAUNCH(e, launch_interleave16(e->mat, g, 1, e->matI, st));
  }
  {
    StageTimer t(e, st, ST_INIT);   // the seed the caller left in disp (patchmatch_gpu.cu:354-355)
    PM_LAUNCH(e, launch_set_disp(e->dcA, g, 1, d_disp, (int)dp, 0, st));
  }
  if (int rc = run_iterations(e, 0, 1, st)) return rc;
  {
    StageTimer t(e, st, ST_MASK);   // MaskBackground, :406-410; the result replaces the seed
    PM_LAUNCH(e, launch_mask_background(e->ref, e->mat, e->dcA, g, 1, e->p.cost_alpha,
                                        e->p.cost_improve_factor, 1, d_disp, (int)dp, 0, st));
  }
  return ws_release(e, st);
}

int pm_mesh_vertices_host(pm_engine* e, const float* disp, int width, int height,
                          size_t disp_stride_bytes, const uint8_t* mask, size_t mask_stride_bytes,
                          const float* keypoints_xy, int n, const pm_stereo_rig* rig,
                          double scale_factor, float* vertex_disps, float* vertices_xyz) {
  if (!e) return PM_ERR_INVALID_ARG;
  if (!disp || !keypoints_xy || !vertex_disps || !vertices_xyz || n < 1 || width < 1 || height < 1 ||
      disp_stride_bytes < (size_t)width * sizeof(float) || disp_stride_bytes % sizeof(float) ||
      (mask && mask_stride_bytes < (size_t)width))
    return fail(e, PM_ERR_INVALID_ARG, "pm_mesh_vertices: null pointer or bad size/stride");
  if (int rc = check_rig(e, rig, scale_factor)) return rc;
  PM_CUDA(e, cudaSetDevice(e->device));
  const size_t dp = (size_t)round_up(width, 32), mp = (size_t)round_up(width, 16);
  char* d = nullptr;
  const size_t bytes_d = dp * height * sizeof(float), bytes_m = mask ? mp * height : 0;
  PM_CUDA(e, cudaMalloc(&d, bytes_d + bytes_m + (size_t)n * (8 + 4 + 12) + 256));
  float* dd = (float*)d;
  uint8_t* dm = mask ? (uint8_t*)(d + bytes_d) : nullptr;
  float2* dk = (float2*)(d + ((bytes_d + bytes_m + 15) & ~(size_t)15));
  float* od = (float*)(dk + n);
  float* ox = od + n;
  cudaStream_t st = e->stream;
  cudaError_t ce = cudaMemcpy2DAsync(dd, dp * 4, disp, disp_stride_bytes, (size_t)width * 4, height,
                                     cudaMemcpyHostToDevice, st);
  if (ce == cudaSuccess && mask)
    ce = cudaMemcpy2DAsync(dm, mp, mask, mask_stride_bytes, width, height, cudaMemcpyHostToDevice, st);
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(dk, keypoints_xy, (size_t)n * 8, cudaMemcpyHostToDevice, st);
  if (ce == cudaSuccess &&
      launch_mesh_vertices(dd, width, height, dp, dm, mp, dk, n, rig->fx, rig->fy, rig->cx, rig->cy,
                           rig->baseline, scale_factor, od, ox, st) < 0)
    ce = cudaGetLastError();
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(vertex_disps, od, (size_t)n * 4, cudaMemcpyDeviceToHost, st);
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(vertices_xyz, ox, (size_t)n * 12, cudaMemcpyDeviceToHost, st);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
  cudaFree(d);
  if (ce != cudaSuccess) return fail(e, PM_ERR_CUDA, "pm_mesh_vertices: %s", cudaGetErrorString(ce));
  e->launches += 1;
  return PM_OK;
}

int pm_foreground_texture_mask_host(pm_engine* e, const uint8_t* gray, int width, int height,
                                    size_t stride_bytes, int ksize, double min_grad, int downsize,
                                    uint8_t* mask, size_t mask_stride_bytes) {
  if (!e) return PM_ERR_INVALID_ARG;
  if (!gray || !mask || width < 2 || height < 2 || stride_bytes < (size_t)width ||
      mask_stride_bytes < (size_t)width)
    return fail(e, PM_ERR_INVALID_ARG, "ForegroundTextureMask: null pointer or bad size/stride");
  if (downsize < 1 || downsize > 8)   // CHECK(downsize >= 1 && downsize <= 8), patchmatch.cpp:26
    return fail(e, PM_ERR_INVALID_ARG, "Use a downsize argument (int) between 1 and 8");
  const int sk = ksize / downsize;
  if (sk <= 1) return fail(e, PM_ERR_INVALID_ARG, "ksize too small for downsize");   // :28
  if (sk > 15) return fail(e, PM_ERR_UNSUPPORTED, "ksize / downsize %d exceeds the kernel's tile (15)", sk);
  if (downsize > 2 || (downsize == 2 && ((width | height) & 1)))
    return fail(e, PM_ERR_UNSUPPORTED, "downsize %d on %dx%d: cv::resize is restated for exact halving "
                "only (downsize 1, or 2 with even sizes; 2 is the reference's default)", downsize, width, height);
  PM_CUDA(e, cudaSetDevice(e->device));
  const int sw = width / downsize, sh = height / downsize;
  const size_t p0 = (size_t)round_up(width, 16), p1 = (size_t)round_up(sw, 16);
  uint8_t* d = nullptr;
  PM_CUDA(e, cudaMalloc(&d, 2 * p0 * height + 2 * p1 * sh));
  uint8_t *dg = d, *dm = d + p0 * height, *ds = dm + p0 * height, *dsm = ds + p1 * sh;
  cudaStream_t st = e->stream;
  int rc = PM_OK, n = 0;
  auto L = [&](int k) { if (k < 0 && rc == PM_OK) rc = PM_ERR_CUDA; else n += k; };
  cudaError_t ce = cudaMemcpy2DAsync(dg, p0, gray, stride_bytes, width, height, cudaMemcpyHostToDevice, st);
  if (ce == cudaSuccess) {
    if (downsize == 1) {
      L(launch_morph_gradient_mask(dg, width, height, p0, sk, (float)min_grad, dm, p0, st));
    } else {
      L(launch_downscale2(dg, width, height, p0, 0, ds, p1, 0, 1, st));
      L(launch_morph_gradient_mask(ds, sw, sh, p1, sk, (float)min_grad, dsm, p1, st));
      L(launch_resize_up2_u8(dsm, sw, sh, p1, dm, p0, st));
    }
    ce = cudaMemcpy2DAsync(mask, mask_stride_bytes, dm, p0, width, height, cudaMemcpyDeviceToHost, st);
  }
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
  cudaFree(d);
  if (ce != cudaSuccess || rc != PM_OK)
    return fail(e, PM_ERR_CUDA, "ForegroundTextureMask: %s", cudaGetErrorString(ce != cudaSuccess ? ce : cudaGetLastError()));
  e->launches += n;
  return PM_OK;
}

int pm_measure_fp32_peak(pm_engine* e, double* tflops) {
  if (!e || !tflops) return PM_ERR_INVALID_ARG;
  PM_CUDA(e, cudaSetDevice(e->device));
  cudaDeviceProp prop;
  PM_CUDA(e, cudaGetDeviceProperties(&prop, e->device));
  float* scratch = nullptr;
  PM_CUDA(e, cudaMalloc(&scratch, 256));
  cudaEvent_t a = nullptr, b = nullptr;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  const int blocks = prop.multiProcessorCount * 2, iters = 1 << 15;
  double best = 0.0;
  int rc = PM_OK;
  for (int rep = 0; rep < 5 && rc == PM_OK; ++rep) {
    cudaEventRecord(a, e->stream);
    if (launch_fma_peak(scratch, blocks, iters, e->stream) < 0) rc = PM_ERR_CUDA;
    cudaEventRecord(b, e->stream);
    if (cudaEventSynchronize(b) != cudaSuccess) rc = PM_ERR_CUDA;
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, a, b);
    if (rep > 0 && ms > 0.0f)   // the first launch warms up
      best = std::max(best, 2.0 * 16.0 * iters * 1024.0 * blocks / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  cudaFree(scratch);
  if (rc != PM_OK) return fail(e, rc, "fp32 peak probe: %s", cudaGetErrorString(cudaGetLastError()));
  e->launches += 5;
  *tflops = best;
  return PM_OK;
}

int pm_synchronize(pm_engine* e, void* stream) {
  if (!e) return PM_ERR_INVALID_ARG;
  PM_CUDA(e, cudaSetDevice(e->device));
  PM_CUDA(e, cudaStreamSynchronize(stream ? (cudaStream_t)stream : e->stream));
  if (e->p2p.region) {    // a neighbour band never published its rows (pm_band_p2p_*)
    int err = 0;
    PM_CUDA(e, cudaMemcpy(&err, p2p_err(e->p2p.region, e->p2p.buf_bytes), sizeof(int), cudaMemcpyDeviceToHost));
    if (err) {
      PM_CUDA(e, cudaMemset(p2p_err(e->p2p.region, e->p2p.buf_bytes), 0, sizeof(int)));
      return fail(e, PM_ERR_STATE, "row-band exchange over peer memory timed out waiting for a neighbour");
    }
  }
  if (e->seed.status) {   // deferred status of the device SparseInit (candidate buffer overflow)
    int st = 0;
    PM_CUDA(e, cudaMemcpy(&st, e->seed.status, sizeof(int), cudaMemcpyDeviceToHost));
    if (st & 1) {
      PM_CUDA(e, cudaMemset(e->seed.status, 0, sizeof(int)));
      return fail(e, PM_ERR_UNSUPPORTED, "more corner candidates than the sort buffer holds (plateaus of "
                  "equal responses)");
    }
  }
  return PM_OK;
}

// taper: the blocking call has nothing to overlap its first upload and its last download with, so its
// pass sizes start small and halve towards the end (e.g. 512 pairs: 32, 128, 128, 112, 56, 28, 16, 12):
// 1.1 ms of upload and 1.7 ms of download stay exposed instead of 4.5 + 17 ms. A stream of asynchronous
// calls overlaps those with its neighbours and keeps the uniform, larger passes.
static int match_batch_host_impl(pm_engine* e, int n, const uint8_t* left, const uint8_t* right, int width,
                                 int height, size_t stride_bytes, const float* seed_l,
                                 const float* seed_r, uint32_t first_pair_index, float* disp_l,
                                 float* disp_r, size_t disp_stride_bytes, bool taper);

int pm_match_batch_host_async(pm_engine* e, int n, const uint8_t* left, const uint8_t* right, int width,
                              int height, size_t stride_bytes, const float* seed_l,
                              const float* seed_r, uint32_t first_pair_index, float* disp_l,
                              float* disp_r, size_t disp_stride_bytes) {
  return match_batch_host_impl(e, n, left, right, width, height, stride_bytes, seed_l, seed_r,
                               first_pair_index, disp_l, disp_r, disp_stride_bytes, false);
}

static int match_batch_host_impl(pm_engine* e, int n, const uint8_t* left, const uint8_t* right, int width,
                                 int height, size_t stride_bytes, const float* seed_l,
                                 const float* seed_r, uint32_t first_pair_index, float* disp_l,
                                 float* disp_r, size_t disp_stride_bytes, bool taper) {
  if (int rc = check_io(e, n, left, right, width, height, stride_bytes, seed_l, seed_r, disp_l,
                        disp_r, disp_stride_bytes)) return rc;
  PM_CUDA(e, cudaSetDevice(e->device));
  const bool seeds = e->p.init_mode == PM_INIT_SPARSE && seed_l != nullptr;
  const int nb = auto_batch(e, width, height, n, true);
  // a size change re-allocates the workspace: ensure_workspace drains the device first, which
  // also completes any call still in flight
  if (int rc = ensure_workspace(e, width, height, nb, true, seeds)) return rc;
  const Level& L0 = e->lv[0];
  if (int rc = ws_acquire(e, e->stream)) return rc;
  const size_t iplane = stride_bytes * height, oplane = disp_stride_bytes * height;
  const size_t dpitch = (size_t)L0.npitch * sizeof(float), dplane = dpitch * height;
  const int min_pass = std::max(8, nb / 8);
  static const bool taper_on = [] { const char* v = getenv("PM_HOST_TAPER"); return !(v && v[0] == '0'); }();
  taper = taper && taper_on;   // diagnosis switch
  for (int i = 0, m = 0; i < n; i += m) {
    const int left_n = n - i;
    m = std::min(nb, left_n);
    if (taper && n > nb && nb >= 16) {
      if (i == 0) m = std::min(left_n, std::max(min_pass, nb / 4));
      else m = std::max(std::min(min_pass, left_n), left_n / 2);
      m = std::min(m, nb);
    }
    const int s = (int)(e->host_pass & 1);
    const size_t rows = (size_t)m * height;
    // the slot's previous occupant (two passes ago, possibly of the previous call) must have
    // been consumed by the kernels (inputs) and downloaded (outputs)
    if (e->host_pass >= 2) {
      PM_CUDA(e, cudaStreamWaitEvent(e->s_in, e->ev_done[s], 0));
      PM_CUDA(e, cudaStreamWaitEvent(e->stream, e->ev_out[s], 0));
    }
    ++e->host_pass;
    PM_CUDA(e, cudaMemcpy2DAsync(e->d_in[s][0], L0.pitch8, left + i * iplane, stride_bytes, width,
                                 rows, cudaMemcpyHostToDevice, e->s_in));
    PM_CUDA(e, cudaMemcpy2DAsync(e->d_in[s][1], L0.pitch8, right + i * iplane, stride_bytes, width,
                                 rows, cudaMemcpyHostToDevice, e->s_in));
    if (seeds) {
      PM_CUDA(e, cudaMemcpy2DAsync(e->d_seed[s][0], dpitch, (const char*)seed_l + i * oplane,
                                   disp_stride_bytes, width * sizeof(float), rows,
                                   cudaMemcpyHostToDevice, e->s_in));
      PM_CUDA(e, cudaMemcpy2DAsync(e->d_seed[s][1], dpitch, (const char*)seed_r + i * oplane,
                                   disp_stride_bytes, width * sizeof(float), rows,
                                   cudaMemcpyHostToDevice, e->s_in));
    }
    PM_CUDA(e, cudaEventRecord(e->ev_in[s], e->s_in));
    PM_CUDA(e, cudaStreamWaitEvent(e->stream, e->ev_in[s], 0));
    if (int rc = run_device(e, m, e->d_in[s][0], e->d_in[s][1], L0.pitch8, L0.plane8,
                            seeds ? e->d_seed[s][0] : nullptr, seeds ? e->d_seed[s][1] : nullptr,
                            L0.npitch, (size_t)L0.npitch * height, first_pair_index + (uint32_t)i,
                            e->d_out[s][0], e->d_out[s][1], dpitch, dplane, e->stream))
      return rc;
    PM_CUDA(e, cudaEventRecord(e->ev_done[s], e->stream));
    PM_CUDA(e, cudaStreamWaitEvent(e->s_out, e->ev_done[s], 0));
    PM_CUDA(e, cudaMemcpy2DAsync((char*)disp_l + i * oplane, disp_stride_bytes, e->d_out[s][0],
                                 dpitch, width * sizeof(float), rows, cudaMemcpyDeviceToHost,
                                 e->s_out));
    PM_CUDA(e, cudaMemcpy2DAsync((char*)disp_r + i * oplane, disp_stride_bytes, e->d_out[s][1],
                                 dpitch, width * sizeof(float), rows, cudaMemcpyDeviceToHost,
                                 e->s_out));
    PM_CUDA(e, cudaEventRecord(e->ev_out[s], e->s_out));
  }
  e->host_pending = true;
  e->host_pending_seedcheck = e->host_pending_seedcheck || (e->p.init_mode == PM_INIT_SPARSE && !seeds);
  return ws_release(e, e->stream);
}

int pm_wait(pm_engine* e) {
  if (!e) return PM_ERR_INVALID_ARG;
  PM_CUDA(e, cudaSetDevice(e->device));
  PM_CUDA(e, cudaStreamSynchronize(e->s_out));
  PM_CUDA(e, cudaStreamSynchronize(e->stream));
  const bool check = e->host_pending_seedcheck;
  e->host_pending = e->host_pending_seedcheck = false;
  return check ? seed_status(e) : PM_OK;
}

int pm_match_batch_host(pm_engine* e, int n, const uint8_t* left, const uint8_t* right, int width,
                        int height, size_t stride_bytes, const float* seed_l, const float* seed_r,
                        uint32_t first_pair_index, float* disp_l, float* disp_r,
                        size_t disp_stride_bytes) {
  if (int rc = match_batch_host_impl(e, n, left, right, width, height, stride_bytes, seed_l, seed_r,
                                     first_pair_index, disp_l, disp_r, disp_stride_bytes, true)) {
    if (e) { cudaStreamSynchronize(e->s_out); cudaStreamSynchronize(e->stream); e->host_pending = false; }
    return rc;
  }
  return pm_wait(e);
}

int pm_match_host(pm_engine* e, const uint8_t* left, const uint8_t* right, int width, int height,
                  size_t stride_bytes, const float* seed_l, const float* seed_r,
                  uint32_t pair_index, float* disp_l, float* disp_r, size_t disp_stride_bytes) {
  return pm_match_batch_host(e, 1, left, right, width, height, stride_bytes, seed_l, seed_r,
                             pair_index, disp_l, disp_r, disp_stride_bytes);
}

// ------------------------------------------------------------- row-band mode

namespace {

constexpr int kBandExtra = 4;  // halo = overlap + kBandExtra rows of images on each side

struct BandRows { int own_lo, own_hi, load_lo, load_hi, k_lo, nk; };

// world must divide sweep_chunks; the chunk length must fit the lock-step schedule.
int band_rows(const pm_params* p, int frame_h, int rank, int world, BandRows* b, std::string* why) {
  char m[200];
  if (world < 1 || rank < 0 || rank >= world || frame_h < 8) {
    snprintf(m, sizeof(m), "band request rank %d of %d, frame height %d", rank, world, frame_h);
    *why = m;
    return PM_ERR_INVALID_ARG;
  }
  if (p->sweep_chunks % world) {
    snprintf(m, sizeof(m), "world size %d does not divide sweep_chunks = %d: bands are whole "
             "column-sweep chunks", world, p->sweep_chunks);
    *why = m;
    return PM_ERR_UNSUPPORTED;
  }
  if (p->pyramid_levels != 1) {
    *why = "row-band mode runs a single pyramid level";
    return PM_ERR_UNSUPPORTED;
  }
  if (p->patch_size != 3 || p->random_search_k != 0) {
    *why = "row-band mode runs the reference's stage list: patch_size 3, no random search";
    return PM_ERR_UNSUPPORTED;
  }
  const int cs = frame_h / p->sweep_chunks;
  if (cs < 2 * p->sweep_overlap + 2) {
    snprintf(m, sizeof(m), "frame height %d gives column chunks of %d rows, shorter than "
             "2*overlap+2", frame_h, cs);
    *why = m;
    return PM_ERR_UNSUPPORTED;
  }
  b->nk = p->sweep_chunks / world;
  b->k_lo = rank * b->nk;
  b->own_lo = b->k_lo * cs;
  b->own_hi = rank == world - 1 ? frame_h : (b->k_lo + b->nk) * cs;
  const int halo = p->sweep_overlap + kBandExtra;
  b->load_lo = std::max(b->own_lo - halo, 0);
  b->load_hi = std::min(b->own_hi + halo, frame_h);
  return PM_OK;
}

// Frame rows swapped after a column sweep of direction dir (see include/pm_b200.h):
// after the sweep this band holds final values for rows [f_lo, f_hi); the rest of the rows
// it keeps current, [own_lo-ov-2, own_hi+ov+2), comes from its neighbours.
void band_exchange_rows(const pm_params* p, const BandRows& b, int rank, int world, int frame_h,
                        int dir, int r[8]) {
  const int ov = p->sweep_overlap;
  const int f_lo = dir > 0 ? b.own_lo + ov : b.own_lo - ov + 1;
  const int f_hi = dir > 0 ? b.own_hi + ov : b.own_hi - ov + 1;
  for (int i = 0; i < 8; ++i) r[i] = 0;
  if (rank > 0) {
    r[0] = f_lo; r[1] = b.own_lo + ov + 2;         // send_prev
    r[2] = b.own_lo - ov - 2; r[3] = f_lo;         // recv_prev
  }
  if (rank < world - 1) {
    r[4] = b.own_hi - ov - 2; r[5] = f_hi;         // send_next
    r[6] = f_hi; r[7] = b.own_hi + ov + 2;         // recv_next
  }
  for (int i = 0; i < 8; ++i) r[i] = std::min(std::max(r[i], 0), frame_h);
}

// rows [lo, hi) of both views of dcA <-> a packed buffer
int band_copy_rows(pm_engine* e, float2* buf, int lo, int hi, bool pack) {
  if (hi <= lo) return PM_OK;
  const Level& L = e->lv[0];
  const pm_engine::Band& b = e->band;
  const size_t row_bytes = (size_t)L.pitch * sizeof(float2);
  const size_t seg = (size_t)(hi - lo) * row_bytes;
  float2* plane = e->dcA + (size_t)(lo - b.load_lo) * L.pitch;
  if (pack)
    PM_CUDA(e, cudaMemcpy2DAsync(buf, seg, plane, L.plane * sizeof(float2), seg, 2,
                                 cudaMemcpyDeviceToDevice, b.st));
  else
    PM_CUDA(e, cudaMemcpy2DAsync(plane, L.plane * sizeof(float2), buf, seg, seg, 2,
                                 cudaMemcpyDeviceToDevice, b.st));
  return PM_OK;
}

}  // namespace

int pm_band_p2p_export(pm_engine* e, int width, void* ipc_handle, void** region) {
  if (!e || width < 8) return PM_ERR_INVALID_ARG;
  PM_CUDA(e, cudaSetDevice(e->device));
  pm_engine::P2P& P = e->p2p;
  p2p_close(e);
  if (P.region && P.width != width) {
    PM_CUDA(e, cudaDeviceSynchronize());
    cudaFree(P.region);
    P.region = nullptr;
  }
  if (!P.region) {
    const size_t pitch = (size_t)round_up(width + 1, 16);
    P.buf_bytes = ((size_t)2 * e->p.sweep_overlap + 3) * pitch * sizeof(float2) * 2;
    PM_CUDA(e, cudaMalloc(&P.region, 4 * P.buf_bytes + kP2PFlagBytes));
    P.width = width;
  }
  PM_CUDA(e, cudaMemset(P.region, 0, 4 * P.buf_bytes + kP2PFlagBytes));
  P.seq = 0;
  if (ipc_handle) {
    cudaIpcMemHandle_t h;
    PM_CUDA(e, cudaIpcGetMemHandle(&h, P.region));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    std::memcpy(ipc_handle, &h, sizeof(h));
  }
  if (region) *region = P.region;
  return PM_OK;
}

int pm_band_p2p_connect(pm_engine* e, const void* handle_prev, const void* handle_next,
                        void* region_prev, void* region_next) {
  if (!e) return PM_ERR_INVALID_ARG;
  pm_engine::P2P& P = e->p2p;
  if (!P.region) return fail(e, PM_ERR_STATE, "pm_band_p2p_connect before pm_band_p2p_export");
  PM_CUDA(e, cudaSetDevice(e->device));
  p2p_close(e);
  auto open = [&](const void* h, void* raw, char** out, bool* ipc) -> int {
    *out = nullptr;
    *ipc = false;
    if (raw) { *out = (char*)raw; return PM_OK; }          // a neighbour in this process
    if (!h) return PM_OK;
    cudaIpcMemHandle_t ih;
    std::memcpy(&ih, h, sizeof(ih));
    void* p = nullptr;
    PM_CUDA(e, cudaIpcOpenMemHandle(&p, ih, cudaIpcMemLazyEnablePeerAccess));
    *out = (char*)p;
    *ipc = true;
    return PM_OK;
  };
  if (int rc = open(handle_prev, region_prev, &P.peer_prev, &P.ipc_prev)) return rc;
  if (int rc = open(handle_next, region_next, &P.peer_next, &P.ipc_next)) return rc;
  P.on = true;
  return PM_OK;
}

int pm_band_p2p_disable(pm_engine* e) {
  if (!e) return PM_ERR_INVALID_ARG;
  PM_CUDA(e, cudaSetDevice(e->device));
  PM_CUDA(e, cudaDeviceSynchronize());
  p2p_close(e);
  return PM_OK;
}

int pm_band_plan(const pm_params* p, int frame_height, int rank, int world, pm_band_layout* out) {
  if (!p || !out) return PM_ERR_INVALID_ARG;
  BandRows b;
  std::string why;
  if (int rc = band_rows(p, frame_height, rank, world, &b, &why)) return fail(nullptr, rc, "%s", why.c_str());
  out->own_lo = b.own_lo; out->own_hi = b.own_hi;
  out->load_lo = b.load_lo; out->load_hi = b.load_hi;
  out->k_lo = b.k_lo; out->nk = b.nk;
  return PM_OK;
}

int pm_band_exchange_rows(const pm_params* p, int frame_height, int rank, int world, int dir,
                          int rows[8]) {
  if (!p || !rows || (dir != 1 && dir != -1)) return PM_ERR_INVALID_ARG;
  BandRows b;
  std::string why;
  if (int rc = band_rows(p, frame_height, rank, world, &b, &why)) return fail(nullptr, rc, "%s", why.c_str());
  band_exchange_rows(p, b, rank, world, frame_height, dir, rows);
  return PM_OK;
}

int pm_band_begin(pm_engine* e, const uint8_t* d_left, const uint8_t* d_right, int width,
                  size_t stride_bytes, int frame_height, int rank, int world,
                  const float* d_seed_l, const float* d_seed_r, size_t seed_stride_bytes,
                  uint32_t pair_index, void* stream) {
  if (!e) return PM_ERR_INVALID_ARG;
  if (!d_left || !d_right || width < 1 || stride_bytes < (size_t)width)
    return fail(e, PM_ERR_INVALID_ARG, "pm_band_begin: null image or bad stride");
  if (e->p.init_mode == PM_INIT_SPARSE && (!d_seed_l || !d_seed_r || seed_stride_bytes % sizeof(float)))
    return fail(e, PM_ERR_INVALID_ARG, "row-band mode needs the seed maps of the band's rows (SparseInit "
                "selects keypoints over the whole frame) or init_mode = random");
  BandRows b;
  std::string why;
  if (int rc = band_rows(&e->p, frame_height, rank, world, &b, &why)) return fail(e, rc, "%s", why.c_str());
  PM_CUDA(e, cudaSetDevice(e->device));
  e->band.running = false;
  if (int rc = ensure_workspace(e, width, b.load_hi - b.load_lo, 1, false, false, frame_height, b.load_lo))
    return rc;
  pm_engine::Band& B = e->band;
  B.rank = rank; B.world = world;
  B.own_lo = b.own_lo; B.own_hi = b.own_hi; B.load_lo = b.load_lo; B.load_hi = b.load_hi;
  B.k_lo = b.k_lo; B.nk = b.nk;
  const Level& L = e->lv[0];
  if (!B.out[0]) {
    B.xrows = (size_t)2 * e->p.sweep_overlap + 3;
    const size_t xb = B.xrows * L.pitch * sizeof(float2) * 2;
    PM_CUDA(e, cudaMalloc(&B.send_prev, xb));
    PM_CUDA(e, cudaMalloc(&B.recv_prev, xb));
    PM_CUDA(e, cudaMalloc(&B.send_next, xb));
    PM_CUDA(e, cudaMalloc(&B.recv_next, xb));
    PM_CUDA(e, cudaMalloc(&B.out[0], (size_t)L.npitch * L.h * sizeof(float)));
    PM_CUDA(e, cudaMalloc(&B.out[1], (size_t)L.npitch * L.h * sizeof(float)));
  }
  B.st = stream ? (cudaStream_t)stream : e->stream;
  if (int rc = ws_acquire(e, B.st)) return rc;
  if (e->p2p.on) {
    if (e->p2p.width != width)
      return fail(e, PM_ERR_STATE, "the peer-memory exchange was set up for width %d", e->p2p.width);
    if ((rank > 0 && !e->p2p.peer_prev) || (rank < world - 1 && !e->p2p.peer_next))
      return fail(e, PM_ERR_STATE, "pm_band_p2p_connect did not receive the region of a neighbour band");
  }
  B.dL = d_left; B.dR = d_right; B.ipitch = stride_bytes;
  B.seedL = d_seed_l; B.seedR = d_seed_r; B.spitch = seed_stride_bytes / sizeof(float);
  B.pair_index = pair_index;
  B.op = 0;
  B.nops = 1 + 5 * e->p.patchmatch_iters;
  B.pending_dir = 0;
  B.running = true;
  return PM_OK;
}

int pm_band_step(pm_engine* e, pm_band_xfer* x) {
  if (!e) return PM_ERR_INVALID_ARG;
  pm_engine::Band& B = e->band;
  if (!B.running) return fail(e, PM_ERR_STATE, "pm_band_step without pm_band_begin");
  PM_CUDA(e, cudaSetDevice(e->device));
  const Level& L = e->lv[0];
  const BandRows br{B.own_lo, B.own_hi, B.load_lo, B.load_hi, B.k_lo, B.nk};
  int r[8];
  if (B.pending_dir) {  // the caller has moved the buffers: unpack what arrived
    band_exchange_rows(&e->p, br, B.rank, B.world, B.frame_h, B.pending_dir, r);
    if (int rc = band_copy_rows(e, B.recv_prev, r[2], r[3], false)) return rc;
    if (int rc = band_copy_rows(e, B.recv_next, r[6], r[7], false)) return rc;
    B.pending_dir = 0;
  }
  while (B.op < B.nops) {
    const int op = B.op++;
    if (op == 0) {
      if (int rc = setup_level(e, 0, 1, B.dL, B.dR, B.ipitch, 0, B.seedL, B.seedR, B.spitch, 0,
                               B.pair_index, B.st)) return rc;
      if (e->p.patchmatch_iters == 0)
        if (int rc = run_noise(e, 0, 2, -1, B.st)) return rc;
      continue;
    }
    const int it = (op - 1) / 5, s = (op - 1) % 5;
    if (s == 0) {
      if (int rc = run_noise(e, 0, 2, it, B.st)) return rc;
      continue;
    }
    const int along_x = (s == 1 || s == 3), dir = s <= 2 ? +1 : -1;
    if (int rc = run_sweep(e, L, 2, 0, along_x, dir, B.st)) return rc;
    if (!along_x && B.world > 1 && e->p2p.on) {
      // push the rows straight into the neighbours' receive buffers over NVLink, publish the
      // exchange's sequence number there, wait for theirs, unpack: all on the band's stream
      pm_engine::P2P& P = e->p2p;
      band_exchange_rows(&e->p, br, B.rank, B.world, B.frame_h, dir, r);
      const unsigned long long seq = ++P.seq;
      const int par = (int)(seq & 1);
      const bool hp = B.rank > 0, hn = B.rank < B.world - 1;
      StageTimer t(e, B.st, ST_XCHG);
      static const bool fused = [] { const char* v = getenv("PM_BAND_FUSED"); return !(v && v[0] == '0'); }();
      if (fused) {   // two launches: {push, push, signal} and {wait, unpack, unpack}
        PM_LAUNCH(e, launch_band_push_signal(
            e->dcA, L.plane, L.pitch, r[0] - B.load_lo, hp ? r[1] - r[0] : 0,
            hp ? P.peer_prev + (size_t)(2 + par) * P.buf_bytes : nullptr, r[4] - B.load_lo,
            hn ? r[5] - r[4] : 0, hn ? P.peer_next + (size_t)par * P.buf_bytes : nullptr,
            hp ? p2p_flag(P.peer_prev, P.buf_bytes, 2 + par) : nullptr,
            hn ? p2p_flag(P.peer_next, P.buf_bytes, par) : nullptr, seq,
            p2p_ticket(P.region, P.buf_bytes), B.st));
        PM_LAUNCH(e, launch_band_wait_unpack(
            hp ? p2p_flag(P.region, P.buf_bytes, par) : nullptr,
            hn ? p2p_flag(P.region, P.buf_bytes, 2 + par) : nullptr, seq, 2000000000ull,
            p2p_err(P.region, P.buf_bytes), e->dcA, L.plane, L.pitch, r[2] - B.load_lo,
            hp ? r[3] - r[2] : 0, P.region + (size_t)par * P.buf_bytes, r[6] - B.load_lo,
            hn ? r[7] - r[6] : 0, P.region + (size_t)(2 + par) * P.buf_bytes, B.st));
        continue;
      }
      if (hp) PM_LAUNCH(e, launch_band_push(e->dcA, L.plane, L.pitch, r[0] - B.load_lo, r[1] - r[0],
                                            P.peer_prev + (size_t)(2 + par) * P.buf_bytes, B.st));
      if (hn) PM_LAUNCH(e, launch_band_push(e->dcA, L.plane, L.pitch, r[4] - B.load_lo, r[5] - r[4],
                                            P.peer_next + (size_t)par * P.buf_bytes, B.st));
      PM_LAUNCH(e, launch_band_signal(hp ? p2p_flag(P.peer_prev, P.buf_bytes, 2 + par) : nullptr,
                                      hn ? p2p_flag(P.peer_next, P.buf_bytes, par) : nullptr, seq, B.st));
      PM_LAUNCH(e, launch_band_wait(hp ? p2p_flag(P.region, P.buf_bytes, par) : nullptr,
                                    hn ? p2p_flag(P.region, P.buf_bytes, 2 + par) : nullptr, seq,
                                    2000000000ull, p2p_err(P.region, P.buf_bytes), B.st));
      if (int rc = band_copy_rows(e, (float2*)(P.region + (size_t)par * P.buf_bytes), r[2], r[3], false)) return rc;
      if (int rc = band_copy_rows(e, (float2*)(P.region + (size_t)(2 + par) * P.buf_bytes), r[6], r[7], false)) return rc;
      continue;
    }
    if (!along_x && B.world > 1) {
      band_exchange_rows(&e->p, br, B.rank, B.world, B.frame_h, dir, r);
      if (int rc = band_copy_rows(e, B.send_prev, r[0], r[1], true)) return rc;
      if (int rc = band_copy_rows(e, B.send_next, r[4], r[5], true)) return rc;
      B.pending_dir = dir;
      if (x) {
        const size_t rb = (size_t)L.pitch * sizeof(float2) * 2;
        x->send_prev = B.send_prev; x->send_prev_bytes = (size_t)(r[1] - r[0]) * rb;
        x->recv_prev = B.recv_prev; x->recv_prev_bytes = (size_t)(r[3] - r[2]) * rb;
        x->send_next = B.send_next; x->send_next_bytes = (size_t)(r[5] - r[4]) * rb;
        x->recv_next = B.recv_next; x->recv_next_bytes = (size_t)(r[7] - r[6]) * rb;
      }
      return 1;
    }
  }
  return 0;
}

int pm_band_finish(pm_engine* e, float* d_disp_l, float* d_disp_r, size_t disp_stride_bytes) {
  if (!e) return PM_ERR_INVALID_ARG;
  pm_engine::Band& B = e->band;
  if (!B.running) return fail(e, PM_ERR_STATE, "pm_band_finish without pm_band_begin");
  if (B.op < B.nops || B.pending_dir)
    return fail(e, PM_ERR_STATE, "pm_band_finish before pm_band_step returned 0");
  const Level& L = e->lv[0];
  if (!d_disp_l || !d_disp_r || disp_stride_bytes < (size_t)L.w * sizeof(float))
    return fail(e, PM_ERR_INVALID_ARG, "pm_band_finish: null output or bad stride");
  PM_CUDA(e, cudaSetDevice(e->device));
  const size_t tp = (size_t)L.npitch * sizeof(float);
  if (int rc = finish_level0(e, 1, B.out[0], B.out[1], tp, tp * L.h, B.st)) return rc;
  const size_t skip = (size_t)(B.own_lo - B.load_lo) * L.npitch;
  PM_CUDA(e, cudaMemcpy2DAsync(d_disp_l, disp_stride_bytes, B.out[0] + skip, tp,
                               L.w * sizeof(float), B.own_hi - B.own_lo, cudaMemcpyDeviceToDevice, B.st));
  PM_CUDA(e, cudaMemcpy2DAsync(d_disp_r, disp_stride_bytes, B.out[1] + skip, tp,
                               L.w * sizeof(float), B.own_hi - B.own_lo, cudaMemcpyDeviceToDevice, B.st));
  B.running = false;
  return ws_release(e, B.st);
}

int pm_match_band_host(pm_engine* e, const uint8_t* left, const uint8_t* right, int width,
                       size_t stride_bytes, int frame_height, int rank, int world,
                       const float* seed_l, const float* seed_r, uint32_t pair_index,
                       float* disp_l, float* disp_r, size_t disp_stride_bytes,
                       pm_band_exchange_fn exchange, void* user) {
  if (!e) return PM_ERR_INVALID_ARG;
  if (!left || !right || !disp_l || !disp_r || (world > 1 && !exchange))
    return fail(e, PM_ERR_INVALID_ARG, "pm_match_band_host: null pointer");
  BandRows b;
  std::string why;
  if (int rc = band_rows(&e->p, frame_height, rank, world, &b, &why)) return fail(e, rc, "%s", why.c_str());
  const bool seeds = e->p.init_mode == PM_INIT_SPARSE;
  if (seeds && (!seed_l || !seed_r))
    return fail(e, PM_ERR_INVALID_ARG, "row-band mode needs seed maps (or init_mode = random)");
  PM_CUDA(e, cudaSetDevice(e->device));
  const int hl = b.load_hi - b.load_lo, ho = b.own_hi - b.own_lo;
  const size_t ip = (size_t)round_up(width, 16), fp = (size_t)round_up(width, 32) * sizeof(float);
  uint8_t* dimg = nullptr;
  float* dflt = nullptr;
  PM_CUDA(e, cudaMalloc(&dimg, 2 * ip * hl));
  cudaError_t st = cudaMalloc(&dflt, fp * ((seeds ? 2 * hl : 0) + 2 * ho));
  if (st != cudaSuccess) { cudaFree(dimg); return fail(e, PM_ERR_OOM, "pm_match_band_host: %s", cudaGetErrorString(st)); }
  float* dseed = dflt;
  float* dout = (float*)((char*)dflt + fp * (seeds ? 2 * hl : 0));
  int rc = PM_OK;
  auto CU = [&](cudaError_t s) { if (s != cudaSuccess && rc == PM_OK) rc = fail(e, PM_ERR_CUDA, "pm_match_band_host: %s", cudaGetErrorString(s)); };
  cudaStream_t s = e->stream;
  CU(cudaMemcpy2DAsync(dimg, ip, left, stride_bytes, width, hl, cudaMemcpyHostToDevice, s));
  CU(cudaMemcpy2DAsync(dimg + ip * hl, ip, right, stride_bytes, width, hl, cudaMemcpyHostToDevice, s));
  if (seeds) {
    CU(cudaMemcpy2DAsync(dseed, fp, seed_l, disp_stride_bytes, width * sizeof(float), hl, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpy2DAsync((char*)dseed + fp * hl, fp, seed_r, disp_stride_bytes, width * sizeof(float), hl, cudaMemcpyHostToDevice, s));
  }
  if (rc == PM_OK)
    rc = pm_band_begin(e, dimg, dimg + ip * hl, width, ip, frame_height, rank, world,
                       seeds ? dseed : nullptr, seeds ? (float*)((char*)dseed + fp * hl) : nullptr, fp,
                       pair_index, s);
  while (rc == PM_OK) {
    pm_band_xfer x;
    const int r = pm_band_step(e, &x);
    if (r < 0) { rc = r; break; }
    if (r == 0) break;
    if (exchange(user, &x, (void*)s) != 0) rc = fail(e, PM_ERR_STATE, "the halo exchange callback failed");
  }
  if (rc == PM_OK) rc = pm_band_finish(e, dout, (float*)((char*)dout + fp * ho), fp);
  if (rc == PM_OK) {
    CU(cudaMemcpy2DAsync(disp_l, disp_stride_bytes, dout, fp, width * sizeof(float), ho, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpy2DAsync(disp_r, disp_stride_bytes, (char*)dout + fp * ho, fp, width * sizeof(float), ho, cudaMemcpyDeviceToHost, s));
  }
  cudaError_t sy = cudaStreamSynchronize(s);
  if (rc == PM_OK) CU(sy);
  e->band.running = false;
  cudaFree(dimg);
  cudaFree(dflt);
  return rc;
}

// ------------------------------------------------------------------ stage API

#define PM_STAGE_GUARD(e, view)                                                            \
  if (!(e)) return PM_ERR_INVALID_ARG;                                                     \
  if (!(e)->stage_loaded) return fail(e, PM_ERR_STATE, "call pm_stage_load_pair first");   \
  if ((view) < 0 || (view) > 1) return fail(e, PM_ERR_INVALID_ARG, "view %d", view);       \
  PM_CUDA(e, cudaSetDevice((e)->device));                                                  \
  if (int _rc = ws_acquire(e, (e)->stream)) return _rc;

int pm_stage_load_pair(pm_engine* e, const uint8_t* left, const uint8_t* right, int width,
                       int height, size_t stride_bytes) {
  if (!e || !left || !right) return PM_ERR_INVALID_ARG;
  PM_CUDA(e, cudaSetDevice(e->device));
  e->stage_loaded = false;
  if (int rc = ensure_workspace(e, width, height, std::max(1, e->nb * (e->w == width && e->h == height)),
                                true, false)) return rc;
  if (int rc = ws_acquire(e, e->stream)) return rc;
  const Level& L0 = e->lv[0];
  PM_CUDA(e, cudaMemcpy2DAsync(e->d_in[0][0], L0.pitch8, left, stride_bytes, width, height,
                               cudaMemcpyHostToDevice, e->stream));
  PM_CUDA(e, cudaMemcpy2DAsync(e->d_in[0][1], L0.pitch8, right, stride_bytes, width, height,
                               cudaMemcpyHostToDevice, e->stream));
  PM_LAUNCH(e, launch_preprocess(e->d_in[0][0], e->d_in[0][1], L0.pitch8, L0.plane8, e->ref,
                                 e->mat, geom(e, L0), 1, e->stream));
  if (L0.row_smem || L0.row_T)
    PM_LAUNCH(e, launch_transpose2(e->ref, L0.w, L0.h, L0.pitch, L0.plane, e->refT, L0.pitchT,
                                   L0.planeT, 2, e->stream));
  if (L0.row_T)
    PM_LAUNCH(e, launch_transpose2(e->mat, L0.w + 1, L0.h, L0.pitch, L0.plane, e->matT, L0.pitchT,
                                   L0.planeT, 2, e->stream));
  if (L0.row_il) PM_LAUNCH(e, launch_interleave16(e->mat, geom(e, L0), 2, e->matI, e->stream));
  PM_CUDA(e, cudaStreamSynchronize(e->stream));
  e->stage_loaded = true;
  return PM_OK;
}

static int download_plane(pm_engine* e, const float* d, int pitch, float* out) {
  const Level& L0 = e->lv[0];
  PM_CUDA(e, cudaMemcpy2DAsync(out, L0.w * sizeof(float), d, pitch * sizeof(float),
                               L0.w * sizeof(float), L0.h, cudaMemcpyDeviceToHost, e->stream));
  PM_CUDA(e, cudaStreamSynchronize(e->stream));
  return PM_OK;
}

int pm_stage_get_planes(pm_engine* e, int view, float* i_ref, float* g_ref, float* i_mat,
                        float* g_mat) {
  PM_STAGE_GUARD(e, view);
  const Level& L0 = e->lv[0];
  const ViewGeom g = geom(e, L0);
  float* outs[4] = {i_ref, g_ref, i_mat, g_mat};
  for (int k = 0; k < 4; ++k) {
    if (!outs[k]) continue;
    const float2* src = (k < 2 ? e->ref : e->mat) + (size_t)view * L0.plane;
    if (k % 2 == 0) PM_LAUNCH(e, launch_extract_disp(src, g, 1, e->dispv, L0.pitch, L0.plane, e->stream));
    else PM_LAUNCH(e, launch_extract_cost(src, g, 1, e->dispv, L0.pitch, L0.plane, e->stream));
    if (int rc = download_plane(e, e->dispv, L0.pitch, outs[k])) return rc;
  }
  return PM_OK;
}

int pm_stage_noise_image(pm_engine* e, int width, int height, float* out) {
  if (!e || !out || width < 1 || height < 1) return PM_ERR_INVALID_ARG;
  PM_CUDA(e, cudaSetDevice(e->device));
  float* d = nullptr;
  PM_CUDA(e, cudaMalloc(&d, (size_t)width * height * sizeof(float)));
  int n = launch_noise_image(d, width, height, width, e->p.seed, 0, e->stream);
  cudaError_t st = n < 0 ? cudaGetLastError() : cudaSuccess;
  if (st == cudaSuccess)
    st = cudaMemcpyAsync(out, d, (size_t)width * height * sizeof(float), cudaMemcpyDeviceToHost, e->stream);
  if (st == cudaSuccess) st = cudaStreamSynchronize(e->stream);
  cudaFree(d);
  if (st != cudaSuccess) return fail(e, PM_ERR_CUDA, "noise image: %s", cudaGetErrorString(st));
  e->launches += 1;
  return PM_OK;
}

static float2* view_dc(pm_engine* e, int view) { return e->dcA + (size_t)view * e->lv[0].plane; }

static int eval_cost(pm_engine* e, int view) {
  const Level& L0 = e->lv[0];
  const size_t vo = (size_t)view * L0.plane;
  PM_LAUNCH(e, launch_noise_cost(e->ref + vo, e->mat + vo, e->dcA + vo, geom(e, L0), 1, L0.noise,
                                 L0.npitch, 0.0f, INFINITY, 0, e->p.cost_alpha, e->stream));
  return PM_OK;
}

int pm_stage_set_disp(pm_engine* e, int view, const float* disp) {
  PM_STAGE_GUARD(e, view);
  if (!disp) return PM_ERR_INVALID_ARG;
  const Level& L0 = e->lv[0];
  PM_CUDA(e, cudaMemcpy2DAsync(e->dispv, L0.pitch * sizeof(float), disp, L0.w * sizeof(float),
                               L0.w * sizeof(float), L0.h, cudaMemcpyHostToDevice, e->stream));
  PM_LAUNCH(e, launch_set_disp(view_dc(e, view), geom(e, L0), 1, e->dispv, L0.pitch, L0.plane, e->stream));
  if (int rc = eval_cost(e, view)) return rc;
  PM_CUDA(e, cudaStreamSynchronize(e->stream));
  return PM_OK;
}

int pm_stage_get_disp(pm_engine* e, int view, float* disp, float* cost) {
  PM_STAGE_GUARD(e, view);
  const Level& L0 = e->lv[0];
  if (disp) {
    PM_LAUNCH(e, launch_extract_disp(view_dc(e, view), geom(e, L0), 1, e->dispv, L0.pitch, L0.plane, e->stream));
    if (int rc = download_plane(e, e->dispv, L0.pitch, disp)) return rc;
  }
  if (cost) {
    PM_LAUNCH(e, launch_extract_cost(view_dc(e, view), geom(e, L0), 1, e->dispv, L0.pitch, L0.plane, e->stream));
    if (int rc = download_plane(e, e->dispv, L0.pitch, cost)) return rc;
  }
  return PM_OK;
}

int pm_stage_add_noise(pm_engine* e, int view, float scale) {
  PM_STAGE_GUARD(e, view);
  const Level& L0 = e->lv[0];
  const size_t vo = (size_t)view * L0.plane;
  const float dmax = e->p.clamp_disp ? (float)e->p.max_disp : INFINITY;
  PM_LAUNCH(e, launch_noise_cost(e->ref + vo, e->mat + vo, e->dcA + vo, geom(e, L0), 1, L0.noise,
                                 L0.npitch, scale, dmax, e->p.noise_accept, e->p.cost_alpha, e->stream));
  PM_CUDA(e, cudaStreamSynchronize(e->stream));
  return PM_OK;
}

int pm_stage_propagate(pm_engine* e, int view, int along_x, int direction) {
  PM_STAGE_GUARD(e, view);
  if (direction != 1 && direction != -1) return fail(e, PM_ERR_INVALID_ARG, "direction %d", direction);
  const Level& L0 = e->lv[0];
  const size_t vo = (size_t)view * L0.plane;
  // both views live in dcA: sweep this view A -> B and copy the result back
  bool in_place = false;
  if (int rc = sweep_views(e, L0, 1, (size_t)view, along_x != 0, direction, e->dcA, e->dcB, e->stream,
                           0.0f, 0.0f, &in_place))
    return rc;
  if (!in_place)
    PM_CUDA(e, cudaMemcpyAsync(e->dcA + vo, e->dcB + vo, L0.plane * sizeof(float2),
                               cudaMemcpyDeviceToDevice, e->stream));
  PM_CUDA(e, cudaStreamSynchronize(e->stream));
  return PM_OK;
}

int pm_stage_mask_background(pm_engine* e, int view) {
  PM_STAGE_GUARD(e, view);
  const Level& L0 = e->lv[0];
  const size_t vo = (size_t)view * L0.plane;
  PM_LAUNCH(e, launch_mask_background(e->ref + vo, e->mat + vo, e->dcA + vo, geom(e, L0), 1,
                                      e->p.cost_alpha, e->p.cost_improve_factor, 1, e->dispv,
                                      L0.pitch, L0.plane, e->stream));
  PM_LAUNCH(e, launch_set_disp(view_dc(e, view), geom(e, L0), 1, e->dispv, L0.pitch, L0.plane, e->stream));
  if (int rc = eval_cost(e, view)) return rc;
  PM_CUDA(e, cudaStreamSynchronize(e->stream));
  return PM_OK;
}

int pm_stage_subpixel(pm_engine* e, int view) {
  PM_STAGE_GUARD(e, view);
  const Level& L0 = e->lv[0];
  const size_t vo = (size_t)view * L0.plane;
  PM_LAUNCH(e, launch_extract_disp(view_dc(e, view), geom(e, L0), 1, e->dispv, L0.pitch, L0.plane, e->stream));
  PM_LAUNCH(e, launch_subpixel(e->ref + vo, e->mat + vo, geom(e, L0), 1, e->p.cost_alpha, e->dispv,
                               L0.pitch, L0.plane, e->stream));
  PM_LAUNCH(e, launch_set_disp(view_dc(e, view), geom(e, L0), 1, e->dispv, L0.pitch, L0.plane, e->stream));
  if (int rc = eval_cost(e, view)) return rc;
  PM_CUDA(e, cudaStreamSynchronize(e->stream));
  return PM_OK;
}

int pm_stage_random_init(pm_engine* e, int view, uint32_t pair_index, uint32_t level, float range) {
  PM_STAGE_GUARD(e, view);
  const Level& L0 = e->lv[0];
  // launch over two views of one pair and keep the requested one
  ViewGeom g = geom(e, L0);
  PM_LAUNCH(e, launch_init_random(e->dcB, g, 2, e->p.seed, pair_index, level, range, e->stream));
  PM_CUDA(e, cudaMemcpyAsync(view_dc(e, view), e->dcB + (size_t)view * L0.plane,
                             L0.plane * sizeof(float2), cudaMemcpyDeviceToDevice, e->stream));
  if (int rc = eval_cost(e, view)) return rc;
  PM_CUDA(e, cudaStreamSynchronize(e->stream));
  return PM_OK;
}

int pm_stage_mask_occlusions(pm_engine* e, float* disp_l, const float* disp_r, int width, int height) {
  if (!e || !disp_l || !disp_r || width < 1 || height < 1) return PM_ERR_INVALID_ARG;
  PM_CUDA(e, cudaSetDevice(e->device));
  const size_t bytes = (size_t)width * height * sizeof(float);
  float* d = nullptr;
  PM_CUDA(e, cudaMalloc(&d, 2 * bytes));
  cudaError_t st = cudaMemcpyAsync(d, disp_l, bytes, cudaMemcpyHostToDevice, e->stream);
  if (st == cudaSuccess) st = cudaMemcpyAsync((char*)d + bytes, disp_r, bytes, cudaMemcpyHostToDevice, e->stream);
  if (st == cudaSuccess && launch_mask_occlusions(d, (float*)((char*)d + bytes), width, height, e->p.lr_mode, e->stream) < 0)
    st = cudaGetLastError();
  if (st == cudaSuccess) st = cudaMemcpyAsync(disp_l, d, bytes, cudaMemcpyDeviceToHost, e->stream);
  if (st == cudaSuccess) st = cudaStreamSynchronize(e->stream);
  cudaFree(d);
  if (st != cudaSuccess) return fail(e, PM_ERR_CUDA, "mask_occlusions: %s", cudaGetErrorString(st));
  e->launches += 1;
  return PM_OK;
}

int pm_stage_downscale2(pm_engine* e, const uint8_t* src, int width, int height, uint8_t* dst) {
  if (!e || !src || !dst || width < 2 || height < 2) return PM_ERR_INVALID_ARG;
  PM_CUDA(e, cudaSetDevice(e->device));
  const int dw = width / 2, dh = height / 2;
  uint8_t* d = nullptr;
  PM_CUDA(e, cudaMalloc(&d, (size_t)width * height + (size_t)dw * dh));
  uint8_t* dd = d + (size_t)width * height;
  cudaError_t st = cudaMemcpyAsync(d, src, (size_t)width * height, cudaMemcpyHostToDevice, e->stream);
  if (st == cudaSuccess && launch_downscale2(d, width, height, width, 0, dd, dw, 0, 1, e->stream) < 0)
    st = cudaGetLastError();
  if (st == cudaSuccess) st = cudaMemcpyAsync(dst, dd, (size_t)dw * dh, cudaMemcpyDeviceToHost, e->stream);
  if (st == cudaSuccess) st = cudaStreamSynchronize(e->stream);
  cudaFree(d);
  if (st != cudaSuccess) return fail(e, PM_ERR_CUDA, "downscale2: %s", cudaGetErrorString(st));
  e->launches += 1;
  return PM_OK;
}

int pm_stage_median(pm_engine* e, const float* src, int width, int height, int ksize, float* dst) {
  if (!e || !src || !dst || width < 1 || height < 1) return PM_ERR_INVALID_ARG;
  if (ksize != 3 && ksize != 5) return fail(e, PM_ERR_INVALID_ARG, "median ksize %d", ksize);
  PM_CUDA(e, cudaSetDevice(e->device));
  const size_t bytes = (size_t)width * height * sizeof(float);
  float* d = nullptr;
  PM_CUDA(e, cudaMalloc(&d, 2 * bytes));
  float* dd = (float*)((char*)d + bytes);
  cudaError_t st = cudaMemcpyAsync(d, src, bytes, cudaMemcpyHostToDevice, e->stream);
  if (st == cudaSuccess && launch_median(d, dd, width, height, width * sizeof(float), bytes, 1, ksize, e->stream) < 0)
    st = cudaGetLastError();
  if (st == cudaSuccess) st = cudaMemcpyAsync(dst, dd, bytes, cudaMemcpyDeviceToHost, e->stream);
  if (st == cudaSuccess) st = cudaStreamSynchronize(e->stream);
  cudaFree(d);
  if (st != cudaSuccess) return fail(e, PM_ERR_CUDA, "median: %s", cudaGetErrorString(st));
  e->launches += 1;
  return PM_OK;
}

// ------------------------------------------- stereo::Patchmatch stage library

static float* cpu_disp(pm_engine* e) { return e->dispv; }  // view-0 plane of dispv

int pm_cpu_set_disp(pm_engine* e, const float* disp) {
  PM_STAGE_GUARD(e, 0);
  if (!disp) return PM_ERR_INVALID_ARG;
  const Level& L0 = e->lv[0];
  PM_CUDA(e, cudaMemcpy2DAsync(cpu_disp(e), L0.pitch * sizeof(float), disp, L0.w * sizeof(float),
                               L0.w * sizeof(float), L0.h, cudaMemcpyHostToDevice, e->stream));
  PM_CUDA(e, cudaStreamSynchronize(e->stream));
  return PM_OK;
}

int pm_cpu_get_disp(pm_engine* e, float* disp) {
  PM_STAGE_GUARD(e, 0);
  if (!disp) return PM_ERR_INVALID_ARG;
  return download_plane(e, cpu_disp(e), e->lv[0].pitch, disp);
}

int pm_cpu_add_noise(pm_engine* e, float amount) {
  PM_STAGE_GUARD(e, 0);
  const Level& L0 = e->lv[0];
  // a fresh cv::RNG(123) on every call (patchmatch.cpp:146); dprev is free scratch here
  PM_LAUNCH(e, launch_rng_uniform(e->dprev, L0.w, L0.h, L0.pitch, 123, -amount, amount, 0, e->stream));
  PM_LAUNCH(e, launch_c_add_noise(cpu_disp(e), e->dprev, L0.w, L0.h, L0.pitch, L0.pitch, e->stream));
  PM_CUDA(e, cudaStreamSynchronize(e->stream));
  return PM_OK;
}

int pm_cpu_propagate(pm_engine* e, int patch_height, int patch_width, int pass) {
  PM_STAGE_GUARD(e, 0);
  if (pass < -1 || pass > 3) return fail(e, PM_ERR_INVALID_ARG, "pass %d", pass);
  const Level& L0 = e->lv[0];
  for (int ps = (pass < 0 ? 0 : pass); ps <= (pass < 0 ? 3 : pass); ++ps) {
    int n = launch_c_propagate_pass(e->ref, e->mat, cpu_disp(e), L0.w, L0.h, L0.pitch, L0.pitch,
                                    patch_height, patch_width, ps, e->stream);
    if (n < 0) return fail(e, PM_ERR_UNSUPPORTED, "patch %dx%d: odd sizes up to 5 are supported",
                           patch_width, patch_height);
    e->launches += n;
  }
  PM_CUDA(e, cudaStreamSynchronize(e->stream));
  return PM_OK;
}

int pm_cpu_remove_background(pm_engine* e, int patch_height, int patch_width, float win_by_factor) {
  PM_STAGE_GUARD(e, 0);
  const Level& L0 = e->lv[0];
  int n = launch_c_remove_background(e->ref, e->mat, cpu_disp(e), L0.w, L0.h, L0.pitch, L0.pitch,
                                     patch_height, patch_width, win_by_factor, e->stream);
  if (n < 0) return fail(e, PM_ERR_UNSUPPORTED, "patch %dx%d: odd sizes up to 5 are supported",
                         patch_width, patch_height);
  e->launches += n;
  PM_CUDA(e, cudaStreamSynchronize(e->stream));
  return PM_OK;
}

int pm_cpu_estimate_disparity(pm_engine* e, const uint8_t* left, const uint8_t* right, int width,
                              int height, size_t stride_bytes, const float* seed, float* disp,
                              size_t disp_stride_bytes) {
  if (!e || !seed || !disp) return PM_ERR_INVALID_ARG;
  if (int rc = pm_stage_load_pair(e, left, right, width, height, stride_bytes)) return rc;
  const Level& L0 = e->lv[0];
  PM_CUDA(e, cudaMemcpy2DAsync(cpu_disp(e), L0.pitch * sizeof(float), seed, disp_stride_bytes,
                               L0.w * sizeof(float), L0.h, cudaMemcpyHostToDevice, e->stream));
  static const float amount[4] = {32.0f, 8.0f, 2.0f, 0.5f};  // patchmatch_test.cpp:173-180
  static const int patch[4] = {5, 5, 3, 3};
  for (int s = 0; s < 4; ++s) {
    if (int rc = pm_cpu_add_noise(e, amount[s])) return rc;
    if (int rc = pm_cpu_propagate(e, patch[s], patch[s], -1)) return rc;
  }
  if (int rc = pm_cpu_remove_background(e, 3, 3, 1.5f)) return rc;  // :183
  PM_CUDA(e, cudaMemcpy2DAsync(disp, disp_stride_bytes, cpu_disp(e), L0.pitch * sizeof(float),
                               L0.w * sizeof(float), L0.h, cudaMemcpyDeviceToHost, e->stream));
  PM_CUDA(e, cudaStreamSynchronize(e->stream));
  return PM_OK;
}

int pm_cpu_cost(pm_engine* e, int n, const int* xs, const int* ys, const float* ds,
                const int* patch, float* out) {
  PM_STAGE_GUARD(e, 0);
  if (n < 1 || !xs || !ys || !ds || !patch || !out) return PM_ERR_INVALID_ARG;
  const Level& L0 = e->lv[0];
  for (int i = 0; i < n; ++i)
    if (patch[i] < 1 || patch[i] > 5 || !(patch[i] & 1) || xs[i] < 0 || xs[i] >= L0.w || ys[i] < 0 ||
        ys[i] >= L0.h)
      return fail(e, PM_ERR_INVALID_ARG, "sample %d out of range", i);
  char* d = nullptr;
  PM_CUDA(e, cudaMalloc(&d, (size_t)n * 20));
  int* dx = (int*)d; int* dy = dx + n; float* dd = (float*)(dy + n); int* dp = (int*)(dd + n);
  float* dout = (float*)(dp + n);
  cudaError_t st = cudaMemcpyAsync(dx, xs, n * 4, cudaMemcpyHostToDevice, e->stream);
  if (st == cudaSuccess) st = cudaMemcpyAsync(dy, ys, n * 4, cudaMemcpyHostToDevice, e->stream);
  if (st == cudaSuccess) st = cudaMemcpyAsync(dd, ds, n * 4, cudaMemcpyHostToDevice, e->stream);
  if (st == cudaSuccess) st = cudaMemcpyAsync(dp, patch, n * 4, cudaMemcpyHostToDevice, e->stream);
  if (st == cudaSuccess && launch_c_cost_list(e->ref, e->mat, L0.w, L0.h, L0.pitch, dx, dy, dd, dp, n,
                                              dout, e->stream) < 0) st = cudaGetLastError();
  if (st == cudaSuccess) st = cudaMemcpyAsync(out, dout, n * 4, cudaMemcpyDeviceToHost, e->stream);
  if (st == cudaSuccess) st = cudaStreamSynchronize(e->stream);
  cudaFree(d);
  if (st != cudaSuccess) return fail(e, PM_ERR_CUDA, "pm_cpu_cost: %s", cudaGetErrorString(st));
  e->launches += 1;
  return PM_OK;
}

// --------------------------------------------------------------- sparse seeding

// Uploads one pair (right == nullptr: the left image twice) into the first input slot.
static int seed_upload(pm_engine* e, const uint8_t* left, const uint8_t* right, int width, int height,
                       size_t stride_bytes) {
  if (!left || width < 1 || height < 1 || stride_bytes < (size_t)width)
    return fail(e, PM_ERR_INVALID_ARG, "null image or bad size/stride");
  PM_CUDA(e, cudaSetDevice(e->device));
  e->stage_loaded = false;
  if (int rc = ensure_workspace(e, width, height, std::max(1, e->nb * (e->w == width && e->h == height)),
                                true, false, 0, 0, false)) return rc;
  if (int rc = ws_acquire(e, e->stream)) return rc;
  const Level& L0 = e->lv[0];
  PM_CUDA(e, cudaMemcpy2DAsync(e->d_in[0][0], L0.pitch8, left, stride_bytes, width, height,
                               cudaMemcpyHostToDevice, e->stream));
  PM_CUDA(e, cudaMemcpy2DAsync(e->d_in[0][1], L0.pitch8, right ? right : left, stride_bytes, width,
                               height, cudaMemcpyHostToDevice, e->stream));
  return PM_OK;
}

static int seed_status(pm_engine* e) {
  int st = 0;
  PM_CUDA(e, cudaMemcpyAsync(&st, e->seed.status, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  PM_CUDA(e, cudaStreamSynchronize(e->stream));
  if (st & 1) {
    PM_CUDA(e, cudaMemsetAsync(e->seed.status, 0, sizeof(int), e->stream));
    return fail(e, PM_ERR_UNSUPPORTED, "more corner candidates than the sort buffer holds (plateaus of "
                "equal responses)");
  }
  return PM_OK;
}

// one seeding problem (left against right) painted into seed_map[0] and downloaded
static int seed_map_host(pm_engine* e, const uint8_t* left, const uint8_t* right, int width, int height,
                         size_t stride_bytes, int radius, int ow, int oh, float div, float* seeds,
                         size_t seeds_stride_bytes) {
  if (!e) return PM_ERR_INVALID_ARG;
  if (!right || !seeds || ow < 1 || oh < 1 || seeds_stride_bytes < (size_t)ow * sizeof(float))
    return fail(e, PM_ERR_INVALID_ARG, "null pointer or bad output size/stride");
  if (int rc = seed_upload(e, left, right, width, height, stride_bytes)) return rc;
  const Level& L0 = e->lv[0];
  if (int rc = ensure_seed_ws(e, 1, true)) return rc;
  if (int rc = seed_keypoints(e, 1, e->d_in[0][0], e->d_in[0][1], L0.pitch8, L0.plane8, true, e->stream))
    return rc;
  PM_LAUNCH(e, launch_seed_paint(1, L0.w, L0.h, radius, ow, oh, div, e->p.fd_max_features_per_frame,
                                 e->seed, e->seed_map[0], e->seed_map[1], L0.npitch,
                                 (size_t)L0.npitch * L0.h, e->stream));
  PM_CUDA(e, cudaMemcpy2DAsync(seeds, seeds_stride_bytes, e->seed_map[0], L0.npitch * sizeof(float),
                               ow * sizeof(float), oh, cudaMemcpyDeviceToHost, e->stream));
  return seed_status(e);
}

int pm_sparse_init_host(pm_engine* e, const uint8_t* left, const uint8_t* right, int width,
                        int height, size_t stride_bytes, int dilate_factor, float* seeds,
                        size_t seeds_stride_bytes) {
  if (!e) return PM_ERR_INVALID_ARG;
  if (dilate_factor < 0 || dilate_factor > 12) return fail(e, PM_ERR_INVALID_ARG, "dilate_factor %d", dilate_factor);
  return seed_map_host(e, left, right, width, height, stride_bytes, (1 << dilate_factor) + 1, width,
                       height, 1.0f, seeds, seeds_stride_bytes);
}

int pm_cpu_initialize(pm_engine* e, const uint8_t* left, const uint8_t* right, int width, int height,
                      size_t stride_bytes, int downsample_factor, float* seeds,
                      size_t seeds_stride_bytes) {
  if (!e) return PM_ERR_INVALID_ARG;
  const int f = downsample_factor;
  if (f < 1 || f > 12 || width / f < 1 || height / f < 1)
    return fail(e, PM_ERR_INVALID_ARG, "downsample_factor %d", f);
  return seed_map_host(e, left, right, width, height, stride_bytes, (1 << (f - 1)) + 1, width / f,
                       height / f, (float)(1 << f), seeds, seeds_stride_bytes);
}

int pm_stage_corner_response(pm_engine* e, const uint8_t* img, int width, int height,
                             size_t stride_bytes, float* out) {
  if (!e) return PM_ERR_INVALID_ARG;
  if (!out) return fail(e, PM_ERR_INVALID_ARG, "null output");
  if (int rc = seed_upload(e, img, nullptr, width, height, stride_bytes)) return rc;
  const Level& L0 = e->lv[0];
  if (int rc = seed_keypoints(e, 1, e->d_in[0][0], e->d_in[0][1], L0.pitch8, L0.plane8, false, e->stream))
    return rc;
  if (int rc = download_plane(e, e->dispv, L0.pitch, out)) return rc;
  return seed_status(e);
}

int pm_stage_detect(pm_engine* e, const uint8_t* img, int width, int height, size_t stride_bytes,
                    int max_out, int* xy, int* n, int* n_candidates) {
  if (!e) return PM_ERR_INVALID_ARG;
  if (!xy || !n || max_out < 0) return fail(e, PM_ERR_INVALID_ARG, "null output");
  if (int rc = seed_upload(e, img, nullptr, width, height, stride_bytes)) return rc;
  const Level& L0 = e->lv[0];
  if (int rc = seed_keypoints(e, 1, e->d_in[0][0], e->d_in[0][1], L0.pitch8, L0.plane8, false, e->stream))
    return rc;
  int nk = 0, nc = 0;
  PM_CUDA(e, cudaMemcpyAsync(&nk, e->seed.nkp, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  PM_CUDA(e, cudaMemcpyAsync(&nc, e->seed.ncand, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  if (int rc = seed_status(e)) return rc;
  nk = std::min(nk, max_out);
  if (nk > 0)
    PM_CUDA(e, cudaMemcpy(xy, e->seed.kps, sizeof(int2) * nk, cudaMemcpyDeviceToHost));
  *n = nk;
  if (n_candidates) *n_candidates = nc;
  return PM_OK;
}

int pm_stage_match_rectified(pm_engine* e, const uint8_t* left, const uint8_t* right, int width,
                             int height, size_t stride_bytes, const int* xy, int n, double* disps) {
  if (!e) return PM_ERR_INVALID_ARG;
  if (!right || !xy || !disps || n < 0) return fail(e, PM_ERR_INVALID_ARG, "null pointer");
  for (int i = 0; i < n; ++i)
    if (xy[2 * i] < 0 || xy[2 * i] >= width || xy[2 * i + 1] < 0 || xy[2 * i + 1] >= height)
      return fail(e, PM_ERR_INVALID_ARG, "keypoint %d outside the image", i);
  if (int rc = seed_upload(e, left, right, width, height, stride_bytes)) return rc;
  const Level& L0 = e->lv[0];
  if (int rc = check_seed_params(e, L0.w, L0.h)) return rc;
  if (int rc = ensure_seed_ws(e, 1, false)) return rc;
  const int maxf = e->p.fd_max_features_per_frame;
  const SeedImages im{e->d_in[0][0], e->d_in[0][1], (size_t)L0.pitch8, L0.plane8, L0.w, L0.h};
  const SeedMatch mp{e->p.sm_templ_cols, e->p.sm_templ_rows, e->p.sm_max_disp, e->p.sm_max_matching_cost};
  std::vector<float> d(maxf);
  for (int i = 0; i < n; i += maxf) {
    const int m = std::min(maxf, n - i);
    PM_CUDA(e, cudaMemcpyAsync(e->seed.kps, xy + 2 * i, sizeof(int2) * m, cudaMemcpyHostToDevice, e->stream));
    PM_CUDA(e, cudaMemcpyAsync(e->seed.nkp, &m, sizeof(int), cudaMemcpyHostToDevice, e->stream));
    PM_LAUNCH(e, launch_seed_match(im, 1, mp, maxf, e->seed, e->stream));
    PM_CUDA(e, cudaMemcpyAsync(d.data(), e->seed.kpd, sizeof(float) * m, cudaMemcpyDeviceToHost, e->stream));
    PM_CUDA(e, cudaStreamSynchronize(e->stream));
    for (int j = 0; j < m; ++j) disps[i + j] = (double)d[j];
  }
  return PM_OK;
}

// ------------------------------------------------------ disparity -> depth / points

static int check_rig(pm_engine* e, const pm_stereo_rig* rig, double scale) {
  if (!rig || !(rig->fx > 0) || !(rig->fy > 0) || !(rig->baseline > 0) || !(scale > 0))
    return fail(e, PM_ERR_INVALID_ARG, "stereo rig needs fx, fy, baseline and scale_factor > 0");
  return PM_OK;
}

int pm_disp_to_depth_device(pm_engine* e, int n, const float* d_disp, int width, int height,
                            size_t disp_stride_bytes, const pm_stereo_rig* rig, double scale_factor,
                            float* d_depth, size_t depth_stride_bytes, float* d_xyz, void* stream) {
  if (!e) return PM_ERR_INVALID_ARG;
  if (n < 1 || !d_disp || width < 1 || height < 1 || disp_stride_bytes < (size_t)width * sizeof(float) ||
      disp_stride_bytes % sizeof(float) || (!d_depth && !d_xyz) ||
      (d_depth && (depth_stride_bytes < (size_t)width * sizeof(float) || depth_stride_bytes % sizeof(float))))
    return fail(e, PM_ERR_INVALID_ARG, "pm_disp_to_depth: null pointer or bad size/stride");
  if (int rc = check_rig(e, rig, scale_factor)) return rc;
  PM_CUDA(e, cudaSetDevice(e->device));
  const size_t dp = disp_stride_bytes / sizeof(float), op = depth_stride_bytes / sizeof(float);
  PM_LAUNCH(e, launch_disp_to_depth(d_disp, width, height, dp, dp * height, n, rig->fx, rig->fy, rig->cx,
                                    rig->cy, rig->baseline, scale_factor, d_depth, op, op * height,
                                    d_xyz, stream ? (cudaStream_t)stream : e->stream));
  return PM_OK;
}

int pm_disp_to_depth_host(pm_engine* e, const float* disp, int width, int height,
                          size_t disp_stride_bytes, const pm_stereo_rig* rig, double scale_factor,
                          float* depth, float* xyz) {
  if (!e) return PM_ERR_INVALID_ARG;
  if (!disp || width < 1 || height < 1 || disp_stride_bytes < (size_t)width * sizeof(float) ||
      (!depth && !xyz))
    return fail(e, PM_ERR_INVALID_ARG, "pm_disp_to_depth: null pointer or bad size/stride");
  if (int rc = check_rig(e, rig, scale_factor)) return rc;
  PM_CUDA(e, cudaSetDevice(e->device));
  const size_t px = (size_t)width * height;
  float* d = nullptr;
  PM_CUDA(e, cudaMalloc(&d, px * sizeof(float) * 5));
  float* dz = d + px;
  float* dxyz = d + 2 * px;
  cudaError_t st = cudaMemcpy2DAsync(d, width * sizeof(float), disp, disp_stride_bytes,
                                     width * sizeof(float), height, cudaMemcpyHostToDevice, e->stream);
  if (st == cudaSuccess &&
      launch_disp_to_depth(d, width, height, width, px, 1, rig->fx, rig->fy, rig->cx, rig->cy,
                           rig->baseline, scale_factor, depth ? dz : nullptr, width, px,
                           xyz ? dxyz : nullptr, e->stream) < 0)
    st = cudaGetLastError();
  if (st == cudaSuccess && depth)
    st = cudaMemcpyAsync(depth, dz, px * sizeof(float), cudaMemcpyDeviceToHost, e->stream);
  if (st == cudaSuccess && xyz)
    st = cudaMemcpyAsync(xyz, dxyz, px * 3 * sizeof(float), cudaMemcpyDeviceToHost, e->stream);
  if (st == cudaSuccess) st = cudaStreamSynchronize(e->stream);
  cudaFree(d);
  if (st != cudaSuccess) return fail(e, PM_ERR_CUDA, "pm_disp_to_depth: %s", cudaGetErrorString(st));
  e->launches += 1;
  return PM_OK;
}

}  // extern "C"
