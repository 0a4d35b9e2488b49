/*
 * pm_b200.h -- C ABI of the B200-native PatchMatch stereo engine.
 *
 * This is the drop-in boundary for the reference's dense PatchMatch path.
 * The reference has no FFI layer: its boundary is the C++ class
 * bm::pm::PatchmatchGpu in libvehicle_pm_gpu
 * (src/vehicle/patchmatch_gpu/patchmatch_gpu.h:77-124).  The functions below
 * are what a binding of that class needs; include/patchmatch_gpu.h is the C++
 * class with the reference's signatures written on top of them, and
 * ocean-perception_b200/engine.py is the ctypes binding the tests and bench use.
 *
 * Conventions (patchmatch_gpu.cu:331-376):
 *   - images: 8-bit, single channel, rectified, same size, row-major with a
 *     row stride in BYTES;
 *   - disparity: float32 pixels, d = x_left - x_right >= 0, 0 = invalid/background,
 *     row stride in BYTES; the left map is occlusion-masked, the right map is in
 *     right-image coordinates and is not (patchmatch_gpu.cu:368-375);
 *   - every function returns PM_OK (0) or a negative pm_status, never throws or
 *     aborts; pm_last_error() gives the message;
 *   - an engine is bound to one CUDA device and is not thread-safe (one engine
 *     per thread/GPU, like the reference's mutable scratch members).
 *
 * No function here falls back to the CPU: without a usable CUDA device
 * pm_create fails with PM_ERR_CUDA.
 *
 * Citations are relative to /root/reference.
 */
#ifndef PM_B200_H
#define PM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PM_B200_ABI_VERSION 2

typedef enum pm_status {
  PM_OK = 0,
  PM_ERR_INVALID_ARG = -1,   /* null pointer, bad size/stride, bad enum value */
  PM_ERR_UNSUPPORTED = -2,   /* image too small for the sweep schedule, patch size not 3 or 5, ... */
  PM_ERR_CUDA = -3,          /* CUDA runtime/driver error (message has the cudaError) */
  PM_ERR_OOM = -4,           /* device or pinned-host allocation failed */
  PM_ERR_YAML = -5,          /* YAML file missing, malformed, or a required key absent */
  PM_ERR_STATE = -6          /* stage call without pm_stage_load_pair, size mismatch, ... */
} pm_status;

/* PM_INIT_SPARSE: the reference's seeding, PatchmatchGpu::SparseInit (patchmatch_gpu.cu:414-442):
 * run on the device when the caller passes no seed maps, else the caller's maps are used.
 * PM_INIT_SEEDS is the old name of the same value. */
enum { PM_INIT_SPARSE = 0, PM_INIT_SEEDS = 0, PM_INIT_RANDOM = 1 };
/* PM_COST_L1GRAD_X5: L1GradientCost3x3, the five taps the reference evaluates (patchmatch_gpu.cu:72-114).
 * PM_COST_L1GRAD_FULL: L1GradientCost with the full 3x3 patch (patchmatch_gpu.cu:45-69, dead code in
 * the reference library); runs on the one-thread-per-chain kernels, not the tuned block kernels. */
/* PM_COST_CENSUS (extension; the reference has no census cost): census transform of the patch_size^2
 * window on the intensity, Hamming distance, samples taken like GetSubpixel; defined by the oracle. */
enum { PM_COST_L1GRAD_X5 = 0, PM_COST_L1GRAD_FULL = 1, PM_COST_CENSUS = 2 };
enum { PM_LR_RATIO = 0, PM_LR_ABS1PX = 1 };
enum { PM_NOISE_ALWAYS = 0, PM_NOISE_IMPROVE = 1 };

/* PatchmatchGpu::Params (patchmatch_gpu.h:79-92) flattened, plus the keys the
 * reference hard-codes at its launch sites. Defaults (pm_params_default)
 * reproduce the reference. */
typedef struct pm_params {
  /* --- reference fields --- */
  float cost_alpha;           /* 0.9  patchmatch_gpu.h:85 */
  int   patchmatch_iters;     /* 3    patchmatch_gpu.h:86 */
  int   init_dilate_factor;   /* 4    patchmatch_gpu.h:87 (sparse seeding only) */
  float cost_improve_factor;  /* 0.8  patchmatch_gpu.h:88 */
  /* --- StereoMatcher sub-tree (feature_tracking/stereo_matcher.hpp:20-25) --- */
  int    sm_templ_cols;        /* 31 */
  int    sm_templ_rows;        /* 11 */
  int    sm_max_disp;          /* 128 */
  double sm_max_matching_cost; /* 0.15 */
  int    sm_bidirectional;     /* 0 (parsed, unused: stereo_matcher.cpp:17) */
  int    sm_subpixel_refinement; /* 0 */
  /* --- FeatureDetector sub-tree (feature_tracking/feature_detector.hpp:29-47) --- */
  int    fd_max_features_per_frame; /* 200 */
  int    fd_min_distance;           /* 20 */
  double fd_gftt_quality_level;     /* 0.01 */
  int    fd_gftt_block_size;        /* 5 */
  int    fd_gftt_use_harris;        /* 0 */
  double fd_gftt_k;                 /* 0.04 */
  /* --- literals of the reference's launch sites, now parameters --- */
  int   patch_size;           /* 3    patchmatch_gpu.cu:397-408; 3 or 5: the kernels' patch_radius (borders,
                               * fmaxf(x-d, r), x-r clamp) and the window of l1grad_full / census */
  int   sweep_chunks;         /* 16   patchmatch_gpu.cu:385-386 */
  int   sweep_overlap;        /* 5    patchmatch_gpu.cu:143-144 */
  float noise_scale0;         /* 32   patchmatch_gpu.cu:395 (scale = noise_scale0 / 2^iter) */
  uint64_t seed;              /* 123  patchmatch_gpu.cu:341 */
  /* --- extensions; 0 / default = reference behaviour --- */
  int   init_mode;            /* PM_INIT_SPARSE (reference) | PM_INIT_RANDOM */
  int   max_disp;             /* 128: range of the random init; clamp when clamp_disp */
  int   clamp_disp;           /* 0: only the reference's d <= x-1 clamp */
  int   pyramid_levels;       /* 1 */
  int   cost_mode;            /* PM_COST_L1GRAD_X5 | PM_COST_L1GRAD_FULL | PM_COST_CENSUS */
  int   lr_mode;              /* PM_LR_RATIO */
  int   noise_accept;         /* PM_NOISE_ALWAYS */
  int   subpixel;             /* 0 */
  int   median_ksize;         /* 0 | 3 | 5 */
  /* --- execution (no effect on results) --- */
  int   max_batch;            /* pairs processed per device pass (0 = auto) */
  /* --- more extensions (ABI 2) --- */
  int   random_search_k;      /* 0: K random-search candidates per pixel after the sweeps of an iteration:
                               * d + (2u-1) * noise_scale(iter) / 2^(k+1), u Philox-keyed, improve-only */
} pm_params;

typedef struct pm_engine pm_engine;

/* Fills *p with the reference's defaults (patchmatch_gpu.h:85-88 and the
 * nested Params defaults). */
int pm_params_default(pm_params* p);

/* PatchmatchGpu::Params(const std::string& filepath): reads a %YAML:1.0 file
 * with `FeatureDetector:` / `StereoMatcher:` sub-trees (patchmatch_gpu.cu:11-15;
 * shape: config/auv/lcm_nodes/ObjectMesherLcm.yaml:37-59).  The sub-trees'
 * keys are required, like the reference's CHECK in yaml_parser.cpp:82; the
 * extension keys at the top level are optional.  `subtree` may name a nested
 * map that holds the PatchMatch tree ("" or NULL = file root).
 * On failure returns PM_ERR_YAML and writes a message to err (if non-NULL). */
int pm_params_load_yaml(const char* path, const char* subtree, pm_params* p,
                        char* err, size_t err_len);

/* PatchmatchGpu::PatchmatchGpu(const Params&), patchmatch_gpu.cu:322-328. */
int pm_create(const pm_params* params, int device, pm_engine** out);
int pm_destroy(pm_engine* e);
const char* pm_last_error(const pm_engine* e); /* e == NULL: last pm_create error */
int pm_get_params(const pm_engine* e, pm_params* out);
int pm_abi_version(void);

/* PatchmatchGpu::Match(const Image1b&, const Image1b&, Image1f&, Image1f&),
 * patchmatch_gpu.cu:331-376, with HOST buffers. With init_mode = PM_INIT_SPARSE and
 * seed_l == seed_r == NULL the engine runs SparseInit (patchmatch_gpu.cu:414-442) for both
 * views on the device, as the reference's Match does on the host (:335, :362-365).
 * Non-NULL seed_l / seed_r replace it: SparseInit outputs in left- / right-image
 * coordinates (the right one flipped back) with the same stride as the outputs. Both
 * are ignored when init_mode is PM_INIT_RANDOM. pair_index keys the random init. */
int pm_match_host(pm_engine* e, const uint8_t* left, const uint8_t* right,
                  int width, int height, size_t stride_bytes,
                  const float* seed_l, const float* seed_r, uint32_t pair_index,
                  float* disp_l, float* disp_r, size_t disp_stride_bytes);

/* n independent pairs, HOST buffers, images back to back (pair i at
 * base + i*height*stride). Host->device and device->host copies are pipelined
 * with the kernels; returns when all outputs are written. The batch is cut into device
 * passes whose sizes start small and taper towards the end (512 pairs: 32, 128, 128, 112,
 * 56, 28, 16, 12), so that only a short first upload and a short last download are exposed. */
int pm_match_batch_host(pm_engine* e, int n, const uint8_t* left, const uint8_t* right,
                        int width, int height, size_t stride_bytes,
                        const float* seed_l, const float* seed_r, uint32_t first_pair_index,
                        float* disp_l, float* disp_r, size_t disp_stride_bytes);

/* The same without waiting: everything (uploads, kernels, downloads) is enqueued on the engine's
 * streams and the call returns. The input buffers must stay valid and the outputs must not be read
 * until pm_wait() returns. Several calls may be issued back to back (a stream of batches): the
 * upload of a call's first pass then overlaps the kernels and downloads of the previous call, so
 * the copies that a single synchronous call leaves exposed at its head and tail disappear.
 * Use pinned host memory (pm_host_alloc), or the copies are not asynchronous. */
int pm_match_batch_host_async(pm_engine* e, int n, const uint8_t* left, const uint8_t* right,
                              int width, int height, size_t stride_bytes,
                              const float* seed_l, const float* seed_r, uint32_t first_pair_index,
                              float* disp_l, float* disp_r, size_t disp_stride_bytes);
/* Waits for every pm_match_batch_host_async call issued so far and returns their status. */
int pm_wait(pm_engine* e);

/* PatchmatchGpu::Match(const cu::GpuMat& ...), patchmatch_gpu.cu:379-411, lifted
 * to whole pairs: DEVICE pointers, asynchronous on `stream` (a cudaStream_t
 * passed as void*; NULL = the engine's own stream). */
int pm_match_batch_device(pm_engine* e, int n, const uint8_t* d_left, const uint8_t* d_right,
                          int width, int height, size_t stride_bytes,
                          const float* d_seed_l, const float* d_seed_r,
                          uint32_t first_pair_index,
                          float* d_disp_l, float* d_disp_r, size_t disp_stride_bytes,
                          void* stream);

/* PatchmatchGpu::Match(const cu::GpuMat& iml, imr, Gl, Gr, cu::GpuMat& disp), patchmatch_gpu.h:104-108,
 * patchmatch_gpu.cu:379-411: ONE view on caller-owned float32 DEVICE planes (intensity 0..255 and
 * gradient magnitude of the reference and of the matched image, one row stride for the four),
 * `d_disp` holds the seed on entry and the background-masked result on exit (in place, like the
 * reference). patchmatch_iters x {AddForegroundNoise, PropagateRow(+1), PropagateCol(+1),
 * PropagateRow(-1), PropagateCol(-1)}, then MaskBackground; no flip, no occlusion mask (those belong
 * to the host overload). The noise image is the engine's cv::RNG(seed) image of this size.
 * Asynchronous on `stream` (NULL = the engine's own). pyramid_levels must be 1. */
int pm_match_planes_device(pm_engine* e, const float* d_il, const float* d_ir, const float* d_gl,
                           const float* d_gr, int width, int height, size_t plane_stride_bytes,
                           float* d_disp, size_t disp_stride_bytes, void* stream);

/* Waits for everything the engine enqueued on `stream` (NULL = the engine's own stream, the one
 * pm_match_batch_device uses when called with stream == NULL) and returns the deferred status of
 * the asynchronous calls: PM_ERR_UNSUPPORTED when the device SparseInit overflowed its candidate
 * buffer (the seeds of that call are then incomplete), else PM_OK.
 * The engine owns ONE workspace: calls on different streams are ordered against each other with
 * events (a call waits for the previous call's kernels), so they never overlap on the device. */
int pm_synchronize(pm_engine* e, void* stream);

/* ------------------------------------------------------------------------
 * One very large frame split into row bands, one band per GPU.
 *
 * The reference is single-GPU; its column sweeps cut every column into
 * sweep_chunks chunks (patchmatch_gpu.cu:196-202). A band is a whole number of
 * those chunks (world must divide sweep_chunks), so a band's column sweep is the
 * frame's column sweep restricted to its chunks; the rows where neighbouring
 * chunks overlap are swapped with the neighbour band after every column sweep
 * (two exchanges of <= 2*overlap+3 rows of {disparity, cost} per iteration).
 * Row sweeps, noise and the masks are row-local. The result is bit-identical to
 * pm_match_* on the whole frame. pyramid_levels must be 1.
 *
 * The engine never talks to the interconnect: it packs the rows to send into
 * device buffers and unpacks received ones; the caller moves them (ncclSend /
 * ncclRecv on the same stream, or torch.distributed P2P ops).
 * ---------------------------------------------------------------------- */
typedef struct pm_band_layout {
  int own_lo, own_hi;    /* frame rows this rank produces: [own_lo, own_hi) */
  int load_lo, load_hi;  /* frame rows this rank must be given: own rows + halo */
  int k_lo, nk;          /* column-sweep chunks of the frame this rank runs */
} pm_band_layout;

typedef struct pm_band_xfer {
  void*  send_prev; size_t send_prev_bytes;  /* to rank-1   (bytes == 0: nothing) */
  void*  recv_prev; size_t recv_prev_bytes;  /* from rank-1 */
  void*  send_next; size_t send_next_bytes;  /* to rank+1 */
  void*  recv_next; size_t recv_next_bytes;  /* from rank+1 */
} pm_band_xfer;

/* Host-only (no GPU needed). rows[8] = frame-row intervals [lo, hi) swapped after a
 * column sweep of direction dir: send_prev, recv_prev, send_next, recv_next. */
int pm_band_plan(const pm_params* p, int frame_height, int rank, int world, pm_band_layout* out);
int pm_band_exchange_rows(const pm_params* p, int frame_height, int rank, int world, int dir,
                          int rows[8]);

/* DEVICE images holding rows [load_lo, load_hi) of the frame (row 0 of the buffer is
 * frame row load_lo); seeds likewise (NULL with PM_INIT_RANDOM). Everything is
 * enqueued on `stream` (NULL = the engine's own). */
int pm_band_begin(pm_engine* e, const uint8_t* d_left, const uint8_t* d_right, int width,
                  size_t stride_bytes, int frame_height, int rank, int world,
                  const float* d_seed_l, const float* d_seed_r, size_t seed_stride_bytes,
                  uint32_t pair_index, void* stream);
/* Runs the schedule up to the next exchange point. Returns 1 with *x filled: move the
 * buffers (stream-ordered after this call), then call pm_band_step again; 0: the
 * iterations are done; < 0: error. */
int pm_band_step(pm_engine* e, pm_band_xfer* x);
/* MaskBackground, LR check (+ extensions) and copy-out of rows [own_lo, own_hi) to DEVICE
 * maps whose row 0 is frame row own_lo. */
int pm_band_finish(pm_engine* e, float* d_disp_l, float* d_disp_r, size_t disp_stride_bytes);

/* ---- halo exchange over PEER MEMORY instead of the caller's transport --------------------------
 * With this set up, pm_band_step never returns 1: after every column sweep the engine itself writes
 * the rows its neighbours need straight into THEIR receive buffers (peer-mapped device memory: NVLink
 * stores from the pack kernel), publishes the exchange's sequence number in their memory, spins on
 * its own flags until both neighbours have published theirs, and unpacks - four small kernels on
 * the band's stream, no NCCL call and no staging copy. Bit-identical to the NCCL path.
 *
 *   1. every rank: pm_band_p2p_export(e, width, handle, &region)   allocates this rank's receive
 *      region for frames of `width` and returns its 64-byte CUDA IPC handle (and device pointer);
 *   2. the ranks exchange the handles (any host-side channel, e.g. an all-gather);
 *   3. every rank: pm_band_p2p_connect(e, handle_of_rank-1, handle_of_rank+1, NULL, NULL) (NULL where
 *      there is no neighbour). Bands living in ONE process pass the neighbours' `region` pointers
 *      instead (last two arguments) - CUDA IPC handles cannot be opened by the exporting process.
 * All ranks must run the same number of exchanges (they count them alike). A neighbour that never
 * shows up makes the wait give up after 2 s; pm_synchronize then returns PM_ERR_STATE. */
int pm_band_p2p_export(pm_engine* e, int width, void* ipc_handle_64_bytes, void** region);
int pm_band_p2p_connect(pm_engine* e, const void* handle_prev, const void* handle_next,
                        void* region_prev, void* region_next);
int pm_band_p2p_disable(pm_engine* e);

/* The same in one call with HOST buffers (left/right: rows [load_lo, load_hi); disp_*:
 * rows [own_lo, own_hi)); `exchange` is called at every exchange point with the
 * engine's stream and must enqueue the four transfers on it (or complete them). */
typedef int (*pm_band_exchange_fn)(void* user, const pm_band_xfer* x, void* stream);
int pm_match_band_host(pm_engine* e, const uint8_t* left, const uint8_t* right, int width,
                       size_t stride_bytes, int frame_height, int rank, int world,
                       const float* seed_l, const float* seed_r, uint32_t pair_index,
                       float* disp_l, float* disp_r, size_t disp_stride_bytes,
                       pm_band_exchange_fn exchange, void* user);

/* Pinned host memory for pm_match_batch_host callers (cudaHostAlloc). */
int pm_host_alloc(size_t bytes, void** out);
int pm_host_free(void* p);

/* Measures the FP32 FMA rate of the engine's device with a dependent-free FFMA kernel (CUDA events,
 * best of four launches): the ceiling bench.py's ALU roofline is quoted against. */
int pm_measure_fp32_peak(pm_engine* e, double* tflops);

/* Kernel launches issued by this engine since the last reset (bench "gpu_launches"). */
int pm_launch_count(const pm_engine* e, uint64_t* out);
int pm_launch_count_reset(pm_engine* e);

/* Per-stage device time: with profiling on, every stage of every pass is bracketed by
 * CUDA events on the launching stream (no synchronisation is added). pm_last_stage_ms
 * waits for the recorded events and returns, per stage, the milliseconds and the number
 * of spans accumulated since pm_set_profiling(e, 1). Arrays of length PM_N_STAGES;
 * spans may be NULL. */
#define PM_N_STAGES 10
int pm_set_profiling(pm_engine* e, int on);
int pm_last_stage_ms(pm_engine* e, float* ms, uint32_t* spans);
const char* pm_stage_name(int i);

/* ------------------------------------------------------------------------
 * Per-stage entry points for parity tests. HOST pointers, synchronous, dense
 * (stride = width). pm_stage_load_pair builds the level-0 planes of both views
 * (view 0 = left reference; view 1 = right reference, i.e. flipped and swapped
 * planes, patchmatch_gpu.cu:357-367); the other calls act on that state.
 * ---------------------------------------------------------------------- */

/* upload + convertTo(CV_32FC1) + GradientMagnitude + flips, patchmatch_gpu.cu:346-360 */
int pm_stage_load_pair(pm_engine* e, const uint8_t* left, const uint8_t* right,
                       int width, int height, size_t stride_bytes);
int pm_stage_get_planes(pm_engine* e, int view, float* i_ref, float* g_ref,
                        float* i_mat, float* g_mat);
/* the U(-1,1) image of cv::RNG(seed), patchmatch_gpu.cu:339-344 */
int pm_stage_noise_image(pm_engine* e, int width, int height, float* out);
/* sets the current disparity of a view and evaluates its cost plane
 * (L1GradientCost3x3, patchmatch_gpu.cu:72-114) */
int pm_stage_set_disp(pm_engine* e, int view, const float* disp);
int pm_stage_get_disp(pm_engine* e, int view, float* disp, float* cost);
/* AddForegroundNoise, patchmatch_gpu.cu:298-304 (+ cost refresh) */
int pm_stage_add_noise(pm_engine* e, int view, float scale);
/* PropagateRow (along_x = 1) / PropagateCol (along_x = 0), direction +1 / -1,
 * patchmatch_gpu.cu:116-230 */
int pm_stage_propagate(pm_engine* e, int view, int along_x, int direction);
/* MaskBackground, patchmatch_gpu.cu:233-270 */
int pm_stage_mask_background(pm_engine* e, int view);
/* MaskOcclusions, patchmatch_gpu.cu:273-295; disp_l is updated in place */
int pm_stage_mask_occlusions(pm_engine* e, float* disp_l, const float* disp_r,
                             int width, int height);
/* cv::resize(size/2), patchmatch_gpu_test.cpp:62-64 */
int pm_stage_downscale2(pm_engine* e, const uint8_t* src, int width, int height, uint8_t* dst);
/* ---- sparse seeding (the step before the path) ----
 * PatchmatchGpu::SparseInit(iml, imr, dilate_factor), patchmatch_gpu.cu:414-442: GFTT keypoints of
 * `left`, template-matched along the epipolar line of `right`, scattered and dilated with a
 * (2*(2^dilate_factor+1)+1)^2 rectangle. HOST buffers, synchronous. The reference calls it with
 * (flip(imr), flip(iml)) for the right view (:362-365); so can the caller. */
int pm_sparse_init_host(pm_engine* e, const uint8_t* left, const uint8_t* right, int width,
                        int height, size_t stride_bytes, int dilate_factor, float* seeds,
                        size_t seeds_stride_bytes);
/* Patchmatch::Initialize(iml, imr, downsample_factor), patchmatch.cpp:52-87: the same keypoint
 * disparities dilated with radius 2^(f-1)+1, resized to (w/f) x (h/f) with INTER_NEAREST and
 * divided by 2^f. `seeds` has height/f rows of width/f floats. */
int pm_cpu_initialize(pm_engine* e, const uint8_t* left, const uint8_t* right, int width,
                      int height, size_t stride_bytes, int downsample_factor, float* seeds,
                      size_t seeds_stride_bytes);
/* cv::cornerMinEigenVal / cornerHarris as goodFeaturesToTrack evaluates it (block size, Harris
 * switch and k from the params), dense width x height floats. */
int pm_stage_corner_response(pm_engine* e, const uint8_t* img, int width, int height,
                             size_t stride_bytes, float* out);
/* FeatureDetector::Detect(img, {}, new_kp), feature_detector.cpp:89-122: up to max_out keypoints
 * as (x, y) int pairs in selection order; *n_candidates (may be NULL) = local maxima above the
 * quality threshold. */
int pm_stage_detect(pm_engine* e, const uint8_t* img, int width, int height, size_t stride_bytes,
                    int max_out, int* xy, int* n, int* n_candidates);
/* StereoMatcher::MatchRectified(left, right, keypoints), stereo_matcher.cpp:119-131: n keypoints
 * as (x, y) int pairs -> n disparities (-1 = no match). */
int pm_stage_match_rectified(pm_engine* e, const uint8_t* left, const uint8_t* right, int width,
                             int height, size_t stride_bytes, const int* xy, int n, double* disps);

/* ---- the consumer of the path: disparity -> metric depth / points ----
 * StereoCamera::DispToDepth (vision_core/stereo_camera.cpp:49-53: fx * baseline / disp) and
 * PinholeCamera::Backproject (vision_core/pinhole_camera.cpp:41-45: depth * K^-1 * (x, y, 1)) for
 * every pixel of n maps, with ObjectMesher's resolution handling (mesher/object_mesher.cpp:147-150):
 * pixel coordinates and disparity are divided by scale_factor = map height / rig height first.
 * Double arithmetic, float32 results; disparity <= 0 (invalid; the reference CHECK-fails) gives 0.
 * DEVICE pointers, asynchronous on `stream`; d_depth (stride in bytes) and d_xyz ([n][h][w][3]
 * floats, dense) may each be NULL. */
typedef struct pm_stereo_rig {
  double fx, fy, cx, cy;  /* left camera intrinsics (config/shared/ZEDMini.yaml: intrinsics) */
  double baseline;        /* metres */
} pm_stereo_rig;
int pm_disp_to_depth_device(pm_engine* e, int n, const float* d_disp, int width, int height,
                            size_t disp_stride_bytes, const pm_stereo_rig* rig, double scale_factor,
                            float* d_depth, size_t depth_stride_bytes, float* d_xyz, void* stream);
/* the same for one map in HOST memory (depth and xyz dense) */
int pm_disp_to_depth_host(pm_engine* e, const float* disp, int width, int height,
                          size_t disp_stride_bytes, const pm_stereo_rig* rig, double scale_factor,
                          float* depth, float* xyz);

/* ForegroundTextureMask(gray, mask, ksize = 7, min_grad = 35.0, downsize = 2), stereo_matching/
 * patchmatch.hpp:22-26, patchmatch.cpp:19-49 (never called in the reference; the texture gate the dense
 * depth was meant to pass through): cv::morphologyEx(MORPH_GRADIENT) of the down-sized
 * gray image with a (2*(ksize/downsize)+1)^2 rectangle, `> min_grad` -> 255 / 0, cv::resize(INTER_LINEAR)
 * back to the full size (the mask holds the interpolated values at region borders, like the reference's).
 * HOST buffers, synchronous. downsize 1, or 2 with even sizes (exact halving); the arguments the
 * reference CHECK-fails on return PM_ERR_INVALID_ARG. */
int pm_foreground_texture_mask_host(pm_engine* e, const uint8_t* gray, int width, int height,
                                    size_t stride_bytes, int ksize, double min_grad, int downsize,
                                    uint8_t* mask, size_t mask_stride_bytes);

/* Mesher-facing adapter: what ObjectMesher::BuildTriangleMesh (mesher/object_mesher.cpp:
 * 139-150) computes per mesh vertex - Backproject(pixel / scale_factor, DispToDepth(disp / scale_factor)) -
 * with the disparity taken from a DENSE map at n keypoints (x, y float pairs in the map's pixels; the map
 * is read at the rounded pixel) instead of the sparse matcher, optionally gated by a foreground mask
 * (non-zero = foreground, e.g. pm_foreground_texture_mask_host). vertex_disps[i] = 0 and a zero vertex
 * where the map is invalid, masked out or the keypoint lies outside. HOST buffers, synchronous. */
int pm_mesh_vertices_host(pm_engine* e, const float* disp, int width, int height,
                          size_t disp_stride_bytes, const uint8_t* mask, size_t mask_stride_bytes,
                          const float* keypoints_xy, int n, const pm_stereo_rig* rig,
                          double scale_factor, float* vertex_disps, float* vertices_xyz);

/* extensions */
int pm_stage_random_init(pm_engine* e, int view, uint32_t pair_index, uint32_t level, float range);
int pm_stage_subpixel(pm_engine* e, int view);
int pm_stage_median(pm_engine* e, const float* src, int width, int height, int ksize, float* dst);

/* ------------------------------------------------------------------------
 * The reference's CPU stage library, stereo::Patchmatch
 * (src/vehicle/stereo_matching/patchmatch.hpp:29-81), on the GPU: strict raster passes
 * and the cost functor of its only driver (test/stereo_matching/patchmatch_test.cpp:30-45).
 * HOST pointers, synchronous, dense (stride = width) unless a stride is given. The
 * functions act on the left view of the pair loaded with pm_stage_load_pair and on the
 * current disparity set with pm_cpu_set_disp.
 * ---------------------------------------------------------------------- */
int pm_cpu_set_disp(pm_engine* e, const float* disp);
int pm_cpu_get_disp(pm_engine* e, float* disp);
/* Patchmatch::AddNoise(disp, amount, disp > 0), patchmatch.cpp:143-155 */
int pm_cpu_add_noise(pm_engine* e, float amount);
/* Patchmatch::Propagate(..., L1GradientCostFunction, patch_height, patch_width),
 * patchmatch.cpp:248-311; pass = -1 runs all four raster passes, 0..3 one of them */
int pm_cpu_propagate(pm_engine* e, int patch_height, int patch_width, int pass);
/* Patchmatch::RemoveBackground, patchmatch.cpp:314-360 */
int pm_cpu_remove_background(pm_engine* e, int patch_height, int patch_width, float win_by_factor);
/* Patchmatch::EstimateDisparity (declared at patchmatch.hpp:48, never defined): defined as
 * the schedule of the reference's only driver (patchmatch_test.cpp:156-183): noise
 * 32/8/2/0.5 with patches 5,5,3,3, then RemoveBackground(3,3,1.5). `seed` is the output of
 * Patchmatch::Initialize (patchmatch.cpp:52-87). */
int pm_cpu_estimate_disparity(pm_engine* e, const uint8_t* left, const uint8_t* right, int width,
                              int height, size_t stride_bytes, const float* seed, float* disp,
                              size_t disp_stride_bytes);
/* the functor on n (x, y, d, patch) samples (known-answer tests) */
int pm_cpu_cost(pm_engine* e, int n, const int* xs, const int* ys, const float* ds,
                const int* patch, float* out);

#ifdef __cplusplus
}
#endif
#endif /* PM_B200_H */
