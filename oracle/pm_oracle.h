/*
 * pm_oracle.h -- CPU restatement of the reference PatchMatch stereo path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under ocean-perception_b200/ may include,
 * link or call this.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker or the
 * timed CPU baseline.
 *
 * Two semantics are restated (all citations relative to /root/reference):
 *
 *  (G) the GPU library  src/vehicle/patchmatch_gpu/patchmatch_gpu.cu
 *      with its racy 16-chunk sweeps made deterministic by the lock-step
 *      schedule a Pascal warp executed them in (DESIGN.md "sweep schedule"),
 *      and nvcc's FMA contraction pinned to the forms nvcc emits for the
 *      reference's expression shapes (profiles/contraction_evidence.txt).
 *
 *  (C) the CPU stage library src/vehicle/stereo_matching/patchmatch.cpp driven
 *      by the cost functor + schedule of test/stereo_matching/patchmatch_test.cpp,
 *      including the OpenCV 3.4 primitives it leans on (cv::RNG, getRectSubPix,
 *      Sobel, resize/2, dilate), restated from their published algorithms and
 *      pinned against cv2 4.13 outputs (tests/golden/, oracle/gen_goldens.py).
 *
 * Parity status: the reference's own tests hold no golden vectors for this
 * path (zero assertions), and the reference cannot be compiled here (OpenCV
 * 3.4.0-CUDA / CUDA 10.2 EXACT).  The oracle is pinned against fixtures made by
 * a literal cv2 transliteration (oracle/t0_literal.py) run on the reference's
 * fsl1/fsr1 fixture; see DESIGN.md section "Oracle".
 */
#ifndef PM_ORACLE_H
#define PM_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------- params */

typedef struct pmo_params {
  /* reference fields, patchmatch_gpu.h:85-88 */
  float cost_alpha;          /* 0.9 */
  int   patchmatch_iters;    /* 3   */
  float cost_improve_factor; /* 0.8 */
  /* sweep schedule, patchmatch_gpu.cu:385-386,143-144 */
  int   sweep_chunks;        /* 16 */
  int   sweep_overlap;       /* 5  */
  /* noise, patchmatch_gpu.cu:395 : scale = noise_scale0 / 2^iter */
  float noise_scale0;        /* 32 */
  int   noise_accept;        /* 0 = always (reference), 1 = only if cost improves */
  uint64_t seed;             /* 123 (cv::RNG seed, patchmatch_gpu.cu:341) */
  /* extensions (defaults reproduce the reference) */
  int   init_mode;           /* 0 = seed maps supplied, 1 = per-pixel random */
  int   max_disp;            /* 128: range of the random init */
  int   clamp_disp;          /* 0 = no upper clamp besides x-1 (reference) */
  int   pyramid_levels;      /* 1 */
  int   lr_mode;             /* 0 = ratio test (reference), 1 = |dl-dr|<=1 */
  int   subpixel;            /* 0 */
  int   median_ksize;        /* 0 (off), 3 or 5 */
  int   cost_mode;           /* 0 = L1GradientCost3x3 (5 taps, reference), 1 = L1GradientCost ph x pw
                              * (full patch, patchmatch_gpu.cu:45-69), 2 = census + Hamming (extension) */
  int   patch_size;          /* 3 (the reference's launch sites) or 5: radius of the borders/clamps and the
                              * window of cost modes 1 and 2 */
  int   random_search_k;     /* 0; K random-search candidates per pixel and iteration (extension) */
} pmo_params;

void pmo_params_default(pmo_params* p);

/* ------------------------------------------------- OpenCV primitives (C) */

/* cv::RNG(seed).fill(UNIFORM, lo, hi) for CV_32F (multiply-with-carry,
 * A = 4164903690; patchmatch_gpu.cu:341-342, patchmatch.cpp:146-147). */
void pmo_rng_uniform_f32(uint64_t seed, float lo, float hi, float* out, size_t n);

/* cv::resize(src, dst, size/2) INTER_LINEAR at an exact factor 2 on u8
 * == (a+b+c+d+2)>>2 over 2x2 blocks (patchmatch_gpu_test.cpp:62-64). */
void pmo_resize_half_u8(const uint8_t* src, int w, int h, uint8_t* dst);

/* sqrt(Sobel_x^2 + Sobel_y^2), 3x3, BORDER_REFLECT_101
 * (patchmatch_gpu.cu:307-319, patchmatch_test.cpp:48-64). Exact integers
 * under the sqrt, correctly rounded sqrtf. */
void pmo_gradient_mag_u8(const uint8_t* im, int w, int h, float* g);

/* cv::getRectSubPix, 8u->8u (16-bit fixed point weights) and 32f->32f. */
void pmo_get_rect_subpix_u8(const uint8_t* src, int w, int h, int pw, int ph,
                            float cx, float cy, uint8_t* dst);
void pmo_get_rect_subpix_f32(const float* src, int w, int h, int pw, int ph,
                             float cx, float cy, float* dst);

/* cv::dilate with a (2r+1)^2 rectangle anchored at its centre
 * (patchmatch_gpu.cu:436-439). */
void pmo_dilate_rect_f32(const float* src, int w, int h, int r, float* dst);

void pmo_flip_h_u8(const uint8_t* src, int w, int h, uint8_t* dst);
void pmo_flip_h_f32(const float* src, int w, int h, float* dst);
void pmo_u8_to_f32(const uint8_t* src, size_t n, float* dst);

/* Philox-4x32-10 counter RNG used by the extensions (random init). Returns
 * U[0,1) with 24 bits: ((x>>8) * 2^-24). */
float pmo_philox_u01(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3);

/* ------------------------------------------- (G) GPU-library semantics */

/* L1GradientCost3x3 (patchmatch_gpu.cu:72-114) with GetSubpixel (:18-42);
 * yr is integral at every call site so only xr is fractional. */
/* Selects what pmo_g_cost5 (and every (G) stage built on it) evaluates: 0 = the 5-tap
 * L1GradientCost3x3 the reference runs, 1 = the full 3x3 L1GradientCost (patchmatch_gpu.cu:45-69).
 * pmo_g_match sets it from pmo_params.cost_mode. Not thread-safe (test infrastructure). */
void pmo_set_cost_mode(int mode);
/* patch_size / 2 = the patch_radius of the reference's kernels (borders, fmaxf(x-d, r), x-r clamp) and
 * the half window of cost modes 1 and 2. pmo_g_match sets it from pmo_params.patch_size. */
void pmo_set_patch_size(int patch_size);
float pmo_g_cost5(const float* Il, const float* Ir, const float* Gl, const float* Gr,
                  int w, int h, int yl, int xl, float xr, float alpha);

/* cost of a whole disparity map at xr = fmaxf(x - d, 1) (the sweeps' sampling, :161-162); border 0 */
void pmo_g_cost_map(const float* Il, const float* Ir, const float* Gl, const float* Gr, int w, int h,
                    const float* disp, float alpha, float* cost);

/* AddForegroundNoise (patchmatch_gpu.cu:298-304). */
void pmo_g_add_noise(float* disp, const float* unit_noise, size_t n, float scale);

/* PropagateRow / PropagateCol (patchmatch_gpu.cu:116-230), lock-step schedule. */
void pmo_g_propagate_row(const float* Il, const float* Ir, const float* Gl, const float* Gr,
                         int w, int h, float* disp, int dir, float alpha,
                         int chunks, int overlap);
void pmo_g_propagate_col(const float* Il, const float* Ir, const float* Gl, const float* Gr,
                         int w, int h, float* disp, int dir, float alpha,
                         int chunks, int overlap);

/* The sweep as independent chains over {d, cost} planes (what pm_sweep.cu computes);
 * tests prove it equal to the lock-step schedule. */
void pmo_g_sweep_chains(const float* Il, const float* Ir, const float* Gl, const float* Gr,
                        int w, int h, const float* d_in, const float* c_in, float* d_out,
                        float* c_out, int along_x, int dir, float alpha, int chunks, int ov);

/* MaskBackground (patchmatch_gpu.cu:233-270). */
void pmo_g_mask_background(const float* Il, const float* Ir, const float* Gl, const float* Gr,
                           int w, int h, float* disp, float alpha, float improve);

/* MaskOcclusions (patchmatch_gpu.cu:273-295). lr_mode 0 = reference. */
void pmo_g_mask_occlusions(float* displ, const float* dispr, int w, int h, int lr_mode);

/* PatchmatchGpu::Match (device overload, patchmatch_gpu.cu:379-411) on one
 * view: planes f32, disp holds the seed on entry and the result on exit.
 * unit_noise is the w*h U(-1,1) image; iter0 shifts the noise schedule
 * (pyramid); do_mask runs MaskBackground at the end. */
void pmo_g_match_view(const pmo_params* p, const float* Il, const float* Ir,
                      const float* Gl, const float* Gr, int w, int h,
                      const float* unit_noise, float level_scale, int iter0,
                      int do_mask, float* disp);

/* The same with the keys of the random search (extension): pair index, view (0 left / 1 right) and
 * pyramid level of this call. */
void pmo_g_match_view_ex(const pmo_params* p, const float* Il, const float* Ir,
                         const float* Gl, const float* Gr, int w, int h,
                         const float* unit_noise, float level_scale, int iter0,
                         int do_mask, uint32_t pair_index, uint32_t view, uint32_t level,
                         float* disp);
void pmo_x_random_search(const pmo_params* p, const float* Il, const float* Ir, const float* Gl,
                         const float* Gr, int w, int h, uint32_t pair_index, uint32_t view,
                         uint32_t level, uint32_t iter_global, float scale, float dmax, float* disp);

/* PatchmatchGpu::Match (host overload, patchmatch_gpu.cu:331-376).
 * seed_l is in left-image coordinates, seed_r in right-image coordinates
 * (the reference's SparseInit(flip(R), flip(L)) flipped back); both may be
 * NULL when init_mode == 1. pair_index keys the random init.
 * Outputs disp_l (occlusion-masked) and disp_r (right-image coordinates). */
int pmo_g_match(const pmo_params* p, const uint8_t* L, const uint8_t* R, int w, int h,
                const float* seed_l, const float* seed_r, uint32_t pair_index,
                float* disp_l, float* disp_r);

/* extension stages shared by the GPU path */
void pmo_x_random_init(const pmo_params* p, int w, int h, uint32_t pair_index,
                       uint32_t view, uint32_t level, float range, float* disp);
void pmo_x_upsample2(const float* src, int sw, int sh, int w, int h, float* dst);
void pmo_x_median(const float* src, int w, int h, int k, float* dst);
void pmo_x_subpixel(const float* Il, const float* Ir, const float* Gl, const float* Gr,
                    int w, int h, float alpha, float* disp);

/* ------------------------------------------- (C) CPU stage-library semantics */

/* test functor L1GradientCostFunction (patchmatch_test.cpp:30-45) on patches
 * fetched as PropagateNeighbors does (patchmatch.cpp:98-111,171-180). */
float pmo_c_cost(const uint8_t* Il, const uint8_t* Ir, const float* Gl, const float* Gr,
                 int w, int h, int x, int y, float d, int pw, int ph);

/* Patchmatch::AddNoise (patchmatch.cpp:143-155) with mask = disp > 0. */
void pmo_c_add_noise(float* disp, int w, int h, float amount);

/* Patchmatch::Propagate (patchmatch.cpp:248-311). */
void pmo_c_propagate(const uint8_t* Il, const uint8_t* Ir, const float* Gl, const float* Gr,
                     int w, int h, float* disp, int ph, int pw);

/* One of the four raster passes of Propagate (pass 0..3), for stage tests. */
void pmo_c_propagate_pass(const uint8_t* Il, const uint8_t* Ir, const float* Gl, const float* Gr,
                          int w, int h, float* disp, int ph, int pw, int pass);

/* Patchmatch::RemoveBackground (patchmatch.cpp:314-360). */
void pmo_c_remove_background(const uint8_t* Il, const uint8_t* Ir, const float* Gl, const float* Gr,
                             int w, int h, float* disp, int ph, int pw, float win_by_factor);

/* Patchmatch::EstimateDisparity -- declared at patchmatch.hpp:48 and never
 * defined; defined here as the only driver the reference has
 * (patchmatch_test.cpp:156-183): gradient, then noise 32/8/2/0.5 with patches
 * 5,5,3,3, then RemoveBackground(3,3,1.5). disp holds the Initialize() seed. */
void pmo_c_estimate_disparity(const uint8_t* Il, const uint8_t* Ir, int w, int h, float* disp);

/* ForegroundTextureMask (stereo_matching/patchmatch.cpp:19-49, declared patchmatch.hpp:22-26, never
 * called in the reference): morphological gradient of the (down-sized) gray image with a
 * (2*(ksize/downsize)+1)^2 rectangle, thresholded at min_grad, resized back with INTER_LINEAR (so the
 * mask holds 0, 255 and the interpolated values in between at region borders). downsize 1 or 2
 * (even image sizes); returns 0, or < 0 for the cases the reference CHECK-fails / not restated. */
int pmo_c_foreground_texture_mask(const uint8_t* gray, int w, int h, int ksize, double min_grad,
                                  int downsize, uint8_t* mask);

/* StereoCamera::DispToDepth + PinholeCamera::Backproject per pixel (vision_core/
 * stereo_camera.cpp:49-53, pinhole_camera.cpp:41-45, mesher/object_mesher.cpp:147-150).
 * depth / xyz ([h][w][3]) may each be NULL. */
void pmo_x_disp_to_depth(const float* disp, int w, int h, double fx, double fy, double cx,
                         double cy, double baseline, double scale, float* depth, float* xyz);

/* ------------------------------------------- (S) sparse seeding semantics
 * (oracle/pm_oracle_seed.c) */

typedef struct pmo_seed_params {
  /* FeatureDetector::Params, feature_tracking/feature_detector.hpp:35-42 */
  int    max_features;       /* 200 */
  int    min_distance;       /* 20  */
  double quality_level;      /* 0.01 */
  int    block_size;         /* 5 */
  int    use_harris;         /* 0 */
  double harris_k;           /* 0.04 */
  /* StereoMatcher::Params, feature_tracking/stereo_matcher.hpp:21-24 */
  int    templ_cols;         /* 31 */
  int    templ_rows;         /* 11 */
  int    max_disp;           /* 128 */
  double max_matching_cost;  /* 0.15 */
} pmo_seed_params;

void pmo_seed_params_default(pmo_seed_params* p);

/* cv::cornerMinEigenVal / cv::cornerHarris (aperture 3, BORDER_REFLECT_101), exact
 * integer sums rounded once. */
void pmo_s_corner_response(const uint8_t* im, int w, int h, int block, int harris, double k,
                           float* out);

/* FeatureDetector::Detect with no tracked keypoints (feature_detector.cpp:89-122) ==
 * cv::goodFeaturesToTrack. kx/ky have room for max_features entries. */
int pmo_s_good_features(const uint8_t* im, int w, int h, const pmo_seed_params* sp, int* kx,
                        int* ky, int* n_candidates);

/* StereoMatcher::MatchRectified, one keypoint (stereo_matcher.cpp:22-116). */
double pmo_s_match_rectified(const uint8_t* L, const uint8_t* R, int w, int h,
                             const pmo_seed_params* sp, int kpx, int kpy);

/* PatchmatchGpu::SparseInit (patchmatch_gpu.cu:414-442). */
void pmo_s_sparse_init(const uint8_t* L, const uint8_t* R, int w, int h,
                       const pmo_seed_params* sp, int dilate_factor, float* seeds);

/* Patchmatch::Initialize (patchmatch.cpp:52-87); seeds is (w/f) x (h/f). */
void pmo_c_initialize(const uint8_t* L, const uint8_t* R, int w, int h,
                      const pmo_seed_params* sp, int downsample_factor, float* seeds);

/* Both seed maps of PatchmatchGpu::Match (patchmatch_gpu.cu:335,357-365); seed_r in
 * right-image coordinates. */
void pmo_s_match_seeds(const uint8_t* L, const uint8_t* R, int w, int h,
                       const pmo_seed_params* sp, int dilate_factor, float* seed_l,
                       float* seed_r);

#ifdef __cplusplus
}
#endif
#endif /* PM_ORACLE_H */
