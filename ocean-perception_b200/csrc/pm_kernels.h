// pm_kernels.h -- host-callable launchers of the sm_100a kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "pm_device.cuh"

namespace pm {

struct SweepParams {
  int chunks, overlap;
  float alpha;
};

// Every launcher enqueues on `st`, returns the number of kernel launches it made
// (for pm_launch_count) or a negative value after recording cudaGetLastError.

// cv::resize(size/2) on u8: (a+b+c+d+2)>>2 (patchmatch_gpu_test.cpp:62-64). n images.
int launch_downscale2(const uint8_t* src, int sw, int sh, size_t spitch, size_t splane,
                      uint8_t* dst, size_t dpitch, size_t dplane, int n, cudaStream_t st);

// upload/convertTo/GradientMagnitude/flip (patchmatch_gpu.cu:346-360) for n pairs:
// writes the {I,G} planes of views 2p (left reference) and 2p+1 (right reference,
// flipped and swapped).
int launch_preprocess(const uint8_t* L, const uint8_t* R, size_t ipitch, size_t iplane,
                      float2* ref, float2* mat, ViewGeom g, int npairs, cudaStream_t st);
// The same plus the derived layouts of the shared-memory row sweeps in one pass: refT (transposed
// reference planes) and matI (slot-interleaved matched planes, `cols` columns per row group).
bool preprocess_fused_supported(const uint8_t* L, const uint8_t* R, size_t ipitch, size_t iplane,
                                ViewGeom g, int cols);
int launch_preprocess_fused(const uint8_t* L, const uint8_t* R, size_t ipitch, size_t iplane,
                            float2* ref, float2* mat, float2* refT, int pitchT, size_t planeT,
                            float2* matI, int cols, size_t planeI, ViewGeom g, int npairs,
                            cudaStream_t st);

// cv::RNG(seed).fill(UNIFORM,-1,1) (patchmatch_gpu.cu:339-344) generated on the device
// by jumping the multiply-with-carry generator ahead.
int launch_noise_image(float* noise, int w, int h, int pitch, uint64_t seed, long first,
                       cudaStream_t st);

// cv::RNG(seed).fill(UNIFORM, lo, hi) for any range (Patchmatch::AddNoise, patchmatch.cpp:146-147)
int launch_rng_uniform(float* out, int w, int h, int pitch, uint64_t seed, float lo, float hi,
                       long first, cudaStream_t st);

// Initial disparity of nviews views, written to dc.x: random (Philox), from seed
// maps (image coordinates, view 1 flipped; level = pyramid level of dc), or the
// previous level upsampled (x2).
int launch_init_random(float2* dc, ViewGeom g, int nviews, uint64_t seed, uint32_t first_pair,
                       uint32_t level, float range, cudaStream_t st);
int launch_init_seeds(float2* dc, ViewGeom g, int npairs, const float* seed_l, const float* seed_r,
                      size_t spitch, size_t splane, int level, cudaStream_t st);
// pstride 2: `prev` is the coarser level's {d, cost} plane itself (its .x is read)
int launch_upsample2(float2* dc, ViewGeom g, int nviews, const float* prev, int pw, int ph,
                     int ppitch, size_t pplane, cudaStream_t st, int pstride = 1);
int launch_extract_disp(const float2* dc, ViewGeom g, int nviews, float* out, int opitch,
                        size_t oplane, cudaStream_t st);
int launch_set_disp(float2* dc, ViewGeom g, int nviews, const float* in, int ipitch,
                    size_t iplane, cudaStream_t st);

// float I and G planes (ipitch in elements) -> one interleaved {I, G} plane (view 0 of `out`)
int launch_interleave_ig(const float* I, const float* G, size_t ipitch, float2* out, ViewGeom g,
                         cudaStream_t st);

// AddForegroundNoise (patchmatch_gpu.cu:298-304) fused with the refresh of the cost
// plane: dc <- {d', cost(d')}. scale == 0 evaluates the cost of the current d only.
int launch_noise_cost(const float2* ref, const float2* mat, float2* dc, ViewGeom g, int nviews,
                      const float* noise, int npitch, float scale, float dmax, int improve,
                      float alpha, cudaStream_t st);

// Random-search refinement (extension): K Philox-keyed candidates per foreground pixel, in place on
// the {d, cost} plane; view v belongs to pair first_pair + v/2.
int launch_random_search(const float2* ref, const float2* mat, float2* dc, ViewGeom g, int nviews,
                         uint64_t seed, uint32_t first_pair, uint32_t level, uint32_t iter_global,
                         int K, float scale, float dmax, float alpha, cudaStream_t st);

// PropagateRow / PropagateCol (patchmatch_gpu.cu:116-230), lock-step schedule, dc_in -> dc_out.
// Column sweeps of a row band run chunks [k_lo, k_lo + nk) of the frame's chunking
// (nk = 0: all chunks).
int launch_sweep(const float2* ref, const float2* mat, const float2* dc_in, float2* dc_out,
                 ViewGeom g, int nviews, int along_x, int dir, SweepParams sp, cudaStream_t st,
                 int k_lo = 0, int nk = 0);

// Block-per-line sweeps (pm_sweep.cu): every output pixel is written, no pre-copy.
// Row sweep: reads the TRANSPOSED reference and {d, cost} planes, matched rows staged in
// shared memory, writes the row-major {d, cost} plane. Column sweep: all planes row-major.
bool sweep_row_supported(int w, int chunks, int ov);
bool sweep_col_supported(int h, int chunks, int ov);
size_t sweep_row_smem_bytes(int w, int chunks);
// noiseT != nullptr: AddForegroundNoise + cost refresh (launch_noise_cost with improve = 0) are
// done inside the sweep; dcT_in is then the plane before the noise. noiseT is the level's
// noise image transposed with the geometry of dcT. Check sweep_row_fuses_noise first.
bool sweep_row_fuses_noise(int w, int chunks, int ov);
// dc_rm != nullptr: the pre-sweep {d, cost} plane is read row-major from dc_rm (no transposed copy
// needed, dcT_in is ignored); check sweep_row_reads_rowmajor first.
bool sweep_row_reads_rowmajor(int w, int chunks, int ov);
int launch_sweep_row(const float2* refT, const float2* mat, const float2* dcT_in, float2* dc_out,
                     ViewGeom g, int pitchT, size_t planeT, int nviews, int dir, SweepParams sp,
                     cudaStream_t st, const float* noiseT = nullptr, float noise_scale = 0.0f,
                     float noise_dmax = 0.0f, const float2* dc_rm = nullptr,
                     const float2* matI = nullptr, size_t planeI = 0);
// matI != nullptr: the block's matched rows are staged from the slot-interleaved copy of the matched
// plane ([view][row group of 16][column][16 rows], launch_interleave16): bank-conflict-free gathers.
bool sweep_row_interleaved(int w, int chunks, int ov);
size_t sweep_row_interleaved_plane(int w, int h);   // float2 elements per view
int sweep_row_interleaved_cols(int w);              // columns per row group of that plane
int launch_interleave16(const float2* mat, ViewGeom g, int nviews, float2* matI, cudaStream_t st);
// float plane [h][pitch] -> [w][pitchT]
int launch_transpose1(const float* src, int w, int h, int pitch, float* dst, int pitchT,
                      cudaStream_t st);
int launch_sweep_col(const float2* ref, const float2* mat, const float2* dc_in, float2* dc_out,
                     ViewGeom g, int nviews, int dir, SweepParams sp, cudaStream_t st);

// Row sweeps of images too wide for the shared-memory kernel: the column kernel on transposed
// planes (refT, matT incl. the pad column as its last row, dcT in -> dcT out).
bool sweep_rowT_supported(int w, int chunks, int ov);
// Column sweep IN PLACE on the {d, cost} plane (third-generation block kernel, 5-tap cost, radius 1).
// sweep_col_inplace_supported(lines, line length, ...): (w, h) for column sweeps, (h, w) for the
// transposed row sweep below.
bool sweep_col_inplace_supported(int nl, int len, int chunks, int ov);
int launch_sweep_col_inplace(const float2* ref, const float2* mat, float2* dc, ViewGeom g, int nviews,
                             int dir, SweepParams sp, cudaStream_t st);
// The same kernel as a ROW sweep in place on transposed planes (wide frames, row bands).
int launch_sweep_rowT_inplace(const float2* refT, const float2* matT, float2* dcT, ViewGeom g, int pitchT,
                              size_t planeT, int nviews, int dir, SweepParams sp, cudaStream_t st);
int launch_sweep_rowT(const float2* refT, const float2* matT, const float2* dcT_in, float2* dcT_out,
                      ViewGeom g, int pitchT, size_t planeT, int nviews, int dir, SweepParams sp,
                      cudaStream_t st);

// float2 plane transposition [h][pitch] -> [w][pitchT], n planes.
int launch_transpose2(const float2* src, int w, int h, int pitch, size_t plane, float2* dst,
                      int pitchT, size_t planeT, int n, cudaStream_t st);

// MaskBackground (patchmatch_gpu.cu:233-270): dc -> plain disparity planes.
int launch_mask_background(const float2* ref, const float2* mat, const float2* dc, ViewGeom g,
                           int nviews, float alpha, float improve, int do_mask, float* out,
                           int opitch, size_t oplane, cudaStream_t st);

// parabola refinement (extension) on plain disparity planes
int launch_subpixel(const float2* ref, const float2* mat, ViewGeom g, int nviews, float alpha,
                    float* disp, int dpitch, size_t dplane, cudaStream_t st);

// cu::flip of the right result + MaskOcclusions (patchmatch_gpu.cu:368-372) + copy-out.
int launch_finalize(const float* dispv, int vpitch, size_t vplane, int w, int h, int npairs,
                    int lr_mode, float* out_l, float* out_r, size_t opitch_bytes,
                    size_t oplane_bytes, cudaStream_t st);

// MaskOcclusions alone on dense maps (stage test).
int launch_mask_occlusions(float* disp_l, const float* disp_r, int w, int h, int lr_mode,
                           cudaStream_t st);

// k x k median (extension), borders copied.
int launch_median(const float* src, float* dst, int w, int h, size_t pitch_bytes,
                  size_t plane_bytes, int n, int k, cudaStream_t st);

// stereo::Patchmatch stage library (pm_cpu_semantics.cu): in-place on a plain f32 disparity plane
int launch_c_propagate_pass(const float2* ref, const float2* mat, float* disp, int w, int h,
                            int pitch, int dpitch, int ph, int pw, int pass, cudaStream_t st);
int launch_c_remove_background(const float2* ref, const float2* mat, float* disp, int w, int h,
                               int pitch, int dpitch, int ph, int pw, float win_by_factor,
                               cudaStream_t st);
int launch_c_add_noise(float* disp, const float* noise, int w, int h, int dpitch, int npitch,
                       cudaStream_t st);
int launch_c_cost_list(const float2* ref, const float2* mat, int w, int h, int pitch, const int* xs,
                       const int* ys, const float* ds, const int* pws, int n, float* out,
                       cudaStream_t st);

// StereoCamera::DispToDepth / PinholeCamera::Backproject over n dense maps (pitches in elements);
// depth and xyz ([n][h][w][3]) may each be null.
int launch_disp_to_depth(const float* disp, int w, int h, size_t dpitch, size_t dplane, int n,
                         double fx, double fy, double cx, double cy, double baseline, double scale,
                         float* depth, size_t opitch, size_t oplane, float* xyz, cudaStream_t st);

// ForegroundTextureMask pieces (stereo_matching/patchmatch.cpp:19-49): morphological gradient with a
// (2k+1)^2 rectangle thresholded at min_grad (255 / 0), and OpenCV's INTER_LINEAR u8 resize x2
int launch_morph_gradient_mask(const uint8_t* src, int w, int h, size_t pitch, int k, float min_grad,
                               uint8_t* dst, size_t dpitch, cudaStream_t st);
int launch_resize_up2_u8(const uint8_t* src, int sw, int sh, size_t spitch, uint8_t* dst, size_t dpitch,
                         cudaStream_t st);

// mesh vertices from a dense disparity map at n keypoints (x, y floats), optional u8 gate mask
int launch_mesh_vertices(const float* disp, int w, int h, size_t dpitch, const uint8_t* mask,
                         size_t mpitch, const float2* kps, int n, double fx, double fy, double cx,
                         double cy, double baseline, double scale, float* out_disp, float* out_xyz,
                         cudaStream_t st);

// row bands over peer memory: push rows of both views into a (remote) buffer, publish / await a
// sequence number (flags live in the receiver's memory)
int launch_band_push_signal(const float2* dc, size_t plane, int pitch, int row_p, int n_p, void* dst_p,
                            int row_n, int n_n, void* dst_n, unsigned long long* flag_p,
                            unsigned long long* flag_n, unsigned long long seq, unsigned* ticket,
                            cudaStream_t st);
int launch_band_wait_unpack(const unsigned long long* flag_p, const unsigned long long* flag_n,
                            unsigned long long seq, unsigned long long timeout_ns, int* err, float2* dc,
                            size_t plane, int pitch, int row_p, int n_p, const void* src_p, int row_n,
                            int n_n, const void* src_n, cudaStream_t st);
int launch_band_push(const float2* dc, size_t plane, int pitch, int row0, int nrows, void* dst,
                     cudaStream_t st);
int launch_band_signal(unsigned long long* flag_a, unsigned long long* flag_b, unsigned long long seq,
                       cudaStream_t st);
int launch_band_wait(const unsigned long long* flag_a, const unsigned long long* flag_b,
                     unsigned long long seq, unsigned long long timeout_ns, int* err, cudaStream_t st);

// dependent-free FFMA probe: blocks x 1024 threads x iters x 16 FMAs
int launch_fma_peak(float* scratch, int blocks, int iters, cudaStream_t st);

// ---- sparse seeding (pm_seed.cu): PatchmatchGpu::SparseInit, patchmatch_gpu.cu:414-442
constexpr int kMaxSeedFeatures = 1024;  // FeatureDetector max_features_per_frame, upper bound

// u8 images of the seeding problems: problem v uses pair v>>1; even v: reference L against R;
// odd v: R against L, both read right-to-left (patchmatch_gpu.cu:362-365).
struct SeedImages {
  const uint8_t* L;
  const uint8_t* R;
  size_t ipitch, iplane;
  int w, h;
};
struct SeedDetect {  // FeatureDetector::Params (feature_detector.hpp:35-42)
  int max_features, min_distance, block_size, use_harris;
  double quality_level, harris_k;
};
struct SeedMatch {   // StereoMatcher::Params (stereo_matcher.hpp:21-24)
  int templ_cols, templ_rows, max_disp;
  double max_matching_cost;
};
struct SeedState {   // per-view outputs, device memory
  int2* kps;         // [nviews][max_features] keypoints in selection order (view coordinates)
  float* kpd;        // [nviews][max_features] keypoint disparity, -1 = no match
  int* nkp;          // [nviews]
  int* ncand;        // [nviews] local-maximum candidates
  unsigned* vmax;    // [nviews] maximum response (ordered-integer form)
  int* status;       // bit 0: candidate buffer overflow
};

// FeatureDetector::Detect with no tracked keypoints (feature_detector.cpp:89-122). resp is a
// scratch plane of floats ([nviews][h][pitch]); keys holds `cap` (a power of two) 64-bit sort
// keys per view, kplane keys apart.
int launch_seed_detect(const SeedImages& im, int nviews, const SeedDetect& sp, float* resp,
                       int pitch, size_t plane, unsigned long long* keys, size_t kplane, int cap,
                       SeedState s, cudaStream_t st);
// StereoMatcher::MatchRectified of every keypoint (stereo_matcher.cpp:22-116)
int launch_seed_match(const SeedImages& im, int nviews, const SeedMatch& mp, int max_features,
                      SeedState s, cudaStream_t st);
// scatter + rectangular dilate of radius `radius` (+ nearest resize to ow x oh and division)
int launch_seed_paint(int nviews, int w, int h, int radius, int ow, int oh, float div,
                      int max_features, SeedState s, float* out_l, float* out_r, size_t opitch,
                      size_t oplane, cudaStream_t st);

}  // namespace pm
