// patchmatch_gpu.h -- C++ host side of the B200 PatchMatch stereo engine.
//
// Same class shape as the reference's bm::pm::PatchmatchGpu
// (src/vehicle/patchmatch_gpu/patchmatch_gpu.h:77-124 in /root/reference): a Params
// struct with nested detector/matcher params that can be built from a YAML file, an
// engine constructed from Params, and Match() overloads returning float32 left and
// right disparity (0 = invalid). Header-only over the C ABI (pm_b200.h); link with
// libpm_b200.so. With OpenCV headers present Image1b/Image1f are the reference's
// cv::Mat_ typedefs (vision_core/cv_types.hpp:8-22); without them a minimal owning
// image type with the same rows/cols/step/data members is used.
//
// Differences a maintainer should know (INTEGRATION.md has the full list):
//   * errors are exceptions (std::runtime_error with the engine's message) instead of
//     glog CHECK aborts and ignored CUDA errors;
//   * Match(iml, imr, disp, dispr) seeds both views with SparseInit like the reference
//     (patchmatch_gpu.cu:335, 362-365), but on the device; SparseInit() is also exposed, and the
//     seeded overload takes maps computed elsewhere;
//   * the device overload takes raw device pointers (images as uint8, not float planes):
//     gradients are computed by the engine.
#pragma once

#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "pm_b200.h"

#if defined(PM_B200_USE_OPENCV) && __has_include(<opencv2/core.hpp>)
#include <opencv2/core.hpp>
#define PM_B200_HAVE_OPENCV 1
#endif

namespace bm {

#ifdef PM_B200_HAVE_OPENCV
typedef cv::Mat_<uint8_t> Image1b;
typedef cv::Mat_<float> Image1f;
namespace pm_detail {
inline void create(Image1f& m, int rows, int cols) { m.create(rows, cols); }
inline size_t step_bytes(const Image1b& m) { return m.step; }
inline size_t step_bytes(const Image1f& m) { return m.step; }
}  // namespace pm_detail
#else
// Minimal stand-in for cv::Mat_<T>: row-major, owning, `step` in bytes.
template <typename T>
struct Image {
  int rows = 0, cols = 0;
  size_t step = 0;
  T* data = nullptr;
  Image() {}
  Image(int r, int c) { create(r, c); }
  Image(int r, int c, T v) { create(r, c); for (auto& e : store_) e = v; }
  void create(int r, int c) {
    rows = r; cols = c; step = sizeof(T) * (size_t)c;
    store_.assign((size_t)r * c, T());
    data = store_.data();
  }
  bool empty() const { return rows == 0 || cols == 0; }
  T& operator()(int y, int x) { return data[(size_t)y * cols + x]; }
  const T& operator()(int y, int x) const { return data[(size_t)y * cols + x]; }
  Image(const Image& o) { *this = o; }
  Image& operator=(const Image& o) {
    rows = o.rows; cols = o.cols; step = o.step; store_ = o.store_; data = store_.data();
    return *this;
  }
 private:
  std::vector<T> store_;
};
typedef Image<uint8_t> Image1b;
typedef Image<float> Image1f;
namespace pm_detail {
inline void create(Image1f& m, int rows, int cols) { if (m.rows != rows || m.cols != cols) m.create(rows, cols); }
inline size_t step_bytes(const Image1b& m) { return m.step; }
inline size_t step_bytes(const Image1f& m) { return m.step; }
}  // namespace pm_detail
#endif

namespace ft {
// ft::FeatureDetector::Params, feature_tracking/feature_detector.hpp:26-51
struct FeatureDetectorParams {
  int max_features_per_frame = 200;
  int min_distance_btw_tracked_and_detected_features = 20;
  double gftt_quality_level = 0.01;
  int gftt_block_size = 5;
  bool gftt_use_harris_corner_detector = false;
  double gftt_k = 0.04;
};
// ft::StereoMatcher::Params, feature_tracking/stereo_matcher.hpp:18-30
struct StereoMatcherParams {
  int templ_cols = 31;
  int templ_rows = 11;
  int max_disp = 128;
  double max_matching_cost = 0.15;
  bool bidirectional = false;
  bool subpixel_refinement = false;
};
}  // namespace ft

namespace pm {

class PatchmatchGpu final {
 public:
  // PatchmatchGpu::Params, patchmatch_gpu.h:79-92. The extension fields default to the
  // reference's behaviour (pm_b200.h documents each one).
  struct Params final {
    ft::FeatureDetectorParams detector_params;
    ft::StereoMatcherParams matcher_params;
    float cost_alpha = 0.9f;
    int patchmatch_iters = 3;
    int init_dilate_factor = 4;
    float cost_improve_factor = 0.8f;
    // literals of the reference's launch sites
    int patch_size = 3, sweep_chunks = 16, sweep_overlap = 5;
    float noise_scale0 = 32.0f;
    uint64_t seed = 123;
    // extensions
    int init_mode = PM_INIT_SEEDS, max_disp = 128, clamp_disp = 0, pyramid_levels = 1;
    int cost_mode = PM_COST_L1GRAD_X5, lr_mode = PM_LR_RATIO, noise_accept = PM_NOISE_ALWAYS;
    int subpixel = 0, median_ksize = 0, max_batch = 0, random_search_k = 0;

    Params() {}
    // MACRO_PARAMS_STRUCT_CONSTRUCTORS(Params) -> Params(filepath): core/macros.hpp:20-24.
    // `subtree` names the map that holds the FeatureDetector / StereoMatcher sub-trees.
    explicit Params(const std::string& filepath, const std::string& subtree = "") {
      pm_params c;
      char err[512];
      if (pm_params_load_yaml(filepath.c_str(), subtree.c_str(), &c, err, sizeof(err)) != PM_OK)
        throw std::runtime_error(err);
      from_c(c);
    }

    pm_params to_c() const {
      pm_params c;
      pm_params_default(&c);
      c.cost_alpha = cost_alpha; c.patchmatch_iters = patchmatch_iters;
      c.init_dilate_factor = init_dilate_factor; c.cost_improve_factor = cost_improve_factor;
      c.sm_templ_cols = matcher_params.templ_cols; c.sm_templ_rows = matcher_params.templ_rows;
      c.sm_max_disp = matcher_params.max_disp;
      c.sm_max_matching_cost = matcher_params.max_matching_cost;
      c.sm_bidirectional = matcher_params.bidirectional;
      c.sm_subpixel_refinement = matcher_params.subpixel_refinement;
      c.fd_max_features_per_frame = detector_params.max_features_per_frame;
      c.fd_min_distance = detector_params.min_distance_btw_tracked_and_detected_features;
      c.fd_gftt_quality_level = detector_params.gftt_quality_level;
      c.fd_gftt_block_size = detector_params.gftt_block_size;
      c.fd_gftt_use_harris = detector_params.gftt_use_harris_corner_detector;
      c.fd_gftt_k = detector_params.gftt_k;
      c.patch_size = patch_size; c.sweep_chunks = sweep_chunks; c.sweep_overlap = sweep_overlap;
      c.noise_scale0 = noise_scale0; c.seed = seed; c.init_mode = init_mode; c.max_disp = max_disp;
      c.clamp_disp = clamp_disp; c.pyramid_levels = pyramid_levels; c.cost_mode = cost_mode;
      c.lr_mode = lr_mode; c.noise_accept = noise_accept; c.subpixel = subpixel;
      c.median_ksize = median_ksize; c.max_batch = max_batch; c.random_search_k = random_search_k;
      return c;
    }

    void from_c(const pm_params& c) {
      cost_alpha = c.cost_alpha; patchmatch_iters = c.patchmatch_iters;
      init_dilate_factor = c.init_dilate_factor; cost_improve_factor = c.cost_improve_factor;
      matcher_params.templ_cols = c.sm_templ_cols; matcher_params.templ_rows = c.sm_templ_rows;
      matcher_params.max_disp = c.sm_max_disp;
      matcher_params.max_matching_cost = c.sm_max_matching_cost;
      matcher_params.bidirectional = c.sm_bidirectional != 0;
      matcher_params.subpixel_refinement = c.sm_subpixel_refinement != 0;
      detector_params.max_features_per_frame = c.fd_max_features_per_frame;
      detector_params.min_distance_btw_tracked_and_detected_features = c.fd_min_distance;
      detector_params.gftt_quality_level = c.fd_gftt_quality_level;
      detector_params.gftt_block_size = c.fd_gftt_block_size;
      detector_params.gftt_use_harris_corner_detector = c.fd_gftt_use_harris != 0;
      detector_params.gftt_k = c.fd_gftt_k;
      patch_size = c.patch_size; sweep_chunks = c.sweep_chunks; sweep_overlap = c.sweep_overlap;
      noise_scale0 = c.noise_scale0; seed = c.seed; init_mode = c.init_mode; max_disp = c.max_disp;
      clamp_disp = c.clamp_disp; pyramid_levels = c.pyramid_levels; cost_mode = c.cost_mode;
      lr_mode = c.lr_mode; noise_accept = c.noise_accept; subpixel = c.subpixel;
      median_ksize = c.median_ksize; max_batch = c.max_batch; random_search_k = c.random_search_k;
    }
  };

  // MACRO_DELETE_COPY_CONSTRUCTORS(PatchmatchGpu), patchmatch_gpu.h:94
  PatchmatchGpu(const PatchmatchGpu&) = delete;
  void operator=(const PatchmatchGpu&) = delete;

  // PatchmatchGpu(const Params&), patchmatch_gpu.h:96
  explicit PatchmatchGpu(const Params& params, int device = 0) : params_(params) {
    const pm_params c = params.to_c();
    if (pm_create(&c, device, &engine_) != PM_OK) throw std::runtime_error(pm_last_error(nullptr));
  }
  ~PatchmatchGpu() { pm_destroy(engine_); }

  // Match(const Image1b&, const Image1b&, Image1f&, Image1f&), patchmatch_gpu.h:99-102.
  // Outputs are (re)allocated by the callee like the reference's download().
  void Match(const Image1b& iml, const Image1b& imr, Image1f& disp, Image1f& dispr,
             uint32_t pair_index = 0) {
    MatchImpl(iml, imr, nullptr, nullptr, disp, dispr, pair_index);
  }

  // Same, with the outputs of SparseInit (patchmatch_gpu.cu:414-442) supplied by the caller:
  // seed_l in left-image coordinates, seed_r in right-image coordinates.
  void Match(const Image1b& iml, const Image1b& imr, const Image1f& seed_l, const Image1f& seed_r,
             Image1f& disp, Image1f& dispr) {
    if (seed_l.rows != iml.rows || seed_l.cols != iml.cols || seed_r.rows != iml.rows ||
        seed_r.cols != iml.cols)
      throw std::runtime_error("Match: seed maps must have the image size");
    MatchImpl(iml, imr, &seed_l, &seed_r, disp, dispr, 0);
  }

  // Image1f SparseInit(const Image1b& iml, const Image1b& imr, int dilate_factor),
  // patchmatch_gpu.h:110-112, patchmatch_gpu.cu:414-442.
  Image1f SparseInit(const Image1b& iml, const Image1b& imr, int dilate_factor) {
    if (iml.rows != imr.rows || iml.cols != imr.cols ||
        pm_detail::step_bytes(iml) != pm_detail::step_bytes(imr))
      throw std::runtime_error("SparseInit: images must share size and stride");
    Image1f seeds;
    pm_detail::create(seeds, iml.rows, iml.cols);
    Check(pm_sparse_init_host(engine_, (const uint8_t*)iml.data, (const uint8_t*)imr.data, iml.cols,
                              iml.rows, pm_detail::step_bytes(iml), dilate_factor,
                              (float*)seeds.data, pm_detail::step_bytes(seeds)));
    return seeds;
  }

  // Match(const cu::GpuMat& ...), patchmatch_gpu.h:104-108, lifted to n whole pairs held in
  // device memory (uint8 images, float32 outputs, strides in bytes); asynchronous on `stream`.
  void Match(int n, const uint8_t* d_left, const uint8_t* d_right, int width, int height,
             size_t stride_bytes, float* d_disp, float* d_dispr, size_t disp_stride_bytes,
             void* stream = nullptr, const float* d_seed_l = nullptr,
             const float* d_seed_r = nullptr, uint32_t first_pair_index = 0) {
    Check(pm_match_batch_device(engine_, n, d_left, d_right, width, height, stride_bytes, d_seed_l,
                                d_seed_r, first_pair_index, d_disp, d_dispr, disp_stride_bytes,
                                stream));
  }

  // Match(const cu::GpuMat& iml, imr, Gl, Gr, cu::GpuMat& disp), patchmatch_gpu.h:104-108, as the
  // reference has it: ONE view on caller-owned float32 device planes (GpuMat::ptr<float>() and
  // GpuMat::step of Il, Ir, Gl, Gr), `d_disp` seed -> background-masked result in place.
  void Match(const float* d_iml, const float* d_imr, const float* d_Gl, const float* d_Gr, int width,
             int height, size_t plane_step_bytes, float* d_disp, size_t disp_step_bytes,
             void* stream = nullptr) {
    Check(pm_match_planes_device(engine_, d_iml, d_imr, d_Gl, d_Gr, width, height, plane_step_bytes,
                                 d_disp, disp_step_bytes, stream));
  }
  // Waits for an asynchronous Match on `stream` and reports its deferred status.
  void Synchronize(void* stream = nullptr) { Check(pm_synchronize(engine_, stream)); }

  // n pairs in host memory, images back to back.
  void MatchBatch(int n, const uint8_t* left, const uint8_t* right, int width, int height,
                  size_t stride_bytes, float* disp, float* dispr, size_t disp_stride_bytes,
                  const float* seed_l = nullptr, const float* seed_r = nullptr,
                  uint32_t first_pair_index = 0) {
    Check(pm_match_batch_host(engine_, n, left, right, width, height, stride_bytes, seed_l, seed_r,
                              first_pair_index, disp, dispr, disp_stride_bytes));
  }

  // One band of a frame that is split across GPUs in row bands (pm_b200.h, "row bands"):
  // `iml`/`imr` hold rows [layout.load_lo, layout.load_hi) of the frame, the results are rows
  // [layout.own_lo, layout.own_hi). `exchange` moves the halo buffers to/from ranks rank-1 and
  // rank+1 on the given stream, e.g. with grouped ncclSend/ncclRecv. Bit-identical to Match on
  // the whole frame.
  static pm_band_layout BandLayout(const Params& params, int frame_height, int rank, int world) {
    const pm_params c = params.to_c();
    pm_band_layout lay;
    if (pm_band_plan(&c, frame_height, rank, world, &lay) != PM_OK)
      throw std::runtime_error(pm_last_error(nullptr));
    return lay;
  }
  void MatchBand(const Image1b& iml, const Image1b& imr, int frame_height, int rank, int world,
                 pm_band_exchange_fn exchange, void* user, Image1f& disp, Image1f& dispr,
                 uint32_t pair_index = 0) {
    const pm_band_layout lay = BandLayout(params_, frame_height, rank, world);
    if (iml.rows != lay.load_hi - lay.load_lo || imr.rows != iml.rows || imr.cols != iml.cols)
      throw std::runtime_error("MatchBand: the images must hold rows [load_lo, load_hi) of the frame");
    if (pm_detail::step_bytes(iml) != pm_detail::step_bytes(imr))
      throw std::runtime_error("MatchBand: left and right images must share one row stride");
    pm_detail::create(disp, lay.own_hi - lay.own_lo, iml.cols);
    pm_detail::create(dispr, lay.own_hi - lay.own_lo, iml.cols);
    Check(pm_match_band_host(engine_, (const uint8_t*)iml.data, (const uint8_t*)imr.data, iml.cols,
                             pm_detail::step_bytes(iml), frame_height, rank, world, nullptr, nullptr,
                             pair_index, (float*)disp.data, (float*)dispr.data,
                             pm_detail::step_bytes(disp), exchange, user));
  }

  const Params& params() const { return params_; }
  pm_engine* handle() { return engine_; }

 private:
  void Check(int rc) {
    if (rc != PM_OK) throw std::runtime_error(pm_last_error(engine_));
  }
  void MatchImpl(const Image1b& iml, const Image1b& imr, const Image1f* sl, const Image1f* sr,
                 Image1f& disp, Image1f& dispr, uint32_t pair_index) {
    if (iml.rows != imr.rows || iml.cols != imr.cols || iml.rows <= 0)
      throw std::runtime_error("Match: left and right images must have the same, non-zero size");
    pm_detail::create(disp, iml.rows, iml.cols);
    pm_detail::create(dispr, iml.rows, iml.cols);
    if (pm_detail::step_bytes(disp) != pm_detail::step_bytes(dispr) ||
        (sl && (pm_detail::step_bytes(*sl) != pm_detail::step_bytes(disp) ||
                pm_detail::step_bytes(*sr) != pm_detail::step_bytes(disp))))
      throw std::runtime_error("Match: disparity and seed maps must share one row stride");
    if (pm_detail::step_bytes(iml) != pm_detail::step_bytes(imr))
      throw std::runtime_error("Match: left and right images must share one row stride");
    Check(pm_match_host(engine_, (const uint8_t*)iml.data, (const uint8_t*)imr.data, iml.cols,
                        iml.rows, pm_detail::step_bytes(iml),
                        sl ? (const float*)sl->data : nullptr, sr ? (const float*)sr->data : nullptr,
                        pair_index, (float*)disp.data, (float*)dispr.data,
                        pm_detail::step_bytes(disp)));
  }

  Params params_;
  pm_engine* engine_ = nullptr;
};

}  // namespace pm
}  // namespace bm
