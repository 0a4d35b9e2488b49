// patchmatch.hpp -- C++ host side of the reference's CPU stage library on the GPU.
//
// Same method set as bm::stereo::Patchmatch (src/vehicle/stereo_matching/patchmatch.hpp:29-81
// in /root/reference): Initialize, AddNoise, Propagate, RemoveBackground, and EstimateDisparity -- which the
// reference declares (patchmatch.hpp:48) but never defines; it is defined as the schedule of
// the reference's only driver (test/stereo_matching/patchmatch_test.cpp:156-183).
// The cost functor is that driver's L1GradientCostFunction (patchmatch_test.cpp:30-45); a
// caller-supplied std::function cannot run on the device, so the functor argument is gone.
#pragma once

#include "patchmatch_gpu.h"

namespace bm {
namespace stereo {

class Patchmatch final {
 public:
  typedef pm::PatchmatchGpu::Params Params;  // patchmatch.hpp:31-40 has the same two sub-trees

  Patchmatch(const Patchmatch&) = delete;
  void operator=(const Patchmatch&) = delete;

  explicit Patchmatch(const Params& params, int device = 0) : gpu_(params, device) {}

  // Image1f Initialize(const Image1b& iml, const Image1b& imr, int downsample_factor),
  // patchmatch.hpp:43-45, patchmatch.cpp:52-87: (rows/f) x (cols/f) seed map.
  Image1f Initialize(const Image1b& iml, const Image1b& imr, int downsample_factor) {
    if (downsample_factor < 1) throw std::runtime_error("Initialize: downsample_factor < 1");
    Image1f seeds;
    pm_detail::create(seeds, iml.rows / downsample_factor, iml.cols / downsample_factor);
    Check(pm_cpu_initialize(gpu_.handle(), (const uint8_t*)iml.data, (const uint8_t*)imr.data,
                            iml.cols, iml.rows, pm_detail::step_bytes(iml), downsample_factor,
                            (float*)seeds.data, pm_detail::step_bytes(seeds)));
    return seeds;
  }

  // Image1f EstimateDisparity(const Image1b& iml, const Image1b& imr), patchmatch.hpp:48:
  // Initialize(iml, imr, 1) as the driver does (patchmatch_test.cpp:142), then the schedule.
  Image1f EstimateDisparity(const Image1b& iml, const Image1b& imr) {
    return EstimateDisparity(iml, imr, Initialize(iml, imr, 1));
  }

  // The same with the Initialize() seed map supplied by the caller.
  Image1f EstimateDisparity(const Image1b& iml, const Image1b& imr, const Image1f& seed) {
    Image1f disp;
    pm_detail::create(disp, iml.rows, iml.cols);
    Check(pm_cpu_estimate_disparity(gpu_.handle(), (const uint8_t*)iml.data, (const uint8_t*)imr.data,
                                    iml.cols, iml.rows, pm_detail::step_bytes(iml),
                                    (const float*)seed.data, (float*)disp.data,
                                    pm_detail::step_bytes(disp)));
    return disp;
  }

  // Stage methods (patchmatch.hpp:51-74) on the pair given to Load().
  void Load(const Image1b& iml, const Image1b& imr) {
    Check(pm_stage_load_pair(gpu_.handle(), (const uint8_t*)iml.data, (const uint8_t*)imr.data,
                             iml.cols, iml.rows, pm_detail::step_bytes(iml)));
    rows_ = iml.rows; cols_ = iml.cols;
  }
  // AddNoise(disp, amount, mask = disp > 0), patchmatch.cpp:143-155
  void AddNoise(Image1f& disp, float amount) {
    Upload(disp); Check(pm_cpu_add_noise(gpu_.handle(), amount)); Download(disp);
  }
  // Propagate(iml, imr, Gl, Gr, disp, f, patch_height, patch_width), patchmatch.cpp:248-311
  void Propagate(Image1f& disp, int patch_height, int patch_width) {
    Upload(disp); Check(pm_cpu_propagate(gpu_.handle(), patch_height, patch_width, -1)); Download(disp);
  }
  // RemoveBackground(..., win_by_factor = 2.0), patchmatch.cpp:314-360
  void RemoveBackground(Image1f& disp, int patch_height, int patch_width, float win_by_factor = 2.0f) {
    Upload(disp);
    Check(pm_cpu_remove_background(gpu_.handle(), patch_height, patch_width, win_by_factor));
    Download(disp);
  }

 private:
  void Check(int rc) { if (rc != PM_OK) throw std::runtime_error(pm_last_error(gpu_.handle())); }
  void Upload(const Image1f& d) {
    if (d.rows != rows_ || d.cols != cols_ || pm_detail::step_bytes(d) != sizeof(float) * (size_t)cols_)
      throw std::runtime_error("Patchmatch: dense disparity map of the loaded image size expected");
    Check(pm_cpu_set_disp(gpu_.handle(), (const float*)d.data));
  }
  void Download(Image1f& d) { Check(pm_cpu_get_disp(gpu_.handle(), (float*)d.data)); }

  pm::PatchmatchGpu gpu_;
  int rows_ = 0, cols_ = 0;
};

}  // namespace stereo
}  // namespace bm
