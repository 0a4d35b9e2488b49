// Drives the C++ host class (include/patchmatch_gpu.h) the way the reference's gtest drives
// bm::pm::PatchmatchGpu (test/stereo_matching/patchmatch_gpu_test.cpp:47-92 in the reference).
// Usage: shim_test <yaml> <width> <height> <left.raw> <right.raw> <disp_out.raw> <dispr_out.raw>
//        shim_test --no-gpu <yaml>     (checks Params parsing and the "no CPU path" error)
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>

#include "patchmatch.hpp"
#include "patchmatch_gpu.h"

using namespace bm;
using namespace bm::pm;

static std::vector<char> slurp(const char* path) {
  std::ifstream f(path, std::ios::binary);
  return std::vector<char>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

int main(int argc, char** argv) {
  if (argc >= 3 && std::string(argv[1]) == "--no-gpu") {
    PatchmatchGpu::Params params(argv[2], "PatchmatchGpu");
    std::printf("params cost_alpha=%.2f iters=%d templ=%dx%d max_disp=%d levels=%d\n", params.cost_alpha,
                params.patchmatch_iters, params.matcher_params.templ_cols,
                params.matcher_params.templ_rows, params.matcher_params.max_disp, params.pyramid_levels);
    try {
      PatchmatchGpu pm(params);
      std::printf("engine created\n");
    } catch (const std::runtime_error& e) {
      std::printf("create failed: %s\n", e.what());
    }
    try {
      PatchmatchGpu::Params bad("/nonexistent.yaml");
      return 1;
    } catch (const std::runtime_error& e) {
      std::printf("yaml error: %s\n", e.what());
    }
    return 0;
  }
  if (argc >= 7 && std::string(argv[1]) == "--band") {
    // shim_test --band <yaml> <width> <height> <left.raw> <right.raw>: the whole frame as ONE
    // band through MatchBand must equal Match (more bands need an interconnect: see
    // tests/test_gpu_bands.py and tools/band_bench.py)
    PatchmatchGpu::Params params(argv[2], "PatchmatchGpu");
    params.pyramid_levels = 1;
    const int w = std::atoi(argv[3]), h = std::atoi(argv[4]);
    Image1b il(h, w), ir(h, w);
    const std::vector<char> a = slurp(argv[5]), b = slurp(argv[6]);
    if ((int)a.size() != w * h || (int)b.size() != w * h) return 3;
    std::memcpy(il.data, a.data(), a.size());
    std::memcpy(ir.data, b.data(), b.size());
    PatchmatchGpu pm(params);
    Image1f d0, r0, d1, r1;
    pm.Match(il, ir, d0, r0);
    const pm_band_layout lay = PatchmatchGpu::BandLayout(params, h, 0, 1);
    if (lay.own_lo != 0 || lay.own_hi != h || lay.load_hi != h) return 4;
    pm.MatchBand(il, ir, h, 0, 1, nullptr, nullptr, d1, r1);
    const bool same = std::memcmp(d0.data, d1.data, sizeof(float) * w * h) == 0 &&
                      std::memcmp(r0.data, r1.data, sizeof(float) * w * h) == 0;
    std::printf(same ? "band ok\n" : "band differs\n");
    try {
      PatchmatchGpu::BandLayout(params, h, 0, 3);   // 3 does not divide 16 chunks
      return 5;
    } catch (const std::runtime_error& e) {
      std::printf("band error: %s\n", e.what());
    }
    return same ? 0 : 6;
  }
  if (argc >= 9 && std::string(argv[1]) == "--sparse") {
    // shim_test --sparse <width> <height> <left.raw> <right.raw> <disp.raw> <dispr.raw> <seed.raw>
    //           <cpu.raw>: the reference's own drivers with their default params
    // (patchmatch_gpu_test.cpp:47-92: Match seeds itself with SparseInit; patchmatch_test.cpp:
    // 116-188: stereo::Patchmatch Initialize + schedule = EstimateDisparity)
    PatchmatchGpu::Params params;
    const int w = std::atoi(argv[2]), h = std::atoi(argv[3]);
    Image1b il(h, w), ir(h, w);
    const std::vector<char> a = slurp(argv[4]), b = slurp(argv[5]);
    if ((int)a.size() != w * h || (int)b.size() != w * h) return 3;
    std::memcpy(il.data, a.data(), a.size());
    std::memcpy(ir.data, b.data(), b.size());
    PatchmatchGpu pm(params);
    Image1f disp, dispr;
    pm.Match(il, ir, disp, dispr);
    const Image1f seeds = pm.SparseInit(il, ir, params.init_dilate_factor);
    stereo::Patchmatch cpu(params);
    const Image1f cpu_disp = cpu.EstimateDisparity(il, ir);
    std::ofstream(argv[6], std::ios::binary).write((const char*)disp.data, sizeof(float) * w * h);
    std::ofstream(argv[7], std::ios::binary).write((const char*)dispr.data, sizeof(float) * w * h);
    std::ofstream(argv[8], std::ios::binary).write((const char*)seeds.data, sizeof(float) * w * h);
    std::ofstream(argv[9], std::ios::binary).write((const char*)cpu_disp.data, sizeof(float) * w * h);
    return 0;
  }
  if (argc < 8) return 2;
  PatchmatchGpu::Params params(argv[1], "PatchmatchGpu");
  const int w = std::atoi(argv[2]), h = std::atoi(argv[3]);
  Image1b il(h, w), ir(h, w);
  const std::vector<char> a = slurp(argv[4]), b = slurp(argv[5]);
  if ((int)a.size() != w * h || (int)b.size() != w * h) return 3;
  std::memcpy(il.data, a.data(), a.size());
  std::memcpy(ir.data, b.data(), b.size());
  PatchmatchGpu pm(params);
  Image1f disp, dispr;
  for (int i = 0; i < 2; ++i) pm.Match(il, ir, disp, dispr);  // the reference test loops too
  std::ofstream(argv[6], std::ios::binary).write((const char*)disp.data, sizeof(float) * w * h);
  std::ofstream(argv[7], std::ios::binary).write((const char*)dispr.data, sizeof(float) * w * h);
  std::printf("ok %dx%d\n", disp.cols, disp.rows);
  return 0;
}
