// pm_yaml.cpp -- PatchmatchGpu::Params from a YAML file.
//
// The reference parses its params with cv::FileStorage through a home-grown
// YamlParser (src/vehicle/params/yaml_parser.hpp:56-77, yaml_parser.cpp:64-82):
// `%YAML:1.0` files, nested maps named after classes, one scalar per key, and a
// glog CHECK abort on a missing key. OpenCV is not a dependency of this library,
// so the subset those files use is read here: block maps by indentation, scalar
// values, `#` comments, the `%YAML` directive and `---`. A missing required key
// is an error return instead of an abort.
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/pm_b200.h"

namespace {

typedef std::map<std::string, std::string> Flat;  // "A/B/key" -> scalar text

std::string trim(const std::string& s) {
  size_t a = 0, b = s.size();
  while (a < b && std::isspace((unsigned char)s[a])) ++a;
  while (b > a && std::isspace((unsigned char)s[b - 1])) --b;
  return s.substr(a, b - a);
}

std::string strip_comment(const std::string& s) {
  bool sq = false, dq = false;
  for (size_t i = 0; i < s.size(); ++i) {
    const char c = s[i];
    if (c == '\'' && !dq) sq = !sq;
    else if (c == '"' && !sq) dq = !dq;
    else if (c == '#' && !sq && !dq && (i == 0 || std::isspace((unsigned char)s[i - 1])))
      return s.substr(0, i);
  }
  return s;
}

bool parse_yaml(const std::string& path, Flat* out, std::string* err) {
  std::ifstream f(path.c_str());
  if (!f) { *err = "cannot open " + path; return false; }
  std::vector<std::pair<int, std::string> > stack;  // (indent, key)
  std::string line;
  int lineno = 0;
  while (std::getline(f, line)) {
    ++lineno;
    if (!line.empty() && line[line.size() - 1] == '\r') line.erase(line.size() - 1);
    if (line.compare(0, 5, "%YAML") == 0 || trim(line) == "---" || trim(line) == "...") continue;
    line = strip_comment(line);
    if (trim(line).empty()) continue;
    if (line.substr(0, line.find_first_not_of(" \t")).find('\t') != std::string::npos) {
      std::ostringstream m; m << path << ":" << lineno << ": tab in indentation";
      *err = m.str(); return false;
    }
    const int indent = (int)line.find_first_not_of(' ');
    std::string body = trim(line);
    if (body[0] == '-') continue;  // block sequences are not part of the subset: ignored
    const size_t colon = body.find(':');
    if (colon == std::string::npos) {
      std::ostringstream m; m << path << ":" << lineno << ": expected `key: value`";
      *err = m.str(); return false;
    }
    std::string key = trim(body.substr(0, colon));
    std::string val = trim(body.substr(colon + 1));
    if (key.size() >= 2 && (key[0] == '"' || key[0] == '\'')) key = key.substr(1, key.size() - 2);
    while (!stack.empty() && stack.back().first >= indent) stack.pop_back();
    if (val.empty()) {
      stack.push_back(std::make_pair(indent, key));
      continue;
    }
    if (val.size() >= 2 && (val[0] == '"' || val[0] == '\'') && val[val.size() - 1] == val[0])
      val = val.substr(1, val.size() - 2);
    std::string full;
    for (size_t i = 0; i < stack.size(); ++i) full += stack[i].second + "/";
    (*out)[full + key] = val;
  }
  return true;
}

struct Reader {
  const Flat& flat;
  std::string prefix, path, err;
  bool ok = true;
  Reader(const Flat& f, const std::string& pre, const std::string& p) : flat(f), prefix(pre), path(p) {}

  const std::string* find(const std::string& key) const {
    Flat::const_iterator it = flat.find(prefix + key);
    return it == flat.end() ? nullptr : &it->second;
  }
  void missing(const std::string& key) {
    if (ok) err = path + ": required key `" + prefix + key + "` not found";
    ok = false;
  }
  void bad(const std::string& key, const std::string& v) {
    if (ok) err = path + ": key `" + prefix + key + "`: cannot parse `" + v + "`";
    ok = false;
  }
  bool num(const std::string& key, double* out, bool required) {
    const std::string* v = find(key);
    if (!v) { if (required) missing(key); return false; }
    std::string s = *v;
    if (s == "true" || s == "True") { *out = 1; return true; }
    if (s == "false" || s == "False") { *out = 0; return true; }
    char* end = nullptr;
    const double d = std::strtod(s.c_str(), &end);
    if (end == s.c_str() || *end) { bad(key, s); return false; }
    *out = d;
    return true;
  }
  template <class T> void get(const std::string& key, T* field, bool required) {
    double d;
    if (num(key, &d, required)) *field = (T)d;
  }
  void get_u64(const std::string& key, uint64_t* field) {
    const std::string* v = find(key);
    if (!v) return;
    char* end = nullptr;
    const unsigned long long u = std::strtoull(v->c_str(), &end, 0);
    if (end == v->c_str() || *end) { bad(key, *v); return; }
    *field = (uint64_t)u;
  }
  // enum given by name or by number
  void get_enum(const std::string& key, int* field, const char* const* names, int n) {
    const std::string* v = find(key);
    if (!v) return;
    for (int i = 0; i < n; ++i)
      if (*v == names[i]) { *field = i; return; }
    char* end = nullptr;
    const long l = std::strtol(v->c_str(), &end, 10);
    if (end == v->c_str() || *end || l < 0 || l >= n) { bad(key, *v); return; }
    *field = (int)l;
  }
};

}  // namespace

extern "C" int pm_params_load_yaml(const char* path, const char* subtree, pm_params* p, char* err,
                                   size_t err_len) {
  auto report = [&](const std::string& m) {
    if (err && err_len) { std::snprintf(err, err_len, "%s", m.c_str()); }
    return PM_ERR_YAML;
  };
  if (!path || !p) return report("pm_params_load_yaml: null argument");
  Flat flat;
  std::string e;
  if (!parse_yaml(path, &flat, &e)) return report(e);
  pm_params_default(p);
  std::string prefix = subtree ? subtree : "";
  if (!prefix.empty() && prefix[prefix.size() - 1] != '/') prefix += "/";

  // LoadParams (patchmatch_gpu.cu:11-15): the two sub-trees, every key required
  // (feature_detector.cpp:33-39, stereo_matcher.cpp:12-19; CHECK at yaml_parser.cpp:82).
  Reader fd(flat, prefix + "FeatureDetector/", path);
  fd.get("max_features_per_frame", &p->fd_max_features_per_frame, true);
  fd.get("min_distance_btw_tracked_and_detected_features", &p->fd_min_distance, true);
  fd.get("gftt_quality_level", &p->fd_gftt_quality_level, true);
  fd.get("gftt_block_size", &p->fd_gftt_block_size, true);
  fd.get("gftt_use_harris_corner_detector", &p->fd_gftt_use_harris, true);
  fd.get("gftt_k", &p->fd_gftt_k, false);
  if (!fd.ok) return report(fd.err);
  Reader sm(flat, prefix + "StereoMatcher/", path);
  sm.get("templ_cols", &p->sm_templ_cols, true);
  sm.get("templ_rows", &p->sm_templ_rows, true);
  sm.get("max_disp", &p->sm_max_disp, true);
  sm.get("max_matching_cost", &p->sm_max_matching_cost, true);
  sm.get("bidirectional", &p->sm_bidirectional, true);
  sm.get("subpixel_refinement", &p->sm_subpixel_refinement, true);
  if (!sm.ok) return report(sm.err);

  // struct fields the reference never reads from YAML (patchmatch_gpu.h:85-88) and the
  // extension keys: optional, top level of the PatchMatch tree.
  Reader r(flat, prefix, path);
  p->max_disp = p->sm_max_disp;  // `max_disp` defaults to StereoMatcher/max_disp
  r.get("cost_alpha", &p->cost_alpha, false);
  r.get("patchmatch_iters", &p->patchmatch_iters, false);
  r.get("init_dilate_factor", &p->init_dilate_factor, false);
  r.get("cost_improve_factor", &p->cost_improve_factor, false);
  r.get("patch_size", &p->patch_size, false);
  r.get("sweep_chunks", &p->sweep_chunks, false);
  r.get("sweep_overlap", &p->sweep_overlap, false);
  r.get("noise_scale0", &p->noise_scale0, false);
  r.get_u64("seed", &p->seed);
  r.get("max_disp", &p->max_disp, false);
  r.get("clamp_disp", &p->clamp_disp, false);
  r.get("pyramid_levels", &p->pyramid_levels, false);
  r.get("subpixel", &p->subpixel, false);
  r.get("median_ksize", &p->median_ksize, false);
  r.get("max_batch", &p->max_batch, false);
  r.get("random_search_k", &p->random_search_k, false);
  static const char* const init_names[] = {"sparse", "random"};
  static const char* const cost_names[] = {"l1grad_x5", "l1grad_full", "census"};
  static const char* const lr_names[] = {"ratio", "abs1px"};
  static const char* const noise_names[] = {"always", "improve"};
  r.get_enum("init_mode", &p->init_mode, init_names, 2);
  r.get_enum("cost_mode", &p->cost_mode, cost_names, 3);
  r.get_enum("lr_mode", &p->lr_mode, lr_names, 2);
  r.get_enum("noise_accept", &p->noise_accept, noise_names, 2);
  if (!r.ok) return report(r.err);
  if (err && err_len) err[0] = 0;
  return PM_OK;
}
