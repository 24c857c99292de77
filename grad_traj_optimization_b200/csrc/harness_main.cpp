// gtop_harness — ROS-free driver for the B200 GTOP inner loop (stands in for the reference's demo nodes
// src/opti_node.cpp and src/example_text_input.cpp, which need ROS + RViz).
//
//   gtop_harness <config.cfg> [--maxeval N] [--steps 1,2] [--device D] [--nlopt /path/libnlopt.so.0 | --builtin-mma] [--dump out.txt]
//
// Loads a fixture file (configs/*.cfg: map, obstacle recipe, waypoints, parameters), builds the SDF on the GPU
// (initSDFMap + updateSDFMap), sets the path, prints cost / gradient norm at the initial derivatives, and — when
// an NLopt shared library is available — runs optimizeTrajectory for the requested step sequence (the demos run
// OPT_FIRST_STEP then OPT_SECOND_STEP: example_text_input.cpp:121-127; opti_node runs OPT_SECOND_STEP only:
// opti_node.cpp:104).  All evaluations of the objective happen on the GPU through the C ABI.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/gtop/grad_traj_optimizer.hpp"

namespace {

struct Config {
  gtop::Vec3 map_size, map_origin;
  double resolution = 0.1, pillar_z_max = 3.5;
  int pillar_half_width = 2;
  std::vector<std::pair<double, double>> pillars;
  std::vector<std::vector<double>> loop_boxes;
  std::vector<gtop::Vec3> waypoints;
  std::map<std::string, double> params;
};

bool load(const std::string& path, Config& c) {
  std::ifstream f(path);
  if (!f) return false;
  std::string line;
  while (std::getline(f, line)) {
    const size_t h = line.find('#');
    if (h != std::string::npos) line.resize(h);
    std::istringstream is(line);
    std::string key;
    if (!(is >> key)) continue;
    std::vector<double> v;
    std::string tok, name;
    if (key == "param") {
      is >> name;
      double x;
      is >> x;
      c.params[name] = x;
      continue;
    }
    double x;
    while (is >> x) v.push_back(x);
    if (key == "map_size" && v.size() == 3) c.map_size = gtop::Vec3(v[0], v[1], v[2]);
    else if (key == "map_origin" && v.size() == 3) c.map_origin = gtop::Vec3(v[0], v[1], v[2]);
    else if (key == "resolution" && v.size() == 1) c.resolution = v[0];
    else if (key == "pillar_half_width" && v.size() == 1) c.pillar_half_width = (int)v[0];
    else if (key == "pillar_z_max" && v.size() == 1) c.pillar_z_max = v[0];
    else if (key == "pillar" && v.size() == 2) c.pillars.emplace_back(v[0], v[1]);
    else if (key == "loop_box" && v.size() == 9) c.loop_boxes.push_back(v);
    else if (key == "waypoint" && v.size() == 3) c.waypoints.emplace_back(v[0], v[1], v[2]);
    else {
      std::fprintf(stderr, "%s: bad line '%s'\n", path.c_str(), line.c_str());
      return false;
    }
  }
  return true;
}

// obstacle recipes, expanded exactly as the demos do
std::vector<gtop::Vec3> obstacles(const Config& c) {
  std::vector<gtop::Vec3> pts;
  const int th = c.pillar_half_width;
  if (!c.pillars.empty())
    for (float z = 0; z < c.pillar_z_max; z += c.resolution)  // float accumulation: example_text_input.cpp:61
      for (auto& p : c.pillars)
        for (int m = -th; m <= th; m++)
          for (int n = -th; n <= th; n++) pts.emplace_back(p.first + m * c.resolution, p.second + n * c.resolution, z);
  auto loop = [](double s, double e, double d, std::vector<double>& out) {
    if (d > 0) for (double v = s; v <= e; v += d) out.push_back(v);
    else for (double v = s; v >= e; v += d) out.push_back(v);
  };
  for (auto& b : c.loop_boxes) {  // opti_node.cpp:66-78
    std::vector<double> xs, ys, zs;
    loop(b[0], b[1], b[2], xs), loop(b[3], b[4], b[5], ys), loop(b[6], b[7], b[8], zs);
    for (double x : xs)
      for (double y : ys)
        for (double z : zs) pts.emplace_back(x, y, z);
  }
  return pts;
}

gtop_params to_params(const Config& c) {
  gtop_params p;
  gtop_params_default(&p);
  auto get = [&](const char* k, double& dst) {
    auto it = c.params.find(k);
    if (it != c.params.end()) dst = it->second;
  };
  get("time_limit_1", p.time_limit_1), get("time_limit_2", p.time_limit_2), get("dt", p.deltat);
  get("ws", p.w_smooth), get("wc", p.w_collision);
  get("alpha", p.alpha), get("r", p.r), get("d0", p.d0);
  get("alpha_v", p.alpha_v), get("r_v", p.r_v), get("v0", p.v0);
  get("alpha_a", p.alpha_a), get("r_a", p.r_a), get("a0", p.a0);
  get("bos", p.bos), get("vos", p.vos), get("aos", p.aos);
  get("mean_v", p.mean_v), get("mean_a", p.mean_a), get("init_time", p.init_time);
  double alg = p.algorithm, ns = p.samples_per_segment;
  get("alg", alg), get("samples_per_segment", ns);
  p.algorithm = (int)alg, p.samples_per_segment = (int)ns;
  return p;
}

double inf_norm(const std::vector<double>& g) {
  double m = 0;
  for (double v : g) m = std::max(m, std::fabs(v));
  return m;
}

}  // namespace

int main(int argc, char** argv) {
  if (argc < 2) {
    std::fprintf(stderr, "usage: %s <config.cfg> [--maxeval N] [--steps 1,2] [--device D] [--nlopt PATH | --builtin-mma] [--dump FILE]\n", argv[0]);
    return 2;
  }
  int maxeval = 100, device = 0;
  bool builtin = false;
  std::string steps = "2", nlopt_path, dump;
  for (int i = 2; i < argc; ++i) {
    const std::string a = argv[i];
    auto next = [&]() -> const char* { return (i + 1 < argc) ? argv[++i] : ""; };
    if (a == "--maxeval") maxeval = std::atoi(next());
    else if (a == "--steps") steps = next();
    else if (a == "--device") device = std::atoi(next());
    else if (a == "--nlopt") nlopt_path = next();
    else if (a == "--builtin-mma") builtin = true;
    else if (a == "--dump") dump = next();
  }
  Config cfg;
  if (!load(argv[1], cfg)) {
    std::fprintf(stderr, "cannot load %s\n", argv[1]);
    return 2;
  }
  try {
    gtop::GradTrajOptimizer opt(to_params(cfg), device);
    opt.initSDFMap(cfg.map_size, cfg.map_origin, cfg.resolution);
    const std::vector<gtop::Vec3> obs = obstacles(cfg);
    opt.updateSDFMap(obs);
    float ms[4];
    gtop_sdf_last_edt_ms(opt.sdf(), ms);
    int32_t grid[3];
    gtop_sdf_grid(opt.sdf(), grid);
    std::printf("map %dx%dx%d, %zu obstacle points, GPU EDT %.3f ms (z %.3f, y %.3f, x %.3f)\n", grid[0], grid[1], grid[2],
                obs.size(), ms[3], ms[0], ms[1], ms[2]);
    opt.setPath(cfg.waypoints);
    std::vector<double> T;
    opt.getSegmentTime(T);
    std::printf("segments %d, variables %zu, T =", opt.numSegments(), opt.derivatives().size());
    for (double t : T) std::printf(" %.6f", t);
    std::printf("\n");
    std::vector<double> g;
    for (int step : {1, 2}) {
      const double c = opt.evaluate(opt.derivatives(), g, step);
      std::printf("x0 step %d: cost %.15g  |grad|_inf %.15g\n", step, c, inf_norm(g));
    }
    if (!nlopt_path.empty()) opt.setNloptLibrary(nlopt_path);
    opt.setMaxEval(maxeval);
    opt.useBuiltinMma(builtin);
    bool solved = true;
    std::stringstream ss(steps);
    std::string tok;
    while (std::getline(ss, tok, ',')) {
      const int step = std::atoi(tok.c_str());
      try {
        opt.optimizeTrajectory(step);
      } catch (const std::exception& e) {
        std::printf("solve skipped: %s\n", e.what());
        solved = false;
        break;
      }
      std::vector<double> cc, ct;
      opt.getCostCurve(cc, ct);
      std::printf("optimizeTrajectory(%d): nlopt result %d, min f %.12g after %zu evaluations (%.3f ms in the objective)\n", step,
                  opt.lastResult(), opt.lastMinF(), cc.size(), ct.empty() ? 0.0 : ct.back() * 1e3);
    }
    std::vector<double> coeff;
    opt.getCoefficient(coeff);
    if (!dump.empty()) {
      std::ofstream o(dump);
      o.precision(17);
      for (int s = 0; s < opt.numSegments(); ++s) {
        for (int k = 0; k < 18; ++k) o << coeff[(size_t)s * 18 + k] << (k == 17 ? '\n' : ' ');
      }
    }
    std::printf("%s; first segment x-coefficients:", solved ? "done" : "done (no solve)");
    for (int k = 0; k < 6; ++k) std::printf(" %.9g", coeff[k]);
    std::printf("\n");
    // the evaluation opti_node runs on the result (src/opti_node.cpp:136-142, PolynomialTraj)
    double met[GTOP_NUM_TRAJ_METRICS];
    if (gtop_traj_metrics_host(opt.problems(), opt.derivatives().data(), 1, 0, 1, met) == GTOP_OK)
      std::printf("trajectory: time %.6f s, length %.6f m, jerk %.6g, mean/max vel %.6g / %.6g, mean/max acc %.6g / %.6g\n", met[0],
                  met[1], met[2], met[3], met[4], met[5], met[6]);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}
