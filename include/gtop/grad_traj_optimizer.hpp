// grad_traj_optimizer.hpp — ROS-free, Eigen-free host mirror of the reference's GradTrajOptimizer
// (include/grad_traj_optimization/grad_traj_optimizer.h:20-131, src/grad_traj_optimizer.cpp upstream)
// on top of the C ABI of libgtop_b200 (include/gtop_c.h).  Header-only; link with -lgtop_b200 -ldl.
//
// Same public method names and argument meaning as upstream:
//   GradTrajOptimizer(), setPath, setKinoPath, optimizeTrajectory(step), getCoefficient, getSegmentTime,
//   initSDFMap, updateSDFMap, getCostCurve, and the static NLopt callback costFunc.
// Differences, all forced by the missing dependencies:
//   * parameters come from a gtop_params struct instead of ros::param::get (grad_traj_optimizer.cpp:5-32);
//   * vectors are gtop::Vec3 / std::vector<double>; Eigen overloads are compiled only when <Eigen/Eigen> exists;
//   * NLopt is reached through its C ABI with dlopen (no nlopt.hpp needed): the library path comes from
//     setNloptLibrary(), $GTOP_NLOPT_LIB, or the default soname libnlopt.so.0 / libnlopt.so;
//   * the only stopping rule upstream is wall clock (set_maxtime, :144-148); setMaxEval() adds a deterministic
//     one for reproducible solves.  maxtime keeps working as upstream.
// The objective itself — getCostAndGradient — runs on the GPU: costFunc forwards to gtop_objective.
#pragma once

#include <dlfcn.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

#include "../gtop_c.h"

#if defined(__has_include)
#if __has_include(<Eigen/Eigen>)
#include <Eigen/Eigen>
#define GTOP_HAVE_EIGEN 1
#endif
#endif

#define OPT_INITIAL_TRY 0
#define OPT_FIRST_STEP 1
#define OPT_SECOND_STEP 2

namespace gtop {

struct Vec3 {
  double x = 0, y = 0, z = 0;
  Vec3() {}
  Vec3(double x_, double y_, double z_) : x(x_), y(y_), z(z_) {}
};

class GradTrajOptimizer {
 public:
  explicit GradTrajOptimizer(const gtop_params& params, int device = 0) : prm_(params), device_(device) {}
  GradTrajOptimizer(const GradTrajOptimizer&) = delete;
  GradTrajOptimizer& operator=(const GradTrajOptimizer&) = delete;
  ~GradTrajOptimizer() {
    gtop_objective_ctx_destroy(ctx_);
    gtop_problems_destroy(problems_);
    gtop_sdf_destroy(sdf_);
  }

  // ---- upstream API -------------------------------------------------------------------------------
  void initSDFMap(const Vec3& map_size_3d, const Vec3& origin, double resolution) {  // :112-115
    // (the reference simply reassigns its SDFMap and may do so between solves: the callback context of the last solve
    // holds the old map and goes with it)
    gtop_objective_ctx_destroy(ctx_);
    ctx_ = nullptr;
    gtop_sdf_destroy(sdf_);
    sdf_ = nullptr;
    const double o[3] = {origin.x, origin.y, origin.z}, s[3] = {map_size_3d.x, map_size_3d.y, map_size_3d.z};
    check(gtop_sdf_create(o, resolution, s, device_, &sdf_));
  }

  void updateSDFMap(const std::vector<Vec3>& obs) {  // :117-126
    need(sdf_, "initSDFMap first");
    std::vector<double> xyz(obs.size() * 3);
    for (size_t i = 0; i < obs.size(); ++i) xyz[3 * i] = obs[i].x, xyz[3 * i + 1] = obs[i].y, xyz[3 * i + 2] = obs[i].z;
    check(gtop_sdf_reset(sdf_));
    check(gtop_sdf_set_occupancy(sdf_, xyz.data(), (int64_t)obs.size()));
    check(gtop_sdf_update_esdf(sdf_));
  }

  void setPath(const std::vector<Vec3>& way_points) {  // :67-110
    if (way_points.size() < 3) throw std::invalid_argument("setPath: need at least 3 waypoints (2 segments)");
    std::vector<double> w(way_points.size() * 3);
    for (size_t i = 0; i < way_points.size(); ++i)
      w[3 * i] = way_points[i].x, w[3 * i + 1] = way_points[i].y, w[3 * i + 2] = way_points[i].z;
    resetProblem();
    check(gtop_problems_create_from_paths(1, (int32_t)way_points.size() - 1, w.data(), &prm_, device_, &problems_));
    finishSetup();
  }

  // Pos/Vel/Acc: (m+1) x 3 row-major, Time: m   (:35-65; start vel/acc of Df are zero — SURVEY H14)
  void setKinoPath(const std::vector<double>& Pos, const std::vector<double>& Vel, const std::vector<double>& Acc,
                   const std::vector<double>& Time) {
    const size_t m = Time.size();
    if (m < 2 || Pos.size() != 3 * (m + 1) || Vel.size() != Pos.size() || Acc.size() != Pos.size())
      throw std::invalid_argument("setKinoPath: shape mismatch");
    resetProblem();
    check(gtop_problems_create_from_kino(1, (int32_t)m, Pos.data(), Vel.data(), Acc.data(), Time.data(), &prm_, device_,
                                         &problems_));
    finishSetup();
  }

  bool optimizeTrajectory(int step) {  // :128-243
    if (step != 0 && step != 1 && step != 2) std::printf("step number error, step should be 0, 1 or 2\n");
    need(problems_, "setPath first"), need(sdf_, "initSDFMap / updateSDFMap first");
    step_ = step;
    if (!ctx_) check(gtop_objective_ctx_create(sdf_, problems_, 0, step, &ctx_));
    check(gtop_objective_ctx_set_step(ctx_, step));
    if (builtin_mma_) {  // the library's own LD_MMA loop (csrc/mma_core.cuh): no libnlopt needed
      if (prm_.algorithm != 24) throw std::runtime_error("the built-in solver only implements algorithm 24 (LD_MMA)");
      if (max_eval_ <= 0) throw std::runtime_error("the built-in solver stops on setMaxEval(n), n > 0");
      std::vector<double> xb = dp_;
      double fb = 0;
      int nev = 0;
      vec_cost_.clear(), vec_time_.clear();
      time_start_ = std::chrono::steady_clock::now();
      check(gtop_mma_minimize((unsigned)n_, &GradTrajOptimizer::nloptTrampoline, this, lb_.data(), ub_.data(), xb.data(), &fb,
                              max_eval_, &nev));
      last_result_ = 5;  // NLOPT_MAXEVAL_REACHED
      dp_ = xb;
      coeff_.assign((size_t)m_ * 18, 0.0);
      check(gtop_coeff_from_x(problems_, 0, dp_.data(), coeff_.data()));
      last_min_f_ = fb;
      if (step == 2) iter_num_ = 0, total_time_ = 0;
      return true;
    }
    Nlopt& nl = nlopt();
    void* opt = nl.create(prm_.algorithm, (unsigned)n_);  // nlopt::opt(algorithm(alg), 3*num_dp)  :137-138
    if (!opt) throw std::runtime_error("nlopt_create failed");
    nl.set_min_objective(opt, &GradTrajOptimizer::nloptTrampoline, this);  // :140
    if (step == OPT_FIRST_STEP) nl.set_maxtime(opt, prm_.time_limit_1);   // :144-148
    else if (step == OPT_SECOND_STEP) nl.set_maxtime(opt, prm_.time_limit_2);
    if (max_eval_ > 0) nl.set_maxeval(opt, max_eval_);
    nl.set_lower_bounds(opt, lb_.data());  // :151-179
    nl.set_upper_bounds(opt, ub_.data());
    std::vector<double> x = dp_;           // :182-187
    double min_f = 0;
    vec_cost_.clear(), vec_time_.clear();
    time_start_ = std::chrono::steady_clock::now();
    last_result_ = nl.optimize(opt, x.data(), &min_f);  // :195
    nl.destroy(opt);
    dp_ = x;  // :202-207
    coeff_.assign((size_t)m_ * 18, 0.0);
    check(gtop_coeff_from_x(problems_, 0, dp_.data(), coeff_.data()));  // :230
    last_min_f_ = min_f;
    if (step == 2) iter_num_ = 0, total_time_ = 0;  // :236-239
    return true;                                     // :242
  }

  void getCoefficient(std::vector<double>& coeff) const { coeff = coeff_; }        // m x 18 row-major (:245-247)
  void getSegmentTime(std::vector<double>& seg_time) const { seg_time = seg_time_; }  // :249-251
  void getCostCurve(std::vector<double>& cost, std::vector<double>& time) const { cost = vec_cost_, time = vec_time_; }

  // NLopt C++ vfunc shape (grad_traj_optimizer.h:90-91, .cpp:554-562)
  static double costFunc(const std::vector<double>& x, std::vector<double>& grad, void* func_data) {
    GradTrajOptimizer* self = reinterpret_cast<GradTrajOptimizer*>(func_data);
    grad.resize(x.size());
    return nloptTrampoline((unsigned)x.size(), x.data(), grad.data(), self);
  }

#ifdef GTOP_HAVE_EIGEN
  void setPath(const std::vector<Eigen::Vector3d>& way_points) {
    std::vector<Vec3> w;
    for (auto& p : way_points) w.emplace_back(p(0), p(1), p(2));
    setPath(w);
  }
  void getCoefficient(Eigen::MatrixXd& coeff) const {
    coeff.resize(m_, 18);
    for (int i = 0; i < m_; ++i)
      for (int j = 0; j < 18; ++j) coeff(i, j) = coeff_[(size_t)i * 18 + j];
  }
  void getSegmentTime(Eigen::VectorXd& t) const {
    t.resize(m_);
    for (int i = 0; i < m_; ++i) t(i) = seg_time_[i];
  }
  void initSDFMap(Eigen::Vector3d s, Eigen::Vector3d o, double res) { initSDFMap(Vec3(s(0), s(1), s(2)), Vec3(o(0), o(1), o(2)), res); }
  void updateSDFMap(const std::vector<Eigen::Vector3d>& obs) {
    std::vector<Vec3> w;
    for (auto& p : obs) w.emplace_back(p(0), p(1), p(2));
    updateSDFMap(w);
  }
#endif

  // ---- additions for the ROS-free harness ---------------------------------------------------------
  void setMaxEval(int n) { max_eval_ = n; }                   // deterministic stop
  void useBuiltinMma(bool on) { builtin_mma_ = on; }          // gtop_mma_minimize instead of libnlopt (LD_MMA only)
  void setNloptLibrary(const std::string& path) { nlopt_path_ = path; }
  double evaluate(const std::vector<double>& x, std::vector<double>& grad, int step) {  // one getCostAndGradient
    need(problems_, "setPath first"), need(sdf_, "initSDFMap / updateSDFMap first");
    if (!ctx_) check(gtop_objective_ctx_create(sdf_, problems_, 0, step, &ctx_));
    check(gtop_objective_ctx_set_step(ctx_, step));
    grad.resize(x.size());
    return gtop_objective((unsigned)x.size(), x.data(), grad.data(), ctx_);
  }
  const std::vector<double>& derivatives() const { return dp_; }  // Dp flattened [x | y | z]
  int numSegments() const { return m_; }
  int lastResult() const { return last_result_; }
  double lastMinF() const { return last_min_f_; }
  int iterNum() const { return iter_num_; }
  double totalTime() const { return total_time_; }
  gtop_sdf* sdf() { return sdf_; }
  gtop_problems* problems() { return problems_; }

 private:
  struct Nlopt {  // hand-declared NLopt C ABI (NLopt >= 2.5)
    void* h = nullptr;
    void* (*create)(int, unsigned) = nullptr;
    void (*destroy)(void*) = nullptr;
    int (*set_min_objective)(void*, double (*)(unsigned, const double*, double*, void*), void*) = nullptr;
    int (*set_lower_bounds)(void*, const double*) = nullptr;
    int (*set_upper_bounds)(void*, const double*) = nullptr;
    int (*set_maxeval)(void*, int) = nullptr;
    int (*set_maxtime)(void*, double) = nullptr;
    int (*optimize)(void*, double*, double*) = nullptr;
  };

  Nlopt& nlopt() {
    if (nl_.h) return nl_;
    std::vector<std::string> cand;
    if (!nlopt_path_.empty()) cand.push_back(nlopt_path_);
    if (const char* e = std::getenv("GTOP_NLOPT_LIB")) cand.push_back(e);
    cand.push_back("libnlopt.so.0");
    cand.push_back("libnlopt.so");
    for (auto& c : cand)
      if ((nl_.h = dlopen(c.c_str(), RTLD_NOW | RTLD_LOCAL))) break;
    if (!nl_.h) throw std::runtime_error("NLopt shared library not found (set GTOP_NLOPT_LIB)");
#define GTOP_NL(f, name) *(void**)(&nl_.f) = dlsym(nl_.h, name)
    GTOP_NL(create, "nlopt_create");
    GTOP_NL(destroy, "nlopt_destroy");
    GTOP_NL(set_min_objective, "nlopt_set_min_objective");
    GTOP_NL(set_lower_bounds, "nlopt_set_lower_bounds");
    GTOP_NL(set_upper_bounds, "nlopt_set_upper_bounds");
    GTOP_NL(set_maxeval, "nlopt_set_maxeval");
    GTOP_NL(set_maxtime, "nlopt_set_maxtime");
    GTOP_NL(optimize, "nlopt_optimize");
#undef GTOP_NL
    if (!nl_.create || !nl_.optimize || !nl_.set_min_objective) throw std::runtime_error("NLopt library lacks the C API");
    return nl_;
  }

  static double nloptTrampoline(unsigned n, const double* x, double* grad, void* data) {
    GradTrajOptimizer* self = reinterpret_cast<GradTrajOptimizer*>(data);
    const auto t0 = std::chrono::steady_clock::now();
    self->iter_num_++;  // :284
    const double cost = gtop_objective(n, x, grad, self->ctx_);
    const auto t1 = std::chrono::steady_clock::now();
    self->total_time_ += std::chrono::duration<double>(t1 - t0).count();  // :435-436
    self->vec_time_.push_back(std::chrono::duration<double>(t1 - self->time_start_).count());
    if (self->vec_cost_.empty() || self->vec_cost_.back() > cost) self->vec_cost_.push_back(cost);  // :438-447
    else self->vec_cost_.push_back(self->vec_cost_.back());
    return cost;
  }

  void resetProblem() {
    gtop_objective_ctx_destroy(ctx_);
    ctx_ = nullptr;
    gtop_problems_destroy(problems_);
    problems_ = nullptr;
  }

  void finishSetup() {
    int32_t np = 0, m = 0, n = 0;
    check(gtop_problems_dims(problems_, &np, &m, &n));
    m_ = m, n_ = n;
    seg_time_.assign(m, 0.0), dp_.assign(n, 0.0), lb_.assign(n, 0.0), ub_.assign(n, 0.0);
    check(gtop_problems_seg_time(problems_, seg_time_.data()));
    check(gtop_problems_x0(problems_, dp_.data()));
    check(gtop_problems_bounds(problems_, lb_.data(), ub_.data()));
    coeff_.assign((size_t)m * 18, 0.0);
    check(gtop_coeff_from_x(problems_, 0, dp_.data(), coeff_.data()));  // initial straight-line coefficients (:89)
  }

  static void check(int rc) {
    if (rc != GTOP_OK) throw std::runtime_error(std::string("gtop: ") + gtop_last_error());
  }
  static void need(const void* p, const char* what) {
    if (!p) throw std::logic_error(what);
  }

  gtop_params prm_;
  int device_ = 0;
  gtop_sdf* sdf_ = nullptr;
  gtop_problems* problems_ = nullptr;
  gtop_objective_ctx* ctx_ = nullptr;
  Nlopt nl_;
  std::string nlopt_path_;
  int m_ = 0, n_ = 0, step_ = 1, max_eval_ = 0, last_result_ = 0, iter_num_ = 0;
  bool builtin_mma_ = false;
  double last_min_f_ = 0, total_time_ = 0;
  std::vector<double> seg_time_, dp_, lb_, ub_, coeff_, vec_cost_, vec_time_;
  std::chrono::steady_clock::time_point time_start_;
};

}  // namespace gtop
