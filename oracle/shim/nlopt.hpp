// oracle/shim/nlopt.hpp — the part of NLopt's C++ wrapper that the reference calls
// (grad_traj_optimizer.cpp:135-195: nlopt::srand, nlopt::opt(algorithm, n), set_min_objective(vfunc, data),
// set_maxtime, set_lower_bounds / set_upper_bounds(std::vector), optimize(x, f)), on top of the C API of the NLopt
// 2.5.0 binary the reference vendors (lib/libnlopt.so.0.9.0, loaded with dlopen; NLopt's own headers are not in the
// reference tree).  TEST INFRASTRUCTURE ONLY (see oracle/shim/Eigen/Eigen).
//
// One extension, because the reference's only stopping rule is wall-clock time:
// nlopt::shim::max_evaluations(), when positive, is applied with nlopt_set_maxeval before every optimize() so that
// the wrapper (oracle/ref_capi.cpp) can run reproducible solves; the reference's set_maxtime call still goes through.
#ifndef GTOP_ORACLE_SHIM_NLOPT
#define GTOP_ORACLE_SHIM_NLOPT
#include <dlfcn.h>

#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

namespace nlopt {

enum algorithm { LD_MMA = 24, NUM_ALGORITHMS = 44 };
enum result {
  FAILURE = -1, INVALID_ARGS = -2, OUT_OF_MEMORY = -3, ROUNDOFF_LIMITED = -4, FORCED_STOP = -5,
  SUCCESS = 1, STOPVAL_REACHED = 2, FTOL_REACHED = 3, XTOL_REACHED = 4, MAXEVAL_REACHED = 5, MAXTIME_REACHED = 6
};
typedef double (*vfunc)(const std::vector<double>& x, std::vector<double>& grad, void* data);

namespace shim {
typedef double (*c_func)(unsigned n, const double* x, double* grad, void* data);
struct Api {
  void* (*create)(int, unsigned);
  void (*destroy)(void*);
  int (*set_min_objective)(void*, c_func, void*);
  int (*set_lower_bounds)(void*, const double*);
  int (*set_upper_bounds)(void*, const double*);
  int (*set_maxtime)(void*, double);
  int (*set_maxeval)(void*, int);
  int (*optimize)(void*, double*, double*);
  void (*srand)(unsigned long);
};
inline std::string& library_path() {
  static std::string p = [] {
    const char* e = std::getenv("GTOP_NLOPT_LIB");
    return std::string(e ? e : "libnlopt.so.0");
  }();
  return p;
}
inline const Api& api() {
  static Api a = [] {
    void* h = dlopen(library_path().c_str(), RTLD_NOW | RTLD_LOCAL);
    if (!h) throw std::runtime_error(std::string("shim nlopt: cannot load ") + library_path() + ": " + dlerror());
    Api r;
    auto sym = [&](const char* name) {
      void* s = dlsym(h, name);
      if (!s) throw std::runtime_error(std::string("shim nlopt: missing symbol ") + name);
      return s;
    };
    r.create = (void* (*)(int, unsigned))sym("nlopt_create");
    r.destroy = (void (*)(void*))sym("nlopt_destroy");
    r.set_min_objective = (int (*)(void*, c_func, void*))sym("nlopt_set_min_objective");
    r.set_lower_bounds = (int (*)(void*, const double*))sym("nlopt_set_lower_bounds");
    r.set_upper_bounds = (int (*)(void*, const double*))sym("nlopt_set_upper_bounds");
    r.set_maxtime = (int (*)(void*, double))sym("nlopt_set_maxtime");
    r.set_maxeval = (int (*)(void*, int))sym("nlopt_set_maxeval");
    r.optimize = (int (*)(void*, double*, double*))sym("nlopt_optimize");
    r.srand = (void (*)(unsigned long))sym("nlopt_srand");
    return r;
  }();
  return a;
}
inline int& max_evaluations() {
  static int n = 0;
  return n;
}
}  // namespace shim

inline void srand(unsigned long seed) { shim::api().srand(seed); }

// Value semantics like the real wrapper's, without nlopt_copy: the reference copy-assigns a freshly built opt
// (`optimizer = opt;`, :139) BEFORE any setter is called, so the settings are recorded here and applied to a C object
// built inside optimize().
class opt {
 public:
  opt() : alg_(LD_MMA), n_(0), f_(nullptr), data_(nullptr), maxtime_(0.0) {}
  opt(algorithm a, unsigned n) : alg_(a), n_(n), f_(nullptr), data_(nullptr), maxtime_(0.0) {}
  void set_min_objective(vfunc f, void* data) { f_ = f, data_ = data; }
  void set_maxtime(double t) { maxtime_ = t; }
  void set_lower_bounds(const std::vector<double>& lb) { lb_ = lb; }
  void set_upper_bounds(const std::vector<double>& ub) { ub_ = ub; }
  result optimize(std::vector<double>& x, double& opt_f) {
    const shim::Api& A = shim::api();
    if (x.size() != n_) throw std::invalid_argument("shim nlopt: dimension mismatch");
    void* o = A.create((int)alg_, n_);
    if (!o) throw std::bad_alloc();
    Tramp tr{f_, data_, std::vector<double>(n_), std::vector<double>(n_)};
    A.set_min_objective(o, &opt::trampoline, &tr);
    if (lb_.size() == n_) A.set_lower_bounds(o, lb_.data());
    if (ub_.size() == n_) A.set_upper_bounds(o, ub_.data());
    if (maxtime_ > 0.0) A.set_maxtime(o, maxtime_);
    if (shim::max_evaluations() > 0) A.set_maxeval(o, shim::max_evaluations());
    const int r = A.optimize(o, x.data(), &opt_f);
    A.destroy(o);
    return (result)r;
  }

 private:
  struct Tramp {
    vfunc f;
    void* data;
    std::vector<double> x, g;
  };
  // the real wrapper copies x into a std::vector and the gradient back out around the vfunc, exactly like this
  static double trampoline(unsigned n, const double* x, double* grad, void* d) {
    Tramp* t = (Tramp*)d;
    t->x.assign(x, x + n);
    const double v = t->f(t->x, t->g, t->data);
    if (grad)
      for (unsigned i = 0; i < n; ++i) grad[i] = t->g[i];
    return v;
  }
  algorithm alg_;
  unsigned n_;
  vfunc f_;
  void* data_;
  double maxtime_;
  std::vector<double> lb_, ub_;
};

}  // namespace nlopt
#endif
