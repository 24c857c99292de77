// gtop_oracle_capi.cpp — extern "C" surface of the CPU ORACLE for ctypes.
// TEST INFRASTRUCTURE ONLY (see gtop_oracle.hpp).  PARITY UNPINNED upstream.
#include <atomic>
#include <chrono>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "gtop_oracle.hpp"

#include <algorithm>

using namespace gtop_oracle;

static thread_local std::string g_err;
#define GTO_TRY try {
#define GTO_CATCH(ret)              \
  }                                 \
  catch (const std::exception& e) { \
    g_err = e.what();               \
    return ret;                     \
  }

extern "C" {

const char* gto_last_error() { return g_err.c_str(); }

// ---- map -----------------------------------------------------------------------------------
void* gto_map_create(const double origin[3], double res, const double size[3]) {
  GTO_TRY return new SDFMap(origin, res, size);
  GTO_CATCH(nullptr)
}
void gto_map_destroy(void* h) { delete (SDFMap*)h; }
void gto_map_grid(void* h, int out[3]) {
  for (int i = 0; i < 3; ++i) out[i] = ((SDFMap*)h)->grid_size[i];
}
void gto_map_reset(void* h) { ((SDFMap*)h)->resetBuffer(); }
void gto_map_reset_range(void* h, const double* min_pos, const double* max_pos) { ((SDFMap*)h)->resetBuffer(min_pos, max_pos); }
void gto_map_set_update_range(void* h, const double* min_pos, const double* max_pos) {
  ((SDFMap*)h)->setUpdateRange(min_pos, max_pos);
}
void gto_map_update_range(void* h, int out[6]) {
  for (int i = 0; i < 3; ++i) out[i] = ((SDFMap*)h)->min_vec[i], out[3 + i] = ((SDFMap*)h)->max_vec[i];
}
void gto_map_set_occupancy(void* h, const double* xyz, int64_t n) {
  SDFMap* m = (SDFMap*)h;
  for (int64_t i = 0; i < n; ++i) m->setOccupancy(xyz + 3 * i, 1);
}
void gto_map_update_esdf(void* h) { ((SDFMap*)h)->updateESDF3d(); }
void gto_map_get_distance(void* h, double* out) {
  SDFMap* m = (SDFMap*)h;
  std::memcpy(out, m->distance_buffer.data(), m->distance_buffer.size() * sizeof(double));
}
void gto_map_set_distance(void* h, const double* in) {
  SDFMap* m = (SDFMap*)h;
  std::memcpy(m->distance_buffer.data(), in, m->distance_buffer.size() * sizeof(double));
}
void gto_map_get_occupancy(void* h, int32_t* out) {
  SDFMap* m = (SDFMap*)h;
  std::memcpy(out, m->occupancy_buffer.data(), m->occupancy_buffer.size() * sizeof(int));
}
void gto_map_set_occupancy_buffer(void* h, const int32_t* in) {
  SDFMap* m = (SDFMap*)h;
  std::memcpy(m->occupancy_buffer.data(), in, m->occupancy_buffer.size() * sizeof(int));
}
void gto_map_brute_edt(void* h, double* out) {
  std::vector<double> v;
  ((SDFMap*)h)->bruteForceEDT(v);
  std::memcpy(out, v.data(), v.size() * sizeof(double));
}
// trilinear query; returns number of out-of-map points
int64_t gto_map_query(void* h, const double* pos, int64_t n, double* dist, double* grad) {
  SDFMap* m = (SDFMap*)h;
  int64_t oob = 0;
  for (int64_t i = 0; i < n; ++i) {
    bool in = true;
    dist[i] = m->getDistWithGradTrilinear(pos + 3 * i, grad + 3 * i, &in);
    if (!in) ++oob;
  }
  return oob;
}

// ---- optimizer -----------------------------------------------------------------------------
void* gto_opt_create(const Params* p) { return new GradTrajOptimizer(*p); }
void gto_opt_destroy(void* h) { delete (GradTrajOptimizer*)h; }
void gto_opt_set_map(void* h, void* map) { ((GradTrajOptimizer*)h)->setMap((SDFMap*)map); }
void gto_opt_set_params(void* h, const Params* p) { ((GradTrajOptimizer*)h)->prm = *p; }
int gto_opt_set_path(void* h, const double* wp, int npts) {
  GTO_TRY((GradTrajOptimizer*)h)->setPath(std::vector<double>(wp, wp + 3 * (size_t)npts));
  return 0;
  GTO_CATCH(-1)
}
// Pos/Vel/Acc: (m+1) x 3 row-major; Time: m
int gto_opt_set_kino_path(void* h, const double* pos, const double* vel, const double* acc, const double* time, int m) {
  GTO_TRY Mat P(m + 1, 3), V(m + 1, 3), A(m + 1, 3);
  for (int i = 0; i <= m; ++i)
    for (int a = 0; a < 3; ++a) P(i, a) = pos[3 * i + a], V(i, a) = vel[3 * i + a], A(i, a) = acc[3 * i + a];
  ((GradTrajOptimizer*)h)->setKinoPath(P, V, A, std::vector<double>(time, time + m));
  return 0;
  GTO_CATCH(-1)
}
void gto_opt_dims(void* h, int* m, int* num_dp) {
  GradTrajOptimizer* g = (GradTrajOptimizer*)h;
  *m = (int)g->segment_time.size();
  *num_dp = g->num_dp;
}
static void copyMat(const Mat& M, double* out) { std::memcpy(out, M.a.data(), M.a.size() * sizeof(double)); }
void gto_opt_get_L(void* h, double* out) { copyMat(((GradTrajOptimizer*)h)->L, out); }      // 6m x (3m+3) col-major
void gto_opt_get_R(void* h, double* out) { copyMat(((GradTrajOptimizer*)h)->R, out); }      // (3m+3)^2 col-major
void gto_opt_get_A(void* h, double* out) { copyMat(((GradTrajOptimizer*)h)->A, out); }      // (6m)^2 col-major
void gto_opt_get_C(void* h, double* out) { copyMat(((GradTrajOptimizer*)h)->C, out); }      // (3m+3) x 6m col-major
void gto_opt_get_Df(void* h, double* out) {  // 3 x 6 ROW-major (axis-major) for convenience
  const Mat& Df = ((GradTrajOptimizer*)h)->Df;
  for (int a = 0; a < 3; ++a)
    for (int j = 0; j < 6; ++j) out[6 * a + j] = Df(a, j);
}
void gto_opt_get_seg_time(void* h, double* out) {
  GradTrajOptimizer* g = (GradTrajOptimizer*)h;
  std::memcpy(out, g->segment_time.data(), g->segment_time.size() * sizeof(double));
}
void gto_opt_get_x0(void* h, double* out) {
  std::vector<double> x = ((GradTrajOptimizer*)h)->x0();
  std::memcpy(out, x.data(), x.size() * sizeof(double));
}
void gto_opt_get_bounds(void* h, double* lb, double* ub) {
  std::vector<double> l, u;
  ((GradTrajOptimizer*)h)->bounds(l, u);
  std::memcpy(lb, l.data(), l.size() * sizeof(double));
  std::memcpy(ub, u.data(), u.size() * sizeof(double));
}
void gto_opt_get_coeff(void* h, double* out) {  // m x 18 ROW-major
  const Mat& c = ((GradTrajOptimizer*)h)->coeff;
  for (int i = 0; i < c.rows; ++i)
    for (int j = 0; j < 18; ++j) out[18 * i + j] = c(i, j);
}
void gto_opt_coeff_from_x(void* h, const double* x, double* out) {  // m x 18 ROW-major
  GradTrajOptimizer* g = (GradTrajOptimizer*)h;
  Mat c;
  g->getCoefficientFromDerivative(c, std::vector<double>(x, x + 3 * (size_t)g->num_dp));
  for (int i = 0; i < c.rows; ++i)
    for (int j = 0; j < 18; ++j) out[18 * i + j] = c(i, j);
}

// single evaluation; terms (nullable) = [smooth, collision, vel, acc] unweighted
double gto_opt_eval(void* h, const double* x, double* grad, int step, int64_t* oob, double* terms) {
  GradTrajOptimizer* g = (GradTrajOptimizer*)h;
  g->step = step;
  std::vector<double> xv(x, x + 3 * (size_t)g->num_dp), gv;
  double cost;
  g->getCostAndGradient(xv, cost, gv, oob, terms);
  if (grad) std::memcpy(grad, gv.data(), gv.size() * sizeof(double));
  return cost;
}

// batch over B start points of ONE problem, nthreads host threads (0 = hardware_concurrency).
// Returns wall seconds of the evaluation loop.
double gto_opt_eval_batch(void* h, const double* X, int64_t B, double* cost, double* grad, int step, int nthreads,
                          int64_t* oob_total) {
  GradTrajOptimizer* g = (GradTrajOptimizer*)h;
  g->step = step;
  const size_t n = 3 * (size_t)g->num_dp;
  if (nthreads <= 0) nthreads = (int)std::thread::hardware_concurrency();
  if (nthreads <= 0) nthreads = 1;
  std::atomic<int64_t> next(0), oob_sum(0);
  auto t0 = std::chrono::steady_clock::now();
  auto work = [&]() {
    std::vector<double> xv(n), gv;
    int64_t oob = 0;
    for (;;) {
      int64_t b = next.fetch_add(1);
      if (b >= B) break;
      std::memcpy(xv.data(), X + b * n, n * sizeof(double));
      double c;
      g->getCostAndGradient(xv, c, gv, &oob, nullptr);
      cost[b] = c;
      if (grad) std::memcpy(grad + b * n, gv.data(), n * sizeof(double));
    }
    oob_sum += oob;
  };
  std::vector<std::thread> th;
  for (int i = 1; i < nthreads; ++i) th.emplace_back(work);
  work();
  for (auto& t : th) t.join();
  auto t1 = std::chrono::steady_clock::now();
  if (oob_total) *oob_total = oob_sum.load();
  return std::chrono::duration<double>(t1 - t0).count();
}

// NLopt solve (vendored NLopt 2.5.0 binary at nlopt_path).  f_override (nullable) replaces the
// oracle objective with the caller's (used to drive the GPU objective through the same solver).
int gto_opt_solve(void* h, int step, int maxeval, double maxtime, double xtol_abs, const char* nlopt_path,
                  nlopt_func_t f_override, void* f_data, double* x_out, double* min_f, int* n_evals) {
  GTO_TRY GradTrajOptimizer* g = (GradTrajOptimizer*)h;
  SolveOptions so;
  so.maxeval = maxeval, so.maxtime = maxtime, so.xtol_abs = xtol_abs;
  g->iter_num = 0;
  int r = g->optimizeTrajectory(step, so, nlopt_path, f_override, f_data);
  std::vector<double> x = g->x0();
  if (x_out) std::memcpy(x_out, x.data(), x.size() * sizeof(double));
  if (min_f) *min_f = g->last_min_f;
  if (n_evals) *n_evals = (int)g->vec_cost.size();
  return r;
  GTO_CATCH(-1002)
}
int64_t gto_opt_cost_curve(void* h, double* out, int64_t cap) {
  GradTrajOptimizer* g = (GradTrajOptimizer*)h;
  int64_t n = (int64_t)g->vec_cost.size();
  for (int64_t i = 0; i < n && i < cap; ++i) out[i] = g->vec_cost[i];
  return n;
}

// PolynomialTraj as opti_node builds it (src/opti_node.cpp:108-144): coeff = getCoefficient (m x 18 row-major, ascending
// powers per axis), reversed into descending order per segment.  out[8] = time_sum, length, jerk, mean_v, max_v, mean_a,
// max_a, acc_cost; traj (optional, cap triples) receives getTraj(); returns the number of trajectory points.
int64_t gto_traj_metrics(const double* coeff, const double* seg_time, int m, double* out, double* traj, int64_t cap) {
  PolynomialTraj pt;
  pt.reset();
  for (int s = 0; s < m; ++s) {
    std::vector<double> c[3];
    for (int a = 0; a < 3; ++a) {
      c[a].assign(coeff + (size_t)s * 18 + a * 6, coeff + (size_t)s * 18 + a * 6 + 6);
      std::reverse(c[a].begin(), c[a].end());
    }
    pt.addSegment(c[0], c[1], c[2], seg_time[s]);
  }
  pt.init();
  const std::vector<double>& tr = pt.getTraj();
  const int64_t npts = (int64_t)tr.size() / 3;
  if (traj)
    for (int64_t i = 0; i < std::min<int64_t>(npts, cap) * 3; ++i) traj[i] = tr[i];
  out[0] = pt.getTimeSum();
  out[1] = pt.getLength();
  out[2] = pt.getJerk();
  pt.getMeanAndMaxVel(out[3], out[4]);
  pt.getMeanAndMaxAcc(out[5], out[6]);
  out[7] = pt.getAccCost();
  return npts;
}

// EDTEnvironment::evaluateEDTWithGrad / evaluateCoarseEDT for n query points; boxes: nbox x 9 (c0[3], c1[3], scale[3])
void gto_edt_env_query(void* map, const double* boxes, int nbox, const double* pos, int64_t n, double time, double* dist,
                       double* grad, double* coarse) {
  EDTEnvironment env;
  env.setMap((SDFMap*)map);
  std::vector<MovingBox> bx(nbox);
  for (int i = 0; i < nbox; ++i)
    for (int k = 0; k < 3; ++k) bx[i].c0[k] = boxes[9 * i + k], bx[i].c1[k] = boxes[9 * i + 3 + k], bx[i].scale[k] = boxes[9 * i + 6 + k];
  env.setObjects(bx);
  for (int64_t i = 0; i < n; ++i) {
    env.evaluateEDTWithGrad(pos + 3 * i, time, dist[i], grad + 3 * i);
    if (coarse) coarse[i] = env.evaluateCoarseEDT(pos + 3 * i, time);
  }
}

int gto_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
