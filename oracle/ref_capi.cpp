// ref_capi.cpp — extern "C" surface over the REFERENCE'S OWN hot-path sources, compiled unmodified from
// /root/reference/src/{sdf_map,qp_generator,grad_traj_optimizer}.cpp against the stand-in headers of oracle/shim/
// (Eigen / roscpp / nlopt.hpp are not in this image) into oracle/_ref/libgtop_reference.so by oracle/Makefile.
//
// TEST INFRASTRUCTURE ONLY.  It exists to pin the oracle restatement (oracle/gtop_oracle.cpp) to the reference itself:
// tests/test_reference_build.py compares the two on the reference's fixtures, and tests/golden/make_golden.py stores
// outputs of THIS library so that the pins travel to the GPU box (which has no /root/reference).  Nothing in the
// product may load it.
//
// The reference keeps the objective private (grad_traj_optimizer.h:83-91); this one translation unit includes the header
// with `private` spelled `public` to reach getCostAndGradient / costFunc — the reference's own translation units are
// compiled as they are.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#define private public
#include "grad_traj_optimization/grad_traj_optimizer.h"
#undef private

static thread_local std::string g_err;
#define REF_TRY try {
#define REF_CATCH(ret)              \
  }                                 \
  catch (const std::exception& e) { \
    g_err = e.what();               \
    return ret;                     \
  }

static Eigen::Vector3d v3(const double* p) { return Eigen::Vector3d(p[0], p[1], p[2]); }
static void dump(const Eigen::MatrixXd& m, double* out) { std::memcpy(out, m.data(), sizeof(double) * (size_t)m.size()); }

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

// ---- SDFMap (include/grad_traj_optimization/sdf_map.h:34-56) ------------------------------------------------------
void* ref_map_create(const double origin[3], double res, const double size[3]) {
  REF_TRY return new SDFMap(v3(origin), res, v3(size));
  REF_CATCH(nullptr)
}
void ref_map_destroy(void* h) { delete (SDFMap*)h; }
void ref_map_grid(void* h, int out[3]) {
  for (int i = 0; i < 3; ++i) out[i] = ((SDFMap*)h)->grid_size(i);
}
void ref_map_reset(void* h) { ((SDFMap*)h)->resetBuffer(); }
void ref_map_reset_range(void* h, const double* lo, const double* hi) { ((SDFMap*)h)->resetBuffer(v3(lo), v3(hi)); }
void ref_map_set_update_range(void* h, const double* lo, const double* hi) { ((SDFMap*)h)->setUpdateRange(v3(lo), v3(hi)); }
void ref_map_update_range(void* h, int out[6]) {
  for (int i = 0; i < 3; ++i) out[i] = ((SDFMap*)h)->min_vec(i), out[3 + i] = ((SDFMap*)h)->max_vec(i);
}
void ref_map_set_occupancy(void* h, const double* xyz, int64_t n) {
  SDFMap* m = (SDFMap*)h;
  for (int64_t i = 0; i < n; ++i) m->setOccupancy(v3(xyz + 3 * i), 1);
}
void ref_map_update_esdf(void* h) { ((SDFMap*)h)->updateESDF3d(); }
void ref_map_get_distance(void* h, double* out) {
  SDFMap* m = (SDFMap*)h;
  std::memcpy(out, m->distance_buffer.data(), m->distance_buffer.size() * sizeof(double));
}
void ref_map_get_occupancy(void* h, int32_t* out) {
  SDFMap* m = (SDFMap*)h;
  for (size_t i = 0; i < m->occupancy_buffer.size(); ++i) out[i] = m->occupancy_buffer[i];
}
// getDistWithGradTrilinear (sdf_map.cpp:185-242); grad is zero-initialised by the shim's Vector3d (H3)
void ref_map_query(void* h, const double* pos, int64_t n, double* dist, double* grad) {
  SDFMap* m = (SDFMap*)h;
  for (int64_t i = 0; i < n; ++i) {
    Eigen::Vector3d g;
    dist[i] = m->getDistWithGradTrilinear(v3(pos + 3 * i), g);
    grad[3 * i] = g(0), grad[3 * i + 1] = g(1), grad[3 * i + 2] = g(2);
  }
}
double ref_map_max_distance(void* h) { return ((SDFMap*)h)->getMaxDistance(); }

// ---- GradTrajOptimizer (include/grad_traj_optimization/grad_traj_optimizer.h) --------------------------------------
// params: the 21 values the constructor reads from the parameter server (grad_traj_optimizer.cpp:5-32), in the
// order alg, time_limit_1, time_limit_2, dt, ws, wc, alpha, r, d0, alpha_v, r_v, v0, alpha_a, r_a, a0, bos, vos,
// aos, mean_v, mean_a, init_time
void* ref_opt_create(const double params[21]) {
  static const char* names[21] = {"alg", "time_limit_1", "time_limit_2", "dt", "ws", "wc", "alpha", "r", "d0",
                                  "alpha_v", "r_v", "v0", "alpha_a", "r_a", "a0", "bos", "vos", "aos",
                                  "mean_v", "mean_a", "init_time"};
  REF_TRY
  for (int i = 0; i < 21; ++i) ros::param::table()[std::string("/traj_opti_node1/") + names[i]] = params[i];
  return new GradTrajOptimizer();
  REF_CATCH(nullptr)
}
void ref_opt_destroy(void* h) { delete (GradTrajOptimizer*)h; }
int ref_opt_init_map(void* h, const double size[3], const double origin[3], double res) {
  REF_TRY((GradTrajOptimizer*)h)->initSDFMap(v3(size), v3(origin), res);
  return 0;
  REF_CATCH(-1)
}
int ref_opt_update_map(void* h, const double* xyz, int64_t n) {
  REF_TRY std::vector<Eigen::Vector3d> obs;
  obs.reserve((size_t)n);
  for (int64_t i = 0; i < n; ++i) obs.push_back(v3(xyz + 3 * i));
  ((GradTrajOptimizer*)h)->updateSDFMap(obs);
  return 0;
  REF_CATCH(-1)
}
// the field of the optimizer's own map / replace it (synthetic maps are built once and shared)
void ref_opt_get_distance(void* h, double* out) { ref_map_get_distance(&((GradTrajOptimizer*)h)->sdf_map, out); }
void ref_opt_set_distance(void* h, const double* in) {
  SDFMap& m = ((GradTrajOptimizer*)h)->sdf_map;
  std::memcpy(m.distance_buffer.data(), in, m.distance_buffer.size() * sizeof(double));
}
int ref_opt_set_path(void* h, const double* wp, int npts) {
  REF_TRY std::vector<Eigen::Vector3d> pts;
  for (int i = 0; i < npts; ++i) pts.push_back(v3(wp + 3 * i));
  ((GradTrajOptimizer*)h)->setPath(pts);
  return 0;
  REF_CATCH(-1)
}
// Pos / Vel / Acc: (m+1) x 3 row-major, Time: m (setKinoPath, grad_traj_optimizer.cpp:35-65).  qp_generator.cpp keeps
// startVel / startAcc in file-scope globals that PolyKinoGeneration never sets (H14): they are whatever the last
// PolyQPGeneration of the process left (zero after setPath, zero-initialised before).
int ref_opt_set_kino_path(void* h, const double* pos, const double* vel, const double* acc, const double* time, int m) {
  REF_TRY Eigen::MatrixXd P(m + 1, 3), V(m + 1, 3), A(m + 1, 3);
  Eigen::VectorXd T(m);
  for (int i = 0; i <= m; ++i)
    for (int a = 0; a < 3; ++a) P(i, a) = pos[3 * i + a], V(i, a) = vel[3 * i + a], A(i, a) = acc[3 * i + a];
  for (int i = 0; i < m; ++i) T(i) = time[i];
  ((GradTrajOptimizer*)h)->setKinoPath(P, V, A, T);
  return 0;
  REF_CATCH(-1)
}
void ref_opt_dims(void* h, int out[3]) {
  GradTrajOptimizer* o = (GradTrajOptimizer*)h;
  out[0] = o->num_point - 1, out[1] = o->num_dp, out[2] = o->num_df;
}
void ref_opt_seg_time(void* h, double* out) { dump(((GradTrajOptimizer*)h)->segment_time, out); }
// column-major, as Eigen stores them: L 6m x (3m+3), R (3m+3)^2, Df 3 x 6, Dp 3 x num_dp, coeff m x 18
void ref_opt_L(void* h, double* out) { dump(((GradTrajOptimizer*)h)->L, out); }
void ref_opt_R(void* h, double* out) { dump(((GradTrajOptimizer*)h)->R, out); }
void ref_opt_Df(void* h, double* out) { dump(((GradTrajOptimizer*)h)->Df, out); }
void ref_opt_Dp(void* h, double* out) { dump(((GradTrajOptimizer*)h)->Dp, out); }
void ref_opt_coeff(void* h, double* out) { dump(((GradTrajOptimizer*)h)->coeff, out); }
// x0 in NLopt's variable order (grad_traj_optimizer.cpp:182-187)
void ref_opt_x0(void* h, double* out) {
  GradTrajOptimizer* o = (GradTrajOptimizer*)h;
  for (int i = 0; i < o->num_dp; ++i)
    for (int a = 0; a < 3; ++a) out[i + a * o->num_dp] = o->Dp(a, i);
}
// one call of the NLopt callback (costFunc, :554-562) with the given step rule (:412-415)
int ref_opt_eval(void* h, int step, const double* x, int64_t B, double* cost, double* grad) {
  REF_TRY GradTrajOptimizer* o = (GradTrajOptimizer*)h;
  const int n = 3 * o->num_dp;
  o->step = step;
  std::vector<double> xv(n), gv;
  for (int64_t b = 0; b < B; ++b) {
    xv.assign(x + b * n, x + (b + 1) * n);
    cost[b] = GradTrajOptimizer::costFunc(xv, gv, o);
    std::memcpy(grad + b * n, gv.data(), sizeof(double) * n);
  }
  return 0;
  REF_CATCH(-1)
}
void ref_opt_coeff_from_x(void* h, const double* x, double* out) {
  GradTrajOptimizer* o = (GradTrajOptimizer*)h;
  std::vector<double> xv(x, x + 3 * o->num_dp);
  Eigen::MatrixXd c;
  o->getCoefficientFromDerivative(c, xv);
  dump(c, out);
}
// optimizeTrajectory(step) (:128-243) with `maxeval` evaluations as the stopping rule (H7); the result is the
// optimised Dp in NLopt's variable order.  evals <- number of objective calls made.
int ref_opt_optimize(void* h, int step, int maxeval, double* x_out, int* evals) {
  REF_TRY GradTrajOptimizer* o = (GradTrajOptimizer*)h;
  nlopt::shim::max_evaluations() = maxeval;
  const int before = o->iter_num;
  o->optimizeTrajectory(step);
  if (evals) *evals = (step == 2 ? (int)o->vec_cost.size() : o->iter_num - before);
  ref_opt_x0(h, x_out);
  return 0;
  REF_CATCH(-1)
}

}  // extern "C"
