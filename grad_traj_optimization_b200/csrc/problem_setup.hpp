// problem_setup.hpp — host-side set-up of one GTOP trajectory problem for the B200 kernels.
//
// Two ways in, one record out (ProblemHost -> packed device record, gtop_types.cuh):
//   * from_dense : the drop-in path.  Takes the dense L, R, Df, segment_time that the reference's
//     GradTrajOptimizer::setPath leaves in its members (src/grad_traj_optimizer.cpp:88-102) and
//     extracts the per-segment 6x6 blocks of L and the block-tridiagonal band of R — the only
//     non-zeros those matrices have (SURVEY §10 structure checks).
//   * from_path  : the native path.  Restates setPath (time allocation :73-81) and builds the same
//     blocks segment by segment from T_s alone — A_s (src/qp_generator.cpp:185-195), the min-jerk
//     Q_s (:226-234), M_s = A_s^-T Q_s A_s^-1 — in O(m), never forming the (6m)^2 matrices the
//     reference inverts three times (:390-393).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/gtop_c.h"

namespace gtop {

struct ProblemHost {
  int m = 0;
  int ns_cap = 30;
  std::vector<double> seg_time;  // m
  std::vector<double> Linv;      // m*36, [s][i][r] r in natural order (p0,v0,a0,pT,vT,aT)
  std::vector<double> Rc;        // (3m+3)*9 column band, waypoint order
  std::vector<double> Rg;        // (3m-3)*9 gradient band
  std::vector<double> Df;        // 18, [axis][p0,v0,a0,pEnd,vEnd,aEnd]
  std::vector<double> dt, ns;    // m each
  std::vector<double> tt;        // m*5*ns_cap
  bool has_path = false;
  std::vector<double> x0, lb, ub;      // n each (native path only)
  std::vector<double> Mseg;            // m*36 (native path only): M_s in natural order, for dense_R
};

// dense column index (in d = [df(6); dp(P)]) of waypoint-order entry (wp, q)
inline int dense_col(int m, int wp, int q) { return wp == 0 ? q : (wp == m ? 3 + q : 6 + 3 * (wp - 1) + q); }

// returns empty string on success, else an error message
std::string setup_from_dense(int m, const double* seg_time, const double* L, const double* R, const double* Df_colmajor,
                             const gtop_params& prm, ProblemHost& out);
std::string setup_from_path(int m, const double* waypoints /* (m+1) x 3 */, const gtop_params& prm, ProblemHost& out);
std::string setup_from_kino(int m, const double* pos, const double* vel, const double* acc /* (m+1) x 3 each */,
                            const double* time /* m */, const gtop_params& prm, ProblemHost& out);

// sample-time table of one problem (dt, ns, tt), from seg_time and samples_per_segment
void build_time_table(ProblemHost& p);

// dense views assembled from the blocks (API parity with getL()/getR()), column-major
void dense_L(const ProblemHost& p, double* L /* 6m x (3m+3) */);
void dense_R(const ProblemHost& p, double* R /* (3m+3)^2 */);

// pack into the device record layout
void pack_record(const ProblemHost& p, double* rec /* layout.stride doubles */);

// coefficients of one trajectory (getCoefficientFromDerivative), m x 18 row-major
void coeff_from_x(const ProblemHost& p, const double* x, double* coeff);

}  // namespace gtop
