// setup_core.cuh — the arithmetic of one GTOP problem's set-up, shared by the host path (problem_setup.cpp) and the
// device path (setup_kernels.cu) so that both run the same operations in the same order.
//
// Restates, segment by segment, what GradTrajOptimizer::setPath / setKinoPath leave in the optimizer's members
// (src/grad_traj_optimizer.cpp:35-110) through TrajectoryGenerator (src/qp_generator.cpp:156-451):
//   time allocation :73-81, A_s rows :185-195, min-jerk Q_s :226-234, M_s = A_s^-T Q_s A_s^-1 and its scatter into
//   R (StackOptiDep :357-405), initial derivatives (getInitialD :407-451), NLopt box (:151-177).
// The only operation that is NOT shared is pow(): the host calls libm's (as the reference does), the device a
// correctly rounded integer power (setup_pow below); they agree except where libm's pow is not correctly rounded.
#pragma once
#include <cmath>

#include "gtop_types.cuh"

namespace gtop {

// t^q for q = 0..5.  Host: std::pow, the call the reference makes (grad_traj_optimizer.cpp:457-485, :549;
// qp_generator.cpp:192, :232).  Device: product in double-double arithmetic (error-free products through fma),
// rounded once — the correctly rounded power up to a 2^-100 relative error before the final rounding.
GTOP_HD inline double setup_pow(double t, int q) {
#if defined(__CUDA_ARCH__)
  if (q == 0) return 1.0;
  double h = t, l = 0.0;
  for (int i = 2; i <= q; ++i) {
    const double p = h * t;
    double e = fma(h, t, -p);
    e = fma(l, t, e);
    const double h2 = p + e;
    l = e - (h2 - p);
    h = h2;
  }
  return h;
#else
  return std::pow(t, q);
#endif
}

GTOP_HD inline int setup_factorial(int x) {
  int f = 1;
  for (int i = 2; i <= x; ++i) f *= i;
  return f;
}

// 6x6 inverse by partial-pivot LU (row-major a[r][c]).  Returns false when singular.
GTOP_HD inline bool setup_invert6(const double A[36], double Inv[36]) {
  double lu[36];
  int perm[6];
  for (int i = 0; i < 36; ++i) lu[i] = A[i];
  for (int i = 0; i < 6; ++i) perm[i] = i;
  for (int k = 0; k < 6; ++k) {
    int piv = k;
    double best = fabs(lu[k * 6 + k]);
    for (int i = k + 1; i < 6; ++i)
      if (fabs(lu[i * 6 + k]) > best) best = fabs(lu[i * 6 + k]), piv = i;
    if (best == 0.0) return false;
    if (piv != k) {
      for (int j = 0; j < 6; ++j) {
        const double t = lu[k * 6 + j];
        lu[k * 6 + j] = lu[piv * 6 + j], lu[piv * 6 + j] = t;
      }
      const int t = perm[k];
      perm[k] = perm[piv], perm[piv] = t;
    }
    for (int i = k + 1; i < 6; ++i) {
      if (lu[i * 6 + k] == 0.0) continue;
      const double l = (lu[i * 6 + k] /= lu[k * 6 + k]);
      for (int j = k + 1; j < 6; ++j) lu[i * 6 + j] -= l * lu[k * 6 + j];
    }
  }
  for (int c = 0; c < 6; ++c) {
    double y[6];
    for (int i = 0; i < 6; ++i) {
      double s = perm[i] == c ? 1.0 : 0.0;
      for (int j = 0; j < i; ++j) s -= lu[i * 6 + j] * y[j];
      y[i] = s;
    }
    for (int i = 5; i >= 0; --i) {
      double s = y[i];
      for (int j = i + 1; j < 6; ++j) s -= lu[i * 6 + j] * Inv[j * 6 + c];
      Inv[i * 6 + c] = s / lu[i * 6 + i];
    }
  }
  return true;
}

// reference local row order is [p0,pT,v0,vT,a0,aT] (qp_generator.cpp:187-192); natural order used
// by the kernels is [p0,v0,a0,pT,vT,aT].
GTOP_HD inline int setup_ref_row_of_natural(int r) { return r < 3 ? 2 * r : 2 * (r - 3) + 1; }

// time allocation of segment i (grad_traj_optimizer.cpp:73-81; init_time on the first segment only, H9)
GTOP_HD inline double setup_segment_time(const double* wp, int i, double mean_v, double init_time) {
  const double dx = wp[3 * i] - wp[3 * i + 3], dy = wp[3 * i + 1] - wp[3 * i + 4], dz = wp[3 * i + 2] - wp[3 * i + 5];
  const double len = sqrt(dx * dx + dy * dy + dz * dz);
  return (i == 0) ? len / mean_v + init_time : len / mean_v;
}

// per-segment blocks from T alone: Li[i][r] = A_s^-1 (row = power, column = local derivative in natural order),
// M = A_s^-T Q_s A_s^-1 in natural order.  Returns false when A_s is singular.
GTOP_HD inline bool setup_segment_blocks(double T, double* Li /* 36 */, double* M /* 36 */) {
  double A[36], Ainv[36], Q[36];
  for (int i = 0; i < 36; ++i) A[i] = 0.0, Q[i] = 0.0;
  for (int i = 0; i < 3; ++i) {  // qp_generator.cpp:187-192
    A[(2 * i) * 6 + i] = setup_factorial(i);
    for (int j = i; j < 6; ++j) A[(2 * i + 1) * 6 + j] = setup_factorial(j) / setup_factorial(j - i) * setup_pow(T, j - i);
  }
  for (int i = 3; i < 6; ++i)  // :227-233 (integer prefactor, exact)
    for (int j = 3; j < 6; ++j)
      Q[i * 6 + j] = i * (i - 1) * (i - 2) * j * (j - 1) * (j - 2) / (i + j - 5) * setup_pow(T, i + j - 5);
  if (!setup_invert6(A, Ainv)) return false;
  for (int i = 0; i < 6; ++i)
    for (int r = 0; r < 6; ++r) Li[i * 6 + r] = Ainv[i * 6 + setup_ref_row_of_natural(r)];
  // M = (Ainv^T Q) Ainv in natural order
  double T1[36];
  for (int r = 0; r < 6; ++r)
    for (int j = 0; j < 6; ++j) {
      double acc = 0.0;
      for (int i = 0; i < 6; ++i) acc += Li[i * 6 + r] * Q[i * 6 + j];
      T1[r * 6 + j] = acc;
    }
  for (int r = 0; r < 6; ++r)
    for (int c = 0; c < 6; ++c) {
      double acc = 0.0;
      for (int j = 0; j < 6; ++j) acc += T1[r * 6 + j] * Li[j * 6 + c];
      M[r * 6 + c] = acc;
    }
  return true;
}

// R(u, v) in waypoint order = sum of the segment forms scattered on (waypoint s, waypoint s+1)
GTOP_HD inline double setup_R_entry(const double* Mseg, int m, int u, int v) {
  const int wu = u / 3, wv = v / 3, qu = u % 3, qv = v % 3;
  double r = 0.0;
  if (wu == wv) {
    if (wu >= 1) r += Mseg[(size_t)(wu - 1) * 36 + (3 + qu) * 6 + 3 + qv];
    if (wu <= m - 1) r += Mseg[(size_t)wu * 36 + qu * 6 + qv];
  } else if (wv == wu + 1) {
    r = Mseg[(size_t)wu * 36 + qu * 6 + 3 + qv];
  } else if (wv == wu - 1) {
    r = Mseg[(size_t)wv * 36 + (3 + qu) * 6 + qv];
  }
  return r;
}
// entry j of row v of the column band Rc [3m+3][9] (smoothness cost)
GTOP_HD inline double setup_Rc_entry(const double* Mseg, int m, int v, int j) {
  const int u = 3 * (v / 3 - 1) + j;
  return (u >= 0 && u < 3 * m + 3) ? setup_R_entry(Mseg, m, u, v) : 0.0;
}
// entry j of row i of the gradient band Rg [3m-3][9] (2 Rfp' df + 2 Rpp dp)
GTOP_HD inline double setup_Rg_entry(const double* Mseg, int m, int i, int j) {
  const int w = i / 3 + 1, self = 3 * w + i % 3;
  const int nb = 3 * (w - 1) + j, wpn = nb / 3;
  const bool fixed = (wpn == 0 || wpn == m);
  return fixed ? setup_R_entry(Mseg, m, nb, self) : setup_R_entry(Mseg, m, self, nb);
}

// x0 / NLopt box of variable (axis a, i = 3(w-1)+q): getInitialD (qp_generator.cpp:407-451, straight-line init
// :204-221 or the front end's derivatives) and grad_traj_optimizer.cpp:151-177
GTOP_HD inline void setup_x0_box(const double* pos, const double* vel, const double* acc, int a, int i, double bos,
                                 double vos, double aos, double* x0, double* lb, double* ub) {
  const int w = i / 3 + 1, q = i % 3;
  const double c = pos[3 * w + a];
  *x0 = (q == 0) ? c : (q == 1 ? (vel ? vel[3 * w + a] : 0.0) : (acc ? acc[3 * w + a] : 0.0));
  const double half = q == 0 ? bos : (q == 1 ? vos : aos);
  *lb = (q == 0 ? c : 0.0) - half;  // box centred on the path / on zero (:154-177)
  *ub = (q == 0 ? c : 0.0) + half;
}

// Pair schedule of k_eval_mma (gtop_types.cuh): where sample k of segment s sits.  ns_first = sample count of the
// pair's first segment.  Returns the TQ / PL iteration u and slot j; *pl_record = the PL record (u, or nit2 for the
// second segment's lanes in the iteration that mixes both segments).
GTOP_HD inline void setup_pair_slot(int s, int k, int ns_first, int nit2, int* u, int* j, int* pl_record) {
  const int second = s & 1;
  const int q = second ? ns_first + k : k;
  *u = q >> 2, *j = q & 3;
  const bool mixed_iter = (ns_first & 3) != 0 && *u == (ns_first >> 2);
  *pl_record = (second && mixed_iter) ? nit2 : *u;
}

// B fragments of the gradient DMMAs for one sample and one local derivative r (gtop_types.cuh, PL):
// (dt * phi_r(t), dt * phi_r'(t)), phi_r(t) = sum_i t^i Linv[i][r]; pw[i] = t^i
GTOP_HD inline void setup_basis_pair(const double pw[6], const double* Li /* 36 */, int r, double dt, double* a_out,
                                     double* b_out) {
  double a = 0.0, b = 0.0;
  for (int i = 0; i < 6; ++i) a += pw[i] * Li[i * 6 + r];
  for (int i = 1; i < 6; ++i) b += (double)i * pw[i - 1] * Li[i * 6 + r];
  *a_out = dt * a, *b_out = dt * b;
}

}  // namespace gtop
