// mma_core.cuh — the optimiser loop that GradTrajOptimizer::optimizeTrajectory runs
// (src/grad_traj_optimizer.cpp:135-195 of the reference: nlopt::opt(nlopt::algorithm(24) = LD_MMA, n),
// set_min_objective, box bounds, optimize), restated for a BATCH of independent trajectories in lock-step.
//
// The algorithm lives in a third-party dependency that is not under /root/reference as source: NLopt 2.5.0
// (the reference vendors the binary lib/libnlopt.so.0.9.0), algorithm LD_MMA = Svanberg's conservative convex
// separable approximations as implemented in NLopt's mma.c.  Restated from the published algorithm
// (K. Svanberg, SIAM J. Optim. 12 (2002) 555-573; NLopt documentation "MMA") for the case this path uses —
// box bounds, NO nonlinear constraints — where the dual problem is empty and every variable has a closed-form
// update:
//
//   state   x (best point so far, approximation point), dfdx = grad f(x), fval = f(x),
//           sigma_j (asymptote distance, starts at (ub_j - lb_j)/2), rho (conservativeness, starts at 1),
//           xcur (last evaluated point), xprev / xprevprev (last points of the two previous outer iterations)
//   candidate (inner iteration):
//           u = dfdx_j sigma_j^2,  v = |dfdx_j| sigma_j + rho/2,
//           dx = (u/v) / (-1 - sqrt|1 - (u/(v sigma_j))^2|),  xcur_j = clip(x_j + dx, [lb,ub], x_j +- 0.9 sigma_j)
//           g  = fval + sum_j (dfdx_j sigma_j^2 dx + (|dfdx_j| sigma_j + rho/2) dx^2) / (sigma_j^2 - dx^2)
//           w  = sum_j dx^2 / (2 (sigma_j^2 - dx^2))
//   after evaluating fcur = f(xcur):
//           fcur < fval              -> x, dfdx, fval := xcur, grad f(xcur), fcur   (also inside inner iterations)
//           g >= fcur (conservative) -> outer step: rho := max(rho/10, 1e-5); from the second outer iteration on
//                                       sigma_j *= 0.7 / 1.2 / 1 by the sign of (xcur-xprev)_j (xprev-xprevprev)_j,
//                                       clamped to [0.01, 10] (ub_j - lb_j); xprevprev := xprev, xprev := xcur
//           else                     -> rho := min(10 rho, 1.1 (rho + (fcur - g)/w))
//   stop    after `maxeval` objective evaluations (the first one at x0 included), as the harness does
//           (the reference's only rule is wall clock — SURVEY H7).
//
// Every trajectory costs exactly one objective evaluation per step whatever branch it takes, so a batch advances
// in lock-step: [candidate for all] -> [ONE batched k_eval launch] -> [bookkeeping for all].  Parity is anchored on
// the reference's call site: tests/test_mma.py drives this code and the vendored NLopt binary with the same
// objective from the same start and compares the iterates.
#pragma once
#include <cmath>

#if defined(__CUDACC__)
#define GTOP_MMA_HD __host__ __device__ __forceinline__
#else
#define GTOP_MMA_HD inline
#endif

namespace gtop {

constexpr double kMmaRhoMin = 1e-5;

// Scalar state of one trajectory's solve.
struct MmaScalars {
  double rho;    // conservativeness parameter
  double fval;   // f at the approximation point x (the best value so far)
  double fcur;   // f at xcur
  double gval;   // value of the approximation at xcur
  double wval;   // sum_j dx^2 / (2 (sigma^2 - dx^2))
  int k;         // outer iteration counter
  int evals;     // objective evaluations so far
};

// Strided views so that host code can use contiguous vectors and the batch kernel a [variable][trajectory]
// layout (coalesced across the trajectories of a warp).
struct MmaVec {
  double* p;
  long long stride;
  GTOP_MMA_HD double& operator[](int j) const { return p[j * stride]; }
};
struct MmaCVec {
  const double* p;
  long long stride;
  GTOP_MMA_HD const double& operator[](int j) const { return p[j * stride]; }
};

// ---- per-variable pieces (the batch kernel spreads the variables of a trajectory over the lanes of a warp; the host
// loop runs them in index order, which reproduces NLopt's arithmetic exactly) ------------------------------------

GTOP_MMA_HD double mma_sigma0(double lb, double ub) { return (std::isinf(lb) || std::isinf(ub)) ? 1.0 : 0.5 * (ub - lb); }

// closed-form minimiser of variable j's term of the separable approximation; adds its share to gsum / wsum
GTOP_MMA_HD double mma_candidate_var(double xj, double g, double sg, double lo, double hi, double rho, double& gsum,
                                     double& wsum) {
  if (sg == 0.0) return xj;  // lb == ub: the variable does not move
  const double sigma2 = sg * sg;
  const double v = fabs(g) * sg + 0.5 * rho;
  const double u = g * sigma2;
  const double q = u / (v * sg);
  double dx = (u / v) / (-1.0 - sqrt(fabs(1.0 - q * q)));
  double xc = xj + dx;
  if (xc > hi) xc = hi;
  else if (xc < lo) xc = lo;
  if (xc > xj + 0.9 * sg) xc = xj + 0.9 * sg;
  else if (xc < xj - 0.9 * sg) xc = xj - 0.9 * sg;
  dx = xc - xj;
  const double dx2 = dx * dx;
  const double denominv = 1.0 / (sigma2 - dx2);
  gsum += (g * (sigma2 * dx) + v * dx2) * denominv;
  wsum += 0.5 * dx2 * denominv;
  return xc;
}

// asymptote update of variable j at the end of an outer iteration (from the second one on)
GTOP_MMA_HD double mma_sigma_update(double sg, double xc, double xp, double xpp, double lo, double hi) {
  const double dx2 = (xc - xp) * (xp - xpp);
  sg *= dx2 < 0.0 ? 0.7 : (dx2 > 0.0 ? 1.2 : 1.0);
  if (!std::isinf(lo) && !std::isinf(hi)) {
    sg = fmin(sg, 10.0 * (hi - lo));
    sg = fmax(sg, 0.01 * (hi - lo));
  }
  return sg;
}

// What the per-variable bookkeeping has to do after an evaluation; decided from the scalars alone.
struct MmaDecision {
  bool accept;       // xcur becomes the approximation point (x, dfdx := xcur, dfdx_cur)
  bool outer_done;   // the outer iteration ended: xprevprev := xprev, xprev := xcur
  bool sigma_update; // ... and sigma is updated first (second outer iteration onwards)
};

// Scalar part of the bookkeeping after f(xcur) = fcur: updates rho / fval / k / evals, returns what to do per variable.
GTOP_MMA_HD MmaDecision mma_after_eval_scalars(double fcur, MmaScalars& s) {
  MmaDecision d{false, false, false};
  ++s.evals;
  s.fcur = fcur;
  if (s.evals == 1) {  // f(x0): approximation point = x0; the first outer iteration starts (xprev := x0)
    s.fval = fcur, s.k = 1;
    d.accept = true, d.outer_done = true;
    return d;
  }
  const bool inner_done = s.gval >= fcur;
  if (fcur < s.fval) s.fval = fcur, d.accept = true;  // a better point, inside inner iterations too
  if (!inner_done) {
    if (fcur > s.gval) s.rho = fmin(10.0 * s.rho, 1.1 * (s.rho + (fcur - s.gval) / s.wval));
    return d;
  }
  // the outer iteration k ends: new rho and sigma for iteration k+1, then iteration k+1 begins
  s.rho = fmax(0.1 * s.rho, kMmaRhoMin);
  d.outer_done = true, d.sigma_update = s.k > 1;
  ++s.k;
  return d;
}

// ---- whole-vector forms (host) -------------------------------------------------------------------------------------

GTOP_MMA_HD void mma_init(int n, MmaCVec lb, MmaCVec ub, MmaVec sigma, MmaScalars& s) {
  for (int j = 0; j < n; ++j) sigma[j] = mma_sigma0(lb[j], ub[j]);
  s.rho = 1.0, s.k = 0, s.evals = 0, s.fval = s.fcur = s.gval = s.wval = 0.0;
}

// The closed-form minimiser of the separable approximation around (x, dfdx, fval): writes xcur, s.gval, s.wval.
GTOP_MMA_HD void mma_candidate(int n, MmaCVec x, MmaCVec dfdx, MmaCVec sigma, MmaCVec lb, MmaCVec ub, MmaVec xcur,
                               MmaScalars& s) {
  double gval = s.fval, wval = 0.0;
  for (int j = 0; j < n; ++j) xcur[j] = mma_candidate_var(x[j], dfdx[j], sigma[j], lb[j], ub[j], s.rho, gval, wval);
  s.gval = gval, s.wval = wval;
}

// Bookkeeping after f and grad f have been evaluated at xcur (fcur, dfdx_cur).  On return the state is ready for the
// next mma_candidate.  The very first evaluation (at x0 = xcur) only seeds the state.
GTOP_MMA_HD void mma_after_eval(int n, double fcur, MmaCVec dfdx_cur, MmaVec x, MmaVec dfdx, MmaVec sigma, MmaCVec lb,
                                MmaCVec ub, MmaCVec xcur, MmaVec xprev, MmaVec xprevprev, MmaScalars& s) {
  const MmaDecision d = mma_after_eval_scalars(fcur, s);
  for (int j = 0; j < n; ++j) {
    if (d.accept) x[j] = xcur[j], dfdx[j] = dfdx_cur[j];
    if (d.sigma_update) sigma[j] = mma_sigma_update(sigma[j], xcur[j], xprev[j], xprevprev[j], lb[j], ub[j]);
    if (d.outer_done) xprevprev[j] = xprev[j], xprev[j] = xcur[j];
  }
}

}  // namespace gtop
