// problem_setup.cpp — see problem_setup.hpp.
#include "problem_setup.hpp"

#include <cmath>
#include <cstring>

#include "gtop_types.cuh"

namespace gtop {

namespace {

int factorial(int x) {
  int f = 1;
  for (int i = 2; i <= x; ++i) f *= i;
  return f;
}

// 6x6 inverse by partial-pivot LU (row-major a[r][c]).  Returns false when singular.
bool invert6(const double A[36], double Inv[36]) {
  double lu[36];
  int perm[6];
  std::memcpy(lu, A, sizeof(lu));
  for (int i = 0; i < 6; ++i) perm[i] = i;
  for (int k = 0; k < 6; ++k) {
    int piv = k;
    double best = std::fabs(lu[k * 6 + k]);
    for (int i = k + 1; i < 6; ++i)
      if (std::fabs(lu[i * 6 + k]) > best) best = std::fabs(lu[i * 6 + k]), piv = i;
    if (best == 0.0) return false;
    if (piv != k) {
      for (int j = 0; j < 6; ++j) std::swap(lu[k * 6 + j], lu[piv * 6 + j]);
      std::swap(perm[k], perm[piv]);
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
inline int ref_row_of_natural(int r) { return r < 3 ? 2 * r : 2 * (r - 3) + 1; }

}  // namespace

void build_time_table(ProblemHost& p) {
  const int m = p.m, NS = p.ns_cap;
  p.dt.assign(m, 0.0);
  p.ns.assign(m, 0.0);
  p.tt.assign((size_t)m * 5 * NS, 0.0);
  for (int s = 0; s < m; ++s) {
    const double T = p.seg_time[s];
    const double dt = T / (double)NS;  // grad_traj_optimizer.cpp:351
    p.dt[s] = dt;
    int k = 0;
    // for (double t = 1e-3; t < T; t += dt)  (:353): sequential accumulation reproduced
    for (double t = 1e-3; t < T && k < NS; t += dt, ++k) {
      double* tt = &p.tt[(size_t)s * 5 * NS];
      tt[k] = t;
      for (int q = 2; q <= 5; ++q) tt[(size_t)(q - 1) * NS + k] = std::pow(t, q);  // :457-485, :549
    }
    p.ns[s] = (double)k;
  }
}

std::string setup_from_dense(int m, const double* seg_time, const double* L, const double* R, const double* Dfc,
                             const gtop_params& prm, ProblemHost& out) {
  if (m < 2) return "need at least 2 segments (the reference's StackOptiDep indexes out of range for m = 1)";
  if (prm.samples_per_segment < 1) return "samples_per_segment must be >= 1";
  out = ProblemHost();
  out.m = m;
  out.ns_cap = prm.samples_per_segment;
  out.seg_time.assign(seg_time, seg_time + m);
  const int rowsL = 6 * m, nd = 3 * m + 3, P = 3 * m - 3;
  // ---- blocks of L --------------------------------------------------------------------------
  out.Linv.assign((size_t)m * 36, 0.0);
  std::vector<char> used(nd);
  for (int s = 0; s < m; ++s) {
    std::fill(used.begin(), used.end(), 0);
    for (int r = 0; r < 6; ++r) {
      const int col = r < 3 ? dense_col(m, s, r) : dense_col(m, s + 1, r - 3);
      used[col] = 1;
      for (int i = 0; i < 6; ++i) out.Linv[(size_t)s * 36 + i * 6 + r] = L[(size_t)col * rowsL + 6 * s + i];
    }
    for (int c = 0; c < nd; ++c)
      if (!used[c])
        for (int i = 0; i < 6; ++i)
          if (L[(size_t)c * rowsL + 6 * s + i] != 0.0)
            return "L is not block-structured (non-zero outside a segment's 6x6 block): not a StackOptiDep mapping";
  }
  // ---- bands of R ---------------------------------------------------------------------------
  auto Rw = [&](int u, int v) {  // waypoint-order access
    return R[(size_t)dense_col(m, v / 3, v % 3) * nd + dense_col(m, u / 3, u % 3)];
  };
  for (int u = 0; u < nd; ++u)
    for (int v = 0; v < nd; ++v) {
      const int du = u / 3 - v / 3;
      if ((du > 1 || du < -1) && Rw(u, v) != 0.0) return "R is not block-tridiagonal in waypoint order";
    }
  out.Rc.assign((size_t)nd * 9, 0.0);
  for (int v = 0; v < nd; ++v)
    for (int j = 0; j < 9; ++j) {
      const int u = 3 * (v / 3 - 1) + j;
      if (u >= 0 && u < nd) out.Rc[(size_t)v * 9 + j] = Rw(u, v);
    }
  out.Rg.assign((size_t)P * 9, 0.0);
  for (int i = 0; i < P; ++i) {
    const int wp = i / 3 + 1, self = 3 * wp + i % 3;
    for (int j = 0; j < 9; ++j) {
      const int nb = 3 * (wp - 1) + j, wpn = nb / 3;
      const bool fixed = (wpn == 0 || wpn == m);
      out.Rg[(size_t)i * 9 + j] = fixed ? Rw(nb, self) /* Rfp(f,i) */ : Rw(self, nb) /* Rpp(i,j) */;
    }
  }
  out.Df.assign(18, 0.0);
  for (int a = 0; a < 3; ++a)
    for (int j = 0; j < 6; ++j) out.Df[a * 6 + j] = Dfc[j * 3 + a];  // 3x6 column-major in
  build_time_table(out);
  return "";
}

// Shared tail of the native set-up: per-segment blocks, banded R, Df, x0, bounds, time table.
// pos: (m+1) x 3 waypoint positions; vel/acc: (m+1) x 3 or nullptr (zeros).
static std::string setup_common(int m, const double* pos, const double* vel, const double* acc, const gtop_params& prm,
                                ProblemHost& out) {
  const int nd = 3 * m + 3, P = 3 * m - 3, n = 3 * P;
  const double* wp = pos;
  for (int i = 0; i < m; ++i)
    if (!(out.seg_time[i] > 0.0)) return "non-positive segment time (coincident waypoints?)";
  out.Linv.assign((size_t)m * 36, 0.0);
  out.Mseg.assign((size_t)m * 36, 0.0);
  for (int s = 0; s < m; ++s) {
    const double T = out.seg_time[s];
    double A[36] = {0}, Ainv[36], Q[36] = {0};
    for (int i = 0; i < 3; ++i) {  // qp_generator.cpp:187-192
      A[(2 * i) * 6 + i] = factorial(i);
      for (int j = i; j < 6; ++j) A[(2 * i + 1) * 6 + j] = factorial(j) / factorial(j - i) * std::pow(T, j - i);
    }
    for (int i = 3; i < 6; ++i)  // :227-233 (integer prefactor, exact)
      for (int j = 3; j < 6; ++j)
        Q[i * 6 + j] = i * (i - 1) * (i - 2) * j * (j - 1) * (j - 2) / (i + j - 5) * std::pow(T, i + j - 5);
    if (!invert6(A, Ainv)) return "singular segment mapping matrix";
    double* Li = &out.Linv[(size_t)s * 36];
    for (int i = 0; i < 6; ++i)
      for (int r = 0; r < 6; ++r) Li[i * 6 + r] = Ainv[i * 6 + ref_row_of_natural(r)];
    // M = (Ainv^T Q) Ainv in natural order
    double T1[36];
    for (int r = 0; r < 6; ++r)
      for (int j = 0; j < 6; ++j) {
        double acc = 0.0;
        for (int i = 0; i < 6; ++i) acc += Li[i * 6 + r] * Q[i * 6 + j];
        T1[r * 6 + j] = acc;
      }
    double* M = &out.Mseg[(size_t)s * 36];
    for (int r = 0; r < 6; ++r)
      for (int c = 0; c < 6; ++c) {
        double acc = 0.0;
        for (int j = 0; j < 6; ++j) acc += T1[r * 6 + j] * Li[j * 6 + c];
        M[r * 6 + c] = acc;
      }
  }
  // R in waypoint order = sum of the segment forms scattered on (wp s, wp s+1)
  auto Rw = [&](int u, int v) {
    const int wu = u / 3, wv = v / 3, qu = u % 3, qv = v % 3;
    double r = 0.0;
    if (wu == wv) {
      if (wu >= 1) r += out.Mseg[(size_t)(wu - 1) * 36 + (3 + qu) * 6 + 3 + qv];
      if (wu <= m - 1) r += out.Mseg[(size_t)wu * 36 + qu * 6 + qv];
    } else if (wv == wu + 1) {
      r = out.Mseg[(size_t)wu * 36 + qu * 6 + 3 + qv];
    } else if (wv == wu - 1) {
      r = out.Mseg[(size_t)wv * 36 + (3 + qu) * 6 + qv];
    }
    return r;
  };
  out.Rc.assign((size_t)nd * 9, 0.0);
  for (int v = 0; v < nd; ++v)
    for (int j = 0; j < 9; ++j) {
      const int u = 3 * (v / 3 - 1) + j;
      if (u >= 0 && u < nd) out.Rc[(size_t)v * 9 + j] = Rw(u, v);
    }
  out.Rg.assign((size_t)P * 9, 0.0);
  for (int i = 0; i < P; ++i) {
    const int w = i / 3 + 1, self = 3 * w + i % 3;
    for (int j = 0; j < 9; ++j) {
      const int nb = 3 * (w - 1) + j, wpn = nb / 3;
      const bool fixed = (wpn == 0 || wpn == m);
      out.Rg[(size_t)i * 9 + j] = fixed ? Rw(nb, self) : Rw(self, nb);
    }
  }
  // Df / x0 (getInitialD, qp_generator.cpp:407-451 with the straight-line init :204-221, zero start
  // velocity / acceleration :83-85) and the NLopt box (grad_traj_optimizer.cpp:151-177)
  out.Df.assign(18, 0.0);
  for (int a = 0; a < 3; ++a) out.Df[a * 6 + 0] = wp[a], out.Df[a * 6 + 3] = wp[3 * m + a];
  out.x0.assign(n, 0.0), out.lb.assign(n, 0.0), out.ub.assign(n, 0.0);
  for (int a = 0; a < 3; ++a)
    for (int i = 0; i < P; ++i) {
      const int w = i / 3 + 1, q = i % 3;
      const double c = wp[3 * w + a];
      out.x0[a * P + i] = (q == 0) ? c : (q == 1 ? (vel ? vel[3 * w + a] : 0.0) : (acc ? acc[3 * w + a] : 0.0));
      const double half = q == 0 ? prm.bos : (q == 1 ? prm.vos : prm.aos);
      out.lb[a * P + i] = (q == 0 ? c : 0.0) - half;  // box centred on the path / on zero (:154-177)
      out.ub[a * P + i] = (q == 0 ? c : 0.0) + half;
    }
  out.has_path = true;
  build_time_table(out);
  return "";
}

std::string setup_from_path(int m, const double* wp, const gtop_params& prm, ProblemHost& out) {
  if (m < 2) return "need at least 2 segments";
  if (prm.samples_per_segment < 1) return "samples_per_segment must be >= 1";
  if (!(prm.mean_v > 0.0)) return "mean_v must be > 0";
  out = ProblemHost();
  out.m = m;
  out.ns_cap = prm.samples_per_segment;
  // time allocation (grad_traj_optimizer.cpp:73-81; init_time on the first segment only, H9)
  out.seg_time.assign(m, 0.0);
  for (int i = 0; i < m; ++i) {
    const double dx = wp[3 * i] - wp[3 * i + 3], dy = wp[3 * i + 1] - wp[3 * i + 4], dz = wp[3 * i + 2] - wp[3 * i + 5];
    const double len = std::sqrt(dx * dx + dy * dy + dz * dz);
    out.seg_time[i] = (i == 0) ? len / prm.mean_v + prm.init_time : len / prm.mean_v;
  }
  return setup_common(m, wp, nullptr, nullptr, prm, out);
}

// setKinoPath (grad_traj_optimizer.cpp:35-65) + PolyKinoGeneration (qp_generator.cpp:23-154): segment times
// are given; x0 carries the front end's velocities / accelerations at the interior waypoints; Df keeps zero
// start/end velocity and acceleration (startVel/startAcc are stale zero globals upstream — SURVEY H14).
std::string setup_from_kino(int m, const double* pos, const double* vel, const double* acc, const double* time,
                            const gtop_params& prm, ProblemHost& out) {
  if (m < 2) return "need at least 2 segments";
  if (prm.samples_per_segment < 1) return "samples_per_segment must be >= 1";
  out = ProblemHost();
  out.m = m;
  out.ns_cap = prm.samples_per_segment;
  out.seg_time.assign(time, time + m);
  return setup_common(m, pos, vel, acc, prm, out);
}

void dense_L(const ProblemHost& p, double* L) {
  const int m = p.m, rows = 6 * m, nd = 3 * m + 3;
  std::memset(L, 0, sizeof(double) * (size_t)rows * nd);
  for (int s = 0; s < m; ++s)
    for (int r = 0; r < 6; ++r) {
      const int col = r < 3 ? dense_col(m, s, r) : dense_col(m, s + 1, r - 3);
      for (int i = 0; i < 6; ++i) L[(size_t)col * rows + 6 * s + i] = p.Linv[(size_t)s * 36 + i * 6 + r];
    }
}

void dense_R(const ProblemHost& p, double* R) {
  const int m = p.m, nd = 3 * m + 3;
  std::memset(R, 0, sizeof(double) * (size_t)nd * nd);
  for (int v = 0; v < nd; ++v)
    for (int j = 0; j < 9; ++j) {
      const int u = 3 * (v / 3 - 1) + j;
      if (u < 0 || u >= nd) continue;
      R[(size_t)dense_col(m, v / 3, v % 3) * nd + dense_col(m, u / 3, u % 3)] = p.Rc[(size_t)v * 9 + j];
    }
}

void pack_record(const ProblemHost& p, double* rec) {
  const ProbLayout L = make_layout(p.m, p.ns_cap);
  std::memset(rec, 0, sizeof(double) * (size_t)L.stride);
  std::memcpy(rec + L.off_dt, p.dt.data(), sizeof(double) * p.m);
  std::memcpy(rec + L.off_ns, p.ns.data(), sizeof(double) * p.m);
  std::memcpy(rec + L.off_Linv, p.Linv.data(), sizeof(double) * p.Linv.size());
  std::memcpy(rec + L.off_tt, p.tt.data(), sizeof(double) * p.tt.size());
  std::memcpy(rec + L.off_Rc, p.Rc.data(), sizeof(double) * p.Rc.size());
  std::memcpy(rec + L.off_Rg, p.Rg.data(), sizeof(double) * p.Rg.size());
  std::memcpy(rec + L.off_Df, p.Df.data(), sizeof(double) * 18);
  for (int s = 0; s < p.m; ++s)
    for (int i = 0; i < 6; ++i)
      for (int r = 0; r < 6; ++r) rec[L.off_LinvT + s * 36 + r * 6 + i] = p.Linv[(size_t)s * 36 + i * 6 + r];
  // ---- tables of the multi-start (DMMA) kernel ------------------------------------------------
  const int NS = p.ns_cap, NSP = L.nsp;
  for (int s = 0; s < p.m; ++s) {
    const double* tt = &p.tt[(size_t)s * 5 * NS];
    const int ns = (int)p.ns[s];
    // TQ / PL (gtop_types.cuh): sample slot k = 4 it + j; padding slots stay zero (memset above)
    for (int k = 0; k < NSP && k < ns; ++k) {
      const int it = k >> 2, j = k & 3;
      double pw[6] = {1.0, 0, 0, 0, 0, 0};  // tau_i = t^i (:549)
      for (int q = 1; q <= 5; ++q) pw[q] = tt[(size_t)(q - 1) * NS + k];
      double* tq = rec + L.off_TQ + ((size_t)s * (NSP / 4) + it) * 24;
      for (int part = 0; part < 3; ++part) tq[(part * 4 + j) * 2] = pw[2 * part], tq[(part * 4 + j) * 2 + 1] = pw[2 * part + 1];
      // (dt * phi_r(t_k), dt * phi_r'(t_k)), phi_r(t) = sum_i t^i Linv[i][r]: T Ldp and T V Ldp of :376-381 times dt (:381)
      double* pl = rec + L.off_PL + ((size_t)s * (NSP / 4) + it) * 64;
      for (int g = 0; g < 8; ++g) {
        const int r = deriv_of_col(g, s);
        if (r < 0) continue;
        double a = 0.0, b = 0.0;
        for (int i = 0; i < 6; ++i) a += pw[i] * p.Linv[(size_t)s * 36 + i * 6 + r];
        for (int i = 1; i < 6; ++i) b += (double)i * pw[i - 1] * p.Linv[(size_t)s * 36 + i * 6 + r];
        pl[(4 * g + j) * 2] = p.dt[s] * a, pl[(4 * g + j) * 2 + 1] = p.dt[s] * b;
      }
    }
    for (int pp = 0; pp < 3; ++pp)
      for (int r = 0; r < 6; ++r) {
        rec[L.off_LiTP + s * 36 + (pp * 6 + r) * 2 + 0] = p.Linv[(size_t)s * 36 + (2 * pp) * 6 + r];
        rec[L.off_LiTP + s * 36 + (pp * 6 + r) * 2 + 1] = p.Linv[(size_t)s * 36 + (2 * pp + 1) * 6 + r];
      }
  }
}

void coeff_from_x(const ProblemHost& p, const double* x, double* coeff) {
  const int m = p.m, P = 3 * m - 3;
  for (int a = 0; a < 3; ++a)
    for (int s = 0; s < m; ++s) {
      double d[6];
      for (int r = 0; r < 6; ++r) {
        const int w = r < 3 ? s : s + 1, q = r % 3;
        d[r] = w == 0 ? p.Df[a * 6 + q] : (w == m ? p.Df[a * 6 + 3 + q] : x[a * P + 3 * (w - 1) + q]);
      }
      const bool last = (s == m - 1);
      for (int i = 0; i < 6; ++i) {
        const double* Li = &p.Linv[(size_t)s * 36 + i * 6];
        double acc = 0.0;
        for (int k = 0; k < 6; ++k) {
          const int r = last ? (k + 3) % 6 : k;  // ascending dense-column order of L*d
          const double prod = Li[r] * d[r];
          acc = (k == 0) ? prod : acc + prod;
        }
        coeff[(size_t)s * 18 + a * 6 + i] = acc;
      }
    }
}

}  // namespace gtop
