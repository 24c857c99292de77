// problem_setup.cpp — see problem_setup.hpp.
#include "problem_setup.hpp"

#include <cmath>
#include <cstring>

#include "gtop_types.cuh"
#include "setup_core.cuh"

namespace gtop {


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
      for (int q = 2; q <= 5; ++q) tt[(size_t)(q - 1) * NS + k] = setup_pow(t, q);  // :457-485, :549
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
    if (!setup_segment_blocks(T, &out.Linv[(size_t)s * 36], &out.Mseg[(size_t)s * 36])) return "singular segment mapping matrix";
  }
  // R in waypoint order = sum of the segment forms scattered on (wp s, wp s+1)
  out.Rc.assign((size_t)nd * 9, 0.0);
  for (int v = 0; v < nd; ++v)
    for (int j = 0; j < 9; ++j) out.Rc[(size_t)v * 9 + j] = setup_Rc_entry(out.Mseg.data(), m, v, j);
  out.Rg.assign((size_t)P * 9, 0.0);
  for (int i = 0; i < P; ++i)
    for (int j = 0; j < 9; ++j) out.Rg[(size_t)i * 9 + j] = setup_Rg_entry(out.Mseg.data(), m, i, j);
  // Df / x0 (getInitialD, qp_generator.cpp:407-451 with the straight-line init :204-221, zero start
  // velocity / acceleration :83-85) and the NLopt box (grad_traj_optimizer.cpp:151-177)
  out.Df.assign(18, 0.0);
  for (int a = 0; a < 3; ++a) out.Df[a * 6 + 0] = wp[a], out.Df[a * 6 + 3] = wp[3 * m + a];
  out.x0.assign(n, 0.0), out.lb.assign(n, 0.0), out.ub.assign(n, 0.0);
  for (int a = 0; a < 3; ++a)
    for (int i = 0; i < P; ++i)
      setup_x0_box(wp, vel, acc, a, i, prm.bos, prm.vos, prm.aos, &out.x0[a * P + i], &out.lb[a * P + i], &out.ub[a * P + i]);
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
  for (int i = 0; i < m; ++i) out.seg_time[i] = setup_segment_time(wp, i, prm.mean_v, prm.init_time);
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
  std::memcpy(rec + L.off_T, p.seg_time.data(), sizeof(double) * p.m);
  for (int i = 0; i < 3 * p.m - 3; ++i)
    for (int j = 0; j < 9; ++j) rec[L.off_RgP + i * 10 + j] = p.Rg[(size_t)i * 9 + j];
  for (int s = 0; s < p.m; ++s)
    for (int i = 0; i < 6; ++i)
      for (int r = 0; r < 6; ++r) rec[L.off_LinvT + s * 36 + r * 6 + i] = p.Linv[(size_t)s * 36 + i * 6 + r];
  // ---- tables of the multi-start (DMMA) kernel ------------------------------------------------
  const int NS = p.ns_cap;
  for (int s = 0; s < p.m; ++s) {
    const double* tt = &p.tt[(size_t)s * 5 * NS];
    const int ns = (int)p.ns[s];
    const int ns_first = (s & 1) ? (int)p.ns[s - 1] : ns;
    // TQ / PL in the pair schedule (gtop_types.cuh); idle slots stay zero (memset above)
    for (int k = 0; k < ns; ++k) {
      int u, j, rec_pl;
      setup_pair_slot(s, k, ns_first, L.nit2, &u, &j, &rec_pl);
      double pw[6] = {1.0, 0, 0, 0, 0, 0};  // tau_i = t^i (:549)
      for (int q = 1; q <= 5; ++q) pw[q] = tt[(size_t)(q - 1) * NS + k];
      double* tq = rec + L.off_TQ + ((size_t)(s >> 1) * L.nit2 + u) * 24;
      for (int part = 0; part < 3; ++part) tq[(part * 4 + j) * 2] = pw[2 * part], tq[(part * 4 + j) * 2 + 1] = pw[2 * part + 1];
      tq[j * 2] = p.dt[s];  // t^0 is not read: its place carries the slot's quadrature weight dt_s (0 for idle slots)
      // (dt * phi_r(t_k), dt * phi_r'(t_k)), phi_r(t) = sum_i t^i Linv[i][r]: T Ldp and T V Ldp of :376-381 times dt (:381)
      double* pl = rec + L.off_PL + ((size_t)(s >> 1) * (L.nit2 + 1) + rec_pl) * 64;
      for (int g = 0; g < 8; ++g) {
        const int r = deriv_of_col(g, s);
        if (r < 0) continue;
        setup_basis_pair(pw, &p.Linv[(size_t)s * 36], r, p.dt[s], &pl[(4 * g + j) * 2], &pl[(4 * g + j) * 2 + 1]);
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
