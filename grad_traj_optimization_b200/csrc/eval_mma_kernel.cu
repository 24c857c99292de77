// eval_mma_kernel.cu — cost + gradient of GTOP trajectories for MULTI-START batches on sm_100a:
// groups of 8 consecutive trajectories share one problem (same waypoints / segment times), which is
// BASELINE configs[2] (65,536 = 64 waypoint sets x 1,024 starts) and configs[4].
//
// Replaces GradTrajOptimizer::getCostAndGradient (src/grad_traj_optimizer.cpp:281-448 of the
// reference), same maths as eval_kernel.cu; what changes is the work decomposition:
//
//   warp = 8 trajectories x 4 lanes.  Segments are walked in pairs whose samples are numbered consecutively;
//   lane (g, j) = (lane >> 2, lane & 3) evaluates slot 4*it + j of trajectory g in iteration `it` of the current
//   pair (15 iterations per pair at N_s = 30: 75 per 10 segments where a per-segment loop needs 80, because 30
//   samples leave two of the four lanes idle in every eighth iteration).
//
//   * The per-sample contribution to the gradient rows of the segment's two waypoints,
//         G[a][r] += dt * (sA_a * phi_r(t) + sB_a * phi_r'(t)),   phi_r(t) = sum_i t^i Linv[i][r]
//     (the reference's (gd grad_a cd |v| T + cd v_a/|v| T V) Ldp dt, :376-381, with T Ldp = phi and T V Ldp = phi'),
//     is a (8 trajectories x 4 samples) x (4 samples x 8 columns) product per axis and per kind, because
//     phi / phi' only depend on the problem: it is issued as FP64 tensor-core tiles
//     (mma.sync.m8n8k4.f64, SASS DMMA) — A fragment = the lane's own sA_a (or sB_a), B fragment = a
//     host-built table of dt*phi_r(t_k), dt*phi_r'(t_k), C fragment = 2 accumulators per axis.  The DMMA sums
//     over the 4 lanes of a trajectory, so there is NO cross-lane shuffle reduction, no separate projection
//     through Linv, and the accumulators shrink from 19 to 6 doubles per lane.  DMMA shares the FP64 pipe
//     with DFMA on B200 (tools/fp64_peak.cu: 36.9 vs 34.0 TFLOP/s, mixed = sum of both), so this does not
//     buy FP64 throughput; it buys issue slots (6 DMMA instead of ~50 DFMA/DMUL per sample) and removes the
//     shuffle/shared-memory traffic of the reduction — the general kernel is L1TEX- and issue-bound, not
//     FP64-bound (profiles/).
//   * The columns alternate with the segment parity (deriv_of_col): the end-derivative columns of
//     segment s are the start-derivative columns of segment s+1, so the gradient rows of an interior
//     waypoint accumulate in place in the C fragment across the two segments that touch it and are
//     written once.
//   * Coefficients c = L d (:253-279) stay scalar, in the reference's summation order, unfused.
//   * SDF corners come from the patch-ordered xy-quad copy of the grid: 2 x LDG.256 per sample.
//   * Per-problem tables are stored in the order the lanes read them (TQ: powers of the sample time, PL: B
//     fragments), 7 L1 wavefronts per iteration instead of 14 for per-sample records; the time record of the NEXT
//     iteration is requested as soon as the current one is dead (after the velocity polynomial), so its L2
//     latency hides behind the second half of the iteration.
//   * The smoothness cost is accumulated from the gradient rows (d_u * (R d)_u) while they are formed;
//     the rows of the fixed start / end state are added before / after the segment loop.
//   * The boundary derivatives live in a six-waypoint ring per trajectory (shared memory independent of m),
//     refilled from x by cp.async one pair of segments ahead.
//   * The tables carry the pair schedule (which sample a slot holds, its quadrature weight, zero weights for idle
//     slots): the loop only picks each lane's coefficient set and, once per pair, moves the first segment's
//     waypoint rows out of the C fragments (in the one iteration that holds samples of both segments the DMMAs are
//     issued twice, once per segment, against B fragments that are zero for the other segment's lanes).
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>

#include "eval_common.cuh"
#include "gtop_types.cuh"

namespace gtop {

// resident CTAs per SM asked of the compiler: the kernel runs 20 warps per SM (96 registers: the register file is
// per scheduler, 16384 / (96 * 32) = 5.3 -> 5 warps each); 17..20 all compile to 96 registers, with different
// schedules — 18 measured fastest (0.3425 ms against 0.3456 at 19 / 20, 0.3451 at 17); 16 -> 126 registers, no
// spills, 16 warps: 0.350; 24 -> 80 registers, 24 warps, 98 B of spills: 0.382
#ifndef GTOP_MMA_MINB
#define GTOP_MMA_MINB 18
#endif
// one warp per CTA: the kernel has no block-level cooperation, and small CTAs pack SMs evenly
#ifndef GTOP_MMA_WARPS
#define GTOP_MMA_WARPS 1
#endif
constexpr int kMmaMaxWarps = GTOP_MMA_WARPS;

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// per-warp shared memory, in doubles:
//   ring [8][55]        boundary derivatives of SIX consecutive waypoints per trajectory, [slot][axis][p,v,a] with
//                       slot = waypoint % 6: a pair of segments (s, s+1) needs waypoints s-1 .. s+2 (coefficients and
//                       the smoothness rows of waypoints s, s+1), and waypoints s+3, s+4 are copied from x with
//                       cp.async WHILE the pair runs, so the read of x never stalls a warp.  Independent of m (a full
//                       copy of x was 8 x 9(m+1) doubles: 6.3 KB per warp at m = 10, 18 KB at m = 30); stride 55 is
//                       odd -> the 8 trajectories of a warp sit in distinct banks
//   coef [2][8][15][2]  coefficient pairs of the two segments of the current pair, 15 16-byte elements per trajectory:
//                       per axis (c0,c1) (c2,c3) (c4,c5) (2c2,3c3) (4c4,5c5).  The 240-byte trajectory
//                       stride puts the 8 trajectories of a warp in the 8 distinct 16-byte bank groups
//                       both for the loads (fixed element) and the stores (element varies with the lane)
//   scr  [8][9]         finished collision-gradient rows of one waypoint (3 axes x p,v,a)
constexpr int kRingSlots = 6;
constexpr int kRing = kRingSlots * 9 + 1;
__host__ __device__ inline int mma_warp_smem_doubles() { return 8 * kRing + 2 * 240 + 72; }

// coefficients c[a][i] = sum_r Linv[i][r] * dseg[a][r], summed in ascending dense-column order of L
// (start derivatives first, except for the last segment whose end derivatives live in columns 3..5
// of d) with unfused mul/add — the same 6-term sum as `L * d` (:273).  Lane j < 3 takes the power pair
// (2j, 2j+1) of all three axes: its six Linv pairs are loaded once (3 distinct addresses per warp
// instruction), and the four lanes of a trajectory read the same waypoint values (shared-memory broadcast).
template <bool kLast>
__device__ __forceinline__ void coef_phase(const double* __restrict__ wp0, const double* __restrict__ wp1,
                                           const double2* __restrict__ LiTP, double2* __restrict__ coef, int j) {
  if (j >= 3) return;
  const int p = j;
  double2 L[6];
#pragma unroll
  for (int r = 0; r < 6; ++r) L[r] = ldg_tab(LiTP + p * 6 + r);
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const double *ws = wp0 + 3 * a, *we = wp1 + 3 * a;  // (p,v,a) of the segment's start / end waypoint
    double q0[6], q1[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      const double w = r < 3 ? ws[r] : we[r - 3];
      q0[r] = __dmul_rn(L[r].x, w), q1[r] = __dmul_rn(L[r].y, w);
    }
    double c0, c1;
    if (kLast) {
      c0 = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(q0[3], q0[4]), q0[5]), q0[0]), q0[1]), q0[2]);
      c1 = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(q1[3], q1[4]), q1[5]), q1[0]), q1[1]), q1[2]);
    } else {
      c0 = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(q0[0], q0[1]), q0[2]), q0[3]), q0[4]), q0[5]);
      c1 = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(q1[0], q1[1]), q1[2]), q1[3]), q1[4]), q1[5]);
    }
    coef[a * 5 + p] = make_double2(c0, c1);
    // e_k = k*c_k for the velocity polynomial (2*coeff(s,2), 3*coeff(s,3), ... :477-485)
    if (p >= 1) coef[a * 5 + 2 + p] = make_double2(__dmul_rn((double)(2 * p), c0), __dmul_rn((double)(2 * p + 1), c1));
  }
}

// kNoAlloc: L1 policy of the gathers (see ldg_quad), chosen from the sampling density.
template <bool kNoAlloc>
__global__ void __launch_bounds__(kMmaMaxWarps * 32, GTOP_MMA_MINB)
k_eval_mma(SdfView S, ProbLayout PL, const double* __restrict__ prob, int num_problems, EvalScalars E,
           const double* __restrict__ X, int64_t B, int64_t traj_offset, int64_t starts_per_problem,
           double* __restrict__ cost_out, double* __restrict__ grad_out, unsigned long long* __restrict__ oob_counter) {
  extern __shared__ __align__(16) double smem[];
  const int m = PL.m, P = PL.P, n = PL.n;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
  const int g = lane >> 2, j = lane & 3;  // trajectory slot of the warp, sample slot of the iteration
  const int per_warp = mma_warp_smem_doubles();

  // 2^(j/16) table for exp2_tab: the only block-level synchronisation of the kernel
  double* xtab = smem + wpc * per_warp;
  exp2_tab_init(xtab, threadIdx.x, blockDim.x);
  __syncthreads();

  // Groups of 8 are aligned on the GLOBAL trajectory index (traj_offset + b), so that any split of a
  // batch evaluates every trajectory in the same lane group with the same instructions: results are
  // bit-identical for every partition of the batch across calls / GPUs.
  const int64_t gb0 = traj_offset - (traj_offset & 7) + ((int64_t)blockIdx.x * wpc + warp) * 8;
  if (gb0 >= traj_offset + B) return;                // the whole warp is past the end
  const int64_t b_raw = gb0 + g - traj_offset;
  const bool live = b_raw >= 0 && b_raw < B;         // ragged head / tail: dead slots shadow a live
  const int64_t b = b_raw < 0 ? 0 : (b_raw < B ? b_raw : B - 1);  // trajectory of the same group

  double* wbase = smem + warp * per_warp;
  double* ring = wbase + g * kRing;
  double2* coef = reinterpret_cast<double2*>(wbase + 8 * kRing) + g * 15;  // this trajectory's 15 pairs
  double* scr = wbase + 8 * kRing + 2 * 240 + g * 9;

  // all 8 trajectories of the group belong to one problem (host-checked: starts_per_problem % 8 == 0)
  const int64_t pid = (gb0 / starts_per_problem) % num_problems;
  const double* __restrict__ pr = prob + pid * (int64_t)PL.stride;
  const double* __restrict__ x = X + b * n;

  // ---- 1. boundary derivatives of waypoints 0, 1 (and 2, in flight) into the ring ------------------------------
  // waypoint w: the fixed start / end state (Df) for w = 0 / m, the free derivatives x[a*P + 3(w-1) + q] otherwise;
  // copied global -> shared with cp.async (no register staging, no stall until the values are needed)
  auto fetch_waypoint = [&](int w) {  // (a waypoint past the end still commits its — empty — group)
    if (w <= m) {
      double* dst = ring + (w % kRingSlots) * 9;
      for (int i = j; i < 9; i += 4) {
        const int a = i / 3, q = i - 3 * a;
        const double* src = w == 0 ? pr + PL.off_Df + a * 6 + q : (w == m ? pr + PL.off_Df + a * 6 + 3 + q : x + a * P + 3 * (w - 1) + q);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst + i)), "l"(src)
                     : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");  // one group per waypoint, in waypoint order
  };
  fetch_waypoint(0);
  fetch_waypoint(1);
  fetch_waypoint(2);
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncwarp();

  double Jc = 0.0, js = 0.0;
  // smoothness cost, rows of the fixed start waypoint (column band of R, :326-327): needs waypoints 0 and 1
  if (j < 3) {
    const double* __restrict__ band = pr + PL.off_Rc + j * 9;  // row u = j of waypoint 0; band entries 3..8 <-> wp 0, 1
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      double y = 0.0;
#pragma unroll
      for (int t = 0; t < 6; ++t) y = fma(ring[(t / 3) * 9 + a * 3 + t % 3], band[3 + t], y);
      js = fma(y, ring[a * 3 + j], js);
    }
  }

  double G[3][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};  // gradient rows of the two waypoints in flight (C fragments)
  unsigned oob = 0;
  const double kA = -E.alpha * E.alpha_over_r * S.res_inv;  // gd*cd = -(alpha^2/r) e^2; grad scale 1/res
  double* __restrict__ gout = grad_out + b * n;

  // rows of waypoint w (the start derivatives of segment w, which also collected the end derivatives of segment
  // w-1 in place) are complete: out of the C fragments, then smoothness rows and output
  auto flush_rows = [&](int w) {  // collision rows of waypoint w out of the C fragments
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int r = deriv_of_col(2 * j + e, w);
      if (r >= 0 && r < 3) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          scr[a * 3 + r] = G[a][e];  // w = 0: the fixed start state, never read
          G[a][e] = 0.0;
        }
      }
    }
    __syncwarp();
  };
  auto combine_rows = [&](int w) {
    if (w >= 1) {
      const double* wpm = ring + ((w - 1) % kRingSlots) * 9;  // waypoint w-1
      const double* wp0 = ring + (w % kRingSlots) * 9;        // waypoint w
      const double* wp1 = ring + ((w + 1) % kRingSlots) * 9;  // waypoint w+1
      // smoothness rows 2 Rfp' df + 2 Rpp dp (:330-332) from the band of R, combined with the collision
      // rows and written out (:425-431).  The same row products give the interior rows of the smoothness
      // cost d' R d (:326-327).  Lane r = j < 3 takes derivative r of the three axes: one band row per lane,
      // the window of the ring read as shared-memory broadcasts.
      if (j < 3) {
        const int r = j, jj = 3 * (w - 1) + r;
        const double2* __restrict__ band = reinterpret_cast<const double2*>(pr + PL.off_RgP + jj * 10);  // padded row
        double bnd[10];
#pragma unroll
        for (int u = 0; u < 5; ++u) {
          const double2 v = ldg_tab(band + u);
          bnd[2 * u] = v.x, bnd[2 * u + 1] = v.y;
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          double gs = 0.0;  // window = waypoints w-1, w, w+1 of axis a
#pragma unroll
          for (int u = 0; u < 3; ++u) gs = fma(wpm[a * 3 + u], bnd[u], gs);
#pragma unroll
          for (int u = 0; u < 3; ++u) gs = fma(wp0[a * 3 + u], bnd[3 + u], gs);
#pragma unroll
          for (int u = 0; u < 3; ++u) gs = fma(wp1[a * 3 + u], bnd[6 + u], gs);
          js = fma(gs, wp0[a * 3 + r], js);
          const double val = (E.ws * (2.0 * gs) + E.wc * scr[a * 3 + r]) + 1e-5;
          if (live) gout[a * P + jj] = val;
        }
      }
      __syncwarp();
    }
  };

  // Segments are walked in pairs (sa, sb) = (2p, 2p+1): the ns_a + ns_b samples of a pair are numbered consecutively
  // and iteration `it` evaluates slots 4 it + j, so a sample count that is not a multiple of 4 (30 upstream) idles
  // lanes once per pair instead of once per segment (75 instead of 80 iterations per 10 segments).  The tables carry
  // the schedule (which sample a slot holds, its quadrature weight, zero weights for idle slots); the loop only has to
  // pick each lane's coefficient set and to move the first segment's waypoint rows out of the C fragments when its
  // last sample has gone in.
  double2* coefB = coef + 8 * 15;  // coefficients of the pair's second segment (same layout, next 240 doubles)
  for (int sa = 0; sa < m; sa += 2) {
    const int sb = sa + 1;
    const bool hasb = sb < m;
    // waypoints sa+3, sa+4 (for the next pair) start flying; everything older — sa .. sa+2 — has landed
    fetch_waypoint(sa + 3);
    fetch_waypoint(sa + 4);
    asm volatile("cp.async.wait_group 2;" ::: "memory");
    __syncwarp();
    bool flushed = false;
    if (!E.skip_collision) {
      {
        const double2* LiTP = reinterpret_cast<const double2*>(pr + PL.off_LiTP + sa * 36);
        const double* wp0 = ring + (sa % kRingSlots) * 9;
        const double* wp1 = ring + ((sa + 1) % kRingSlots) * 9;
        if (sa == m - 1) coef_phase<true>(wp0, wp1, LiTP, coef, j);
        else coef_phase<false>(wp0, wp1, LiTP, coef, j);
      }
      if (hasb) {
        const double2* LiTP = reinterpret_cast<const double2*>(pr + PL.off_LiTP + sb * 36);
        const double* wp0 = ring + (sb % kRingSlots) * 9;
        const double* wp1 = ring + ((sb + 1) % kRingSlots) * 9;
        if (sb == m - 1) coef_phase<true>(wp0, wp1, LiTP, coefB, j);
        else coef_phase<false>(wp0, wp1, LiTP, coefB, j);
      }
      __syncwarp();

      const int nsa = (int)pr[PL.off_ns + sa], nsb = hasb ? (int)pr[PL.off_ns + sb] : 0;
      const int nit = (nsa + nsb + 3) >> 2;
      const bool mixed = nsb > 0 && (nsa & 3) != 0;  // one iteration holds samples of both segments
      // this lane samples the second segment from iteration ubl on (the mixed iteration splits the lanes)
      const int ubl = nsb > 0 ? (nsa >> 2) + ((mixed && j < (nsa & 3)) ? 1 : 0) : 0x7fffffff;
      // the first segment's last sample goes in at iteration uf (its rows are flushed after that iteration's DMMAs)
      const int uf = nsb > 0 ? (mixed ? (nsa >> 2) : (nsa >> 2) - 1) : -1;
      if (nsb > 0 && uf < 0) flush_rows(sa), flushed = true;  // (degenerate: the first segment has no samples)
      // lane-ordered tables of the pair (gtop_types.cuh): the 4 sample slots of an iteration read 64 contiguous bytes
      // per load of the time record (1 L1 wavefront), the 32 lanes 512 contiguous bytes of B fragments (4 wavefronts)
      const double2* __restrict__ tq = reinterpret_cast<const double2*>(pr + PL.off_TQ) + (size_t)(sa >> 1) * PL.nit2 * 12 + j;
      const double2* __restrict__ pl = reinterpret_cast<const double2*>(pr + PL.off_PL) + (size_t)(sa >> 1) * (PL.nit2 + 1) * 32 + lane;
      double2 ta, tb, tc;  // (dt,t) (t2,t3) (t4,t5) of the current sample, requested one iteration ahead
      if (nit > 0) ta = ldg_tab(tq), tb = ldg_tab(tq + 4), tc = ldg_tab(tq + 8);
      const double2* cf = 0 >= ubl ? coefB : coef;  // this lane's coefficient set in the current iteration
      for (int it = 0; it < nit; ++it, tq += 12, pl += 32) {
        const double t = ta.y, t2 = tb.x, t3 = tb.y, t4 = tc.x, t5 = tc.y;
        const double wdt = ta.x;  // dt of the slot's segment; 0 for idle slots, which evaluate a segment start in vain
        double pos[3], vel[3], c1[3];
        float pf[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          const double2 c01 = cf[a * 5 + 0], c23 = cf[a * 5 + 1], c45 = cf[a * 5 + 2];
          c1[a] = c01.y;
          // getPositionFromCoeff (:457-465): left-to-right, unfused, then float
          double pa = __dadd_rn(c01.x, __dmul_rn(c01.y, t));
          pa = __dadd_rn(pa, __dmul_rn(c23.x, t2));
          pa = __dadd_rn(pa, __dmul_rn(c23.y, t3));
          pa = __dadd_rn(pa, __dmul_rn(c45.x, t4));
          pa = __dadd_rn(pa, __dmul_rn(c45.y, t5));
          pf[a] = __double2float_rn(pa);
          pos[a] = (double)pf[a];
        }
        SdfCell cell;
        sdf_gather<kNoAlloc>(S, pos[0], pos[1], pos[2], pf[0], pf[1], pf[2], cell);  // 2 x LDG.256 in flight from here
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          const double2 e01 = cf[a * 5 + 3], e23 = cf[a * 5 + 4];
          // getVelocityFromCoeff (:477-485)
          double va = __dadd_rn(c1[a], __dmul_rn(e01.x, t));
          va = __dadd_rn(va, __dmul_rn(e01.y, t2));
          va = __dadd_rn(va, __dmul_rn(e23.x, t3));
          va = __dadd_rn(va, __dmul_rn(e23.y, t4));
          vel[a] = f32_round_trip(va);
        }
        // B fragments of this sample (column g of the lane), then the time record of the next sample
        const double2 bAB = ldg_tab(pl);
        if (it + 1 < nit) ta = ldg_tab(tq + 12), tb = ldg_tab(tq + 16), tc = ldg_tab(tq + 20);

        const double vn2 = fma(vel[2], vel[2], fma(vel[1], vel[1], fma(vel[0], vel[0], 1e-300)));
        const double vnorm = sqrt_pos(vn2) + 1e-5;  // :358 (the 1e-300 keeps v = 0 finite; absorbed by 1e-5)
        const double inv_vn = rcp_pos(vnorm);

        double dist, gd[3];
        // (the sample loop is warp-uniform; out-of-map samples of real — non-idle — slots are counted)
        sdf_finish<true>(cell, dist, gd[0], gd[1], gd[2], &oob, __double2hiint(wdt) != 0);

        // exp(-(d-d0)/r) (:509/:514) = 2^((d0-d)*log2(e)/r)
        const double ee = exp2_tab(fma(dist, E.neg_log2e_over_r, E.d0_log2e_over_r), xtab);
        const double ev = ee * vnorm;
        Jc = fma(ev, wdt, Jc);  // cost: cd*|v|*dt = alpha*ee*vnorm*dt (:374; alpha applied at the end)
        // sA_a = gd*grad_a*cd*|v| = kA * ee^2 * vnorm * g_a      (:378, extra cd reproduced — H4)
        // sB_a = cd*v_a/|v|       = alpha * ee * v_a / vnorm     (:379)
        const double sAc = kA * (ee * ev);
        const double sBc = E.alpha * (ee * inv_vn);
        double fa[3], fb[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) fa[a] = sAc * gd[a], fb[a] = sBc * vel[a];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          dmma884(G[a][0], G[a][1], fa[a], bAB.x);   // += sA_a(sample) * dt phi_r(sample), straight into G
          dmma884(G[a][0], G[a][1], fb[a], bAB.y);
        }
        cf = it + 1 >= ubl ? coefB : coef;
        if (__builtin_expect(it == uf, 0)) {
          // the first segment is complete: its waypoint rows leave the C fragments (they are combined with the
          // smoothness rows after the loop).  In a mixed iteration record `it` carried the first segment's lanes only;
          // the second segment's lanes go in now through the pair's extra record
          flush_rows(sa), flushed = true;
          if (mixed) {
            const double2 bX = ldg_tab(pl + (size_t)(PL.nit2 - it) * 32);
#pragma unroll
            for (int a = 0; a < 3; ++a) {
              dmma884(G[a][0], G[a][1], fa[a], bX.x);
              dmma884(G[a][0], G[a][1], fb[a], bX.y);
            }
          }
        }
      }
    }
    if (!flushed) flush_rows(sa);
    combine_rows(sa);
    if (hasb) {
      flush_rows(sb);
      combine_rows(sb);
    }
  }
  if (oob_counter != nullptr) {
    oob = __reduce_add_sync(kFull, live ? oob : 0u);
    if (lane == 0 && oob) atomicAdd(oob_counter, (unsigned long long)oob);
  }

  // ---- smoothness cost, rows of the fixed end waypoint (column band of R, :326-327): waypoints m-1 and m, which
  // the ring still holds after the last segment
  if (j < 3) {
    const double* __restrict__ band = pr + PL.off_Rc + (3 * m + j) * 9;  // band entries 0..5 <-> waypoints m-1, m
    const double* wa = ring + ((m - 1) % kRingSlots) * 9;
    const double* wb = ring + (m % kRingSlots) * 9;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      double y = 0.0;
#pragma unroll
      for (int t = 0; t < 3; ++t) y = fma(wa[a * 3 + t], band[t], y);
#pragma unroll
      for (int t = 0; t < 3; ++t) y = fma(wb[a * 3 + t], band[3 + t], y);
      js = fma(y, wb[a * 3 + j], js);
    }
  }
#pragma unroll
  for (int o = 2; o > 0; o >>= 1) {
    js += __shfl_xor_sync(kFull, js, o);
    Jc += __shfl_xor_sync(kFull, Jc, o);
  }
  if (live && j == 0) cost_out[b] = E.ws * js + E.wc * (E.alpha * Jc) + 1e-3;  // :417-418
}

bool eval_mma_applicable(const ProbLayout& PL, int64_t traj_offset, int64_t starts_per_problem) {
  (void)traj_offset;
  // tuning knob: segment-count limit of this kernel (its shared memory no longer depends on m)
  static const int max_m = [] { const char* e = std::getenv("GTOP_MMA_MAX_M"); return e ? std::atoi(e) : 4096; }();  // tuning knob
  return starts_per_problem % 8 == 0 && PL.m <= max_m;
}

cudaError_t launch_eval_mma(const SdfView& S, const ProbLayout& PL, const double* prob, int num_problems,
                            const EvalScalars& E, const double* X, int64_t B, int64_t traj_offset,
                            int64_t starts_per_problem, double* cost, double* grad, unsigned long long* oob_counter,
                            cudaStream_t stream) {
  if (B <= 0) return cudaSuccess;
  // warps per CTA: as many as keep the CTA's shared memory modest (several CTAs per SM)
  const size_t per_warp = sizeof(double) * (size_t)mma_warp_smem_doubles();
  int wpc = kMmaMaxWarps;
  while (wpc > 1 && per_warp * wpc + 512 > 40 * 1024) wpc >>= 1;
  size_t smem = per_warp * wpc + kExpTab * sizeof(double);
  {  // tuning knob: extra dynamic shared memory per CTA limits the resident CTAs per SM
    static const long pad = [] { const char* e = std::getenv("GTOP_MMA_SMEM_PAD"); return e ? std::atol(e) : 0L; }();
    smem += (size_t)pad;
  }
  auto kern = PL.ns_cap <= 64 ? k_eval_mma<true> : k_eval_mma<false>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  const int64_t per_cta = (int64_t)wpc * 8;
  const int64_t grid = ((traj_offset & 7) + B + per_cta - 1) / per_cta;
  kern<<<(unsigned)grid, wpc * 32, smem, stream>>>(S, PL, prob, num_problems, E, X, B, traj_offset, starts_per_problem,
                                                   cost, grad, oob_counter);
  return cudaGetLastError();
}

}  // namespace gtop
