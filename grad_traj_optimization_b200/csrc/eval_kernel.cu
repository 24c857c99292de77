// eval_kernel.cu — batched cost + gradient of GTOP trajectories on sm_100a.
//
// Replaces GradTrajOptimizer::getCostAndGradient (src/grad_traj_optimizer.cpp:281-448 of the
// reference) for a batch of independent trajectories:
//   (a) smoothness  J_s = sum_a d_a' R d_a and 2 Rfp' df + 2 Rpp dp        (:302-336)  — banded R
//   (b) collision   sum over 30 (N_s) samples per segment of cd*|v|*dt     (:345-381)
//       with SDFMap::getDistWithGradTrilinear (src/sdf_map.cpp:185-242) gathers from HBM/L2
//   (c) coefficient map c = L d (getCoefficientFromDerivative :253-279) in per-segment 6x6 blocks
//
// This is the GENERAL kernel (every trajectory may belong to a different problem); multi-start batches
// whose groups of 8 consecutive trajectories share a problem go to k_eval_mma (eval_mma_kernel.cu).
//
// Work decomposition (v4): a QUARTER-WARP (8 lanes) PER TRAJECTORY, four trajectories per warp
// walking their segments in lock-step; there is no block-level barrier anywhere.  Inside a segment
// lane j of the quarter takes samples j, j+8, j+16, ... (4 per lane at N_s = 30), so
//   * the 8 lanes of a quarter gather from 8 consecutive samples (~0.5 m of path: a few 128-B lines
//     of the patch-ordered quad copy of the grid per 256-bit load instruction),
//   * every lane has several independent sample iterations in flight (latency hiding by ILP; the
//     kernel needs ~128 registers, i.e. 16 warps/SM),
//   * the per-segment serial steps — coefficient staging, the cross-lane reduction of the 3x6
//     accumulators U[a][i] = sum_k (sA_ak tau_ik + sB_ak tau'_ik), the projection through the
//     segment's 6x6 block — are paid once per FOUR segments and the reduction spans only 8 lanes
//     (18 exchanges of a transposing shuffle reduction instead of 19 x 5 butterflies).
// The projection through the 6x6 block replaces the reference's dense per-sample products
// (:376-381): O(m N_s) work instead of O(m^2 N_s).
//
// Rounding discipline: everything that feeds the float truncation of position /
// velocity (:457-485) and the voxel index / diff arithmetic (sdf_map.cpp:201-209) is evaluated in
// the reference's operation order with __dmul_rn/__dadd_rn (no FMA contraction), from the same
// host-computed pow() table, so cell selection and the fp32 round trip are bit-identical to the
// CPU path.  Everything downstream is smooth and free to fuse / reorder.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>

#include "eval_common.cuh"
#include "gtop_types.cuh"

namespace gtop {

#ifndef GTOP_EVAL_MINB
#define GTOP_EVAL_MINB 4
#endif
constexpr int kMaxWarpsPerCta = 4;
constexpr int kTrajPerWarp = 4;  // quarter-warps

// One step of the transposing reduction: every lane holds 2*H values; lanes whose `bit` is clear
// keep the low H (and give away the high H), the others the opposite.  After the step each lane
// holds H values, each the sum over the lane pair.
template <int H>
__device__ __forceinline__ void xreduce_step(double (&v)[20], int lane, int bit) {
  const bool hi = (lane & bit) != 0;
#pragma unroll
  for (int i = 0; i < H; ++i) {
    const double send = hi ? v[i] : v[i + H];
    const double keep = hi ? v[i + H] : v[i];
    v[i] = keep + __shfl_xor_sync(kFull, send, bit);
  }
}

// per-trajectory (quarter-warp) shared memory, in doubles:
//   Wd   [3][nrow]   boundary derivatives per axis in waypoint order (nrow = 3(m+1))
//   ga   [3][nrow]   collision gradient accumulated in the same order
//   cb   [32]        c[3][6] + e[3][4] of the current segment
//   rs   [20]        the 19 reduced sums of the current segment
__host__ __device__ inline int traj_smem_doubles(int m) { return 2 * 9 * (m + 1) + 32 + 20; }

// One segment of one trajectory, evaluated by a quarter-warp (lane j of 8): coefficients, the N_s samples (lane j
// takes samples j, j+8, ...), the 8-lane transposing reduction of the 20 sums into rs[], and the back-projection
// of the 3x6 moment block through the segment's 6x6 block: emit(o, value) hands out contribution o = a*6 + r (axis a,
// local derivative r: start derivatives 0..2, end derivatives 3..5), three per lane.  rs[18] / rs[19] then hold the
// segment's sums of cd |v| (and of the velocity / acceleration penalties) before the factor dt.
// Shared by k_eval_traj (a quarter-warp walks the segments of its trajectory) and k_eval_traj_seg (one warp per segment).
template <bool kVelAcc, bool kPlain, class Emit>
__device__ __forceinline__ void segment_pass(const SdfView& S, const ProbLayout& PL, const EvalScalars& E,
                                             const double* __restrict__ pr, int s, const double* Wd, double* cb, double* rs,
                                             const double* xtab, double kA, int lane, int j, unsigned& oob, Emit emit) {
  const int m = PL.m, nrow = 3 * (m + 1), ns_cap = PL.ns_cap;
  // which reduced sums this lane holds after the 8-lane transposing reduction (3 per lane)
  const int red_base = 10 * ((j >> 2) & 1) + 5 * ((j >> 1) & 1) + 3 * (j & 1);
  const int red_cnt = (j & 1) ? 2 : 3;  // odd lanes hold (3,4,pad) of their group of five
  const double* __restrict__ Li = pr + PL.off_Linv + s * 36;
  const double* __restrict__ LiT = pr + PL.off_LinvT + s * 36;
  // ---- coefficients c[a][i] = sum_r Linv[i][r] * dseg[a][r], summed in ascending dense-column
  // order of L (start derivatives first, except for the last segment whose end derivatives live
  // in columns 3..5 of d) with unfused mul/add — the same 6-term sum as `L * d` (:273).
  const bool last = (s == m - 1);
#pragma unroll
  for (int o = j; o < 18; o += 8) {
    const int a = o / 6, i = o % 6;
    const double* w0 = Wd + a * nrow + 3 * s;  // p0,v0,a0,pT,vT,aT contiguous in waypoint order
    double p[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) p[r] = __dmul_rn(LiT[r * 6 + i], w0[r]);
    const double q0 = last ? p[3] : p[0], q1 = last ? p[4] : p[1], q2 = last ? p[5] : p[2];
    const double q3 = last ? p[0] : p[3], q4 = last ? p[1] : p[4], q5 = last ? p[2] : p[5];
    const double c = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(q0, q1), q2), q3), q4), q5);
    cb[o] = c;
    // e_k = k*c_k for the velocity polynomial (2*coeff(s,2), 3*coeff(s,3), ... :477-485)
    if (i >= 2) cb[18 + a * 4 + (i - 2)] = __dmul_rn((double)i, c);
  }
  __syncwarp();

  const int ns = (int)pr[PL.off_ns + s];
  const double dt = pr[PL.off_dt + s];
  const double* __restrict__ tt = pr + PL.off_tt + s * 5 * ns_cap;

  double v[20];  // U[a][i] at a*6+i, cost at 18, velocity/acceleration penalty cost at 19
#pragma unroll
  for (int i = 0; i < 20; ++i) v[i] = 0.0;

  for (int k = j; k < ns; k += 8) {
    const double t = tt[k], t2 = tt[ns_cap + k], t3 = tt[2 * ns_cap + k], t4 = tt[3 * ns_cap + k],
                 t5 = tt[4 * ns_cap + k];
    double pos[3], vel[3];
    float pf[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const double* c = cb + a * 6;
      // getPositionFromCoeff (:457-465): left-to-right, unfused, then float
      double pa = __dadd_rn(c[0], __dmul_rn(c[1], t));
      pa = __dadd_rn(pa, __dmul_rn(c[2], t2));
      pa = __dadd_rn(pa, __dmul_rn(c[3], t3));
      pa = __dadd_rn(pa, __dmul_rn(c[4], t4));
      pa = __dadd_rn(pa, __dmul_rn(c[5], t5));
      pf[a] = __double2float_rn(pa);
      pos[a] = (double)pf[a];
    }
    SdfCell cell;
    sdf_gather<false, kPlain>(S, pos[0], pos[1], pos[2], pf[0], pf[1], pf[2], cell);  // 2 x LDG.256 in flight from here
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const double* c = cb + a * 6;
      const double* e = cb + 18 + a * 4;
      // getVelocityFromCoeff (:477-485)
      double va = __dadd_rn(c[1], __dmul_rn(e[0], t));
      va = __dadd_rn(va, __dmul_rn(e[1], t2));
      va = __dadd_rn(va, __dmul_rn(e[2], t3));
      va = __dadd_rn(va, __dmul_rn(e[3], t4));
      vel[a] = f32_round_trip(va);
    }
    const double vn2 = fma(vel[2], vel[2], fma(vel[1], vel[1], fma(vel[0], vel[0], 1e-300)));
    const double vnorm = sqrt_pos(vn2) + 1e-5;  // :358 (the 1e-300 keeps v = 0 finite; absorbed by 1e-5)
    const double inv_vn = rcp_pos(vnorm);

    double dist, g[3];
    sdf_finish<false>(cell, dist, g[0], g[1], g[2]);  // lane-dependent trip count: no votes here
    oob += cell.in_map ? 0u : 1u;

    // exp(-(d-d0)/r) (:509/:514) = 2^((d0-d)*log2(e)/r)
    const double ee = exp2_tab(fma(dist, E.neg_log2e_over_r, E.d0_log2e_over_r), xtab);
    const double ev = ee * vnorm;
    v[18] += ev;                       // cost: cd*|v| = alpha*ee*vnorm (alpha, dt applied later)
    // sA_a = gd*grad_a*cd*|v| = kA * ee^2 * vnorm * g_a      (:378, extra cd reproduced — H4)
    // sB_a = cd*v_a/|v|       = alpha * ee * v_a / vnorm     (:379)
    const double sAc = kA * (ee * ev);
    const double sBc = E.alpha * (ee * inv_vn);
    const double tp2 = 2.0 * t, tp3 = 3.0 * t2, tp4 = 4.0 * t3, tp5 = 5.0 * t4;  // T*V (:379)
    double sBv[3] = {0.0, 0.0, 0.0}, sCv[3] = {0.0, 0.0, 0.0};
    if (kVelAcc) {
      // getAccelerationFromCoeff (:491-505): 2c2 + 6c3 t + 12c4 t^2 + 20c5 t^3, left to right, then float.
      // 6c3 = 2*(3c3), 12c4 = 3*(4c4), 20c5 = 4*(5c5) round identically from the e_k = k*c_k already held.
      double cv[3], ca[3], gv[3], gacc[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const double* e = cb + 18 + a * 4;
        double aa = __dadd_rn(e[0], __dmul_rn(2.0 * e[1], t));
        aa = __dadd_rn(aa, __dmul_rn(__dmul_rn(3.0, e[2]), t2));
        aa = __dadd_rn(aa, __dmul_rn(4.0 * e[3], t3));
        const double acc = f32_round_trip(aa);
        const double xv = exp2_tab(fma(fabs(vel[a]), E.log2e_over_rv, E.neg_v0_log2e_over_rv), xtab);
        const double xa = exp2_tab(fma(fabs(acc), E.log2e_over_ra, E.neg_a0_log2e_over_ra), xtab);
        cv[a] = E.alpha_v * xv, gv[a] = E.alpha_v_over_r * xv;    // :517-525
        ca[a] = E.alpha_a * xa, gacc[a] = E.alpha_a_over_r * xa;  // :527-535
      }
      v[19] += ((cv[0] + cv[1]) + cv[2] + (ca[0] + ca[1]) + ca[2]) * vnorm;  // cost_vel + cost_acc (:388-392)
      const double carried = (cv[2] + ca[2]) * inv_vn;  // `cv`, `ca` keep their k = 2 values in the second loop
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        // (gv*|v| + cv*v_a/|v| + ca*v_a/|v|) T V Ldp  and  ga*|v| T V V Ldp   (:396-404), weights wv = wa = 1
        sBv[a] = E.inv_wc * fma(gv[a], vnorm, carried * vel[a]);
        sCv[a] = E.inv_wc * (gacc[a] * vnorm);
      }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const double sA = sAc * g[a], sB = sBc * vel[a] + sBv[a];
      v[a * 6 + 0] += sA;
      v[a * 6 + 1] += fma(sA, t, sB);
      v[a * 6 + 2] += fma(sA, t2, sB * tp2);
      v[a * 6 + 3] += fma(sA, t3, sB * tp3);
      v[a * 6 + 4] += fma(sA, t4, sB * tp4);
      v[a * 6 + 5] += fma(sA, t5, sB * tp5);
      if (kVelAcc) {  // T*V*V = (0, 0, 2, 6t, 12t^2, 20t^3)
        v[a * 6 + 2] = fma(sCv[a], 2.0, v[a * 6 + 2]);
        v[a * 6 + 3] = fma(sCv[a], 3.0 * tp2, v[a * 6 + 3]);
        v[a * 6 + 4] = fma(sCv[a], 4.0 * tp3, v[a * 6 + 4]);
        v[a * 6 + 5] = fma(sCv[a], 5.0 * tp4, v[a * 6 + 5]);
      }
    }
  }

  // ---- transposing reduction inside the quarter: 20 -> 10 -> 5 -> 3 values per lane ----------
  xreduce_step<10>(v, lane, 4);
  xreduce_step<5>(v, lane, 2);
  v[5] = 0.0;
  xreduce_step<3>(v, lane, 1);
#pragma unroll
  for (int i = 0; i < 3; ++i)
    if (i < red_cnt && red_base + i < 20) rs[red_base + i] = v[i];
  __syncwarp();
  // ---- back-projection through the segment block: g[a][r] = sum_i U[a][i]*Linv[i][r] (T*Ldp, :378-379)
#pragma unroll
  for (int o = j; o < 18; o += 8) {
    const int a = o / 6, r = o % 6;
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < 6; ++i) acc = fma(rs[a * 6 + i], Li[i * 6 + r], acc);
    emit(o, __dmul_rn(acc, dt));  // (rounded here: the caller's `+=` must not fuse it, or the two kernels below would differ)
  }
  __syncwarp();
}

// kVelAcc: also evaluate the velocity / acceleration penalty block that upstream HEAD has commented out
// (:383-407, penalties :517-535) — gtop_params.vel_acc_penalty, default off; literal restatement including
// the `cv` / `ca` values carried from the k = 2 iteration of the cost loop into the gradient loop.
// kPlain: the map keeps no quad copy; corners come straight from the distance field (eval_common.cuh).
template <bool kVelAcc, bool kPlain>
__global__ void __launch_bounds__(kMaxWarpsPerCta * 32, GTOP_EVAL_MINB)
k_eval_traj(SdfView S, ProbLayout PL, const double* __restrict__ prob, int num_problems, EvalScalars E,
            const double* __restrict__ X, int64_t B, int64_t traj_offset, int64_t starts_per_problem,
            double* __restrict__ cost_out, double* __restrict__ grad_out, unsigned long long* __restrict__ oob_counter) {
  extern __shared__ double smem[];
  const int m = PL.m, P = PL.P, n = PL.n, nrow = 3 * (m + 1);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
  const int q = lane >> 3, j = lane & 7;  // trajectory slot of the warp, lane within the quarter

  // 2^(j/16) table for exp2_tab: the only block-level synchronisation of the kernel
  double* xtab = smem + wpc * kTrajPerWarp * traj_smem_doubles(m);
  exp2_tab_init(xtab, threadIdx.x, blockDim.x);
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // programmatic dependent launch, see k_eval_traj_seg
  asm volatile("griddepcontrol.wait;" ::: "memory");
  __syncthreads();

  const int64_t b_raw = ((int64_t)blockIdx.x * wpc + warp) * kTrajPerWarp + q;
  if (b_raw - q >= B) return;              // the whole warp is past the end
  const bool live = b_raw < B;             // ragged tail: dead quarters shadow the last trajectory
  const int64_t b = live ? b_raw : B - 1;  // (same control flow, outputs suppressed)

  double* Wd = smem + (warp * kTrajPerWarp + q) * traj_smem_doubles(m);
  double* ga = Wd + 3 * nrow;
  double* cb = ga + 3 * nrow;
  double* rs = cb + 32;

  const int64_t pid = ((traj_offset + b) / starts_per_problem) % num_problems;
  const double* __restrict__ pr = prob + pid * (int64_t)PL.stride;
  const double* __restrict__ x = X + b * n;

  // ---- 1. boundary derivatives in waypoint order; clear the gradient accumulators ---------------
  for (int jj = j; jj < P; jj += 8) {
#pragma unroll
    for (int a = 0; a < 3; ++a) Wd[a * nrow + 3 + jj] = x[a * P + jj];
  }
  for (int o = j; o < 18; o += 8) {
    const int a = o / 6, r = o % 6;
    Wd[a * nrow + (r < 3 ? r : 3 * m + (r - 3))] = pr[PL.off_Df + o];
  }
  for (int jj = j; jj < 3 * nrow; jj += 8) ga[jj] = 0.0;
  __syncwarp();

  double Jc_total = 0.0, Jva_total = 0.0;
  unsigned oob = 0;
  const double kA = -E.alpha * E.alpha_over_r * S.res_inv;  // gd*cd = -(alpha^2/r) e^2; grad scale 1/res

  if (!E.skip_collision) {
    for (int s = 0; s < m; ++s) {
      segment_pass<kVelAcc, kPlain>(S, PL, E, pr, s, Wd, cb, rs, xtab, kA, lane, j, oob,
                                    [&](int o, double val) { ga[(o / 6) * nrow + 3 * s + o % 6] += val; });
      const double dt = pr[PL.off_dt + s];
      Jc_total = fma(rs[18], dt, Jc_total);
      if (kVelAcc) Jva_total = fma(rs[19], dt, Jva_total);
      __syncwarp();
    }
  }
  if (oob_counter != nullptr) {
    oob = __reduce_add_sync(kFull, live ? oob : 0u);
    if (lane == 0 && oob) atomicAdd(oob_counter, (unsigned long long)oob);
  }

  // ---- 3. smoothness cost: banded column sums of R in waypoint order (:326-327) -------------------
  double js = 0.0;
  for (int u = j; u < nrow; u += 8) {
    const double* __restrict__ band = pr + PL.off_Rc + u * 9;
    const int base = 3 * (u / 3 - 1);
    const int j0 = base < 0 ? 3 : 0, j1 = (base + 9 > nrow) ? 6 : 9;
    double y0 = 0.0, y1 = 0.0, y2 = 0.0;
    for (int jj = j0; jj < j1; ++jj) {
      const double bv = band[jj];
      y0 = fma(Wd[base + jj], bv, y0);
      y1 = fma(Wd[nrow + base + jj], bv, y1);
      y2 = fma(Wd[2 * nrow + base + jj], bv, y2);
    }
    js = fma(y0, Wd[u], js);
    js = fma(y1, Wd[nrow + u], js);
    js = fma(y2, Wd[2 * nrow + u], js);
  }
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) js += __shfl_xor_sync(kFull, js, o);

  // ---- 4. outputs (:417-431) ----------------------------------------------------------------------
  if (live) {
    double* __restrict__ gout = grad_out + b * n;
    for (int jj = j; jj < P; jj += 8) {
      const double* __restrict__ band = pr + PL.off_Rg + jj * 9;  // 2 Rfp' df + 2 Rpp dp row (:330-332)
      const double* Wb = Wd + (jj - jj % 3);                      // 3*(wp-1), wp = jj/3 + 1
      double g0 = 0.0, g1 = 0.0, g2 = 0.0;
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const double bv = band[t];
        g0 = fma(Wb[t], bv, g0);
        g1 = fma(Wb[nrow + t], bv, g1);
        g2 = fma(Wb[2 * nrow + t], bv, g2);
      }
      gout[jj] = (E.ws * (2.0 * g0) + E.wc * ga[3 + jj]) + 1e-5;
      gout[P + jj] = (E.ws * (2.0 * g1) + E.wc * ga[nrow + 3 + jj]) + 1e-5;
      gout[2 * P + jj] = (E.ws * (2.0 * g2) + E.wc * ga[2 * nrow + 3 + jj]) + 1e-5;
    }
    // wv = wa = 1 (:412)
    if (j == 0) cost_out[b] = E.ws * js + E.wc * (E.alpha * Jc_total) + (kVelAcc ? Jva_total : 0.0) + 1e-3;
  }
}

// ---- the same evaluation, segments in parallel: small batches ---------------------------------------------------------
// k_eval_traj gives a trajectory to a quarter-warp, which walks its m segments one after the other: a batch of B
// trajectories is B/4 warps of m dependent segment passes — 256 warps for the 1,024 seven-waypoint problems of
// BASELINE configs[1], on 148 SMs.  Here a CTA owns four trajectories (the four quarters of a warp) and its warps take
// the SEGMENTS: each runs the same segment_pass as k_eval_traj, so every per-segment number is the same bit for bit;
// the contributions of the two segments that meet in a waypoint are added (two terms: the order does not matter) and
// the per-segment cost sums are folded in segment order by the epilogue exactly as the sequential kernel folds them.
// Results are therefore IDENTICAL to k_eval_traj's, and launch_eval may choose between the two by batch size.
//   shared memory, in doubles: Wd [4][3 nrow] | slot [4][m][18] | jseg [4][m][2] | cb [W][4][32] | rs [W][4][20] | xtab
constexpr int kSegMaxWarps = 16;
__host__ __device__ inline int seg_smem_doubles(int m, int W) {
  return 4 * 9 * (m + 1) + 4 * m * 18 + 4 * m * 2 + W * 4 * 52 + kExpTab;
}

template <bool kVelAcc, bool kPlain>
__global__ void __launch_bounds__(kSegMaxWarps * 32, 1)
k_eval_traj_seg(SdfView S, ProbLayout PL, const double* __restrict__ prob, int num_problems, EvalScalars E,
                const double* __restrict__ X, int64_t B, int64_t traj_offset, int64_t starts_per_problem,
                double* __restrict__ cost_out, double* __restrict__ grad_out, unsigned long long* __restrict__ oob_counter) {
  extern __shared__ double smem[];
  const int m = PL.m, P = PL.P, n = PL.n, nrow = 3 * (m + 1);
  const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5, W = nthr >> 5;
  const int q = lane >> 3, j = lane & 7;  // trajectory slot of the CTA, lane within the quarter
  double* Wd_all = smem;
  double* slot_all = Wd_all + 4 * 3 * nrow;
  double* jseg_all = slot_all + 4 * m * 18;
  double* cb = jseg_all + 4 * m * 2 + (warp * 4 + q) * 52;
  double* rs = cb + 32;
  double* xtab = jseg_all + 4 * m * 2 + W * 4 * 52;

  auto traj_of = [&](int qq) {  // ragged tail: dead quarters shadow the last trajectory (outputs suppressed)
    const int64_t b = (int64_t)blockIdx.x * 4 + qq;
    return b < B ? b : B - 1;
  };
  auto problem_of = [&](int64_t b) { return prob + (((traj_offset + b) / starts_per_problem) % num_problems) * (int64_t)PL.stride; };

  // ---- 1. boundary derivatives of the four trajectories in waypoint order ------------------------------------
  exp2_tab_init(xtab, tid, nthr);
  // programmatic dependent launch: the next kernel of the stream may take free SM slots now; X and the problem records
  // may come from the preceding kernel and are only read after the wait (worth ~2 us of a 15 us launch)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  for (int i = tid; i < 4 * n; i += nthr) {
    const int qq = i / n, k = i - qq * n, a = k / P, jj = k - a * P;
    Wd_all[qq * 3 * nrow + a * nrow + 3 + jj] = X[traj_of(qq) * n + k];
  }
  for (int i = tid; i < 4 * 18; i += nthr) {
    const int qq = i / 18, o = i - 18 * qq, a = o / 6, r = o % 6;
    Wd_all[qq * 3 * nrow + a * nrow + (r < 3 ? r : 3 * m + (r - 3))] = problem_of(traj_of(qq))[PL.off_Df + o];
  }
  __syncthreads();

  const int64_t b = traj_of(q);
  const bool live = (int64_t)blockIdx.x * 4 + q < B;
  const double* __restrict__ pr = problem_of(b);
  const double* Wd = Wd_all + q * 3 * nrow;
  double* slot = slot_all + q * m * 18;
  double* jseg = jseg_all + q * m * 2;

  // ---- 2. one segment per warp (and round) --------------------------------------------------------------------
  if (!E.skip_collision) {
    const double kA = -E.alpha * E.alpha_over_r * S.res_inv;  // gd*cd = -(alpha^2/r) e^2; grad scale 1/res
    unsigned oob = 0;
    for (int s = warp; s < m; s += W) {
      segment_pass<kVelAcc, kPlain>(S, PL, E, pr, s, Wd, cb, rs, xtab, kA, lane, j, oob,
                                    [&](int o, double val) { slot[s * 18 + o] = val; });
      if (j == 0) jseg[2 * s] = rs[18], jseg[2 * s + 1] = rs[19];
      __syncwarp();
    }
    if (oob_counter != nullptr) {
      oob = __reduce_add_sync(kFull, live ? oob : 0u);
      if (lane == 0 && oob) atomicAdd(oob_counter, (unsigned long long)oob);
    }
  }
  __syncthreads();

  // ---- 3. smoothness cost and the cost output: warp 0, one quarter per trajectory, the sequential kernel's code ----
  if (warp == 0) {
    double js = 0.0;
    for (int u = j; u < nrow; u += 8) {
      const double* __restrict__ band = pr + PL.off_Rc + u * 9;
      const int base = 3 * (u / 3 - 1);
      const int j0 = base < 0 ? 3 : 0, j1 = (base + 9 > nrow) ? 6 : 9;
      double y0 = 0.0, y1 = 0.0, y2 = 0.0;
      for (int jj = j0; jj < j1; ++jj) {
        const double bv = band[jj];
        y0 = fma(Wd[base + jj], bv, y0);
        y1 = fma(Wd[nrow + base + jj], bv, y1);
        y2 = fma(Wd[2 * nrow + base + jj], bv, y2);
      }
      js = fma(y0, Wd[u], js);
      js = fma(y1, Wd[nrow + u], js);
      js = fma(y2, Wd[2 * nrow + u], js);
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) js += __shfl_xor_sync(kFull, js, o);
    if (live && j == 0) {
      double Jc_total = 0.0, Jva_total = 0.0;
      if (!E.skip_collision)
        for (int s = 0; s < m; ++s) {  // folded in segment order, as k_eval_traj does
          const double dt = pr[PL.off_dt + s];
          Jc_total = fma(jseg[2 * s], dt, Jc_total);
          if (kVelAcc) Jva_total = fma(jseg[2 * s + 1], dt, Jva_total);
        }
      cost_out[b] = E.ws * js + E.wc * (E.alpha * Jc_total) + (kVelAcc ? Jva_total : 0.0) + 1e-3;  // wv = wa = 1 (:412)
    }
  }

  // ---- 4. gradient rows (:425-431): all threads, one (trajectory, row) each ------------------------------------
  for (int i = tid; i < 4 * P; i += nthr) {
    const int qq = i / P, jj = i - qq * P;
    if ((int64_t)blockIdx.x * 4 + qq >= B) continue;
    const int64_t bb = (int64_t)blockIdx.x * 4 + qq;
    const double* __restrict__ prq = problem_of(bb);
    const double* __restrict__ band = prq + PL.off_Rg + jj * 9;  // 2 Rfp' df + 2 Rpp dp row (:330-332)
    const double* Wb = Wd_all + qq * 3 * nrow + (jj - jj % 3);   // 3*(wp-1), wp = jj/3 + 1
    double g0 = 0.0, g1 = 0.0, g2 = 0.0;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const double bv = band[t];
      g0 = fma(Wb[t], bv, g0);
      g1 = fma(Wb[nrow + t], bv, g1);
      g2 = fma(Wb[2 * nrow + t], bv, g2);
    }
    // row 3 + jj of the waypoint order = start derivative r of segment s = (3 + jj) / 3, end derivative of s - 1
    const int s = (3 + jj) / 3, r = (3 + jj) - 3 * s;
    const double* sl = slot_all + qq * m * 18;
    double c[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) c[a] = E.skip_collision ? 0.0 : sl[(s - 1) * 18 + a * 6 + 3 + r] + sl[s * 18 + a * 6 + r];
    double* __restrict__ gout = grad_out + bb * n;
    gout[jj] = (E.ws * (2.0 * g0) + E.wc * c[0]) + 1e-5;
    gout[P + jj] = (E.ws * (2.0 * g1) + E.wc * c[1]) + 1e-5;
    gout[2 * P + jj] = (E.ws * (2.0 * g2) + E.wc * c[2]) + 1e-5;
  }
}

// warps per CTA of the segment-parallel kernel: one per segment while a single wave holds the whole batch (latency),
// else the count that keeps the most warps of an SM on segments (16 resident warps of 128 registers)
static int seg_warps(int m, int64_t ctas) {
  static const int force = [] { const char* e = std::getenv("GTOP_SEG_WARPS"); return e ? std::atoi(e) : 0; }();  // tuning knob
  if (force > 0) return std::min(std::min(force, kSegMaxWarps), m);
  if (m <= kSegMaxWarps && ctas <= (int64_t)148 * (16 / m)) return m;
  int best = 1;
  double best_busy = -1.0;
  for (int W = 1; W <= kSegMaxWarps && W <= m; ++W) {
    const int rounds = (m + W - 1) / W;
    const double busy = (double)(16 / W) * m / rounds - 1e-3 * rounds;
    if (busy > best_busy) best_busy = busy, best = W;
  }
  return best;
}

bool eval_mma_applicable(const ProbLayout& PL, int64_t B, int num_problems, int64_t starts_per_problem);
cudaError_t launch_eval_mma(const SdfView& S, const ProbLayout& PL, const double* prob, int num_problems,
                            const EvalScalars& E, const double* X, int64_t B, int64_t traj_offset,
                            int64_t starts_per_problem, double* cost, double* grad, unsigned long long* oob_counter,
                            cudaStream_t stream);

// GTOP_EVAL_KERNEL=general forces the quarter-warp kernel (tests exercise both); default: the DMMA
// multi-start kernel whenever groups of 8 consecutive trajectories share a problem.
static bool force_general() {
  static const bool v = [] {
    const char* e = std::getenv("GTOP_EVAL_KERNEL");
    return e != nullptr && std::strcmp(e, "general") == 0;
  }();
  return v;
}

// launch with programmatic stream serialization (the kernels above call griddepcontrol.wait before their first read of
// anything a preceding kernel may have written); GTOP_PDL=0 turns it off
template <class Kern, class... Args>
static cudaError_t launch_pdl(Kern kern, unsigned grid, unsigned block, size_t smem, cudaStream_t stream, Args... args) {
  static const bool pdl = [] { const char* e = std::getenv("GTOP_PDL"); return !e || std::atoi(e) != 0; }();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, args...);
}

// would launch_eval take the multi-start kernel (whose x / gradient blocks move by bulk copies)?
bool eval_takes_mma_kernel(const ProbLayout& PL, const EvalScalars& E, int64_t B, int num_problems, int64_t starts_per_problem) {
  return !force_general() && !E.vel_acc && eval_mma_applicable(PL, B, num_problems, starts_per_problem);
}

cudaError_t launch_eval(const SdfView& S, const ProbLayout& PL, const double* prob, int num_problems,
                        const EvalScalars& E, const double* X, int64_t B, int64_t traj_offset,
                        int64_t starts_per_problem, double* cost, double* grad, unsigned long long* oob_counter,
                        int max_ctas, cudaStream_t stream) {
  (void)max_ctas;
  if (B <= 0) return cudaSuccess;
  if (!force_general() && !E.vel_acc && eval_mma_applicable(PL, B, num_problems, starts_per_problem))
    return launch_eval_mma(S, PL, prob, num_problems, E, X, B, traj_offset, starts_per_problem, cost, grad, oob_counter,
                           stream);
  // small batches: segments in parallel (same bits, see k_eval_traj_seg)
  static const int64_t seg_below = [] { const char* e = std::getenv("GTOP_SEG_BELOW"); return e ? std::atol(e) : 16384L; }();  // tuning knob
  if (B < seg_below && PL.m >= 2) {
    const int64_t ctas = (B + 3) / 4;
    const int W = seg_warps(PL.m, ctas);
    const size_t smem = sizeof(double) * (size_t)seg_smem_doubles(PL.m, W);
    if (smem <= 200 * 1024) {
      const bool plain = S.quad == nullptr;
      auto kern = E.vel_acc ? (plain ? k_eval_traj_seg<true, true> : k_eval_traj_seg<true, false>)
                            : (plain ? k_eval_traj_seg<false, true> : k_eval_traj_seg<false, false>);
      if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
      }
      return launch_pdl(kern, (unsigned)ctas, W * 32, smem, stream, S, PL, prob, num_problems, E, X, B, traj_offset,
                        starts_per_problem, cost, grad, oob_counter);
    }
  }
  // warps per CTA: as many as keep the CTA's shared memory modest (4 CTAs/SM at 16 warps/SM)
  const size_t per_warp = sizeof(double) * (size_t)traj_smem_doubles(PL.m) * kTrajPerWarp;
  int wpc = kMaxWarpsPerCta;
  while (wpc > 1 && per_warp * wpc + 512 > 48 * 1024) wpc >>= 1;
  const size_t smem = per_warp * wpc + kExpTab * sizeof(double);
  const bool plain = S.quad == nullptr;
  auto kern = E.vel_acc ? (plain ? k_eval_traj<true, true> : k_eval_traj<true, false>)
                        : (plain ? k_eval_traj<false, true> : k_eval_traj<false, false>);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  const int64_t per_cta = (int64_t)wpc * kTrajPerWarp;
  const int64_t grid = (B + per_cta - 1) / per_cta;
  return launch_pdl(kern, (unsigned)grid, wpc * 32, smem, stream, S, PL, prob, num_problems, E, X, B, traj_offset,
                    starts_per_problem, cost, grad, oob_counter);
}

}  // namespace gtop
