// eval_kernel.cu — batched cost + gradient of GTOP trajectories on sm_100a.
//
// Replaces GradTrajOptimizer::getCostAndGradient (src/grad_traj_optimizer.cpp:281-448 of the
// reference) for a batch of independent trajectories:
//   (a) smoothness  J_s = sum_a d_a' R d_a and 2 Rfp' df + 2 Rpp dp        (:302-336)  — banded R
//   (b) collision   sum over 30 (N_s) samples per segment of cd*|v|*dt     (:345-381)
//       with SDFMap::getDistWithGradTrilinear (src/sdf_map.cpp:185-242) gathers from HBM/L2
//   (c) coefficient map c = L d (getCoefficientFromDerivative :253-279) in per-segment 6x6 blocks
//
// Work decomposition (v2): ONE WARP PER TRAJECTORY, no block-level barrier anywhere.  The warp walks
// the trajectory's segments; inside a segment one lane = one sample (30 of 32 lanes busy), so the
// lanes of a gather instruction hit neighbouring voxels (samples are 1/30 of a segment apart: a
// handful of 128-B lines of the z-fastest grid per load instruction).  Per-sample work stays in
// registers; the 3x6 per-segment accumulators U[a][i] = sum_k (sA_ak tau_ik + sB_ak tau'_ik) and
// the cost are reduced across the warp with a transposing shuffle reduction (21 exchanges for 19
// values instead of 19 x 5 butterflies) and projected once through the segment's 6x6 block —
// O(m N_s) instead of the reference's dense O(m^2 N_s) products (:376-381).
//
// Rounding discipline: everything that feeds the float truncation of position /
// velocity (:457-485) and the voxel index / diff arithmetic (sdf_map.cpp:201-209) is evaluated in
// the reference's operation order with __dmul_rn/__dadd_rn (no FMA contraction), from the same
// host-computed pow() table, so cell selection and the fp32 round trip are bit-identical to the
// CPU path.  Everything downstream is smooth and free to fuse / reorder.
#include <cuda_runtime.h>

#include <cstdint>

#include "gtop_types.cuh"

namespace gtop {

constexpr int kWarpsPerCta = 4;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ double f32_round_trip(double x) { return (double)__double2float_rn(x); }

// SDFMap::getDistWithGradTrilinear (sdf_map.cpp:185-242).  Out of map: dist=-1, grad=0 (H3).
// The gradient is returned WITHOUT the 1/resolution factor (folded into the caller's constants).
// Interpolation is evaluated as nested lerps a + d*(b-a): the corner differences are shared between
// the value and the gradient (25 FP64 ops instead of 46); it differs from the reference's
// (1-d)*a + d*b form only by rounding (~1e-16 relative).
__device__ __forceinline__ bool sdf_trilinear(const SdfView& S, double px, double py, double pz, double& dist,
                                              double& gx, double& gy, double& gz) {
  if (px < S.lox || py < S.loy || pz < S.loz || px > S.hix || py > S.hiy || pz > S.hiz) {
    dist = -1.0;
    gx = gy = gz = 0.0;
    return false;
  }
  // pos_m = pos - 0.5*res; idx = floor((pos_m - origin)*res_inv)           (:201-204, :71-74)
  const double fx = floor(__dmul_rn(__dsub_rn(__dsub_rn(px, S.half_res), S.ox), S.res_inv));
  const double fy = floor(__dmul_rn(__dsub_rn(__dsub_rn(py, S.half_res), S.oy), S.res_inv));
  const double fz = floor(__dmul_rn(__dsub_rn(__dsub_rn(pz, S.half_res), S.oz), S.res_inv));
  // idx_pos = (idx + 0.5)*res + origin; diff = (pos - idx_pos)*res_inv     (:76-78, :209)
  const double dx = __dmul_rn(__dsub_rn(px, __dadd_rn(__dmul_rn(__dadd_rn(fx, 0.5), S.res), S.ox)), S.res_inv);
  const double dy = __dmul_rn(__dsub_rn(py, __dadd_rn(__dmul_rn(__dadd_rn(fy, 0.5), S.res), S.oy)), S.res_inv);
  const double dz = __dmul_rn(__dsub_rn(pz, __dadd_rn(__dmul_rn(__dadd_rn(fz, 0.5), S.res), S.oz)), S.res_inv);
  const int ix = (int)fx, iy = (int)fy, iz = (int)fz;
  // clamped corner indices (getDistance(Vector3i) :166-174)
  const int x0 = max(min(ix, S.nx - 1), 0), x1 = max(min(ix + 1, S.nx - 1), 0);
  const int y0 = max(min(iy, S.ny - 1), 0), y1 = max(min(iy + 1, S.ny - 1), 0);
  const int z0 = max(min(iz, S.nz - 1), 0), z1 = max(min(iz + 1, S.nz - 1), 0);
  const int syz = S.ny * S.nz;
  const double* __restrict__ D = S.dist;
  const int o00 = x0 * syz + y0 * S.nz, o01 = x0 * syz + y1 * S.nz;  // nvox < 2^31 checked on the host
  const int o10 = x1 * syz + y0 * S.nz, o11 = x1 * syz + y1 * S.nz;
  const double v000 = __ldg(D + o00 + z0), v001 = __ldg(D + o00 + z1);
  const double v010 = __ldg(D + o01 + z0), v011 = __ldg(D + o01 + z1);
  const double v100 = __ldg(D + o10 + z0), v101 = __ldg(D + o10 + z1);
  const double v110 = __ldg(D + o11 + z0), v111 = __ldg(D + o11 + z1);
  // x-lerps; the x-differences feed grad x
  const double e00 = v100 - v000, e01 = v101 - v001, e10 = v110 - v010, e11 = v111 - v011;
  const double v00 = fma(dx, e00, v000), v01 = fma(dx, e01, v001);
  const double v10 = fma(dx, e10, v010), v11 = fma(dx, e11, v011);
  // y-lerps; y-differences feed grad y
  const double f0 = v10 - v00, f1 = v11 - v01;
  const double v0 = fma(dy, f0, v00), v1 = fma(dy, f1, v01);
  const double h = v1 - v0;
  dist = fma(dz, h, v0);                 // :229
  gz = h;                                // :231
  gy = fma(dz, f1 - f0, f0);             // :232-233
  const double ex0 = fma(dy, e10 - e00, e00), ex1 = fma(dy, e11 - e01, e01);
  gx = fma(dz, ex1 - ex0, ex0);          // :234-239
  return true;
}

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

// per-warp shared memory (doubles):
//   Wd   [3][nrow]   boundary derivatives per axis in waypoint order (nrow = 3(m+1))
//   ga   [3][nrow]   collision gradient accumulated in the same order
//   cb   [32]        staging: c[3][6] + e[3][4] per segment; then the 19 reduced sums
__host__ __device__ inline int warp_smem_doubles(int m) { return 2 * 9 * (m + 1) + 32; }

__global__ void __launch_bounds__(kWarpsPerCta * 32)
k_eval_traj(SdfView S, ProbLayout PL, const double* __restrict__ prob, int num_problems, EvalScalars E,
            const double* __restrict__ X, int64_t B, int64_t traj_offset, int64_t starts_per_problem,
            double* __restrict__ cost_out, double* __restrict__ grad_out, unsigned long long* __restrict__ oob_counter) {
  extern __shared__ double smem[];
  const int m = PL.m, P = PL.P, n = PL.n, nrow = 3 * (m + 1);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* Wd = smem + warp * warp_smem_doubles(m);
  double* ga = Wd + 3 * nrow;
  double* cb = ga + 3 * nrow;

  const int64_t b = (int64_t)blockIdx.x * kWarpsPerCta + warp;
  if (b >= B) return;  // whole warp exits; there is no block-level synchronisation in this kernel

  const int64_t pid = ((traj_offset + b) / starts_per_problem) % num_problems;
  const double* __restrict__ pr = prob + pid * (int64_t)PL.stride;
  const double* __restrict__ x = X + b * n;

  // ---- 1. boundary derivatives in waypoint order; clear the gradient accumulators ---------------
  for (int j = lane; j < P; j += 32) {
#pragma unroll
    for (int a = 0; a < 3; ++a) Wd[a * nrow + 3 + j] = x[a * P + j];
  }
  if (lane < 18) {
    const int a = lane / 6, q = lane % 6;
    Wd[a * nrow + (q < 3 ? q : 3 * m + (q - 3))] = pr[PL.off_Df + lane];
  }
  for (int j = lane; j < 3 * nrow; j += 32) ga[j] = 0.0;
  __syncwarp();

  // lane roles for the per-segment 18-wide steps: (axis, index)
  const int la = (lane < 18) ? lane / 6 : 0, li = (lane < 18) ? lane % 6 : 0;
  // which of the 19 reduced sums this lane ends up holding after the transposing reduction
  const int b2 = (lane >> 2) & 1, b1 = (lane >> 1) & 1, b0 = lane & 1;
  const int sub = 3 * b2 + 2 * b1 + b0;
  const bool red_valid = !((b1 & b0) | (b2 & b1));          // sub in {0,1,2,3,4}
  const int red_idx = 10 * ((lane >> 4) & 1) + 5 * ((lane >> 3) & 1) + sub;

  double Jc_total = 0.0;
  unsigned oob = 0;
  const double kA = -E.alpha * E.alpha_over_r * S.res_inv;  // gd*cd = -(alpha^2/r) e^2; grad scale 1/res
  const int ns_cap = PL.ns_cap;

  if (!E.skip_collision) {
    for (int s = 0; s < m; ++s) {
      const double* __restrict__ Li = pr + PL.off_Linv + s * 36;
      // ---- coefficients c[a][i] = sum_r Linv[i][r] * dseg[a][r], summed in ascending dense-column
      // order of L (start derivatives first, except for the last segment whose end derivatives live
      // in columns 3..5 of d) with unfused mul/add — the same 6-term sum as `L * d` (:273).
      if (lane < 18) {
        const double* w0 = Wd + la * nrow + 3 * s;  // p0,v0,a0,pT,vT,aT contiguous in waypoint order
        double p[6];
#pragma unroll
        for (int r = 0; r < 6; ++r) p[r] = __dmul_rn(Li[li * 6 + r], w0[r]);
        const bool last = (s == m - 1);
        const double q0 = last ? p[3] : p[0], q1 = last ? p[4] : p[1], q2 = last ? p[5] : p[2];
        const double q3 = last ? p[0] : p[3], q4 = last ? p[1] : p[4], q5 = last ? p[2] : p[5];
        const double c = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(q0, q1), q2), q3), q4), q5);
        cb[lane] = c;
        // e_k = k*c_k for the velocity polynomial (2*coeff(s,2), 3*coeff(s,3), ... :477-485)
        if (li >= 2) cb[18 + la * 4 + (li - 2)] = __dmul_rn((double)li, c);
      }
      __syncwarp();

      const int ns = (int)pr[PL.off_ns + s];
      const double dt = pr[PL.off_dt + s];
      const double* __restrict__ tt = pr + PL.off_tt + s * 5 * ns_cap;

      double v[20];  // U[a][i] at a*6+i, cost at 18, pad at 19
#pragma unroll
      for (int i = 0; i < 20; ++i) v[i] = 0.0;

      for (int k = lane; k < ns; k += 32) {
        const double t = tt[k], t2 = tt[ns_cap + k], t3 = tt[2 * ns_cap + k], t4 = tt[3 * ns_cap + k],
                     t5 = tt[4 * ns_cap + k];
        double pos[3], vel[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          const double* c = cb + a * 6;
          const double* e = cb + 18 + a * 4;
          // getPositionFromCoeff (:457-465): left-to-right, unfused, then float
          double pa = __dadd_rn(c[0], __dmul_rn(c[1], t));
          pa = __dadd_rn(pa, __dmul_rn(c[2], t2));
          pa = __dadd_rn(pa, __dmul_rn(c[3], t3));
          pa = __dadd_rn(pa, __dmul_rn(c[4], t4));
          pa = __dadd_rn(pa, __dmul_rn(c[5], t5));
          pos[a] = f32_round_trip(pa);
          // getVelocityFromCoeff (:477-485)
          double va = __dadd_rn(c[1], __dmul_rn(e[0], t));
          va = __dadd_rn(va, __dmul_rn(e[1], t2));
          va = __dadd_rn(va, __dmul_rn(e[2], t3));
          va = __dadd_rn(va, __dmul_rn(e[3], t4));
          vel[a] = f32_round_trip(va);
        }
        const double vn2 = fma(vel[2], vel[2], fma(vel[1], vel[1], vel[0] * vel[0]));
        const double vnorm = sqrt(vn2) + 1e-5;  // :358

        double dist, g[3];
        if (!sdf_trilinear(S, pos[0], pos[1], pos[2], dist, g[0], g[1], g[2])) ++oob;

        const double ee = exp((E.d0 - dist) * E.inv_r);  // exp(-(d-d0)/r), :509/:514
        const double ev = ee * vnorm;
        v[18] += ev;                       // cost: cd*|v| = alpha*ee*vnorm (alpha, dt applied later)
        // sA_a = gd*grad_a*cd*|v| = kA * ee^2 * vnorm * g_a      (:378, extra cd reproduced — H4)
        // sB_a = cd*v_a/|v|       = alpha * ee * v_a / vnorm     (:379)
        const double sAc = kA * (ee * ev);
        const double sBc = E.alpha * (ee / vnorm);
        const double tp2 = 2.0 * t, tp3 = 3.0 * t2, tp4 = 4.0 * t3, tp5 = 5.0 * t4;  // T*V (:379)
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          const double sA = sAc * g[a], sB = sBc * vel[a];
          v[a * 6 + 0] += sA;
          v[a * 6 + 1] += fma(sA, t, sB);
          v[a * 6 + 2] += fma(sA, t2, sB * tp2);
          v[a * 6 + 3] += fma(sA, t3, sB * tp3);
          v[a * 6 + 4] += fma(sA, t4, sB * tp4);
          v[a * 6 + 5] += fma(sA, t5, sB * tp5);
        }
      }

      // ---- transposing warp reduction: 20 -> 10 -> 5 -> 3 -> 2 -> 1 values per lane ---------------
      xreduce_step<10>(v, lane, 16);
      xreduce_step<5>(v, lane, 8);
      v[5] = 0.0;
      xreduce_step<3>(v, lane, 4);
      v[3] = 0.0;
      xreduce_step<2>(v, lane, 2);
      xreduce_step<1>(v, lane, 1);
      __syncwarp();  // everybody is done reading c/e from cb
      if (red_valid && red_idx < 19) cb[red_idx] = v[0];
      __syncwarp();
      Jc_total = fma(cb[18], dt, Jc_total);
      // ---- back-projection through the segment block: g[a][r] = sum_i U[a][i]*Linv[i][r] (T*Ldp, :378-379)
      if (lane < 18) {
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < 6; ++i) acc = fma(cb[la * 6 + i], Li[i * 6 + li], acc);
        ga[la * nrow + 3 * s + li] += acc * dt;
      }
      __syncwarp();
    }
  }
  if (oob_counter != nullptr) {
    oob = __reduce_add_sync(kFull, oob);
    if (lane == 0 && oob) atomicAdd(oob_counter, (unsigned long long)oob);
  }

  // ---- 3. smoothness cost: banded column sums of R in waypoint order (:326-327) -------------------
  double js = 0.0;
  for (int u = lane; u < nrow; u += 32) {
    const double* __restrict__ band = pr + PL.off_Rc + u * 9;
    const int base = 3 * (u / 3 - 1);
    const int j0 = base < 0 ? 3 : 0, j1 = (base + 9 > nrow) ? 6 : 9;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const double* Wa = Wd + a * nrow;
      double y = 0.0;
      for (int j = j0; j < j1; ++j) y = fma(Wa[base + j], band[j], y);
      js = fma(y, Wa[u], js);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) js += __shfl_xor_sync(kFull, js, o);

  // ---- 4. outputs (:417-431) ----------------------------------------------------------------------
  double* __restrict__ gout = grad_out + b * n;
  for (int j = lane; j < P; j += 32) {
    const double* __restrict__ band = pr + PL.off_Rg + j * 9;  // 2 Rfp' df + 2 Rpp dp row (:330-332)
    const int base = j - j % 3;                                // 3*(wp-1), wp = j/3 + 1
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const double* Wa = Wd + a * nrow + base;
      double gs = 0.0;
#pragma unroll
      for (int jj = 0; jj < 9; ++jj) gs = fma(Wa[jj], band[jj], gs);
      gout[a * P + j] = (E.ws * (2.0 * gs) + E.wc * ga[a * nrow + 3 + j]) + 1e-5;
    }
  }
  if (lane == 0) cost_out[b] = E.ws * js + E.wc * (E.alpha * Jc_total) + 1e-3;
}

cudaError_t launch_eval(const SdfView& S, const ProbLayout& PL, const double* prob, int num_problems,
                        const EvalScalars& E, const double* X, int64_t B, int64_t traj_offset,
                        int64_t starts_per_problem, double* cost, double* grad, unsigned long long* oob_counter,
                        int max_ctas, cudaStream_t stream) {
  (void)max_ctas;
  if (B <= 0) return cudaSuccess;
  const size_t smem = sizeof(double) * (size_t)warp_smem_doubles(PL.m) * kWarpsPerCta;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k_eval_traj, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  const int64_t grid = (B + kWarpsPerCta - 1) / kWarpsPerCta;
  k_eval_traj<<<(unsigned)grid, kWarpsPerCta * 32, smem, stream>>>(S, PL, prob, num_problems, E, X, B, traj_offset,
                                                                   starts_per_problem, cost, grad, oob_counter);
  return cudaGetLastError();
}

}  // namespace gtop
