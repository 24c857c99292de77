// traj_metrics_kernels.cu — post-hoc evaluation of a batch of optimised trajectories on the device.
//
// Replaces, for a whole batch, what opti_node does with the result of one solve (src/opti_node.cpp:104-144):
//   GradTrajOptimizer::getCoefficient              src/grad_traj_optimizer.cpp:245-251 (coefficients = L * d, :253-279)
//   PolynomialTraj::getTraj / getLength            include/grad_traj_optimization/polynomial_traj.hpp:70-91
//   PolynomialTraj::getJerk / getAccCost           :95-139
//   PolynomialTraj::getMeanAndMaxVel / ...Acc      :141-203
// so that the multi-starts of gtop_solve_batch_* can be ranked without copying 65,536 coefficient sets to the host.
//
//   warp = one trajectory.  Coefficients of all segments are formed in shared memory in the reference's summation
//   order (bit-identical to gtop_coeff_from_x); the 0.01 s sampling clock of getTraj is the reference's sequential
//   accumulation `eval_t += 0.01` (every lane runs the additions, lane l evaluates samples l, l+32, ...), so the number
//   of trajectory points is exactly the reference's; polynomial values, norms and sums are smooth and agree to
//   rounding (pow() is the correctly rounded setup_pow, sums are warp reductions).
//
// Reference quirks reproduced: getMeanAndMaxVel / Acc evaluate at the segment DURATION for every step of
// their clock (pow(ts, i), :158, :190) — the end velocity / acceleration of each segment weighted by its number of
// 0.01 s steps; evaluate() has no bound on the segment index for t == time_sum (:47-50) — clamped to the last segment.
//
// Compiled with -fmad=false (coefficients and clocks round as on the host).
#include <cuda_runtime.h>

#include <cstdint>

#include "gtop_types.cuh"
#include "setup_core.cuh"

namespace gtop {

constexpr int kMetricsWarps = 4;
constexpr unsigned kFullMask = 0xffffffffu;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(kFullMask, v, o));
  return v;
}

// out[b][8] = time_sum, length, jerk, mean_v, max_v, mean_a, max_a, acc_cost
__global__ void __launch_bounds__(kMetricsWarps * 32)
k_traj_metrics(ProbLayout PL, const double* __restrict__ prob, int num_problems, const double* __restrict__ X, int64_t B,
               int64_t traj_offset, int64_t starts_per_problem, double* __restrict__ out) {
  extern __shared__ __align__(16) double sm[];
  const int m = PL.m, P = PL.P, n = PL.n;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t b = (int64_t)blockIdx.x * kMetricsWarps + warp;
  if (b >= B) return;
  double* coef = sm + (size_t)warp * (19 * m);  // [m][18] ascending powers per axis (getCoefficient), then T[m]
  double* Ts = coef + 18 * m;
  const int64_t pid = ((traj_offset + b) / starts_per_problem) % num_problems;
  const double* __restrict__ pr = prob + pid * (int64_t)PL.stride;
  const double* __restrict__ x = X + b * n;

  // ---- coefficients: c = L d per segment, ascending dense-column order, unfused (:273; gtop_coeff_from_x) ----
  for (int e = lane; e < 18 * m; e += 32) {
    const int s = e / 18, q = e - 18 * s, a = q / 6, i = q - 6 * a;
    const double* Li = pr + PL.off_Linv + s * 36 + i * 6;
    const bool last = (s == m - 1);
    double acc = 0.0;
    for (int k = 0; k < 6; ++k) {
      const int r = last ? (k + 3) % 6 : k;
      const int w = r < 3 ? s : s + 1, dq = r % 3;
      const double d = w == 0 ? pr[PL.off_Df + a * 6 + dq] : (w == m ? pr[PL.off_Df + a * 6 + 3 + dq] : x[a * P + 3 * (w - 1) + dq]);
      const double prod = Li[r] * d;
      acc = (k == 0) ? prod : acc + prod;
    }
    coef[e] = acc;
  }
  for (int s = lane; s < m; s += 32) Ts[s] = pr[PL.off_T + s];
  __syncwarp();

  double time_sum = 0.0;  // PolynomialTraj::init (:37-43)
  for (int s = 0; s < m; ++s) time_sum += Ts[s];

  // ---- getTraj + getLength (:70-91): points every 0.01 s on the reference's accumulated clock --------------------
  double eval_t = 0.0;
  for (int k = 0; k < lane; ++k) eval_t += 0.01;
  double length = 0.0;
  double carry[3] = {0.0, 0.0, 0.0};  // last point of the previous round (lane 31's)
  bool have_carry = false;
  for (;;) {
    const bool valid = eval_t <= time_sum;
    if (!__any_sync(kFullMask, valid)) break;
    double pt[3] = {0.0, 0.0, 0.0};
    if (valid) {
      double t = eval_t;  // evaluate (:45-65)
      int idx = 0;
      while (idx < m - 1 && Ts[idx] <= t) t -= Ts[idx], ++idx;
      const double* c = coef + idx * 18;
      double pw[6];
#pragma unroll
      for (int i = 0; i < 6; ++i) pw[i] = setup_pow(t, i);
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        double v = 0.0;
#pragma unroll
        for (int i = 5; i >= 0; --i) v += pw[i] * c[a * 6 + i];  // tv.dot(cx): descending powers first
        pt[a] = v;
      }
    }
    // segment (k-1, k): previous point from lane-1, or from the previous round's lane 31
    double prev[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const double up = __shfl_up_sync(kFullMask, pt[a], 1);
      prev[a] = lane == 0 ? carry[a] : up;
    }
    const bool has_prev = lane == 0 ? have_carry : true;
    if (valid && has_prev) {
      const double dx = pt[0] - prev[0], dy = pt[1] - prev[1], dz = pt[2] - prev[2];
      length += sqrt(dx * dx + dy * dy + dz * dz);
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) carry[a] = __shfl_sync(kFullMask, pt[a], 31);
    have_carry = true;
    for (int k = 0; k < 32; ++k) eval_t += 0.01;
  }
  length = warp_sum(length);

  // ---- per-segment quantities: lanes over segments --------------------------------------------------------------
  double jerk = 0.0, acc_cost = 0.0, sum_v = 0.0, sum_a = 0.0, max_v = -1.0, max_a = -1.0, num = 0.0;
  for (int s = lane; s < m; s += 32) {
    const double* c = coef + s * 18;
    const double ts = Ts[s];
    double pw[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) pw[i] = setup_pow(ts, i);
    // getJerk (:109-139): c' J c with J[i][j] = i(i-1)(i-2) j(j-1)(j-2) ts^(i+j-5) / (i+j-5), i,j = 3..5
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      double q = 0.0;
#pragma unroll
      for (int j = 3; j < 6; ++j) {
        double col = 0.0;
#pragma unroll
        for (int i = 3; i < 6; ++i) {
          const double J = (double)(i * (i - 1) * (i - 2)) * (double)(j * (j - 1) * (j - 2)) * setup_pow(ts, i + j - 5) / (double)(i + j - 5);
          col += c[a * 6 + i] * J;
        }
        q += col * c[a * 6 + j];
      }
      jerk += q;
    }
    // getAccCost (:95-107): |2 c_2|^2 ts
    {
      const double ux = 2 * c[2], uy = 2 * c[6 + 2], uz = 2 * c[12 + 2];
      acc_cost += (ux * ux + uy * uy + uz * uz) * ts;
    }
    // getMeanAndMaxVel / Acc (:141-203): polynomial at ts (H16), once per 0.01 s step of the segment's clock
    double v[3], ac[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      double vv = 0.0, aa = 0.0;
#pragma unroll
      for (int i = 0; i < 5; ++i) vv += pw[i] * ((double)(i + 1) * c[a * 6 + i + 1]);
#pragma unroll
      for (int i = 0; i < 4; ++i) aa += pw[i] * ((double)((i + 2) * (i + 1)) * c[a * 6 + i + 2]);
      v[a] = vv, ac[a] = aa;
    }
    const double vn = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    const double an = sqrt(ac[0] * ac[0] + ac[1] * ac[1] + ac[2] * ac[2]);
    double steps = 0.0;
    for (double et = 0.0; et < ts; et += 0.01) steps += 1.0;
    sum_v += steps * vn, sum_a += steps * an, num += steps;
    if (steps > 0.0) max_v = fmax(max_v, vn), max_a = fmax(max_a, an);
  }
  jerk = warp_sum(jerk), acc_cost = warp_sum(acc_cost), sum_v = warp_sum(sum_v), sum_a = warp_sum(sum_a), num = warp_sum(num);
  max_v = warp_max(max_v), max_a = warp_max(max_a);
  if (lane == 0) {
    double* o = out + b * 8;
    o[0] = time_sum, o[1] = length, o[2] = jerk, o[3] = sum_v / num, o[4] = max_v, o[5] = sum_a / num, o[6] = max_a,
    o[7] = acc_cost;
  }
}

cudaError_t launch_traj_metrics(const ProbLayout& PL, const double* prob, int num_problems, const double* X, int64_t B,
                                int64_t traj_offset, int64_t starts_per_problem, double* out, cudaStream_t st) {
  if (B <= 0) return cudaSuccess;
  const size_t smem = sizeof(double) * (size_t)kMetricsWarps * 19 * PL.m;
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k_traj_metrics, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  const int64_t grid = (B + kMetricsWarps - 1) / kMetricsWarps;
  k_traj_metrics<<<(unsigned)grid, kMetricsWarps * 32, smem, st>>>(PL, prob, num_problems, X, B, traj_offset, starts_per_problem,
                                                                  out);
  return cudaGetLastError();
}

}  // namespace gtop
