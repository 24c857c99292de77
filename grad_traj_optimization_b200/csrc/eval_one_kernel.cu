// eval_one_kernel.cu — cost + gradient of a FEW trajectories at minimum latency: the kernel behind gtop_objective,
// the NLopt callback that replaces GradTrajOptimizer::costFunc (src/grad_traj_optimizer.cpp:554-562 -> :281-448).
//
// NLopt calls the objective synchronously, one trajectory at a time, so what matters on this path is
// the latency of ONE evaluation, not throughput.  The batch kernels walk the m segments of a trajectory one after the
// other inside a (quarter-)warp: 10 segments x 4 sample iterations of dependent latency.  Here the segments are
// independent given x, so a CTA takes one trajectory and warp s takes segment s with one sample per lane
// (N_s = 30 -> a single sample iteration); the 19 sums of a segment are reduced with a shuffle butterfly, projected
// through the segment's 6x6 block, and the block combines the per-segment rows after one barrier.
// Same arithmetic as k_eval_traj (sample_accumulate in eval_common.cuh, same rounding discipline).
#include <cuda_runtime.h>

#include <cstdint>

#include "eval_common.cuh"
#include "gtop_types.cuh"

namespace gtop {

constexpr int kOneMaxSegments = 16;  // one warp per segment; 512 threads keep 128 registers per thread

// shared memory (doubles): Wd[3][nrow] | gseg[m][18] | jc[m] | jva[m] | cb[m][32] | xtab[16] ; then oob counters
__host__ __device__ inline int one_smem_doubles(int m) { return 9 * (m + 1) + 18 * m + 2 * m + 32 * m + kExpTab; }

template <bool kVelAcc, bool kPlain>
__global__ void __launch_bounds__(kOneMaxSegments * 32)
k_eval_one(SdfView S, ProbLayout PL, const double* __restrict__ prob, int num_problems, EvalScalars E,
           const double* __restrict__ X, int64_t B, int64_t traj_offset, int64_t starts_per_problem,
           double* __restrict__ cost_out, double* __restrict__ grad_out, unsigned long long* __restrict__ oob_counter) {
  extern __shared__ __align__(16) double smem[];
  const int m = PL.m, P = PL.P, n = PL.n, nrow = 3 * (m + 1);
  const int lane = threadIdx.x & 31, s = threadIdx.x >> 5;  // warp s <-> segment s
  const int64_t b = blockIdx.x;
  double* Wd = smem;
  double* gseg = Wd + 3 * nrow;
  double* jc = gseg + 18 * m;
  double* jva = jc + m;
  double* cb = jva + m + 32 * s;  // this segment's c[3][6], e[3][4]
  double* xtab = jva + m + 32 * m;
  __shared__ unsigned oob_block;

  const int64_t pid = ((traj_offset + b) / starts_per_problem) % num_problems;
  const double* __restrict__ pr = prob + pid * (int64_t)PL.stride;
  const double* __restrict__ x = X + b * n;

  // ---- boundary derivatives in waypoint order ------------------------------------------------------
  exp2_tab_init(xtab, threadIdx.x, blockDim.x);
  if (threadIdx.x == 0) oob_block = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int a = i / P, jj = i - a * P;
    Wd[a * nrow + 3 + jj] = x[i];
  }
  for (int o = threadIdx.x; o < 18; o += blockDim.x) {
    const int a = o / 6, r = o % 6;
    Wd[a * nrow + (r < 3 ? r : 3 * m + (r - 3))] = pr[PL.off_Df + o];
  }
  __syncthreads();

  unsigned oob = 0;
  if (!E.skip_collision) {
    // ---- coefficients of segment s: lanes 0..17, reference summation order (:273), unfused ----------
    const double* __restrict__ LiT = pr + PL.off_LinvT + s * 36;
    if (lane < 18) {
      const int a = lane / 6, i = lane % 6;
      const double* w0 = Wd + a * nrow + 3 * s;
      const bool last = (s == m - 1);
      double p[6];
#pragma unroll
      for (int r = 0; r < 6; ++r) p[r] = __dmul_rn(LiT[r * 6 + i], w0[r]);
      const double q0 = last ? p[3] : p[0], q1 = last ? p[4] : p[1], q2 = last ? p[5] : p[2];
      const double q3 = last ? p[0] : p[3], q4 = last ? p[1] : p[4], q5 = last ? p[2] : p[5];
      const double c = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(q0, q1), q2), q3), q4), q5);
      cb[lane] = c;
      if (i >= 2) cb[18 + a * 4 + (i - 2)] = __dmul_rn((double)i, c);  // e_k = k*c_k (:477-485)
    }
    __syncwarp();

    const int ns = (int)pr[PL.off_ns + s];
    const double dt = pr[PL.off_dt + s];
    const double* __restrict__ tt = pr + PL.off_tt + s * 5 * PL.ns_cap;
    const double kA = -E.alpha * E.alpha_over_r * S.res_inv;
    double v[20];
#pragma unroll
    for (int i = 0; i < 20; ++i) v[i] = 0.0;
    for (int k = lane; k < ns; k += 32) {
      const int nc = PL.ns_cap;
      sample_accumulate<kVelAcc, false, kPlain>(S, E, kA, cb, tt[k], tt[nc + k], tt[2 * nc + k], tt[3 * nc + k], tt[4 * nc + k], xtab,
                                        v, oob);
    }
    // ---- the 20 sums of the segment over its 32 lanes -------------------------------------------------
#pragma unroll
    for (int i = 0; i < 20; ++i) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(kFull, v[i], o);
    }
    // ---- back-projection through the segment block (T*Ldp, :378-379): lanes 0..17 ---------------------
    const double* __restrict__ Li = pr + PL.off_Linv + s * 36;
    if (lane < 18) {
      const int a = lane / 6, r = lane % 6;
      double acc = 0.0;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        // v[a*6+i] with a runtime a: select among the three axes
        const double u = a == 0 ? v[i] : (a == 1 ? v[6 + i] : v[12 + i]);
        acc = fma(u, Li[i * 6 + r], acc);
      }
      gseg[s * 18 + lane] = acc * dt;
    }
    if (lane == 0) jc[s] = v[18] * dt, jva[s] = v[19] * dt;
    oob = __reduce_add_sync(kFull, oob);
    if (lane == 0 && oob) atomicAdd(&oob_block, oob);
  }
  __syncthreads();

  // ---- outputs (:417-431): thread jj < P takes row jj of the three axes --------------------------------
  double* __restrict__ gout = grad_out + b * n;
  for (int jj = threadIdx.x; jj < P; jj += blockDim.x) {
    const double* __restrict__ band = pr + PL.off_Rg + jj * 9;  // 2 Rfp' df + 2 Rpp dp row (:330-332)
    const int w = jj / 3 + 1, q = jj % 3;                       // interior waypoint w, derivative q
    const double* Wb = Wd + 3 * (w - 1);
    double g0 = 0.0, g1 = 0.0, g2 = 0.0;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const double bv = band[t];
      g0 = fma(Wb[t], bv, g0);
      g1 = fma(Wb[nrow + t], bv, g1);
      g2 = fma(Wb[2 * nrow + t], bv, g2);
    }
    double c0 = 0.0, c1 = 0.0, c2 = 0.0;  // end rows of segment w-1 + start rows of segment w
    if (!E.skip_collision) {
      const double* ge = gseg + (w - 1) * 18 + 3 + q;
      const double* gs = gseg + w * 18 + q;
      c0 = ge[0] + gs[0], c1 = ge[6] + gs[6], c2 = ge[12] + gs[12];
    }
    gout[jj] = (E.ws * (2.0 * g0) + E.wc * c0) + 1e-5;
    gout[P + jj] = (E.ws * (2.0 * g1) + E.wc * c1) + 1e-5;
    gout[2 * P + jj] = (E.ws * (2.0 * g2) + E.wc * c2) + 1e-5;
  }
  // ---- cost: smoothness d' R d from the column band (:326-327) + collision sums, warp 0 ---------------
  if (s == 0) {
    double js = 0.0;
    for (int u = lane; u < nrow; u += 32) {
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
    double Jc = 0.0, Jva = 0.0;
    if (!E.skip_collision)
      for (int k = lane; k < m; k += 32) Jc += jc[k], Jva += jva[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      js += __shfl_xor_sync(kFull, js, o);
      Jc += __shfl_xor_sync(kFull, Jc, o);
      Jva += __shfl_xor_sync(kFull, Jva, o);
    }
    if (lane == 0) {
      cost_out[b] = E.ws * js + E.wc * (E.alpha * Jc) + (kVelAcc ? Jva : 0.0) + 1e-3;  // wv = wa = 1 (:412)
      if (oob_counter != nullptr && oob_block) atomicAdd(oob_counter, (unsigned long long)oob_block);
    }
  }
}

bool eval_one_applicable(const ProbLayout& PL) { return PL.m >= 2 && PL.m <= kOneMaxSegments; }

cudaError_t launch_eval_one(const SdfView& S, const ProbLayout& PL, const double* prob, int num_problems,
                            const EvalScalars& E, const double* X, int64_t B, int64_t traj_offset,
                            int64_t starts_per_problem, double* cost, double* grad, unsigned long long* oob_counter,
                            cudaStream_t stream) {
  if (B <= 0) return cudaSuccess;
  const size_t smem = sizeof(double) * (size_t)one_smem_doubles(PL.m);
  const bool plain = S.quad == nullptr;
  auto kern = E.vel_acc ? (plain ? k_eval_one<true, true> : k_eval_one<true, false>)
                        : (plain ? k_eval_one<false, true> : k_eval_one<false, false>);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  kern<<<(unsigned)B, PL.m * 32, smem, stream>>>(S, PL, prob, num_problems, E, X, B, traj_offset, starts_per_problem, cost,
                                                 grad, oob_counter);
  return cudaGetLastError();
}

}  // namespace gtop
