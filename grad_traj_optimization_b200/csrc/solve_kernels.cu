// solve_kernels.cu — batched lock-step MMA bookkeeping around the cost+gradient kernel (mma_core.cuh).
//
// One WARP per trajectory runs the per-trajectory state machine between two batched evaluations: the scalar
// decisions (accept / rho / outer iteration) are taken redundantly by every lane, the per-variable work — copies,
// asymptote update, the closed-form minimiser of the next approximation — is spread over the lanes (all buffers are
// [trajectory][variable]: coalesced), and the two sums of the approximation (g, w) are combined with a shuffle
// butterfly.  (The host loop of gtop_mma_minimize adds them in index order, which is NLopt's order; on the device
// the objective already differs from the CPU's at the 1e-15 level, so the order of these sums is immaterial.)
#include <cuda_runtime.h>

#include <cstdint>

#include "mma_core.cuh"

namespace gtop {

struct SolveView {
  int n;
  int64_t B, traj_offset, starts_per_problem;
  int num_problems;
  const double* lb;    // [num_problems][n]
  const double* ub;
  double* xcur;        // [B][n]  evaluation input
  const double* grad;  // [B][n]  evaluation output
  const double* cost;  // [B]
  double* x;           // [B][n]  best point
  double* dfdx;        // [B][n]
  double* sigma;       // [B][n]
  double* xprev;       // [B][n]
  double* xprevprev;   // [B][n]
  MmaScalars* sc;      // [B]
};

constexpr int kSolveWarps = 4;

// xcur := clip(start, lb, ub); sigma, rho, counters initialised
__global__ void __launch_bounds__(kSolveWarps * 32) k_mma_init(SolveView V) {
  const int lane = threadIdx.x & 31;
  const int64_t b = (int64_t)blockIdx.x * kSolveWarps + (threadIdx.x >> 5);
  if (b >= V.B) return;
  const int64_t pid = ((V.traj_offset + b) / V.starts_per_problem) % V.num_problems;
  const double *lb = V.lb + pid * V.n, *ub = V.ub + pid * V.n;
  for (int j = lane; j < V.n; j += 32) {
    const int64_t i = b * V.n + j;
    V.xcur[i] = fmin(fmax(V.xcur[i], lb[j]), ub[j]);
    V.sigma[i] = mma_sigma0(lb[j], ub[j]);
  }
  if (lane == 0) {
    MmaScalars s;
    s.rho = 1.0, s.k = 0, s.evals = 0, s.fval = s.fcur = s.gval = s.wval = 0.0;
    V.sc[b] = s;
  }
}

// after an evaluation of xcur: bookkeeping, then the next candidate into xcur
__global__ void __launch_bounds__(kSolveWarps * 32) k_mma_step(SolveView V, int want_candidate) {
  const int lane = threadIdx.x & 31;
  const int64_t b = (int64_t)blockIdx.x * kSolveWarps + (threadIdx.x >> 5);
  if (b >= V.B) return;
  const int64_t pid = ((V.traj_offset + b) / V.starts_per_problem) % V.num_problems;
  const double *lb = V.lb + pid * V.n, *ub = V.ub + pid * V.n;
  MmaScalars s = V.sc[b];
  const MmaDecision d = mma_after_eval_scalars(V.cost[b], s);
  double gsum = 0.0, wsum = 0.0;
  for (int j = lane; j < V.n; j += 32) {
    const int64_t i = b * V.n + j;
    const double xc = V.xcur[i], lo = lb[j], hi = ub[j];
    double xj, gj, sg = V.sigma[i];
    if (d.accept) {
      xj = xc, gj = V.grad[i];
      V.x[i] = xj, V.dfdx[i] = gj;
    } else {
      xj = V.x[i], gj = V.dfdx[i];
    }
    if (d.outer_done) {
      const double xp = V.xprev[i];
      if (d.sigma_update) {
        sg = mma_sigma_update(sg, xc, xp, V.xprevprev[i], lo, hi);
        V.sigma[i] = sg;
      }
      V.xprevprev[i] = xp, V.xprev[i] = xc;
    }
    if (want_candidate) V.xcur[i] = mma_candidate_var(xj, gj, sg, lo, hi, s.rho, gsum, wsum);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    gsum += __shfl_xor_sync(0xffffffffu, gsum, o);
    wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
  }
  if (want_candidate) s.gval = s.fval + gsum, s.wval = wsum;
  if (lane == 0) V.sc[b] = s;
}

// best point back into the caller's [B][n] buffer, best value into minf
__global__ void __launch_bounds__(256) k_mma_finish(SolveView V, double* __restrict__ X, double* __restrict__ minf) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < V.B * V.n) X[i] = V.x[i];
  if (minf && i < V.B) minf[i] = V.sc[i].fval;
}

cudaError_t launch_mma_init(const SolveView& V, cudaStream_t st) {
  k_mma_init<<<(unsigned)((V.B + kSolveWarps - 1) / kSolveWarps), kSolveWarps * 32, 0, st>>>(V);
  return cudaGetLastError();
}
cudaError_t launch_mma_step(const SolveView& V, int want_candidate, cudaStream_t st) {
  k_mma_step<<<(unsigned)((V.B + kSolveWarps - 1) / kSolveWarps), kSolveWarps * 32, 0, st>>>(V, want_candidate);
  return cudaGetLastError();
}
cudaError_t launch_mma_finish(const SolveView& V, double* X, double* minf, cudaStream_t st) {
  const int64_t tot = V.B * V.n;
  k_mma_finish<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(V, X, minf);
  return cudaGetLastError();
}

// ---- FP64 ceilings of the device (bench.py reports the evaluation kernel against them) ---------------
// mode 0: independent DFMA chains; mode 1: independent mma.sync.m8n8k4.f64 (DMMA) chains
template <int MODE>
__global__ void __launch_bounds__(256) k_fp64_peak(double* out, int iters, double seed) {
  double x[8], c[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = seed + i + threadIdx.x, c[i] = seed * i;
  const double a = seed * 1.0000001, b = seed * 0.9999999;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
    } else {
#pragma unroll
      for (int i = 0; i < 8; i += 2)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c[i]), "+d"(c[i + 1]) : "d"(a), "d"(b));
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i] + c[i];
  if (s == 12345.678) out[0] = s;
}

// returns TFLOP/s (DFMA = 2 flops per lane, DMMA m8n8k4 = 512 flops per warp instruction)
cudaError_t measure_fp64_peak(double* dfma_tflops, double* dmma_tflops, cudaStream_t st) {
  double* out = nullptr;
  cudaError_t e = cudaMalloc(&out, 64);
  if (e != cudaSuccess) return e;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = sms * 8, iters = 8000;
  const double warps = (double)grid * 8;
  float ms = 0;
  for (int mode = 0; mode < 2 && e == cudaSuccess; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {  // first repetition warms up
      cudaEventRecord(e0, st);
      if (mode == 0) k_fp64_peak<0><<<grid, 256, 0, st>>>(out, iters, 1.0);
      else k_fp64_peak<1><<<grid, 256, 0, st>>>(out, iters, 1.0);
      cudaEventRecord(e1, st);
      e = cudaStreamSynchronize(st);
      if (e != cudaSuccess) break;
      cudaEventElapsedTime(&ms, e0, e1);
    }
    const double winst = warps * iters * (mode == 0 ? 32.0 : 4.0);
    const double tf = winst * (mode == 0 ? 64.0 : 512.0) / (ms * 1e-3) / 1e12;
    if (mode == 0) *dfma_tflops = tf;
    else *dmma_tflops = tf;
  }
  cudaEventDestroy(e0), cudaEventDestroy(e1);
  cudaFree(out);
  return e;
}

}  // namespace gtop
