// solve_kernels.cu — batched lock-step MMA bookkeeping around the cost+gradient kernel (mma_core.cuh).
//
// One thread per trajectory runs the per-trajectory state machine between two batched evaluations:
//   mma_after_eval (accept / rho / sigma / outer-iteration bookkeeping) then mma_candidate (next point).
// State vectors (x, dfdx, sigma, xprev, xprevprev) are stored [variable][trajectory] so that the threads of a
// warp touch consecutive addresses; the evaluation kernel's own buffers (xcur, grad) stay [trajectory][variable].
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
  double* x;           // [n][B]  best point
  double* dfdx;        // [n][B]
  double* sigma;       // [n][B]
  double* xprev;       // [n][B]
  double* xprevprev;   // [n][B]
  MmaScalars* sc;      // [B]
};

// xcur := clip(start, lb, ub); sigma, rho, counters initialised
__global__ void __launch_bounds__(128) k_mma_init(SolveView V) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= V.B) return;
  const int64_t pid = ((V.traj_offset + b) / V.starts_per_problem) % V.num_problems;
  const MmaCVec lb{V.lb + pid * V.n, 1}, ub{V.ub + pid * V.n, 1};
  double* xc = V.xcur + b * V.n;
  for (int j = 0; j < V.n; ++j) xc[j] = fmin(fmax(xc[j], lb[j]), ub[j]);
  MmaScalars s;
  mma_init(V.n, lb, ub, MmaVec{V.sigma + b, V.B}, s);
  V.sc[b] = s;
}

// after an evaluation of xcur: bookkeeping, then the next candidate into xcur
__global__ void __launch_bounds__(128) k_mma_step(SolveView V, int want_candidate) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= V.B) return;
  const int64_t pid = ((V.traj_offset + b) / V.starts_per_problem) % V.num_problems;
  const MmaCVec lb{V.lb + pid * V.n, 1}, ub{V.ub + pid * V.n, 1};
  MmaScalars s = V.sc[b];
  const MmaVec x{V.x + b, V.B}, dfdx{V.dfdx + b, V.B}, sigma{V.sigma + b, V.B}, xprev{V.xprev + b, V.B},
      xprevprev{V.xprevprev + b, V.B}, xcur{V.xcur + b * V.n, 1};
  mma_after_eval(V.n, V.cost[b], MmaCVec{V.grad + b * V.n, 1}, x, dfdx, sigma, lb, ub, MmaCVec{xcur.p, 1}, xprev, xprevprev, s);
  if (want_candidate)
    mma_candidate(V.n, MmaCVec{x.p, V.B}, MmaCVec{dfdx.p, V.B}, MmaCVec{sigma.p, V.B}, lb, ub, xcur, s);
  V.sc[b] = s;
}

// best point back into the caller's [B][n] buffer, best value into minf
__global__ void __launch_bounds__(128) k_mma_finish(SolveView V, double* __restrict__ X, double* __restrict__ minf) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= V.B) return;
  for (int j = 0; j < V.n; ++j) X[b * V.n + j] = V.x[(int64_t)j * V.B + b];
  if (minf) minf[b] = V.sc[b].fval;
}

cudaError_t launch_mma_init(const SolveView& V, cudaStream_t st) {
  k_mma_init<<<(unsigned)((V.B + 127) / 128), 128, 0, st>>>(V);
  return cudaGetLastError();
}
cudaError_t launch_mma_step(const SolveView& V, int want_candidate, cudaStream_t st) {
  k_mma_step<<<(unsigned)((V.B + 127) / 128), 128, 0, st>>>(V, want_candidate);
  return cudaGetLastError();
}
cudaError_t launch_mma_finish(const SolveView& V, double* X, double* minf, cudaStream_t st) {
  k_mma_finish<<<(unsigned)((V.B + 127) / 128), 128, 0, st>>>(V, X, minf);
  return cudaGetLastError();
}

}  // namespace gtop
