// setup_kernels.cu — set-up of a batch of GTOP problems on the device: one waypoint set per
// trajectory (BASELINE configs 2 and 4) makes the set-up, not the evaluation, the cost of a planning cycle when it
// runs on the host (tools/setup_bench.py: 171 ms for 4,096 problems of 10 segments on 16 host threads, against
// 0.08 ms for evaluating them).
//
// Replaces GradTrajOptimizer::setPath (src/grad_traj_optimizer.cpp:67-110) + TrajectoryGenerator::PolyQPGeneration
// / StackOptiDep (src/qp_generator.cpp:156-236, :357-405) for a whole batch: the reference inverts three dense
// (6m)^2 matrices per problem; here every segment's 6x6 block is built from T_s alone (setup_core.cuh, the same code
// the host path runs) and the packed per-problem record the evaluation kernels read (gtop_types.cuh) is written in
// place — only the waypoints cross PCIe.
//
//   CTA = one problem.  Phase 1: thread s builds the blocks of segment s (A_s^-1, M_s) into shared memory and walks
//   the segment's sample times (the reference's sequential `t += dt` loop, :353 — H8).  Phase 2: all threads fill the
//   record's sections entry by entry from shared memory (bands of R, transposes, power / fragment tables).
//
// Compiled with -fmad=false: every product / sum rounds as on the host, so the record only differs from the host
// path's where libm's pow() is not correctly rounded (setup_pow; 1 ulp in ~0.08 % of the calls) and in what is
// computed from such a power — rounding level, <= 1e-12 relative (tests/test_gpu_setup.py).
#include <cuda_runtime.h>

#include <cstdint>

#include "gtop_types.cuh"
#include "setup_core.cuh"

namespace gtop {

struct SetupArgs {
  ProbLayout PL;
  double mean_v, init_time;
  const double* pos;   // [np][m+1][3] waypoints
  const double* time;  // [np][m] given segment times (setKinoPath), or nullptr: allocated from the path (setPath)
  double* rec;         // [np][PL.stride], zero-filled by the caller (padding slots stay zero)
  int* status;         // [np]: 0 ok, 1 non-positive segment time, 2 singular mapping matrix
};

// shared memory (doubles): Mseg[m][36] | Linv[m][36] | dt[m] | ns[m]
__host__ __device__ inline size_t setup_smem_doubles(int m) { return (size_t)m * 74; }

__global__ void __launch_bounds__(128) k_setup_problems(SetupArgs A) {
  extern __shared__ __align__(16) double sm[];
  const ProbLayout& L = A.PL;
  const int m = L.m, NS = L.ns_cap, nd = 3 * m + 3, P = L.P;
  double* Mseg = sm;
  double* Linv = Mseg + (size_t)m * 36;
  double* dts = Linv + (size_t)m * 36;
  double* nss = dts + m;
  const int p = blockIdx.x;
  const double* wp = A.pos + (size_t)p * (m + 1) * 3;
  double* rec = A.rec + (size_t)p * L.stride;
  __shared__ int bad;
  if (threadIdx.x == 0) bad = 0;
  __syncthreads();

  // ---- phase 1: per-segment blocks and sample times ----------------------------------------------------------
  for (int s = threadIdx.x; s < m; s += blockDim.x) {
    const double T = A.time ? A.time[(size_t)p * m + s] : setup_segment_time(wp, s, A.mean_v, A.init_time);
    if (!(T > 0.0)) {
      atomicMax(&bad, 1);
      dts[s] = 0.0, nss[s] = 0.0;
      continue;
    }
    if (!setup_segment_blocks(T, Linv + (size_t)s * 36, Mseg + (size_t)s * 36)) atomicMax(&bad, 2);
    const double dt = T / (double)NS;  // grad_traj_optimizer.cpp:351
    int k = 0;
    double* tt = rec + L.off_tt + (size_t)s * 5 * NS;
    // for (double t = 1e-3; t < T; t += dt)  (:353): sequential accumulation reproduced
    for (double t = 1e-3; t < T && k < NS; t += dt, ++k) tt[k] = t;
    dts[s] = dt, nss[s] = (double)k;
    rec[L.off_T + s] = T;
  }
  __syncthreads();
  if (bad) {
    if (threadIdx.x == 0) A.status[p] = bad;
    return;
  }
  if (threadIdx.x == 0) A.status[p] = 0;

  // ---- phase 2: the record, entry by entry --------------------------------------------------------------------
  for (int s = threadIdx.x; s < m; s += blockDim.x) rec[L.off_dt + s] = dts[s], rec[L.off_ns + s] = nss[s];
  for (int e = threadIdx.x; e < m * 36; e += blockDim.x) {
    const int s = e / 36, q = e - 36 * s, i = q / 6, r = q - 6 * i;
    const double v = Linv[e];  // Linv[s][i][r]
    rec[L.off_Linv + e] = v;
    rec[L.off_LinvT + s * 36 + r * 6 + i] = v;
    rec[L.off_LiTP + s * 36 + ((i >> 1) * 6 + r) * 2 + (i & 1)] = v;
  }
  for (int e = threadIdx.x; e < nd * 9; e += blockDim.x) rec[L.off_Rc + e] = setup_Rc_entry(Mseg, m, e / 9, e % 9);
  for (int e = threadIdx.x; e < P * 9; e += blockDim.x) {
    const double v = setup_Rg_entry(Mseg, m, e / 9, e % 9);
    rec[L.off_Rg + e] = v, rec[L.off_RgP + (e / 9) * 10 + e % 9] = v;
  }
  // Df (getInitialD, qp_generator.cpp:410-431): start / end position, zero start / end velocity and acceleration
  if (threadIdx.x < 3) {
    const int a = threadIdx.x;
    rec[L.off_Df + a * 6 + 0] = wp[a], rec[L.off_Df + a * 6 + 3] = wp[3 * m + a];
  }
  // power and fragment tables, one sample slot per thread
  for (int e = threadIdx.x; e < m * NS; e += blockDim.x) {
    const int s = e / NS, k = e - s * NS;
    if (k >= (int)nss[s]) continue;
    double* tt = rec + L.off_tt + (size_t)s * 5 * NS;
    double pw[6];
    pw[0] = 1.0, pw[1] = tt[k];
    for (int q = 2; q <= 5; ++q) pw[q] = setup_pow(pw[1], q), tt[(size_t)(q - 1) * NS + k] = pw[q];  // :457-485, :549
    int u, j, rec_pl;
    setup_pair_slot(s, k, (s & 1) ? (int)nss[s - 1] : (int)nss[s], L.nit2, &u, &j, &rec_pl);
    double* tq = rec + L.off_TQ + ((size_t)(s >> 1) * L.nit2 + u) * 24;
    for (int part = 0; part < 3; ++part) tq[(part * 4 + j) * 2] = pw[2 * part], tq[(part * 4 + j) * 2 + 1] = pw[2 * part + 1];
    tq[j * 2] = dts[s];  // the slot's quadrature weight in the place of t^0
    double* pl = rec + L.off_PL + ((size_t)(s >> 1) * (L.nit2 + 1) + rec_pl) * 64;
    for (int g = 0; g < 8; ++g) {
      const int r = deriv_of_col(g, s);
      if (r < 0) continue;
      setup_basis_pair(pw, Linv + (size_t)s * 36, r, dts[s], &pl[(4 * g + j) * 2], &pl[(4 * g + j) * 2 + 1]);
    }
  }
}

// pos_dev: [np][m+1][3]; time_dev: [np][m] or nullptr; rec_dev zero-filled here; status_dev: [np]
cudaError_t launch_setup_problems(const ProbLayout& PL, double mean_v, double init_time, const double* pos_dev,
                                  const double* time_dev, int np, double* rec_dev, int* status_dev, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(rec_dev, 0, sizeof(double) * (size_t)np * PL.stride, st);
  if (e != cudaSuccess) return e;
  const size_t smem = sizeof(double) * setup_smem_doubles(PL.m);
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  if (smem > 48 * 1024) {
    e = cudaFuncSetAttribute(k_setup_problems, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  SetupArgs A{PL, mean_v, init_time, pos_dev, time_dev, rec_dev, status_dev};
  k_setup_problems<<<(unsigned)np, 128, smem, st>>>(A);
  return cudaGetLastError();
}

}  // namespace gtop
