// fp64_mix.cu — do FP64<->FP32 conversions (F2F), FRND/F2I.F64 and DMMA share execution time with DFMA on this GPU?
// Runs each instruction stream alone and then interleaved in one warp; overlap => mixed ~ max, shared => mixed ~ sum.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_mix fp64_mix.cu
#include <cuda_runtime.h>

#include <cstdio>

template <int A, int Bk>  // A: DFMA stream on/off; Bk: 0 none, 1 F2F round trips, 2 floor+cvt (FRND/F2I), 3 int IMAD stream, 4 LDS stream
__global__ void __launch_bounds__(256) k_mix(double* out, int iters, double seed) {
  __shared__ double sm[256];
  sm[threadIdx.x] = seed + threadIdx.x;
  __syncthreads();
  double x[8], y[4];
  int q[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = seed + i + threadIdx.x * 1e-3;
#pragma unroll
  for (int i = 0; i < 4; ++i) y[i] = seed * 3.3 + i + threadIdx.x * 1e-2, q[i] = threadIdx.x + i;
  const double a = 1.0000001, b = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if (A) {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);  // 8 DFMA
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {  // 4 "B" ops (8 instructions for the round trip)
        if (Bk == 1) y[i] = (double)__double2float_rn(y[i]);
        if (Bk == 2) y[i] = floor(y[i]) + (double)__double2int_rd(y[i]) * 1e-300;
        if (Bk == 3) q[i] = q[i] * 3 + it;
        if (Bk == 4) y[i] += sm[(q[i] + it) & 255];
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
#pragma unroll
  for (int i = 0; i < 4; ++i) s += y[i] + q[i];
  if (s == 12345.678) out[0] = s;
}

template <int A, int Bk>
static double run(double* out) {
  const int iters = 4000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  k_mix<A, Bk><<<148 * 2, 256>>>(out, 100, 1.0);  // 16 warps per SM = 4 per sub-partition
  cudaEventRecord(e0);
  k_mix<A, Bk><<<148 * 2, 256>>>(out, iters, 1.0);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms * 1e-3 * 1.965e9 / (iters * 4.0) / 4.0;  // cycles per (8 DFMA + 4 B-ops) group per warp, divided by warps/SMSP
}

int main() {
  double* out;
  cudaMalloc(&out, 64);
  printf("cycles per SMSP per group of [8 DFMA | 4 B-ops] (4 warps per SMSP)\n");
  printf("DFMA alone            : %.2f\n", run<1, 0>(out));
  printf("F2F rt alone          : %.2f   mixed with DFMA: %.2f\n", run<0, 1>(out), run<1, 1>(out));
  printf("floor+cvt alone       : %.2f   mixed with DFMA: %.2f\n", run<0, 2>(out), run<1, 2>(out));
  printf("IMAD alone            : %.2f   mixed with DFMA: %.2f\n", run<0, 3>(out), run<1, 3>(out));
  printf("LDS alone             : %.2f   mixed with DFMA: %.2f\n", run<0, 4>(out), run<1, 4>(out));
  return cudaGetLastError() != cudaSuccess;
}
