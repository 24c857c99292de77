// fp64_peak.cu — measures the FP64 issue ceilings of the GPU the eval kernel runs on (BASELINE.md §2:
// "% of measured FP64 peak" next to "% of measured HBM"):
//   dfma  : independent DFMA chains (vector FP64 pipe)
//   dmma  : independent mma.sync.m8n8k4.f64 chains (SASS DMMA)
//   mixed : both interleaved in one instruction stream (do they share a pipe?)
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_peak fp64_peak.cu ; prints one JSON line.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

template <int MODE>  // 0 dfma, 1 dmma, 2 mixed
__global__ void __launch_bounds__(256) k_peak(double* out, int iters, double seed) {
  double x[8], c[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = seed + i + threadIdx.x, c[i] = seed * i;
  const double a = seed * 1.0000001, b = seed * 0.9999999;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0 || MODE == 2) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
    }
    if (MODE == 1 || MODE == 2) {
#pragma unroll
      for (int i = 0; i < 8; i += 2) dmma(c[i], c[i + 1], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i] + c[i];
  if (s == 12345.678) out[0] = s;
}

template <int MODE>
static double run(int iters, double* out) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  const int grid = 148 * 8;
  k_peak<MODE><<<grid, 256>>>(out, iters / 10, 1.0);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k_peak<MODE><<<grid, 256>>>(out, iters, 1.0);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  double* out;
  cudaMalloc(&out, 64);
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int iters = 20000;
  const double warps = 148.0 * 8 * 8;
  const double ms0 = run<0>(iters, out), ms1 = run<1>(iters, out), ms2 = run<2>(iters, out);
  const double dfma_winst = warps * iters * 32.0, dmma_winst = warps * iters * 4.0;
  // flops: DFMA warp instruction = 64, DMMA m8n8k4 = 512
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"dfma_ms\": %.4f, \"dfma_tflops\": %.3f, \"dfma_winst_per_ns\": %.3f, "
         "\"dmma_ms\": %.4f, \"dmma_tflops\": %.3f, \"dmma_winst_per_ns\": %.3f, "
         "\"mixed_ms\": %.4f, \"mixed_tflops\": %.3f, \"mixed_vs_sum\": %.3f}\n",
         p.name, p.multiProcessorCount, ms0, dfma_winst * 64 / ms0 / 1e9, dfma_winst / ms0 / 1e6, ms1,
         dmma_winst * 512 / ms1 / 1e9, dmma_winst / ms1 / 1e6, ms2, (dfma_winst * 64 + dmma_winst * 512) / ms2 / 1e9,
         ms2 / (ms0 + ms1));
  if (cudaGetLastError() != cudaSuccess) return 1;
  return 0;
}
