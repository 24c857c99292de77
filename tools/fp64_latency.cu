// fp64_latency.cu — dependent-issue latency and throughput of FP64 ops on the eval kernel's GPU, as a function
// of independent chains per warp (ILP) and warps per SM sub-partition.  Prints a small table.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_latency fp64_latency.cu
#include <cuda_runtime.h>

#include <cstdio>

template <int ILP, int OP>  // OP 0: DFMA, 1: DADD, 2: DMUL(unfused, __dmul_rn), 3: F2F round trip, 4: DMMA chain
__global__ void k_lat(double* out, int iters, double seed) {
  double x[ILP];
  double c1 = seed;
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = seed + i + threadIdx.x * 1e-3;
  const double a = 1.0000001, b = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 16; ++r)
#pragma unroll
      for (int i = 0; i < ILP; ++i) {
        if (OP == 0) x[i] = fma(x[i], a, b);
        if (OP == 1) x[i] = __dadd_rn(x[i], b);
        if (OP == 2) x[i] = __dmul_rn(x[i], a);
        if (OP == 3) x[i] = (double)__double2float_rn(x[i]) ;
        if (OP == 4) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(x[i]), "+d"(c1) : "d"(a), "d"(b));
      }
  }
  double s = c1;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  if (s == 12345.678) out[0] = s;
}

template <int ILP, int OP>
static void run(const char* name, double* out) {
  const int iters = 4000;
  for (int wps = 1; wps <= 8; wps *= 2) {  // warps per SM sub-partition
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    const int threads = 32 * 4 * wps;  // one CTA per SM
    k_lat<ILP, OP><<<148, threads>>>(out, 100, 1.0);
    cudaEventRecord(e0);
    k_lat<ILP, OP><<<148, threads>>>(out, iters, 1.0);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double cyc = ms * 1e-3 * 1.965e9;
    const double ops_per_warp = (double)iters * 16 * ILP;
    printf("%-6s ILP=%d warps/SMSP=%d: %.2f cycles per op per warp (chain step %.2f), SMSP throughput %.3f warp-ops/cycle\n", name, ILP,
           wps, cyc / ops_per_warp, cyc / (iters * 16.0), ops_per_warp * wps / cyc);
  }
}

int main() {
  double* out;
  cudaMalloc(&out, 64);
  run<1, 0>("DFMA", out), run<2, 0>("DFMA", out), run<4, 0>("DFMA", out);
  run<1, 1>("DADD", out), run<1, 2>("DMUL", out);
  run<1, 3>("F2Frt", out), run<4, 3>("F2Frt", out);
  run<1, 4>("DMMA", out), run<3, 4>("DMMA", out);
  return cudaGetLastError() != cudaSuccess;
}
