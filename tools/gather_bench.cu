// gather_bench.cu — how many L1TEX data-pipe cycles does the trilinear corner gather cost per sample, for
// different HBM layouts of the distance field?  (experiment behind DESIGN.md §2 "cell-row layout")
//   A: z-pair layout, 16 B/voxel, 4 x LDG.128 per sample (columns (x0,y0),(x0,y1),(x1,y0),(x1,y1))
//   B: x-pair+z-pair layout, 32 B/voxel, 2 x LDG.256 per sample (rows y0, y1)
//   C: xy-quad layout, 32 B/voxel, 2 x LDG.256 per sample (z0, z0+1: adjacent in memory)
//   D: full cell, 64 B/voxel, 2 x LDG.256 per sample (same 64-B block)
// Access pattern mimics k_eval_traj: 8 trajectories per warp, 4 lanes per trajectory on consecutive samples
// half a voxel apart along a random direction.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

struct __align__(32) d4 { double x, y, z, w; };
__device__ __forceinline__ d4 ldg256(const void* p) {
  d4 r;
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ double2 ldg128(const void* p) { return __ldg((const double2*)p); }

constexpr int NX = 200, NY = 200, NZ = 50;

__device__ __forceinline__ unsigned hash(unsigned x) {
  x ^= x >> 16, x *= 0x7feb352du, x ^= x >> 15, x *= 0x846ca68bu, x ^= x >> 16;
  return x;
}

template <int MODE>
__global__ void __launch_bounds__(128) k_gather(const char* __restrict__ tab, double* out, int iters) {
  const int lane = threadIdx.x & 31, t = lane & 3;
  const unsigned traj = (blockIdx.x * blockDim.x + threadIdx.x) >> 2;
  double acc = 0.0;
  for (int it = 0; it < iters; ++it) {
    const unsigned h = hash(traj * 977u + (it >> 3) * 131071u);  // a new segment every 8 iterations
    const float bx = 20.f + (h & 127), by = 20.f + ((h >> 7) & 127), bz = 12.f + ((h >> 14) & 15);
    const float ang = (float)((h >> 18) & 1023) * (6.2831853f / 1024.f);
    const float k = (float)((it & 7) * 4 + t) * 0.5f;
    const float px = bx + k * __cosf(ang), py = by + k * __sinf(ang), pz = bz + k * 0.1f;
    const int x0 = (int)px, y0 = (int)py, z0 = (int)pz;
    if (MODE == 0) {
      const char* p = tab + 16 * (size_t)((x0 * NY + y0) * NZ + z0);
      const double2 a = ldg128(p), b = ldg128(p + 16 * NZ), c = ldg128(p + 16 * NY * NZ), d = ldg128(p + 16 * (NY * NZ + NZ));
      acc += (a.x + a.y) + (b.x + b.y) + (c.x + c.y) + (d.x + d.y);
    } else if (MODE == 1) {
      const char* p = tab + 32 * (size_t)((x0 * NY + y0) * NZ + z0);
      const d4 a = ldg256(p), b = ldg256(p + 32 * NZ);
      acc += (a.x + a.y) + (a.z + a.w) + (b.x + b.y) + (b.z + b.w);
    } else if (MODE == 2) {
      const char* p = tab + 32 * (size_t)((x0 * NY + y0) * NZ + z0);
      const d4 a = ldg256(p), b = ldg256(p + 32);
      acc += (a.x + a.y) + (a.z + a.w) + (b.x + b.y) + (b.z + b.w);
    } else if (MODE == 4) {  // quads, 128-B line = 2x2 xy patch at one z: [x/2][y/2][z][x%2][y%2]
      const char* p = tab + 32 * (size_t)(((((x0 >> 1) * (NY / 2) + (y0 >> 1)) * NZ + z0) << 2) + ((x0 & 1) << 1) + (y0 & 1));
      const d4 a = ldg256(p), b = ldg256(p + 128);
      acc += (a.x + a.y) + (a.z + a.w) + (b.x + b.y) + (b.z + b.w);
    } else if (MODE == 5) {  // quads, y fastest: [x][z][y]
      const char* p = tab + 32 * (size_t)((x0 * NZ + z0) * NY + y0);
      const d4 a = ldg256(p), b = ldg256(p + 32 * NY);
      acc += (a.x + a.y) + (a.z + a.w) + (b.x + b.y) + (b.z + b.w);
    } else if (MODE == 6) {  // x-pair+z-pair elements, y fastest: rows y0,y1 adjacent: [x][z][y]
      const char* p = tab + 32 * (size_t)((x0 * NZ + z0) * NY + y0);
      const d4 a = ldg256(p), b = ldg256(p + 32);
      acc += (a.x + a.y) + (a.z + a.w) + (b.x + b.y) + (b.z + b.w);
    } else {
      const char* p = tab + 64 * (size_t)((x0 * NY + y0) * NZ + z0);
      const d4 a = ldg256(p), b = ldg256(p + 32);
      acc += (a.x + a.y) + (a.z + a.w) + (b.x + b.y) + (b.z + b.w);
    }
  }
  if (acc == 1234.5) out[0] = acc;
}

template <int MODE>
static double run(const char* tab, double* out, int iters, int grid) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  k_gather<MODE><<<grid, 128>>>(tab, out, iters);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k_gather<MODE><<<grid, 128>>>(tab, out, iters);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  const size_t nvox = (size_t)NX * NY * NZ;
  char* tab;
  cudaMalloc(&tab, nvox * 64 + 4096);
  cudaMemset(tab, 0, nvox * 64 + 4096);
  double* out;
  cudaMalloc(&out, 64);
  const int iters = 320, grid = 148 * 16 * 4;  // 4 waves of 16 CTAs x 4 warps per SM
  const double samples = (double)grid * 128 * iters;
  const double ms[7] = {run<0>(tab, out, iters, grid), run<1>(tab, out, iters, grid), run<2>(tab, out, iters, grid),
                        run<3>(tab, out, iters, grid), run<4>(tab, out, iters, grid), run<5>(tab, out, iters, grid),
                        run<6>(tab, out, iters, grid)};
  const char* names[7] = {"A_zpair_4xLDG128", "B_xzpair_2xLDG256", "C_xyquad_2xLDG256", "D_cell64_2xLDG256",
                          "E_quad_xypatch", "F_quad_yfastest", "G_xzpair_yfastest"};
  printf("{");
  for (int i = 0; i < 7; ++i)
    printf("\"%s\": {\"ms\": %.4f, \"Gsamples_per_s\": %.2f, \"sm_cycles_per_warp_iter_at_1965MHz\": %.1f}%s", names[i], ms[i],
           samples / ms[i] / 1e6, ms[i] * 1e-3 * 1.965e9 * 148 / (samples / 32), i < 6 ? ", " : "");
  printf("}\n");
  return cudaGetLastError() != cudaSuccess;
}
