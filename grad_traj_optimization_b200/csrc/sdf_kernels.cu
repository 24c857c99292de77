// sdf_kernels.cu — voxel-map kernels: occupancy scatter, exact Euclidean distance transform,
// batched trilinear query.
//
// Replaces (reference paths relative to EpicOne1/grad_traj_optimization):
//   SDFMap::resetBuffer       src/sdf_map.cpp:26-53   (whole grid: k_fill; a box: k_box_reset)
//   SDFMap::setUpdateRange    src/sdf_map.cpp:244-264 (updateESDF3d restricted to a box: launch_edt_window)
//   SDFMap::setOccupancy      src/sdf_map.cpp:80-99   (isInMap :55-69, posToIndex :71-74)
//   SDFMap::updateESDF3d      src/sdf_map.cpp:310-368 (fillESDF :266-308)
//   SDFMap::getDistWithGradTrilinear   src/sdf_map.cpp:185-242  (k_query, test / utility path)
//
// EDT design.  The reference runs Felzenszwalb–Huttenlocher lower envelopes in fp64 along z, y, x.
// The result is the exact unsigned EDT: d(v) = res * sqrt(min over occupied u of |v-u|^2) capped at
// 10000, and 10000 everywhere for an empty map.  Here all three passes work on exact
// int32 squared voxel distances, so the only floating-point step is the final res*sqrt(double(d2))
// (IEEE-exact on both sides) and the output is bit-identical to the CPU path:
//   pass z : warp per (x,y) column (z is the contiguous axis): ballot + clz/ffs give the nearest
//            occupied voxel below and above each z in two coalesced sweeps.
//   pass y : thread per (x,z) row, threads adjacent in z so every step's loads/stores coalesce;
//            integer lower envelope (Meijster's formulation of the same parabola envelope) with the
//            stack in per-thread local memory.
//   pass x : same with thread per (y,z) row; fused epilogue min(res*sqrt(d2), old) -> fp64 buffer.
#include <cuda_runtime.h>

#include <cstdint>
#include <type_traits>

#include "gtop_types.cuh"

namespace gtop {

static constexpr int kInf = 0x3fffffff;  // "no occupied voxel" marker for int32 squared distances

// ---------------------------------------------------------------------------------------------
__global__ void k_fill(double* __restrict__ dist, int64_t ndist, uint8_t* __restrict__ occ, int64_t n, double dval) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < ndist; i += stride) {
    dist[i] = dval;  // (brick padding included: never read)
    if (i < n) occ[i] = 0;
  }
}
// distance field between the linear order (host side, reference order) and the map's own layout
__global__ void k_relayout(const double* __restrict__ src, double* __restrict__ dst, DistLayout lay, int nx, bool to_linear) {
  const int64_t n = (int64_t)nx * lay.ny * lay.nz;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int z = (int)(i % lay.nz);
  const int64_t c = i / lay.nz;
  const int64_t a = lay.at((int)(c / lay.ny), (int)(c % lay.ny), z);
  if (to_linear) dst[i] = src[a];
  else dst[a] = src[i];
}

// xy-quad copy of the distance field for the evaluation kernels' 256-bit gathers (see SdfView::quad).
// Thread per (2x2 patch, z), adjacent threads on adjacent z: the 3x3 neighbourhood columns are read
// coalesced along z, each thread writes the patch's four quads = one 128-byte line.
__global__ void __launch_bounds__(256)
k_build_quads(const double* __restrict__ dist, double2* __restrict__ quad, int nx, int ny, int nz, int nyh, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int z = (int)(i % nz);
  const int64_t patch = i / nz;
  const int x = 2 * (int)(patch / nyh), y = 2 * (int)(patch % nyh);
  double v[3][3];
#pragma unroll
  for (int dx = 0; dx < 3; ++dx)
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int xx = min(x + dx, nx - 1), yy = min(y + dy, ny - 1);
      v[dx][dy] = dist[((int64_t)xx * ny + yy) * nz + z];
    }
  double2* q = quad + i * 8;  // 4 quads x 2 double2
#pragma unroll
  for (int sx = 0; sx < 2; ++sx)
#pragma unroll
    for (int sy = 0; sy < 2; ++sy) {
      q[(sx * 2 + sy) * 2 + 0] = make_double2(v[sx][sy], v[sx][sy + 1]);
      q[(sx * 2 + sy) * 2 + 1] = make_double2(v[sx + 1][sy], v[sx + 1][sy + 1]);
    }
}

// the same for the quads that read a voxel of a box: patches [pxa, pxa+npx) x [pya, pya+npy), heights [z0, z0+wz)
__global__ void __launch_bounds__(256)
k_build_quads_box(const double* __restrict__ dist, double2* __restrict__ quad, int nx, int ny, int nz, int nyh, int pxa,
                  int pya, int z0, int npy, int wz, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int z = z0 + (int)(i % wz);
  const int64_t pp = i / wz;
  const int px = pxa + (int)(pp / npy), py = pya + (int)(pp % npy);
  const int x = 2 * px, y = 2 * py;
  double v[3][3];
#pragma unroll
  for (int dx = 0; dx < 3; ++dx)
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int xx = min(x + dx, nx - 1), yy = min(y + dy, ny - 1);
      v[dx][dy] = dist[((int64_t)xx * ny + yy) * nz + z];
    }
  double2* q = quad + (((int64_t)px * nyh + py) * nz + z) * 8;
#pragma unroll
  for (int sx = 0; sx < 2; ++sx)
#pragma unroll
    for (int sy = 0; sy < 2; ++sy) {
      q[(sx * 2 + sy) * 2 + 0] = make_double2(v[sx][sy], v[sx][sy + 1]);
      q[(sx * 2 + sy) * 2 + 1] = make_double2(v[sx + 1][sy], v[sx + 1][sy + 1]);
    }
}

// voxel i of a box (z fastest) -> index in the full grid
__device__ __forceinline__ int64_t box_voxel(const VoxelBox& b, int ny, int nz, int64_t i) {
  const int z = (int)(i % b.wz);
  const int64_t c = i / b.wz;
  const int y = (int)(c % b.wy), x = (int)(c / b.wy);
  return ((int64_t)(b.x0 + x) * ny + (b.y0 + y)) * nz + (b.z0 + z);
}
// SDFMap::resetBuffer(min_pos, max_pos) (sdf_map.cpp:28-53): occupancy 0, distance 10000 inside the box
// the same voxel in the distance field's own layout
__device__ __forceinline__ int64_t box_voxel_dist(const VoxelBox& b, const DistLayout& lay, int64_t i) {
  const int z = (int)(i % b.wz);
  const int64_t c = i / b.wz;
  return lay.at(b.x0 + (int)(c / b.wy), b.y0 + (int)(c % b.wy), b.z0 + z);
}
__global__ void k_box_reset(double* __restrict__ dist, uint8_t* __restrict__ occ, VoxelBox b, DistLayout lay, double dval,
                            int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  dist[box_voxel_dist(b, lay, i)] = dval;
  occ[box_voxel(b, lay.ny, lay.nz, i)] = 0;
}
__global__ void k_box_gather_occ(const uint8_t* __restrict__ occ, uint8_t* __restrict__ wocc, VoxelBox b, int ny, int nz,
                                 int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) wocc[i] = occ[box_voxel(b, ny, nz, i)];
}
// min(resolution*sqrt(val), old) (sdf_map.cpp:356-360) for the voxels of the box
__global__ void k_box_merge(const double* __restrict__ wdist, double* __restrict__ dist, VoxelBox b, DistLayout lay,
                            int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t a = box_voxel_dist(b, lay, i);
  dist[a] = fmin(wdist[i], dist[a]);
}

__global__ void k_set_occupancy(SdfView S, uint8_t* __restrict__ occ, const double* __restrict__ pts, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double px = pts[3 * i], py = pts[3 * i + 1], pz = pts[3 * i + 2];
  if (px < S.lox || py < S.loy || pz < S.loz || px > S.hix || py > S.hiy || pz > S.hiz) return;  // isInMap
  const int ix = (int)floor(__dmul_rn(__dsub_rn(px, S.ox), S.res_inv));                         // posToIndex
  const int iy = (int)floor(__dmul_rn(__dsub_rn(py, S.oy), S.res_inv));
  const int iz = (int)floor(__dmul_rn(__dsub_rn(pz, S.oz), S.res_inv));
  if (ix < 0 || iy < 0 || iz < 0 || ix >= S.nx || iy >= S.ny || iz >= S.nz) return;  // cannot happen in-map
  occ[((int64_t)ix * S.ny + iy) * S.nz + iz] = 1;
}

__global__ void k_set_boxes(int nx, int ny, int nz, uint8_t* __restrict__ occ, const int32_t* __restrict__ boxes,
                            int64_t nboxes) {
  // one CTA per box; threads sweep the box with z fastest
  for (int64_t bi = blockIdx.x; bi < nboxes; bi += gridDim.x) {
    const int32_t* b = boxes + 6 * bi;
    const int x0 = max(b[0], 0), y0 = max(b[1], 0), z0 = max(b[2], 0);
    const int x1 = min(b[3], nx - 1), y1 = min(b[4], ny - 1), z1 = min(b[5], nz - 1);
    if (x1 < x0 || y1 < y0 || z1 < z0) continue;
    const int64_t dx = x1 - x0 + 1, dy = y1 - y0 + 1, dz = z1 - z0 + 1, tot = dx * dy * dz;
    for (int64_t i = threadIdx.x; i < tot; i += blockDim.x) {
      const int z = z0 + (int)(i % dz), y = y0 + (int)((i / dz) % dy), x = x0 + (int)(i / (dz * dy));
      occ[((int64_t)x * ny + y) * nz + z] = 1;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// pass z: warp per column.  out[z] = (z - nearest occupied z')^2, or kInf when the column is empty.
__global__ void __launch_bounds__(256)
k_edt_z(const uint8_t* __restrict__ occ, int32_t* __restrict__ out, int64_t ncols, int nz) {
  const int lane = threadIdx.x & 31;
  const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int nchunk = (nz + 31) >> 5;
  for (int64_t col = wid; col < ncols; col += nw) {
    const uint8_t* __restrict__ o = occ + col * nz;
    int32_t* __restrict__ d = out + col * nz;
    // forward sweep: distance to the last occupied voxel at or below z
    int last = -kInf;  // "none yet"
    for (int c = 0; c < nchunk; ++c) {
      const int z = (c << 5) + lane;
      const bool oc = z < nz && o[z] != 0;
      const unsigned mask = __ballot_sync(0xffffffffu, oc);
      const unsigned below = mask & (0xffffffffu >> (31 - lane));
      const int near = below ? ((c << 5) + 31 - __clz(below)) : last;
      if (z < nz) d[z] = (near == -kInf) ? kInf : (z - near);
      if (mask) last = (c << 5) + 31 - __clz(mask);
    }
    // backward sweep: nearest occupied at or above z, combine, square
    int next = kInf;
    for (int c = nchunk - 1; c >= 0; --c) {
      const int z = (c << 5) + lane;
      const bool oc = z < nz && o[z] != 0;
      const unsigned mask = __ballot_sync(0xffffffffu, oc);
      const unsigned above = mask & (0xffffffffu << lane);
      const int near = above ? ((c << 5) + __ffs(above) - 1) : next;
      if (z < nz) {
        const int down = d[z];
        const int up = (near == kInf) ? kInf : (near - z);
        const int best = min(down, up);
        d[z] = (best >= kInf) ? kInf : best * best;
      }
      if (mask) next = (c << 5) + __ffs(mask) - 1;
    }
  }
}

// pass z for columns up to 32*NCH voxels: the column's occupancy lives in NCH ballot words held by every
// lane, both sweeps run on registers and every voxel is written once (the two-sweep kernel above writes,
// re-reads and re-writes its output: 3x the traffic).
template <int NCH>
__global__ void __launch_bounds__(256)
k_edt_z_reg(const uint8_t* __restrict__ occ, int32_t* __restrict__ out, int64_t ncols, int nz) {
  const int lane = threadIdx.x & 31;
  const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t col = wid; col < ncols; col += nw) {
    const uint8_t* __restrict__ o = occ + col * nz;
    int32_t* __restrict__ d = out + col * nz;
    unsigned mask[NCH];
    int down[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int z = (c << 5) + lane;
      mask[c] = __ballot_sync(0xffffffffu, z < nz && o[z] != 0);
    }
    int last = -kInf;  // nearest occupied voxel at or below, "none yet"
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const unsigned below = mask[c] & (0xffffffffu >> (31 - lane));
      const int near = below ? ((c << 5) + 31 - __clz(below)) : last;
      down[c] = (near == -kInf) ? kInf : ((c << 5) + lane - near);
      if (mask[c]) last = (c << 5) + 31 - __clz(mask[c]);
    }
    int next = kInf;  // nearest occupied voxel at or above
#pragma unroll
    for (int c = NCH - 1; c >= 0; --c) {
      const int z = (c << 5) + lane;
      const unsigned above = mask[c] & (0xffffffffu << lane);
      const int near = above ? ((c << 5) + __ffs(above) - 1) : next;
      const int up = (near == kInf) ? kInf : (near - z);
      const int best = min(down[c], up);
      if (z < nz) d[z] = (best >= kInf) ? kInf : best * best;
      if (mask[c]) next = (c << 5) + __ffs(mask[c]) - 1;
    }
  }
}

// pass z, vectorised: 4 voxels per lane (one 32-bit occupancy load, one 128-bit store), 128 per warp step, for
// columns with nz % 4 == 0 and nz <= 128*NCH.  The nearest occupied voxel below / above comes from a per-lane scan of
// its 4 voxels plus, for the other lanes, a bit scan of the ballot of occupied lanes and one shuffle of that lane's last /
// first occupied index, instead of one ballot round per 32 voxels: well under one warp instruction per voxel instead of
// 2.2 (the ballot kernel is issue-bound: 79 % issue utilisation at 24 % of DRAM bandwidth on the 2000x2000x400 map).
template <int NCH>
__global__ void __launch_bounds__(256)
k_edt_z_vec(const uint8_t* __restrict__ occ, int32_t* __restrict__ out, int64_t ncols, int nz) {
  const int lane = threadIdx.x & 31;
  const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  constexpr int kNone = 1 << 28;  // "no occupied voxel on that side"
  for (int64_t col = wid; col < ncols; col += nw) {
    const uint32_t* __restrict__ o = reinterpret_cast<const uint32_t*>(occ + col * nz);
    int4* __restrict__ d = reinterpret_cast<int4*>(out + col * nz);
    uint32_t w[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int z0 = c * 128 + lane * 4;
      w[c] = z0 < nz ? o[z0 >> 2] : 0u;
    }
    // lanes that hold an occupied voxel, per chunk: the nearest occupied voxel in ANOTHER lane is then one bit scan of
    // this mask plus one shuffle (a 5-step shuffle scan per sweep was a quarter of the kernel's instructions)
    unsigned m[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) m[c] = __ballot_sync(0xffffffffu, w[c] != 0u);
    // backward sweep: index of the nearest occupied voxel at or above each voxel (kNone if none)
    int up[NCH][4];
    int carry = kNone;
#pragma unroll
    for (int c = NCH - 1; c >= 0; --c) {
      const int z0 = c * 128 + lane * 4;
      // first occupied voxel of this lane
      int first = kNone;
#pragma unroll
      for (int j = 3; j >= 0; --j)
        if ((w[c] >> (8 * j)) & 0xffu) first = z0 + j;
      const unsigned higher = m[c] & (0xfffffffeu << lane);  // occupied lanes above this one
      const int v = __shfl_sync(0xffffffffu, first, higher ? __ffs(higher) - 1 : lane);
      int cur = higher ? v : carry;
#pragma unroll
      for (int j = 3; j >= 0; --j) {
        if ((w[c] >> (8 * j)) & 0xffu) cur = z0 + j;
        up[c][j] = cur;
      }
      const int lowest = __shfl_sync(0xffffffffu, first, m[c] ? __ffs(m[c]) - 1 : 0);
      carry = m[c] ? lowest : carry;
    }
    // forward sweep: nearest occupied at or below, combine, square, store
    carry = -kNone;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int z0 = c * 128 + lane * 4;
      int last = -kNone;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if ((w[c] >> (8 * j)) & 0xffu) last = z0 + j;
      const unsigned lower = m[c] & ((1u << lane) - 1u);  // occupied lanes below this one
      const int v = __shfl_sync(0xffffffffu, last, lower ? 31 - __clz(lower) : lane);
      int cur = lower ? v : carry;
      int r[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if ((w[c] >> (8 * j)) & 0xffu) cur = z0 + j;
        const int z = z0 + j;
        const int best = min(z - cur, up[c][j] - z);  // either side may be ~kNone away
        r[j] = best >= (kNone >> 1) ? kInf : best * best;
      }
      if (z0 < nz) d[z0 >> 2] = make_int4(r[0], r[1], r[2], r[3]);
      const int highest = __shfl_sync(0xffffffffu, last, m[c] ? 31 - __clz(m[c]) : 0);
      carry = m[c] ? highest : carry;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// ---- cheap exact arithmetic for the envelope passes (they are issue-bound: profiles/r1_microbench.md) ------------
// floor(num / den) for 0 <= num < 2^31, 0 < den < 2^31: quotient estimate from a Newton-refined FP64 reciprocal
// (relative error 2^-40, so the estimate is within 1 of the quotient), then an exact remainder fix-up.  ~13
// instructions against ~30 for the 32-bit integer division sequence.
__device__ __forceinline__ int floor_div_pos(int num, int den) {
  const double d = (double)den;
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  y = fma(y, fma(-d, y, 1.0), y);
  int q = __double2int_rz((double)num * y);
  const int r = num - q * den;
  q += (r >= den) ? 1 : 0;
  q -= (r < 0) ? 1 : 0;
  return q;
}
// sqrt((double)v) correctly rounded for integers 1 <= v <= kFastSqrtMax (= 3 * 8192^2, every value an envelope row
// of up to 8192 voxels can produce): Goldschmidt step from the MUFU seed plus one exact-residual correction.  The
// result differs from the true root by < 2^-75 relative before the final rounding; that it IS the correctly rounded
// value for every v in the range is checked exhaustively on the device (gtop_debug_check_sqrt, tests/test_gpu_sdf.py).
constexpr int kFastSqrtMax = 3 * 8192 * 8192;
__device__ __forceinline__ double sqrt_small_int(int v) {
  const double a = (double)v;
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  double t = a * y, h = 0.5 * y;
  const double r = fma(-t, h, 0.5);
  t = fma(t, r, t), h = fma(h, r, h);
  return fma(fma(-t, t, a), h, t);
}
// res * sqrt(val) as the reference computes it (sdf_map.cpp:356-360), bit for bit
// kLong: rows longer than 8192 voxels, whose squared distances may exceed kFastSqrtMax (a compile-time fact of the row
// bucket: the test costs 3 instructions per voxel)
template <bool kLong>
__device__ __forceinline__ double edt_distance(double res, int val) {
  if (kLong && val > kFastSqrtMax) return res * sqrt((double)val);
  const double r = sqrt_small_int(max(val, 1));            // occupied voxel: sqrt(0) = 0 without the seed's infinity
  return res * (val == 0 ? 0.0 : r);
}

// exhaustive check of sqrt_small_int against the IEEE square root: counts the integers in [1, vmax] where they differ
__global__ void k_check_sqrt(long long vmax, unsigned long long* mismatches) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  unsigned long long bad = 0;
  for (long long v = 1 + (long long)blockIdx.x * blockDim.x + threadIdx.x; v <= vmax; v += stride)
    bad += (sqrt_small_int((int)v) != sqrt((double)v)) ? 1ull : 0ull;
  if (bad) atomicAdd(mismatches, bad);
}
cudaError_t launch_check_sqrt(long long vmax, unsigned long long* mismatches_dev, cudaStream_t st) {
  k_check_sqrt<<<148 * 16, 256, 0, st>>>(vmax, mismatches_dev);
  return cudaGetLastError();
}

// 1-D lower envelope of parabolas on one strided row, exact integer arithmetic.
//   out[u] = min_i ( (u-i)^2 + g[i] ),   g[i] in [0, kInf]  (kInf = +inf)
// Hull entries (first index where the site wins, link to the site below, g(site)) packed in ONE 8-byte
// local-memory entry per site (rows are < 65535 long: 16-bit start and link), with the top entry held in
// registers: a push is one coalesced 8-byte store, a pop one 8-byte load, the common "compare with the top"
// touches no memory.
// (Three separate arrays cost three uncoalesced local accesses per step: the x pass of the 2000x2000x400
// map moved 184 GB through L1/L2 for 19 GB of useful traffic.)  F(x,i) = (x-i)^2 + g(i).
// Rows up to 8192 long keep every intermediate below 2^31 (u^2 + 3 n^2 <= 2.7e8), so the intersection
// is a 32-bit division; longer rows use 64-bit arithmetic.
// Output: out_i (int32, same stride as the input) or, FINAL, the fp64 distances: out_d points at the row's LAST element
// and walks down by gstride (linear field) or, BRICK, through the bricks of DistLayout: from x to x - 1 the address drops
// by 4 doubles inside a chunk (x odd), by 60 to the previous chunk of the brick, by `ojump` to the previous brick.
template <int MAXN, bool FINAL, bool BRICK = false>
__device__ __forceinline__ void envelope_row(const int32_t* __restrict__ g, int64_t gstride, int n,
                                             int32_t* __restrict__ out_i, double* __restrict__ out_d, double res,
                                             int fresh, int64_t ojump = 0) {
  typedef typename std::conditional<(MAXN <= 8192), int, long long>::type wide_t;
  // Hull sites as a linked list embedded in a per-site array: st[u] = (start_u | prev_u << 16, g(u)) is written when
  // site u is pushed — every lane of a warp is at the same u, so the store is coalesced (a compacted stack indexed by
  // depth scatters: the depth differs from lane to lane) — and a pop follows the link to a recently written entry.
  int2 st[MAXN];
  constexpr int kNoSite = 0xffff;
  int ts = kNoSite, tt = 0, tg = 0, tp = kNoSite;  // top site, its start, g(top), link to the site below
  // The row loads do not depend on the stack: fetch them kPF elements ahead so that several strided
  // loads are in flight while the envelope logic runs.
  constexpr int kPF = 8;
  int gbuf[kPF];
  const int32_t* __restrict__ gp = g;  // element u0 + kPF + k of the row is gp[0] when it is fetched (running pointer:
                                       // a 64-bit multiply per element was 10 % of the pass)
#pragma unroll
  for (int k = 0; k < kPF; ++k) {
    gbuf[k] = (k < n) ? *gp : kInf;
    gp += gstride;
  }
  for (int u0 = 0; u0 < n; u0 += kPF) {
    int gcur[kPF];
#pragma unroll
    for (int k = 0; k < kPF; ++k) {
      gcur[k] = gbuf[k];
      gbuf[k] = (u0 + kPF + k < n) ? *gp : kInf;
      gp += gstride;
    }
#pragma unroll
    for (int k = 0; k < kPF; ++k) {
      const int u = u0 + k, gu = gcur[k];
      if (gu >= kInf) continue;  // an infinite parabola never wins (also pads past the row end)
      // pop while the new parabola is below the top one at the top's start
      while (ts != kNoSite) {
        const int dq = tt - ts, du = tt - u;
        if ((wide_t)dq * dq + tg > (wide_t)du * du + gu) {
          ts = tp;
          if (ts != kNoSite) {
            const int2 e = st[ts];
            tt = e.x & 0xffff, tp = (unsigned)e.x >> 16, tg = e.y;
          }
        } else {
          break;
        }
      }
      if (ts == kNoSite) {
        ts = u, tt = 0, tg = gu, tp = kNoSite;
        st[u] = make_int2(tt | (tp << 16), tg);
      } else {
        // Sep(i,u) = floor((u^2 - i^2 + g(u) - g(i)) / (2(u-i))): last x where site i still wins
        // (num >= 0: the pops above leave a top with F(tt, ts) <= F(tt, u), i.e. num >= 2 tt (u - ts))
        const wide_t num = (wide_t)u * u - (wide_t)ts * ts + (wide_t)gu - (wide_t)tg;
        // (neighbouring sites are the common case on dense rows: a shift instead of the division sequence)
        const int w = 1 + ((u - ts == 1) ? (int)(num >> 1)
                                         : ((MAXN <= 8192) ? floor_div_pos((int)num, 2 * (u - ts)) : (int)(num / (wide_t)(2 * (u - ts)))));
        if (w < n) {
          tp = ts, ts = u, tt = w, tg = gu;
          st[u] = make_int2(tt | (tp << 16), tg);
        }
      }
    }
  }
  if (ts == kNoSite) {  // whole row infinite: res*sqrt(DBL_MAX) upstream, never below the old value (<= 10000)
    double* od = out_d;
    for (int u = n - 1; u >= 0; --u) {
      if (FINAL) {
        if (fresh) *od = 10000.0;
        od -= BRICK ? ((u & 1) ? (int64_t)4 : ((u & 7) ? (int64_t)60 : ojump)) : gstride;
      } else {
        out_i[(int64_t)u * gstride] = kInf;
      }
    }
    return;
  }
  double* od = FINAL ? out_d : nullptr;
  int32_t* oi = FINAL ? nullptr : out_i + (int64_t)(n - 1) * gstride;
  // min(resolution*sqrt(val), old)  (sdf_map.cpp:356-360).  `fresh` (warp-uniform) picks one of three copies of the
  // loop: 0 = read the old value; 1 = the buffer holds 10000 everywhere (just reset): no read; 2 = as 1, and no distance
  // of this grid can reach 10000: plain store (the min is 3 of ~110 instructions per voxel)
  auto fill = [&](auto mode_tag) {
    constexpr int kMode = decltype(mode_tag)::value;
    auto emit = [&](int u) {  // voxel u of the row: pop down to its site, evaluate, store (the caller moves od / oi)
      while (tp != kNoSite && tt > u) {
        ts = tp;
        const int2 e = st[ts];
        tt = e.x & 0xffff, tp = (unsigned)e.x >> 16, tg = e.y;
      }
      const int d = u - ts;
      const int val = d * d + tg;
      if (FINAL) {
        const double dnew = edt_distance<(MAXN > 8192)>(res, val);
        *od = kMode == 2 ? dnew : fmin(dnew, kMode == 1 ? 10000.0 : *od);
      } else {
        *oi = val;
      }
    };
    int u = n - 1;
    if (FINAL && BRICK) {
      // pairs (odd x, even x) share a 64-byte chunk: the step inside a pair is a constant, one select per pair
      if (!(u & 1)) {
        emit(u);
        od -= (u & 7) ? (int64_t)60 : ojump;
        --u;
      }
      for (; u >= 1; u -= 2) {
        emit(u);
        od -= 4;
        emit(u - 1);
        od -= ((u - 1) & 7) ? (int64_t)60 : ojump;
      }
    } else {
      for (; u >= 0; --u) {
        emit(u);
        if (FINAL) od -= gstride;
        else oi -= gstride;
      }
    }
  };
  if (FINAL && fresh == 2) fill(std::integral_constant<int, 2>{});
  else if (FINAL && fresh == 1) fill(std::integral_constant<int, 1>{});
  else fill(std::integral_constant<int, 0>{});
}

// pass y: row = (x,z); element y at ((x*ny + y)*nz + z)
template <int MAXN>
__global__ void __launch_bounds__(128)
k_edt_y(const int32_t* __restrict__ in, int32_t* __restrict__ out, int nx, int ny, int nz) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= (int64_t)nx * nz) return;
  const int x = (int)(r / nz), z = (int)(r % nz);
  const int64_t base = (int64_t)x * ny * nz + z;
  envelope_row<MAXN, false>(in + base, nz, ny, out + base, nullptr, 0.0, 0);
}

// pass x: row = (y,z); element x at (x*ny*nz + (y*nz + z)); writes the fp64 distance buffer
template <int MAXN>
__global__ void __launch_bounds__(128)
k_edt_x(const int32_t* __restrict__ in, double* __restrict__ dist, int nx, int ny, int nz, double res, int fresh) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t plane = (int64_t)ny * nz;
  if (r >= plane) return;
  envelope_row<MAXN, true>(in + r, plane, nx, nullptr, dist + r + (int64_t)(nx - 1) * plane, res, fresh);
}

// pass x into a field kept in 8 x 8 x 4 bricks (DistLayout).  A warp takes 4 y x 8 z rows: its loads of the linear int32
// input are four full 32-byte sectors per x, its stores eight full sectors (2 y x 2 z each) of two bricks; the other
// half of their 64-byte chunks follows with the next x.
template <int MAXN>
__global__ void __launch_bounds__(128)
k_edt_x_brick(const int32_t* __restrict__ in, double* __restrict__ dist, int nx, DistLayout lay, double res, int fresh) {
  const int lane = threadIdx.x & 31;
  const int nz8 = (lay.nz + 7) >> 3;
  const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);  // warp -> (group of 4 y, block of 8 z)
  const int y = (int)(w / nz8) * 4 + (lane >> 3), z = (int)(w % nz8) * 8 + (lane & 7);
  if (y >= lay.ny || z >= lay.nz) return;
  const int64_t plane = (int64_t)lay.ny * lay.nz;
  envelope_row<MAXN, true, true>(in + (int64_t)y * lay.nz + z, plane, nx, nullptr, dist + lay.at(nx - 1, y, z), res, fresh,
                                 ((int64_t)lay.nby * lay.nbz << 8) - (3 * 64 + 4));
}

template <int MAXN>
static cudaError_t launch_yx(const int32_t* d1, int32_t* d2, double* dist, int nx, int ny, int nz, double res,
                             bool fresh, cudaStream_t st, cudaEvent_t ev_mid, const DistLayout* lay) {
  const int threads = 128;
  // fill mode of the x pass (envelope_row): after a reset nothing has to be read, and if no distance of this grid can
  // reach the 10000 cap the min goes as well
  const double dmax = res * sqrt((double)nx * nx + (double)ny * ny + (double)nz * nz);
  const int mode = fresh ? (dmax * 1.000001 < 10000.0 ? 2 : 1) : 0;
  {
    const int64_t rows = (int64_t)nx * nz;
    k_edt_y<MAXN><<<(unsigned)((rows + threads - 1) / threads), threads, 0, st>>>(d1, d2, nx, ny, nz);
  }
  if (ev_mid) cudaEventRecord(ev_mid, st);
  if (lay && lay->bricks) {
    const int64_t warps = (int64_t)((ny + 3) / 4) * ((nz + 7) / 8);
    k_edt_x_brick<MAXN><<<(unsigned)((warps + 3) / 4), threads, 0, st>>>(d2, dist, nx, *lay, res, mode);
  } else {
    const int64_t rows = (int64_t)ny * nz;
    k_edt_x<MAXN><<<(unsigned)((rows + threads - 1) / threads), threads, 0, st>>>(d2, dist, nx, ny, nz, res, mode);
  }
  return cudaGetLastError();
}

cudaError_t launch_fill(double* dist, int64_t ndist, uint8_t* occ, int64_t n, double dval, cudaStream_t st) {
  k_fill<<<148 * 8, 256, 0, st>>>(dist, ndist, occ, n, dval);
  return cudaGetLastError();
}

cudaError_t launch_relayout(const double* src, double* dst, const DistLayout& lay, int nx, bool to_linear, cudaStream_t st) {
  const int64_t n = (int64_t)nx * lay.ny * lay.nz;
  k_relayout<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, dst, lay, nx, to_linear);
  return cudaGetLastError();
}

cudaError_t launch_build_quads(const SdfView& S, const double* dist, void* quad, cudaStream_t st) {
  const int64_t n = (int64_t)S.nxh * S.nyh * S.nz;
  k_build_quads<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dist, (double2*)quad, S.nx, S.ny, S.nz, S.nyh, n);
  return cudaGetLastError();
}

cudaError_t launch_build_quads_box(const SdfView& S, const double* dist, void* quad, const VoxelBox& b, cudaStream_t st) {
  // quads of voxels (x,y) with x in [x0-1, x0+wx-1], y likewise read a voxel of the box (x+ = x+1, y+ = y+1)
  const int xa = b.x0 > 0 ? b.x0 - 1 : 0, ya = b.y0 > 0 ? b.y0 - 1 : 0;
  const int pxa = xa >> 1, pxb = (b.x0 + b.wx - 1) >> 1, pya = ya >> 1, pyb = (b.y0 + b.wy - 1) >> 1;
  const int npx = pxb - pxa + 1, npy = pyb - pya + 1;
  const int64_t n = (int64_t)npx * npy * b.wz;
  k_build_quads_box<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dist, (double2*)quad, S.nx, S.ny, S.nz, S.nyh, pxa, pya,
                                                                  b.z0, npy, b.wz, n);
  return cudaGetLastError();
}

cudaError_t launch_box_reset(double* dist, uint8_t* occ, const VoxelBox& b, const DistLayout& lay, double dval, cudaStream_t st) {
  const int64_t n = (int64_t)b.wx * b.wy * b.wz;
  k_box_reset<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dist, occ, b, lay, dval, n);
  return cudaGetLastError();
}

cudaError_t launch_set_occupancy(const SdfView& S, uint8_t* occ, const double* pts_dev, int64_t n, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  k_set_occupancy<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(S, occ, pts_dev, n);
  return cudaGetLastError();
}

cudaError_t launch_set_boxes(int nx, int ny, int nz, uint8_t* occ, const int32_t* boxes_dev, int64_t n,
                             cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const unsigned grid = (unsigned)(n < 148 * 16 ? n : 148 * 16);
  k_set_boxes<<<grid, 256, 0, st>>>(nx, ny, nz, occ, boxes_dev, n);
  return cudaGetLastError();
}

// full EDT: occ -> d1 (z) -> d2 (y) -> dist (x).  ev[0..3] (nullable) bracket the three passes.  lay: layout of `dist`
// (nullptr = linear)
cudaError_t launch_edt(const uint8_t* occ, int32_t* d1, int32_t* d2, double* dist, int nx, int ny, int nz, double res,
                       bool fresh, cudaStream_t st, cudaEvent_t* ev, const DistLayout* lay) {
  const int64_t ncols = (int64_t)nx * ny;
  if (ev) cudaEventRecord(ev[0], st);
  {
    const int threads = 256;
    int64_t blocks = (ncols * 32 + threads - 1) / threads;
    if (blocks > 148 * 64) blocks = 148 * 64;
    if (nz % 4 == 0 && nz > 128 && nz <= 512) {
      if (nz <= 256) k_edt_z_vec<2><<<(unsigned)blocks, threads, 0, st>>>(occ, d1, ncols, nz);
      else k_edt_z_vec<4><<<(unsigned)blocks, threads, 0, st>>>(occ, d1, ncols, nz);
    } else if (nz <= 64) k_edt_z_reg<2><<<(unsigned)blocks, threads, 0, st>>>(occ, d1, ncols, nz);
    else if (nz <= 128) k_edt_z_reg<4><<<(unsigned)blocks, threads, 0, st>>>(occ, d1, ncols, nz);
    else k_edt_z<<<(unsigned)blocks, threads, 0, st>>>(occ, d1, ncols, nz);  // taller columns: the re-read of the
    // two-sweep kernel hits L1, and it measured faster than 8/16 ballot words in registers (4.05 vs 4.74 ms at nz=400)
  }
  if (ev) cudaEventRecord(ev[1], st);
  // (columns taller than 8192 voxels can produce squared distances above kFastSqrtMax: they take the long-row bucket,
  // the only one that keeps the range test of edt_distance)
  const int maxn = nz > 8192 ? 16384 : (nx > ny ? nx : ny);
  cudaError_t e;
  if (maxn <= 256)
    e = launch_yx<256>(d1, d2, dist, nx, ny, nz, res, fresh, st, ev ? ev[2] : nullptr, lay);
  else if (maxn <= 1024)
    e = launch_yx<1024>(d1, d2, dist, nx, ny, nz, res, fresh, st, ev ? ev[2] : nullptr, lay);
  else if (maxn <= 2048)
    e = launch_yx<2048>(d1, d2, dist, nx, ny, nz, res, fresh, st, ev ? ev[2] : nullptr, lay);
  else if (maxn <= 4096)
    e = launch_yx<4096>(d1, d2, dist, nx, ny, nz, res, fresh, st, ev ? ev[2] : nullptr, lay);
  else if (maxn <= 16384)
    e = launch_yx<16384>(d1, d2, dist, nx, ny, nz, res, fresh, st, ev ? ev[2] : nullptr, lay);
  else
    return cudaErrorInvalidValue;
  if (ev) cudaEventRecord(ev[3], st);
  return e;
}

// updateESDF3d restricted to the box [min_vec, max_vec] that setUpdateRange left (sdf_map.cpp:244-264, :310-368):
// the three envelope passes only run over the rows of the box and only see its occupied voxels, and the result is
// min'ed into the distance buffer inside the box.  Here: the occupancy of the box is copied into a compact
// (wx, wy, wz) grid, the SAME exact-integer EDT kernels run on that grid (`fresh`: values capped at 10000), and
// k_box_merge folds min(new, old) back — identical to the reference's arithmetic because the squared distances are
// exact integers whatever the coordinate origin of the rows.
cudaError_t launch_edt_window(const uint8_t* occ, uint8_t* wocc, int32_t* d1, int32_t* d2, double* wdist, double* dist,
                              const VoxelBox& b, const DistLayout& lay, double res, cudaStream_t st, cudaEvent_t* ev) {
  const int ny = lay.ny, nz = lay.nz;
  const int64_t n = (int64_t)b.wx * b.wy * b.wz;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  k_box_gather_occ<<<blocks, 256, 0, st>>>(occ, wocc, b, ny, nz, n);
  cudaError_t e = launch_edt(wocc, d1, d2, wdist, b.wx, b.wy, b.wz, res, true, st, ev, nullptr);  // compact grid: linear
  if (e != cudaSuccess) return e;
  k_box_merge<<<blocks, 256, 0, st>>>(wdist, dist, b, lay, n);
  if (ev) cudaEventRecord(ev[3], st);  // the merge belongs to the x pass
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// batched getDistWithGradTrilinear (reference formula, :221-239); same rounding discipline as the
// evaluation kernel for the index / diff arithmetic.
__global__ void k_query(SdfView S, const double* __restrict__ pos, int64_t n, double* __restrict__ dist,
                        double* __restrict__ grad, unsigned long long* __restrict__ oob) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double px = pos[3 * i], py = pos[3 * i + 1], pz = pos[3 * i + 2];
  if (px < S.lox || py < S.loy || pz < S.loz || px > S.hix || py > S.hiy || pz > S.hiz) {
    dist[i] = -1.0;
    grad[3 * i] = grad[3 * i + 1] = grad[3 * i + 2] = 0.0;
    atomicAdd(oob, 1ULL);
    return;
  }
  const double fx = floor(__dmul_rn(__dsub_rn(__dsub_rn(px, S.half_res), S.ox), S.res_inv));
  const double fy = floor(__dmul_rn(__dsub_rn(__dsub_rn(py, S.half_res), S.oy), S.res_inv));
  const double fz = floor(__dmul_rn(__dsub_rn(__dsub_rn(pz, S.half_res), S.oz), S.res_inv));
  const double dx = __dmul_rn(__dsub_rn(px, __dadd_rn(__dmul_rn(__dadd_rn(fx, 0.5), S.res), S.ox)), S.res_inv);
  const double dy = __dmul_rn(__dsub_rn(py, __dadd_rn(__dmul_rn(__dadd_rn(fy, 0.5), S.res), S.oy)), S.res_inv);
  const double dz = __dmul_rn(__dsub_rn(pz, __dadd_rn(__dmul_rn(__dadd_rn(fz, 0.5), S.res), S.oz)), S.res_inv);
  const int ix = (int)fx, iy = (int)fy, iz = (int)fz;
  const int x0 = max(min(ix, S.nx - 1), 0), x1 = max(min(ix + 1, S.nx - 1), 0);
  const int y0 = max(min(iy, S.ny - 1), 0), y1 = max(min(iy + 1, S.ny - 1), 0);
  const int z0 = max(min(iz, S.nz - 1), 0), z1 = max(min(iz + 1, S.nz - 1), 0);
  const double* D = S.dist;
  const DistLayout& L = S.lay;
  const double v000 = D[L.at(x0, y0, z0)], v001 = D[L.at(x0, y0, z1)];
  const double v010 = D[L.at(x0, y1, z0)], v011 = D[L.at(x0, y1, z1)];
  const double v100 = D[L.at(x1, y0, z0)], v101 = D[L.at(x1, y0, z1)];
  const double v110 = D[L.at(x1, y1, z0)], v111 = D[L.at(x1, y1, z1)];
  // unfused, reference order: results are bit-identical to the CPU path
  const double ex = __dsub_rn(1.0, dx), ey = __dsub_rn(1.0, dy), ez = __dsub_rn(1.0, dz);
  const double v00 = __dadd_rn(__dmul_rn(ex, v000), __dmul_rn(dx, v100));
  const double v01 = __dadd_rn(__dmul_rn(ex, v001), __dmul_rn(dx, v101));
  const double v10 = __dadd_rn(__dmul_rn(ex, v010), __dmul_rn(dx, v110));
  const double v11 = __dadd_rn(__dmul_rn(ex, v011), __dmul_rn(dx, v111));
  const double v0 = __dadd_rn(__dmul_rn(ey, v00), __dmul_rn(dy, v10));
  const double v1 = __dadd_rn(__dmul_rn(ey, v01), __dmul_rn(dy, v11));
  dist[i] = __dadd_rn(__dmul_rn(ez, v0), __dmul_rn(dz, v1));
  grad[3 * i + 2] = __dmul_rn(__dsub_rn(v1, v0), S.res_inv);
  grad[3 * i + 1] = __dmul_rn(
      __dadd_rn(__dmul_rn(ez, __dsub_rn(v10, v00)), __dmul_rn(dz, __dsub_rn(v11, v01))), S.res_inv);
  double g0 = __dmul_rn(__dmul_rn(ez, ey), __dsub_rn(v100, v000));
  g0 = __dadd_rn(g0, __dmul_rn(__dmul_rn(ez, dy), __dsub_rn(v110, v010)));
  g0 = __dadd_rn(g0, __dmul_rn(__dmul_rn(dz, ey), __dsub_rn(v101, v001)));
  g0 = __dadd_rn(g0, __dmul_rn(__dmul_rn(dz, dy), __dsub_rn(v111, v011)));
  grad[3 * i] = __dmul_rn(g0, S.res_inv);
}

// ---------------------------------------------------------------------------------------------
// batched EDTEnvironment::evaluateEDTWithGrad (src/edt_environment.cpp:75-122): the static field min'ed,
// at the eight corner voxels of the cell, with the distance to constant-velocity boxes at `time`
// (distToBox :27-44, evaluateConstVel obj_predictor.h:57-66), then the reference's trilinear value and gradient.
// Corner values go through getDistance(Vector3d) (sdf_map.cpp:155-164): -1 for a corner outside the map, not clamped.
// getInterpolationData is absent from the reference and taken from getDistWithGradTrilinear (sdf_map.cpp:196-209).
// boxes: nbox x 9 = centre at t = 0, velocity, full edge lengths.  Unfused, reference order: bit-identical to the oracle.
__global__ void k_query_dynamic(SdfView S, const double* __restrict__ boxes, int nbox, double time,
                                const double* __restrict__ pos, int64_t n, double* __restrict__ dist,
                                double* __restrict__ grad, double* __restrict__ coarse) {
  extern __shared__ double bx[];  // per box: min[3], max[3] at `time`
  for (int e = threadIdx.x; e < nbox * 3; e += blockDim.x) {
    const int b = e / 3, a = e - 3 * b;
    const double c = __dadd_rn(boxes[9 * b + a], __dmul_rn(time, boxes[9 * b + 3 + a]));  // 1.0*c0 + pow(t,1)*c1
    const double h = __dmul_rn(0.5, boxes[9 * b + 6 + a]);
    bx[6 * b + a] = __dsub_rn(c, h), bx[6 * b + 3 + a] = __dadd_rn(c, h);
  }
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  auto in_map = [&](double x, double y, double z) {
    return !(x < S.lox || y < S.loy || z < S.loz || x > S.hix || y > S.hiy || z > S.hiz);
  };
  auto min_box = [&](double x, double y, double z) {  // minDistToAllBox (:61-73)
    double best = 10000000.0;
    for (int b = 0; b < nbox; ++b) {
      const double* q = bx + 6 * b;
      const double ax = (x >= q[0] && x <= q[3]) ? 0.0 : fmin(fabs(__dsub_rn(x, q[0])), fabs(__dsub_rn(x, q[3])));
      const double ay = (y >= q[1] && y <= q[4]) ? 0.0 : fmin(fabs(__dsub_rn(y, q[1])), fabs(__dsub_rn(y, q[4])));
      const double az = (z >= q[2] && z <= q[5]) ? 0.0 : fmin(fabs(__dsub_rn(z, q[2])), fabs(__dsub_rn(z, q[5])));
      const double d = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay)), __dmul_rn(az, az)));
      if (d < best) best = d;
    }
    return best;
  };
  auto voxel_value = [&](double x, double y, double z) {  // getDistance(Vector3d)
    if (!in_map(x, y, z)) return -1.0;
    const int vx = (int)floor(__dmul_rn(__dsub_rn(x, S.ox), S.res_inv)), vy = (int)floor(__dmul_rn(__dsub_rn(y, S.oy), S.res_inv)),
              vz = (int)floor(__dmul_rn(__dsub_rn(z, S.oz), S.res_inv));
    return S.dist[S.lay.at(vx, vy, vz)];
  };
  const double px = pos[3 * i], py = pos[3 * i + 1], pz = pos[3 * i + 2];
  if (coarse) {  // evaluateCoarseEDT (:124-137)
    const double d1 = voxel_value(px, py, pz);
    coarse[i] = time < 0.0 ? d1 : fmin(d1, min_box(px, py, pz));
  }
  const double fx = floor(__dmul_rn(__dsub_rn(__dsub_rn(px, S.half_res), S.ox), S.res_inv));
  const double fy = floor(__dmul_rn(__dsub_rn(__dsub_rn(py, S.half_res), S.oy), S.res_inv));
  const double fz = floor(__dmul_rn(__dsub_rn(__dsub_rn(pz, S.half_res), S.oz), S.res_inv));
  const double dx = __dmul_rn(__dsub_rn(px, __dadd_rn(__dmul_rn(__dadd_rn(fx, 0.5), S.res), S.ox)), S.res_inv);
  const double dy = __dmul_rn(__dsub_rn(py, __dadd_rn(__dmul_rn(__dadd_rn(fy, 0.5), S.res), S.oy)), S.res_inv);
  const double dz = __dmul_rn(__dsub_rn(pz, __dadd_rn(__dmul_rn(__dadd_rn(fz, 0.5), S.res), S.oz)), S.res_inv);
  double v[2][2][2];
#pragma unroll
  for (int cx = 0; cx < 2; ++cx)
#pragma unroll
    for (int cy = 0; cy < 2; ++cy)
#pragma unroll
      for (int cz = 0; cz < 2; ++cz) {
        // corner position = indexToPos(idx + corner) (:76-78)
        const double qx = __dadd_rn(__dmul_rn(__dadd_rn(fx + cx, 0.5), S.res), S.ox);
        const double qy = __dadd_rn(__dmul_rn(__dadd_rn(fy + cy, 0.5), S.res), S.oy);
        const double qz = __dadd_rn(__dmul_rn(__dadd_rn(fz + cz, 0.5), S.res), S.oz);
        const double d1 = voxel_value(qx, qy, qz);
        v[cx][cy][cz] = time < 0.0 ? d1 : fmin(d1, min_box(qx, qy, qz));
      }
  const double ex = __dsub_rn(1.0, dx), ey = __dsub_rn(1.0, dy), ez = __dsub_rn(1.0, dz);
  const double v00 = __dadd_rn(__dmul_rn(ex, v[0][0][0]), __dmul_rn(dx, v[1][0][0]));
  const double v01 = __dadd_rn(__dmul_rn(ex, v[0][0][1]), __dmul_rn(dx, v[1][0][1]));
  const double v10 = __dadd_rn(__dmul_rn(ex, v[0][1][0]), __dmul_rn(dx, v[1][1][0]));
  const double v11 = __dadd_rn(__dmul_rn(ex, v[0][1][1]), __dmul_rn(dx, v[1][1][1]));
  const double v0 = __dadd_rn(__dmul_rn(ey, v00), __dmul_rn(dy, v10));
  const double v1 = __dadd_rn(__dmul_rn(ey, v01), __dmul_rn(dy, v11));
  dist[i] = __dadd_rn(__dmul_rn(ez, v0), __dmul_rn(dz, v1));
  grad[3 * i + 2] = __dmul_rn(__dsub_rn(v1, v0), S.res_inv);
  grad[3 * i + 1] = __dmul_rn(__dadd_rn(__dmul_rn(ez, __dsub_rn(v10, v00)), __dmul_rn(dz, __dsub_rn(v11, v01))), S.res_inv);
  double g0 = __dmul_rn(__dmul_rn(ez, ey), __dsub_rn(v[1][0][0], v[0][0][0]));
  g0 = __dadd_rn(g0, __dmul_rn(__dmul_rn(ez, dy), __dsub_rn(v[1][1][0], v[0][1][0])));
  g0 = __dadd_rn(g0, __dmul_rn(__dmul_rn(dz, ey), __dsub_rn(v[1][0][1], v[0][0][1])));
  g0 = __dadd_rn(g0, __dmul_rn(__dmul_rn(dz, dy), __dsub_rn(v[1][1][1], v[0][1][1])));
  grad[3 * i] = __dmul_rn(g0, S.res_inv);
}

cudaError_t launch_query_dynamic(const SdfView& S, const double* boxes_dev, int nbox, double time, const double* pos_dev,
                                 int64_t n, double* dist_dev, double* grad_dev, double* coarse_dev, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const size_t smem = sizeof(double) * 6 * (size_t)(nbox > 0 ? nbox : 1);
  if (smem > 48 * 1024) return cudaErrorInvalidValue;
  k_query_dynamic<<<(unsigned)((n + 255) / 256), 256, smem, st>>>(S, boxes_dev, nbox, time, pos_dev, n, dist_dev, grad_dev,
                                                                  coarse_dev);
  return cudaGetLastError();
}

cudaError_t launch_query(const SdfView& S, const double* pos_dev, int64_t n, double* dist_dev, double* grad_dev,
                         unsigned long long* oob_dev, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  k_query<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(S, pos_dev, n, dist_dev, grad_dev, oob_dev);
  return cudaGetLastError();
}

}  // namespace gtop
