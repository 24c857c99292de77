// eval_kernel.cu — batched cost + gradient of GTOP trajectories on sm_100a.
//
// Replaces GradTrajOptimizer::getCostAndGradient (src/grad_traj_optimizer.cpp:281-448 of the
// reference) for a batch of independent trajectories:
//   (a) smoothness  J_s = sum_a d_a' R d_a and 2 Rfp' df + 2 Rpp dp        (:302-336)  — banded R
//   (b) collision   sum over 30 (N_s) samples per segment of cd*|v|*dt     (:345-381)
//       with SDFMap::getDistWithGradTrilinear (src/sdf_map.cpp:185-242) gathers from HBM/L2
//   (c) coefficient map c = L d (getCoefficientFromDerivative :253-279) in per-segment 6x6 blocks
//
// Work decomposition: one CTA per trajectory, one warp per segment, one lane per sample.
// The 30 samples of a segment sit 1/30 of a segment apart, so the lanes of a warp gather from
// neighbouring voxels (shared 128-B lines of the z-fastest grid).  Per-sample work stays in
// registers; the 3x6 per-segment accumulators U[a][i] = sum_k (sA_ak tau_ik + sB_ak tau'_ik) are
// reduced across the warp and projected once through the segment's 6x6 block (O(m N_s) instead
// of the reference's dense O(m^2 N_s) products, :376-381).
//
// Rounding discipline: everything that feeds the float truncation of position /
// velocity (:457-485) and the voxel index / diff arithmetic (sdf_map.cpp:201-209) is evaluated in
// the reference's operation order with __dmul_rn/__dadd_rn (no FMA contraction), from the same
// host-computed pow() table, so cell selection and the fp32 round trip are bit-identical to the
// CPU path.  Everything downstream is smooth and free to fuse.
#include <cuda_runtime.h>

#include <cstdint>

#include "gtop_types.cuh"

namespace gtop {

__device__ __forceinline__ double f32_round_trip(double x) { return (double)__double2float_rn(x); }

// SDFMap::getDistWithGradTrilinear (sdf_map.cpp:185-242).  Out of map: dist=-1, grad=0 (H3).
__device__ __forceinline__ bool sdf_trilinear(const SdfView& S, double px, double py, double pz, double& dist,
                                              double& gx, double& gy, double& gz) {
  if (px < S.lox || py < S.loy || pz < S.loz || px > S.hix || py > S.hiy || pz > S.hiz) {
    dist = -1.0;
    gx = gy = gz = 0.0;
    return false;
  }
  // pos_m = pos - 0.5*res; idx = floor((pos_m - origin)*res_inv)           (:201-204, :71-74)
  const double fx = floor(__dmul_rn(__dsub_rn(__dsub_rn(px, S.half_res), S.ox), S.res_inv));
  const double fy = floor(__dmul_rn(__dsub_rn(__dsub_rn(py, S.half_res), S.oy), S.res_inv));
  const double fz = floor(__dmul_rn(__dsub_rn(__dsub_rn(pz, S.half_res), S.oz), S.res_inv));
  // idx_pos = (idx + 0.5)*res + origin; diff = (pos - idx_pos)*res_inv     (:76-78, :209)
  const double dx = __dmul_rn(__dsub_rn(px, __dadd_rn(__dmul_rn(__dadd_rn(fx, 0.5), S.res), S.ox)), S.res_inv);
  const double dy = __dmul_rn(__dsub_rn(py, __dadd_rn(__dmul_rn(__dadd_rn(fy, 0.5), S.res), S.oy)), S.res_inv);
  const double dz = __dmul_rn(__dsub_rn(pz, __dadd_rn(__dmul_rn(__dadd_rn(fz, 0.5), S.res), S.oz)), S.res_inv);
  const int ix = (int)fx, iy = (int)fy, iz = (int)fz;
  // clamped corner indices (getDistance(Vector3i) :166-174)
  const int x0 = max(min(ix, S.nx - 1), 0), x1 = max(min(ix + 1, S.nx - 1), 0);
  const int y0 = max(min(iy, S.ny - 1), 0), y1 = max(min(iy + 1, S.ny - 1), 0);
  const int z0 = max(min(iz, S.nz - 1), 0), z1 = max(min(iz + 1, S.nz - 1), 0);
  const int64_t sx = (int64_t)S.ny * S.nz, sy = S.nz;
  const double* __restrict__ D = S.dist;
  const double* r00 = D + x0 * sx + y0 * sy;
  const double* r01 = D + x0 * sx + y1 * sy;
  const double* r10 = D + x1 * sx + y0 * sy;
  const double* r11 = D + x1 * sx + y1 * sy;
  // values[x][y][z]
  const double v000 = __ldg(r00 + z0), v001 = __ldg(r00 + z1);
  const double v010 = __ldg(r01 + z0), v011 = __ldg(r01 + z1);
  const double v100 = __ldg(r10 + z0), v101 = __ldg(r10 + z1);
  const double v110 = __ldg(r11 + z0), v111 = __ldg(r11 + z1);

  const double ex = 1.0 - dx, ey = 1.0 - dy, ez = 1.0 - dz;
  const double v00 = ex * v000 + dx * v100;  // :221-224
  const double v01 = ex * v001 + dx * v101;
  const double v10 = ex * v010 + dx * v110;
  const double v11 = ex * v011 + dx * v111;
  const double v0 = ey * v00 + dy * v10;  // :226-227
  const double v1 = ey * v01 + dy * v11;
  dist = ez * v0 + dz * v1;  // :229
  gz = (v1 - v0) * S.res_inv;                                  // :231
  gy = (ez * (v10 - v00) + dz * (v11 - v01)) * S.res_inv;      // :232-233
  double g0 = ez * ey * (v100 - v000);                         // :234-239
  g0 += ez * dy * (v110 - v010);
  g0 += dz * ey * (v101 - v001);
  g0 += dz * dy * (v111 - v011);
  gx = g0 * S.res_inv;
  return true;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// dynamic shared memory layout (doubles):
//   W     [3][m+1][3]   boundary derivatives per axis in waypoint order (wp 0 = start, wp m = end)
//   gseg  [m][3][6]     collision gradient w.r.t. each segment's 6 local derivatives (natural order)
//   jc    [m]           collision cost per segment
//   cbuf  [nwarps][32]  per-warp coefficient staging: c[3][6] then e[3][4] (k*c_k), then U staging
//   red   [nwarps]      block reduction scratch
__global__ void __launch_bounds__(512, 1)
k_eval_traj(SdfView S, ProbLayout PL, const double* __restrict__ prob, int num_problems, EvalScalars E,
            const double* __restrict__ X, int64_t B, int64_t traj_offset, int64_t starts_per_problem,
            double* __restrict__ cost_out, double* __restrict__ grad_out, unsigned long long* __restrict__ oob_counter) {
  extern __shared__ double smem[];
  const int m = PL.m, P = PL.P, n = PL.n;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  double* W = smem;
  double* gseg = W + 3 * (m + 1) * 3;
  double* jc = gseg + m * 18;
  double* cbuf = jc + m;
  double* red = cbuf + nwarps * 32;

  for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
    const int64_t pid = ((traj_offset + b) / starts_per_problem) % num_problems;
    const double* __restrict__ pr = prob + pid * (int64_t)PL.stride;
    const double* __restrict__ x = X + b * n;

    // ---- 1. boundary derivatives in waypoint order --------------------------------------------
    for (int i = tid; i < 3 * (m + 1) * 3; i += blockDim.x) {
      const int a = i / ((m + 1) * 3), r = i % ((m + 1) * 3), wp = r / 3, q = r % 3;
      double v;
      if (wp == 0)
        v = pr[PL.off_Df + a * 6 + q];
      else if (wp == m)
        v = pr[PL.off_Df + a * 6 + 3 + q];
      else
        v = x[a * P + 3 * (wp - 1) + q];
      W[i] = v;
    }
    __syncthreads();

    // ---- 2. collision term: warp per segment, lane per sample ---------------------------------
    for (int s = warp; s < m; s += nwarps) {
      double* cb = cbuf + warp * 32;
      if (E.skip_collision) {
        if (lane < 18) gseg[s * 18 + lane] = 0.0;
        if (lane == 0) jc[s] = 0.0;
        continue;
      }
      const double* __restrict__ Li = pr + PL.off_Linv + s * 36;
      // coefficients c[a][i] = sum_r Linv[i][r] * dseg[a][r], summed in ascending dense-column
      // order of L (start derivatives first, except for the last segment whose end derivatives
      // live in columns 3..5 of d) with unfused mul/add — the same 6-term sum as `L * d` (:273).
      if (lane < 18) {
        const int a = lane / 6, i = lane % 6;
        const double* w0 = W + (a * (m + 1) + s) * 3;  // start wp: p,v,a ; end wp follows
        double p[6];
#pragma unroll
        for (int r = 0; r < 6; ++r) p[r] = __dmul_rn(Li[i * 6 + r], w0[r]);
        const bool last = (s == m - 1);
        double q0 = last ? p[3] : p[0], q1 = last ? p[4] : p[1], q2 = last ? p[5] : p[2];
        double q3 = last ? p[0] : p[3], q4 = last ? p[1] : p[4], q5 = last ? p[2] : p[5];
        double c = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(q0, q1), q2), q3), q4), q5);
        cb[lane] = c;
        // e_k = k*c_k for the velocity polynomial (2*coeff(s,2), 3*coeff(s,3), ... :477-485)
        if (i >= 2) cb[18 + a * 4 + (i - 2)] = __dmul_rn((double)i, c);
      }
      __syncwarp();

      const int ns = (int)pr[PL.off_ns + s];
      const double dt = pr[PL.off_dt + s];
      const double* __restrict__ tt = pr + PL.off_tt + s * 5 * PL.ns_cap;

      double U[3][6];
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int i = 0; i < 6; ++i) U[a][i] = 0.0;
      double Jc = 0.0;
      unsigned oob = 0;

      for (int k = lane; k < ns; k += 32) {
        const double t = tt[k], t2 = tt[PL.ns_cap + k], t3 = tt[2 * PL.ns_cap + k], t4 = tt[3 * PL.ns_cap + k],
                     t5 = tt[4 * PL.ns_cap + k];
        double pos[3], vel[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          const double* c = cb + a * 6;
          const double* e = cb + 18 + a * 4;
          // getPositionFromCoeff (:457-465): left-to-right, unfused, then float
          double pa = __dadd_rn(c[0], __dmul_rn(c[1], t));
          pa = __dadd_rn(pa, __dmul_rn(c[2], t2));
          pa = __dadd_rn(pa, __dmul_rn(c[3], t3));
          pa = __dadd_rn(pa, __dmul_rn(c[4], t4));
          pa = __dadd_rn(pa, __dmul_rn(c[5], t5));
          pos[a] = f32_round_trip(pa);
          // getVelocityFromCoeff (:477-485)
          double va = __dadd_rn(c[1], __dmul_rn(e[0], t));
          va = __dadd_rn(va, __dmul_rn(e[1], t2));
          va = __dadd_rn(va, __dmul_rn(e[2], t3));
          va = __dadd_rn(va, __dmul_rn(e[3], t4));
          vel[a] = f32_round_trip(va);
        }
        const double vnorm = sqrt(vel[0] * vel[0] + vel[1] * vel[1] + vel[2] * vel[2]) + 1e-5;  // :358

        double dist, g[3];
        if (!sdf_trilinear(S, pos[0], pos[1], pos[2], dist, g[0], g[1], g[2])) ++oob;

        const double ee = exp(-(dist - E.d0) * E.inv_r);
        const double cd = E.alpha * ee;           // :509
        const double gd = -E.alpha_over_r * ee;   // :514
        Jc += cd * vnorm;                         // :373 (dt applied per segment)

        const double tp[6] = {0.0, 1.0, 2.0 * t, 3.0 * t2, 4.0 * t3, 5.0 * t4};  // T*V (:379)
        const double ta[6] = {1.0, t, t2, t3, t4, t5};                           // T   (:544-551)
        const double inv_vn = 1.0 / vnorm;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          const double sA = gd * g[a] * cd * vnorm;  // :378 (extra cd reproduced, H4)
          const double sB = cd * (vel[a] * inv_vn);  // :379
#pragma unroll
          for (int i = 0; i < 6; ++i) U[a][i] += sA * ta[i] + sB * tp[i];
        }
      }

      // warp reduction of the 3x6 accumulators and the cost
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int i = 0; i < 6; ++i) U[a][i] = warp_sum(U[a][i]);
      Jc = warp_sum(Jc);
      if (oob_counter != nullptr) {
        oob = __reduce_add_sync(0xffffffffu, oob);
        if (lane == 0 && oob) atomicAdd(oob_counter, (unsigned long long)oob);
      }
      __syncwarp();
      if (lane == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int i = 0; i < 6; ++i) cb[a * 6 + i] = U[a][i];
        jc[s] = Jc * dt;
      }
      __syncwarp();
      // back-projection through the segment block: g[a][r] = sum_i U[a][i] * Linv[i][r]  (T*Ldp, :378-379)
      if (lane < 18) {
        const int a = lane / 6, r = lane % 6;
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < 6; ++i) acc += cb[a * 6 + i] * Li[i * 6 + r];
        gseg[s * 18 + lane] = acc * dt;
      }
      __syncwarp();
    }
    __syncthreads();

    // ---- 3. smoothness cost: banded column sums of R in waypoint order ------------------------
    double js_part = 0.0;
    const int nrow = 3 * (m + 1);
    for (int i = tid; i < 3 * nrow; i += blockDim.x) {
      const int a = i / nrow, u = i % nrow, wp = u / 3;
      const double* band = pr + PL.off_Rc + u * 9;
      const double* Wa = W + a * nrow;
      double y = 0.0;
#pragma unroll
      for (int j = 0; j < 9; ++j) {
        const int v = 3 * (wp - 1) + j;
        if (v >= 0 && v < nrow) y += Wa[v] * band[j];
      }
      js_part += y * Wa[u];
    }
    js_part = warp_sum(js_part);
    if (lane == 0) red[warp] = js_part;
    __syncthreads();

    // ---- 4. outputs ---------------------------------------------------------------------------
    double* __restrict__ gout = grad_out + b * n;
    for (int i = tid; i < n; i += blockDim.x) {
      const int a = i / P, j = i % P, wp = j / 3 + 1, q = j % 3;
      // smoothness gradient row (2 Rfp' df + 2 Rpp dp, :330-332)
      const double* band = pr + PL.off_Rg + j * 9;
      const double* Wa = W + a * nrow;
      double gs = 0.0;
#pragma unroll
      for (int jj = 0; jj < 9; ++jj) gs += Wa[3 * (wp - 1) + jj] * band[jj];
      gs *= 2.0;
      const double gc = gseg[(wp - 1) * 18 + a * 6 + 3 + q] + gseg[wp * 18 + a * 6 + q];
      gout[i] = (E.ws * gs + E.wc * gc) + 1e-5;  // :425-431
    }
    if (tid == 0) {
      double Js = 0.0, Jc = 0.0;
      for (int w = 0; w < nwarps; ++w) Js += red[w];
      for (int s = 0; s < m; ++s) Jc += jc[s];
      cost_out[b] = E.ws * Js + E.wc * Jc + 1e-3;  // :417-418
    }
    __syncthreads();
  }
}

size_t eval_smem_bytes(int m, int nwarps) {
  return sizeof(double) * (size_t)(3 * (m + 1) * 3 + m * 18 + m + nwarps * 32 + nwarps);
}

cudaError_t launch_eval(const SdfView& S, const ProbLayout& PL, const double* prob, int num_problems,
                        const EvalScalars& E, const double* X, int64_t B, int64_t traj_offset,
                        int64_t starts_per_problem, double* cost, double* grad, unsigned long long* oob_counter,
                        int max_ctas, cudaStream_t stream) {
  if (B <= 0) return cudaSuccess;
  const int nwarps = PL.m < 16 ? PL.m : 16;
  const int threads = nwarps * 32;
  const size_t smem = eval_smem_bytes(PL.m, nwarps);
  int64_t grid = B;
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
  k_eval_traj<<<(unsigned)grid, threads, smem, stream>>>(S, PL, prob, num_problems, E, X, B, traj_offset,
                                                          starts_per_problem, cost, grad, oob_counter);
  return cudaGetLastError();
}

}  // namespace gtop
