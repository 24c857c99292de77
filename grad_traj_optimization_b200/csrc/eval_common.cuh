// eval_common.cuh — device helpers shared by the two cost+gradient kernels (eval_kernel.cu: one
// quarter-warp per trajectory, any batch; eval_mma_kernel.cu: multi-start batches on FP64 DMMA).
//
// Rounding discipline: everything that feeds the float truncation of position /
// velocity (grad_traj_optimizer.cpp:457-485) and the voxel index / diff arithmetic
// (sdf_map.cpp:201-209) is evaluated in the reference's operation order with __dmul_rn/__dadd_rn
// (no FMA contraction), so cell selection and the fp32 round trip are bit-identical to the CPU
// path.  Everything downstream is smooth and free to fuse / reorder.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "gtop_types.cuh"

namespace gtop {

constexpr unsigned kFull = 0xffffffffu;

// ---- branch-free FP64 helpers for the smooth part of the sample pass (no special-case paths: the
// library sqrt/div/exp carry slow-path branches that split the sample body into basic blocks and
// keep the scheduler from overlapping the gather latency with arithmetic) ---------------------------
__device__ __forceinline__ double rcp_seed(double a) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));  // MUFU.RCP64H, ~20 good bits
  return r;
}
__device__ __forceinline__ double rsqrt_seed(double a) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));  // MUFU.RSQ64H, ~20 good bits
  return r;
}
// sqrt(a) for a in the normal range (a >= 1e-300): one coupled Newton (Goldschmidt) step from the
// 20-bit seed (-> ~2^-40) and one residual correction (-> < 1 ulp)
__device__ __forceinline__ double sqrt_pos(double a) {
  const double y = rsqrt_seed(a);
  double t = a * y, h = 0.5 * y;
  const double r = fma(-t, h, 0.5);
  t = fma(t, r, t), h = fma(h, r, h);
  return fma(fma(-t, t, a), h, t);
}
// 1/a for a in the normal range: seed (2^-20), cubic step e + e^2 (-> 2^-60)
__device__ __forceinline__ double rcp_pos(double a) {
  const double y = rcp_seed(a);
  const double e = fma(-a, y, 1.0);
  return fma(y, fma(e, e, e), y);
}

// 2^y for |y| < 2^26: y = k + j/16 + r, |r| <= 1/32; 2^y = 2^k * T[j] * 2^r with a degree-6 Taylor
// polynomial for 2^r (truncation (ln2/32)^7/7! = 4e-16).  T = 2^(j/16), 16 doubles = one conflict-free
// shared-memory wavefront per warp.  The exponent k is clamped to [-1000, 1000] in the integer domain
// (r stays the true remainder): results below 2^-1000 come out as ~2^-1000 instead of 0, which is
// far below anything the cost can resolve, and there is no FP64 min/max (a 6-instruction emulation).
constexpr int kExpTab = 16;
// ln2^k/k!, k = 6..1, in the constant bank: the DFMAs take them as c[][] operands (as immediates the
// compiler re-materialises each 64-bit constant with two UMOVs per use)
__constant__ double kExp2Poly[6] = {1.5403530393381609e-4, 1.3333558146428443e-3, 9.6181291076284772e-3,
                                    5.5504108664821580e-2, 2.4022650695910071e-1, 6.9314718055994531e-1};
// 2^(j/16), correctly rounded (a copy in shared memory serves the lanes' different j in one wavefront; the constant
// bank would serialise them)
__constant__ double kExp2Tab[kExpTab] = {
    0x1.0000000000000p+0, 0x1.0b5586cf9890fp+0, 0x1.172b83c7d517bp+0, 0x1.2387a6e756238p+0,
    0x1.306fe0a31b715p+0, 0x1.3dea64c123422p+0, 0x1.4bfdad5362a27p+0, 0x1.5ab07dd485429p+0,
    0x1.6a09e667f3bcdp+0, 0x1.7a11473eb0187p+0, 0x1.8ace5422aa0dbp+0, 0x1.9c49182a3f090p+0,
    0x1.ae89f995ad3adp+0, 0x1.c199bdd85529cp+0, 0x1.d5818dcfba487p+0, 0x1.ea4afa2a490dap+0};
__device__ __forceinline__ void exp2_tab_init(double* T, int tid, int nthreads) {
  for (int i = tid; i < kExpTab; i += nthreads) T[i] = kExp2Tab[i];
}
__device__ __forceinline__ double exp2_tab(double y, const double* __restrict__ T) {
  const double kMagic = 422212465065984.0;  // 1.5 * 2^48: ulp = 1/16
  const double z = y + kMagic;
  const int N = max(min(__double2loint(z), 16000), -16000);  // round(y*16), clamped
  const double r = y - (z - kMagic);
  double p = kExp2Poly[0];
#pragma unroll
  for (int i = 1; i < 6; ++i) p = fma(p, r, kExp2Poly[i]);
  const double tj = T[N & (kExpTab - 1)];
  const double m = fma(tj * r, p, tj);        // T[j] * (1 + r*p)
  return __hiloint2double(__double2hiint(m) + ((N >> 4) << 20), __double2loint(m));
}

__device__ __forceinline__ double f32_round_trip(double x) { return (double)__double2float_rn(x); }

// ---- 256-bit gather of one xy-quad ----------------------------------------------------------------
struct __align__(32) Quad {
  double v00, v01, v10, v11;  // (x0,y0) (x0,y1) (x1,y0) (x1,y1) at one height
};
// L1 policy of the corner gathers.  kNoAlloc: `L1::no_allocate` — with ~30 samples per segment the lines a
// warp gathers are not touched again by its next iteration (samples 4 apart are ~2.5 voxels apart), and
// keeping ~8 KB per warp and iteration out of L1 leaves it to the per-problem tables.  With dense sampling
// (200 samples per segment: 0.1 voxel apart) the next iterations hit the same lines, so those shapes allocate.
template <bool kNoAlloc>
__device__ __forceinline__ Quad ldg_quad(const void* p) {
  Quad q;
  if (kNoAlloc)
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(q.v00), "=d"(q.v01), "=d"(q.v10), "=d"(q.v11) : "l"(p));
  else
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(q.v00), "=d"(q.v01), "=d"(q.v10), "=d"(q.v11) : "l"(p));
  return q;
}
#ifndef GTOP_TABLE_L1
#define GTOP_TABLE_L1 1
#endif
// 128-bit load from a per-problem table (sample times, fragments): kept in L1 with evict_last
__device__ __forceinline__ double2 ldg_tab(const double2* p) {
#if GTOP_TABLE_L1 == 1
  double2 v;
  asm volatile("ld.global.nc.L1::evict_last.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
#else
  return __ldg(p);
#endif
}

// SDFMap::getDistWithGradTrilinear (sdf_map.cpp:185-242), split in two so that the caller can put
// independent arithmetic between the gathers and their use:
//   sdf_gather  : in-map test, cell index / diff (reference op order, unfused), 2 x LDG.256
//   sdf_finish  : trilinear value and gradient; out of map -> dist=-1, grad=0 (H3), branch-free
// The gradient is returned WITHOUT the 1/resolution factor (folded into the caller's constants).
// Interpolation is evaluated as nested lerps a + d*(b-a): the corner differences are shared between
// the value and the gradient; it differs from the reference's (1-d)*a + d*b form only by rounding.
//
// Clamped corners (getDistance(Vector3i) :166-174).  Inside the map idx >= -1 on every axis.  The
// quad element of voxel (x,y) already carries the upper clamp (x+ = min(x+1,nx-1)); for idx = -1
// both corners are voxel 0, which the quad of voxel 0 does not provide — `xlow`/`ylow` then evaluate
// the cell at diff = 0 (value of corner 0) and zero that gradient component, which is what two equal
// corners give.  Along z the two quads are addressed separately (z0, z1 clamped like the reference).
struct SdfCell {
  Quad q0, q1;  // heights z0, z1
  double dx, dy, dz;
  bool in_map, xlow, ylow;
  bool special;  // !in_map || xlow || ylow: needs the select path of sdf_finish (rare: map border)
};

// kPlain: the map keeps no quad copy (maps too big for one: the copy is 4x the field and has to be rebuilt after every
// update): the eight corners come straight from `dist`, which such maps keep in 8 x 8 x 4 bricks (DistLayout) — eight
// 64-bit loads where the quad copy needs two 256-bit loads; same values, same clamping.
template <bool kNoAlloc, bool kPlain = false>
__device__ __forceinline__ void sdf_gather(const SdfView& S, double px, double py, double pz, float fx32, float fy32,
                                           float fz32, SdfCell& c) {
  // isInMap (:55-69) on the float images of the bounds: exact for float-valued positions
  c.in_map = !(fx32 < S.lofx || fy32 < S.lofy || fz32 < S.lofz || fx32 > S.hifx || fy32 > S.hify || fz32 > S.hifz);
  // pos_m = pos - 0.5*res; idx = floor((pos_m - origin)*res_inv)           (:201-204, :71-74)
  const double tx = __dmul_rn(__dsub_rn(__dsub_rn(px, S.half_res), S.ox), S.res_inv);
  const double ty = __dmul_rn(__dsub_rn(__dsub_rn(py, S.half_res), S.oy), S.res_inv);
  const double tz = __dmul_rn(__dsub_rn(__dsub_rn(pz, S.half_res), S.oz), S.res_inv);
  const double fx = floor(tx), fy = floor(ty), fz = floor(tz);
#ifdef GTOP_DIFF_REFERENCE_ORDER
  // idx_pos = (idx + 0.5)*res + origin; diff = (pos - idx_pos)*res_inv     (:76-78, :209)
  c.dx = __dmul_rn(__dsub_rn(px, __dadd_rn(__dmul_rn(__dadd_rn(fx, 0.5), S.res), S.ox)), S.res_inv);
  c.dy = __dmul_rn(__dsub_rn(py, __dadd_rn(__dmul_rn(__dadd_rn(fy, 0.5), S.res), S.oy)), S.res_inv);
  c.dz = __dmul_rn(__dsub_rn(pz, __dadd_rn(__dmul_rn(__dadd_rn(fz, 0.5), S.res), S.oz)), S.res_inv);
#else
  // diff = (pos - ((idx + 0.5)*res + origin))*res_inv (:76-78, :209) is the fractional part of the scaled position
  // whose floor is idx: t - floor(t), exact as a subtraction.  It differs from the reference's five-operation form by
  // the rounding of t (<= 2^-53 * |t| ~ 2e-14 voxel at 200 voxels) and no discrete decision hangs on it — the
  // interpolation is continuous — so it is not kept in the reference's operation order (12 FP64 ops per sample less).
  c.dx = tx - fx, c.dy = ty - fy, c.dz = tz - fz;
#endif
  const int ix = (int)fx, iy = (int)fy, iz = (int)fz;
  c.xlow = ix < 0, c.ylow = iy < 0;
  c.special = !c.in_map || ((ix | iy) < 0);
  // the clamps also keep the loads of an out-of-map sample inside the buffer: no branch needed
  const int x0 = max(min(ix, S.nx - 1), 0), y0 = max(min(iy, S.ny - 1), 0);
  const int z0 = max(min(iz, S.nz - 1), 0), z1 = max(min(iz + 1, S.nz - 1), 0);
  if (kPlain) {
    const int x1 = min(x0 + 1, S.nx - 1), y1 = min(y0 + 1, S.ny - 1);
    // brick address = column part (x,y) + height part (z), see DistLayout
    const int bx0 = (x0 >> 3) * S.lay.nby, bx1 = (x1 >> 3) * S.lay.nby;
    const size_t c00 = ((size_t)(bx0 + (y0 >> 3)) * S.lay.nbz << 8) + DistLayout::col(x0, y0), c01 = ((size_t)(bx0 + (y1 >> 3)) * S.lay.nbz << 8) + DistLayout::col(x0, y1);
    const size_t c10 = ((size_t)(bx1 + (y0 >> 3)) * S.lay.nbz << 8) + DistLayout::col(x1, y0), c11 = ((size_t)(bx1 + (y1 >> 3)) * S.lay.nbz << 8) + DistLayout::col(x1, y1);
    const int h0 = ((z0 >> 2) << 8) | DistLayout::hgt(z0), h1 = ((z1 >> 2) << 8) | DistLayout::hgt(z1);
    const double* __restrict__ D = S.dist;
    c.q0.v00 = __ldg(D + c00 + h0), c.q0.v01 = __ldg(D + c01 + h0), c.q0.v10 = __ldg(D + c10 + h0), c.q0.v11 = __ldg(D + c11 + h0);
    c.q1.v00 = __ldg(D + c00 + h1), c.q1.v01 = __ldg(D + c01 + h1), c.q1.v10 = __ldg(D + c10 + h1), c.q1.v11 = __ldg(D + c11 + h1);
    return;
  }
  const int e0 = ((((x0 >> 1) * S.nyh + (y0 >> 1)) * S.nz + z0) << 2) + ((x0 & 1) << 1) + (y0 & 1);
  const char* p = (const char*)S.quad + (size_t)(unsigned)e0 * 32u;
  c.q0 = ldg_quad<kNoAlloc>(p);
  c.q1 = ldg_quad<kNoAlloc>(p + (z1 - z0) * 128);
}

// returns dist and the three gradient components (unscaled by 1/res).  The border cases (xlow / ylow /
// out of map) are patched behind one branch that is almost never taken.  kWarpUniform: the caller
// guarantees that all 32 lanes reach this call together (then the branch is made warp-uniform with a
// vote); otherwise every lane decides for itself.
template <bool kWarpUniform>
__device__ __forceinline__ void sdf_finish(const SdfCell& c, double& dist, double& gx, double& gy, double& gz,
                                           unsigned* oob = nullptr, bool counted = true) {
  double dx = c.dx, dy = c.dy;
  const bool any_special = kWarpUniform ? __any_sync(kFull, c.special) : c.special;
  if (__builtin_expect(any_special, 0)) {
    dx = c.xlow ? 0.0 : dx;
    dy = c.ylow ? 0.0 : dy;
  }
  // x-lerps; the x-differences feed grad x
  const double e00 = c.q0.v10 - c.q0.v00, e10 = c.q0.v11 - c.q0.v01;  // (y0,z0), (y1,z0)
  const double e01 = c.q1.v10 - c.q1.v00, e11 = c.q1.v11 - c.q1.v01;  // (y0,z1), (y1,z1)
  const double v00 = fma(dx, e00, c.q0.v00), v10 = fma(dx, e10, c.q0.v01);
  const double v01 = fma(dx, e01, c.q1.v00), v11 = fma(dx, e11, c.q1.v01);
  // y-lerps; y-differences feed grad y
  const double f0 = v10 - v00, f1 = v11 - v01;
  const double v0 = fma(dy, f0, v00), v1 = fma(dy, f1, v01);
  gz = v1 - v0;                                           // :231
  dist = fma(c.dz, gz, v0);                               // :229
  gy = fma(c.dz, f1 - f0, f0);                            // :232-233
  const double ex0 = fma(dy, e10 - e00, e00), ex1 = fma(dy, e11 - e01, e01);
  gx = fma(c.dz, ex1 - ex0, ex0);                         // :234-239
  if (__builtin_expect(any_special, 0)) {
    if (oob) *oob += (!c.in_map && counted) ? 1u : 0u;     // out-of-map samples are counted on the rare path only
    dist = c.in_map ? dist : -1.0;                        // :187, H3
    gx = (c.in_map && !c.xlow) ? gx : 0.0;
    gy = (c.in_map && !c.ylow) ? gy : 0.0;
    gz = c.in_map ? gz : 0.0;
  }
}


// One collision sample of getCostAndGradient (:353-381, optional block :383-407): position / velocity from the
// segment's coefficients `cb` (c[3][6] then e[3][4] = k*c_k), trilinear SDF, penalties, and the sample's
// contribution to the 20 per-segment sums v[] (U[a][i] at a*6+i, cost at 18, velocity/acceleration penalty cost
// at 19).  Shared by the general kernel and the single-trajectory kernel.
template <bool kVelAcc, bool kWarpUniform, bool kPlain = false>
__device__ __forceinline__ void sample_accumulate(const SdfView S, const EvalScalars E, const double kA,
                                                  const double* __restrict__ cb, double t, double t2, double t3, double t4,
                                                  double t5, const double* __restrict__ xtab, double (&v)[20],
                                                  unsigned& oob) {
  double pos[3], vel[3];
  float pf[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const double* c = cb + a * 6;
    // getPositionFromCoeff (:457-465): left-to-right, unfused, then float
    double pa = __dadd_rn(c[0], __dmul_rn(c[1], t));
    pa = __dadd_rn(pa, __dmul_rn(c[2], t2));
    pa = __dadd_rn(pa, __dmul_rn(c[3], t3));
    pa = __dadd_rn(pa, __dmul_rn(c[4], t4));
    pa = __dadd_rn(pa, __dmul_rn(c[5], t5));
    pf[a] = __double2float_rn(pa);
    pos[a] = (double)pf[a];
  }
  SdfCell cell;
  sdf_gather<false, kPlain>(S, pos[0], pos[1], pos[2], pf[0], pf[1], pf[2], cell);  // 2 x LDG.256 in flight from here
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const double* c = cb + a * 6;
    const double* e = cb + 18 + a * 4;
    // getVelocityFromCoeff (:477-485)
    double va = __dadd_rn(c[1], __dmul_rn(e[0], t));
    va = __dadd_rn(va, __dmul_rn(e[1], t2));
    va = __dadd_rn(va, __dmul_rn(e[2], t3));
    va = __dadd_rn(va, __dmul_rn(e[3], t4));
    vel[a] = f32_round_trip(va);
  }
  const double vn2 = fma(vel[2], vel[2], fma(vel[1], vel[1], fma(vel[0], vel[0], 1e-300)));
  const double vnorm = sqrt_pos(vn2) + 1e-5;  // :358 (the 1e-300 keeps v = 0 finite; absorbed by 1e-5)
  const double inv_vn = rcp_pos(vnorm);

  double dist, g[3];
  sdf_finish<kWarpUniform>(cell, dist, g[0], g[1], g[2]);
  oob += cell.in_map ? 0u : 1u;

  // exp(-(d-d0)/r) (:509/:514) = 2^((d0-d)*log2(e)/r)
  const double ee = exp2_tab(fma(dist, E.neg_log2e_over_r, E.d0_log2e_over_r), xtab);
  const double ev = ee * vnorm;
  v[18] += ev;                       // cost: cd*|v| = alpha*ee*vnorm (alpha, dt applied later)
  // sA_a = gd*grad_a*cd*|v| = kA * ee^2 * vnorm * g_a      (:378, extra cd reproduced — H4)
  // sB_a = cd*v_a/|v|       = alpha * ee * v_a / vnorm     (:379)
  const double sAc = kA * (ee * ev);
  const double sBc = E.alpha * (ee * inv_vn);
  const double tp2 = 2.0 * t, tp3 = 3.0 * t2, tp4 = 4.0 * t3, tp5 = 5.0 * t4;  // T*V (:379)
  double sBv[3] = {0.0, 0.0, 0.0}, sCv[3] = {0.0, 0.0, 0.0};
  if (kVelAcc) {
    // getAccelerationFromCoeff (:491-505): 2c2 + 6c3 t + 12c4 t^2 + 20c5 t^3, left to right, then float.
    // 6c3 = 2*(3c3), 12c4 = 3*(4c4), 20c5 = 4*(5c5) round identically from the e_k = k*c_k already held.
    double cv[3], ca[3], gv[3], gacc[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const double* e = cb + 18 + a * 4;
      double aa = __dadd_rn(e[0], __dmul_rn(2.0 * e[1], t));
      aa = __dadd_rn(aa, __dmul_rn(__dmul_rn(3.0, e[2]), t2));
      aa = __dadd_rn(aa, __dmul_rn(4.0 * e[3], t3));
      const double acc = f32_round_trip(aa);
      const double xv = exp2_tab(fma(fabs(vel[a]), E.log2e_over_rv, E.neg_v0_log2e_over_rv), xtab);
      const double xa = exp2_tab(fma(fabs(acc), E.log2e_over_ra, E.neg_a0_log2e_over_ra), xtab);
      cv[a] = E.alpha_v * xv, gv[a] = E.alpha_v_over_r * xv;    // :517-525
      ca[a] = E.alpha_a * xa, gacc[a] = E.alpha_a_over_r * xa;  // :527-535
    }
    v[19] += ((cv[0] + cv[1]) + cv[2] + (ca[0] + ca[1]) + ca[2]) * vnorm;  // cost_vel + cost_acc (:388-392)
    const double carried = (cv[2] + ca[2]) * inv_vn;  // `cv`, `ca` keep their k = 2 values in the second loop
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      // (gv*|v| + cv*v_a/|v| + ca*v_a/|v|) T V Ldp  and  ga*|v| T V V Ldp   (:396-404), weights wv = wa = 1
      sBv[a] = E.inv_wc * fma(gv[a], vnorm, carried * vel[a]);
      sCv[a] = E.inv_wc * (gacc[a] * vnorm);
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const double sA = sAc * g[a], sB = sBc * vel[a] + sBv[a];
    v[a * 6 + 0] += sA;
    v[a * 6 + 1] += fma(sA, t, sB);
    v[a * 6 + 2] += fma(sA, t2, sB * tp2);
    v[a * 6 + 3] += fma(sA, t3, sB * tp3);
    v[a * 6 + 4] += fma(sA, t4, sB * tp4);
    v[a * 6 + 5] += fma(sA, t5, sB * tp5);
    if (kVelAcc) {  // T*V*V = (0, 0, 2, 6t, 12t^2, 20t^3)
      v[a * 6 + 2] = fma(sCv[a], 2.0, v[a * 6 + 2]);
      v[a * 6 + 3] = fma(sCv[a], 3.0 * tp2, v[a * 6 + 3]);
      v[a * 6 + 4] = fma(sCv[a], 4.0 * tp3, v[a * 6 + 4]);
      v[a * 6 + 5] = fma(sCv[a], 5.0 * tp4, v[a * 6 + 5]);
    }
  }
}

}  // namespace gtop
