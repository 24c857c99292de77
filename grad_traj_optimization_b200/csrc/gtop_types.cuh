// gtop_types.cuh — PODs shared by the host side of libgtop_b200 and its sm_100a kernels.
#pragma once
#include <vector_types.h>

#include <cstdint>

namespace gtop {

// View of an SDFMap resident in HBM.  Index of voxel (x,y,z) is x*ny*nz + y*nz + z, as in the
// reference (sdf_map.cpp:92,97-98): z is the fastest axis.
struct SdfView {
  const double* dist;  // nx*ny*nz fp64 distances
  // z-pair copy of the same field: pair[i] = (dist[x,y,z], dist[x,y,min(z+1,nz-1)]) as one 16-byte
  // element, so the two z-corners of a trilinear cell come from ONE aligned 128-bit load whatever
  // the parity of z (z is the fastest axis).  2x the footprint (32 MB for the 200x200x50 map:
  // still L2-resident), half the gather instructions / L1 wavefronts.
  const double2* pair;
  int nx, ny, nz;
  double ox, oy, oz;        // origin
  double res, res_inv;      // resolution, 1/resolution (sdf_map.cpp:6-7)
  double half_res;          // 0.5*resolution (sdf_map.cpp:201)
  double lox, loy, loz;     // min_range + 1e-4   (isInMap, sdf_map.cpp:56-57)
  double hix, hiy, hiz;     // max_range - 1e-4   (sdf_map.cpp:62-63)
};

// Per-problem record in HBM: `stride` doubles per problem, sections at the offsets below.
//   dt    [m]            T_s / N_s                               (grad_traj_optimizer.cpp:351)
//   ns    [m]            number of samples of segment s (as double) — the `t < T_s` loop count (:353)
//   Linv  [m][6][6]      per-segment mapping block A_s^-1: row = power i, col = local derivative in
//                        NATURAL order (p0,v0,a0,pT,vT,aT)      (nonzero 6x6 block of L = A^-1 C^T)
//   tt    [m][5][ns_cap] sample-time table: t_k, t_k^2..t_k^5, host libm pow() (bit-exact with the
//                        reference's pow calls, :457-485,549)
//   Rc    [3m+3][9]      column band of R in waypoint order (smoothness cost, :326-327)
//   Rg    [3m-3][9]      band of (2 Rfp^T df + 2 Rpp dp) rows (smoothness gradient, :330-332)
//   Df    [3][6]         fixed derivatives per axis: p0,v0,a0,pEnd,vEnd,aEnd
//   LinvT [m][6][6]      transpose of Linv (col = power): the coefficient step reads 6 contiguous
//                        doubles per derivative instead of a stride-6 column
struct ProbLayout {
  int m;       // segments
  int P;       // free derivatives per axis = 3(m-1)
  int n;       // NLopt variables = 3P
  int ns_cap;  // samples_per_segment (table capacity)
  int off_dt, off_ns, off_Linv, off_tt, off_Rc, off_Rg, off_Df, off_LinvT;
  int stride;
};

inline ProbLayout make_layout(int m, int ns_cap) {
  ProbLayout L;
  L.m = m;
  L.P = 3 * (m - 1);
  L.n = 3 * L.P;
  L.ns_cap = ns_cap;
  int o = 0;
  L.off_dt = o, o += m;
  L.off_ns = o, o += m;
  L.off_Linv = o, o += m * 36;
  L.off_tt = o, o += m * 5 * ns_cap;
  L.off_Rc = o, o += (3 * m + 3) * 9;
  L.off_Rg = o, o += (3 * m - 3) * 9;
  L.off_Df = o, o += 18;
  L.off_LinvT = o, o += m * 36;
  L.stride = (o + 1) & ~1;  // keep every record 16-byte aligned
  return L;
}

// Scalars of one evaluation call.
struct EvalScalars {
  double ws, wc;          // weights after the step rule (ws = 0 in OPT_FIRST_STEP, :412-415)
  double alpha, d0;       // distance penalty (:507-515)
  double inv_r;           // 1/r
  double alpha_over_r;    // alpha/r
  double neg_log2e_over_r;  // -log2(e)/r      } exp(-(d-d0)/r) = 2^(d*neg_log2e_over_r + d0_log2e_over_r)
  double d0_log2e_over_r;   // d0*log2(e)/r    }
  int skip_collision;     // |wc| < 1e-4 (:346)
};

}  // namespace gtop
