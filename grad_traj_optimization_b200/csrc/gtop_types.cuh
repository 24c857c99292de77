// gtop_types.cuh — PODs shared by the host side of libgtop_b200 and its sm_100a kernels.
#pragma once
#include <vector_types.h>

#include <cstdint>

#if defined(__CUDACC__)
#define GTOP_HD __host__ __device__
#else
#define GTOP_HD
#endif

namespace gtop {

// Where voxel (x,y,z) of the fp64 distance field lives.
//   linear (bricks == 0): x*ny*nz + y*nz + z, as in the reference (sdf_map.cpp:92,97-98), z fastest — maps that also keep
//                         the xy-quad copy below (the evaluation kernels then never read `dist`);
//   bricks (bricks == 1): 8 x 8 x 4 (x,y,z) voxel bricks of 2 KB made of 2 x 2 x 2 voxel chunks of 64 bytes (what one L2
//                         miss fetches from HBM): [x>>3][y>>3][z>>2] [x>>1&3][y>>1&3][z>>1&1] [x&1][y&1][z&1] — maps too
//                         big for a quad copy, whose trilinear gathers go straight to the field in HBM.  In the linear
//                         order the four (x,y) columns of a cell are 8*nz bytes and 8*ny*nz bytes (6.4 MB on the
//                         2000x2000x400 map) apart: four sectors from four DRAM pages per sample.  Here a cell lies in
//                         1-8 chunks of one or two neighbouring bricks (3.4 on average, 1 when it is aligned), and
//                         consecutive samples of a trajectory (about a voxel apart) reuse them.  Measured
//                         (tools/bigmap_gather_bench.cu, 4,096 x 300 samples on 12.8 GB): 117 us linear, 57 us for
//                         bricks in plain [x&7][y&7][z&3] order, 49 us for this order (40 us when the same samples are
//                         gathered again: the touched lines fit L2).
// The occupancy and the EDT's int32 intermediates stay linear.
struct DistLayout {
  int ny, nz;    // grid
  int nby, nbz;  // bricks along y and z (ceil(ny/8), ceil(nz/4))
  int bricks;
  // offset of column (x,y) / of height z inside their brick (bit-disjoint)
  GTOP_HD static int col(int x, int y) { return (((x >> 1) & 3) << 6) | (((y >> 1) & 3) << 4) | ((x & 1) << 2) | ((y & 1) << 1); }
  GTOP_HD static int hgt(int z) { return (((z >> 1) & 1) << 3) | (z & 1); }
  GTOP_HD int64_t at(int x, int y, int z) const {
    if (bricks) return ((((int64_t)(x >> 3) * nby + (y >> 3)) * nbz + (z >> 2)) << 8) + (col(x, y) | hgt(z));
    return ((int64_t)x * ny + y) * nz + z;
  }
};
// doubles to allocate for a field of nx x ny x nz voxels (brick-padded)
inline int64_t dist_elems(int nx, int ny, int nz, bool bricks) {
  return bricks ? (int64_t)((nx + 7) / 8) * ((ny + 7) / 8) * ((nz + 3) / 4) * 256 : (int64_t)nx * ny * nz;
}

// View of an SDFMap resident in HBM.
struct SdfView {
  const double* dist;  // fp64 distances, laid out as `lay` says
  DistLayout lay;
  // xy-quad copy of the same field for the evaluation kernels' 256-bit gathers: one 32-byte element
  // per voxel, quad(x,y,z) = (D[x,y,z], D[x,y+,z], D[x+,y,z], D[x+,y+,z]) with x+ = min(x+1,nx-1),
  // y+ = min(y+1,ny-1): the four (x,y) corners of a trilinear cell at one height are ONE aligned
  // LDG.256, a sample needs two (z0, z1).  Elements are stored in 2x2 (x,y) patches,
  //   index = ((((x>>1)*nyh + (y>>1))*nz + z) << 2) + ((x&1)<<1) + (y&1),
  // so a 128-byte line holds the quads of a 2x2 patch at one height and the z-neighbour is the next
  // line: consecutive samples of a trajectory (about half a voxel apart, mostly horizontal motion)
  // fall into the same lines.  Measured on B200 (tools/gather_bench.cu): 35 L1 wavefronts per
  // 32-sample gather against 76 for four 128-bit loads from a z-pair layout.  4x the footprint of
  // `dist` (64 MB for the 200x200x50 map).
  const void* quad;
  int nx, ny, nz;
  int nxh, nyh;             // ceil(nx/2), ceil(ny/2): patch grid of the quad copy
  double ox, oy, oz;        // origin
  double res, res_inv;      // resolution, 1/resolution (sdf_map.cpp:6-7)
  double half_res;          // 0.5*resolution (sdf_map.cpp:201)
  double lox, loy, loz;     // min_range + 1e-4   (isInMap, sdf_map.cpp:56-57)
  double hix, hiy, hiz;     // max_range - 1e-4   (sdf_map.cpp:62-63)
  // the same bounds for a float-valued position p: p < lo  <=>  p < lof (smallest float >= lo),
  // p > hi <=> p > hif (largest float <= hi).  Sample positions are float-valued (:457-467).
  float lofx, lofy, lofz, hifx, hify, hifz;
};

// a box of voxels [x0, x0+wx) x [y0, y0+wy) x [z0, z0+wz): the update window of SDFMap (min_vec..max_vec,
// sdf_map.cpp:13-14, :244-264)
struct VoxelBox {
  int x0, y0, z0, wx, wy, wz;
};

// number of 32-byte quad elements of a grid (patch-padded)
inline int64_t quad_elems(int nx, int ny, int nz) { return (int64_t)((nx + 1) / 2) * ((ny + 1) / 2) * nz * 4; }

// Per-problem record in HBM: `stride` doubles per problem, sections at the offsets below.
//   dt    [m]            T_s / N_s                               (grad_traj_optimizer.cpp:351)
//   ns    [m]            number of samples of segment s (as double) — the `t < T_s` loop count (:353)
//   Linv  [m][6][6]      per-segment mapping block A_s^-1: row = power i, col = local derivative in
//                        NATURAL order (p0,v0,a0,pT,vT,aT)      (nonzero 6x6 block of L = A^-1 C^T)
//   tt    [m][5][ns_cap] sample-time table: t_k, t_k^2..t_k^5, host libm pow() (bit-exact with the
//                        reference's pow calls, :457-485,549)
//   Rc    [3m+3][9]      column band of R in waypoint order (smoothness cost, :326-327)
//   Rg    [3m-3][9]      band of (2 Rfp^T df + 2 Rpp dp) rows (smoothness gradient, :330-332)
//   Df    [3][6]         fixed derivatives per axis: p0,v0,a0,pEnd,vEnd,aEnd (16-byte aligned)
//   LinvT [m][6][6]      transpose of Linv (col = power): the coefficient step reads 6 contiguous
//                        doubles per derivative instead of a stride-6 column
// Sections read by the multi-start kernel k_eval_mma (eval_mma_kernel.cu).  The kernel walks the segments in PAIRS
// (2p, 2p+1): the ns_a + ns_b samples of a pair are numbered consecutively (slot q < ns_a: sample q of the first
// segment, else sample q - ns_a of the second) and iteration u of the pair evaluates slots 4u + j (j = lane & 3), so
// that a sample count that is not a multiple of 4 (30 in the reference) leaves idle slots once per pair, not once
// per segment.  npair = ceil(m/2), nit2 = ceil(2 ns_cap / 4) iterations per pair:
//   TQ    [npair][nit2][3][4][2]   powers of the sample times in the order the lanes read them: part 0/1/2 of slot j =
//                        (dt_s,t) (t^2,t^3) (t^4,t^5) — the quadrature weight dt_s of the slot's segment sits where t^0
//                        would —, the four slots of a part contiguous (one 64-byte L1 wavefront per
//                        128-bit load).  Idle slots are all zero: they evaluate a segment's start point.
//   PL    [npair][nit2+1][32][2]   B fragments of the gradient DMMAs, one 16-byte pair per lane = 4 g + j of iteration u:
//                        (dt_s * phi_r(t_k), dt_s * phi_r'(t_k)) of the slot's segment s and sample k, phi_r(t) =
//                        sum_i t^i Linv_s[i][r] the basis function of local derivative r = deriv_of_col(g, s) (T Ldp
//                        and T V Ldp of :376-381, which depend on the problem only).  Column g <-> derivative r
//                        alternates with the segment parity so that the end-derivative columns of segment s are the
//                        start-derivative columns of s+1.  Idle slots and columns 6, 7 are zero: they contribute
//                        nothing.  In the one iteration that mixes both segments of a pair (u = ns_a / 4 when ns_a is
//                        not a multiple of 4) record u carries the first segment's lanes only and record nit2 (the
//                        extra one) the second segment's: the kernel issues the DMMAs once per record and closes the
//                        first segment's waypoint rows in between.
//   LiTP  [m][3][6][2]   (Linv[2p][r], Linv[2p+1][r]): coefficient step, two powers per 128-bit load
//   T     [m]            segment times T_s (segment_time, grad_traj_optimizer.cpp:73-81): post-hoc metrics kernel
//   RgP   [3m-3][10]     the rows of Rg padded to 80 bytes: five aligned 128-bit loads per row in k_eval_mma
struct ProbLayout {
  int m;       // segments
  int P;       // free derivatives per axis = 3(m-1)
  int n;       // NLopt variables = 3P
  int ns_cap;  // samples_per_segment (table capacity)
  int nsp;     // ns_cap rounded up to a multiple of 4
  int npair;   // segment pairs = ceil(m/2)
  int nit2;    // iterations of k_eval_mma per pair at full sample count = ceil(2 ns_cap / 4)
  int off_dt, off_ns, off_Linv, off_tt, off_Rc, off_Rg, off_Df, off_LinvT;
  int off_LiTP, off_TQ, off_PL, off_T, off_RgP;
  int stride;
};

// local derivative (natural order p0,v0,a0,pT,vT,aT) carried by column c of the gradient DMMAs
// of segment s; -1 for the two padding columns
GTOP_HD inline int deriv_of_col(int c, int s) { return c >= 6 ? -1 : ((s & 1) ? (c + 3) % 6 : c); }

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
  o = (o + 1) & ~1;  // 16-byte aligned: k_eval_mma fetches Df with a bulk copy
  L.off_Df = o, o += 18;
  L.off_LinvT = o, o += m * 36;
  o = (o + 1) & ~1;
  L.nsp = (ns_cap + 3) & ~3;
  o = (o + 7) & ~7;  // 64-byte aligned slots (relative to the 16-byte aligned record)
  L.off_LiTP = o, o += m * 36;
  o = (o + 7) & ~7;
  L.npair = (m + 1) / 2;
  L.nit2 = (2 * ns_cap + 3) / 4;
  L.off_TQ = o, o += L.npair * L.nit2 * 24;
  L.off_PL = o, o += L.npair * (L.nit2 + 1) * 64;
  L.off_T = o, o += m;
  o = (o + 1) & ~1;
  L.off_RgP = o, o += (3 * m - 3) * 10;
  L.stride = (o + 7) & ~7;  // keep every record 64-byte aligned
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
  // optional velocity / acceleration penalties (the block commented out upstream, :383-407, :517-535);
  // vel_acc != 0 only when gtop_params.vel_acc_penalty is set and step == OPT_SECOND_STEP
  int vel_acc;
  double alpha_v, alpha_v_over_r, log2e_over_rv, neg_v0_log2e_over_rv;  // cv = alpha_v * 2^(|v|*k + c)
  double alpha_a, alpha_a_over_r, log2e_over_ra, neg_a0_log2e_over_ra;
  double inv_wc;          // 1/wc: the penalty gradients (weight 1) ride in the collision accumulators (weight wc)
};

}  // namespace gtop
