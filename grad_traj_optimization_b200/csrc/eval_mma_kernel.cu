// eval_mma_kernel.cu — cost + gradient of GTOP trajectories for MULTI-START batches on sm_100a, segment pairs
// in parallel: a CTA owns an aligned group of 8 consecutive trajectories (which share one problem: the same
// waypoints / segment times — BASELINE configs[2] = 64 waypoint sets x 1,024 starts, configs[4]) and its warps take
// the PAIRS of segments of those trajectories.
//
// Replaces GradTrajOptimizer::getCostAndGradient (src/grad_traj_optimizer.cpp:281-448 of the reference).
//
// Why pairs in parallel: with one warp walking all segments of 8 trajectories (the round-1 kernel) a batch of B
// trajectories is B/8 warps of 7.5 m dependent sample iterations each; the 8,192-trajectory shard a GPU gets when the
// 65,536 batch is split over 8 GPUs is 1,024 warps on 148 SMs x 20 resident warps: a third of one wave.  Here the
// same shard is 1,024 groups x 5 warps of 15 iterations.  The segments of a trajectory only meet in the gradient rows
// of the waypoints they share and in two scalar sums.
//
//   in      one bulk copy (TMA, cp.async.bulk + mbarrier) brings the 8 x n doubles of the group's x and the problem's
//           fixed boundary state into shared memory: one instruction, no register staging, no address arithmetic
//   pair    warp = pair p = segments (2p, 2p+1): coefficients c = L d (:253-279, reference summation order, unfused),
//           then the sample loop.  Lane (g, j) = (lane >> 2, lane & 3) evaluates slot 4 it + j of trajectory g; the
//           samples of the two segments are numbered consecutively, so 30 samples per segment idle two lanes once per
//           pair (15 iterations, not 16).  Per sample: position / velocity with the reference's float truncation
//           (:451-488), trilinear SDF value and gradient from two 256-bit gathers of the patch-ordered xy-quad copy of
//           the grid (sdf_map.cpp:185-242), exponential penalty (:507-515), and the gradient contribution
//               G[a][r] += dt (sA_a phi_r(t) + sB_a phi_r'(t)),  phi_r(t) = sum_i t^i Linv[i][r]
//           ((gd grad_a cd |v| T + cd v_a/|v| T V) Ldp dt of :376-381) as six FP64 tensor-core tiles
//           (mma.sync.m8n8k4.f64, SASS DMMA): A fragment = the lane's own sA_a / sB_a, B fragment = a per-problem table
//           of dt phi_r(t_k), dt phi_r'(t_k), C fragment = the rows in flight.  The DMMA sums over the 4 lanes of a
//           trajectory: no shuffle reduction, no projection through L.  The columns alternate with the segment parity,
//           so the rows of the waypoint INSIDE the pair (2p+1) accumulate in place across its two segments.
//   rows    the pair owns the output rows of waypoints 2p+1 and 2p+2: smoothness rows from the band of R
//           (2 Rfp' df + 2 Rpp dp, :330-332; their product with d is the rows' share of d'Rd, :326-327) plus the
//           collision rows (:425-431).  Waypoint 2p+2 also collects the START derivatives of segment 2p+2, which the
//           next pair publishes in shared memory behind a flag half-way through its own samples: the warps of a CTA
//           never meet at a barrier.
//   out     gradient rows are staged in shared memory in x's layout; the last warp to finish adds up the cost and
//           sends the rows off with one bulk store.
// One CTA per group (a persistent variant that walked several groups per CTA with the next x prefetched measured
// slower — 0.380 against 0.357 ms per 65,536 — and the grid order keeps co-resident CTAs on neighbouring starts).
//
// Every sum has ONE shape, fixed by the problem alone (per-pair partials, combined in pair order), so neither the
// number of warps per CTA nor the split of a batch over calls or GPUs changes a bit of the result.
//
// DMMA shares the FP64 pipe with DFMA on B200 (tools/fp64_peak.cu: 37.1 vs 34.2 TFLOP/s, mixed = sum of both): the
// tiles buy issue slots and registers (6 DMMA instead of ~50 DFMA/DMUL + shuffles per sample), not FLOPs.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>

#include "eval_common.cuh"
#include "gtop_types.cuh"

namespace gtop {

constexpr int kPairMaxWarps = 10;  // 96 registers x 20 resident warps per SM: 2 CTAs of 10, 4 of 5, ...

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// ---- TMA bulk copies and their mbarrier (PTX ISA: cp.async.bulk, mbarrier) -------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// x and the gradient stream through once: evict-first in L2, so that they do not push out the map and the per-problem
// tables, which every launch re-reads
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src_gmem, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_load_stream(void* dst_smem, const void* src_gmem, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(l2_evict_first_policy())
               : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst_gmem, const void* src_smem, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst_gmem),
               "r"(smem_u32(src_smem)), "r"(bytes), "l"(l2_evict_first_policy())
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

// shared memory of a CTA, in doubles (coef first: 16-byte aligned pairs; every section starts 16-byte aligned)
//   coef [W][2][8][15][2]  coefficient pairs of the two segments a warp is working on, 15 16-byte elements per
//                          trajectory: per axis (c0,c1) (c2,c3) (c4,c5) (2c2,3c3) (4c4,5c5); the 240-byte trajectory
//                          stride puts the 8 trajectories of a warp in 8 distinct 16-byte bank groups
//   xin  [8 n + 18]        x of the group's 8 trajectories as it lies in global memory (trajectory stride n), then the
//                          problem's fixed boundary state Df[3][6]
//   out  [8 n]             gradient rows of the group in x's layout
//   scr  [W][2][8][9]      collision rows of waypoints 2p+1 and 2p+2 (its END-derivative part), [axis][p,v,a] per
//                          trajectory (warp-private)
//   exch [npair][8][9]     collision rows of waypoint 2p from the START derivatives of segment 2p: read by pair p-1
//   jsp  [8][npair]        d_u (R d)_u summed over the rows a pair owns
//   jcp  [8][npair]        sum of cd |v| dt over the samples of a pair
//   xtab [16]              2^(j/16)
//   mbar [1]               mbarrier: x has landed
//   flag [npair + 1]       (32-bit) exch[p] is published; the last entry counts the warps that are done
struct PairSmem {
  int coef, xin, out, scr, exch, jsp, jcp, xtab, mbar, flag, total;
};
__host__ __device__ inline PairSmem pair_smem(int m, int W) {
  PairSmem s;
  const int npair = (m + 1) / 2, n = 9 * (m - 1);
  int o = 0;
  s.coef = o, o += W * 480;
  s.xin = o, o += (8 * n + 18 + 1) & ~1;
  s.out = o, o += (8 * n + 1) & ~1;
  s.scr = o, o += W * 144;
  s.exch = o, o += npair * 72;
  s.jsp = o, o += 8 * npair;
  s.jcp = o, o += 8 * npair;
  s.xtab = o, o += kExpTab;
  s.mbar = o, o += 2;
  s.flag = o, o += (npair + 2) / 2;
  s.total = (o + 1) & ~1;
  return s;
}

// boundary derivatives of one waypoint of one trajectory: value (axis a, derivative r) = p[a * s + r].  Interior
// waypoints live in x (axis stride P), the fixed start / end state in Df (axis stride 6).
struct Wp {
  const double* p;
  int s;
};

// coefficients c[a][i] = sum_r Linv[i][r] * dseg[a][r], summed in ascending dense-column order of L (start
// derivatives first, except for the last segment whose end derivatives live in columns 3..5 of d) with unfused
// mul/add — the same 6-term sum as `L * d` (:273).  Lane j < 3 takes the power pair (2j, 2j+1) of all three axes: its six
// Linv pairs are loaded once, and the four lanes of a trajectory read the same waypoint values (broadcast).
// Lpre: the lane's six Linv pairs if the caller has loaded them already (under the copy of x), else nullptr
template <bool kLast>
__device__ __forceinline__ void pair_coef_phase(const Wp wp0, const Wp wp1, const double2* __restrict__ LiTP,
                                                double2* __restrict__ coef, int j, const double2* Lpre = nullptr) {
  if (j >= 3) return;
  const int p = j;
  double2 L[6];
#pragma unroll
  for (int r = 0; r < 6; ++r) L[r] = Lpre ? Lpre[r] : ldg_tab(LiTP + p * 6 + r);
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const double *ws = wp0.p + a * wp0.s, *we = wp1.p + a * wp1.s;  // (p,v,a) of the segment's start / end waypoint
    double q0[6], q1[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      const double w = r < 3 ? ws[r] : we[r - 3];
      q0[r] = __dmul_rn(L[r].x, w), q1[r] = __dmul_rn(L[r].y, w);
    }
    double c0, c1;
    if (kLast) {
      c0 = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(q0[3], q0[4]), q0[5]), q0[0]), q0[1]), q0[2]);
      c1 = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(q1[3], q1[4]), q1[5]), q1[0]), q1[1]), q1[2]);
    } else {
      c0 = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(q0[0], q0[1]), q0[2]), q0[3]), q0[4]), q0[5]);
      c1 = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(q1[0], q1[1]), q1[2]), q1[3]), q1[4]), q1[5]);
    }
    coef[a * 5 + p] = make_double2(c0, c1);
    // e_k = k*c_k for the velocity polynomial (2*coeff(s,2), 3*coeff(s,3), ... :477-485)
    if (p >= 1) coef[a * 5 + 2 + p] = make_double2(__dmul_rn((double)(2 * p), c0), __dmul_rn((double)(2 * p + 1), c1));
  }
}

// shared-memory flags between the warps of a CTA (PTX memory model, CTA scope)
__device__ __forceinline__ void st_release_cta(unsigned* p, unsigned v) {
  asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_cta(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ unsigned atom_add_acq_rel_cta(unsigned* p, unsigned v) {
  unsigned old;
  asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(smem_u32(p)), "r"(v) : "memory");
  return old;
}

// kNoAlloc: L1 policy of the gathers (see ldg_quad), chosen from the sampling density.  kOneRound: as many warps as
// pairs (no loop around the pair body: the compiler then keeps the sample loop's ~40 map / penalty constants in uniform
// registers instead of re-reading them from the constant bank in every iteration).  kPlain: the map keeps no quad copy.
// GTOP_POLY_HORNER (compile-time experiment, not built by default): position / velocity polynomials by Horner's rule
// with FMA — 27 FP64 instructions per sample less and one 128-bit table load instead of three (0.317 ms against
// 0.335 ms per 65,536, profiles/r2_microbench.md) — but the float image of a position is then no longer the
// reference's in ~13 of 20 million samples, and such a flip moves that sample's penalty by 1e-6 relative.
#ifdef GTOP_POLY_HORNER
constexpr bool kHorner = true;
#else
constexpr bool kHorner = false;
#endif
template <bool kNoAlloc, bool kOneRound, bool kPlain>
__global__ void __launch_bounds__(kPairMaxWarps * 32, 2)
k_eval_mma(SdfView S, ProbLayout PL, PairSmem sm, const double* __restrict__ prob, unsigned num_problems,
                 EvalScalars E, const double* __restrict__ X, int64_t B, int head, unsigned group0, unsigned gpp,
                 int bulk_ok, double* __restrict__ cost_out, double* __restrict__ grad_out,
                 unsigned long long* __restrict__ oob_counter) {
  extern __shared__ __align__(16) double smem[];
  const int m = PL.m, P = PL.P, n = PL.n, npair = PL.npair;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, W = nthr >> 5;
  const int g = lane >> 2, j = lane & 3;  // trajectory slot of the warp, sample slot of the iteration
  double* xtab = smem + sm.xtab;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + sm.mbar);
  unsigned* flags = reinterpret_cast<unsigned*>(smem + sm.flag);
  const unsigned xbytes = (unsigned)(8 * n) * 8u;

  // Groups of 8 are aligned on the GLOBAL trajectory index (traj_offset + b): any split of a batch evaluates every
  // trajectory in the same lane group with the same instructions.  Group grp of this launch holds the batch's
  // trajectories 8 grp - head .. 8 grp - head + 7 (head = traj_offset & 7: a ragged first group).
  const int64_t ngroups = (head + B + 7) >> 3;
  // all 8 trajectories of a group belong to one problem (host-checked: starts_per_problem % 8 == 0); gpp groups per
  // problem, group0 = the launch's first global group index modulo the period gpp * num_problems (32-bit arithmetic)
  auto problem_of = [&](int64_t grp) { return prob + (((group0 + (unsigned)grp) / gpp) % num_problems) * (int64_t)PL.stride; };
  const int64_t grp = blockIdx.x;
  const int64_t b0 = grp * 8 - head;  // the group's first trajectory inside the batch
  // A group whose 8 trajectories all lie inside the batch moves by bulk copies (bulk_ok: the host checked the 16-byte
  // alignment of X, grad and of the groups inside them); ragged head / tail groups take plain loads and stores (dead
  // slots shadow a live trajectory of the batch)
  const bool whole = bulk_ok && b0 >= 0 && b0 + 8 <= B;
  const double* __restrict__ pr = prob + (((group0 + (unsigned)grp) / gpp) % num_problems) * (int64_t)PL.stride;
  double* xin = smem + sm.xin;
  double* outb = smem + sm.out;

  // ---- set-up that reads nothing a preceding kernel may have written -----------------------------------------------
  if (tid == 0) {
    mbar_init(mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  exp2_tab_init(xtab, tid, nthr);
  for (int i = tid; i <= npair; i += nthr) flags[i] = 0u;
  __syncthreads();  // the only block-wide barrier of the kernel (whole groups)
  // programmatic dependent launch: the next kernel of the stream may start to take free SM slots (and run the set-up
  // above) now; X and the problem records may come from the preceding kernel and are only read after the wait
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (whole) {
    if (tid == 0) {
      mbar_expect_tx(mbar, xbytes + 144u);
      bulk_load_stream(xin, X + b0 * n, xbytes, mbar);
      bulk_load(xin + 8 * n, pr + PL.off_Df, 144u, mbar);
    }
  } else {
    for (int i = tid; i < 8 * n; i += nthr) {
      const int gg = i / n;
      const int64_t bb = b0 + gg;
      xin[i] = X[(bb < 0 ? 0 : (bb < B ? bb : B - 1)) * n + (i - gg * n)];
    }
    if (tid < 18) xin[8 * n + tid] = pr[PL.off_Df + tid];
    __syncthreads();
  }
  // One pair per warp: what the pair reads from the problem record before its first sample — the sample counts and the
  // Linv pairs of its two segments — is requested NOW, under the copy of x.  Taken one after the other behind x these were
  // two more L2 round trips in the life of a CTA, and a small launch is made of those: one CTA per SM with ONE sample
  // iteration per pair took 7.7 us (tools/kbench.py --ns 2; 7.3 with this), and the full batch gains 0.7 % (0.3317 ->
  // 0.3294 ms).  Hoisting the first time record as well made the prologue shorter still and the sample loop slower
  // (a different schedule); an L1 prefetch of the first records and of the epilogue's band rows changed nothing.
  double2 preLa[6], preLb[6];
  int pre_nsa = 0, pre_nsb = 0;
  if (kOneRound && !E.skip_collision) {
    const int sa = 2 * warp, sb = sa + 1;
    pre_nsa = (int)pr[PL.off_ns + sa];
    pre_nsb = sb < m ? (int)pr[PL.off_ns + sb] : 0;
    if (j < 3) {
      const double2* La = reinterpret_cast<const double2*>(pr + PL.off_LiTP + sa * 36) + j * 6;
      const double2* Lb = reinterpret_cast<const double2*>(pr + PL.off_LiTP + (sb < m ? sb : sa) * 36) + j * 6;
#pragma unroll
      for (int r = 0; r < 6; ++r) preLa[r] = ldg_tab(La + r), preLb[r] = ldg_tab(Lb + r);
    }
  }
  if (whole) mbar_wait(mbar, 0u);

  const double kA = -E.alpha * E.alpha_over_r * S.res_inv;  // gd*cd = -(alpha^2/r) e^2; grad scale 1/res
  double2* coef = reinterpret_cast<double2*>(smem + sm.coef + warp * 480) + g * 15;  // this trajectory's 15 pairs
  double2* coefB = coef + 8 * 15;  // coefficients of the pair's second segment (same layout, next 240 doubles)
  double* scr = smem + sm.scr + warp * 144 + g * 9;
  unsigned* fl = flags;
  {
    const double* xg = xin + g * n;
    const double* dfs = xin + 8 * n;
    // boundary derivatives of waypoint w (0..m+1; m+1 is never read): pointer and axis stride, no branches
    auto waypoint = [&](int w) {
      Wp r;
      const bool fixed = w == 0 || w >= m;
      r.p = fixed ? (w == 0 ? dfs : dfs + 3) : xg + 3 * (w - 1);
      r.s = fixed ? 6 : P;
      return r;
    };

    // ---- pairs of this warp, highest first (pair p waits for a part of pair p+1, never the other way) ----------------
    unsigned oob = 0;
    auto do_pair = [&](const int p) {
      const int sa = 2 * p, sb = sa + 1;
      const bool hasb = sb < m;
      const Wp wp0 = waypoint(sa), wp1 = waypoint(sa + 1), wp2 = waypoint(sa + 2), wp3 = waypoint(sa + 3);
      double* exch = smem + sm.exch + p * 72 + g * 9;
      double Jc = 0.0;
      if (!E.skip_collision) {
        {
          const double2* LiTP = reinterpret_cast<const double2*>(pr + PL.off_LiTP + sa * 36);
          if (sa == m - 1) pair_coef_phase<true>(wp0, wp1, LiTP, coef, j, kOneRound ? preLa : nullptr);
          else pair_coef_phase<false>(wp0, wp1, LiTP, coef, j, kOneRound ? preLa : nullptr);
        }
        if (hasb) {
          const double2* LiTP = reinterpret_cast<const double2*>(pr + PL.off_LiTP + sb * 36);
          if (sb == m - 1) pair_coef_phase<true>(wp1, wp2, LiTP, coefB, j, kOneRound ? preLb : nullptr);
          else pair_coef_phase<false>(wp1, wp2, LiTP, coefB, j, kOneRound ? preLb : nullptr);
        }
        __syncwarp();

        double G[3][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};  // gradient rows in flight (C fragments)
        // the columns that carry the START derivatives of a segment of parity w leave the C fragments for `dst`
        auto flush_rows = [&](int w, double* dst) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int r = deriv_of_col(2 * j + e, w);
            if (r >= 0 && r < 3) {
#pragma unroll
              for (int a = 0; a < 3; ++a) {
                dst[a * 3 + r] = G[a][e];
                G[a][e] = 0.0;
              }
            }
          }
        };
        // the rows of waypoint 2p from this pair's first segment go to pair p-1 (which owns that waypoint)
        auto publish_start_rows = [&]() {
          flush_rows(sa, exch);
          __syncwarp();
          if (lane == 0) st_release_cta(fl + p, 1u);
        };

        const int nsa = kOneRound ? pre_nsa : (int)pr[PL.off_ns + sa], nsb = kOneRound ? pre_nsb : (hasb ? (int)pr[PL.off_ns + sb] : 0);
        const int nit = (nsa + nsb + 3) >> 2;
        const bool mixed = nsb > 0 && (nsa & 3) != 0;  // one iteration holds samples of both segments
        // this lane samples the second segment from iteration ubl on (the mixed iteration splits the lanes)
        const int ubl = nsb > 0 ? (nsa >> 2) + ((mixed && j < (nsa & 3)) ? 1 : 0) : 0x7fffffff;
        // the first segment's last sample goes in at iteration uf (its rows are flushed after that iteration's DMMAs)
        const int uf = nsb > 0 ? (mixed ? (nsa >> 2) : (nsa >> 2) - 1) : -1;
        bool flushed = false;
        if (nsb > 0 && uf < 0) publish_start_rows(), flushed = true;  // (degenerate: the first segment has no samples)
        // lane-ordered tables of the pair (gtop_types.cuh): the 4 sample slots of an iteration read 64 contiguous bytes
        // per load of the time record (1 L1 wavefront), the 32 lanes 512 contiguous bytes of B fragments (4 wavefronts)
        const double2* __restrict__ tq = reinterpret_cast<const double2*>(pr + PL.off_TQ) + (size_t)p * PL.nit2 * 12 + j;
        const double2* __restrict__ pl = reinterpret_cast<const double2*>(pr + PL.off_PL) + (size_t)p * (PL.nit2 + 1) * 32 + lane;
        double2 ta, tb = make_double2(0.0, 0.0), tc = tb;  // (dt,t) (t2,t3) (t4,t5) of the current sample, requested one iteration ahead
        if (nit > 0) {
          ta = ldg_tab(tq);
          if (!kHorner) tb = ldg_tab(tq + 4), tc = ldg_tab(tq + 8);
        }
        const double2* cf = 0 >= ubl ? coefB : coef;  // this lane's coefficient set in the current iteration
        for (int it = 0; it < nit; ++it, tq += 12, pl += 32) {
          const double t = ta.y;
          double t2 = 0.0, t3 = 0.0, t4 = 0.0, t5 = 0.0;
          if (!kHorner) t2 = tb.x, t3 = tb.y, t4 = tc.x, t5 = tc.y;
          const double wdt = ta.x;  // dt of the slot's segment; 0 for idle slots, which evaluate a segment start in vain
          double pos[3], vel[3], c1[3];
          float pf[3];
#pragma unroll
          for (int a = 0; a < 3; ++a) {
            const double2 c01 = cf[a * 5 + 0], c23 = cf[a * 5 + 1], c45 = cf[a * 5 + 2];
            c1[a] = c01.y;
            double pa;
            if (kHorner) {  // experiment: 5 DFMA instead of 5 DMUL + 5 DADD; differs from the reference by rounding only
              pa = fma(fma(fma(fma(fma(c45.y, t, c45.x), t, c23.y), t, c23.x), t, c01.y), t, c01.x);
            } else {
              // getPositionFromCoeff (:457-465): left-to-right, unfused, then float
              pa = __dadd_rn(c01.x, __dmul_rn(c01.y, t));
              pa = __dadd_rn(pa, __dmul_rn(c23.x, t2));
              pa = __dadd_rn(pa, __dmul_rn(c23.y, t3));
              pa = __dadd_rn(pa, __dmul_rn(c45.x, t4));
              pa = __dadd_rn(pa, __dmul_rn(c45.y, t5));
            }
            pf[a] = __double2float_rn(pa);
            pos[a] = (double)pf[a];
          }
          SdfCell cell;
          sdf_gather<kNoAlloc, kPlain>(S, pos[0], pos[1], pos[2], pf[0], pf[1], pf[2], cell);  // 2 x LDG.256 in flight from here
#pragma unroll
          for (int a = 0; a < 3; ++a) {
            const double2 e01 = cf[a * 5 + 3], e23 = cf[a * 5 + 4];
            double va;
            if (kHorner) {
              va = fma(fma(fma(fma(e23.y, t, e23.x), t, e01.y), t, e01.x), t, c1[a]);
            } else {
              // getVelocityFromCoeff (:477-485)
              va = __dadd_rn(c1[a], __dmul_rn(e01.x, t));
              va = __dadd_rn(va, __dmul_rn(e01.y, t2));
              va = __dadd_rn(va, __dmul_rn(e23.x, t3));
              va = __dadd_rn(va, __dmul_rn(e23.y, t4));
            }
            vel[a] = f32_round_trip(va);
          }
          // B fragments of this sample (column g of the lane), then the time record of the next sample
          const double2 bAB = ldg_tab(pl);
          if (it + 1 < nit) {
            ta = ldg_tab(tq + 12);
            if (!kHorner) tb = ldg_tab(tq + 16), tc = ldg_tab(tq + 20);
          }

          const double vn2 = fma(vel[2], vel[2], fma(vel[1], vel[1], fma(vel[0], vel[0], 1e-300)));
          const double vnorm = sqrt_pos(vn2) + 1e-5;  // :358 (the 1e-300 keeps v = 0 finite; absorbed by 1e-5)
          const double inv_vn = rcp_pos(vnorm);

          double dist, gd[3];
          // (the sample loop is warp-uniform; out-of-map samples of real — non-idle — slots are counted)
          sdf_finish<true>(cell, dist, gd[0], gd[1], gd[2], &oob, __double2hiint(wdt) != 0);

          // exp(-(d-d0)/r) (:509/:514) = 2^((d0-d)*log2(e)/r)
          const double ee = exp2_tab(fma(dist, E.neg_log2e_over_r, E.d0_log2e_over_r), xtab);
          const double ev = ee * vnorm;
          Jc = fma(ev, wdt, Jc);  // cost: cd*|v|*dt = alpha*ee*vnorm*dt (:374; alpha applied at the end)
          // sA_a = gd*grad_a*cd*|v| = kA * ee^2 * vnorm * g_a      (:378, extra cd reproduced — H4)
          // sB_a = cd*v_a/|v|       = alpha * ee * v_a / vnorm     (:379)
          const double sAc = kA * (ee * ev);
          const double sBc = E.alpha * (ee * inv_vn);
          double fa[3], fb[3];
#pragma unroll
          for (int a = 0; a < 3; ++a) fa[a] = sAc * gd[a], fb[a] = sBc * vel[a];
#pragma unroll
          for (int a = 0; a < 3; ++a) {
            dmma884(G[a][0], G[a][1], fa[a], bAB.x);  // += sA_a(sample) * dt phi_r(sample), straight into G
            dmma884(G[a][0], G[a][1], fb[a], bAB.y);
          }
          cf = it + 1 >= ubl ? coefB : coef;
          if (__builtin_expect(it == uf, 0)) {
            // the first segment is complete: the rows of its start waypoint leave the C fragments.  In a mixed
            // iteration record `it` carried the first segment's lanes only; the second segment's lanes go in now
            // through the pair's extra record
            publish_start_rows(), flushed = true;
            if (mixed) {
              const double2 bX = ldg_tab(pl + (size_t)(PL.nit2 - it) * 32);
#pragma unroll
              for (int a = 0; a < 3; ++a) {
                dmma884(G[a][0], G[a][1], fa[a], bX.x);
                dmma884(G[a][0], G[a][1], fb[a], bX.y);
              }
            }
          }
        }
        if (!flushed) publish_start_rows();
        flush_rows(sb, scr);           // waypoint 2p+1: end derivatives of segment 2p (+ start derivatives of 2p+1, in place)
        flush_rows(sb + 1, scr + 72);  // waypoint 2p+2: end derivatives of segment 2p+1
        __syncwarp();
      }

      // ---- rows the pair owns: waypoints 2p+1 and 2p+2 (and the fixed start / end state's share of d'Rd) -------------
      double js = 0.0;
      if (j < 3) {
        const int r = j;
        // rows of a fixed boundary state: column band of R (:326-327); entries 3..8 <-> waypoints 0, 1 for w = 0,
        // entries 0..5 <-> waypoints m-1, m for w = m
        auto fixed_rows = [&](const double* __restrict__ band, const Wp w0, const Wp w1, const Wp ww) {
#pragma unroll
          for (int a = 0; a < 3; ++a) {
            double y = 0.0;
#pragma unroll
            for (int u = 0; u < 3; ++u) y = fma(w0.p[a * w0.s + u], band[u], y);
#pragma unroll
            for (int u = 0; u < 3; ++u) y = fma(w1.p[a * w1.s + u], band[3 + u], y);
            js = fma(y, ww.p[a * ww.s + r], js);
          }
        };
        // smoothness rows 2 Rfp' df + 2 Rpp dp (:330-332) of interior waypoint w from the band of R; the same row
        // products give the rows' share of d' R d (:326-327)
        auto smooth_rows = [&](int w, const Wp wm, const Wp w0, const Wp w1, double (&gs)[3]) {
          const double2* __restrict__ band = reinterpret_cast<const double2*>(pr + PL.off_RgP + (3 * (w - 1) + r) * 10);
          double bnd[10];
#pragma unroll
          for (int u = 0; u < 5; ++u) {
            const double2 v = ldg_tab(band + u);
            bnd[2 * u] = v.x, bnd[2 * u + 1] = v.y;
          }
#pragma unroll
          for (int a = 0; a < 3; ++a) {
            double s = 0.0;
#pragma unroll
            for (int u = 0; u < 3; ++u) s = fma(wm.p[a * wm.s + u], bnd[u], s);
#pragma unroll
            for (int u = 0; u < 3; ++u) s = fma(w0.p[a * w0.s + u], bnd[3 + u], s);
#pragma unroll
            for (int u = 0; u < 3; ++u) s = fma(w1.p[a * w1.s + u], bnd[6 + u], s);
            js = fma(s, w0.p[a * w0.s + r], js);
            gs[a] = s;
          }
        };
        const int w1 = sa + 1, w2 = sa + 2;
        if (p == 0) fixed_rows(pr + PL.off_Rc + r * 9 + 3, wp0, wp1, wp0);
        if (w1 < m) {
          double gs[3];
          smooth_rows(w1, wp0, wp1, wp2, gs);
#pragma unroll
          for (int a = 0; a < 3; ++a)
            outb[g * n + a * P + 3 * (w1 - 1) + r] = (E.ws * (2.0 * gs[a]) + E.wc * (E.skip_collision ? 0.0 : scr[a * 3 + r])) + 1e-5;
        } else {  // w1 == m: the pair is the odd last segment
          fixed_rows(pr + PL.off_Rc + (3 * m + r) * 9, wp0, wp1, wp1);
        }
        if (w2 < m) {
          double gs[3];
          smooth_rows(w2, wp1, wp2, wp3, gs);
          // the START-derivative part of these rows comes from pair p+1 (published half a pair ago, as a rule)
          const double* ex = exch + 72 + r;
          if (!E.skip_collision)
            while (ld_acquire_cta(fl + p + 1) == 0u) {
            }
#pragma unroll
          for (int a = 0; a < 3; ++a) {
            const double cr = E.skip_collision ? 0.0 : scr[72 + a * 3 + r] + ex[a * 3];
            outb[g * n + a * P + 3 * (w2 - 1) + r] = (E.ws * (2.0 * gs[a]) + E.wc * cr) + 1e-5;
          }
        } else if (w2 == m) {
          fixed_rows(pr + PL.off_Rc + (3 * m + r) * 9, wp1, wp2, wp2);
        }
      }
      js += __shfl_xor_sync(kFull, js, 1);
      js += __shfl_xor_sync(kFull, js, 2);
      Jc += __shfl_xor_sync(kFull, Jc, 1);
      Jc += __shfl_xor_sync(kFull, Jc, 2);
      if (j == 0) smem[sm.jsp + g * npair + p] = js, smem[sm.jcp + g * npair + p] = Jc;
      __syncwarp();  // scr and the coefficient buffers are rewritten by the warp's next pair
    };
    if (kOneRound) {
      do_pair(warp);
    } else {
      int p_top = warp;  // this warp's highest pair
      while (p_top + W < npair) p_top += W;
      for (int p = p_top; p >= 0; p -= W) do_pair(p);
    }
    const int64_t br = b0 + g;  // this lane's trajectory inside the batch (ragged groups: may be outside)
    const bool live = br >= 0 && br < B;
    if (oob_counter != nullptr) {
      oob = __reduce_add_sync(kFull, live ? oob : 0u);
      if (lane == 0 && oob) atomicAdd(oob_counter, (unsigned long long)oob);
    }

    // ---- the warp is done with the group; the LAST warp of the CTA to get here closes it -------------------------------
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the rows written above -> visible to the bulk store
    __syncwarp();
    unsigned last = 0;
    if (lane == 0) last = atom_add_acq_rel_cta(fl + npair, 1u) + 1u == (unsigned)W;
    last = __shfl_sync(kFull, last, 0);
    if (last) {
      // cost (:417-418): pair partials in pair order
      double js = 0.0, Jc = 0.0;
      for (int p = j; p < npair; p += 4)
        js += smem[sm.jsp + g * npair + p], Jc += smem[sm.jcp + g * npair + p];
#pragma unroll
      for (int o = 2; o > 0; o >>= 1) {
        js += __shfl_xor_sync(kFull, js, o);
        Jc += __shfl_xor_sync(kFull, Jc, o);
      }
      if (j == 0 && live) cost_out[br] = E.ws * js + E.wc * (E.alpha * Jc) + 1e-3;
      // out: one bulk store of the group's gradient rows (ragged groups: guarded plain stores)
      if (whole) {
        if (lane == 0) {
          bulk_store(grad_out + b0 * n, outb, xbytes);
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // shared memory must outlive the reads
        }
      } else {
        for (int i = lane; i < 8 * n; i += 32) {
          const int gg = i / n;
          const int64_t bb = b0 + gg;
          if (bb >= 0 && bb < B) grad_out[bb * n + (i - gg * n)] = outb[i];
        }
      }
    }
  }
}

// warps per CTA: the candidate that keeps the most warps of an SM busy — resident CTAs (20 warps of 96 registers, 227 KB
// of shared memory per SM) x pairs per CTA / rounds a CTA needs for its pairs; ties go to the fewest rounds
static int pair_warps(int m) {
  static const int force = [] { const char* e = std::getenv("GTOP_PAIR_WARPS"); return e ? std::atoi(e) : 0; }();  // tuning knob
  const int npair = (m + 1) / 2;
  if (force > 0) return std::min(std::min(force, kPairMaxWarps), npair);
  int best = 1;
  double best_busy = -1.0;
  for (int W = 1; W <= kPairMaxWarps && W <= npair; ++W) {
    const int rounds = (npair + W - 1) / W;
    const size_t bytes = sizeof(double) * (size_t)pair_smem(m, W).total + 1024;
    const int ctas = (int)std::min<size_t>(20 / W, (227 * 1024) / bytes);
    const double busy = (double)ctas * npair / rounds - 1e-3 * rounds;
    if (busy > best_busy) best_busy = busy, best = W;
  }
  return best;
}

bool eval_mma_applicable(const ProbLayout& PL, int64_t B, int num_problems, int64_t starts_per_problem) {
  if (starts_per_problem % 8 != 0 || PL.m < 2) return false;
  // 32-bit group arithmetic in the kernel
  if ((starts_per_problem / 8) * (int64_t)num_problems + B / 8 + 2 >= (int64_t)1 << 31) return false;
  const PairSmem sm = pair_smem(PL.m, pair_warps(PL.m));
  return (size_t)sm.total * sizeof(double) <= 100 * 1024;  // two CTAs per SM at least
}

cudaError_t launch_eval_mma(const SdfView& S, const ProbLayout& PL, const double* prob, int num_problems,
                                  const EvalScalars& E, const double* X, int64_t B, int64_t traj_offset,
                                  int64_t starts_per_problem, double* cost, double* grad,
                                  unsigned long long* oob_counter, cudaStream_t stream) {
  if (B <= 0) return cudaSuccess;
  const int W = pair_warps(PL.m);
  const PairSmem sm = pair_smem(PL.m, W);
  const size_t smem = sizeof(double) * (size_t)sm.total;
  const bool one = W == PL.npair;
  const bool na = PL.ns_cap <= 64;
  auto kern = k_eval_mma<false, false, false>;
  if (S.quad == nullptr) kern = na ? (one ? k_eval_mma<true, true, true> : k_eval_mma<true, false, true>)
                                   : (one ? k_eval_mma<false, true, true> : k_eval_mma<false, false, true>);
  else kern = na ? (one ? k_eval_mma<true, true, false> : k_eval_mma<true, false, false>)
                 : (one ? k_eval_mma<false, true, false> : k_eval_mma<false, false, false>);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  // bulk copies need 16-byte aligned addresses: a group's block starts (8 k - (traj_offset & 7)) * n doubles into X / grad
  const int bulk_ok = ((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(grad)) & 15u) == 0 &&
                      (((traj_offset & 7) * PL.n) & 1) == 0 && (PL.off_Df & 1) == 0 && (PL.stride & 1) == 0 &&
                      (reinterpret_cast<uintptr_t>(prob) & 15u) == 0;
  const int64_t ngroups = ((traj_offset & 7) + B + 7) / 8;
  // problem of a group in 32-bit arithmetic: the launch's first global group index modulo the period of the mapping
  const uint64_t gpp = (uint64_t)(starts_per_problem / 8), period = gpp * (uint64_t)num_problems;
  const unsigned group0 = (unsigned)((uint64_t)(traj_offset >> 3) % period);
  static const bool pdl = [] { const char* e = std::getenv("GTOP_PDL"); return !e || std::atoi(e) != 0; }();  // tuning knob
  static const bool nobulk = [] { const char* e = std::getenv("GTOP_NO_BULK"); return e && std::atoi(e) != 0; }();  // test knob
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)ngroups);
  cfg.blockDim = dim3(W * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, S, PL, sm, prob, (unsigned)num_problems, E, X, B, (int)(traj_offset & 7), group0,
                            (unsigned)gpp, nobulk ? 0 : bulk_ok, cost, grad, oob_counter);
}

}  // namespace gtop
