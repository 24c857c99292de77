/* gtop_c.h — C ABI of the B200-native GTOP inner loop (libgtop_b200.so).
 *
 * This is the drop-in boundary: host C++ (the reference's grad_traj_optimizer.cpp, or this repo's
 * ROS-free GradTrajOptimizer mirror in include/gtop/) reaches the sm_100a CUDA kernels ONLY through
 * these entry points.  Plain pointers and sizes, no C++/torch types, no exceptions across the ABI.
 *
 * Each entry point names the reference interface it replaces (paths relative to the upstream repo
 * EpicOne1/grad_traj_optimization):
 *
 *   gtop_sdf_create            SDFMap::SDFMap(origin, resolution, map_size)        src/sdf_map.cpp:3-24
 *                              GradTrajOptimizer::initSDFMap                        src/grad_traj_optimizer.cpp:112-115
 *   gtop_sdf_reset             SDFMap::resetBuffer()                                src/sdf_map.cpp:26-53
 *   gtop_sdf_set_occupancy     SDFMap::setOccupancy(pos, 1) per point               src/sdf_map.cpp:80-99 (+isInMap :55-69, posToIndex :71-74)
 *   gtop_sdf_update_esdf       SDFMap::updateESDF3d()                               src/sdf_map.cpp:310-368 (+fillESDF :266-308)
 *                              (reset + set_occupancy + update_esdf == GradTrajOptimizer::updateSDFMap, grad_traj_optimizer.cpp:117-126)
 *   gtop_sdf_query             SDFMap::getDistWithGradTrilinear per point           src/sdf_map.cpp:185-242
 *   gtop_problems_create_dense the members L, R, Df, segment_time left by setPath   src/grad_traj_optimizer.cpp:67-110
 *                              (TrajectoryGenerator::StackOptiDep/getInitialD       src/qp_generator.cpp:357-451)
 *   gtop_problems_create_from_paths  setPath itself (time allocation :73-81 + PolyQPGeneration/StackOptiDep/getInitialD)
 *   gtop_problems_create_from_kino   setKinoPath                                   src/grad_traj_optimizer.cpp:35-65 (PolyKinoGeneration src/qp_generator.cpp:23-154)
 *   gtop_eval_batch_device / _host   GradTrajOptimizer::getCostAndGradient          src/grad_traj_optimizer.cpp:281-448, batched
 *   gtop_objective             GradTrajOptimizer::costFunc (NLopt callback)         src/grad_traj_optimizer.cpp:554-562,
 *                              registered by set_min_objective(costFunc, this)      src/grad_traj_optimizer.cpp:140
 *   gtop_coeff_from_x          GradTrajOptimizer::getCoefficientFromDerivative      src/grad_traj_optimizer.cpp:253-279
 *   gtop_sdf_reset_range       SDFMap::resetBuffer(min_pos, max_pos)                src/sdf_map.cpp:28-53
 *   gtop_sdf_set_update_range  SDFMap::setUpdateRange(min_pos, max_pos)             src/sdf_map.cpp:244-264
 *   gtop_sdf_query_dynamic     EDTEnvironment::evaluateEDTWithGrad / evaluateCoarseEDT   src/edt_environment.cpp:75-137
 *   gtop_problems_create_from_paths_device / _from_kino_device   setPath / setKinoPath for a batch, on the device
 *   gtop_solve_batch_device / _host, gtop_mma_minimize   the nlopt::opt(LD_MMA) call of optimizeTrajectory   src/grad_traj_optimizer.cpp:135-195
 *   gtop_traj_metrics_device / _host   PolynomialTraj evaluation of the result      src/opti_node.cpp:104-144,
 *                                                                                   include/grad_traj_optimization/polynomial_traj.hpp:70-203
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error (GTOP_E_*); gtop_last_error() gives the
 *     thread-local message.  gtop_objective returns HUGE_VAL after recording an error.
 *   - handles are opaque, own their device memory, and are usable from one host thread at a time;
 *     distinct handles / streams may be used concurrently.
 *   - x / grad layout is the reference's: n = 3*P doubles, P = 3*(m-1), [x-axis P | y-axis P | z-axis P],
 *     within an axis [pos, vel, acc] per interior waypoint (grad_traj_optimizer.cpp:309-313).
 *   - dense matrices are COLUMN-major (Eigen::MatrixXd default), Df is 3x6 column-major too unless
 *     stated; the distance / occupancy buffers use the reference index x*ny*nz + y*nz + z.
 *   - there is no CPU fallback: every entry point that computes needs a CUDA device and fails
 *     with GTOP_E_CUDA otherwise.
 */
#ifndef GTOP_C_H_
#define GTOP_C_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GTOP_ABI_VERSION 1

#define GTOP_OK 0
#define GTOP_E_INVALID (-1) /* bad argument */
#define GTOP_E_CUDA (-2)    /* CUDA runtime error / no device */
#define GTOP_E_ALLOC (-3)
#define GTOP_E_STATE (-4) /* call order / handle state */

#define GTOP_OPT_INITIAL_TRY 0 /* grad_traj_optimizer.h:14-16 */
#define GTOP_OPT_FIRST_STEP 1
#define GTOP_OPT_SECOND_STEP 2

typedef struct gtop_sdf gtop_sdf;                     /* SDFMap on one GPU */
typedef struct gtop_problems gtop_problems;           /* a set of trajectory problems (same m) on one GPU */
typedef struct gtop_objective_ctx gtop_objective_ctx; /* `this` for the NLopt callback */

/* The ROS parameters of grad_traj_optimizer.cpp:5-32 as a plain struct (doubles first, then ints). */
typedef struct gtop_params {
  double time_limit_1, time_limit_2, deltat;
  double w_smooth, w_collision; /* ws, wc */
  double alpha, r, d0;
  double alpha_v, r_v, v0;
  double alpha_a, r_a, a0;
  double bos, vos, aos;
  double mean_v, mean_a, init_time;
  int32_t algorithm;           /* NLopt algorithm id (24 = LD_MMA) */
  int32_t samples_per_segment; /* hard-coded 30 upstream (grad_traj_optimizer.cpp:351) */
  int32_t vel_acc_penalty;     /* 0 = HEAD behaviour (block :383-407 commented out) */
  int32_t reserved;
} gtop_params;

int gtop_abi_version(void);
const char* gtop_last_error(void);
int gtop_device_count(void);
void gtop_params_default(gtop_params* p);

/* pinned host memory for the *_host entry points (optional; any host pointer works, pinned is faster) */
int gtop_host_alloc(void** ptr, size_t bytes);
int gtop_host_free(void* ptr);

/* ---------------------------------------------------------------- SDF map ---- */
int gtop_sdf_create(const double origin[3], double resolution, const double map_size[3], int device, gtop_sdf** out);
int gtop_sdf_destroy(gtop_sdf* sdf);
int gtop_sdf_grid(const gtop_sdf* sdf, int32_t grid[3]);
int gtop_sdf_reset(gtop_sdf* sdf);
/* SDFMap::resetBuffer(min_pos, max_pos) (src/sdf_map.cpp:28-53): occupancy 0 and distance 10000 inside the metric box
 * (clamped to the map; voxels posToIndex(min_pos) .. posToIndex(max_pos - res/2)).  GTOP_E_INVALID for an empty box or a box
 * outside the map (undefined behaviour in the reference). */
int gtop_sdf_reset_range(gtop_sdf* sdf, const double min_pos[3], const double max_pos[3]);
/* SDFMap::setUpdateRange (src/sdf_map.cpp:244-264): the NEXT gtop_sdf_update_esdf only runs over this box and only
 * sees its occupied voxels; afterwards the window is the whole grid again (:366-367). */
int gtop_sdf_set_update_range(gtop_sdf* sdf, const double min_pos[3], const double max_pos[3]);
/* current update window (SDFMap::min_vec, max_vec): lo_hi = {x0,y0,z0, x1,y1,z1}, inclusive voxel indices */
int gtop_sdf_update_range(const gtop_sdf* sdf, int32_t lo_hi[6]);
int gtop_sdf_set_occupancy(gtop_sdf* sdf, const double* xyz /* n x 3, host */, int64_t n);
/* extension for large synthetic maps: mark whole voxel boxes [lo, hi] (inclusive, clipped), n x 6 int32 host */
int gtop_sdf_set_occupancy_boxes(gtop_sdf* sdf, const int32_t* lo_hi, int64_t n);
int gtop_sdf_update_esdf(gtop_sdf* sdf);
/* device time of the last update_esdf: ms[0..2] = z/y/x pass kernels, ms[3] = total */
int gtop_sdf_last_edt_ms(const gtop_sdf* sdf, float ms[4]);
/* the same, through to "ready for the evaluation kernels": ms[0..2] = z/y/x passes, ms[3] = their sum, ms[4] = rebuild of
 * the gather copy (0 for maps that keep none), ms[5] = start of the update to evaluator-ready */
int gtop_sdf_last_update_ms(const gtop_sdf* sdf, float ms[6]);
/* how the evaluation kernels read the field: has_quad_copy = 1 -> from the xy-quad copy (4x the field, two 256-bit loads
 * per sample; maps whose copy stays under GTOP_QUAD_LIMIT_MB, default 1024), 0 -> straight from the distance field, which
 * such maps keep in 8 x 8 x 4 voxel bricks instead of the reference's linear order (big maps: eight 64-bit loads per
 * sample, no copy to keep or rebuild; up / download convert).  resident_bytes = device memory the map holds. */
int gtop_sdf_gather_layout(const gtop_sdf* sdf, int32_t* has_quad_copy, int64_t* resident_bytes);
int gtop_sdf_download_distance(const gtop_sdf* sdf, double* host);
int gtop_sdf_upload_distance(gtop_sdf* sdf, const double* host);
int gtop_sdf_download_occupancy(const gtop_sdf* sdf, int32_t* host);
int gtop_sdf_upload_occupancy(gtop_sdf* sdf, const int32_t* host);
/* device pointer of the fp64 distance buffer (nx*ny*nz doubles, x*ny*nz + y*nz + z), for callers that share it;
 * GTOP_E_INVALID for a map without a quad copy (its field is bricked, see gtop_sdf_gather_layout) */
int gtop_sdf_distance_device_ptr(const gtop_sdf* sdf, const double** dev_ptr);
/* batched getDistWithGradTrilinear: host in/out; out-of-map -> dist=-1, grad=0; returns
 * the out-of-map count through *n_out_of_map (nullable) */
int gtop_sdf_query(const gtop_sdf* sdf, const double* pos /* n x 3 */, int64_t n, double* dist, double* grad /* n x 3 */,
                   int64_t* n_out_of_map);
/* EDTEnvironment::evaluateEDTWithGrad / evaluateCoarseEDT (src/edt_environment.cpp:75-137) for n points: the static
 * field min'ed at the eight corner voxels of the cell with the distance to nbox constant-velocity boxes at `time`
 * (time < 0: static field only).  boxes: nbox x 9 = centre at t = 0, velocity, full edge lengths
 * (PolynomialPrediction::evaluateConstVel + ObjScale, obj_predictor.h:57-66).  Corner voxels outside the map contribute
 * -1 (getDistance(Vector3d), sdf_map.cpp:155-164), they are not clamped.  coarse may be NULL.  Host buffers. */
int gtop_sdf_query_dynamic(const gtop_sdf* sdf, const double* boxes, int32_t nbox, double time, const double* pos /* n x 3 */,
                           int64_t n, double* dist, double* grad /* n x 3 */, double* coarse);

/* ---------------------------------------------------------------- problems ---- */
/* Drop-in path: take the matrices the reference's setPath leaves in the optimizer.
 *   seg_time : num_problems x m
 *   L        : num_problems x [6m x (3m+3)] column-major     (getL, = A^-1 C^T)
 *   R        : num_problems x [(3m+3) x (3m+3)] column-major (getR)
 *   Df       : num_problems x [3 x 6] column-major           (getInitialD().second)
 */
int gtop_problems_create_dense(int32_t num_problems, int32_t m, const double* seg_time, const double* L,
                               const double* R, const double* Df, const gtop_params* params, int device,
                               gtop_problems** out);
/* Native path: waypoints (num_problems x (m+1) x 3, row-major points) -> time allocation, per-segment
 * mapping blocks, banded R, Df, x0 and bounds, built block-wise in O(m) (no dense (6m)^2 inverse). */
int gtop_problems_create_from_paths(int32_t num_problems, int32_t m, const double* waypoints, const gtop_params* params,
                                    int device, gtop_problems** out);
/* Front-end hand-off (setKinoPath, grad_traj_optimizer.cpp:35-65): positions / velocities / accelerations at
 * the m+1 waypoints (num_problems x (m+1) x 3 each, row-major) and the segment times (num_problems x m) of a
 * kinodynamic search (getKinoTrajMat).  x0 carries the interior velocities / accelerations. */
int gtop_problems_create_from_kino(int32_t num_problems, int32_t m, const double* pos, const double* vel,
                                   const double* acc, const double* time, const gtop_params* params, int device,
                                   gtop_problems** out);
/* The same two set-ups run ON THE DEVICE, one CTA per problem (csrc/setup_kernels.cu): only the waypoints / times are
 * uploaded and the per-problem records are written in place — for batches with one waypoint set per trajectory, where
 * the host set-up above costs far more than evaluating the batch.  Same arithmetic as the host path (shared source,
 * no FMA contraction) except pow(), which is correctly rounded on the device and libm's on the host (1 ulp apart in
 * about 0.08 % of the calls): the records agree to rounding level (<= 1e-12 relative).  All accessors work; dense_L / dense_R /
 * coeff_from_x set the requested problem up on the host on first use. */
int gtop_problems_create_from_paths_device(int32_t num_problems, int32_t m, const double* waypoints,
                                           const gtop_params* params, int device, gtop_problems** out);
int gtop_problems_create_from_kino_device(int32_t num_problems, int32_t m, const double* pos, const double* vel,
                                          const double* acc, const double* time, const gtop_params* params, int device,
                                          gtop_problems** out);
int gtop_problems_destroy(gtop_problems* p);
int gtop_problems_dims(const gtop_problems* p, int32_t* num_problems, int32_t* m, int32_t* n /* = 9(m-1) */);
int gtop_problems_set_params(gtop_problems* p, const gtop_params* params); /* weights / penalty constants only */
/* host copies of set-up products (native or dense path) */
int gtop_problems_seg_time(const gtop_problems* p, double* seg_time /* np x m */);
int gtop_problems_x0(const gtop_problems* p, double* x0 /* np x n */);          /* Dp flattened (:182-187) */
int gtop_problems_bounds(const gtop_problems* p, double* lb, double* ub);       /* np x n each (:151-177) */
/* dense views for API parity with getL()/getR()/Df (col-major, per problem); native path only */
int gtop_problems_dense_L(const gtop_problems* p, int32_t problem, double* L);
int gtop_problems_dense_R(const gtop_problems* p, int32_t problem, double* R);
/* inspection: the packed device record of one problem (csrc/gtop_types.cuh, ProbLayout) copied to the host; *len = its
 * length in doubles (rec may be NULL to query it) */
int gtop_problems_record(const gtop_problems* p, int32_t problem, double* rec, int64_t cap, int64_t* len);
int gtop_problems_Df(const gtop_problems* p, int32_t problem, double* Df /* 3x6 col-major */);

/* ---------------------------------------------------------------- evaluation ---- */
/* Trajectory b (0 <= b < B) of this call is global trajectory (traj_offset + b) and belongs to
 * problem ((traj_offset + b) / starts_per_problem) % num_problems.
 * X: B x n, cost: B, grad: B x n.  `stream` is a cudaStream_t (NULL = the handle's own stream).
 * The *_device form takes device pointers and only enqueues (asynchronous w.r.t. the host);
 * the *_host form copies X in, evaluates, copies cost/grad out and synchronises.
 * n_out_of_map (nullable, host) receives the number of samples that fell outside the map. */
int gtop_eval_batch_device(const gtop_sdf* sdf, const gtop_problems* problems, const double* X_dev, int64_t B,
                           int64_t traj_offset, int64_t starts_per_problem, int step, double* cost_dev,
                           double* grad_dev, void* stream);
int gtop_eval_batch_host(const gtop_sdf* sdf, const gtop_problems* problems, const double* X, int64_t B,
                         int64_t traj_offset, int64_t starts_per_problem, int step, double* cost, double* grad,
                         int64_t* n_out_of_map);

/* ---------------------------------------------------------------- post-hoc metrics ---- */
/* What opti_node computes from one optimised trajectory through PolynomialTraj (src/opti_node.cpp:104-144,
 * include/grad_traj_optimization/polynomial_traj.hpp:70-203), for a batch: per trajectory GTOP_NUM_TRAJ_METRICS doubles
 *   [0] time_sum  [1] length (getTraj every 0.01 s + getLength)  [2] jerk (getJerk)  [3] mean_v  [4] max_v
 *   [5] mean_a  [6] max_a (getMeanAndMaxVel / Acc, literal: evaluated at the segment duration)  [7] acc cost (getAccCost)
 * from the variables X (B x n, as the optimiser returns them) and the problems' mapping blocks (getCoefficient). */
#define GTOP_NUM_TRAJ_METRICS 8
int gtop_traj_metrics_device(const gtop_problems* problems, const double* X_dev, int64_t B, int64_t traj_offset,
                             int64_t starts_per_problem, double* metrics_dev /* B x 8 */, void* stream);
int gtop_traj_metrics_host(const gtop_problems* problems, const double* X, int64_t B, int64_t traj_offset,
                           int64_t starts_per_problem, double* metrics /* B x 8 */);

/* cumulative out-of-map sample counter of the problems handle (reset on read) */
int gtop_problems_take_out_of_map(gtop_problems* p, int64_t* n);
/* number of kernel launches issued by this library since load (bench.py's gpu_launches claim) */
int64_t gtop_kernel_launch_count(void);

/* coefficients (m x 18 row-major: [x0..x5 | y0..y5 | z0..z5], ascending powers) of one trajectory */
int gtop_coeff_from_x(const gtop_problems* problems, int32_t problem, const double* x /* n, host */, double* coeff);

/* ---------------------------------------------------------------- NLopt callback ---- */
/* ctx plays the role of `this` in set_min_objective(GradTrajOptimizer::costFunc, this). */
int gtop_objective_ctx_create(const gtop_sdf* sdf, const gtop_problems* problems, int32_t problem, int step,
                              gtop_objective_ctx** out);
int gtop_objective_ctx_set_step(gtop_objective_ctx* ctx, int step);
int gtop_objective_ctx_destroy(gtop_objective_ctx* ctx);
int64_t gtop_objective_ctx_evals(const gtop_objective_ctx* ctx);
/* nlopt_func signature: double f(unsigned n, const double* x, double* grad, void* f_data) */
double gtop_objective(unsigned n, const double* x, double* grad, void* ctx);

/* Measured FP64 ceilings of the device in TFLOP/s: independent DFMA chains and independent mma.sync.m8n8k4.f64 (DMMA)
 * chains (about 25 ms).  bench.py reports the evaluation kernel's FP64 work against them. */
int gtop_measure_fp64_peak(int device, double* dfma_tflops, double* dmma_tflops);
/* Self-check of the EDT's fast integer square root: the number of integers v in [1, vmax] for which it differs from
 * the IEEE sqrt((double)v) that SDFMap::updateESDF3d applies (sdf_map.cpp:356-360).  Must be 0 for vmax = 3*8192^2. */
int gtop_debug_check_sqrt(int64_t vmax, int64_t* mismatches);
/* Host-only inspection of the packed per-problem record the evaluation kernels read (csrc/gtop_types.cuh): set-up of one
 * path as gtop_problems_create_from_paths does it, no device involved.  layout = {m, ns_cap, nsp, npair, nit2, stride,
 * off_dt, off_ns, off_tt, off_TQ, off_PL, off_LiTP, off_Linv, off_T, off_Rc, off_Rg}; rec may be NULL to query it. */
int gtop_debug_pack_record(int32_t m, const double* waypoints, const gtop_params* params, double* rec, int64_t cap,
                           int32_t layout[16]);

/* ---------------------------------------------------------------- optimiser loop ---- */
/* Replaces the solver call of GradTrajOptimizer::optimizeTrajectory (grad_traj_optimizer.cpp:135-195:
 * nlopt::opt(nlopt::algorithm(24) = LD_MMA, n), set_lower/upper_bounds, set_min_objective, optimize) without the
 * NLopt dependency: Svanberg's MMA for box bounds and no nonlinear constraints, as NLopt 2.5.0 runs it on this
 * path (csrc/mma_core.cuh), stopping after `maxeval` objective evaluations.
 *
 * gtop_mma_minimize: one problem on the host, any objective with NLopt's nlopt_func signature (for example
 * gtop_objective).  x: in = start (inside the bounds), out = best point; *minf = its value. */
typedef double (*gtop_nlopt_func)(unsigned n, const double* x, double* grad, void* f_data);
int gtop_mma_minimize(unsigned n, gtop_nlopt_func f, void* f_data, const double* lb, const double* ub, double* x,
                      double* minf, int maxeval, int* evals_done);
/* The same loop for a batch of independent trajectories in lock-step on the GPU: every step is one batched
 * cost+gradient evaluation (gtop_eval_batch_device) plus one bookkeeping kernel; x never leaves the device.
 * Bounds are the problems' own NLopt box (gtop_problems_bounds; native / kino path).  X: B x n, in = starts
 * (clipped into the box), out = best points; minf: B (nullable).  The *_host form copies X in and out. */
int gtop_solve_batch_device(const gtop_sdf* sdf, const gtop_problems* problems, double* X_dev, int64_t B,
                            int64_t traj_offset, int64_t starts_per_problem, int step, int maxeval, double* minf_dev,
                            void* stream);
int gtop_solve_batch_host(const gtop_sdf* sdf, const gtop_problems* problems, double* X, int64_t B, int64_t traj_offset,
                          int64_t starts_per_problem, int step, int maxeval, double* minf);

#ifdef __cplusplus
}
#endif
#endif /* GTOP_C_H_ */
