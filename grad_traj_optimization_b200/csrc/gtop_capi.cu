// gtop_capi.cu — implementation of the C ABI declared in include/gtop_c.h.
// Host logic only: handle management, set-up packing, stream/copy orchestration; every compute
// step is a launch of a kernel in eval_kernel.cu / sdf_kernels.cu.  No CPU fallback.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/gtop_c.h"
#include "gtop_types.cuh"
#include "mma_core.cuh"
#include "problem_setup.hpp"
#include "setup_core.cuh"

namespace gtop {
cudaError_t launch_eval(const SdfView& S, const ProbLayout& PL, const double* prob, int num_problems,
                        const EvalScalars& E, const double* X, int64_t B, int64_t traj_offset,
                        int64_t starts_per_problem, double* cost, double* grad, unsigned long long* oob_counter,
                        int max_ctas, cudaStream_t stream);
bool eval_takes_mma_kernel(const ProbLayout& PL, const EvalScalars& E, int64_t B, int num_problems, int64_t starts_per_problem);
cudaError_t launch_fill(double* dist, int64_t ndist, uint8_t* occ, int64_t n, double dval, cudaStream_t st);
cudaError_t launch_relayout(const double* src, double* dst, const DistLayout& lay, int nx, bool to_linear, cudaStream_t st);
cudaError_t launch_build_quads(const SdfView& S, const double* dist, void* quad, cudaStream_t st);
cudaError_t launch_set_occupancy(const SdfView& S, uint8_t* occ, const double* pts_dev, int64_t n, cudaStream_t st);
cudaError_t launch_set_boxes(int nx, int ny, int nz, uint8_t* occ, const int32_t* boxes_dev, int64_t n,
                             cudaStream_t st);
cudaError_t launch_edt(const uint8_t* occ, int32_t* d1, int32_t* d2, double* dist, int nx, int ny, int nz, double res,
                       bool fresh, cudaStream_t st, cudaEvent_t* ev, const DistLayout* lay);
cudaError_t launch_edt_window(const uint8_t* occ, uint8_t* wocc, int32_t* d1, int32_t* d2, double* wdist, double* dist,
                              const VoxelBox& b, const DistLayout& lay, double res, cudaStream_t st, cudaEvent_t* ev);
cudaError_t launch_build_quads_box(const SdfView& S, const double* dist, void* quad, const VoxelBox& b, cudaStream_t st);
cudaError_t launch_setup_problems(const ProbLayout& PL, double mean_v, double init_time, const double* pos_dev,
                                  const double* time_dev, int np, double* rec_dev, int* status_dev, cudaStream_t st);
cudaError_t launch_traj_metrics(const ProbLayout& PL, const double* prob, int num_problems, const double* X, int64_t B,
                                int64_t traj_offset, int64_t starts_per_problem, double* out, cudaStream_t st);
cudaError_t launch_query_dynamic(const SdfView& S, const double* boxes_dev, int nbox, double time, const double* pos_dev,
                                 int64_t n, double* dist_dev, double* grad_dev, double* coarse_dev, cudaStream_t st);
cudaError_t launch_box_reset(double* dist, uint8_t* occ, const VoxelBox& b, const DistLayout& lay, double dval, cudaStream_t st);
cudaError_t measure_fp64_peak(double* dfma_tflops, double* dmma_tflops, cudaStream_t st);
cudaError_t launch_check_sqrt(long long vmax, unsigned long long* mismatches_dev, cudaStream_t st);
bool eval_one_applicable(const ProbLayout& PL);
cudaError_t launch_eval_one(const SdfView& S, const ProbLayout& PL, const double* prob, int num_problems,
                            const EvalScalars& E, const double* X, int64_t B, int64_t traj_offset,
                            int64_t starts_per_problem, double* cost, double* grad, unsigned long long* oob_counter,
                            cudaStream_t stream);
cudaError_t launch_query(const SdfView& S, const double* pos_dev, int64_t n, double* dist_dev, double* grad_dev,
                         unsigned long long* oob_dev, cudaStream_t st);
struct SolveView {
  int n;
  int64_t B, traj_offset, starts_per_problem;
  int num_problems;
  const double* lb;
  const double* ub;
  double* xcur;
  const double* grad;
  const double* cost;
  double* x;
  double* dfdx;
  double* sigma;
  double* xprev;
  double* xprevprev;
  MmaScalars* sc;
};
cudaError_t launch_mma_init(const SolveView& V, cudaStream_t st);
cudaError_t launch_mma_step(const SolveView& V, int want_candidate, cudaStream_t st);
cudaError_t launch_mma_finish(const SolveView& V, double* X, double* minf, cudaStream_t st);
}  // namespace gtop

using namespace gtop;

// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static std::atomic<int64_t> g_launches{0};

static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define CU(call)                                                                                  \
  do {                                                                                            \
    cudaError_t e__ = (call);                                                                     \
    if (e__ != cudaSuccess)                                                                       \
      return fail(GTOP_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));              \
  } while (0)
#define REQUIRE(cond, msg) \
  do {                     \
    if (!(cond)) return fail(GTOP_E_INVALID, msg); \
  } while (0)

struct gtop_sdf {
  int device = 0;
  SdfView view{};
  double origin[3], map_size[3];
  int64_t nvox = 0;
  int64_t ndist = 0;  // doubles in `dist` (brick-padded for maps without a quad copy)
  double* dist = nullptr;
  void* quad = nullptr;  // xy-quad copy of dist (SdfView::quad), rebuilt whenever dist changes; nullptr for big maps
  uint8_t* occ = nullptr;
  int32_t *d1 = nullptr, *d2 = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // update start, after z, y, x, after the quad build
  float last_ms[6] = {0, 0, 0, 0, 0, 0};                                // z, y, x, z+y+x, quad build, call to evaluator-ready
  bool dist_fresh = false;  // distance buffer is 10000 everywhere (just created / reset)
  // update window of the next gtop_sdf_update_esdf (SDFMap::min_vec / max_vec, sdf_map.cpp:13-14): voxel indices,
  // inclusive; the whole grid unless gtop_sdf_set_update_range was called since the last update (:366-367)
  int win_lo[3] = {0, 0, 0}, win_hi[3] = {0, 0, 0};
  uint8_t* wocc = nullptr;  // compact occupancy / distances of a partial window (grown on demand)
  double* wdist = nullptr;
  int64_t wcap = 0;
};

struct gtop_problems {
  int device = 0;
  int np = 0;
  ProbLayout layout{};
  gtop_params prm{};
  // host copies of the problems.  Sets created on the device (gtop_problems_create_*_device) only carry the cheap
  // parts here (segment times, Df, x0, NLopt box); the per-segment blocks of problem i are set up on the host the
  // first time an accessor needs them (dense_L / dense_R / coeff_from_x) from the retained inputs below.
  mutable std::vector<ProblemHost> host;
  bool device_setup = false;
  std::vector<double> in_pos, in_vel, in_acc, in_time;  // inputs of a device set-up ([np][m+1][3], [np][m])
  double* rec = nullptr;               // np * layout.stride doubles
  unsigned long long* oob = nullptr;   // device counter
  unsigned long long* h_oob = nullptr;  // pinned landing slot of the counter for the *_host pipeline
  cudaStream_t streams[3] = {nullptr, nullptr, nullptr};  // [0] compute (and default), [1] H2D, [2] D2H
  cudaEvent_t ev_in[16] = {}, ev_done[16] = {};            // per-chunk hand-offs of gtop_eval_batch_host
  // staging for the *_host entry point (grown on demand)
  double *dX = nullptr, *dcost = nullptr, *dgrad = nullptr;
  int64_t cap = 0;
  // workspace of gtop_solve_batch_* (grown on demand): bounds [np][n]; xcur/grad [B][n]; cost [B];
  // five state vectors [n][B]; scalars [B]
  double *s_lb = nullptr, *s_ub = nullptr, *s_xcur = nullptr, *s_grad = nullptr, *s_cost = nullptr, *s_state = nullptr;
  MmaScalars* s_sc = nullptr;
  int64_t s_cap = 0;
};

struct gtop_objective_ctx {
  const gtop_sdf* sdf = nullptr;
  const gtop_problems* problems = nullptr;
  int problem = 0, step = 2, n = 0;
  double *hx = nullptr, *hout = nullptr;  // pinned: x[n]; out = cost + grad[n]
  double *hx_dev = nullptr, *hout_dev = nullptr;  // the same buffers as the device addresses them
  double *dx = nullptr, *dout = nullptr;
  cudaStream_t stream = nullptr;
  int64_t evals = 0;
};

static EvalScalars make_scalars(const gtop_params& p, int step) {
  EvalScalars E;
  E.ws = (step == GTOP_OPT_FIRST_STEP) ? 0.0 : p.w_smooth;  // grad_traj_optimizer.cpp:412-415
  E.wc = p.w_collision;
  E.alpha = p.alpha;
  E.d0 = p.d0;
  E.inv_r = 1.0 / p.r;
  E.alpha_over_r = p.alpha / p.r;
  const double log2e = 1.4426950408889634074;
  E.neg_log2e_over_r = -log2e / p.r;
  E.d0_log2e_over_r = p.d0 * log2e / p.r;
  E.skip_collision = std::fabs(p.w_collision) < 1e-4 ? 1 : 0;  // :346
  // the optional block sits inside the sample loop, under `if (step == 2)` (:384)
  E.vel_acc = (p.vel_acc_penalty && step == GTOP_OPT_SECOND_STEP && !E.skip_collision) ? 1 : 0;
  E.alpha_v = p.alpha_v, E.alpha_v_over_r = p.alpha_v / p.r_v;
  E.log2e_over_rv = log2e / p.r_v, E.neg_v0_log2e_over_rv = -p.v0 * log2e / p.r_v;
  E.alpha_a = p.alpha_a, E.alpha_a_over_r = p.alpha_a / p.r_a;
  E.log2e_over_ra = log2e / p.r_a, E.neg_a0_log2e_over_ra = -p.a0 * log2e / p.r_a;
  E.inv_wc = E.skip_collision ? 0.0 : 1.0 / p.w_collision;
  return E;
}

extern "C" {

int gtop_abi_version(void) { return GTOP_ABI_VERSION; }
const char* gtop_last_error(void) { return g_err.c_str(); }
int64_t gtop_kernel_launch_count(void) { return g_launches.load(); }

int gtop_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

void gtop_params_default(gtop_params* p) {
  if (!p) return;
  std::memset(p, 0, sizeof(*p));
  p->time_limit_1 = 0.04, p->time_limit_2 = 0.06, p->deltat = 0.2;
  p->w_smooth = 1.0, p->w_collision = 1.0;
  p->alpha = 1.0, p->r = 1.0, p->d0 = 1.0;
  p->alpha_v = 0.0, p->r_v = 1.0, p->v0 = 1.0;
  p->alpha_a = 0.0, p->r_a = 1.0, p->a0 = 1.0;
  p->bos = 3.0, p->vos = 8.0, p->aos = 10.0;
  p->mean_v = 1.0, p->mean_a = 1.0, p->init_time = 0.0;
  p->algorithm = 24, p->samples_per_segment = 30, p->vel_acc_penalty = 0;
}

int gtop_host_alloc(void** ptr, size_t bytes) {
  REQUIRE(ptr, "null ptr");
  CU(cudaMallocHost(ptr, bytes));
  return GTOP_OK;
}
int gtop_host_free(void* ptr) {
  if (ptr) CU(cudaFreeHost(ptr));
  return GTOP_OK;
}

// ================================================================================ SDF map
int gtop_sdf_create(const double origin[3], double resolution, const double map_size[3], int device, gtop_sdf** out) {
  REQUIRE(origin && map_size && out, "null argument");
  REQUIRE(resolution > 0.0, "resolution must be > 0");
  REQUIRE(gtop_device_count() > device && device >= 0, "no such CUDA device (this library has no CPU fallback)");
  CU(cudaSetDevice(device));
  gtop_sdf* s = new gtop_sdf();
  s->device = device;
  int g[3];
  for (int i = 0; i < 3; ++i) {
    s->origin[i] = origin[i], s->map_size[i] = map_size[i];
    g[i] = (int)std::ceil(map_size[i] / resolution);  // sdf_map.cpp:9
    if (g[i] < 1 || g[i] > 16384) {
      delete s;
      return fail(GTOP_E_INVALID, "grid dimension out of range [1, 16384]");
    }
  }
  if ((int64_t)g[0] * g[1] * g[2] > 2147483647LL || quad_elems(g[0], g[1], g[2]) > 2147483647LL) {
    delete s;
    return fail(GTOP_E_INVALID, "more than 2^31-1 voxels");
  }
  SdfView& v = s->view;
  v.nx = g[0], v.ny = g[1], v.nz = g[2];
  v.nxh = (g[0] + 1) / 2, v.nyh = (g[1] + 1) / 2;
  v.ox = origin[0], v.oy = origin[1], v.oz = origin[2];
  v.res = resolution, v.res_inv = 1 / resolution, v.half_res = 0.5 * resolution;
  v.lox = origin[0] + 1e-4, v.loy = origin[1] + 1e-4, v.loz = origin[2] + 1e-4;  // sdf_map.cpp:56-57
  v.hix = (origin[0] + map_size[0]) - 1e-4, v.hiy = (origin[1] + map_size[1]) - 1e-4,
  v.hiz = (origin[2] + map_size[2]) - 1e-4;  // :62-63 with max_range = origin + map_size (:12)
  {  // float images of the isInMap bounds (exact for float-valued positions, see SdfView)
    auto up = [](double b) { float f = (float)b; return (double)f < b ? std::nextafterf(f, INFINITY) : f; };
    auto down = [](double b) { float f = (float)b; return (double)f > b ? std::nextafterf(f, -INFINITY) : f; };
    v.lofx = up(v.lox), v.lofy = up(v.loy), v.lofz = up(v.loz);
    v.hifx = down(v.hix), v.hify = down(v.hiy), v.hifz = down(v.hiz);
  }
  s->nvox = (int64_t)g[0] * g[1] * g[2];
  for (int i = 0; i < 3; ++i) s->win_lo[i] = 0, s->win_hi[i] = g[i] - 1;
  // The xy-quad copy (4x the field) halves the L1 wavefronts of a gather and pays while it stays cache-resident; a map
  // whose copy would not (default: above 1 GiB) is gathered straight from `dist`: no 4x footprint, and no rebuild of the
  // copy — more bytes than the whole EDT moves — between an update and the first evaluation.
  const char* qenv = std::getenv("GTOP_QUAD_LIMIT_MB");  // read per map: tests build both layouts in one process
  const double quad_limit_mb = qenv ? std::atof(qenv) : 1024.0;
  const size_t quad_bytes = 32 * (size_t)quad_elems(g[0], g[1], g[2]);
  const bool with_quad = (double)quad_bytes <= quad_limit_mb * 1048576.0;
  // ... and keeps the field itself in 8 x 8 x 4 bricks (DistLayout) instead of the reference's linear order
  v.lay.ny = g[1], v.lay.nz = g[2], v.lay.nby = (g[1] + 7) / 8, v.lay.nbz = (g[2] + 3) / 4, v.lay.bricks = with_quad ? 0 : 1;
  s->ndist = dist_elems(g[0], g[1], g[2], !with_quad);
  cudaError_t e = cudaMalloc(&s->dist, sizeof(double) * s->ndist);
  if (e == cudaSuccess && with_quad) e = cudaMalloc(&s->quad, quad_bytes);
  if (e == cudaSuccess) e = cudaMalloc(&s->occ, s->nvox);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
  for (int i = 0; i < 5 && e == cudaSuccess; ++i) e = cudaEventCreate(&s->ev[i]);
  if (e != cudaSuccess) {
    gtop_sdf_destroy(s);
    return fail(GTOP_E_ALLOC, std::string("sdf allocation: ") + cudaGetErrorString(e));
  }
  v.dist = s->dist;
  v.quad = s->quad;
  e = launch_fill(s->dist, s->ndist, s->occ, s->nvox, 10000.0, s->stream);  // sdf_map.cpp:22-23
  s->dist_fresh = true;
  if (e == cudaSuccess && s->quad) e = launch_build_quads(v, s->dist, s->quad, s->stream);
  g_launches += 2;
  if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
  if (e != cudaSuccess) {
    gtop_sdf_destroy(s);
    return fail(GTOP_E_CUDA, std::string("sdf initialisation: ") + cudaGetErrorString(e));
  }
  *out = s;
  return GTOP_OK;
}

int gtop_sdf_destroy(gtop_sdf* s) {
  if (!s) return GTOP_OK;
  cudaSetDevice(s->device);
  cudaFree(s->dist), cudaFree(s->quad), cudaFree(s->occ), cudaFree(s->d1), cudaFree(s->d2);
  cudaFree(s->wocc), cudaFree(s->wdist);
  for (auto& e : s->ev)
    if (e) cudaEventDestroy(e);
  if (s->stream) cudaStreamDestroy(s->stream);
  delete s;
  return GTOP_OK;
}

int gtop_sdf_grid(const gtop_sdf* s, int32_t grid[3]) {
  REQUIRE(s && grid, "null argument");
  grid[0] = s->view.nx, grid[1] = s->view.ny, grid[2] = s->view.nz;
  return GTOP_OK;
}

int gtop_sdf_reset(gtop_sdf* s) {
  REQUIRE(s, "null sdf");
  CU(cudaSetDevice(s->device));
  CU(launch_fill(s->dist, s->ndist, s->occ, s->nvox, 10000.0, s->stream));
  s->dist_fresh = true;
  if (s->quad) CU(launch_build_quads(s->view, s->dist, s->quad, s->stream));
  g_launches += 2;
  CU(cudaStreamSynchronize(s->stream));
  return GTOP_OK;
}

// Voxel box of a metric box as resetBuffer / setUpdateRange compute it (sdf_map.cpp:29-43, :248-259): clamp to the map,
// posToIndex(min_pos), posToIndex(max_pos - res/2) with posToIndex = floor((pos - origin) * res_inv) (:71-74).
static int metric_box_to_voxels(const gtop_sdf* s, const double min_pos[3], const double max_pos[3], int lo[3], int hi[3]) {
  const int g[3] = {s->view.nx, s->view.ny, s->view.nz};
  for (int i = 0; i < 3; ++i) {
    const double min_range = s->origin[i], max_range = s->origin[i] + s->map_size[i];  // :11-12
    const double a = std::max(min_pos[i], min_range);
    const double b = std::min(max_pos[i], max_range) - s->view.res / 2;
    lo[i] = (int)std::floor((a - s->origin[i]) * s->view.res_inv);
    hi[i] = (int)std::floor((b - s->origin[i]) * s->view.res_inv);
    // the reference indexes its buffers with whatever comes out (undefined behaviour for a box outside the map
    // or an inverted one); here that is an error
    if (!(a == a) || !(b == b) || lo[i] < 0 || hi[i] > g[i] - 1 || lo[i] > hi[i]) return GTOP_E_INVALID;
  }
  return GTOP_OK;
}
static VoxelBox window_box(const int lo[3], const int hi[3]) {
  return VoxelBox{lo[0], lo[1], lo[2], hi[0] - lo[0] + 1, hi[1] - lo[1] + 1, hi[2] - lo[2] + 1};
}

int gtop_sdf_reset_range(gtop_sdf* s, const double min_pos[3], const double max_pos[3]) {
  REQUIRE(s && min_pos && max_pos, "null argument");
  int lo[3], hi[3];
  if (metric_box_to_voxels(s, min_pos, max_pos, lo, hi) != GTOP_OK)
    return fail(GTOP_E_INVALID, "reset range: empty box or box outside the map");
  CU(cudaSetDevice(s->device));
  const VoxelBox b = window_box(lo, hi);
  CU(launch_box_reset(s->dist, s->occ, b, s->view.lay, 10000.0, s->stream));  // sdf_map.cpp:45-52
  if (s->quad) CU(launch_build_quads_box(s->view, s->dist, s->quad, b, s->stream));
  g_launches += 2;
  s->dist_fresh = false;  // (conservative: only a whole-grid reset re-arms the read-free x pass)
  CU(cudaStreamSynchronize(s->stream));
  return GTOP_OK;
}

int gtop_sdf_set_update_range(gtop_sdf* s, const double min_pos[3], const double max_pos[3]) {
  REQUIRE(s && min_pos && max_pos, "null argument");
  int lo[3], hi[3];
  if (metric_box_to_voxels(s, min_pos, max_pos, lo, hi) != GTOP_OK)
    return fail(GTOP_E_INVALID, "update range: empty box or box outside the map");
  for (int i = 0; i < 3; ++i) s->win_lo[i] = lo[i], s->win_hi[i] = hi[i];
  return GTOP_OK;
}

int gtop_sdf_update_range(const gtop_sdf* s, int32_t lo_hi[6]) {
  REQUIRE(s && lo_hi, "null argument");
  for (int i = 0; i < 3; ++i) lo_hi[i] = s->win_lo[i], lo_hi[3 + i] = s->win_hi[i];
  return GTOP_OK;
}

int gtop_sdf_set_occupancy(gtop_sdf* s, const double* xyz, int64_t n) {
  REQUIRE(s && (xyz || n == 0) && n >= 0, "bad argument");
  if (n == 0) return GTOP_OK;
  CU(cudaSetDevice(s->device));
  double* d = nullptr;
  CU(cudaMalloc(&d, sizeof(double) * 3 * n));
  cudaError_t e = cudaMemcpyAsync(d, xyz, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, s->stream);
  if (e == cudaSuccess) e = launch_set_occupancy(s->view, s->occ, d, n, s->stream);
  ++g_launches;
  if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
  cudaFree(d);
  CU(e);
  return GTOP_OK;
}

int gtop_sdf_set_occupancy_boxes(gtop_sdf* s, const int32_t* lo_hi, int64_t n) {
  REQUIRE(s && (lo_hi || n == 0) && n >= 0, "bad argument");
  if (n == 0) return GTOP_OK;
  CU(cudaSetDevice(s->device));
  int32_t* d = nullptr;
  CU(cudaMalloc(&d, sizeof(int32_t) * 6 * n));
  cudaError_t e = cudaMemcpyAsync(d, lo_hi, sizeof(int32_t) * 6 * n, cudaMemcpyHostToDevice, s->stream);
  if (e == cudaSuccess) e = launch_set_boxes(s->view.nx, s->view.ny, s->view.nz, s->occ, d, n, s->stream);
  ++g_launches;
  if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
  cudaFree(d);
  CU(e);
  return GTOP_OK;
}

int gtop_sdf_update_esdf(gtop_sdf* s) {
  REQUIRE(s, "null sdf");
  CU(cudaSetDevice(s->device));
  if (!s->d1 || !s->d2) {
    cudaFree(s->d1), cudaFree(s->d2);
    s->d1 = s->d2 = nullptr;
    cudaError_t e = cudaMalloc(&s->d1, sizeof(int32_t) * s->nvox);
    if (e == cudaSuccess) e = cudaMalloc(&s->d2, sizeof(int32_t) * s->nvox);
    if (e != cudaSuccess) {
      cudaFree(s->d1), cudaFree(s->d2);
      s->d1 = s->d2 = nullptr;
      return fail(GTOP_E_ALLOC, std::string("edt scratch: ") + cudaGetErrorString(e));
    }
  }
  const int g[3] = {s->view.nx, s->view.ny, s->view.nz};
  const bool whole = s->win_lo[0] == 0 && s->win_lo[1] == 0 && s->win_lo[2] == 0 && s->win_hi[0] == g[0] - 1 &&
                     s->win_hi[1] == g[1] - 1 && s->win_hi[2] == g[2] - 1;
  if (whole) {
    CU(launch_edt(s->occ, s->d1, s->d2, s->dist, g[0], g[1], g[2], s->view.res, s->dist_fresh, s->stream, s->ev, &s->view.lay));
    if (s->quad) CU(launch_build_quads(s->view, s->dist, s->quad, s->stream));
    g_launches += 4;
  } else {  // the box setUpdateRange left (sdf_map.cpp:244-264)
    const VoxelBox b = window_box(s->win_lo, s->win_hi);
    const int64_t n = (int64_t)b.wx * b.wy * b.wz;
    if (n > s->wcap) {
      cudaFree(s->wocc), cudaFree(s->wdist);
      s->wocc = nullptr, s->wdist = nullptr, s->wcap = 0;
      cudaError_t e = cudaMalloc(&s->wocc, (size_t)n);
      if (e == cudaSuccess) e = cudaMalloc(&s->wdist, sizeof(double) * (size_t)n);
      if (e != cudaSuccess) return fail(GTOP_E_ALLOC, std::string("edt window scratch: ") + cudaGetErrorString(e));
      s->wcap = n;
    }
    CU(launch_edt_window(s->occ, s->wocc, s->d1, s->d2, s->wdist, s->dist, b, s->view.lay, s->view.res, s->stream, s->ev));
    if (s->quad) CU(launch_build_quads_box(s->view, s->dist, s->quad, b, s->stream));
    g_launches += 6;
    for (int i = 0; i < 3; ++i) s->win_lo[i] = 0, s->win_hi[i] = g[i] - 1;  // :366-367
  }
  s->dist_fresh = false;
  CU(cudaEventRecord(s->ev[4], s->stream));  // the map is ready for the evaluation kernels from here
  CU(cudaStreamSynchronize(s->stream));
  for (int i = 0; i < 3; ++i) cudaEventElapsedTime(&s->last_ms[i], s->ev[i], s->ev[i + 1]);
  cudaEventElapsedTime(&s->last_ms[3], s->ev[0], s->ev[3]);
  cudaEventElapsedTime(&s->last_ms[4], s->ev[3], s->ev[4]);
  cudaEventElapsedTime(&s->last_ms[5], s->ev[0], s->ev[4]);
  return GTOP_OK;
}

int gtop_measure_fp64_peak(int device, double* dfma_tflops, double* dmma_tflops) {
  REQUIRE(dfma_tflops && dmma_tflops, "null argument");
  REQUIRE(gtop_device_count() > device && device >= 0, "no such CUDA device (this library has no CPU fallback)");
  CU(cudaSetDevice(device));
  CU(measure_fp64_peak(dfma_tflops, dmma_tflops, nullptr));
  g_launches += 4;
  return GTOP_OK;
}

int gtop_debug_check_sqrt(int64_t vmax, int64_t* mismatches) {
  REQUIRE(mismatches && vmax >= 1 && vmax <= 2147483647LL, "bad argument");
  REQUIRE(gtop_device_count() >= 1, "no CUDA device (this library has no CPU fallback)");
  unsigned long long* d = nullptr;
  CU(cudaMalloc(&d, sizeof(unsigned long long)));
  cudaError_t e = cudaMemset(d, 0, sizeof(unsigned long long));
  if (e == cudaSuccess) e = launch_check_sqrt(vmax, d, nullptr);
  ++g_launches;
  unsigned long long h = 0;
  if (e == cudaSuccess) e = cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost);
  cudaFree(d);
  CU(e);
  *mismatches = (int64_t)h;
  return GTOP_OK;
}

int gtop_sdf_last_edt_ms(const gtop_sdf* s, float ms[4]) {
  REQUIRE(s && ms, "null argument");
  for (int i = 0; i < 4; ++i) ms[i] = s->last_ms[i];
  return GTOP_OK;
}
int gtop_sdf_last_update_ms(const gtop_sdf* s, float ms[6]) {
  REQUIRE(s && ms, "null argument");
  for (int i = 0; i < 6; ++i) ms[i] = s->last_ms[i];
  return GTOP_OK;
}
int gtop_sdf_gather_layout(const gtop_sdf* s, int32_t* has_quad_copy, int64_t* resident_bytes) {
  REQUIRE(s, "null sdf");
  const int64_t quad = s->quad ? 32 * quad_elems(s->view.nx, s->view.ny, s->view.nz) : 0;
  if (has_quad_copy) *has_quad_copy = s->quad ? 1 : 0;
  // distances + occupancy + (EDT scratch once an update has run) + quad copy
  if (resident_bytes) *resident_bytes = s->ndist * 8 + s->nvox + (s->d1 ? s->nvox * 8 : 0) + quad;
  return GTOP_OK;
}

int gtop_sdf_download_distance(const gtop_sdf* s, double* host) {
  REQUIRE(s && host, "null argument");
  CU(cudaSetDevice(s->device));
  if (!s->view.lay.bricks) {
    CU(cudaMemcpy(host, s->dist, sizeof(double) * s->nvox, cudaMemcpyDeviceToHost));
    return GTOP_OK;
  }
  // bricks -> the reference's linear order through a device staging buffer
  double* tmp = nullptr;
  CU(cudaMalloc(&tmp, sizeof(double) * s->nvox));
  cudaError_t e = launch_relayout(s->dist, tmp, s->view.lay, s->view.nx, true, s->stream);
  ++g_launches;
  if (e == cudaSuccess) e = cudaMemcpyAsync(host, tmp, sizeof(double) * s->nvox, cudaMemcpyDeviceToHost, s->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
  cudaFree(tmp);
  CU(e);
  return GTOP_OK;
}
int gtop_sdf_upload_distance(gtop_sdf* s, const double* host) {
  REQUIRE(s && host, "null argument");
  CU(cudaSetDevice(s->device));
  if (!s->view.lay.bricks) {
    CU(cudaMemcpy(s->dist, host, sizeof(double) * s->nvox, cudaMemcpyHostToDevice));
  } else {
    double* tmp = nullptr;
    CU(cudaMalloc(&tmp, sizeof(double) * s->nvox));
    cudaError_t e = cudaMemcpyAsync(tmp, host, sizeof(double) * s->nvox, cudaMemcpyHostToDevice, s->stream);
    if (e == cudaSuccess) e = launch_relayout(tmp, s->dist, s->view.lay, s->view.nx, false, s->stream);
    ++g_launches;
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    cudaFree(tmp);
    CU(e);
  }
  s->dist_fresh = false;
  if (s->quad) CU(launch_build_quads(s->view, s->dist, s->quad, s->stream));
  ++g_launches;
  CU(cudaStreamSynchronize(s->stream));
  return GTOP_OK;
}
int gtop_sdf_download_occupancy(const gtop_sdf* s, int32_t* host) {
  REQUIRE(s && host, "null argument");
  CU(cudaSetDevice(s->device));
  std::vector<uint8_t> tmp(s->nvox);
  CU(cudaMemcpy(tmp.data(), s->occ, s->nvox, cudaMemcpyDeviceToHost));
  for (int64_t i = 0; i < s->nvox; ++i) host[i] = tmp[i];
  return GTOP_OK;
}
int gtop_sdf_upload_occupancy(gtop_sdf* s, const int32_t* host) {
  REQUIRE(s && host, "null argument");
  CU(cudaSetDevice(s->device));
  std::vector<uint8_t> tmp(s->nvox);
  for (int64_t i = 0; i < s->nvox; ++i) tmp[i] = host[i] == 1 ? 1 : 0;
  CU(cudaMemcpy(s->occ, tmp.data(), s->nvox, cudaMemcpyHostToDevice));
  return GTOP_OK;
}
int gtop_sdf_distance_device_ptr(const gtop_sdf* s, const double** p) {
  REQUIRE(s && p, "null argument");
  REQUIRE(!s->view.lay.bricks, "the map keeps its field in 8x8x4 bricks (no quad copy): use gtop_sdf_download_distance");
  *p = s->dist;
  return GTOP_OK;
}

int gtop_sdf_query(const gtop_sdf* s, const double* pos, int64_t n, double* dist, double* grad, int64_t* n_oob) {
  REQUIRE(s && n >= 0 && (n == 0 || (pos && dist && grad)), "bad argument");
  if (n_oob) *n_oob = 0;
  if (n == 0) return GTOP_OK;
  CU(cudaSetDevice(s->device));
  double *dp = nullptr, *dd = nullptr, *dg = nullptr;
  unsigned long long* dc = nullptr;
  cudaError_t e = cudaMalloc(&dp, sizeof(double) * 3 * n);
  if (e == cudaSuccess) e = cudaMalloc(&dd, sizeof(double) * n);
  if (e == cudaSuccess) e = cudaMalloc(&dg, sizeof(double) * 3 * n);
  if (e == cudaSuccess) e = cudaMalloc(&dc, sizeof(unsigned long long));
  unsigned long long hc = 0;
  if (e == cudaSuccess) e = cudaMemsetAsync(dc, 0, sizeof(unsigned long long), s->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dp, pos, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, s->stream);
  if (e == cudaSuccess) e = launch_query(s->view, dp, n, dd, dg, dc, s->stream);
  ++g_launches;
  if (e == cudaSuccess) e = cudaMemcpyAsync(dist, dd, sizeof(double) * n, cudaMemcpyDeviceToHost, s->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(grad, dg, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, s->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(&hc, dc, sizeof(hc), cudaMemcpyDeviceToHost, s->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
  cudaFree(dp), cudaFree(dd), cudaFree(dg), cudaFree(dc);
  CU(e);
  if (n_oob) *n_oob = (int64_t)hc;
  return GTOP_OK;
}

int gtop_sdf_query_dynamic(const gtop_sdf* s, const double* boxes, int32_t nbox, double time, const double* pos, int64_t n,
                           double* dist, double* grad, double* coarse) {
  REQUIRE(s && n >= 0 && nbox >= 0 && (nbox == 0 || boxes) && (n == 0 || (pos && dist && grad)), "bad argument");
  REQUIRE(nbox <= 1024, "at most 1024 boxes");
  if (n == 0) return GTOP_OK;
  CU(cudaSetDevice(s->device));
  double *dp = nullptr, *dd = nullptr, *dg = nullptr, *dc = nullptr, *db = nullptr;
  cudaError_t e = cudaMalloc(&dp, sizeof(double) * 3 * n);
  if (e == cudaSuccess) e = cudaMalloc(&dd, sizeof(double) * n);
  if (e == cudaSuccess) e = cudaMalloc(&dg, sizeof(double) * 3 * n);
  if (e == cudaSuccess && coarse) e = cudaMalloc(&dc, sizeof(double) * n);
  if (e == cudaSuccess && nbox > 0) e = cudaMalloc(&db, sizeof(double) * 9 * nbox);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dp, pos, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, s->stream);
  if (e == cudaSuccess && nbox > 0) e = cudaMemcpyAsync(db, boxes, sizeof(double) * 9 * nbox, cudaMemcpyHostToDevice, s->stream);
  if (e == cudaSuccess) e = launch_query_dynamic(s->view, db, nbox, time, dp, n, dd, dg, dc, s->stream);
  ++g_launches;
  if (e == cudaSuccess) e = cudaMemcpyAsync(dist, dd, sizeof(double) * n, cudaMemcpyDeviceToHost, s->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(grad, dg, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, s->stream);
  if (e == cudaSuccess && coarse) e = cudaMemcpyAsync(coarse, dc, sizeof(double) * n, cudaMemcpyDeviceToHost, s->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
  cudaFree(dp), cudaFree(dd), cudaFree(dg), cudaFree(dc), cudaFree(db);
  CU(e);
  return GTOP_OK;
}

// ================================================================================ problems
static int finish_handle(gtop_problems* p, cudaError_t e, gtop_problems** out);
static int finish_problems(gtop_problems* p, gtop_problems** out) {
  const ProbLayout& L = p->layout;
  std::vector<double> rec((size_t)p->np * L.stride);
  {  // independent records: pack them on all host threads
    std::atomic<int> next(0);
    auto work = [&]() {
      for (int i; (i = next.fetch_add(1)) < p->np;) pack_record(p->host[i], rec.data() + (size_t)i * L.stride);
    };
    int nt = (int)std::thread::hardware_concurrency();
    nt = std::max(1, std::min(nt, p->np / 64));
    std::vector<std::thread> th;
    for (int i = 1; i < nt; ++i) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
  }
  cudaError_t e = cudaMalloc(&p->rec, sizeof(double) * rec.size());
  if (e == cudaSuccess) e = cudaMemcpy(p->rec, rec.data(), sizeof(double) * rec.size(), cudaMemcpyHostToDevice);
  return finish_handle(p, e, out);
}

// counters, streams and events of a problem set whose record is in place (or failed with `e`)
static int finish_handle(gtop_problems* p, cudaError_t e, gtop_problems** out) {
  if (e == cudaSuccess) e = cudaMalloc(&p->oob, sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMemset(p->oob, 0, sizeof(unsigned long long));
  if (e == cudaSuccess && cudaMallocHost(&p->h_oob, sizeof(unsigned long long)) != cudaSuccess) p->h_oob = nullptr, cudaGetLastError();
  for (int i = 0; i < 3 && e == cudaSuccess; ++i) e = cudaStreamCreateWithFlags(&p->streams[i], cudaStreamNonBlocking);
  for (int i = 0; i < 16 && e == cudaSuccess; ++i) {
    e = cudaEventCreateWithFlags(&p->ev_in[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ev_done[i], cudaEventDisableTiming);
  }
  if (e != cudaSuccess) {
    gtop_problems_destroy(p);
    return fail(GTOP_E_ALLOC, std::string("problem upload: ") + cudaGetErrorString(e));
  }
  *out = p;
  return GTOP_OK;
}

// segment-count limit: the record layout (gtop_types.cuh) is indexed with 32-bit offsets
static bool layout_fits(int m, int ns) { return m >= 2 && m <= 4096 && (int64_t)m * ns * 110 + (int64_t)m * 400 < ((int64_t)1 << 30); }

static int check_params(const gtop_params* prm) {
  REQUIRE(prm, "null params");
  REQUIRE(prm->samples_per_segment >= 1 && prm->samples_per_segment <= 4096, "samples_per_segment out of range");
  REQUIRE(prm->r != 0.0, "r must be non-zero");
  // the table exp2 of the kernels takes |y| < 2^26 with y = (d0 - d) log2(e) / r, |d| <= 10000 (unseen voxels)
  REQUIRE((std::fabs(prm->d0) + 1.0e4) * 1.4427 / std::fabs(prm->r) < 3.0e7, "|d0| / r or 1e4 / r too large for the penalty's range");
  if (prm->vel_acc_penalty)
    REQUIRE(prm->r_v != 0.0 && prm->r_a != 0.0 && (std::fabs(prm->v0) + 1.0e3) * 1.4427 / std::fabs(prm->r_v) < 3.0e7 &&
                (std::fabs(prm->a0) + 1.0e4) * 1.4427 / std::fabs(prm->r_a) < 3.0e7,
            "r_v / r_a must be non-zero and v0 / r_v, a0 / r_a in the penalty's range with vel_acc_penalty");
  if (prm->vel_acc_penalty) REQUIRE(prm->r_v != 0.0 && prm->r_a != 0.0, "r_v and r_a must be non-zero with vel_acc_penalty");
  return GTOP_OK;
}

int gtop_problems_create_dense(int32_t np, int32_t m, const double* seg_time, const double* L, const double* R,
                               const double* Df, const gtop_params* prm, int device, gtop_problems** out) {
  REQUIRE(np >= 1 && m >= 2 && seg_time && L && R && Df && out, "bad argument");
  REQUIRE(prm && layout_fits(m, prm->samples_per_segment), "segments x samples_per_segment out of range");
  if (int rc = check_params(prm)) return rc;
  REQUIRE(gtop_device_count() > device && device >= 0, "no such CUDA device (this library has no CPU fallback)");
  CU(cudaSetDevice(device));
  gtop_problems* p = new gtop_problems();
  p->device = device, p->np = np, p->prm = *prm;
  p->layout = make_layout(m, prm->samples_per_segment);
  p->host.resize(np);
  const size_t szL = (size_t)6 * m * (3 * m + 3), szR = (size_t)(3 * m + 3) * (3 * m + 3);
  for (int i = 0; i < np; ++i) {
    std::string err =
        setup_from_dense(m, seg_time + (size_t)i * m, L + i * szL, R + i * szR, Df + (size_t)i * 18, *prm, p->host[i]);
    if (!err.empty()) {
      delete p;
      return fail(GTOP_E_INVALID, "problem " + std::to_string(i) + ": " + err);
    }
  }
  return finish_problems(p, out);
}

int gtop_problems_create_from_paths(int32_t np, int32_t m, const double* waypoints, const gtop_params* prm, int device,
                                    gtop_problems** out) {
  REQUIRE(np >= 1 && m >= 2 && waypoints && out, "bad argument");
  REQUIRE(prm && layout_fits(m, prm->samples_per_segment), "segments x samples_per_segment out of range");
  if (int rc = check_params(prm)) return rc;
  REQUIRE(gtop_device_count() > device && device >= 0, "no such CUDA device (this library has no CPU fallback)");
  CU(cudaSetDevice(device));
  gtop_problems* p = new gtop_problems();
  p->device = device, p->np = np, p->prm = *prm;
  p->layout = make_layout(m, prm->samples_per_segment);
  p->host.resize(np);
  // independent problems: set them up on all host threads
  std::atomic<int> next(0), bad(-1);
  std::vector<std::string> errs(np);
  auto work = [&]() {
    for (;;) {
      const int i = next.fetch_add(1);
      if (i >= np) break;
      errs[i] = setup_from_path(m, waypoints + (size_t)i * (m + 1) * 3, *prm, p->host[i]);
      if (!errs[i].empty()) bad.store(i);
    }
  };
  int nt = (int)std::thread::hardware_concurrency();
  nt = std::max(1, std::min(nt, np / 64));
  std::vector<std::thread> th;
  for (int i = 1; i < nt; ++i) th.emplace_back(work);
  work();
  for (auto& t : th) t.join();
  if (bad.load() >= 0) {
    const int i = bad.load();
    std::string msg = "problem " + std::to_string(i) + ": " + errs[i];
    delete p;
    return fail(GTOP_E_INVALID, msg);
  }
  return finish_problems(p, out);
}

int gtop_problems_create_from_kino(int32_t np, int32_t m, const double* pos, const double* vel, const double* acc,
                                   const double* time, const gtop_params* prm, int device, gtop_problems** out) {
  REQUIRE(np >= 1 && m >= 2 && pos && vel && acc && time && out, "bad argument");
  REQUIRE(prm && layout_fits(m, prm->samples_per_segment), "segments x samples_per_segment out of range");
  if (int rc = check_params(prm)) return rc;
  REQUIRE(gtop_device_count() > device && device >= 0, "no such CUDA device (this library has no CPU fallback)");
  CU(cudaSetDevice(device));
  gtop_problems* p = new gtop_problems();
  p->device = device, p->np = np, p->prm = *prm;
  p->layout = make_layout(m, prm->samples_per_segment);
  p->host.resize(np);
  const size_t w = (size_t)(m + 1) * 3;
  for (int i = 0; i < np; ++i) {
    std::string err = setup_from_kino(m, pos + i * w, vel + i * w, acc + i * w, time + (size_t)i * m, *prm, p->host[i]);
    if (!err.empty()) {
      delete p;
      return fail(GTOP_E_INVALID, "problem " + std::to_string(i) + ": " + err);
    }
  }
  return finish_problems(p, out);
}

// Set-up on the device (setup_kernels.cu): only the waypoints (and, for the kino form, the segment times) are uploaded.
static int create_on_device(int32_t np, int32_t m, const double* pos, const double* vel, const double* acc,
                            const double* time, const gtop_params* prm, int device, gtop_problems** out) {
  if (int rc = check_params(prm)) return rc;
  REQUIRE(time || prm->mean_v > 0.0, "mean_v must be > 0");
  REQUIRE(gtop_device_count() > device && device >= 0, "no such CUDA device (this library has no CPU fallback)");
  CU(cudaSetDevice(device));
  gtop_problems* p = new gtop_problems();
  p->device = device, p->np = np, p->prm = *prm, p->device_setup = true;
  p->layout = make_layout(m, prm->samples_per_segment);
  const ProbLayout& L = p->layout;
  const size_t w = (size_t)(m + 1) * 3;
  p->in_pos.assign(pos, pos + (size_t)np * w);
  if (vel) p->in_vel.assign(vel, vel + (size_t)np * w);
  if (acc) p->in_acc.assign(acc, acc + (size_t)np * w);
  if (time) p->in_time.assign(time, time + (size_t)np * m);
  // the cheap host parts: segment times (same IEEE operations as the kernel), Df, x0 and the NLopt box
  p->host.resize(np);
  for (int i = 0; i < np; ++i) {
    ProblemHost& h = p->host[i];
    const double* wp = pos + i * w;
    h.m = m, h.ns_cap = prm->samples_per_segment;
    h.seg_time.resize(m);
    for (int s = 0; s < m; ++s) {
      h.seg_time[s] = time ? time[(size_t)i * m + s] : setup_segment_time(wp, s, prm->mean_v, prm->init_time);
      if (!(h.seg_time[s] > 0.0)) {
        delete p;
        return fail(GTOP_E_INVALID, "problem " + std::to_string(i) + ": non-positive segment time (coincident waypoints?)");
      }
    }
    h.Df.assign(18, 0.0);
    for (int a = 0; a < 3; ++a) h.Df[a * 6 + 0] = wp[a], h.Df[a * 6 + 3] = wp[3 * m + a];
    h.x0.resize(L.n), h.lb.resize(L.n), h.ub.resize(L.n);
    for (int a = 0; a < 3; ++a)
      for (int q = 0; q < L.P; ++q)
        setup_x0_box(wp, vel ? vel + i * w : nullptr, acc ? acc + i * w : nullptr, a, q, prm->bos, prm->vos, prm->aos,
                     &h.x0[a * L.P + q], &h.lb[a * L.P + q], &h.ub[a * L.P + q]);
    h.has_path = true;
  }
  double *d_pos = nullptr, *d_time = nullptr;
  int* d_status = nullptr;
  std::vector<int> status(np, 0);
  cudaError_t e = cudaMalloc(&p->rec, sizeof(double) * (size_t)np * L.stride);
  if (e == cudaSuccess) e = cudaMalloc(&d_pos, sizeof(double) * (size_t)np * w);
  if (e == cudaSuccess && time) e = cudaMalloc(&d_time, sizeof(double) * (size_t)np * m);
  if (e == cudaSuccess) e = cudaMalloc(&d_status, sizeof(int) * (size_t)np);
  if (e == cudaSuccess) e = cudaMemcpy(d_pos, pos, sizeof(double) * (size_t)np * w, cudaMemcpyHostToDevice);
  if (e == cudaSuccess && time) e = cudaMemcpy(d_time, time, sizeof(double) * (size_t)np * m, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = launch_setup_problems(L, prm->mean_v, prm->init_time, d_pos, d_time, np, p->rec, d_status, nullptr);
  ++g_launches;
  if (e == cudaSuccess) e = cudaMemcpy(status.data(), d_status, sizeof(int) * (size_t)np, cudaMemcpyDeviceToHost);
  cudaFree(d_pos), cudaFree(d_time), cudaFree(d_status);
  if (e == cudaSuccess)
    for (int i = 0; i < np; ++i)
      if (status[i]) {
        gtop_problems_destroy(p);
        return fail(GTOP_E_INVALID, "problem " + std::to_string(i) + (status[i] == 2 ? ": singular segment mapping matrix"
                                                                                    : ": non-positive segment time"));
      }
  return finish_handle(p, e, out);
}

int gtop_problems_create_from_paths_device(int32_t np, int32_t m, const double* waypoints, const gtop_params* prm,
                                           int device, gtop_problems** out) {
  REQUIRE(np >= 1 && m >= 2 && waypoints && out, "bad argument");
  REQUIRE(prm && layout_fits(m, prm->samples_per_segment), "segments x samples_per_segment out of range");
  return create_on_device(np, m, waypoints, nullptr, nullptr, nullptr, prm, device, out);
}

int gtop_problems_create_from_kino_device(int32_t np, int32_t m, const double* pos, const double* vel, const double* acc,
                                          const double* time, const gtop_params* prm, int device, gtop_problems** out) {
  REQUIRE(np >= 1 && m >= 2 && pos && vel && acc && time && out, "bad argument");
  REQUIRE(prm && layout_fits(m, prm->samples_per_segment), "segments x samples_per_segment out of range");
  return create_on_device(np, m, pos, vel, acc, time, prm, device, out);
}

// the per-segment blocks of problem i on the host (device-built sets set them up on first use)
static int ensure_blocks(const gtop_problems* p, int i) {
  if (!p->device_setup || !p->host[i].Linv.empty()) return GTOP_OK;
  const int m = p->layout.m;
  const size_t w = (size_t)(m + 1) * 3;
  ProblemHost h;
  const std::string err =
      p->in_time.empty() ? setup_from_path(m, p->in_pos.data() + i * w, p->prm, h)
                         : setup_from_kino(m, p->in_pos.data() + i * w, p->in_vel.data() + i * w, p->in_acc.data() + i * w,
                                           p->in_time.data() + (size_t)i * m, p->prm, h);
  if (!err.empty()) return fail(GTOP_E_INVALID, "problem " + std::to_string(i) + ": " + err);
  p->host[i] = std::move(h);
  return GTOP_OK;
}

int gtop_problems_destroy(gtop_problems* p) {
  if (!p) return GTOP_OK;
  cudaSetDevice(p->device);
  cudaFree(p->rec), cudaFree(p->oob), cudaFree(p->dX), cudaFree(p->dcost), cudaFree(p->dgrad);
  if (p->h_oob) cudaFreeHost(p->h_oob);
  cudaFree(p->s_lb), cudaFree(p->s_ub), cudaFree(p->s_xcur), cudaFree(p->s_grad), cudaFree(p->s_cost), cudaFree(p->s_state);
  cudaFree(p->s_sc);
  for (auto& s : p->streams)
    if (s) cudaStreamDestroy(s);
  for (int i = 0; i < 16; ++i) {
    if (p->ev_in[i]) cudaEventDestroy(p->ev_in[i]);
    if (p->ev_done[i]) cudaEventDestroy(p->ev_done[i]);
  }
  delete p;
  return GTOP_OK;
}

int gtop_problems_dims(const gtop_problems* p, int32_t* np, int32_t* m, int32_t* n) {
  REQUIRE(p, "null problems");
  if (np) *np = p->np;
  if (m) *m = p->layout.m;
  if (n) *n = p->layout.n;
  return GTOP_OK;
}

int gtop_problems_set_params(gtop_problems* p, const gtop_params* prm) {
  REQUIRE(p, "null problems");
  if (int rc = check_params(prm)) return rc;
  REQUIRE(prm->samples_per_segment == p->prm.samples_per_segment, "samples_per_segment is fixed at creation");
  // baked into the device records, x0 and the NLopt box when the set was created: only the weights, the penalty
  // constants, the solver settings and the velocity / acceleration flag can change on a live set
  REQUIRE(prm->mean_v == p->prm.mean_v && prm->init_time == p->prm.init_time && prm->bos == p->prm.bos &&
              prm->vos == p->prm.vos && prm->aos == p->prm.aos,
          "mean_v, init_time, bos, vos, aos are fixed at creation: create a new problem set");
  p->prm = *prm;
  return GTOP_OK;
}

int gtop_problems_seg_time(const gtop_problems* p, double* out) {
  REQUIRE(p && out, "null argument");
  for (int i = 0; i < p->np; ++i) std::memcpy(out + (size_t)i * p->layout.m, p->host[i].seg_time.data(), sizeof(double) * p->layout.m);
  return GTOP_OK;
}
int gtop_problems_x0(const gtop_problems* p, double* out) {
  REQUIRE(p && out, "null argument");
  if (!p->host[0].has_path) return fail(GTOP_E_STATE, "x0 is only known for problems created from paths");
  for (int i = 0; i < p->np; ++i) std::memcpy(out + (size_t)i * p->layout.n, p->host[i].x0.data(), sizeof(double) * p->layout.n);
  return GTOP_OK;
}
int gtop_problems_bounds(const gtop_problems* p, double* lb, double* ub) {
  REQUIRE(p && lb && ub, "null argument");
  if (!p->host[0].has_path) return fail(GTOP_E_STATE, "bounds are only known for problems created from paths");
  for (int i = 0; i < p->np; ++i) {
    std::memcpy(lb + (size_t)i * p->layout.n, p->host[i].lb.data(), sizeof(double) * p->layout.n);
    std::memcpy(ub + (size_t)i * p->layout.n, p->host[i].ub.data(), sizeof(double) * p->layout.n);
  }
  return GTOP_OK;
}
int gtop_problems_dense_L(const gtop_problems* p, int32_t i, double* L) {
  REQUIRE(p && L && i >= 0 && i < p->np, "bad argument");
  if (int rc = ensure_blocks(p, i)) return rc;
  dense_L(p->host[i], L);
  return GTOP_OK;
}
int gtop_problems_dense_R(const gtop_problems* p, int32_t i, double* R) {
  REQUIRE(p && R && i >= 0 && i < p->np, "bad argument");
  if (int rc = ensure_blocks(p, i)) return rc;
  dense_R(p->host[i], R);
  return GTOP_OK;
}
int gtop_problems_Df(const gtop_problems* p, int32_t i, double* Df) {
  REQUIRE(p && Df && i >= 0 && i < p->np, "bad argument");
  for (int a = 0; a < 3; ++a)
    for (int j = 0; j < 6; ++j) Df[j * 3 + a] = p->host[i].Df[a * 6 + j];
  return GTOP_OK;
}

// host-only: the packed record of one problem set up from a path, and the offsets of its sections (no device involved)
int gtop_debug_pack_record(int32_t m, const double* waypoints, const gtop_params* prm, double* rec, int64_t cap,
                           int32_t layout[16]) {
  REQUIRE(m >= 2 && waypoints && layout, "bad argument");
  if (int rc = check_params(prm)) return rc;
  const ProbLayout L = make_layout(m, prm->samples_per_segment);
  const int32_t lay[16] = {L.m,      L.ns_cap,  L.nsp,   L.npair,  L.nit2,   L.stride, L.off_dt, L.off_ns,
                           L.off_tt, L.off_TQ,  L.off_PL, L.off_LiTP, L.off_Linv, L.off_T, L.off_Rc, L.off_Rg};
  std::memcpy(layout, lay, sizeof(lay));
  if (!rec) return GTOP_OK;
  REQUIRE(cap >= L.stride, "record buffer too small");
  ProblemHost h;
  const std::string err = setup_from_path(m, waypoints, *prm, h);
  if (!err.empty()) return fail(GTOP_E_INVALID, err);
  pack_record(h, rec);
  return GTOP_OK;
}

int gtop_problems_record(const gtop_problems* p, int32_t i, double* rec, int64_t cap, int64_t* len) {
  REQUIRE(p && i >= 0 && i < p->np, "bad argument");
  const int64_t n = p->layout.stride;
  if (len) *len = n;
  if (!rec) return GTOP_OK;
  REQUIRE(cap >= n, "record buffer too small");
  CU(cudaSetDevice(p->device));
  CU(cudaMemcpy(rec, p->rec + (size_t)i * n, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost));
  return GTOP_OK;
}

int gtop_coeff_from_x(const gtop_problems* p, int32_t i, const double* x, double* coeff) {
  REQUIRE(p && x && coeff && i >= 0 && i < p->np, "bad argument");
  if (int rc = ensure_blocks(p, i)) return rc;
  coeff_from_x(p->host[i], x, coeff);
  return GTOP_OK;
}

// ================================================================================ evaluation
static int check_eval(const gtop_sdf* sdf, const gtop_problems* pr, int64_t B, int64_t off, int64_t starts, int step) {
  REQUIRE(sdf && pr, "null handle");
  REQUIRE(sdf->device == pr->device, "sdf and problems live on different devices");
  REQUIRE(B >= 0 && starts >= 1, "bad batch shape");
  REQUIRE(off >= 0, "traj_offset must be >= 0 (it selects the problem of every trajectory)");
  REQUIRE(B <= (int64_t)1 << 40 && off <= (int64_t)1 << 60, "batch too large");
  REQUIRE(step >= 0 && step <= 2, "step must be 0, 1 or 2");
  return GTOP_OK;
}

int gtop_eval_batch_device(const gtop_sdf* sdf, const gtop_problems* pr, const double* X, int64_t B, int64_t off,
                           int64_t starts, int step, double* cost, double* grad, void* stream) {
  if (int rc = check_eval(sdf, pr, B, off, starts, step)) return rc;
  if (B == 0) return GTOP_OK;
  REQUIRE(X && cost && grad, "null device buffer");
  CU(cudaSetDevice(pr->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : pr->streams[0];
  CU(launch_eval(sdf->view, pr->layout, pr->rec, pr->np, make_scalars(pr->prm, step), X, B, off, starts, cost, grad,
                 pr->oob, 0, st));
  ++g_launches;
  return GTOP_OK;
}

// ================================================================================ post-hoc metrics
int gtop_traj_metrics_device(const gtop_problems* pr, const double* X, int64_t B, int64_t off, int64_t starts,
                             double* metrics, void* stream) {
  REQUIRE(pr, "null problems");
  REQUIRE(B >= 0 && starts >= 1 && off >= 0, "bad batch shape");
  if (B == 0) return GTOP_OK;
  REQUIRE(X && metrics, "null device buffer");
  CU(cudaSetDevice(pr->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : pr->streams[0];
  CU(launch_traj_metrics(pr->layout, pr->rec, pr->np, X, B, off, starts, metrics, st));
  ++g_launches;
  return GTOP_OK;
}

int gtop_traj_metrics_host(const gtop_problems* pr, const double* X, int64_t B, int64_t off, int64_t starts,
                           double* metrics) {
  REQUIRE(pr, "null problems");
  REQUIRE(B >= 0 && starts >= 1, "bad batch shape");
  if (B == 0) return GTOP_OK;
  REQUIRE(X && metrics, "null host buffer");
  CU(cudaSetDevice(pr->device));
  const int n = pr->layout.n;
  double *dX = nullptr, *dM = nullptr;
  cudaError_t e = cudaMalloc(&dX, sizeof(double) * (size_t)B * n);
  if (e == cudaSuccess) e = cudaMalloc(&dM, sizeof(double) * (size_t)B * GTOP_NUM_TRAJ_METRICS);
  cudaStream_t st = pr->streams[0];
  if (e == cudaSuccess) e = cudaMemcpyAsync(dX, X, sizeof(double) * (size_t)B * n, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = launch_traj_metrics(pr->layout, pr->rec, pr->np, dX, B, off, starts, dM, st);
  ++g_launches;
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(metrics, dM, sizeof(double) * (size_t)B * GTOP_NUM_TRAJ_METRICS, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(dX), cudaFree(dM);
  CU(e);
  return GTOP_OK;
}

int gtop_eval_batch_host(const gtop_sdf* sdf, const gtop_problems* cpr, const double* X, int64_t B, int64_t off,
                         int64_t starts, int step, double* cost, double* grad, int64_t* n_oob) {
  if (int rc = check_eval(sdf, cpr, B, off, starts, step)) return rc;
  if (n_oob) *n_oob = 0;
  if (B == 0) return GTOP_OK;
  REQUIRE(X && cost && grad, "null host buffer");
  gtop_problems* pr = const_cast<gtop_problems*>(cpr);  // staging buffers are a cache
  CU(cudaSetDevice(pr->device));
  const int n = pr->layout.n;
  const EvalScalars E = make_scalars(pr->prm, step);
  // Pinned (or registered) host buffers are addressable from the device: the multi-start kernel then reads its x blocks
  // and writes its gradient blocks straight across PCIe with the bulk copies it uses anyway (5 KB per group of 8
  // trajectories, hundreds of groups in flight) — one launch, both directions busy from the first block to the last, no
  // staging copies and no cross-stream hand-offs (each costs ~10 us on the chunked pipeline below): 1.14 against 1.29 ms
  // per 65,536 batch.  Pageable buffers, and batches for the general kernels (8-byte loads), take the staged pipeline.
  static const bool zero_copy = [] { const char* e = std::getenv("GTOP_HOST_ZEROCOPY"); return !e || std::atoi(e) != 0; }();
  if (zero_copy && eval_takes_mma_kernel(pr->layout, E, B, pr->np, starts)) {
    auto device_view = [](const void* h) -> void* {
      cudaPointerAttributes a;
      if (cudaPointerGetAttributes(&a, h) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
      }
      return a.type == cudaMemoryTypeHost ? a.devicePointer : nullptr;
    };
    const double* zx = static_cast<const double*>(device_view(X));
    double* zc = static_cast<double*>(device_view(cost));
    double* zg = static_cast<double*>(device_view(grad));
    if (zx && zc && zg) {
      // (measured and dropped: the tail of the batch brought in by DMA meanwhile and evaluated by a second launch — the
      // link, not the initiator, is the limit: 1.20-1.33 ms against 1.14; x by DMA in chunks with direct stores: 1.21-1.49)
      cudaStream_t s_k = pr->streams[0];
      cudaError_t err = launch_eval(sdf->view, pr->layout, pr->rec, pr->np, E, zx, B, off, starts, zc, zg, pr->oob, 0, s_k);
      ++g_launches;
      if (err == cudaSuccess && n_oob && pr->h_oob) {
        err = cudaMemcpyAsync(pr->h_oob, pr->oob, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s_k);
        if (err == cudaSuccess) err = cudaMemsetAsync(pr->oob, 0, sizeof(unsigned long long), s_k);
      }
      const cudaError_t es = cudaStreamSynchronize(s_k);
      if (err == cudaSuccess) err = es;
      CU(err);
      if (n_oob) {
        if (!pr->h_oob) return gtop_problems_take_out_of_map(pr, n_oob);
        *n_oob = (int64_t)*pr->h_oob;
      }
      return GTOP_OK;
    }
  }
  if (pr->cap < B) {
    cudaFree(pr->dX), cudaFree(pr->dcost), cudaFree(pr->dgrad);
    pr->dX = pr->dcost = pr->dgrad = nullptr, pr->cap = 0;
    cudaError_t e = cudaMalloc(&pr->dX, sizeof(double) * B * n);
    if (e == cudaSuccess) e = cudaMalloc(&pr->dgrad, sizeof(double) * B * n);
    if (e == cudaSuccess) e = cudaMalloc(&pr->dcost, sizeof(double) * B);
    if (e != cudaSuccess) return fail(GTOP_E_ALLOC, std::string("eval staging: ") + cudaGetErrorString(e));
    pr->cap = B;
  }
  // chunked pipeline over three streams: all H2D copies of x back to back on the copy-in stream, the kernels
  // on the compute stream as their chunk lands, cost/gradient copied out on the copy-out stream as each
  // kernel finishes (the two PCIe directions and the SMs are busy at the same time)
  static const int64_t max_chunks = [] { const char* e = std::getenv("GTOP_HOST_CHUNKS"); return e ? std::max(1L, std::min(16L, std::atol(e))) : 16L; }();
  const int64_t nchunk = std::min<int64_t>(max_chunks, (B + 4095) / 4096);
  const int64_t per = ((B + nchunk - 1) / nchunk + 63) & ~int64_t(63);  // whole lane groups per chunk
  cudaStream_t s_k = pr->streams[0], s_in = pr->streams[1], s_out = pr->streams[2];
  cudaError_t err = cudaSuccess;
  auto ok = [&](cudaError_t e) { return err == cudaSuccess && (err = e) == cudaSuccess; };
  for (int64_t c = 0; c < nchunk && err == cudaSuccess; ++c) {
    const int64_t b0 = c * per, bn = std::min(per, B - b0);
    if (bn <= 0) break;
    ok(cudaMemcpyAsync(pr->dX + b0 * n, X + b0 * n, sizeof(double) * bn * n, cudaMemcpyHostToDevice, s_in)) &&
        ok(cudaEventRecord(pr->ev_in[c], s_in)) && ok(cudaStreamWaitEvent(s_k, pr->ev_in[c], 0)) &&
        ok(launch_eval(sdf->view, pr->layout, pr->rec, pr->np, E, pr->dX + b0 * n, bn, off + b0, starts, pr->dcost + b0,
                       pr->dgrad + b0 * n, pr->oob, 0, s_k)) &&
        ok(cudaEventRecord(pr->ev_done[c], s_k)) && ok(cudaStreamWaitEvent(s_out, pr->ev_done[c], 0)) &&
        ok(cudaMemcpyAsync(grad + b0 * n, pr->dgrad + b0 * n, sizeof(double) * bn * n, cudaMemcpyDeviceToHost, s_out));
    // (one copy of all costs behind the last gradient block measured SLOWER than these small per-chunk copies: 1.46
    // against 1.30 ms per 65,536 batch with 16 chunks, same box)
    ok(cudaMemcpyAsync(cost + b0, pr->dcost + b0, sizeof(double) * bn, cudaMemcpyDeviceToHost, s_out));
    ++g_launches;
  }
  // the out-of-map counter leaves behind the last block (no separate synchronous read)
  if (n_oob && err == cudaSuccess && pr->h_oob)
    ok(cudaMemcpyAsync(pr->h_oob, pr->oob, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s_out)) &&
        ok(cudaMemsetAsync(pr->oob, 0, sizeof(unsigned long long), s_out));
  // whatever happened, nothing may still be writing into the caller's buffers when this returns
  for (auto& st : pr->streams) {
    const cudaError_t e = cudaStreamSynchronize(st);
    if (err == cudaSuccess) err = e;
  }
  CU(err);
  if (n_oob) {
    if (!pr->h_oob) return gtop_problems_take_out_of_map(pr, n_oob);
    *n_oob = (int64_t)*pr->h_oob;
  }
  return GTOP_OK;
}

int gtop_problems_take_out_of_map(gtop_problems* p, int64_t* n) {
  REQUIRE(p && n, "null argument");
  CU(cudaSetDevice(p->device));
  unsigned long long h = 0;
  CU(cudaMemcpy(&h, p->oob, sizeof(h), cudaMemcpyDeviceToHost));
  CU(cudaMemset(p->oob, 0, sizeof(h)));
  *n = (int64_t)h;
  return GTOP_OK;
}

// ================================================================================ optimiser loop
int gtop_mma_minimize(unsigned n, gtop_nlopt_func f, void* f_data, const double* lb, const double* ub, double* x,
                      double* minf, int maxeval, int* evals_done) {
  REQUIRE(n > 0 && f && lb && ub && x && maxeval >= 1, "bad argument");
  for (unsigned j = 0; j < n; ++j) REQUIRE(lb[j] <= ub[j] && x[j] >= lb[j] && x[j] <= ub[j], "start outside the bounds");
  std::vector<double> xcur(x, x + n), g(n), xb(n), gb(n), sigma(n), xp(n), xpp(n);
  MmaScalars s;
  const MmaCVec vlb{lb, 1}, vub{ub, 1};
  mma_init((int)n, vlb, vub, MmaVec{sigma.data(), 1}, s);
  for (int e = 0; e < maxeval; ++e) {
    const double fc = f(n, xcur.data(), g.data(), f_data);
    if (!(fc < HUGE_VAL)) return fail(GTOP_E_STATE, "objective returned HUGE_VAL / NaN (forced stop)");
    mma_after_eval((int)n, fc, MmaCVec{g.data(), 1}, MmaVec{xb.data(), 1}, MmaVec{gb.data(), 1}, MmaVec{sigma.data(), 1}, vlb,
                   vub, MmaCVec{xcur.data(), 1}, MmaVec{xp.data(), 1}, MmaVec{xpp.data(), 1}, s);
    if (e + 1 < maxeval)
      mma_candidate((int)n, MmaCVec{xb.data(), 1}, MmaCVec{gb.data(), 1}, MmaCVec{sigma.data(), 1}, vlb, vub,
                    MmaVec{xcur.data(), 1}, s);
  }
  std::memcpy(x, xb.data(), sizeof(double) * n);
  if (minf) *minf = s.fval;
  if (evals_done) *evals_done = s.evals;
  return GTOP_OK;
}

static int solve_workspace(gtop_problems* pr, int64_t B) {
  const int n = pr->layout.n;
  if (!pr->host[0].has_path) return fail(GTOP_E_STATE, "the optimiser loop needs the problems' NLopt box (create them from paths)");
  if (!pr->s_lb) {
    std::vector<double> lb((size_t)pr->np * n), ub((size_t)pr->np * n);
    if (int rc = gtop_problems_bounds(pr, lb.data(), ub.data())) return rc;
    cudaError_t e = cudaMalloc(&pr->s_lb, sizeof(double) * lb.size());
    if (e == cudaSuccess) e = cudaMalloc(&pr->s_ub, sizeof(double) * ub.size());
    if (e == cudaSuccess) e = cudaMemcpy(pr->s_lb, lb.data(), sizeof(double) * lb.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(pr->s_ub, ub.data(), sizeof(double) * ub.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return fail(GTOP_E_ALLOC, std::string("solver bounds: ") + cudaGetErrorString(e));
  }
  if (pr->s_cap < B) {
    cudaFree(pr->s_xcur), cudaFree(pr->s_grad), cudaFree(pr->s_cost), cudaFree(pr->s_state), cudaFree(pr->s_sc);
    pr->s_xcur = pr->s_grad = pr->s_cost = pr->s_state = nullptr, pr->s_sc = nullptr, pr->s_cap = 0;
    cudaError_t e = cudaMalloc(&pr->s_xcur, sizeof(double) * B * n);
    if (e == cudaSuccess) e = cudaMalloc(&pr->s_grad, sizeof(double) * B * n);
    if (e == cudaSuccess) e = cudaMalloc(&pr->s_cost, sizeof(double) * B);
    if (e == cudaSuccess) e = cudaMalloc(&pr->s_state, sizeof(double) * 5 * B * n);
    if (e == cudaSuccess) e = cudaMalloc(&pr->s_sc, sizeof(MmaScalars) * B);
    if (e != cudaSuccess) return fail(GTOP_E_ALLOC, std::string("solver workspace: ") + cudaGetErrorString(e));
    pr->s_cap = B;
  }
  return GTOP_OK;
}

int gtop_solve_batch_device(const gtop_sdf* sdf, const gtop_problems* cpr, double* X, int64_t B, int64_t off,
                            int64_t starts, int step, int maxeval, double* minf, void* stream) {
  if (int rc = check_eval(sdf, cpr, B, off, starts, step)) return rc;
  REQUIRE(maxeval >= 1, "maxeval must be >= 1");
  if (B == 0) return GTOP_OK;
  REQUIRE(X, "null device buffer");
  gtop_problems* pr = const_cast<gtop_problems*>(cpr);  // the workspace is a cache
  CU(cudaSetDevice(pr->device));
  if (int rc = solve_workspace(pr, B)) return rc;
  const int n = pr->layout.n;
  cudaStream_t st = stream ? (cudaStream_t)stream : pr->streams[0];
  SolveView V;
  V.n = n, V.B = B, V.traj_offset = off, V.starts_per_problem = starts, V.num_problems = pr->np;
  V.lb = pr->s_lb, V.ub = pr->s_ub, V.xcur = pr->s_xcur, V.grad = pr->s_grad, V.cost = pr->s_cost;
  V.x = pr->s_state, V.dfdx = pr->s_state + (size_t)B * n, V.sigma = pr->s_state + 2 * (size_t)B * n;
  V.xprev = pr->s_state + 3 * (size_t)B * n, V.xprevprev = pr->s_state + 4 * (size_t)B * n, V.sc = pr->s_sc;
  CU(cudaMemcpyAsync(V.xcur, X, sizeof(double) * B * n, cudaMemcpyDeviceToDevice, st));
  CU(launch_mma_init(V, st));
  const EvalScalars E = make_scalars(pr->prm, step);
  for (int e = 0; e < maxeval; ++e) {
    CU(launch_eval(sdf->view, pr->layout, pr->rec, pr->np, E, V.xcur, B, off, starts, pr->s_cost, pr->s_grad, pr->oob, 0, st));
    CU(launch_mma_step(V, e + 1 < maxeval ? 1 : 0, st));
    g_launches += 2;
  }
  CU(launch_mma_finish(V, X, minf, st));
  g_launches += 2;
  return GTOP_OK;
}

int gtop_solve_batch_host(const gtop_sdf* sdf, const gtop_problems* cpr, double* X, int64_t B, int64_t off, int64_t starts,
                          int step, int maxeval, double* minf) {
  if (int rc = check_eval(sdf, cpr, B, off, starts, step)) return rc;
  if (B == 0) return GTOP_OK;
  REQUIRE(X, "null host buffer");
  gtop_problems* pr = const_cast<gtop_problems*>(cpr);
  CU(cudaSetDevice(pr->device));
  const int n = pr->layout.n;
  // x lives on the device for the whole loop: it goes in and comes back once, through the staging buffers the handle keeps
  // for the *_host entry points (pinned host buffers copy at the link's rate; nothing is allocated per call)
  if (pr->cap < B) {
    cudaFree(pr->dX), cudaFree(pr->dcost), cudaFree(pr->dgrad);
    pr->dX = pr->dcost = pr->dgrad = nullptr, pr->cap = 0;
    cudaError_t ea = cudaMalloc(&pr->dX, sizeof(double) * B * n);
    if (ea == cudaSuccess) ea = cudaMalloc(&pr->dgrad, sizeof(double) * B * n);
    if (ea == cudaSuccess) ea = cudaMalloc(&pr->dcost, sizeof(double) * B);
    if (ea != cudaSuccess) return fail(GTOP_E_ALLOC, std::string("solve staging: ") + cudaGetErrorString(ea));
    pr->cap = B;
  }
  double *dX = pr->dX, *dmin = pr->dcost;
  cudaStream_t st = pr->streams[0];
  cudaError_t e = cudaMemcpyAsync(dX, X, sizeof(double) * B * n, cudaMemcpyHostToDevice, st);
  int rc = GTOP_OK;
  if (e == cudaSuccess) rc = gtop_solve_batch_device(sdf, cpr, dX, B, off, starts, step, maxeval, dmin, st);
  if (e == cudaSuccess && rc == GTOP_OK) e = cudaMemcpyAsync(X, dX, sizeof(double) * B * n, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && rc == GTOP_OK && minf) e = cudaMemcpyAsync(minf, dmin, sizeof(double) * B, cudaMemcpyDeviceToHost, st);
  const cudaError_t es = cudaStreamSynchronize(st);  // whatever happened, nothing still writes into the caller's buffers
  if (e == cudaSuccess) e = es;
  if (rc != GTOP_OK) return rc;
  CU(e);
  return GTOP_OK;
}

// ================================================================================ NLopt callback
int gtop_objective_ctx_create(const gtop_sdf* sdf, const gtop_problems* pr, int32_t problem, int step,
                              gtop_objective_ctx** out) {
  if (int rc = check_eval(sdf, pr, 1, 0, 1, step)) return rc;
  REQUIRE(out && problem >= 0 && problem < pr->np, "bad argument");
  CU(cudaSetDevice(pr->device));
  gtop_objective_ctx* c = new gtop_objective_ctx();
  c->sdf = sdf, c->problems = pr, c->problem = problem, c->step = step, c->n = pr->layout.n;
  cudaError_t e = cudaHostAlloc(&c->hx, sizeof(double) * c->n, cudaHostAllocMapped);
  if (e == cudaSuccess) e = cudaHostAlloc(&c->hout, sizeof(double) * (c->n + 1), cudaHostAllocMapped);
  if (e == cudaSuccess) e = cudaHostGetDevicePointer(&c->hx_dev, c->hx, 0);
  if (e == cudaSuccess) e = cudaHostGetDevicePointer(&c->hout_dev, c->hout, 0);
  if (e == cudaSuccess) e = cudaMalloc(&c->dx, sizeof(double) * c->n);
  if (e == cudaSuccess) e = cudaMalloc(&c->dout, sizeof(double) * (c->n + 1));
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    gtop_objective_ctx_destroy(c);
    return fail(GTOP_E_ALLOC, std::string("objective ctx: ") + cudaGetErrorString(e));
  }
  *out = c;
  return GTOP_OK;
}

int gtop_objective_ctx_set_step(gtop_objective_ctx* c, int step) {
  REQUIRE(c && step >= 0 && step <= 2, "bad argument");
  c->step = step;
  return GTOP_OK;
}

int gtop_objective_ctx_destroy(gtop_objective_ctx* c) {
  if (!c) return GTOP_OK;
  cudaSetDevice(c->problems->device);
  cudaFreeHost(c->hx), cudaFreeHost(c->hout), cudaFree(c->dx), cudaFree(c->dout);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return GTOP_OK;
}

int64_t gtop_objective_ctx_evals(const gtop_objective_ctx* c) { return c ? c->evals : 0; }

double gtop_objective(unsigned n, const double* x, double* grad, void* vctx) {
  gtop_objective_ctx* c = (gtop_objective_ctx*)vctx;
  if (!c || !x || (int)n != c->n) {
    g_err = "gtop_objective: bad context or dimension";
    return HUGE_VAL;
  }
  const gtop_problems* pr = c->problems;
  cudaError_t e = cudaSetDevice(pr->device);
  // zero-copy: x is read from, and cost/gradient are written to, pinned host memory that the device addresses
  // directly (unified addressing) — one launch and one synchronisation per call, no copy commands
  std::memcpy(c->hx, x, sizeof(double) * n);
  if (e == cudaSuccess) {
    // one trajectory, latency matters: CTA per trajectory, warp per segment (eval_one_kernel.cu) when m <= 16
    const EvalScalars E = make_scalars(pr->prm, c->step);
    if (eval_one_applicable(pr->layout))
      e = launch_eval_one(c->sdf->view, pr->layout, pr->rec, pr->np, E, c->hx_dev, 1, c->problem, 1, c->hout_dev,
                          c->hout_dev + 1, pr->oob, c->stream);
    else
      e = launch_eval(c->sdf->view, pr->layout, pr->rec, pr->np, E, c->hx_dev, 1, c->problem, 1, c->hout_dev,
                      c->hout_dev + 1, pr->oob, 0, c->stream);
  }
  ++g_launches;
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  if (e != cudaSuccess) {
    g_err = std::string("gtop_objective: ") + cudaGetErrorString(e);
    return HUGE_VAL;
  }
  ++c->evals;
  if (grad) std::memcpy(grad, c->hout + 1, sizeof(double) * n);
  return c->hout[0];
}

}  // extern "C"
