// gtop_oracle.hpp — CPU ORACLE for the GTOP trajectory-optimization inner loop.
//
// *** TEST INFRASTRUCTURE, NOT PRODUCT CODE. ***
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library.  The product (grad_traj_optimization_b200/, include/) never links, loads or
// calls anything in oracle/.
//
// PARITY UNPINNED: the reference (EpicOne1/grad_traj_optimization) ships no tests, no golden
// vectors and cannot be compiled in this image (needs Eigen3, ROS, nlopt.hpp — all absent), so
// this restatement is pinned only by (a) the survey-computed sanity numbers of SURVEY.md §10 and
// (b) its own self-checks (brute-force EDT, finite differences of the exact-formula variant).
//
// What it restates (paths relative to /root/reference):
//   src/qp_generator.cpp:156-236,317-355   PolyQPGeneration (type 2, straight-line init)
//   src/qp_generator.cpp:23-154            PolyKinoGeneration
//   src/qp_generator.cpp:357-405           StackOptiDep  (C, L = A^-1 C^T, R and its blocks)
//   src/qp_generator.cpp:407-451           getInitialD
//   src/sdf_map.cpp:3-99                   SDFMap ctor, resetBuffer, isInMap, posToIndex, setOccupancy
//   src/sdf_map.cpp:155-242                getDistance, getDistWithGradTrilinear
//   src/sdf_map.cpp:266-368                fillESDF, updateESDF3d
//   src/grad_traj_optimizer.cpp:35-110     setKinoPath, setPath
//   src/grad_traj_optimizer.cpp:128-243    optimizeTrajectory (NLopt driver)
//   src/grad_traj_optimizer.cpp:253-279    getCoefficientFromDerivative
//   src/grad_traj_optimizer.cpp:281-448    getCostAndGradient
//   src/grad_traj_optimizer.cpp:451-562    evaluators, penalties, costFunc
// Rulings on the reference's undefined / quirky behaviour follow SURVEY.md §9 (H1..H15) and are
// repeated next to the code they affect.
//
// Build: oracle/Makefile, -O2 -ffp-contract=off, no -march=native (the reference build has no FMA
// contraction: CMakeLists.txt:5-7).
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

namespace gtop_oracle {

// ---------------------------------------------------------------------------------------------
// Tiny dense column-major matrix (stands in for Eigen::MatrixXd, which is column-major).
// ---------------------------------------------------------------------------------------------
struct Mat {
  int rows = 0, cols = 0;
  std::vector<double> a;
  Mat() {}
  Mat(int r, int c) : rows(r), cols(c), a((size_t)r * c, 0.0) {}
  double& operator()(int i, int j) { return a[(size_t)j * rows + i]; }
  double operator()(int i, int j) const { return a[(size_t)j * rows + i]; }
};
Mat matmul(const Mat& A, const Mat& B);
Mat transpose(const Mat& A);
Mat inverse(const Mat& A);  // partial-pivot LU, as Eigen's MatrixXd::inverse() for dynamic sizes
Mat block(const Mat& A, int r0, int c0, int nr, int nc);

// ---------------------------------------------------------------------------------------------
// Parameters: the ROS params of grad_traj_optimizer.cpp:5-32 as a plain struct.
// Layout is mirrored by ctypes in tests/ (all doubles first, then ints).
// ---------------------------------------------------------------------------------------------
struct Params {
  double time_limit_1 = 0.04, time_limit_2 = 0.06, deltat = 0.2;
  double w_smooth = 1.0, w_collision = 1.0;
  double alpha = 1.0, r = 1.0, d0 = 1.0;
  double alpha_v = 0.0, r_v = 1.0, v0 = 1.0;
  double alpha_a = 0.0, r_a = 1.0, a0 = 1.0;
  double bos = 3.0, vos = 8.0, aos = 10.0;
  double mean_v = 1.0, mean_a = 1.0, init_time = 0.0;
  int32_t algorithm = 24;            // NLopt LD_MMA
  int32_t samples_per_segment = 30;  // hard-coded 30.0 at grad_traj_optimizer.cpp:351
  int32_t vel_acc_penalty = 0;       // 0 = HEAD behaviour (block :383-407 commented out)
  int32_t exact_gradient = 0;        // 1 = test-only variant: true gradient of the cost (H4)
};

// ---------------------------------------------------------------------------------------------
// SDFMap restatement.
// ---------------------------------------------------------------------------------------------
class SDFMap {
 public:
  SDFMap(const double origin[3], double resolution, const double map_size[3]);
  void resetBuffer();
  // windowed forms (sdf_map.cpp:28-53, :244-264): reset / restrict the next updateESDF3d to a box given in metres
  void resetBuffer(const double min_pos[3], const double max_pos[3]);
  void setUpdateRange(const double min_pos[3], const double max_pos[3]);
  void setOccupancy(const double pos[3], int occ = 1);
  void updateESDF3d();
  // brute-force O(N * N_occ) unsigned EDT, same output convention; self-check only.
  void bruteForceEDT(std::vector<double>& out) const;
  double getDistance(int x, int y, int z) const;  // clamped
  double getDistance(const double pos[3]) const;  // :155-164: -1 outside the map, else the voxel's value
  // returns dist; grad set to 0 and *in_map=false when outside (H3 ruling)
  double getDistWithGradTrilinear(const double pos[3], double grad[3], bool* in_map) const;
  bool isInMap(const double pos[3]) const;

  int grid_size[3];
  double origin[3], map_size[3], min_range[3], max_range[3];
  double resolution, resolution_inv;
  int min_vec[3], max_vec[3];
  std::vector<int> occupancy_buffer;
  std::vector<double> distance_buffer, tmp_buffer1, tmp_buffer2;

 public:
  void posToIndex(const double pos[3], int id[3]) const;

 private:
  template <typename FG, typename FS>
  void fillESDF(FG f_get_val, FS f_set_val, int start, int end, int dim);
};

// ---------------------------------------------------------------------------------------------
// TrajectoryGenerator restatement (globals m/startVel/startAcc of qp_generator.cpp:12-16 become
// members — H14).
// ---------------------------------------------------------------------------------------------
class TrajectoryGenerator {
 public:
  Mat PolyQPGeneration(const Mat& Path, const double Vel[3], const double Acc[3],
                       const std::vector<double>& Time, int type);
  Mat PolyKinoGeneration(const Mat& Pos, const Mat& Vel, const Mat& Acc,
                         const std::vector<double>& Time);
  void StackOptiDep();
  void getInitialD(Mat& Dp, Mat& Df) const;
  Mat _A, _Q, _C, _L, _R, _Rff, _Rpp, _Rpf, _Rfp;
  std::vector<double> _Dx, _Dy, _Dz;
  int m = 3;
  double startVel[3] = {0, 0, 0}, startAcc[3] = {0, 0, 0};
};

// ---------------------------------------------------------------------------------------------
// GradTrajOptimizer restatement.
// ---------------------------------------------------------------------------------------------
typedef double (*nlopt_func_t)(unsigned n, const double* x, double* grad, void* data);

struct SolveOptions {
  int maxeval = 0;         // >0: deterministic stop (H7); the reference has none
  double maxtime = 0.0;    // >0: wall-clock stop as in the reference (:144-148)
  double xtol_abs = 0.0;   // >0: the commented-out :141
};

class GradTrajOptimizer {
 public:
  explicit GradTrajOptimizer(const Params& p) : prm(p) {}
  void setMap(const SDFMap* map) { sdf_map = map; }
  void setPath(const std::vector<double>& way_points_xyz);  // npts*3, row-major points
  void setKinoPath(const Mat& Pos, const Mat& Vel, const Mat& Acc, const std::vector<double>& Time);
  // NLopt through its C ABI, loaded with dlopen from nlopt_path (NLopt 2.5.0, vendored binary).
  int optimizeTrajectory(int step, const SolveOptions& opt, const char* nlopt_path,
                         nlopt_func_t override_f = nullptr, void* override_data = nullptr);
  void getCoefficientFromDerivative(Mat& coefficient, const std::vector<double>& dp) const;
  void getCostAndGradient(const std::vector<double>& dp, double& cost, std::vector<double>& grad,
                          int64_t* out_of_map = nullptr, double* terms = nullptr) const;
  void bounds(std::vector<double>& lb, std::vector<double>& ub) const;
  std::vector<double> x0() const;  // Dp flattened [x-axis P | y-axis P | z-axis P]

  Params prm;
  const SDFMap* sdf_map = nullptr;
  Mat coeff, A, C, L, R, Rff, Rpp, Rpf, Rfp, V, Df, Dp, path;
  std::vector<double> segment_time;
  int num_dp = 0, num_df = 0, num_point = 0;
  int step = 1;
  mutable int iter_num = 0;
  double last_min_f = 0.0;
  int last_result = 0;
  std::vector<double> vec_cost;  // monotone best-so-far curve (:438-447); times omitted
  static double costFunc(unsigned n, const double* x, double* grad, void* data);

 private:
  void finishSetup(TrajectoryGenerator& gen);
};

// ---------------------------------------------------------------------------------------------
// PolynomialTraj restatement (include/grad_traj_optimization/polynomial_traj.hpp): the post-hoc evaluation
// opti_node runs on the optimised trajectory (src/opti_node.cpp:108-144).  Coefficients are stored in
// DESCENDING powers per segment (opti_node reverses the rows of getCoefficient, :117-119).
// Reference quirks reproduced (H16): getMeanAndMaxVel / getMeanAndMaxAcc evaluate their polynomials at the
// segment DURATION ts for every step of eval_t (pow(ts, i), :170, :202), i.e. the end velocity / acceleration of
// the segment weighted by its number of 0.01 s steps; evaluate() would index past the last segment for
// t == time_sum (:47-50) — clamped here.
// ---------------------------------------------------------------------------------------------
class PolynomialTraj {
 public:
  void reset();
  void addSegment(const std::vector<double>& cx, const std::vector<double>& cy, const std::vector<double>& cz, double t);
  void init();
  void evaluate(double t, double pt[3]) const;                  // :45-65
  double getTimeSum() const { return time_sum; }
  const std::vector<double>& getTraj();                         // :70-79, xyz triples every 0.01 s
  double getLength();                                           // :81-91 (needs getTraj first)
  double getAccCost() const;                                    // :95-107
  double getJerk() const;                                       // :109-139
  void getMeanAndMaxVel(double& mean_v, double& max_v) const;   // :141-171
  void getMeanAndMaxAcc(double& mean_a, double& max_a) const;   // :173-203

  std::vector<double> times;
  std::vector<std::vector<double>> cxs, cys, czs;
  double time_sum = 0.0, length = 0.0;
  int num_seg = 0;
  std::vector<double> traj_vec3d;
};

// ---------------------------------------------------------------------------------------------
// EDTEnvironment restatement (src/edt_environment.cpp, include/.../edt_environment.h; moving boxes from
// obj_predictor.h:57-66): the static ESDF min'ed, at the eight corners of the trilinear cell, with the distance to
// constant-velocity boxes at time `time` (:75-122).  PARITY UNPINNED and the callee is absent: the file is not
// compiled by the reference's CMakeLists (:65-73) and calls SDFMap::getInterpolationData, which the reference's
// sdf_map.h does not declare; it is restated here from its sibling getDistWithGradTrilinear (sdf_map.cpp:196-209:
// pos_m = pos - res/2, idx = posToIndex(pos_m), diff = (pos - indexToPos(idx)) * res_inv, corner positions
// indexToPos(idx + {0,1}^3)).  Corner values go through getDistance(Vector3d) (:155-164): -1 outside the map (NOT
// clamped, unlike getDistWithGradTrilinear).
// ---------------------------------------------------------------------------------------------
struct MovingBox {
  double c0[3], c1[3];  // PolynomialPrediction::evaluateConstVel: centre(t) = c0 + t * c1 (obj_predictor.h:57-66)
  double scale[3];      // ObjScale: full edge lengths
};

class EDTEnvironment {
 public:
  void setMap(const SDFMap* map) { sdf_map_ = map, resolution_inv_ = 1 / map->resolution; }  // :11-15
  void setObjects(const std::vector<MovingBox>& boxes) { boxes_ = boxes; }                    // :17-25
  double distToBox(int idx, const double pos[3], double time) const;                          // :27-44
  double minDistToAllBox(const double pos[3], double time) const;                             // :61-73
  void evaluateEDTWithGrad(const double pos[3], double time, double& dist, double grad[3]) const;  // :75-122
  double evaluateCoarseEDT(const double pos[3], double time) const;                           // :124-137

 private:
  const SDFMap* sdf_map_ = nullptr;
  std::vector<MovingBox> boxes_;
  double resolution_inv_ = 0.0;
};

}  // namespace gtop_oracle
