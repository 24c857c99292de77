// gtop_oracle.cpp — CPU ORACLE (test infrastructure; see gtop_oracle.hpp for scope and rulings).
// PARITY UNPINNED by the reference (no tests / golden vectors upstream); pinned by SURVEY.md §10
// sanity numbers and self-checks only.
#include "gtop_oracle.hpp"

#include <dlfcn.h>

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <stdexcept>

namespace gtop_oracle {

// =============================================================================================
// dense helpers
// =============================================================================================
Mat matmul(const Mat& A, const Mat& B) {
  if (A.cols != B.rows) throw std::runtime_error("matmul: shape mismatch");
  Mat Cm(A.rows, B.cols);
  for (int j = 0; j < B.cols; ++j)
    for (int k = 0; k < A.cols; ++k) {
      const double b = B(k, j);
      if (b == 0.0) continue;  // adding exact zeros does not change any partial sum
      for (int i = 0; i < A.rows; ++i) Cm(i, j) += A(i, k) * b;
    }
  return Cm;
}

Mat transpose(const Mat& A) {
  Mat T(A.cols, A.rows);
  for (int i = 0; i < A.rows; ++i)
    for (int j = 0; j < A.cols; ++j) T(j, i) = A(i, j);
  return T;
}

Mat block(const Mat& A, int r0, int c0, int nr, int nc) {
  Mat Bm(nr, nc);
  for (int j = 0; j < nc; ++j)
    for (int i = 0; i < nr; ++i) Bm(i, j) = A(r0 + i, c0 + j);
  return Bm;
}

// Eigen's MatrixXd::inverse() for dynamic sizes goes through PartialPivLU; restated as an
// unblocked partial-pivot LU followed by a solve against the identity.
Mat inverse(const Mat& A) {
  if (A.rows != A.cols) throw std::runtime_error("inverse: not square");
  const int n = A.rows;
  Mat LU = A;
  std::vector<int> perm(n);
  for (int i = 0; i < n; ++i) perm[i] = i;
  for (int k = 0; k < n; ++k) {
    int piv = k;
    double best = std::fabs(LU(k, k));
    for (int i = k + 1; i < n; ++i) {
      double v = std::fabs(LU(i, k));
      if (v > best) best = v, piv = i;
    }
    if (best == 0.0) throw std::runtime_error("inverse: singular matrix");
    if (piv != k) {
      for (int j = 0; j < n; ++j) std::swap(LU(k, j), LU(piv, j));
      std::swap(perm[k], perm[piv]);
    }
    const double d = LU(k, k);
    for (int i = k + 1; i < n; ++i) {
      if (LU(i, k) == 0.0) continue;
      LU(i, k) /= d;
      const double l = LU(i, k);
      for (int j = k + 1; j < n; ++j) LU(i, j) -= l * LU(k, j);
    }
  }
  Mat Inv(n, n);
  std::vector<double> y(n);
  for (int c = 0; c < n; ++c) {
    // solve L U x = P e_c
    for (int i = 0; i < n; ++i) {
      double s = (perm[i] == c) ? 1.0 : 0.0;
      for (int j = 0; j < i; ++j) s -= LU(i, j) * y[j];
      y[i] = s;
    }
    for (int i = n - 1; i >= 0; --i) {
      double s = y[i];
      for (int j = i + 1; j < n; ++j) s -= LU(i, j) * Inv(j, c);
      Inv(i, c) = s / LU(i, i);
    }
  }
  return Inv;
}

// =============================================================================================
// SDFMap  (src/sdf_map.cpp)
// =============================================================================================
SDFMap::SDFMap(const double org[3], double res, const double msize[3]) {  // :3-24
  resolution = res;
  resolution_inv = 1 / res;
  for (int i = 0; i < 3; ++i) {
    origin[i] = org[i];
    map_size[i] = msize[i];
    grid_size[i] = (int)std::ceil(msize[i] / res);
    min_range[i] = org[i];
    max_range[i] = org[i] + msize[i];
    min_vec[i] = 0;
    max_vec[i] = grid_size[i] - 1;
  }
  const size_t n = (size_t)grid_size[0] * grid_size[1] * grid_size[2];
  occupancy_buffer.assign(n, 0);
  distance_buffer.assign(n, 10000.0);
  tmp_buffer1.assign(n, 0.0);
  tmp_buffer2.assign(n, 0.0);
}

void SDFMap::posToIndex(const double pos[3], int id[3]) const {  // :71-74
  for (int i = 0; i < 3; ++i) id[i] = (int)std::floor((pos[i] - origin[i]) * resolution_inv);
}

bool SDFMap::isInMap(const double pos[3]) const {  // :55-69
  if (pos[0] < min_range[0] + 1e-4 || pos[1] < min_range[1] + 1e-4 || pos[2] < min_range[2] + 1e-4)
    return false;
  if (pos[0] > max_range[0] - 1e-4 || pos[1] > max_range[1] - 1e-4 || pos[2] > max_range[2] - 1e-4)
    return false;
  return true;
}

void SDFMap::resetBuffer() { resetBuffer(min_range, max_range); }  // :26

void SDFMap::resetBuffer(const double min_pos_in[3], const double max_pos_in[3]) {  // :28-53
  double min_pos[3], max_pos[3];
  int min_id[3], max_id[3];
  for (int i = 0; i < 3; ++i) {
    min_pos[i] = std::max(min_pos_in[i], min_range[i]);
    max_pos[i] = std::min(max_pos_in[i], max_range[i]) - resolution / 2;
  }
  posToIndex(min_pos, min_id);
  posToIndex(max_pos, max_id);
  for (int x = min_id[0]; x <= max_id[0]; ++x)
    for (int y = min_id[1]; y <= max_id[1]; ++y)
      for (int z = min_id[2]; z <= max_id[2]; ++z) {
        const size_t a = (size_t)x * grid_size[1] * grid_size[2] + (size_t)y * grid_size[2] + z;
        occupancy_buffer[a] = 0;
        distance_buffer[a] = 10000;
      }
}

void SDFMap::setUpdateRange(const double min_pos_in[3], const double max_pos_in[3]) {  // :244-264
  double min_pos[3], max_pos[3];
  for (int i = 0; i < 3; ++i) {
    min_pos[i] = std::max(min_pos_in[i], min_range[i]);
    max_pos[i] = std::min(max_pos_in[i], max_range[i]) - resolution / 2;
  }
  posToIndex(min_pos, min_vec);
  posToIndex(max_pos, max_vec);
}

void SDFMap::setOccupancy(const double pos[3], int occ) {  // :80-99
  if (occ != 1 && occ != 0) return;
  if (!isInMap(pos)) return;  // H13: points within 1e-4 of a face are dropped silently
  int id[3];
  posToIndex(pos, id);
  occupancy_buffer[(size_t)id[0] * grid_size[1] * grid_size[2] + (size_t)id[1] * grid_size[2] + id[2]] = occ;
}

double SDFMap::getDistance(const double pos[3]) const {  // :155-164
  if (!isInMap(pos)) return -1;
  int id[3];
  posToIndex(pos, id);
  return distance_buffer[(size_t)id[0] * grid_size[1] * grid_size[2] + (size_t)id[1] * grid_size[2] + id[2]];
}

double SDFMap::getDistance(int x, int y, int z) const {  // :166-183 (clamped)
  x = std::max(std::min(x, grid_size[0] - 1), 0);
  y = std::max(std::min(y, grid_size[1] - 1), 0);
  z = std::max(std::min(z, grid_size[2] - 1), 0);
  return distance_buffer[(size_t)x * grid_size[1] * grid_size[2] + (size_t)y * grid_size[2] + z];
}

double SDFMap::getDistWithGradTrilinear(const double pos[3], double grad[3], bool* in_map) const {  // :185-242
  if (!isInMap(pos)) {
    // H3: the reference returns -1 and leaves grad uninitialised; ruling: grad = 0.
    grad[0] = grad[1] = grad[2] = 0.0;
    if (in_map) *in_map = false;
    return -1;
  }
  if (in_map) *in_map = true;
  double pos_m[3], idx_pos[3], diff[3];
  int idx[3];
  for (int i = 0; i < 3; ++i) pos_m[i] = pos[i] - 0.5 * resolution * 1.0;
  posToIndex(pos_m, idx);
  for (int i = 0; i < 3; ++i) idx_pos[i] = (idx[i] + 0.5) * resolution + origin[i];  // indexToPos :76-78
  for (int i = 0; i < 3; ++i) diff[i] = (pos[i] - idx_pos[i]) * resolution_inv;

  double values[2][2][2];
  for (int x = 0; x < 2; x++)
    for (int y = 0; y < 2; y++)
      for (int z = 0; z < 2; z++) values[x][y][z] = getDistance(idx[0] + x, idx[1] + y, idx[2] + z);

  double v00 = (1 - diff[0]) * values[0][0][0] + diff[0] * values[1][0][0];
  double v01 = (1 - diff[0]) * values[0][0][1] + diff[0] * values[1][0][1];
  double v10 = (1 - diff[0]) * values[0][1][0] + diff[0] * values[1][1][0];
  double v11 = (1 - diff[0]) * values[0][1][1] + diff[0] * values[1][1][1];
  double v0 = (1 - diff[1]) * v00 + diff[1] * v10;
  double v1 = (1 - diff[1]) * v01 + diff[1] * v11;
  double dist = (1 - diff[2]) * v0 + diff[2] * v1;

  grad[2] = (v1 - v0) * resolution_inv;
  grad[1] = ((1 - diff[2]) * (v10 - v00) + diff[2] * (v11 - v01)) * resolution_inv;
  grad[0] = (1 - diff[2]) * (1 - diff[1]) * (values[1][0][0] - values[0][0][0]);
  grad[0] += (1 - diff[2]) * diff[1] * (values[1][1][0] - values[0][1][0]);
  grad[0] += diff[2] * (1 - diff[1]) * (values[1][0][1] - values[0][0][1]);
  grad[0] += diff[2] * diff[1] * (values[1][1][1] - values[0][1][1]);
  grad[0] *= resolution_inv;
  return dist;
}

// Felzenszwalb–Huttenlocher 1-D lower envelope, :266-308 (VLAs replaced by vectors).
template <typename FG, typename FS>
void SDFMap::fillESDF(FG f_get_val, FS f_set_val, int start, int end, int dim) {
  std::vector<int> v(grid_size[dim]);
  std::vector<double> z(grid_size[dim] + 1);
  int k = start;
  v[start] = start;
  z[start] = -std::numeric_limits<double>::max();
  z[start + 1] = std::numeric_limits<double>::max();
  for (int q = start + 1; q <= end; q++) {
    k++;
    double s;
    do {
      k--;
      s = ((f_get_val(q) + q * q) - (f_get_val(v[k]) + v[k] * v[k])) / (2 * q - 2 * v[k]);
    } while (s <= z[k]);
    k++;
    v[k] = q;
    z[k] = s;
    z[k + 1] = std::numeric_limits<double>::max();
  }
  k = start;
  for (int q = start; q <= end; q++) {
    while (z[k + 1] < q) k++;
    double val = (q - v[k]) * (q - v[k]) + f_get_val(v[k]);
    f_set_val(q, val);
  }
}

void SDFMap::updateESDF3d() {  // :310-368
  const size_t sy = grid_size[2], sx = (size_t)grid_size[1] * grid_size[2];
  for (int x = min_vec[0]; x <= max_vec[0]; x++)
    for (int y = min_vec[1]; y <= max_vec[1]; y++)
      fillESDF(
          [&](int z) {
            return occupancy_buffer[x * sx + y * sy + z] == 1 ? 0 : std::numeric_limits<double>::max();
          },
          [&](int z, double val) { tmp_buffer1[x * sx + y * sy + z] = val; }, min_vec[2], max_vec[2], 2);
  for (int x = min_vec[0]; x <= max_vec[0]; x++)
    for (int z = min_vec[2]; z <= max_vec[2]; z++)
      fillESDF([&](int y) { return tmp_buffer1[x * sx + y * sy + z]; },
               [&](int y, double val) { tmp_buffer2[x * sx + y * sy + z] = val; }, min_vec[1], max_vec[1], 1);
  for (int y = min_vec[1]; y <= max_vec[1]; y++)
    for (int z = min_vec[2]; z <= max_vec[2]; z++)
      fillESDF([&](int x) { return tmp_buffer2[x * sx + y * sy + z]; },
               [&](int x, double val) {
                 double& d = distance_buffer[x * sx + y * sy + z];
                 d = std::min(resolution * std::sqrt(val), d);
               },
               min_vec[0], max_vec[0], 0);
  for (int i = 0; i < 3; ++i) min_vec[i] = 0, max_vec[i] = grid_size[i] - 1;
}

void SDFMap::bruteForceEDT(std::vector<double>& out) const {
  const int nx = grid_size[0], ny = grid_size[1], nz = grid_size[2];
  std::vector<int> ox, oy, oz;
  for (int x = 0; x < nx; ++x)
    for (int y = 0; y < ny; ++y)
      for (int z = 0; z < nz; ++z)
        if (occupancy_buffer[((size_t)x * ny + y) * nz + z] == 1) ox.push_back(x), oy.push_back(y), oz.push_back(z);
  out.assign((size_t)nx * ny * nz, 10000.0);
  if (ox.empty()) return;
  for (int x = 0; x < nx; ++x)
    for (int y = 0; y < ny; ++y)
      for (int z = 0; z < nz; ++z) {
        long best = std::numeric_limits<long>::max();
        for (size_t i = 0; i < ox.size(); ++i) {
          long dx = x - ox[i], dy = y - oy[i], dz = z - oz[i];
          long d2 = dx * dx + dy * dy + dz * dz;
          if (d2 < best) best = d2;
        }
        out[((size_t)x * ny + y) * nz + z] = std::min(resolution * std::sqrt((double)best), 10000.0);
      }
}

// =============================================================================================
// TrajectoryGenerator  (src/qp_generator.cpp)
// =============================================================================================
static int Factorial(int x) {
  int fac = 1;
  for (int i = x; i > 0; i--) fac = fac * i;
  return fac;
}

static Mat buildA(const std::vector<double>& Time) {  // :181-197 / :41-54
  const int m = (int)Time.size();
  Mat A(m * 6, m * 6);
  for (int k = 0; k < m; k++)
    for (int i = 0; i < 3; i++) {
      A(k * 6 + 2 * i, k * 6 + i) = Factorial(i);
      for (int j = i; j < 6; j++)
        A(k * 6 + 2 * i + 1, k * 6 + j) = Factorial(j) / Factorial(j - i) * std::pow(Time[k], j - i);
    }
  return A;
}

static Mat buildQ(const std::vector<double>& Time) {  // :223-236 / :100-112 (min-jerk Hessian)
  const int m = (int)Time.size();
  Mat H(m * 6, m * 6);
  for (int k = 0; k < m; k++)
    for (int i = 3; i < 6; i++)
      for (int j = 3; j < 6; j++)
        H(k * 6 + i, k * 6 + j) =
            i * (i - 1) * (i - 2) * j * (j - 1) * (j - 2) / (i + j - 5) * std::pow(Time[k], (i + j - 5));
  return H;
}

static std::vector<double> matvec(const Mat& A, const std::vector<double>& x) {
  std::vector<double> y(A.rows, 0.0);
  for (int j = 0; j < A.cols; ++j) {
    if (x[j] == 0.0) continue;
    for (int i = 0; i < A.rows; ++i) y[i] += A(i, j) * x[j];
  }
  return y;
}

Mat TrajectoryGenerator::PolyQPGeneration(const Mat& Path, const double Vel[3], const double Acc[3],
                                          const std::vector<double>& Time, int type) {
  if (type != 2) throw std::runtime_error("oracle: only the straight-line init (type 2) used by setPath is restated");
  for (int i = 0; i < 3; ++i) startVel[i] = Vel[i], startAcc[i] = Acc[i];
  m = (int)Time.size();
  _A = buildA(Time);
  _Dx.assign(m * 6, 0.0), _Dy.assign(m * 6, 0.0), _Dz.assign(m * 6, 0.0);
  for (int k = 1; k < m + 1; k++) {  // :204-221
    _Dx[(k - 1) * 6] = Path(k - 1, 0), _Dx[(k - 1) * 6 + 1] = Path(k, 0);
    _Dy[(k - 1) * 6] = Path(k - 1, 1), _Dy[(k - 1) * 6 + 1] = Path(k, 1);
    _Dz[(k - 1) * 6] = Path(k - 1, 2), _Dz[(k - 1) * 6 + 1] = Path(k, 2);
    if (k == 1) {
      _Dx[2] = Vel[0], _Dy[2] = Vel[1], _Dz[2] = Vel[2];
      _Dx[4] = Acc[0], _Dy[4] = Acc[1], _Dz[4] = Acc[2];
    }
  }
  _Q = buildQ(Time);
  Mat Ainv = inverse(_A);  // :334-336
  std::vector<double> Px = matvec(Ainv, _Dx), Py = matvec(Ainv, _Dy), Pz = matvec(Ainv, _Dz);
  Mat PolyCoeff(m, 18);
  for (int i = 0; i < m; i++)
    for (int j = 0; j < 6; ++j)
      PolyCoeff(i, j) = Px[i * 6 + j], PolyCoeff(i, 6 + j) = Py[i * 6 + j], PolyCoeff(i, 12 + j) = Pz[i * 6 + j];
  return PolyCoeff;
}

Mat TrajectoryGenerator::PolyKinoGeneration(const Mat& Pos, const Mat& Vel, const Mat& Acc,
                                            const std::vector<double>& Time) {  // :23-154
  m = (int)Time.size();  // startVel/startAcc deliberately untouched (H14)
  _A = buildA(Time);
  _Dx.assign(m * 6, 0.0), _Dy.assign(m * 6, 0.0), _Dz.assign(m * 6, 0.0);
  for (int k = 0; k < m; k++) {
    _Dx[k * 6] = Pos(k, 0), _Dx[k * 6 + 1] = Pos(k + 1, 0);
    _Dy[k * 6] = Pos(k, 1), _Dy[k * 6 + 1] = Pos(k + 1, 1);
    _Dz[k * 6] = Pos(k, 2), _Dz[k * 6 + 1] = Pos(k + 1, 2);
    _Dx[k * 6 + 2] = Vel(k, 0), _Dx[k * 6 + 3] = Vel(k + 1, 0);
    _Dy[k * 6 + 2] = Vel(k, 1), _Dy[k * 6 + 3] = Vel(k + 1, 1);
    _Dz[k * 6 + 2] = Vel(k, 2), _Dz[k * 6 + 3] = Vel(k + 1, 2);
    _Dx[k * 6 + 4] = Acc(k, 0), _Dx[k * 6 + 5] = Acc(k + 1, 0);
    _Dy[k * 6 + 4] = Acc(k, 1), _Dy[k * 6 + 5] = Acc(k + 1, 1);
    _Dz[k * 6 + 4] = Acc(k, 2), _Dz[k * 6 + 5] = Acc(k + 1, 2);
  }
  _Q = buildQ(Time);
  Mat Ainv = inverse(_A);
  std::vector<double> Px = matvec(Ainv, _Dx), Py = matvec(Ainv, _Dy), Pz = matvec(Ainv, _Dz);
  Mat PolyCoeff(m, 18);
  for (int i = 0; i < m; i++)
    for (int j = 0; j < 6; ++j)
      PolyCoeff(i, j) = Px[i * 6 + j], PolyCoeff(i, 6 + j) = Py[i * 6 + j], PolyCoeff(i, 12 + j) = Pz[i * 6 + j];
  return PolyCoeff;
}

void TrajectoryGenerator::StackOptiDep() {  // :357-405
  if (m < 2) throw std::runtime_error("oracle: StackOptiDep needs >= 2 segments (the reference indexes out of range)");
  const int num_f = 6, num_p = 3 * m - 3, num_d = 6 * m;
  Mat Ct(num_d, num_f + num_p);
  Ct(0, 0) = 1, Ct(2, 1) = 1, Ct(4, 2) = 1;
  Ct(1, 6) = 1, Ct(3, 7) = 1, Ct(5, 8) = 1;
  Ct(6 * (m - 1) + 0, 3 * m + 0) = 1, Ct(6 * (m - 1) + 2, 3 * m + 1) = 1, Ct(6 * (m - 1) + 4, 3 * m + 2) = 1;
  Ct(6 * (m - 1) + 1, 3) = 1, Ct(6 * (m - 1) + 3, 4) = 1, Ct(6 * (m - 1) + 5, 5) = 1;
  for (int j = 2; j < m; j++) {
    Ct(6 * (j - 1) + 0, 6 + 3 * (j - 2) + 0) = 1;
    Ct(6 * (j - 1) + 1, 6 + 3 * (j - 1) + 0) = 1;
    Ct(6 * (j - 1) + 2, 6 + 3 * (j - 2) + 1) = 1;
    Ct(6 * (j - 1) + 3, 6 + 3 * (j - 1) + 1) = 1;
    Ct(6 * (j - 1) + 4, 6 + 3 * (j - 2) + 2) = 1;
    Ct(6 * (j - 1) + 5, 6 + 3 * (j - 1) + 2) = 1;
  }
  _C = transpose(Ct);
  Mat Ainv = inverse(_A);
  _L = matmul(Ainv, Ct);
  // _R = _C * B.transpose() * _Q * (_A.inverse()) * Ct, evaluated left to right (:392-393)
  _R = matmul(matmul(matmul(matmul(_C, transpose(Ainv)), _Q), Ainv), Ct);
  _Rff = block(_R, 0, 0, 6, 6);
  _Rfp = block(_R, 0, 6, 6, num_p);
  _Rpf = block(_R, 6, 0, num_p, 6);
  _Rpp = block(_R, 6, 6, num_p, num_p);
}

void TrajectoryGenerator::getInitialD(Mat& Dp, Mat& Df) const {  // :407-451
  Dp = Mat(3, 3 * m - 3);
  Df = Mat(3, 6);
  const std::vector<double>* D[3] = {&_Dx, &_Dy, &_Dz};
  for (int a = 0; a < 3; ++a) {
    const std::vector<double>& Da = *D[a];
    Df(a, 0) = Da[0];
    Df(a, 3) = Da[Da.size() - 5];
    Df(a, 1) = startVel[a];
    Df(a, 2) = startAcc[a];
    for (int k = 1; k < m; k++)
      for (int i = 0; i < 3; i++) Dp(a, (k - 1) * 3 + i) = Da[(k - 1) * 6 + 2 * i + 1];
  }
}

// =============================================================================================
// GradTrajOptimizer  (src/grad_traj_optimizer.cpp)
// =============================================================================================
void GradTrajOptimizer::finishSetup(TrajectoryGenerator& gen) {
  gen.StackOptiDep();
  R = gen._R, Rff = gen._Rff, Rpp = gen._Rpp, Rpf = gen._Rpf, Rfp = gen._Rfp;
  L = gen._L, A = gen._A, C = gen._C;
  gen.getInitialD(Dp, Df);
  // H2: the reference never zeroes V (only the super-diagonal is written, :104-105); ruling: zeros.
  V = Mat(6, 6);
  for (int i = 0; i < 5; ++i) V(i, i + 1) = i + 1;
  num_dp = Dp.cols;
  num_df = Df.cols;
  num_point = (int)segment_time.size() + 1;
}

void GradTrajOptimizer::setPath(const std::vector<double>& wp) {  // :67-110
  const int npts = (int)(wp.size() / 3);
  path = Mat(npts, 3);
  for (int i = 0; i < npts; ++i)
    for (int a = 0; a < 3; ++a) path(i, a) = wp[3 * i + a];
  segment_time.assign(npts - 1, 0.0);
  for (int i = 0; i < (int)segment_time.size(); ++i) {
    double dx = path(i, 0) - path(i + 1, 0), dy = path(i, 1) - path(i + 1, 1), dz = path(i, 2) - path(i + 1, 2);
    double len = std::sqrt(dx * dx + dy * dy + dz * dz);
    if (i == 0 || i == (int)segment_time.size())  // H9: second test is unreachable
      segment_time[i] = len / prm.mean_v + prm.init_time;
    else
      segment_time[i] = len / prm.mean_v;
  }
  double vel[3] = {0, 0, 0}, acc[3] = {0, 0, 0};
  TrajectoryGenerator gen;
  coeff = gen.PolyQPGeneration(path, vel, acc, segment_time, 2);
  finishSetup(gen);
}

void GradTrajOptimizer::setKinoPath(const Mat& Pos, const Mat& Vel, const Mat& Acc,
                                    const std::vector<double>& Time) {  // :35-65
  path = Pos;
  segment_time = Time;
  TrajectoryGenerator gen;
  coeff = gen.PolyKinoGeneration(Pos, Vel, Acc, Time);
  finishSetup(gen);
}

std::vector<double> GradTrajOptimizer::x0() const {  // :182-187
  std::vector<double> x(3 * num_dp);
  for (int i = 0; i < num_dp; ++i) x[i] = Dp(0, i), x[i + num_dp] = Dp(1, i), x[i + 2 * num_dp] = Dp(2, i);
  return x;
}

void GradTrajOptimizer::bounds(std::vector<double>& lb, std::vector<double>& ub) const {  // :151-177
  lb.assign(3 * num_dp, 0.0), ub.assign(3 * num_dp, 0.0);
  for (int i = 0; i < num_dp; ++i) {
    if (i % 3 == 0) {
      for (int a = 0; a < 3; ++a) {
        lb[i + a * num_dp] = path(i / 3 + 1, a) - prm.bos;
        ub[i + a * num_dp] = path(i / 3 + 1, a) + prm.bos;
      }
    } else if (i % 3 == 1) {
      for (int a = 0; a < 3; ++a) lb[i + a * num_dp] = -prm.vos, ub[i + a * num_dp] = prm.vos;
    } else {
      for (int a = 0; a < 3; ++a) lb[i + a * num_dp] = -prm.aos, ub[i + a * num_dp] = prm.aos;
    }
  }
}

void GradTrajOptimizer::getCoefficientFromDerivative(Mat& coefficient, const std::vector<double>& _dp) const {  // :253-279
  coefficient = Mat(num_point - 1, 18);
  std::vector<double> d(num_df + num_dp);
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 6; ++j) d[j] = Df(i, j);
    for (int j = 0; j < num_dp; j++) d[6 + j] = _dp[j + num_dp * i];
    // coe = L * d : row sums taken over columns in ascending order (the nonzeros of a row of L
    // are a 6-entry block, so this is a 6-term unfused sum; zeros contribute exactly nothing)
    for (int r = 0; r < L.rows; ++r) {
      double s = 0.0;
      for (int c = 0; c < L.cols; ++c) {
        const double l = L(r, c);
        if (l != 0.0) s += l * d[c];
      }
      coefficient(r / 6, 6 * i + r % 6) = s;
    }
  }
}

// getPositionFromCoeff / getVelocityFromCoeff / getAccelerationFromCoeff (:451-505): computed in
// double, stored in float (H1), widened back.
static inline void posFromCoeff(double pos[3], const Mat& c, int s, double t, bool no_float) {
  for (int a = 0; a < 3; ++a) {
    const int o = 6 * a;
    double x = c(s, o + 0) + c(s, o + 1) * t + c(s, o + 2) * std::pow(t, 2) + c(s, o + 3) * std::pow(t, 3) +
               c(s, o + 4) * std::pow(t, 4) + c(s, o + 5) * std::pow(t, 5);
    pos[a] = no_float ? x : (double)(float)x;
  }
}
static inline void velFromCoeff(double vel[3], const Mat& c, int s, double t, bool no_float) {
  for (int a = 0; a < 3; ++a) {
    const int o = 6 * a;
    double v = c(s, o + 1) + 2 * c(s, o + 2) * std::pow(t, 1) + 3 * c(s, o + 3) * std::pow(t, 2) +
               4 * c(s, o + 4) * std::pow(t, 3) + 5 * c(s, o + 5) * std::pow(t, 4);
    vel[a] = no_float ? v : (double)(float)v;
  }
}
static inline void accFromCoeff(double acc[3], const Mat& c, int s, double t, bool no_float) {
  for (int a = 0; a < 3; ++a) {
    const int o = 6 * a;
    double v = 2 * c(s, o + 2) + 6 * c(s, o + 3) * std::pow(t, 1) + 12 * c(s, o + 4) * std::pow(t, 2) +
               20 * c(s, o + 5) * std::pow(t, 3);
    acc[a] = no_float ? v : (double)(float)v;
  }
}

// row(1xP) += s * (T(1x6) * M(6xP))   — dense, as the reference's per-sample Eigen products
static inline void addScaledRowProduct(double* row, double s, const double T6[6], const double* M /*6xP colmajor*/,
                                       int P) {
  double sT[6];
  for (int i = 0; i < 6; ++i) sT[i] = s * T6[i];
  for (int j = 0; j < P; ++j) {
    const double* col = M + 6 * (size_t)j;
    double acc = 0.0;
    for (int i = 0; i < 6; ++i) acc += sT[i] * col[i];
    row[j] = acc;
  }
}

void GradTrajOptimizer::getCostAndGradient(const std::vector<double>& dp, double& cost, std::vector<double>& _grad,
                                           int64_t* out_of_map, double* terms) const {  // :281-448
  const int P = num_dp, ND = num_dp + num_df;
  double cost_smooth = 0, cost_colli = 0, cost_vel = 0, cost_acc = 0;
  std::vector<double> g_smooth(3 * (size_t)P, 0.0), g_colli(3 * (size_t)P, 0.0), g_vel(3 * (size_t)P, 0.0),
      g_acc(3 * (size_t)P, 0.0);
  const bool exact = prm.exact_gradient != 0;

  // smoothness: fs = d' R d per axis (:326-327); gradient 2 Rfp' df + 2 Rpp dp (:330-332, H11)
  std::vector<double> d(ND), tmp(ND);
  for (int a = 0; a < 3; ++a) {
    for (int j = 0; j < 6; ++j) d[j] = Df(a, j);
    for (int j = 0; j < P; ++j) d[6 + j] = dp[j + (size_t)P * a];
    for (int j = 0; j < ND; ++j) {  // tmp = d' R
      double s = 0.0;
      for (int i = 0; i < ND; ++i) s += d[i] * R(i, j);
      tmp[j] = s;
    }
    double q = 0.0;
    for (int j = 0; j < ND; ++j) q += tmp[j] * d[j];
    cost_smooth += q;
    for (int i = 0; i < P; ++i) {
      double s1 = 0.0, s2 = 0.0;
      for (int f = 0; f < 6; ++f) s1 += Rfp(f, i) * d[f];
      for (int j = 0; j < P; ++j) s2 += Rpp(i, j) * d[6 + j];
      g_smooth[(size_t)a * P + i] = 2 * s1 + 2 * s2;
    }
  }

  Mat coe;
  getCoefficientFromDerivative(coe, dp);

  const int NS = prm.samples_per_segment;
  std::vector<double> Ldp(6 * (size_t)P), VLdp(6 * (size_t)P), VVLdp(6 * (size_t)P), r1(P), r2(P);
  int64_t oob = 0;
  for (int s = 0; s < (int)segment_time.size(); s++) {
    if (std::fabs(prm.w_collision) < 1e-4) break;  // :346
    for (int j = 0; j < P; ++j)
      for (int i = 0; i < 6; ++i) Ldp[6 * (size_t)j + i] = L(6 * s + i, 6 + j);  // L.block(6s, 6, 6, P)
    // T*V*Ldp is evaluated per sample as (T*V)*Ldp in the reference; V*Ldp is only pre-formed here
    // for the optional vel/acc block's T*V*V*Ldp.
    const double dt = segment_time[s] / (double)NS;  // :351 (30.0 hard-coded upstream)
    for (double t = 1e-3; t < segment_time[s]; t += dt) {  // :353, H8
      double pos[3], vel[3];
      posFromCoeff(pos, coe, s, t, exact);
      velFromCoeff(vel, coe, s, t, exact);
      const double vnorm_raw = std::sqrt(vel[0] * vel[0] + vel[1] * vel[1] + vel[2] * vel[2]);
      const double vel_norm = vnorm_raw + 1e-5;  // :358

      double dist = 0, gd = 0, cd = 0, grad[3];
      bool in_map = true;
      dist = sdf_map->getDistWithGradTrilinear(pos, grad, &in_map);
      if (!in_map) ++oob;
      cd = prm.alpha * std::exp(-(dist - prm.d0) / prm.r);               // :509
      gd = -(prm.alpha / prm.r) * std::exp(-(dist - prm.d0) / prm.r);    // :514

      double T6[6], TV[6];
      for (int i = 0; i < 6; ++i) T6[i] = std::pow(t, i);  // :544-551
      // T*V with V = superdiag(1..5): (T*V)_j = T_{j-1} * j
      TV[0] = 0.0;
      for (int j = 1; j < 6; ++j) TV[j] = T6[j - 1] * (double)j;

      cost_colli += cd * vel_norm * dt;  // :373

      for (int k = 0; k < 3; k++) {  // :376-381
        double sA, sB;
        if (!exact) {
          sA = gd * grad[k] * cd * vel_norm;  // H4: extra cd factor reproduced
          sB = cd * (vel[k] / vel_norm);
        } else {
          sA = gd * grad[k] * vel_norm;
          sB = vnorm_raw > 0 ? cd * (vel[k] / vnorm_raw) : 0.0;
        }
        addScaledRowProduct(r1.data(), sA, T6, Ldp.data(), P);
        addScaledRowProduct(r2.data(), sB, TV, Ldp.data(), P);
        double* row = &g_colli[(size_t)k * P];
        for (int j = 0; j < P; ++j) row[j] = row[j] + (r1[j] + r2[j]) * dt;
      }

      if (prm.vel_acc_penalty && step == 2) {  // commented-out block :383-407, literal (H5, a9)
        double cv = 0, ca = 0, gv = 0, ga = 0, acc[3];
        accFromCoeff(acc, coe, s, t, exact);
        for (int k = 0; k < 3; k++) {
          cv = prm.alpha_v * std::exp((std::fabs(vel[k]) - prm.v0) / prm.r_v);
          cost_vel += cv * vel_norm * dt;
          ca = prm.alpha_a * std::exp((std::fabs(acc[k]) - prm.a0) / prm.r_a);
          cost_acc += ca * vel_norm * dt;
        }
        double TVV[6];  // T*V*V: (TV*V)_j = TV_{j-1} * j
        TVV[0] = 0.0;
        for (int j = 1; j < 6; ++j) TVV[j] = TV[j - 1] * (double)j;
        for (int k = 0; k < 3; k++) {
          gv = (prm.alpha_v / prm.r_v) * std::exp((std::fabs(vel[k]) - prm.v0) / prm.r_v);
          addScaledRowProduct(r1.data(), gv * vel_norm, TV, Ldp.data(), P);
          addScaledRowProduct(r2.data(), cv * (vel[k] / vel_norm), TV, Ldp.data(), P);  // cv carried from k=2
          double* rv = &g_vel[(size_t)k * P];
          for (int j = 0; j < P; ++j) rv[j] = rv[j] + (r1[j] + r2[j]) * dt;
          ga = (prm.alpha_a / prm.r_a) * std::exp((std::fabs(acc[k]) - prm.a0) / prm.r_a);
          addScaledRowProduct(r1.data(), ga * vel_norm, TVV, Ldp.data(), P);
          addScaledRowProduct(r2.data(), ca * (vel[k] / vel_norm), TV, Ldp.data(), P);  // ca carried from k=2
          double* ra = &g_acc[(size_t)k * P];
          for (int j = 0; j < P; ++j) ra[j] = ra[j] + (r1[j] + r2[j]) * dt;
        }
      }
    }
  }

  double ws = prm.w_smooth, wc = prm.w_collision, wv = 1.0, wa = 1.0;  // :412
  if (step == 1) ws = 0.0;                                              // OPT_FIRST_STEP
  cost = ws * cost_smooth + wc * cost_colli + wv * cost_vel + wa * cost_acc + 1e-3;  // :417-418, H10
  _grad.resize((size_t)P * 3);
  for (size_t i = 0; i < (size_t)P * 3; ++i)
    _grad[i] = (ws * g_smooth[i] + wc * g_colli[i] + wv * g_vel[i] + wa * g_acc[i]) + 1e-5;  // :425-431
  if (out_of_map) *out_of_map += oob;
  if (terms) terms[0] = cost_smooth, terms[1] = cost_colli, terms[2] = cost_vel, terms[3] = cost_acc;
}

double GradTrajOptimizer::costFunc(unsigned n, const double* x, double* grad, void* data) {  // :554-562
  GradTrajOptimizer* g = reinterpret_cast<GradTrajOptimizer*>(data);
  std::vector<double> xv(x, x + n), gv;
  double cost;
  g->iter_num++;  // :284 (kept out of getCostAndGradient so batch evaluation stays thread-safe)
  g->getCostAndGradient(xv, cost, gv);
  if (grad) std::memcpy(grad, gv.data(), n * sizeof(double));
  // cost curve bookkeeping (:438-447)
  if (g->vec_cost.empty() || g->vec_cost.back() > cost)
    g->vec_cost.push_back(cost);
  else
    g->vec_cost.push_back(g->vec_cost.back());
  return cost;
}

// NLopt 2.5.0 through its C ABI (prototypes hand-declared: nlopt.h is not in the image).
namespace {
typedef struct nlopt_opt_s* nlopt_opt;
struct NloptApi {
  void* h = nullptr;
  nlopt_opt (*create)(int, unsigned) = nullptr;
  void (*destroy)(nlopt_opt) = nullptr;
  int (*set_min_objective)(nlopt_opt, nlopt_func_t, void*) = nullptr;
  int (*set_lower_bounds)(nlopt_opt, const double*) = nullptr;
  int (*set_upper_bounds)(nlopt_opt, const double*) = nullptr;
  int (*set_maxeval)(nlopt_opt, int) = nullptr;
  int (*set_maxtime)(nlopt_opt, double) = nullptr;
  int (*set_xtol_abs1)(nlopt_opt, double) = nullptr;
  int (*optimize)(nlopt_opt, double*, double*) = nullptr;
  void (*srand_)(unsigned long) = nullptr;
};
bool loadNlopt(const char* path, NloptApi& api) {
  api.h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
  if (!api.h) return false;
#define GTOP_SYM(field, name) *(void**)(&api.field) = dlsym(api.h, name)
  GTOP_SYM(create, "nlopt_create");
  GTOP_SYM(destroy, "nlopt_destroy");
  GTOP_SYM(set_min_objective, "nlopt_set_min_objective");
  GTOP_SYM(set_lower_bounds, "nlopt_set_lower_bounds");
  GTOP_SYM(set_upper_bounds, "nlopt_set_upper_bounds");
  GTOP_SYM(set_maxeval, "nlopt_set_maxeval");
  GTOP_SYM(set_maxtime, "nlopt_set_maxtime");
  GTOP_SYM(set_xtol_abs1, "nlopt_set_xtol_abs1");
  GTOP_SYM(optimize, "nlopt_optimize");
  GTOP_SYM(srand_, "nlopt_srand");
#undef GTOP_SYM
  return api.create && api.destroy && api.set_min_objective && api.set_lower_bounds && api.set_upper_bounds &&
         api.set_maxeval && api.set_maxtime && api.optimize;
}
}  // namespace

int GradTrajOptimizer::optimizeTrajectory(int step_, const SolveOptions& so, const char* nlopt_path,
                                          nlopt_func_t override_f, void* override_data) {  // :128-243
  if (step_ != 0 && step_ != 1 && step_ != 2) std::fprintf(stderr, "step number error, step should be 0, 1 or 2\n");
  this->step = step_;
  NloptApi api;
  if (!loadNlopt(nlopt_path, api)) return -1000;
  if (api.srand_) api.srand_(12345);  // the reference seeds from ros::Time (:135); irrelevant for MMA
  const unsigned n = 3 * (unsigned)num_dp;
  nlopt_opt opt = api.create(prm.algorithm, n);
  if (!opt) return -1001;
  api.set_min_objective(opt, override_f ? override_f : &GradTrajOptimizer::costFunc,
                        override_f ? override_data : (void*)this);
  if (so.xtol_abs > 0 && api.set_xtol_abs1) api.set_xtol_abs1(opt, so.xtol_abs);
  if (so.maxtime > 0) api.set_maxtime(opt, so.maxtime);
  if (so.maxeval > 0) api.set_maxeval(opt, so.maxeval);
  std::vector<double> lb, ub;
  bounds(lb, ub);
  api.set_lower_bounds(opt, lb.data());
  api.set_upper_bounds(opt, ub.data());
  std::vector<double> x = x0();
  double min_f = 0.0;
  vec_cost.clear();
  int result = api.optimize(opt, x.data(), &min_f);
  api.destroy(opt);
  for (int i = 0; i < num_dp; ++i) Dp(0, i) = x[i], Dp(1, i) = x[i + num_dp], Dp(2, i) = x[i + 2 * num_dp];  // :202-207
  getCoefficientFromDerivative(this->coeff, x);  // :230
  last_min_f = min_f;
  last_result = result;
  if (step_ == 2) iter_num = 0;
  return result;
}

}  // namespace gtop_oracle

// =============================================================================================
// PolynomialTraj (polynomial_traj.hpp)
// =============================================================================================
namespace gtop_oracle {

void PolynomialTraj::reset() {  // :27-30
  times.clear(), cxs.clear(), cys.clear(), czs.clear();
  time_sum = 0.0, num_seg = 0;
}

void PolynomialTraj::addSegment(const std::vector<double>& cx, const std::vector<double>& cy, const std::vector<double>& cz,
                                double t) {  // :32-35
  cxs.push_back(cx), cys.push_back(cy), czs.push_back(cz), times.push_back(t);
}

void PolynomialTraj::init() {  // :37-43
  num_seg = (int)times.size();
  time_sum = 0.0;
  for (size_t i = 0; i < times.size(); ++i) time_sum += times[i];
}

void PolynomialTraj::evaluate(double t, double pt[3]) const {  // :45-65
  int idx = 0;
  while (idx < num_seg - 1 && times[idx] <= t) {  // (the reference has no bound on idx — H16)
    t -= times[idx];
    ++idx;
  }
  const int order = (int)cxs[idx].size();
  std::vector<double> tv(order);
  for (int i = 0; i < order; ++i) tv[order - 1 - i] = std::pow(t, double(i));
  pt[0] = pt[1] = pt[2] = 0.0;
  for (int k = 0; k < order; ++k) pt[0] += tv[k] * cxs[idx][k], pt[1] += tv[k] * cys[idx][k], pt[2] += tv[k] * czs[idx][k];
}

const std::vector<double>& PolynomialTraj::getTraj() {  // :70-79
  double eval_t = 0.0;
  traj_vec3d.clear();
  while (eval_t <= time_sum) {
    double pt[3];
    evaluate(eval_t, pt);
    traj_vec3d.insert(traj_vec3d.end(), pt, pt + 3);
    eval_t += 0.01;
  }
  return traj_vec3d;
}

double PolynomialTraj::getLength() {  // :81-91
  length = 0.0;
  for (size_t i = 1; i < traj_vec3d.size() / 3; ++i) {
    const double dx = traj_vec3d[3 * i] - traj_vec3d[3 * i - 3], dy = traj_vec3d[3 * i + 1] - traj_vec3d[3 * i - 2],
                 dz = traj_vec3d[3 * i + 2] - traj_vec3d[3 * i - 1];
    length += std::sqrt(dx * dx + dy * dy + dz * dz);
  }
  return length;
}

double PolynomialTraj::getAccCost() const {  // :95-107
  double cost = 0.0;
  const int order = (int)cxs[0].size();
  for (size_t s = 0; s < times.size(); ++s) {
    const double ux = 2 * cxs[s][order - 3], uy = 2 * cys[s][order - 3], uz = 2 * czs[s][order - 3];
    cost += (ux * ux + uy * uy + uz * uz) * times[s];
  }
  return cost;
}

double PolynomialTraj::getJerk() const {  // :109-139
  double jerk = 0.0;
  for (size_t s = 0; s < times.size(); ++s) {
    const int order = (int)cxs[s].size();
    std::vector<double> cxv(order), cyv(order), czv(order);
    for (int j = 0; j < order; ++j) cxv[j] = cxs[s][order - 1 - j], cyv[j] = cys[s][order - 1 - j], czv[j] = czs[s][order - 1 - j];
    const double ts = times[s];
    std::vector<double> mat((size_t)order * order, 0.0);
    for (double i = 3; i < order; i += 1)
      for (double j = 3; j < order; j += 1)
        mat[(size_t)i * order + (size_t)j] = i * (i - 1) * (i - 2) * j * (j - 1) * (j - 2) * std::pow(ts, i + j - 5) / (i + j - 5);
    auto quad = [&](const std::vector<double>& c) {
      double q = 0.0;
      for (int j = 0; j < order; ++j) {
        double col = 0.0;
        for (int i = 0; i < order; ++i) col += c[i] * mat[(size_t)i * order + j];
        q += col * c[j];
      }
      return q;
    };
    jerk += quad(cxv);
    jerk += quad(cyv);
    jerk += quad(czv);
  }
  return jerk;
}

void PolynomialTraj::getMeanAndMaxVel(double& mean_v, double& max_v) const {  // :141-171
  int num = 0;
  mean_v = 0.0, max_v = -1.0;
  for (size_t s = 0; s < times.size(); ++s) {
    const int order = (int)cxs[s].size();
    std::vector<double> vx(order - 1), vy(order - 1), vz(order - 1);
    for (int i = 0; i < order - 1; ++i) {
      vx[i] = double(i + 1) * cxs[s][order - 2 - i];
      vy[i] = double(i + 1) * cys[s][order - 2 - i];
      vz[i] = double(i + 1) * czs[s][order - 2 - i];
    }
    const double ts = times[s];
    double eval_t = 0.0;
    while (eval_t < ts) {
      double v[3] = {0, 0, 0};
      for (int i = 0; i < order - 1; ++i) {
        const double tv = std::pow(ts, i);  // ts, not eval_t (:158 — H16)
        v[0] += tv * vx[i], v[1] += tv * vy[i], v[2] += tv * vz[i];
      }
      const double vn = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
      mean_v += vn;
      if (vn > max_v) max_v = vn;
      ++num;
      eval_t += 0.01;
    }
  }
  mean_v = mean_v / double(num);
}

void PolynomialTraj::getMeanAndMaxAcc(double& mean_a, double& max_a) const {  // :173-203
  int num = 0;
  mean_a = 0.0, max_a = -1.0;
  for (size_t s = 0; s < times.size(); ++s) {
    const int order = (int)cxs[s].size();
    std::vector<double> ax(order - 2), ay(order - 2), az(order - 2);
    for (int i = 0; i < order - 2; ++i) {
      ax[i] = double((i + 2) * (i + 1)) * cxs[s][order - 3 - i];
      ay[i] = double((i + 2) * (i + 1)) * cys[s][order - 3 - i];
      az[i] = double((i + 2) * (i + 1)) * czs[s][order - 3 - i];
    }
    const double ts = times[s];
    double eval_t = 0.0;
    while (eval_t < ts) {
      double a[3] = {0, 0, 0};
      for (int i = 0; i < order - 2; ++i) {
        const double tv = std::pow(ts, i);  // ts, not eval_t (:190 — H16)
        a[0] += tv * ax[i], a[1] += tv * ay[i], a[2] += tv * az[i];
      }
      const double an = std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
      mean_a += an;
      if (an > max_a) max_a = an;
      ++num;
      eval_t += 0.01;
    }
  }
  mean_a = mean_a / double(num);
}

}  // namespace gtop_oracle

// =============================================================================================
// EDTEnvironment (edt_environment.cpp)
// =============================================================================================
namespace gtop_oracle {

double EDTEnvironment::distToBox(int idx, const double pos[3], double time) const {  // :27-44
  const MovingBox& b = boxes_[idx];
  double d[3];
  for (int i = 0; i < 3; ++i) {
    const double pos_box = 1.0 * b.c0[i] + std::pow(time, 1) * b.c1[i];  // evaluateConstVel (obj_predictor.h:57-66)
    const double box_max = pos_box + 0.5 * b.scale[i];
    const double box_min = pos_box - 0.5 * b.scale[i];
    d[i] = pos[i] >= box_min && pos[i] <= box_max ? 0.0 : std::min(std::fabs(pos[i] - box_min), std::fabs(pos[i] - box_max));
  }
  return std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
}

double EDTEnvironment::minDistToAllBox(const double pos[3], double time) const {  // :61-73
  double dist = 10000000.0;
  for (size_t i = 0; i < boxes_.size(); i++) {
    const double di = distToBox((int)i, pos, time);
    if (di < dist) dist = di;
  }
  return dist;
}

void EDTEnvironment::evaluateEDTWithGrad(const double pos[3], double time, double& dist, double grad[3]) const {  // :75-122
  // SDFMap::getInterpolationData (absent from the reference; restated from sdf_map.cpp:196-209)
  const SDFMap& M = *sdf_map_;
  double pos_m[3], idx_pos[3], diff[3];
  int idx[3];
  for (int i = 0; i < 3; ++i) pos_m[i] = pos[i] - 0.5 * M.resolution * 1.0;
  M.posToIndex(pos_m, idx);
  for (int i = 0; i < 3; ++i) idx_pos[i] = (idx[i] + 0.5) * M.resolution + M.origin[i];
  for (int i = 0; i < 3; ++i) diff[i] = (pos[i] - idx_pos[i]) * M.resolution_inv;
  double values[2][2][2];
  for (int x = 0; x < 2; x++)
    for (int y = 0; y < 2; y++)
      for (int z = 0; z < 2; z++) {
        const int c[3] = {idx[0] + x, idx[1] + y, idx[2] + z};
        double pt[3];
        for (int i = 0; i < 3; ++i) pt[i] = (c[i] + 0.5) * M.resolution + M.origin[i];  // indexToPos
        const double d1 = M.getDistance(pt);
        if (time < 0.0) {
          values[x][y][z] = d1;
        } else {
          const double d2 = minDistToAllBox(pt, time);
          values[x][y][z] = std::min(d1, d2);
        }
      }
  const double v00 = (1 - diff[0]) * values[0][0][0] + diff[0] * values[1][0][0];
  const double v01 = (1 - diff[0]) * values[0][0][1] + diff[0] * values[1][0][1];
  const double v10 = (1 - diff[0]) * values[0][1][0] + diff[0] * values[1][1][0];
  const double v11 = (1 - diff[0]) * values[0][1][1] + diff[0] * values[1][1][1];
  const double v0 = (1 - diff[1]) * v00 + diff[1] * v10;
  const double v1 = (1 - diff[1]) * v01 + diff[1] * v11;
  dist = (1 - diff[2]) * v0 + diff[2] * v1;
  grad[2] = (v1 - v0) * resolution_inv_;
  grad[1] = ((1 - diff[2]) * (v10 - v00) + diff[2] * (v11 - v01)) * resolution_inv_;
  grad[0] = (1 - diff[2]) * (1 - diff[1]) * (values[1][0][0] - values[0][0][0]);
  grad[0] += (1 - diff[2]) * diff[1] * (values[1][1][0] - values[0][1][0]);
  grad[0] += diff[2] * (1 - diff[1]) * (values[1][0][1] - values[0][0][1]);
  grad[0] += diff[2] * diff[1] * (values[1][1][1] - values[0][1][1]);
  grad[0] *= resolution_inv_;
}

double EDTEnvironment::evaluateCoarseEDT(const double pos[3], double time) const {  // :124-137
  const double d1 = sdf_map_->getDistance(pos);
  if (time < 0.0) return d1;
  return std::min(d1, minDistToAllBox(pos, time));
}

}  // namespace gtop_oracle
