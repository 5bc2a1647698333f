// oracle_math.cpp -- TEST INFRASTRUCTURE (see vloam_oracle.h header).
// Third-party semantics the reference relies on, restated from their published
// algorithms: Eigen quaternion algebra, PCL VoxelGrid,
// FLANN exact kNN (brute force truth + single KD-tree for timing), a 3x3
// symmetric eigen-solver and a 5x3 column-pivoted Householder least squares.
#include <math.h>
#include <float.h>
#include <string.h>
#include <algorithm>
#include <numeric>
#include "vloam_oracle.h"

namespace vo {

// ---------------------------------------------------------------------------
// Eigen 3.3 quaternion semantics, storage x,y,z,w.
// ---------------------------------------------------------------------------
void q_mul(const double a[4], const double b[4], double o[4]) {
  const double ax = a[0], ay = a[1], az = a[2], aw = a[3];
  const double bx = b[0], by = b[1], bz = b[2], bw = b[3];
  const double w = aw * bw - ax * bx - ay * by - az * bz;
  const double x = aw * bx + ax * bw + ay * bz - az * by;
  const double y = aw * by + ay * bw + az * bx - ax * bz;
  const double z = aw * bz + az * bw + ax * by - ay * bx;
  o[0] = x; o[1] = y; o[2] = z; o[3] = w;
}

// Eigen QuaternionBase::_transformVector: uv = u x v; uv += uv; v + w*uv + u x uv
// (used by LO.cpp:167 and LM.cpp:158; no normalisation of q).
void q_rot(const double q[4], const double v[3], double o[3]) {
  const double ux = q[0], uy = q[1], uz = q[2], w = q[3];
  double uvx = uy * v[2] - uz * v[1];
  double uvy = uz * v[0] - ux * v[2];
  double uvz = ux * v[1] - uy * v[0];
  uvx += uvx; uvy += uvy; uvz += uvz;
  const double cx = uy * uvz - uz * uvy;
  const double cy = uz * uvx - ux * uvz;
  const double cz = ux * uvy - uy * uvx;
  o[0] = (v[0] + w * uvx) + cx;
  o[1] = (v[1] + w * uvy) + cy;
  o[2] = (v[2] + w * uvz) + cz;
}

void q_inv(const double q[4], double o[4]) {  // conjugate / squaredNorm (LM.cpp:149)
  const double n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  o[0] = -q[0] / n2; o[1] = -q[1] / n2; o[2] = -q[2] / n2; o[3] = q[3] / n2;
}

// ---------------------------------------------------------------------------
// pcl::VoxelGrid<PointXYZI>::applyFilter, downsample_all_data,
// min_points_per_voxel = 0.  Canonical refinement of the unstable std::sort:
// order (voxel idx, point index); f32 sums in that order; centroid = sum / n.
// ---------------------------------------------------------------------------
void voxel_grid(const Cloud& in, float leaf, Cloud& out) {
  out.clear();
  if (in.empty()) return;
  const float inv = 1.0f / leaf;  // inverse_leaf_size_ = 1 / leaf_size_ (float)
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (const P4& p : in) {  // getMinMax3D
    mn[0] = std::min(mn[0], p.x); mx[0] = std::max(mx[0], p.x);
    mn[1] = std::min(mn[1], p.y); mx[1] = std::max(mx[1], p.y);
    mn[2] = std::min(mn[2], p.z); mx[2] = std::max(mx[2], p.z);
  }
  // overflow guard: leaf too small for the extent -> output = input
  const int64_t dx = (int64_t)((mx[0] - mn[0]) * inv) + 1;
  const int64_t dy = (int64_t)((mx[1] - mn[1]) * inv) + 1;
  const int64_t dz = (int64_t)((mx[2] - mn[2]) * inv) + 1;
  if (dx * dy * dz > (int64_t)INT32_MAX) { out = in; return; }
  int minb[3], maxb[3], divb[3];
  for (int a = 0; a < 3; ++a) {
    minb[a] = (int)floorf(mn[a] * inv);
    maxb[a] = (int)floorf(mx[a] * inv);
    divb[a] = maxb[a] - minb[a] + 1;
  }
  const int mul[3] = {1, divb[0], divb[0] * divb[1]};
  struct Key { unsigned idx; unsigned pt; };
  std::vector<Key> keys(in.size());
  for (size_t k = 0; k < in.size(); ++k) {
    const P4& p = in[k];
    const int i0 = (int)(floorf(p.x * inv) - (float)minb[0]);
    const int i1 = (int)(floorf(p.y * inv) - (float)minb[1]);
    const int i2 = (int)(floorf(p.z * inv) - (float)minb[2]);
    keys[k].idx = (unsigned)(i0 * mul[0] + i1 * mul[1] + i2 * mul[2]);
    keys[k].pt = (unsigned)k;
  }
  std::sort(keys.begin(), keys.end(), [](const Key& a, const Key& b) {
    return a.idx != b.idx ? a.idx < b.idx : a.pt < b.pt;
  });
  size_t s = 0;
  while (s < keys.size()) {
    size_t e = s + 1;
    while (e < keys.size() && keys[e].idx == keys[s].idx) ++e;
    float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;  // CentroidPoint accumulators (float)
    for (size_t k = s; k < e; ++k) {
      const P4& p = in[keys[k].pt];
      sx += p.x; sy += p.y; sz += p.z; si += p.i;
    }
    const float n = (float)(e - s);
    P4 c = {sx / n, sy / n, sz / n, si / n};
    out.push_back(c);
    s = e;
  }
}

// ---------------------------------------------------------------------------
// Exact kNN.  FLANN L2_Simple<float>: acc = 0; acc += d*d for x, y, z (f32).
// Canonical refinement of FLANN's traversal-order ties: (d2, index) ascending.
// ---------------------------------------------------------------------------
static inline float dist2(const P4& a, const P4& q) {
  float acc = 0.f, d;
  d = q.x - a.x; acc += d * d;
  d = q.y - a.y; acc += d * d;
  d = q.z - a.z; acc += d * d;
  return acc;
}

struct TopK {
  int k, n;
  int* idx;
  float* d2;
  inline bool better(float d, int i, int slot) const { return d < d2[slot] || (d == d2[slot] && i < idx[slot]); }
  inline void push(float d, int i) {
    if (n == k && !better(d, i, k - 1)) return;
    int pos = (n < k) ? n++ : k - 1;
    while (pos > 0 && better(d, i, pos - 1)) { d2[pos] = d2[pos - 1]; idx[pos] = idx[pos - 1]; --pos; }
    d2[pos] = d; idx[pos] = i;
  }
  inline float worst() const { return n < k ? FLT_MAX : d2[k - 1]; }
};

int brute_knn(const Cloud& c, const P4& q, int k, int* idx, float* d2) {
  TopK t{k, 0, idx, d2};
  for (int i = 0; i < (int)c.size(); ++i) t.push(dist2(c[i], q), i);
  return t.n;
}

struct KdNode { int left, right, dim; float split; int lo, hi; };
struct KdTree {
  std::vector<KdNode> nodes;
  std::vector<int> perm;
  std::vector<P4> pts;  // reordered copy for locality (FLANN reorder=true)
};

static int kd_build_rec(KdTree* t, const Cloud& c, int lo, int hi) {
  const int id = (int)t->nodes.size();
  t->nodes.push_back(KdNode{-1, -1, 0, 0.f, lo, hi});
  if (hi - lo <= 15) return id;  // KDTreeSingleIndexParams(15)
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int k = lo; k < hi; ++k) {
    const P4& p = c[t->perm[k]];
    const float v[3] = {p.x, p.y, p.z};
    for (int a = 0; a < 3; ++a) { mn[a] = std::min(mn[a], v[a]); mx[a] = std::max(mx[a], v[a]); }
  }
  int dim = 0;
  if (mx[1] - mn[1] > mx[dim] - mn[dim]) dim = 1;
  if (mx[2] - mn[2] > mx[dim] - mn[dim]) dim = 2;
  const int mid = (lo + hi) / 2;
  auto coord = [&](int i) { const P4& p = c[i]; return dim == 0 ? p.x : (dim == 1 ? p.y : p.z); };
  std::nth_element(t->perm.begin() + lo, t->perm.begin() + mid, t->perm.begin() + hi,
                   [&](int a, int b) { return coord(a) < coord(b); });
  const float split = coord(t->perm[mid]);
  const int l = kd_build_rec(t, c, lo, mid);
  const int r = kd_build_rec(t, c, mid, hi);
  KdNode& n = t->nodes[id];
  n.left = l; n.right = r; n.dim = dim; n.split = split;
  return id;
}

KdTree* kd_build(const Cloud& c) {
  KdTree* t = new KdTree();
  t->perm.resize(c.size());
  std::iota(t->perm.begin(), t->perm.end(), 0);
  if (!c.empty()) {
    t->nodes.reserve(c.size() / 4 + 16);
    kd_build_rec(t, c, 0, (int)c.size());
  }
  t->pts.resize(c.size());
  for (size_t k = 0; k < c.size(); ++k) t->pts[k] = c[t->perm[k]];
  return t;
}
void kd_free(KdTree* t) { delete t; }

static void kd_search(const KdTree* t, int node, const P4& q, TopK& top) {
  const KdNode& n = t->nodes[node];
  if (n.left < 0) {
    for (int k = n.lo; k < n.hi; ++k) top.push(dist2(t->pts[k], q), t->perm[k]);
    return;
  }
  const float qv = n.dim == 0 ? q.x : (n.dim == 1 ? q.y : q.z);
  const float diff = qv - n.split;
  const int near = diff < 0.f ? n.left : n.right;
  const int far = diff < 0.f ? n.right : n.left;
  kd_search(t, near, q, top);
  // f32 lower bound of any d2 on the far side; <= keeps (d2, index) ties exact
  if (diff * diff <= top.worst()) kd_search(t, far, q, top);
}

int kd_knn(const KdTree* t, const Cloud& c, const P4& q, int k, int* idx, float* d2) {
  (void)c;
  TopK top{k, 0, idx, d2};
  if (!t->nodes.empty()) kd_search(t, 0, q, top);
  return top.n;
}

// ---------------------------------------------------------------------------
// Eigen::SelfAdjointEigenSolver<Matrix3d>::compute (LM.cpp:583-591), restated as Eigen 3.3 runs it
// (Eigen/src/Eigenvalues/SelfAdjointEigenSolver.h: compute() -> tridiagonalization_inplace (the closed
// 3x3 special case of Tridiagonalization.h) -> computeFromTridiagonal_impl: implicit symmetric QR steps
// with Wilkinson shift and Givens rotations (tridiagonal_qr_step, Jacobi.h makeGivens), deflation test
// |e_i| <= eps * sqrt(|d_i| + |d_i+1|), selection sort of the eigenvalues, ascending):
//   1. scale = max |a_ij| (1 if the matrix is zero); work on A / scale; eigenvalues are scaled back.
//   2. one Householder-like reflection zeroes a(2,0).
//   3. QR steps on the largest unreduced trailing block until every sub-diagonal entry deflates
//      (at most 30 * n steps).
// Eigenvalues ascending, unit eigenvectors in columns, sign unspecified by the contract.  The CUDA
// path uses a cyclic Jacobi solver instead: this routine is its INDEPENDENT
// witness -- a different algorithm, so agreement (accept flags, factors to ~1e-15) is evidence and
// not an identity.  Eigen itself is not in this container: restated from its published source.
// ---------------------------------------------------------------------------
namespace {
struct Givens { double c, s; };
inline Givens make_givens(double p, double q) {  // Eigen JacobiRotation<double>::makeGivens (real case)
  Givens g;
  if (q == 0.0) { g.c = p < 0.0 ? -1.0 : 1.0; g.s = 0.0; }
  else if (p == 0.0) { g.c = 0.0; g.s = q < 0.0 ? 1.0 : -1.0; }
  else if (fabs(p) > fabs(q)) {
    const double t = q / p;
    double u = sqrt(1.0 + t * t);
    if (p < 0.0) u = -u;
    g.c = 1.0 / u; g.s = -t * g.c;
  } else {
    const double t = p / q;
    double u = sqrt(1.0 + t * t);
    if (q < 0.0) u = -u;
    g.s = -1.0 / u; g.c = -t * g.s;
  }
  return g;
}
// one implicit QR step on the unreduced block [start, end] of the tridiagonal (diag, subdiag); Q <- Q G_k
inline void tridiagonal_qr_step(double* diag, double* subdiag, int start, int end, double Q[3][3]) {
  const double td = (diag[end - 1] - diag[end]) * 0.5;
  const double e = subdiag[end - 1];
  double mu = diag[end];
  if (td == 0.0) mu -= fabs(e);
  else if (e != 0.0) {
    const double e2 = e * e;
    const double h = hypot(td, e);
    if (e2 == 0.0) mu -= e / ((td + (td > 0.0 ? h : -h)) / e);
    else mu -= e2 / (td + (td > 0.0 ? h : -h));
  }
  double x = diag[start] - mu;
  double z = subdiag[start];
  for (int k = start; k < end && z != 0.0; ++k) {
    const Givens r = make_givens(x, z);
    const double sdk = r.s * diag[k] + r.c * subdiag[k];
    const double dkp1 = r.s * subdiag[k] + r.c * diag[k + 1];
    diag[k] = r.c * (r.c * diag[k] - r.s * subdiag[k]) - r.s * (r.c * subdiag[k] - r.s * diag[k + 1]);
    diag[k + 1] = r.s * sdk + r.c * dkp1;
    subdiag[k] = r.c * sdk - r.s * dkp1;
    if (k > start) subdiag[k - 1] = r.c * subdiag[k - 1] - r.s * z;
    x = subdiag[k];
    if (k < end - 1) { z = -r.s * subdiag[k + 1]; subdiag[k + 1] = r.c * subdiag[k + 1]; }
    for (int i = 0; i < 3; ++i) {  // applyOnTheRight(k, k + 1, rot)
      const double xi = Q[i][k], yi = Q[i][k + 1];
      Q[i][k] = r.c * xi - r.s * yi;
      Q[i][k + 1] = r.s * xi + r.c * yi;
    }
  }
}
}  // namespace

void sym_eig3(const double Ain[9], double evals[3], double evecs[9]) {
  // (Eigen reads the lower triangle only)
  double m[3][3];
  double scale = 0.0;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j <= i; ++j) scale = fmax(scale, fabs(Ain[i * 3 + j]));
  if (scale == 0.0) scale = 1.0;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) m[i][j] = Ain[(i >= j ? i * 3 + j : j * 3 + i)] / scale;
  double diag[3], subdiag[2], Q[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  // tridiagonalization_inplace_selector<MatrixType, 3, false>
  diag[0] = m[0][0];
  const double v1norm2 = m[2][0] * m[2][0];
  if (v1norm2 <= DBL_MIN) {
    diag[1] = m[1][1]; diag[2] = m[2][2];
    subdiag[0] = m[1][0]; subdiag[1] = m[2][1];
  } else {
    const double beta = sqrt(m[1][0] * m[1][0] + v1norm2);
    const double invBeta = 1.0 / beta;
    const double m01 = m[1][0] * invBeta, m02 = m[2][0] * invBeta;
    const double q = 2.0 * m01 * m[2][1] + m02 * (m[2][2] - m[1][1]);
    diag[1] = m[1][1] + m02 * q;
    diag[2] = m[2][2] - m02 * q;
    subdiag[0] = beta;
    subdiag[1] = m[2][1] - m01 * q;
    Q[1][1] = m01; Q[1][2] = m02; Q[2][1] = m02; Q[2][2] = -m01;
  }
  // computeFromTridiagonal_impl
  const int n = 3, maxIterations = 30;
  int end = n - 1, start = 0, iter = 0;
  const double considerAsZero = DBL_MIN, precision_inv = 1.0 / DBL_EPSILON;
  while (end > 0) {
    for (int i = start; i < end; ++i) {
      if (fabs(subdiag[i]) < considerAsZero) subdiag[i] = 0.0;
      else {
        const double scaled = precision_inv * subdiag[i];
        if (scaled * scaled <= fabs(diag[i]) + fabs(diag[i + 1])) subdiag[i] = 0.0;
      }
    }
    while (end > 0 && subdiag[end - 1] == 0.0) end--;
    if (end <= 0) break;
    if (++iter > maxIterations * n) break;
    start = end - 1;
    while (start > 0 && subdiag[start - 1] != 0.0) start--;
    tridiagonal_qr_step(diag, subdiag, start, end, Q);
  }
  for (int i = 0; i < n - 1; ++i) {  // selection sort, ascending
    int k = 0;
    for (int j = 1; j < n - i; ++j) if (diag[i + j] < diag[i + k]) k = j;
    if (k > 0) {
      std::swap(diag[i], diag[i + k]);
      for (int r = 0; r < 3; ++r) std::swap(Q[r][i], Q[r][i + k]);
    }
  }
  for (int c = 0; c < 3; ++c) {
    evals[c] = diag[c] * scale;
    for (int r = 0; r < 3; ++r) evecs[r * 3 + c] = Q[r][c];
  }
}

// The two fits of LaserMapping::solveMapping on five f32 neighbours (LM.cpp:559-603, 637-680), callable on their own
// (vloam_oracle_fit): returns the accept flag and writes the factor parameters {a[3], b[3]} (edge: the two synthetic
// line points; plane: unit normal and {d, 0, 0}).
bool fit_line5(const float near_f[15], double a[3], double b[3]) {
  double c[3] = {0, 0, 0}, near[5][3];
  for (int j = 0; j < 5; j++) {
    for (int k = 0; k < 3; ++k) { near[j][k] = near_f[j * 3 + k]; c[k] = c[k] + near[j][k]; }
  }
  for (int k = 0; k < 3; ++k) c[k] = c[k] / 5.0;
  double cov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int j = 0; j < 5; j++) {
    const double z[3] = {near[j][0] - c[0], near[j][1] - c[1], near[j][2] - c[2]};
    for (int p = 0; p < 3; ++p) for (int q = 0; q < 3; ++q) cov[p * 3 + q] = cov[p * 3 + q] + z[p] * z[q];
  }
  double ev[3], evec[9]; sym_eig3(cov, ev, evec);
  if (!(ev[2] > 3 * ev[1])) return false;
  for (int k = 0; k < 3; ++k) {
    const double u = evec[k * 3 + 2];
    a[k] = 0.1 * u + c[k];
    b[k] = -0.1 * u + c[k];
  }
  return true;
}
bool fit_plane5(const float near_f[15], double nrm[3], double* d) {
  double A[15], B[5] = {-1, -1, -1, -1, -1};
  for (int j = 0; j < 15; j++) A[j] = near_f[j];
  double norm[3]; colpiv_qr_solve_5x3(A, B, norm);
  const double nn = sqrt(norm[0] * norm[0] + norm[1] * norm[1] + norm[2] * norm[2]);
  const double negative_OA_dot_norm = 1 / nn;
  if (nn > 0) { norm[0] /= nn; norm[1] /= nn; norm[2] /= nn; }  // Eigen normalize()
  for (int j = 0; j < 5; j++)
    if (fabs(norm[0] * A[j * 3] + norm[1] * A[j * 3 + 1] + norm[2] * A[j * 3 + 2] + negative_OA_dot_norm) > 0.2) return false;
  for (int k = 0; k < 3; ++k) nrm[k] = norm[k];
  *d = negative_OA_dot_norm;
  return true;
}

// ---------------------------------------------------------------------------
// Eigen ColPivHouseholderQR<Matrix<double,5,3>>::solve (LM.cpp:655): least
// squares with column pivoting; pivots below eps * 3 * |max pivot| are treated
// as rank deficiency (those unknowns are set to zero).
// ---------------------------------------------------------------------------
bool colpiv_qr_solve_5x3(const double Ain[15], const double bin[5], double x[3]) {
  double A[5][3], b[5];
  for (int i = 0; i < 5; ++i) { b[i] = bin[i]; for (int j = 0; j < 3; ++j) A[i][j] = Ain[i * 3 + j]; }
  int perm[3] = {0, 1, 2};
  double maxpivot = 0.0;
  int rank = 3;
  const double thresh = DBL_EPSILON * 3.0;
  double diag[3] = {0, 0, 0};
  for (int k = 0; k < 3; ++k) {
    int best = k; double bestn = -1.0;
    for (int j = k; j < 3; ++j) {
      double n2 = 0; for (int i = k; i < 5; ++i) n2 += A[i][j] * A[i][j];
      if (n2 > bestn) { bestn = n2; best = j; }
    }
    if (best != k) {
      for (int i = 0; i < 5; ++i) std::swap(A[i][k], A[i][best]);
      std::swap(perm[k], perm[best]);
    }
    // Householder on column k, rows k..4
    double tail2 = 0; for (int i = k + 1; i < 5; ++i) tail2 += A[i][k] * A[i][k];
    const double c0 = A[k][k];
    double beta, tau, v[5] = {0, 0, 0, 0, 0};
    if (tail2 <= DBL_MIN) { tau = 0; beta = c0; }
    else {
      beta = sqrt(c0 * c0 + tail2);
      if (c0 >= 0) beta = -beta;
      for (int i = k + 1; i < 5; ++i) v[i] = A[i][k] / (c0 - beta);
      tau = (beta - c0) / beta;
    }
    v[k] = 1.0;
    if (tau != 0) {
      for (int j = k + 1; j < 3; ++j) {
        double s = 0; for (int i = k; i < 5; ++i) s += v[i] * A[i][j];
        s *= tau; for (int i = k; i < 5; ++i) A[i][j] -= s * v[i];
      }
      double s = 0; for (int i = k; i < 5; ++i) s += v[i] * b[i];
      s *= tau; for (int i = k; i < 5; ++i) b[i] -= s * v[i];
    }
    A[k][k] = beta; for (int i = k + 1; i < 5; ++i) A[i][k] = 0;
    diag[k] = beta;
    if (fabs(beta) > maxpivot) maxpivot = fabs(beta);
  }
  rank = 0;
  for (int k = 0; k < 3; ++k) if (fabs(diag[k]) > thresh * maxpivot) ++rank;
  double y[3] = {0, 0, 0};
  for (int k = rank - 1; k >= 0; --k) {
    double s = b[k];
    for (int j = k + 1; j < rank; ++j) s -= A[k][j] * y[j];
    y[k] = s / A[k][k];
  }
  for (int k = 0; k < 3; ++k) x[perm[k]] = y[k];
  return rank == 3;
}

}  // namespace vo
