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
// Eigen::SelfAdjointEigenSolver<Matrix3d> contract (LM.cpp:583-591): eigenvalues
// ascending, unit eigenvectors in columns, sign unspecified.  Restated with the
// cyclic Jacobi method (same contract; agrees with Eigen's tridiagonal QR to a
// few ulp -- the departure is documented in DESIGN.md).
// ---------------------------------------------------------------------------
void sym_eig3(const double Ain[9], double evals[3], double evecs[9]) {
  double a[3][3], v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) a[i][j] = Ain[i * 3 + j];
  for (int sweep = 0; sweep < 30; ++sweep) {
    const double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
    if (off == 0.0) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (a[p][q] == 0.0) continue;
        const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        const double apq = a[p][q];
        a[p][p] -= t * apq;
        a[q][q] += t * apq;
        a[p][q] = a[q][p] = 0.0;
        const int r = 3 - p - q;
        const double arp = a[r][p], arq = a[r][q];
        a[r][p] = a[p][r] = c * arp - s * arq;
        a[r][q] = a[q][r] = s * arp + c * arq;
        for (int k = 0; k < 3; ++k) {
          const double vkp = v[k][p], vkq = v[k][q];
          v[k][p] = c * vkp - s * vkq;
          v[k][q] = s * vkp + c * vkq;
        }
      }
  }
  int ord[3] = {0, 1, 2};
  std::sort(ord, ord + 3, [&](int i, int j) { return a[i][i] < a[j][j]; });
  for (int c = 0; c < 3; ++c) {
    evals[c] = a[ord[c]][ord[c]];
    for (int r = 0; r < 3; ++r) evecs[r * 3 + c] = v[r][ord[c]];
  }
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
