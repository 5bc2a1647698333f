// oracle_ceres.cpp -- TEST INFRASTRUCTURE (see vloam_oracle.h header).
// Restatement of ceres::Solve as the reference configures it (LO.cpp:500-509,
// LM.cpp:710-717): Ceres 2.0.0 TRUST_REGION + LEVENBERG_MARQUARDT, DENSE_QR,
// max_num_iterations = 4, HuberLoss(0.1) on every block, the quaternion block
// on EigenQuaternionParameterization, Jacobi scaling on, everything else at
// its default.  Residuals are the functors of
// lidarFactor.hpp differentiated with forward-mode dual numbers exactly as
// ceres::AutoDiffCostFunction does (LF.hpp:14-144).
#include <math.h>
#include <float.h>
#include <algorithm>
#include <vector>
#include "vloam_oracle.h"

namespace vo {

// ---- dual numbers (ceres::Jet<double,7>) ----------------------------------
struct Jet {
  double a;
  double v[7];
  Jet() : a(0) { for (double& x : v) x = 0; }
  Jet(double s) : a(s) { for (double& x : v) x = 0; }  // NOLINT
  static Jet var(double s, int k) { Jet j(s); j.v[k] = 1.0; return j; }
};
static inline Jet operator+(const Jet& x, const Jet& y) { Jet r; r.a = x.a + y.a; for (int k = 0; k < 7; ++k) r.v[k] = x.v[k] + y.v[k]; return r; }
static inline Jet operator-(const Jet& x, const Jet& y) { Jet r; r.a = x.a - y.a; for (int k = 0; k < 7; ++k) r.v[k] = x.v[k] - y.v[k]; return r; }
static inline Jet operator*(const Jet& x, const Jet& y) { Jet r; r.a = x.a * y.a; for (int k = 0; k < 7; ++k) r.v[k] = y.a * x.v[k] + x.a * y.v[k]; return r; }
static inline Jet operator/(const Jet& x, const Jet& y) {
  Jet r; const double inv = 1.0 / y.a; r.a = x.a * inv; const double q = x.a * inv;
  for (int k = 0; k < 7; ++k) r.v[k] = (x.v[k] - q * y.v[k]) * inv;
  return r;
}
static inline Jet jsqrt(const Jet& x) { Jet r; r.a = sqrt(x.a); const double h = 1.0 / (2.0 * r.a); for (int k = 0; k < 7; ++k) r.v[k] = x.v[k] * h; return r; }
static inline double jsqrt(double x) { return sqrt(x); }
static inline double val(double x) { return x; }
static inline double val(const Jet& x) { return x.a; }

template <typename T> struct V3 { T x, y, z; };
template <typename T> static inline V3<T> cross(const V3<T>& a, const V3<T>& b) {
  return V3<T>{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
// Eigen Quaternion<T> * Vector3 (no normalisation), SURVEY A.5.
template <typename T> static inline V3<T> rot(const T q[4], const V3<T>& v) {
  const V3<T> u{q[0], q[1], q[2]};
  V3<T> uv = cross(u, v);
  uv = V3<T>{uv.x + uv.x, uv.y + uv.y, uv.z + uv.z};
  const V3<T> c = cross(u, uv);
  return V3<T>{(v.x + q[3] * uv.x) + c.x, (v.y + q[3] * uv.y) + c.y, (v.z + q[3] * uv.z) + c.z};
}

// ceres::Jet overloads of the functions Eigen's slerp calls (jet.h: abs, acos, sin)
static inline Jet jabs(const Jet& x) { Jet r; r.a = fabs(x.a); const double sg = x.a < 0.0 ? -1.0 : 1.0; for (int k = 0; k < 7; ++k) r.v[k] = sg * x.v[k]; return r; }
static inline Jet jacos(const Jet& x) { Jet r; r.a = acos(x.a); const double h = -1.0 / sqrt(1.0 - x.a * x.a); for (int k = 0; k < 7; ++k) r.v[k] = h * x.v[k]; return r; }
static inline Jet jsin(const Jet& x) { Jet r; r.a = sin(x.a); const double c = cos(x.a); for (int k = 0; k < 7; ++k) r.v[k] = c * x.v[k]; return r; }
static inline double jabs(double x) { return fabs(x); }
static inline double jacos(double x) { return acos(x); }
static inline double jsin(double x) { return sin(x); }

// Eigen::QuaternionBase::slerp(t, other) with *this = Identity (Eigen/src/Geometry/Quaternion.h), on T = double or Jet:
//   d = this.dot(other) = w;  |d| >= 1 - eps: scale0 = 1 - t, scale1 = t;  else theta = acos(|d|),
//   scale0 = sin((1 - t) theta) / sin(theta), scale1 = sin(t theta) / sin(theta);  d < 0: scale1 = -scale1;
//   result coeffs = scale0 * Identity + scale1 * other.   (Jet comparisons look at the value part only.)
template <typename T> static inline void slerp_identity(double t, const T q[4], T out[4]) {
  const double one = 1.0 - DBL_EPSILON;
  const T d = q[3];
  const T absD = jabs(d);
  T scale0, scale1;
  if (val(absD) >= one) { scale0 = T(1.0 - t); scale1 = T(t); }
  else {
    const T theta = jacos(absD);
    const T sinTheta = jsin(theta);
    scale0 = jsin(T(1.0 - t) * theta) / sinTheta;
    scale1 = jsin(T(t) * theta) / sinTheta;
  }
  if (val(d) < 0.0) scale1 = T(0.0) - scale1;
  out[0] = scale1 * q[0]; out[1] = scale1 * q[1]; out[2] = scale1 * q[2];
  out[3] = scale0 * T(1.0) + scale1 * q[3];
}
void slerp_identity_d(double t, const double q[4], double out[4]) { slerp_identity<double>(t, q, out); }

// The functors pass q through Identity.slerp(T(s), q) and scale t by s (LF.hpp:29-33, 86-90).  With s == 1 (DISTORTION
// false, LO.h:90; mapping always passes 1.0, LM.cpp:600) slerp returns scale0*I + scale1*q with scale0 == 0 and
// scale1 == +-1 exactly, with exactly-zero derivatives of the scales, so value and Jacobian equal the plain q*cp + t
//  and the restatement applies q directly; factors built under DISTORTION (f.slerp) take the literal path.
template <typename T> static int eval_factor(const Factor& f, const T q[4], const T t[3], T r[3]) {
  const V3<T> cp{T(f.p[0]), T(f.p[1]), T(f.p[2])};
  V3<T> lp;
  if (f.slerp && f.type != 2) {
    T qs[4]; slerp_identity<T>(f.s, q, qs);
    lp = rot(qs, cp);
    lp = V3<T>{lp.x + T(f.s) * t[0], lp.y + T(f.s) * t[1], lp.z + T(f.s) * t[2]};
  } else {
    lp = rot(q, cp);
    lp = V3<T>{lp.x + t[0], lp.y + t[1], lp.z + t[2]};
  }
  if (f.type == 0) {  // LidarEdgeFactor LF.hpp:22-50
    const V3<T> lpa{T(f.a[0]), T(f.a[1]), T(f.a[2])}, lpb{T(f.b[0]), T(f.b[1]), T(f.b[2])};
    const V3<T> nu = cross(V3<T>{lp.x - lpa.x, lp.y - lpa.y, lp.z - lpa.z},
                           V3<T>{lp.x - lpb.x, lp.y - lpb.y, lp.z - lpb.z});
    const V3<T> de{lpa.x - lpb.x, lpa.y - lpb.y, lpa.z - lpb.z};
    const T n = jsqrt(de.x * de.x + de.y * de.y + de.z * de.z);
    r[0] = nu.x / n; r[1] = nu.y / n; r[2] = nu.z / n;
    return 3;
  } else if (f.type == 1) {  // LidarPlaneFactor LF.hpp:78-99
    const V3<T> lpj{T(f.a[0]), T(f.a[1]), T(f.a[2])}, n{T(f.b[0]), T(f.b[1]), T(f.b[2])};
    r[0] = (lp.x - lpj.x) * n.x + (lp.y - lpj.y) * n.y + (lp.z - lpj.z) * n.z;
    return 1;
  } else {  // LidarPlaneNormFactor LF.hpp:121-133
    const V3<T> n{T(f.a[0]), T(f.a[1]), T(f.a[2])};
    r[0] = (n.x * lp.x + n.y * lp.y + n.z * lp.z) + T(f.b[0]);
    return 1;
  }
}

// ceres::HuberLoss(a).Evaluate
static inline void huber(double a, double s, double rho[3]) {
  const double b = a * a;
  if (s > b) {
    const double r = sqrt(s);
    rho[0] = 2.0 * a * r - b;
    rho[1] = std::max(DBL_MIN, a / r);
    rho[2] = -rho[1] / (2.0 * s);
  } else { rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; }
}

// EigenQuaternionParameterization::ComputeJacobian (4x3 row-major), SURVEY A.4
static inline void plus_jacobian(const double x[4], double P[12]) {
  P[0] = x[3];  P[1] = x[2];  P[2] = -x[1];
  P[3] = -x[2]; P[4] = x[3];  P[5] = x[0];
  P[6] = x[1];  P[7] = -x[0]; P[8] = x[3];
  P[9] = -x[0]; P[10] = -x[1]; P[11] = -x[2];
}
// EigenQuaternionParameterization::Plus + identity on t
static inline void plus(const double x[7], const double d[6], double o[7]) {
  const double n = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  if (n > 0.0) {
    const double s = sin(n) / n;
    const double dq[4] = {s * d[0], s * d[1], s * d[2], cos(n)};
    q_mul(dq, x, o);
  } else { o[0] = x[0]; o[1] = x[1]; o[2] = x[2]; o[3] = x[3]; }
  o[4] = x[4] + d[3]; o[5] = x[5] + d[4]; o[6] = x[6] + d[5];
}

struct Eval {
  double cost;
  std::vector<double> r;  // robustified residuals (m rows)
  std::vector<double> J;  // robustified local Jacobian, m x 6 row-major
  double g[6];            // J^T r
};

// ProgramEvaluator::Evaluate + ResidualBlock::Evaluate (loss correction with
// rho'' <= 0: residual and Jacobian rows scaled by sqrt(rho')).
static void evaluate(const std::vector<Factor>& fs, const double x[7], bool with_jac, Eval* e) {
  double cost = 0;
  if (!with_jac) {
    for (const Factor& f : fs) {
      double r[3];
      const int nr = eval_factor<double>(f, x, x + 4, r);
      double s = 0; for (int k = 0; k < nr; ++k) s += r[k] * r[k];
      double rho[3]; huber(0.1, s, rho);
      cost += 0.5 * rho[0];
    }
    e->cost = cost;
    return;
  }
  size_t rows = 0;
  for (const Factor& f : fs) rows += (f.type == 0) ? 3 : 1;
  e->r.assign(rows, 0.0);
  e->J.assign(rows * 6, 0.0);
  for (double& v : e->g) v = 0;
  double P[12]; plus_jacobian(x, P);
  Jet q[4], t[3];
  for (int k = 0; k < 4; ++k) q[k] = Jet::var(x[k], k);
  for (int k = 0; k < 3; ++k) t[k] = Jet::var(x[4 + k], 4 + k);
  size_t row = 0;
  for (const Factor& f : fs) {
    Jet r[3];
    const int nr = eval_factor<Jet>(f, q, t, r);
    double s = 0; for (int k = 0; k < nr; ++k) s += r[k].a * r[k].a;
    double rho[3]; huber(0.1, s, rho);
    cost += 0.5 * rho[0];
    const double sr = sqrt(rho[1]);
    for (int k = 0; k < nr; ++k, ++row) {
      double* Jr = &e->J[row * 6];
      for (int c = 0; c < 3; ++c) {
        double acc = 0;
        for (int m = 0; m < 4; ++m) acc += r[k].v[m] * P[m * 3 + c];
        Jr[c] = acc * sr;
      }
      for (int c = 0; c < 3; ++c) Jr[3 + c] = r[k].v[4 + c] * sr;
      e->r[row] = r[k].a * sr;
      for (int c = 0; c < 6; ++c) e->g[c] += Jr[c] * e->r[row];
    }
  }
  e->cost = cost;
}

void evaluate_normal_eq(const std::vector<Factor>& f, const double x[7], double* cost, double H[36], double g[6]) {
  Eval e; evaluate(f, x, true, &e);
  *cost = e.cost;
  for (int i = 0; i < 36; ++i) H[i] = 0;
  const size_t m = e.r.size();
  for (size_t r = 0; r < m; ++r)
    for (int i = 0; i < 6; ++i)
      for (int j = 0; j < 6; ++j) H[i * 6 + j] += e.J[r * 6 + i] * e.J[r * 6 + j];
  for (int i = 0; i < 6; ++i) g[i] = e.g[i];
}

// DenseQRSolver: min || [J; D] y - [r; 0] ||  via Householder QR (Eigen householderQr().solve)
static bool dense_qr_solve(const std::vector<double>& J, const std::vector<double>& r, const double D[6], double y[6]) {
  const size_t m = r.size(), M = m + 6;
  std::vector<double> A(M * 6, 0.0), b(M, 0.0);
  std::copy(J.begin(), J.end(), A.begin());
  std::copy(r.begin(), r.end(), b.begin());
  for (int c = 0; c < 6; ++c) A[(m + c) * 6 + c] = D[c];
  double Rd[6];
  for (int k = 0; k < 6; ++k) {
    double tail2 = 0;
    for (size_t i = k + 1; i < M; ++i) tail2 += A[i * 6 + k] * A[i * 6 + k];
    const double c0 = A[k * 6 + k];
    double beta, tau;
    if (tail2 <= DBL_MIN) { tau = 0; beta = c0; }
    else {
      beta = sqrt(c0 * c0 + tail2);
      if (c0 >= 0) beta = -beta;
      const double inv = 1.0 / (c0 - beta);
      for (size_t i = k + 1; i < M; ++i) A[i * 6 + k] *= inv;  // essential part of v
      tau = (beta - c0) / beta;
    }
    if (tau != 0) {
      for (int j = k + 1; j < 6; ++j) {
        double s = A[k * 6 + j];
        for (size_t i = k + 1; i < M; ++i) s += A[i * 6 + k] * A[i * 6 + j];
        s *= tau;
        A[k * 6 + j] -= s;
        for (size_t i = k + 1; i < M; ++i) A[i * 6 + j] -= s * A[i * 6 + k];
      }
      double s = b[k];
      for (size_t i = k + 1; i < M; ++i) s += A[i * 6 + k] * b[i];
      s *= tau;
      b[k] -= s;
      for (size_t i = k + 1; i < M; ++i) b[i] -= s * A[i * 6 + k];
    }
    Rd[k] = beta;
  }
  for (int k = 5; k >= 0; --k) {
    double s = b[k];
    for (int j = k + 1; j < 6; ++j) s -= A[k * 6 + j] * y[j];
    y[k] = s / Rd[k];
  }
  for (int k = 0; k < 6; ++k) if (!std::isfinite(y[k])) return false;
  return true;
}

// TrustRegionMinimizer::Minimize + LevenbergMarquardtStrategy (Ceres 2.0.0).
void ceres_solve(const std::vector<Factor>& fs, double x_user[7], SolveLog* log) {
  SolveLog dummy; if (!log) log = &dummy;
  *log = SolveLog();
  if (fs.empty()) return;  // no residual blocks: parameter blocks are removed, nothing to do
  const int kMaxIter = 4;
  double radius = 1e4, decrease_factor = 2.0;
  const double min_radius = 1e-32, max_radius = 1e16;
  const double min_diag = 1e-6, max_diag = 1e32;
  const double min_relative_decrease = 1e-3;
  const double function_tol = 1e-6, gradient_tol = 1e-10, parameter_tol = 1e-8;
  bool reuse_diagonal = false;

  double x[7]; for (int k = 0; k < 7; ++k) x[k] = x_user[k];
  double x_norm = 0; for (int k = 0; k < 7; ++k) x_norm += x[k] * x[k]; x_norm = sqrt(x_norm);

  Eval e; evaluate(fs, x, true, &e);  // IterationZero
  const size_t m = e.r.size();
  double scale[6];
  for (int c = 0; c < 6; ++c) {
    double n2 = 0; for (size_t r = 0; r < m; ++r) n2 += e.J[r * 6 + c] * e.J[r * 6 + c];
    scale[c] = 1.0 / (1.0 + sqrt(n2));
  }
  auto scale_columns = [&](Eval& ev) { for (size_t r = 0; r < m; ++r) for (int c = 0; c < 6; ++c) ev.J[r * 6 + c] *= scale[c]; };
  scale_columns(e);
  auto grad_max_norm = [&](const Eval& ev) {
    double ng[6]; for (int c = 0; c < 6; ++c) ng[c] = -ev.g[c];
    double xp[7]; plus(x, ng, xp);
    double mx = 0; for (int k = 0; k < 7; ++k) mx = std::max(mx, fabs(x[k] - xp[k]));
    return mx;
  };
  double x_cost = e.cost, minimum_cost = e.cost;
  log->initial_cost = x_cost;
  double diag[6];
  bool last_successful = false;
  double gmax = grad_max_norm(e);

  for (int iter = 1; iter <= kMaxIter; ++iter) {
    // FinalizeIterationAndCheckIfMinimizerCanContinue of the previous iteration
    if (last_successful && gmax <= gradient_tol) break;
    if (radius < min_radius) break;
    log->iterations = iter;
    last_successful = false;
    // LevenbergMarquardtStrategy::ComputeStep
    if (!reuse_diagonal) {
      for (int c = 0; c < 6; ++c) {
        double n2 = 0; for (size_t r = 0; r < m; ++r) n2 += e.J[r * 6 + c] * e.J[r * 6 + c];
        diag[c] = std::min(std::max(n2, min_diag), max_diag);
      }
    }
    double D[6]; for (int c = 0; c < 6; ++c) D[c] = sqrt(diag[c] / radius);
    double y[6];
    const bool ok = dense_qr_solve(e.J, e.r, D, y);
    reuse_diagonal = true;
    double step[6]; for (int c = 0; c < 6; ++c) step[c] = -y[c];
    double model_cost_change = 0;
    if (ok) {
      // -(J s)' (r + J s / 2)
      for (size_t r = 0; r < m; ++r) {
        double js = 0; for (int c = 0; c < 6; ++c) js += e.J[r * 6 + c] * step[c];
        model_cost_change -= js * (e.r[r] + js / 2.0);
      }
    }
    if (!ok || !(model_cost_change > 0.0)) {  // HandleInvalidStep -> StepRejected(0)
      radius /= decrease_factor; decrease_factor *= 2.0;
      log->cost_trace.push_back(x_cost);
      continue;
    }
    double delta[6]; for (int c = 0; c < 6; ++c) delta[c] = step[c] * scale[c];
    double xc[7]; plus(x, delta, xc);
    Eval ec; evaluate(fs, xc, false, &ec);
    const double cand_cost = ec.cost;
    // ParameterToleranceReached
    double sn = 0; for (int k = 0; k < 7; ++k) sn += (x[k] - xc[k]) * (x[k] - xc[k]); sn = sqrt(sn);
    if (sn <= parameter_tol * (x_norm + parameter_tol)) break;
    // FunctionToleranceReached
    const double cost_change = x_cost - cand_cost;
    if (fabs(cost_change) <= function_tol * x_cost) break;
    const double rho = cost_change / model_cost_change;
    if (rho > min_relative_decrease) {  // HandleSuccessfulStep
      for (int k = 0; k < 7; ++k) x[k] = xc[k];
      x_norm = 0; for (int k = 0; k < 7; ++k) x_norm += x[k] * x[k]; x_norm = sqrt(x_norm);
      evaluate(fs, x, true, &e);
      scale_columns(e);
      x_cost = e.cost;
      gmax = grad_max_norm(e);
      radius = radius / std::max(1.0 / 3.0, 1.0 - pow(2.0 * rho - 1.0, 3));
      radius = std::min(max_radius, radius);
      decrease_factor = 2.0;
      reuse_diagonal = false;
      last_successful = true;
      log->successful++;
      if (x_cost < minimum_cost) { minimum_cost = x_cost; for (int k = 0; k < 7; ++k) x_user[k] = x[k]; }
    } else {  // HandleUnsuccessfulStep
      radius /= decrease_factor; decrease_factor *= 2.0;
    }
    log->cost_trace.push_back(x_cost);
  }
  log->final_cost = minimum_cost;
}

}  // namespace vo
