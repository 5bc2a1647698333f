// lm_solver.cu -- ceres::Solve as the reference configures it (laser_odometry.cpp:500-509,
// laser_mapping.cpp:710-717), entirely on the device.
//
// One kernel per Levenberg-Marquardt evaluation: every thread evaluates residuals and
// analytic Jacobians of its factors (lidarFactor.hpp:14-144) in f64, applies the Huber
// corrector, and the CTA reduces the 21 + 6 + 1 numbers of the robustified normal
// equations (upper-triangular J^T J, J^T r, cost) with warp shuffles.  The last CTA to
// finish adds the per-CTA partials in a fixed order (deterministic) and runs the
// trust-region bookkeeping of Ceres 2.0.0 (Jacobi scaling, LM diagonal clamp, 6x6
// solve, Plus on the quaternion manifold, step acceptance, radius update, the three
// tolerances) in one thread, leaving the next candidate in the state struct.  The host
// never sees the 6x6 system; it only queues 1 + 4 evaluation kernels per solve.
#include <float.h>
#include <cooperative_groups.h>
#include "common.cuh"

namespace cg = cooperative_groups;

#define LM_BLOCK 256
#define LM_MAX_BLOCKS 512

__device__ unsigned int g_lm_counter_dummy;

struct FactorRow { double r[3]; double J[3][6]; int nr; };

// residual + local (6-dof) Jacobian of one factor at pose x = {q (xyzw), t}
__device__ __forceinline__ void lm_factor(const double* __restrict__ f, const double* __restrict__ x, FactorRow& o) {
  const int type = (int)f[0];
  double rp[3];
  vl_qrot(x, f[1], f[2], f[3], rp);
  const double lp[3] = {rp[0] + x[4], rp[1] + x[5], rp[2] + x[6]};
  if (type == 0) {  // LidarEdgeFactor (LF.hpp:22-50): r = ((lp-a) x (lp-b)) / |a-b|
    const double a[3] = {f[4], f[5], f[6]}, b[3] = {f[7], f[8], f[9]};
    const double u[3] = {lp[0] - a[0], lp[1] - a[1], lp[2] - a[2]}, v[3] = {lp[0] - b[0], lp[1] - b[1], lp[2] - b[2]};
    const double nu[3] = {u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0]};
    const double de[3] = {a[0] - b[0], a[1] - b[1], a[2] - b[2]};
    const double inv = 1.0 / sqrt(de[0] * de[0] + de[1] * de[1] + de[2] * de[2]);
    o.nr = 3;
    o.r[0] = nu[0] * inv; o.r[1] = nu[1] * inv; o.r[2] = nu[2] * inv;
    // d r / d lp = [w]x with w = (b - a) / |a - b|;  d lp / d delta = -2 [rp]x;  d lp / d t = I
    // => J_rot = -2 [w]x [rp]x = -2 (rp w^T - (w . rp) I),  J_t = [w]x
    const double w[3] = {-de[0] * inv, -de[1] * inv, -de[2] * inv};
    const double wr = w[0] * rp[0] + w[1] * rp[1] + w[2] * rp[2];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
      for (int cidx = 0; cidx < 3; ++cidx) o.J[i][cidx] = -2.0 * (i == cidx ? rp[i] * w[cidx] - wr : rp[i] * w[cidx]);
    }
    o.J[0][3] = 0.0;   o.J[0][4] = -w[2]; o.J[0][5] = w[1];
    o.J[1][3] = w[2];  o.J[1][4] = 0.0;   o.J[1][5] = -w[0];
    o.J[2][3] = -w[1]; o.J[2][4] = w[0];  o.J[2][5] = 0.0;
  } else {
    double n[3];
    o.nr = 1;
    if (type == 1) {  // LidarPlaneFactor (LF.hpp:78-99): r = (lp - j) . ljm_norm
      n[0] = f[7]; n[1] = f[8]; n[2] = f[9];
      o.r[0] = (lp[0] - f[4]) * n[0] + (lp[1] - f[5]) * n[1] + (lp[2] - f[6]) * n[2];
    } else {          // LidarPlaneNormFactor (LF.hpp:121-133): r = n . lp + d
      n[0] = f[4]; n[1] = f[5]; n[2] = f[6];
      o.r[0] = (n[0] * lp[0] + n[1] * lp[1] + n[2] * lp[2]) + f[7];
    }
    // J_rot = -2 (n x rp)^T, J_t = n^T
    o.J[0][0] = -2.0 * (n[1] * rp[2] - n[2] * rp[1]);
    o.J[0][1] = -2.0 * (n[2] * rp[0] - n[0] * rp[2]);
    o.J[0][2] = -2.0 * (n[0] * rp[1] - n[1] * rp[0]);
    o.J[0][3] = n[0]; o.J[0][4] = n[1]; o.J[0][5] = n[2];
  }
}

// DISTORTION == true (LO.cpp:368-372, 472-476): the functors evaluate lp = Identity.slerp(s, q) * cp + s t (LF.hpp:29-36, 86-93),
// and Ceres differentiates THROUGH the slerp with respect to the four quaternion coefficients before contracting with the
// 4x3 plus-Jacobian.  Analytic form of that chain (validated against the oracle's dual numbers, which run Eigen's slerp on
// ceres::Jet-like duals): q_s = sc0 e_w + sc1 q with sc0, sc1 functions of w only (theta = acos|w|);
//   V(U, W) = p + 2 W (U x p) + 2 U x (U x p)   (Eigen's quaternion * vector, no normalisation)
//   dV/dW = 2 (U x p),   dV/dU = -2 W [p]x - 2 [U x p]x - 2 [U]x [p]x
//   G = dV/dq (3x4): columns x,y,z = dV/dU * sc1; column w = dV/dU u sc1' + dV/dW (sc0' + sc1 + w sc1')
//   local Jacobian = D [ G P(q) | s I ] with D = d r / d lp ([w_dir]x for the edge factor, n^T for the plane factor).
__device__ void lm_factor_deskew(const double* __restrict__ f, const double s, const double* __restrict__ x, FactorRow& o) {
  const int type = (int)f[0];
  const double p[3] = {f[1], f[2], f[3]};
  const double qx = x[0], qy = x[1], qz = x[2], qw = x[3];
  const double one = 1.0 - DBL_EPSILON;
  const double absD = fabs(qw);
  double sc0, sc1, dsc0 = 0.0, dsc1 = 0.0;
  if (absD >= one) { sc0 = 1.0 - s; sc1 = s; }
  else {
    const double th = acos(absD), sth = sin(th), cth = cos(th);
    const double a0 = (1.0 - s) * th, a1 = s * th;
    const double s0 = sin(a0), s1 = sin(a1);
    sc0 = s0 / sth; sc1 = s1 / sth;
    const double dth = (qw < 0.0 ? 1.0 : -1.0) / sqrt(1.0 - absD * absD);  // d acos|w| / dw
    dsc0 = (((1.0 - s) * cos(a0)) * sth - s0 * cth) / (sth * sth) * dth;
    dsc1 = ((s * cos(a1)) * sth - s1 * cth) / (sth * sth) * dth;
  }
  if (qw < 0.0) { sc1 = -sc1; dsc1 = -dsc1; }
  const double qs[4] = {sc1 * qx, sc1 * qy, sc1 * qz, sc0 + sc1 * qw};
  double rp[3];
  vl_qrot(qs, p[0], p[1], p[2], rp);
  const double lp[3] = {rp[0] + s * x[4], rp[1] + s * x[5], rp[2] + s * x[6]};
  const double U[3] = {qs[0], qs[1], qs[2]}, W = qs[3];
  const double up[3] = {U[1] * p[2] - U[2] * p[1], U[2] * p[0] - U[0] * p[2], U[0] * p[1] - U[1] * p[0]};  // U x p
  // A = dV/dU = -2 (W [p]x + [U x p]x + [U]x [p]x);  [U]x [p]x = p U^T - (U . p) I
  const double udp = U[0] * p[0] + U[1] * p[1] + U[2] * p[2];
  double A[3][3];
  const double m[3] = {W * p[0] + up[0], W * p[1] + up[1], W * p[2] + up[2]};  // [m]x = W [p]x + [U x p]x
  A[0][0] = -2.0 * (p[0] * U[0] - udp);            A[0][1] = -2.0 * (-m[2] + p[0] * U[1]);        A[0][2] = -2.0 * (m[1] + p[0] * U[2]);
  A[1][0] = -2.0 * (m[2] + p[1] * U[0]);           A[1][1] = -2.0 * (p[1] * U[1] - udp);          A[1][2] = -2.0 * (-m[0] + p[1] * U[2]);
  A[2][0] = -2.0 * (-m[1] + p[2] * U[0]);          A[2][1] = -2.0 * (m[0] + p[2] * U[1]);         A[2][2] = -2.0 * (p[2] * U[2] - udp);
  double G[3][4];
  const double kw = dsc0 + sc1 + qw * dsc1;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    G[i][0] = A[i][0] * sc1; G[i][1] = A[i][1] * sc1; G[i][2] = A[i][2] * sc1;
    G[i][3] = (A[i][0] * qx + A[i][1] * qy + A[i][2] * qz) * dsc1 + 2.0 * up[i] * kw;
  }
  // L = G P(q), P rows: [w, z, -y; -z, w, x; y, -x, w; -x, -y, -z]  (EigenQuaternionParameterization)
  double L[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    L[i][0] = G[i][0] * qw - G[i][1] * qz + G[i][2] * qy - G[i][3] * qx;
    L[i][1] = G[i][0] * qz + G[i][1] * qw - G[i][2] * qx - G[i][3] * qy;
    L[i][2] = -G[i][0] * qy + G[i][1] * qx + G[i][2] * qw - G[i][3] * qz;
  }
  if (type == 0) {
    const double a[3] = {f[4], f[5], f[6]}, b[3] = {f[7], f[8], f[9]};
    const double u[3] = {lp[0] - a[0], lp[1] - a[1], lp[2] - a[2]}, v[3] = {lp[0] - b[0], lp[1] - b[1], lp[2] - b[2]};
    const double nu[3] = {u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0]};
    const double de[3] = {a[0] - b[0], a[1] - b[1], a[2] - b[2]};
    const double inv = 1.0 / sqrt(de[0] * de[0] + de[1] * de[1] + de[2] * de[2]);
    o.nr = 3;
    o.r[0] = nu[0] * inv; o.r[1] = nu[1] * inv; o.r[2] = nu[2] * inv;
    const double w[3] = {-de[0] * inv, -de[1] * inv, -de[2] * inv};
    const double D[3][3] = {{0.0, -w[2], w[1]}, {w[2], 0.0, -w[0]}, {-w[1], w[0], 0.0}};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
      for (int cidx = 0; cidx < 3; ++cidx) {
        o.J[k][cidx] = D[k][0] * L[0][cidx] + D[k][1] * L[1][cidx] + D[k][2] * L[2][cidx];
        o.J[k][3 + cidx] = D[k][cidx] * s;
      }
    }
  } else {  // type 1 (LidarPlaneFactor); the plane-norm factor of the mapping stage has no s (LF.hpp:121-133)
    const double n[3] = {f[7], f[8], f[9]};
    o.nr = 1;
    o.r[0] = (lp[0] - f[4]) * n[0] + (lp[1] - f[5]) * n[1] + (lp[2] - f[6]) * n[2];
#pragma unroll
    for (int cidx = 0; cidx < 3; ++cidx) {
      o.J[0][cidx] = n[0] * L[0][cidx] + n[1] * L[1][cidx] + n[2] * L[2][cidx];
      o.J[0][3 + cidx] = n[cidx] * s;
    }
  }
}

// EigenQuaternionParameterization::Plus on q, plain addition on t
__device__ void lm_plus(const double x[7], const double d[6], double o[7]) {
  const double n = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  if (n > 0.0) {
    double sn, cs;
    sincos(n, &sn, &cs);
    const double s = sn / n;
    const double dq[4] = {s * d[0], s * d[1], s * d[2], cs};
    vl_qmul(dq, x, o);
  } else { o[0] = x[0]; o[1] = x[1]; o[2] = x[2]; o[3] = x[3]; }
  o[4] = x[4] + d[3]; o[5] = x[5] + d[4]; o[6] = x[6] + d[5];
}

__device__ __forceinline__ int tri(int i, int j) { return i <= j ? i * 6 - i * (i - 1) / 2 + (j - i) : j * 6 - j * (j - 1) / 2 + (i - j); }

// Solve A y = b for symmetric positive definite 6x6 A by a right-looking, fully unrolled L D L^T
// (unit lower L): no square roots, and the six reciprocals are the only long-latency operations on
// the dependent chain (f64 sqrt ~ 90 and div ~ 130 cycles on B200, measured).  Returns false on breakdown.
__device__ __forceinline__ bool lm_ldl6(double A[6][6], const double b[6], double y[6]) {
  double inv[6];
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    const double d = A[j][j];
    ok = ok && (d > 0.0);
    inv[j] = __drcp_rn(d);
    double l[6];
#pragma unroll
    for (int i = j + 1; i < 6; ++i) l[i] = A[i][j] * inv[j];
#pragma unroll
    for (int i = j + 1; i < 6; ++i) {
#pragma unroll
      for (int k = j + 1; k <= i; ++k) A[i][k] -= l[i] * A[k][j];  // A[k][j] still holds L[k][j] * d
    }
#pragma unroll
    for (int i = j + 1; i < 6; ++i) A[i][j] = l[i];
  }
  double z[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double t = b[i];
#pragma unroll
    for (int k = 0; k < i; ++k) t -= A[i][k] * z[k];
    z[i] = t;
  }
#pragma unroll
  for (int i = 5; i >= 0; --i) {
    double t = z[i] * inv[i];
#pragma unroll
    for (int k = i + 1; k < 6; ++k) t -= A[k][i] * y[k];
    y[i] = t;
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) ok = ok && isfinite(y[i]);
  return ok;
}

__device__ double lm_gmax(const double x[7], const double g[6]) {
  // The translation part of x - Plus(x, -g) is g[3..5] itself; when that alone exceeds the gradient
  // tolerance the max norm cannot pass the 1e-10 test and the quaternion part need not be formed.
  const double gt = fmax(fabs(g[3]), fmax(fabs(g[4]), fabs(g[5])));
  if (gt > 1e-6) return gt;
  double ng[6], xp[7];
#pragma unroll
  for (int k = 0; k < 6; ++k) ng[k] = -g[k];
  lm_plus(x, ng, xp);
  double m = 0;
#pragma unroll
  for (int k = 0; k < 7; ++k) m = fmax(m, fabs(x[k] - xp[k]));
  return m;
}

#define LG_TRACE(slot) do { if (tr && lane == 0) tr[slot] = clock64(); } while (0)
// ---- TrustRegionMinimizer bookkeeping, warp-cooperative -------------------------------------------------
// Runs on warp 0 of every CTA after each evaluation.  f64 sqrt / div are ~90 / ~130-cycle dependent
// sequences on B200 and a lone thread issues a dependent f64 op only every 8 cycles, so a one-thread
// version of this costs ~3 us per evaluation -- as much as the evaluation itself.  Here the 21 + 6
// entries of the normal equations stay distributed over the lanes that summed them (lane t < 21 owns
// entry t = tri(i, j), lane 21 + k owns g[k], lane 27 the cost): Jacobi scales, the scaled system, the
// LM diagonal and the model cost change are computed one entry per lane; only the 6x6 L D L^T itself
// (a strictly dependent chain) is replicated on every lane.  Scalars are replicated: every lane takes
// the same branches, and every CTA of the cluster computes the same bits.
struct LmWarp {
  double x[7], xc[7];
  double sc[6];                      // Jacobi scaling, fixed at iteration 0
  double cost, min_cost, initial_cost, model_cost_change, radius, decrease_factor, x_norm, gmax;
  int reuse_diagonal, done, iter, last_successful;
  double H, Hs, diag;                // this lane's entry: accepted-point system (unscaled, scaled), LM diagonal
  int li, lj;                        // this lane's (row, column); vector lanes: (k, -1)
};

__device__ __forceinline__ double lm_sel6(const double (&v)[6], int k) {
  double r = v[0];
#pragma unroll
  for (int q = 1; q < 6; ++q) r = (k == q) ? v[q] : r;
  return r;
}

__device__ __forceinline__ double lm_gmax_warp(const LmWarp& S) {
  double g[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) g[k] = __shfl_sync(0xffffffffu, S.H, 21 + k);
  return lm_gmax(S.x, g);
}

// e: this lane's entry of the evaluation (at S.x when iter == 0, at S.xc otherwise).  best[7]: shared memory.
__device__ __forceinline__ void lm_logic_warp(LmWarp& S, const double e, const int lane, double* __restrict__ best, long long* tr) {
  const double ftol = 1e-6, gtol = 1e-10, ptol = 1e-8, min_rel = 1e-3;
  const double min_radius = 1e-32, max_radius = 1e16, min_diag = 1e-6, max_diag = 1e32;
  const int kMaxIter = 4;
  const double cost = __shfl_sync(0xffffffffu, e, 27);
  const bool isDiag = S.li == S.lj;
  if (S.iter == 0) {  // IterationZero
    S.H = e;
    S.cost = cost; S.min_cost = cost; S.initial_cost = cost;
    const double s = 1.0 / (1.0 + sqrt(fabs(e)));  // meaningful on the six diagonal lanes (0, 6, 11, 15, 18, 20)
#pragma unroll
    for (int k = 0; k < 6; ++k) S.sc[k] = __shfl_sync(0xffffffffu, s, tri(k, k));
    double xn = 0;
#pragma unroll
    for (int k = 0; k < 7; ++k) xn += S.x[k] * S.x[k];
    S.x_norm = sqrt(xn);
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < 7; ++k) best[k] = S.x[k];
    }
    S.gmax = lm_gmax_warp(S);
    S.radius = 1e4; S.decrease_factor = 2.0; S.reuse_diagonal = 0; S.last_successful = 0;
    S.Hs = (S.H * lm_sel6(S.sc, S.li)) * (S.lj >= 0 ? lm_sel6(S.sc, S.lj) : 1.0);
  } else {
    double sn = 0, xn = 0;
#pragma unroll
    for (int k = 0; k < 7; ++k) { sn += (S.x[k] - S.xc[k]) * (S.x[k] - S.xc[k]); xn += S.xc[k] * S.xc[k]; }
    // the four long operations of this branch are independent of each other up to the radius update
    sn = sqrt(sn);
    const double xcn = sqrt(xn);
    const double cost_change = S.cost - cost;
    const double rho = cost_change / S.model_cost_change;
    if (sn <= ptol * (S.x_norm + ptol)) { S.done = 1; return; }       // ParameterToleranceReached
    if (fabs(cost_change) <= ftol * S.cost) { S.done = 1; return; }   // FunctionToleranceReached
    if (rho > min_rel) {  // HandleSuccessfulStep
#pragma unroll
      for (int k = 0; k < 7; ++k) S.x[k] = S.xc[k];
      S.x_norm = xcn;
      S.H = e;
      S.Hs = (S.H * lm_sel6(S.sc, S.li)) * (S.lj >= 0 ? lm_sel6(S.sc, S.lj) : 1.0);
      S.cost = cost;
      S.gmax = lm_gmax_warp(S);
      const double t3 = 2.0 * rho - 1.0;
      S.radius = fmin(max_radius, S.radius / fmax(1.0 / 3.0, 1.0 - t3 * t3 * t3));
      S.decrease_factor = 2.0; S.reuse_diagonal = 0; S.last_successful = 1;
      if (S.cost < S.min_cost) {
        S.min_cost = S.cost;
        if (lane == 0) {
#pragma unroll
          for (int k = 0; k < 7; ++k) best[k] = S.x[k];
        }
      }
    } else {  // HandleUnsuccessfulStep
      S.radius /= S.decrease_factor; S.decrease_factor *= 2.0; S.last_successful = 0;
    }
  }
  LG_TRACE(9);
  // next trust-region step(s); an invalid step consumes an iteration without an evaluation
  while (true) {
    if (S.last_successful && S.gmax <= gtol) { S.done = 1; return; }
    if (S.radius < min_radius) { S.done = 1; return; }
    if (S.iter >= kMaxIter) { S.done = 1; return; }
    S.iter++;
    S.last_successful = 0;
    if (!S.reuse_diagonal) S.diag = fmin(fmax(S.Hs, min_diag), max_diag);  // used on the diagonal lanes only
    const double invr = 1.0 / S.radius;
    const double a = isDiag ? S.Hs + S.diag * invr : S.Hs;  // D^2 = diag / radius
    double A[6][6], gs[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
#pragma unroll
      for (int j = i; j < 6; ++j) { const double v = __shfl_sync(0xffffffffu, a, tri(i, j)); A[i][j] = v; A[j][i] = v; }
      gs[i] = __shfl_sync(0xffffffffu, a, 21 + i);
    }
    LG_TRACE(10);
    double y[6];
    const bool ok = lm_ldl6(A, gs, y);
    LG_TRACE(11);
    S.reuse_diagonal = 1;
    // model_cost_change = -s'gs - s'Hs s / 2 with s = -y  =  y'gs - y'Hs y / 2, one term per lane
    double term = 0.0;
    if (lane < 21) term = ((isDiag ? -0.5 : -1.0) * S.Hs) * (lm_sel6(y, S.li) * lm_sel6(y, S.lj));
    else if (lane < 27) term = lm_sel6(y, S.li) * S.Hs;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) term += __shfl_xor_sync(0xffffffffu, term, d);  // commutative pairs: the same bits on every lane
    const double mcc = ok ? term : 0.0;
    if (!ok || !(mcc > 0.0)) { S.radius /= S.decrease_factor; S.decrease_factor *= 2.0; continue; }
    S.model_cost_change = mcc;
    double delta[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) delta[k] = -y[k] * S.sc[k];
    LG_TRACE(12);
    lm_plus(S.x, delta, S.xc);
    LG_TRACE(13);
    return;
  }
}

// Huber-corrected contribution of one factor to the 21 + 6 + 1 sums
__device__ __forceinline__ void lm_accumulate(const FactorRow& fr, double* __restrict__ acc) {
  double s = fr.r[0] * fr.r[0];
  if (fr.nr == 3) s = (s + fr.r[1] * fr.r[1]) + fr.r[2] * fr.r[2];
  double rho0, rho1;  // ceres::HuberLoss(0.1)
  if (s > 0.01) { const double r = sqrt(s); rho0 = 2.0 * 0.1 * r - 0.01; rho1 = fmax(DBL_MIN, 0.1 / r); }
  else { rho0 = s; rho1 = 1.0; }
  acc[27] += 0.5 * rho0;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    if (k >= fr.nr) break;
    int t = 0;
#pragma unroll
    for (int a = 0; a < 6; ++a) {
      const double wa = rho1 * fr.J[k][a];
#pragma unroll
      for (int b = a; b < 6; ++b) { acc[t] = fma(wa, fr.J[k][b], acc[t]); ++t; }  // explicit DFMA: the library is built with -fmad=false
      acc[21 + a] = fma(wa, fr.r[k], acc[21 + a]);
    }
  }
}

// One robustified evaluation at xEval (vloam_b200_evaluate: the parity tests compare the normal
// equations themselves with the oracle's).  The last CTA to finish adds the per-CTA partials in order.
__global__ void __launch_bounds__(LM_BLOCK) lm_eval(const double* __restrict__ factors, const int* __restrict__ valid, int nslots,
                                                    const double* __restrict__ xEval, EvalOut* __restrict__ partials,
                                                    EvalOut* __restrict__ evalOut, unsigned int* __restrict__ counter, const double* __restrict__ sArr) {
  VL_PDL_WAIT();

  double x[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) x[k] = xEval[k];
  double acc[28];
#pragma unroll
  for (int k = 0; k < 28; ++k) acc[k] = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nslots; i += gridDim.x * blockDim.x) {
    if (!valid[i]) continue;
    FactorRow fr;
    if (sArr && (int)factors[(size_t)i * 10] != 2) lm_factor_deskew(factors + (size_t)i * 10, sArr[i], x, fr);
    else lm_factor(factors + (size_t)i * 10, x, fr);
    lm_accumulate(fr, acc);
  }
  __shared__ double red[LM_BLOCK / 32][28];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 28; ++k) {
    double v = acc[k];
    for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
    if (lane == 0) red[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < 28) {
    double v = 0;
    for (int w = 0; w < LM_BLOCK / 32; ++w) v += red[w][threadIdx.x];
    partials[blockIdx.x].v[threadIdx.x] = v;
  }
  __shared__ int isLast;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) isLast = (atomicAdd(counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!isLast) return;
  __threadfence();
  if (threadIdx.x < 28) {
    double v = 0;
    for (unsigned b = 0; b < gridDim.x; ++b) v += ((volatile EvalOut*)partials)[b].v[threadIdx.x];
    evalOut->v[threadIdx.x] = v;
  }
  if (threadIdx.x == 0) *counter = 0;
}


// ---- the whole ceres::Solve in ONE launch: a thread-block cluster keeps the iterate on chip ----
// One cluster of 8 (odometry) or 16 (mapping) CTAs.  Every thread loads its <= FPT factors ONCE into
// registers together with their pose-independent parts (|a-b|, (b-a)/|a-b|): cluster.sync invalidates
// L1D, so re-reading the factors at each of the 5 evaluations costs an L2 round trip every time.
// Per evaluation: residuals + Jacobians from registers, a transposed butterfly reduction (31 double
// shuffles per warp instead of 28 x 5), per-CTA partials in shared memory, ONE cluster barrier, then
// EVERY CTA pulls all partials through distributed shared memory, adds them in rank order and runs the
// trust-region bookkeeping redundantly (same inputs, same instructions => the same candidate in every
// CTA).  Nothing has to be published back, so the second cluster barrier and the remote read of the
// next evaluation point of the previous design are gone; the partials are double-buffered by
// evaluation parity, which the single barrier is enough to protect.
#ifndef LMC_MIN_CTAS
#define LMC_MIN_CTAS 1  // (2 caps the kernel at 128 registers so that a solver CTA fits beside another stream's resident CTA; the spills
#endif                  // cost more -- 130 -> 160 us of solves per frame -- than the co-residency gains)
#define LMC_CTAS 16  // one cluster of the non-portable maximum size: the f64 pipes of 16 SMs

struct LmFactorReg {  // one factor with its pose-independent parts hoisted
  int type;           // -1: empty / invalid slot
  double p[3];        // current point
  double a[3];        // edge: a; plane: j; plane-norm: n
  double b[3];        // edge: b; plane: ljm_norm; plane-norm: {d, -, -}
  double inv;         // edge: 1 / |a - b|
  double w[3];        // edge: (b - a) / |a - b|
};

__device__ __forceinline__ void lm_factor_load(const double* __restrict__ f, bool ok, LmFactorReg& o) {
  // the ten loads are issued unconditionally (the slot array is allocated up to the host bound), so they overlap
  // the loads of the slot count and of the validity flag instead of waiting for them
  const double t = f[0];
  o.p[0] = f[1]; o.p[1] = f[2]; o.p[2] = f[3];
  o.a[0] = f[4]; o.a[1] = f[5]; o.a[2] = f[6];
  o.b[0] = f[7]; o.b[1] = f[8]; o.b[2] = f[9];
  o.type = ok ? (int)t : -1;
  o.inv = 0; o.w[0] = o.w[1] = o.w[2] = 0;
  if (o.type == 0) {
    const double de[3] = {o.a[0] - o.b[0], o.a[1] - o.b[1], o.a[2] - o.b[2]};
    o.inv = 1.0 / sqrt(de[0] * de[0] + de[1] * de[1] + de[2] * de[2]);
    o.w[0] = -de[0] * o.inv; o.w[1] = -de[1] * o.inv; o.w[2] = -de[2] * o.inv;
  }
}

// same arithmetic, in the same order, as lm_factor
__device__ __forceinline__ void lm_factor_eval(const LmFactorReg& f, const double* __restrict__ x, FactorRow& o) {
  double rp[3];
  vl_qrot(x, f.p[0], f.p[1], f.p[2], rp);
  const double lp[3] = {rp[0] + x[4], rp[1] + x[5], rp[2] + x[6]};
  if (f.type == 0) {
    const double u[3] = {lp[0] - f.a[0], lp[1] - f.a[1], lp[2] - f.a[2]}, v[3] = {lp[0] - f.b[0], lp[1] - f.b[1], lp[2] - f.b[2]};
    const double nu[3] = {u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0]};
    o.nr = 3;
    o.r[0] = nu[0] * f.inv; o.r[1] = nu[1] * f.inv; o.r[2] = nu[2] * f.inv;
    const double wr = f.w[0] * rp[0] + f.w[1] * rp[1] + f.w[2] * rp[2];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
      for (int cidx = 0; cidx < 3; ++cidx) o.J[i][cidx] = -2.0 * (i == cidx ? rp[i] * f.w[cidx] - wr : rp[i] * f.w[cidx]);
    }
    o.J[0][3] = 0.0;     o.J[0][4] = -f.w[2]; o.J[0][5] = f.w[1];
    o.J[1][3] = f.w[2];  o.J[1][4] = 0.0;     o.J[1][5] = -f.w[0];
    o.J[2][3] = -f.w[1]; o.J[2][4] = f.w[0];  o.J[2][5] = 0.0;
  } else {
    double n[3];
    o.nr = 1;
    if (f.type == 1) {
      n[0] = f.b[0]; n[1] = f.b[1]; n[2] = f.b[2];
      o.r[0] = (lp[0] - f.a[0]) * n[0] + (lp[1] - f.a[1]) * n[1] + (lp[2] - f.a[2]) * n[2];
    } else {
      n[0] = f.a[0]; n[1] = f.a[1]; n[2] = f.a[2];
      o.r[0] = (n[0] * lp[0] + n[1] * lp[1] + n[2] * lp[2]) + f.b[0];
    }
    o.J[0][0] = -2.0 * (n[1] * rp[2] - n[2] * rp[1]);
    o.J[0][1] = -2.0 * (n[2] * rp[0] - n[0] * rp[2]);
    o.J[0][2] = -2.0 * (n[0] * rp[1] - n[1] * rp[0]);
    o.J[0][3] = n[0]; o.J[0][4] = n[1]; o.J[0][5] = n[2];
  }
}

// Sum v[0..32) over the warp; afterwards v[0] of lane L holds the total of entry L.  Each level halves
// the entries a lane is responsible for: 16 + 8 + 4 + 2 + 1 = 31 double shuffles.
__device__ __forceinline__ void lm_warp_transpose_reduce(double (&v)[32], int lane) {
#pragma unroll
  for (int h = 16; h >= 1; h >>= 1) {
    const bool up = (lane & h) != 0;
#pragma unroll
    for (int k = 0; k < h; ++k) {
      const double lo = v[k], hi = v[k + h];
      const double send = up ? lo : hi, keep = up ? hi : lo;
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, h);
    }
  }
}

#ifndef LM_TRACE_IT
#define LM_TRACE_IT 1  // which evaluation the phase stamps describe (0 = the first, which also waits for the factor loads)
#endif
#define LM_TRACE(slot) do { if (trace && rank == 0 && threadIdx.x == 0) trace[slot] = clock64(); } while (0)

// FPT: factor slots held in registers per thread (0 = read them from memory at every evaluation).
// Slot i belongs to CTA i % 16 (thread (i / 16) % THREADS): the edge factors, which cost three times a
// plane factor and sit at the front of the slot array, are spread evenly over the CTAs.
// NCTA: CTAs of the cluster -- 16 (non-portable maximum: a whole GPC) for the mapping stage's ~7000 factors, 8 for the odometry's ~1800
// (a cluster of 8 is placed as soon as 8 SMs of a GPC are free: the look-ahead odometry runs beside the mapping stage's wide kernels).
template <int FPT, int THREADS, int NCTA>
__global__ void __launch_bounds__(THREADS, LMC_MIN_CTAS)
lm_solve_cluster(const double* __restrict__ factors, const int* __restrict__ valid, int nslotsBound, const int* __restrict__ d_nslots,
                 double* __restrict__ x_inout, LmSolveState* __restrict__ st_out, long long* __restrict__ trace, const double* __restrict__ sArr) {
  VL_PDL_WAIT(); vl_chain_stamp(NCTA == 8 ? 14 : 4);

  cg::cluster_group cluster = cg::this_cluster();
  const unsigned rank = cluster.block_rank();
  // raw slots first: these loads do not depend on the slot count or the validity flags
  double raw[FPT > 0 ? FPT : 1][10];
  int vflag[FPT > 0 ? FPT : 1];
  if (FPT > 0) {
#pragma unroll
    for (int j = 0; j < FPT; ++j) {
      const int i = (j * THREADS + threadIdx.x) * NCTA + (int)rank;
      const bool inb = i < nslotsBound;
      vflag[j] = inb ? valid[i] : 0;
#pragma unroll
      for (int q = 0; q < 10; ++q) raw[j][q] = inb ? factors[(size_t)i * 10 + q] : 0.0;
    }
  }
  const int nslots = d_nslots ? min(nslotsBound, *d_nslots) : nslotsBound;
  if (nslots <= 0) return;  // no residual blocks: Ceres leaves the parameters untouched (uniform over the cluster)
  __shared__ double part[2][28];
  __shared__ double red[THREADS / 32][28];
  __shared__ double gath[NCTA][28];
  __shared__ double xnext[8];
  __shared__ double best[7];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  LM_TRACE(0);
  double x[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) x[k] = x_inout[k];
  LmWarp S;  // live in warp 0 only
#pragma unroll
  for (int k = 0; k < 7; ++k) { S.x[k] = x[k]; S.xc[k] = x[k]; }
  S.iter = 0; S.done = 0; S.initial_cost = 0; S.cost = 0; S.min_cost = 0; S.reuse_diagonal = 0; S.last_successful = 0;
  S.model_cost_change = 0; S.radius = 1e4; S.decrease_factor = 2.0; S.x_norm = 0; S.gmax = 0; S.H = 0; S.Hs = 0; S.diag = 0;
#pragma unroll
  for (int k = 0; k < 6; ++k) S.sc[k] = 1.0;
  S.li = 0; S.lj = -1;
  if (lane < 21) { int t = lane, i = 0; while (t >= 6 - i) { t -= 6 - i; ++i; } S.li = i; S.lj = i + t; }
  else if (lane < 27) S.li = lane - 21;
  // the host picks FPT from a hint (last frame's count); the actual count decides here, uniformly over the cluster
  const bool inreg = FPT > 0 && nslots <= FPT * THREADS * NCTA;
  LmFactorReg fr_[FPT > 0 ? FPT : 1];
  if (inreg) {
#pragma unroll
    for (int j = 0; j < FPT; ++j) {
      const int i = (j * THREADS + threadIdx.x) * NCTA + (int)rank;
      lm_factor_load(raw[j], i < nslots && vflag[j] != 0, fr_[j]);
    }
  }
  LM_TRACE(14);
  for (int it = 0; it < 5; ++it) {
    if (it == LM_TRACE_IT) LM_TRACE(1);
    double acc[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) acc[k] = 0.0;
    if (inreg) {
#pragma unroll
      for (int j = 0; j < FPT; ++j) {
        if (fr_[j].type < 0) continue;
        FactorRow fr;
        lm_factor_eval(fr_[j], x, fr);
        lm_accumulate(fr, acc);
      }
    } else {
      for (int i = threadIdx.x * NCTA + (int)rank; i < nslots; i += THREADS * NCTA) {
        if (!valid[i]) continue;
        FactorRow fr;
        if (FPT == 0 && sArr && (int)factors[(size_t)i * 10] != 2) lm_factor_deskew(factors + (size_t)i * 10, sArr[i], x, fr);  // DISTORTION: only the streaming variant
        else lm_factor(factors + (size_t)i * 10, x, fr);
        lm_accumulate(fr, acc);
      }
    }
    if (it == LM_TRACE_IT) LM_TRACE(2);
    lm_warp_transpose_reduce(acc, lane);
    if (lane < 28) red[warp][lane] = acc[0];
    __syncthreads();
    double* mine = part[it & 1];
    if (threadIdx.x < 28) {
      double v = 0;
#pragma unroll
      for (int w = 0; w < THREADS / 32; ++w) v += red[w][threadIdx.x];
      mine[threadIdx.x] = v;
    }
    if (it == LM_TRACE_IT) LM_TRACE(3);
    cluster.sync();  // every CTA's partial of this evaluation is visible cluster-wide
    if (it == LM_TRACE_IT) LM_TRACE(4);
    for (int t = threadIdx.x; t < 28 * NCTA; t += THREADS) {
      const int r = t / 28, k = t - r * 28;
      gath[r][k] = cluster.map_shared_rank(mine, r)[k];
    }
    __syncthreads();
    if (it == LM_TRACE_IT) LM_TRACE(5);
    if (warp == 0) {
      double v = 0;
      if (lane < 28) {
#pragma unroll
        for (int r = 0; r < NCTA; ++r) v += gath[r][lane];  // rank order: deterministic
      }
      lm_logic_warp(S, v, lane, best, (it == LM_TRACE_IT && rank == 0) ? trace : nullptr);
      if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 7; ++k) xnext[k] = S.xc[k];
        xnext[7] = (double)S.done;
      }
    }
    __syncthreads();
    if (it == LM_TRACE_IT) LM_TRACE(6);
    if (xnext[7] != 0.0) break;  // identical in every CTA
#pragma unroll
    for (int k = 0; k < 7; ++k) x[k] = xnext[k];
  }
  LM_TRACE(7);
  cluster.sync();  // nobody may exit while another CTA can still read its partials
  if (rank == 0 && warp == 0) {
    __syncwarp();
    if (lane < 7) x_inout[lane] = best[lane];
    if (lane == 0 && st_out) {  // what the debug getters read (costs, iterations)
      st_out->initial_cost = S.initial_cost; st_out->final_cost = S.min_cost; st_out->min_cost = S.min_cost; st_out->cost = S.cost;
      st_out->iter = S.iter; st_out->done = S.done; st_out->nfactors = nslots; st_out->radius = S.radius;
    }
  }
  LM_TRACE(8);
  vl_chain_stamp(NCTA == 8 ? 24 : 9);  // (CTA 0 leaves: the pose is final)
}

static long long* g_solver_trace = nullptr;  // device buffer of 16 clock64 stamps, allocated on first request
int vl_solver_trace(vloam_b200_ctx* c, long long* out16) {
  if (!g_solver_trace) { VL_CUDA(cudaMalloc(&g_solver_trace, 16 * sizeof(long long))); VL_CUDA(cudaMemset(g_solver_trace, 0, 16 * sizeof(long long))); }
  VL_CUDA(cudaStreamSynchronize(c->stream));
  VL_CUDA(cudaMemcpy(out16, g_solver_trace, 16 * sizeof(long long), cudaMemcpyDeviceToHost));
  return VLOAM_OK;
}

template <int FPT, int THREADS, int NCTA = LMC_CTAS>
static cudaError_t lm_launch(cudaLaunchConfig_t& cfg, const double* cf, const int* cv, int nslots, const int* d_nslots, double* x,
                             LmSolveState* so, long long* trace, const double* sArr = nullptr) {
  cfg.blockDim = dim3(THREADS);
  cfg.gridDim = dim3(NCTA);
  cfg.attrs[0].val.clusterDim.x = NCTA;
  return cudaLaunchKernelEx(&cfg, lm_solve_cluster<FPT, THREADS, NCTA>, cf, cv, nslots, d_nslots, x, so, trace, sArr);
}

// function attributes are per device: set when a context is created on it (vloam_b200_create)
int vl_chain_trace_arm_solver(void* dev) { return cudaMemcpyToSymbol(g_chain_trace, &dev, sizeof(void*)) == cudaSuccess ? VLOAM_OK : VLOAM_E_CUDA; }
int vl_solver_set_attrs(vloam_b200_ctx* c) {
  VL_CUDA(cudaFuncSetAttribute((lm_solve_cluster<1, 256, 16>), cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  VL_CUDA(cudaFuncSetAttribute((lm_solve_cluster<2, 256, 16>), cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  VL_CUDA(cudaFuncSetAttribute((lm_solve_cluster<4, 256, 16>), cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  VL_CUDA(cudaFuncSetAttribute((lm_solve_cluster<0, 256, 16>), cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  // lazy module loading would otherwise charge each kernel's first launch (~1 ms apiece) to the first sweeps
  cudaFuncAttributes fa_;
  VL_CUDA(cudaFuncGetAttributes(&fa_, lm_eval));
  VL_CUDA(cudaFuncGetAttributes(&fa_, (lm_solve_cluster<1, 256, 16>)));
  VL_CUDA(cudaFuncGetAttributes(&fa_, (lm_solve_cluster<2, 256, 16>)));
  VL_CUDA(cudaFuncGetAttributes(&fa_, (lm_solve_cluster<4, 256, 16>)));
  VL_CUDA(cudaFuncGetAttributes(&fa_, (lm_solve_cluster<0, 256, 16>)));
  VL_CUDA(cudaFuncGetAttributes(&fa_, (lm_solve_cluster<1, 256, 8>)));
  VL_CUDA(cudaFuncGetAttributes(&fa_, (lm_solve_cluster<2, 256, 8>)));
  return VLOAM_OK;
}

int vl_solve(vloam_b200_ctx* c, int nslots, const int* d_nslots, double* d_x_inout, double* costs2, int hint, const double* d_s) {
  return vl_solve_buf(c, c->factors.p, c->factorValid.p, nslots, d_nslots, d_x_inout, costs2, hint, d_s, 16);
}

int vl_solve_buf(vloam_b200_ctx* c, const double* cf, const int* cv, int nslots, const int* d_nslots, double* d_x_inout, double* costs2, int hint,
                 const double* d_s, int ncta) {
  cudaStream_t st = VL_STREAM(c);
  if (nslots > 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(LMC_CTAS); cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = LMC_CTAS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 2;
    const bool prof = c->prof_name[0] && vl_prof_match(c, "lm_solve_cluster") && c->prof_n < VL_PROF_MAX;
    if (prof) cudaEventRecord(c->prof_ev[c->prof_n][0], st);
    LmSolveState* so = costs2 ? c->lms : nullptr;
    long long* tr = g_solver_trace;
    // nslots is a host BOUND (buffer capacity when the count still lives on the device); `hint` is the last known
    // actual count.  The variant only decides how many slots a thread can keep in registers: a solve whose
    // actual count exceeds it streams the factors from memory at every evaluation instead.
    const int est = hint > 0 ? min(nslots, hint + hint / 32 + 64) : nslots;  // counts move by a few per cent between sweeps; a miss only costs speed
    if (d_s) VL_CUDA((lm_launch<0, 256>(cfg, cf, cv, nslots, d_nslots, d_x_inout, so, tr, d_s)));  // DISTORTION: factors + s streamed from memory
    // The cluster size fixes the order of the sums (slot i belongs to CTA i % NCTA), so it must not follow the hint: the caller
    // names it -- 8 for the odometry stage, whichever path (plain or look-ahead, with different hints) queues the solve.  The
    // register variants of one cluster size are bit-identical to each other.
    else if (ncta == 8 && est <= 8 * 256) VL_CUDA((lm_launch<1, 256, 8>(cfg, cf, cv, nslots, d_nslots, d_x_inout, so, tr)));
    else if (ncta == 8) VL_CUDA((lm_launch<2, 256, 8>(cfg, cf, cv, nslots, d_nslots, d_x_inout, so, tr)));  // (streams from memory above 4096 slots)
    else if (est <= LMC_CTAS * 256) VL_CUDA((lm_launch<1, 256>(cfg, cf, cv, nslots, d_nslots, d_x_inout, so, tr)));
    else if (est <= LMC_CTAS * 512) VL_CUDA((lm_launch<2, 256>(cfg, cf, cv, nslots, d_nslots, d_x_inout, so, tr)));
    else if (est <= LMC_CTAS * 1024) VL_CUDA((lm_launch<4, 256>(cfg, cf, cv, nslots, d_nslots, d_x_inout, so, tr)));
    else VL_CUDA((lm_launch<0, 256>(cfg, cf, cv, nslots, d_nslots, d_x_inout, so, tr)));
    // algorithmic bytes: the factor slots (80 B + flag) are read once; the 5 evaluations run from registers
    if (prof) { cudaEventRecord(c->prof_ev[c->prof_n][1], st); c->prof_kname[c->prof_n] = "lm_solve_cluster"; c->prof_kbytes[c->prof_n] = 84.0 * min(nslots, max(hint, 1)); c->prof_kstream[c->prof_n] = st;
                c->prof_n++; c->prof_bytes += 84.0 * min(nslots, max(hint, 1)); }
    __atomic_fetch_add(&c->launches, 1LL, __ATOMIC_RELAXED);
  }
  if (costs2) {
    if (nslots > 0) {
      VL_CUDA(cudaMemcpyAsync(c->h_lms, c->lms, sizeof(LmSolveState), cudaMemcpyDeviceToHost, st));
      VL_CUDA(cudaStreamSynchronize(st));
      costs2[0] = c->h_lms->initial_cost; costs2[1] = c->h_lms->final_cost;
    } else { costs2[0] = costs2[1] = 0; }
  }
  return VLOAM_OK;
}

int vl_evaluate_once(vloam_b200_ctx* c, int nslots, const double* d_x, EvalOut* d_out, const double* d_s) {
  const int nb = max(1, min(vl_div_up(nslots, LM_BLOCK), LM_MAX_BLOCKS));
  VL_TRY(vl_reserve(c, c->evalPartials, LM_MAX_BLOCKS));
  unsigned int* counter = reinterpret_cast<unsigned int*>(c->vScalars + 60);
  VL_LAUNCH(lm_eval, nb, LM_BLOCK, 0, c->factors.p, c->factorValid.p, nslots, d_x, c->evalPartials.p, d_out, counter, d_s);
  VL_CUDA(cudaGetLastError());
  return VLOAM_OK;
}
