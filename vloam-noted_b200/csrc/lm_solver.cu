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
    const double n = sqrt(de[0] * de[0] + de[1] * de[1] + de[2] * de[2]);
    o.nr = 3;
    o.r[0] = nu[0] / n; o.r[1] = nu[1] / n; o.r[2] = nu[2] / n;
    // d r / d lp = [w]x with w = (b - a) / n;  d lp / d delta = -2 [rp]x;  d lp / d t = I
    // => J_rot = -2 [w]x [rp]x = -2 (rp w^T - (w . rp) I),  J_t = [w]x
    const double w[3] = {-de[0] / n, -de[1] / n, -de[2] / n};
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

// Solve A y = b for symmetric positive definite 6x6 A (in-place Cholesky, fully unrolled so
// every entry lives in a register); returns false on breakdown.
__device__ __forceinline__ bool lm_chol6(double A[6][6], const double b[6], double y[6]) {
  double inv[6];
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    double s = A[j][j];
#pragma unroll
    for (int k = 0; k < j; ++k) s -= A[j][k] * A[j][k];
    ok = ok && (s > 0.0);
    const double d = sqrt(s);
    A[j][j] = d;
    inv[j] = 1.0 / d;
#pragma unroll
    for (int i = j + 1; i < 6; ++i) {
      double t = A[i][j];
#pragma unroll
      for (int k = 0; k < j; ++k) t -= A[i][k] * A[j][k];
      A[i][j] = t * inv[j];
    }
  }
  double z[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double t = b[i];
#pragma unroll
    for (int k = 0; k < i; ++k) t -= A[i][k] * z[k];
    z[i] = t * inv[i];
  }
#pragma unroll
  for (int i = 5; i >= 0; --i) {
    double t = z[i];
#pragma unroll
    for (int k = i + 1; k < 6; ++k) t -= A[k][i] * y[k];
    y[i] = t * inv[i];
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
  for (int k = 0; k < 6; ++k) ng[k] = -g[k];
  lm_plus(x, ng, xp);
  double m = 0;
  for (int k = 0; k < 7; ++k) m = fmax(m, fabs(x[k] - xp[k]));
  return m;
}

// TrustRegionMinimizer bookkeeping after an evaluation `e` (at st->x when iter == 0, at st->xc otherwise).
__device__ void lm_logic(LmSolveState* st, const double* e) {
  const double ftol = 1e-6, gtol = 1e-10, ptol = 1e-8, min_rel = 1e-3;
  const double min_radius = 1e-32, max_radius = 1e16, min_diag = 1e-6, max_diag = 1e32;
  const int kMaxIter = 4;
  const double cost = e[27];
  if (st->iter == 0) {  // IterationZero
    for (int k = 0; k < 21; ++k) st->H[k] = e[k];
    for (int k = 0; k < 6; ++k) st->g[k] = e[21 + k];
    st->cost = cost; st->min_cost = cost; st->initial_cost = cost;
    for (int k = 0; k < 6; ++k) st->scale[k] = 1.0 / (1.0 + sqrt(st->H[tri(k, k)]));
    double xn = 0; for (int k = 0; k < 7; ++k) { xn += st->x[k] * st->x[k]; st->best[k] = st->x[k]; }
    st->x_norm = sqrt(xn);
    st->gmax = lm_gmax(st->x, st->g);
    st->radius = 1e4; st->decrease_factor = 2.0; st->reuse_diagonal = 0; st->last_successful = 0;
  } else {
    st->cand_cost = cost;
    double sn = 0; for (int k = 0; k < 7; ++k) sn += (st->x[k] - st->xc[k]) * (st->x[k] - st->xc[k]);
    sn = sqrt(sn);
    if (sn <= ptol * (st->x_norm + ptol)) { st->done = 1; return; }     // ParameterToleranceReached
    const double cost_change = st->cost - cost;
    if (fabs(cost_change) <= ftol * st->cost) { st->done = 1; return; }  // FunctionToleranceReached
    const double rho = cost_change / st->model_cost_change;
    if (rho > min_rel) {  // HandleSuccessfulStep
      double xn = 0;
      for (int k = 0; k < 7; ++k) { st->x[k] = st->xc[k]; xn += st->x[k] * st->x[k]; }
      st->x_norm = sqrt(xn);
      for (int k = 0; k < 21; ++k) st->H[k] = e[k];
      for (int k = 0; k < 6; ++k) st->g[k] = e[21 + k];
      st->cost = cost;
      st->gmax = lm_gmax(st->x, st->g);
      const double tr = 2.0 * rho - 1.0;
      st->radius = fmin(max_radius, st->radius / fmax(1.0 / 3.0, 1.0 - tr * tr * tr));
      st->decrease_factor = 2.0; st->reuse_diagonal = 0; st->last_successful = 1;
      if (st->cost < st->min_cost) { st->min_cost = st->cost; for (int k = 0; k < 7; ++k) st->best[k] = st->x[k]; }
    } else {  // HandleUnsuccessfulStep
      st->radius /= st->decrease_factor; st->decrease_factor *= 2.0; st->last_successful = 0;
    }
  }
  // next trust-region step(s); an invalid step consumes an iteration without an evaluation
  while (true) {
    if (st->last_successful && st->gmax <= gtol) { st->done = 1; return; }
    if (st->radius < min_radius) { st->done = 1; return; }
    if (st->iter >= kMaxIter) { st->done = 1; return; }
    st->iter++;
    st->last_successful = 0;
    double Hs[6][6], gs[6], sc[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) sc[i] = st->scale[i];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      gs[i] = st->g[i] * sc[i];
#pragma unroll
      for (int j = 0; j < 6; ++j) Hs[i][j] = st->H[i <= j ? i * 6 - i * (i - 1) / 2 + (j - i) : j * 6 - j * (j - 1) / 2 + (i - j)] * sc[i] * sc[j];
    }
    if (!st->reuse_diagonal) {
#pragma unroll
      for (int k = 0; k < 6; ++k) st->diag[k] = fmin(fmax(Hs[k][k], min_diag), max_diag);
    }
    double A[6][6];
    const double invr = 1.0 / st->radius;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
#pragma unroll
      for (int j = 0; j < 6; ++j) A[i][j] = Hs[i][j];
      A[i][i] += st->diag[i] * invr;  // D^2 = diag / radius
    }
    double y[6];
    const bool ok = lm_chol6(A, gs, y);
    st->reuse_diagonal = 1;
    double mcc = 0;
    if (ok) {  // model_cost_change = -s'gs - s'Hs s / 2 with s = -y
      double sHs = 0, sg = 0;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        sg += -y[i] * gs[i];
        double row = 0;
#pragma unroll
        for (int j = 0; j < 6; ++j) row += Hs[i][j] * y[j];
        sHs += y[i] * row;
      }
      mcc = -sg - 0.5 * sHs;
    }
    if (!ok || !(mcc > 0.0)) { st->radius /= st->decrease_factor; st->decrease_factor *= 2.0; continue; }
    st->model_cost_change = mcc;
    double delta[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) delta[k] = -y[k] * sc[k];
    lm_plus(st->x, delta, st->xc);
    return;
  }
}

// mode 0: solver step on st; mode 1: evaluate at xEval only, sums to evalOut.
__global__ void __launch_bounds__(LM_BLOCK) lm_eval(const double* __restrict__ factors, const int* __restrict__ valid, int nslots,
                                                    LmSolveState* st, const double* __restrict__ xEval, EvalOut* __restrict__ partials,
                                                    EvalOut* __restrict__ evalOut, unsigned int* __restrict__ counter, int mode) {
  VL_PDL_WAIT();

  if (mode == 0 && st->done) return;
  double x[7];
  const double* xs = mode == 1 ? xEval : (st->iter == 0 ? st->x : st->xc);
#pragma unroll
  for (int k = 0; k < 7; ++k) x[k] = xs[k];
  double acc[28];
#pragma unroll
  for (int k = 0; k < 28; ++k) acc[k] = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nslots; i += gridDim.x * blockDim.x) {
    if (!valid[i]) continue;
    FactorRow fr;
    lm_factor(factors + (size_t)i * 10, x, fr);
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
        for (int b = a; b < 6; ++b) acc[t++] += wa * fr.J[k][b];
        acc[21 + a] += wa * fr.r[k];
      }
    }
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
  __shared__ double total[28];
  if (threadIdx.x < 28) {
    double v = 0;
    for (unsigned b = 0; b < gridDim.x; ++b) v += ((volatile EvalOut*)partials)[b].v[threadIdx.x];
    total[threadIdx.x] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    *counter = 0;
    if (mode == 1) { for (int k = 0; k < 28; ++k) evalOut->v[k] = total[k]; }
    else lm_logic(st, total);
  }
}


// ---- the whole ceres::Solve in ONE launch: a thread-block cluster keeps the iterate on chip ----
// 8 CTAs (one cluster) evaluate the factors; per-CTA partial sums stay in shared memory and CTA 0
// adds them through distributed shared memory in rank order (deterministic), runs the trust-region
// bookkeeping and publishes the next evaluation point in its own shared memory, which the other
// CTAs read back through DSMEM.  Two cluster barriers per evaluation replace the kernel boundary
// (and the global-memory round trip) of the lm_eval chain: 1 launch instead of 7 per solve.
#define LMC_CTAS 8
#define LMC_THREADS 256

#define LMC_MAX_CTAS 16
__global__ void __launch_bounds__(LMC_THREADS)
lm_solve_cluster(const double* __restrict__ factors, const int* __restrict__ valid, int nslotsBound, const int* __restrict__ d_nslots,
                 double* __restrict__ x_inout, LmSolveState* __restrict__ st_out) {
  VL_PDL_WAIT();

  const int nslots = d_nslots ? min(nslotsBound, *d_nslots) : nslotsBound;
  if (nslots <= 0) return;  // no residual blocks: Ceres leaves the parameters untouched (uniform over the cluster)
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned rank = cluster.block_rank();
  const int nct = (int)cluster.num_blocks();
  __shared__ double part[28];
  __shared__ double xs[8];
  __shared__ int sdone;
  __shared__ double total[28];
  __shared__ LmSolveState st;
  __shared__ double red[LMC_THREADS / 32][28];
  __shared__ double gath[LMC_MAX_CTAS][28];
  __shared__ double xloc[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (rank == 0 && threadIdx.x == 0) {
    for (int k = 0; k < 7; ++k) { st.x[k] = x_inout[k]; st.xc[k] = x_inout[k]; st.best[k] = x_inout[k]; xs[k] = x_inout[k]; }
    st.iter = 0; st.done = 0; st.nfactors = nslots; st.initial_cost = 0; st.final_cost = 0; st.cost = 0; st.min_cost = 0;
    sdone = 0;
  }
  cluster.sync();
  const double* xsrc = cluster.map_shared_rank(xs, 0);
  const int* dsrc = cluster.map_shared_rank(&sdone, 0);
  for (int it = 0; it < 5; ++it) {
    // one round of remote reads: 7 coordinates + the done flag, then a local broadcast
    if (threadIdx.x < 7) xloc[threadIdx.x] = xsrc[threadIdx.x];
    else if (threadIdx.x == 7) xloc[7] = (double)*dsrc;
    __syncthreads();
    if (xloc[7] != 0.0) break;  // uniform over the cluster: written before the last barrier
    double x[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) x[k] = xloc[k];
    double acc[28];
#pragma unroll
    for (int k = 0; k < 28; ++k) acc[k] = 0.0;
    for (int i = rank * LMC_THREADS + threadIdx.x; i < nslots; i += nct * LMC_THREADS) {
      if (!valid[i]) continue;
      FactorRow fr;
      lm_factor(factors + (size_t)i * 10, x, fr);
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
          for (int b = a; b < 6; ++b) acc[t++] += wa * fr.J[k][b];
          acc[21 + a] += wa * fr.r[k];
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 28; ++k) {
      double v = acc[k];
      for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
      if (lane == 0) red[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < 28) {
      double v = 0;
      for (int w = 0; w < LMC_THREADS / 32; ++w) v += red[w][threadIdx.x];
      part[threadIdx.x] = v;
    }
    cluster.sync();  // every CTA's partial is visible cluster-wide
    if (rank == 0) {
      // pull all 8 x 28 partials in one round of remote reads (one DSMEM latency, not eight), then add
      // them in rank order so the sum is deterministic
      for (int t = threadIdx.x; t < 28 * nct; t += LMC_THREADS) {
        const int r = t / 28, k = t - r * 28;
        gath[r][k] = cluster.map_shared_rank(part, r)[k];
      }
      __syncthreads();
      if (threadIdx.x < 28) {
        double v = 0;
        for (int r = 0; r < nct; ++r) v += gath[r][threadIdx.x];
        total[threadIdx.x] = v;
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        lm_logic(&st, total);
        for (int k = 0; k < 7; ++k) xs[k] = st.xc[k];
        sdone = st.done;
      }
    }
    cluster.sync();  // the next evaluation point / done flag is published
  }
  cluster.sync();  // nobody may exit while another CTA can still read its shared memory (the done flag lives in CTA 0)
  if (rank == 0 && threadIdx.x == 0) {
    for (int k = 0; k < 7; ++k) x_inout[k] = st.best[k];
    st.final_cost = st.min_cost;
    if (st_out) *st_out = st;
  }
}

__global__ void lm_begin(LmSolveState* st, const double* __restrict__ x, int nfactorsHint) {
  VL_PDL_WAIT();

  if (threadIdx.x != 0) return;
  for (int k = 0; k < 7; ++k) { st->x[k] = x[k]; st->xc[k] = x[k]; st->best[k] = x[k]; }
  st->iter = 0; st->done = 0; st->nfactors = nfactorsHint;
  st->initial_cost = 0; st->final_cost = 0; st->cost = 0; st->min_cost = 0;
}

// No residual blocks: Ceres removes the parameter blocks and returns without touching x.
__global__ void lm_end(LmSolveState* st, double* __restrict__ x, const int* __restrict__ valid, int nslots) {
  VL_PDL_WAIT();

  if (threadIdx.x != 0) return;
  for (int k = 0; k < 7; ++k) x[k] = st->best[k];
  st->final_cost = st->min_cost;
}

int vl_solve(vloam_b200_ctx* c, int nslots, const int* d_nslots, double* d_x_inout, double* costs2) {
  if (nslots > 0) {
    // one cluster: 8 CTAs for the small odometry problems, 16 (non-portable size) for the mapping ones
    const int nct = nslots > 4096 ? LMC_MAX_CTAS : LMC_CTAS;
    static bool attr = false;
    if (!attr) { VL_CUDA(cudaFuncSetAttribute(lm_solve_cluster, cudaFuncAttributeNonPortableClusterSizeAllowed, 1)); attr = true; }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(nct); cfg.blockDim = dim3(LMC_THREADS); cfg.dynamicSmemBytes = 0; cfg.stream = c->stream;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = nct; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 2;
    const bool prof = c->prof_name[0] && vl_prof_match(c, "lm_solve_cluster") && c->prof_n < VL_PROF_MAX;
    if (prof) cudaEventRecord(c->prof_ev[c->prof_n][0], c->stream);
    const double* cf = c->factors.p; const int* cv = c->factorValid.p; LmSolveState* so = costs2 ? c->lms : nullptr;
    VL_CUDA(cudaLaunchKernelEx(&cfg, lm_solve_cluster, cf, cv, nslots, d_nslots, d_x_inout, so));
    if (prof) { cudaEventRecord(c->prof_ev[c->prof_n][1], c->stream); c->prof_kname[c->prof_n] = "lm_solve_cluster"; c->prof_kbytes[c->prof_n] = 84.0 * nslots * 5;
                c->prof_n++; c->prof_bytes += 84.0 * nslots * 5; }
    c->launches++;
  }
  if (costs2) {
    if (nslots > 0) {
      VL_CUDA(cudaMemcpyAsync(c->h_lms, c->lms, sizeof(LmSolveState), cudaMemcpyDeviceToHost, c->stream));
      VL_CUDA(cudaStreamSynchronize(c->stream));
      costs2[0] = c->h_lms->initial_cost; costs2[1] = c->h_lms->final_cost;
    } else { costs2[0] = costs2[1] = 0; }
  }
  return VLOAM_OK;
}

int vl_evaluate_once(vloam_b200_ctx* c, int nslots, const double* d_x, EvalOut* d_out) {
  const int nb = max(1, min(vl_div_up(nslots, LM_BLOCK), LM_MAX_BLOCKS));
  VL_TRY(vl_reserve(c, c->evalPartials, LM_MAX_BLOCKS));
  unsigned int* counter = reinterpret_cast<unsigned int*>(c->vScalars + 60);
  VL_LAUNCH(lm_eval, nb, LM_BLOCK, 0, c->factors.p, c->factorValid.p, nslots, c->lms, d_x, c->evalPartials.p, d_out, counter, 1);
  VL_CUDA(cudaGetLastError());
  return VLOAM_OK;
}
