// laser_odometry.cu -- LaserOdometry::solveLO (laser_odometry.cpp:199-584) as CUDA.
//
// Per outer pass (2 per frame, LO.cpp:224):
//   lo_assoc<false>  sharp points  -> 1-NN in cornerLast + ring-window scan -> LidarEdgeFactor slots
//   lo_assoc<true>   flat points   -> 1-NN in surfLast  + ring-window scan -> LidarPlaneFactor slots
//   vl_solve         5 evaluation kernels (lm_solver.cu) on para_q / para_t
// then lo_accumulate (LO.cpp:524-525) and the cloud swap (LO.cpp:558-564, a pointer flip).
//
// The 1-NN is exact brute force: a CTA takes 8 queries and streams the target cloud
// once through registers (8 distances per loaded point), so L2 traffic is 1/8 of a
// query-per-CTA scan; ties resolve by (d2, index) like the oracle.  The ring-window
// scans follow the reference's visit order literally (forward ascending, then backward
// descending, strict <), one warp per query, 128 points per step, with the `break`
// position found by ballot so nothing is assumed about the monotonicity of
// int(intensity).
#include "common.cuh"

#define LO_QPB 8  // queries per block == warps per block

struct Best { float d; int j; };

__device__ __forceinline__ Best best_min_lo(Best a, Best b) {  // smaller d, then smaller index
  return (b.d < a.d || (b.d == a.d && b.j < a.j)) ? b : a;
}
__device__ __forceinline__ Best best_min_hi(Best a, Best b) {  // smaller d, then larger index (descending visit order)
  return (b.d < a.d || (b.d == a.d && b.j > a.j)) ? b : a;
}
__device__ __forceinline__ Best warp_best(Best v, bool preferLow) {
  for (int d = 16; d > 0; d >>= 1) {
    Best o; o.d = __shfl_xor_sync(0xffffffffu, v.d, d); o.j = __shfl_xor_sync(0xffffffffu, v.j, d);
    v = preferLow ? best_min_lo(v, o) : best_min_hi(v, o);
  }
  return v;
}

// LO.cpp:319-322: ((tx-sx)^2 + (ty-sy)^2) + (tz-sz)^2 in f32
__device__ __forceinline__ float lo_sqdis(const float4 t, float sx, float sy, float sz) {
  const float dx = __fsub_rn(t.x, sx), dy = __fsub_rn(t.y, sy), dz = __fsub_rn(t.z, sz);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

#define LO_TBL 144  // ring-value table: first index whose int(intensity) >= v, v in [0, LO_TBL)

// int(intensity) is the ring id and the "last" clouds are ring-major, so the `break` positions of
// the reference's scans are table look-ups -- provided the sequence really is non-decreasing.  This
// kernel builds the table and verifies that; a cloud that fails the check (a negative relTime makes
// int(intensity) = ring - 1) takes the ballot path that assumes nothing.
// tbl layout per cloud: [0, LO_TBL) first-index table, [LO_TBL] = 1 if monotone.
__global__ void __launch_bounds__(256) lo_ring_table(const float4* __restrict__ corner, int nc, const float4* __restrict__ surf, int ns,
                                                     int* __restrict__ tbl) {
  const int which = blockIdx.y;
  const float4* cl = which ? surf : corner;
  const int n = which ? ns : nc;
  int* t = tbl + which * (LO_TBL + 1);
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (n == 0) { if (j <= LO_TBL) t[j] = j == LO_TBL ? 1 : 0; return; }
  if (j >= n) return;
  const int v = min(max((int)cl[j].w, 0), LO_TBL - 1);
  const int raw = (int)cl[j].w;
  if (j == 0) { for (int q = 0; q <= v; ++q) t[q] = 0; }
  else {
    const int rawp = (int)cl[j - 1].w;
    const int vp = min(max(rawp, 0), LO_TBL - 1);
    if (raw < rawp || raw < 0) t[LO_TBL] = 0;  // not monotone (or negative): table unusable
    for (int q = vp + 1; q <= v; ++q) t[q] = j;
  }
  if (j == n - 1) for (int q = v + 1; q < LO_TBL; ++q) t[q] = n;
}
__global__ void lo_ring_table_init(int* __restrict__ tbl) { if (threadIdx.x < 2) tbl[threadIdx.x * (LO_TBL + 1) + LO_TBL] = 1; }

template <bool SURF>
__global__ void __launch_bounds__(LO_QPB * 32) lo_assoc(const float4* __restrict__ query, int nq, const float4* __restrict__ target, int nt,
                                                        const double* __restrict__ pose, const int* __restrict__ tbl, int* __restrict__ outIdx,
                                                        double* __restrict__ factors, int* __restrict__ valid, int slotBase) {
  __shared__ float sq[LO_QPB][3];
  __shared__ Best sbest[LO_QPB][LO_QPB];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * LO_QPB;
  if (threadIdx.x < LO_QPB) {
    const int qi = q0 + threadIdx.x;
    float sx = 0.f, sy = 0.f, sz = 0.f;
    if (qi < nq) {  // TransformToStart, DISTORTION == false (LO.cpp:152-173): f64 math, f32 store
      const float4 p = query[qi];
      double r[3];
      vl_qrot(pose, (double)p.x, (double)p.y, (double)p.z, r);
      sx = (float)(r[0] + pose[4]); sy = (float)(r[1] + pose[5]); sz = (float)(r[2] + pose[6]);
    }
    sq[threadIdx.x][0] = sx; sq[threadIdx.x][1] = sy; sq[threadIdx.x][2] = sz;
  }
  __syncthreads();
  // ---- phase 1: exact 1-NN for the block's 8 queries (kdtree nearestKSearch k = 1)
  float qx[LO_QPB], qy[LO_QPB], qz[LO_QPB];
  Best b[LO_QPB];
#pragma unroll
  for (int k = 0; k < LO_QPB; ++k) { qx[k] = sq[k][0]; qy[k] = sq[k][1]; qz[k] = sq[k][2]; b[k].d = 3.0e38f; b[k].j = -1; }
#pragma unroll 4
  for (int j = threadIdx.x; j < nt; j += LO_QPB * 32) {
    const float4 t = __ldg(&target[j]);
#pragma unroll
    for (int k = 0; k < LO_QPB; ++k) {
      const float d = vl_dist2(qx[k], qy[k], qz[k], t.x, t.y, t.z);
      if (d < b[k].d) { b[k].d = d; b[k].j = j; }
    }
  }
#pragma unroll
  for (int k = 0; k < LO_QPB; ++k) {
    if (b[k].j < 0) b[k].j = 0x7fffffff;
    const Best w = warp_best(b[k], true);
    if (lane == 0) sbest[k][warp] = w;
  }
  __syncthreads();
  // ---- phase 2: warp `warp` owns query q0 + warp
  const int qi = q0 + warp;
  if (qi >= nq) return;
  Best nn = sbest[warp][0];
#pragma unroll
  for (int w = 1; w < LO_QPB; ++w) nn = best_min_lo(nn, sbest[warp][w]);
  const float sx = sq[warp][0], sy = sq[warp][1], sz = sq[warp][2];
  int closest = -1, ind2 = -1, ind3 = -1;
  if (nt > 0 && nn.j != 0x7fffffff && (double)nn.d < 25.0) {  // DISTANCE_SQ_THRESHOLD (LO.cpp:299, 397)
    closest = nn.j;
    const int id = (int)target[closest].w;  // closestPointScanID
    Best f2{25.0f, -1}, f3{25.0f, -1};
    Best g2{25.0f, -1}, g3{25.0f, -1};
    if (tbl[LO_TBL]) {
      // monotone ring values: the scans stop at table positions, every point in between is visited
      const int F = tbl[min(id + 3, LO_TBL - 1)];            // first j with int(intensity) >= id + 3  (> id + 2.5)
      const int Bq = id - 2 <= 0 ? 0 : tbl[min(id - 2, LO_TBL - 1)];  // points below it have int(intensity) <= id - 3
#pragma unroll 4
      for (int j = closest + 1 + lane; j < F; j += 32) {   // LO.cpp:309-331, 407-430
        const float4 t = __ldg(&target[j]);
        const int v = (int)t.w;
        const float d = lo_sqdis(t, sx, sy, sz);
        if (SURF) {
          if (v <= id) { if (d < f2.d) { f2.d = d; f2.j = j; } }
          else if (d < f3.d) { f3.d = d; f3.j = j; }
        } else if (v > id) { if (d < f2.d) { f2.d = d; f2.j = j; } }
      }
#pragma unroll 4
      for (int j = closest - 1 - lane; j >= Bq; j -= 32) {  // LO.cpp:334-355, 433-456
        const float4 t = __ldg(&target[j]);
        const int v = (int)t.w;
        const float d = lo_sqdis(t, sx, sy, sz);
        if (SURF) {
          if (v >= id) { if (d < g2.d) { g2.d = d; g2.j = j; } }
          else if (d < g3.d) { g3.d = d; g3.j = j; }
        } else if (v < id) { if (d < g2.d) { g2.d = d; g2.j = j; } }
      }
    } else {
    // forward: j = closest+1 .. ; break at the first int(intensity) > id + 2.5 (LO.cpp:309-331, 407-430)
    for (int base = closest + 1; base < nt; base += 128) {
      bool stop = false;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = base + u * 32 + lane;
        const bool in = j < nt;
        const float4 t = in ? __ldg(&target[j]) : make_float4(0.f, 0.f, 0.f, 0.f);
        const int v = (int)t.w;
        const bool brk = in && (double)v > (double)id + 2.5;
        const unsigned bb = __ballot_sync(0xffffffffu, brk);
        const bool live = in && (bb == 0 || lane < __ffs(bb) - 1);
        if (live) {
          const float d = lo_sqdis(t, sx, sy, sz);
          if (SURF) {
            if (v <= id) { if (d < f2.d) { f2.d = d; f2.j = j; } }
            else if (d < f3.d) { f3.d = d; f3.j = j; }
          } else if (v > id) { if (d < f2.d) { f2.d = d; f2.j = j; } }
        }
        if (bb != 0 || base + (u + 1) * 32 >= nt) { stop = true; break; }
      }
      if (stop) break;
    }
    // backward: j = closest-1 .. 0 ; break at the first int(intensity) < id - 2.5 (LO.cpp:334-355, 433-456)
    for (int base = closest - 1; base >= 0; base -= 128) {
      bool stop = false;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = base - u * 32 - lane;
        const bool in = j >= 0;
        const float4 t = in ? __ldg(&target[j]) : make_float4(0.f, 0.f, 0.f, 0.f);
        const int v = (int)t.w;
        const bool brk = in && (double)v < (double)id - 2.5;
        const unsigned bb = __ballot_sync(0xffffffffu, brk);
        const bool live = in && (bb == 0 || lane < __ffs(bb) - 1);
        if (live) {
          const float d = lo_sqdis(t, sx, sy, sz);
          if (SURF) {
            if (v >= id) { if (d < g2.d) { g2.d = d; g2.j = j; } }
            else if (d < g3.d) { g3.d = d; g3.j = j; }
          } else if (v < id) { if (d < g2.d) { g2.d = d; g2.j = j; } }
        }
        if (bb != 0 || base - (u + 1) * 32 < 0) { stop = true; break; }
      }
      if (stop) break;
    }
    }
    f2 = warp_best(f2, true);
    g2 = warp_best(g2, false);
    if (SURF) { f3 = warp_best(f3, true); g3 = warp_best(g3, false); }
    // forward candidates were visited first: backward wins only when strictly nearer
    ind2 = (g2.j >= 0 && g2.d < f2.d) ? g2.j : f2.j;
    if (ind2 < 0 && g2.j >= 0) ind2 = g2.j;
    if (SURF) { ind3 = (g3.j >= 0 && g3.d < f3.d) ? g3.j : f3.j; if (ind3 < 0 && g3.j >= 0) ind3 = g3.j; }
  }
  if (lane != 0) return;
  const int slot = slotBase + qi;
  double* f = factors + (size_t)slot * 10;
  const float4 cp = query[qi];
  if (SURF) {
    outIdx[qi * 3] = closest; outIdx[qi * 3 + 1] = ind2; outIdx[qi * 3 + 2] = ind3;
    const bool ok = closest >= 0 && ind2 >= 0 && ind3 >= 0;
    valid[slot] = ok ? 1 : 0;
    if (ok) {  // LidarPlaneFactor ctor (LF.hpp:73-74): ljm_norm = normalize((j-l) x (j-m))
      const float4 J = target[closest], L = target[ind2], M = target[ind3];
      const double u[3] = {(double)J.x - (double)L.x, (double)J.y - (double)L.y, (double)J.z - (double)L.z};
      const double w[3] = {(double)J.x - (double)M.x, (double)J.y - (double)M.y, (double)J.z - (double)M.z};
      double n[3] = {u[1] * w[2] - u[2] * w[1], u[2] * w[0] - u[0] * w[2], u[0] * w[1] - u[1] * w[0]};
      const double n2 = n[0] * n[0] + n[1] * n[1] + n[2] * n[2];
      if (n2 > 0.0) { const double nn2 = sqrt(n2); n[0] /= nn2; n[1] /= nn2; n[2] /= nn2; }
      f[0] = 1.0; f[1] = cp.x; f[2] = cp.y; f[3] = cp.z;
      f[4] = J.x; f[5] = J.y; f[6] = J.z; f[7] = n[0]; f[8] = n[1]; f[9] = n[2];
    }
  } else {
    outIdx[qi * 2] = closest; outIdx[qi * 2 + 1] = ind2;
    const bool ok = ind2 >= 0;
    valid[slot] = ok ? 1 : 0;
    if (ok) {  // LidarEdgeFactor(curr, a = closest, b = second, s = 1) (LO.cpp:360-381)
      const float4 A = target[closest], B = target[ind2];
      f[0] = 0.0; f[1] = cp.x; f[2] = cp.y; f[3] = cp.z;
      f[4] = A.x; f[5] = A.y; f[6] = A.z; f[7] = B.x; f[8] = B.y; f[9] = B.z;
    }
  }
}

__global__ void lo_accumulate(LoScalars* s) {  // LO.cpp:524-525
  if (threadIdx.x != 0) return;
  double r[3];
  vl_qrot(s->q_w, s->para_t[0], s->para_t[1], s->para_t[2], r);
  s->t_w[0] = s->t_w[0] + r[0]; s->t_w[1] = s->t_w[1] + r[1]; s->t_w[2] = s->t_w[2] + r[2];
  double qn[4];
  vl_qmul(s->q_w, s->para_q, qn);
  for (int k = 0; k < 4; ++k) s->q_w[k] = qn[k];
}

__global__ void lo_set_prior(LoScalars* s, const double* __restrict__ prior) {  // LO.cpp:237-250
  if (threadIdx.x < 4) s->para_q[threadIdx.x] = prior[threadIdx.x];
  else if (threadIdx.x < 7) s->para_t[threadIdx.x - 4] = prior[threadIdx.x];
}

static int lo_ring_tables(vloam_b200_ctx* c, const float4* cornerLast, int nCL, const float4* surfLast, int nSL) {
  VL_LAUNCH(lo_ring_table_init, 1, 32, 0, c->loRingTbl);
  VL_LAUNCH(lo_ring_table, dim3(vl_div_up(max(max(nCL, nSL), LO_TBL + 1), 256), 2), 256, 0, cornerLast, nCL, surfLast, nSL, c->loRingTbl);
  return VLOAM_OK;
}

static int lo_associate(vloam_b200_ctx* c, const double* d_pose, const float4* cornerLast, int nCL, const float4* surfLast, int nSL) {
  const int nS = c->nSharp, nF = c->nFlat;
  VL_TRY(vl_reserve(c, c->loCornerIdx, (size_t)max(nS, 1) * 2));
  VL_TRY(vl_reserve(c, c->loSurfIdx, (size_t)max(nF, 1) * 3));
  VL_TRY(vl_reserve(c, c->factors, (size_t)max(nS + nF, 1) * 10));
  VL_TRY(vl_reserve(c, c->factorValid, (size_t)max(nS + nF, 1)));
  if (nS > 0)
    VL_LAUNCH(lo_assoc<false>, vl_div_up(nS, LO_QPB), LO_QPB * 32, 0, c->sharp.p, nS, cornerLast, nCL, d_pose, c->loRingTbl, c->loCornerIdx.p,
              c->factors.p, c->factorValid.p, 0);
  if (nF > 0)
    VL_LAUNCH(lo_assoc<true>, vl_div_up(nF, LO_QPB), LO_QPB * 32, 0, c->flat.p, nF, surfLast, nSL, d_pose, c->loRingTbl + (LO_TBL + 1), c->loSurfIdx.p,
              c->factors.p, c->factorValid.p, nS);
  VL_CUDA(cudaGetLastError());
  return VLOAM_OK;
}

extern bool vl_debug_capture(const vloam_b200_ctx* c);

int vl_lo_run(vloam_b200_ctx* c, const double* prior_q, const double* prior_t, int use_prior) {
  VL_TRY(vl_sr_sync_counts(c));  // sync point S1
  double* d_pose = c->los->para_q;  // para_q[4] + para_t[3] are contiguous
  if (c->lo_inited) {  // LO.cpp:209-217: the first frame only initialises
    const float4* cornerLast = c->cornerLastPtr; const float4* surfLast = c->surfLastPtr;
    double* d_prior = reinterpret_cast<double*>(c->vScalars + 32);  // 8-byte aligned scratch (7 doubles)
    if (use_prior) {
      double h[7] = {prior_q[0], prior_q[1], prior_q[2], prior_q[3], prior_t[0], prior_t[1], prior_t[2]};
      VL_CUDA(cudaMemcpyAsync(d_prior, h, sizeof h, cudaMemcpyHostToDevice, c->stream));
      VL_CUDA(cudaStreamSynchronize(c->stream));  // h is a stack buffer
    }
    VL_TRY(lo_ring_tables(c, cornerLast, c->nCornerLast, surfLast, c->nSurfLast));
    for (int pass = 0; pass < 2; ++pass) {  // LO.cpp:224
      if (use_prior) VL_LAUNCH(lo_set_prior, 1, 32, 0, c->los, d_prior);
      VL_TRY(lo_associate(c, d_pose, cornerLast, c->nCornerLast, surfLast, c->nSurfLast));
      if (vl_debug_capture(c)) {
        VL_TRY(vl_reserve(c, c->dbgLoCorner[pass], (size_t)max(c->nSharp, 1) * 2));
        VL_TRY(vl_reserve(c, c->dbgLoSurf[pass], (size_t)max(c->nFlat, 1) * 3));
        VL_CUDA(cudaMemcpyAsync(c->dbgLoCorner[pass].p, c->loCornerIdx.p, sizeof(int) * 2 * c->nSharp, cudaMemcpyDeviceToDevice, c->stream));
        VL_CUDA(cudaMemcpyAsync(c->dbgLoSurf[pass].p, c->loSurfIdx.p, sizeof(int) * 3 * c->nFlat, cudaMemcpyDeviceToDevice, c->stream));
      }
      VL_TRY(vl_solve(c, c->nSharp + c->nFlat, d_pose, vl_debug_capture(c) ? &c->dbgLoCost[pass * 2] : nullptr));
    }
    VL_LAUNCH(lo_accumulate, 1, 32, 0, c->los);
  }
  c->lo_inited = true;
  // LO.cpp:558-574: this frame's less-sharp / less-flat clouds become the "last" clouds
  c->cornerLastPtr = c->lessSharp[c->cur].p; c->surfLastPtr = c->lessFlat[c->cur].p;
  c->nCornerLast = c->nLessSharp; c->nSurfLast = c->nLessFlat;
  c->lo_frameCount++;
  c->skip_frame = (c->lo_frameCount % c->prm.mapping_skip_frame) != 0;  // LO.cpp:668-678
  VL_CUDA(cudaGetLastError());
  return VLOAM_OK;
}

// Association only, at an explicit pose; nothing in the context's odometry state changes.
int vl_lo_associate_only(vloam_b200_ctx* c, const double* x, int* corner_idx, int* surf_idx) {
  VL_TRY(vl_sr_sync_counts(c));
  double* d_x = reinterpret_cast<double*>(c->vScalars + 32);
  VL_CUDA(cudaMemcpyAsync(d_x, x, 7 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  VL_CUDA(cudaStreamSynchronize(c->stream));
  VL_TRY(lo_ring_tables(c, c->cornerLastPtr, c->nCornerLast, c->surfLastPtr, c->nSurfLast));
  VL_TRY(lo_associate(c, d_x, c->cornerLastPtr, c->nCornerLast, c->surfLastPtr, c->nSurfLast));
  if (corner_idx && c->nSharp) VL_CUDA(cudaMemcpyAsync(corner_idx, c->loCornerIdx.p, sizeof(int) * 2 * c->nSharp, cudaMemcpyDeviceToHost, c->stream));
  if (surf_idx && c->nFlat) VL_CUDA(cudaMemcpyAsync(surf_idx, c->loSurfIdx.p, sizeof(int) * 3 * c->nFlat, cudaMemcpyDeviceToHost, c->stream));
  VL_CUDA(cudaStreamSynchronize(c->stream));
  return VLOAM_OK;
}
