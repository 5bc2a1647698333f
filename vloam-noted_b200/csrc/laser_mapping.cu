// laser_mapping.cu -- LaserMapping::input / solveMapping (laser_mapping.cpp:178-814) as CUDA.
//
// Map layout in HBM.  The reference keeps 2 x 4851 per-cube PCL clouds (LM.h:117-150).
// Here each cube of each kind (corner / surf) is a segment [start, start+count) of one
// big float4 pool plus a `sorted` length: the prefix whose voxel keys are strictly
// increasing (what a previous VoxelGrid pass left behind); anything after it is an
// unsorted tail (raw appends to cubes outside the 5x5x3 window).  Rolling the 21x21x11
// window (LM.cpp:252-444) permutes the 4851-entry table, it never moves points.
//
// Per frame:
//   lm_prepare        input() + centre cube + roll + valid-cube list + gather offsets
//   lm_gather         the 75 valid cubes -> contiguous cornerFromMap / surfFromMap (canonical kNN ids)
//   vl_voxel_grid     current less-sharp / less-flat -> cornerStack / surfStack (LM.cpp:492-500)
//   [sync S2: Mc, Ms, Qc, Qs, tail sizes]
//   grid build        counting sort of both sub-maps into 2 m cells over the 250x250x150 m window
//   2 x { lm_knn (5-NN), lm_fit (PCA line / QR plane -> factor slots), vl_solve }
//   lm_transform_update, then the map update:
//   rf_*              VoxelGrid re-filter of the valid cubes (LM.cpp:795-808) WITHOUT re-sorting the
//                     map: only tails + this frame's points are sorted, then merged into the
//                     already-ordered prefixes (exact: a prefix has one point per voxel and
//                     the lowest point indices, so every run is [prefix point?] + tail points)
#include <limits.h>
#include <float.h>
#include <math_constants.h>
#include <cooperative_groups.h>
#include "common.cuh"
#include "bitonic.cuh"
namespace cg = cooperative_groups;
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

// ---- helper thread ------------------------------------------------------------------------------------
// After sync point S2 the caller has its pose; what remains of the frame is issuing ~25 launches of the map
// update and of the next frame's speculative sub-map on stream3 -- ~100 us of host time during which neither
// the caller nor the next sweep's scan registration made progress.  One helper thread per context issues
// them instead.  It spins for a millisecond after each task (a replay hands it work every ~0.4 ms) and
// sleeps on a condition variable otherwise (a live 10 Hz feed must not burn a core).
thread_local cudaStream_t vl_tls_stream = nullptr;

struct VlWorker {
  std::thread th;
  std::mutex m;
  std::condition_variable cv;
  std::function<int()> task;
  std::atomic<int> state{0};   // 0 idle, 1 task posted, 2 running
  std::atomic<bool> sleeping{false};
  bool quit = false;
  int rc = VLOAM_OK;
  int device = 0;
};

static void lm_worker_main(VlWorker* w) {
  cudaSetDevice(w->device);
  for (;;) {
    const auto t0 = std::chrono::steady_clock::now();
    while (w->state.load(std::memory_order_acquire) != 1) {
      if (std::chrono::steady_clock::now() - t0 > std::chrono::milliseconds(1)) {
        std::unique_lock<std::mutex> lk(w->m);
        w->sleeping.store(true);
        w->cv.wait(lk, [&] { return w->state.load(std::memory_order_acquire) == 1 || w->quit; });
        w->sleeping.store(false);
        if (w->quit) return;
        break;
      }
      // busy-wait briefly, then give the core away between polls: with one process per GPU on a 16-core box the
      // helper threads of eight ranks must not starve the threads that feed them
      if (std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(50)) std::this_thread::yield();
#if defined(__x86_64__)
      else __builtin_ia32_pause();
#endif
    }
    w->state.store(2, std::memory_order_relaxed);
    w->rc = w->task();
    w->state.store(0, std::memory_order_release);
  }
}

int vl_lm_join(vloam_b200_ctx* c) {
  VlWorker* w = c->worker;
  if (!w) return VLOAM_OK;
  while (w->state.load(std::memory_order_acquire) != 0) {
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
  }
  const int rc = w->rc;
  w->rc = VLOAM_OK;
  return rc;
}

void vl_lm_shutdown(vloam_b200_ctx* c) {
  VlWorker* w = c->worker;
  if (!w) return;
  vl_lm_join(c);
  { std::lock_guard<std::mutex> lk(w->m); w->quit = true; }
  w->cv.notify_all();
  // a spinning worker notices quit only through the condition variable path: post nothing, it falls asleep within 1 ms
  w->th.join();
  delete w;
  c->worker = nullptr;
}

static int lm_submit(vloam_b200_ctx* c, std::function<int()> f) {
  if (!c->worker) {
    VlWorker* w = new VlWorker();
    w->device = c->device;
    w->th = std::thread(lm_worker_main, w);
    c->worker = w;
  }
  VlWorker* w = c->worker;
  VL_TRY(vl_lm_join(c));
  w->task = std::move(f);
  // Publish under the worker's mutex and always notify: a worker that is between its last poll of `state` and
  // cv.wait() holds the mutex, so it either sees state == 1 in the wait predicate or receives this notification
  // (checking `sleeping` without the lock could miss both: the store and the load may be reordered on x86).
  { std::lock_guard<std::mutex> lk(w->m); w->state.store(1, std::memory_order_seq_cst); }
  w->cv.notify_one();
  return VLOAM_OK;
}


#define LM_CELL 2.0f
#define LM_GX 125
#define LM_GY 125
#define LM_GZ 75
#define LM_NCELL (LM_GX * LM_GY * LM_GZ)
#define LM_NSEG 250  // (kind, valid slot)

struct RfWork {
  int slotOfCube[VL_CUBE_NUM];          // valid slot of a cube index or -1
  int gatherOff[2][VL_MAX_VALID + 1];   // exclusive offsets of the valid cubes inside fromMap
  int tailOff[LM_NSEG + 1];             // exclusive offsets of existing-tail keys
  int prefOff[LM_NSEG + 1];             // exclusive offsets of prefix points
  int tailBegin[LM_NSEG + 1];           // per segment range in the sorted key array
  int segCount[LM_NSEG + 1];            // keys per segment (histogram of rf_keys; zeroed by lm_prepare)
  int segFill[LM_NSEG + 1];             // scatter cursors
  int outOff[LM_NSEG + 1];              // staging offsets
  int outCount[LM_NSEG];
  int firstViolation[LM_NSEG];
  int nKeysValid;
  int nq;                               // Qc + Qs when the optimisation runs, else 0
  float gridOrigin[3];
  int Qc, Qs;
};

// Everything the sub-map gather and the search grid need to know about one valid-cube window.  Two
// copies live on the device: `real`, written by lm_prepare from this frame's pose, and `spec`, written
// right after the previous frame's map update for the window that frame used (lm_spec_prepare).
struct LmSub {
  int validNum, Mc, Ms;
  int cI, cJ, cK, cenW, cenH, cenD;     // window centre and laserCloudCen* it was built for
  int ok;                               // spec: the descriptor (and the structures built from it) is complete
  float gridOrigin[3];
  int validInd[VL_MAX_VALID];
  int gatherOff[2][VL_MAX_VALID + 1];
};

__device__ __forceinline__ int lm_cube_of(float v, int cen) {  // LM.cpp:747-756
  const double t = (double)v + 25.0;
  int c = (int)(t / 50.0) + cen;
  if (t < 0) c--;
  return c;
}

// voxel key of a map point inside cube (ci,cj,ck): 10 bits per axis, (iz, iy, ix) order ==
// ascending pcl::VoxelGrid linear index inside any bounding box of that cube's points.
__device__ __forceinline__ unsigned lm_vox_key(const float4 p, float inv, int ci, int cj, int ck, int cenW, int cenH, int cenD) {
  const int bx = (int)floorf(__fmul_rn((float)(50 * (ci - cenW) - 25), inv)) - 2;
  const int by = (int)floorf(__fmul_rn((float)(50 * (cj - cenH) - 25), inv)) - 2;
  const int bz = (int)floorf(__fmul_rn((float)(50 * (ck - cenD) - 25), inv)) - 2;
  const int ix = min(max((int)floorf(__fmul_rn(p.x, inv)) - bx, 0), 1023);
  const int iy = min(max((int)floorf(__fmul_rn(p.y, inv)) - by, 0), 1023);
  const int iz = min(max((int)floorf(__fmul_rn(p.z, inv)) - bz, 0), 1023);
  return ((unsigned)iz << 20) | ((unsigned)iy << 10) | (unsigned)ix;
}
__device__ __forceinline__ unsigned lm_vox_key_cube(const float4 p, float inv, int cube, const LmScalars* s) {
  const int ci = cube % VL_CUBE_W, cj = (cube / VL_CUBE_W) % VL_CUBE_H, ck = cube / (VL_CUBE_W * VL_CUBE_H);
  return lm_vox_key(p, inv, ci, cj, ck, s->cenW, s->cenH, s->cenD);
}

// exclusive scan of one int per thread over the first 256 threads of the block (all threads must call)
__device__ __forceinline__ int lm_scan256(int v, int* buf, int* total) {
  // exclusive scan over threads 0..255 (block of 256 or 1024 threads; every thread must call): warp shuffles, then
  // the eight warp totals through shared memory -- two barriers instead of the eighteen of a Hillis-Steele ladder
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  int inc = t < 256 ? v : 0;
  for (int d = 1; d < 32; d <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += u; }
  if (t < 256 && lane == 31) buf[warp] = inc;
  __syncthreads();
  int before = 0, all = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) { const int sw = buf[w]; if (w < warp) before += sw; all += sw; }
  if (total) *total = all;
  __syncthreads();
  return t < 256 ? before + inc - v : 0;
}

// ---- LaserMapping::input (LM.cpp:178-209) + centre cube / roll / valid list (LM.cpp:228-466)
__global__ void __launch_bounds__(1024) lm_prepare(LmScalars* __restrict__ s, const LoScalars* __restrict__ lo, MapCubeTable* __restrict__ tc,
                                                   MapCubeTable* __restrict__ ts, RfWork* __restrict__ w, int skip, int resetValid,
                                                   LmSub* __restrict__ real, const LmSub* __restrict__ spec, int specQueued, int* __restrict__ specOK) {
  VL_PDL_WAIT();

  __shared__ int shift[3];
  __shared__ int center[3];
  __shared__ int sFresh;
  if (threadIdx.x == 0) {
    if (resetValid) s->validNum = 0;  // LaserMapping::reset (LM.cpp:132-136)
    sFresh = s->validNum == 0;
    for (int k = 0; k < 4; ++k) s->q_wodom[k] = lo->q_w[k];
    for (int k = 0; k < 3; ++k) s->t_wodom[k] = lo->t_w[k];
    double r[3];
    vl_qrot(s->q_wmap_wodom, s->t_wodom[0], s->t_wodom[1], s->t_wodom[2], r);
    if (skip) {
      vl_qmul(s->q_wmap_wodom, s->q_wodom, s->q_hf);
      for (int k = 0; k < 3; ++k) s->t_hf[k] = r[k] + s->t_wmap_wodom[k];
    } else {
      double qn[4];
      vl_qmul(s->q_wmap_wodom, s->q_wodom, qn);
      for (int k = 0; k < 4; ++k) s->pose[k] = qn[k];
      for (int k = 0; k < 3; ++k) s->pose[4 + k] = r[k] + s->t_wmap_wodom[k];
    }
    int cI = (int)((s->pose[4] + 25.0) / 50.0) + s->cenW;
    int cJ = (int)((s->pose[5] + 25.0) / 50.0) + s->cenH;
    int cK = (int)((s->pose[6] + 25.0) / 50.0) + s->cenD;
    if (s->pose[4] + 25.0 < 0) cI--;
    if (s->pose[5] + 25.0 < 0) cJ--;
    if (s->pose[6] + 25.0 < 0) cK--;
    int sI = 0, sJ = 0, sK = 0;
    if (!skip) {  // net effect of the six while loops: contents move by +s along an axis
      while (cI < 3) { cI++; sI++; }
      while (cI >= VL_CUBE_W - 3) { cI--; sI--; }
      while (cJ < 3) { cJ++; sJ++; }
      while (cJ >= VL_CUBE_H - 3) { cJ--; sJ--; }
      while (cK < 3) { cK++; sK++; }
      while (cK >= VL_CUBE_D - 3) { cK--; sK--; }
      s->cenW += sI; s->cenH += sJ; s->cenD += sK;
    }
    shift[0] = sI; shift[1] = sJ; shift[2] = sK;
    center[0] = cI; center[1] = cJ; center[2] = cK;
  }
  __syncthreads();
  if (skip) return;  // (a skipped frame neither consumes nor invalidates the speculative sub-map)
  const int sI = shift[0], sJ = shift[1], sK = shift[2];
  if (sI != 0 || sJ != 0 || sK != 0) {
    // new[i,j,k] = old[i-sI, j-sJ, k-sK]; entries whose source wrapped are the cleared slabs (they
    // keep the storage of the entry that fell off the other side).
    int v[5][8]; bool wrap[5];
    for (int q = 0; q < 5; ++q) {
      const int d = threadIdx.x + q * 1024;
      if (d >= VL_CUBE_NUM) break;
      const int i = d % VL_CUBE_W, j = (d / VL_CUBE_W) % VL_CUBE_H, k = d / (VL_CUBE_W * VL_CUBE_H);
      const int si = i - sI, sj = j - sJ, sk = k - sK;
      wrap[q] = si < 0 || si >= VL_CUBE_W || sj < 0 || sj >= VL_CUBE_H || sk < 0 || sk >= VL_CUBE_D;
      // saturating modulo: multi-cube jumps still give a bijection
      const int mi = ((si % VL_CUBE_W) + VL_CUBE_W) % VL_CUBE_W, mj = ((sj % VL_CUBE_H) + VL_CUBE_H) % VL_CUBE_H,
                mk = ((sk % VL_CUBE_D) + VL_CUBE_D) % VL_CUBE_D;
      const int src = mi + VL_CUBE_W * mj + VL_CUBE_W * VL_CUBE_H * mk;
      v[q][0] = tc->start[src]; v[q][1] = tc->count[src]; v[q][2] = tc->cap[src]; v[q][3] = tc->sorted[src];
      v[q][4] = ts->start[src]; v[q][5] = ts->count[src]; v[q][6] = ts->cap[src]; v[q][7] = ts->sorted[src];
    }
    __syncthreads();
    for (int q = 0; q < 5; ++q) {
      const int d = threadIdx.x + q * 1024;
      if (d >= VL_CUBE_NUM) break;
      tc->start[d] = v[q][0]; tc->count[d] = wrap[q] ? 0 : v[q][1]; tc->cap[d] = v[q][2]; tc->sorted[d] = wrap[q] ? 0 : v[q][3];
      ts->start[d] = v[q][4]; ts->count[d] = wrap[q] ? 0 : v[q][5]; ts->cap[d] = v[q][6]; ts->sorted[d] = wrap[q] ? 0 : v[q][7];
    }
  }
  __shared__ int sTot[2];
  if (threadIdx.x < 2) sTot[threadIdx.x] = 0;
  __syncthreads();
  if (threadIdx.x <= LM_NSEG) { w->segCount[threadIdx.x] = 0; w->segFill[threadIdx.x] = 0; }  // histogram / bucket cursors of this frame's update keys
  int myC = 0, myS = 0;
  for (int d = threadIdx.x; d < VL_CUBE_NUM; d += 1024) { w->slotOfCube[d] = -1; myC += tc->count[d]; myS += ts->count[d]; }
  for (int o = 16; o > 0; o >>= 1) { myC += __shfl_xor_sync(0xffffffffu, myC, o); myS += __shfl_xor_sync(0xffffffffu, myS, o); }
  if ((threadIdx.x & 31) == 0) { atomicAdd(&sTot[0], myC); atomicAdd(&sTot[1], myS); }
  __shared__ int sbuf[256];
  __shared__ int snv;
  __syncthreads();
  if (threadIdx.x == 0) { s->totalC = sTot[0]; s->totalS = sTot[1]; }
  if (threadIdx.x == 0) {
    const int cI = center[0], cJ = center[1], cK = center[2];
    int nv = s->validNum;  // LaserMapping::reset zeroes it once per frame (LM.cpp:132-136)
    for (int i = cI - 2; i <= cI + 2; i++)
      for (int j = cJ - 2; j <= cJ + 2; j++)
        for (int k = cK - 1; k <= cK + 1; k++)
          if (i >= 0 && i < VL_CUBE_W && j >= 0 && j < VL_CUBE_H && k >= 0 && k < VL_CUBE_D && nv < VL_MAX_VALID)
            s->validInd[nv++] = i + VL_CUBE_W * j + VL_CUBE_W * VL_CUBE_H * k;
    s->validNum = nv;
    snv = nv;
    w->gridOrigin[0] = (float)(50 * (cI - 2 - s->cenW) - 25);
    w->gridOrigin[1] = (float)(50 * (cJ - 2 - s->cenH) - 25);
    w->gridOrigin[2] = (float)(50 * (cK - 1 - s->cenD) - 25);
    real->validNum = nv; real->cI = cI; real->cJ = cJ; real->cK = cK; real->cenW = s->cenW; real->cenH = s->cenH; real->cenD = s->cenD;
    for (int k = 0; k < 3; ++k) real->gridOrigin[k] = w->gridOrigin[k];
  }
  __syncthreads();
  const int nv = snv;
  const int t = threadIdx.x;
  // thread t < 250 owns segment (kind, slot) = (t / 125, t % 125)
  const int kind = t / VL_MAX_VALID, slot = t % VL_MAX_VALID;
  int cnt = 0, srt = 0;
  if (t < LM_NSEG && slot < nv) {
    const int cb = s->validInd[slot];
    const MapCubeTable* tb = kind ? ts : tc;
    cnt = tb->count[cb]; srt = tb->sorted[cb];
    if (kind == 0) w->slotOfCube[cb] = slot;
  }
  int total = 0;
  const int offCnt = lm_scan256(cnt, sbuf, &total);          // corner segments first, then surf
  __shared__ int sMc;
  if (t == VL_MAX_VALID) sMc = offCnt;                        // exclusive offset of the first surf segment == Mc
  __syncthreads();
  if (t < LM_NSEG) { w->gatherOff[kind][slot] = kind ? offCnt - sMc : offCnt; real->gatherOff[kind][slot] = kind ? offCnt - sMc : offCnt; }
  if (t < VL_MAX_VALID) real->validInd[t] = t < nv ? s->validInd[t] : 0;
  if (t == 0) {
    s->Mc = sMc; s->Ms = total - sMc; w->gatherOff[0][VL_MAX_VALID] = sMc; w->gatherOff[1][VL_MAX_VALID] = total - sMc;
    real->Mc = sMc; real->Ms = total - sMc; real->gatherOff[0][VL_MAX_VALID] = sMc; real->gatherOff[1][VL_MAX_VALID] = total - sMc; real->ok = 1;
    // The sub-map gathered and cell-sorted after the previous map update is this frame's sub-map when the
    // window did not move: same centre cube, no roll, a freshly reset valid list.  The cube tables have not
    // changed since, so equal windows mean equal offsets; the sizes are compared as a last line of defence.
    *specOK = (specQueued && spec->ok && sFresh && shift[0] == 0 && shift[1] == 0 && shift[2] == 0 && spec->cI == center[0] &&
               spec->cJ == center[1] && spec->cK == center[2] && spec->cenW == s->cenW && spec->cenH == s->cenH && spec->cenD == s->cenD &&
               spec->validNum == nv && spec->Mc == sMc && spec->Ms == total - sMc) ? 1 : 0;
  }
  int tailTotal = 0, prefTotal = 0;
  const int tl = cnt - srt;
  const int offTail = lm_scan256(tl, sbuf, &tailTotal);
  const int offPref = lm_scan256(srt, sbuf, &prefTotal);
  if (t < LM_NSEG) { w->tailOff[t] = offTail; w->prefOff[t] = offPref; }
  __shared__ int sTailC;
  if (t == VL_MAX_VALID) sTailC = offTail;
  __syncthreads();
  if (t == 0) { w->tailOff[LM_NSEG] = tailTotal; w->prefOff[LM_NSEG] = prefTotal; s->tailC = sTailC; s->tailS = tailTotal - sTailC; }
}

// The window the next frame will most likely use is the one this frame used (a cube is 50 m wide): right
// after the map update, on the update's stream, its descriptor is rebuilt from the updated cube tables.
__global__ void __launch_bounds__(256) lm_spec_prepare(const LmSub* __restrict__ last, const MapCubeTable* __restrict__ tc,
                                                       const MapCubeTable* __restrict__ ts, LmSub* __restrict__ spec) {
  VL_PDL_WAIT();

  __shared__ int sbuf[256];
  __shared__ int snv;
  if (threadIdx.x == 0) {
    const int cI = last->cI, cJ = last->cJ, cK = last->cK;
    int nv = 0;
    for (int i = cI - 2; i <= cI + 2; i++)  // LM.cpp:448-466, on a freshly reset list
      for (int j = cJ - 2; j <= cJ + 2; j++)
        for (int k = cK - 1; k <= cK + 1; k++)
          if (i >= 0 && i < VL_CUBE_W && j >= 0 && j < VL_CUBE_H && k >= 0 && k < VL_CUBE_D && nv < VL_MAX_VALID)
            spec->validInd[nv++] = i + VL_CUBE_W * j + VL_CUBE_W * VL_CUBE_H * k;
    spec->validNum = nv; snv = nv;
    spec->cI = cI; spec->cJ = cJ; spec->cK = cK; spec->cenW = last->cenW; spec->cenH = last->cenH; spec->cenD = last->cenD;
    spec->gridOrigin[0] = (float)(50 * (cI - 2 - last->cenW) - 25);
    spec->gridOrigin[1] = (float)(50 * (cJ - 2 - last->cenH) - 25);
    spec->gridOrigin[2] = (float)(50 * (cK - 1 - last->cenD) - 25);
  }
  __syncthreads();
  const int nv = snv, t = threadIdx.x;
  const int kind = t / VL_MAX_VALID, slot = t % VL_MAX_VALID;
  int cnt = 0;
  if (t < LM_NSEG && slot < nv) cnt = (kind ? ts : tc)->count[spec->validInd[slot]];
  int total = 0;
  const int offCnt = lm_scan256(cnt, sbuf, &total);
  __shared__ int sMc;
  if (t == VL_MAX_VALID) sMc = offCnt;
  __syncthreads();
  if (t < LM_NSEG) spec->gatherOff[kind][slot] = kind ? offCnt - sMc : offCnt;
  if (t == 0) { spec->Mc = sMc; spec->Ms = total - sMc; spec->gatherOff[0][VL_MAX_VALID] = sMc; spec->gatherOff[1][VL_MAX_VALID] = total - sMc; spec->ok = 1; }
}

// zero the two cell-counter arrays of the search grid (a kernel, not a memset, so that it can be skipped)
__device__ __forceinline__ void lm_dev_zero(int* __restrict__ a, int* __restrict__ b, int n) {
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < n; g += gridDim.x * blockDim.x) { a[g] = 0; b[g] = 0; }
}
__global__ void __launch_bounds__(256) lm_grid_zero(int* __restrict__ a, int* __restrict__ b, int n, const int* __restrict__ skip) {
  VL_PDL_WAIT();

  if (skip && *skip) return;
  lm_dev_zero(a, b, n);
}

// LM.cpp:476-485: concatenate the valid cubes (loop order of LM.cpp:448-452) into the sub-map clouds.
// skip (may be null): device flag "the speculative build already produced exactly this" -> nothing to do.
__device__ __forceinline__ void lm_dev_gather(const LmSub* __restrict__ sub, const MapCubeTable* __restrict__ tc, const MapCubeTable* __restrict__ ts,
                                              const float4* __restrict__ poolC, const float4* __restrict__ poolS,
                                              float4* __restrict__ outC, float4* __restrict__ outS) {
  const int nv = sub->validNum, mc = sub->Mc, total = sub->Mc + sub->Ms;
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < total; g += gridDim.x * blockDim.x) {
    const int kind = g >= mc;
    const int e = kind ? g - mc : g;
    const int* off = sub->gatherOff[kind];
    int lo = 0, hi = nv;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (off[mid] <= e) lo = mid; else hi = mid; }
    const int cb = sub->validInd[lo];
    if (kind) outS[e] = poolS[ts->start[cb] + (e - off[lo])];
    else outC[e] = poolC[tc->start[cb] + (e - off[lo])];
  }
}
// ---- search grid: counting sort of both sub-maps into 2 m cells --------------------------------
__device__ __forceinline__ int lm_cell_coord(float v, float o, int n) {
  const int cidx = (int)floorf(__fmul_rn(__fsub_rn(v, o), 1.0f / LM_CELL));
  return min(max(cidx, 0), n - 1);
}
__device__ __forceinline__ void lm_dev_count(const LmSub* __restrict__ sub, const float4* __restrict__ mapC, const float4* __restrict__ mapS,
                                             int* __restrict__ cellCount, int* __restrict__ cellOfPoint) {
  const int mc = sub->Mc, total = sub->Mc + sub->Ms;
  const float ox = sub->gridOrigin[0], oy = sub->gridOrigin[1], oz = sub->gridOrigin[2];
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < total; g += gridDim.x * blockDim.x) {
    const int kind = g >= mc;
    const float4 p = kind ? mapS[g - mc] : mapC[g];
    const int cell = kind * LM_NCELL + lm_cell_coord(p.x, ox, LM_GX) + LM_GX * (lm_cell_coord(p.y, oy, LM_GY) + LM_GY * lm_cell_coord(p.z, oz, LM_GZ));
    cellOfPoint[g] = cell;
    atomicAdd(&cellCount[cell], 1);
  }
}
// Speculative path: gather and cell count in one pass -- each point is read from the pool once, written to the
// sub-map cloud and binned (one launch and one re-read of the gathered cloud less on the update -> sub-map chain).
__global__ void __launch_bounds__(256) lm_gather_count(const LmSub* __restrict__ sub, const MapCubeTable* __restrict__ tc,
                                                       const MapCubeTable* __restrict__ ts, const float4* __restrict__ poolC,
                                                       const float4* __restrict__ poolS, float4* __restrict__ outC, float4* __restrict__ outS,
                                                       int* __restrict__ cellCount, int* __restrict__ cellOfPoint) {
  VL_PDL_WAIT();

  const int nv = sub->validNum, mc = sub->Mc, total = sub->Mc + sub->Ms;
  const float ox = sub->gridOrigin[0], oy = sub->gridOrigin[1], oz = sub->gridOrigin[2];
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < total; g += gridDim.x * blockDim.x) {
    const int kind = g >= mc;
    const int e = kind ? g - mc : g;
    const int* off = sub->gatherOff[kind];
    int lo = 0, hi = nv;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (off[mid] <= e) lo = mid; else hi = mid; }
    const int cb = sub->validInd[lo];
    const float4 p = kind ? poolS[ts->start[cb] + (e - off[lo])] : poolC[tc->start[cb] + (e - off[lo])];
    if (kind) outS[e] = p; else outC[e] = p;
    const int cell = kind * LM_NCELL + lm_cell_coord(p.x, ox, LM_GX) + LM_GX * (lm_cell_coord(p.y, oy, LM_GY) + LM_GY * lm_cell_coord(p.z, oz, LM_GZ));
    cellOfPoint[g] = cell;
    atomicAdd(&cellCount[cell], 1);
  }
}
// exclusive scan over 2*LM_NCELL counts, three-phase form for the cooperative in-line build: tile sums (1024 per
// block) -> scan of tile sums -> apply
__device__ __forceinline__ void lm_dev_scan_tile(const int* __restrict__ in, int n, int* __restrict__ tileSum, int tile) {  // 256 threads
  int acc = 0;
  const int base = tile * 1024;
  for (int q = 0; q < 4; ++q) { const int t = base + q * 256 + threadIdx.x; if (t < n) acc += in[t]; }
  __shared__ int ws[8];
  for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
  __syncthreads();  // (ws may still be read by the previous tile of a looping caller)
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) { int v = 0; for (int k = 0; k < 8; ++k) v += ws[k]; tileSum[tile] = v; }
}
// exclusive scan of the tile sums in place by ONE block of T threads (T a power of two <= 1024)
template <int T>
__device__ __forceinline__ void lm_dev_scan_sums(int* __restrict__ tileSum, int nTiles) {
  __shared__ int ws[32];
  int carry = 0;  // the same value in every thread
  for (int base = 0; base < nTiles; base += T) {
    const int b = base + threadIdx.x;
    const int own = b < nTiles ? tileSum[b] : 0;
    int tot = 0;
    const int ex = vl_block_excl_scan<T>(own, ws, &tot);
    if (b < nTiles) tileSum[b] = carry + ex;
    carry += tot;
  }
}
__device__ __forceinline__ void lm_dev_scan_apply(const int* __restrict__ in, int n, const int* __restrict__ tileSum, int* __restrict__ out,
                                                  int tile, int nTiles) {  // 256 threads
  __shared__ int ws[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int running = tileSum[tile];
  for (int q = 0; q < 4; ++q) {
    const int t = tile * 1024 + q * 256 + threadIdx.x;
    const int v = t < n ? in[t] : 0;
    int inc = v;
    for (int d = 1; d < 32; d <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += u; }
    __syncthreads();  // (ws may still be read by the previous round / tile)
    if (lane == 31) ws[warp] = inc;
    __syncthreads();
    int before = 0, all = 0;
    for (int k = 0; k < 8; ++k) { if (k < warp) before += ws[k]; all += ws[k]; }
    if (t < n) out[t] = running + before + inc - v;
    running += all;
  }
  if (tile == nTiles - 1 && threadIdx.x == 0) out[n] = running;  // total
}
// ---- single-launch exclusive scan (chained tiles with look-back) --------------------------------------
// out[0..n] = exclusive scan of in[0..n) (out[n] = total).  A tile is 4096 counts (16 consecutive per thread, four
// 128-bit loads and stores each).  Tiles take their index from a ticket counter, so every predecessor of a running
// tile is itself running or done; a tile publishes its own sum at once ("aggregate"), warp 0 then walks the status
// words of the tiles before it, 64 at a time, until it meets one that already knows its inclusive prefix, and
// publishes its own.  Status word: [63:34] epoch of this launch | [33:32] 1 aggregate / 2 inclusive | [31:0] value --
// written and read as one 64-bit word, and stale words of earlier launches simply read as "not ready", so nothing
// has to be cleared between launches.  The three-kernel version (tile sums, scan of sums, apply) cost two more
// dependent launches and a second read of the counts on the update -> sub-map chain.
#define SCAN_TILE 4096  // 256 threads x 16 consecutive counts
__global__ void __launch_bounds__(256) lm_scan_chained(const int* __restrict__ in, int n, unsigned long long* __restrict__ state,
                                                       unsigned epoch, int* __restrict__ out, const int* __restrict__ skip) {
  VL_PDL_WAIT();

  if (skip && *skip) return;
  __shared__ int sTile, sPrefix;
  __shared__ int ws[32];
  int* ticket = reinterpret_cast<int*>(state + gridDim.x);
  if (threadIdx.x == 0) {
    const int t = atomicAdd(ticket, 1);
    if (t == (int)gridDim.x - 1) *ticket = 0;  // every tile has its ticket: ready for the next launch
    sTile = t;
  }
  __syncthreads();
  const int tile = sTile;
  const int base = tile * SCAN_TILE + threadIdx.x * 16;
  int v[16];
  if (base + 15 < n) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int4 x = *reinterpret_cast<const int4*>(in + base + 4 * q);
      v[4 * q] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w;
    }
  } else {
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = base + k < n ? in[base + k] : 0;
  }
  int own = 0;
#pragma unroll
  for (int k = 0; k < 16; ++k) own += v[k];
  int tot = 0;
  const int ex = vl_block_excl_scan<256>(own, ws, &tot);
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    const unsigned long long tag = (unsigned long long)epoch << 34;
    volatile unsigned long long* st = state;
    int excl = 0;
    if (tile > 0) {
      if (lane == 0) st[tile] = tag | (1ull << 32) | (unsigned)tot;
      // Tiles start together, so a far tile may walk a long way back over aggregates: 64 status words per step
      // (group 0: the 32 nearest tiles, group 1: the 32 before them).
      int look = tile - 1;
      for (;;) {
        const int i0 = look - lane, i1 = look - 32 - lane;
        const unsigned long long w0 = i0 >= 0 ? st[i0] : (tag | (2ull << 32));  // before tile 0: inclusive prefix 0
        const unsigned long long w1 = i1 >= 0 ? st[i1] : (tag | (2ull << 32));
        const int s0 = (w0 >> 34) == (unsigned long long)epoch ? (int)((w0 >> 32) & 3) : 0;
        const int s1 = (w1 >> 34) == (unsigned long long)epoch ? (int)((w1 >> 32) & 3) : 0;
        const unsigned nr0 = __ballot_sync(0xffffffffu, s0 == 0), in0 = __ballot_sync(0xffffffffu, s0 == 2);
        const unsigned nr1 = __ballot_sync(0xffffffffu, s1 == 0), in1 = __ballot_sync(0xffffffffu, s1 == 2);
        int val; bool done = false; int step = 0;
        if (in0) {  // an inclusive prefix among the 32 nearest: lanes 0..f of group 0 close the walk
          const int f = __ffs(in0) - 1;
          const unsigned need = f >= 31 ? 0xffffffffu : ((2u << f) - 1u);
          if (nr0 & need) continue;
          val = lane <= f ? (int)(unsigned)w0 : 0; done = true;
        } else {
          if (nr0) continue;  // group 0 is all aggregates once everything there is published
          val = (int)(unsigned)w0; step = 32;
          const int f = in1 ? __ffs(in1) - 1 : 32;
          const unsigned need = f >= 31 ? 0xffffffffu : ((2u << f) - 1u);
          if (!(nr1 & need)) { val += lane <= f ? (int)(unsigned)w1 : 0; step = 64; done = f < 32; }
        }
        for (int d = 16; d > 0; d >>= 1) val += __shfl_xor_sync(0xffffffffu, val, d);
        excl += val;
        if (done) break;
        look -= step;
      }
    }
    if (lane == 0) { st[tile] = tag | (2ull << 32) | (unsigned)(excl + tot); sPrefix = excl; }
  }
  __syncthreads();
  int run = sPrefix + ex;
#pragma unroll
  for (int k = 0; k < 16; ++k) { const int t = v[k]; v[k] = run; run += t; }
  if (base + 15 < n) {
#pragma unroll
    for (int q = 0; q < 4; ++q) *reinterpret_cast<int4*>(out + base + 4 * q) = make_int4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
  } else {
#pragma unroll
    for (int k = 0; k < 16; ++k) if (base + k < n) out[base + k] = v[k];
  }
  if (tile == (int)gridDim.x - 1 && threadIdx.x == 0) out[n] = sPrefix + tot;
}
int vl_scan_alloc(VlScan* sc, int n) {
  const int tiles = vl_div_up(n, SCAN_TILE);
  if (sc->state && sc->tiles == tiles) return VLOAM_OK;
  vl_scan_free(sc);
  if (cudaMalloc(&sc->state, sizeof(unsigned long long) * (tiles + 1)) != cudaSuccess) return VLOAM_E_CUDA;
  if (cudaMemset(sc->state, 0, sizeof(unsigned long long) * (tiles + 1)) != cudaSuccess) return VLOAM_E_CUDA;  // epoch 0 = never used
  if (cudaDeviceSynchronize() != cudaSuccess) return VLOAM_E_CUDA;  // (first use only) the memset is not ordered with the non-blocking streams
  sc->tiles = tiles; sc->epoch = 0;
  return VLOAM_OK;
}
void vl_scan_free(VlScan* sc) { if (sc->state) cudaFree(sc->state); sc->state = nullptr; sc->tiles = 0; }
int vl_scan_exclusive(vloam_b200_ctx* c, const int* in, int n, VlScan* sc, int* out, const int* d_skip) {
  if (!sc->state || sc->tiles != vl_div_up(n, SCAN_TILE) || (((uintptr_t)in | (uintptr_t)out) & 15)) return VLOAM_E_INVALID;
  if (sc->epoch >= (1u << 30) - 1) {  // the 30-bit epoch wraps: forget every old status word first
    VL_CUDA(cudaMemsetAsync(sc->state, 0, sizeof(unsigned long long) * (sc->tiles + 1), VL_STREAM(c)));
    sc->epoch = 0;
  }
  ++sc->epoch;
  // (a scan over cell counters is search-structure overhead: SURVEY 8(d) assigns it no algorithmic bytes -> achieved GB/s is not reported for it)
  VL_LAUNCH(lm_scan_chained, sc->tiles, 256, 0, in, n, sc->state, sc->epoch, out, d_skip);
  return VLOAM_OK;
}

__device__ __forceinline__ void lm_dev_fill(const LmSub* __restrict__ sub, const float4* __restrict__ mapC, const float4* __restrict__ mapS,
                                            const int* __restrict__ cellOfPoint, const int* __restrict__ cellStart, int* __restrict__ cellFill,
                                            float4* __restrict__ sortedPts) {
  const int mc = sub->Mc, total = sub->Mc + sub->Ms;
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < total; g += gridDim.x * blockDim.x) {
    const int kind = g >= mc;
    const int id = kind ? g - mc : g;  // canonical index inside its own sub-map cloud
    const float4 p = kind ? mapS[id] : mapC[id];
    const int cell = cellOfPoint[g];
    const int pos = cellStart[cell] + atomicAdd(&cellFill[cell], 1);
    sortedPts[pos] = make_float4(p.x, p.y, p.z, __int_as_float(id));
  }
}
__global__ void __launch_bounds__(256) lm_grid_fill(const LmSub* __restrict__ sub, const int* __restrict__ skip, const float4* __restrict__ mapC,
                                                    const float4* __restrict__ mapS, const int* __restrict__ cellOfPoint,
                                                    const int* __restrict__ cellStart, int* __restrict__ cellFill,
                                                    float4* __restrict__ sortedPts) {
  VL_PDL_WAIT();

  if (skip && *skip) return;
  lm_dev_fill(sub, mapC, mapS, cellOfPoint, cellStart, cellFill, sortedPts);
}

// The whole in-line sub-map build (gather, zero, count, scan, fill) as ONE cooperative launch.  It runs every
// frame on the main stream but does real work only when the window moved: with the speculative build valid it
// is a single grid that returns at once, where eight separately launched early-exit kernels cost ~20 us of
// dependent launch latency.  Grid barriers separate the phases when it does run.
struct LmInlineArgs {
  const LmSub* sub; const int* skip; const MapCubeTable* tc; const MapCubeTable* ts; const float4* poolC; const float4* poolS;
  float4* outC; float4* outS; int* cellCount; int* cellFill; int* cellStart; int* tileSum; int* cellOfPoint; float4* sortedPts; int nCells;
};
__global__ void __launch_bounds__(256) lm_inline_build(LmInlineArgs a) {
  if (*a.skip) return;  // uniform over the grid
  cg::grid_group grid = cg::this_grid();
  lm_dev_gather(a.sub, a.tc, a.ts, a.poolC, a.poolS, a.outC, a.outS);
  lm_dev_zero(a.cellCount, a.cellFill, a.nCells + 1);
  grid.sync();
  lm_dev_count(a.sub, a.outC, a.outS, a.cellCount, a.cellOfPoint);
  grid.sync();
  const int nTiles = (a.nCells + 1023) / 1024;
  for (int tile = blockIdx.x; tile < nTiles; tile += gridDim.x) lm_dev_scan_tile(a.cellCount, a.nCells, a.tileSum, tile);
  grid.sync();
  if (blockIdx.x == 0) lm_dev_scan_sums<256>(a.tileSum, nTiles);
  grid.sync();
  for (int tile = blockIdx.x; tile < nTiles; tile += gridDim.x) lm_dev_scan_apply(a.cellCount, a.nCells, a.tileSum, a.cellStart, tile, nTiles);
  grid.sync();
  lm_dev_fill(a.sub, a.outC, a.outS, a.cellOfPoint, a.cellStart, a.cellFill, a.sortedPts);
}

// ---- fits -------------------------------------------------------------------------------------
// Eigen::SelfAdjointEigenSolver<Matrix3d> contract (ascending eigenvalues, unit eigenvectors) by the cyclic Jacobi method: branch-light
// and register-resident.  The oracle restates Eigen's own tridiagonal-QR algorithm (oracle_math.cpp), so the two are independent witnesses.
__device__ void lm_sym_eig3(double a[3][3], double evals[3], double evecs[3][3]) {
  double v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int sweep = 0; sweep < 30; ++sweep) {
    const double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
    if (off == 0.0) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (a[p][q] == 0.0) continue;
        const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double cs = 1.0 / sqrt(t * t + 1.0), sn = t * cs;
        const double apq = a[p][q];
        a[p][p] -= t * apq;
        a[q][q] += t * apq;
        a[p][q] = a[q][p] = 0.0;
        const int r = 3 - p - q;
        const double arp = a[r][p], arq = a[r][q];
        a[r][p] = a[p][r] = cs * arp - sn * arq;
        a[r][q] = a[q][r] = sn * arp + cs * arq;
        for (int k = 0; k < 3; ++k) {
          const double vkp = v[k][p], vkq = v[k][q];
          v[k][p] = cs * vkp - sn * vkq;
          v[k][q] = sn * vkp + cs * vkq;
        }
      }
  }
  int ord[3] = {0, 1, 2};  // stable insertion sort by eigenvalue (matches std::sort on 3 distinct keys)
  for (int i = 1; i < 3; ++i) for (int j = i; j > 0 && a[ord[j]][ord[j]] < a[ord[j - 1]][ord[j - 1]]; --j) { const int t = ord[j]; ord[j] = ord[j - 1]; ord[j - 1] = t; }
  for (int cidx = 0; cidx < 3; ++cidx) {
    evals[cidx] = a[ord[cidx]][ord[cidx]];
    for (int r = 0; r < 3; ++r) evecs[r][cidx] = v[r][ord[cidx]];
  }
}

// ColPivHouseholderQR<5x3>::solve restated exactly as oracle_math.cpp does.
__device__ void lm_qr_solve_5x3(double A[5][3], double b[5], double x[3]) {
  int perm[3] = {0, 1, 2};
  double maxpivot = 0.0, diag[3] = {0, 0, 0};
  for (int k = 0; k < 3; ++k) {
    int best = k; double bestn = -1.0;
    for (int j = k; j < 3; ++j) { double n2 = 0; for (int i = k; i < 5; ++i) n2 += A[i][j] * A[i][j]; if (n2 > bestn) { bestn = n2; best = j; } }
    if (best != k) { for (int i = 0; i < 5; ++i) { const double t = A[i][k]; A[i][k] = A[i][best]; A[i][best] = t; } const int t = perm[k]; perm[k] = perm[best]; perm[best] = t; }
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
        double sm = 0; for (int i = k; i < 5; ++i) sm += v[i] * A[i][j];
        sm *= tau; for (int i = k; i < 5; ++i) A[i][j] -= sm * v[i];
      }
      double sm = 0; for (int i = k; i < 5; ++i) sm += v[i] * b[i];
      sm *= tau; for (int i = k; i < 5; ++i) b[i] -= sm * v[i];
    }
    A[k][k] = beta; for (int i = k + 1; i < 5; ++i) A[i][k] = 0;
    diag[k] = beta;
    if (fabs(beta) > maxpivot) maxpivot = fabs(beta);
  }
  int rank = 0;
  for (int k = 0; k < 3; ++k) if (fabs(diag[k]) > DBL_EPSILON * 3.0 * maxpivot) ++rank;
  double y[3] = {0, 0, 0};
  for (int k = rank - 1; k >= 0; --k) { double sm = b[k]; for (int j = k + 1; j < rank; ++j) sm -= A[k][j] * y[j]; y[k] = sm / A[k][k]; }
  for (int k = 0; k < 3; ++k) x[perm[k]] = y[k];
}

// One warp per downsampled feature: 5-NN in the 3x3x3 cell neighbourhood (exact inside the 1 m
// acceptance ball), ordered by (d2, canonical id); then the line / plane fit.
__global__ void __launch_bounds__(256) lm_knn(const LmScalars* __restrict__ s, const RfWork* __restrict__ w,
                                                  const float4* __restrict__ stackC, const float4* __restrict__ stackS,
                                                  const float4* __restrict__ mapC, const float4* __restrict__ mapS,
                                                  const int* __restrict__ cellStart, const float4* __restrict__ sortedPts,
                                                  int* __restrict__ knnIdx, float* __restrict__ knnD2, int* __restrict__ knnOk,
                                                  double* __restrict__ factors, int* __restrict__ valid) {
  VL_PDL_WAIT();

  const int lane = threadIdx.x & 31;
  const int Qc = s->Qc, Qs = s->Qs, opt = s->optimized;
  double pose[7];  // loaded with the counts: one memory latency instead of two
#pragma unroll
  for (int k = 0; k < 7; ++k) pose[k] = s->pose[k];
  const float ox = w->gridOrigin[0], oy = w->gridOrigin[1], oz = w->gridOrigin[2];
  if (!opt) return;
  const int nWarps = (gridDim.x * blockDim.x) >> 5;
  for (int qi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; qi < Qc + Qs; qi += nWarps) {
  const int kind = qi >= Qc;
  const float4 po = kind ? stackS[qi - Qc] : stackC[qi];
  double r[3];
  vl_qrot(pose, (double)po.x, (double)po.y, (double)po.z, r);  // pointAssociateToMap (LM.cpp:154-164)
  const float sx = (float)(r[0] + pose[4]), sy = (float)(r[1] + pose[5]), sz = (float)(r[2] + pose[6]);
  const int cx = (int)floorf(__fmul_rn(__fsub_rn(sx, ox), 1.0f / LM_CELL));
  const int cy = (int)floorf(__fmul_rn(__fsub_rn(sy, oy), 1.0f / LM_CELL));
  const int cz = (int)floorf(__fmul_rn(__fsub_rn(sz, oz), 1.0f / LM_CELL));
  float bd[5]; int bi[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) { bd[k] = CUDART_INF_F; bi[k] = 0x7fffffff; }
  {  // 9 (y,z) rows; the 3 x-adjacent cells of a row are contiguous.  Lane r < 9 looks up row r.
    int rb = 0, rl = 0;
    if (lane < 9) {
      const int yy = cy - 1 + lane % 3, zz = cz - 1 + lane / 3;
      const int x0 = max(cx - 1, 0), x1 = min(cx + 1, LM_GX - 1);
      if (yy >= 0 && yy < LM_GY && zz >= 0 && zz < LM_GZ && x0 <= x1) {
        const int c0 = kind * LM_NCELL + x0 + LM_GX * (yy + LM_GY * zz);
        rb = cellStart[c0];
        rl = cellStart[c0 + (x1 - x0) + 1] - rb;
      }
    }
    VL_WARP_VISIT_FLAT(rb, rl, lane, sortedPts, {
      const float d = vl_dist2(sx, sy, sz, t.x, t.y, t.z);
      const int id = __float_as_int(t.w);
      if (d < bd[4] || (d == bd[4] && id < bi[4])) {  // insertion into the lane-local sorted top-5
        bd[4] = d; bi[4] = id;
#pragma unroll
        for (int k = 4; k > 0; --k)
          if (bd[k] < bd[k - 1] || (bd[k] == bd[k - 1] && bi[k] < bi[k - 1])) {
            const float td = bd[k]; bd[k] = bd[k - 1]; bd[k - 1] = td;
            const int ti = bi[k]; bi[k] = bi[k - 1]; bi[k - 1] = ti;
          }
      }
    });
  }
  // merge the 32 lane-local lists: pop the global minimum of (d2, id) five times.  d2 >= 0, so its bit pattern
  // orders like the value: two hardware warp reductions per pop (smallest d2, then the smallest id holding it)
  float nd[5]; int ni[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const unsigned hb = __float_as_uint(bd[0]);  // +inf (empty list) sorts last
    const unsigned gmin = __reduce_min_sync(0xffffffffu, hb);
    const unsigned imin = __reduce_min_sync(0xffffffffu, hb == gmin ? (unsigned)bi[0] : 0x7fffffffu);
    nd[k] = __uint_as_float(gmin); ni[k] = (int)imin;
    if (hb == gmin && (unsigned)bi[0] == imin && imin != 0x7fffffffu) {
#pragma unroll
      for (int q = 0; q < 4; ++q) { bd[q] = bd[q + 1]; bi[q] = bi[q + 1]; }
      bd[4] = CUDART_INF_F; bi[4] = 0x7fffffff;
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 5; ++k) { knnIdx[qi * 5 + k] = (ni[k] == 0x7fffffff) ? -1 : ni[k]; knnD2[qi * 5 + k] = nd[k]; }
  }
  }
}

// The two fits of LM.cpp:559-603 / 637-680 on five neighbours (f64 from f32 coordinates).  o6: edge -> {a, b} (the two
// synthetic line points), plane -> {unit normal, d, 0, 0}.  Returns the accept flag.
__device__ __forceinline__ bool lm_fit_one(int kind, double P[5][3], double o6[6]) {
  if (!kind) {  // LM.cpp:559-603: PCA line test
    double cen[3] = {0, 0, 0};
    for (int j = 0; j < 5; ++j) for (int k = 0; k < 3; ++k) cen[k] = cen[k] + P[j][k];
    for (int k = 0; k < 3; ++k) cen[k] = cen[k] / 5.0;
    double cov[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int j = 0; j < 5; ++j) {
      const double z[3] = {P[j][0] - cen[0], P[j][1] - cen[1], P[j][2] - cen[2]};
      for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) cov[a][b] = cov[a][b] + z[a] * z[b];
    }
    double ev[3], evec[3][3];
    lm_sym_eig3(cov, ev, evec);
    if (!(ev[2] > 3 * ev[1])) return false;
    for (int k = 0; k < 3; ++k) { const double u = evec[k][2]; o6[k] = 0.1 * u + cen[k]; o6[3 + k] = -0.1 * u + cen[k]; }
    return true;
  }
  // LM.cpp:637-680: least-squares plane n.p + 1 = 0, 0.2 m flatness check
  double A[5][3], B[5] = {-1, -1, -1, -1, -1}, nrm[3];
  for (int j = 0; j < 5; ++j) for (int k = 0; k < 3; ++k) A[j][k] = P[j][k];
  lm_qr_solve_5x3(A, B, nrm);
  const double nn = sqrt(nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2]);
  const double d = 1 / nn;
  if (nn > 0) { nrm[0] /= nn; nrm[1] /= nn; nrm[2] /= nn; }
  for (int j = 0; j < 5; ++j)
    if (fabs(nrm[0] * P[j][0] + nrm[1] * P[j][1] + nrm[2] * P[j][2] + d) > 0.2) return false;
  o6[0] = nrm[0]; o6[1] = nrm[1]; o6[2] = nrm[2]; o6[3] = d; o6[4] = 0; o6[5] = 0;
  return true;
}

// Line / plane fit of one feature per THREAD (the f64 eigen / QR work of 32 features shares a warp's
// issue slots instead of idling 31 lanes behind lane 0 of the search kernel).
__global__ void __launch_bounds__(128) lm_fit(const LmScalars* __restrict__ s, const float4* __restrict__ stackC, const float4* __restrict__ stackS,
                                              const float4* __restrict__ mapC, const float4* __restrict__ mapS, const int* __restrict__ knnIdx,
                                              const float* __restrict__ knnD2, int* __restrict__ knnOk, double* __restrict__ factors,
                                              int* __restrict__ valid) {
  VL_PDL_WAIT();

  const int Qc = s->Qc, Qs = s->Qs;
  if (!s->optimized) return;
  for (int qi = blockIdx.x * blockDim.x + threadIdx.x; qi < Qc + Qs; qi += gridDim.x * blockDim.x) {
  const int kind = qi >= Qc;
  const float4 po = kind ? stackS[qi - Qc] : stackC[qi];
  int ni[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) ni[k] = knnIdx[qi * 5 + k];
  const bool have5 = ni[4] >= 0;
  const float d4 = knnD2[qi * 5 + 4];
  bool ok = false;
  double* f = factors + (size_t)qi * 10;
  if (have5 && (double)d4 < 1.0) {
    const float4* map = kind ? mapS : mapC;
    double P[5][3];
#pragma unroll
    for (int j = 0; j < 5; ++j) { const float4 t = map[ni[j]]; P[j][0] = t.x; P[j][1] = t.y; P[j][2] = t.z; }
    double o6[6];
    ok = lm_fit_one(kind, P, o6);
    if (ok) {
      f[0] = kind ? 2.0 : 0.0; f[1] = po.x; f[2] = po.y; f[3] = po.z;
#pragma unroll
      for (int k = 0; k < 6; ++k) f[4 + k] = o6[k];
    }
  }
  valid[qi] = ok ? 1 : 0;
  knnOk[qi] = ok ? 1 : 0;
  }
}

// vloam_b200_fit: the same fits on caller-supplied five-point sets (parity tests against an independent witness)
__global__ void __launch_bounds__(128) lm_fit_sets(const float* __restrict__ near, int n, int kind, int* __restrict__ ok, double* __restrict__ prm) {
  VL_PDL_WAIT();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double P[5][3], o6[6] = {0, 0, 0, 0, 0, 0};
  for (int j = 0; j < 5; ++j) for (int k = 0; k < 3; ++k) P[j][k] = near[(size_t)i * 15 + j * 3 + k];
  const bool a = lm_fit_one(kind, P, o6);
  ok[i] = a ? 1 : 0;
  for (int k = 0; k < 6; ++k) prm[(size_t)i * 6 + k] = a ? o6[k] : 0.0;
}
int vl_lm_fit_sets(vloam_b200_ctx* c, const float* d_near, int n, int kind, int* d_ok, double* d_prm) {
  VL_LAUNCH(lm_fit_sets, vl_div_up(max(n, 1), 128), 128, 0, d_near, n, kind, d_ok, d_prm);
  VL_CUDA(cudaGetLastError());
  return VLOAM_OK;
}

__global__ void lm_transform_update(LmScalars* s) {
  VL_PDL_WAIT();
  // LM.cpp:147-151
  if (threadIdx.x != 0) return;
  const double* q = s->q_wodom;
  const double n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  const double qi[4] = {-q[0] / n2, -q[1] / n2, -q[2] / n2, q[3] / n2};
  vl_qmul(s->pose, qi, s->q_wmap_wodom);
  double r[3];
  vl_qrot(s->q_wmap_wodom, s->t_wodom[0], s->t_wodom[1], s->t_wodom[2], r);
  for (int k = 0; k < 3; ++k) s->t_wmap_wodom[k] = s->pose[4 + k] - r[k];
}

// ---- map update: insert (LM.cpp:741-788) + per-cube VoxelGrid re-filter (LM.cpp:795-808) -------
// sort key: [63:56] segment (kind*125 + valid slot) | [55:26] voxel (iz,iy,ix) | [25:0] order
// order: existing tail point -> its position in the cube; new point -> 2^25 + stack index.
__device__ __forceinline__ float lm_leaf_inv(const vloam_b200_params& p, int kind) { return __fdiv_rn(1.0f, kind ? p.plane_res : p.line_res); }

__global__ void __launch_bounds__(256) rf_keys(const LmScalars* __restrict__ s, RfWork* __restrict__ w, vloam_b200_params prm,
                                               const MapCubeTable* __restrict__ tc, const MapCubeTable* __restrict__ ts,
                                               const float4* __restrict__ poolC, const float4* __restrict__ poolS,
                                               const float4* __restrict__ stackC, const float4* __restrict__ stackS,
                                               float4* __restrict__ newPts, int* __restrict__ newCube,
                                               unsigned long long* __restrict__ keys, int P) {
  VL_PDL_WAIT();

  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= P) return;
  const int tailTotal = w->tailOff[LM_NSEG];
  const int Qc = s->Qc, Qs = s->Qs;
  unsigned long long key = ~0ull;
  if (g < tailTotal) {  // an existing tail point of a valid cube
    int lo = 0, hi = LM_NSEG;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (w->tailOff[mid] <= g) lo = mid; else hi = mid; }
    const int kind = lo / VL_MAX_VALID, slot = lo % VL_MAX_VALID;
    const int cb = s->validInd[slot];
    const MapCubeTable* t = kind ? ts : tc;
    const int pos = t->sorted[cb] + (g - w->tailOff[lo]);
    const float4 p = (kind ? poolS : poolC)[t->start[cb] + pos];
    const unsigned vk = lm_vox_key_cube(p, lm_leaf_inv(prm, kind), cb, s);
    key = ((unsigned long long)lo << 56) | ((unsigned long long)vk << 26) | (unsigned long long)pos;
  } else if (g < tailTotal + Qc + Qs) {  // this frame's point i of the stacks
    const int i = g - tailTotal;
    const int kind = i >= Qc;
    const float4 po = kind ? stackS[i - Qc] : stackC[i];
    double r[3];
    vl_qrot(s->pose, (double)po.x, (double)po.y, (double)po.z, r);
    const float4 p = make_float4((float)(r[0] + s->pose[4]), (float)(r[1] + s->pose[5]), (float)(r[2] + s->pose[6]), po.w);
    const int ci = lm_cube_of(p.x, s->cenW), cj = lm_cube_of(p.y, s->cenH), ck = lm_cube_of(p.z, s->cenD);
    int cb = -1;
    if (ci >= 0 && ci < VL_CUBE_W && cj >= 0 && cj < VL_CUBE_H && ck >= 0 && ck < VL_CUBE_D) cb = ci + VL_CUBE_W * cj + VL_CUBE_W * VL_CUBE_H * ck;
    newPts[i] = p;
    const int slot = cb >= 0 ? w->slotOfCube[cb] : -1;
    newCube[i] = (cb >= 0 && slot < 0) ? cb : -1;  // only cubes outside the window need the raw append path
    if (slot >= 0) {
      const unsigned vk = lm_vox_key(p, lm_leaf_inv(prm, kind), ci, cj, ck, s->cenW, s->cenH, s->cenD);
      key = ((unsigned long long)(kind * VL_MAX_VALID + slot) << 56) | ((unsigned long long)vk << 26) | (unsigned long long)((1u << 25) + (unsigned)i);
    }
  }
  keys[g] = key;
  if (key != ~0ull) atomicAdd(&w->segCount[(int)(key >> 56)], 1);
}

// ---- segmented sort of the update keys ----------------------------------------------------------------
// The keys are (segment | voxel | order) and every later step works per segment (one segment = one valid cube
// of one kind).  Instead of one bitonic network over all keys (37 us for 8k keys: ~48 barrier rounds of a
// 1024-thread CTA pair), the keys are bucketed by segment -- histogram in rf_keys, a 250-entry scan that IS the
// per-segment range table, one scatter -- and each bucket is sorted by its own CTA in shared memory.
// Every block of the scatter scans the 250 counts for itself (one 256-thread scan, ~1 us) instead of waiting for a
// separate single-block launch; block 0 publishes the range table for the later steps.  segFill is zeroed by lm_prepare.
__global__ void __launch_bounds__(256) rf_seg_scatter(const unsigned long long* __restrict__ in, int n, RfWork* __restrict__ w,
                                                      unsigned long long* __restrict__ out) {
  VL_PDL_WAIT();

  __shared__ int sb[256];
  __shared__ int sBegin[256];
  const int t = threadIdx.x;
  const int cnt = t < LM_NSEG ? w->segCount[t] : 0;
  int total = 0;
  const int off = lm_scan256(cnt, sb, &total);
  sBegin[t] = off;
  if (blockIdx.x == 0) {
    if (t < LM_NSEG) { w->tailBegin[t] = off; w->firstViolation[t] = INT_MAX; }
    if (t == 0) { w->tailBegin[LM_NSEG] = total; w->nKeysValid = total; }
  }
  __syncthreads();
  const int g = blockIdx.x * blockDim.x + t;
  if (g >= n) return;
  const unsigned long long key = in[g];
  if (key == ~0ull) return;
  const int sg = (int)(key >> 56);
  out[sBegin[sg] + atomicAdd(&w->segFill[sg], 1)] = key;
}
#define RF_SEG_THREADS 512
#define RF_SEG_CAP 8192  // keys of one segment sorted in shared memory; larger segments use `scratch` (global)
__global__ void __launch_bounds__(RF_SEG_THREADS) rf_seg_sort(unsigned long long* __restrict__ keys, const RfWork* __restrict__ w,
                                                              unsigned long long* __restrict__ scratch, int cap) {
  VL_PDL_WAIT();

  extern __shared__ unsigned long long sk[];
  const int sg = blockIdx.x;
  const int beg = w->tailBegin[sg], n = w->tailBegin[sg + 1] - beg;
  if (n <= 1) return;
  int P = 8; while (P < n) P <<= 1;
  if (P <= cap) {
    for (int t = threadIdx.x; t < P; t += RF_SEG_THREADS) sk[t] = t < n ? keys[beg + t] : ~0ull;
    __syncthreads();
    bt_smem_sort<RF_SEG_THREADS>(sk, P);
    for (int t = threadIdx.x; t < n; t += RF_SEG_THREADS) keys[beg + t] = sk[t];
  } else {  // P <= 2 n - 1: the segments' scratch areas [2 beg, 2 beg + P) do not overlap
    unsigned long long* g = scratch + (size_t)2 * beg;
    for (int t = threadIdx.x; t < P; t += RF_SEG_THREADS) g[t] = t < n ? keys[beg + t] : ~0ull;
    __syncthreads();
    bt_sort_batched<RF_SEG_THREADS>(g, P, 1);
    for (int t = threadIdx.x; t < n; t += RF_SEG_THREADS) keys[beg + t] = g[t];
  }
}

__device__ __forceinline__ float4 rf_key_point(unsigned long long key, const LmScalars* s, const MapCubeTable* tc, const MapCubeTable* ts,
                                               const float4* poolC, const float4* poolS, const float4* newPts) {
  const unsigned ord = (unsigned)(key & 0x3ffffffull);
  if (ord >= (1u << 25)) return newPts[ord - (1u << 25)];
  const int sg = (int)(key >> 56);
  const int kind = sg / VL_MAX_VALID, cb = s->validInd[sg % VL_MAX_VALID];
  return (kind ? poolS : poolC)[(kind ? ts : tc)->start[cb] + (int)ord];
}
#define RF_VOX(key) ((unsigned)(((key) >> 26) & 0x3fffffffull))

// For each sorted tail key: is it the head of a run that no prefix point owns?
__global__ void __launch_bounds__(256) rf_match(const unsigned long long* __restrict__ keys, const LmScalars* __restrict__ s,
                                                const RfWork* __restrict__ w, vloam_b200_params prm, const MapCubeTable* __restrict__ tc,
                                                const MapCubeTable* __restrict__ ts, const float4* __restrict__ poolC,
                                                const float4* __restrict__ poolS, int* __restrict__ unmatchedHead) {
  VL_PDL_WAIT();

  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= w->nKeysValid) return;
  const unsigned long long key = keys[t];
  const int sg = (int)(key >> 56);
  const unsigned vk = RF_VOX(key);
  const bool head = (t == w->tailBegin[sg]) || RF_VOX(keys[t - 1]) != vk;
  int um = 0;
  if (head) {
    const int kind = sg / VL_MAX_VALID, cb = s->validInd[sg % VL_MAX_VALID];
    const MapCubeTable* tb = kind ? ts : tc;
    const float4* pool = (kind ? poolS : poolC) + tb->start[cb];
    const float inv = lm_leaf_inv(prm, kind);
    int lo = 0, hi = tb->sorted[cb];
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (lm_vox_key_cube(pool[mid], inv, cb, s) < vk) lo = mid + 1; else hi = mid; }
    const bool matched = lo < tb->sorted[cb] && lm_vox_key_cube(pool[lo], inv, cb, s) == vk;
    um = matched ? 0 : 1;
  }
  unmatchedHead[t] = um;
}

__global__ void __launch_bounds__(1024) rf_scan_layout(int* __restrict__ unmatched, const LmScalars* __restrict__ s, RfWork* __restrict__ w,
                                                       const MapCubeTable* __restrict__ tc, const MapCubeTable* __restrict__ ts) {
  VL_PDL_WAIT();

  // exclusive scan of unmatched[0..n) in place, unmatched[n] = total; 8 consecutive flags per thread and round
  __shared__ int ws[32];
  const int n = w->nKeysValid;
  int carry = 0;  // the same value in every thread
  for (int base = 0; base < n; base += 8192) {
    const int b = base + threadIdx.x * 8;
    int v[8];
    if (b + 7 < n) {
      const int4 lo = *reinterpret_cast<const int4*>(unmatched + b), hi = *reinterpret_cast<const int4*>(unmatched + b + 4);
      v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w; v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = b + k < n ? unmatched[b + k] : 0;
    }
    int own = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) own += v[k];
    int tot = 0;
    int run = carry + vl_block_excl_scan<1024>(own, ws, &tot);
#pragma unroll
    for (int k = 0; k < 8; ++k) { const int t = v[k]; v[k] = run; run += t; }
    if (b + 7 < n) {
      *reinterpret_cast<int4*>(unmatched + b) = make_int4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<int4*>(unmatched + b + 4) = make_int4(v[4], v[5], v[6], v[7]);
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) if (b + k < n) unmatched[b + k] = v[k];
    }
    carry += tot;
  }
  if (threadIdx.x == 0) unmatched[n] = carry;
  __syncthreads();
  __shared__ int sb2[256];
  const int sg = threadIdx.x;
  int cnt = 0;
  if (sg < LM_NSEG) {
    const int kind = sg / VL_MAX_VALID, slot = sg % VL_MAX_VALID;
    if (slot < s->validNum) {
      const int cb = s->validInd[slot];
      cnt = (kind ? ts : tc)->sorted[cb] + (unmatched[w->tailBegin[sg + 1]] - unmatched[w->tailBegin[sg]]);
    }
  }
  int tot = 0;
  const int off = lm_scan256(cnt, sb2, &tot);
  if (sg < LM_NSEG) { w->outOff[sg] = off; w->outCount[sg] = cnt; }
  if (sg == 0) w->outOff[LM_NSEG] = tot;
}

__device__ __forceinline__ float4 rf_fold(float4 acc, const float4 p) {
  return make_float4(__fadd_rn(acc.x, p.x), __fadd_rn(acc.y, p.y), __fadd_rn(acc.z, p.z), __fadd_rn(acc.w, p.w));
}
__device__ __forceinline__ float4 rf_centroid(const float4 acc, int n) {
  const float fn = (float)n;
  return make_float4(__fdiv_rn(acc.x, fn), __fdiv_rn(acc.y, fn), __fdiv_rn(acc.z, fn), __fdiv_rn(acc.w, fn));
}

// new voxels: runs of tail keys that no prefix point owns
__device__ __forceinline__ void rf_dev_emit_new(int blk, const unsigned long long* __restrict__ keys, const int* __restrict__ uScan,
                                                const LmScalars* __restrict__ s, const RfWork* __restrict__ w, const vloam_b200_params& prm,
                                                const MapCubeTable* __restrict__ tc, const MapCubeTable* __restrict__ ts,
                                                const float4* __restrict__ poolC, const float4* __restrict__ poolS,
                                                const float4* __restrict__ newPts, float4* __restrict__ staging) {
  const int t = blk * blockDim.x + threadIdx.x;
  const int n = w->nKeysValid;
  if (t >= n) return;
  if (uScan[t + 1] == uScan[t]) return;  // not an unmatched head
  const unsigned long long key = keys[t];
  const int sg = (int)(key >> 56);
  const unsigned vk = RF_VOX(key);
  const int segEnd = w->tailBegin[sg + 1];
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  int cnt = 0;
  for (int q = t; q < segEnd && RF_VOX(keys[q]) == vk; ++q) { acc = rf_fold(acc, rf_key_point(keys[q], s, tc, ts, poolC, poolS, newPts)); ++cnt; }
  const int kind = sg / VL_MAX_VALID, cb = s->validInd[sg % VL_MAX_VALID];
  const MapCubeTable* tb = kind ? ts : tc;
  const float4* pool = (kind ? poolS : poolC) + tb->start[cb];
  const float inv = lm_leaf_inv(prm, kind);
  int lo = 0, hi = tb->sorted[cb];
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (lm_vox_key_cube(pool[mid], inv, cb, s) < vk) lo = mid + 1; else hi = mid; }
  staging[w->outOff[sg] + lo + (uScan[t] - uScan[w->tailBegin[sg]])] = rf_centroid(acc, cnt);
}

// prefix points: shifted by the new voxels that sort before them; absorb a matching tail run
__device__ __forceinline__ void rf_dev_emit_prefix(int blk, int nblk, const unsigned long long* __restrict__ keys, const int* __restrict__ uScan,
                                                   const LmScalars* __restrict__ s, const RfWork* __restrict__ w, const vloam_b200_params& prm,
                                                   const MapCubeTable* __restrict__ tc, const MapCubeTable* __restrict__ ts,
                                                   const float4* __restrict__ poolC, const float4* __restrict__ poolS,
                                                   const float4* __restrict__ newPts, float4* __restrict__ staging) {
  const int total = w->prefOff[LM_NSEG];
  for (int g = blk * blockDim.x + threadIdx.x; g < total; g += nblk * blockDim.x) {
    int lo = 0, hi = LM_NSEG;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (w->prefOff[mid] <= g) lo = mid; else hi = mid; }
    const int sg = lo, i = g - w->prefOff[sg];
    const int kind = sg / VL_MAX_VALID, cb = s->validInd[sg % VL_MAX_VALID];
    const MapCubeTable* tb = kind ? ts : tc;
    const float4 p = (kind ? poolS : poolC)[tb->start[cb] + i];
    const unsigned vk = lm_vox_key_cube(p, lm_leaf_inv(prm, kind), cb, s);
    int a = w->tailBegin[sg], b = w->tailBegin[sg + 1];
    const int segBegin = a, segEnd = b;
    while (a < b) { const int mid = (a + b) >> 1; if (RF_VOX(keys[mid]) < vk) a = mid + 1; else b = mid; }
    float4 acc = rf_fold(make_float4(0.f, 0.f, 0.f, 0.f), p);
    int cnt = 1;
    for (int q = a; q < segEnd && RF_VOX(keys[q]) == vk; ++q) { acc = rf_fold(acc, rf_key_point(keys[q], s, tc, ts, poolC, poolS, newPts)); ++cnt; }
    staging[w->outOff[sg] + i + (uScan[a] - uScan[segBegin])] = rf_centroid(acc, cnt);
  }
}

// Both emitters in one launch: they write disjoint staging entries and neither reads what the other writes, so the
// first nbNew blocks handle the new voxels (a few long per-run loops) while the rest stream the prefix points.
__global__ void __launch_bounds__(256) rf_emit(int nbNew, const unsigned long long* __restrict__ keys, const int* __restrict__ uScan,
                                               const LmScalars* __restrict__ s, const RfWork* __restrict__ w, vloam_b200_params prm,
                                               const MapCubeTable* __restrict__ tc, const MapCubeTable* __restrict__ ts,
                                               const float4* __restrict__ poolC, const float4* __restrict__ poolS,
                                               const float4* __restrict__ newPts, float4* __restrict__ staging) {
  VL_PDL_WAIT();

  if ((int)blockIdx.x < nbNew) rf_dev_emit_new(blockIdx.x, keys, uScan, s, w, prm, tc, ts, poolC, poolS, newPts, staging);
  else rf_dev_emit_prefix(blockIdx.x - nbNew, gridDim.x - nbNew, keys, uScan, s, w, prm, tc, ts, poolC, poolS, newPts, staging);
}

// grow cube storage where the re-filtered cloud no longer fits (bump allocation from the pool top;
// a block scan keeps the layout deterministic)
__global__ void __launch_bounds__(256) rf_alloc(LmScalars* __restrict__ s, const RfWork* __restrict__ w, MapCubeTable* __restrict__ tc,
                                                MapCubeTable* __restrict__ ts, int poolCapC, int poolCapS) {
  VL_PDL_WAIT();

  __shared__ int sb[256];
  const int sg = threadIdx.x;
  const int kind = sg / VL_MAX_VALID, slot = sg % VL_MAX_VALID;
  int ncapC = 0, ncapS = 0, cb = -1;
  if (sg < LM_NSEG && slot < s->validNum) {
    cb = s->validInd[slot];
    const MapCubeTable* tb = kind ? ts : tc;
    const int need = w->outCount[sg];
    if (need > tb->cap[cb]) { if (kind) ncapS = max(2 * need, 256); else ncapC = max(2 * need, 256); }
  }
  int totC = 0, totS = 0;
  const int offC = lm_scan256(ncapC, sb, &totC);
  const int offS = lm_scan256(ncapS, sb, &totS);
  const int topC = s->poolTopC, topS = s->poolTopS;
  const bool okC = topC + totC <= poolCapC, okS = topS + totS <= poolCapS;
  if (ncapC > 0 && okC) { tc->start[cb] = topC + offC; tc->cap[cb] = ncapC; }
  if (ncapS > 0 && okS) { ts->start[cb] = topS + offS; ts->cap[cb] = ncapS; }
  __syncthreads();
  if (sg == 0) {
    if (okC) s->poolTopC = topC + totC; else s->overflow = 1;
    if (okS) s->poolTopS = topS + totS; else s->overflow = 1;
  }
}

__global__ void __launch_bounds__(256) rf_commit(const float4* __restrict__ staging, LmScalars* __restrict__ s, RfWork* __restrict__ w,
                                                 vloam_b200_params prm, MapCubeTable* __restrict__ tc, MapCubeTable* __restrict__ ts,
                                                 float4* __restrict__ poolC, float4* __restrict__ poolS) {
  VL_PDL_WAIT();

  const int total = w->outOff[LM_NSEG];
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < total; g += gridDim.x * blockDim.x) {
    int lo = 0, hi = LM_NSEG;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (w->outOff[mid] <= g) lo = mid; else hi = mid; }
    const int sg = lo, k = g - w->outOff[sg];
    const int kind = sg / VL_MAX_VALID, cb = s->validInd[sg % VL_MAX_VALID];
    MapCubeTable* tb = kind ? ts : tc;
    if (w->outCount[sg] > tb->cap[cb]) continue;  // pool exhausted (overflow flag is set)
    const float4 p = staging[g];
    (kind ? poolS : poolC)[tb->start[cb] + k] = p;
    if (k > 0) {  // does the re-filtered cloud keep one point per voxel in ascending order?
      const float inv = lm_leaf_inv(prm, kind);
      if (lm_vox_key_cube(staging[g - 1], inv, cb, s) >= lm_vox_key_cube(p, inv, cb, s)) atomicMin(&w->firstViolation[sg], k);
    }
  }
}

__global__ void rf_finish(LmScalars* __restrict__ s, const RfWork* __restrict__ w, MapCubeTable* __restrict__ tc, MapCubeTable* __restrict__ ts) {
  VL_PDL_WAIT();

  const int sg = threadIdx.x;
  if (sg >= LM_NSEG) return;
  const int kind = sg / VL_MAX_VALID, slot = sg % VL_MAX_VALID;
  if (slot >= s->validNum) return;
  const int cb = s->validInd[slot];
  MapCubeTable* tb = kind ? ts : tc;
  if (w->outCount[sg] > tb->cap[cb]) return;
  tb->count[cb] = w->outCount[sg];
  tb->sorted[cb] = min(w->outCount[sg], w->firstViolation[sg]);
}

// Points that land in a cube outside the 5x5x3 window are appended raw, in stack order (LM.cpp:762, 786).
// One CTA walks the stacks in chunks of 1024: outside points are compacted in order, ranked among
// the points of the same (kind, cube) inside the chunk, cube storage is grown where needed, and every
// point is written to count + rank.  The common case (no outside point at all) is one pass of flags.
__global__ void __launch_bounds__(1024) rf_append_outside(LmScalars* __restrict__ s, const float4* __restrict__ newPts,
                                                          const int* __restrict__ newCube, MapCubeTable* __restrict__ tc,
                                                          MapCubeTable* __restrict__ ts, float4* __restrict__ poolC, float4* __restrict__ poolS,
                                                          int poolCapC, int poolCapS) {
  VL_PDL_WAIT();

  const int Qc = s->Qc, total = s->Qc + s->Qs;
  __shared__ int lIdx[1024], lKey[1024];
  __shared__ int gKey[1024], gOld[1024], gNew[1024], gCnt[1024];
  __shared__ int warpSum[32];
  __shared__ int nGrow;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int any = 0;
  for (int i = threadIdx.x; i < total; i += 1024) any |= (newCube[i] >= 0);
  if (__syncthreads_or(any) == 0) return;
  for (int base = 0; base < total; base += 1024) {
    const int i = base + threadIdx.x;
    const int cb = i < total ? newCube[i] : -1;
    const bool f = cb >= 0;
    const unsigned bal = __ballot_sync(0xffffffffu, f);
    if (lane == 0) warpSum[warp] = __popc(bal);
    if (threadIdx.x == 0) nGrow = 0;
    __syncthreads();
    int before = 0, n = 0;
    for (int w = 0; w < 32; ++w) { const int v = warpSum[w]; if (w < warp) before += v; n += v; }
    if (f) { const int pos = before + __popc(bal & ((1u << lane) - 1u)); lIdx[pos] = i; lKey[pos] = (i >= Qc ? VL_CUBE_NUM : 0) + cb; }
    __syncthreads();
    if (n == 0) continue;
    // rank inside the chunk among entries of the same (kind, cube); the first one is the leader
    const int e = threadIdx.x;
    int key = -1, rank = 0, tot = 0;
    if (e < n) {
      key = lKey[e];
      for (int q = 0; q < n; ++q) { const bool same = lKey[q] == key; rank += (same && q < e); tot += same; }
      if (rank == 0) {
        const int kind = key >= VL_CUBE_NUM, c2 = key - kind * VL_CUBE_NUM;
        const MapCubeTable* tb = kind ? ts : tc;
        if (tb->count[c2] + tot > tb->cap[c2]) { const int g = atomicAdd(&nGrow, 1); gKey[g] = key; gCnt[g] = tb->count[c2] + tot; }
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int g = 0; g < nGrow; ++g) {
        const int kind = gKey[g] >= VL_CUBE_NUM, c2 = gKey[g] - kind * VL_CUBE_NUM;
        MapCubeTable* tb = kind ? ts : tc;
        const int ncap = max(2 * gCnt[g], 256);
        int& top = kind ? s->poolTopS : s->poolTopC;
        gOld[g] = tb->start[c2];
        if (top + ncap > (kind ? poolCapS : poolCapC)) { s->overflow = 1; gNew[g] = -1; continue; }
        gNew[g] = top; tb->start[c2] = top; tb->cap[c2] = ncap; top += ncap;
      }
    }
    __syncthreads();
    for (int g = 0; g < nGrow; ++g) {
      if (gNew[g] < 0) continue;
      const int kind = gKey[g] >= VL_CUBE_NUM, c2 = gKey[g] - kind * VL_CUBE_NUM;
      float4* pool = kind ? poolS : poolC;
      const int cnt = (kind ? ts : tc)->count[c2];
      for (int q = threadIdx.x; q < cnt; q += 1024) pool[gNew[g] + q] = pool[gOld[g] + q];
    }
    __syncthreads();
    if (e < n) {
      const int kind = key >= VL_CUBE_NUM, c2 = key - kind * VL_CUBE_NUM;
      MapCubeTable* tb = kind ? ts : tc;
      if (tb->count[c2] + tot <= tb->cap[c2]) (kind ? poolS : poolC)[tb->start[c2] + tb->count[c2] + rank] = newPts[lIdx[e]];
    }
    __syncthreads();
    if (e < n && rank == 0) {
      const int kind = key >= VL_CUBE_NUM, c2 = key - kind * VL_CUBE_NUM;
      MapCubeTable* tb = kind ? ts : tc;
      if (tb->count[c2] + tot <= tb->cap[c2]) tb->count[c2] += tot;
    }
    __syncthreads();
  }
}

// after an import: longest strictly increasing voxel-key prefix of every cube
__global__ void __launch_bounds__(256) lm_scan_sorted(const LmScalars* __restrict__ s, vloam_b200_params prm, MapCubeTable* __restrict__ t,
                                                      const float4* __restrict__ pool, int kind) {
  VL_PDL_WAIT();

  const int cb = blockIdx.x;
  __shared__ int firstBad;
  if (threadIdx.x == 0) firstBad = INT_MAX;
  __syncthreads();
  const int n = t->count[cb];
  const float inv = lm_leaf_inv(prm, kind);
  const float4* p = pool + t->start[cb];
  for (int k = 1 + threadIdx.x; k < n; k += blockDim.x)
    if (lm_vox_key_cube(p[k - 1], inv, cb, s) >= lm_vox_key_cube(p[k], inv, cb, s)) atomicMin(&firstBad, k);
  __syncthreads();
  if (threadIdx.x == 0) t->sorted[cb] = min(n, firstBad);
}

__global__ void lm_set_counts(LmScalars* s, RfWork* w, const int* qc, const int* qs) {
  VL_PDL_WAIT();

  if (threadIdx.x != 0) return;
  s->Qc = *qc; s->Qs = *qs;
  // LM.cpp:514: optimise only against a sub-map with > 10 corner and > 50 surf points
  s->optimized = (s->Mc > 10 && s->Ms > 50 && *qc + *qs > 0) ? 1 : 0;
  w->nq = s->optimized ? *qc + *qs : 0;  // factor slots the solver looks at
}

// -----------------------------------------------------------------------------------------------
#define LM_POOL_C (16 << 20)
#define LM_POOL_S (48 << 20)

struct LmDevice {  // extra device state owned by this file
  RfWork* work;
  int* cellCount;   // 2*LM_NCELL + 1 (scan output in place: cellStart)
  int* cellStart;
  int* cellFill;
  int* tileSum;  // in-line (cooperative) build only
  VlScan scan;   // speculative build: single-launch scan
  DBuf<int> cellOfPoint;
  DBuf<float4> sortedPts;
  DBuf<float4> newPts; DBuf<int> newCube;
  DBuf<int> unmatched;
  int* dQ;          // 2 x two device ints: Qc, Qs from the voxel filters (pair ctx::stackSel is current)
  long long hMapUpperC, hMapUpperS;  // host upper bounds on the total map size
  LmSub* subReal; LmSub* subSpec;    // sub-map window descriptors (this frame's / the speculative one)
  int* specOK;                       // device flag written by lm_prepare: the speculative sub-map is this frame's
  bool specQueued;                   // host: a speculative build was queued after the last map update and nothing touched the map since
  bool specEnabled;
  int inlineGrid;                    // co-resident grid of lm_inline_build (cooperative launch)
};
static LmDevice* lmdev(vloam_b200_ctx* c) { return reinterpret_cast<LmDevice*>(c->gridPrm); }

int vl_lm_init(vloam_b200_ctx* c) {
  LmDevice* d = new LmDevice();
  c->gridPrm = reinterpret_cast<GridParams*>(d);
  VL_CUDA(cudaMalloc(&d->work, sizeof(RfWork)));
  VL_CUDA(cudaMemset(d->work, 0, sizeof(RfWork)));
  VL_CUDA(cudaMalloc(&d->cellCount, sizeof(int) * (2 * LM_NCELL + 1)));
  VL_CUDA(cudaMalloc(&d->cellStart, sizeof(int) * (2 * LM_NCELL + 1)));
  VL_CUDA(cudaMalloc(&d->cellFill, sizeof(int) * (2 * LM_NCELL + 1)));
  VL_CUDA(cudaMalloc(&d->tileSum, sizeof(int) * (vl_div_up(2 * LM_NCELL, 1024) + 1)));
  VL_TRY(vl_scan_alloc(&d->scan, 2 * LM_NCELL));
  VL_CUDA(cudaMalloc(&d->dQ, sizeof(int) * 4));
  VL_CUDA(cudaMemset(d->dQ, 0, sizeof(int) * 4));
  VL_CUDA(cudaMalloc(&d->subReal, sizeof(LmSub))); VL_CUDA(cudaMemset(d->subReal, 0, sizeof(LmSub)));
  VL_CUDA(cudaMalloc(&d->subSpec, sizeof(LmSub))); VL_CUDA(cudaMemset(d->subSpec, 0, sizeof(LmSub)));
  VL_CUDA(cudaMalloc(&d->specOK, sizeof(int))); VL_CUDA(cudaMemset(d->specOK, 0, sizeof(int)));
  d->specQueued = false;
  d->specEnabled = getenv("VLOAM_NO_SPECULATION") == nullptr;
  VL_CUDA(cudaFuncSetAttribute(rf_seg_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, RF_SEG_CAP * 8));
  int perSm = 0;
  VL_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, lm_inline_build, 256, 0));
  d->inlineGrid = c->num_sms * max(1, min(perSm, 4));
  d->hMapUpperC = d->hMapUpperS = 0;
  // Map pools: bump allocation with doubling per cube; when a pool is half full it is doubled (copy) at sync point
  // S2, so a long drive degrades into a rare ~ms stall instead of VLOAM_E_CAPACITY.  VLOAM_POOL_POINTS: initial
  // size of the corner pool in points (the surf pool gets 3x), for tests of the growth path.
  const char* pp = getenv("VLOAM_POOL_POINTS");
  if (pp) {  // exact sizes (vl_reserve rounds small requests up)
    const size_t pc = (size_t)max(atoll(pp), 4096LL);
    VL_CUDA(cudaMalloc(&c->poolC.p, pc * sizeof(float4))); c->poolC.cap = pc;
    VL_CUDA(cudaMalloc(&c->poolS.p, 3 * pc * sizeof(float4))); c->poolS.cap = 3 * pc;
  } else {
    VL_TRY(vl_reserve(c, c->poolC, LM_POOL_C));
    VL_TRY(vl_reserve(c, c->poolS, LM_POOL_S));
  }
  VL_CUDA(cudaMemset(c->cubeC, 0, sizeof(MapCubeTable)));
  VL_CUDA(cudaMemset(c->cubeS, 0, sizeof(MapCubeTable)));
  LmScalars h;
  memset(&h, 0, sizeof h);
  h.cenW = 10; h.cenH = 10; h.cenD = 5;  // LM.h:75-78
  h.pose[3] = 1.0; h.q_wmap_wodom[3] = 1.0; h.q_wodom[3] = 1.0; h.q_hf[3] = 1.0;
  VL_CUDA(cudaMemcpy(c->lmm, &h, sizeof h, cudaMemcpyHostToDevice));
  *c->h_lmm = h;
  return VLOAM_OK;
}

void vl_lm_free(vloam_b200_ctx* c) {  // everything vl_lm_init and this file's reserves own (the pools are freed by capi.cu)
  LmDevice* d = lmdev(c);
  if (!d) return;
  void* dev[] = {d->work, d->cellCount, d->cellStart, d->cellFill, d->tileSum, d->dQ, d->subReal, d->subSpec, d->specOK,
                 d->cellOfPoint.p, d->sortedPts.p, d->newPts.p, d->newCube.p, d->unmatched.p};
  for (void* p : dev) if (p) cudaFree(p);
  vl_scan_free(&d->scan);
  delete d;
  c->gridPrm = nullptr;
}

extern bool vl_debug_capture(const vloam_b200_ctx* c);

// LM.cpp:492-500: VoxelGrid of this frame's less-sharp / less-flat clouds.  They depend on scan
// registration only, so the odometry stage enqueues them on the side stream right after its first sync
// point and they run underneath the odometry kernels; solveMapping waits on evStacks.
// The filters only conflict with the previous map update through rf_keys, the one kernel that reads the previous
// stacks (evKeys), not with the rest of the update or the speculative sub-map behind it.
// early = true (the counts were known when the sweep arrived: look-ahead scan registration): the helper thread
// issues the ~14 launches while the caller queues the odometry, so the stacks are ready well before the mapping
// front needs them; whoever needs evStacks joins the helper first (vl_lm_run does).
static int lm_issue_stacks(vloam_b200_ctx* c, const float4* corner, int nc, const float4* surf, int ns, bool toNext) {
  LmDevice* d = lmdev(c);
  struct Restore { cudaStream_t prev; ~Restore() { vl_tls_stream = prev; } } restore{vl_tls_stream};
  DBuf<float4>& dstS = toNext ? c->stackSN : c->stackS;
  DBuf<float4>& dstC = toNext ? c->stackCN : c->stackC;
  int* dq = d->dQ + 2 * (toNext ? c->stackSel ^ 1 : c->stackSel);
  // surf filter on stream2 (scratch lane 0), corner filter beside it on stream4 (scratch lane 1)
  vl_tls_stream = c->stream2;
  // the buffers about to be overwritten were last read by rf_keys of the update that used this pair (stream3)
  const int pairW = toNext ? c->stackSel ^ 1 : c->stackSel;
  cudaStreamWaitEvent(c->stream2, toNext ? c->evKeysSel[pairW] : c->evKeys, 0);
  int r = vl_reserve(c, dstS, (size_t)max(ns, 1));
  if (r == VLOAM_OK) r = vl_voxel_grid_device(c, surf, ns, nullptr, c->prm.plane_res, dstS.p, dq + 1, 0);
  if (c->timing) cudaEventRecord(c->evx[3], c->stream2);
  vl_tls_stream = c->stream4;
  cudaStreamWaitEvent(c->stream4, toNext ? c->evKeysSel[pairW] : c->evKeys, 0);
  if (r == VLOAM_OK) r = vl_reserve(c, dstC, (size_t)max(nc, 1));
  if (r == VLOAM_OK) r = vl_voxel_grid_device(c, corner, nc, nullptr, c->prm.line_res, dstC.p, dq, 1);
  if (c->timing) cudaEventRecord(c->evx[4], c->stream4);
  if (r != VLOAM_OK) return r;
  VL_CUDA(cudaEventRecord(c->evStacks, c->stream2));
  VL_CUDA(cudaEventRecord(c->evStacksC, c->stream4));
  if (!toNext) c->stacksReady = true;
  return VLOAM_OK;
}
int vl_lm_enqueue_stacks(vloam_b200_ctx* c, const float4* corner, int nc, const float4* surf, int ns, bool early) {
  VL_TRY(vl_lm_join(c));  // evKeys of the previous frame's update must have been recorded before it is waited on
  static const bool noWorker = getenv("VLOAM_NO_WORKER") != nullptr;
  if (early && !noWorker && !c->prof_name[0] && !vl_debug_capture(c))
    return lm_submit(c, [=]() -> int { return lm_issue_stacks(c, corner, nc, surf, ns, false); });
  return lm_issue_stacks(c, corner, nc, surf, ns, false);
}
// The next sweep's stacks, filtered while this sweep's mapping still runs (its scan registration finished early): they
// go to the spare buffers and the spare pair of counts, so nothing of this sweep is disturbed -- this sweep's mapping has
// been queued already (it holds its own wait on evStacks), its map update reads the current buffers.
int vl_lm_enqueue_stacks_next(vloam_b200_ctx* c, const float4* corner, int nc, const float4* surf, int ns) {
  c->stacksNextReady = false;
  static const bool noWorker = getenv("VLOAM_NO_WORKER") != nullptr;
  if (noWorker) VL_TRY(lm_issue_stacks(c, corner, nc, surf, ns, true));
  else VL_TRY(lm_submit(c, [=]() -> int { return lm_issue_stacks(c, corner, nc, surf, ns, true); }));
  c->stacksNextReady = true;
  return VLOAM_OK;
}
int vl_lm_adopt_stacks_next(vloam_b200_ctx* c) {
  VL_TRY(vl_lm_join(c));  // the helper may still be issuing them (it owns the spare DBufs until it is done)
  { DBuf<float4> t_ = c->stackC; c->stackC = c->stackCN; c->stackCN = t_; }
  { DBuf<float4> t_ = c->stackS; c->stackS = c->stackSN; c->stackSN = t_; }
  c->stackSel ^= 1;
  c->stacksReady = true;
  return VLOAM_OK;
}

// in-line sub-map build of this frame (one cooperative launch; returns at once on the device when *specOK)
static int lm_inline_launch(vloam_b200_ctx* c, LmDevice* d, long long totalBound) {
  LmInlineArgs a;
  a.sub = d->subReal; a.skip = d->specOK; a.tc = c->cubeC; a.ts = c->cubeS; a.poolC = c->poolC.p; a.poolS = c->poolS.p;
  a.outC = c->fromMapC.p; a.outS = c->fromMapS.p; a.cellCount = d->cellCount; a.cellFill = d->cellFill; a.cellStart = d->cellStart;
  a.tileSum = d->tileSum; a.cellOfPoint = d->cellOfPoint.p; a.sortedPts = d->sortedPts.p; a.nCells = 2 * LM_NCELL;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(d->inlineGrid); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = VL_STREAM(c);
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative; at[0].val.cooperative = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  const bool prof = c->prof_name[0] && vl_prof_match(c, "lm_inline_build") && c->prof_n < VL_PROF_MAX;
  if (prof) cudaEventRecord(c->prof_ev[c->prof_n][0], VL_STREAM(c));
  VL_CUDA(cudaLaunchKernelEx(&cfg, lm_inline_build, a));
  if (prof) { cudaEventRecord(c->prof_ev[c->prof_n][1], VL_STREAM(c)); c->prof_kname[c->prof_n] = "lm_inline_build";
              c->prof_kbytes[c->prof_n] = 100.0 * (double)totalBound; c->prof_kstream[c->prof_n] = VL_STREAM(c); c->prof_n++; }
  __atomic_fetch_add(&c->launches, 1LL, __ATOMIC_RELAXED);
  return VLOAM_OK;
}

int vl_lm_run(vloam_b200_ctx* c) {
  LmDevice* d = lmdev(c);
  VL_HOST_MARK(3);
  VL_TRY(vl_lm_join(c));  // the previous frame's map update has been issued (evMap recorded, host bounds updated)
  VL_HOST_MARK(4);
  const int skip = c->skip_frame ? 1 : 0;
  VL_CUDA(cudaStreamWaitEvent(c->stream, c->evMap, 0));  // the previous frame's map update (stream3) must be complete
  VL_CUDA(cudaStreamWaitEvent(c->stream, c->evAux, 0));  // ... including its outside appends (streamAux)
  VL_LAUNCH(lm_prepare, 1, 1024, 0, c->lmm, c->los, c->cubeC, c->cubeS, d->work, skip, c->lm_reset_pending ? 1 : 0, d->subReal, d->subSpec,
            d->specQueued ? 1 : 0, d->specOK);
  c->lm_reset_pending = false;
  if (skip) { VL_TRY(vl_lo_flush_deferred(c)); VL_CUDA(cudaGetLastError()); return VLOAM_OK; }
  const int gsGrid = c->num_sms * 8;
  VL_TRY(vl_reserve(c, c->fromMapC, (size_t)max(d->hMapUpperC, 1LL), false, (size_t)d->hMapUpperC / 2 + (1 << 20)));
  VL_TRY(vl_reserve(c, c->fromMapS, (size_t)max(d->hMapUpperS, 1LL), false, (size_t)d->hMapUpperS / 2 + (1 << 20)));
  if (!c->stacksReady) {  // laser_mapping called without this frame's laser_odometry having queued them
    VL_CUDA(cudaStreamSynchronize(c->stream));
    VL_TRY(vl_lm_enqueue_stacks(c, c->cornerLastPtr, c->nCornerLast, c->surfLastPtr, c->nSurfLast));
  }
  c->stacksReady = false;
  // The search / fit / solve kernels read the sizes (Mc, Ms, Qc, Qs) and the LM.cpp:514 decision on the
  // device, so they are queued without waiting for the host; sync point S2 sits after the solve, where
  // the pose has to be final anyway.  (Debug snapshots need host counts first and sync here.)
  const bool capture = vl_debug_capture(c);
  int Qc = 0, Qs = 0;
  const long long totalBound = d->hMapUpperC + d->hMapUpperS;
  const int nqBound = max(c->nCornerLast + c->nSurfLast, 1);  // a voxel filter never grows a cloud
  {
    VL_TRY(vl_reserve(c, d->cellOfPoint, (size_t)max(totalBound, 1LL), false, (size_t)totalBound / 2 + (1 << 20)));
    VL_TRY(vl_reserve(c, d->sortedPts, (size_t)max(totalBound, 1LL), false, (size_t)totalBound / 2 + (1 << 20)));
    VL_TRY(lm_inline_launch(c, d, totalBound));  // returns at once on the device when lm_prepare found the speculative build valid
    if (c->timing) VL_CUDA(cudaEventRecord(c->evx[0], c->stream));
    // only now are this frame's downsampled stacks needed (they were filtered on the side streams)
    VL_CUDA(cudaStreamWaitEvent(c->stream, c->evStacks, 0));
    VL_CUDA(cudaStreamWaitEvent(c->stream, c->evStacksC, 0));
    VL_LAUNCH(lm_set_counts, 1, 32, 0, c->lmm, d->work, d->dQ + 2 * c->stackSel, d->dQ + 2 * c->stackSel + 1);
    if (c->timing) VL_CUDA(cudaEventRecord(c->evx[1], c->stream));
    if (capture) {
      VL_CUDA(cudaMemcpyAsync(c->h_lmm, c->lmm, sizeof(LmScalars), cudaMemcpyDeviceToHost, c->stream));
      VL_CUDA(cudaStreamSynchronize(c->stream));
      Qc = c->h_lmm->Qc; Qs = c->h_lmm->Qs;
    }
    VL_TRY(vl_reserve(c, c->knnIdx, (size_t)nqBound * 5));
    VL_TRY(vl_reserve(c, c->knnD2, (size_t)nqBound * 5));
    VL_TRY(vl_reserve(c, c->knnOk, (size_t)nqBound));
    VL_TRY(vl_reserve(c, c->factors, (size_t)nqBound * 10));
    VL_TRY(vl_reserve(c, c->factorValid, (size_t)nqBound));
    for (int pass = 0; pass < 2; ++pass) {  // LM.cpp:526
      VL_BYTES(16.0 * 7000 * 6);  // query + 5 neighbours, typical Qc + Qs
      VL_LAUNCH(lm_knn, c->num_sms * 8, 256, 0, c->lmm, d->work, c->stackC.p, c->stackS.p, c->fromMapC.p,
                c->fromMapS.p, d->cellStart, d->sortedPts.p, c->knnIdx.p, c->knnD2.p, c->knnOk.p, c->factors.p, c->factorValid.p);
      VL_BYTES((16.0 * 6 + 24.0 + 84.0) * 7000);
      VL_LAUNCH(lm_fit, c->num_sms, 128, 0, c->lmm, c->stackC.p, c->stackS.p, c->fromMapC.p, c->fromMapS.p, c->knnIdx.p, c->knnD2.p,
                c->knnOk.p, c->factors.p, c->factorValid.p);
      if (capture) {
        for (int kind = 0; kind < 2; ++kind) {
          const int n = kind ? Qs : Qc, off = kind ? Qc : 0;
          VL_TRY(vl_reserve(c, c->dbgKnnIdx[pass][kind], (size_t)max(n, 1) * 5));
          VL_TRY(vl_reserve(c, c->dbgKnnD2[pass][kind], (size_t)max(n, 1) * 5));
          VL_TRY(vl_reserve(c, c->dbgKnnOk[pass][kind], (size_t)max(n, 1)));
          if (n == 0) continue;
          VL_CUDA(cudaMemcpyAsync(c->dbgKnnIdx[pass][kind].p, c->knnIdx.p + (size_t)off * 5, sizeof(int) * 5 * n, cudaMemcpyDeviceToDevice, c->stream));
          VL_CUDA(cudaMemcpyAsync(c->dbgKnnD2[pass][kind].p, c->knnD2.p + (size_t)off * 5, sizeof(float) * 5 * n, cudaMemcpyDeviceToDevice, c->stream));
          VL_CUDA(cudaMemcpyAsync(c->dbgKnnOk[pass][kind].p, c->knnOk.p + off, sizeof(int) * n, cudaMemcpyDeviceToDevice, c->stream));
        }
      }
      VL_TRY(vl_solve(c, nqBound, &d->work->nq, c->lmm->pose, capture ? &c->dbgLmCost[pass * 2] : nullptr, c->h_lmm->Qc + c->h_lmm->Qs));
      if (c->timing && pass == 0) VL_CUDA(cudaEventRecord(c->evx[2], c->stream));
    }
  }
  VL_LAUNCH(lm_transform_update, 1, 32, 0, c->lmm);  // LM.cpp:737 (runs even when the optimisation was skipped)
  // ---- map update (LM.cpp:741-808).  The pose is final here; the update runs on stream3 so that the caller can
  // read the pose, and the next frame's scan registration + odometry can start, while the map is brought up to date.
  VL_CUDA(cudaEventRecord(c->evPose, c->stream));
  // ---- sync point S2: sizes for the map update (and the pose, which is final now)
  VL_CUDA(cudaMemcpyAsync(c->h_los, c->los, sizeof(LoScalars), cudaMemcpyDeviceToHost, c->stream));
  VL_CUDA(cudaMemcpyAsync(c->h_lmm, c->lmm, sizeof(LmScalars), cudaMemcpyDeviceToHost, c->stream));
  VL_CUDA(cudaEventRecord(c->evS2, c->stream));
  VL_HOST_MARK(5);
  // While the device finishes this sweep's mapping, the next sweep's odometry solve is queued behind it (replays with a
  // registered look-ahead sweep): the device then runs on without the S2 -> caller -> next call round trip (~40 us).
  VL_TRY(vl_lo_flush_deferred(c));
  if (!capture) VL_TRY(vl_lo_lookahead(c));
  VL_CUDA(cudaEventSynchronize(c->evS2));
  c->s2Done = true;
  VL_HOST_MARK(6);
  const int Mc = c->h_lmm->Mc, Ms = c->h_lmm->Ms;
  Qc = c->h_lmm->Qc; Qs = c->h_lmm->Qs;
  if (c->h_lmm->overflow) { snprintf(c->err, sizeof c->err, "map pool exhausted"); return VLOAM_E_CAPACITY; }
  // Pool head room for this frame's update: a cube that outgrows its segment gets a new one of twice its new size,
  // so the update allocates at most 2 x (points in the map + inserts) + 256 per touched cube.  When that no longer
  // fits, the pool is doubled (copy) now: the previous update is complete (this stream waited on evMap), the next
  // one has not been issued, and the copy keeps every cube's offset valid.
  {
    const size_t needC = 2 * ((size_t)c->h_lmm->totalC + Qc) + 256 * (size_t)(VL_MAX_VALID + min(Qc, VL_CUBE_NUM));
    const size_t needS = 2 * ((size_t)c->h_lmm->totalS + Qs) + 256 * (size_t)(VL_MAX_VALID + min(Qs, VL_CUBE_NUM));
    if ((size_t)c->h_lmm->poolTopC + needC > c->poolC.cap) VL_TRY(vl_reserve(c, c->poolC, 2 * ((size_t)c->h_lmm->poolTopC + needC), true));
    if ((size_t)c->h_lmm->poolTopS + needS > c->poolS.cap) VL_TRY(vl_reserve(c, c->poolS, 2 * ((size_t)c->h_lmm->poolTopS + needS), true));
  }
  const int tailTotal = c->h_lmm->tailC + c->h_lmm->tailS;
  const int nq = Qc + Qs;
  c->lm_optimized = c->h_lmm->optimized;
  const long long totalC = c->h_lmm->totalC, totalS = c->h_lmm->totalS;
  const float4* const stackCp = c->stackC.p; const float4* const stackSp = c->stackS.p;  // (the next frame may swap the buffers while the helper issues this)
  const int stackSelNow = c->stackSel;
  auto update = [=]() -> int {
  vl_tls_stream = c->stream3;  // every launch helper below issues on the update's stream, whichever thread runs this
  struct Restore { ~Restore() { vl_tls_stream = nullptr; } } restore;
  VL_CUDA(cudaStreamWaitEvent(c->stream3, c->evPose, 0));
  const bool spec = d->specEnabled && !capture;
  auto zeroSpecGrid = [&]() -> int {  // (issued after the update's first kernel: that one heads the critical chain)
  if (spec) {  // the cell counters of the speculative grid are zeroed beside the update, not behind it
    VL_CUDA(cudaStreamWaitEvent(c->streamAux, c->evPose, 0));
    vl_tls_stream = c->streamAux;
    VL_BYTES(8.0 * (2 * LM_NCELL + 1));
    VL_LAUNCH(lm_grid_zero, gsGrid, 256, 0, d->cellCount, d->cellFill, 2 * LM_NCELL + 1, (const int*)nullptr);
    vl_tls_stream = c->stream3;
    VL_CUDA(cudaEventRecord(c->evAuxZero, c->streamAux));
  }
  return VLOAM_OK;
  };
  const int nKeys = tailTotal + nq;
  if (nKeys == 0) VL_TRY(zeroSpecGrid());
  if (nKeys > 0) {
    const size_t N = ((size_t)nKeys + 255) & ~(size_t)255;
    VL_TRY(vl_reserve(c, c->tailKeys, 4 * N, false, 4 * N));  // [0, N) unsorted keys | [N, 2N) bucketed + sorted | [2N, 4N) scratch for oversize segments
    unsigned long long* keysIn = c->tailKeys.p;
    unsigned long long* keysSorted = c->tailKeys.p + N;
    VL_TRY(vl_reserve(c, d->newPts, (size_t)max(nq, 1)));
    VL_TRY(vl_reserve(c, d->newCube, (size_t)max(nq, 1)));
    VL_TRY(vl_reserve(c, d->unmatched, (size_t)nKeys + 2, false, (size_t)nKeys + (1 << 16)));
    VL_TRY(vl_reserve(c, c->staging, (size_t)Mc + Ms + nKeys + 1, false, (size_t)(Mc + Ms) / 2 + (1 << 20)));
    VL_LAUNCH(rf_keys, vl_div_up(nKeys, 256), 256, 0, c->lmm, d->work, c->prm, c->cubeC, c->cubeS, c->poolC.p, c->poolS.p, stackCp, stackSp,
              d->newPts.p, d->newCube.p, keysIn, nKeys);
    VL_CUDA(cudaEventRecord(c->evKeys, c->stream3));  // nothing below reads the stacks any more: the next sweep's filters may overwrite them
    VL_CUDA(cudaEventRecord(c->evKeysSel[stackSelNow], c->stream3));
    VL_TRY(zeroSpecGrid());
    VL_LAUNCH(rf_seg_scatter, vl_div_up(nKeys, 256), 256, 0, keysIn, nKeys, d->work, keysSorted);
    VL_BYTES(16.0 * nKeys);
    static const int segCap = getenv("VLOAM_SEG_CAP") ? max(8, min(atoi(getenv("VLOAM_SEG_CAP")), RF_SEG_CAP)) : RF_SEG_CAP;  // tests force the global path
    VL_LAUNCH(rf_seg_sort, LM_NSEG, RF_SEG_THREADS, (size_t)RF_SEG_CAP * 8, keysSorted, d->work, c->tailKeys.p + 2 * N, segCap);
    VL_LAUNCH(rf_match, vl_div_up(nKeys, 256), 256, 0, keysSorted, c->lmm, d->work, c->prm, c->cubeC, c->cubeS, c->poolC.p, c->poolS.p, d->unmatched.p);
    VL_LAUNCH(rf_scan_layout, 1, 1024, 0, d->unmatched.p, c->lmm, d->work, c->cubeC, c->cubeS);
    VL_BYTES(32.0 * (Mc + Ms));  // read every prefix point once, write it once to staging
    VL_LAUNCH(rf_emit, vl_div_up(nKeys, 256) + gsGrid, 256, 0, vl_div_up(nKeys, 256), keysSorted, d->unmatched.p, c->lmm, d->work, c->prm, c->cubeC,
              c->cubeS, c->poolC.p, c->poolS.p, d->newPts.p, c->staging.p);
    VL_LAUNCH(rf_alloc, 1, 256, 0, c->lmm, d->work, c->cubeC, c->cubeS, (int)c->poolC.cap, (int)c->poolS.cap);
    VL_BYTES(32.0 * (Mc + Ms + nq));  // staging -> pool copy
    VL_LAUNCH(rf_commit, gsGrid, 256, 0, c->staging.p, c->lmm, d->work, c->prm, c->cubeC, c->cubeS, c->poolC.p, c->poolS.p);
    VL_LAUNCH(rf_finish, 1, 256, 0, c->lmm, d->work, c->cubeC, c->cubeS);
    if (nq > 0) {
      // Points that fell outside the 5x5x3 window go to cubes the speculative sub-map never reads: this single-CTA
      // kernel (~20 us) runs beside the sub-map build; the next solveMapping waits on both (evMap, evAux).
      VL_CUDA(cudaEventRecord(c->evUpd, c->stream3));
      VL_CUDA(cudaStreamWaitEvent(c->streamAux, c->evUpd, 0));
      vl_tls_stream = c->streamAux;
      VL_LAUNCH(rf_append_outside, 1, 1024, 0, c->lmm, d->newPts.p, d->newCube.p, c->cubeC, c->cubeS, c->poolC.p, c->poolS.p, (int)c->poolC.cap,
                (int)c->poolS.cap);
      vl_tls_stream = c->stream3;
    }
  }
  VL_CUDA(cudaEventRecord(c->evAux, c->streamAux));
  // the map after this frame's update holds at most the points it held before plus this frame's inserts
  d->hMapUpperC = totalC + Qc; d->hMapUpperS = totalS + Qs;
  // ---- speculative sub-map of the NEXT frame.  Gathering the 75 valid cubes and cell-sorting ~1M points is
  // ~80 us of dependent kernels that only depend on the pose through the window centre, and the window moves
  // once per 50 m.  So the work is done here, behind the map update on its side stream (underneath the next
  // sweep's scan registration and odometry), for the window this frame used; the next lm_prepare checks the
  // window and lets the in-line build run only when it moved.  Debug snapshots read this frame's sub-map
  // after the call returns, so capture mode keeps the in-line build.
  d->specQueued = false;
  if (spec) {
    const long long tb = d->hMapUpperC + d->hMapUpperS;
    VL_TRY(vl_reserve(c, c->fromMapC, (size_t)max(d->hMapUpperC, 1LL), false, (size_t)d->hMapUpperC / 2 + (1 << 20)));
    VL_TRY(vl_reserve(c, c->fromMapS, (size_t)max(d->hMapUpperS, 1LL), false, (size_t)d->hMapUpperS / 2 + (1 << 20)));
    VL_TRY(vl_reserve(c, d->cellOfPoint, (size_t)max(tb, 1LL), false, (size_t)tb / 2 + (1 << 20)));
    VL_TRY(vl_reserve(c, d->sortedPts, (size_t)max(tb, 1LL), false, (size_t)tb / 2 + (1 << 20)));
    VL_LAUNCH(lm_spec_prepare, 1, 256, 0, d->subReal, c->cubeC, c->cubeS, d->subSpec);
    VL_CUDA(cudaStreamWaitEvent(c->stream3, c->evAuxZero, 0));  // the cell counters were zeroed on streamAux
    VL_BYTES(56.0 * (double)tb);
    VL_LAUNCH(lm_gather_count, gsGrid, 256, 0, d->subSpec, c->cubeC, c->cubeS, c->poolC.p, c->poolS.p, c->fromMapC.p, c->fromMapS.p, d->cellCount,
              d->cellOfPoint.p);
    VL_TRY(vl_scan_exclusive(c, d->cellCount, 2 * LM_NCELL, &d->scan, d->cellStart, nullptr));
    VL_BYTES(44.0 * (double)tb);
    VL_LAUNCH(lm_grid_fill, gsGrid, 256, 0, d->subSpec, (const int*)nullptr, c->fromMapC.p, c->fromMapS.p, d->cellOfPoint.p, d->cellStart, d->cellFill,
              d->sortedPts.p);
    d->specQueued = true;
  }
  VL_CUDA(cudaEventRecord(c->evMap, c->stream3));
  return VLOAM_OK;
  };
  c->lm_frameCount++;
  // profiling and debug snapshots serialise everything; otherwise the helper thread issues the update while
  // the caller returns with its pose (whoever needs the map next joins it first: vl_lm_join)
  static const bool noWorker = getenv("VLOAM_NO_WORKER") != nullptr;
  if (noWorker || capture || c->prof_name[0]) { const int rmap = update(); if (rmap != VLOAM_OK) return rmap; }
  else VL_TRY(lm_submit(c, update));
  VL_HOST_MARK(7);
  VL_CUDA(cudaGetLastError());
  return VLOAM_OK;
}

// LM.cpp:901-905 (LaserMapping::publish): laserCloudFullRes through pointAssociateToMap
__global__ void __launch_bounds__(256) lm_register_full(const float4* __restrict__ in, int n, const LmScalars* __restrict__ s, int skip,
                                                        float4* __restrict__ out) {
  VL_PDL_WAIT();

  const double* q = skip ? s->q_hf : s->pose;
  const double* t = skip ? s->t_hf : s->pose + 4;
  const double pose[7] = {q[0], q[1], q[2], q[3], t[0], t[1], t[2]};
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < n; g += gridDim.x * blockDim.x) {
    const float4 p = in[g];
    double r[3];
    vl_qrot(pose, (double)p.x, (double)p.y, (double)p.z, r);  // LM.cpp:154-164
    out[g] = make_float4((float)(r[0] + pose[4]), (float)(r[1] + pose[5]), (float)(r[2] + pose[6]), p.w);
  }
}

int vl_lm_register_full(vloam_b200_ctx* c, const float4* d_in, int n, float4* d_out) {
  VL_BYTES(32.0 * n);
  VL_LAUNCH(lm_register_full, c->num_sms * 4, 256, 0, d_in, n, c->lmm, c->skip_frame ? 1 : 0, d_out);
  VL_CUDA(cudaGetLastError());
  return VLOAM_OK;
}

int vl_lm_rescan_sorted(vloam_b200_ctx* c) {
  lmdev(c)->specQueued = false;  // the state behind the speculative sub-map was edited from outside
  VL_LAUNCH(lm_scan_sorted, VL_CUBE_NUM, 256, 0, c->lmm, c->prm, c->cubeC, c->poolC.p, 0);
  VL_LAUNCH(lm_scan_sorted, VL_CUBE_NUM, 256, 0, c->lmm, c->prm, c->cubeS, c->poolS.p, 1);
  VL_CUDA(cudaStreamSynchronize(c->stream));
  return VLOAM_OK;
}

// blob = int32 counts[4851] followed by the points of all cubes in cube-index order
int vl_lm_export_map(vloam_b200_ctx* c, int which, void* out, long cap, long* bytes) {
  MapCubeTable h;
  VL_CUDA(cudaStreamSynchronize(c->stream));
  VL_CUDA(cudaMemcpy(&h, which ? c->cubeS : c->cubeC, sizeof h, cudaMemcpyDeviceToHost));
  long total = 0;
  for (int i = 0; i < VL_CUBE_NUM; ++i) total += h.count[i];
  *bytes = (long)VL_CUBE_NUM * 4 + total * 16;
  if (!out || cap < *bytes) return VLOAM_OK;
  memcpy(out, h.count, (size_t)VL_CUBE_NUM * 4);
  char* p = (char*)out + (size_t)VL_CUBE_NUM * 4;
  const float4* pool = which ? c->poolS.p : c->poolC.p;
  for (int i = 0; i < VL_CUBE_NUM; ++i) {
    if (h.count[i] == 0) continue;
    VL_CUDA(cudaMemcpy(p, pool + h.start[i], (size_t)h.count[i] * 16, cudaMemcpyDeviceToHost));
    p += (size_t)h.count[i] * 16;
  }
  return VLOAM_OK;
}

int vl_lm_import_map(vloam_b200_ctx* c, int which, const void* data, long bytes) {
  LmDevice* d = lmdev(c);
  d->specQueued = false;  // the map behind the speculative sub-map is replaced
  if (bytes < (long)VL_CUBE_NUM * 4) { snprintf(c->err, sizeof c->err, "map blob too short"); return VLOAM_E_INVALID; }
  const int* counts = (const int*)data;
  long long total = 0;
  for (int i = 0; i < VL_CUBE_NUM; ++i) total += counts[i];
  if (bytes != (long)VL_CUBE_NUM * 4 + total * 16) { snprintf(c->err, sizeof c->err, "map blob size mismatch"); return VLOAM_E_INVALID; }
  DBuf<float4>& pool = which ? c->poolS : c->poolC;
  // fresh layout: every cube gets 25 % head room (at least 256 points)
  MapCubeTable h;
  memset(&h, 0, sizeof h);
  long long top = 0;
  for (int i = 0; i < VL_CUBE_NUM; ++i) {
    if (counts[i] == 0) continue;
    h.start[i] = (int)top; h.count[i] = counts[i]; h.cap[i] = counts[i] + counts[i] / 4 + 256;
    top += h.cap[i];
  }
  VL_CUDA(cudaStreamSynchronize(c->stream));
  if ((size_t)top * 2 > pool.cap) VL_TRY(vl_reserve(c, pool, (size_t)top * 2));  // (contents are replaced below)
  const char* p = (const char*)data + (size_t)VL_CUBE_NUM * 4;
  for (int i = 0; i < VL_CUBE_NUM; ++i) {
    if (counts[i] == 0) continue;
    VL_CUDA(cudaMemcpy(pool.p + h.start[i], p, (size_t)counts[i] * 16, cudaMemcpyHostToDevice));
    p += (size_t)counts[i] * 16;
  }
  MapCubeTable* dt = which ? c->cubeS : c->cubeC;
  VL_CUDA(cudaMemcpy(dt, &h, sizeof h, cudaMemcpyHostToDevice));
  LmScalars hs;
  VL_CUDA(cudaMemcpy(&hs, c->lmm, sizeof hs, cudaMemcpyDeviceToHost));
  if (which) hs.poolTopS = (int)top; else hs.poolTopC = (int)top;
  VL_CUDA(cudaMemcpy(c->lmm, &hs, sizeof hs, cudaMemcpyHostToDevice));
  VL_LAUNCH(lm_scan_sorted, VL_CUBE_NUM, 256, 0, c->lmm, c->prm, dt, pool.p, which);
  VL_CUDA(cudaStreamSynchronize(c->stream));
  if (which) d->hMapUpperS = total; else d->hMapUpperC = total;
  return VLOAM_OK;
}

int vl_lm_preload(vloam_b200_ctx* c) {  // see vl_sr_set_attrs: load every kernel of this file when a context is created
  cudaFuncAttributes fa_;
  VL_CUDA(cudaFuncGetAttributes(&fa_, lm_prepare));
  VL_CUDA(cudaFuncGetAttributes(&fa_, lm_spec_prepare));
  VL_CUDA(cudaFuncGetAttributes(&fa_, lm_grid_zero));
  VL_CUDA(cudaFuncGetAttributes(&fa_, lm_gather_count));
  VL_CUDA(cudaFuncGetAttributes(&fa_, lm_scan_chained));
  VL_CUDA(cudaFuncGetAttributes(&fa_, lm_grid_fill));
  VL_CUDA(cudaFuncGetAttributes(&fa_, lm_inline_build));
  VL_CUDA(cudaFuncGetAttributes(&fa_, lm_knn));
  VL_CUDA(cudaFuncGetAttributes(&fa_, lm_fit));
  VL_CUDA(cudaFuncGetAttributes(&fa_, lm_fit_sets));
  VL_CUDA(cudaFuncGetAttributes(&fa_, lm_transform_update));
  VL_CUDA(cudaFuncGetAttributes(&fa_, rf_keys));
  VL_CUDA(cudaFuncGetAttributes(&fa_, rf_seg_scatter));
  VL_CUDA(cudaFuncGetAttributes(&fa_, rf_seg_sort));
  VL_CUDA(cudaFuncGetAttributes(&fa_, rf_match));
  VL_CUDA(cudaFuncGetAttributes(&fa_, rf_scan_layout));
  VL_CUDA(cudaFuncGetAttributes(&fa_, rf_emit));
  VL_CUDA(cudaFuncGetAttributes(&fa_, rf_alloc));
  VL_CUDA(cudaFuncGetAttributes(&fa_, rf_commit));
  VL_CUDA(cudaFuncGetAttributes(&fa_, rf_finish));
  VL_CUDA(cudaFuncGetAttributes(&fa_, rf_append_outside));
  VL_CUDA(cudaFuncGetAttributes(&fa_, lm_scan_sorted));
  VL_CUDA(cudaFuncGetAttributes(&fa_, lm_set_counts));
  VL_CUDA(cudaFuncGetAttributes(&fa_, lm_register_full));
  return VLOAM_OK;
}
