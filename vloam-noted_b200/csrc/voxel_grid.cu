// voxel_grid.cu -- device-wide sort and pcl::VoxelGrid<PointXYZI> restated as CUDA.
//
// pcl::VoxelGrid (used at scan_registration.cpp:497-501, laser_mapping.cpp:492-500,
// 795-808) is: bounding box -> linear voxel index per point -> sort (index, point)
// pairs -> one f32 centroid per run, output in ascending voxel index.  The canonical
// refinement of PCL's unstable sort is ascending point index
// inside a voxel, so keys are (voxel idx << 32 | point index) and any correct sort
// of those unique 64-bit keys reproduces the oracle bit for bit.
//
// Sort: bitonic network, tiles of 4096 keys sorted / merged in shared memory, the
// strides above a tile as one global compare-exchange kernel each.  The arrays here
// are <= a few hundred thousand keys (L2-resident), where this beats a radix sort's
// fixed pass count and needs no scratch buffer.
#include <limits.h>
#include <math_constants.h>
#include <cooperative_groups.h>
#include "common.cuh"
#include "bitonic.cuh"

namespace cg = cooperative_groups;

#define BT_TILE 4096
#define BT_THREADS 512
#define VG_BLOCK 256
#define VG_TILE 256   // sorted keys per CTA in vg_head_count / vg_centroid (1024 left the 50k-point stack filter with 49 CTAs on 148 SMs)

__global__ void __launch_bounds__(BT_THREADS) bt_tile_sort(unsigned long long* __restrict__ keys, int tile) {
  VL_PDL_WAIT();

  __shared__ unsigned long long s[BT_TILE];
  const size_t base = (size_t)blockIdx.x * tile;
  for (int t = threadIdx.x; t < tile; t += BT_THREADS) s[t] = keys[base + t];
  __syncthreads();
  const int half = tile >> 1;
  for (int k = 2; k <= tile; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int u = threadIdx.x; u < half; u += BT_THREADS) {
        const int i = ((u & ~(j - 1)) << 1) | (u & (j - 1));
        const int l = i | j;
        const unsigned long long a = s[i], b = s[l];
        const bool up = ((base + i) & (size_t)k) == 0;
        if ((a > b) == up) { s[i] = b; s[l] = a; }
      }
      __syncthreads();
    }
  for (int t = threadIdx.x; t < tile; t += BT_THREADS) keys[base + t] = s[t];
}

__global__ void __launch_bounds__(256) bt_global_step(unsigned long long* __restrict__ keys, int halfN, int j, int k) {
  VL_PDL_WAIT();

  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= halfN) return;
  const int i = ((u & ~(j - 1)) << 1) | (u & (j - 1));
  const int l = i | j;
  const unsigned long long a = keys[i], b = keys[l];
  const bool up = (i & k) == 0;
  if ((a > b) == up) { keys[i] = b; keys[l] = a; }
}

__global__ void __launch_bounds__(BT_THREADS) bt_tile_merge(unsigned long long* __restrict__ keys, int tile, int k) {
  VL_PDL_WAIT();

  __shared__ unsigned long long s[BT_TILE];
  const size_t base = (size_t)blockIdx.x * tile;
  for (int t = threadIdx.x; t < tile; t += BT_THREADS) s[t] = keys[base + t];
  __syncthreads();
  const int half = tile >> 1;
  const bool up = (base & (size_t)k) == 0;  // k > tile: the whole tile shares one direction
  for (int j = tile >> 1; j > 0; j >>= 1) {
    for (int u = threadIdx.x; u < half; u += BT_THREADS) {
      const int i = ((u & ~(j - 1)) << 1) | (u & (j - 1));
      const int l = i | j;
      const unsigned long long a = s[i], b = s[l];
      if ((a > b) == up) { s[i] = b; s[l] = a; }
    }
    __syncthreads();
  }
  for (int t = threadIdx.x; t < tile; t += BT_THREADS) keys[base + t] = s[t];
}


// Strides j = jtop, jtop/2, ..., 1 of one bitonic merge level inside a shared-memory tile, two strides
// per barrier: a thread owns the 4 elements {i, i+h, i+j, i+j+h} (h = j/2), which are closed under both
// compare-exchange steps, so the pair of stages needs one __syncthreads instead of two.
template <int THREADS>
__device__ __forceinline__ void bt_smem_strides(unsigned long long* s, int tile, int jtop, size_t base, int k, bool uniformUp, bool upAll) {
  int j = jtop;
  while (j >= 2) {
    const int h = j >> 1;
    const int lh = __ffs(h) - 1;
    for (int u = threadIdx.x; u < (tile >> 2); u += THREADS) {
      const int i0 = ((u >> lh) << (lh + 2)) | (u & (h - 1));
      const bool up = uniformUp ? upAll : (((base + i0) & (size_t)k) == 0);
      unsigned long long a = s[i0], b = s[i0 + h], c = s[i0 + j], d = s[i0 + j + h];
      unsigned long long t;
      if ((a > c) == up) { t = a; a = c; c = t; }
      if ((b > d) == up) { t = b; b = d; d = t; }
      if ((a > b) == up) { t = a; a = b; b = t; }
      if ((c > d) == up) { t = c; c = d; d = t; }
      s[i0] = a; s[i0 + h] = b; s[i0 + j] = c; s[i0 + j + h] = d;
    }
    __syncthreads();
    j >>= 2;
  }
  if (j == 1) {
    for (int u = threadIdx.x; u < (tile >> 1); u += THREADS) {
      const int i = u << 1;
      const bool up = uniformUp ? upAll : (((base + i) & (size_t)k) == 0);
      const unsigned long long a = s[i], b = s[i + 1];
      if ((a > b) == up) { s[i] = b; s[i + 1] = a; }
    }
    __syncthreads();
  }
}

// ---- single-launch sort: one thread-block cluster, tiles in distributed shared memory -----------
// Up to 16 CTAs x 8192 keys.  Every CTA bitonic-sorts its tile in shared memory; the strides that
// cross tiles read the partner tile through DSMEM (cluster.map_shared_rank) and write the result
// into the second half of a double buffer, so a cross-tile stage costs one cluster barrier instead
// of a kernel launch plus a global-memory round trip.
#define CS_THREADS 1024

__global__ void __launch_bounds__(CS_THREADS) bt_cluster_sort(unsigned long long* __restrict__ keys, int tile) {
  VL_PDL_WAIT();

  extern __shared__ unsigned long long cs[];  // [2][tile]
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned rank = cluster.block_rank();
  const int n = tile * (int)cluster.num_blocks();
  const size_t base = (size_t)rank * tile;
  unsigned long long* cur = cs;
  unsigned long long* alt = cs + tile;
  for (int t = threadIdx.x; t < tile; t += CS_THREADS) cur[t] = keys[base + t];
  __syncthreads();
  if (tile >= 8) {
    bt_smem_init4<CS_THREADS>(cur, tile);
    for (int k = 8; k <= tile; k <<= 1) bt_smem_level<CS_THREADS>(cur, tile, k >> 1, base, k);
  } else {  // 2 or 4 keys in all (a single CTA)
    for (int k = 2; k <= tile; k <<= 1) bt_smem_strides<CS_THREADS>(cur, tile, k >> 1, base, k, false, false);
  }
  for (int k = tile << 1; k <= n; k <<= 1) {
    for (int j = k >> 1; j >= tile; j >>= 1) {
      cluster.sync();  // partner tiles are complete in `cur`
      const unsigned prank = rank ^ (unsigned)(j / tile);
      const unsigned long long* rem = cluster.map_shared_rank(cur, prank);
      const bool low = (base & (size_t)j) == 0;
      const bool up = (base & (size_t)k) == 0;
      const bool takeMin = (low == up);
      for (int t = threadIdx.x; t < tile; t += CS_THREADS) {
        const unsigned long long a = cur[t], b = rem[t];
        alt[t] = takeMin ? (a < b ? a : b) : (a > b ? a : b);
      }
      unsigned long long* tmp = cur; cur = alt; alt = tmp;
    }
    // every CTA flipped buffers the same number of times, so `cur` means the same half everywhere;
    // the barrier below also keeps a fast CTA from overwriting `alt` while a partner still reads it
    cluster.sync();
    bt_smem_level<CS_THREADS>(cur, tile, tile >> 1, base, k);
  }
  for (int t = threadIdx.x; t < tile; t += CS_THREADS) keys[base + t] = cur[t];
  cluster.sync();  // no CTA exits while a partner may still read its shared memory
}

// function attributes are per device: set when a context is created on it (vloam_b200_create)
int vl_sort_set_attrs(vloam_b200_ctx* c) {
  VL_CUDA(cudaFuncSetAttribute(bt_cluster_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 8192 * 8));
  VL_CUDA(cudaFuncSetAttribute(bt_cluster_sort, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  // lazy module loading would otherwise charge each kernel's first launch (~1 ms apiece) to the first sweeps
  cudaFuncAttributes fa_;
  VL_CUDA(cudaFuncGetAttributes(&fa_, bt_tile_sort));
  VL_CUDA(cudaFuncGetAttributes(&fa_, bt_global_step));
  VL_CUDA(cudaFuncGetAttributes(&fa_, bt_tile_merge));
  VL_CUDA(cudaFuncGetAttributes(&fa_, bt_cluster_sort));
  return VLOAM_OK;
}

static int vl_sort_cluster(vloam_b200_ctx* c, unsigned long long* d_keys, int n_pow2) {
  int tile = n_pow2 <= 4096 ? n_pow2 : 4096;
  int nct = n_pow2 / tile;
  if (nct > 16) { tile = 8192; nct = n_pow2 / tile; }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(nct); cfg.blockDim = dim3(CS_THREADS); cfg.dynamicSmemBytes = (size_t)2 * tile * 8; cfg.stream = VL_STREAM(c);
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = nct; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = 2;
  const bool prof = c->prof_name[0] && vl_prof_match(c, "bt_cluster_sort") && c->prof_n < VL_PROF_MAX;
  if (prof) cudaEventRecord(c->prof_ev[c->prof_n][0], VL_STREAM(c));
  VL_CUDA(cudaLaunchKernelEx(&cfg, bt_cluster_sort, d_keys, tile));
  if (prof) { cudaEventRecord(c->prof_ev[c->prof_n][1], VL_STREAM(c)); c->prof_kname[c->prof_n] = "bt_cluster_sort"; c->prof_kbytes[c->prof_n] = 16.0 * n_pow2; c->prof_kstream[c->prof_n] = VL_STREAM(c);
              c->prof_n++; c->prof_bytes += 16.0 * n_pow2; }
  __atomic_fetch_add(&c->launches, 1LL, __ATOMIC_RELAXED);
  return VLOAM_OK;
}

// Ascending sort of n_pow2 unique 64-bit keys in place (n_pow2 a power of two >= 2).
int vl_sort_u64(vloam_b200_ctx* c, unsigned long long* d_keys, int n_pow2) {
  if (n_pow2 < 2) return VLOAM_OK;
  if (n_pow2 <= 16 * 8192) return vl_sort_cluster(c, d_keys, n_pow2);
  const int tile = n_pow2 < BT_TILE ? n_pow2 : BT_TILE;
  VL_LAUNCH(bt_tile_sort, n_pow2 / tile, BT_THREADS, 0, d_keys, tile);
  for (int k = tile << 1; k <= n_pow2; k <<= 1) {
    for (int j = k >> 1; j >= tile; j >>= 1)
      VL_LAUNCH(bt_global_step, vl_div_up(n_pow2 / 2, 256), 256, 0, d_keys, n_pow2 / 2, j, k);
    VL_LAUNCH(bt_tile_merge, n_pow2 / tile, BT_THREADS, 0, d_keys, tile, k);
  }
  VL_CUDA(cudaGetLastError());
  return VLOAM_OK;
}

// ---- voxel grid over one large cloud ------------------------------------------------
struct VgBox { int minb[3], mul1, mul2, guard, n; float inv; };

__device__ __forceinline__ unsigned vg_idx(const float4 p, const VgBox& b) {
  const int i0 = (int)__fsub_rn(floorf(__fmul_rn(p.x, b.inv)), (float)b.minb[0]);
  const int i1 = (int)__fsub_rn(floorf(__fmul_rn(p.y, b.inv)), (float)b.minb[1]);
  const int i2 = (int)__fsub_rn(floorf(__fmul_rn(p.z, b.inv)), (float)b.minb[2]);
  return (unsigned)(i0 + i1 * b.mul1 + i2 * b.mul2);
}

// getMinMax3D: per-block partial bounds -> partial[block*6 + {minx,miny,minz,maxx,maxy,maxz}]
__global__ void __launch_bounds__(VG_BLOCK) vg_bbox(const float4* __restrict__ in, int nBound, const int* __restrict__ dN,
                                                    float* __restrict__ partial) {
  VL_PDL_WAIT();

  const int n = dN ? min(*dN, nBound) : nBound;
  float mn[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F}, mx[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 p = in[i];
    mn[0] = fminf(mn[0], p.x); mx[0] = fmaxf(mx[0], p.x);
    mn[1] = fminf(mn[1], p.y); mx[1] = fmaxf(mx[1], p.y);
    mn[2] = fminf(mn[2], p.z); mx[2] = fmaxf(mx[2], p.z);
  }
  __shared__ float red[6][VG_BLOCK / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float lo = mn[a], hi = mx[a];
    for (int d = 16; d > 0; d >>= 1) { lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, d)); hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, d)); }
    if (lane == 0) { red[a][warp] = lo; red[3 + a][warp] = hi; }
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    float v = red[threadIdx.x][0];
    for (int w = 1; w < VG_BLOCK / 32; ++w) v = threadIdx.x < 3 ? fminf(v, red[threadIdx.x][w]) : fmaxf(v, red[threadIdx.x][w]);
    partial[blockIdx.x * 6 + threadIdx.x] = v;
  }
}

__global__ void vg_box(const float* __restrict__ partial, int nPartial, int nBound, const int* __restrict__ dN, float leaf,
                       VgBox* __restrict__ box, int* __restrict__ dCount) {
  VL_PDL_WAIT();

  // (launched with one warp: the <= 256 partial boxes are reduced by all 32 lanes -- one lane walking 1536 values took ~20 us)
  const int n = dN ? min(*dN, nBound) : nBound;
  float mn[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F}, mx[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
  for (int b = threadIdx.x; b < nPartial; b += 32)
    for (int a = 0; a < 3; ++a) { mn[a] = fminf(mn[a], partial[b * 6 + a]); mx[a] = fmaxf(mx[a], partial[b * 6 + 3 + a]); }
#pragma unroll
  for (int a = 0; a < 3; ++a)
    for (int d = 16; d > 0; d >>= 1) { mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], d)); mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], d)); }
  if (threadIdx.x != 0) return;
  VgBox bx;
  bx.n = n;
  const float inv = __fdiv_rn(1.0f, leaf);
  bx.inv = inv;
  bx.guard = 0; bx.mul1 = bx.mul2 = 0; bx.minb[0] = bx.minb[1] = bx.minb[2] = 0;
  if (n > 0) {
    const long long dx = (long long)__fmul_rn(__fsub_rn(mx[0], mn[0]), inv) + 1;
    const long long dy = (long long)__fmul_rn(__fsub_rn(mx[1], mn[1]), inv) + 1;
    const long long dz = (long long)__fmul_rn(__fsub_rn(mx[2], mn[2]), inv) + 1;
    bx.guard = (dx * dy * dz > (long long)INT_MAX) ? 1 : 0;
    int maxb[3];
    for (int a = 0; a < 3; ++a) { bx.minb[a] = (int)floorf(__fmul_rn(mn[a], inv)); maxb[a] = (int)floorf(__fmul_rn(mx[a], inv)); }
    const int d0 = maxb[0] - bx.minb[0] + 1, d1 = maxb[1] - bx.minb[1] + 1;
    bx.mul1 = d0; bx.mul2 = d0 * d1;
  }
  *box = bx;
  if (n == 0) *dCount = 0;
}

__global__ void __launch_bounds__(VG_BLOCK) vg_keys(const float4* __restrict__ in, const VgBox* __restrict__ box,
                                                    unsigned long long* __restrict__ keys, int P) {
  VL_PDL_WAIT();

  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const VgBox b = *box;
  keys[i] = (i < b.n && !b.guard) ? (((unsigned long long)vg_idx(in[i], b) << 32) | (unsigned)i) : ~0ull;
}

// Tile of VG_TILE sorted keys per block: count run heads.
__global__ void __launch_bounds__(VG_BLOCK) vg_head_count(const unsigned long long* __restrict__ keys, const VgBox* __restrict__ box,
                                                          int* __restrict__ blockCnt) {
  VL_PDL_WAIT();

  const int n = box->guard ? 0 : box->n;
  int cnt = 0;
  const int base = blockIdx.x * VG_TILE;
  for (int q = 0; q < VG_TILE / VG_BLOCK; ++q) {
    const int t = base + q * VG_BLOCK + threadIdx.x;
    if (t < n && (t == 0 || (unsigned)(keys[t] >> 32) != (unsigned)(keys[t - 1] >> 32))) ++cnt;
  }
  __shared__ int ws[VG_BLOCK / 32];
  for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) { int s = 0; for (int w = 0; w < VG_BLOCK / 32; ++w) s += ws[w]; blockCnt[blockIdx.x] = s; }
}

__global__ void __launch_bounds__(1024) vg_block_scan(int* __restrict__ blockCnt, int nBlocks, const VgBox* __restrict__ box,
                                                      int* __restrict__ dCount) {
  VL_PDL_WAIT();

  __shared__ int ws[32];
  int carry = 0;  // the same value in every thread
  for (int base = 0; base < nBlocks; base += 1024) {
    const int b = base + threadIdx.x;
    const int own = b < nBlocks ? blockCnt[b] : 0;
    int tot = 0;
    const int ex = vl_block_excl_scan<1024>(own, ws, &tot);
    if (b < nBlocks) blockCnt[b] = carry + ex;
    carry += tot;
  }
  if (threadIdx.x == 0) *dCount = box->guard ? box->n : carry;
}

#define VG_HALO 64
__global__ void __launch_bounds__(VG_BLOCK) vg_centroid(const float4* __restrict__ in, const unsigned long long* __restrict__ keys,
                                                        const VgBox* __restrict__ box, const int* __restrict__ blockOff,
                                                        float4* __restrict__ out) {
  VL_PDL_WAIT();

  const VgBox b = *box;
  if (b.guard) {  // leaf too small for the extent: output = input
    for (int i = blockIdx.x * VG_TILE + threadIdx.x; i < min(b.n, (int)(blockIdx.x + 1) * VG_TILE); i += VG_BLOCK) out[i] = in[i];
    return;
  }
  const int n = b.n;
  // stage the tile's voxel ids and points in shared memory: the serial f32 folds below then run at
  // shared-memory latency instead of chasing global loads (a run may continue past the tile; the tail
  // is read from global, which is rare)
  // VG_HALO entries beyond the tile are staged too: the last run of nearly every tile continues into the next
  // one, and finishing it from global memory is a chain of dependent key -> point loads (~1 us per 2 points).
  __shared__ unsigned svox[VG_TILE + VG_HALO + 1];
  __shared__ float4 spt[VG_TILE + VG_HALO];
  __shared__ int ws[VG_BLOCK / 32];
  const int base = blockIdx.x * VG_TILE;
  for (int t = threadIdx.x; t < VG_TILE + VG_HALO; t += VG_BLOCK) {
    const int g = base + t;
    if (g < n) { const unsigned long long k = keys[g]; svox[t] = (unsigned)(k >> 32); spt[t] = in[(int)(unsigned)(k & 0xffffffffull)]; }
    else svox[t] = 0xffffffffu;
  }
  __syncthreads();
  int running = blockOff[blockIdx.x];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int q = 0; q < VG_TILE / VG_BLOCK; ++q) {
    const int lt = q * VG_BLOCK + threadIdx.x;
    const int t = base + lt;
    const unsigned vox = svox[lt];
    const bool head = t < n && (lt == 0 ? (t == 0 || (unsigned)(keys[t - 1] >> 32) != vox) : svox[lt - 1] != vox);
    const unsigned bal = __ballot_sync(0xffffffffu, head);
    if (lane == 0) ws[warp] = __popc(bal);
    __syncthreads();
    int before = 0, all = 0;
    for (int w = 0; w < VG_BLOCK / 32; ++w) { const int v = ws[w]; if (w < warp) before += v; all += v; }
    if (head) {
      float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f; int nrun = 0;
      int r = lt;
      for (; r < VG_TILE + VG_HALO && svox[r] == vox; ++r) {
        const float4 p = spt[r];
        sx = __fadd_rn(sx, p.x); sy = __fadd_rn(sy, p.y); sz = __fadd_rn(sz, p.z); si = __fadd_rn(si, p.w);
        ++nrun;
      }
      if (r == VG_TILE + VG_HALO)
        for (int g = base + VG_TILE + VG_HALO; g < n && (unsigned)(keys[g] >> 32) == vox; ++g) {
          const float4 p = in[(int)(unsigned)(keys[g] & 0xffffffffull)];
          sx = __fadd_rn(sx, p.x); sy = __fadd_rn(sy, p.y); sz = __fadd_rn(sz, p.z); si = __fadd_rn(si, p.w);
          ++nrun;
        }
      const float fn = (float)nrun;
      out[running + before + __popc(bal & ((1u << lane) - 1u))] =
          make_float4(__fdiv_rn(sx, fn), __fdiv_rn(sy, fn), __fdiv_rn(sz, fn), __fdiv_rn(si, fn));
    }
    running += all;
    __syncthreads();
  }
}

int vl_voxel_grid_device(vloam_b200_ctx* c, const float4* d_in, int n, const int* d_n, float leaf, float4* d_out, int* d_count, int lane) {
  if (n <= 0) { VL_CUDA(cudaMemsetAsync(d_count, 0, sizeof(int), VL_STREAM(c))); return VLOAM_OK; }
  int P = 2; while (P < n) P <<= 1;
  const int nb = min(vl_div_up(n, VG_BLOCK), 256);
  const int nTiles = vl_div_up(n, VG_TILE);
  DBuf<unsigned long long>& vKeys = lane ? c->vKeys2 : c->vKeys;
  DBuf<int>& vScan = lane ? c->vScan2 : c->vScan;
  VL_TRY(vl_reserve(c, vKeys, (size_t)P));
  VL_TRY(vl_reserve(c, vScan, (size_t)nTiles + 256 * 6 + 64));
  float* partial = reinterpret_cast<float*>(vScan.p + nTiles);
  VgBox* box = reinterpret_cast<VgBox*>(c->vScalars + (lane ? 96 : 0));
  VL_LAUNCH(vg_bbox, nb, VG_BLOCK, 0, d_in, n, d_n, partial);
  VL_LAUNCH(vg_box, 1, 32, 0, partial, nb, n, d_n, leaf, box, d_count);
  VL_LAUNCH(vg_keys, vl_div_up(P, VG_BLOCK), VG_BLOCK, 0, d_in, box, vKeys.p, P);
  VL_TRY(vl_sort_u64(c, vKeys.p, P));
  VL_LAUNCH(vg_head_count, nTiles, VG_BLOCK, 0, vKeys.p, box, vScan.p);
  VL_LAUNCH(vg_block_scan, 1, 1024, 0, vScan.p, nTiles, box, d_count);
  VL_BYTES(40.0 * n);  // key + gathered point in, centroid out
  VL_LAUNCH(vg_centroid, nTiles, VG_BLOCK, 0, d_in, vKeys.p, box, vScan.p, d_out);
  VL_CUDA(cudaGetLastError());
  return VLOAM_OK;
}

int vl_vg_preload(vloam_b200_ctx* c) {  // see vl_sr_set_attrs: load the voxel-filter kernels when a context is created
  cudaFuncAttributes fa_;
  VL_CUDA(cudaFuncGetAttributes(&fa_, vg_bbox)); VL_CUDA(cudaFuncGetAttributes(&fa_, vg_box)); VL_CUDA(cudaFuncGetAttributes(&fa_, vg_keys));
  VL_CUDA(cudaFuncGetAttributes(&fa_, vg_head_count)); VL_CUDA(cudaFuncGetAttributes(&fa_, vg_block_scan)); VL_CUDA(cudaFuncGetAttributes(&fa_, vg_centroid));
  return VLOAM_OK;
}
