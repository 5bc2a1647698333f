// bitonic.cuh -- block-level bitonic sort of 64-bit keys (shared or global memory), two strides per barrier.
#pragma once

// Sorts `nArr` independent arrays of P keys each (P a power of two >= 4, array a at s + a*P) ascending.
// A thread owns the 4 keys {i, i+h, i+j, i+j+h} (h = j/2): they are closed under the compare-exchange
// steps of strides j and h, so a pair of network stages costs one __syncthreads.
template <int THREADS>
__device__ __forceinline__ void bt_sort_batched(unsigned long long* s, int P, int nArr) {
  const int quarter = P >> 2, halfP = P >> 1;
  for (int k = 2; k <= P; k <<= 1) {
    int j = k >> 1;
    while (j >= 2) {
      const int h = j >> 1;
      const int lh = __ffs(h) - 1;
      for (int t = threadIdx.x; t < nArr * quarter; t += THREADS) {
        const int arr = t / quarter, u = t - arr * quarter;
        const int i0 = ((u >> lh) << (lh + 2)) | (u & (h - 1));
        const bool up = (i0 & k) == 0;
        unsigned long long* p = s + (size_t)arr * P + i0;
        unsigned long long a = p[0], b = p[h], c = p[j], d = p[j + h], x;
        if ((a > c) == up) { x = a; a = c; c = x; }
        if ((b > d) == up) { x = b; b = d; d = x; }
        if ((a > b) == up) { x = a; a = b; b = x; }
        if ((c > d) == up) { x = c; c = d; d = x; }
        p[0] = a; p[h] = b; p[j] = c; p[j + h] = d;
      }
      __syncthreads();
      j >>= 2;
    }
    if (j == 1) {
      for (int t = threadIdx.x; t < nArr * halfP; t += THREADS) {
        const int arr = t / halfP, u = t - arr * halfP;
        const int i = u << 1;
        const bool up = (i & k) == 0;
        unsigned long long* p = s + (size_t)arr * P + i;
        const unsigned long long a = p[0], b = p[1];
        if ((a > b) == up) { p[0] = b; p[1] = a; }
      }
      __syncthreads();
    }
  }
}


// ---- single array in SHARED memory, tuned rounds ---------------------------------------------------------
// Call these with a pointer the compiler can see is shared memory (derived from the __shared__ array in the
// calling kernel, not selected at run time between shared and global): the accesses then compile to LDS / STS.
// Strides 2 and 1 of every merge level -- and the whole k = 2, k = 4 levels -- are done on four CONSECUTIVE
// keys held in registers (two 128-bit accesses per thread, conflict-free), strides >= 8 in pairs per barrier
// ({i, i+h, i+j, i+j+h} is closed under strides j and h = j/2), a left-over stride 4 on its own.
// Direction of element i at level k: ascending iff ((gbase + i) & k) == 0 (gbase: global index of s[0]).
__device__ __forceinline__ void bt_cswap(unsigned long long& a, unsigned long long& b, bool up) {
  if ((a > b) == up) { const unsigned long long t = a; a = b; b = t; }
}

template <int THREADS>
__device__ __forceinline__ void bt_smem_init4(unsigned long long* s, int n) {  // levels k = 2 and k = 4
  for (int u = threadIdx.x; u < (n >> 2); u += THREADS) {
    ulonglong2* p = reinterpret_cast<ulonglong2*>(s + (u << 2));
    ulonglong2 v0 = p[0], v1 = p[1];
    bt_cswap(v0.x, v0.y, true); bt_cswap(v1.x, v1.y, false);           // k = 2: (i & 2) == 0 ascending
    const bool up = ((u << 2) & 4) == 0;                                // k = 4 (gbase is a multiple of 4)
    bt_cswap(v0.x, v1.x, up); bt_cswap(v0.y, v1.y, up);
    bt_cswap(v0.x, v0.y, up); bt_cswap(v1.x, v1.y, up);
    p[0] = v0; p[1] = v1;
  }
  __syncthreads();
}

template <int THREADS>
__device__ __forceinline__ void bt_smem_level(unsigned long long* s, int n, int jtop, size_t gbase, int k) {  // k >= 8, strides jtop .. 1
  int j = jtop;
  while (j >= 8) {
    const int h = j >> 1, lh = __ffs(h) - 1;
    for (int u = threadIdx.x; u < (n >> 2); u += THREADS) {
      const int i0 = ((u >> lh) << (lh + 2)) | (u & (h - 1));
      const bool up = ((gbase + i0) & (size_t)k) == 0;
      unsigned long long a = s[i0], b = s[i0 + h], c = s[i0 + j], d = s[i0 + j + h];
      bt_cswap(a, c, up); bt_cswap(b, d, up); bt_cswap(a, b, up); bt_cswap(c, d, up);
      s[i0] = a; s[i0 + h] = b; s[i0 + j] = c; s[i0 + j + h] = d;
    }
    __syncthreads();
    j >>= 2;
  }
  if (j == 4) {
    for (int u = threadIdx.x; u < (n >> 1); u += THREADS) {
      const int i = ((u & ~3) << 1) | (u & 3);
      const bool up = ((gbase + i) & (size_t)k) == 0;
      unsigned long long a = s[i], b = s[i + 4];
      bt_cswap(a, b, up);
      s[i] = a; s[i + 4] = b;
    }
    __syncthreads();
  }
  for (int u = threadIdx.x; u < (n >> 2); u += THREADS) {  // strides 2 and 1 in registers
    ulonglong2* p = reinterpret_cast<ulonglong2*>(s + (u << 2));
    ulonglong2 v0 = p[0], v1 = p[1];
    const bool up = ((gbase + (size_t)(u << 2)) & (size_t)k) == 0;
    bt_cswap(v0.x, v1.x, up); bt_cswap(v0.y, v1.y, up);
    bt_cswap(v0.x, v0.y, up); bt_cswap(v1.x, v1.y, up);
    p[0] = v0; p[1] = v1;
  }
  __syncthreads();
}

// ascending sort of n keys (power of two >= 8) in shared memory
template <int THREADS>
__device__ __forceinline__ void bt_smem_sort(unsigned long long* s, int n) {
  bt_smem_init4<THREADS>(s, n);
  for (int k = 8; k <= n; k <<= 1) bt_smem_level<THREADS>(s, n, k >> 1, 0, k);
}
