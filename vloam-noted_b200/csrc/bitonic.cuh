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
