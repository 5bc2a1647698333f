// Ad-hoc microbenchmarks (B200): dependent-chain latency of REDUX (__reduce_max_sync), SHFL butterflies, ballot,
// shared-memory loads through a shared-typed and a generic pointer.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(long long* out, unsigned* sink, unsigned char* gptr) {
  __shared__ unsigned sm[1024];
  __shared__ unsigned char sb[4096];
  const int lane = threadIdx.x;
  for (int i = lane; i < 1024; i += 32) sm[i] = (i * 7 + 3) & 1023;
  for (int i = lane; i < 4096; i += 32) sb[i] = (unsigned char)((i * 5 + 1) & 31);
  __syncwarp();
  unsigned v = lane * 17 + 1;
  const int n = 1000;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) v = __reduce_max_sync(0xffffffffu, v + lane) & 1023;
  long long t1 = clock64();
  out[0] = t1 - t0;
  t0 = clock64();
  for (int i = 0; i < n; ++i) { unsigned w = v + lane; for (int d = 16; d > 0; d >>= 1) w = max(w, __shfl_xor_sync(0xffffffffu, w, d)); v = w & 1023; }
  t1 = clock64();
  out[1] = t1 - t0;
  t0 = clock64();
  for (int i = 0; i < n; ++i) v = (__ballot_sync(0xffffffffu, ((v + i + lane) & 3) == 0) * 3) & 1023;
  t1 = clock64();
  out[2] = t1 - t0;
  t0 = clock64();
  for (int i = 0; i < n; ++i) v = sm[v];
  t1 = clock64();
  out[3] = t1 - t0;
  const unsigned char* gp = gptr ? gptr : sb;  // generic pointer (shared at run time)
  unsigned u = v & 31;
  t0 = clock64();
  for (int i = 0; i < n; ++i) u = gp[u * 32 + lane];
  t1 = clock64();
  out[4] = t1 - t0;
  sink[lane] = v + u;
}
int main() {
  long long* out; unsigned* sink; cudaMalloc(&out, 64); cudaMalloc(&sink, 256);
  k<<<1, 32>>>(out, sink, nullptr);
  long long h[5]; cudaMemcpy(h, out, 40, cudaMemcpyDeviceToHost);
  printf("cycles per dependent op: REDUX.max %.1f | 5-level SHFL butterfly max %.1f | ballot %.1f | LDS (shared ptr) %.1f | LD (generic ptr to shared) %.1f\n",
         h[0] / 1000.0, h[1] / 1000.0, h[2] / 1000.0, h[3] / 1000.0, h[4] / 1000.0);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
