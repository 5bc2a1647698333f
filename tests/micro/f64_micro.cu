// Ad-hoc microbenchmarks (B200): f64 dependent-chain latency and per-SM throughput, f64 div / sqrt cost.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void chain(double* out, long long* cyc, int n, double a, double b) {
  double x = a + threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) { x = x * b; x = x + a; }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void chain_fma(double* out, long long* cyc, int n, double a, double b) {
  double x = a + threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) { x = fma(x, b, a); x = fma(x, b, a); }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void indep(double* out, long long* cyc, int n, double a, double b) {
  double x[8];
  for (int k = 0; k < 8; ++k) x[k] = a + threadIdx.x + k;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) { x[k] = x[k] * b; x[k] = x[k] + a; }
  }
  long long t1 = clock64();
  double s = 0; for (int k = 0; k < 8; ++k) s += x[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void divs(double* out, long long* cyc, int n, double a, double b) {
  double x = a + threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) { x = b / x; }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void sqrts(double* out, long long* cyc, int n, double a, double b) {
  double x = a + threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) { x = sqrt(x) + b; }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024);
  long long h;
  const int n = 2000;
  for (int threads : {32, 128, 256, 512, 1024}) {
    chain<<<1, threads>>>(out, cyc, n, 1.0000001, 0.9999999); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("threads %4d: dep chain mul+add %.2f cyc/op", threads, (double)h / (2.0 * n));
    chain_fma<<<1, threads>>>(out, cyc, n, 1.0000001, 0.9999999); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf(" | dep fma %.2f cyc/op", (double)h / (2.0 * n));
    indep<<<1, threads>>>(out, cyc, n, 1.0000001, 0.9999999); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf(" | 8-way ILP %.2f cyc/op/thread => %.1f lane-ops/cyc/SM", (double)h / (16.0 * n), 16.0 * n * threads / (double)h);
    divs<<<1, threads>>>(out, cyc, n, 1.5, 3.0); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf(" | div %.1f cyc", (double)h / n);
    sqrts<<<1, threads>>>(out, cyc, n, 1.5, 3.0); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf(" | sqrt+add %.1f cyc\n", (double)h / n);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
