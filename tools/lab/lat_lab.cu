// LAB: dependent-issue latencies on B200 (one warp): DFMA, DADD, DMUL, DSETP+FSEL (max), STS->LDS round trip.
#include <cstdio>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)
constexpr int N = 4096;
__global__ void k(double* out, long long* cyc, double a, double b) {
  __shared__ double sh[64];
  double x = threadIdx.x * 1e-3 + a;
  long long t0, t1;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = fma(x, b, a);
  t1 = clock64(); if (threadIdx.x == 0) cyc[0] = t1 - t0;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = x + a;
  t1 = clock64(); if (threadIdx.x == 0) cyc[1] = t1 - t0;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = x * b;
  t1 = clock64(); if (threadIdx.x == 0) cyc[2] = t1 - t0;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = fmax(x, a) + b;   // max + add
  t1 = clock64(); if (threadIdx.x == 0) cyc[3] = t1 - t0;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) { sh[threadIdx.x] = x; __syncwarp(); x = sh[(threadIdx.x + 1) & 31] + a; __syncwarp(); }
  t1 = clock64(); if (threadIdx.x == 0) cyc[4] = t1 - t0;
  // two independent chains: throughput-limited?
  double y = x + 1.0;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) { x = fma(x, b, a); y = fma(y, b, a); }
  t1 = clock64(); if (threadIdx.x == 0) cyc[5] = t1 - t0;
  float f = (float)x;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) f = fmaf(f, (float)b, (float)a);
  t1 = clock64(); if (threadIdx.x == 0) cyc[6] = t1 - t0;
  out[threadIdx.x] = x + y + f;
}
int main() {
  double* d; long long* c; long long h[8];
  CK(cudaMalloc(&d, 64 * 8)); CK(cudaMalloc(&c, 64));
  for (int rep = 0; rep < 2; ++rep) { k<<<1, 32>>>(d, c, 1e-9, 0.999999); CK(cudaDeviceSynchronize()); }
  CK(cudaMemcpy(h, c, 56, cudaMemcpyDeviceToHost));
  const char* n[] = {"DFMA", "DADD", "DMUL", "DMAX+DADD", "STS+sync+LDS+DADD", "2xDFMA (independent)", "FFMA"};
  for (int i = 0; i < 7; ++i) printf("%-22s %.1f cycles / step\n", n[i], (double)h[i] / N);
  return 0;
}
