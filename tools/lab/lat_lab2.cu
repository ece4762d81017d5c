// LAB: single-warp issue throughput of independent FP64 instructions, and select/max sequences.
#include <cstdio>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)
constexpr int N = 2048;
__device__ __forceinline__ double dmaxsel(double a, double b) { return a > b ? a : b; }
__device__ __forceinline__ double relu_bits(double v) {
  int hi = __double2hiint(v), lo = __double2loint(v);
  const int m = ~(hi >> 31);
  return __hiloint2double(hi & m, lo & m);
}
template <int W>
__global__ void k(double* out, long long* cyc, double a, double b) {
  double x[8];
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3 + a * i;
  long long t0, t1;
  __syncthreads();
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = fma(x[j], b, a);
  }
  t1 = clock64(); if (threadIdx.x == 0) cyc[0] = t1 - t0;
  __syncthreads();
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = dmaxsel(x[j], a) * b;
  }
  t1 = clock64(); if (threadIdx.x == 0) cyc[1] = t1 - t0;
  __syncthreads();
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = relu_bits(x[j]) * b - a;
  }
  t1 = clock64(); if (threadIdx.x == 0) cyc[2] = t1 - t0;
  // dependent: select-max + add, bit relu + add
  double y = x[0];
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) y = dmaxsel(y, a) + b;
  t1 = clock64(); if (threadIdx.x == 0) cyc[3] = t1 - t0;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) y = relu_bits(y) - b;
  t1 = clock64(); if (threadIdx.x == 0) cyc[4] = t1 - t0;
  double sacc = 0;
  for (int i = 0; i < 8; ++i) sacc += x[i];
  out[threadIdx.x + blockIdx.x * blockDim.x] = sacc + y;
}
template <int W>
int run(const char* tag) {
  double* d; long long* c; long long h[8];
  CK(cudaMalloc(&d, 32 * W * 8)); CK(cudaMalloc(&c, 64));
  for (int rep = 0; rep < 2; ++rep) { k<W><<<1, 32 * W>>>(d, c, 1e-9, 0.999999); CK(cudaDeviceSynchronize()); }
  CK(cudaMemcpy(h, c, 40, cudaMemcpyDeviceToHost));
  printf("%s: 8 indep DFMA %.2f cyc/instr | 8 indep (cmp-select max + DMUL) %.2f cyc/pair | 8 indep (bit relu + DFMA) %.2f cyc/pair | dep selmax+DADD %.1f | dep bitrelu+DADD %.1f\n",
         tag, (double)h[0] / N / 8, (double)h[1] / N / 8, (double)h[2] / N / 8, (double)h[3] / N, (double)h[4] / N);
  return 0;
}
int main() { return run<1>("1 warp ") || run<4>("4 warps (1/SMSP)") || run<12>("12 warps (3/SMSP)"); }
