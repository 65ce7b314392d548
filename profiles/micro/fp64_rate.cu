// fp64 issue-rate probe for sm_100a: SIMT DFMA vs mma.sync.m8n8k4.f64 (per SM, per clock)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dfma(double* out, int iters) {
  double a[8];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x + i;
  const double x = 1.0000001, y = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fma(a[i], x, y);
  }
  double s = 0;
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void dmma(double* out, int iters) {
  double c[4][2];
  for (int i = 0; i < 4; ++i) c[i][0] = c[i][1] = 0;
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1e-3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1])
                   : "d"(a), "d"(b));
  }
  double s = 0;
  for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  double* out;
  cudaMalloc(&out, sizeof(double) * 148 * 1024 * 2);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  int clk;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  for (int threads : {128, 256, 512, 1024}) {
    const int iters = 20000;
    float ms;
    dfma<<<148, threads>>>(out, iters);
    cudaEventRecord(e0);
    dfma<<<148, threads>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    const double fma_total = 148.0 * threads * 8.0 * iters;
    printf("dfma threads=%d  %.3f ms  %.2f TFLOP/s  (%.1f fma/clk/SM at nominal %d MHz)\n", threads, ms,
           2 * fma_total / ms * 1e-9, fma_total / 148 / (ms * 1e-3 * clk * 1e3), clk / 1000);
    dmma<<<148, threads>>>(out, iters);
    cudaEventRecord(e0);
    dmma<<<148, threads>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    const double mfma = 148.0 * (threads / 32) * 4.0 * iters * 256.0;
    printf("dmma threads=%d  %.3f ms  %.2f TFLOP/s  (%.1f fma/clk/SM)\n", threads, ms, 2 * mfma / ms * 1e-9,
           mfma / 148 / (ms * 1e-3 * clk * 1e3));
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
