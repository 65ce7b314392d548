// dependent-chain latencies (cycles) on sm_100a: DFMA, DADD, 1/x, a/b, rsqrt, sqrt, 64-bit shuffle, LDS
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void chain(double* out, long long* cyc, double seed, int n) {
  __shared__ double sh[64];
  sh[threadIdx.x & 63] = seed;
  __syncthreads();
  double x = seed + threadIdx.x * 1e-9, y = 1.0000001;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
    if (OP == 0) x = fma(x, y, 1e-9);
    if (OP == 1) x = x + y;
    if (OP == 2) x = 1.0 / x + 0.5;
    if (OP == 3) x = y / x + 0.5;
    if (OP == 4) x = rsqrt(x) + 1.0;
    if (OP == 5) x = sqrt(x) + 1.0;
    if (OP == 6) x = __shfl_xor_sync(0xffffffffu, x, 1);
    if (OP == 7) x = x + __shfl_xor_sync(0xffffffffu, x, 1);
    if (OP == 8) x = sh[((int)x) & 63] + 1.0;
    if (OP == 9) x = (double)((float)x * 1.0001f);
    if (OP == 10) x = x * y;
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
template <int OP>
void run(const char* name, double seed) {
  double* out;
  long long* cyc;
  cudaMalloc(&out, 8 * 64);
  cudaMalloc(&cyc, 8);
  const int n = 4096;
  chain<OP><<<1, 32>>>(out, cyc, seed, n);
  chain<OP><<<1, 32>>>(out, cyc, seed, n);
  long long h;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-28s %.1f cycles/iter\n", name, (double)h / n);
}
int main() {
  run<0>("dfma", 1.0);
  run<1>("dadd", 1.0);
  run<10>("dmul", 1.0);
  run<2>("1/x + dadd", 1.3);
  run<3>("y/x + dadd", 1.3);
  run<4>("rsqrt + dadd", 1.3);
  run<5>("sqrt + dadd", 1.3);
  run<6>("shfl64", 1.0);
  run<7>("shfl64 + dadd", 1.0);
  run<8>("f2i + lds + dadd", 1.0);
  run<9>("f64->f32 mul ->f64", 1.0);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
