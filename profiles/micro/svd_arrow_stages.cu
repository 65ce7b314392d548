// Stage timing of svd_arrow (clock64 stamps between stages) on a synthetic restart matrix:
// B = [diag(D), z e_k | 0, T], m = 95.  Build: nvcc -DADOBO_SA_TIMING -I../../include ... (see run line in profiles/README)
#define ADOBO_SA_TIMING
#include "../../adobo_b200/csrc/svd_arrow.cu"
#include <cstdarg>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
namespace adobo {
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vfprintf(stderr, fmt, ap);
  va_end(ap);
  fputc('\n', stderr);
}
void prof_begin(adobo_ctx*, const char*) {}
void prof_end(adobo_ctx*) {}
void pool_free(void* p) { cudaFree(p); }
}  // namespace adobo
int main() {
  using namespace adobo;
  const int m = 95;
  adobo_ctx ctx;
  for (int k : {92, 90, 85, 75}) {
    std::vector<double> B((size_t)m * m, 0.0);
    srand(7);
    for (int i = 0; i < k; ++i) {
      B[(size_t)i * m + i] = 120.0 / (1.0 + 0.08 * i) + 1e-3 * (rand() % 1000);
      B[(size_t)i * m + k] = (i < 70 ? 1e-6 : 1e-2) * ((rand() % 2000) - 1000) / 1000.0;
    }
    for (int j = k; j < m; ++j) {
      B[(size_t)j * m + j] = 5.0 + 0.1 * (rand() % 100);
      if (j + 1 < m) B[(size_t)j * m + j + 1] = 2.0 + 0.05 * (rand() % 100);
    }
    double *dB, *dSu, *dsv, *dSv;
    int* dinfo;
    cudaMalloc(&dB, sizeof(double) * m * m);
    cudaMalloc(&dSu, sizeof(double) * m * m);
    cudaMalloc(&dSv, sizeof(double) * m * m);
    cudaMalloc(&dsv, sizeof(double) * m);
    cudaMalloc(&dinfo, sizeof(int) * 4);
    cudaMemcpy(dB, B.data(), sizeof(double) * m * m, cudaMemcpyHostToDevice);
    svd_arrow_prepare(m);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    float ms = 0;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      launch_svd_arrow(&ctx, dB, m, k, dSu, dsv, dSv, dinfo, 0, LastDiag());
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1);
    }
    long long st[16];
    cudaMemcpyFromSymbol(st, sa_stamps, sizeof(st));
    int info[4];
    cudaMemcpy(info, dinfo, sizeof(info), cudaMemcpyDeviceToHost);
    printf("k=%d t=%d  %.1f us  ok=%d na=%d  total %lld clk | stages:", k, m - k, ms * 1e3, info[3], -info[2], st[12] - st[0]);
    for (int i = 0; i < 12; ++i) printf(" S%d=%lld", i, st[i + 1] - st[i]);
    printf("  | S5 root 3: setup %lld clk, loop %lld clk, %lld evaluations\n", st[13], st[14], st[15]);
    long long z[16] = {0};
    cudaMemcpyToSymbol(sa_stamps, z, sizeof(z));
    std::vector<double> U((size_t)m * m), V((size_t)m * m), sv(m);
    cudaMemcpy(U.data(), dSu, sizeof(double) * m * m, cudaMemcpyDeviceToHost);
    cudaMemcpy(V.data(), dSv, sizeof(double) * m * m, cudaMemcpyDeviceToHost);
    cudaMemcpy(sv.data(), dsv, sizeof(double) * m, cudaMemcpyDeviceToHost);
    double rec = 0, ortu = 0, ortv = 0;
    for (int i = 0; i < m; ++i)
      for (int j = 0; j < m; ++j) {
        double b = 0, uu = 0, vv = 0;
        for (int r = 0; r < m; ++r) {
          b += sv[r] * U[(size_t)r * m + i] * V[(size_t)r * m + j];
          uu += U[(size_t)i * m + r] * U[(size_t)j * m + r];
          vv += V[(size_t)i * m + r] * V[(size_t)j * m + r];
        }
        rec = fmax(rec, fabs(b - B[(size_t)i * m + j]));
        ortu = fmax(ortu, fabs(uu - (i == j)));
        ortv = fmax(ortv, fabs(vv - (i == j)));
      }
    printf("  |U S V^T - B| = %.2e  |UU^T - I| = %.2e  |VV^T - I| = %.2e  sv[0]=%.6f sv[m-1]=%.3e  %s\n", rec, ortu, ortv, sv[0],
           sv[m - 1], cudaGetErrorString(cudaDeviceSynchronize()));
  }
  return 0;
}
