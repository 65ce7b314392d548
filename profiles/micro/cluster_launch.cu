// launch + barrier cost of thread-block clusters on sm_100a: dependent chain of tiny kernels in one stream
// (a) 8 x 1024 threads, no cluster  (b) cluster of 8, no barrier  (c) cluster of 8, two cluster.sync + DSMEM push
// (d) 148 x 1024 threads with a last-CTA atomic reduction (the pattern of the grid-wide kernels)
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;
__global__ void k_plain(double* x) {
  if (threadIdx.x == 0 && blockIdx.x == 0) x[0] += 1.0;
}
__global__ void k_cluster(double* x, int syncs) {
  cg::cluster_group cl = cg::this_cluster();
  __shared__ double box[8];
  for (int s = 0; s < syncs; ++s) {
    if (threadIdx.x < cl.num_blocks()) cl.map_shared_rank(box, threadIdx.x)[cl.block_rank()] = x[blockIdx.x] + s;
    cl.sync();
  }
  if (threadIdx.x == 0 && blockIdx.x == 0) x[0] += box[0] * 1e-30 + 1.0;
}
__global__ void k_lastcta(double* x, double* part, unsigned* counter) {
  __shared__ bool last;
  if (threadIdx.x == 0) {
    part[blockIdx.x] = x[1] + blockIdx.x;
    __threadfence();
    last = atomicAdd(counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last) {
    double s = 0;
    for (int b = threadIdx.x; b < gridDim.x; b += blockDim.x) s += __ldcg(part + b);
    if (threadIdx.x == 0) {
      x[0] += s * 1e-30 + 1.0;
      *counter = 0;
    }
  }
}
template <typename F>
float timeit(F f, int n) {
  cudaEvent_t a, b;
  cudaEventCreate(&a), cudaEventCreate(&b);
  cudaStream_t st;
  cudaStreamCreate(&st);
  // captured into a graph, like the product runs them
  cudaGraph_t g;
  cudaGraphExec_t ge;
  cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
  for (int i = 0; i < n; ++i) f(st);
  cudaStreamEndCapture(st, &g);
  cudaGraphInstantiate(&ge, g, 0);
  cudaGraphLaunch(ge, st);
  cudaStreamSynchronize(st);
  cudaEventRecord(a, st);
  cudaGraphLaunch(ge, st);
  cudaEventRecord(b, st);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms * 1e3f / n;
}
int main() {
  double *x, *part;
  unsigned* counter;
  cudaMalloc(&x, 8 * 1024);
  cudaMalloc(&part, 8 * 1024);
  cudaMalloc(&counter, 4);
  cudaMemset(x, 0, 8 * 1024);
  cudaMemset(counter, 0, 4);
  const int n = 400;
  printf("graph of %d dependent launches, us per launch:\n", n);
  printf("  8 x 1024, no cluster            %.2f\n", timeit([&](cudaStream_t st) { k_plain<<<8, 1024, 0, st>>>(x); }, n));
  printf("  148 x 1024, no cluster          %.2f\n", timeit([&](cudaStream_t st) { k_plain<<<148, 1024, 0, st>>>(x); }, n));
  for (int syncs : {0, 1, 2, 3}) {
    float t = timeit(
        [&](cudaStream_t st) {
          cudaLaunchConfig_t cfg = {};
          cfg.gridDim = dim3(8);
          cfg.blockDim = dim3(1024);
          cfg.stream = st;
          cudaLaunchAttribute at[1];
          at[0].id = cudaLaunchAttributeClusterDimension;
          at[0].val.clusterDim.x = 8, at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = 1;
          cfg.attrs = at, cfg.numAttrs = 1;
          cudaLaunchKernelEx(&cfg, k_cluster, x, syncs);
        },
        n);
    printf("  cluster of 8 x 1024, %d sync(s)   %.2f\n", syncs, t);
  }
  printf("  148 x 1024, last-CTA reduction  %.2f\n",
         timeit([&](cudaStream_t st) { k_lastcta<<<148, 1024, 0, st>>>(x, part, counter); }, n));
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
