// Seeded synthetic count matrices generated ON the device -- the data source of bench.py for the configurations a
// host generator cannot feed in reasonable time (C3 68 k x 33 k, C4 1.3 M x 28 k; the numpy generator
// of adobo_b200/synth.py needs 43 s for 33 M non-zeros).  Not part of the reference's path: test / bench plumbing.
//
// Model (the one of synth.py): lambda(cell, gene) = table[cluster(cell)][gene] * size(cell), count ~ Poisson(lambda *
// Gamma(2, 1/2)); table = exp(base + off + lfc[cluster]) comes from the host (8 x g floats).  Every draw is a pure
// function of (seed, global cell index, gene) through Philox4x32-10, so any contiguous shard of cells can be
// generated on its own rank and the count pass and the fill pass see the same matrix.  A cell without any count
// gets one at gene (7919 cell) mod g (library size 0 divides by zero in normalize.standard, normalize.py:285-286).
#include "common.cuh"

namespace adobo {

struct Philox {
  uint32_t k0, k1;
  __device__ __forceinline__ static void round(uint32_t (&c)[4], uint32_t ka, uint32_t kb) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ ka, n2 = hi0 ^ c[3] ^ kb;
    c[0] = n0, c[1] = lo1, c[2] = n2, c[3] = lo0;
  }
  __device__ __forceinline__ void operator()(uint32_t a, uint32_t b, uint32_t cc, uint32_t d, uint32_t (&out)[4]) const {
    uint32_t c[4] = {a, b, cc, d};
    uint32_t ka = k0, kb = k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      round(c, ka, kb);
      ka += 0x9E3779B9u;
      kb += 0xBB67AE85u;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) out[i] = c[i];
  }
};
__device__ __forceinline__ float u01(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }  // (0, 1)

// per-cell cluster and size factor exp(N(0, 0.35))
__device__ __forceinline__ void cell_params(const Philox& rng, i64 cell, int n_clusters, int& cl, float& size) {
  uint32_t r[4];
  rng((uint32_t)cell, (uint32_t)(cell >> 32), 0xC311u, 0u, r);
  cl = (int)(r[0] % (uint32_t)n_clusters);
  const float z = sqrtf(-2.0f * __logf(u01(r[1]))) * __cosf(6.2831853f * u01(r[2]));
  size = __expf(0.35f * z);
}

__device__ __forceinline__ int draw_count(const Philox& rng, i64 cell, int gene, float lam) {
  uint32_t r[4];
  rng((uint32_t)cell, (uint32_t)(cell >> 32), (uint32_t)gene, 1u, r);
  const float gam = -0.5f * __logf(u01(r[0]) * u01(r[1]));  // Gamma(2, scale 1/2)
  const float m = lam * gam;
  const float u = u01(r[2]);
  if (m < 24.0f) {  // inversion of the Poisson CDF
    float p = __expf(-m), cdf = p;
    int k = 0;
    while (u > cdf && k < 200) {
      ++k;
      p *= m / (float)k;
      cdf += p;
    }
    return k;
  }
  const float z = sqrtf(-2.0f * __logf(u)) * __cosf(6.2831853f * u01(r[3]));  // normal approximation, large means
  const float v = m + sqrtf(m) * z;
  return v > 0.f ? (int)(v + 0.5f) : 0;
}

// FILL = false: non-zeros per row;  FILL = true: ordered compaction of every row into (indices, counts)
template <bool FILL>
__global__ void __launch_bounds__(256) synth_rows(i64 n_rows, int g, i64 row_begin, Philox rng, const float* __restrict__ table,
                                                  int n_clusters, i64* __restrict__ rowcnt, const i64* __restrict__ indptr,
                                                  int* __restrict__ indices, int* __restrict__ counts) {
  const int lane = threadIdx.x & 31;
  const i64 warps = ((i64)gridDim.x * blockDim.x) >> 5;
  for (i64 row = (((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5); row < n_rows; row += warps) {
    const i64 cell = row_begin + row;
    int cl;
    float size;
    cell_params(rng, cell, n_clusters, cl, size);
    const float* tab = table + (size_t)cl * g;
    i64 out = FILL ? indptr[row] : 0;
    const i64 out0 = out;
    int total = 0;
    for (int base = 0; base < g; base += 32) {  // warp-uniform trip count
      const int gene = base + lane;
      int c = 0;
      if (gene < g) c = draw_count(rng, cell, gene, __ldg(tab + gene) * size);
      const unsigned keep = __ballot_sync(0xffffffffu, c > 0);
      if (FILL && c > 0) {
        const i64 pos = out + __popc(keep & ((1u << lane) - 1u));
        indices[pos] = gene;
        counts[pos] = c;
      }
      out += __popc(keep);
      total += __popc(keep);
    }
    if (total == 0) {  // forced count
      if (FILL && lane == 0) {
        indices[out0] = (int)((cell * 7919) % g);
        counts[out0] = 1;
      }
      total = 1;
    }
    if (!FILL && lane == 0) rowcnt[row] = total;
  }
}

void run_synth_counts(adobo_ctx* c, i64 n_rows, int g, i64 row_begin, uint64_t seed, const float* table_host, int n_clusters,
                      i64* out_nnz) {
  REQUIRE(n_rows >= 0 && g >= 1 && n_clusters >= 1 && table_host, ADOBO_ERR_ARG, "synth_counts: bad arguments");
  DevBuf<float> table;
  table.ensure((size_t)n_clusters * g);
  CUDA_CHECK(cudaMemcpyAsync(table.p, table_host, sizeof(float) * (size_t)n_clusters * g, cudaMemcpyHostToDevice, c->stream));
  Philox rng{(uint32_t)seed, (uint32_t)(seed >> 32) ^ 0xADB0B200u};
  DevBuf<i64> rowcnt;
  rowcnt.ensure((size_t)n_rows + 2);
  CUDA_CHECK(cudaMemsetAsync(rowcnt.p, 0, sizeof(i64) * (n_rows + 2), c->stream));
  c->indptr.ensure((size_t)n_rows + 2);
  const unsigned blocks = (unsigned)std::max<i64>(1, std::min<i64>((n_rows + 7) / 8, (i64)c->sm_count * 16));
  if (n_rows > 0)
    LAUNCH(c, "synth_rows", (synth_rows<false><<<blocks, 256, 0, c->stream>>>(n_rows, g, row_begin, rng, table.p, n_clusters,
                                                                             rowcnt.p, nullptr, nullptr, nullptr)));
  exclusive_scan_i64(c, rowcnt.p, c->indptr.p, n_rows);
  i64 nnz = 0;
  CUDA_CHECK(cudaMemcpyAsync(&nnz, c->indptr.p + n_rows, sizeof(i64), cudaMemcpyDeviceToHost, c->stream));
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  c->indices.ensure((size_t)nnz);
  c->counts.ensure((size_t)nnz);
  if (n_rows > 0)
    LAUNCH(c, "synth_rows", (synth_rows<true><<<blocks, 256, 0, c->stream>>>(n_rows, g, row_begin, rng, table.p, n_clusters,
                                                                            nullptr, c->indptr.p, c->indices.p, c->counts.p)));
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  c->n = n_rows;
  c->g = g;
  c->nnz = nnz;
  if (out_nnz) *out_nnz = nnz;
}

}  // namespace adobo
