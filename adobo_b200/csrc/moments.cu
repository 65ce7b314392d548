// K2  csr_gene_moments -- per-gene sum(x), sum(x^2) of the normalized values (column reduction
// of a row-major CSR shard) and the host-side Seurat binning / z-scoring / ranking.
//
// Replaces adobo/hvg.py:60 (data.mean(axis=1)), :65 (var/mean per gene) and :62-71 (pd.cut into
// 20 equal-width bins, per-bin z-score, sort, head).  The device part is the O(nnz) column
// reduction; the O(g) binning (g <= ~33k genes) runs on the host inside this library, is
// deterministic, and is replicated on every rank after the all-reduce of the 2 g sums.
#include <algorithm>
#include <cmath>
#include <limits>

#include "common.cuh"

namespace adobo {

// One warp per cell; lanes stride the row (coalesced 128 B requests of ids and values); fp64
// reductions straight into L2 (RED.ADD.F64, no return value).
__global__ void __launch_bounds__(256) csr_gene_moments(const i64* __restrict__ indptr,
                                                        const int* __restrict__ indices,
                                                        const float* __restrict__ vals, double* __restrict__ gsum,
                                                        double* __restrict__ gsq, i64 n) {
  const int lane = threadIdx.x & 31;
  const i64 warps = ((i64)gridDim.x * blockDim.x) >> 5;
  for (i64 row = (((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5); row < n; row += warps) {
    const i64 start = indptr[row], end = indptr[row + 1];
#pragma unroll 4
    for (i64 e = start + lane; e < end; e += 32) {
      const int g = __ldg(indices + e);
      const double x = (double)__ldg(vals + e);
      atomicAdd(gsum + g, x);
      atomicAdd(gsq + g, x * x);
    }
  }
}

void run_moments(adobo_ctx* c) {
  REQUIRE(c->have_norm, ADOBO_ERR_STATE, "gene moments: run adobo_normalize / adobo_upload_normalized first");
  c->gsum.ensure((size_t)c->g);
  c->gsq.ensure((size_t)c->g);
  CUDA_CHECK(cudaMemsetAsync(c->gsum.p, 0, sizeof(double) * c->g, c->stream));
  CUDA_CHECK(cudaMemsetAsync(c->gsq.p, 0, sizeof(double) * c->g, c->stream));
  if (c->n > 0) {
    i64 blocks = std::min<i64>((c->n + 7) / 8, (i64)c->sm_count * 8);
    WORK(c, (double)c->nnz * 8.0 + 16.0 * c->g + 8.0 * (c->n + 1));
    LAUNCH(c, "csr_gene_moments", (csr_gene_moments<<<(unsigned)blocks, 256, 0, c->stream>>>(
                                      c->indptr.p, c->indices.p, c->vals.p, c->gsum.p, c->gsq.p, c->n)));
  }
  allreduce_sum_f64(c, c->gsum.p, (size_t)c->g);
  allreduce_sum_f64(c, c->gsq.p, (size_t)c->g);
  std::vector<double> s1(c->g), s2(c->g);
  CUDA_CHECK(cudaMemcpyAsync(s1.data(), c->gsum.p, sizeof(double) * c->g, cudaMemcpyDeviceToHost, c->stream));
  CUDA_CHECK(cudaMemcpyAsync(s2.data(), c->gsq.p, sizeof(double) * c->g, cudaMemcpyDeviceToHost, c->stream));
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  const double n = (double)c->n_total;
  c->h_mean.resize(c->g);
  c->h_var.resize(c->g);
  for (int i = 0; i < c->g; ++i) {
    double m = s1[i] / n;  // hvg.py:60: mean over ALL cells
    double ss = s2[i] - n * m * m;
    if (ss < 0) ss = 0;
    c->h_mean[i] = m;
    c->h_var[i] = ss / (n - 1.0);  // ddof = 1 (pandas var), NaN for a single cell like pandas
  }
  c->have_moments = true;
}

// hvg.py:62-71.  `order` = gene ids in the reference's output order (all genes that fall in a
// bin; the caller takes head(ngenes)); z = per-gene z-score (NaN where pandas gives NaN).
void run_seurat_host(adobo_ctx* c, int ngenes, int num_bins, std::vector<int>& order, std::vector<double>& z) {
  REQUIRE(c->have_moments, ADOBO_ERR_STATE, "seurat: gene moments missing");
  REQUIRE(num_bins >= 1, ADOBO_ERR_ARG, "seurat: num_bins must be >= 1");
  const int g = c->g;
  const double NaN = std::numeric_limits<double>::quiet_NaN();
  const std::vector<double>& mean = c->h_mean;
  const std::vector<double>& var = c->h_var;
  // pandas.cut(gene_mean, num_bins): equal-width, right-closed, lowest edge lowered by 0.1 % of range
  double mn = std::numeric_limits<double>::infinity(), mx = -mn;
  for (int i = 0; i < g; ++i)
    if (!std::isnan(mean[i])) {
      mn = std::min(mn, mean[i]);
      mx = std::max(mx, mean[i]);
    }
  std::vector<double> edges(num_bins + 1);
  auto linspace = [&](double a, double b) {
    // numpy.linspace: start + arange(num) * step, last point forced to `b`
    double step = (b - a) / num_bins;
    for (int i = 0; i <= num_bins; ++i) edges[i] = a + i * step;
    edges[num_bins] = b;
  };
  if (g == 0 || !(mn <= mx)) {
    order.clear();
    z.assign(g, NaN);
    return;
  }
  if (mn == mx) {
    double lo = mn - (mn != 0 ? 0.001 * std::fabs(mn) : 0.001);
    double hi = mx + (mx != 0 ? 0.001 * std::fabs(mx) : 0.001);
    linspace(lo, hi);
  } else {
    linspace(mn, mx);
    edges[0] -= (mx - mn) * 0.001;
  }
  std::vector<int> bin(g, -1);
  for (int i = 0; i < g; ++i) {
    if (std::isnan(mean[i])) continue;
    // searchsorted(edges, x, side='left'): first index with edges[idx] >= x
    int idx = (int)(std::lower_bound(edges.begin(), edges.end(), mean[i]) - edges.begin());
    if (idx == 0 || idx == num_bins + 1) continue;
    bin[i] = idx - 1;
  }
  z.assign(g, NaN);
  std::vector<std::vector<int>> members(num_bins);
  for (int i = 0; i < g; ++i)
    if (bin[i] >= 0) members[bin[i]].push_back(i);
  std::vector<int> concat;
  concat.reserve(g);
  for (int b = 0; b < num_bins; ++b) {
    const std::vector<int>& mem = members[b];
    if (mem.empty()) continue;
    std::vector<double> d(mem.size());
    double sum = 0;
    int cnt = 0;
    for (size_t t = 0; t < mem.size(); ++t) {
      d[t] = var[mem[t]] / mean[mem[t]];  // dispersion, hvg.py:65 (0/0 -> NaN)
      if (!std::isnan(d[t])) {
        sum += d[t];
        ++cnt;
      }
    }
    double m = cnt > 0 ? sum / cnt : NaN;
    double sd = NaN;
    if (cnt > 1) {
      double ss = 0;
      for (size_t t = 0; t < mem.size(); ++t)
        if (!std::isnan(d[t])) ss += (d[t] - m) * (d[t] - m);
      sd = std::sqrt(ss / (cnt - 1));
    }
    for (size_t t = 0; t < mem.size(); ++t) {
      z[mem[t]] = (d[t] - m) / sd;  // hvg.py:66
      concat.push_back(mem[t]);
    }
  }
  // sort_values(ascending=False): NaN last; ties keep concat order (the reference's own tie
  // order is arbitrary -- numpy quicksort)
  std::vector<int> good, bad;
  for (int id : concat) (std::isnan(z[id]) ? bad : good).push_back(id);
  std::stable_sort(good.begin(), good.end(), [&](int a, int b) { return z[a] > z[b]; });
  order = good;
  order.insert(order.end(), bad.begin(), bad.end());
  (void)ngenes;
}

}  // namespace adobo
