// K1  csr_libsize_scale_log -- normalize.standard + log step on the CSR count shard.
//
// Replaces adobo/normalize.py:285-286 (col_sums; (data / col_sums) * scaling_factor) and
// :559-560 (log_func(norm + small_const)).  One warp per cell (CSR row):
//   pass 1  128-bit loads of the row's counts, int64 library size by warp shuffle;
//   LUT     lane l evaluates y(l+1) = log((l+1)/S * scale + c) ONCE in fp64 -- counts are small
//           integers, so >99 % of the non-zeros take their value from a lane by shuffle and the
//           fp64 log never limits the kernel (it is HBM bound: 4 B in + 4 B out per non-zero);
//   pass 2  re-reads the row (L2 hit), writes fp32 values with 128-bit stores (and fp64 values for
//           the host-facing API when asked).
// The device copy holds y - z0 with z0 = log(small_const), the value every implicit zero takes (0 for the
// reference's default small_const = 1; preproc.impute logs raw + 1.01, preproc.py:472): the matrix stays sparse for
// any small_const, gene means get z0 back (moments.cu), centring removes it from the PCA.
#include "common.cuh"

namespace adobo {

__device__ __forceinline__ double norm_value(int c, double S, double scaling, int log_mode, double sc) {
  double x = ((double)c / S) * scaling;  // divide first, then multiply: normalize.py:286
  switch (log_mode) {
    case ADOBO_LOG_2: return log2(x + sc);
    case ADOBO_LOG_E: return log(x + sc);
    case ADOBO_LOG_10: return log10(x + sc);
    default: return x;
  }
}

template <bool OUT64>
__global__ void __launch_bounds__(256) csr_libsize_scale_log(const i64* __restrict__ indptr,
                                                             const int* __restrict__ counts,
                                                             float* __restrict__ vals, double* __restrict__ out64,
                                                             i64* __restrict__ libsize, i64 n, double scaling,
                                                             int log_mode, double small_const, double z0) {
  const int lane = threadIdx.x & 31;
  const i64 warps = ((i64)gridDim.x * blockDim.x) >> 5;
  for (i64 row = (((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5); row < n; row += warps) {
    const i64 start = indptr[row], end = indptr[row + 1];
    i64 a0 = (start + 3) & ~(i64)3;
    if (a0 > end) a0 = end;
    i64 a1 = end & ~(i64)3;
    if (a1 < a0) a1 = a0;
    // ---- pass 1: library size ----
    i64 sum = 0;
    for (i64 e = start + lane; e < a0; e += 32) sum += counts[e];
#pragma unroll 4
    for (i64 e = a0 + 4 * lane; e < a1; e += 128) {
      int4 v = *reinterpret_cast<const int4*>(counts + e);
      sum += (i64)v.x + (i64)v.y + (i64)v.z + (i64)v.w;
    }
    for (i64 e = a1 + lane; e < end; e += 32) sum += counts[e];
    sum = warp_sum(sum);
    if (lane == 0 && libsize) libsize[row] = sum;
    const double S = (double)sum;
    // ---- per-row table: lane l holds y(count = l + 1) ----
    const double lut64 = norm_value(lane + 1, S, scaling, log_mode, small_const);
    const float lut32 = (float)(lut64 - z0);
    // ---- pass 2 ----
    // head / tail (at most 3 elements each): direct evaluation
    for (i64 e = start + lane; e < a0; e += 32) {
      double y = norm_value(counts[e], S, scaling, log_mode, small_const);
      vals[e] = (float)(y - z0);
      if (OUT64) out64[e] = y;
    }
    for (i64 e = a1 + lane; e < end; e += 32) {
      double y = norm_value(counts[e], S, scaling, log_mode, small_const);
      vals[e] = (float)(y - z0);
      if (OUT64) out64[e] = y;
    }
    for (i64 base = a0; base < a1; base += 128) {  // warp-uniform trip count: shuffles inside
      const i64 e = base + 4 * lane;
      const bool act = e < a1;
      int4 v = make_int4(1, 1, 1, 1);
      if (act) v = ld_stream(reinterpret_cast<const int4*>(counts + e));
      int cc[4] = {v.x, v.y, v.z, v.w};
      float y32[4];
      double y64[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int src = (cc[q] - 1) & 31;
        y32[q] = __shfl_sync(0xffffffffu, lut32, src);
        if (OUT64) y64[q] = __shfl_sync(0xffffffffu, lut64, src);
        if (cc[q] < 1 || cc[q] > 32) {  // rare: large count, evaluate directly
          double y = norm_value(cc[q], S, scaling, log_mode, small_const);
          y32[q] = (float)(y - z0);
          if (OUT64) y64[q] = y;
        }
      }
      if (act) {
        st_stream(reinterpret_cast<float4*>(vals + e), make_float4(y32[0], y32[1], y32[2], y32[3]));
        if (OUT64) {
          double2* o = reinterpret_cast<double2*>(out64 + e);
          o[0] = make_double2(y64[0], y64[1]);
          o[1] = make_double2(y64[2], y64[3]);
        }
      }
    }
  }
}

// fp64 values supplied by the caller (hvg/pca on an existing normalization) -> fp32 device copy
__global__ void cast_f64_f32(const double* __restrict__ in, float* __restrict__ out, i64 n) {
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x)
    out[i] = (float)in[i];
}

void run_normalize(adobo_ctx* c, double scaling, int log_mode, double small_const, i64* out_lib, double* out_vals) {
  REQUIRE(c->have_counts, ADOBO_ERR_STATE, "adobo_normalize: no count matrix uploaded");
  c->vals.ensure((size_t)c->nnz);
  DevBuf<i64> lib;
  DevBuf<double> v64;
  if (out_lib) lib.ensure((size_t)c->n);
  if (out_vals) v64.ensure((size_t)c->nnz);
  i64 blocks = (c->n + 7) / 8;
  i64 cap = (i64)c->sm_count * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  const double z0 = log_mode == ADOBO_LOG_2 ? log2(small_const) : log_mode == ADOBO_LOG_E ? log(small_const)
                  : log_mode == ADOBO_LOG_10 ? log10(small_const) : 0.0;
  REQUIRE(z0 == z0 && z0 > -1e300, ADOBO_ERR_ARG, "normalize: small_const must be positive when log is applied");
  c->zero_level = z0;
  if (c->n > 0) {
    WORK(c, (double)c->nnz * (out_vals ? 16.0 : 8.0) + 8.0 * (c->n + 1));
    if (out_vals)
      LAUNCH(c, "csr_libsize_scale_log", (csr_libsize_scale_log<true><<<(unsigned)blocks, 256, 0, c->stream>>>(
                                             c->indptr.p, c->counts.p, c->vals.p, v64.p, lib.p, c->n, scaling,
                                             log_mode, small_const, z0)));
    else
      LAUNCH(c, "csr_libsize_scale_log", (csr_libsize_scale_log<false><<<(unsigned)blocks, 256, 0, c->stream>>>(
                                             c->indptr.p, c->counts.p, c->vals.p, nullptr, lib.p, c->n, scaling,
                                             log_mode, small_const, z0)));
  }
  c->have_norm = true;
  c->have_moments = false;
  {  // bound of the values for the fixed-point gene moments: count / S <= 1
    const double top = scaling;
    c->val_bound = (log_mode == ADOBO_LOG_2 ? log2(top + small_const) : log_mode == ADOBO_LOG_E ? log(top + small_const)
                  : log_mode == ADOBO_LOG_10 ? log10(top + small_const) : top) - z0;
    if (!(c->val_bound > 0)) c->val_bound = 0;  // measured instead
  }
  if (out_lib) CUDA_CHECK(cudaMemcpyAsync(out_lib, lib.p, sizeof(i64) * c->n, cudaMemcpyDeviceToHost, c->stream));
  if (out_vals)
    CUDA_CHECK(cudaMemcpyAsync(out_vals, v64.p, sizeof(double) * c->nnz, cudaMemcpyDeviceToHost, c->stream));
  if (out_lib || out_vals) CUDA_CHECK(cudaStreamSynchronize(c->stream));
}

// ---- dataset metadata (data.py:75-84): reads per cell, detected genes per cell, expressing cells per gene ------------
// warp per cell, 128-bit loads where the row is aligned; the per-gene counts are integer REDs (order independent).
__global__ void __launch_bounds__(256) csr_count_stats(const i64* __restrict__ indptr, const int* __restrict__ indices,
                                                       const int* __restrict__ counts, i64 n, i64* __restrict__ total,
                                                       i64* __restrict__ detected, int* __restrict__ expressed) {
  const int lane = threadIdx.x & 31;
  const i64 nw = ((i64)gridDim.x * blockDim.x) >> 5;
  for (i64 r = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n; r += nw) {
    const i64 b = indptr[r], e = indptr[r + 1];
    i64 s = 0;
    int d = 0;
    for (i64 i = b + lane; i < e; i += 32) {
      const int v = __ldg(counts + i);
      s += v;
      if (v > 0) {
        ++d;
        atomicAdd(expressed + __ldg(indices + i), 1);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      d += __shfl_xor_sync(0xffffffffu, d, o);
    }
    if (lane == 0) total[r] = s, detected[r] = d;
  }
}

void run_count_stats(adobo_ctx* c, i64* out_total, i64* out_detected, i64* out_expressed) {
  REQUIRE(c->have_counts, ADOBO_ERR_STATE, "adobo_count_stats: no count matrix on the device");
  DevBuf<i64> tot, det;
  DevBuf<int> expr;
  tot.ensure((size_t)std::max<i64>(c->n, 1));
  det.ensure((size_t)std::max<i64>(c->n, 1));
  expr.ensure((size_t)std::max<i64>(c->g, 1));
  CUDA_CHECK(cudaMemsetAsync(expr.p, 0, sizeof(int) * std::max<i64>(c->g, 1), c->stream));
  if (c->n > 0) {
    const i64 blocks = std::max<i64>(1, std::min<i64>((c->n + 7) / 8, (i64)c->sm_count * 8));
    WORK(c, 8.0 * (double)c->nnz + 24.0 * c->n);
    LAUNCH(c, "csr_count_stats", (csr_count_stats<<<(unsigned)blocks, 256, 0, c->stream>>>(
                                     c->indptr.p, c->indices.p, c->counts.p, c->n, tot.p, det.p, expr.p)));
  }
  std::vector<int> he((size_t)c->g);
  if (out_total) CUDA_CHECK(cudaMemcpyAsync(out_total, tot.p, sizeof(i64) * c->n, cudaMemcpyDeviceToHost, c->stream));
  if (out_detected) CUDA_CHECK(cudaMemcpyAsync(out_detected, det.p, sizeof(i64) * c->n, cudaMemcpyDeviceToHost, c->stream));
  if (out_expressed && c->g > 0)
    CUDA_CHECK(cudaMemcpyAsync(he.data(), expr.p, sizeof(int) * c->g, cudaMemcpyDeviceToHost, c->stream));
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  if (out_expressed)
    for (i64 j = 0; j < c->g; ++j) out_expressed[j] = he[(size_t)j];
}

void upload_normalized_values(adobo_ctx* c, const double* values) {
  DevBuf<double> tmp;
  tmp.ensure((size_t)c->nnz);
  c->vals.ensure((size_t)c->nnz);
  CUDA_CHECK(cudaMemcpyAsync(tmp.p, values, sizeof(double) * c->nnz, cudaMemcpyHostToDevice, c->stream));
  if (c->nnz > 0)
    LAUNCH(c, "cast_f64_f32", (cast_f64_f32<<<c->sm_count * 4, 256, 0, c->stream>>>(tmp.p, c->vals.p, c->nnz)));
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  c->have_norm = true;
  c->have_moments = false;
  c->val_bound = 0;  // unknown values: run_moments measures max |x|
  c->zero_level = 0;
}

}  // namespace adobo
