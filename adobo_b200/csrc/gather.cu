// K3  HVG column subset of the normalized CSR shard, in the two layouts the Lanczos mat-vecs
// read (CSR for A.v, CSC for A^T.u), plus the per-gene centring / scaling statistics.
//
// Replaces adobo/dr.py:319 (data[data.index.isin(hvg)] -- original gene order is kept) and the
// statistics of sklearn.preprocessing.scale as dr.irlb calls it (dr.py:156-160): column mean and
// POPULATION std (ddof=0), std < 10*eps -> 1.  The reference densifies and rewrites the matrix;
// here the subset stays sparse and (mu, 1/sigma) are applied inside the mat-vecs.
#include <cub/cub.cuh>

#include "common.cuh"

namespace adobo {

void exclusive_scan_i64(adobo_ctx* c, const i64* in, i64* out, i64 n) {
  // scans n+1 items; the caller guarantees in[n] exists (its value is ignored: out is exclusive)
  size_t bytes = 0;
  CUDA_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, (int)(n + 1), c->stream));
  c->scratch.ensure(bytes);
  CUDA_CHECK(cub::DeviceScan::ExclusiveSum(c->scratch.p, bytes, in, out, (int)(n + 1), c->stream));
  c->launches++;
}

// per row: number of entries whose gene is selected (colmap[g] >= 0)
__global__ void __launch_bounds__(256) subset_count(const i64* __restrict__ indptr, const int* __restrict__ indices,
                                                    const int* __restrict__ colmap, i64* __restrict__ rowcnt,
                                                    i64 n) {
  const int lane = threadIdx.x & 31;
  const i64 warps = ((i64)gridDim.x * blockDim.x) >> 5;
  for (i64 row = (((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5); row < n; row += warps) {
    const i64 start = indptr[row], end = indptr[row + 1];
    int cnt = 0;
    for (i64 e = start + lane; e < end; e += 32) cnt += (colmap[__ldg(indices + e)] >= 0);
    cnt = warp_sum(cnt);
    if (lane == 0) rowcnt[row] = cnt;
  }
}

// ordered compaction of each row into the subset CSR (columns renumbered 0..h-1, still ascending)
__global__ void __launch_bounds__(256) subset_fill(const i64* __restrict__ indptr, const int* __restrict__ indices,
                                                   const float* __restrict__ vals, const int* __restrict__ colmap,
                                                   const i64* __restrict__ rptr, int* __restrict__ cidx,
                                                   float* __restrict__ rval, int* __restrict__ rowof, i64 n) {
  const int lane = threadIdx.x & 31;
  const i64 warps = ((i64)gridDim.x * blockDim.x) >> 5;
  for (i64 row = (((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5); row < n; row += warps) {
    const i64 start = indptr[row], end = indptr[row + 1];
    i64 out = rptr[row];
    for (i64 base = start; base < end; base += 32) {  // warp-uniform trip count
      const i64 e = base + lane;
      int nc = -1;
      float v = 0.f;
      if (e < end) {
        nc = colmap[__ldg(indices + e)];
        v = __ldg(vals + e);
      }
      const unsigned keep = __ballot_sync(0xffffffffu, nc >= 0);
      if (nc >= 0) {
        const i64 pos = out + __popc(keep & ((1u << lane) - 1u));
        cidx[pos] = nc;
        rval[pos] = v;
        rowof[pos] = (int)row;
      }
      out += __popc(keep);
    }
  }
}

__global__ void iota_u32(unsigned* p, i64 n) {
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x)
    p[i] = (unsigned)i;
}

__global__ void col_histogram(const int* __restrict__ cidx, i64 nnz, i64* __restrict__ hist) {
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (i64)gridDim.x * blockDim.x)
    atomicAdd((unsigned long long*)(hist + cidx[i]), 1ull);
}

template <typename VT>
__global__ void csc_permute(const unsigned* __restrict__ perm, const int* __restrict__ rowof,
                            const VT* __restrict__ rval, int* __restrict__ ridx, VT* __restrict__ cval, i64 nnz) {
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (i64)gridDim.x * blockDim.x) {
    const unsigned e = perm[i];
    ridx[i] = rowof[e];
    cval[i] = rval[e];
  }
}

__global__ void rows_of_entries(const i64* __restrict__ rptr, int* __restrict__ rowof, i64 n) {
  const int lane = threadIdx.x & 31;
  const i64 warps = ((i64)gridDim.x * blockDim.x) >> 5;
  for (i64 row = (((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5); row < n; row += warps)
    for (i64 e = rptr[row] + lane; e < rptr[row + 1]; e += 32) rowof[e] = (int)row;
}

// CSC copy of a CSR matrix with rows ascending inside every column: stable radix sort of the
// entry ids by column (one-time format conversion, not a hot op).
template <typename VT>
void build_csc_impl(adobo_ctx* c, SparsePair<VT>& A, const int* rowof) {
  REQUIRE(A.nnz < ((i64)1 << 31), ADOBO_ERR_LIMIT, "subset has %lld non-zeros on one GPU (limit 2^31)",
          (long long)A.nnz);
  A.cptr.ensure((size_t)A.cols + 2);
  A.ridx.ensure((size_t)A.nnz);
  A.cval.ensure((size_t)A.nnz);
  DevBuf<i64> hist;
  hist.ensure((size_t)A.cols + 2);
  CUDA_CHECK(cudaMemsetAsync(hist.p, 0, sizeof(i64) * (A.cols + 2), c->stream));
  if (A.nnz > 0) {
    LAUNCH(c, "col_histogram", (col_histogram<<<c->sm_count * 4, 256, 0, c->stream>>>(A.cidx.p, A.nnz, hist.p)));
  }
  exclusive_scan_i64(c, hist.p, A.cptr.p, A.cols);
  if (A.nnz == 0) return;
  DevBuf<int> keys_out;
  DevBuf<unsigned> perm_in, perm_out;
  keys_out.ensure((size_t)A.nnz);
  perm_in.ensure((size_t)A.nnz);
  perm_out.ensure((size_t)A.nnz);
  LAUNCH(c, "iota_u32", (iota_u32<<<c->sm_count * 4, 256, 0, c->stream>>>(perm_in.p, A.nnz)));
  int bits = 1;
  while ((1ll << bits) < A.cols) ++bits;
  size_t bytes = 0;
  CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, bytes, A.cidx.p, keys_out.p, perm_in.p, perm_out.p,
                                             (int)A.nnz, 0, bits, c->stream));
  c->scratch.ensure(bytes);
  CUDA_CHECK(cub::DeviceRadixSort::SortPairs(c->scratch.p, bytes, A.cidx.p, keys_out.p, perm_in.p, perm_out.p,
                                             (int)A.nnz, 0, bits, c->stream));
  c->launches += 3;
  LAUNCH(c, "csc_permute", (csc_permute<VT><<<c->sm_count * 4, 256, 0, c->stream>>>(perm_out.p, rowof, A.rval.p,
                                                                                  A.ridx.p, A.cval.p, A.nnz)));
  CUDA_CHECK(cudaStreamSynchronize(c->stream));  // temporaries die here
}

template <typename VT>
void build_csc(adobo_ctx* c, SparsePair<VT>& A) {
  DevBuf<int> rowof;
  rowof.ensure((size_t)A.nnz);
  if (A.rows > 0 && A.nnz > 0) {
    i64 blocks = std::min<i64>((A.rows + 7) / 8, (i64)c->sm_count * 8);
    LAUNCH(c, "rows_of_entries", (rows_of_entries<<<(unsigned)blocks, 256, 0, c->stream>>>(A.rptr.p, rowof.p, A.rows)));
  }
  build_csc_impl<VT>(c, A, rowof.p);
}
template void build_csc<float>(adobo_ctx*, SparsePair<float>&);
template void build_csc<double>(adobo_ctx*, SparsePair<double>&);

// ---- per-column statistics on the CSC copy: one CTA per column, deterministic ----------------
template <typename VT>
__global__ void __launch_bounds__(256) col_sum(const i64* __restrict__ cptr, const VT* __restrict__ cval,
                                               double* __restrict__ out) {
  __shared__ double red[8];
  const int col = blockIdx.x;
  double s = 0;
  for (i64 e = cptr[col] + threadIdx.x; e < cptr[col + 1]; e += 256) s += (double)cval[e];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int w = 0; w < 8; ++w) t += red[w];
    out[col] = t;
  }
}

// sum over this rank's cells of (x - mu)^2, implicit zeros included
template <typename VT>
__global__ void __launch_bounds__(256) col_sqdev(const i64* __restrict__ cptr, const VT* __restrict__ cval,
                                                 const double* __restrict__ colsum, double n_total, i64 n_local,
                                                 double* __restrict__ out) {
  __shared__ double red[8];
  const int col = blockIdx.x;
  const double mu = colsum[col] / n_total;
  double s = 0;
  const i64 b = cptr[col], e1 = cptr[col + 1];
  for (i64 e = b + threadIdx.x; e < e1; e += 256) {
    double d = (double)cval[e] - mu;
    s += d * d;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int w = 0; w < 8; ++w) t += red[w];
    out[col] = t + (double)(n_local - (e1 - b)) * mu * mu;
  }
}

__global__ void finish_scale_stats(const double* __restrict__ colsum, const double* __restrict__ sqdev,
                                   double n_total, int h, int scale, double* __restrict__ mu,
                                   double* __restrict__ rs) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= h) return;
  if (!scale) {
    mu[i] = 0.0;
    rs[i] = 1.0;
    return;
  }
  mu[i] = colsum[i] / n_total;
  double sd = sqrt(sqdev[i] / n_total);              // np.nanstd, ddof = 0
  if (sd < 10.0 * 2.220446049250313e-16) sd = 1.0;   // sklearn _handle_zeros_in_scale
  rs[i] = 1.0 / sd;
}

void compute_scale_stats(adobo_ctx* c, SparsePair<float>& A, bool scale, DevBuf<double>& mu, DevBuf<double>& rs) {
  const int h = A.cols;
  mu.ensure((size_t)h);
  rs.ensure((size_t)h);
  DevBuf<double> colsum, sqdev;
  colsum.ensure((size_t)h);
  sqdev.ensure((size_t)h);
  if (scale) {
    LAUNCH(c, "col_sum", (col_sum<float><<<h, 256, 0, c->stream>>>(A.cptr.p, A.cval.p, colsum.p)));
    allreduce_sum_f64(c, colsum.p, (size_t)h);
    LAUNCH(c, "col_sqdev", (col_sqdev<float><<<h, 256, 0, c->stream>>>(A.cptr.p, A.cval.p, colsum.p,
                                                                      (double)c->n_total, A.rows, sqdev.p)));
    allreduce_sum_f64(c, sqdev.p, (size_t)h);
  }
  LAUNCH(c, "finish_scale_stats", (finish_scale_stats<<<(h + 255) / 256, 256, 0, c->stream>>>(
                                      colsum.p, sqdev.p, (double)c->n_total, h, scale ? 1 : 0, mu.p, rs.p)));
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
}

// ---- dr.jackstraw: the subset with some genes permuted across cells (dr.py:606-613) ---------------------------------
// One CTA per gene column of the CSC copy: entry (row r, gene j) gets the sort key (new row << 32 | j); unselected
// genes keep their rows.  A radix sort by key is the permuted matrix in CSR order (columns ascending inside a row:
// the summation order of the mat-vecs stays a pure function of the matrix).
__global__ void __launch_bounds__(256) perm_keys(const i64* __restrict__ cptr, const int* __restrict__ ridx,
                                                 const float* __restrict__ cval, const int* __restrict__ selmap,
                                                 unsigned long long seed, i64 n, unsigned long long* __restrict__ keys,
                                                 float* __restrict__ vals) {
  const int col = blockIdx.x;
  const bool sel = selmap[col] != 0;
  const PermKey k = perm_key(seed, col);
  for (i64 e = cptr[col] + threadIdx.x; e < cptr[col + 1]; e += 256) {
    const i64 r = ridx[e];
    const i64 nr = sel ? perm_apply(k, n, r) : r;
    keys[e] = ((unsigned long long)nr << 32) | (unsigned)col;
    vals[e] = cval[e];
  }
}

// sorted (row << 32 | col) keys -> CSR: column ids, and row pointers from the row boundaries (empty rows included)
__global__ void csr_from_sorted(const unsigned long long* __restrict__ keys, i64 nnz, i64 n, i64* __restrict__ rptr,
                                int* __restrict__ cidx) {
  for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += (i64)gridDim.x * blockDim.x) {
    const unsigned long long k = keys[e];
    const i64 row = (i64)(k >> 32);
    cidx[e] = (int)(k & 0xffffffffull);
    const i64 prev = e == 0 ? -1 : (i64)(keys[e - 1] >> 32);
    for (i64 r = prev + 1; r <= row; ++r) rptr[r] = e;   // rows (prev, row] start at e
    if (e == nnz - 1)
      for (i64 r = row + 1; r <= n; ++r) rptr[r] = nnz;
  }
}

void run_permute_subset(adobo_ctx* c, const int* perm_cols, int n_perm, uint64_t seed, bool need_csc) {
  SparsePair<float>& A = c->sub;
  SparsePair<float>& P = c->perm;
  REQUIRE(c->world == 1, ADOBO_ERR_STATE, "jackstraw permutations run on one GPU (give every GPU its own permutations)");
  REQUIRE(A.rows > 0 && A.cols > 0, ADOBO_ERR_STATE, "jackstraw: no gene subset on the device (adobo_gather_subset / adobo_irlb first)");
  const int h = A.cols;
  std::vector<int> sel(h, 0);
  for (int i = 0; i < n_perm; ++i) {
    REQUIRE(perm_cols[i] >= 0 && perm_cols[i] < h, ADOBO_ERR_ARG, "jackstraw: gene position %d outside 0..%d", perm_cols[i], h - 1);
    sel[perm_cols[i]] = 1;
  }
  DevBuf<int> dsel;
  dsel.ensure((size_t)h);
  CUDA_CHECK(cudaMemcpyAsync(dsel.p, sel.data(), sizeof(int) * h, cudaMemcpyHostToDevice, c->stream));
  P.rows = A.rows;
  P.cols = h;
  P.nnz = A.nnz;
  P.rptr.ensure((size_t)A.rows + 2);
  P.cidx.ensure((size_t)std::max<i64>(A.nnz, 1));
  P.rval.ensure((size_t)std::max<i64>(A.nnz, 1));
  if (A.nnz == 0) {
    CUDA_CHECK(cudaMemsetAsync(P.rptr.p, 0, sizeof(i64) * (A.rows + 1), c->stream));
  } else {
    DevBuf<unsigned long long> k_in, k_out;
    DevBuf<float> v_in;
    k_in.ensure((size_t)A.nnz);
    k_out.ensure((size_t)A.nnz);
    v_in.ensure((size_t)A.nnz);
    LAUNCH(c, "perm_keys", (perm_keys<<<h, 256, 0, c->stream>>>(A.cptr.p, A.ridx.p, A.cval.p, dsel.p, seed, A.rows, k_in.p, v_in.p)));
    int rbits = 1;
    while ((1ll << rbits) < A.rows) ++rbits;
    size_t bytes = 0;
    CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, bytes, k_in.p, k_out.p, v_in.p, P.rval.p, (int)A.nnz, 0, 32 + rbits, c->stream));
    c->scratch.ensure(bytes);
    CUDA_CHECK(cub::DeviceRadixSort::SortPairs(c->scratch.p, bytes, k_in.p, k_out.p, v_in.p, P.rval.p, (int)A.nnz, 0, 32 + rbits, c->stream));
    c->launches += 3;
    LAUNCH(c, "csr_from_sorted", (csr_from_sorted<<<c->sm_count * 4, 256, 0, c->stream>>>(k_out.p, A.nnz, A.rows, P.rptr.p, P.cidx.p)));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));  // temporaries die here
  }
  if (need_csc) build_csc<float>(c, P);
}

void run_gather_subset(adobo_ctx* c, const std::vector<int>& genes_sorted, bool scale) {
  REQUIRE(c->have_norm, ADOBO_ERR_STATE, "pca: no normalized matrix on the device");
  const int h = (int)genes_sorted.size();
  std::vector<int> colmap(c->g, -1);
  for (int i = 0; i < h; ++i) {
    REQUIRE(genes_sorted[i] >= 0 && genes_sorted[i] < c->g, ADOBO_ERR_ARG, "gene index %d out of range", genes_sorted[i]);
    REQUIRE(i == 0 || genes_sorted[i] > genes_sorted[i - 1], ADOBO_ERR_ARG, "duplicate gene index %d", genes_sorted[i]);
    colmap[genes_sorted[i]] = i;
  }
  DevBuf<int> dmap;
  dmap.ensure((size_t)c->g);
  CUDA_CHECK(cudaMemcpyAsync(dmap.p, colmap.data(), sizeof(int) * c->g, cudaMemcpyHostToDevice, c->stream));
  c->sub_genes.ensure((size_t)h);
  CUDA_CHECK(cudaMemcpyAsync(c->sub_genes.p, genes_sorted.data(), sizeof(int) * h, cudaMemcpyHostToDevice, c->stream));
  c->h_sub_genes = genes_sorted;
  SparsePair<float>& A = c->sub;
  A.rows = c->n;
  A.cols = h;
  DevBuf<i64> rowcnt;
  rowcnt.ensure((size_t)c->n + 2);
  CUDA_CHECK(cudaMemsetAsync(rowcnt.p, 0, sizeof(i64) * (c->n + 2), c->stream));
  A.rptr.ensure((size_t)c->n + 2);
  const i64 blocks = std::max<i64>(1, std::min<i64>((c->n + 7) / 8, (i64)c->sm_count * 8));
  if (c->n > 0)
    LAUNCH(c, "subset_count", (subset_count<<<(unsigned)blocks, 256, 0, c->stream>>>(c->indptr.p, c->indices.p, dmap.p,
                                                                                    rowcnt.p, c->n)));
  exclusive_scan_i64(c, rowcnt.p, A.rptr.p, c->n);
  i64 nnz = 0;
  CUDA_CHECK(cudaMemcpyAsync(&nnz, A.rptr.p + c->n, sizeof(i64), cudaMemcpyDeviceToHost, c->stream));
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  A.nnz = nnz;
  A.cidx.ensure((size_t)nnz);
  A.rval.ensure((size_t)nnz);
  DevBuf<int> rowof;
  rowof.ensure((size_t)nnz);
  if (c->n > 0 && nnz > 0)
    LAUNCH(c, "subset_fill", (subset_fill<<<(unsigned)blocks, 256, 0, c->stream>>>(
                                 c->indptr.p, c->indices.p, c->vals.p, dmap.p, A.rptr.p, A.cidx.p, A.rval.p, rowof.p, c->n)));
  build_csc_impl<float>(c, A, rowof.p);
  compute_scale_stats(c, A, scale, c->mu, c->rs);
  c->sub_scaled = scale;
}

}  // namespace adobo
