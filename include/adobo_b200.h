/*
 * adobo_b200 -- C-ABI of the B200 (sm_100a) implementation of adobo's cell-embedding hot path:
 *
 *     normalize.norm('standard') -> hvg.find_hvg('seurat') -> dr.pca('irlb') -> clustering knn/snn
 *
 * The reference (oscar-franzen/adobo v0.2.70) is pure Python; its only FFI precedent is the
 * ctypes binding of adobo/libs/pdf.c (adobo/preproc.py:446-459): plain pointers + lengths,
 * caller-allocated outputs.  This header keeps that idiom.  Every entry point names the
 * reference function it replaces (paths relative to the reference checkout).
 *
 * Conventions
 *   - all pointers are HOST pointers unless the name ends in _dev; the caller owns them;
 *   - the count matrix is CSR, cells x genes (a pandas sparse genes x cells frame is physically
 *     this layout, adobo/data.py:67): indptr int64[n+1], indices int32[nnz] ascending inside a
 *     row, counts int32[nnz];
 *   - functions return 0 on success, <0 on error; adobo_last_error() gives the message of the
 *     last failure on the calling thread;
 *   - one context per GPU, one host thread per context; under torchrun each rank owns the
 *     contiguous block of cells it uploaded and the contexts are joined with adobo_comm_init;
 *   - there is no CPU fallback: every call fails with ADOBO_ERR_CUDA when no sm_100 device
 *     is usable.
 */
#ifndef ADOBO_B200_H
#define ADOBO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ADOBO_OK 0
#define ADOBO_ERR_ARG (-1)
#define ADOBO_ERR_CUDA (-2)
#define ADOBO_ERR_STATE (-3)
#define ADOBO_ERR_NCCL (-4)
#define ADOBO_ERR_LIMIT (-5)

#define ADOBO_LOG_NONE 0  /* norm(log=False) */
#define ADOBO_LOG_2 1     /* log_func=np.log2  (reference default, normalize.py:437) */
#define ADOBO_LOG_E 2     /* np.log   */
#define ADOBO_LOG_10 3    /* np.log10 */

typedef struct adobo_ctx adobo_ctx;

int adobo_version(void);
const char* adobo_last_error(void);

/* ---- context / device plumbing ------------------------------------------------------------ */
int adobo_ctx_create(int device, adobo_ctx** out);
void adobo_ctx_destroy(adobo_ctx* ctx);
int adobo_sync(adobo_ctx* ctx);
/* NCCL over NVLink: rank r of `world` holds cells [row_offset, row_offset + n) of n_total.
 * id128 is the ncclUniqueId made by rank 0 with adobo_comm_unique_id and broadcast by the host
 * (torch.distributed). */
int adobo_comm_unique_id(void* id128);
int adobo_comm_init(adobo_ctx* ctx, int world, int rank, const void* id128);

/* CUDA-event stopwatch on the context's stream (bench.py times on the launching stream). */
int adobo_event_record(adobo_ctx* ctx, int slot);                   /* slot 0..63 */
int adobo_event_elapsed_ms(adobo_ctx* ctx, int slot_a, int slot_b, float* ms);
/* Accumulated device time (ms) and launch count per kernel class since the last reset; only
 * collected while profiling is on (adds one event pair per launch). */
int adobo_profile(adobo_ctx* ctx, int on);
int adobo_profile_read(adobo_ctx* ctx, int max_items, char* names /* max_items x 32 */, float* ms,
                       int64_t* launches, double* work /* algorithmic bytes (flops: knn_shortlist) */,
                       int* n_items);
int64_t adobo_launch_count(adobo_ctx* ctx);                         /* kernels launched so far */
int adobo_l2_flush(adobo_ctx* ctx);                                 /* writes a 256 MB buffer */
/* Page-lock / unlock a caller buffer so that the copies of the host-facing calls are true DMA. */
int adobo_host_register(void* ptr, size_t bytes);
int adobo_host_unregister(void* ptr);

/* ---- data in ------------------------------------------------------------------------------- */
/* Replaces: the count_data frame handed to normalize.norm (adobo/normalize.py:540) after
 * clean_matrix (normalize.py:413-433) -- the caller applies the cell / gene masks. */
int adobo_upload_counts(adobo_ctx* ctx, int64_t n_cells, int32_t n_genes, const int64_t* indptr,
                        const int32_t* indices, const int32_t* counts);
/* Same, from an already normalized matrix (hvg.find_hvg / dr.pca called on an existing
 * obj.norm_data[...]['data'], hvg.py:471, dr.py:308). values: float64[nnz]. */
int adobo_upload_normalized(adobo_ctx* ctx, int64_t n_cells, int32_t n_genes, const int64_t* indptr,
                            const int32_t* indices, const double* values);

/* Test / bench data source (no counterpart in the reference): seeded synthetic counts generated on the device,
 * model of SURVEY.md 8(d) -- count ~ Poisson(lam_table[cluster(cell)][gene] * size(cell) * Gamma(2, 1/2)), every draw a
 * pure function of (seed, global cell index, gene).  Generates cells [row_begin, row_begin + n_rows) of the
 * conceptual matrix and leaves them in the context exactly as adobo_upload_counts would (any rank can generate its
 * own shard).  lam_table: float32 [n_clusters x n_genes] (host).  adobo_download_counts copies the resident count
 * matrix back (indptr int64[n+1], indices / counts int32[nnz]; any may be NULL) for the CPU checkers. */
int adobo_synth_counts(adobo_ctx* ctx, int64_t n_rows, int32_t n_genes, int64_t row_begin, uint64_t seed,
                       const float* lam_table, int32_t n_clusters, int64_t* out_nnz);
int adobo_download_counts(adobo_ctx* ctx, int64_t* indptr, int32_t* indices, int32_t* counts);

/* The summary statistics adobo.data.dataset computes when it is constructed (adobo/data.py:75-84), from the resident
 * count matrix: total_reads[n] = raw_mat.sum(axis=0), detected_genes[n] = (raw_mat > 0).sum(axis=0) per cell (this
 * rank's cells), expressed[g] = number of cells with a positive count per gene (summed over all ranks).  Integers:
 * exact and order independent.  Any output may be NULL. */
int adobo_count_stats(adobo_ctx* ctx, int64_t* out_total_reads, int64_t* out_detected_genes, int64_t* out_expressed);

/* ---- normalize.standard + log step (adobo/normalize.py:285-287, 559-560) ------------------- */
/* y = log_f((c / S_cell) * scaling_factor + small_const) on the non-zeros; zeros stay zero.
 * out_libsize int64[n] and out_values float64[nnz] may be NULL (device-resident use). */
int adobo_normalize(adobo_ctx* ctx, double scaling_factor, int log_mode, double small_const,
                    int64_t* out_libsize, double* out_values);

/* ---- hvg.seurat (adobo/hvg.py:31-72) ------------------------------------------------------- */
/* Per-gene mean over all cells and sample variance (ddof=1) of the normalized values
 * (hvg.py:60, 65).  Multi-GPU: sums are all-reduced, every rank gets the same vectors. */
int adobo_gene_moments(adobo_ctx* ctx, double* out_mean, double* out_var);
/* Equal-width mean bins, per-bin z-score of var/mean, top `ngenes` by z (hvg.py:62-71).
 * out_idx int32[ngenes] in the reference's output order; *out_n = number returned
 * (min(ngenes, genes in any bin)); out_z float64[n_genes] may be NULL. */
int adobo_hvg_seurat(adobo_ctx* ctx, int32_t ngenes, int32_t num_bins, int32_t* out_idx,
                     int32_t* out_n, double* out_z);

/* ---- dr.irlb / irlbpy.lanczos (adobo/dr.py:110-172, adobo/irlbpy/irlb.py:95-258) ----------- */
/* Truncated SVD of the (cells x h) column subset `gene_idx` (any order; used in ascending gene
 * order like dr.py:319) of the normalized matrix, centred and scaled per gene as
 * sklearn.preprocessing.scale does when scale != 0 (dr.py:156-160), by augmented implicitly
 * restarted Lanczos bidiagonalization.  v0: float64[h] start vector BEFORE normalization
 * (np.random.seed(seed); np.random.randn(h), irlb.py:151-152), in ascending gene order.
 * Outputs (may be NULL): out_comp float64[n x ncomp] row-major = U diag(s) if var_weigh else U
 * (dr.py:164-168), this rank's cells; out_s float64[ncomp]; out_contr float64[h x ncomp]
 * row-major = V (dr.py:171); out_steps / out_nmult = LanczosResult.steps / .nmult. */
int adobo_irlb(adobo_ctx* ctx, const int32_t* gene_idx, int32_t h, int32_t ncomp, double tol,
               int32_t maxit, const double* v0, int32_t scale, int32_t var_weigh, double* out_comp,
               double* out_s, double* out_contr, int32_t* out_steps, int32_t* out_nmult);
/* irlbpy.lanczos(A, nval, tol, maxit) on a caller-supplied sparse matrix (rows x cols CSR,
 * float64 values, no centring/scaling: center=None, scale=None as adobo calls it).
 * out_U float64[rows x nval], out_V float64[cols x nval] row-major. */
int adobo_lanczos_csr(adobo_ctx* ctx, int64_t rows, int32_t cols, const int64_t* indptr,
                      const int32_t* indices, const double* values, int32_t nval, double tol,
                      int32_t maxit, const double* v0, double* out_U, double* out_s, double* out_V,
                      int32_t* out_steps, int32_t* out_nmult);

/* ---- dr.jackstraw (adobo/dr.py:604-620): the 500 x irlb() multiplier of the hot path ------------------------ */
/* adobo_gather_subset leaves the HVG column subset (ascending gene order) and its centring / scaling statistics on
 * the device without running the PCA (what adobo_irlb does first).  adobo_irlb_permuted then runs ONE jackstraw
 * permutation entirely on the device: the genes at positions perm_cols[0..n_perm) of the subset are permuted across
 * cells -- gene j's entry of cell r moves to cell adobo_perm_row(n_cells, perm_seed, j, r), a keyed bijection of
 * [0, n_cells) -- the permuted CSR is rebuilt by a key sort, and IRLB runs on it with the UNPERMUTED statistics (a
 * permutation keeps a gene's mean and sigma, so this equals dr.py:606-617: permute rows of the scaled matrix, then
 * irlb(scale=False)).  out_contr float64[h x ncomp] = V.  Single GPU (several GPUs take different permutations). */
int adobo_gather_subset(adobo_ctx* ctx, const int32_t* gene_idx, int32_t h, int32_t scale);
int adobo_irlb_permuted(adobo_ctx* ctx, const int32_t* perm_cols, int32_t n_perm, uint64_t perm_seed, int32_t ncomp,
                        double tol, int32_t maxit, const double* v0, int32_t var_weigh, double* out_contr,
                        int32_t* out_steps, int32_t* out_nmult);
/* The bijection itself (host function, same arithmetic as the device): where cell `row`'s entry of gene `col` goes. */
int64_t adobo_perm_row(int64_t n_cells, uint64_t perm_seed, int32_t col, int64_t row);

/* ---- clustering.knn (adobo/clustering.py:25-52) -------------------------------------------- */
/* Exact Euclidean k nearest neighbours of every row of comp among all rows (self included),
 * ascending by fp64 distance, 1-BASED like the reference (clustering.py:51).  comp float64
 * [n x p] row-major, or NULL to use the embedding left on the device by adobo_irlb.
 * out_idx int64[n x k] (this rank's cells; indices are global). */
int adobo_knn(adobo_ctx* ctx, const double* comp, int64_t n, int32_t p, int32_t k, int64_t* out_idx);

/* The other metrics clustering.knn(distance=...) forwards to sklearn's ball tree (clustering.py:47-48): the same call
 * with a metric code.  ADOBO_METRIC_EUCLIDEAN is adobo_knn itself (tensor-core shortlist + fp64 re-rank); the others
 * are an exact fp64 brute force on the CUDA cores that accumulates the coordinates in index order like sklearn's
 * DistanceMetric (same distances, bit for bit), ordered by (distance, index) with the query itself first.
 * ADOBO_METRIC_EUCLIDEAN_FP64 is that brute force with squared differences (a cross-check of the tensor path). */
#define ADOBO_METRIC_EUCLIDEAN 0      /* 'euclidean', 'l2', 'minkowski' (p = 2, the ball tree's default) */
#define ADOBO_METRIC_MANHATTAN 1      /* 'manhattan', 'cityblock', 'l1' */
#define ADOBO_METRIC_CHEBYSHEV 2      /* 'chebyshev', 'infinity' */
#define ADOBO_METRIC_HAMMING 3        /* 'hamming' */
#define ADOBO_METRIC_CANBERRA 4       /* 'canberra' */
#define ADOBO_METRIC_BRAYCURTIS 5     /* 'braycurtis' */
#define ADOBO_METRIC_EUCLIDEAN_FP64 6
int adobo_knn_metric(adobo_ctx* ctx, const double* comp, int64_t n, int32_t p, int32_t k, int32_t metric,
                     int64_t* out_idx);

/* Leaves a caller-supplied embedding (this rank's cells, float64 [n x p] row-major) on the device as if adobo_irlb had
 * produced it: adobo_knn(ctx, NULL, ...) then runs on it (obj.norm_data[..]['dr']['pca']['comp'] of an earlier
 * session, clustering.py:337-341; the kNN / SNN sweep of bench.py). */
int adobo_upload_embedding(adobo_ctx* ctx, const double* comp, int64_t n, int32_t p);

/* ---- clustering.snn (adobo/clustering.py:55-128) ------------------------------------------- */
/* Shared-nearest-neighbour graph: M has ones at (nn_idx[i,0]-1, nn_idx[i,c]-1), c>=1, and on the
 * diagonal; V = M M^T; edge (i,j) kept iff V_ij / (k + (k - V_ij)) > prune_snn.  nn_idx int64
 * [n x k] 1-based, or NULL to use the result of adobo_knn left on the device.  Returns counts;
 * the edge list (0-based, sorted by source then target, nodes without any kept edge appended as
 * self loops, clustering.py:114-118) is then read with adobo_snn_fetch. */
int adobo_snn(adobo_ctx* ctx, const int64_t* nn_idx, int64_t n, int32_t k, double prune_snn,
              int64_t* out_n_edges, int64_t* out_pruned, int64_t* out_total_links);
int adobo_snn_fetch(adobo_ctx* ctx, int32_t* out_source, int32_t* out_target);
/* The integer edge weights the reference computes and then drops (clustering.py:99-112 keeps only the
 * coordinates): out_v[e] = V_ij = (M M^T)_ij of edge e of adobo_snn_fetch, 1..k; 0 for the self loops appended for
 * nodes without a kept edge. */
int adobo_snn_fetch_counts(adobo_ctx* ctx, int32_t* out_v);
/* Sizes of what the context holds, so that callers can size the buffers of the fetch calls:
 * out8 = { local cells, PCA components, PCA genes h, k of the kNN result, rows of the kNN result, SNN edges,
 *          non-zeros of the matrix shard, genes }. */
int adobo_result_sizes(adobo_ctx* ctx, int64_t* out8);

/* ---- whole path, device resident (bench `value`) ------------------------------------------- */
/* normalize -> moments -> seurat -> irlb -> knn -> snn without leaving the device.  v0 has
 * length ngenes.  Results stay in the context (fetch with the out_* == NULL-tolerant calls
 * above or adobo_fetch_*). */
int adobo_embed_path(adobo_ctx* ctx, double scaling_factor, int32_t ngenes, int32_t ncomp,
                     const double* v0, int32_t k, double prune_snn, int32_t* out_steps,
                     int32_t* out_nmult, int64_t* out_n_edges);
int adobo_fetch_pca(adobo_ctx* ctx, double* out_comp, double* out_s, double* out_contr,
                    int32_t* out_genes);
int adobo_fetch_knn(adobo_ctx* ctx, int64_t* out_idx);

#ifdef __cplusplus
}
#endif
#endif
