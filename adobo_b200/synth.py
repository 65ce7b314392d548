"""Seeded synthetic count matrices and embeddings of the benchmark shapes.

The generator is the single source of inputs for the parity tests, the golden vectors and
bench.py: the same arrays feed the CUDA path and its CPU checkers.  Cells are drawn
in blocks of 1,024 with a per-block stream so that any contiguous cell shard of a matrix can
be regenerated on its own rank without generating the rest.

Model (negative-binomial-like, with cluster structure so that PCA / kNN have something to
find): gene base log-mean ~ N(off, 2.0); 8 latent cell clusters with log-fold-changes N(0, 1)
on 5 % of genes; per-cell size factor exp(N(0, 0.35)); count ~ Poisson(lambda * Gamma(2, 1/2)).
`off` is solved so that the expected fraction of non-zeros equals `density`.  Every cell is
forced to hold at least one count (library size 0 would divide by zero in
normalize.standard, adobo/normalize.py:285-286).
"""
import numpy as np
import scipy.sparse as sp

BLOCK = 1024


class CountModel:
    def __init__(self, n_genes, density, seed, n_clusters=8):
        rng = np.random.default_rng([seed, 0xADB0])
        self.n_genes = int(n_genes)
        self.n_clusters = n_clusters
        self.seed = seed
        self.base = rng.normal(0.0, 2.0, n_genes)
        mask = rng.random((n_clusters, n_genes)) < 0.05
        self.lfc = rng.normal(0.0, 1.0, (n_clusters, n_genes)) * mask
        # P(count == 0 | lambda) for Poisson(lambda * Gamma(2, 1/2)) is (1 + lambda/2)^-2
        lo, hi = -30.0, 10.0
        for _ in range(60):
            mid = 0.5 * (lo + hi)
            lam = np.exp(self.base[None, :] + self.lfc + mid)
            d = 1.0 - np.mean((1.0 + 0.5 * lam) ** -2)
            if d < density:
                lo = mid
            else:
                hi = mid
        self.off = 0.5 * (lo + hi)

    def lam_table(self):
        """Expected count per (cluster, gene) before the per-cell size factor: the table the device generator
        (csrc/synth.cu, Engine.synth_counts) draws from."""
        return np.exp(self.base[None, :] + self.off + self.lfc).astype(np.float32)

    def block(self, b, n_rows):
        """Dense int32 counts of cell block `b` (n_rows <= BLOCK rows)."""
        rng = np.random.default_rng([self.seed, 1 + b])
        cl = rng.integers(0, self.n_clusters, BLOCK)
        size = np.exp(rng.normal(0.0, 0.35, BLOCK))
        lam = np.exp(self.base[None, :] + self.off + self.lfc[cl]) * size[:, None]
        lam *= rng.gamma(2.0, 0.5, lam.shape)
        c = rng.poisson(lam).astype(np.int32)
        empty = np.flatnonzero(c.sum(axis=1) == 0)
        c[empty, (empty * 7919) % self.n_genes] = 1
        return c[:n_rows]


def synth_counts(n_cells, n_genes, density=0.064, seed=43, row_start=0, row_stop=None):
    """CSR (cells x genes) int32 count matrix; rows [row_start, row_stop) of the full matrix.

    indices int32, data int32 (scipy picks the indptr width; Engine widens it to int64 for the C-ABI).
    """
    if row_stop is None:
        row_stop = n_cells
    model = CountModel(n_genes, density, seed)
    parts = []
    b0, b1 = row_start // BLOCK, (row_stop + BLOCK - 1) // BLOCK
    for b in range(b0, b1):
        lo, hi = b * BLOCK, min((b + 1) * BLOCK, n_cells)
        blk = model.block(b, hi - lo)
        s, e = max(lo, row_start) - lo, min(hi, row_stop) - lo
        parts.append(sp.csr_matrix(blk[s:e]))
    m = sp.vstack(parts, format='csr') if parts else sp.csr_matrix((0, n_genes), dtype=np.int32)
    m.sort_indices()
    return sp.csr_matrix((m.data.astype(np.int32), m.indices.astype(np.int32),
                          m.indptr.astype(np.int64)), shape=(row_stop - row_start, n_genes))


def synth_embedding(n, p=50, n_blobs=16, seed=5):
    """C5 input: Gaussian blobs + unit noise, float64 (n x p)."""
    rng = np.random.default_rng([seed, 0xE3B])
    centers = rng.normal(0.0, 4.0, (n_blobs, p))
    out = np.empty((n, p))
    for lo in range(0, n, 1 << 16):
        hi = min(n, lo + (1 << 16))
        r = np.random.default_rng([seed, 1 + lo // (1 << 16)])
        lab = r.integers(0, n_blobs, hi - lo)
        out[lo:hi] = centers[lab] + r.normal(0.0, 1.0, (hi - lo, p))
    return out


def counts_to_frame(csr, prefix_gene='g', prefix_cell='c'):
    """genes x cells dense int64 DataFrame, the input adobo.data.dataset expects (data.py:54-67)."""
    import pandas as pd
    dense = np.asarray(csr.T.todense()).astype(np.int64)
    return pd.DataFrame(dense, index=['%s%d' % (prefix_gene, i) for i in range(csr.shape[1])],
                        columns=['%s%d' % (prefix_cell, i) for i in range(csr.shape[0])])
