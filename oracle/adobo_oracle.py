"""TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of adobo's cell-embedding hot path.

    normalize.norm('standard') -> hvg.seurat -> dr.irlb / irlbpy.lanczos -> clustering.knn / snn

Each function restates the arithmetic of the cited reference lines (paths relative to the
reference checkout, oscar-franzen/adobo v0.2.70) on scipy.sparse CSR (cells x genes) instead of
pandas sparse frames, so that it also scales to shapes pandas cannot hold.  It is the CHECKER for
the CUDA path: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import it; nothing under adobo_b200/ does.

Pinning: the reference ships no tests and no golden vectors, and the notebook
inputs are missing from the checkout.  The oracle is therefore pinned against OUTPUTS OF THE
REFERENCE ITSELF: tests/golden/make_golden.py runs the unmodified reference (oracle/ref_loader.py)
on seeded synthetic matrices and freezes every stage into tests/golden/*.npz;
tests/test_oracle_golden.py checks this file against those vectors, and
tests/test_oracle_vs_reference.py against the live reference when /root/reference is present.

Third-party arithmetic the reference delegates to (not in its tree; the reference pins none of
them, setup.py:17-36) is restated from the installed versions: pandas 3.0.2 (`mean`, `var`,
`cut`, `sort_values`), scikit-learn 1.9.0 (`preprocessing.scale`,
`NearestNeighbors(algorithm='ball_tree')`), scipy 1.18.1 (sparse product), numpy 2.3.5
(`linalg.svd`, legacy `random.seed/randn`).
"""
import numpy as np
import scipy.sparse as sp

# --------------------------------------------------------------------------------------------
# normalize.standard + log step
# --------------------------------------------------------------------------------------------


def normalize_standard(counts, scaling_factor=10000, log=True, log_func=np.log2, small_const=1):
    """adobo/normalize.py:285-286 (`col_sums`, divide THEN multiply) and :559-560 (log step).

    counts: CSR cells x genes, integer.  Returns (CSR float64 same pattern, library sizes int64).
    Zeros stay zero: the reference logs log_func(0 + small_const) on the fill value and the
    frame is re-sparsified with fill 0 (normalize.py:567-568); for small_const == 1 that is 0.
    """
    counts = sp.csr_matrix(counts)
    lib = np.asarray(counts.sum(axis=1)).ravel().astype(np.int64)
    rows = np.repeat(np.arange(counts.shape[0]), np.diff(counts.indptr))
    x = (counts.data.astype(np.float64) / lib[rows].astype(np.float64)) * scaling_factor
    if log:
        x = log_func(x + small_const)
    out = sp.csr_matrix((x, counts.indices.copy(), counts.indptr.copy()), shape=counts.shape)
    return out, lib


# --------------------------------------------------------------------------------------------
# hvg.seurat
# --------------------------------------------------------------------------------------------


def gene_moments(norm):
    """Per-gene mean over ALL cells (hvg.py:60) and sample variance ddof=1 (hvg.py:65).

    pandas' nanvar is two-pass: sum((mean - x)^2) / (count - 1); the implicit zeros contribute
    (n - nnz_g) * mean^2.
    """
    norm = sp.csc_matrix(norm)
    n = norm.shape[0]
    s1 = np.asarray(norm.sum(axis=0)).ravel()
    mean = s1 / n
    cols = np.repeat(np.arange(norm.shape[1]), np.diff(norm.indptr))
    dev = norm.data - mean[cols]
    ss = np.bincount(cols, weights=dev * dev, minlength=norm.shape[1])
    nz = np.diff(norm.indptr)
    ss = ss + (n - nz) * mean * mean
    with np.errstate(invalid='ignore', divide='ignore'):
        var = ss / (n - 1)
    return mean, var


def cut_equal_width(x, nbins):
    """pandas.cut(x, nbins) (reshape/tile.py `_nbins_to_bins` + `_bins_to_cuts`), right-closed:
    edges = linspace(min, max, nbins+1); edges[0] -= 0.1 % of the range;
    bin = searchsorted(edges, x, 'left') - 1; values outside (edge0, edgeN] -> -1."""
    mn, mx = np.nanmin(x), np.nanmax(x)
    if mn == mx:
        mn -= 0.001 * abs(mn) if mn != 0 else 0.001
        mx += 0.001 * abs(mx) if mx != 0 else 0.001
        edges = np.linspace(mn, mx, nbins + 1, endpoint=True)
    else:
        edges = np.linspace(mn, mx, nbins + 1, endpoint=True)
        edges[0] -= (mx - mn) * 0.001
    ids = edges.searchsorted(x, side='left')
    bad = np.isnan(x) | (ids == len(edges)) | (ids == 0)
    b = ids - 1
    b[bad] = -1
    return b, edges


def seurat_zscores(mean, var, num_bins=20):
    """hvg.py:62-66: dispersion = var/mean, z-scored inside each of the equal-width mean bins
    with NaN-skipping mean and std(ddof=1)."""
    bins, _ = cut_equal_width(mean, num_bins)
    with np.errstate(invalid='ignore', divide='ignore'):
        disp = var / mean
    z = np.full(mean.shape[0], np.nan)
    for b in range(num_bins):
        idx = np.flatnonzero(bins == b)
        if idx.size == 0:
            continue
        d = disp[idx]
        ok = ~np.isnan(d)
        cnt = int(ok.sum())
        with np.errstate(invalid='ignore', divide='ignore'):
            m = d[ok].sum() / cnt if cnt > 0 else np.nan
            if cnt > 1:
                sd = np.sqrt(((d[ok] - m) ** 2).sum() / (cnt - 1))
            else:
                sd = np.nan
            z[idx] = (d - m) / sd
    return z, bins


def seurat_order(z, bins, ngenes, num_bins=20):
    """hvg.py:68-71: concat bins in order, sort z descending (NaN last), head(ngenes).
    Returns gene indices in the reference's output order.  (pandas sorts descending as the
    reverse of an ascending sort of the reversed array; ties are unordered in the reference.)"""
    order = np.concatenate([np.flatnonzero(bins == b) for b in range(num_bins)] or [np.zeros(0, int)])
    zz = z[order]
    good = np.flatnonzero(~np.isnan(zz))
    rev = good[::-1]
    asc = rev[np.argsort(zz[rev], kind='quicksort')]
    desc = asc[::-1]
    full = np.concatenate([desc, np.flatnonzero(np.isnan(zz))])
    return order[full[:ngenes]]


def seurat(norm, ngenes=1000, num_bins=20):
    """hvg.seurat (hvg.py:31-72) on CSR cells x genes.  Returns ordered gene indices."""
    mean, var = gene_moments(norm)
    z, bins = seurat_zscores(mean, var, num_bins)
    return seurat_order(z, bins, ngenes, num_bins)


# --------------------------------------------------------------------------------------------
# dr.irlb: sklearn.preprocessing.scale + irlbpy.lanczos
# --------------------------------------------------------------------------------------------


def scale_stats(sub):
    """Column mean and population std of the (cells x h) subset as sklearn.preprocessing.scale
    computes them (dr.py:156-160; sklearn/preprocessing/_data.py `scale`): nanmean, nanstd
    (ddof=0), std < 10*eps -> 1."""
    dense = np.asarray(sp.csr_matrix(sub).todense(), dtype=np.float64)
    mu = dense.mean(axis=0)
    sd = dense.std(axis=0)
    sd[sd < 10 * np.finfo(np.float64).eps] = 1.0
    return mu, sd


def scale_dense(sub):
    """The dense matrix dr.irlb hands to lanczos (dr.py:156-161)."""
    dense = np.array(sp.csr_matrix(sub).todense(), dtype=np.float64)
    mu = dense.mean(axis=0)
    sd = dense.std(axis=0)
    dense -= mu
    m1 = dense.mean(axis=0)
    if not np.allclose(m1, 0):
        dense -= m1
    sd[sd < 10 * np.finfo(np.float64).eps] = 1.0
    dense /= sd
    m2 = dense.mean(axis=0)
    if not np.allclose(m2, 0):
        dense -= m2
    return dense


class DenseOp:
    """A as an explicit array (what the reference does)."""

    def __init__(self, a):
        self.a = np.asarray(a, dtype=np.float64)
        self.shape = self.a.shape

    def mv(self, v):
        return self.a.dot(v)          # multA(A, x)        irlb.py:34

    def rmv(self, u):
        return u.dot(self.a)          # multA(A, x, TP)    irlb.py:33


class ScaledCenteredOp:
    """A_s = (X - 1 mu^T) diag(1/sigma) applied implicitly to a sparse X -- the form the CUDA
    path uses.  A_s v = X (v/sigma) - (mu . (v/sigma)) 1 ;  A_s^T u = (X^T u - mu (1^T u)) / sigma.
    `allreduce` sums an array over ranks when X holds only this rank's cells."""

    def __init__(self, x, mu, sigma, allreduce=None):
        self.x = sp.csr_matrix(x, dtype=np.float64)
        self.xt = self.x.T.tocsr()
        self.mu, self.sigma = mu, sigma
        self.shape = self.x.shape
        self.allreduce = allreduce or (lambda a: a)

    def mv(self, v):
        vs = v / self.sigma
        return self.x.dot(vs) - self.mu.dot(vs)

    def rmv(self, u):
        t = np.concatenate([self.xt.dot(u), [u.sum()]])
        t = self.allreduce(t)
        return (t[:-1] - self.mu * t[-1]) / self.sigma


def _inv_or_zero(x):
    """invcheck, irlb.py:84-92."""
    return 1.0 / x if x > 2 * np.finfo(np.float64).eps else 0.0


def _cgs(y, basis, allreduce):
    """orthog, irlb.py:73-79: one pass of classical Gram-Schmidt."""
    c = allreduce(y.dot(basis))
    return y - basis.dot(c)


def lanczos(op, nval, tol=0.0001, maxit=1000, seed=None, v0=None, allreduce=None, trace=None):
    """irlbpy.lanczos (irlb.py:95-258) for a 2-D operator, center=None, scale=None as dr.irlb
    calls it (dr.py:163).  Rows of W (and of the operator) may be sharded over ranks: then
    `allreduce` must sum over ranks; V, B and every scalar are replicated.

    Returns dict U, s, V, steps, nmult.
    """
    ar = allreduce or (lambda a: a)
    m, n = op.shape
    nu = nval
    m_b = min(nu + 20, 3 * nu, n)                         # irlb.py:138
    V = np.zeros((n, m_b))
    W = np.zeros((m, m_b))
    B = np.zeros((m_b, m_b))
    if v0 is None:
        np.random.seed(seed)                              # irlb.py:151
        v0 = np.random.randn(n)
    V[:, 0] = v0 / np.linalg.norm(v0)                     # irlb.py:153 (V is zero elsewhere)
    nmult, it, k, smax = 0, 0, nu, 1.0
    norm = lambda w: np.sqrt(ar(np.array([w.dot(w)]))[0])
    while it < maxit:
        j = k if it > 0 else 0                            # irlb.py:156-157
        w = op.mv(V[:, j]); nmult += 1                    # irlb.py:165
        if it > 0:
            w = _cgs(w, W[:, :j], ar)                     # irlb.py:173-175
        s = norm(w)                                       # irlb.py:176
        W[:, j] = _inv_or_zero(s) * w
        while j < m_b:                                    # irlb.py:181
            F = op.rmv(W[:, j]); nmult += 1               # irlb.py:182
            F = F - s * V[:, j]                           # irlb.py:189
            F = F - V[:, :j + 1].dot(F.dot(V[:, :j + 1]))  # irlb.py:190 (V replicated)
            fn = np.linalg.norm(F)                        # irlb.py:191
            F = _inv_or_zero(fn) * F
            if j < m_b - 1:
                V[:, j + 1] = F                           # irlb.py:195-197
                B[j, j] = s
                B[j, j + 1] = fn
                w = op.mv(F); nmult += 1                  # irlb.py:204
                w = w - fn * W[:, j]                      # irlb.py:214
                w = _cgs(w, W[:, :j + 1], ar)             # irlb.py:216
                s = norm(w)                               # irlb.py:217
                W[:, j + 1] = _inv_or_zero(s) * w
            else:
                B[j, j] = s                               # irlb.py:221
            j += 1
        Su, sv, Svt = np.linalg.svd(B)                    # irlb.py:224
        R = fn * Su[m_b - 1, :]                           # irlb.py:225
        smax = sv[0] if it == 0 else max(sv[0], smax)     # irlb.py:226-229
        conv = int(np.sum(np.abs(R[:nu]) < tol * smax))   # irlb.py:231
        if trace is not None:
            trace.append(dict(it=it, k=k, conv=conv, smax=smax, sv=sv.copy()))
        if conv >= nu:
            break
        k = min(max(conv + nu, k), m_b - 3)               # irlb.py:233-234
        V[:, :k] = V.dot(Svt.T[:, :k])                    # irlb.py:238
        V[:, k] = F                                       # irlb.py:239
        B = np.zeros((m_b, m_b))
        B[np.arange(k), np.arange(k)] = sv[:k]            # irlb.py:242-243
        B[:k, k] = R[:k]                                  # irlb.py:244
        W[:, :k] = W.dot(Su[:, :k])                       # irlb.py:246
        it += 1
    U = W.dot(Su[:, :nu])                                 # irlb.py:249
    Vout = V.dot(Svt.T[:, :nu])                           # irlb.py:250
    return dict(U=U, s=sv[:nu].copy(), V=Vout, steps=it, nmult=nmult)


def irlb_pca(norm, hvg_idx, ncomp=75, scale=True, var_weigh=True, seed=42, implicit=True,
             tol=0.0001, maxit=1000):
    """dr.pca(method='irlb') numerics (dr.py:319, 110-172): HVG column subset IN ORIGINAL GENE
    ORDER, sklearn scale, lanczos, comp = U diag(s), contr = V.

    Returns dict comp (n x p), contr (h x p), s, steps, nmult, genes (sorted subset indices).
    """
    genes = np.sort(np.asarray(hvg_idx))
    sub = sp.csr_matrix(norm)[:, genes]
    if scale:
        if implicit:
            mu, sd = scale_stats(sub)
            op = ScaledCenteredOp(sub, mu, sd)
        else:
            op = DenseOp(scale_dense(sub))
    else:
        op = DenseOp(np.asarray(sub.todense()))
    r = lanczos(op, ncomp, tol=tol, maxit=maxit, seed=seed)
    comp = r['U'] * r['s'][None, :] if var_weigh else r['U']      # dr.py:164-168
    return dict(comp=comp, contr=r['V'], s=r['s'], steps=r['steps'], nmult=r['nmult'], genes=genes)


# --------------------------------------------------------------------------------------------
# clustering.knn / clustering.snn
# --------------------------------------------------------------------------------------------


def knn(comp, k=10, block=256, return_dist=False):
    """clustering.knn (clustering.py:47-52): exact Euclidean k-NN of every row among all rows,
    ascending by distance, 1-based.  Distances are the direct sum of squared differences in
    fp64 (what sklearn's ball-tree metric evaluates); candidates come from an fp64 Gram
    shortlist of k+16 and are re-ranked with the direct form.  Ties are broken by index -- the
    reference's tie order is unspecified (documented tie = equal fp64 squared distance)."""
    x = np.ascontiguousarray(np.asarray(comp, dtype=np.float64))
    n = x.shape[0]
    kk = min(n, k + 16)
    sq = np.einsum('ij,ij->i', x, x)
    idx = np.empty((n, k), dtype=np.int64)
    dist = np.empty((n, k))
    for lo in range(0, n, block):
        hi = min(n, lo + block)
        g = sq[lo:hi, None] + sq[None, :] - 2.0 * x[lo:hi].dot(x.T)
        cand = np.argpartition(g, kk - 1, axis=1)[:, :kk] if kk < n else np.tile(np.arange(n), (hi - lo, 1))
        diff = x[lo:hi, None, :] - x[cand]
        d2 = np.einsum('ijk,ijk->ij', diff, diff)
        order = np.lexsort((cand, d2), axis=1)[:, :k]
        idx[lo:hi] = np.take_along_axis(cand, order, axis=1)
        dist[lo:hi] = np.take_along_axis(d2, order, axis=1)
    if return_dist:
        return idx + 1, dist
    return idx + 1


def knn_metric(comp, k=10, distance='manhattan'):
    """clustering.knn with another ball-tree metric (clustering.py:47-48 forwards `distance` to
    sklearn.neighbors.NearestNeighbors(algorithm='ball_tree', metric=distance)).  Restates the parameter-free metrics
    of scikit-learn 1.9.0's DistanceMetric (sklearn/metrics/_dist_metrics.pyx.tp: ManhattanDistance sum |x - y|,
    ChebyshevDistance max |x - y|, HammingDistance mean(x != y), CanberraDistance sum |x - y| / (|x| + |y|) skipping
    0 / 0, BrayCurtisDistance sum |x - y| / sum (|x| + |y|) -- not scipy's |x + y|), coordinates accumulated in index order, every pair, ties
    by index with the point itself first.  1-based like the reference.  Pinned against the ball tree itself by
    tests/test_oracle_golden.py::test_knn_metric_is_the_ball_tree."""
    x = np.ascontiguousarray(np.asarray(comp, dtype=np.float64))
    n, p = x.shape
    out = np.empty((n, k), dtype=np.int64)
    ids = np.arange(n)
    for i in range(n):
        d = np.zeros(n)
        den = np.zeros(n)
        for t in range(p):                                   # index order: the same rounding as the metric's loop
            a = np.abs(x[i, t] - x[:, t])
            if distance in ('manhattan', 'cityblock', 'l1'):
                d += a
            elif distance in ('chebyshev', 'infinity'):
                d = np.maximum(d, a)
            elif distance == 'hamming':
                d += (x[i, t] != x[:, t])
            elif distance == 'canberra':
                q = np.abs(x[i, t]) + np.abs(x[:, t])
                d += np.divide(a, q, out=np.zeros(n), where=q > 0)
            elif distance == 'braycurtis':
                d += a
                den += np.abs(x[i, t]) + np.abs(x[:, t])
            elif distance in ('euclidean', 'l2', 'minkowski'):
                d += a * a
            else:
                raise ValueError("Unrecognized metric '%s'" % distance)
        if distance == 'braycurtis':
            d = np.divide(d, den, out=np.zeros(n), where=den > 0)
        key = ids.copy()
        key[i] = -1
        out[i] = ids[np.lexsort((key, d))[:k]]
    return out + 1


def metric_dist_rows(comp, rows, idx1, distance):
    """Distances (same definitions as knn_metric, up to the monotone final step) of x[rows] to x[idx1 - 1]."""
    x = np.asarray(comp, dtype=np.float64)
    a = x[rows][:, None, :]
    b = x[np.asarray(idx1) - 1]
    if distance in ('manhattan', 'cityblock', 'l1'):
        return np.abs(a - b).sum(axis=2)
    if distance in ('chebyshev', 'infinity'):
        return np.abs(a - b).max(axis=2)
    if distance == 'hamming':
        return (a != b).sum(axis=2).astype(np.float64)
    if distance == 'canberra':
        q = np.abs(a) + np.abs(b)
        return np.divide(np.abs(a - b), q, out=np.zeros(q.shape), where=q > 0).sum(axis=2)
    if distance == 'braycurtis':
        den = (np.abs(a) + np.abs(b)).sum(axis=2)
        return np.divide(np.abs(a - b).sum(axis=2), den, out=np.zeros(den.shape), where=den > 0)
    return ((a - b) ** 2).sum(axis=2)


def knn_direct(comp, k=10):
    """Small-case cross-check of knn(): every pair by direct differences."""
    x = np.asarray(comp, dtype=np.float64)
    n = x.shape[0]
    out = np.empty((n, k), dtype=np.int64)
    for i in range(n):
        d = ((x - x[i]) ** 2).sum(axis=1)
        out[i] = np.lexsort((np.arange(n), d))[:k]
    return out + 1


def snn_counts(nn_idx):
    """clustering.py:86-99: M has a one at (nn_idx[i,0]-1, nn_idx[i,c]-1) for c >= 1 and on the
    diagonal (duplicates add up, as coo->csr does); returns V = M M^T as CSR int64."""
    nn = np.asarray(nn_idx, dtype=np.int64)
    n, k = nn.shape
    rows = np.concatenate([np.repeat(nn[:, 0], k - 1), np.arange(1, n + 1)]) - 1
    cols = np.concatenate([nn[:, 1:].ravel(), np.arange(1, n + 1)]) - 1
    m = sp.coo_matrix((np.ones(rows.shape[0], dtype=np.int64), (rows, cols)), shape=(n, n)).tocsr()
    v = (m @ m.T).tocsr()
    v.sort_indices()
    return v


def snn(nn_idx, k=10, prune_snn=0.067):
    """clustering.snn (clustering.py:55-128).  Returns (edges E x 2 int64 sorted by (source,
    target), pruned_count, total_links).  strength = v / (k + (k - v)) must EXCEED prune_snn
    (clustering.py:108-109)."""
    v = snn_counts(nn_idx).tocoo()
    strength = v.data / (k + (k - v.data))
    keep = strength > prune_snn
    src, dst = v.row[keep], v.col[keep]
    n = np.asarray(nn_idx).shape[0]
    present = np.zeros(n, dtype=bool)
    present[src] = True
    present[dst] = True
    missing = np.flatnonzero(~present)                    # clustering.py:114-118
    src = np.concatenate([src, missing]); dst = np.concatenate([dst, missing])
    order = np.lexsort((dst, src))
    edges = np.stack([src[order], dst[order]], axis=1).astype(np.int64)
    return edges, int((~keep).sum()), int(v.data.shape[0])


# --------------------------------------------------------------------------------------------
# whole path
# --------------------------------------------------------------------------------------------


def embed_path(counts, ngenes=1000, ncomp=75, k=10, prune_snn=0.067, seed=42, timings=None):
    """normalize -> seurat -> irlb pca -> knn -> snn on a CSR count matrix."""
    import time
    t = [time.perf_counter()]
    norm, lib = normalize_standard(counts); t.append(time.perf_counter())
    hv = seurat(norm, ngenes); t.append(time.perf_counter())
    pca = irlb_pca(norm, hv, ncomp=ncomp, seed=seed); t.append(time.perf_counter())
    nn = knn(pca['comp'], k); t.append(time.perf_counter())
    edges, pruned, total = snn(nn, k, prune_snn); t.append(time.perf_counter())
    if timings is not None:
        timings.update(dict(zip(['norm', 'hvg', 'pca', 'knn', 'snn'], np.diff(t))))
    return dict(norm=norm, lib=lib, hvg=hv, pca=pca, nn_idx=nn, edges=edges, pruned=pruned, total=total)
