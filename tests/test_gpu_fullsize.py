"""Size-independent properties at BASELINE.json's large shapes (the oracle cannot be run in full there):
C5  kNN / SNN on a 1,000,000 x 50 embedding (k = 10 and 30) -- self first, ascending distances, a random
    sample of queries exact against fp64 brute force, SNN edge set symmetric with one self loop per cell
    and edge weights (recomputed from the neighbour lists) above the pruning threshold;
C4-shard-sized normalize / gene moments (100,000 cells x 28,000 genes, 140 M non-zeros) -- integer library
    sizes exact, linearity sum_g (2^y - 1) = scaling_factor per cell, per-gene sums against numpy."""
import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def eng():
    from adobo_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


@pytest.mark.parametrize('k', [10, 30])
def test_c5_knn_snn_one_million_points(eng, k):
    from adobo_b200.synth import synth_embedding
    n, p = 1_000_000, 50
    x = synth_embedding(n, p, seed=5)
    nn = eng.knn(x, k)
    assert nn.shape == (n, k) and nn.min() >= 1 and nn.max() <= n
    assert np.array_equal(nn[:, 0], np.arange(1, n + 1))                      # self first
    rng = np.random.default_rng(1)
    rows = rng.choice(n, 192, replace=False)
    d_ret = np.einsum('ijk,ijk->ij', x[rows][:, None, :] - x[nn[rows] - 1], x[rows][:, None, :] - x[nn[rows] - 1])
    assert np.all(np.diff(d_ret, axis=1) >= 0)                                # ascending
    for lo in range(0, rows.size, 48):                                        # exact vs fp64 brute force
        r = rows[lo:lo + 48]
        d = ((x[r] ** 2).sum(1)[:, None] + (x ** 2).sum(1)[None, :] - 2.0 * x[r] @ x.T)
        cand = np.argpartition(d, k + 8, axis=1)[:, :k + 8]
        dd = np.einsum('ijk,ijk->ij', x[r][:, None, :] - x[cand], x[r][:, None, :] - x[cand])
        order = np.lexsort((cand, dd), axis=1)[:, :k]
        ref = np.take_along_axis(cand, order, axis=1) + 1
        dref = np.take_along_axis(dd, order, axis=1)
        got = nn[r]
        mism = got != ref
        assert np.all(~mism | (np.abs(d_ret[lo:lo + 48] - dref) <= 1e-12 * np.maximum(dref, 1e-300)))
    # ---- SNN on the device-resident neighbour lists ----
    edges, pruned, total, ne = eng.snn(None, k, 0.067)
    assert ne == edges.shape[0] and total - pruned == ne                      # no isolated cells here
    src, dst = edges[:, 0].astype(np.int64), edges[:, 1].astype(np.int64)
    assert np.all(np.diff(src) >= 0)
    key = src * n + dst
    assert np.all(np.diff(key) > 0)                                           # sorted, no duplicates
    assert np.array_equal(np.sort(dst * n + src), key)                        # symmetric
    assert np.count_nonzero(src == dst) == n                                  # one self loop per cell
    samp = rng.choice(ne, 2000, replace=False)
    nb = nn - 1
    nb[:, 0] = np.arange(n)
    for e in samp:
        v = np.intersect1d(nb[src[e]], nb[dst[e]]).size
        assert v / (k + (k - v)) > 0.067


def test_c4_shard_normalize_and_moments(eng):
    n, g, per_row = 100_000, 28_000, 1400
    rng = np.random.default_rng(7)
    start = rng.integers(0, 20, n)
    jitter = rng.integers(0, 3, (n, per_row), dtype=np.int64)
    cols = start[:, None] + 19 * np.arange(per_row)[None, :] + jitter         # ascending, unique, < g
    assert cols.max() < g
    counts = rng.geometric(0.55, (n, per_row)).astype(np.int32)
    counts[rng.random((n, per_row)) < 0.001] += 40                             # a few large counts (direct path)
    indptr = np.arange(0, n * per_row + 1, per_row, dtype=np.int64)
    m = sp.csr_matrix((counts.ravel(), cols.ravel().astype(np.int32), indptr), shape=(n, g))
    eng.upload_counts(m)
    vals, lib = eng.normalize(10000, 1, 1.0)
    assert np.array_equal(lib, counts.sum(axis=1, dtype=np.int64))            # integer work: exact
    y = vals.data.reshape(n, per_row)
    assert np.allclose(np.exp2(y).sum(axis=1) - per_row, 10000.0, rtol=1e-11)  # linearity of the scaling
    ref = np.log2(counts / lib[:, None].astype(np.float64) * 10000 + 1)
    assert np.max(np.abs(y - ref) / ref) < 1e-14
    mean, var = eng.gene_moments()
    s1 = np.bincount(cols.ravel(), weights=ref.ravel(), minlength=g)
    s2 = np.bincount(cols.ravel(), weights=(ref * ref).ravel(), minlength=g)
    rm = s1 / n
    rv = (s2 - n * rm * rm) / (n - 1)
    ok = rm > 0
    assert np.max(np.abs(mean[ok] - rm[ok]) / rm[ok]) < 1e-6
    assert np.max(np.abs(var[ok] - rv[ok]) / rv[ok]) < 1e-5


def test_irlb_long_shard_kernels_against_the_oracle(eng):
    """IRLB on a shard long enough (>= 65,536 cells) to select the long-shard variants of the step kernels
    (two-CTA-per-SM w_step2, streamed rotate_basis): same tolerances as the golden cases, against the oracle's
    implicit-operator Lanczos on the same normalized matrix, and against the exact SVD for the leading values."""
    from adobo_b200.synth import synth_counts
    from oracle import adobo_oracle as O
    from tests import parity as P
    n, g, h, ncomp = 70_000, 500, 160, 8
    counts = synth_counts(n, g, 0.12, 31)
    norm, _ = O.normalize_standard(counts)
    hv = np.sort(O.seurat(norm, h))
    eng.upload_normalized(norm)
    r = eng.irlb(hv, ncomp=ncomp, seed=42)
    o = O.irlb_pca(norm, hv, ncomp=ncomp, seed=42)
    np.testing.assert_allclose(r['s'], o['s'], rtol=1e-6)
    ex_s = np.linalg.svd(O.scale_dense(norm[:, hv]), compute_uv=False)[:ncomp]
    np.testing.assert_allclose(r['s'], ex_s, rtol=1e-6)
    assert np.max(P.component_err(r['comp'], o['comp'])[:ncomp - 2]) < 1e-4
    assert np.max(P.component_err(r['contr'], o['contr'])[:ncomp - 2]) < 1e-4
