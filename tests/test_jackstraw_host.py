"""dr.jackstraw routed through the path's boundary.  Host logic only (no GPU): the engine is
replaced by the oracle's IRLB, so what is tested is the sparse permutation trick, the empirical p-values and the
per-component test -- the GPU IRLB itself is covered by tests/test_gpu_parity.py."""
import numpy as np
import pandas as pd
import scipy.sparse as sp

from adobo_b200 import dr
from adobo_b200.data import dataset
from adobo_b200.synth import synth_counts
from oracle import adobo_oracle as O


class OracleEngine:
    """The two engine calls jackstraw makes, answered by the CPU oracle."""
    resident_key = None

    def upload_normalized(self, norm, key=None):
        self.norm = sp.csr_matrix(norm)

    def irlb(self, genes, ncomp=75, maxit=1000, seed=None, scale=True, var_weigh=True, **kw):
        return O.irlb_pca(self.norm, np.asarray(genes), ncomp=ncomp, scale=scale, var_weigh=var_weigh, seed=seed)


def test_sparse_permutation_equals_permuting_the_scaled_dense_rows():
    """dr.py:606-613 permutes rows of the dense SCALED matrix; here the unscaled sparse gene is permuted and scaled
    afterwards.  With the same cell assignment both give the same scaled matrix."""
    rng = np.random.default_rng(3)
    a = sp.random(60, 12, density=0.3, random_state=5, format='csc')
    cols = [2, 7, 11]
    p = dr.permute_gene_columns(a, cols, np.random.default_rng(9))
    dense_scaled = O.scale_dense(sp.csr_matrix(a))
    got = O.scale_dense(sp.csr_matrix(p))
    for j in range(12):
        if j in cols:
            assert sorted(p[:, j].data.tolist()) == sorted(a[:, j].data.tolist())       # same values, moved
            assert np.allclose(np.sort(got[:, j]), np.sort(dense_scaled[:, j]))           # a permutation of the scaled gene
            assert not np.array_equal(p[:, j].toarray(), a[:, j].toarray())
        else:
            assert np.array_equal(got[:, j], dense_scaled[:, j])
    assert p.has_sorted_indices or (np.diff(p.indptr) <= 1).all() or all(
        np.all(np.diff(p.indices[p.indptr[j]:p.indptr[j + 1]]) > 0) for j in range(12))
    del rng


def test_keyed_bijection_matches_the_library_and_permutes():
    """The permutation the DEVICE applies (adobo_irlb_permuted) is a keyed bijection of [0, n): its numpy restatement
    (dr.perm_rows, used by the host path and as the checker) equals the library's own host function bit for bit."""
    from adobo_b200 import _lib
    lib = _lib.load()
    for n in (1, 2, 3, 17, 1000, 20000):
        for col in (0, 7, 999):
            rows = np.arange(n)
            got = dr.perm_rows(n, 0x1234567890ABCDEF, col, rows)
            assert np.array_equal(np.sort(got), rows)                                   # a permutation of the cells
            ref = np.array([lib.adobo_perm_row(n, 0x1234567890ABCDEF, col, int(r)) for r in rows[:500]])
            assert np.array_equal(got[:500], ref)
    a = dr.perm_rows(20000, 1, 0, np.arange(20000))
    b = dr.perm_rows(20000, 1, 1, np.arange(20000))
    c = dr.perm_rows(20000, 2, 0, np.arange(20000))
    assert np.mean(a == b) < 0.01 and np.mean(a == c) < 0.01 and np.mean(a == np.arange(20000)) < 0.01
    assert abs(np.corrcoef(a, np.arange(20000))[0, 1]) < 0.05                           # no visible structure
    m = sp.random(300, 9, density=0.2, random_state=4, format='csc')
    p = dr.permute_gene_columns_keyed(m, [1, 4], 99)
    for j in range(9):
        if j in (1, 4):
            assert sorted(p[:, j].data.tolist()) == sorted(m[:, j].data.tolist())
            assert np.array_equal(p[:, j].toarray().ravel()[dr.perm_rows(300, 99, j, np.arange(300))], m[:, j].toarray().ravel())
        else:
            assert np.array_equal(p[:, j].toarray(), m[:, j].toarray())


def test_bh_adjustment_matches_the_reference_formula():
    p = np.array([0.01, 0.04, 0.03, 0.20, np.nan, 0.5])
    q = dr._p_adjust_bh(p)
    ok = ~np.isnan(p)
    m = ok.sum()
    order = np.argsort(p[ok])
    exp = np.minimum.accumulate((p[ok][order] * m / np.arange(1, m + 1))[::-1])[::-1]
    assert np.allclose(q[ok][order], np.minimum(exp, 1)) and np.isnan(q[4])


def test_jackstraw_flags_the_structured_components_only():
    rng = np.random.default_rng(8)
    n, g, planted = 300, 60, 15
    base = rng.poisson(0.6, (n, g)).astype(np.float64)                 # sparse-ish noise
    z = (rng.random(n) < 0.5).astype(np.float64)
    base[:, :planted] += 3.0 * z[:, None] * (rng.random((n, planted)) < 0.9)   # one planted factor on 15 genes
    norm = sp.csr_matrix(np.log2(base + 1))
    genes = np.array(['g%d' % i for i in range(g)])
    cells = ['c%d' % i for i in range(n)]
    hv = np.arange(g)
    obj = dataset.from_csr(sp.csr_matrix(base.astype(np.int64)), genes=list(genes), cells=cells)
    frame = pd.DataFrame.sparse.from_spmatrix(sp.csc_matrix(norm.T), index=genes, columns=cells)
    o = O.irlb_pca(norm, hv, ncomp=6, seed=42)
    obj.norm_data['standard'] = {'data': frame, 'hvg': {'genes': genes},
                                 'dr': {'pca': {'comp': pd.DataFrame(o['comp']),
                                                'contr': pd.DataFrame(o['contr'], index=genes)}}}
    res, final = dr.jackstraw(obj, permutations=25, ncomp=6, subset_frac_genes=0.1, score_thr=0.05, fdr=0.05,
                              seed=1, eng=OracleEngine())
    assert res.shape == (g, 6) and list(res.columns) == ['PC%d' % i for i in range(6)]
    assert ((res.values >= 0) & (res.values <= 1)).all()
    assert list(final.columns) == ['PC', 'chi2_p', 'chi2_p_adj', 'significant']
    # the planted genes beat every permuted loading on the first component; the noise components look like the null
    assert (res['PC0'].values[:planted] < 0.05).all()
    assert (res['PC0'].values[planted:] < 0.05).mean() < 0.2
    assert (res.iloc[:, -3:].values < 0.05).mean() < 0.25
    assert final['significant'].iloc[0] and not final['significant'].iloc[-1]
    assert obj.norm_data['standard']['dr']['jackstraw']['score_mat'] is res
