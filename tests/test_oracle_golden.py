"""Pins the oracle (oracle/adobo_oracle.py) to the golden vectors frozen from the unmodified
reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest

from oracle import adobo_oracle as O
from tests import parity as P


@pytest.fixture(scope='module', params=P.CASES)
def case(request):
    return P.load_case(request.param)


def test_normalize_bit_identical(case):
    norm, lib = O.normalize_standard(case['counts_csr'])
    assert np.array_equal(norm.data, case['norm_values'])          # fp64, bit-for-bit
    assert np.array_equal(lib, np.asarray(case['counts_csr'].sum(axis=1)).ravel())


def test_seurat_ordered_list(case):
    mean, var = O.gene_moments(case['norm_csr'])
    z, bins = O.seurat_zscores(mean, var)
    hv = O.seurat_order(z, bins, case['ngenes'])
    P.assert_hvg_equal(hv, case['hvg_ordered'], z)
    assert np.array_equal(np.sort(hv), np.sort(case['hvg_ordered'])) or np.isnan(z[case['hvg_ordered'][-1]]) \
        or (z[hv] == z[case['hvg_ordered'][-1]]).any()


@pytest.mark.parametrize('implicit', [False, True])
def test_irlb_pca(case, implicit):
    r = O.irlb_pca(case['norm_csr'], case['hvg_ordered'], ncomp=case['ncomp'], implicit=implicit)
    assert np.array_equal(r['genes'], case['contr_genes'])          # original gene order, dr.py:319
    ex_comp, ex_contr, ex_s = P.exact_pca(case['norm_csr'], case['hvg_ordered'], case['ncomp'])
    P.assert_pca_close(r['comp'], r['contr'], case['comp'], case['contr'], ex_comp, ex_contr)
    np.testing.assert_allclose(r['s'], case['s'], rtol=1e-6)
    assert abs(r['steps'] - int(case['steps'])) <= max(2, int(0.15 * int(case['steps'])))


def test_knn(case):
    nn = O.knn(case['comp'], case['k'])
    P.assert_knn_equal(nn, case['nn_idx'], case['comp'])
    if case['n'] <= 1000:
        P.assert_knn_equal(O.knn_direct(case['comp'], case['k']), case['nn_idx'], case['comp'])


def test_snn_edges_and_pruned_count(case):
    e, pruned, total = O.snn(case['nn_idx'], case['k'], 0.067)
    assert np.array_equal(e, case['edges'])
    msg = str(case['snn_message']).splitlines()[0]
    assert msg == '%.2f%% (n=%s) of links pruned' % (pruned / total * 100, '{:,}'.format(pruned))


def test_formula_kat_from_notebook():
    """choroidPlexus.ipynb:1338-1346 prints normalized values of one cell (AAACCTGAGCCAGTTT):
    Tcea1 1.235510 and Atp6v1h 1.891144.  Under y = log2(c * 10000 / S + 1)
    (normalize.py:285-286, 559-560) the ratio (2^y2 - 1) / (2^y1 - 1) is the ratio of the two
    integer counts: it must be 2/1, and S follows from either value."""
    y1, y2 = 1.235510, 1.891144
    assert abs((2 ** y2 - 1) / (2 ** y1 - 1) - 2.0) < 1e-5
    S = 10000.0 / (2 ** y1 - 1)
    import scipy.sparse as sp
    lib = int(round(S))
    counts = sp.csr_matrix(np.array([[1, 2, lib - 3]], dtype=np.int64))
    norm, _ = O.normalize_standard(counts)
    assert abs(norm.data[0] - y1) < 2e-5 and abs(norm.data[1] - y2) < 2e-5


@pytest.mark.parametrize('distance', ['manhattan', 'chebyshev', 'canberra', 'braycurtis', 'hamming', 'l1', 'infinity',
                                      'minkowski'])
def test_knn_metric_is_the_ball_tree(distance):
    """oracle.knn_metric restates sklearn's metrics: the reference's own call (clustering.py:47-50) on the same
    points gives the same neighbours (up to ties in the metric)."""
    from functools import partial
    from sklearn.neighbors import NearestNeighbors
    rng = np.random.default_rng(11)
    x = rng.normal(size=(300, 7))
    if distance == 'hamming':
        x = np.round(x)                                      # coordinates that can be equal
    x[40:44] = x[40]                                         # duplicate cells
    ref = NearestNeighbors(n_neighbors=9, algorithm='ball_tree', metric=distance).fit(x).kneighbors(x)[1] + 1
    ours = O.knn_metric(x, 9, distance)
    tied = P.assert_knn_equal(ours, ref, x, dist_rows=partial(O.metric_dist_rows, distance=distance))
    if distance != 'hamming':
        assert tied <= 100                                   # of 2700 entries: only around the duplicate cells
