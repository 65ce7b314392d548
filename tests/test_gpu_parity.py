"""GPU parity tests: the CUDA path, called through the C-ABI (adobo_b200.engine.Engine is a thin
ctypes wrapper), against the golden vectors frozen from the unmodified reference and against the
oracle on seeded inputs.  Each stage is fed the reference's upstream output (stage-wise parity)."""
import numpy as np
import pytest
import scipy.sparse as sp

from tests import parity as P

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def eng():
    from adobo_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


@pytest.fixture(scope='module', params=P.CASES)
def case(request):
    return P.load_case(request.param)


def test_normalize_values_and_library_size(eng, case):
    eng.upload_counts(case['counts_csr'])
    vals, lib = eng.normalize(10000, 1, 1.0)
    assert np.array_equal(lib, np.asarray(case['counts_csr'].sum(axis=1)).ravel())        # integer: exact
    P.assert_values_close(vals.data, case['norm_values'], rtol=1e-5, what='normalized values')
    # fp64 path is far tighter than the tolerance: a few ulp of CUDA's log2
    assert np.max(np.abs(vals.data - case['norm_values']) / case['norm_values']) < 1e-14


def test_normalize_large_counts_and_log_modes(eng):
    from oracle import adobo_oracle as O
    rng = np.random.default_rng(3)
    dense = rng.poisson(0.2, (300, 500)) * rng.integers(1, 100000, (300, 500))
    dense[np.arange(300), rng.integers(0, 500, 300)] += 1
    counts = sp.csr_matrix(dense.astype(np.int64))
    eng.upload_counts(counts)
    for mode, fn in ((1, np.log2), (2, np.log), (3, np.log10)):
        vals, lib = eng.normalize(10000, mode, 1.0)
        ref, reflib = O.normalize_standard(counts, 10000, True, fn, 1)
        assert np.array_equal(lib, reflib)
        P.assert_values_close(vals.data, ref.data, rtol=1e-12, what='mode %d' % mode)
    vals, _ = eng.normalize(5000, 0, 1.0)
    ref, _ = O.normalize_standard(counts, 5000, log=False)
    P.assert_values_close(vals.data, ref.data, rtol=1e-14, what='no log')


def test_gene_moments_and_hvg_ranking(eng, case):
    from oracle import adobo_oracle as O
    eng.upload_normalized(case['norm_csr'])
    mean, var = eng.gene_moments()
    rmean, rvar = O.gene_moments(case['norm_csr'])
    P.assert_values_close(mean, rmean, rtol=1e-5, what='gene mean')
    ok = rvar > 0
    P.assert_values_close(var[ok], rvar[ok], rtol=1e-5, what='gene var')
    hv, z = eng.hvg_seurat(case['ngenes'], 20)
    rz, _ = O.seurat_zscores(rmean, rvar)
    fin = np.isfinite(rz)
    assert np.array_equal(np.isnan(z), np.isnan(rz))
    assert np.max(np.abs(z[fin] - rz[fin]) / np.maximum(1.0, np.abs(rz[fin]))) < 1e-5
    P.assert_hvg_equal(hv, case['hvg_ordered'], rz, rtol=1e-5)


def test_irlb_pca(eng, case):
    eng.upload_normalized(case['norm_csr'])
    r = eng.irlb(case['hvg_ordered'], ncomp=case['ncomp'], seed=42)
    assert np.array_equal(r['genes'], case['contr_genes'])
    ex_comp, ex_contr, _ = P.exact_pca(case['norm_csr'], case['hvg_ordered'], case['ncomp'])
    P.assert_pca_close(r['comp'], r['contr'], case['comp'], case['contr'], ex_comp, ex_contr)
    np.testing.assert_allclose(r['s'], case['s'], rtol=1e-6)
    assert abs(r['steps'] - int(case['steps'])) <= max(2, int(0.15 * int(case['steps'])))


def test_irlb_unscaled_matches_oracle(eng):
    from oracle import adobo_oracle as O
    case = P.load_case('small')
    eng.upload_normalized(case['norm_csr'])
    r = eng.irlb(case['hvg_ordered'], ncomp=10, seed=7, scale=False, var_weigh=False)
    o = O.irlb_pca(case['norm_csr'], case['hvg_ordered'], ncomp=10, scale=False, var_weigh=False, seed=7)
    assert np.max(P.component_err(r['comp'], o['comp'])) < 1e-4
    assert np.max(P.component_err(r['contr'], o['contr'])) < 1e-4
    np.testing.assert_allclose(r['s'], o['s'], rtol=1e-6)


def test_lanczos_on_explicit_matrix(eng):
    """irlbpy.lanczos(A, nval) on the dense scaled matrix, as dr.irlb calls it (dr.py:163)."""
    from oracle import adobo_oracle as O
    case = P.load_case('small')
    A = O.scale_dense(case['norm_csr'][:, np.sort(case['hvg_ordered'])])
    r = eng.lanczos_csr(sp.csr_matrix(A), case['ncomp'], maxit=1000, seed=42)
    ex_comp, ex_contr, _ = P.exact_pca(case['norm_csr'], case['hvg_ordered'], case['ncomp'])
    P.assert_pca_close(r.U * r.s[None, :], r.V, case['comp'], case['contr'], ex_comp, ex_contr)
    np.testing.assert_allclose(r.s, case['s'], rtol=1e-6)
    assert r.nmult > 0 and r.steps >= 0


def test_knn_exact(eng, case):
    nn = eng.knn(case['comp'], case['k'])
    assert nn.dtype == np.int64 and nn.shape == case['nn_idx'].shape
    P.assert_knn_equal(nn, case['nn_idx'], case['comp'])
    d = P.knn_dist(case['comp'], nn)
    assert np.all(np.diff(d, axis=1) >= 0)                                # ascending by distance


@pytest.mark.parametrize('k', [1, 16, 17, 30, 48])
def test_knn_other_k(eng, k):
    from oracle import adobo_oracle as O
    from adobo_b200.synth import synth_embedding
    x = synth_embedding(1500, 20, seed=9)
    nn = eng.knn(x, k)
    P.assert_knn_equal(nn, O.knn(x, k), x)
    assert np.array_equal(nn[:, 0], np.arange(1, 1501))                   # self first


def test_knn_duplicates_and_tiny(eng):
    from oracle import adobo_oracle as O
    rng = np.random.default_rng(5)
    x = rng.normal(size=(400, 8))
    x[100:180] = x[100]                                                   # 80 identical cells (> shortlist)
    nn = eng.knn(x, 10)
    ref = O.knn(x, 10)
    P.assert_knn_equal(nn, ref, x)
    assert np.array_equal(nn[:, 0], np.arange(1, 401))                    # self wins the zero-distance tie
    y = rng.normal(size=(12, 3))
    P.assert_knn_equal(eng.knn(y, 12), O.knn(y, 12), y)                   # k == n, n < one tile


@pytest.mark.parametrize('distance', ['manhattan', 'chebyshev', 'hamming', 'canberra', 'braycurtis'])
def test_knn_other_metrics(eng, distance):
    """clustering.knn(distance=...) (clustering.py:47-48): the fp64 brute-force kernel against the oracle's restatement
    of sklearn's metrics (itself pinned to the ball tree on the CPU), duplicates and a ragged last tile included."""
    from functools import partial
    from oracle import adobo_oracle as O
    from adobo_b200 import clustering
    rng = np.random.default_rng(13)
    x = rng.normal(size=(1237, 23))
    if distance == 'hamming':
        x = np.round(x)
    x[500:540] = x[500]
    for k in (10, 33):
        nn = clustering.knn(x, k, distance, eng=eng)
        assert nn.dtype == np.int64 and nn.shape == (1237, k)
        P.assert_knn_equal(nn, O.knn_metric(x, k, distance), x, dist_rows=partial(O.metric_dist_rows, distance=distance))
        assert np.array_equal(nn[:, 0], np.arange(1, 1238))               # self first
    with pytest.raises(ValueError):
        clustering.knn(x, 10, 'mahalanobis', eng=eng)


def test_knn_fp64_brute_force_is_the_tensor_path(eng):
    """ADOBO_METRIC_EUCLIDEAN_FP64 (the brute-force kernel with squared differences) and the tcgen05 shortlist + fp64
    re-rank give the same neighbours."""
    from adobo_b200.synth import synth_embedding
    x = synth_embedding(5000, 50, seed=3)
    a = eng.knn(x, 20)
    b = eng.knn(x, 20, metric=6)
    P.assert_knn_equal(a, b, x)


def test_snn_edges_bit_exact(eng, case):
    edges, pruned, total, ne = eng.snn(case['nn_idx'], case['k'], 0.067)
    assert np.array_equal(P.edges_sorted(edges), case['edges'])
    assert np.array_equal(edges, P.edges_sorted(edges))                   # already sorted (source, target)
    msg = str(case['snn_message']).splitlines()[0]
    assert msg == '%.2f%% (n=%s) of links pruned' % (pruned / total * 100, '{:,}'.format(pruned))


@pytest.mark.parametrize('prune', [0.0, 0.2, 1.0])
def test_snn_prune_levels_and_hub_rows(eng, prune):
    from oracle import adobo_oracle as O
    rng = np.random.default_rng(8)
    n, k = 3000, 10
    nn = np.empty((n, k), dtype=np.int64)
    for i in range(n):
        others = rng.choice(n - 1, k - 2, replace=False) + 1
        others = others[others != i + 1][:k - 2]
        while others.size < k - 2:
            others = np.append(others, (others[-1] % n) + 1)
        nn[i] = np.concatenate([[i + 1, 1 if i else 2], others])         # node 1 is everybody's neighbour: hub
    edges, pruned, total, ne = eng.snn(nn, k, prune)
    ref, rpruned, rtotal = O.snn(nn, k, prune)
    assert (pruned, total) == (rpruned, rtotal)
    assert np.array_equal(P.edges_sorted(edges), ref)


def test_embed_path_end_to_end(eng):
    """Whole device-resident path vs the oracle run end to end (not stage-wise): the graph must be
    essentially the same graph."""
    from oracle import adobo_oracle as O
    case = P.load_case('medium')
    eng.upload_counts(case['counts_csr'])
    np.random.seed(42)
    v0 = np.random.randn(case['ngenes'])
    r = eng.embed_path(v0, ngenes=case['ngenes'], ncomp=case['ncomp'], k=case['k'])
    pca = eng.fetch_pca(case['ngenes'], case['ncomp'])
    assert np.array_equal(pca['genes'], case['contr_genes'])
    e = P.component_err(pca['comp'], case['comp'])
    assert np.max(e[:case['ncomp'] - 8]) < 1e-4
    nn = eng.fetch_knn(case['k'])
    edges = eng.fetch_edges(r['n_edges'])
    ours = set(map(tuple, edges.tolist()))
    ref = set(map(tuple, case['edges'].tolist()))
    assert len(ours & ref) / len(ours | ref) > 0.97
    assert np.mean(nn == case['nn_idx']) > 0.97


def test_api_drop_in(eng):
    """adobo's own call sequence through the adobo_b200 modules (same names / defaults)."""
    import adobo_b200 as ad
    from adobo_b200 import engine as E
    E.set_default_engine(eng)
    case = P.load_case('small')
    exp = ad.data.dataset.from_csr(case['counts_csr'])
    ad.normalize.norm(exp)
    nd = exp.norm_data['standard']['data']
    assert str(nd.dtypes.iloc[0]) == 'Sparse[float64, 0]' and nd.shape == (case['g'], case['n'])
    got = ad.data.frame_to_csr(nd)
    P.assert_values_close(got.data, case['norm_values'], rtol=1e-5)
    ad.hvg.find_hvg(exp, ngenes=case['ngenes'])
    names = exp.norm_data['standard']['hvg']['genes']
    from oracle import adobo_oracle as O
    rz, _ = O.seurat_zscores(*O.gene_moments(case['norm_csr']))
    P.assert_hvg_equal(np.array([int(s[1:]) for s in names]), case['hvg_ordered'], rz, rtol=1e-5)   # ordered list, GPU path
    ad.dr.pca(exp, ncomp=case['ncomp'])
    pca = exp.norm_data['standard']['dr']['pca']
    assert list(pca['comp'].index) == list(nd.columns)
    assert [int(s[1:]) for s in pca['contr'].index] == case['contr_genes'].tolist()
    assert np.max(P.component_err(pca['comp'].values, case['comp'])[:10]) < 1e-4
    ad.clustering.generate(exp, k=case['k'], clust_alg=None)
    g = exp.norm_data['standard']['graph']
    assert list(g.columns) == ['source_node', 'target_node'] and len(g) > case['n']
    assert exp.get_assay('norm') and exp.get_assay('find_hvg') and exp.get_assay('pca') and exp.get_assay('clustering')
    # error behaviour mirrors the reference
    with pytest.raises(Exception, match='Run normalization first'):
        ad.hvg.find_hvg(ad.data.dataset.from_csr(case['counts_csr']))
    with pytest.raises(ValueError):
        ad.dr.pca(exp, genes=5)


def test_knn_blobs_30k_tensor_core_path(eng):
    """C5-shaped input (Gaussian blobs, 50 dims) at 30,000 points: many reference tiles per query tile,
    exercising the TMA ring / TMEM double buffer of the tcgen05 shortlist; exact vs the oracle."""
    from oracle import adobo_oracle as O
    from adobo_b200.synth import synth_embedding
    x = synth_embedding(30000, 50, seed=5)
    for k in (10, 30):
        nn = eng.knn(x, k)
        ref = O.knn(x, k, block=512)
        P.assert_knn_equal(nn, ref, x)
        assert np.array_equal(nn[:, 0], np.arange(1, 30001))


@pytest.mark.parametrize('hook', ['svd_jacobi', 'svd_fallback', 'no_graph', 'no_fork', 'rotate_simt', 'no_fusedt',
                                  'no_fusedt,no_vone', 'no_zpath', 'no_vstep', 'no_fusedt,no_wfold', 'no_fusedt,no_split',
                                  'no_pdl', 'no_runahead', 'no_fusedt,no_pdl', 'no_lazy', 'no_lazy,no_graph', 'SVD_CLUSTER'])
def test_irlb_alternative_paths(eng, hook, monkeypatch):
    """Every fallback of the Lanczos driver (ADOBO_B200_IRLB hooks, csrc/irlb.cu: the kernels for sizes the two-kernel
    step does not cover, the generic Jacobi SVD of B, the synchronous forms of the asynchronous mechanisms) must give
    the same PCA as the default path."""
    case = P.load_case('medium')
    eng.upload_normalized(case['norm_csr'])
    base = eng.irlb(case['hvg_ordered'], ncomp=case['ncomp'], seed=42)
    if hook == 'SVD_CLUSTER':
        monkeypatch.setenv('ADOBO_B200_SVD_CLUSTER', '1')
    else:
        monkeypatch.setenv('ADOBO_B200_IRLB', hook)
    eng.upload_normalized(case['norm_csr'])
    alt = eng.irlb(case['hvg_ordered'], ncomp=case['ncomp'], seed=42)
    ex_comp, ex_contr, _ = P.exact_pca(case['norm_csr'], case['hvg_ordered'], case['ncomp'])
    P.assert_pca_close(alt['comp'], alt['contr'], case['comp'], case['contr'], ex_comp, ex_contr)
    np.testing.assert_allclose(alt['s'], base['s'], rtol=1e-7)  # converged to tol^2; the paths round differently
    assert np.max(P.component_err(alt['comp'], base['comp'])[:case['ncomp'] - 8]) < 1e-6


@pytest.mark.parametrize('maxit', [1, 2, 7, 1000])
def test_irlb_runahead_is_the_synchronous_loop(eng, maxit, monkeypatch):
    """The host enqueues outer iterations ahead of the convergence read-back (device-side halt flag) and the
    step kernels are launched programmatically; neither changes arithmetic: same restarts, mat-vecs and bits as
    the loop that waits for every read-back -- converged or cut off by maxit (irlb.py:155, 238-247)."""
    case = P.load_case('medium')
    eng.upload_normalized(case['norm_csr'])
    fast = eng.irlb(case['hvg_ordered'], ncomp=case['ncomp'], seed=42, maxit=maxit)
    monkeypatch.setenv('ADOBO_B200_IRLB', 'no_runahead,no_pdl')
    eng.upload_normalized(case['norm_csr'])
    slow = eng.irlb(case['hvg_ordered'], ncomp=case['ncomp'], seed=42, maxit=maxit)
    assert (fast['steps'], fast['nmult']) == (slow['steps'], slow['nmult'])
    assert np.array_equal(fast['s'], slow['s'])
    assert np.array_equal(fast['comp'], slow['comp']) and np.array_equal(fast['contr'], slow['contr'])


def test_api_on_csr_backed_dataset(eng):
    """f1: the same calls on a dataset that holds no pandas frame (CsrFrame) give the frame-backed results; find_hvg
    / pca after norm reuse the matrix left on the device (ResidentKey), a different frame of the same shape does not."""
    import adobo_b200 as ad
    from adobo_b200 import engine as E
    from adobo_b200.data import CsrFrame, frame_to_csr
    E.set_default_engine(eng)
    case = P.load_case('small')
    a = ad.data.dataset.from_csr(case['counts_csr'], frame=True)
    b = ad.data.dataset.from_csr(case['counts_csr'], frame=False)
    for exp in (a, b):
        ad.normalize.norm(exp)
        ad.hvg.find_hvg(exp, ngenes=case['ngenes'])
        ad.dr.pca(exp, ncomp=case['ncomp'])
        ad.clustering.generate(exp, k=case['k'], clust_alg=None)
    na, nb = a.norm_data['standard'], b.norm_data['standard']
    assert isinstance(nb['data'], CsrFrame)
    assert np.array_equal(frame_to_csr(na['data']).data, nb['data'].csr.data)
    assert list(na['hvg']['genes']) == list(nb['hvg']['genes'])
    assert np.array_equal(na['dr']['pca']['comp'].values, nb['dr']['pca']['comp'].values)          # same bits
    assert list(nb['dr']['pca']['contr'].index) == list(na['dr']['pca']['contr'].index)
    assert na['graph'].equals(nb['graph'])
    # a fresh frame of the same shape must be uploaded, not mistaken for the resident one
    other = CsrFrame(nb['data'].csr.power(3), nb['data'].index, nb['data'].columns)
    g1 = ad.hvg.seurat(nb['data'], case['ngenes'])
    g2 = ad.hvg.seurat(other, case['ngenes'])
    g3 = ad.hvg.seurat(nb['data'], case['ngenes'])
    assert list(g1) == list(g3) and list(g1) == list(nb['hvg']['genes'])
    assert list(g2) != list(g1)                     # other values: other dispersions, other ranking


def test_device_metadata_and_lazy_normalized_frame(eng):
    """A CsrFrame-backed dataset made while an engine exists: counts uploaded once, data.py:75-84's statistics from
    the device (adobo_count_stats) equal the host ones; norm() leaves the values on the device (DeviceCsrFrame) and
    they reach the host intact on first use, or when another matrix is about to replace the resident one."""
    import adobo_b200 as ad
    from adobo_b200 import engine as E
    from adobo_b200.data import DeviceCsrFrame
    case = P.load_case('medium')
    counts = case['counts_csr'].copy()
    counts.data[::97] = 0                                   # explicit zeros: not 'detected', not 'expressed'
    E.set_default_engine(None)
    host = ad.data.dataset.from_csr(counts, frame=False)    # no engine: host statistics
    E.set_default_engine(eng)
    dev = ad.data.dataset.from_csr(counts, frame=False)
    for col in ('total_reads', 'detected_genes'):
        assert np.array_equal(host.meta_cells[col].values, dev.meta_cells[col].values), col
    assert np.array_equal(host.meta_genes['expressed'].values, dev.meta_genes['expressed'].values)
    dense = np.asarray(counts.todense())
    assert np.array_equal(dev.meta_cells['detected_genes'].values, (dense > 0).sum(axis=1))      # data.py:78
    assert np.array_equal(dev.meta_genes['expressed'].values, (dense > 0).sum(axis=0))           # data.py:83-84
    l0 = eng.launch_count()
    ad.normalize.norm(dev)
    ad.normalize.norm(dev, name='nolog', log=False)         # a second normalization of the same resident counts
    d1, d2 = dev.norm_data['standard']['data'], dev.norm_data['nolog']['data']
    assert isinstance(d1, DeviceCsrFrame) and not d1.materialized and not d2.materialized
    ad.hvg.find_hvg(dev, normalization='nolog', ngenes=200)
    ref1, _ = _eager_norm(eng.__class__, counts, 1)
    ref2, _ = _eager_norm(eng.__class__, counts, 0)
    # another dataset takes the device: both lazy frames are brought to the host first
    other = ad.data.dataset.from_csr(case['counts_csr'][:500], frame=False)
    assert d1.materialized and d2.materialized
    assert np.array_equal(d1.csr.data, ref1.data) and np.array_equal(d2.csr.data, ref2.data)
    assert np.array_equal(d1.csr.indices, counts.indices)
    hv = list(dev.norm_data['nolog']['hvg']['genes'])
    ad.hvg.find_hvg(dev, normalization='nolog', ngenes=200)  # re-uploaded from the host copy: same list
    assert list(dev.norm_data['nolog']['hvg']['genes']) == hv
    del other
    assert eng.launch_count() > l0


def _eager_norm(engine_cls, counts, log_mode):
    e = engine_cls(0)
    e.upload_counts(counts)
    out = e.normalize(10000, log_mode, 1.0)
    e.close()
    return out


def test_snn_two_pass_fallback_is_the_single_pass(eng, monkeypatch):
    """The single pass (kept pairs staged, snn_emit) and the two-pass scheme it falls back to when the staging buffer
    overflows give the same edge list, counts included."""
    case = P.load_case('medium')
    e1, p1, t1, _ = eng.snn(case['nn_idx'], case['k'], 0.067)
    _, v1 = eng.fetch_edges(counts=True)
    monkeypatch.setenv('ADOBO_B200_SNN_TWOPASS', '1')
    e2, p2, t2, _ = eng.snn(case['nn_idx'], case['k'], 0.067)
    _, v2 = eng.fetch_edges(counts=True)
    assert np.array_equal(e1, e2) and (p1, t1) == (p2, t2) and np.array_equal(v1, v2)
    assert np.array_equal(P.edges_sorted(e1), case['edges'])


def test_snn_counts_are_the_reference_matrix_entries(eng):
    """north_star: 'SNN edge weights bit-exact' -- V_ij = (M M^T)_ij of every kept edge against the oracle's sparse
    product (the reference computes V and keeps only its coordinates, clustering.py:99-112)."""
    from oracle import adobo_oracle as O
    for name in ('ragged', 'medium'):
        case = P.load_case(name)
        edges, pruned, total, ne = eng.snn(case['nn_idx'], case['k'], 0.067)
        e2, v = eng.fetch_edges(counts=True)
        assert np.array_equal(e2, edges)
        V = O.snn_counts(case['nn_idx'])
        ref = np.asarray(V[edges[:, 0], edges[:, 1]]).ravel()
        appended = (edges[:, 0] == edges[:, 1]) & (v == 0)            # nodes without any kept edge (clustering.py:114-118)
        assert np.array_equal(v[~appended], ref[~appended]) and v[~appended].min() >= 1   # duplicates in nn_idx add up


def test_impute_subpopulations_front_half(eng):
    """f4: preproc.impute's subpopulation estimate (preproc.py:461-488) through the GPU path against the oracle run on
    the DENSE lnorm = log10(cpm + 1.01) the reference forms (every zero becomes log10(1.01))."""
    import adobo_b200 as ad
    from oracle import adobo_oracle as O
    case = P.load_case('small')
    counts = case['counts_csr']
    exp = ad.data.dataset.from_csr(counts, frame=False)
    out = ad.preproc.impute_subpopulations(exp, ncomp=20, seed=11, clust_alg=None, verbose=False, eng=eng)
    lib = np.asarray(counts.sum(axis=1)).ravel().astype(np.float64)
    dense = np.log10(np.asarray(counts.todense(), dtype=np.float64) * (10 ** 6 / lib)[:, None] + 1.01)   # preproc.py:470-472
    lnorm = sp.csr_matrix(dense)
    mean, var = O.gene_moments(lnorm)
    z, bins = O.seurat_zscores(mean, var)
    hv = O.seurat_order(z, bins, 1000)
    assert sorted(int(g[1:]) for g in out['hvg']) == sorted(hv.tolist())
    o = O.irlb_pca(lnorm, hv, ncomp=20, seed=11)
    assert np.max(P.component_err(out['comp'].values, o['comp'])[:12]) < 1e-4
    P.assert_knn_equal(eng.knn(o['comp'], 10), O.knn(o['comp'], 10), o['comp'])
    assert list(out['graph'].columns) == ['source_node', 'target_node'] and out['nn_idx'].shape == (case['n'], 10)


def test_pca_contributors_table_per_cluster(eng):
    """f4: plotting.pca_contributors' data half; the per-cluster case re-runs irlb on the cluster's cells
    (plotting.py:298-303)."""
    import adobo_b200 as ad
    import pandas as pd
    from adobo_b200 import engine as E
    E.set_default_engine(eng)
    case = P.load_case('small')
    exp = ad.data.dataset.from_csr(case['counts_csr'], frame=True)
    ad.normalize.norm(exp)
    ad.hvg.find_hvg(exp, ngenes=case['ngenes'])
    ad.dr.pca(exp, ncomp=case['ncomp'])
    nd = exp.norm_data['standard']
    tab = ad.plotting.pca_contributors_table(exp, dim=[0, 3], top=5)
    s0 = nd['dr']['pca']['contr'][0].sort_values(ascending=False)
    assert list(tab[0].index) == list(s0.head(5).index) + list(s0.tail(5).index)
    cl = pd.Series(np.arange(case['n']) % 3, index=nd['data'].columns)
    nd['clusters']['leiden'] = {'membership': cl}
    tab1 = ad.plotting.pca_contributors_table(exp, cluster=1, dim=[0], top=5)
    X_ss = nd['data'].loc[:, (cl == 1).to_numpy()]
    X_ss = X_ss[X_ss.index.isin(nd['hvg']['genes'])]
    np.random.seed(None)
    _, contr = ad.dr.irlb(X_ss, seed=None)
    # the start vector is random (seed=None, like the reference): compare the leading loadings up to sign
    a, b = tab1[0], contr[0]
    assert set(a.index) <= set(b.index)
    assert np.allclose(np.abs(a.values), np.abs(b.loc[a.index].values), rtol=1e-3, atol=1e-6)


def test_jackstraw_on_the_device_matches_the_oracle_engine(eng):
    """f3: dr.jackstraw with every permutation made and decomposed on the device (adobo_irlb_permuted: keyed bijection
    per gene, key sort back to CSR, IRLB with the unpermuted statistics) against the same call answered by the oracle
    (host permutation with the numpy restatement of the bijection + CPU IRLB), same seed: same p-value table."""
    import adobo_b200 as ad
    from adobo_b200 import engine as E
    from tests.test_jackstraw_host import OracleEngine
    E.set_default_engine(eng)
    case = P.load_case('small')
    exp = ad.data.dataset.from_csr(case['counts_csr'], frame=True)
    ad.normalize.norm(exp)
    ad.hvg.find_hvg(exp, ngenes=case['ngenes'])
    ad.dr.pca(exp, ncomp=8)
    args = dict(permutations=12, ncomp=8, subset_frac_genes=0.1, score_thr=0.05, fdr=0.05, seed=5)
    res_d, fin_d = ad.dr.jackstraw(exp, eng=eng, **args)
    res_h, fin_h = ad.dr.jackstraw(exp, eng=OracleEngine(), **args)
    nnull = 12 * int(round(case['ngenes'] * 0.1))
    # empirical p-values are counts / nnull: a loading within 1e-4 of a null loading may fall on either side
    assert np.max(np.abs(res_d.values - res_h.values)) <= 3.0 / nnull
    assert np.mean(res_d.values == res_h.values) > 0.98
    assert list(fin_d['significant']) == list(fin_h['significant'])
    # the real PCA and the resident matrix survive the permutations
    pca = eng.fetch_pca()
    assert np.array_equal(pca['comp'], exp.norm_data['standard']['dr']['pca']['comp'].values)
