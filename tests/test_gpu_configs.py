"""Parity at BASELINE.json's own shapes: c1 (8,400 x 33,000, the reference's CPU-runnable case) and c2 (20,000 x
25,000, bench.py's `--config c2`), against golden vectors frozen from the UNMODIFIED reference
(tests/golden/make_golden_configs.py).  Stage-wise -- every stage fed the reference's own upstream output -- and
end to end.  The per-component PCA error vectors are printed (pytest -s) and summarised in DESIGN.md."""
import numpy as np
import pytest

from tests import parity as P

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def eng():
    from adobo_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


@pytest.fixture(scope='module', params=P.CONFIG_CASES)
def cfg(request):
    return P.load_config_case(request.param)


def test_normalize_at_config_shape(eng, cfg):
    eng.upload_counts(cfg['counts_csr'])
    vals, lib = eng.normalize(10000, 1, 1.0)
    assert np.array_equal(lib, cfg['libsize'])                                             # integer: exact
    got = vals.data[cfg['norm_sample_pos']]
    assert np.max(np.abs(got - cfg['norm_sample']) / cfg['norm_sample']) < 1e-14           # north_star asks 1e-5
    assert abs(vals.data.sum() - float(cfg['norm_sum'])) <= 1e-12 * float(cfg['norm_sum'])


def test_moments_and_hvg_at_config_shape(eng, cfg, monkeypatch):
    from oracle import adobo_oracle as O
    eng.upload_counts(cfg['counts_csr'])
    eng.normalize(10000, 1, 1.0, want_values=False)
    mean, var = eng.gene_moments()
    P.assert_values_close(mean, cfg['gene_mean'], rtol=1e-5, what='gene mean')
    ok = cfg['gene_var'] > 0
    P.assert_values_close(var[ok], cfg['gene_var'][ok], rtol=1e-5, what='gene var')
    hv, z = eng.hvg_seurat(cfg['ngenes'], 20)
    rz, _ = O.seurat_zscores(cfg['gene_mean'], cfg['gene_var'])
    P.assert_hvg_equal(hv, cfg['hvg_ordered'], rz, rtol=1e-5)                              # the ORDERED list
    # bit-reproducible: a second pass gives the same bits (fixed-point integer sums are order independent)
    eng.normalize(10000, 1, 1.0, want_values=False)
    mean2, var2 = eng.gene_moments()
    assert np.array_equal(mean, mean2) and np.array_equal(var, var2, equal_nan=True)
    # the tiled kernel (private accumulators in shared memory, no atomics) and the atomic one (small matrices) add the
    # same integers: same bits
    monkeypatch.setenv('ADOBO_B200_MOMENTS_ATOMIC', '1')
    eng.normalize(10000, 1, 1.0, want_values=False)
    mean3, var3 = eng.gene_moments()
    assert np.array_equal(mean, mean3) and np.array_equal(var, var3, equal_nan=True)


def test_irlb_at_config_shape(eng, cfg, capsys):
    eng.upload_counts(cfg['counts_csr'])
    eng.normalize(10000, 1, 1.0, want_values=False)
    r = eng.irlb(cfg['hvg_ordered'], ncomp=cfg['ncomp'], seed=42)
    assert np.array_equal(r['genes'], cfg['contr_genes'])
    e, n_over, bad = P.pca_report(r['comp'], cfg['comp'], cfg['exact_comp_err'])
    ec, nc_over, badc = P.pca_report(r['contr'], cfg['contr'], cfg['exact_contr_err'])
    with capsys.disabled():
        print('\n[%d x %d] restarts %d (reference %d), mat-vecs %d (%d); comp: %d of %d components > 1e-4 from the '
              "reference's (its own error vs the exact SVD exceeds 1e-4 on %d); contr: %d"
              % (cfg['n'], cfg['g'], r['steps'], int(cfg['steps']), r['nmult'], int(cfg['nmult']), n_over, e.size,
                 int(np.sum(cfg['exact_comp_err'] > 1e-4)), nc_over))
        print('comp err  ', np.array2string(e, precision=1, max_line_width=200))
        print('ref vs SVD', np.array2string(cfg['exact_comp_err'], precision=1, max_line_width=200))
    assert bad.size == 0 and badc.size == 0, (bad, e[bad], cfg['exact_comp_err'][bad])
    np.testing.assert_allclose(r['s'], cfg['s'], rtol=1e-6)
    # trajectory parity is not bit-level: the restart count of the reference algorithm moves by a few per cent
    # under 1-ulp changes of the arithmetic
    assert abs(r['steps'] - int(cfg['steps'])) <= max(2, int(0.15 * int(cfg['steps'])))


def test_knn_snn_stagewise_at_config_shape(eng, cfg):
    nn = eng.knn(cfg['comp'], cfg['k'])
    P.assert_knn_equal(nn, cfg['nn_idx'], cfg['comp'])
    edges, pruned, total, ne = eng.snn(cfg['nn_idx'], cfg['k'], 0.067)
    assert np.array_equal(P.edges_sorted(edges), cfg['edges'])
    msg = str(cfg['snn_message']).splitlines()[0]
    assert msg == '%.2f%% (n=%s) of links pruned' % (pruned / total * 100, '{:,}'.format(pruned))


def test_whole_path_at_config_shape(eng, cfg, capsys):
    eng.upload_counts(cfg['counts_csr'])
    np.random.seed(42)
    v0 = np.random.randn(cfg['ngenes'])
    r = eng.embed_path(v0, ngenes=cfg['ngenes'], ncomp=cfg['ncomp'], k=cfg['k'])
    pca = eng.fetch_pca(cfg['ngenes'], cfg['ncomp'])
    assert np.array_equal(pca['genes'], cfg['contr_genes'])                               # same HVG set end to end
    e, n_over, bad = P.pca_report(pca['comp'], cfg['comp'], cfg['exact_comp_err'])
    nn = eng.fetch_knn(cfg['k'])
    edges = eng.fetch_edges(r['n_edges'])
    ours, ref = set(map(tuple, edges.tolist())), set(map(tuple, cfg['edges'].tolist()))
    jac = len(ours & ref) / len(ours | ref)
    agree = float(np.mean(nn == cfg['nn_idx']))
    with capsys.disabled():
        print('\n[%d x %d] end to end: restarts %d (reference %d), %d components > 1e-4, kNN entries equal %.4f, edges %d '
              '(reference %d), Jaccard %.4f' % (cfg['n'], cfg['g'], r['steps'], int(cfg['steps']), n_over, agree, len(ours),
                                               len(ref), jac))
    assert bad.size == 0, (bad, e[bad])
    assert agree > 0.97 and jac > 0.97
