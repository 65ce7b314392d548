"""The oracle against the LIVE unmodified reference (only where the read-only checkout exists: the build
container).  tests/golden/*.npz are what travels to the GPU box; this test proves they are what the reference
produces today -- it re-runs tests/golden/make_golden.py's procedure for the smallest case and compares every array
with the committed file bit for bit -- and runs the oracle beside the reference on a fresh seed that has no golden
file."""
import os
import sys

import numpy as np
import pytest

from oracle import adobo_oracle as O
from oracle import ref_loader
from tests import parity as P

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason='reference checkout not present')


@pytest.fixture(scope='module')
def ref():
    return ref_loader.load()


def _make_golden():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
    import make_golden
    return make_golden


def test_committed_golden_is_what_the_reference_produces(ref):
    mg = _make_golden()
    out = mg.run_case(ref, *mg.CASES['ragged'])
    G = dict(np.load(os.path.join(P.GOLDEN_DIR, 'ragged.npz')))
    for key in ('norm_values', 'hvg_ordered', 'contr_genes', 'comp', 'contr', 's', 'nn_idx', 'edges'):
        assert np.array_equal(np.asarray(out[key]), G[key]), key
    assert int(out['steps']) == int(G['steps']) and int(out['nmult']) == int(G['nmult'])
    assert str(out['snn_message']) == str(G['snn_message'])


def test_oracle_beside_the_reference_on_a_fresh_seed(ref):
    mg = _make_golden()
    n, g, ngenes, ncomp, k = 310, 420, 80, 8, 7
    out = mg.run_case(ref, n, g, 0.09, 77, ngenes, ncomp, k)
    import scipy.sparse as sp
    counts = sp.csr_matrix((out['counts'], out['indices'], out['indptr']), shape=(n, g))
    norm, _ = O.normalize_standard(counts)
    assert np.array_equal(norm.data, out['norm_values'])                                  # bit for bit
    mean, var = O.gene_moments(norm)
    z, bins = O.seurat_zscores(mean, var)
    P.assert_hvg_equal(O.seurat_order(z, bins, ngenes), out['hvg_ordered'], z, rtol=1e-12)   # two-gene bins: z = +-1/sqrt(2) up to an ulp
    r = O.irlb_pca(norm, out['hvg_ordered'], ncomp=ncomp, implicit=False)
    assert np.max(P.component_err(r['comp'], out['comp'])) < 1e-8                         # same algorithm, dense operator
    assert (r['steps'], r['nmult']) == (int(out['steps']), int(out['nmult']))
    P.assert_knn_equal(O.knn(out['comp'], k), out['nn_idx'], out['comp'])
    e, pruned, total = O.snn(out['nn_idx'], k, 0.067)
    assert np.array_equal(e, out['edges'])
