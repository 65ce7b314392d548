"""CPU-side checks: the C-ABI library loads and exports every symbol the header declares, the
host-side plumbing (frame <-> CSR, clean_matrix, argument checking) behaves like the reference,
and computing calls fail loudly without a GPU."""
import os
import re

import numpy as np
import pandas as pd
import pytest
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_cuda():
    import torch
    return torch.cuda.is_available()


def test_library_exports_every_declared_symbol():
    from adobo_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(ROOT, 'include', 'adobo_b200.h')).read()
    declared = set(re.findall(r'\b(adobo_[a-z0-9_]+)\s*\(', header))
    declared -= {'adobo_ctx'}
    assert declared, 'no declarations parsed'
    for name in sorted(declared):
        assert hasattr(lib, name), 'missing symbol ' + name
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    assert lib.adobo_version() >= 100


def test_no_cpu_fallback_without_gpu():
    if _has_cuda():
        pytest.skip('GPU present')
    from adobo_b200.engine import Engine
    from adobo_b200._lib import AdoboB200Error
    with pytest.raises(AdoboB200Error, match='no CPU fallback'):
        Engine(0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, 'adobo_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'import\s+oracle|from\s+oracle|oracle[./]', src), f + ' uses the oracle'


def test_frame_csr_round_trip_and_dataset():
    from adobo_b200.data import dataset, frame_to_csr, csr_to_frame
    from adobo_b200.synth import synth_counts
    counts = synth_counts(50, 70, 0.1, seed=1)
    exp = dataset.from_csr(counts)
    assert exp.count_data.shape == (70, 50)
    back = frame_to_csr(exp.count_data)
    assert (back != counts).nnz == 0
    assert np.array_equal(exp.meta_cells['total_reads'].values, np.asarray(counts.sum(axis=1)).ravel())
    f = csr_to_frame(counts.astype(np.float64), exp.count_data.index, exp.count_data.columns)
    assert str(f.dtypes.iloc[0]) == 'Sparse[float64, 0]'
    exp.norm_data['x'] = {'data': f, 'dr': {'pca': 1}, 'clusters': {'leiden': 2}, 'hvg': {}}
    exp.delete(('clusters', 'dr'))
    assert exp.norm_data['x']['dr'] == {} and exp.norm_data['x']['clusters'] == {}


def test_clean_matrix_masks_like_reference():
    from adobo_b200.data import dataset
    from adobo_b200.normalize import clean_matrix
    from adobo_b200.synth import synth_counts
    exp = dataset.from_csr(synth_counts(20, 30, 0.2, seed=2))
    exp.meta_cells.loc[exp.meta_cells.index[3], 'status'] = 'EXCLUDE'
    exp.meta_genes.loc[exp.meta_genes.index[5], 'status'] = 'EXCLUDE'
    exp.meta_genes.loc[exp.meta_genes.index[7], 'mitochondrial'] = True
    out = clean_matrix(exp.count_data, exp)
    assert out.shape == (28, 19)
    assert 'c3' not in out.columns and 'g5' not in out.index and 'g7' not in out.index


def test_argument_errors_match_reference_text():
    import adobo_b200 as ad
    from adobo_b200.synth import synth_counts
    exp = ad.data.dataset.from_csr(synth_counts(20, 30, 0.2, seed=2))
    with pytest.raises(Exception, match='Run normalization first before running find_hvg'):
        ad.hvg.find_hvg(exp)
    with pytest.raises(Exception, match='Run normalization first before running pca'):
        ad.dr.pca(exp)
    with pytest.raises(Exception, match='Supported community detection algorithms'):
        ad.clustering.generate(exp, clust_alg='nope')
    with pytest.raises(Exception, match='Unknown normalization method'):
        ad.normalize.norm(exp, method='bogus')
    with pytest.raises(Exception, match='gene_lengths'):
        ad.normalize.norm(exp, method='rpkm')
    with pytest.raises(Exception, match='small_const'):
        ad.normalize.norm(exp, small_const=2)


def test_synth_blocks_are_shard_reproducible():
    from adobo_b200.synth import synth_counts
    full = synth_counts(2500, 120, 0.08, seed=4)
    part = synth_counts(2500, 120, 0.08, seed=4, row_start=1000, row_stop=2300)
    assert (full[1000:2300] != part).nnz == 0
    assert full.indices.dtype == np.int32 and full.data.dtype == np.int32
    assert np.all(np.asarray(full.sum(axis=1)).ravel() > 0)


def test_csr_backed_dataset_matches_frame_backed():
    """f1: a dataset without any pandas frame (CsrFrame) carries the same metadata and masks as the frame-backed one."""
    from adobo_b200.data import dataset, CsrFrame, frame_to_csr
    from adobo_b200.normalize import clean_matrix, ResidentKey
    from adobo_b200.synth import synth_counts
    counts = synth_counts(40, 60, 0.15, seed=3)
    a = dataset.from_csr(counts, frame=True)
    b = dataset.from_csr(counts, frame=False)
    assert isinstance(b.count_data, CsrFrame) and b.count_data.shape == a.count_data.shape == (60, 40)
    assert list(b.count_data.index) == list(a.count_data.index) and list(b.count_data.columns) == list(a.count_data.columns)
    pd.testing.assert_frame_equal(a.meta_cells, b.meta_cells)
    pd.testing.assert_frame_equal(a.meta_genes, b.meta_genes)
    for exp in (a, b):
        exp.meta_cells.loc[exp.meta_cells.index[3], 'status'] = 'EXCLUDE'
        exp.meta_genes.loc[exp.meta_genes.index[5], 'status'] = 'EXCLUDE'
        exp.meta_genes.loc[exp.meta_genes.index[7], 'mitochondrial'] = True
    ca, cb = clean_matrix(a.count_data, a), clean_matrix(b.count_data, b)
    assert cb.shape == ca.shape == (58, 39) and list(cb.index) == list(ca.index) and list(cb.columns) == list(ca.columns)
    assert (frame_to_csr(ca) != frame_to_csr(cb)).nnz == 0
    sel = cb[cb.index.isin(['g1', 'g9', 'g11'])]
    assert sel.shape == (3, 39) and (sel.csr != cb.csr[:, cb.index.get_indexer(['g1', 'g9', 'g11'])]).nnz == 0
    assert (frame_to_csr(cb.to_frame()) != cb.csr).nnz == 0
    # resident-matrix key: identity, not id(); content changes are seen
    k = ResidentKey(cb)
    assert k.matches(cb) and not k.matches(CsrFrame(cb.csr.copy(), cb.index, cb.columns))
    cb.csr.data[0] += 1
    assert not k.matches(cb)
    fa = ca
    kf = ResidentKey(fa)
    assert kf.matches(fa) and not kf.matches(fa.copy())


class _StubEngine:
    """Host-logic stand-in for engine.Engine (no GPU): keeps the 'resident' counts, normalizes with the oracle, counts
    what the host layer asks for.  Only the bookkeeping of norm() / dataset() / DeviceCsrFrame is under test here."""

    def __init__(self):
        from adobo_b200.engine import Engine
        self.world, self.rank, self.h = 1, 0, object()
        self.resident_key = self.counts_key = None
        self._lazy, self._pattern = [], None
        self.uploads = self.norm_calls = self.value_downloads = 0
        self._evict = Engine._evict.__get__(self)            # the real eviction logic

    def upload_counts(self, counts, key=None):
        self._evict()
        self.counts = sp.csr_matrix(counts).copy()
        self.n, self.g, self.nnz = self.counts.shape[0], self.counts.shape[1], self.counts.nnz
        self.uploads += 1
        self.resident_key, self.counts_key = None, key

    def upload_normalized(self, norm, key=None):
        self._evict()
        self.counts = None
        self.resident_key = key

    def count_stats(self):
        c = self.counts
        return (np.asarray(c.sum(axis=1)).ravel().astype(np.int64), np.asarray((c > 0).sum(axis=1)).ravel().astype(np.int64),
                np.asarray((c > 0).sum(axis=0)).ravel().astype(np.int64))

    def normalize(self, scaling_factor=10000, log_mode=1, small_const=1.0, want_values=True):
        from oracle import adobo_oracle as O
        assert self.counts is not None, 'normalize without resident counts'
        self.norm_calls += 1
        lib = np.asarray(self.counts.sum(axis=1)).ravel()
        if not want_values:
            return None, lib
        self.value_downloads += 1
        vals, _ = O.normalize_standard(self.counts, scaling_factor, log=(log_mode != 0))
        return vals, lib


def test_device_csr_frame_bookkeeping():
    """A CsrFrame-backed dataset made with an engine: one upload (dataset), none in norm(); the normalized values stay
    'on the device' until read; a second normalization keeps the first frame valid; replacing the resident matrix
    brings every lazy frame to the host first; a materialised frame that was the resident one stays the resident one."""
    from adobo_b200.data import dataset, DeviceCsrFrame
    from adobo_b200 import normalize as N
    from adobo_b200.synth import synth_counts
    from oracle import adobo_oracle as O
    counts = synth_counts(50, 70, 0.2, seed=8)
    eng = _StubEngine()
    exp = dataset.from_csr(counts, frame=False, eng=eng)
    assert eng.uploads == 1 and eng.counts_key.matches(exp.count_data)
    assert np.array_equal(exp.meta_cells['total_reads'].values, np.asarray(counts.sum(axis=1)).ravel())
    N.norm(exp, eng=eng)
    N.norm(exp, name='nolog', log=False, eng=eng)
    assert eng.uploads == 1 and eng.value_downloads == 0             # counts found resident, no values copied
    d1, d2 = exp.norm_data['standard']['data'], exp.norm_data['nolog']['data']
    assert isinstance(d1, DeviceCsrFrame) and not d1.materialized and d1.shape == (70, 50)
    assert eng.resident_key.matches(d2) and not eng.resident_key.matches(d1)
    assert 'values on the device' in repr(d1)
    v2 = d2.csr                                                      # read: one download, still the resident frame
    assert eng.value_downloads == 1 and d2.materialized and eng.resident_key.matches(d2)
    assert np.array_equal(v2.data, O.normalize_standard(counts, 10000, log=False)[0].data)
    other = dataset.from_csr(synth_counts(20, 70, 0.2, seed=9), frame=False, eng=eng)   # replaces the resident counts
    assert d1.materialized and eng.value_downloads == 2 and eng.uploads == 2
    assert np.array_equal(d1.csr.data, O.normalize_standard(counts, 10000)[0].data)
    assert not eng.counts_key.matches(exp.count_data) and eng.counts_key.matches(other.count_data)
    N.norm(exp, name='again', eng=eng)                               # the first dataset again: its counts are uploaded again
    assert eng.uploads == 3
    # without an engine the statistics are computed on the host (pandas-frame datasets always are): same numbers
    host = dataset.from_csr(counts, frame=False)
    pd.testing.assert_frame_equal(host.meta_cells, exp.meta_cells)
    pd.testing.assert_frame_equal(host.meta_genes, exp.meta_genes)


def test_partitioner_hand_off_builds_the_reference_graph(monkeypatch):
    """f2: `_igraph_from_edges` (arrays straight into igraph) executed against a recording stand-in for igraph --
    the package is absent here -- and compared with what the reference's construction hands over
    (clustering.py:154-162: vertex count from the first column, 1-based names, edges as tuples from itertuples)."""
    import types
    from adobo_b200 import clustering

    class Graph:
        def __init__(self):
            self.n = 0
            self.vs = {}
            self.edges = None

        def add_vertices(self, n):
            self.n += n

        def add_edges(self, e):
            self.edges = [tuple(int(v) for v in row) for row in e]

    monkeypatch.setitem(__import__('sys').modules, 'igraph', types.SimpleNamespace(Graph=Graph))
    rng = np.random.default_rng(0)
    src = np.sort(rng.integers(0, 50, 400))
    snn_graph = pd.DataFrame({'source_node': src, 'target_node': rng.integers(0, 50, 400)})
    g = clustering._igraph_from_edges(snn_graph)
    # the reference: nodes = unique(first column); names 1..n; edges = [(row[0], row[1]) for row in itertuples]
    nodes = np.unique(snn_graph[snn_graph.columns[0]])
    ref_edges = [(int(r[0]), int(r[1])) for r in snn_graph.itertuples(index=False)]
    assert g.n == len(nodes) and g.vs['name'] == list(range(1, len(nodes) + 1)) and g.edges == ref_edges
