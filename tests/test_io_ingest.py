"""CSR-native ingest and the partitioner hand-off, against the unmodified reference
(golden frame frozen by tests/golden/make_io_golden.py; live comparison when /root/reference is present)."""
import importlib
import os

import numpy as np
import pandas as pd
import pytest

from adobo_b200 import IO
from adobo_b200.data import dataset, frame_to_csr

HERE = os.path.dirname(os.path.abspath(__file__))
ARCHIVE = os.path.join(HERE, 'golden', 'mm_small.tar.gz')


def test_load_matrix_market_equals_reference_frame():
    exp = np.load(os.path.join(HERE, 'golden', 'mm_small_expected.npz'), allow_pickle=False)
    df = IO.load_matrix_market(ARCHIVE)
    assert all(isinstance(t, pd.SparseDtype) for t in df.dtypes.unique())      # never densified
    assert list(df.index) == list(exp['genes']) and list(df.columns) == list(exp['cells'])
    assert df.index[3] == 'GENE3'                                               # second field, quotes removed (IO.py:253-262)
    assert np.array_equal(np.asarray(df.sparse.to_dense().values), exp['values'])


def test_dataset_from_matrix_market_is_csr_native():
    exp = np.load(os.path.join(HERE, 'golden', 'mm_small_expected.npz'), allow_pickle=False)
    obj = dataset.from_matrix_market(ARCHIVE)
    csr = frame_to_csr(obj.count_data)                                          # cells x genes, what the GPU path uploads
    assert np.array_equal(csr.toarray(), exp['values'].T)
    assert np.array_equal(obj.meta_cells['total_reads'].values, exp['values'].sum(axis=0))     # data.py:76
    assert np.array_equal(obj.meta_genes['expressed'].values, (exp['values'] > 0).sum(axis=1))  # data.py:83-84
    assert obj.input_file == ARCHIVE


def test_missing_and_malformed_archives_raise_like_the_reference(tmp_path):
    with pytest.raises(Exception, match='not found'):
        IO.load_matrix_market(str(tmp_path / 'nope.tar.gz'))
    import tarfile
    bad = tmp_path / 'bad.tar.gz'
    with tarfile.open(bad, 'w:gz') as tar:
        tar.add(os.path.join(HERE, 'golden', 'make_io_golden.py'), arcname='matrix.mtx.gz')
    with pytest.raises(Exception, match='exactly three files'):
        IO.load_matrix_market(str(bad))


def test_live_reference_when_present():
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip('reference checkout absent')
    ref_loader.load()
    ref = importlib.import_module('adobo.IO').load_matrix_market(ARCHIVE)
    ours = IO.load_matrix_market(ARCHIVE)
    assert list(ref.index) == list(ours.index) and list(ref.columns) == list(ours.columns)
    assert np.array_equal(np.asarray(ref.values), np.asarray(ours.sparse.to_dense().values))


def test_edge_arrays_is_the_reference_graph_construction():
    """clustering.py:154-162: vertices = distinct source nodes, edges = the rows as tuples."""
    import pandas as pd
    from adobo_b200.clustering import edge_arrays
    g = pd.DataFrame({'source_node': [0, 0, 1, 2, 2, 2, 4], 'target_node': [0, 2, 1, 0, 2, 4, 4]})
    n, edges = edge_arrays(g)
    ll = [tuple(i) for i in g.itertuples(index=False)]                 # the reference's loop
    assert n == len(set(g[g.columns[0]])) == 4
    assert edges.dtype == np.int32 and edges.flags['C_CONTIGUOUS'] and [tuple(e) for e in edges.tolist()] == ll
