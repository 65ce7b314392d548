"""CSR-native ingest for the hot path.

`load_matrix_market` reads the same 10x-style archive as adobo.IO.load_matrix_market
(adobo/IO.py:204-266: a .tar.gz holding barcodes.tsv.gz, genes.tsv.gz and matrix.mtx.gz) and
returns the same genes x cells frame -- same labels, same values -- but never densifies
(the reference calls `.todense()` at IO.py:238, which is what makes a 68 k-cell matrix
unloadable): the frame has `Sparse[int64, 0]` columns, which is the layout `data.dataset`
keeps anyway (data.py:67) and which `frame_to_csr` turns into the cells x genes CSR the
GPU path uploads without a copy of the values.  `read_matrix_market_csr` skips pandas
altogether for matrices that are too large for a frame.
"""
import gzip
import os
import tarfile
import tempfile

import numpy as np
import pandas as pd
import scipy.sparse as sp
from scipy.io import mmread

_MEMBERS = ('barcodes.tsv.gz', 'genes.tsv.gz', 'matrix.mtx.gz')
_ERR = 'Input file %s should contain exactly three files with the following \
file names: barcodes.tsv.gz, genes.tsv.gz, matrix.mtx.gz'


def _labels(path):
    """Second tab-separated field when there are several and the first is not empty, else the
    first; double quotes removed (IO.py:240-262)."""
    out = []
    with gzip.open(path, 'rt') as f:
        for z in f.read().splitlines():
            z = z.replace('"', '').split('\t')
            if len(z) > 1:
                if z[0] != '':
                    out.append(z[1])
            else:
                out.append(z[0])
    return out


def read_matrix_market_csr(filename):
    """-> (counts CSR cells x genes, gene labels, cell labels); no dense or pandas step."""
    if not os.path.exists(filename):
        raise Exception('%s not found' % filename)
    with tarfile.open(filename, 'r:gz') as tar, tempfile.TemporaryDirectory() as tmp:
        names = [os.path.basename(m.name) for m in tar.getmembers() if m.isfile()]
        if len(names) != 3 or any(n not in names for n in _MEMBERS):
            raise Exception(_ERR)
        for m in tar.getmembers():
            if m.isfile():
                m.name = os.path.basename(m.name)        # flat, inside the private directory
                tar.extract(m, tmp)
        with gzip.open(os.path.join(tmp, 'matrix.mtx.gz'), 'r') as f:
            gm = mmread(f)                               # genes x cells, COO
        cells = _labels(os.path.join(tmp, 'barcodes.tsv.gz'))
        genes = _labels(os.path.join(tmp, 'genes.tsv.gz'))
    gm = sp.coo_matrix(gm)
    if gm.shape != (len(genes), len(cells)):
        raise Exception('matrix.mtx.gz is %d x %d but there are %d gene and %d cell labels'
                        % (gm.shape[0], gm.shape[1], len(genes), len(cells)))
    counts = gm.T.tocsr()                                # duplicates are summed, as in todense()
    counts.sort_indices()
    return counts, genes, cells


def load_matrix_market(filename):
    """adobo.IO.load_matrix_market (IO.py:204-266) without the dense intermediate: genes x cells
    frame, index = gene symbols, columns = barcodes, `Sparse[<mtx dtype>, 0]` columns."""
    counts, genes, cells = read_matrix_market_csr(filename)
    df = pd.DataFrame.sparse.from_spmatrix(counts.T.tocsc(), index=genes, columns=cells)
    return df.astype(pd.SparseDtype(counts.dtype, 0))
