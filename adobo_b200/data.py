"""dataset container compatible with adobo.data.dataset (adobo/data.py:21-352).

The hot-path functions only need the attributes below, so a reference `adobo.data.dataset`
object can be passed to them unchanged; this class exists so that the package works without the
reference installed and so that matrices too large for a dense pandas frame can be loaded from
a scipy CSR matrix (`from_csr`).
"""
import numpy as np
import pandas as pd
import scipy.sparse as sp


class dataset:
    def __init__(self, raw_mat, desc='no desc set', output_file=None, input_file=None, sparse=True, verbose=False):
        self.sparse = sparse
        self.desc = desc
        self.output_file = output_file
        self.input_file = input_file
        self._assays = {}
        if sparse and not all(isinstance(t, pd.SparseDtype) for t in raw_mat.dtypes.unique()):
            self.count_data = raw_mat.astype(pd.SparseDtype('int', 0))      # data.py:67
        else:
            self.count_data = raw_mat
        self.imp_count_data = pd.DataFrame()
        self._norm_data = {}
        csr = frame_to_csr(self.count_data)                                  # cells x genes
        self.meta_cells = pd.DataFrame(index=raw_mat.columns)                # data.py:75-78
        self.meta_cells['total_reads'] = np.asarray(csr.sum(axis=1)).ravel()
        self.meta_cells['status'] = ['OK'] * raw_mat.shape[1]
        self.meta_cells['detected_genes'] = np.diff(csr.indptr)
        self.meta_genes = pd.DataFrame(index=raw_mat.index)                  # data.py:80-89
        expressed = np.bincount(csr.indices[csr.data > 0], minlength=raw_mat.shape[0])
        self.meta_genes['expressed'] = expressed
        self.meta_genes['expressed_perc'] = expressed / raw_mat.shape[1] * 100
        self.meta_genes['status'] = ['OK'] * raw_mat.shape[0]
        self.meta_genes['mitochondrial'] = [None] * raw_mat.shape[0]
        self.meta_genes['ERCC'] = [None] * raw_mat.shape[0]
        self.version = 'adobo_b200'

    @classmethod
    def from_csr(cls, counts, genes=None, cells=None, **kw):
        """counts: scipy CSR cells x genes."""
        counts = sp.csr_matrix(counts)
        genes = genes if genes is not None else ['g%d' % i for i in range(counts.shape[1])]
        cells = cells if cells is not None else ['c%d' % i for i in range(counts.shape[0])]
        df = pd.DataFrame.sparse.from_spmatrix(counts.T.tocsc(), index=genes, columns=cells)
        df = df.astype(pd.SparseDtype(counts.dtype, 0))
        return cls(df, **kw)

    @classmethod
    def from_matrix_market(cls, filename, **kw):
        """10x-style archive (see IO.load_matrix_market) -> dataset, CSR all the way."""
        from .IO import read_matrix_market_csr
        counts, genes, cells = read_matrix_market_csr(filename)
        kw.setdefault('input_file', filename)
        return cls.from_csr(counts, genes=genes, cells=cells, **kw)

    @property
    def norm_data(self):
        return self._norm_data

    @norm_data.setter
    def norm_data(self, val):
        self._norm_data = val

    def get_assay(self, name, lang=False):
        if lang:
            return ('No', 'Yes')[name in self._assays]
        return name in self._assays

    def set_assay(self, name, key=1):                                        # data.py:156-158
        self._assays[name] = key

    def delete(self, what):                                                  # data.py:343-350
        try:
            for n in list(self.norm_data.keys()):
                if n == what or n in what:
                    del self.norm_data[n]
                else:
                    for k in self.norm_data[n].keys():
                        if k == what or k in what:
                            self.norm_data[n][k] = {}
        except Exception:
            pass


def frame_to_csr(df):
    """genes x cells frame (pandas sparse or dense) -> scipy CSR cells x genes.  A pandas sparse
    frame stores one SparseArray per cell, i.e. it already is CSR(cells x genes)."""
    if len(df.dtypes) and all(isinstance(t, pd.SparseDtype) for t in df.dtypes.unique()):
        m = df.sparse.to_coo().T.tocsr()
    else:
        m = sp.csr_matrix(np.asarray(df.values).T)
    m.sort_indices()
    return m


def csr_to_frame(m, genes, cells):
    """scipy CSR cells x genes (float64) -> genes x cells frame with Sparse[float64, 0] columns
    (normalize.py:567-568)."""
    df = pd.DataFrame.sparse.from_spmatrix(sp.csc_matrix(m.T), index=genes, columns=cells)
    return df.astype(pd.SparseDtype('float64', 0))
