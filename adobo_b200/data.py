"""dataset container compatible with adobo.data.dataset (adobo/data.py:21-352).

The hot-path functions only need the attributes below, so a reference `adobo.data.dataset`
object can be passed to them unchanged; this class exists so that the package works without the
reference installed and so that matrices too large for a dense pandas frame can be loaded from
a scipy CSR matrix (`from_csr`).
"""
import numpy as np
import pandas as pd
import scipy.sparse as sp


class CsrFrame:
    """genes x cells matrix held as scipy CSR (cells x genes) with pandas labels -- what stands in for the pandas
    frame where a frame cannot exist.  A pandas sparse frame keeps one SparseArray object PER CELL (data.py:67):
    building it takes seconds at 20 k cells and it cannot be built at all at 68 k (18 GB dense input) or 1.3 M cells
    (SURVEY.md 8f, f1).  It offers the frame surface the hot path touches -- `index`, `columns`, `shape`,
    `drop(labels, axis=, errors=)` (clean_matrix, normalize.py:413-433), row selection by a boolean mask (dr.py:319),
    `transpose()` -- and converts to the real thing with `to_frame()` when the matrix is small enough to want one."""

    def __init__(self, csr, genes, cells):
        # an existing CSR object is kept as it is: scipy caches has_sorted_indices on the object, and the check is a
        # pass over all non-zeros (35 ms at 70 M) that a fresh wrapper object would repeat
        self._csr = csr if isinstance(csr, sp.csr_matrix) else sp.csr_matrix(csr)
        if not self._csr.has_sorted_indices:
            self._csr.sort_indices()
        self.index = pd.Index(genes)
        self.columns = pd.Index(cells)
        assert self._csr.shape == (len(self.columns), len(self.index)), 'CSR is cells x genes'

    @property
    def csr(self):
        return self._csr

    @property
    def shape(self):
        return (len(self.index), len(self.columns))

    @property
    def dtypes(self):
        return pd.Series([pd.SparseDtype(self.csr.dtype, 0)] * min(len(self.columns), 1))

    def drop(self, labels, axis=0, errors='raise'):
        labels = pd.Index(labels)
        if axis in (0, 'index'):
            keep = ~self.index.isin(labels)
            if errors == 'raise' and keep.sum() != len(self.index) - len(labels):
                raise KeyError('labels not found in axis')
            if keep.all():
                return self
            return CsrFrame(self.csr[:, np.flatnonzero(keep)], self.index[keep], self.columns)
        keep = ~self.columns.isin(labels)
        if errors == 'raise' and keep.sum() != len(self.columns) - len(labels):
            raise KeyError('labels not found in axis')
        if keep.all():
            return self
        return CsrFrame(self.csr[np.flatnonzero(keep)], self.index, self.columns[keep])

    def __getitem__(self, mask):
        """frame[bool mask over genes] -> the selected genes (dr.py:319: data[data.index.isin(hvg)])."""
        mask = np.asarray(mask)
        if mask.dtype != bool or mask.shape[0] != len(self.index):
            raise KeyError('CsrFrame supports selection of genes by a boolean mask only')
        return CsrFrame(self.csr[:, np.flatnonzero(mask)], self.index[mask], self.columns)

    def select_cells(self, mask):
        """The cells where the boolean mask is set (frame.loc[:, mask] of the reference, plotting.py:298)."""
        mask = np.asarray(mask, dtype=bool)
        return CsrFrame(self.csr[np.flatnonzero(mask)], self.index, self.columns[mask])

    def to_frame(self):
        df = pd.DataFrame.sparse.from_spmatrix(self.csr.T.tocsc(), index=self.index, columns=self.columns)
        return df.astype(pd.SparseDtype(self.csr.dtype, 0))

    def __repr__(self):
        return 'CsrFrame(%d genes x %d cells, %d non-zeros, %s)' % (self.shape + (self.csr.nnz, self.csr.dtype))


class DeviceCsrFrame(CsrFrame):
    """The normalized matrix of normalize.norm on a CsrFrame-backed dataset: labels on the host, VALUES LEFT ON THE
    DEVICE where find_hvg / pca / generate consume them (the resident matrix; normalize.ResidentKey).  The host copy
    -- 8 bytes per non-zero, 0.56 GB at 68 k cells, more than the rest of the call sequence together -- is made when
    `.csr` is first read, or, at the latest, when the engine is about to replace its resident matrix (Engine._evict),
    so the frame never loses its data.  Values are the fp64 ones adobo_normalize delivers (the kernel is re-run on the
    resident counts with the same parameters)."""

    def __init__(self, eng, params, genes, cells):
        import weakref
        self._csr = None
        self._eng = eng
        self._params = params                            # (scaling_factor, log mode, small_const)
        self._shape_csr = (len(cells), len(genes))
        self.nnz = eng.nnz
        self.index = pd.Index(genes)
        self.columns = pd.Index(cells)
        eng._lazy = [r for r in eng._lazy if r() is not None] + [weakref.ref(self)]

    @property
    def materialized(self):
        return self._csr is not None

    def materialize(self):
        if self._csr is None:
            eng, self._eng = self._eng, None
            if eng is None or eng.h is None:
                raise Exception('DeviceCsrFrame: the engine that held the normalized values was closed')
            from .normalize import ResidentKey
            key = eng.resident_key
            mine = isinstance(key, ResidentKey) and key.matches(self)
            self._csr, _ = eng.normalize(*self._params)  # re-runs K1 on the resident counts: the device values are now these
            eng.resident_key = ResidentKey(self) if mine else None
            eng._lazy = [r for r in eng._lazy if r() is not None and r() is not self]
        return self._csr

    @property
    def csr(self):
        return self.materialize()

    def __repr__(self):
        return 'DeviceCsrFrame(%d genes x %d cells, %d non-zeros, %s)' % (
            self.shape + (self.nnz, 'on the host' if self.materialized else 'values on the device'))


def _row_sums(csr):
    """Row sums of a CSR matrix in int64 / float64 without the dtype-converted copy of `data` that scipy's sum makes
    (0.2 s at 70 M non-zeros): one reduceat over the row starts, empty rows set to zero afterwards."""
    n = csr.shape[0]
    out = np.zeros(n, dtype=np.int64 if np.issubdtype(csr.dtype, np.integer) else np.float64)
    if csr.nnz == 0 or n == 0:
        return out
    lens = np.diff(csr.indptr)
    rows = np.flatnonzero(lens > 0)
    out[rows] = np.add.reduceat(csr.data, csr.indptr[rows], dtype=out.dtype)
    return out


class dataset:
    def __init__(self, raw_mat, desc='no desc set', output_file=None, input_file=None, sparse=True, verbose=False,
                 eng=None):
        """adobo.data.dataset (data.py:54-93).  `eng` (extension; default: the process-wide engine IF one exists):
        for a CsrFrame-backed dataset the counts are uploaded once, here, and the summary statistics come from the
        device (adobo_count_stats); normalize.norm then finds the counts resident.  Without an engine, and for pandas
        frames, the statistics are computed on the host like the reference does."""
        self.sparse = sparse
        self.desc = desc
        self.output_file = output_file
        self.input_file = input_file
        self._assays = {}
        if isinstance(raw_mat, CsrFrame):
            self.count_data = raw_mat                                        # no pandas frame at all
        elif sparse and not all(isinstance(t, pd.SparseDtype) for t in raw_mat.dtypes.unique()):
            self.count_data = raw_mat.astype(pd.SparseDtype('int', 0))      # data.py:67
        else:
            self.count_data = raw_mat
        self.imp_count_data = pd.DataFrame()
        self._norm_data = {}
        csr = frame_to_csr(self.count_data)                                  # cells x genes
        if eng is None and isinstance(raw_mat, CsrFrame):
            from . import engine as _engine
            eng = _engine._default
        if eng is not None and isinstance(raw_mat, CsrFrame) and eng.world == 1:
            from .normalize import ResidentKey
            eng.upload_counts(csr, key=ResidentKey(raw_mat))
            total, detected, expressed = eng.count_stats()
        else:
            total = _row_sums(csr)
            allpos = csr.nnz == 0 or csr.data.min() > 0
            detected = np.diff(csr.indptr).astype(np.int64) if allpos else _row_sums(sp.csr_matrix(
                ((csr.data > 0).astype(np.int64), csr.indices, csr.indptr), shape=csr.shape))
            expressed = np.bincount(csr.indices if allpos else csr.indices[csr.data > 0], minlength=raw_mat.shape[0])
        self.meta_cells = pd.DataFrame(index=raw_mat.columns)                # data.py:75-78
        self.meta_cells['total_reads'] = total
        self.meta_cells['status'] = ['OK'] * raw_mat.shape[1]
        self.meta_cells['detected_genes'] = detected
        self.meta_genes = pd.DataFrame(index=raw_mat.index)                  # data.py:80-89
        self.meta_genes['expressed'] = expressed
        self.meta_genes['expressed_perc'] = expressed / raw_mat.shape[1] * 100
        self.meta_genes['status'] = ['OK'] * raw_mat.shape[0]
        self.meta_genes['mitochondrial'] = [None] * raw_mat.shape[0]
        self.meta_genes['ERCC'] = [None] * raw_mat.shape[0]
        self.version = 'adobo_b200'

    @classmethod
    def from_csr(cls, counts, genes=None, cells=None, frame='auto', **kw):
        """counts: scipy CSR cells x genes.  frame=True builds the pandas sparse frame the reference holds
        (data.py:67); frame=False keeps the matrix as a CsrFrame -- no per-cell pandas objects, the only way to hold
        BASELINE's 68 k- and 1.3 M-cell shapes; 'auto' = a pandas frame up to 10,000 cells."""
        if not isinstance(counts, sp.csr_matrix):
            counts = sp.csr_matrix(counts)
        genes = genes if genes is not None else ['g%d' % i for i in range(counts.shape[1])]
        cells = cells if cells is not None else ['c%d' % i for i in range(counts.shape[0])]
        if frame == 'auto':
            frame = counts.shape[0] <= 10000
        cf = CsrFrame(counts, genes, cells)
        return cls(cf.to_frame() if frame else cf, **kw)

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
    """genes x cells frame (pandas sparse or dense, or a CsrFrame) -> scipy CSR cells x genes.  A pandas sparse
    frame stores one SparseArray per cell, i.e. it already is CSR(cells x genes)."""
    if isinstance(df, CsrFrame):
        return df.csr
    if len(df.dtypes) and all(isinstance(t, pd.SparseDtype) for t in df.dtypes.unique()):
        m = df.sparse.to_coo().T.tocsr()
    else:
        m = sp.csr_matrix(np.asarray(df.values).T)
    m.sort_indices()
    return m


def csr_to_frame(m, genes, cells, like=None):
    """scipy CSR cells x genes (float64) -> genes x cells frame with Sparse[float64, 0] columns
    (normalize.py:567-568); a CsrFrame when `like` is one."""
    if isinstance(like, CsrFrame):
        return CsrFrame(m, genes, cells)
    df = pd.DataFrame.sparse.from_spmatrix(sp.csc_matrix(m.T), index=genes, columns=cells)
    return df.astype(pd.SparseDtype('float64', 0))
