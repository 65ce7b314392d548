"""adobo.normalize drop-in for the GPU path: norm(method='standard') and standard().

Signatures, defaults, error messages and side effects follow adobo/normalize.py:256-287, 413-580.
The other normalizers of the reference (vsn, clr, rpkm, fqn) are host statistics outside the
accelerated path and are not provided here.
"""
import sys

import numpy as np
import pandas as pd

from . import engine as _engine
from .data import frame_to_csr, csr_to_frame, CsrFrame, DeviceCsrFrame

_LOG_MODES = {np.log2: (_engine.LOG_2, 0.0), np.log: (_engine.LOG_E, 0.0), np.log10: (_engine.LOG_10, 0.0),
              np.log1p: (_engine.LOG_E, 1.0)}


def clean_matrix(data, obj, remove_low_qual=True, remove_mito=True, meta=False):
    """normalize.py:413-433 (frame in, frame out; cheap metadata masking, stays on the host)."""
    if remove_low_qual:
        remove = obj.meta_cells.status[obj.meta_cells.status != 'OK']
        data = data.drop(remove.index, axis=1, errors='ignore')
        v = np.logical_and(obj.meta_genes.status != 'OK', obj.meta_genes.ERCC != True)  # noqa: E712
        remove = obj.meta_genes.status[v]
        data = data.drop(remove.index, axis=0, errors='ignore')
    if remove_mito:
        remove = obj.meta_genes[obj.meta_genes.mitochondrial == True]  # noqa: E712
        data = data.drop(remove.index, axis=0, errors='ignore')
    if meta:
        md = obj.meta_cells.copy()
        md = md.loc[md.index.isin(data.columns), :]
        md = md.reindex(data.columns)
        return data, md
    return data


def _log_mode(log, log_func, small_const):
    if not log:
        return _engine.LOG_NONE, 1.0
    if log_func not in _LOG_MODES:
        raise Exception('adobo_b200: log_func must be one of np.log2, np.log, np.log10, np.log1p')
    mode, extra = _LOG_MODES[log_func]
    sc = float(small_const) + extra
    if sc != 1.0:
        # log(0 + c) != 0 would turn every implicit zero of the sparse frame into a stored value
        raise Exception('adobo_b200: small_const must keep log_func(0 + small_const) == 0 (sparse output)')
    return mode, sc


def standard(data, scaling_factor=10000, eng=None):
    """normalize.standard (normalize.py:256-287): per-cell library-size scaling, no log.
    `data`: genes x cells frame; returns a frame of the same shape."""
    eng = eng or _engine.default_engine()
    counts = frame_to_csr(data)
    eng.upload_counts(counts)
    vals, _ = eng.normalize(scaling_factor, _engine.LOG_NONE, 1.0)
    return csr_to_frame(vals, data.index, data.columns, like=data)


def norm(obj, method='standard', name=None, use_imputed=False, log=True, log_func=np.log2, small_const=1,
         remove_low_qual=True, remove_mito=True, gene_lengths=None, scaling_factor=10000, axis='genes',
         ngenes=2000, nworkers='auto', retx=False, verbose=False, eng=None):
    """normalize.norm (normalize.py:436-580) for method='standard'."""
    if name is None or name == '':
        name = method
    if method == 'rpkm' and gene_lengths is None:
        raise Exception('The `gene_lengths` parameter needs to be set when method is RPKM.')
    if use_imputed:
        if obj.imp_count_data.shape[0] == 0:
            raise Exception('No imputed data found. Run adobo.preproc.impute() first.')
        data = obj.imp_count_data
    else:
        data = obj.count_data
    data = clean_matrix(data, obj)                      # NB: reference ignores remove_* here too (normalize.py:541)
    if method != 'standard':
        if method in ('rpkm', 'fqn', 'clr', 'vsn'):
            raise Exception('adobo_b200 accelerates method="standard" only; use adobo.normalize for "%s".' % method)
        raise Exception('Unknown normalization method.')
    mode, sc = _log_mode(log, log_func, small_const)
    eng = eng or _engine.default_engine()
    ck = eng.counts_key
    if not (isinstance(ck, ResidentKey) and ck.matches(data)):      # dataset(..., eng) left these counts on the device
        eng.upload_counts(frame_to_csr(data), key=ResidentKey(data) if isinstance(data, CsrFrame) else None)
    if isinstance(data, CsrFrame):
        # no pandas frame on this path: the values stay where find_hvg / pca / generate read them (DeviceCsrFrame);
        # library sizes only come back
        ck = eng.counts_key
        eng.normalize(scaling_factor, mode, sc, want_values=False)
        eng.counts_key = ck
        normf = DeviceCsrFrame(eng, (scaling_factor, mode, sc), data.index, data.columns)
    else:
        vals, _ = eng.normalize(scaling_factor, mode, sc)
        normf = csr_to_frame(vals, data.index, data.columns, like=data)
    ne = None
    ercc = obj.meta_genes.ERCC.reindex(normf.index)
    if np.any(ercc == True):  # noqa: E712             # normalize.py:562-566
        flag = (ercc == True).values  # noqa: E712
        ne = normf[flag]
        normf = normf[np.logical_not(flag)]
    elif not getattr(obj, 'sparse', True) and not isinstance(normf, CsrFrame):
        normf = normf.sparse.to_dense()
    obj.norm_data[name] = {'data': normf, 'method': method, 'log': log, 'norm_ercc': ne, 'dr': {},
                           'clusters': {}, 'slingshot': {}, 'de': {}, 'combat': {}}
    if ne is None:
        eng.resident_key = ResidentKey(normf)         # the matrix on the device is this normalization
    obj.set_assay(sys._getframe().f_code.co_name, 'standard')
    if retx:
        return normf


class ResidentKey:
    """Identifies the normalized frame whose matrix is resident on the device, so that find_hvg / pca after norm
    skip the upload.  A WEAK reference compared with `is` (CPython re-uses id()s of collected frames: a fresh frame
    of the same shape -- dr.jackstraw makes one per permutation -- must never be taken for the resident one) plus a
    content fingerprint of a fixed sample of cells, which catches in-place edits of norm_data[..]['data']."""

    def __init__(self, df):
        import weakref
        try:
            self.ref = weakref.ref(df)
        except TypeError:
            self.ref = None
        self.fp = _fingerprint(df)

    def matches(self, df):
        return self.ref is not None and self.ref() is df and self.fp == _fingerprint(df)


def _fingerprint(df):
    if isinstance(df, DeviceCsrFrame) and not df.materialized:
        return ('device', df.shape, df.nnz)              # nothing on the host that could have been edited
    if isinstance(df, CsrFrame):
        m = df.csr
        pos = np.linspace(0, max(m.nnz - 1, 0), num=min(m.nnz, 64), dtype=np.int64)
        return (df.shape, int(m.nnz), float(m.data[pos].sum()) if m.nnz else 0.0, int(m.indices[pos].sum()) if m.nnz else 0)
    n = df.shape[1]
    cols = np.unique(np.linspace(0, max(n - 1, 0), num=min(n, 16), dtype=np.int64)) if n else []
    acc = [df.shape]
    for c in cols:
        a = df.iloc[:, c].array
        if isinstance(a, pd.arrays.SparseArray):
            acc.append((int(a.sp_index.npoints), float(np.asarray(a.sp_values).sum())))
        else:
            acc.append((len(a), float(np.asarray(a).sum())))
    return tuple(acc)
