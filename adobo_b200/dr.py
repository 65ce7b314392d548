"""adobo.dr drop-in for the GPU path: pca(method='irlb') and irlb()
(adobo/dr.py:110-172, 236-340)."""
import sys

import numpy as np
import pandas as pd

from . import engine as _engine
from .hvg import _ensure_resident


def irlb(data_norm, scale=True, ncomp=75, var_weigh=True, seed=None, eng=None):
    """dr.irlb (dr.py:110-172): `data_norm` genes x cells frame (already restricted to the genes
    to use).  Returns (comp cells x ncomp, contr genes x ncomp) as frames."""
    eng = eng or _engine.default_engine()
    _ensure_resident(eng, data_norm)
    r = eng.irlb(np.arange(data_norm.shape[0]), ncomp=ncomp, maxit=1000, seed=seed, scale=scale,
                 var_weigh=var_weigh)
    comp = pd.DataFrame(r['comp'], index=data_norm.columns)
    contr = pd.DataFrame(r['contr'], index=data_norm.index)
    return comp, contr


def pca(obj, method='irlb', normalization=None, ncomp=75, genes='hvg', scale=True, var_weigh=True,
        use_combat=False, verbose=False, seed=42, eng=None):
    """dr.pca (dr.py:236-340)."""
    if not obj.norm_data:
        raise Exception('Run normalization first before running pca. See here: \
https://oscar-franzen.github.io/adobo/adobo.html#adobo.normalize.norm')
    if not isinstance(genes, (str, list)):
        raise ValueError('"genes" can be a string or list, see help pages.')
    if isinstance(genes, str):
        if genes not in ('hvg', 'all'):
            raise ValueError('If "genes" is a string, then allowed values are \
"hvg" (recommended) or "all".')
    targets = {}
    if normalization is None or normalization == '':
        targets = obj.norm_data
    else:
        targets[normalization] = obj.norm_data[normalization]
    obj.delete(('clusters', 'dr'))                       # dr.py:305
    eng = eng or _engine.default_engine()
    for k in targets:
        item = targets[k]
        data = item['combat'] if use_combat else item['data']
        if isinstance(genes, str) and genes == 'hvg':
            try:
                hvg = item['hvg']['genes']
            except KeyError:
                raise Exception('Run adobo.dr.find_hvg() first.')
            keep = data.index.isin(hvg)                  # dr.py:319: original gene order is kept
        elif isinstance(genes, list):
            keep = data.index.isin(genes)
        else:
            keep = np.ones(data.shape[0], dtype=bool)
        gene_idx = np.flatnonzero(keep)
        if verbose:
            v = (method, k, '{:,}'.format(gene_idx.shape[0]), '{:,}'.format(data.shape[1]))
            print('Running PCA (method=%s) on the %s normalization (dimensions \
%s genes x %s cells)' % v)
        if method == 'irlb':
            _ensure_resident(eng, data)                  # whole normalization stays on the device
            r = eng.irlb(gene_idx, ncomp=ncomp, maxit=1000, seed=seed, scale=scale, var_weigh=var_weigh)
            comp = pd.DataFrame(r['comp'], index=data.columns)
            contr = pd.DataFrame(r['contr'], index=data.index[gene_idx])
        elif method == 'svd':
            raise Exception('adobo_b200 accelerates method="irlb" only; use adobo.dr.svd for the dense SVD.')
        else:
            raise Exception('Unkown PCA method spefified. Valid choices are: irlb and svd')
        comp.index = data.columns
        obj.norm_data[k]['dr']['pca'] = {'comp': comp, 'contr': contr, 'method': method}
        if verbose:
            print('saving %s components' % ncomp)
        obj.set_assay(sys._getframe().f_code.co_name, method)
