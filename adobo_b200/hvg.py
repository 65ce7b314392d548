"""adobo.hvg drop-in for the GPU path: find_hvg(method='seurat') and seurat()
(adobo/hvg.py:31-72, 402-487)."""
import sys

import numpy as np

from . import engine as _engine
from .data import frame_to_csr
from .normalize import ResidentKey


def _ensure_resident(eng, data):
    key = eng.resident_key
    if not (isinstance(key, ResidentKey) and key.matches(data)):
        eng.upload_normalized(frame_to_csr(data), key=ResidentKey(data))


def seurat(data, ngenes=1000, num_bins=20, eng=None):
    """hvg.seurat (hvg.py:31-72): `data` genes x cells normalized frame -> array of gene names."""
    eng = eng or _engine.default_engine()
    _ensure_resident(eng, data)
    eng.gene_moments()
    idx, _ = eng.hvg_seurat(ngenes, num_bins)
    return np.array(data.index[idx])


def find_hvg(obj, method='seurat', normalization=None, ngenes=1000, fdr=0.1, use_combat=False, verbose=False,
             eng=None):
    """hvg.find_hvg (hvg.py:402-487)."""
    if not obj.norm_data:
        raise Exception('Run normalization first before running find_hvg. See here: \
https://oscar-franzen.github.io/adobo/adobo.html#adobo.normalize.norm')
    targets = {}
    norm = normalization
    if norm is None or norm == '':
        targets = obj.norm_data
    else:
        targets[norm] = obj.norm_data[norm]
    obj.delete(('clusters', 'dr'))                       # hvg.py:459
    for k in targets:
        item = targets[k]
        if verbose:
            print('Running on the %s normalization' % k)
        data = item['combat'] if use_combat else item['data']
        if method == 'seurat':
            hvg = seurat(data, ngenes, eng=eng)
        elif method in ('brennecke', 'scran', 'chen2016', 'mm'):
            raise Exception('adobo_b200 accelerates method="seurat" only; use adobo.hvg for "%s".' % method)
        else:
            raise Exception('Unknown HVG method specified. Valid choices are: seurat, \
brennecke, scran, chen2016 and mm')
        obj.norm_data[k]['hvg'] = {'genes': hvg, 'method': method}
    obj.set_assay(sys._getframe().f_code.co_name, method)
