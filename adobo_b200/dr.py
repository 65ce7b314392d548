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


# ---- dr.jackstraw (dr.py:520-654): the 500 x irlb() multiplier of the hot path ------------------------------
def _p_adjust_bh(p):
    """Benjamini-Hochberg adjusted p-values (adobo/_stats.py:133-158)."""
    p = np.asarray(p, dtype=np.float64)
    q = np.full(p.shape, np.nan)
    ok = ~np.isnan(p)
    pv = p[ok]
    m = float(pv.shape[0])
    by_descend = pv.argsort()[::-1]
    by_orig = by_descend.argsort()
    steps = m / np.arange(m, 0, -1)
    q[ok] = np.minimum(1, np.minimum.accumulate(steps * pv[by_descend]))[by_orig]
    return q


def permute_gene_columns(sub_csc, cols, rng):
    """Independent random permutation, across cells, of the given gene columns of a cells x genes CSC matrix.

    The reference permutes rows of the DENSE scaled genes x cells matrix (dr.py:606-613).  Permuting a gene across
    cells leaves its mean and standard deviation untouched, so scaling the permuted gene equals permuting the
    scaled gene: the matrix can stay sparse and unscaled, and the centring / scaling stays inside the mat-vecs of
    the GPU path.  A uniformly random permutation sends the gene's nnz non-zeros to nnz distinct cells chosen
    uniformly, each value to one of them."""
    import scipy.sparse as sp
    m = sp.csc_matrix(sub_csc, copy=True)
    n = m.shape[0]
    for j in cols:
        lo, hi = m.indptr[j], m.indptr[j + 1]
        pos = rng.choice(n, hi - lo, replace=False)          # entry i of the column moves to cell pos[i]
        order = np.argsort(pos, kind='stable')
        m.indices[lo:hi] = pos[order]
        m.data[lo:hi] = m.data[lo:hi][order]
    return m


# ---- the keyed bijection of [0, n) the DEVICE permutes with (csrc/common.cuh perm_key / perm_apply) -------------
_M64 = (1 << 64) - 1


def _splitmix(state):
    state = (state + 0x9E3779B97F4A7C15) & _M64
    z = state
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return state, z ^ (z >> 31)


def perm_rows(n, seed, col, rows):
    """Where the cells `rows` send their entry of gene position `col` under permutation key `seed`: numpy
    restatement of adobo_perm_row (four rounds of odd multiply + add and xor-shift on the smallest power-of-two
    domain holding n, cycle-walked into [0, n)) -- used by the host path of jackstraw and by the tests."""
    st = (int(seed) ^ ((0xD1B54A32D192ED03 * (int(col) + 1)) & _M64)) & _M64
    mul, add = [], []
    for _ in range(4):
        st, z = _splitmix(st)
        mul.append(z | 1)
        st, z = _splitmix(st)
        add.append(z)
    bits = 1
    while (1 << bits) < n:
        bits += 1
    mask = np.uint64((1 << bits) - 1)
    sh = np.uint64(bits // 2 if bits > 1 else 1)
    x = np.asarray(rows, dtype=np.uint64).copy()
    todo = np.ones(x.shape, dtype=bool)
    with np.errstate(over='ignore'):
        while todo.any():
            y = x[todo]
            for r in range(4):
                y = (y * np.uint64(mul[r]) + np.uint64(add[r])) & mask
                y ^= y >> sh
            x[todo] = y
            todo[todo] = y >= np.uint64(n)
    return x.astype(np.int64)


def permute_gene_columns_keyed(sub_csc, cols, seed):
    """The permutation the device applies (adobo_irlb_permuted), on a host CSC matrix: gene column j in `cols` has the
    entry of cell r moved to cell perm_rows(n, seed, j, r)."""
    import scipy.sparse as sp
    m = sp.csc_matrix(sub_csc, copy=True)
    n = m.shape[0]
    for j in cols:
        lo, hi = m.indptr[j], m.indptr[j + 1]
        pos = perm_rows(n, seed, j, m.indices[lo:hi])
        order = np.argsort(pos, kind='stable')
        m.indices[lo:hi] = pos[order]
        m.data[lo:hi] = m.data[lo:hi][order]
    return m


def jackstraw(obj, normalization=None, permutations=500, ncomp=None, subset_frac_genes=0.05, score_thr=1e-03,
              fdr=0.01, retx=True, verbose=False, seed=None, eng=None):
    """dr.jackstraw (dr.py:520-654, Chung & Storey 2015): empirical p-values of every HVG's loading on every
    component from `permutations` PCAs of the matrix with a random 5 % of the genes permuted across cells, then
    one chi-square p-value per component.  Every permutation is one IRLB through the GPU path (scale=True on the
    sparse permuted subset, see `permute_gene_columns`).  `seed` (extension) makes the random draws reproducible;
    the reference draws from the global `random` / `numpy.random` state."""
    from scipy.stats import chi2_contingency
    from .data import frame_to_csr
    if normalization is None or normalization == '':
        norm = list(obj.norm_data.keys())[-1]                # dr.py:577
    else:
        norm = normalization
    item = obj.norm_data[norm]
    try:
        loadings = np.abs(item['dr']['pca']['contr'])
    except KeyError:
        raise Exception('Run `adobo.dr.pca(...)` first.')
    X = item['data']
    if not ncomp:
        ncomp = loadings.shape[1]
    elif ncomp > loadings.shape[1]:
        raise Exception('"ncomp" is higher than the number of available \
components computed by adobo.dr.pca(...)')
    if verbose:
        print('computing for ncomp=%s' % ncomp)
    try:
        hvg = item['hvg']['genes']
    except KeyError:
        raise Exception('Run adobo.dr.find_hvg() first.')
    keep = np.flatnonzero(X.index.isin(hvg))                 # dr.py:597: original gene order
    genes = X.index[keep]
    h = keep.shape[0]
    nrand = int(round(h * subset_frac_genes))
    rng = np.random.default_rng(seed)
    eng = eng or _engine.default_engine()
    on_device = hasattr(eng, 'irlb_permuted')                # the GPU engine permutes and rebuilds the matrix itself
    if on_device:
        _ensure_resident(eng, X)
        eng.gather_subset(keep, scale=True)                  # subset + (mu, sigma) once; every permutation reuses them
    else:
        sub = frame_to_csr(X)[:, keep].tocsc()               # cells x h, sparse, unscaled
    perm_loadings = []
    for perm in range(permutations):
        if verbose:
            print('random set %s ' % perm)
        rand = np.sort(rng.choice(h, nrand, replace=False))  # dr.py:606; rows of contr come back in gene order (:616)
        pseed = int(rng.integers(0, 2 ** 63 - 1))            # key of this permutation's bijections (one per gene)
        vseed = int(rng.integers(0, 2 ** 31 - 1))            # IRLB start vector (irlb.py:151-152)
        if on_device:
            contr = eng.irlb_permuted(rand, pseed, ncomp=ncomp, maxit=1000, seed=vseed, var_weigh=True)
        else:
            eng.upload_normalized(permute_gene_columns_keyed(sub, rand, pseed).tocsr())
            contr = eng.irlb(np.arange(h), ncomp=ncomp, maxit=1000, seed=vseed, scale=True, var_weigh=True)['contr']
        perm_loadings.append(np.abs(contr[rand, :ncomp]))
    null = np.concatenate(perm_loadings, axis=0) if perm_loadings else np.zeros((0, ncomp))
    real = np.asarray(loadings)[:, :ncomp]
    # empirical p-value of a loading: share of the null loadings of that component above it (dr.py:621-624)
    res = np.empty(real.shape)
    for i in range(ncomp):
        srt = np.sort(null[:, i])
        res[:, i] = (srt.shape[0] - np.searchsorted(srt, real[:, i], side='right')) / max(srt.shape[0], 1)
    res = pd.DataFrame(res, columns=['PC%d' % i for i in range(ncomp)])    # dr.py:625-628 (ignore_index: 0..h-1)
    final = []
    for i in range(ncomp):                                   # dr.py:630-643
        pc = res.iloc[:, i].values
        nsign_found = np.sum(pc < score_thr)
        nsign_expected = np.floor(len(pc) * score_thr)
        ct = [[nsign_found, nsign_expected], [len(pc) - nsign_found, len(pc) - nsign_expected]]
        try:
            pv = chi2_contingency(np.array(ct))[1]
        except ValueError:
            pv = 1
        final.append(['PC%d' % i, pv])
    final = pd.DataFrame(final)
    final['p.adj'] = _p_adjust_bh(final[1])
    final.columns = ['PC', 'chi2_p', 'chi2_p_adj']
    final['significant'] = final.chi2_p_adj < fdr
    if not on_device:
        eng.resident_key = None                              # the device holds the last permuted copy, not `X`
    obj.norm_data[norm]['dr']['jackstraw'] = {'score_mat': res, 'results_by_comp': final}
    if retx:
        return res, final
