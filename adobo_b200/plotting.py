"""adobo.plotting.pca_contributors, data half (adobo/plotting.py:262-316): which genes load highest on each
component, for all cells or for the cells of one cluster -- the per-cluster case re-runs dr.irlb on the cluster's
columns (plotting.py:298-303), which goes through the GPU path here.  Drawing stays with matplotlib in the reference."""
import pandas as pd

from . import dr as _dr


def pca_contributors_table(obj, normalization=None, clust_alg=None, cluster=None, all_genes=False, dim=(0, 1, 2), top=10,
                           eng=None):
    """Returns {component: Series of the `top` highest and `top` lowest loadings} (what the bar plot draws)."""
    if normalization is None or normalization == '':
        norm = list(obj.norm_data.keys())[-1]
    else:
        norm = normalization
    try:
        target = obj.norm_data[norm]
    except KeyError:
        raise Exception('"%s" not found' % norm)
    if clust_alg is None or clust_alg == '':
        try:
            clust_alg = list(target['clusters'].keys())[-1]
        except IndexError:
            pass
    if not isinstance(dim, (list, tuple)):
        dim = [dim]
    try:
        contr = target['dr']['pca']['contr']
    except KeyError:
        raise Exception('PCA decomposition not found.')
    if cluster is not None:
        X = target['data']
        try:
            cl = target['clusters'][clust_alg]['membership']
        except KeyError:
            raise Exception('"%s" not found, run adobo.clustering.generate(...)' % clust_alg)
        sel = (cl == cluster).to_numpy()
        X_ss = X.select_cells(sel) if hasattr(X, 'select_cells') else X.loc[:, sel]   # plotting.py:298
        if not all_genes:
            hvg = target['hvg']['genes']
            X_ss = X_ss[X_ss.index.isin(hvg)]
        _, contr = _dr.irlb(X_ss, eng=eng)                                # plotting.py:303
    out = {}
    for d in dim:
        s = contr[d].sort_values(ascending=False)
        out[d] = pd.concat([s.head(top), s.tail(top)])                   # plotting.py:308-311
    return out
