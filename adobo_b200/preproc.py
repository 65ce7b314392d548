"""adobo.preproc.impute, front half (adobo/preproc.py:461-488): the subpopulation estimate that scImpute-style
imputation starts from is the hot path with other constants -- CPM scaling, log10(x + 1.01), seurat HVG (1000),
sklearn scale, irlb, knn, snn, leiden -- so it runs on the GPU through the same C-ABI.  The mixture-model fit that
follows (preproc.py:489-640: per-gene gamma / normal EM in worker processes) is host statistics outside the
accelerated path and stays in the reference."""
import numpy as np
import pandas as pd

from . import engine as _engine
from .data import frame_to_csr


def impute_subpopulations(obj, filtered=True, res=0.5, ncomp=75, k=10, prune_snn=0.067, seed=None, clust_alg='leiden',
                          verbose=True, eng=None):
    """Returns dict(comp, nn_idx, graph, membership, hvg) for the cells impute() works on.

    lnorm = log10(raw * (10**6 / col_sums) + 1.01) (preproc.py:470-472) has no zeros -- every implicit zero becomes
    log10(1.01) -- but it is a constant shift of a sparse matrix: the device keeps lnorm - log10(1.01) sparse, gene
    means get the constant back for the seurat bins and the PCA's centring removes it, so nothing is densified.
    `seed` seeds the IRLB start vector (the reference draws it from numpy's global state, irlb.py:151-152);
    clust_alg=None skips the host partitioner."""
    from . import clustering
    eng = eng or _engine.default_engine()
    raw = obj.count_data
    if filtered:                                              # preproc.py:461-469
        remove = obj.meta_cells.status[obj.meta_cells.status != 'OK']
        raw = raw.drop(remove.index, axis=1)
        remove = obj.meta_genes.status[obj.meta_genes.status != 'OK']
        raw = raw.drop(remove.index, axis=0)
        if verbose:
            print('Running on the quality filtered data (dimensions %sx%s)' % raw.shape)
    eng.upload_counts(frame_to_csr(raw))
    eng.normalize(10 ** 6, _engine.LOG_10, 1.01, want_values=False)      # preproc.py:470-472
    eng.gene_moments()
    idx, _ = eng.hvg_seurat(1000, 20)                                      # preproc.py:475
    hvg = np.array(raw.index[idx])
    r = eng.irlb(idx, ncomp=ncomp, maxit=1000, seed=seed, scale=True, var_weigh=True)   # preproc.py:476-483
    comp = pd.DataFrame(r['comp'], index=raw.columns)
    nn_idx = eng.knn(None, k)                                              # preproc.py:485
    edges, pruned, total, ne = eng.snn(None, k, prune_snn)                 # preproc.py:486
    graph = pd.DataFrame({'source_node': edges[:, 0].astype(np.int64), 'target_node': edges[:, 1].astype(np.int64)})
    eng.resident_key = None
    out = {'comp': comp, 'nn_idx': nn_idx, 'graph': graph, 'hvg': hvg, 'membership': None}
    if clust_alg == 'leiden':
        out['membership'] = np.array(clustering.leiden(graph, res))        # preproc.py:487 (host: leidenalg)
    elif clust_alg is not None:
        raise Exception('impute_subpopulations: clust_alg must be "leiden" or None')
    return out


def impute(obj, filtered=True, res=0.5, drop_thre=0.5, nworkers='auto', verbose=True):
    raise Exception('adobo_b200 accelerates the subpopulation estimate of impute (impute_subpopulations); the '
                    'mixture-model fit stays in adobo.preproc.impute.')
