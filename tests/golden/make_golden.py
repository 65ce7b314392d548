"""Generates tests/golden/*.npz by running the UNMODIFIED reference (oracle/ref_loader.py) on
seeded synthetic count matrices.  Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Every stage of the hot path is frozen: normalized values, the ordered HVG list, PCA comp/contr
(+ singular values / restart and mat-vec counts from a direct irlbpy.lanczos call on the same
scaled input), kNN indices, SNN edge list and the pruned-link count.
"""
import os
import sys
import io
import contextlib

import numpy as np
import pandas as pd
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader                      # noqa: E402
from adobo_b200.synth import synth_counts, counts_to_frame  # noqa: E402

CASES = {
    # name: (n_cells, n_genes, density, data_seed, ngenes, ncomp, k)
    'small': (600, 800, 0.10, 11, 200, 20, 10),
    'ragged': (257, 333, 0.05, 12, 60, 5, 5),
    'medium': (3000, 1500, 0.07, 13, 400, 50, 10),
}


def run_case(ref, n, g, density, seed, ngenes, ncomp, k):
    counts = synth_counts(n, g, density, seed)
    df = counts_to_frame(counts)
    obj = ref.data.dataset(df, sparse=True)
    ref.normalize.norm(obj)                                         # normalize.py:436
    nd = obj.norm_data['standard']['data']
    norm_csr = nd.sparse.to_coo().T.tocsr()
    norm_csr.sort_indices()
    assert (norm_csr.indptr == counts.indptr).all() and (norm_csr.indices == counts.indices).all()
    # pandas >= 3 cannot var() a Sparse frame (hvg.py:65): hand seurat the dense frame
    genes = ref.hvg.seurat(nd.sparse.to_dense(), ngenes)            # hvg.py:31
    obj.norm_data['standard']['hvg'] = {'genes': genes, 'method': 'seurat'}
    hvg_idx = np.array([int(s[1:]) for s in genes], dtype=np.int64)
    ref.dr.pca(obj, ncomp=ncomp)                                    # dr.py:236, seed=42
    pca = obj.norm_data['standard']['dr']['pca']
    comp, contr = pca['comp'], pca['contr']
    contr_genes = np.array([int(s[1:]) for s in contr.index], dtype=np.int64)
    # the same call dr.irlb makes (dr.py:156-163), to record s / steps / nmult
    from sklearn.preprocessing import scale as sklearn_scale
    sub = nd[nd.index.isin(genes)].transpose()
    dense = sklearn_scale(sub.sparse.to_dense(), axis=0, with_mean=True, with_std=True)
    lanc = ref.irlbpy.lanczos(pd.DataFrame(dense), nval=ncomp, maxit=1000, seed=42)
    assert np.array_equal(np.asarray(comp), lanc.U.dot(np.diag(lanc.s)))
    nn_idx = ref.clustering.knn(comp, k)                            # clustering.py:25
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        graph = ref.clustering.snn(nn_idx, k, 0.067, True)          # clustering.py:55
    e = graph.values.astype(np.int64)
    e = e[np.lexsort((e[:, 1], e[:, 0]))]
    return dict(
        shape=np.array([n, g, ngenes, ncomp, k]), density=density, data_seed=seed,
        indptr=counts.indptr, indices=counts.indices, counts=counts.data,
        norm_values=norm_csr.data, hvg_ordered=hvg_idx, contr_genes=contr_genes,
        comp=np.asarray(comp), contr=np.asarray(contr), s=lanc.s, steps=lanc.steps, nmult=lanc.nmult,
        nn_idx=nn_idx.astype(np.int64), edges=e, snn_message=buf.getvalue().strip())


if __name__ == '__main__':
    ref = ref_loader.load()
    only = sys.argv[1:]
    for name, cfg in CASES.items():
        if only and name not in only:
            continue
        out = run_case(ref, *cfg)
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), name + '.npz')
        np.savez_compressed(path, **out)
        print(name, 'steps', out['steps'], 'nmult', out['nmult'], out['snn_message'],
              '%.1f KB' % (os.path.getsize(path) / 1024))
