"""Golden vectors at BASELINE.json's own shapes, from the UNMODIFIED reference (oracle/ref_loader.py).

    python tests/golden/make_golden_configs.py [c1] [c2]        (build container only; minutes per case)

c1 = configs[0]: 8,400 cells x 33,000 genes, 4.2 % non-zeros, 1000 HVG, 75 components, k = 10 (data seed 42)
c2 = configs[1]: 20,000 cells x 25,000 genes, 6.4 % non-zeros, same pipeline (data seed 43) -- bench.py's `c2` workload

The count matrix itself is not stored (adobo_b200.synth regenerates it from the seed; its nnz and checksums are
stored to prove the regeneration), nor the ~100 MB of normalized values (a seeded sample of 200,000 of them is).
Everything downstream is stored whole: library sizes, the reference's per-gene mean / variance, the ordered HVG
list, comp / contr / s / steps / nmult, nn_idx, the SNN edge list and the pruned-links message.
"""
import contextlib
import io
import os
import sys
import time

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader                      # noqa: E402
from adobo_b200.synth import synth_counts, counts_to_frame  # noqa: E402

CONFIGS = {
    # name: (n_cells, n_genes, density, data_seed, ngenes, ncomp, k)
    'c1': (8400, 33000, 0.042, 42, 1000, 75, 10),
    'c2': (20000, 25000, 0.064, 43, 1000, 75, 10),
}


def run_config(ref, n, g, density, seed, ngenes, ncomp, k):
    tm = {}
    t0 = time.perf_counter()
    counts = synth_counts(n, g, density, seed)
    df = counts_to_frame(counts)
    tm['synth'] = time.perf_counter() - t0; t0 = time.perf_counter()
    obj = ref.data.dataset(df, sparse=True)
    del df
    tm['dataset'] = time.perf_counter() - t0; t0 = time.perf_counter()
    ref.normalize.norm(obj)                                         # normalize.py:436
    tm['norm'] = time.perf_counter() - t0; t0 = time.perf_counter()
    nd = obj.norm_data['standard']['data']
    norm_csr = nd.sparse.to_coo().T.tocsr()
    norm_csr.sort_indices()
    assert (norm_csr.indptr == counts.indptr).all() and (norm_csr.indices == counts.indices).all()
    dense = nd.sparse.to_dense()                                    # pandas >= 3 cannot var() a Sparse frame (hvg.py:65)
    tm['to_dense'] = time.perf_counter() - t0; t0 = time.perf_counter()
    genes = ref.hvg.seurat(dense, ngenes)                           # hvg.py:31
    tm['seurat'] = time.perf_counter() - t0; t0 = time.perf_counter()
    gene_mean = np.asarray(dense.mean(axis=1))                      # hvg.py:60
    gene_var = np.asarray(dense.var(axis=1))                        # hvg.py:65 (ddof = 1)
    del dense
    obj.norm_data['standard']['hvg'] = {'genes': genes, 'method': 'seurat'}
    hvg_idx = np.array([int(s[1:]) for s in genes], dtype=np.int64)
    t0 = time.perf_counter()
    ref.dr.pca(obj, ncomp=ncomp)                                    # dr.py:236, seed=42
    tm['pca'] = time.perf_counter() - t0
    pca = obj.norm_data['standard']['dr']['pca']
    comp, contr = pca['comp'], pca['contr']
    contr_genes = np.array([int(s[1:]) for s in contr.index], dtype=np.int64)
    # the same call dr.irlb makes (dr.py:156-163), to record s / steps / nmult
    from sklearn.preprocessing import scale as sklearn_scale
    sub = nd[nd.index.isin(genes)].transpose()
    scaled = sklearn_scale(sub.sparse.to_dense(), axis=0, with_mean=True, with_std=True)
    lanc = ref.irlbpy.lanczos(pd.DataFrame(scaled), nval=ncomp, maxit=1000, seed=42)
    assert np.array_equal(np.asarray(comp), lanc.U.dot(np.diag(lanc.s)))
    # the reference's own error against the exact SVD of the matrix it decomposed (documents which components
    # of ITS result are further than 1e-4 from the truth)
    U, s_ex, Vt = np.linalg.svd(scaled, full_matrices=False)
    exact_comp = U[:, :ncomp] * s_ex[:ncomp]
    exact_contr = Vt[:ncomp].T
    del U, Vt, scaled
    t0 = time.perf_counter()
    nn_idx = ref.clustering.knn(comp, k)                            # clustering.py:25
    tm['knn'] = time.perf_counter() - t0; t0 = time.perf_counter()
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        graph = ref.clustering.snn(nn_idx, k, 0.067, True)          # clustering.py:55
    tm['snn'] = time.perf_counter() - t0
    e = graph.values.astype(np.int64)
    e = e[np.lexsort((e[:, 1], e[:, 0]))]
    rng = np.random.default_rng(1234)
    pos = np.sort(rng.choice(norm_csr.nnz, min(200000, norm_csr.nnz), replace=False))
    return dict(
        shape=np.array([n, g, ngenes, ncomp, k]), density=density, data_seed=seed,
        nnz=np.int64(counts.nnz), counts_sum=np.int64(counts.data.astype(np.int64).sum()),
        indices_sum=np.int64(counts.indices.astype(np.int64).sum()),
        libsize=np.asarray(counts.sum(axis=1)).ravel().astype(np.int64),
        norm_sample_pos=pos, norm_sample=norm_csr.data[pos], norm_sum=np.float64(norm_csr.data.sum()),
        gene_mean=gene_mean, gene_var=gene_var,
        hvg_ordered=hvg_idx, contr_genes=contr_genes,
        comp=np.asarray(comp), contr=np.asarray(contr), s=lanc.s, steps=lanc.steps, nmult=lanc.nmult,
        exact_comp_err=_component_err(np.asarray(comp), exact_comp), exact_contr_err=_component_err(np.asarray(contr), exact_contr),
        s_exact=s_ex[:ncomp],
        nn_idx=nn_idx.astype(np.int32), edges=e.astype(np.int32), snn_message=buf.getvalue().strip(),
        ref_stage_seconds=np.array([tm[k_] for k_ in ('dataset', 'norm', 'to_dense', 'seurat', 'pca', 'knn', 'snn')]))


def _component_err(a, b):
    sg = np.sign((a * b).sum(axis=0))
    sg[sg == 0] = 1
    return np.abs(a * sg - b).max(axis=0) / np.abs(b).max(axis=0)


if __name__ == '__main__':
    ref = ref_loader.load()
    only = sys.argv[1:]
    for name, cfg in CONFIGS.items():
        if only and name not in only:
            continue
        t0 = time.perf_counter()
        out = run_config(ref, *cfg)
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), name + '.npz')
        np.savez_compressed(path, **out)
        print(name, 'steps', out['steps'], 'nmult', out['nmult'], out['snn_message'], 'edges', out['edges'].shape[0],
              'stage s', np.round(out['ref_stage_seconds'], 1), '%.1f MB' % (os.path.getsize(path) / 2**20),
              'total %.0f s' % (time.perf_counter() - t0), flush=True)
