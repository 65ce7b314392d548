"""Multi-rank parity check, run under torchrun with one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/multigpu_check.py

Cells of golden case `medium` are split into contiguous row shards; every stage runs sharded (NCCL
all-reduce inside IRLB / gene moments, all-gather for kNN / SNN) and is compared with the golden
vectors of the unmodified reference, exactly like the single-GPU parity tests."""
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

from tests import parity as P


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from adobo_b200.engine import Engine
    from oracle import adobo_oracle as O
    eng = Engine(local)
    eng.init_comm_from_torch()
    case = P.load_case(os.environ.get('CASE', 'medium'))
    n = case['n']
    bounds = np.linspace(0, n, world + 1).astype(int)
    bounds[1:-1] += 7 * np.arange(1, world)                     # ragged shards on purpose
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    # ---- normalize (embarrassingly parallel) ----
    eng.upload_counts(case['counts_csr'][lo:hi])
    vals, lib = eng.normalize(10000, 1, 1.0)
    P.assert_values_close(vals.data, case['norm_csr'][lo:hi].data, rtol=1e-5)
    # ---- HVG: all-reduced moments, replicated ranking ----
    mean, var = eng.gene_moments()
    rmean, rvar = O.gene_moments(case['norm_csr'])
    P.assert_values_close(mean, rmean, rtol=1e-5, what='mean')
    hv, z = eng.hvg_seurat(case['ngenes'], 20)
    rz, _ = O.seurat_zscores(rmean, rvar)
    P.assert_hvg_equal(hv, case['hvg_ordered'], rz, rtol=1e-5)
    # ---- IRLB: A^T u all-reduced, W row-sharded ----
    eng.upload_normalized(case['norm_csr'][lo:hi])
    r = eng.irlb(case['hvg_ordered'], ncomp=case['ncomp'], seed=42)
    ex_comp, ex_contr, _ = P.exact_pca(case['norm_csr'], case['hvg_ordered'], case['ncomp'])
    P.assert_pca_close(r['comp'], r['contr'], case['comp'][lo:hi], case['contr'], ex_comp[lo:hi], ex_contr)
    np.testing.assert_allclose(r['s'], case['s'], rtol=1e-6)
    # the three-kernel step (the path for h > 1024: two exchanges per step in the last CTAs) must give the same PCA as
    # the default two-kernel step (one exchange per step, by the eight CTAs of the V-side cluster)
    os.environ['ADOBO_B200_IRLB'] = 'no_fusedt'
    eng.upload_normalized(case['norm_csr'][lo:hi])
    r2 = eng.irlb(case['hvg_ordered'], ncomp=case['ncomp'], seed=42)
    del os.environ['ADOBO_B200_IRLB']
    P.assert_pca_close(r2['comp'], r2['contr'], case['comp'][lo:hi], case['contr'], ex_comp[lo:hi], ex_contr)
    np.testing.assert_allclose(r2['s'], r['s'], rtol=1e-7)
    if rank == 0:
        print('three-kernel step ok: restarts %d (two-kernel step %d)' % (r2['steps'], r['steps']))
    # every rank must hold bit-identical replicated results
    sig = torch.tensor([float(np.abs(r['contr']).sum()), float(r['s'].sum()), float(r['steps'])], dtype=torch.float64, device='cuda')
    ref = sig.clone()
    dist.broadcast(ref, 0)
    assert torch.equal(sig, ref), 'replicated V / s differ between ranks'
    # ---- kNN: queries sharded, references all-gathered ----
    nn = eng.knn(np.ascontiguousarray(case['comp'][lo:hi]), case['k'])
    mism = nn != case['nn_idx'][lo:hi]
    if mism.any():
        rows = np.flatnonzero(mism.any(axis=1))
        do = P.knn_dist_rows(case['comp'], lo + rows, nn[rows])
        dr = P.knn_dist_rows(case['comp'], lo + rows, case['nn_idx'][lo + rows])
        assert np.all((np.abs(do - dr) <= 1e-12 * np.maximum(dr, 1e-300)) | ~mism[rows])
    # ---- SNN: rows sharded, neighbour lists all-gathered ----
    edges, pruned, total, ne = eng.snn(np.ascontiguousarray(case['nn_idx'][lo:hi]), case['k'], 0.067)
    gold = case['edges']
    mine = gold[(gold[:, 0] >= lo) & (gold[:, 0] < hi)]
    assert np.array_equal(P.edges_sorted(edges), mine), 'SNN edges of shard %d differ' % rank
    msg = str(case['snn_message']).splitlines()[0]
    assert msg == '%.2f%% (n=%s) of links pruned' % (pruned / total * 100, '{:,}'.format(pruned))
    # ---- whole path, device resident ----
    eng.upload_counts(case['counts_csr'][lo:hi])
    np.random.seed(42)
    v0 = np.random.randn(case['ngenes'])
    info = eng.embed_path(v0, ngenes=case['ngenes'], ncomp=case['ncomp'], k=case['k'])
    nn2 = eng.fetch_knn(case['k'])
    agree = float(np.mean(nn2 == case['nn_idx'][lo:hi]))
    assert agree > 0.97, agree
    dist.barrier()
    if rank == 0:
        print('multigpu_check ok: world %d, irlb steps %d, knn agreement %.4f, edges(rank0) %d' % (world, info['steps'], agree, info['n_edges']))
    eng.close()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
