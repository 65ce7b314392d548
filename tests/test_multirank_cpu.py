"""N>1 host logic on CPU: two gloo ranks run the SHARDED formulation the GPU path uses (cells
row-sharded; gene moments / A^T u / Gram-Schmidt coefficients / norms all-reduced; V, B replicated;
kNN queries sharded against all-gathered references; SNN rows sharded against all-gathered lists)
with the oracle's arithmetic, and must reproduce the single-rank oracle / golden vectors."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import scipy.sparse as sp
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from oracle import adobo_oracle as O
    from tests import parity as P

    def allreduce(a):
        t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64).copy())
        dist.all_reduce(t)
        return t.numpy()

    def allgather_rows(a):
        parts = [None] * world
        dist.all_gather_object(parts, a)
        return np.concatenate(parts, axis=0)

    case = P.load_case('small')
    n = case['n']
    bounds = np.linspace(0, n, world + 1).astype(int)
    bounds[1:-1] += 5
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    norm, _ = O.normalize_standard(case['counts_csr'][lo:hi])
    assert np.array_equal(norm.data, case['norm_csr'][lo:hi].data)
    # gene moments from all-reduced one-pass sums (what csr_gene_moments + NCCL compute)
    s1 = allreduce(np.asarray(norm.sum(axis=0)).ravel())
    s2 = allreduce(np.asarray(norm.multiply(norm).sum(axis=0)).ravel())
    mean = s1 / n
    var = (s2 - n * mean * mean) / (n - 1)
    rmean, rvar = O.gene_moments(case['norm_csr'])
    assert np.allclose(mean, rmean, rtol=1e-12) and np.allclose(var, rvar, rtol=1e-9, atol=1e-14)
    z, bins = O.seurat_zscores(mean, var)
    rz, _ = O.seurat_zscores(rmean, rvar)
    hv = O.seurat_order(z, bins, case['ngenes'])
    P.assert_hvg_equal(hv, case['hvg_ordered'], rz, rtol=1e-9)
    # sharded implicit operator + all-reduced Lanczos
    genes = np.sort(case['hvg_ordered'])
    sub = sp.csr_matrix(norm)[:, genes]
    cs = allreduce(np.asarray(sub.sum(axis=0)).ravel())
    mu = cs / n
    dev = sub.copy()
    dev.data = (dev.data - mu[dev.indices]) ** 2
    sq = np.asarray(dev.sum(axis=0)).ravel() + ((hi - lo) - np.diff(sub.tocsc().indptr)) * mu * mu
    sd = np.sqrt(allreduce(sq) / n)
    sd[sd < 10 * np.finfo(float).eps] = 1.0
    op = O.ScaledCenteredOp(sub, mu, sd, allreduce=allreduce)
    r = O.lanczos(op, case['ncomp'], maxit=1000, seed=42, allreduce=allreduce)
    comp = r['U'] * r['s'][None, :]
    ex_comp, ex_contr, _ = P.exact_pca(case['norm_csr'], case['hvg_ordered'], case['ncomp'])
    P.assert_pca_close(comp, r['V'], case['comp'][lo:hi], case['contr'], ex_comp[lo:hi], ex_contr)
    assert np.allclose(r['s'], case['s'], rtol=1e-6)
    # kNN: sharded queries vs all-gathered references
    full = allgather_rows(case['comp'][lo:hi])
    assert np.array_equal(full, case['comp'])
    x = full
    d = ((x[lo:hi, None, :] - x[None, :, :]) ** 2).sum(-1)
    nn = np.lexsort((np.broadcast_to(np.arange(n), d.shape), d), axis=1)[:, :case['k']] + 1
    P.assert_knn_equal_rows = None
    mism = nn != case['nn_idx'][lo:hi]
    if mism.any():
        rows = np.flatnonzero(mism.any(axis=1))
        do = P.knn_dist_rows(x, lo + rows, nn[rows])
        dr = P.knn_dist_rows(x, lo + rows, case['nn_idx'][lo + rows])
        assert np.all((np.abs(do - dr) <= 1e-12 * np.maximum(dr, 1e-300)) | ~mism[rows])
    # SNN: every rank gathers all lists, emits its rows
    nn_all = allgather_rows(case['nn_idx'][lo:hi])
    e, pruned, total = O.snn(nn_all, case['k'], 0.067)
    mine = e[(e[:, 0] >= lo) & (e[:, 0] < hi)]
    gold = case['edges']
    assert np.array_equal(mine, gold[(gold[:, 0] >= lo) & (gold[:, 0] < hi)])
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, 'ok', r['steps']))


def test_two_rank_sharded_path_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 300)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
    codes = [p.exitcode for p in procs]
    for p in procs:
        if p.is_alive():
            p.terminate()
    assert codes == [0, 0], codes
    got = sorted(q.get(timeout=5) for _ in range(2))
    assert [g[1] for g in got] == ['ok', 'ok']
    assert got[0][2] == got[1][2]                        # replicated control flow: same restart count


def test_shard_bounds_cover_all_cells():
    from adobo_b200.synth import synth_counts
    full = synth_counts(3000, 50, 0.1, seed=3)
    parts = [synth_counts(3000, 50, 0.1, seed=3, row_start=a, row_stop=b) for a, b in ((0, 1100), (1100, 2048), (2048, 3000))]
    import scipy.sparse as sp
    assert (sp.vstack(parts) != full).nnz == 0
