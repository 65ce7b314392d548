#!/usr/bin/env python
"""bench.py -- cells/s of adobo's cell-embedding hot path on B200.

    python bench.py --gpus N --steps K --warmup W [--config c1|c2|c3|c4|c5]   (N>1: launched by torchrun, one rank/GPU)
    python bench.py --impl reference ...                                     (CPU arm: the oracle port on host cores)

A "step" is one pass of the whole path normalize -> HVG(seurat) -> IRLB PCA -> kNN -> SNN over one synthetic count
matrix.  The configurations are BASELINE.json's:

    c1  PBMC-8k shape       8,400 cells x 33,000 genes, 4.2 % non-zeros, 75 components   (host generator; golden vectors
    c2  choroid-plexus      20,000 x 25,000, 6.4 %, 75 components                         from the unmodified reference)
    c3  PBMC-68k shape      68,000 x 33,000, 3 %, 75 components        <- default, the shape the metric is quoted on
    c4  mouse-brain shape   1,300,000 x 28,000, 7 %, 50 components     (8 GPUs; generated on the device, shard by shard)
    c5  kNN / SNN sweep     1,000,000 x 50 embedding, k = 10 / 20 / 30

With N > 1 the SAME matrix is row-sharded over the ranks (strong scaling); IRLB exchanges one vector per Lanczos
step over NVLink peer memory, kNN / SNN all-gather over NCCL.

`value`    device-resident throughput: counts already in HBM when the timed region starts.
`e2e`      same metric through the host-facing C-ABI calls: pinned host CSR in, PCA + graph out, copies inside the
           timed region (wall clock, max over ranks).
`roofline` the dominant kernel of a step (by CUDA-event time, measured in extra profiled steps right after the timed
           region): algorithmic bytes / average launch time vs the measured HBM copy peak.
`cpu_baseline` the oracle port (numpy/scipy restatement of the reference) on this box's host cores, on a bounded
           sample of the workload (its first 20,000 cells).
`parity`   the results of the timed configuration checked in the same run: against the golden vectors of the
           unmodified reference (c1, c2) or against the oracle / exact recomputation (c3: stage by stage).
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SMI_FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')

CONFIGS = {
    'c1': dict(cells=8400, genes=33000, density=0.042, seed=42, ngenes=1000, ncomp=75, k=10, gen='host', golden='c1',
               name='PBMC-8k-shape synthetic counts (BASELINE configs[0])'),
    'c2': dict(cells=20000, genes=25000, density=0.064, seed=43, ngenes=1000, ncomp=75, k=10, gen='host', golden='c2',
               name='choroid-plexus-shape synthetic counts (BASELINE configs[1])'),
    'c3': dict(cells=68000, genes=33000, density=0.030, seed=44, ngenes=1000, ncomp=75, k=10, gen='device', golden=None,
               name='PBMC-68k-shape synthetic counts (BASELINE configs[2])'),
    'c4': dict(cells=1300000, genes=28000, density=0.070, seed=45, ngenes=1000, ncomp=50, k=10, gen='device', golden=None,
               name='mouse-brain-shape synthetic counts (BASELINE configs[3])'),
}
SAMPLE_CELLS = 20000      # CPU arms: the oracle runs on the first SAMPLE_CELLS cells of the workload
METRIC = 'cells/sec normalize->HVG->IRLB PCA->kNN/SNN'


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default='c3', choices=['c1', 'c2', 'c3', 'c4', 'c5'])
    ap.add_argument('--cells', type=int, default=0, help='override the total number of cells of the configuration')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-api', action='store_true', help='skip the api_e2e figure (the Python call sequence)')
    ap.add_argument('--parity', default='fast', choices=['off', 'fast', 'full'],
                    help='c3/c4: fast = stage-wise exact / property checks; full = also the whole oracle on the full matrix')
    ap.add_argument('--profile-steps', type=int, default=2)
    ap.add_argument('--ks', default='', help='c5: comma-separated k values (default 10,20,30)')
    return ap.parse_args()


def peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            p = json.load(f)
        return float(p['hbm_gbs']), float(p['bf16_tflops']), 'measured'
    except Exception:
        return 6650.0, 1590.0, 'fallback'


class ClockSampler:
    def __init__(self, device):
        self.proc = None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(device), '--query-gpu=' + SMI_FIELDS,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out = self.proc.communicate(timeout=5)[0]
        except Exception:
            self.proc.kill()
            out = ''
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': float(max(mx)) if mx else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


def config_of(args):
    cfg = dict(CONFIGS[args.config])
    if args.cells:
        cfg['cells'] = args.cells
    return cfg


def workload_config(cfg, world, key):
    return {'workload': '%s: %d cells x %d genes, %.1f%% non-zeros, standard norm, seurat HVG %d, irlb PCA %d comps, '
                        'knn/snn k=%d prune 0.067' % (cfg['name'], cfg['cells'], cfg['genes'], cfg['density'] * 100,
                                                      cfg['ngenes'], cfg['ncomp'], cfg['k']),
            'config': key, 'cells_total': cfg['cells'], 'genes': cfg['genes'], 'ngenes': cfg['ngenes'],
            'ncomp': cfg['ncomp'], 'k': cfg['k'], 'parallelism': 'cells row-sharded x%d (same matrix for every N)' % world,
            'generator': 'numpy, adobo_b200/synth.py' if cfg['gen'] == 'host' else 'device (Philox), csrc/synth.cu',
            'l2': 'L2 flushed (256 MB memset) between timed steps'}


def shard_bounds(n, world):
    return [(n * r) // world for r in range(world + 1)]


def sample_counts(cfg, eng=None):
    """The first SAMPLE_CELLS cells of the workload as host CSR (the bounded sample of the CPU arms)."""
    from adobo_b200.synth import synth_counts, CountModel
    ns = min(SAMPLE_CELLS, cfg['cells'])
    if cfg['gen'] == 'host':
        return synth_counts(cfg['cells'], cfg['genes'], cfg['density'], cfg['seed'], 0, ns)
    from adobo_b200.engine import Engine
    own = eng is None
    eng = eng or Engine(0)
    eng.synth_counts(CountModel(cfg['genes'], cfg['density'], cfg['seed']), 0, ns)
    m = eng.download_counts()
    if own:
        eng.close()
    return m


def host_threads():
    """All host cores for the CPU arm's BLAS: torchrun exports OMP_NUM_THREADS=1 to every rank, which would time a
    single-threaded reference at N > 1.  Returns the thread count in use."""
    n = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_info, threadpool_limits
        threadpool_limits(limits=n)
        used = [p.get('num_threads', 1) for p in threadpool_info()]
        return max(used) if used else n
    except Exception:
        return int(os.environ.get('OMP_NUM_THREADS', n))


def run_reference(args, rank, world):
    """CPU arm: the oracle port (the reference is pure Python + pandas and cannot travel to the GPU box;
    oracle/adobo_oracle.py restates its arithmetic on scipy.sparse and is pinned to the reference's own outputs by
    tests/golden).  Each step = the whole path on a bounded sample of the workload: its first 20,000 cells."""
    if rank != 0:
        return
    from oracle import adobo_oracle as O
    host_threads()
    if args.config == 'c5':
        return run_reference_c5(args)
    cfg = config_of(args)
    try:
        counts = sample_counts(cfg)
    except Exception as e:                               # device generator without a GPU
        print(json.dumps({'impl': 'reference', 'unavailable': 'cannot generate the %s sample: %s' % (args.config, e)}))
        return
    ns = counts.shape[0]
    times, tm = [], {}
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        O.embed_path(counts, cfg['ngenes'], cfg['ncomp'], cfg['k'], 0.067, 42, tm)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    total = float(np.sum(times))
    value = ns * args.steps / total
    cores = host_threads()
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'cells/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': total / args.steps * 1e3,
        'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(cfg, world, args.config),
        'cpu_baseline': {'value': value, 'unit': 'cells/s', 'cores': cores, 'kind': 'port',
                         'sample': 'first %d of the %d cells per step, whole path; numpy/scipy with up to %d BLAS threads'
                                   % (ns, cfg['cells'], cores),
                         'stage_s': {k: float(v) for k, v in tm.items()}},
        'e2e': {'value': value, 'unit': 'cells/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------------------
# parity of the timed configuration (rank 0, N = 1)
# --------------------------------------------------------------------------------------------------------------
def parity_block(args, cfg, eng, counts_host, info, v0):
    """Checks the results left on the device by the last timed pass.  c1/c2: against the golden vectors frozen from
    the unmodified reference (tests/golden/c1.npz, c2.npz).  c3/c4: stage by stage against the oracle / an exact
    recomputation on this box's CPU."""
    from oracle import adobo_oracle as O
    from tests import parity as P
    import scipy.sparse as sp
    out = {}
    h, p, k = cfg['ngenes'], cfg['ncomp'], cfg['k']
    pca = eng.fetch_pca(h, p)
    nn = eng.fetch_knn(k)
    edges = eng.fetch_edges(info['n_edges'])
    if cfg['golden'] and not args.cells:
        G = dict(np.load(os.path.join(ROOT, 'tests', 'golden', cfg['golden'] + '.npz')))
        out['against'] = 'golden vectors of the unmodified reference (tests/golden/%s.npz)' % cfg['golden']
        out['input_identical'] = bool(counts_host.nnz == int(G['nnz']) and int(counts_host.data.astype(np.int64).sum()) == int(G['counts_sum']))
        ref_genes = np.sort(G['hvg_ordered'])
        out['hvg_set_equal'] = bool(np.array_equal(np.sort(pca['genes']), ref_genes))
        if out['hvg_set_equal']:
            ec = P.component_err(pca['comp'], G['comp'])
            out['pca_comp_err'] = [float('%.3g' % e) for e in ec]
            out['pca_components_within_1e-4'] = int(np.sum(ec <= 1e-4))
            out['pca_components_beyond_reference_own_error_x2.5'] = int(np.sum(ec > np.maximum(1e-4, 2.5 * G['exact_comp_err'])))
            out['reference_own_comp_err_vs_exact_svd_max'] = float(G['exact_comp_err'].max())
            out['s_max_rel_err'] = float(np.max(np.abs(pca['s'] - G['s']) / G['s']))
        out['restarts'] = [int(info['steps']), int(G['steps'])]
        out['matvecs'] = [int(info['nmult']), int(G['nmult'])]
        out['knn_row_agreement'] = float(np.mean(np.all(nn == G['nn_idx'], axis=1)))
        ours = set(map(tuple, edges.tolist()))
        theirs = set(map(tuple, G['edges'].tolist()))
        out['edges'] = [len(ours), len(theirs)]
        out['edge_jaccard'] = len(ours & theirs) / max(1, len(ours | theirs))
        # stage-wise exactness on the reference's own upstream outputs
        nn_sw = eng.knn(G['comp'], k)
        out['knn_stagewise_exact'] = bool(P.assert_knn_equal(nn_sw, G['nn_idx'].astype(np.int64), G['comp']) >= 0)
        e_sw, pruned, total, _ = eng.snn(G['nn_idx'].astype(np.int64), k, 0.067)
        out['snn_stagewise_exact'] = bool(np.array_equal(P.edges_sorted(e_sw), G['edges']))
        return out
    out['against'] = 'oracle / exact recomputation on the host, stage by stage'
    # normalize: every value of 512 sampled cells
    rng = np.random.default_rng(7)
    n = counts_host.shape[0]
    norm_o, _ = O.normalize_standard(counts_host)
    # HVG list: oracle on the full matrix
    t0 = time.perf_counter()
    hv = O.seurat(norm_o, h)
    out['hvg_set_equal'] = bool(np.array_equal(np.sort(hv), np.sort(pca['genes'])))
    out['hvg_oracle_s'] = time.perf_counter() - t0
    # PCA: the triplets must satisfy IRLB's own stopping rule on the exact operator (irlb.py:231: residual < tol smax)
    genes = np.sort(pca['genes'])
    sub = sp.csr_matrix(norm_o)[:, genes]
    mu = np.asarray(sub.mean(axis=0)).ravel()
    ex2 = np.asarray(sub.multiply(sub).mean(axis=0)).ravel()
    sd = np.sqrt(np.maximum(ex2 - mu * mu, 0.0))
    sd[sd < 10 * np.finfo(np.float64).eps] = 1.0
    op = O.ScaledCenteredOp(sub, mu, sd)
    U = pca['comp'] / pca['s'][None, :]
    V = pca['contr']
    res_l = np.array([np.linalg.norm(op.rmv(U[:, i]) - pca['s'][i] * V[:, i]) for i in range(p)])
    res_r = np.array([np.linalg.norm(op.mv(V[:, i]) - pca['s'][i] * U[:, i]) for i in range(p)])
    out['pca_residual_over_tol_smax'] = float(max(res_l.max(), res_r.max()) / (1e-4 * pca['s'][0]))
    out['pca_orthogonality'] = float(max(np.abs(U.T @ U - np.eye(p)).max(), np.abs(V.T @ V - np.eye(p)).max()))
    out['restarts'] = [int(info['steps']), None]
    if args.parity == 'full':
        t0 = time.perf_counter()
        o = O.irlb_pca(norm_o, hv, ncomp=p, seed=42)
        ec = P.component_err(pca['comp'], o['comp'])
        out['pca_comp_err_vs_oracle'] = [float('%.3g' % e) for e in ec]
        out['pca_components_within_1e-4'] = int(np.sum(ec <= 1e-4))
        out['restarts'] = [int(info['steps']), int(o['steps'])]
        out['matvecs'] = [int(info['nmult']), int(o['nmult'])]
        out['irlb_oracle_s'] = time.perf_counter() - t0
    # kNN: sampled queries, exact fp64 brute force on the GPU's own embedding (stage-wise)
    rows = rng.choice(n, min(1024, n), replace=False)
    x = pca['comp']
    sq = np.einsum('ij,ij->i', x, x)
    bad = 0
    for lo in range(0, rows.size, 256):
        r = rows[lo:lo + 256]
        d = sq[r, None] + sq[None, :] - 2.0 * x[r] @ x.T
        cand = np.argpartition(d, k + 8, axis=1)[:, :k + 8]
        dd = np.einsum('ijk,ijk->ij', x[r][:, None, :] - x[cand], x[r][:, None, :] - x[cand])
        order = np.lexsort((cand, dd), axis=1)[:, :k]
        ref = np.take_along_axis(cand, order, axis=1) + 1
        dref = np.take_along_axis(dd, order, axis=1)
        got = nn[r]
        dg = np.einsum('ijk,ijk->ij', x[r][:, None, :] - x[got - 1], x[r][:, None, :] - x[got - 1])
        bad += int(np.sum((got != ref) & ~(np.abs(dg - dref) <= 1e-12 * np.maximum(dref, 1e-300))))
    out['knn_sampled_queries'] = int(rows.size)
    out['knn_sampled_mismatches_not_ties'] = bad
    # SNN: the oracle on the GPU's own neighbour lists, whole edge set
    e_o, pruned_o, total_o = O.snn(nn, k, 0.067)
    out['snn_edges_exact'] = bool(np.array_equal(P.edges_sorted(edges), e_o))
    out['edges'] = [int(edges.shape[0]), int(e_o.shape[0])]
    return out


# --------------------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))

    from adobo_b200.engine import Engine
    from adobo_b200._lib import check as _check
    from adobo_b200.synth import synth_counts, CountModel

    eng = Engine(local_rank)
    if world > 1:
        eng.init_comm_from_torch()

    def barrier():
        eng.sync()
        if dist is not None:
            dist.barrier()
        eng.sync()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    if args.config == 'c5':
        run_c5(args, eng, dist, rank, world, local_rank, barrier, max_over_ranks)
        return
    cfg = config_of(args)
    n_total = cfg['cells']
    bounds = shard_bounds(n_total, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    t_gen = time.perf_counter()
    if cfg['gen'] == 'host':
        counts = synth_counts(n_total, cfg['genes'], cfg['density'], cfg['seed'], lo, hi)
        eng.upload_counts(counts)
    else:
        eng.synth_counts(CountModel(cfg['genes'], cfg['density'], cfg['seed']), lo, hi - lo)
        counts = None
    eng.sync()
    t_gen = time.perf_counter() - t_gen
    np.random.seed(42)                                   # dr.pca default seed (dr.py:237)
    v0 = np.random.randn(cfg['ngenes'])
    path = (v0, 10000, cfg['ngenes'], cfg['ncomp'], cfg['k'], 0.067)

    # ---------------- device-resident throughput ----------------
    info = None
    for _ in range(args.warmup):
        eng.l2_flush()
        info = eng.embed_path(*path)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    l0 = eng.launch_count()
    dev_ms = 0.0
    for _ in range(args.steps):
        eng.l2_flush()
        eng.record(0)
        info = eng.embed_path(*path)
        eng.record(1)
        dev_ms += eng.elapsed_ms(0, 1)
    barrier()
    launches = eng.launch_count() - l0 - args.steps      # minus the L2-flush memsets
    clocks = sampler.stop() if sampler else None
    dev_ms = max_over_ranks(dev_ms)
    value = n_total * args.steps / (dev_ms * 1e-3)

    # ---------------- end to end through the host-facing calls ----------------
    if counts is None:
        counts = eng.download_counts()                   # this rank's shard, once, outside the timed region
    indptr = np.ascontiguousarray(counts.indptr, dtype=np.int64)
    indices = np.ascontiguousarray(counts.indices, dtype=np.int32)
    data = np.ascontiguousarray(counts.data, dtype=np.int32)
    for a in (indptr, indices, data):
        eng.pin(a)
    h = cfg['ngenes']
    e2e_s, d2h = 0.0, 0
    e2e_steps = max(1, min(args.steps, 3 if n_total > 200000 else args.steps))
    # results land in caller-owned page-locked arrays that are reused from pass to pass (the C-ABI's outputs are
    # caller-allocated): allocated after the first, untimed pass has told the number of edges
    out_pca, out_edges, stage = None, None, [0.0, 0.0, 0.0]
    for it in range(1 + e2e_steps):
        barrier()
        t0 = time.perf_counter()
        _check(eng.lib.adobo_upload_counts(eng.h, counts.shape[0], counts.shape[1], indptr, indices, data))
        t1 = time.perf_counter()
        r = eng.embed_path(*path)
        t2 = time.perf_counter()
        pca = eng.fetch_pca(h, cfg['ncomp'], out=out_pca)
        edges = eng.fetch_edges(r['n_edges'], out=out_edges)
        eng.sync()
        t3 = time.perf_counter()
        dt = max_over_ranks(t3 - t0)
        if it >= 1:
            e2e_s += dt
            stage = [stage[0] + t1 - t0, stage[1] + t2 - t1, stage[2] + t3 - t2]
            d2h = pca['comp'].nbytes + pca['contr'].nbytes + pca['s'].nbytes + pca['genes'].nbytes + 2 * edges[0].nbytes
        else:
            out_pca = dict(comp=np.empty((hi - lo, cfg['ncomp'])), contr=np.empty((h, cfg['ncomp'])), s=np.empty(cfg['ncomp']),
                           genes=np.empty(h, dtype=np.int32))
            out_edges = (np.empty(max(r['n_edges'], 1), dtype=np.int32), np.empty(max(r['n_edges'], 1), dtype=np.int32))
            for a in list(out_pca.values()) + list(out_edges):
                eng.pin(a)
    e2e_value = n_total * e2e_steps / e2e_s
    h2d = indptr.nbytes + indices.nbytes + data.nbytes + v0.nbytes
    for a in [indptr, indices, data] + list(out_pca.values()) + list(out_edges):
        eng.unpin(a)
    del pca, edges

    # ---------------- the reference's own call sequence (adobo_b200's Python modules), one GPU ----------------
    # dataset.from_csr(frame=False) -> normalize.norm -> hvg.find_hvg -> dr.pca -> clustering.generate(graph half): host CSR
    # in, norm_data entries (CsrFrame / pandas frames of the PCA and the edge list) out; wall clock of the second call
    api = None
    if world == 1 and not args.no_api:
        try:
            import scipy.sparse as sp
            import adobo_b200 as ad
            from adobo_b200 import engine as E
            E.set_default_engine(eng)
            m = sp.csr_matrix((data, indices, indptr), shape=counts.shape)
            api_s = []
            for it in range(2):
                t0 = time.perf_counter()
                exp = ad.data.dataset.from_csr(m, frame=False)
                ad.normalize.norm(exp)
                ad.hvg.find_hvg(exp, ngenes=cfg['ngenes'])
                ad.dr.pca(exp, ncomp=cfg['ncomp'])
                ad.clustering.generate(exp, k=cfg['k'], clust_alg=None)
                api_s.append(time.perf_counter() - t0)
                n_api_edges = len(exp.norm_data['standard']['graph'])
                del exp
            api = {'value': n_total / api_s[-1], 'unit': 'cells/s', 'ms': api_s[-1] * 1e3, 'first_call_ms': api_s[0] * 1e3,
                   'edges': int(n_api_edges),
                   'path': 'dataset.from_csr(frame=False) -> normalize.norm -> hvg.find_hvg -> dr.pca -> clustering.generate'
                           '(clust_alg=None): host CSR in, norm_data entries out (PCA frames, edge frame; the normalized '
                           'matrix is a DeviceCsrFrame: labels on the host, values left on the device until read)'}
        except Exception as ex:  # the figure is informative; the contract keys above do not depend on it
            api = {'error': '%s: %s' % (type(ex).__name__, ex)}

    # ---------------- per-kernel breakdown (profiled steps, outside the timed region) ----------------
    eng.profile(True)
    for _ in range(args.profile_steps):
        eng.l2_flush()
        eng.embed_path(*path)
    prof = eng.profile_read()
    eng.profile(False)
    hbm_peak, bf16_peak, peak_kind = peaks()
    # Every profiled launch sits between two CUDA events of its own; an EMPTY kernel between the same pair of events
    # (64 of them, launched by adobo_profile) measures what that costs.  An empty kernel itself lasts ~1 us, so the
    # overhead charged to a launch is that time minus 1 us -- conservative for the achieved figures below.
    ov = prof.pop('_event_overhead', None)
    event_overhead_us = max(0.0, ov['ms'] * 1e3 / max(ov['launches'], 1) - 1.0) if ov else 0.0
    kernels = []
    for nm, v in prof.items():
        per_launch_us = v['ms'] * 1e3 / max(v['launches'], 1)
        net_us = max(per_launch_us - event_overhead_us, 0.25 * per_launch_us)
        ach = (v['work'] / max(v['launches'], 1) / (net_us * 1e-6) / 1e9) if v['ms'] > 0 and v['work'] > 0 else None
        kernels.append({'kernel': nm, 'ms_per_step': v['ms'] / args.profile_steps,
                        'launches_per_step': v['launches'] / args.profile_steps, 'us_per_launch': net_us,
                        'us_per_launch_with_events': per_launch_us, 'achieved': ach,
                        'work_per_launch': v['work'] / max(v['launches'], 1)})
    kernels.sort(key=lambda d: -d['ms_per_step'])
    try:
        with open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')) as f:
            traffic = json.load(f).get(args.config, {})
    except Exception:
        traffic = {}
    n_loc = hi - lo

    def roofline_of(kd):
        nm = kd['kernel']
        if nm in ('knn_scores_tc', 'knn_shortlist'):
            issued = (kd['achieved'] or 0) / 1e3             # GFLOP/s -> TFLOP/s (issued: 3 x TF32 split)
            algo = 2.0 * n_loc * n_total * cfg['ncomp'] / (kd['us_per_launch'] * 1e-6) / 1e12 / max(kd['launches_per_step'], 1)
            return {'bound': 'tensor', 'kernel': nm, 'achieved': algo, 'peak': bf16_peak / 2, 'unit': 'TFLOP/s',
                    'frac': algo / (bf16_peak / 2), 'traffic': None, 'peak_kind': peak_kind + ' bf16 / 2 (tf32 dense)',
                    'achieved_issued_3xtf32': issued, 'frac_issued': issued / (bf16_peak / 2),
                    'algorithmic': '2 n_local n p flop', 'us_per_launch': kd['us_per_launch']}
        ach = kd['achieved'] or 0.0
        rf = {'bound': 'hbm', 'kernel': nm, 'achieved': ach, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': ach / hbm_peak,
              'traffic': traffic.get(nm), 'peak_kind': peak_kind, 'us_per_launch': kd['us_per_launch'],
              'algorithmic_bytes_per_launch': kd['work_per_launch']}
        if nm == 'rotate_basis':
            flops = kd['work_per_launch'] / 8.0 * 2.0 * cfg['ncomp']      # ~2 k flop per basis element read
            rf['note'] = ('tall-skinny fp64 GEMM of a restart: bounded by the fp64 FMA pipe (64 FMA/clk/SM = 37 TFLOP/s), '
                          'not by HBM; ~%.1f TFLOP/s here' % (flops / (kd['us_per_launch'] * 1e-6) / 1e12))
        return rf

    roof = roofline_of(kernels[0])
    roof['event_overhead_us'] = event_overhead_us
    shown = ('csr_libsize_scale_log', 'csr_gene_moments', 'gene_moments_tiles', 'w_step_t', 'v_step_t', 'w_step', 'spmvT_csc',
             'v_step', 'rotate_basis', 'rotate_mma', 'knn_scores_tc', 'snn_rows')
    rooflines = [roofline_of(kd) for kd in kernels if kd['kernel'] in shown]

    # ---------------- CPU baseline + parity (rank 0, N=1) ----------------
    cpu = None
    parity = None
    if rank == 0 and world == 1:
        from oracle import adobo_oracle as O
        if not args.no_cpu_baseline:
            ns = min(SAMPLE_CELLS, n_total)
            sample = counts[:ns]
            ncores = host_threads()
            tm = {}
            t0 = time.perf_counter()
            O.embed_path(sample, cfg['ngenes'], cfg['ncomp'], cfg['k'], 0.067, 42, tm)
            dt = time.perf_counter() - t0
            cpu = {'value': ns / dt, 'unit': 'cells/s', 'cores': ncores, 'kind': 'port',
                   'sample': 'first %d of the %d cells, whole path once (%.1f s); numpy/scipy restatement, BLAS threads <= %d'
                             % (ns, n_total, dt, ncores),
                   'stage_s': {k: float(v) for k, v in tm.items()}}
        if args.parity != 'off':
            eng.upload_counts(counts)
            info_p = eng.embed_path(*path)
            t0 = time.perf_counter()
            try:
                parity = parity_block(args, cfg, eng, counts, info_p, v0)
            except AssertionError as e:
                parity = {'failed': str(e)[:300]}
            parity['seconds'] = time.perf_counter() - t0

    if rank == 0:
        line = {
            'metric': METRIC, 'value': value, 'unit': 'cells/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dev_ms / args.steps,
            'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': workload_config(cfg, world, args.config), 'clocks': clocks,
            'e2e': {'value': e2e_value, 'unit': 'cells/s', 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                    'ms_per_step': e2e_s / e2e_steps * 1e3, 'steps': e2e_steps,
                    'stage_ms_rank0': {'upload': stage[0] / e2e_steps * 1e3, 'path': stage[1] / e2e_steps * 1e3,
                                       'fetch': stage[2] / e2e_steps * 1e3}},
            'gpu_launches': int(launches), 'roofline': roof, 'cpu_baseline': cpu, 'parity': parity, 'api_e2e': api,
            'rooflines': rooflines, 'irlb': {'restarts': info['steps'], 'matvecs': info['nmult']}, 'edges': info['n_edges'],
            'nnz_rank0': int(counts.nnz), 'generate_s': t_gen,
            'kernels': kernels[:12],
        }
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
    eng.close()                                          # graphs + NCCL communicator go before torch's group
    if dist is not None:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------------------------
# c5: kNN / SNN sweep on a 1 M x 50 embedding
# --------------------------------------------------------------------------------------------------------------
C5 = dict(n=1000000, p=50, ks=(10, 20, 30), seed=5)


def run_c5(args, eng, dist, rank, world, local_rank, barrier, max_over_ranks):
    from adobo_b200.synth import synth_embedding
    n, p = (args.cells or C5['n']), C5['p']
    x = synth_embedding(n, p, seed=C5['seed'])
    bounds = shard_bounds(n, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    xl = np.ascontiguousarray(x[lo:hi])
    eng.upload_embedding(xl)
    hbm_peak, bf16_peak, peak_kind = peaks()
    sweep = []
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches = 0
    for k in ([int(x) for x in args.ks.split(',')] if args.ks else C5['ks']):
        for _ in range(args.warmup):
            eng.knn(None, k, fetch=False)
            eng.snn(None, k, 0.067, fetch=False)
        barrier()
        l0 = eng.launch_count()
        ms_knn = ms_snn = 0.0
        ne = 0
        for _ in range(args.steps):
            eng.l2_flush()
            eng.record(0)
            eng.knn(None, k, fetch=False)
            eng.record(1)
            _, _, _, ne = eng.snn(None, k, 0.067, fetch=False)
            eng.record(2)
            ms_knn += eng.elapsed_ms(0, 1)
            ms_snn += eng.elapsed_ms(1, 2)
        barrier()
        launches += eng.launch_count() - l0 - args.steps
        ms_knn = max_over_ranks(ms_knn) / args.steps
        ms_snn = max_over_ranks(ms_snn) / args.steps
        # end to end: host embedding in -> neighbour lists + edges out
        barrier()
        t0 = time.perf_counter()
        eng.upload_embedding(xl)
        nn = eng.knn(None, k)
        edges, _, _, _ = eng.snn(None, k, 0.067)
        eng.sync()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        sweep.append({'k': k, 'ms_knn': ms_knn, 'ms_snn': ms_snn, 'cells_per_s': n / ((ms_knn + ms_snn) * 1e-3),
                      'knn_tflops_algorithmic_2n2p': 2.0 * n * n * p / world / (ms_knn * 1e-3) / 1e12,
                      'knn_frac_of_tf32_peak': 2.0 * n * n * p / world / (ms_knn * 1e-3) / 1e12 / (bf16_peak / 2),
                      'edges_rank0': int(ne), 'e2e_cells_per_s': n / e2e_s,
                      'h2d_bytes': int(xl.nbytes), 'd2h_bytes': int(nn.nbytes + edges.nbytes)})
        del nn, edges
    clocks = sampler.stop() if sampler else None
    if rank == 0:
        s0 = sweep[0]
        line = {'metric': 'cells/sec kNN -> SNN graph (C5 sweep; value at k=10)', 'value': s0['cells_per_s'], 'unit': 'cells/s',
                'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': s0['ms_knn'] + s0['ms_snn'],
                'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'tf32x3+f64', 'data': 'synthetic',
                'config': {'workload': 'kNN/SNN graph sweep on a %d x %d PCA embedding (16 Gaussian blobs), k = 10/20/30 '
                                       '(BASELINE configs[4]); queries row-sharded x%d, references all-gathered' % (n, p, world),
                           'config': 'c5', 'cells_total': n, 'p': p, 'l2': 'L2 flushed (256 MB memset) between timed steps'},
                'clocks': clocks, 'sweep': sweep,
                'e2e': {'value': s0['e2e_cells_per_s'], 'unit': 'cells/s', 'h2d_bytes_per_step': s0['h2d_bytes'],
                        'd2h_bytes_per_step': s0['d2h_bytes']},
                'gpu_launches': int(launches),
                'roofline': {'bound': 'tensor', 'kernel': 'knn_scores_tc', 'achieved': s0['knn_tflops_algorithmic_2n2p'],
                             'peak': bf16_peak / 2, 'unit': 'TFLOP/s', 'frac': s0['knn_frac_of_tf32_peak'], 'traffic': None,
                             'peak_kind': peak_kind + ' bf16 / 2 (tf32 dense)',
                             'note': 'algorithmic 2 n^2 p flop over the whole kNN stage (shortlist + fp64 re-rank); the tensor '
                                     'kernel issues 3 x that (TF32 operand split for fp32 accuracy)'},
                'cpu_baseline': None}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
    eng.close()
    if dist is not None:
        dist.destroy_process_group()


def run_reference_c5(args):
    """CPU arm of the sweep: the oracle's exact kNN on a bounded sample of queries (2,048 of the 1 M, against all
    references), extrapolated linearly in the number of queries, plus its SNN on the sampled rows' share."""
    from adobo_b200.synth import synth_embedding
    from oracle import adobo_oracle as O
    n, p = (args.cells or C5['n']), C5['p']
    x = synth_embedding(n, p, seed=C5['seed'])
    q = 2048
    sq = np.einsum('ij,ij->i', x, x)
    t0 = time.perf_counter()
    for lo in range(0, q, 256):
        d = sq[lo:lo + 256, None] + sq[None, :] - 2.0 * x[lo:lo + 256] @ x.T
        cand = np.argpartition(d, 26, axis=1)[:, :26]
        dd = np.einsum('ijk,ijk->ij', x[lo:lo + 256][:, None, :] - x[cand], x[lo:lo + 256][:, None, :] - x[cand])
        np.lexsort((cand, dd), axis=1)
    dt = time.perf_counter() - t0
    value = q / dt
    line = {'impl': 'reference', 'metric': 'cells/sec kNN -> SNN graph (C5 sweep; value at k=10)', 'value': value,
            'unit': 'cells/s', 'n_gpus': args.gpus, 'steps': 1, 'warmup': 0, 'ms_per_step': dt * 1e3, 'higher_is_better': True,
            'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': 'kNN/SNN graph sweep on a %d x %d PCA embedding' % (n, p), 'config': 'c5'},
            'cpu_baseline': {'value': value, 'unit': 'cells/s', 'cores': host_threads(), 'kind': 'port',
                             'sample': 'exact kNN of %d of the %d queries against all references (queries/s; SNN not included)' % (q, n)},
            'e2e': {'value': value, 'unit': 'cells/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}, 'gpu_launches': 0}
    print(json.dumps(line))


if __name__ == '__main__':
    main()
