#!/usr/bin/env python
"""bench.py -- cells/s of adobo's cell-embedding hot path on B200.

    python bench.py --gpus N --steps K --warmup W          (N>1: launched by torchrun, one rank/GPU)
    python bench.py --impl reference ...                   (CPU arm: the oracle port on host cores)

A "step" is one pass of the whole path normalize -> HVG(seurat) -> IRLB PCA -> kNN -> SNN over one
synthetic count matrix.  At N=1 the workload is BASELINE.json configs[1]: choroid-plexus-shape
counts, 20,000 cells x 25,000 genes, 6.4 % non-zeros, 1000 HVG, 75 components, k=10.  At N>1 every
rank holds its own 20,000-cell block of one N*20,000-cell matrix (weak scaling); IRLB all-reduces
over NCCL, kNN/SNN all-gather.

`value`    device-resident throughput: counts already in HBM when the timed region starts.
`e2e`      same metric through the host-facing C-ABI calls: pinned host CSR in, PCA + graph out,
           copies inside the timed region (wall clock, max over ranks).
`roofline` the dominant kernel of a step (by CUDA-event time, measured in extra profiled steps
           right after the timed region): algorithmic bytes / average launch time vs the measured
           HBM copy peak.
`cpu_baseline` the oracle port (numpy/scipy restatement of the reference) on this box's host cores.
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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--cells', type=int, default=20000, help='cells per GPU')
    ap.add_argument('--genes', type=int, default=25000)
    ap.add_argument('--density', type=float, default=0.064)
    ap.add_argument('--ngenes', type=int, default=1000)
    ap.add_argument('--ncomp', type=int, default=75)
    ap.add_argument('--k', type=int, default=10)
    ap.add_argument('--seed', type=int, default=43)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--profile-steps', type=int, default=2)
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


def run_reference(args, rank, world):
    """CPU arm: the oracle port (the reference is pure Python + pandas and cannot travel to the GPU
    box; oracle/adobo_oracle.py restates its arithmetic on scipy.sparse and is pinned to the
    reference's own outputs by tests/golden).  Each step = the same workload as the GPU arm."""
    if rank != 0:
        return
    from adobo_b200.synth import synth_counts
    from oracle import adobo_oracle as O
    counts = synth_counts(args.cells, args.genes, args.density, args.seed)
    times, tm = [], {}
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        O.embed_path(counts, args.ngenes, args.ncomp, args.k, 0.067, 42, tm)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    total = float(np.sum(times))
    value = args.cells * args.steps / total
    cores = os.cpu_count()
    line = {
        'impl': 'reference', 'metric': 'cells/sec normalize->HVG->IRLB PCA->kNN/SNN', 'value': value, 'unit': 'cells/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': total / args.steps * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(args, 1),
        'cpu_baseline': {'value': value, 'unit': 'cells/s', 'cores': cores, 'kind': 'port',
                         'sample': 'whole workload (%d cells), numpy/scipy with up to %d BLAS threads' % (args.cells, cores),
                         'stage_s': {k: float(v) for k, v in tm.items()}},
        'e2e': {'value': value, 'unit': 'cells/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


def workload_config(args, world):
    return {'workload': 'choroid-plexus-shape synthetic counts (BASELINE configs[1]): %d cells x %d genes per GPU, '
                        '%.1f%% non-zeros, standard norm, seurat HVG %d, irlb PCA %d comps, knn/snn k=%d prune 0.067'
                        % (args.cells, args.genes, args.density * 100, args.ngenes, args.ncomp, args.k),
            'cells_total': args.cells * world, 'genes': args.genes, 'ngenes': args.ngenes, 'ncomp': args.ncomp,
            'k': args.k, 'parallelism': 'cells row-sharded x%d' % world,
            'l2': 'L2 flushed (256 MB memset) between timed steps'}


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
    from adobo_b200.synth import synth_counts

    eng = Engine(local_rank)
    if world > 1:
        eng.init_comm_from_torch()
    n_total = args.cells * world
    counts = synth_counts(n_total, args.genes, args.density, args.seed, rank * args.cells, (rank + 1) * args.cells)
    indptr = np.ascontiguousarray(counts.indptr, dtype=np.int64)
    indices = np.ascontiguousarray(counts.indices, dtype=np.int32)
    data = np.ascontiguousarray(counts.data, dtype=np.int32)
    np.random.seed(42)                                   # dr.pca default seed (dr.py:237)
    v0 = np.random.randn(args.ngenes)

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

    # ---------------- device-resident throughput ----------------
    eng.upload_counts(counts)
    info = None
    for _ in range(args.warmup):
        eng.l2_flush()
        info = eng.embed_path(v0, 10000, args.ngenes, args.ncomp, args.k, 0.067)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    l0 = eng.launch_count()
    dev_ms = 0.0
    for _ in range(args.steps):
        eng.l2_flush()
        eng.record(0)
        info = eng.embed_path(v0, 10000, args.ngenes, args.ncomp, args.k, 0.067)
        eng.record(1)
        dev_ms += eng.elapsed_ms(0, 1)
    barrier()
    launches = eng.launch_count() - l0 - args.steps      # minus the L2-flush memsets
    clocks = sampler.stop() if sampler else None
    dev_ms = max_over_ranks(dev_ms)
    value = n_total * args.steps / (dev_ms * 1e-3)

    # ---------------- end to end through the host-facing calls ----------------
    for a in (indptr, indices, data):
        eng.pin(a)
    h = args.ngenes
    e2e_s, d2h = 0.0, 0
    for it in range(2 + args.steps):
        barrier()
        t0 = time.perf_counter()
        eng.lib.adobo_upload_counts(eng.h, counts.shape[0], counts.shape[1], indptr, indices, data)
        r = eng.embed_path(v0, 10000, args.ngenes, args.ncomp, args.k, 0.067)
        pca = eng.fetch_pca(h, args.ncomp)
        edges = eng.fetch_edges(r['n_edges'])
        eng.sync()
        dt = time.perf_counter() - t0
        dt = max_over_ranks(dt)
        if it >= 2:
            e2e_s += dt
            d2h = pca['comp'].nbytes + pca['contr'].nbytes + pca['s'].nbytes + edges.nbytes
    e2e_value = n_total * args.steps / e2e_s
    h2d = indptr.nbytes + indices.nbytes + data.nbytes + v0.nbytes
    for a in (indptr, indices, data):
        eng.unpin(a)

    # ---------------- per-kernel breakdown (profiled steps, outside the timed region) ----------------
    eng.profile(True)
    for _ in range(args.profile_steps):
        eng.l2_flush()
        eng.embed_path(v0, 10000, args.ngenes, args.ncomp, args.k, 0.067)
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
                        'us_per_launch_with_events': per_launch_us, 'achieved': ach})
    kernels.sort(key=lambda d: -d['ms_per_step'])
    try:
        with open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')) as f:
            traffic = json.load(f)
    except Exception:
        traffic = {}

    def roofline_of(kd):
        nm = kd['kernel']
        if nm in ('knn_scores_tc', 'knn_shortlist'):
            ach_t = (kd['achieved'] or 0) / 1e3              # GFLOP/s -> TFLOP/s
            return {'bound': 'tensor', 'kernel': nm, 'achieved': ach_t, 'peak': bf16_peak / 2, 'unit': 'TFLOP/s',
                    'frac': ach_t / (bf16_peak / 2), 'traffic': None, 'peak_kind': peak_kind + ' bf16 / 2 (tf32 dense)',
                    'us_per_launch': kd['us_per_launch']}
        ach = kd['achieved'] or 0.0
        return {'bound': 'hbm', 'kernel': nm, 'achieved': ach, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': ach / hbm_peak,
                'traffic': traffic.get(nm.replace('_v', '') if nm.endswith('_v') else nm), 'peak_kind': peak_kind,
                'us_per_launch': kd['us_per_launch']}

    # the dominant kernel of the step (CUDA-event time, includes ~6 us of event overhead per launch), then the
    # bandwidth-type kernels of the path for context
    roof = roofline_of(kernels[0])
    if kernels[0]['kernel'] in ('svd_arrow', 'jacobi_svd'):
        roof['note'] = ('m_b x m_b SVD of B: bounded by dependent fp64 latency, not by HBM; algorithmic bytes = 24 m_b^2')
    elif kernels[0]['kernel'] in ('w_step', 'spmv_dots', 'spmvT_csc', 'v_step', 'update_norm'):
        roof['note'] = ('Lanczos step kernel at %d cells per GPU: the subset matrix (CSR + CSC) and the basis stay in the 126 MB '
                        'L2 (ncu, warm: < 100 KB of DRAM traffic per launch), so the bytes are served by L2 and the launch is '
                        'bounded by dependent-load latency, shared-memory gathers and the cross-CTA reduction, not by HBM; '
                        'us_per_launch is the CUDA-event time minus the calibrated event overhead (%.1f us, empty kernel '
                        'between the same events); in-graph durations are in profiles/r01_timeline_irlb.md and the HBM-bound '
                        'regime of the same kernels in profiles/r01_scale_rooflines.md' % (args.cells, event_overhead_us))
    roof['event_overhead_us'] = event_overhead_us
    roof['empty_kernel_between_events_us'] = (ov['ms'] * 1e3 / max(ov['launches'], 1)) if ov else None
    rooflines = [roofline_of(kd) for kd in kernels if kd['kernel'] in
                 ('csr_libsize_scale_log', 'csr_gene_moments', 'w_step', 'spmv_dots', 'spmvT_csc', 'update_norm', 'v_step',
                  'rotate_basis', 'knn_scores_tc')]

    # ---------------- CPU baseline (rank 0, N=1) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import adobo_oracle as O
        tm = {}
        t0 = time.perf_counter()
        O.embed_path(counts, args.ngenes, args.ncomp, args.k, 0.067, 42, tm)
        dt = time.perf_counter() - t0
        cpu = {'value': args.cells / dt, 'unit': 'cells/s', 'cores': os.cpu_count(), 'kind': 'port',
               'sample': 'whole workload once (%d cells, %.1f s); numpy/scipy restatement, BLAS threads <= %d'
                         % (args.cells, dt, os.cpu_count()),
               'stage_s': {k: float(v) for k, v in tm.items()}}

    if rank == 0:
        line = {
            'metric': 'cells/sec normalize->HVG->IRLB PCA->kNN/SNN', 'value': value, 'unit': 'cells/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dev_ms / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': workload_config(args, world), 'clocks': clocks,
            'e2e': {'value': e2e_value, 'unit': 'cells/s', 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                    'ms_per_step': e2e_s / args.steps * 1e3},
            'gpu_launches': int(launches), 'roofline': roof, 'cpu_baseline': cpu,
            'rooflines': rooflines, 'irlb': {'restarts': info['steps'], 'matvecs': info['nmult']}, 'edges': info['n_edges'],
            'kernels': kernels[:12],
        }
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
    eng.close()                                          # graphs + NCCL communicator go before torch's group
    if dist is not None:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
