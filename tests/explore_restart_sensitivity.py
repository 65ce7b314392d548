"""Not a test (not collected): how sensitive is the IRLB restart count of the C2 bench workload to
rounding?  Runs the CPU oracle's lanczos on the bench's own matrix with the start vector perturbed
at the 1e-15 level, and with one- vs two-pass Gram-Schmidt, and prints restarts / mat-vecs.

    python tests/explore_restart_sensitivity.py [cells] [nperturb]
"""
import os
import sys
import time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adobo_b200.synth import synth_counts
from oracle import adobo_oracle as O
import scipy.sparse as sp

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
npert = int(sys.argv[2]) if len(sys.argv) > 2 else 4
counts = synth_counts(cells, 25000, 0.064, 43)
norm, _ = O.normalize_standard(counts)
hv = O.seurat(norm, 1000)
genes = np.sort(np.asarray(hv))
sub = sp.csr_matrix(norm)[:, genes]
mu, sd = O.scale_stats(sub)
op = O.ScaledCenteredOp(sub, mu, sd)
np.random.seed(42)
v0 = np.random.randn(len(genes))
rng = np.random.default_rng(0)
cgs1 = O._cgs


def cgs2(y, basis, ar):
    return cgs1(cgs1(y, basis, ar), basis, ar)


for name, fn in (('one-pass CGS (reference)', cgs1), ('two-pass CGS', cgs2)):
    O._cgs = fn
    for p in range(npert):
        v = v0 * (1.0 + (0 if p == 0 else 1e-15) * rng.standard_normal(len(v0)))
        t0 = time.perf_counter()
        r = O.lanczos(op, 75, v0=v)
        print('%-26s perturbation %d: restarts %d, mat-vecs %d, %.1f s' % (name, p, r['steps'], r['nmult'],
                                                                         time.perf_counter() - t0), flush=True)
O._cgs = cgs1

# the GPU path stores the normalised values in fp32 (computes in fp64): same experiment on that matrix
sub32 = sub.copy()
sub32.data = sub32.data.astype(np.float32).astype(np.float64)
mu32, sd32 = O.scale_stats(sub32)
op32 = O.ScaledCenteredOp(sub32, mu32, sd32)
for p in range(npert):
    v = v0 * (1.0 + (0 if p == 0 else 1e-15) * rng.standard_normal(len(v0)))
    r = O.lanczos(op32, 75, v0=v)
    print('fp32-stored values         perturbation %d: restarts %d, mat-vecs %d' % (p, r['steps'], r['nmult']), flush=True)

# per-restart trace of the reference algorithm (compare with profiles/trace_irlb.py on the GPU)
tr = []
O.lanczos(op32, 75, v0=v0, trace=tr)
for e in tr:
    print('oracle it=%d k=%d conv=%d' % (e['it'], e['k'], e['conv']))
