"""Per-kernel achieved bandwidth at a C4-shard-like size (BASELINE configs[3]: 1.3 M cells over 8 GPUs =
162 k cells per GPU): 160,000 cells x 28,000 genes, 1,400 non-zeros per cell (224 M non-zeros), structured
synthetic pattern (fast to generate on the host).  CUDA-event time per kernel class, algorithmic bytes
from the library's own accounting."""
import os
import sys
import json
import numpy as np
import scipy.sparse as sp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adobo_b200.engine import Engine

n, g, per_row = 160_000, 28_000, 1400
rng = np.random.default_rng(7)
start = rng.integers(0, 20, n)
cols = (start[:, None] + 19 * np.arange(per_row)[None, :] + rng.integers(0, 3, (n, per_row))).astype(np.int32)
lam = np.exp(rng.normal(0, 1.0, g))[cols] * np.exp(rng.normal(0, 0.3, n))[:, None] * 0.6
counts = (1 + rng.poisson(lam)).astype(np.int32)
indptr = np.arange(0, n * per_row + 1, per_row, dtype=np.int64)
m = sp.csr_matrix((counts.ravel(), cols.ravel(), indptr), shape=(n, g))
del lam
eng = Engine(0)
eng.upload_counts(m)
hbm = 6550.4
try:
    hbm = json.load(open(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')))['hbm_gbs']
except Exception:
    pass
out = {}
for rep in range(3):
    eng.l2_flush()
    eng.profile(True)
    eng.normalize(10000, 1, 1.0, want_values=False)
    eng.gene_moments()
    hv, z = eng.hvg_seurat(1000, 20)
    r = eng.irlb(hv, ncomp=50, maxit=4, seed=42)
    pr = eng.profile_read()
    eng.profile(False)
    for k, v in pr.items():
        if v['work'] > 0 and v['ms'] > 0:
            out[k] = dict(us_per_launch=v['ms'] * 1e3 / v['launches'], launches=v['launches'],
                          GBps=v['work'] / (v['ms'] * 1e-3) / 1e9)
print('n=%d g=%d nnz=%d; subset nnz per launch accounted inside the library' % (n, g, m.nnz))
for k, v in sorted(out.items(), key=lambda kv: -kv[1]['GBps']):
    print('%-24s %9.1f us/launch  %8.1f GB/s  %5.1f %% of measured HBM peak (%.0f GB/s)' % (k, v['us_per_launch'], v['GBps'], 100 * v['GBps'] / hbm, hbm))
