"""Per-restart trace (k, converged count) of the IRLB on the C2 bench matrix, for comparison with the CPU
restatement's trace (tests/explore_restart_sensitivity.py).  ADOBO_B200_IRLB=trace python profiles/trace_irlb.py"""
import os
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adobo_b200.engine import Engine
from adobo_b200.synth import synth_counts

eng = Engine(0)
counts = synth_counts(20000, 25000, 0.064, 43)
eng.upload_counts(counts)
np.random.seed(42)
v0 = np.random.randn(1000)
rng = np.random.default_rng(0)
for p in range(3):
    v = v0 * (1.0 + (0 if p == 0 else 1e-15) * rng.standard_normal(len(v0)))
    print('== run', p, file=sys.stderr, flush=True)
    r = eng.embed_path(v, ngenes=1000, ncomp=75, k=10)
    print('== result', {k_: r[k_] for k_ in r if k_ in ('steps', 'nmult', 'svd_fallbacks')}, file=sys.stderr, flush=True)
