"""Small driver for ncu captures of the IRLB kernels: one device-resident pass of the hot path on a
20,000-cell x 3,000-gene synthetic matrix (same n, h=1000, ncomp=75 as the C2 bench workload; fewer
genes only to shorten input generation)."""
import os
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adobo_b200.engine import Engine
from adobo_b200.synth import synth_counts

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
eng = Engine(0)
counts = synth_counts(cells, 3000, 0.10, 77)
eng.upload_counts(counts)
np.random.seed(42)
v0 = np.random.randn(1000)
r = eng.embed_path(v0, ngenes=1000, ncomp=75, k=10)
print(r)
