"""One configuration of bench.py, device resident, a fixed number of passes -- the command ncu wraps.

    python profiles/prof_path.py [c2|c3] [passes]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from adobo_b200.engine import Engine  # noqa: E402
from adobo_b200.synth import CountModel, synth_counts  # noqa: E402

cfg = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else 'c3']
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 2
eng = Engine(0)
if cfg['gen'] == 'host':
    eng.upload_counts(synth_counts(cfg['cells'], cfg['genes'], cfg['density'], cfg['seed']))
else:
    eng.synth_counts(CountModel(cfg['genes'], cfg['density'], cfg['seed']), 0, cfg['cells'])
np.random.seed(42)
v0 = np.random.randn(cfg['ngenes'])
for _ in range(passes):
    r = eng.embed_path(v0, 10000, cfg['ngenes'], cfg['ncomp'], cfg['k'], 0.067)
print('restarts %d mat-vecs %d edges %d launches %d' % (r['steps'], r['nmult'], r['n_edges'], eng.launch_count()))
eng.close()
