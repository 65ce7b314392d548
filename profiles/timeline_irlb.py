"""Timeline of the last outer iteration of IRLB inside the replayed CUDA graph (diagnostic build:
ADOBO_B200_CFLAGS=-DADOBO_B200_TIMELINE python -m adobo_b200.build --force; run with ADOBO_B200_TIMELINE=1)."""
import os
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adobo_b200.engine import Engine
from adobo_b200.synth import synth_counts

eng = Engine(0)
if len(sys.argv) > 1 and sys.argv[1] in ('c2', 'c3'):        # a bench.py configuration, generated on the device
    import bench
    from adobo_b200.synth import CountModel
    cfg = bench.CONFIGS[sys.argv[1]]
    if cfg['gen'] == 'host':
        eng.upload_counts(synth_counts(cfg['cells'], cfg['genes'], cfg['density'], cfg['seed']))
    else:
        eng.synth_counts(CountModel(cfg['genes'], cfg['density'], cfg['seed']), 0, cfg['cells'])
else:
    cells = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    counts = synth_counts(cells, 3000, 0.05, 77)
    eng.upload_counts(counts)
np.random.seed(42)
v0 = np.random.randn(1000)
for rep in range(2):
    r = eng.embed_path(v0, ngenes=1000, ncomp=75, k=10)
print(r)
