import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from adobo_b200.engine import Engine
from adobo_b200.synth import CountModel
cfg = bench.CONFIGS['c3']
eng = Engine(0)
eng.synth_counts(CountModel(cfg['genes'], cfg['density'], cfg['seed']), 0, cfg['cells'])
np.random.seed(42)
v0 = np.random.randn(cfg['ngenes'])
for _ in range(2):
    eng.embed_path(v0, 10000, cfg['ngenes'], cfg['ncomp'], cfg['k'], 0.067)
eng.profile(True)
eng.embed_path(v0, 10000, cfg['ngenes'], cfg['ncomp'], cfg['k'], 0.067)
prof = eng.profile_read()
eng.profile(False)
ov = prof.pop('_event_overhead')
o = ov['ms'] / ov['launches']
tot = 0
for nm, v in sorted(prof.items(), key=lambda kv: -kv[1]['ms']):
    net = v['ms'] - o * v['launches'] + 0.001 * v['launches']
    tot += net
    print('%-24s %8.3f ms (net %7.3f) x%d' % (nm, v['ms'], net, v['launches']))
print('sum net', tot, 'event overhead us', o * 1e3)
