"""Wall time of every call of the reference-style Python sequence on the C3 shape (where the host layer spends it)."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import adobo_b200 as ad  # noqa: E402
from adobo_b200 import engine as E  # noqa: E402
from adobo_b200.engine import Engine  # noqa: E402
from adobo_b200.synth import CountModel  # noqa: E402

cfg = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else 'c3']
eng = Engine(0)
eng.synth_counts(CountModel(cfg['genes'], cfg['density'], cfg['seed']), 0, cfg['cells'])
m = eng.download_counts()
E.set_default_engine(eng)


def run(tag):
    t = [time.perf_counter()]
    exp = ad.data.dataset.from_csr(m, frame=False); t.append(time.perf_counter())
    ad.normalize.norm(exp); t.append(time.perf_counter())
    ad.hvg.find_hvg(exp, ngenes=cfg['ngenes']); t.append(time.perf_counter())
    ad.dr.pca(exp, ncomp=cfg['ncomp']); t.append(time.perf_counter())
    ad.clustering.generate(exp, k=cfg['k'], clust_alg=None); t.append(time.perf_counter())
    names = ['from_csr', 'norm', 'find_hvg', 'pca', 'generate']
    print(tag, ' '.join('%s %.1f ms' % (n, (b - a) * 1e3) for n, a, b in zip(names, t[:-1], t[1:])), 'total %.1f ms' % ((t[-1] - t[0]) * 1e3))


run('first ')
run('second')
pr = cProfile.Profile()
pr.enable()
run('profiled')
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
