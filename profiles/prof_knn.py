"""kNN stage alone on a synthetic embedding: per-kernel device times of the shortlist / re-rank kernels.

    python profiles/prof_knn.py <n> <p> <k> [passes]      (the command ncu wraps for knn_scores_tc)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from adobo_b200.engine import Engine  # noqa: E402
from adobo_b200.synth import synth_embedding  # noqa: E402

n, p, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
passes = int(sys.argv[4]) if len(sys.argv) > 4 else 3
eng = Engine(0)
eng.upload_embedding(synth_embedding(n, p, seed=7))
eng.knn(None, k, fetch=False)
ms = []
for _ in range(passes):
    eng.l2_flush()
    eng.record(0)
    eng.knn(None, k, fetch=False)
    eng.record(1)
    ms.append(eng.elapsed_ms(0, 1))
eng.profile(True)
eng.knn(None, k, fetch=False)
prof = eng.profile_read()
eng.profile(False)
print('n=%d p=%d k=%d  knn %.3f ms (min %.3f)  2n^2p/t = %.1f TFLOP/s' % (n, p, k, np.mean(ms), min(ms), 2.0 * n * n * p / (min(ms) * 1e-3) / 1e12))
for name, d in prof.items():
    print('   %-22s %9.3f ms  x%d' % (name, d['ms'], d['launches']))
eng.close()
