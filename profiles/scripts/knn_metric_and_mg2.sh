#!/bin/bash
# new metric kNN tests + a timing at 68 k points, then the 2-rank parity check and C3 bench line
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "other_metrics or fp64_brute" > gpurun_out/r02g_knn_metric_test.log 2>&1
echo "metric tests rc=$?"; tail -3 gpurun_out/r02g_knn_metric_test.log
python - > gpurun_out/r02g_knn_metric_time.log 2>&1 <<'PY'
import time, numpy as np
from adobo_b200.engine import Engine
from adobo_b200.synth import synth_embedding
eng = Engine(0)
x = synth_embedding(68000, 75, seed=5)
for m, nm in ((0, 'euclidean (tensor path)'), (1, 'manhattan'), (2, 'chebyshev'), (4, 'canberra'), (6, 'euclidean fp64 brute force')):
    eng.knn(x, 10, metric=m)
    t0 = time.perf_counter(); eng.knn(x, 10, metric=m); dt = time.perf_counter() - t0
    print('%-28s 68000 x 75, k=10: %.1f ms (host buffers in and out)' % (nm, dt * 1e3))
PY
cat gpurun_out/r02g_knn_metric_time.log
bash profiles/scripts/mg.sh 2
