#!/bin/bash
# one optimisation iteration: GPU tests, then the default bench line (C3, N=1)
TAG=${1:-it}
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_gputest.log 2>&1
echo "gputest rc=$?"; tail -3 gpurun_out/${TAG}_gputest.log
python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench_c3_n1.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; tail -2 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/${TAG}_bench_c3_n1.json') if l.startswith('{')][0])
print('ms_per_step %.3f e2e %.3f restarts %s parity %s' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['irlb'], json.dumps(d['parity'])[:400]))
print('api', d['api_e2e'])
for k in d['kernels']: print('  %-24s ms/step %7.3f launches %6.1f us/launch %8.2f' % (k['kernel'], k['ms_per_step'], k['launches_per_step'], k['us_per_launch']))
PY
