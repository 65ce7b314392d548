"""ncu launch list (--metrics gpu__time_duration.sum --csv --log-file X.csv) -> markdown table by kernel.
    python profiles/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches.md"""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline='') as f:
    lines = [ln for ln in f if not ln.startswith('==')]
rd = csv.reader(lines)
hdr = None
for r in rd:
    if hdr is None:
        if 'Kernel Name' in r:
            hdr = r
        continue
    if len(r) == len(hdr):
        rows.append(dict(zip(hdr, r)))
agg = defaultdict(list)
for r in rows:
    if r.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    v = float(r['Metric Value'].replace(',', ''))
    unit = r.get('Metric Unit', 'ns')
    us = v / 1e3 if unit in ('ns', 'nsecond') else (v if unit in ('us', 'usecond') else v * 1e3)
    name = re.sub(r'^(void )?(adobo::)?', '', r['Kernel Name'])
    name = re.sub(r'\(.*$', '', name)[:60]
    agg[name].append(us)
tot = sum(sum(v) for v in agg.values())
print('Sum of kernel time: %.1f ms over %d launches.\n' % (tot / 1e3, sum(len(v) for v in agg.values())))
print('| kernel | launches | total us | avg us | min us | max us | share |')
print('|---|---|---|---|---|---|---|')
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    if sum(v) < 0.0005 * tot:
        continue
    print('| %s | %d | %.1f | %.2f | %.2f | %.2f | %.1f%% |' % (k, len(v), sum(v), sum(v) / len(v), min(v), max(v), 100 * sum(v) / tot))
