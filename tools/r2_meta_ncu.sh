#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/meta_launches.csv python bench.py --steps 1 --warmup 1 --records 20000000 --meta-records 20000000 --no-ont --no-gz --no-ingest --no-e2e --no-cpu-baseline > gpurun_out/meta_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(l for l in open('gpurun_out/meta_launches.csv') if l.startswith('"'))]
h = rows[0]; ki = h.index('Kernel Name'); vi = h.index('Metric Value')
agg = collections.OrderedDict()
for r in rows[1:]:
    try: v = float(r[vi].replace(',', ''))
    except ValueError: continue
    a = agg.setdefault(r[ki].split('(')[0], [0, 0.0, 0.0]); a[0] += 1; a[1] += v; a[2] = max(a[2], v)
for k, (n, t, mx) in agg.items(): print(f"{k:40s} n={n:4d} total={t/1e6:9.3f} ms max={mx/1e6:8.3f}")
PY
