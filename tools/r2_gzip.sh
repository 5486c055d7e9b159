#!/bin/bash
# on-device gzip: tests, timing, per-kernel launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gzip.py tests/test_gpu_bgzf.py -x -q -m gpu > gpurun_out/gz_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/gz_tests.log
tail -15 gpurun_out/gz_tests.log
timeout 600 python tools/gz_time.py 4000000 100 > gpurun_out/gz_time.log 2>&1; tail -8 gpurun_out/gz_time.log
GZ_HOST=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/gz_launches.csv python tools/gz_time.py 1000000 100 > gpurun_out/gz_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(l for l in open('gpurun_out/gz_launches.csv') if l.startswith('"'))]
h = rows[0]; ki = h.index('Kernel Name'); vi = h.index('Metric Value')
agg = collections.OrderedDict()
for r in rows[1:]:
    try: v = float(r[vi].replace(',', ''))
    except ValueError: continue
    a = agg.setdefault(r[ki].split('(')[0], [0, 0.0]); a[0] += 1; a[1] += v
for k, (n, t) in agg.items(): print(f"{k:40s} n={n:4d} total={t/1e6:9.3f} ms")
PY
