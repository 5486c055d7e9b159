#!/bin/bash
mkdir -p gpurun_out
for v in "" m8 m7; do
  if [ -n "$v" ]; then export FQGPU_LIB=$PWD/seq-collection_b200/variants/libfqgpu_$v.so; else unset FQGPU_LIB; fi
  echo "== variant '$v'"
  GZ_HOST=0 timeout 600 python tools/gz_time.py 4000000 100 2>&1 | grep device
done
unset FQGPU_LIB
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
