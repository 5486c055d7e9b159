#!/bin/bash
for v in "" w256 w128 w64; do
  if [ -n "$v" ]; then export FQGPU_LIB=$PWD/seq-collection_b200/variants/libfqgpu_$v.so; else unset FQGPU_LIB; fi
  for r in 12500000 25000000 100000000; do
  echo -n "variant '$v' "
  timeout 600 python bench.py --records $r --steps 20 --warmup 5 --no-ont --no-gz --no-ingest --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('GB',d['config']['bytes']/1e9,'value',round(d['value'],1),'ms',round(d['ms_per_step'],4),'scan-only',round(d['roofline']['achieved'],1))"
  done
done
