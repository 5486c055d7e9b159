#!/bin/bash
for v in "" b8 b10; do
  if [ -n "$v" ]; then export FQGPU_LIB=$PWD/seq-collection_b200/variants/libfqgpu_$v.so; else unset FQGPU_LIB; fi
  echo "== variant '$v'"
  timeout 600 python tools/bgzf_time.py 4000000 2>&1 | grep "bgzf device"
done
unset FQGPU_LIB
timeout 900 python -m pytest tests/test_gpu_bgzf.py tests/test_gpu_gzip.py -x -q -m gpu 2>&1 | tail -3
