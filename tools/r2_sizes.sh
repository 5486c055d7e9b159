#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out; rm -f gpurun_out/kbench.log
for mb in 4300 34332; do
  for args in "" "--core"; do
    timeout 300 python tools/kbench.py --mb $mb --reps 7 $args >> gpurun_out/kbench.log 2>&1
  done
done
timeout 300 python tools/kbench.py --mb 19000 --reps 5 --workload ont >> gpurun_out/kbench.log 2>&1
timeout 600 python -m pytest tests/test_gpu_scan_tiles.py -q -m gpu --timeout 300 -x -k "not tiny_spans" 2>&1 | tail -3 >> gpurun_out/kbench.log
cat gpurun_out/kbench.log
