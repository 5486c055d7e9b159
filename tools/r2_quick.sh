#!/bin/bash
# quick GPU iteration: kernel timings, a parity subset (whole + tiny spans), optional ncu capture ($1 = tag)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
rm -f gpurun_out/kbench.log
for args in "" "--core" "--workload ont" "--workload ont --core"; do
  timeout 180 python tools/kbench.py --mb 8192 --reps 5 $args >> gpurun_out/kbench.log 2>&1
done
timeout 600 python -m pytest tests/test_gpu_scan_tiles.py tests/test_gpu_parity.py -q -m gpu --timeout 300 -x -k "not tiny_spans and not two_gigabytes" 2>&1 | tail -15 > gpurun_out/t1.log
FQGPU_SPAN_MIN_TILES=1 timeout 600 python -m pytest tests/test_gpu_scan_tiles.py tests/test_gpu_shards.py -q -m gpu --timeout 300 -x -k "not tiny_spans" 2>&1 | tail -15 > gpurun_out/t2.log
if [ -n "$1" ]; then
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:fq_scan_kernel -c 1 -o gpurun_out/r2_$1_full python tools/kbench.py --mb 2048 --reps 1 > gpurun_out/ncu_full.log 2>&1
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:fq_scan_kernel -c 1 -o gpurun_out/r2_$1_core python tools/kbench.py --mb 2048 --reps 1 --core > gpurun_out/ncu_core.log 2>&1
fi
cat gpurun_out/kbench.log; tail -4 gpurun_out/t1.log; tail -4 gpurun_out/t2.log
