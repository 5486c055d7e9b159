#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out; rm -f gpurun_out/kbench.log
for tw in $VARIANTS; do
  export FQGPU_LIB=$PWD/seq-collection_b200/variants/libfqgpu_$tw.so
  echo "== tw=$tw" >> gpurun_out/kbench.log
  for args in "" "--core" "--workload ont"; do
    timeout 180 python tools/kbench.py --mb 8192 --reps 5 $args >> gpurun_out/kbench.log 2>&1
  done
  timeout 300 python -m pytest tests/test_gpu_scan_tiles.py -q -m gpu --timeout 300 -x -k "not tiny_spans" 2>&1 | tail -2 >> gpurun_out/kbench.log
  FQGPU_SPAN_MIN_TILES=1 timeout 300 python -m pytest tests/test_gpu_scan_tiles.py tests/test_gpu_shards.py -q -m gpu --timeout 300 -x -k "not tiny_spans" 2>&1 | tail -2 >> gpurun_out/kbench.log
done
if [ -n "$NCU_TW" ]; then
  export FQGPU_LIB=$PWD/seq-collection_b200/variants/libfqgpu_$NCU_TW.so
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:fq_scan_kernel -c 1 -o gpurun_out/r2_${NCU_TW}_full python tools/kbench.py --mb 2048 --reps 1 > gpurun_out/ncu_full.log 2>&1
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:fq_scan_kernel -c 1 -o gpurun_out/r2_${NCU_TW}_core python tools/kbench.py --mb 2048 --reps 1 --core > gpurun_out/ncu_core.log 2>&1
fi
cat gpurun_out/kbench.log
