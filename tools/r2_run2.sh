#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
export FQGPU_DBG=1
for args in "" "--core"; do
  timeout 180 python tools/kbench.py --mb 8192 --reps 5 $args >> gpurun_out/kbench2.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fq_scan_kernel -c 1 -o gpurun_out/r2_nolb_full python tools/kbench.py --mb 2048 --reps 1 > gpurun_out/ncu_full.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fq_scan_kernel -c 1 -o gpurun_out/r2_nolb_core python tools/kbench.py --mb 2048 --reps 1 --core > gpurun_out/ncu_core.log 2>&1
cat gpurun_out/kbench2.log
