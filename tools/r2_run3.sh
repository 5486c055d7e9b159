#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 120 python -c "
import seq_collection_b200 as fq
from oracle import fq_oracle as O
d = open('tests/golden/fastq/illumina_3.fq','rb').read()
with fq.FqGpu(meta_records=100) as c:
    st = c.count_bytes(d)
    print('smoke', st.reads, st.bases, st.gc_bases, O.count(d,100)['bases'])
" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/smoke.log
rm -f gpurun_out/kbench.log
for args in "" "--core" "--workload ont" "--workload ont --core"; do
  timeout 180 python tools/kbench.py --mb 8192 --reps 5 $args >> gpurun_out/kbench.log 2>&1
done
timeout 1200 python -m pytest tests/test_gpu_scan_tiles.py tests/test_gpu_parity.py tests/test_gpu_shards.py -q -m gpu --timeout 600 -x 2>&1 | tail -60 > gpurun_out/t1.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fq_scan_kernel -c 1 -o gpurun_out/r2_span3_full python tools/kbench.py --mb 2048 --reps 1 > gpurun_out/ncu_full.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fq_scan_kernel -c 1 -o gpurun_out/r2_span3_core python tools/kbench.py --mb 2048 --reps 1 --core > gpurun_out/ncu_core.log 2>&1
cat gpurun_out/smoke.log gpurun_out/kbench.log; tail -25 gpurun_out/t1.log
