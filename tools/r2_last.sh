#!/bin/bash
# last checks of the round: sanitizer over the one-pass gzip path, then the bench line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 240 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_gzip.py -q -m gpu -x --timeout 200 -k "one_pass or batches or (equals_oracle and 6-4) or members" > gpurun_out/r2_sanitizer_memcheck_gzip_one_pass.log 2>&1
echo "exit code $?" >> gpurun_out/r2_sanitizer_memcheck_gzip_one_pass.log
tail -4 gpurun_out/r2_sanitizer_memcheck_gzip_one_pass.log
timeout 300 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/bench_n1.err; tail -2 gpurun_out/bench_n1.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_n1.json').read().strip().splitlines()[-1])
g=d['gz']; print(d['value'], d['core_only']['value'], g['value'], g['seconds'], g['host_zlib']['value'], g['fq_meta_all_reads']['value'], d['ingest']['bgzf_device_inflate']['value'])"
GZ_HOST=0 FQGPU_GZ_TRACE=1 timeout 200 python tools/gz_time.py 4000000 100 > gpurun_out/r2_gzip_trace.txt 2>&1; tail -9 gpurun_out/r2_gzip_trace.txt
