#!/bin/bash
# round-2 evidence on one GPU: every GPU test, the two bench arms, ncu launch list + full captures, sanitizer runs
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 2>&1 | tail -15 > gpurun_out/r2_tests_gpu.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/bench_ref.err
timeout 900 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-ont --no-gz --no-ingest --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fq_scan_kernel -c 1 -o gpurun_out/r2_final_full python tools/kbench.py --mb 2048 --reps 1 > gpurun_out/ncu_full.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fq_scan_kernel -c 1 -o gpurun_out/r2_final_core python tools/kbench.py --mb 2048 --reps 1 --core > gpurun_out/ncu_core.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:fq_scan_kernel -c 1 -o gpurun_out/r2_final_ont python tools/kbench.py --mb 2048 --reps 1 --workload ont > gpurun_out/ncu_ont.log 2>&1
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_scan_tiles.py tests/test_gpu_shards.py tests/test_gpu_dedup.py -q -m gpu -x --timeout 500 -k "not tiny_spans and not many_epochs and (seed0 or seed1 or seed2 or seed3 or read_len or crlf or dense or world or dedup or odd_sizes)" > gpurun_out/r2_sanitizer_memcheck.log 2>&1
echo "exit code $?" >> gpurun_out/r2_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_scan_tiles.py -q -m gpu -x --timeout 800 -k "not tiny_spans and not many_epochs and (seed1 or seed5 or read_len36 or read_len151 or dense)" > gpurun_out/r2_sanitizer_racecheck.log 2>&1
echo "exit code $?" >> gpurun_out/r2_sanitizer_racecheck.log
tail -3 gpurun_out/r2_tests_gpu.log; tail -4 gpurun_out/r2_sanitizer_memcheck.log; tail -4 gpurun_out/r2_sanitizer_racecheck.log; cut -c1-400 gpurun_out/r2_bench_n1.json; tail -3 gpurun_out/bench_n1.err
