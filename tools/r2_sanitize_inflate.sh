#!/bin/bash
# compute-sanitizer over the kernels added late in round 2: gzip / BGZF inflate, the parallel fq-meta fold
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_gzip.py tests/test_gpu_bgzf.py tests/test_gpu_meta_fold.py -q -m gpu -x --timeout 1400 -k "not large_file and not edge_corpus_every and not cuts_everywhere and (chunk_kb or batches or flush or members or broken or equals_oracle or sample_sizes or must_not or long_reads or repetitive)" > gpurun_out/r2_sanitizer_memcheck_inflate.log 2>&1
echo "exit code $?" >> gpurun_out/r2_sanitizer_memcheck_inflate.log
tail -5 gpurun_out/r2_sanitizer_memcheck_inflate.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_gzip.py tests/test_gpu_meta_fold.py -q -m gpu -x --timeout 1100 -k "batches or (equals_oracle and 6-4) or long_reads or sample_sizes" > gpurun_out/r2_sanitizer_racecheck_inflate.log 2>&1
echo "exit code $?" >> gpurun_out/r2_sanitizer_racecheck_inflate.log
tail -5 gpurun_out/r2_sanitizer_racecheck_inflate.log
