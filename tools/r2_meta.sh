#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_meta_fold.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5
GZ_HOST=0 timeout 600 python tools/gz_time.py 4000000 4000000 2>&1 | grep device | tail -1
bash tools/r2_meta_ncu.sh 2>&1 | grep -E "fq_meta|fq_scan"
