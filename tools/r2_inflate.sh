#!/bin/bash
# both device inflaters: tests, end-to-end timing
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_gzip.py tests/test_gpu_bgzf.py -x -q -m gpu --timeout 120 2>&1 | tail -3
GZ_HOST=0 timeout 600 python tools/gz_time.py 4000000 100 2>&1 | grep device
GZ_HOST=0 timeout 600 python tools/gz_time.py 4000000 4000000 2>&1 | grep device
timeout 600 python tools/bgzf_time.py 4000000 2>&1 | grep "bgzf device"
