#!/bin/bash
# one full ncu capture of the gzip decode kernels (count + write) with source
mkdir -p gpurun_out
GZ_HOST=0 timeout 900 ncu --set full --import-source on --clock-control none -k regex:"gz_count_kernel|gz_write_kernel" -c 2 -f -o gpurun_out/gz_decode python tools/gz_time.py 1000000 100 > gpurun_out/gz_ncu_full.log 2>&1
tail -3 gpurun_out/gz_ncu_full.log
ls -la gpurun_out/*.ncu-rep
