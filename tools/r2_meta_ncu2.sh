#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:fq_meta_seg_kernel -s 12 -c 1 -f -o gpurun_out/meta_seg python bench.py --steps 1 --warmup 1 --records 20000000 --meta-records 20000000 --no-ont --no-gz --no-ingest --no-e2e --no-cpu-baseline > gpurun_out/meta_ncu2.log 2>&1
ls -la gpurun_out/meta_seg.ncu-rep
