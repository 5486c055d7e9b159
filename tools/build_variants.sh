#!/bin/bash
# builds libfqgpu variants with different team widths into gpurun_out-free paths: seq-collection_b200/variants/libfqgpu_tw<N>.so
set -e
cd "$(dirname "$0")/.."
mkdir -p seq-collection_b200/variants
for tw in "$@"; do
  d=/tmp/fqv_$tw; rm -rf $d; mkdir -p $d
  for f in fq_scan fq_meta fqgpu_api fq_synth fq_shard fq_index fq_dedup fq_bgzf; do
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-Wall,-Wno-unused-function -DFQ_TW=$tw -c seq-collection_b200/csrc/$f.cu -o $d/$f.o &
  done
  wait
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o seq-collection_b200/variants/libfqgpu_tw$tw.so $d/*.o -lz
  echo built tw=$tw
done
