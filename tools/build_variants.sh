#!/bin/bash
# builds libfqgpu tuning variants: tools/build_variants.sh name:"-DFLAGS" ...  -> seq-collection_b200/variants/libfqgpu_<name>.so
set -e
cd "$(dirname "$0")/.."
mkdir -p seq-collection_b200/variants
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  d=/tmp/fqv_$name; rm -rf $d; mkdir -p $d
  for f in fq_scan fq_meta fqgpu_api fq_synth fq_shard fq_index fq_dedup fq_bgzf fq_gzip; do
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-Wall,-Wno-unused-function $flags -c seq-collection_b200/csrc/$f.cu -o $d/$f.o &
  done
  wait
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o seq-collection_b200/variants/libfqgpu_$name.so $d/*.o -lz -ldl
  echo built $name
done
