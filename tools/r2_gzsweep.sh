#!/bin/bash
for mb in 96 128 192 256 384; do for kb in 16 24 32; do
  echo "== batch $mb MiB, chunk min $kb KiB"
  FQGPU_GZ_BATCH_MB=$mb FQGPU_GZ_CHUNK_KB=$kb GZ_HOST=0 timeout 600 python tools/gz_time.py 4000000 100 2>&1 | grep device | tail -2
done; done
