#!/bin/bash
GZ_HOST=0 FQGPU_GZ_TRACE=1 timeout 600 python tools/gz_time.py 4000000 100 2>&1 | tail -40
