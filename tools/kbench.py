#!/usr/bin/env python
"""Small kernel-only timing harness used while tuning (not the contract bench): scans an
HBM-resident synthetic buffer `reps` times and prints GB/s.  Usage:
    python tools/kbench.py [--mb 4096] [--reps 5] [--workload illumina|ont] [--meta 100]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import seq_collection_b200 as fq

ap = argparse.ArgumentParser()
ap.add_argument("--mb", type=int, default=4096)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--workload", default="illumina")
ap.add_argument("--meta", type=int, default=100)
ap.add_argument("--core", action="store_true", help="FQGPU_F_CORE_ONLY")
ap.add_argument("--index", action="store_true", help="time fqgpu_index_device instead of the scan")
a = ap.parse_args()
n = (a.mb << 20)
buf = torch.empty(n + 4096, dtype=torch.uint8, device="cuda")
ctx = fq.FqGpu(meta_records=a.meta, flags=fq.F_CORE_ONLY if a.core else 0)
if a.workload == "illumina":
    n -= n % 360
    ctx.synth_illumina(buf.data_ptr(), n, 0, n // 360, 20240229)
else:
    n = ctx.synth_ont(buf.data_ptr(), n, 0, n // 27300, 20240301)
if a.index:
    cap = n // 100 + 16
    offs = torch.empty(cap, dtype=torch.int64, device="cuda")
    best = 1e9
    for r in range(a.reps):
        ctx.reset()
        nrec = ctx.index_device(buf.data_ptr(), n, offs.data_ptr(), cap)
        ms, launches = ctx.last_timing()
        best = min(best, ms)
    print(f"index {a.workload} {n/1e9:.2f} GB: best {best:.3f} ms  {n/best/1e6:.1f} GB/s  records {nrec}")
    sys.exit(0)
best = 1e9
for r in range(a.reps):
    st = ctx.count_device(buf.data_ptr(), n)
    ms, launches = ctx.last_timing()
    best = min(best, ms)
print(f"{a.workload} {n/1e9:.2f} GB: best {best:.3f} ms  {n/best/1e6:.1f} GB/s  launches/step {launches}  reads {st.reads} bases {st.bases}")
