#!/usr/bin/env python
"""many .fq.gz files in one `sc fq-count` call.  Writes K files of the config-5 shape
(synthetic Illumina 2x150 bp, gzip -6, one member each), counts them through fqgpu_count_files with 1 host
thread (the sequential loop of sc.nim:115-116) and with one thread per file, checks every row against the
oracle reading the same .gz, and prints one JSON line.
Usage: python tools/batch_gz.py [--files 8] [--records 1000000] [--dir /tmp]"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch

import seq_collection_b200 as fq
import fq_oracle as O  # checker only

ap = argparse.ArgumentParser()
ap.add_argument("--files", type=int, default=8)
ap.add_argument("--records", type=int, default=1_000_000)
ap.add_argument("--dir", default="/tmp")
a = ap.parse_args()
n = 360 * a.records
buf = torch.empty(n, dtype=torch.uint8, device="cuda")
ctx = fq.FqGpu(meta_records=0)
paths, procs = [], []
for k in range(a.files):
    ctx.synth_illumina(buf.data_ptr(), n, k * a.records, a.records, 20240229)
    torch.cuda.synchronize()
    raw = os.path.join(a.dir, f"batch{k}.fq")
    buf.cpu().numpy().tofile(raw)
    procs.append(subprocess.Popen(f"gzip -6 -f {raw}", shell=True))
    paths.append(raw + ".gz")
for p in procs:
    assert p.wait() == 0
ctx.close()
out = {"config": "%d files x gzip -6 of synthetic Illumina 2x150 bp, %d records each (%.2f GB raw in total)" % (a.files, a.records, a.files * n / 1e9),
       "host_cores": os.cpu_count(), "runs": {}}
want = [O.fq_count_row(O.count_file(p, 0)) for p in paths]
for threads in sorted({1, 2, 4, a.files}):
    best = 1e9
    for _ in range(2):
        t0 = time.perf_counter()
        rc, res = fq.count_files(paths, n_threads=threads, flags=fq.F_CORE_ONLY)
        best = min(best, time.perf_counter() - t0)
    assert rc == fq.OK
    assert [fq.fq_count_row(st) for _, st in res] == want
    out["runs"][f"{threads}_threads"] = {"seconds": round(best, 3), "raw_GBps": round(a.files * n / best / 1e9, 3)}
print(json.dumps(out))
for p in paths:
    os.remove(p)
