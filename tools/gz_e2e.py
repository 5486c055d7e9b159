#!/usr/bin/env python
"""gzip-compressed 2x150 bp FASTQ end to end (host zlib inflate into pinned chunks,
the call shape of gzip_stream.nim:16-17, then the GPU scan).  Writes the synthetic set (default 4 M records,
1.44 GB raw) with /usr/bin/gzip -6 as one member and as a two-member concatenation, runs `fq-count` +
`fq-meta` through the C ABI (fqgpu_count_file_as) and checks the rows against the oracle reading the same
.gz files.  Prints one JSON line.  Usage: python tools/gz_e2e.py [--records 4000000] [--dir /tmp]"""
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
ap.add_argument("--records", type=int, default=4_000_000)
ap.add_argument("--dir", default="/tmp")
a = ap.parse_args()
n = 360 * a.records
buf = torch.empty(n, dtype=torch.uint8, device="cuda")
ctx = fq.FqGpu(meta_records=100)
ctx.synth_illumina(buf.data_ptr(), n, 0, a.records, 20240229)
torch.cuda.synchronize()
raw = os.path.join(a.dir, "cfg5.fq")
buf.cpu().numpy().tofile(raw)
half = 360 * (a.records // 2)
t0 = time.perf_counter()
subprocess.run(f"gzip -6 -c {raw} > {raw}.gz", shell=True, check=True)
t_gzip = time.perf_counter() - t0
subprocess.run(f"(head -c {half} {raw} | gzip -6 -c; tail -c +{half + 1} {raw} | gzip -6 -c) > {raw}.2m.gz", shell=True, check=True)
out = {"config": "gzip -6 of synthetic Illumina 2x150 bp, %d records (%.2f GB raw)" % (a.records, n / 1e9),
       "gz_bytes": os.path.getsize(raw + ".gz"), "gzip_seconds_cpu": round(t_gzip, 1), "runs": {}}
for name, path in (("plain", raw), ("gz_1member", raw + ".gz"), ("gz_2member", raw + ".2m.gz")):
    best = 1e9
    for _ in range(2):
        t0 = time.perf_counter()
        st = ctx.count_file(path)
        best = min(best, time.perf_counter() - t0)
    row = fq.fq_count_row(st) + "\t" + "\t".join(map(str, fq.fq_meta_quality_fields(st)))
    want = O.count_file(path, 100)
    wrow = O.fq_count_row(want) + "\t" + "\t".join(map(str, O.fq_meta_quality_fields(want)))
    assert row == wrow, (name, row, wrow)
    assert st.to_dict() == want, name
    out["runs"][name] = {"seconds": round(best, 3), "raw_GBps": round(n / best / 1e9, 3), "row": row}
# host inflate alone (zlib, one thread): the ceiling of the .gz runs
t0 = time.perf_counter()
subprocess.run(f"gzip -dc {raw}.gz > /dev/null", shell=True, check=True)
out["gunzip_only_raw_GBps"] = round(n / (time.perf_counter() - t0) / 1e9, 3)
print(json.dumps(out))
for p in (raw, raw + ".gz", raw + ".2m.gz"):
    os.remove(p)
