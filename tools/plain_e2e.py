#!/usr/bin/env python
"""End to end on an uncompressed FASTQ FILE (page cache -> pinned ring -> H2D -> scan): `sc fq-count reads.fq`.
Writes the synthetic Illumina set (default 12 M records, 4.3 GB) to --dir, counts it through fqgpu_count_file with
1 reader thread (the reference's single read loop) and with the default reader threads, checks the row against the
HBM-resident scan of the same bytes and prints one JSON line.
Usage: python tools/plain_e2e.py [--records 12000000] [--dir /tmp]"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import seq_collection_b200 as fq

ap = argparse.ArgumentParser()
ap.add_argument("--records", type=int, default=12_000_000)
ap.add_argument("--dir", default="/tmp")
a = ap.parse_args()
n = 360 * a.records
buf = torch.empty(n, dtype=torch.uint8, device="cuda")
ctx = fq.FqGpu(meta_records=100)
ctx.synth_illumina(buf.data_ptr(), n, 0, a.records, 20240229)
ref = ctx.count_device(buf.data_ptr(), n)
path = os.path.join(a.dir, "plain_e2e.fq")
buf.cpu().numpy().tofile(path)
ctx.close()
out = {"config": "synthetic Illumina 2x150 bp, %d records (%.2f GB) as a file in the page cache" % (a.records, n / 1e9),
       "host_cores": os.cpu_count(), "runs": {}}
code = ("import sys,time,json; sys.path.insert(0,%r); import seq_collection_b200 as fq\n"
        "c=fq.FqGpu(meta_records=100); best=1e9\n"
        "for _ in range(4):\n"
        "    t0=time.perf_counter(); st=c.count_file(%r); best=min(best,time.perf_counter()-t0)\n"
        "print(json.dumps({'seconds':best,'row':fq.fq_count_row(st)}))\n") % (ROOT, path)
for name, env in (("1_reader_thread", {"FQGPU_READ_THREADS": "1"}), ("default_reader_threads", {})):
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env={**os.environ, **env})
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["row"] == fq.fq_count_row(ref), name
    out["runs"][name] = {"seconds": round(d["seconds"], 3), "GBps": round(n / d["seconds"] / 1e9, 2)}
print(json.dumps(out))
os.remove(path)
