#!/usr/bin/env python
"""One plain FASTQ file over all visible GPUs in ONE process (fqgpu_count_file_sharded, SURVEY 8e file mode) against
the same file through one context.  Writes the synthetic Illumina set (default 24 M records, 8.6 GB) to --dir (page
cache), checks both rows against the HBM-resident scan and prints one JSON line.
Usage: python tools/sharded_e2e.py [--records 24000000] [--dir /tmp]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import seq_collection_b200 as fq

ap = argparse.ArgumentParser()
ap.add_argument("--records", type=int, default=24_000_000)
ap.add_argument("--dir", default="/tmp")
a = ap.parse_args()
n = 360 * a.records
ndev = torch.cuda.device_count()
buf = torch.empty(n, dtype=torch.uint8, device="cuda")
ctx = fq.FqGpu(meta_records=100)
ctx.synth_illumina(buf.data_ptr(), n, 0, a.records, 20240229)
ref = ctx.count_device(buf.data_ptr(), n)
path = os.path.join(a.dir, "sharded_e2e.fq")
buf.cpu().numpy().tofile(path)
del buf
out = {"config": "synthetic Illumina 2x150 bp, %d records (%.2f GB) as a file in the page cache" % (a.records, n / 1e9),
       "gpus": ndev, "host_cores": os.cpu_count(), "runs": {}}
best = 1e9
for _ in range(3):
    t0 = time.perf_counter()
    st = ctx.count_file(path)
    best = min(best, time.perf_counter() - t0)
assert st.to_dict() == ref.to_dict()
out["runs"]["one_context"] = {"seconds": round(best, 3), "GBps": round(n / best / 1e9, 2)}
for world in sorted({ndev, 2 * ndev}):
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        st = fq.count_file_sharded(path, devices=[g % ndev for g in range(world)], meta_records=100)
        best = min(best, time.perf_counter() - t0)
    assert st.to_dict() == ref.to_dict(), world
    out["runs"][f"sharded_{world}_shards"] = {"seconds": round(best, 3), "GBps": round(n / best / 1e9, 2),
                                              "note": "includes creating and destroying one context per shard"}
print(json.dumps(out))
os.remove(path)
