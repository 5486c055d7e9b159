#!/usr/bin/env python
"""BGZF-compressed 2x150 bp FASTQ end to end.  Writes the config-5 set (default 4 M records,
1.44 GB raw) as BGZF (64 KiB members, zlib level 6), counts it through fqgpu_count_file with the device inflate
(csrc/fq_bgzf.cu) and with FQGPU_NO_BGZF=1 (host zlib, the reference's gzip_stream path), checks both rows against
the known tallies of the synthetic set and prints one JSON line.
Usage: python tools/bgzf_e2e.py [--records 4000000] [--dir /tmp]"""
import argparse
import json
import multiprocessing as mp
import os
import struct
import sys
import time
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

BLOCK = 65280


def member(chunk: bytes) -> bytes:
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    comp = co.compress(chunk) + co.flush()
    return (struct.pack("<BBBBIBBH", 0x1F, 0x8B, 8, 4, 0, 0, 0xFF, 6) + b"BC" + struct.pack("<HH", 2, len(comp) + 25)
            + comp + struct.pack("<II", zlib.crc32(chunk), len(chunk)))


def main():
    import torch

    import seq_collection_b200 as fq

    ap = argparse.ArgumentParser()
    ap.add_argument("--records", type=int, default=4_000_000)
    ap.add_argument("--dir", default="/tmp")
    a = ap.parse_args()
    n = 360 * a.records
    buf = torch.empty(n, dtype=torch.uint8, device="cuda")
    ctx = fq.FqGpu(meta_records=100)
    ctx.synth_illumina(buf.data_ptr(), n, 0, a.records, 20240229)
    ref = ctx.count_device(buf.data_ptr(), n)  # the same bytes, HBM-resident
    raw = buf.cpu().numpy().tobytes()
    t0 = time.perf_counter()
    with mp.Pool(min(16, os.cpu_count() or 1)) as pool:
        parts = pool.map(member, [raw[i:i + BLOCK] for i in range(0, n, BLOCK)], chunksize=64)
    path = os.path.join(a.dir, "cfg5.bgzf.fq.gz")
    with open(path, "wb") as f:
        for p in parts:
            f.write(p)
        f.write(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))
    out = {"config": "BGZF (64 KiB members, zlib -6) of synthetic Illumina 2x150 bp, %d records (%.2f GB raw)" % (a.records, n / 1e9),
           "gz_bytes": os.path.getsize(path), "members": len(parts) + 1, "compress_seconds_cpu_pool": round(time.perf_counter() - t0, 1),
           "runs": {}}
    row_ref = fq.fq_count_row(ref) + "\t" + "\t".join(map(str, fq.fq_meta_quality_fields(ref)))
    for name, env in (("device_inflate", None), ("host_zlib", "1")):
        if env:
            os.environ["FQGPU_NO_BGZF"] = env
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            st = ctx.count_file(path)
            best = min(best, time.perf_counter() - t0)
        row = fq.fq_count_row(st) + "\t" + "\t".join(map(str, fq.fq_meta_quality_fields(st)))
        assert row == row_ref and st.to_dict() == ref.to_dict(), name
        out["runs"][name] = {"seconds": round(best, 3), "raw_GBps": round(n / best / 1e9, 3), "members_on_device": ctx.bgzf_members()}
        os.environ.pop("FQGPU_NO_BGZF", None)
    out["row"] = row_ref
    print(json.dumps(out))
    os.remove(path)


if __name__ == "__main__":
    main()
