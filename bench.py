#!/usr/bin/env python
"""bench.py -- FASTQ scanning throughput (GB/s, reads/s) of the fq-count / fq-meta hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--records R] [--workload illumina|ont]

One "step" = one pass of the hot path (reset, scan, counter reduction, 20 KB result read-back)
over the whole synthetic input.  N=1 workload = BASELINE.json configs[1]: synthetic Illumina
2x150 bp uncompressed FASTQ, 100 M reads (36.0 GB), phred+33, resident in HBM.  N>1 = configs[2]:
the same bytes sharded by byte range over the ranks (generated in place), one peer-memory all-gather + device combine of the
counter blocks per step (strong scaling).  The input is far larger than the 126 MB L2, so no flush
is needed between steps.

`value`   : whole-job GB/s, device-timed (CUDA events on the library's stream), max over ranks.
`e2e`     : the same metric through the C ABI with HOST buffers (pinned), H2D inside the timed region.
`roofline`: algorithmic bytes (1 per input byte) / device time, against the measured HBM
            peak of MEASURED_PEAKS.json.
`stats_digest`: sha256 of the fqgpu_stats struct the timed steps produced.  The N-GPU runs scan the same 36.0 GB stream, so
            the digests of N = 1, 2, 4, 8 must be identical (BASELINE configs[2]: "must equal config 2's bit-for-bit"); on
            top of that every run asserts the FULL struct against the generator's own tallies (fqgpu_synth_illumina_tally:
            derived from the random numbers, no byte is read; pinned by the CPU twin oracle/fq_synth_twin.c).
`ont`, `gz` : BASELINE configs[3] and [4] next to the headline (N = 1 only).
`cpu_baseline` / `--impl reference`: the reference's own work shape (oracle/fq_oracle.c
            fqo_ref_fq_count_mem: line reader + three count passes, src/fq_count.nim:38-45) on the host
            cores of this box.  The Nim reference itself cannot be built here (no Nim toolchain), and it is
            single-threaded for this command, so cores = 1.
"""
from __future__ import annotations

import argparse
import glob
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

REC_BYTES = 360
SEED_ILLUMINA = 20240229
SEED_ONT = 20240301
METRIC = "fastq_scan_throughput"


def workload_config(args, world: int) -> dict:
    """The `config` object of the JSON line: identical for the repo arm and the reference arm."""
    total_bytes = args.records * REC_BYTES if args.workload == "illumina" else int(args.ont_gb * 1e9)
    return {"workload": args.workload_name, "bytes": total_bytes, "record_bytes": REC_BYTES,
            "seed": SEED_ILLUMINA if args.workload == "illumina" else SEED_ONT,
            "sharding": (f"byte ranges over {world} rank(s); one launch per rank all-gathers the counter blocks through peer memory "
                         "(NVLink P2P stores) and combines them on the device") if world > 1 else "none",
            "l2": "input >> 126 MB L2, no flush needed", "meta_records": args.meta_records,
            "stats": "full (fq-count + A/C/G/T/N, length tables, quality histogram, per-position sums, fq-meta range)"}


def stats_digest(st) -> str:
    import hashlib

    return hashlib.sha256(bytes(st)).hexdigest()


def twin_sample(nbytes: int):
    """`nbytes` of the synthetic stream from the CPU twin of the generator (oracle/fq_synth_twin.c), in parallel slices."""
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor
    from oracle import fq_oracle as O
    import ctypes as C

    out = np.empty(nbytes, dtype=np.uint8)
    L = O.lib()
    nthr = max(1, min(16, os.cpu_count() or 1))
    step = -(-nbytes // nthr)
    step += -step % REC_BYTES

    def fill(i):
        lo = i * step
        n = max(0, min(step, nbytes - lo))
        if n:
            L.fqo_synth_illumina_bytes(C.c_void_p(out.ctypes.data + lo), lo, n, SEED_ILLUMINA)

    with ThreadPoolExecutor(max_workers=nthr) as pool:
        list(pool.map(fill, range(nthr)))
    return out


def measured_traffic_ratio():
    """DRAM bytes (read + write) per input byte of fq_scan_kernel, from the committed `ncu --set full` capture."""
    best = None
    for p in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json"))):
        try:
            best = (float(json.load(open(p))["dram_bytes_per_input_byte"]), os.path.relpath(p, ROOT))
        except Exception:
            pass
    return best


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_leg(sample, seconds_budget: float = 20.0):
    """Times the reference work shape on `sample` (numpy uint8); returns (GB/s, reads/s, used_bytes, row)."""
    from oracle import fq_oracle as O

    n = sample.size
    probe = min(n, 256 << 20)
    probe -= probe % REC_BYTES
    t0 = time.perf_counter()
    O.ref_fq_count_mem(sample[:probe])
    dt = time.perf_counter() - t0
    rate = probe / dt
    use = int(min(n, max(probe, rate * seconds_budget)))
    use -= use % REC_BYTES
    t0 = time.perf_counter()
    r = O.ref_fq_count_mem(sample[:use])
    dt = time.perf_counter() - t0
    return use / dt / 1e9, r["reads"] / dt, use, r


def run_reference_arm(args):
    """--impl reference: the reference's CPU path (restated; see module docstring) on this box's host cores.  Nothing of
    libfqgpu is loaded: the sample comes from the CPU twin of the generator."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return 0
    # the same synthetic bytes as our arm, bounded sample
    sample_bytes = min(args.records * REC_BYTES, REC_BYTES * ((args.ref_sample_mb << 20) // REC_BYTES))
    sample_bytes -= sample_bytes % REC_BYTES
    host = twin_sample(sample_bytes)

    vals, reads_s = [], []
    per_step_budget = max(2.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    used = 0
    for i in range(args.warmup + args.steps):
        gbs, rps, used, _ = cpu_reference_leg(host, per_step_budget)
        if i >= args.warmup:
            vals.append(gbs); reads_s.append(rps)
    v = statistics.mean(vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "GB/s", "reads_per_s": statistics.mean(reads_s),
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": used / v / 1e6,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(args, max(world, args.gpus)),
        "cpu_baseline": {"value": v, "unit": "GB/s", "cores": 1, "kind": "port",
                         "sample": f"{used / 1e9:.2f} GB prefix of the same synthetic stream (CPU twin of the generator), in host RAM; "
                                   "C restatement of src/fq_count.nim:38-45 (Nim toolchain absent; the reference is single-threaded "
                                   "for this command, so one core is all it can use)"},
        "e2e": {"value": v, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def single_member_gzip(raw: bytes, level: int = 6) -> bytes:
    """ONE gzip member holding `raw` (what `gzip -6` writes, compressed pigz-style by a thread pool: independent raw-deflate
    pieces joined at sync-flush boundaries are one valid DEFLATE stream)."""
    import struct
    import zlib
    from concurrent.futures import ThreadPoolExecutor

    piece = 8 << 20
    chunks = [raw[i:i + piece] for i in range(0, len(raw), piece)] or [b""]

    def comp(args):
        i, c = args
        co = zlib.compressobj(level, zlib.DEFLATED, -15)
        return co.compress(c) + co.flush(zlib.Z_FINISH if i == len(chunks) - 1 else zlib.Z_SYNC_FLUSH)

    with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as pool:
        parts = list(pool.map(comp, enumerate(chunks)))
    return (struct.pack("<BBBBIBB", 0x1F, 0x8B, 8, 0, 0, 0, 0xFF) + b"".join(parts)
            + struct.pack("<II", zlib.crc32(raw) & 0xFFFFFFFF, len(raw) & 0xFFFFFFFF))


def gz_leg(fq, ctx, n_records: int):
    """BASELINE configs[4]: gzip-compressed 2x150 bp FASTQ end to end -- a single-member .fq.gz (what `gzip -6` writes)
    through fqgpu_count_file: file -> pinned memory -> HBM -> inflated ON THE DEVICE (csrc/fq_gzip.cu: chunks of the one
    DEFLATE stream in parallel, CRC-32 and ISIZE checked) -> scanned; fq-count row + fq-meta quality range with
    -n <all records>; reported in GB/s of UNCOMPRESSED bytes, wall clock.  The same file through the reference's path
    (gzread on one host thread, FQGPU_NO_GZIP_DEVICE=1) is timed beside it.  Both results are checked against the
    generator's tallies."""
    import shutil
    import tempfile

    import torch

    tmp = tempfile.mkdtemp(prefix="fqgpu_gz_")
    try:
        n = n_records * REC_BYTES
        dev = torch.empty(n + 256, dtype=torch.uint8, device="cuda")
        ctx.synth_illumina(dev.data_ptr(), n, 0, n_records, SEED_ILLUMINA)
        raw = dev[:n].cpu().numpy().tobytes()
        del dev
        path = os.path.join(tmp, "reads.fq.gz")
        t0 = time.perf_counter()
        blob = single_member_gzip(raw)
        with open(path, "wb") as f:
            f.write(blob)
        t_comp = time.perf_counter() - t0
        with fq.FqGpu(meta_records=100) as g:  # fq-meta's default sample: -n 100 (sc.nim:70)
            want = g.synth_illumina_tally(0, n_records, SEED_ILLUMINA)
            best = 1e9
            for _ in range(3):
                t0 = time.perf_counter()
                st = g.count_file(path)
                best = min(best, time.perf_counter() - t0)
            assert bytes(st) == bytes(want), "gz end to end: result differs from the generator's tallies"
            chunks, false_starts, members = g.gzip_chunks(), g.gzip_false_starts(), g.bgzf_members()
            assert chunks > 0, "gz end to end: the device inflate path was not taken"
            os.environ["FQGPU_NO_GZIP_DEVICE"] = "1"
            try:
                t0 = time.perf_counter()
                st_host = g.count_file(path)
                t_host = time.perf_counter() - t0
                assert g.gzip_chunks() == 0
            finally:
                os.environ.pop("FQGPU_NO_GZIP_DEVICE", None)
            assert bytes(st_host) == bytes(want), "gz end to end (host zlib): result differs from the generator's tallies"
        # the quality range over ALL reads (-n <all reads>) on the same file: the fold (src/fq_meta.nim:100-102 is
        # sequential in the line order) runs many CTAs wide where it is a plain min/max, csrc/fq_meta.cu
        with fq.FqGpu(meta_records=n_records) as g:
            want_all = g.synth_illumina_tally(0, n_records, SEED_ILLUMINA)
            t_all = 1e9
            for _ in range(2):
                t0 = time.perf_counter()
                st_all = g.count_file(path)
                t_all = min(t_all, time.perf_counter() - t0)
            assert bytes(st_all) == bytes(want_all), "gz end to end, -n all: result differs from the generator's tallies"
        return {"value": n / best / 1e9, "unit": "GB/s of uncompressed bytes", "records": n_records, "raw_bytes": n, "gz_bytes": len(blob),
                "seconds": best, "fq_count_row": fq.fq_count_row(st), "fq_meta": {"min_qual": st.meta_qual_min, "max_qual": st.meta_qual_max, "n_lines": st.meta_lines // 4},
                "chunks_inflated_on_device": chunks, "false_block_starts_skipped": false_starts, "bgzf_members": members,
                "host_zlib": {"value": n / t_host / 1e9, "seconds": t_host,
                              "note": "the same file with FQGPU_NO_GZIP_DEVICE=1: gzread on ONE host thread into the pinned ring, the reference's gzip_stream path"},
                "fq_meta_all_reads": {"value": n / t_all / 1e9, "seconds": t_all, "min_qual": st_all.meta_qual_min, "max_qual": st_all.meta_qual_max,
                                      "n_lines": st_all.meta_lines // 4,
                                      "note": "the same with -n <all reads>: the quality-range fold over every record (segments folded in parallel, csrc/fq_meta.cu), best of 2"},
                "note": "single-member gzip (level 6), inflated on the device: block starts guessed per chunk and proven by the chunk before "
                        "landing on them, one decode pass, back-references across chunks carried as markers and resolved through a chain of 32 KiB windows, "
                        "CRC-32 + ISIZE of the member verified; fq-meta sample -n 100 (the reference's default); wall clock from open() to "
                        f"the finished statistics, best of 3 (the file was written in {t_comp:.1f} s)"}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def ont_leg(fq, device: int, gb: float, meta_records: int, steps: int, peak: float):
    """BASELINE configs[3]: synthetic ONT-style long reads (lengths 1-100 kb log-normal), HBM-resident, full statistics."""
    import torch

    cap = int(gb * 1e9)
    buf = torch.empty(cap + (1 << 20), dtype=torch.uint8, device="cuda")
    n_rec = int(cap / 27200)  # mean record ~ 2*13.5 kB + header
    with fq.FqGpu(device=device, meta_records=meta_records) as c:
        nbytes = c.synth_ont(buf.data_ptr(), cap + (1 << 20), 0, n_rec, SEED_ONT)
        for _ in range(2):
            st = c.count_device(buf.data_ptr(), nbytes)
        ms = 0.0
        for _ in range(steps):
            st = c.count_device(buf.data_ptr(), nbytes)
            ms += c.last_timing()[0]
        # invariants of the shape (every line of a record is whole: chunk-edge record carry)
        assert st.reads == n_rec and st.lines == 4 * n_rec and st.bytes == nbytes, (st.reads, st.lines, st.bytes)
        assert st.bases == sum(st.qual_counts) == sum(st.base_counts) and st.seq_lines == st.qual_lines == n_rec
        assert 1000 <= st.seq_len_min <= st.seq_len_max <= 100000 and (st.seq_len_min, st.seq_len_max) == (st.qual_len_min, st.qual_len_max)
        assert sum(st.base_counts[ord(ch)] for ch in "ACGT") == st.bases and st.n_bases == 0
        assert sum(st.qual_pos_sum) == sum(i * v for i, v in enumerate(st.qual_counts))
        # the same stream in two launches cut at an awkward offset (8 MiB + 13) gives the same struct
        c.reset()
        cut = (8 << 20) + 13
        c.scan_device(buf.data_ptr(), cut)
        c.scan_device(buf.data_ptr() + cut, nbytes - cut)
        assert bytes(c.finish()) == bytes(st), "ONT: split scan differs"
        v = nbytes / (ms / steps * 1e6)
        return {"value": v, "unit": "GB/s", "frac_of_hbm_peak": v / peak, "bytes": nbytes, "reads": n_rec, "steps": steps,
                "reads_per_s": n_rec / (ms / steps / 1e3), "stats_digest": stats_digest(st),
                "note": "full statistics, HBM-resident, CUDA events on the library stream; invariants of the shape and a split scan asserted"}


def _bgzf_member(chunk: bytes) -> bytes:
    import struct
    import zlib

    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    comp = co.compress(chunk) + co.flush()
    return (struct.pack("<BBBBIBBH", 0x1F, 0x8B, 8, 4, 0, 0, 0xFF, 6) + b"BC" + struct.pack("<HH", 2, len(comp) + 25)
            + comp + struct.pack("<II", zlib.crc32(chunk), len(chunk)))


def ingest_leg(fq, ctx, buf, nbytes):
    """Wall-clock GB/s of fqgpu_count_file on files holding a prefix of the benchmark stream; every result is
    checked against the HBM-resident scan of the same bytes."""
    import shutil
    import tempfile
    from concurrent.futures import ThreadPoolExecutor

    tmp = tempfile.mkdtemp(prefix="fqgpu_bench_")
    out = {}
    try:
        fctx = fq.FqGpu(meta_records=100)
        # plain file, page cache
        n_plain = min(nbytes, 4 << 30)
        n_plain -= n_plain % REC_BYTES
        want = ctx.count_device(buf.data_ptr(), n_plain)
        path = os.path.join(tmp, "plain.fq")
        buf[:n_plain].cpu().numpy().tofile(path)
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            st = fctx.count_file(path)
            best = min(best, time.perf_counter() - t0)
        assert st.to_dict() == want.to_dict(), "plain file: result differs from the HBM-resident scan"
        out["plain_file"] = {"value": n_plain / best / 1e9, "unit": "GB/s", "bytes": n_plain,
                             "note": "uncompressed .fq in the page cache -> pinned ring (multi-threaded reads) -> H2D -> scan; wall clock, best of 3"}
        os.remove(path)
        # BGZF: 64 KiB members, zlib level 6 (compressed here by a thread pool; zlib releases the GIL)
        n_gz = min(nbytes, 1440 * 1_000_000)
        n_gz -= n_gz % REC_BYTES
        want = ctx.count_device(buf.data_ptr(), n_gz)
        raw = buf[:n_gz].cpu().numpy().tobytes()
        with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as pool:
            parts = list(pool.map(_bgzf_member, [raw[i:i + 65280] for i in range(0, n_gz, 65280)]))
        path = os.path.join(tmp, "bgzf.fq.gz")
        with open(path, "wb") as f:
            for p in parts:
                f.write(p)
            f.write(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))
        for name, env in (("bgzf_device_inflate", None), ("bgzf_host_zlib", "1")):
            if env:  # neither device inflater: gzread on the host, the reference's path
                os.environ["FQGPU_NO_BGZF"] = env
                os.environ["FQGPU_NO_GZIP_DEVICE"] = env
            try:
                best = 1e9
                for _ in range(1 if env else 3):
                    t0 = time.perf_counter()
                    st = fctx.count_file(path)
                    best = min(best, time.perf_counter() - t0)
            finally:
                os.environ.pop("FQGPU_NO_BGZF", None)
                os.environ.pop("FQGPU_NO_GZIP_DEVICE", None)
            assert st.to_dict() == want.to_dict(), name + ": result differs from the HBM-resident scan"
            out[name] = {"value": n_gz / best / 1e9, "unit": "GB/s of uncompressed bytes", "bytes": n_gz,
                         "members_on_device": fctx.bgzf_members()}
        out["bgzf_host_zlib"]["note"] = "the same file with FQGPU_NO_BGZF=1 FQGPU_NO_GZIP_DEVICE=1: zlib on one host thread, the reference's gzip_stream path"
        fctx.close()
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="fqgpu", choices=["fqgpu", "reference"])
    ap.add_argument("--records", type=int, default=100_000_000, help="total reads of the synthetic Illumina set")
    ap.add_argument("--workload", default="illumina", choices=["illumina", "ont"])
    ap.add_argument("--ont-gb", type=float, default=20.0)
    ap.add_argument("--meta-records", type=int, default=100, help="fq-meta sample_n folded in the same pass (CLI default 100)")
    ap.add_argument("--e2e-mb", type=int, default=8192, help="host-resident sample for the end-to-end leg")
    ap.add_argument("--ref-sample-mb", type=int, default=4096)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-ingest", action="store_true", help="skip the file-ingest figures (plain file and BGZF through fqgpu_count_file)")
    ap.add_argument("--no-ont", action="store_true", help="skip the ONT figure (BASELINE configs[3]) of the N=1 line")
    ap.add_argument("--no-gz", action="store_true", help="skip the gzip end-to-end figure (BASELINE configs[4]) of the N=1 line")
    ap.add_argument("--gz-records", type=int, default=4_000_000, help="records of the gzip end-to-end file (SURVEY 8d config 5: 4 M = 1.44 GB raw)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "fqgpu" else args.warmup
    args.workload_name = ("synthetic Illumina 2x150 bp uncompressed FASTQ, %d reads, phred+33" % args.records
                          if args.workload == "illumina" else "synthetic ONT-style long reads, log-normal 1-100 kb, %.0f GB" % args.ont_gb)
    if args.impl == "reference":
        return run_reference_arm(args)

    import numpy as np
    import torch

    import seq_collection_b200 as fq

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    if args.workload == "ont":  # BASELINE configs[3] as the headline of this run (single GPU)
        assert world == 1, "the ONT workload is a single-GPU config (BASELINE.json configs[3])"
        peak, peak_src = measured_hbm_peak()
        sampler = ClockSampler(local_rank)
        sampler.start()
        o = ont_leg(fq, local_rank, args.ont_gb, args.meta_records, args.steps, peak)
        clocks = sampler.stop()
        ms = o["bytes"] / o["value"] / 1e6
        cfg = workload_config(args, 1)
        cfg["bytes"] = o["bytes"]
        print(json.dumps({"metric": METRIC, "value": o["value"], "unit": "GB/s", "reads_per_s": o["reads_per_s"], "n_gpus": 1, "steps": args.steps,
                          "warmup": 2, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
                          "data": "synthetic", "config": cfg, "stats_digest": o["stats_digest"], "clocks": clocks,
                          "roofline": {"bound": "hbm", "achieved": o["value"], "peak": peak, "unit": "GB/s", "frac": o["value"] / peak,
                                       "traffic": None, "peak_source": peak_src}, "gpu_launches": int(args.steps * 4),
                          "e2e": None, "cpu_baseline": None, "note": o["note"]}))
        return 0

    ctx = fq.FqGpu(device=local_rank, meta_records=args.meta_records)

    # ---- synthetic input, generated in place in HBM (rank r holds its byte range of ONE stream) ----
    if args.workload == "illumina":
        total_bytes = args.records * REC_BYTES
        lo = total_bytes * rank // world
        hi = total_bytes * (rank + 1) // world
        lo -= lo % 16 if rank else 0  # keep shard starts 16-byte aligned (not record aligned)
        hi -= hi % 16 if rank + 1 < world else 0
        nbytes = hi - lo
        buf = torch.empty(nbytes + 256, dtype=torch.uint8, device="cuda")
        ctx.synth_illumina_bytes(buf.data_ptr(), lo, nbytes, SEED_ILLUMINA)
        total_reads = args.records
    else:
        assert world == 1, "the ONT workload is a single-GPU config (BASELINE.json configs[3])"
        cap = int(args.ont_gb * 1e9)
        buf = torch.empty(cap + (1 << 20), dtype=torch.uint8, device="cuda")
        n_rec = int(cap / 27200)  # mean record ~ 2*13.5 kB + header
        nbytes = ctx.synth_ont(buf.data_ptr(), cap + (1 << 20), 0, n_rec, SEED_ONT)
        total_bytes = nbytes
        lo = 0
        total_reads = n_rec
    torch.cuda.synchronize()

    def open_exchange(c):
        """The library's own collective (fqgpu_shard_exchange_*): every rank's exchange buffer is opened by the others through
        CUDA IPC; the 64-byte handles are all-gathered once, outside the timed region."""
        h = c.shard_exchange_create(rank, world)
        mine = torch.frombuffer(bytearray(h), dtype=torch.uint8).cuda()
        allh = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allh, mine)
        c.shard_exchange_open(b"".join(bytes(t.cpu().numpy()) for t in allh))

    if world > 1:
        open_exchange(ctx)

    lib_stream = torch.cuda.ExternalStream(ctx.stream)

    def sharded_step(c, scan):
        """shard_begin -> scan -> ONE launch that all-gathers the blocks through peer memory (NVLink P2P stores) and
        combines them on the device -> one read-back.  A wrong phase hypothesis (malformed input) repeats the exchange
        after the first wrong rank has scanned again with the exact carry."""
        c.shard_begin(rank, world)
        scan()
        while True:
            c.shard_exchange_start()
            rc, st_ = c.shard_exchange_finish()
            if rc == 0:
                return st_
            if c.shard_rescan(c.shard_gathered()) == fq.ERETRY:
                scan()

    def step():
        """One pass of the hot path over this rank's bytes; returns Stats."""
        if world == 1:
            return ctx.count_device(buf.data_ptr(), nbytes)
        return sharded_step(ctx, lambda: ctx.scan_device(buf.data_ptr(), nbytes))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        st = step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    dev_ms = 0.0
    launches = 0
    t0 = time.perf_counter()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(lib_stream)
    for _ in range(args.steps):
        st = step()
        ms_i, n_i = ctx.last_timing()  # device time of this step's kernels (events on the library stream)
        dev_ms += ms_i
        launches += n_i
    ev1.record(lib_stream)
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    span_ms = ev0.elapsed_time(ev1)  # device time of the whole timed region on the library's stream
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([span_ms, dev_ms, wall_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    span_ms, dev_ms, wall_ms = [float(x) for x in t.cpu()]
    ms_per_step = span_ms / args.steps
    value = total_bytes / (ms_per_step * 1e6)

    # ---- the result: the FULL fqgpu_stats of the whole stream against the generator's own tallies (derived from the
    # random numbers, no byte read; oracle/fq_synth_twin.c pins the tally kernel) -- at every N, so that the digests of
    # N = 1, 2, 4, 8 are provably the digest of the same struct ----
    digest = stats_digest(st)
    tally = ctx.synth_illumina_tally(0, args.records, SEED_ILLUMINA)
    if bytes(st) != bytes(tally):
        got, want = st.to_dict(), tally.to_dict()
        bad = [k for k in want if got[k] != want[k]]
        raise AssertionError(f"scan result differs from the generator's tallies in {bad[:8]}")

    # ---- the same bytes in core-only mode (FQGPU_F_CORE_ONLY: exactly what `sc fq-count` prints) ----
    # An extra figure next to the headline (which stays the full statistics set); N=1 only.
    core = None
    if world == 1 and args.workload == "illumina":
        cctx = fq.FqGpu(device=local_rank, meta_records=0, flags=fq.F_CORE_ONLY)
        for _ in range(2):
            cst = cctx.count_device(buf.data_ptr(), nbytes)
        cms = 0.0
        ncore = max(3, min(args.steps, 5))
        for _ in range(ncore):
            cst = cctx.count_device(buf.data_ptr(), nbytes)
            cms += cctx.last_timing()[0]
        assert (cst.reads, cst.bases, cst.gc_bases, cst.n_bases) == (st.reads, st.bases, st.gc_bases, st.n_bases)
        assert bytes(cst) == bytes(cctx.synth_illumina_tally(0, args.records, SEED_ILLUMINA)), "core-only result differs from the generator's tallies"
        core = {"value": nbytes / (cms / ncore * 1e6), "unit": "GB/s", "steps": ncore, "frac_of_hbm_peak": nbytes / (cms / ncore * 1e6) / measured_hbm_peak()[0],
                "stats_digest": stats_digest(cst), "fq_count_row": fq.fq_count_row(cst),
                "stats": "reads, bases, G/C/N and sequence-length tables only (quality lines are not examined)"}
        cctx.close()

    # ---- end-to-end: host (pinned) buffers through the C ABI, H2D inside the timed region ----
    # A bounded sample of the same stream (first e2e_mb MiB, byte-range sharded over the ranks like the
    # resident run) sits in pinned host memory; every step copies it to the device in 64 MiB chunks
    # (copy stream overlapping the scan stream), scans, exchanges and reads the ~20 KB result back.
    e2e = None
    if not args.no_e2e and args.workload == "illumina":
        total_sample = min(total_bytes, args.e2e_mb << 20)
        total_sample -= total_sample % REC_BYTES
        slo = total_sample * rank // world
        shi = total_sample * (rank + 1) // world
        sample = shi - slo
        dev = torch.empty(sample + 256, dtype=torch.uint8, device="cuda")
        ctx.synth_illumina_bytes(dev.data_ptr(), slo, sample, SEED_ILLUMINA)
        host = torch.empty(sample, dtype=torch.uint8, pin_memory=True)
        host.copy_(dev[:sample])
        torch.cuda.synchronize()
        # the ceiling of this leg: the pinned host -> device copy rate of this box, measured here (one rank alone would see
        # the PCIe link; with every rank copying at once it is the host memory / root-complex share of each)
        h2d_n = min(sample, 2 << 30)
        barrier()
        h2d_best = 0.0
        for _ in range(3):
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            dev[:h2d_n].copy_(host[:h2d_n], non_blocking=True)
            c1.record()
            torch.cuda.synchronize()
            h2d_best = max(h2d_best, h2d_n / (c0.elapsed_time(c1) * 1e6))
        th = torch.tensor([h2d_best], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(th, op=dist.ReduceOp.SUM)
        h2d_total = float(th.cpu()[0])
        ectx = fq.FqGpu(device=local_rank, meta_records=args.meta_records)
        e_lib_stream = torch.cuda.ExternalStream(ectx.stream)

        if world > 1:
            open_exchange(ectx)

        def e2e_step():
            if world == 1:
                return ectx.count_host_ptr(host.data_ptr(), sample)  # returns after the result is on the host
            return sharded_step(ectx, lambda: ectx.scan_host_ptr(host.data_ptr(), sample))

        for _ in range(2):
            est = e2e_step()
        barrier()
        n_e2e = max(3, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            est = e2e_step()
        barrier()
        e2e_s = (time.perf_counter() - t0) / n_e2e
        te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_s = float(te.cpu()[0])
        assert est.reads == total_sample // REC_BYTES and est.bases == 150 * est.reads, "end-to-end result is wrong"
        if world == 1:
            want = ctx.count_device(buf.data_ptr(), sample)
            assert est.to_dict() == want.to_dict(), "end-to-end result differs from the HBM-resident result"
        e2e = {"value": total_sample / e2e_s / 1e9, "unit": "GB/s", "h2d_bytes_per_step": sample,
               "d2h_bytes_per_step": 8 * 2144 + 80 + 40,
               "reads_per_s": est.reads / e2e_s,
               "roofline": {"bound": "pinned host->device copy", "peak": h2d_total, "unit": "GB/s", "frac": total_sample / e2e_s / 1e9 / h2d_total,
                            "how": f"cudaMemcpyAsync of {h2d_n / 1e9:.2f} GB pinned -> device per rank, CUDA events, best of 3, all {world} rank(s) copying at once; sum over ranks"},
               "sample": f"{total_sample / 1e9:.2f} GB prefix of the stream in pinned host memory ({sample / 1e9:.2f} GB per rank), "
                         "streamed in 64 MiB chunks (copy stream overlaps the scan stream); wall clock, max over ranks"}
        ectx.close()
        del host, dev

    # ---- CPU baseline beside it (bounded sample, rank 0, N=1 only) ----
    cpu = None
    if not args.no_cpu_baseline and world == 1 and rank == 0 and args.workload == "illumina":
        sample = min(nbytes, 2048 << 20)
        sample -= sample % REC_BYTES
        hostnp = buf[:sample].cpu().numpy()
        gbs, rps, used, r = cpu_reference_leg(hostnp, 12.0)
        chk = ctx.count_device(buf.data_ptr(), used)
        assert (r["reads"], r["gc_bases"], r["n_bases"], r["bases"]) == (chk.reads, chk.gc_bases, chk.n_bases, chk.bases)
        cpu = {"value": gbs, "unit": "GB/s", "cores": 1, "kind": "port", "reads_per_s": rps,
               "sample": f"{used / 1e9:.2f} GB prefix of the same stream; C restatement of the reference loop "
                         "(oracle/fq_oracle.c fqo_ref_fq_count_mem); result equals the GPU's on that prefix"}

    # ---- file ingest beside it (N=1 only): the same bytes as FILES through fqgpu_count_file ----
    # A plain .fq from the page cache (multi-threaded reads into the pinned ring) and a BGZF .fq.gz (members inflated
    # on the device, csrc/fq_bgzf.cu) against the same file through host zlib (the reference's gzip_stream path).
    ingest = None
    if not args.no_ingest and world == 1 and rank == 0 and args.workload == "illumina":
        try:
            ingest = ingest_leg(fq, ctx, buf, nbytes)
        except Exception as e:  # never lets the extra figures break the contract line
            ingest = {"error": repr(e)[:200]}

    # ---- the record-offset index beside it (N=1 only): fqgpu_index_device over the same resident bytes ----
    index = None
    if world == 1 and rank == 0 and args.workload == "illumina":
        try:
            n_rec = nbytes // REC_BYTES
            offs = torch.empty(n_rec, dtype=torch.int64, device="cuda")
            best = 1e30
            for _ in range(5):
                ctx.reset()
                got = ctx.index_device(buf.data_ptr(), nbytes, offs.data_ptr(), n_rec)
                best = min(best, ctx.last_timing()[0])
            ok = got == n_rec and bool((offs == torch.arange(n_rec, device=offs.device, dtype=torch.int64) * REC_BYTES).all())
            assert ok, "index differs from the known record starts"
            peak_i = measured_hbm_peak()[0]
            index = {"value": nbytes / best / 1e6, "unit": "GB/s", "ms": best, "records": int(got), "launches": 1,
                     "frac_of_hbm_peak": nbytes / best / 1e6 / peak_i,
                     "note": "fqgpu_index_device (one launch, chained tile prefix; 1 B read per input byte + 8 B written per record), "
                             "CUDA events on the library stream, best of 5; every offset checked (= 360 k)"}
            del offs
        except Exception as e:  # never lets the extra figures break the contract line
            index = {"error": repr(e)[:200]}

    ont = gz = None
    if world == 1 and rank == 0 and args.workload == "illumina":
        del buf
        torch.cuda.empty_cache()
        if not args.no_ont:
            try:
                ont = ont_leg(fq, local_rank, args.ont_gb, args.meta_records, max(3, min(args.steps, 5)), measured_hbm_peak()[0])
            except Exception as e:  # never lets the extra figures break the contract line
                ont = {"error": repr(e)[:300]}
        if not args.no_gz:
            try:
                gz = gz_leg(fq, ctx, args.gz_records)
            except Exception as e:
                gz = {"error": repr(e)[:300]}

    if rank == 0:
        peak, peak_src = measured_hbm_peak()
        tr = measured_traffic_ratio()
        kern_ms_per_step = dev_ms / args.steps
        achieved = (nbytes / 1e9) / (kern_ms_per_step / 1e3)  # this rank's bytes / its device time
        scan_launches_per_step = launches / args.steps
        line = {
            "metric": METRIC, "value": value, "unit": "GB/s", "reads_per_s": total_reads / (ms_per_step / 1e3),
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "wall_ms_per_step": wall_ms / args.steps,
            "higher_is_better": True, "scaling": "strong" if args.workload == "illumina" else "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(args, world),
            "stats_digest": digest,
            "result_check": "the full fqgpu_stats struct equals the generator's tallies (fqgpu_synth_illumina_tally) for the whole stream",
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         # the dominant kernel (pass 0 of fq_scan_kernel) covers this rank's bytes in ONE launch
                         "traffic": (tr[0] * nbytes) if tr else None,
                         "traffic_unit": "bytes per launch", "traffic_source": tr[1] if tr else None,
                         "algorithmic_bytes_per_launch": nbytes,
                         "peak_source": peak_src,
                         "note": "algorithmic bytes = input bytes (1 B read per byte); device time = CUDA events around "
                                 "the two scan launches (every span; the spans whose guessed phase was wrong -- none here) on the library "
                                 "stream, the fq-meta prefix kernel beside them"},
            "clocks": clocks,
            # per step: reset + fq-meta prefix + the two scan launches (+ shard begin and the exchange kernel of a shard)
            "gpu_launches": int(args.steps * (2 + (1 if args.meta_records and rank == 0 else 0) + 1 + (2 if world > 1 else 0))),
            "e2e": e2e,
            "cpu_baseline": cpu,
            "core_only": core,
            "ingest": ingest,
            "index": index,
            "ont": ont,
            "gz": gz,
        }
        print(json.dumps(line))
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
