"""Times fqgpu_count_file on a single-member .fq.gz (device inflate vs host zlib); run on the GPU box."""
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
import seq_collection_b200 as fq  # noqa: E402

n_records = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
meta = int(sys.argv[2]) if len(sys.argv) > 2 else 100
n = n_records * bench.REC_BYTES
with fq.FqGpu(meta_records=meta) as g:
    dev = torch.empty(n + 256, dtype=torch.uint8, device="cuda")
    g.synth_illumina(dev.data_ptr(), n, 0, n_records, bench.SEED_ILLUMINA)
    raw = dev[:n].cpu().numpy().tobytes()
    del dev
    tmp = tempfile.mkdtemp(prefix="gzt_")
    path = os.path.join(tmp, "r.fq.gz")
    t0 = time.perf_counter()
    blob = bench.single_member_gzip(raw)
    open(path, "wb").write(blob)
    print(f"compressed {n/1e9:.2f} GB -> {len(blob)/1e9:.3f} GB in {time.perf_counter()-t0:.1f} s", flush=True)
    want = g.synth_illumina_tally(0, n_records, bench.SEED_ILLUMINA)
    for rep in range(3):
        t0 = time.perf_counter()
        st = g.count_file(path)
        dt = time.perf_counter() - t0
        print(f"device: {dt*1e3:.1f} ms = {n/dt/1e9:.2f} GB/s raw; chunks {g.gzip_chunks()} false starts {g.gzip_false_starts()} equal={bytes(st)==bytes(want)}", flush=True)
    if os.environ.get("GZ_HOST", "1") == "1":
        os.environ["FQGPU_NO_GZIP_DEVICE"] = "1"
        t0 = time.perf_counter()
        st = g.count_file(path)
        dt = time.perf_counter() - t0
        print(f"host zlib: {dt*1e3:.1f} ms = {n/dt/1e9:.2f} GB/s raw equal={bytes(st)==bytes(want)}", flush=True)
        os.environ.pop("FQGPU_NO_GZIP_DEVICE")
    print(path)
