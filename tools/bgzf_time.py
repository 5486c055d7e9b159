"""Times fqgpu_count_file on a BGZF .fq.gz (device inflate); run on the GPU box."""
import os
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
import seq_collection_b200 as fq  # noqa: E402

n_records = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
n = n_records * bench.REC_BYTES
with fq.FqGpu(meta_records=100) as g:
    dev = torch.empty(n + 256, dtype=torch.uint8, device="cuda")
    g.synth_illumina(dev.data_ptr(), n, 0, n_records, bench.SEED_ILLUMINA)
    raw = dev[:n].cpu().numpy().tobytes()
    del dev
    tmp = tempfile.mkdtemp(prefix="bgt_")
    path = os.path.join(tmp, "r.fq.gz")
    with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as pool:
        parts = list(pool.map(bench._bgzf_member, [raw[i:i + 65280] for i in range(0, n, 65280)]))
    with open(path, "wb") as f:
        for p in parts:
            f.write(p)
        f.write(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))
    want = g.synth_illumina_tally(0, n_records, bench.SEED_ILLUMINA)
    for rep in range(4):
        t0 = time.perf_counter()
        st = g.count_file(path)
        dt = time.perf_counter() - t0
        print(f"bgzf device: {dt*1e3:.1f} ms = {n/dt/1e9:.2f} GB/s raw; members {g.bgzf_members()} equal={bytes(st)==bytes(want)}", flush=True)
