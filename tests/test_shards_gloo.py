"""N>1 host path on CPU (world_size 2, gloo): every rank puts the block of its byte range into its slot,
the slots are SUM-all-reduced (the path's one collective), and every rank runs the library's combine step
(fqgpu_shard_combine_host -- pure host arithmetic in libfqgpu.so) and must obtain the oracle's result."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, datasets, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    import seq_collection_b200 as fq
    from tests import shard_emulator as E

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert fq.load_library().fqgpu_shard_block_words() == E.SHARD_WORDS
        out = []
        for data, cuts, wrong in datasets:
            edges = [0] + list(cuts) + [len(data)]
            shard = data[edges[rank]:edges[rank + 1]]
            phase = data[:edges[rank]].count(b"\n") & 3
            hyp = (phase + 1) & 3 if (wrong and rank == wrong) else phase
            slots = torch.zeros(world * E.SHARD_WORDS, dtype=torch.int64)
            blk = E.make_block(shard, rank, hyp)
            slots[rank * E.SHARD_WORDS:(rank + 1) * E.SHARD_WORDS] = torch.from_numpy(blk.view(np.int64))
            dist.all_reduce(slots)  # SUM of disjoint slots == gather
            arr = slots.numpy()
            rc, st = fq.shard_combine_host(world, arr.ctypes.data, 0)
            out.append((rc, st.to_dict() if rc == 0 else None))
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def _datasets():
    from tests import corpus

    rng = np.random.default_rng(77)
    ds = []
    d1 = corpus.random_fastq(rng, 300, min_len=10, max_len=200, crlf=True, final_newline=False)
    for cut in (1, 97, len(d1) // 2, len(d1) // 2 + 1, len(d1) - 1):
        ds.append((d1, [cut], 0))
    cases = corpus.edge_cases()
    d2 = cases["crlf"] + cases["blank_lines"] + cases["qual_starts_with_at"] + cases["trailing_cr_no_lf"]
    for cut in range(0, len(d2) + 1, 3):
        ds.append((d2, [cut], 0))
    ds.append((d1, [len(d1) // 3], 1))  # rank 1 exports a wrong hypothesis -> combine must return ERETRY
    return ds


def test_two_rank_gloo_combine():
    import torch.multiprocessing as mp

    from oracle import fq_oracle as O

    datasets = _datasets()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 400)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, datasets, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    meta_fields = ("meta_qual_min", "meta_qual_max", "meta_lines", "meta_status")
    for i, (data, cuts, wrong) in enumerate(datasets):
        r0, r1 = results[0][i], results[1][i]
        assert r0 == r1, "both ranks must compute the same result"
        if wrong:
            assert r0[0] == 1  # FQGPU_ERETRY
            continue
        assert r0[0] == 0
        want = O.count(data, 0)
        got = r0[1]
        for k in want:
            if k not in meta_fields:
                assert got[k] == want[k], (i, cuts, k)
