"""Multi-GPU shard protocol exercised on ONE device: `world` contexts each scan a byte range
of the same stream with an unknown start phase, export their block into one buffer (standing in for
the SUM all-reduce of disjoint slots), and the combine step must reproduce the oracle bit for bit --
including stitched lines, CRLF pairs split across shard edges and wrong resync hypotheses (rescan)."""
import numpy as np
import pytest

import seq_collection_b200 as fq
from oracle import fq_oracle as O
from tests import corpus
from tests.test_gpu_parity import _adversarial_records, assert_equal_stats

pytestmark = pytest.mark.gpu

META_FIELDS = ("meta_qual_min", "meta_qual_max", "meta_lines", "meta_status")


def check(got: dict, data: bytes, first_cut: int, msg: str, meta: int = 100):
    """The fq-meta range is a prefix scan: exact when the sampled 4*n lines lie inside shard 0, otherwise
    the result carries the 'incomplete' flag (meta_status & 0x100) and the range fields are not compared."""
    want = O.count(data, meta)
    lines_in_shard0 = data[:first_cut].count(b"\n")
    if lines_in_shard0 >= 4 * meta or first_cut >= len(data):
        assert_equal_stats(got, want, msg)
    else:
        assert got["meta_status"] & 0x100, msg + ": incomplete fq-meta scan must be flagged"
        g2, w2 = dict(got), dict(want)
        for k in META_FIELDS:
            g2.pop(k); w2.pop(k)
        for k in O.SCALARS:
            if k not in META_FIELDS:
                assert g2[k] == w2[k], f"{msg}: {k}: gpu={g2[k]} oracle={w2[k]}"
        for k in O.ARRAYS:
            assert g2[k] == w2[k], f"{msg}: {k}"


def sharded_count(data: bytes, cuts, meta=100, max_rounds=None):
    """Both forms of the exchange step must agree: the caller's collective + fqgpu_shard_combine (host arithmetic), and the
    library's own collective (peer-memory all-gather + device combine, fqgpu_shard_exchange_*)."""
    a = _sharded_count_allreduce(data, cuts, meta)
    b = _sharded_count_exchange(data, cuts, meta)
    assert a == b, "the in-library collective disagrees with all-reduce + host combine"
    return a


def _sharded_count_exchange(data: bytes, cuts, meta=100):
    import torch

    world = len(cuts) + 1
    edges = [0] + list(cuts) + [len(data)]
    buf = torch.frombuffer(bytearray(data) if data else bytearray(b"\0"), dtype=torch.uint8).cuda()
    ctxs = [fq.FqGpu(meta_records=meta) for _ in range(world)]
    try:
        for g, c in enumerate(ctxs):
            c.shard_exchange_create(g, world)
        xbufs = [c.shard_xbuf() for c in ctxs]
        for c in ctxs:
            c.shard_exchange_set_peers(xbufs)
        for step in range(2):  # twice: the exchange buffers are double-buffered by the step's parity
            for g, c in enumerate(ctxs):
                c.shard_begin(g, world)
                c.scan_device(buf.data_ptr() + edges[g], edges[g + 1] - edges[g])
            rounds = 0
            while True:
                for c in ctxs:
                    c.shard_exchange_start()  # asynchronous: the kernels of all ranks wait for each other on the device
                results = [c.shard_exchange_finish() for c in ctxs]
                rcs = {rc for rc, _ in results}
                assert len(rcs) == 1, "every rank must reach the same verdict"
                if rcs == {0}:
                    dicts = [st.to_dict() for _, st in results]
                    assert all(d == dicts[0] for d in dicts), "every rank must compute the same stats"
                    break
                rounds += 1
                assert rounds <= world, "rescan did not converge"
                for g, c in enumerate(ctxs):
                    if c.shard_rescan(c.shard_gathered()) == fq.ERETRY:
                        c.scan_device(buf.data_ptr() + edges[g], edges[g + 1] - edges[g])
            if step == 0:
                first = (dicts[0], rounds)
            else:
                assert dicts[0] == first[0], "second step differs"
        return first
    finally:
        for c in ctxs:
            c.close()


def _sharded_count_allreduce(data: bytes, cuts, meta=100, max_rounds=None):
    import torch

    world = len(cuts) + 1
    edges = [0] + list(cuts) + [len(data)]
    buf = torch.frombuffer(bytearray(data) if data else bytearray(b"\0"), dtype=torch.uint8).cuda()
    ctxs = [fq.FqGpu(meta_records=meta) for _ in range(world)]
    bw = ctxs[0].shard_block_words()
    blocks = torch.zeros(world * bw, dtype=torch.int64, device="cuda")
    try:
        for g, c in enumerate(ctxs):
            c.shard_begin(g, world)
            c.scan_device(buf.data_ptr() + edges[g], edges[g + 1] - edges[g])
        rounds = 0
        while True:
            blocks.zero_()
            torch.cuda.synchronize()  # the contexts run on their own non-blocking streams
            for c in ctxs:
                c.shard_export(blocks.data_ptr())
            torch.cuda.synchronize()
            results = [c.shard_combine(blocks.data_ptr()) for c in ctxs]
            rcs = {rc for rc, _ in results}
            assert len(rcs) == 1, "every rank must reach the same verdict"
            if rcs == {0}:
                dicts = [st.to_dict() for _, st in results]
                assert all(d == dicts[0] for d in dicts), "every rank must compute the same stats"
                return dicts[0], rounds
            rounds += 1
            assert rounds <= world, "rescan did not converge"
            for g, c in enumerate(ctxs):
                if c.shard_rescan(blocks.data_ptr()) == fq.ERETRY:
                    c.scan_device(buf.data_ptr() + edges[g], edges[g + 1] - edges[g])
    finally:
        for c in ctxs:
            c.close()


@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_random_fastq_even_shards(world):
    rng = np.random.default_rng(100 + world)
    data = corpus.random_fastq(rng, 3000, min_len=20, max_len=250, crlf=bool(world & 1), final_newline=bool(world & 2))
    cuts = [len(data) * g // world for g in range(1, world)]
    got, rounds = sharded_count(data, cuts)
    check(got, data, cuts[0], f"world={world}")
    assert rounds == 0, "well-formed FASTQ must need exactly one exchange"


def test_every_cut_position_two_ranks():
    cases = corpus.edge_cases()
    data = cases["crlf"] + cases["qual_starts_with_at"] + corpus.rec(b"x", b"ACGTNN", b"IIII55") * 3 + cases["crlf_no_final"]
    for cut in range(0, len(data) + 1):
        got, _ = sharded_count(data, [cut])
        check(got, data, cut, f"cut={cut}")


def test_three_ranks_tiny_shards_and_empty_shard():
    data = corpus.rec(b"a", b"ACGTACGTAC", b"IIIIIIIIII") * 4
    for c1 in range(0, len(data), 7):
        for c2 in (c1, c1 + 1, c1 + 5, len(data)):
            if c2 > len(data):
                continue
            got, _ = sharded_count(data, [c1, c2])
            check(got, data, c1, f"cuts={c1},{c2}")


def test_long_lines_spanning_shards():
    rng = np.random.default_rng(4)
    L = 300_000
    s = bytes(rng.choice(list(b"ACGT"), size=L).astype(np.uint8))
    q = bytes(rng.integers(35, 80, size=L, dtype=np.uint8))
    data = b"@long\n" + s + b"\n+\n" + q + b"\n" + corpus.random_fastq(rng, 200)
    got, _ = sharded_count(data, [len(data) // 4, len(data) // 2, 3 * len(data) // 4])
    check(got, data, len(data) // 4, "long")


def test_wrong_hypothesis_triggers_rescan():
    data = _adversarial_records(4000)
    cuts = [len(data) // 3 + 11, 2 * len(data) // 3 + 5]
    got, rounds = sharded_count(data, cuts)
    check(got, data, cuts[0], "adversarial shards")
    assert rounds >= 1


def test_blank_line_shift_and_no_pattern_shard():
    rng = np.random.default_rng(8)
    body = corpus.random_fastq(rng, 2000, min_len=40, max_len=120)
    data = body[:50000] + b"\n" + body[50000:] + (b"ACGTACGT\n" * 30000)
    cuts = [len(data) // 4, len(data) // 2, len(data) - 100000]
    got, rounds = sharded_count(data, cuts)
    check(got, data, cuts[0], "blank + no pattern")
