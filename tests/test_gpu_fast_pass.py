"""GPU parity of the experimental register-resident pass 0 (FQGPU_SCAN=fast, fq_scan_fast.cuh).

The fast kernel handles well-formed input itself and abandons a span otherwise (bytes >= 0x80, a '\\r' in a
sequence / quality line, three newlines in a 32-byte group, ...); the stitch kernel then sends the span to the
exact second pass.  Either way the result must equal the oracle bit for bit.
"""
import numpy as np
import pytest

import seq_collection_b200 as fq
from oracle import fq_oracle as O
from tests import corpus
from tests.test_gpu_parity import assert_equal_stats

pytestmark = pytest.mark.gpu


@pytest.fixture()
def fast(monkeypatch):
    monkeypatch.setenv("FQGPU_SCAN", "fast")
    with fq.FqGpu(meta_records=100) as c:
        yield c


@pytest.mark.parametrize("name", sorted(corpus.edge_cases()))
def test_fast_edge_corpus(fast, name):
    data = corpus.edge_cases()[name]
    assert_equal_stats(fast.count_bytes(data).to_dict(), O.count(data, 100), name)


@pytest.mark.parametrize("seed,kw", [
    (1, dict(min_len=30, max_len=260)),                 # well-formed, ragged lengths: stays on the fast path
    (2, dict(min_len=1, max_len=40)),                   # short lines: three newlines per group -> abandoned spans
    (3, dict(min_len=30, max_len=260, crlf=True)),      # CRLF -> abandoned
    (4, dict(min_len=500, max_len=3000, final_newline=False)),  # lines beyond the position bins, open last line
])
def test_fast_random_fastq(fast, seed, kw):
    rng = np.random.default_rng(seed)
    data = corpus.random_fastq(rng, 3000, **kw)
    assert_equal_stats(fast.count_bytes(data).to_dict(), O.count(data, 100), f"seed={seed}")


def test_fast_split_scans_and_misaligned_pointers(fast):
    """One stream fed as two device scans cut at awkward offsets, from unaligned pointers."""
    import torch

    rng = np.random.default_rng(11)
    data = corpus.random_fastq(rng, 2000, min_len=60, max_len=200)
    want = O.count(data, 100)
    for misalign in (0, 3, 15):
        buf = torch.zeros(len(data) + 64, dtype=torch.uint8, device="cuda")
        buf[misalign:misalign + len(data)] = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
        for cut in (0, 1, 31, 32, 33, 12288, 12289, 16384, len(data) // 2, len(data) - 1, len(data)):
            fast.reset()
            fast.scan_device(buf.data_ptr() + misalign, cut)
            fast.scan_device(buf.data_ptr() + misalign + cut, len(data) - cut)
            assert_equal_stats(fast.finish().to_dict(), want, f"misalign={misalign} cut={cut}")


def test_fast_core_only(monkeypatch):
    monkeypatch.setenv("FQGPU_SCAN", "fast")
    rng = np.random.default_rng(12)
    data = corpus.random_fastq(rng, 3000, min_len=60, max_len=200)
    with fq.FqGpu(meta_records=100, flags=fq.F_CORE_ONLY) as c, fq.FqGpu(meta_records=100) as full:
        got, ref = c.count_bytes(data), full.count_bytes(data)
        assert (got.reads, got.bases, got.gc_bases, got.n_bases) == (ref.reads, ref.bases, ref.gc_bases, ref.n_bases)
        assert fq.fq_count_row(got) == O.fq_count_row(O.count(data, 100))


def test_fast_synthetic_illumina(fast):
    """The bench shape at moderate size (many spans, every span on the fast path)."""
    import torch

    n = 64 << 20
    n -= n % 360
    buf = torch.empty(n + 4096, dtype=torch.uint8, device="cuda")
    fast.synth_illumina(buf.data_ptr(), n, 0, n // 360, 20240229)
    got = fast.count_device(buf.data_ptr(), n)
    data = bytes(buf[:n].cpu().numpy())
    assert_equal_stats(got.to_dict(), O.count(data, 100), "illumina 64 MiB")
