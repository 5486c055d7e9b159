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


def _soup(rng, size):
    """Random bytes with FASTQ-like structure broken in random ways: the alphabet is heavy with the bytes the scan
    treats specially ('\\n', '\\r', '@', '+'), plus a few high bytes and NULs."""
    mode = int(rng.integers(0, 4))
    if mode == 0:
        alphabet = np.frombuffer(b"\n\n\r@+ACGTNI#5", dtype=np.uint8)
        return bytes(rng.choice(alphabet, size=size))
    data = bytearray(corpus.random_fastq(rng, max(1, size // 160), min_len=1, max_len=300, crlf=bool(rng.integers(0, 2)),
                                         final_newline=bool(rng.integers(0, 2))))
    n_mut = int(rng.integers(0, 12)) if mode < 3 else 0
    for _ in range(n_mut):
        if not data:
            break
        pos = int(rng.integers(0, len(data)))
        kind = int(rng.integers(0, 5))
        if kind == 0:
            data[pos:pos] = b"\n"
        elif kind == 1:
            del data[pos:pos + int(rng.integers(1, 40))]
        elif kind == 2:
            data[pos] = int(rng.choice([0x80, 0xFF, 0x00, 0x0D, 0x40, 0x2B]))
        elif kind == 3:
            data[pos:pos] = b"\r"
        else:
            data[pos:pos] = b"\n" * int(rng.integers(2, 6))
    return bytes(data)


@pytest.mark.parametrize("seed", range(24))
def test_byte_soup_both_kernels(seed, monkeypatch):
    """Differential fuzz: default pass 0 and the fast pass 0 against the oracle, whole and split in two scans."""
    import torch

    rng = np.random.default_rng(1000 + seed)
    data = _soup(rng, int(rng.integers(1, 400_000)))
    want = O.count(data, 100)
    buf = torch.frombuffer(bytearray(data) + bytearray(64), dtype=torch.uint8).cuda()
    cut = int(rng.integers(0, len(data) + 1))
    for mode in ("", "fast"):
        if mode:
            monkeypatch.setenv("FQGPU_SCAN", mode)
        else:
            monkeypatch.delenv("FQGPU_SCAN", raising=False)
        with fq.FqGpu(meta_records=100) as c:
            assert_equal_stats(c.count_bytes(data).to_dict(), want, f"seed={seed} mode={mode or 'default'} whole")
            c.reset()
            c.scan_device(buf.data_ptr(), cut)
            c.scan_device(buf.data_ptr() + cut, len(data) - cut)
            assert_equal_stats(c.finish().to_dict(), want, f"seed={seed} mode={mode or 'default'} cut={cut}")
