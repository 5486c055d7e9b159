"""On-device BGZF inflate (csrc/fq_bgzf.cu) behind fqgpu_count_file*: the statistics must equal the oracle's on
the same .gz (which the oracle reads through zlib, like the reference), for every DEFLATE block type, and anything
that is not well-formed BGZF must behave exactly like the host zlib path."""
import gzip
import os

import numpy as np
import pytest

import seq_collection_b200 as fq
from oracle import fq_oracle as O
from tests import corpus
from tests.test_gpu_parity import assert_equal_stats

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    with fq.FqGpu(meta_records=100) as c:
        yield c


def _write(tmp_path, name, blob):
    p = tmp_path / name
    p.write_bytes(blob)
    return str(p)


@pytest.mark.parametrize("level,block", [(6, 65280), (1, 65280), (9, 30000), (0, 65280), (6, 61), (6, 7)])
def test_bgzf_equals_oracle(ctx, tmp_path, level, block):
    """level 0 = stored blocks, tiny members = fixed Huffman codes, the rest dynamic codes."""
    rng = np.random.default_rng(100 + level + block)
    n = 6000 if block > 1000 else 40
    data = corpus.random_fastq(rng, n, min_len=20, max_len=300, final_newline=(level != 1))
    path = _write(tmp_path, f"l{level}_b{block}.fq.gz", corpus.bgzf_bytes(data, block=block, level=level))
    st = ctx.count_file(path)
    assert ctx.bgzf_members() >= len(data) // block, "the device path was not taken"
    assert_equal_stats(st.to_dict(), O.count(data, 100), f"level={level} block={block}")
    assert_equal_stats(st.to_dict(), O.count_file(path, 100), "oracle through zlib")


def test_bgzf_edge_corpus_and_no_eof_member(ctx, tmp_path):
    for name, data in corpus.edge_cases().items():
        path = _write(tmp_path, name + ".fq.gz", corpus.bgzf_bytes(data, block=4096, eof=(len(name) % 2 == 0)))
        assert_equal_stats(ctx.count_file(path).to_dict(), O.count(data, 100), name)


def test_plain_gzip_is_not_taken_for_bgzf(ctx, tmp_path):
    rng = np.random.default_rng(7)
    data = corpus.random_fastq(rng, 500)
    path = str(tmp_path / "plain.fq.gz")
    with gzip.open(path, "wb") as f:
        f.write(data)
    st = ctx.count_file(path)
    assert ctx.bgzf_members() == 0
    assert_equal_stats(st.to_dict(), O.count(data, 100), "plain gzip")


def test_malformed_bgzf_behaves_like_zlib(ctx, tmp_path, monkeypatch):
    rng = np.random.default_rng(8)
    data = corpus.random_fastq(rng, 3000)
    good = corpus.bgzf_bytes(data, block=20000)
    cases = {
        "corrupt_payload": good[:300] + bytes([good[300] ^ 0x5A, good[301] ^ 0xA5]) + good[302:],
        "truncated": good[: len(good) // 2],
        # sizes stay right, only the CRC-32 of the member tells: a literal byte of a stored block, the CRC field itself
        "stored_byte_flip": (lambda b: b[:40] + bytes([b[40] ^ 1]) + b[41:])(corpus.bgzf_bytes(data, block=20000, level=0)),
        "crc_field_flip": (lambda b, e: b[:e - 8] + bytes([b[e - 8] ^ 0x10]) + b[e - 7:])(good, 18 + (good[16] | good[17] << 8) - 17),
        "bgzf_then_plain_member": good + gzip.compress(b"@x\nACGT\n+\nIIII\n"),
    }
    for name, blob in cases.items():
        path = _write(tmp_path, name + ".fq.gz", blob)

        def outcome():
            try:
                return ("ok", ctx.count_file(path).to_dict())
            except fq.FqGpuError as e:
                return ("error", e.code)

        dev = outcome()
        monkeypatch.setenv("FQGPU_NO_BGZF", "1")
        host = outcome()
        monkeypatch.delenv("FQGPU_NO_BGZF")
        assert dev == host, name


def test_bgzf_many_members_synthetic(ctx, tmp_path):
    """The bench shape: 100k records (36 MB) in 64 KiB members; also through fqgpu_count_files."""
    import torch

    n = 360 * 100_000
    buf = torch.empty(n, dtype=torch.uint8, device="cuda")
    ctx.synth_illumina(buf.data_ptr(), n, 0, 100_000, 20240229)
    data = bytes(buf.cpu().numpy())
    path = _write(tmp_path, "synth.fq.gz", corpus.bgzf_bytes(data, level=1))
    st = ctx.count_file(path)
    assert ctx.bgzf_members() == (n + 65279) // 65280 + 1
    want = O.count(data, 100)
    assert_equal_stats(st.to_dict(), want, "synthetic bgzf")
    rc, res = fq.count_files([path, path], meta_records=100)
    assert rc == fq.OK and all(r == fq.OK for r, _ in res)
    for _, s in res:
        assert_equal_stats(s.to_dict(), want, "count_files bgzf")


def _inflate_stress_payloads():
    """Byte streams that reach the corners of the DEFLATE decoder: literal codes longer than the 10-bit table
    (geometric distribution over 256 byte values), distance codes longer than the 8-bit table and far matches
    (a 30 KB period), overlapping matches with every short period and with periods >= 32, maximal runs,
    incompressible noise (stored blocks at any level), and all of it glued together."""
    rng = np.random.default_rng(77)
    geo = np.minimum(rng.geometric(0.08, size=200_000) - 1, 255).astype(np.uint8).tobytes()
    base = rng.integers(0, 256, size=30_000, dtype=np.uint8).tobytes()
    far = base * 5
    periods = b"".join((bytes(rng.integers(65, 91, size=p, dtype=np.uint8)) * (700 // p + 2))[:700] for p in list(range(1, 40)) + [47, 63, 64, 65, 129, 257, 300])
    runs = b"A" * 70_000 + b"\n" + b"#" * 1000 + b"\n"
    noise = rng.integers(0, 256, size=150_000, dtype=np.uint8).tobytes()
    text = corpus.random_fastq(rng, 800, min_len=30, max_len=300)
    return {"geometric": geo, "far_matches": far, "periods": periods, "runs": runs, "noise": noise, "mix": text + geo[:50_000] + far[:70_000] + periods + text}


@pytest.mark.parametrize("strategy,level", [(0, 9), (0, 1), (1, 6), (2, 6), (3, 6), (4, 9)])  # default, filtered, huffman-only, rle, fixed
def test_bgzf_inflate_stress(ctx, tmp_path, strategy, level):
    for name, data in _inflate_stress_payloads().items():
        for block in (65280, 9973):
            path = _write(tmp_path, f"{name}_{strategy}_{level}_{block}.fq.gz", corpus.bgzf_bytes(data, block=block, level=level, strategy=strategy))
            st = ctx.count_file(path)
            assert ctx.bgzf_members() >= len(data) // block, f"{name}: the device path was not taken (decoder rejected a valid stream)"
            assert_equal_stats(st.to_dict(), O.count(data, 100), f"{name} strategy={strategy} level={level} block={block}")


def test_truncated_gzip_ends_like_the_reference_stream(ctx, tmp_path):
    """A gz file cut in the middle opens fine; the reference's stream simply ends where zlib stops (gzip_stream.nim:16-17)
    and the row of what was read is printed.  Same here: no error, the counts of the inflatable prefix."""
    import zlib

    rng = np.random.default_rng(21)
    data = corpus.random_fastq(rng, 4000)
    blob = gzip.compress(data)
    cut = blob[: len(blob) // 2]
    prefix = zlib.decompressobj(31).decompress(cut)
    assert 0 < len(prefix) < len(data)
    path = _write(tmp_path, "cut.fq.gz", cut)
    assert_equal_stats(ctx.count_file(path).to_dict(), O.count(prefix, 100), "truncated gzip")
