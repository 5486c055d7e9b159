"""On-device inflate of ordinary gzip (csrc/fq_gzip.cu) behind fqgpu_count_file*: one DEFLATE stream cut into chunks
that are decoded in parallel from guessed block starts.  The statistics must equal the oracle's on the same bytes
(which the oracle reads through zlib, like the reference: src/utils/gzip_stream.nim:16-17), for every block type,
compression level, chunk and batch size; whatever the device path cannot prove must behave exactly like the host
zlib path (FQGPU_NO_GZIP_DEVICE=1)."""
import gzip
import zlib

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


def gz_bytes(data: bytes, level: int = 6, flush_every: int = 0, flush_mode: int = zlib.Z_SYNC_FLUSH, strategy: int = 0) -> bytes:
    """One gzip member; optionally with flush points (empty stored blocks / independent blocks) inside."""
    co = zlib.compressobj(level, zlib.DEFLATED, 31, 8, strategy)
    if not flush_every:
        return co.compress(data) + co.flush()
    parts = []
    for i in range(0, len(data), flush_every):
        parts.append(co.compress(data[i:i + flush_every]))
        parts.append(co.flush(flush_mode))
    parts.append(co.flush())
    return b"".join(parts)


def _write(tmp_path, name, blob):
    p = tmp_path / name
    p.write_bytes(blob)
    return str(p)


def _bulk(nbytes: int, first: int = 0) -> bytes:
    return bytes(O.synth_illumina_bytes(first, nbytes))


@pytest.mark.parametrize("level", [1, 4, 6, 9])
@pytest.mark.parametrize("chunk_kb", [1, 4, 16])
def test_gzip_equals_oracle(ctx, tmp_path, monkeypatch, level, chunk_kb):
    monkeypatch.setenv("FQGPU_GZ_CHUNK_KB", str(chunk_kb))
    rng = np.random.default_rng(31 * level + chunk_kb)
    data = corpus.random_fastq(rng, 5000, min_len=20, max_len=300, final_newline=(level != 4)) + _bulk(1_500_000)
    path = _write(tmp_path, f"l{level}.fq.gz", gz_bytes(data, level))
    st = ctx.count_file(path)
    assert ctx.gzip_chunks() > 4, "the device path was not taken"
    assert_equal_stats(st.to_dict(), O.count(data, 100), f"level={level} chunk={chunk_kb} KiB")
    assert_equal_stats(st.to_dict(), O.count_file(path, 100), "oracle through zlib")


def test_gzip_batches_carry_the_window(ctx, tmp_path, monkeypatch):
    """Batches of 1 MiB of compressed bytes: each starts where the one before stopped, with its last 32 KiB."""
    monkeypatch.setenv("FQGPU_GZ_BATCH_MB", "1")
    monkeypatch.setenv("FQGPU_GZ_CHUNK_KB", "8")
    data = _bulk(18_000_000, first=12345)
    blob = gz_bytes(data, 6)
    assert len(blob) > 3 << 20
    path = _write(tmp_path, "batches.fq.gz", blob)
    st = ctx.count_file(path)
    assert ctx.gzip_chunks() > 100
    assert_equal_stats(st.to_dict(), O.count(data, 100), "five batches")


def test_gzip_repetitive_input_long_matches(ctx, tmp_path, monkeypatch):
    """The same few records over and over: nearly every byte is a match, the runs overlap themselves, and a byte of
    the first chunk is copied forward through every later chunk (markers resolved through a chain of windows)."""
    monkeypatch.setenv("FQGPU_GZ_CHUNK_KB", "1")
    unit = corpus.rec(b"@r", b"ACGTNACGTTGCA" * 7, b"I!5#~" * 18 + b"J") + corpus.rec(b"@s/2", b"GGGGGGGGGG", b"##########")
    data = unit * 40000 + b"@t\nAC\n+\nII"
    for level in (1, 9):
        path = _write(tmp_path, f"rep{level}.fq.gz", gz_bytes(data, level))
        st = ctx.count_file(path)
        assert_equal_stats(st.to_dict(), O.count(data, 100), f"repetitive, level {level}")
    run = b"@q\n" + b"A" * 3_000_000 + b"\n+\n" + b"I" * 3_000_000 + b"\n"
    path = _write(tmp_path, "run.fq.gz", gz_bytes(run, 6))
    assert_equal_stats(ctx.count_file(path).to_dict(), O.count(run, 100), "period-1 runs")


@pytest.mark.parametrize("mode", [zlib.Z_SYNC_FLUSH, zlib.Z_FULL_FLUSH])
def test_gzip_flush_points_stored_and_fixed_blocks(ctx, tmp_path, monkeypatch, mode):
    """Flush points put empty stored blocks into the stream; short pieces are written with the fixed codes."""
    monkeypatch.setenv("FQGPU_GZ_CHUNK_KB", "1")
    rng = np.random.default_rng(5)
    data = corpus.random_fastq(rng, 4000, min_len=30, max_len=200)
    for every in (97, 5000, 70000):
        path = _write(tmp_path, f"flush{every}.fq.gz", gz_bytes(data, 6, flush_every=every, flush_mode=mode))
        assert_equal_stats(ctx.count_file(path).to_dict(), O.count(data, 100), f"flush every {every}")


def test_gzip_other_strategies_and_level0(ctx, tmp_path, monkeypatch):
    monkeypatch.setenv("FQGPU_GZ_CHUNK_KB", "2")
    rng = np.random.default_rng(6)
    data = corpus.random_fastq(rng, 6000, min_len=30, max_len=200)
    for name, blob in {
        "stored": gz_bytes(data, 0),
        "huffman_only": gz_bytes(data, 6, strategy=zlib.Z_HUFFMAN_ONLY),
        "rle": gz_bytes(data, 6, strategy=zlib.Z_RLE),
        "fixed": gz_bytes(data, 6, strategy=zlib.Z_FIXED),
    }.items():
        path = _write(tmp_path, name + ".fq.gz", blob)
        assert_equal_stats(ctx.count_file(path).to_dict(), O.count(data, 100), name)


def test_gzip_edge_corpus(ctx, tmp_path):
    for name, data in corpus.edge_cases().items():
        path = str(tmp_path / (name + ".fq.gz"))
        with gzip.open(path, "wb") as f:  # (writes the FNAME header field)
            f.write(data)
        assert_equal_stats(ctx.count_file(path).to_dict(), O.count(data, 100), name)


def test_gzip_concatenated_members_and_header_fields(ctx, tmp_path, monkeypatch):
    monkeypatch.setenv("FQGPU_GZ_CHUNK_KB", "4")
    a, b, c = _bulk(700_000), _bulk(900_000, first=700_000), _bulk(100_000, first=1_600_000)
    # FEXTRA + FNAME + FCOMMENT + FHCRC on the second member
    body = zlib.compressobj(6, zlib.DEFLATED, -15)
    raw_b = body.compress(b) + body.flush()
    hdr = bytes([0x1f, 0x8b, 8, 2 | 4 | 8 | 16, 0, 0, 0, 0, 0, 3]) + (5).to_bytes(2, "little") + b"XX\x01\x00Z" + b"name.fq\0" + b"a comment\0"
    hdr += (zlib.crc32(hdr) & 0xFFFF).to_bytes(2, "little")
    member_b = hdr + raw_b + zlib.crc32(b).to_bytes(4, "little") + (len(b) & 0xFFFFFFFF).to_bytes(4, "little")
    assert gzip.decompress(member_b) == b
    blob = gz_bytes(a, 6) + member_b + gz_bytes(c, 9)
    path = _write(tmp_path, "members.fq.gz", blob)
    st = ctx.count_file(path)
    assert ctx.gzip_chunks() > 10
    assert_equal_stats(st.to_dict(), O.count(a + b + c, 100), "three members")
    # trailing bytes that are not a member are ignored (gzread does the same)
    path = _write(tmp_path, "garbage.fq.gz", gz_bytes(a, 6) + b"\0" * 100 + b"not gzip")
    assert_equal_stats(ctx.count_file(path).to_dict(), O.count(a, 100), "trailing garbage")
    assert_equal_stats(ctx.count_file(path).to_dict(), O.count_file(path, 100), "trailing garbage, oracle through zlib")


def test_broken_gzip_behaves_like_the_zlib_path(ctx, tmp_path, monkeypatch):
    monkeypatch.setenv("FQGPU_GZ_CHUNK_KB", "2")
    data = _bulk(2_000_000)
    good = gz_bytes(data, 6)
    mid = len(good) // 2
    cases = {
        "truncated": good[:mid],
        "truncated_in_trailer": good[:-5],
        "corrupt_payload": good[:mid] + bytes([good[mid] ^ 0x5A, good[mid + 1] ^ 0xA5]) + good[mid + 2:],
        "isize_flip": good[:-1] + bytes([good[-1] ^ 1]),
        "not_gzip_at_all": data[:100_000],
        "empty_file": b"",
        "header_only": good[:10],
    }
    for name, blob in cases.items():
        path = _write(tmp_path, name + ".fq.gz", blob)

        def outcome():
            try:
                return ("ok", ctx.count_file(path).to_dict())
            except fq.FqGpuError as e:
                return ("error", e.code)

        monkeypatch.delenv("FQGPU_NO_GZIP_DEVICE", raising=False)
        dev = outcome()
        monkeypatch.setenv("FQGPU_NO_GZIP_DEVICE", "1")
        host = outcome()
        monkeypatch.delenv("FQGPU_NO_GZIP_DEVICE", raising=False)
        assert dev[0] == host[0], name
        if dev[0] == "ok":
            assert_equal_stats(dev[1], host[1], name)
        else:
            assert dev[1] == host[1], name


def test_gzip_one_pass_arena_and_second_pass(ctx, tmp_path, monkeypatch):
    """By default a batch is decoded once, every chunk's symbols into its own slot of an arena sized for 8 output bytes
    per compressed byte; what does not fit there (data that compresses better, here also an arena made too small on
    purpose) is decoded a second time into exact places, and FQGPU_GZ_TWO_PASS=1 always does that.  Same result."""
    monkeypatch.setenv("FQGPU_GZ_CHUNK_KB", "4")
    data = _bulk(6_000_000, first=4242)
    want = O.count(data, 100)
    path = _write(tmp_path, "plain.fq.gz", gz_bytes(data, 6))
    st = ctx.count_file(path)
    assert ctx.gzip_chunks() > 10 and ctx.gzip_second_passes() == 0
    assert_equal_stats(st.to_dict(), want, "one pass")
    monkeypatch.setenv("FQGPU_GZ_ARENA_RATIO", "2")  # FASTQ compresses about 4.4 : 1
    st = ctx.count_file(path)
    assert ctx.gzip_chunks() > 10 and ctx.gzip_second_passes() >= 1
    assert_equal_stats(st.to_dict(), want, "arena too small: second pass")
    monkeypatch.delenv("FQGPU_GZ_ARENA_RATIO")
    monkeypatch.setenv("FQGPU_GZ_TWO_PASS", "1")
    st = ctx.count_file(path)
    assert ctx.gzip_chunks() > 10 and ctx.gzip_second_passes() == 0
    assert_equal_stats(st.to_dict(), want, "two passes")
    monkeypatch.delenv("FQGPU_GZ_TWO_PASS")
    unit = corpus.rec(b"@r", b"ACGTNACGTTGCA" * 7, b"I!5#~" * 18 + b"J")
    rep = unit * 60000  # compresses far better than 8 : 1
    path = _write(tmp_path, "rep.fq.gz", gz_bytes(rep, 6))
    st = ctx.count_file(path)
    assert ctx.gzip_second_passes() >= 1
    assert_equal_stats(st.to_dict(), O.count(rep, 100), "repetitive: second pass")


def test_gzip_default_chunking_large_file(ctx, tmp_path):
    """Default chunk and batch sizes on a file large enough for thousands of chunks; no false starts expected to
    break the chain (they are counted, and skipped)."""
    data = _bulk(120_000_000, first=999)
    path = _write(tmp_path, "large.fq.gz", gz_bytes(data, 6))
    st = ctx.count_file(path)
    assert ctx.gzip_chunks() > 500
    want = O.count(data, 100)
    assert_equal_stats(st.to_dict(), want, "120 MB")
