"""GPU parity: libfqgpu (through its C ABI) against the CPU oracle, bit-exact on every integer.

Covers the reference's golden fixtures, the edge corpus, every chunk/tile boundary offset, the
pinned streaming ring, unaligned device pointers and the synthetic generators at moderate size.
Nothing here reads /root/reference (it does not exist on the GPU box).
"""
import ctypes as C
import os

import numpy as np
import pytest

import seq_collection_b200 as fq
from oracle import fq_oracle as O
from tests import corpus

pytestmark = pytest.mark.gpu

TILE = 32768


def _torch():
    import torch

    return torch


def assert_equal_stats(got: dict, want: dict, ctxmsg=""):
    for k in O.SCALARS:
        assert got[k] == want[k], f"{ctxmsg}: {k}: gpu={got[k]} oracle={want[k]}"
    for k in O.ARRAYS:
        if got[k] != want[k]:
            bad = [(i, a, b) for i, (a, b) in enumerate(zip(got[k], want[k])) if a != b][:8]
            raise AssertionError(f"{ctxmsg}: {k} differs at (idx, gpu, oracle) {bad}")


@pytest.fixture(scope="module")
def ctx():
    c = fq.FqGpu(meta_records=100)
    yield c
    c.close()


def test_golden_fixtures_rows(ctx, golden_dir):
    """docs/fq-count.md rows + docs/fq-meta.md min/max through the product path (file API, incl. .gz)."""
    lines = open(os.path.join(golden_dir, "fq_count_docs.tsv")).read().splitlines()[1:]
    for ln in lines:
        name, reads, gc_content, gc, n, bases = ln.split("\t")
        path = os.path.join(golden_dir, "fastq", name)
        st = ctx.count_file(path)
        assert (st.reads, st.gc_bases, st.n_bases, st.bases) == (int(reads), int(gc), int(n), int(bases)), name
        assert fq.fq_count_row(st) == O.fq_count_row(O.count_file(path)), name
        assert_equal_stats(st.to_dict(), O.count_file(path, 100), name)
    for ln in open(os.path.join(golden_dir, "fq_meta_docs.tsv")).read().splitlines()[1:]:
        name, fmt, phred, multiple, qmin, qmax, n_lines = ln.split("\t")
        st = ctx.count_file(os.path.join(golden_dir, "fastq", name))
        f = fq.fq_meta_quality_fields(st)
        assert f == [fmt, phred, multiple.lower(), qmin, qmax, n_lines], name


def test_missing_file_is_eio(ctx):
    with pytest.raises(fq.FqGpuError) as ei:
        ctx.count_file("/nonexistent/file.fq")
    assert ei.value.code == fq.EIO and "Unable to open file" in ei.value.msg


@pytest.mark.parametrize("name", sorted(corpus.edge_cases()))
def test_edge_corpus(ctx, name):
    data = corpus.edge_cases()[name]
    assert_equal_stats(ctx.count_bytes(data).to_dict(), O.count(data, 100), name)


@pytest.mark.parametrize("meta_n", [1, 2, 3, 7, 1000])
def test_meta_records_values(meta_n):
    with fq.FqGpu(meta_records=meta_n) as c:
        for name, data in corpus.edge_cases().items():
            assert_equal_stats(c.count_bytes(data).to_dict(), O.count(data, meta_n), f"{name} n={meta_n}")


def test_chunk_edges_every_offset():
    """Feed one stream in two device scans split at every byte offset (carry resolved on device)."""
    torch = _torch()
    cases = corpus.edge_cases()
    data = cases["crlf"] + cases["qual_starts_with_at"] + cases["blank_lines"] + cases["crlf_no_final"]
    want = O.count(data, 100)
    buf = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
    with fq.FqGpu(meta_records=100) as c:
        for cut in range(0, len(data) + 1):
            c.reset()
            c.scan_device(buf.data_ptr(), cut)
            c.scan_device(buf.data_ptr() + cut, len(data) - cut)
            assert_equal_stats(c.finish().to_dict(), want, f"cut={cut}")


def test_streaming_ring_tiny_chunks():
    """fqgpu_acquire/submit with a ring of tiny pinned chunks: every record straddles chunk edges."""
    rng = np.random.default_rng(5)
    data = corpus.random_fastq(rng, 400, max_len=400, crlf=True, final_newline=False)
    want = O.count(data, 100)
    for chunk in (4096, 4096 * 3):
        with fq.FqGpu(meta_records=100, chunk_bytes=chunk, n_buffers=2) as c:
            c.reset()
            c.submit_bytes(data)
            assert_equal_stats(c.finish().to_dict(), want, f"chunk={chunk}")


def test_tile_edges_and_unaligned_pointers():
    """Lines, CRLF pairs and '\\n' bytes placed on and around tile boundaries; unaligned base pointers."""
    torch = _torch()
    rng = np.random.default_rng(9)
    body = corpus.random_fastq(rng, 300, min_len=30, max_len=260, crlf=True)
    with fq.FqGpu(meta_records=50) as c:
        for shift in list(range(0, 20)) + [TILE - 2, TILE - 1, TILE, TILE + 1]:
            data = b"@pad\n" + b"A" * shift + b"\n+\n" + b"I" * shift + b"\n" + body
            want = O.count(data, 50)
            for misalign in (0, 1, 7, 15):
                buf = torch.zeros(len(data) + 64, dtype=torch.uint8, device="cuda")
                buf[misalign:misalign + len(data)] = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
                st = c.count_device(buf.data_ptr() + misalign, len(data))
                assert_equal_stats(st.to_dict(), want, f"shift={shift} misalign={misalign}")


def test_dense_newlines_windowed_path(ctx):
    """More newlines per tile than the newline-index window holds."""
    data = (b"A\n" * 30000) + (b"\n" * 40000) + b"@x\nACGT\n+\nIIII\n"
    assert_equal_stats(ctx.count_bytes(data).to_dict(), O.count(data, 100), "dense")


def test_long_lines_across_many_tiles(ctx):
    rng = np.random.default_rng(3)
    recs = []
    for L in (1, TILE - 7, TILE, 3 * TILE + 5, 70000, 2048, 2049, 100):
        s = bytes(rng.choice(list(b"ACGTN"), size=L).astype(np.uint8))
        q = bytes(rng.integers(33, 90, size=L, dtype=np.uint8))
        recs.append(b"@r\n" + s + b"\n+\n" + q + b"\n")
    data = b"".join(recs)
    assert_equal_stats(ctx.count_bytes(data).to_dict(), O.count(data, 100), "long")


@pytest.mark.parametrize("kind,n_records", [("illumina", 300000), ("ont", 1500)])
def test_synthetic_generators_vs_oracle(kind, n_records):
    """Generated on the device, scanned on the device, copied back and scanned by the oracle."""
    torch = _torch()
    cap = 360 * n_records if kind == "illumina" else 64 << 20
    buf = torch.empty(cap, dtype=torch.uint8, device="cuda")
    with fq.FqGpu(meta_records=100) as c:
        if kind == "illumina":
            n = c.synth_illumina(buf.data_ptr(), cap, 0, n_records, 20240229)
            assert n == 360 * n_records
        else:
            n = c.synth_ont(buf.data_ptr(), cap, 0, n_records, 20240301)
        st = c.count_device(buf.data_ptr(), n)
        host = buf[:n].cpu().numpy()
        want = O.count(host, 100)
        assert_equal_stats(st.to_dict(), want, kind)
        assert st.reads == n_records
        if kind == "illumina":
            assert st.bases == 150 * n_records and st.seq_len_min == 150 == st.seq_len_max
            assert host[:56].tobytes().startswith(b"@A00156:217:HKJWGDSXX:")
            assert abs(st.gc_bases / (st.bases - st.n_bases) - 0.41) < 0.01
        else:
            assert st.seq_len_min >= 1000 and st.seq_len_max <= 100000
        # odd split of the same bytes must not change anything (carry on the device)
        c.reset()
        cut = n // 3 + 13
        c.scan_device(buf.data_ptr(), cut)
        c.scan_device(buf.data_ptr() + cut, n - cut)
        assert_equal_stats(c.finish().to_dict(), want, kind + " split")


def test_illumina_byte_range_generation():
    torch = _torch()
    n = 360 * 1000
    a = torch.empty(n, dtype=torch.uint8, device="cuda")
    b = torch.empty(n, dtype=torch.uint8, device="cuda")
    with fq.FqGpu() as c:
        c.synth_illumina(a.data_ptr(), n, 0, 1000, 7)
        cut = 123457
        c.synth_illumina_bytes(b.data_ptr(), 0, cut, 7)
        c.synth_illumina_bytes(b.data_ptr() + cut, cut, n - cut, 7)
    assert torch.equal(a, b)


def test_size_independent_properties_large():
    """1.44 GB Illumina stream: invariants that need no oracle pass over the full size."""
    torch = _torch()
    n_records = 4_000_000
    n = 360 * n_records
    buf = torch.empty(n, dtype=torch.uint8, device="cuda")
    with fq.FqGpu(meta_records=n_records) as c:
        c.synth_illumina(buf.data_ptr(), n, 0, n_records, 20240229)
        st = c.count_device(buf.data_ptr(), n)
        d = st.to_dict()
        assert d["bytes"] == n and d["lines"] == 4 * n_records and d["reads"] == n_records
        assert d["bases"] == 150 * n_records == sum(d["base_counts"])
        assert sum(d["qual_counts"]) == 150 * n_records
        assert d["seq_len_hist"][150] == n_records and d["qual_len_hist"][150] == n_records
        assert all(d["qual_pos_cnt"][p] == n_records for p in range(150)) and d["qual_pos_cnt"][150] == 0
        assert sum(d["qual_pos_sum"]) == sum(v * cnt for v, cnt in enumerate(d["qual_counts"]))
        assert set(i for i, v in enumerate(d["qual_counts"]) if v) == {ord("F"), ord(":"), ord(","), ord("#")}
        assert (d["meta_qual_min"], d["meta_qual_max"], d["meta_lines"]) == (2, 37, 4 * n_records)
        # linearity: the sum of two halves (independent streams cut at a record edge) equals the whole
        half = 360 * (n_records // 2)
        c2 = fq.FqGpu()
        s1 = c2.count_device(buf.data_ptr(), half).to_dict()
        s2 = c2.count_device(buf.data_ptr() + half, n - half).to_dict()
        c2.close()
        for k in ("reads", "bases", "gc_bases", "n_bases"):
            assert s1[k] + s2[k] == d[k]
        for k in ("base_counts", "qual_counts", "qual_pos_sum"):
            assert [x + y for x, y in zip(s1[k], s2[k])] == d[k]
        # a 64 MB prefix against the oracle
        m = 360 * 180_000
        assert_equal_stats(fq.FqGpu(meta_records=100).count_device(buf.data_ptr(), m).to_dict(),
                           O.count(buf[:m].cpu().numpy(), 100), "prefix")


def _adversarial_records(n):
    """Every quality line starts with '@' and every sequence line with '+': the content resync of a
    span start picks the wrong line as the header, so the stitch kernel's verification must catch the
    wrong phase and the span must be rescanned exactly."""
    out = []
    for i in range(n):
        L = 40 + (i * 7) % 90
        out.append(b"@r%d\n" % i + b"+" + b"ACGT" * (L // 4) + b"\n+\n" + b"@" + b"I" * (L // 4 * 4) + b"\n")
    return b"".join(out)


def test_wrong_resync_guess_is_verified_and_rescanned(ctx):
    data = _adversarial_records(6000)  # ~0.8 MB: dozens of spans
    assert_equal_stats(ctx.count_bytes(data).to_dict(), O.count(data, 100), "adversarial")


def test_blank_line_shifts_phase_large(ctx):
    """A stray blank line makes every following record start at line phase 1: guesses made on the
    '@'/'+' pattern are wrong for all later spans (reference semantics are line-number based)."""
    rng = np.random.default_rng(21)
    body = corpus.random_fastq(rng, 4000, min_len=50, max_len=150)
    data = body[:100000] + b"\n" + body[100000:]
    assert_equal_stats(ctx.count_bytes(data).to_dict(), O.count(data, 100), "blank-shift")


def test_no_resync_possible(ctx):
    """Spans that contain no '@'...'+' pattern at all (unknown phase -> counted only, then rescanned)."""
    data = (b"ACGTNACGTN" * 9 + b"\n") * 40000  # 3.6 MB of anonymous lines
    assert_equal_stats(ctx.count_bytes(data).to_dict(), O.count(data, 100), "no-pattern")
    data = b"@x\n" + b"G" * 2_000_000 + b"\n+\n" + b"5" * 2_000_000 + b"\n" + corpus.random_fastq(np.random.default_rng(2), 500)
    assert_equal_stats(ctx.count_bytes(data).to_dict(), O.count(data, 100), "two huge lines")


def test_span_edges_with_crlf(ctx):
    """CRLF records sized so that '\\r' / '\\n' pairs fall on span and tile boundaries."""
    for pad in range(0, 6):
        rec = b"@h\r\n" + b"ACGT" * 31 + b"\r\n+\r\n" + b"IIII" * 31 + b"\r\n"  # 260 bytes
        data = b"A" * pad + b"\n" + rec * 3000
        assert_equal_stats(ctx.count_bytes(data).to_dict(), O.count(data, 100), f"pad={pad}")


# ---- cases aimed at the mechanisms of the record-driven scan kernel (kernel K2) ----

@pytest.mark.gpu
def test_packed_position_sums_do_not_overflow(ctx):
    """Many quality lines of the highest printable byte: the 16-bit halves of the per-position table are
    flushed before they can carry (126 * 520 lines > 65535)."""
    rec = b"@q\n" + b"ACGT" * 5 + b"\n+\n" + b"~" * 20 + b"\n"       # 48-byte records, 341 quality lines per tile
    data = rec * 60000
    assert_equal_stats(ctx.count_bytes(data).to_dict(), O.count(data, 100), "tilde short")
    rec = b"@q\n" + b"A" * 3 + b"\n+\n" + b"~" * 3 + b"\n"             # 13-byte records: close to the newline-index cap
    data = rec * 200000
    assert_equal_stats(ctx.count_bytes(data).to_dict(), O.count(data, 100), "tilde dense")


@pytest.mark.gpu
def test_lines_around_position_bins_and_long_line_threshold(ctx):
    """Lengths around POS_BINS (512: table / overflow / straddling group), around the long-line threshold
    (32 full groups), and every alignment of the line start inside its 16-byte group."""
    rng = np.random.default_rng(11)
    recs = []
    for L in list(range(490, 560)) + list(range(1, 40)) + [1023, 1024, 1025, 2047, 4096]:
        s = bytes(rng.choice(list(b"ACGTN"), size=L).astype(np.uint8))
        q = bytes(rng.integers(33, 127, size=L, dtype=np.uint8))
        recs.append(b"@" + b"h" * int(rng.integers(1, 18)) + b"\n" + s + b"\n+\n" + q + b"\n")
    data = b"".join(recs) * 7
    assert_equal_stats(ctx.count_bytes(data).to_dict(), O.count(data, 100), "pos bins")


@pytest.mark.gpu
def test_mixed_long_and_short_lines_in_one_tile(ctx):
    """A tile with short reads and one long read: the long line goes through the shared long-line list, the
    short ones through the slot map with a small K."""
    rng = np.random.default_rng(12)
    out = []
    for i in range(400):
        L = 6000 if i % 37 == 5 else int(rng.integers(20, 200))
        s = bytes(rng.choice(list(b"ACGT"), size=L).astype(np.uint8))
        q = bytes(rng.integers(40, 80, size=L, dtype=np.uint8))
        out.append(b"@m%d\n" % i + s + b"\n+\n" + q + b"\n")
    data = b"".join(out)
    assert_equal_stats(ctx.count_bytes(data).to_dict(), O.count(data, 100), "mixed")


@pytest.mark.gpu
def test_line_density_around_the_index_capacity(ctx):
    """Tiles just below / above 1024 newlines (index capacity) alternate between the record path and the walker."""
    for width in (13, 14, 15, 16, 17, 18):
        seq = b"ACGTACGTACGTACGTAC"[:width]
        rec = b"@\n" + seq + b"\n+\n" + (b"I" * width) + b"\n"
        data = rec * (3 * 16384 // len(rec) * 8)
        assert_equal_stats(ctx.count_bytes(data).to_dict(), O.count(data, 100), f"width={width}")


@pytest.mark.gpu
def test_cr_at_every_group_and_tile_offset(ctx):
    """CRLF lines whose '\\r' falls on every offset of a 16-byte group, including the last byte of a tile."""
    rec = b"@h\r\n" + b"ACGTN" * 9 + b"\r\n+\r\n" + b"FFFFF" * 9 + b"\r\n"   # 103 bytes: walks through all alignments
    for pad in (0, 1, 5):
        data = b"x" * pad + b"\n" + rec * 2000
        assert_equal_stats(ctx.count_bytes(data).to_dict(), O.count(data, 100), f"pad={pad}")
    # a lone '\r' inside lines and at line ends without '\n' right after
    data = (b"@h\n" + b"AC\rGT" * 7 + b"\n+\n" + b"II\rII" * 7 + b"\r\n") * 3000
    assert_equal_stats(ctx.count_bytes(data).to_dict(), O.count(data, 100), "inner cr")


QUAL_FIELDS = ("qual_counts", "qual_len_hist", "qual_pos_sum", "qual_pos_cnt", "qual_lines", "qual_len_min", "qual_len_max")


@pytest.mark.gpu
def test_core_only_mode():
    """FQGPU_F_CORE_ONLY (what `sc fq-count` needs): every sequence-line output equals the oracle's, every
    quality-line output is zero, on the same inputs as the full mode (edge corpus, CRLF, long lines, streaming)."""
    rng = np.random.default_rng(21)
    cases = dict(corpus.edge_cases())
    cases["random"] = corpus.random_fastq(rng, 4000, min_len=1, max_len=400)
    cases["crlf"] = corpus.random_fastq(rng, 3000, min_len=20, max_len=260, crlf=True, final_newline=False)
    cases["long"] = b"".join(b"@r\n" + b"ACGTN" * n + b"\n+\n" + b"I" * (5 * n) + b"\n" for n in (1, 300, 7000, 40, 3300))
    cases["dense"] = (b"@\nAC\n+\nII\n") * 30000
    with fq.FqGpu(meta_records=100, flags=fq.F_CORE_ONLY) as c, fq.FqGpu(meta_records=100, flags=fq.F_CORE_ONLY, chunk_bytes=4096, n_buffers=2) as cs:
        for name, data in cases.items():
            want = O.count(data, 100)
            for k in QUAL_FIELDS:
                want[k] = [0] * len(want[k]) if isinstance(want[k], list) else 0
            assert_equal_stats(c.count_bytes(data).to_dict(), want, f"core {name}")
            cs.reset()
            cs.submit_bytes(data)
            assert_equal_stats(cs.finish().to_dict(), want, f"core streamed {name}")


@pytest.mark.gpu
def test_lines_longer_than_two_gigabytes():
    """One record whose sequence and quality lines have 2.2e9 bytes each (closed-form expectations, no oracle):
    the 64-bit carried-line paths, saturated positions and the overflow bin, in one device scan and streamed
    in two device calls that cut both lines."""
    torch = _torch()
    L = 2_200_000_011
    head, mid, tail = b"@x\n", b"\n+\n", b"\n"
    n = len(head) + L + len(mid) + L + len(tail)
    buf = torch.empty(n, dtype=torch.uint8, device="cuda")
    o = 0
    for part in (head, ord("G"), mid, ord("5"), tail):
        if isinstance(part, int):
            buf[o:o + L] = part
            o += L
        else:
            buf[o:o + len(part)] = torch.frombuffer(bytearray(part), dtype=torch.uint8).cuda()
            o += len(part)

    def check(d):
        assert d["bytes"] == n and d["lines"] == 4 and d["reads"] == 1
        assert d["bases"] == L and d["gc_bases"] == L and d["n_bases"] == 0
        assert d["base_counts"][ord("G")] == L and sum(d["base_counts"]) == L
        assert d["qual_counts"][ord("5")] == L and sum(d["qual_counts"]) == L
        assert d["seq_len_min"] == d["seq_len_max"] == L and d["qual_len_min"] == d["qual_len_max"] == L
        assert d["seq_len_hist"][512] == 1 and sum(d["seq_len_hist"]) == 1 and d["qual_len_hist"][512] == 1
        assert d["seq_len_log2"][32] == 1 and sum(d["seq_len_log2"]) == 1          # 2^31 <= L < 2^32
        assert d["qual_pos_sum"][:512] == [ord("5")] * 512 and d["qual_pos_sum"][512] == ord("5") * (L - 512)
        assert d["qual_pos_cnt"][:512] == [1] * 512 and d["qual_pos_cnt"][512] == L - 512

    with fq.FqGpu(meta_records=0) as c:
        check(c.count_device(buf.data_ptr(), n).to_dict())
        c.reset()
        cut = len(head) + L // 3 + 5
        c.scan_device(buf.data_ptr(), cut)
        cut2 = len(head) + L + len(mid) + L // 2 + 1
        c.scan_device(buf.data_ptr() + cut, cut2 - cut)
        c.scan_device(buf.data_ptr() + cut2, n - cut2)
        check(c.finish().to_dict())


@pytest.mark.gpu
def test_meta_fold_parallel_and_serial_kernels():
    """fq-meta quality range (src/fq_meta.nim:94-102,226-248) through both prefix kernels (the one-CTA kernel up to
    1024 sampled records, the single-warp kernel beyond), on bytes outside the table, CR / CRLF, empty quality
    lines and chunk cuts inside the sampled prefix."""
    rng = np.random.default_rng(33)
    body = bytearray(corpus.random_fastq(rng, 2600, min_len=0, max_len=90, qual_lo=33, qual_hi=126))
    for pos in rng.integers(0, len(body), size=400):  # sprinkle bytes outside the table and stray CRs
        if body[pos] != 0x0A:
            body[pos] = int(rng.choice([9, 13, 31, 32, 127, 200]))
    datas = {"weird": bytes(body),
             "crlf": corpus.random_fastq(rng, 1500, min_len=1, max_len=60, crlf=True, final_newline=False, qual_lo=40, qual_hi=110),
             "long": b"".join(b"@r\n" + b"A" * L + b"\n+\n" + bytes(rng.integers(35, 100, size=L, dtype=np.uint8)) + b"\n"
                              for L in (40000, 3, 70001, 1))}
    for name, data in datas.items():
        for n in (1, 2, 7, 100, 1024, 1025, 5000):
            want = O.count(data, n)
            with fq.FqGpu(meta_records=n) as c:
                assert_equal_stats(c.count_bytes(data).to_dict(), want, f"{name} n={n}")
            with fq.FqGpu(meta_records=n, chunk_bytes=4096, n_buffers=2) as c:
                c.reset()
                c.submit_bytes(data)
                assert_equal_stats(c.finish().to_dict(), want, f"{name} n={n} streamed")


def test_many_tiles_whole_and_split(ctx):
    """Inputs of several hundred tiles (every persistent CTA takes several, the chained prefix runs long): whole, and
    as two launches cut at an awkward offset."""
    torch = _torch()
    rng = np.random.default_rng(91)
    cases = {"lf": corpus.random_fastq(rng, 120000, min_len=30, max_len=260),
             "crlf_open": corpus.random_fastq(rng, 60000, min_len=1, max_len=400, crlf=True, final_newline=False),
             "long": b"".join(b"@r\n" + b"ACGTN" * 40000 + b"\n+\n" + b"IIIII" * 40000 + b"\n" for _ in range(40))}
    cases.update({k: v for k, v in corpus.edge_cases().items() if k in ("blank_lines", "qual_starts_with_at", "high_bytes", "long_line_300k")})
    for name, data in cases.items():
        want = O.count(data, 100)
        assert_equal_stats(ctx.count_bytes(data).to_dict(), want, name)
        buf = torch.frombuffer(bytearray(data) + bytearray(64), dtype=torch.uint8).cuda()
        cut = len(data) // 3
        ctx.reset(); ctx.scan_device(buf.data_ptr(), cut); ctx.scan_device(buf.data_ptr() + cut, len(data) - cut)
        assert_equal_stats(ctx.finish().to_dict(), want, name + " split")
