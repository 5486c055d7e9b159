"""The fq-meta quality-range fold (csrc/fq_meta.cu) when the sample is large: segments of 64 KiB are folded in parallel
where the fold is a plain min/max, the sequential walk takes over where it is not (an empty quality line, a byte
outside the table, a lone '\\r', the segment in which the sample ends).  Whatever the mix, the result must be the
oracle's (src/fq_meta.nim:94-102,226-248 restated in oracle/), for every sample size and every cut of the stream."""
import numpy as np
import pytest

import seq_collection_b200 as fq
from oracle import fq_oracle as O
from tests import corpus
from tests.test_gpu_parity import assert_equal_stats

pytestmark = pytest.mark.gpu

SEG = 65536


def run(data: bytes, meta_records: int, cuts=()):
    """The stream through the pinned ring in pieces cut at `cuts` (a launch per piece)."""
    with fq.FqGpu(meta_records=meta_records) as c:
        c.reset()
        last = 0
        for cut in list(cuts) + [len(data)]:
            if cut > last:
                c.submit_bytes(data[last:cut])  # one launch per piece
                last = cut
        return c.finish().to_dict()


def check(data: bytes, meta_records: int, cuts=(), msg=""):
    got = run(data, meta_records, cuts)
    want = O.count(data, meta_records)
    assert_equal_stats(got, want, f"{msg} n={meta_records} cuts={list(cuts)[:4]}")


@pytest.mark.parametrize("crlf", [False, True])
def test_sample_sizes_around_segment_edges(crlf):
    rng = np.random.default_rng(11 + crlf)
    data = corpus.random_fastq(rng, 3000, min_len=50, max_len=250, crlf=crlf)  # ~ 1 MB: 15 segments
    n_lines = data.count(b"\n")
    for n in (1, 7, 100, 250, 1000, 2999, 3000, 3001, 10**9):
        check(data, n, msg=f"crlf={crlf}")
    # the sample ends exactly at / around the line that crosses a segment edge
    pos = 3 * SEG
    lines_before = data[:pos].count(b"\n")
    for n in {max(1, lines_before // 4 - 1), lines_before // 4, lines_before // 4 + 1}:
        check(data, n, msg="edge sample")
    assert n_lines == 12000


def test_cuts_everywhere():
    rng = np.random.default_rng(5)
    data = corpus.random_fastq(rng, 2500, min_len=30, max_len=200, crlf=True, final_newline=False)
    for n in (10**9, 1200):
        for cut in (1, 15, 16, 17, SEG - 1, SEG, SEG + 1, 2 * SEG + 3, len(data) - 1):
            check(data, n, cuts=[cut], msg="one cut")
        # a cut between every '\r' and '\n' of some line ends, and right behind newlines
        ends = [i for i in range(len(data)) if data[i:i + 1] == b"\n"][40:4000:397]
        check(data, n, cuts=ends, msg="cuts before newlines")
        check(data, n, cuts=[e + 1 for e in ends], msg="cuts behind newlines")
        check(data, n, cuts=[e - 1 for e in ends], msg="cuts before the carriage returns")


def _with_line(data: bytes, which: int, new: bytes) -> bytes:
    lines = data.split(b"\n")
    lines[which] = new
    return b"\n".join(lines)


def test_lines_the_parallel_fold_must_not_vouch_for():
    rng = np.random.default_rng(9)
    base = corpus.random_fastq(rng, 2000, min_len=80, max_len=120, qual_lo=40, qual_hi=70)
    nl = base.count(b"\n")
    cases = {}
    for rec in (0, 1, 700, 1999):
        q = 4 * rec + 3
        cases[f"empty quality line {rec}"] = _with_line(base, q, b"")
        cases[f"byte outside the table {rec}"] = _with_line(base, q, b"II\x1fII")
        cases[f"lone carriage return {rec}"] = _with_line(base, q, b"II\rII")
        cases[f"only a carriage return {rec}"] = _with_line(base, q, b"\r")
        cases[f"high byte {rec}"] = _with_line(base, q, b"II\xffII")
        cases[f"empty sequence line {rec}"] = _with_line(base, q - 2, b"")
    cases["negative first, then replaced"] = _with_line(_with_line(base, 3, b" !"), 7, b"5")
    cases["starts with newlines"] = b"\n\n\n" + base
    cases["starts with crlf"] = b"\r\n" + base
    cases["ends with a carriage return"] = base[:-1] + b"\r"
    cases["ends inside a quality line"] = base[:-30]
    for name, data in cases.items():
        for n in (10**9, 1000, nl // 4):
            check(data, n, msg=name)
            check(data, n, cuts=[SEG + 5, 3 * SEG], msg=name + " (cuts)")


def test_long_reads_span_segments():
    rng = np.random.default_rng(3)
    data = corpus.random_fastq(rng, 12, min_len=90_000, max_len=200_000)
    for n in (1, 5, 12, 100):
        check(data, n, msg="long reads")
        check(data, n, cuts=[100_000, 100_001, 777_777], msg="long reads (cuts)")


def test_edge_corpus_every_sample_size():
    for name, data in corpus.edge_cases().items():
        for n in (1, 2, 3, 100):
            check(data, n, msg=name)
