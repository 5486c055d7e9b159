"""fq-dedup (src/fq_dedup.nim:14-84): GPU duplicate marks and the host mirror against the oracle."""
import gzip
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import corpus
import fq_oracle as O
import seq_collection_b200 as fq

pytestmark = pytest.mark.gpu


def _with_dups(rng, n, dup_frac, crlf=False, final_newline=True):
    base = corpus.random_fastq(rng, n, min_len=1, max_len=80, crlf=crlf).split(b"\r\n@r" if crlf else b"\n@r")
    nl = b"\r\n" if crlf else b"\n"
    recs = [base[0] + nl] + [b"@r" + r + nl for r in base[1:]]
    recs[-1] = recs[-1][:-len(nl)] + nl  # every record is now terminated once
    out = []
    for r in recs:
        out.append(r)
        if rng.random() < dup_frac:
            out.append(recs[int(rng.integers(0, len(recs)))])  # an earlier or later record again
    data = b"".join(out)
    return data if final_newline else data[:-len(nl)]


def _check(c, data: bytes, tmp_path, name: str):
    want_out, want_reads, want_dups, want_keep = O.fq_dedup(data)
    keep, nrec, nlines, ndups = c.dedup_bytes(data)
    assert list(keep) == want_keep, name
    assert (nlines // 4, ndups) == (want_reads, want_dups), name
    p = os.path.join(tmp_path, name + ".fq")
    open(p, "wb").write(data)
    out, err = fq.fq_dedup(p, ctx=c)
    assert out == want_out, name
    assert "total_reads: %d\n" % want_reads in err and "duplicates %d\n" % want_dups in err, name
    assert ("No Duplicates Found" in err) == (want_dups == 0), name


def test_dedup_golden_files(golden_dir, tmp_path):
    """scripts/functional-tests.sh:86-91: four '@' lines remain of dup.fq and dup.fq.gz."""
    with fq.FqGpu(meta_records=0) as c:
        for f in ("dup.fq", "dup.fq.gz", "nodup.fq"):
            out, err = fq.fq_dedup(os.path.join(golden_dir, "fastq", f), ctx=c)
            assert out.count(b"@") == 4, f
            raw = open(os.path.join(golden_dir, "fastq", f), "rb").read()
            data = gzip.decompress(raw) if f.endswith(".gz") else raw
            assert out == O.fq_dedup(data)[0], f
        assert fq.fq_dedup(os.path.join(golden_dir, "fastq", "nodup.fq"), ctx=c)[1].startswith("No Duplicates Found")


def test_dedup_random_and_edges(tmp_path):
    rng = np.random.default_rng(44)
    cases = {
        "empty": b"", "one_line": b"@a", "partial": b"@a\nAC\n+\nII\n@a\nAC", "crlf_vs_lf": b"@a\r\nAC\r\n+\r\nII\r\n@a\nAC\n+\nII\n@a\r",
        "empty_headers": b"\nA\n+\nI\n\nC\n+\nI\n", "prefix_ids": b"@ab\nA\n+\nI\n@a\nA\n+\nI\n@abc\nA\n+\nI\n@a\nC\n+\nI\n",
        "long_header": b"@" + b"x" * 5000 + b"\nA\n+\nI\n@" + b"x" * 5000 + b"\nA\n+\nI\n@" + b"x" * 4999 + b"y\nA\n+\nI\n",
        "dups": _with_dups(rng, 3000, 0.4), "dups_crlf": _with_dups(rng, 1500, 0.5, crlf=True, final_newline=False),
        "many_same": (b"@same\nACGT\n+\nIIII\n" * 5000) + b"@other\nA\n+\nI\n",
    }
    with fq.FqGpu(meta_records=0) as c:
        for name, data in cases.items():
            _check(c, data, str(tmp_path), name)


def test_dedup_docs_benchmark_shape():
    """The shape of docs/fq-dedup.md:26-31: 2.5 M reads, 1 M+ duplicates (against the oracle's table)."""
    import torch
    n = 1_400_000
    with fq.FqGpu(meta_records=0) as c:
        buf = torch.empty(360 * n, dtype=torch.uint8, device="cuda")
        c.synth_illumina(buf.data_ptr(), 360 * n, 0, n, 7)
        host = buf.cpu().numpy()
        data = bytes(host) + bytes(host[:360 * 1_100_000])  # 2.5 M records, the last 1.1 M repeat earlier ones
        keep, nrec, nlines, ndups = c.dedup_bytes(data)
        assert (nrec, nlines, ndups) == (2_500_000, 10_000_000, 1_100_000)
        assert keep == b"\x01" * n + b"\x00" * 1_100_000
