"""Pins the CPU oracle on every golden vector the reference holds for this path and
checks its two independent restatements (C, pure Python) against each other on the edge corpus."""
import os

import numpy as np
import pytest

from oracle import fq_oracle as O
from tests import corpus


def _rows(path):
    lines = open(path).read().splitlines()
    head = lines[0].split("\t")
    return [dict(zip(head, ln.split("\t"))) for ln in lines[1:]]


def _sig6(x: float) -> str:
    return "%.6g" % x


def test_fq_count_docs_table(golden_dir):
    """docs/fq-count.md:29-43 -- all 15 rows, integers exact, gc_content equal at the 6 digits the docs print."""
    rows = _rows(os.path.join(golden_dir, "fq_count_docs.tsv"))
    assert len(rows) == 15
    for r in rows:
        d = O.count_file(os.path.join(golden_dir, "fastq", r["file"]))
        assert d["reads"] == int(r["reads"]), r["file"]
        assert d["gc_bases"] == int(r["gc_bases"]), r["file"]
        assert d["n_bases"] == int(r["n_bases"]), r["file"]
        assert d["bases"] == int(r["bases"]), r["file"]
        gc = d["gc_bases"] / (d["bases"] - d["n_bases"])
        assert _sig6(gc) == _sig6(float(r["gc_content"])), r["file"]


def test_fq_count_row_format(golden_dir):
    """`$float` of Nim 1.0.6: 0.3333333333333333, 1.0, 0.0, nan."""
    d = O.count_file(os.path.join(golden_dir, "fastq", "illumina_8.fq"))
    assert O.fq_count_row(d) == "2\t0.3333333333333333\t14\t0\t42"
    d = O.count_file(os.path.join(golden_dir, "fastq", "illumina_2000_2500.fq"))
    assert O.fq_count_row(d) == "1\t1.0\t101\t0\t101"
    d = O.count_file(os.path.join(golden_dir, "fastq", "novaseq.fq"))
    assert O.fq_count_row(d) == "9\t0.0\t0\t0\t9"
    d = O.count_file(os.path.join(golden_dir, "fastq", "sra.fq"))
    assert O.fq_count_row(d) == "2\t0.4305555555555556\t62\t0\t144"
    assert O.fq_count_row(O.count(b"")) == "0\tnan\t0\t0\t0"
    assert O.format_float(0.35) == "0.35"
    assert O.format_float(1e22) == "1e+22"


def test_fq_meta_docs_table(golden_dir):
    """docs/fq-meta.md:34-37 -- min_qual/max_qual/n_lines and the encoding guess at the CLI default -n 100."""
    rows = _rows(os.path.join(golden_dir, "fq_meta_docs.tsv"))
    assert len(rows) == 4
    for r in rows:
        d = O.count_file(os.path.join(golden_dir, "fastq", r["file"]), meta_records=100)
        f = O.fq_meta_quality_fields(d)
        assert f[0] == r["qual_format"], r["file"]
        assert f[1] == r["qual_phred"], r["file"]
        assert f[2].upper() == r["qual_multiple"], r["file"]
        assert f[3] == r["min_qual"] and f[4] == r["max_qual"], r["file"]
        assert f[5] == r["n_lines"], r["file"]


def test_survey_min_max_column(golden_dir):
    """min_q/max_q of all 15 fixtures as tabulated in SURVEY 8c (restatement run during the survey)."""
    expect = {"dup.fq": (32, 41), "dup.fq.gz": (32, 41), "illumina_1.fq": (0, 37), "illumina_2.fq": (0, 37),
              "illumina_2000_2500.fq": (14, 14), "illumina_3.fq": (0, 37), "illumina_3000_4000.fq": (14, 14),
              "illumina_4.fq": (0, 37), "illumina_6.fq": (0, 37), "illumina_7.fq": (0, 37), "illumina_8.fq": (65, 93),
              "illumina_hiseq_x.fq": (0, 37), "nodup.fq": (32, 41), "novaseq.fq": (2, 37), "sra.fq": (8, 40)}
    for name, (lo, hi) in expect.items():
        d = O.count_file(os.path.join(golden_dir, "fastq", name), meta_records=100)
        assert (d["meta_qual_min"], d["meta_qual_max"]) == (lo, hi), name


def test_gz_equals_plain(golden_dir):
    a = O.count_file(os.path.join(golden_dir, "fastq", "dup.fq"), meta_records=20)
    b = O.count_file(os.path.join(golden_dir, "fastq", "dup.fq.gz"), meta_records=20)
    assert a == b


def test_missing_file():
    with pytest.raises(OSError):
        O.count_file("/nonexistent/x.fq")


@pytest.mark.parametrize("name", sorted(corpus.edge_cases()))
def test_c_oracle_equals_python_twin(name):
    data = corpus.edge_cases()[name]
    for n in (0, 1, 2, 100):
        assert O.count(data, n) == O.py_count(data, n), (name, n)


@pytest.mark.parametrize("name", sorted(corpus.edge_cases()))
def test_oracle_chunking_invariance(name):
    data = corpus.edge_cases()[name]
    whole = O.count(data, 100)
    for chunk in (1, 2, 3, 7, 64, 4097):
        assert O.count(data, 100, chunk=chunk) == whole, (name, chunk)


def test_golden_files_python_twin(golden_dir):
    for name in sorted(os.listdir(os.path.join(golden_dir, "fastq"))):
        if name.endswith(".gz"):
            continue
        data = open(os.path.join(golden_dir, "fastq", name), "rb").read()
        assert O.count(data, 100) == O.py_count(data, 100), name


def test_random_fastq_twins():
    rng = np.random.default_rng(7)
    for k in range(6):
        data = corpus.random_fastq(rng, 40, crlf=bool(k & 1), final_newline=bool(k & 2))
        assert O.count(data, 13) == O.py_count(data, 13)


def test_ref_work_shape_matches():
    rng = np.random.default_rng(11)
    data = corpus.random_fastq(rng, 500, final_newline=False)
    d = O.count(data)
    r = O.ref_fq_count_mem(data)
    for k in ("reads", "gc_bases", "n_bases", "bases", "lines"):
        assert r[k] == d[k]


def test_ref_work_shape_file(tmp_path):
    for name, data in corpus.edge_cases().items():
        if b"\x00" in data or name in ("trailing_cr_no_lf",):
            continue  # NUL handling of fgets-based readLine is unpinned (see fq_oracle.c header)
        p = tmp_path / "x.fq"
        p.write_bytes(data)
        d = O.count(data)
        r = O.ref_fq_count_file(str(p))
        for k in ("reads", "gc_bases", "n_bases", "bases", "lines"):
            assert r[k] == d[k], (name, k)


def test_record_offsets_on_golden_files(golden_dir):
    """The index checker: one offset per read of docs/fq-count.md, each record of the well-formed fixtures starts
    with '@' (novaseq.fq is malformed on purpose: mod-4 classing, no '@' check), the brute-force definition agrees."""
    rows = _rows(os.path.join(golden_dir, "fq_count_docs.tsv"))
    for r in rows:
        if r["file"].endswith(".gz"):
            continue
        data = open(os.path.join(golden_dir, "fastq", r["file"]), "rb").read()
        offs = O.record_offsets(data)
        assert len(offs) == int(r["reads"]), r["file"]
        lines = data.split(b"\n")
        if lines and lines[-1] == b"":
            lines.pop()
        brute, pos = [], 0
        for i, ln in enumerate(lines):
            if i % 4 == 0:
                brute.append(pos)
            pos += len(ln) + 1
        assert offs.tolist() == brute, r["file"]
        heads = O.header_lines(data, len(offs))
        assert heads == [ln[:256] for ln in lines[0::4]], r["file"]  # (no CR in the fixtures)
        if r["file"] != "novaseq.fq":
            assert all(h.startswith(b"@") for h in heads), r["file"]
    for name, data in corpus.edge_cases().items():
        offs = O.record_offsets(data)
        assert len(offs) == (O.count(data, 0)["reads"] if data else 0), name


def test_fq_dedup_oracle_on_golden(golden_dir):
    """scripts/functional-tests.sh:86-91 pins `grep -c '@'` == 4 for dup.fq / dup.fq.gz."""
    import gzip
    for f in ("dup.fq", "dup.fq.gz"):
        raw = open(os.path.join(golden_dir, "fastq", f), "rb").read()
        out, n_reads, n_dups, keep = O.fq_dedup(gzip.decompress(raw) if f.endswith(".gz") else raw)
        assert out.count(b"@") == 4 and (n_reads, n_dups) == (8, 4) and keep == [1, 1, 0, 1, 0, 1, 0, 0]
    out, n_reads, n_dups, keep = O.fq_dedup(open(os.path.join(golden_dir, "fastq", "nodup.fq"), "rb").read())
    assert (n_reads, n_dups) == (4, 0) and keep == [1, 1, 1, 1]
