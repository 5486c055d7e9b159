"""The `sc` mirror CLI (seq-collection_b200/csrc/sc_main.cpp): flags, header strings, exit codes and
error texts of the reference (sc.nim:64-79,103-116; helpers.nim:29-34,200-224) on CPU, and -- on the GPU --
the printed rows against the reference's golden tables and its functional tests
(scripts/functional-tests.sh:94-166 pin fq-meta columns 2-3)."""
import os
import subprocess

import pytest

from oracle import fq_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SC = os.path.join(ROOT, "seq-collection_b200", "sc")
FQ = os.path.join(ROOT, "tests", "golden", "fastq")


@pytest.fixture(scope="module", autouse=True)
def build_cli():
    from importlib import import_module

    b = import_module("seq-collection_b200.build")
    b.build_lib()
    b.build_cli()
    assert os.path.exists(SC)


def run(*args, cwd=None):
    p = subprocess.run([SC, *args], capture_output=True, text=True, cwd=cwd)
    return p.returncode, p.stdout, p.stderr


def test_fq_count_header_variants():
    assert run("fq-count", "--header") == (0, "reads\tgc_content\tgc_bases\tn_bases\tbases\n", "")
    assert run("fq-count", "-t", "-b")[1] == "reads\tgc_content\tgc_bases\tn_bases\tbases\tbasename\n"
    assert run("fq-count", "-t", "-b", "-a")[1] == "reads\tgc_content\tgc_bases\tn_bases\tbases\tbasename\tabsolute\n"


def test_fq_meta_header():
    rc, out, _ = run("fq-meta", "-t")
    assert rc == 0
    assert out.rstrip("\n").split("\t") == ["machine", "sequencer", "prob_sequencer", "flowcell", "flowcell_description", "run",
                                           "lane", "sequence_id", "index1", "index2", "qual_format", "qual_phred",
                                           "qual_multiple", "min_qual", "max_qual", "n_lines"]


def test_no_fastq_is_exit_3():
    rc, out, err = run("fq-count", "-b")
    assert rc == 3 and out == "" and "Error 3: No FASTQ specified" in err


def test_help_when_no_arguments():
    rc, out, _ = run()
    assert rc == 0 and "Sequence data utilities (Version 0.0.2)" in out and "fq-count" in out


@pytest.mark.gpu
def test_fq_count_rows_match_docs_table(golden_dir):
    names = [ln.split("\t")[0] for ln in open(os.path.join(golden_dir, "fq_count_docs.tsv")).read().splitlines()[1:]]
    rc, out, err = run("fq-count", "--header", "-b", *[os.path.join(FQ, n) for n in names])
    assert rc == 0, err
    lines = out.splitlines()
    assert lines[0] == "reads\tgc_content\tgc_bases\tn_bases\tbases\tbasename"
    for name, ln in zip(names, lines[1:]):
        want = O.fq_count_row(O.count_file(os.path.join(FQ, name))) + "\t" + name
        assert ln == want
    assert lines[1 + names.index("illumina_8.fq")] == "2\t0.3333333333333333\t14\t0\t42\tillumina_8.fq"
    assert lines[1 + names.index("novaseq.fq")] == "9\t0.0\t0\t0\t9\tnovaseq.fq"


@pytest.mark.gpu
def test_fq_count_errors():
    rc, out, err = run("fq-count", "/nonexistent/reads.fq")
    assert rc == 2 and out == "" and "Error 2: Unable to open file: /nonexistent/reads.fq" in err
    rc, out, err = run("fq-count", "/nonexistent/reads.fq.gz")
    assert rc == 2 and "Unable to open file" in err
    rc, _, _ = run("fq-count", "ab")  # fastq[^3 .. ^1] raises in the reference
    assert rc == 1


@pytest.mark.gpu
def test_fq_count_absolute_and_relative(tmp_path):
    import shutil

    shutil.copy(os.path.join(FQ, "dup.fq"), tmp_path / "x.fq")
    rc, out, _ = run("fq-count", "-b", "-a", "x.fq", cwd=str(tmp_path))
    assert rc == 0
    assert out == f"8\t0.53125\t17\t0\t32\tx.fq\t{tmp_path}/x.fq\n"


@pytest.mark.gpu
@pytest.mark.parametrize("name,sequencer,prob", [
    ("illumina_1.fq", "GenomeAnalyzerIIx", "likely:machine"),
    ("illumina_2.fq", "GenomeAnalyzerIIx", "likely:machine"),
    ("illumina_3.fq", "", ""),
    ("illumina_4.fq", "", ""),
    ("illumina_2000_2500.fq", "HiSeq2000/2500", "high:machine+flowcell"),
    ("illumina_3000_4000.fq", "HiSeq3000/4000", "high:machine+flowcell"),
    ("illumina_hiseq_x.fq", "HiSeqX", "high:machine+flowcell"),
    ("novaseq.fq", "NovaSeq", "high:machine+flowcell"),
])
def test_fq_meta_functional_tests(name, sequencer, prob):
    """scripts/functional-tests.sh:116-166: columns 2 (sequencer) and 3 (confidence)."""
    rc, out, err = run("fq-meta", os.path.join(FQ, name))
    assert rc == 0, err
    cols = out.rstrip("\n").split("\t")
    assert len(cols) == 16
    assert cols[1] == sequencer
    assert cols[2].split(":")[0] == prob.split(":")[0]
    if prob.startswith("high"):
        assert cols[2] == prob


@pytest.mark.gpu
def test_fq_meta_docs_rows(golden_dir):
    """docs/fq-meta.md:34-37 -- every column of the four example rows except the path."""
    want = {
        "illumina_2000_2500.fq": ["D00446", "HiSeq2000/2500", "high:machine+flowcell", "C8HN4ANXX", "High Output (8-lane) v4 flow cell", "1", "8", "", "GCTCGGTA", "", "Sanger;Illumina 1.8+", "Phred+33", "true", "14", "14", "1"],
        "illumina_3000_4000.fq": ["K00100", "HiSeq3000/4000", "high:machine+flowcell", "H300JBBXX", "(8-lane) v1 flow cell", "33", "6", "", "GCCAAT", "", "Sanger;Illumina 1.8+", "Phred+33", "true", "14", "14", "1"],
        "illumina_6.fq": ["D00209", "HiSeq2000/2500", "high:machine+flowcell", "CACDKANXX", "High Output (8-lane) v4 flow cell", "258", "6", "", "CGCAGTT", "", "Sanger;Illumina 1.8+", "Phred+33", "true", "0", "37", "1"],
        "illumina_7.fq": ["D00209", "HiSeq2000/2500", "high:machine+flowcell", "CACDKANXX", "High Output (8-lane) v4 flow cell", "258", "6", "", "GAGCAAG", "", "Sanger;Illumina 1.8+", "Phred+33", "true", "0", "37", "1"],
    }
    for name, cols in want.items():
        rc, out, err = run("fq-meta", "-b", os.path.join(FQ, name))
        assert rc == 0, err
        assert out.rstrip("\n").split("\t") == cols + [name], name


@pytest.mark.gpu
def test_fq_meta_dual_index_barcode(tmp_path):
    """src/fq_meta.nim:210 `[ATCGN\\+\\-]{3,12}+`: the trailing '+' repeats the bounded group, so a dual-index barcode
    (8 + 1 + 8 characters) is reported whole in the index1 column."""
    f = tmp_path / "dual.fq"
    rec = b"@A00156:217:HKJWGDSXX:1:1101:1000:1000 1:N:0:AACGCTTA+GGTTCAGT\nACGT\n+\nFFFF\n"
    f.write_bytes(rec * 3)
    rc, out, err = run("fq-meta", str(f))
    assert rc == 0, err
    assert out.rstrip("\n").split("\t")[8] == "AACGCTTA+GGTTCAGT"
    g = tmp_path / "short.fq"
    g.write_bytes(rec.replace(b"AACGCTTA+GGTTCAGT", b"AC") * 2)  # shorter than 3: no barcode
    rc, out, err = run("fq-meta", str(g))
    assert rc == 0 and out.rstrip("\n").split("\t")[8] == ""


@pytest.mark.gpu
def test_fq_meta_n_option_and_gz_case():
    rc, out, _ = run("fq-meta", "-n", "2", os.path.join(FQ, "illumina_3.fq"))
    assert rc == 0 and out.rstrip("\n").split("\t")[15] == "2"
    rc, out, _ = run("fq-meta", os.path.join(FQ, "dup.fq.gz"))
    assert rc == 0 and out.rstrip("\n").split("\t")[13:16] == ["32", "41", "8"]


@pytest.mark.gpu
def test_fq_dedup_functional_tests(golden_dir, tmp_path):
    """scripts/functional-tests.sh:86-91 (`grep -c '@'` == 4 for dup.fq and dup.fq.gz), the whole stdout against the
    oracle's restatement of src/fq_dedup.nim, the stderr summary (:76-83) and the exit code for a missing file."""
    import gzip
    from oracle import fq_oracle as O
    for f in ("dup.fq", "dup.fq.gz", "nodup.fq"):
        path = os.path.join(golden_dir, "fastq", f)
        raw = open(path, "rb").read()
        want, n_reads, n_dups, _ = O.fq_dedup(gzip.decompress(raw) if f.endswith(".gz") else raw)
        p = subprocess.run([SC, "fq-dedup", path], capture_output=True)
        assert p.returncode == 0, p.stderr
        assert p.stdout == want and p.stdout.count(b"@") == 4, f
        err = p.stderr.decode()
        assert "total_reads: %d\n" % n_reads in err and "duplicates %d\n" % n_dups in err and "false-positive: 0\n" in err
        assert ("No Duplicates Found\nCopying fq to stdout\n" in err) == (n_dups == 0)
        assert err.rstrip().endswith("false-positive-rate: " + ("nan" if n_dups == 0 else "0.0"))
    crlf = tmp_path / "crlf.fq"
    crlf.write_bytes(b"@a\r\nAC\r\n+\r\nII\r\n@a\nGG\n+\nII\n@b\nT")
    p = subprocess.run([SC, "fq-dedup", str(crlf)], capture_output=True)
    assert p.stdout == b"@a\nAC\n+\nII\n@b\nT\n" and b"duplicates 1\n" in p.stderr
    code, out, err = run("fq-dedup", str(tmp_path / "missing.fq"))
    assert code == 2 and "Unable to open file" in err
