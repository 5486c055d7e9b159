"""fqgpu_count_files: many files counted concurrently (the loop of sc.nim:115-116).

Each file is an independent stream with its own context, so every result must equal the oracle's for that
file whatever the thread count, and errors must surface where the sequential loop would have stopped."""
import gzip
import os
import subprocess

import numpy as np
import pytest

import seq_collection_b200 as fq
from oracle import fq_oracle as O
from tests import corpus
from tests.test_gpu_parity import assert_equal_stats

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SC = os.path.join(ROOT, "seq-collection_b200", "sc")


@pytest.fixture(scope="module")
def files(tmp_path_factory, golden_dir):
    d = tmp_path_factory.mktemp("batch")
    rng = np.random.default_rng(21)
    out = []
    for i in range(6):
        data = corpus.random_fastq(rng, 2000 + 700 * i, min_len=20, max_len=300, crlf=(i == 3), final_newline=(i != 4))
        p = d / f"r{i}.fq"
        if i % 2:
            p = d / f"r{i}.fq.gz"
            with gzip.open(p, "wb", compresslevel=1) as f:
                f.write(data)
        else:
            p.write_bytes(data)
        out.append(str(p))
    for name in ("dup.fq.gz", "illumina_2.fq", "novaseq.fq"):
        q = os.path.join(golden_dir, "fastq", name)
        if os.path.exists(q):
            out.append(q)
    empty = d / "empty.fq"
    empty.write_bytes(b"")
    out.append(str(empty))
    return out


@pytest.mark.parametrize("threads", [0, 1, 3])
def test_count_files_equals_oracle(files, threads):
    rc, res = fq.count_files(files, n_threads=threads, meta_records=100)
    assert rc == fq.OK
    for path, (r, st) in zip(files, res):
        assert r == fq.OK, path
        assert_equal_stats(st.to_dict(), O.count_file(path, 100), f"{path} threads={threads}")


def test_count_files_reports_the_first_failure_in_file_order(files):
    paths = files[:2] + ["/nonexistent/x.fq"] + files[2:4] + ["/nonexistent/y.fq.gz"]
    rc, res = fq.count_files(paths, n_threads=4)
    assert rc == fq.EIO
    assert [r for r, _ in res] == [fq.OK, fq.OK, fq.EIO, fq.OK, fq.OK, fq.EIO]
    assert "Unable to open file: /nonexistent/x.fq" in fq.load_library().fqgpu_last_error(None).decode()
    for path, (r, st) in zip(paths, res):
        if r == fq.OK:
            assert_equal_stats(st.to_dict(), O.count_file(path, 0), path)


def test_count_files_empty_list_and_all_devices(files):
    assert fq.count_files([])[0] == fq.OK
    rc, res = fq.count_files(files[:3], device=fq.DEVICE_ALL, flags=fq.F_CORE_ONLY)
    assert rc == fq.OK
    for path, (r, st) in zip(files[:3], res):
        assert fq.fq_count_row(st) == O.fq_count_row(O.count_file(path, 0)), path


def test_fq_count_many_and_cli_rows(files):
    rows = fq.fq_count_many(files, basename=True)
    want = [fq.output_w_fnames(O.fq_count_row(O.count_file(p, 0)), p, True, False) for p in files]
    assert rows == want
    from importlib import import_module

    b = import_module("seq-collection_b200.build")
    b.build_lib()
    b.build_cli()
    par = subprocess.run([SC, "fq-count", "-t", "-b", *files], capture_output=True, text=True)
    seq = subprocess.run([SC, "fq-count", "-t", "-b", *files], capture_output=True, text=True, env={**os.environ, "FQGPU_THREADS": "1"})
    assert par.returncode == 0 and seq.returncode == 0, par.stderr + seq.stderr
    assert par.stdout == seq.stdout == "reads\tgc_content\tgc_bases\tn_bases\tbases\tbasename\n" + "\n".join(want) + "\n"
    # a missing file in the middle: the rows before it, then "Error 2" as in src/fq_count.nim:35-36
    bad = subprocess.run([SC, "fq-count", files[0], "/nonexistent/x.fq", files[1]], capture_output=True, text=True)
    assert bad.returncode == 2 and "Unable to open file: /nonexistent/x.fq" in bad.stderr
    assert bad.stdout == O.fq_count_row(O.count_file(files[0], 0)) + "\n"


def test_multithreaded_file_reads_equal_oracle(tmp_path, monkeypatch):
    """A plain file larger than the threshold of parallel_pread, several chunks per file, odd sizes: the chunk
    filled by 1, 3 and 8 reader threads gives the same stream (FQGPU_READ_THREADS is read once per process, so
    each setting runs in its own interpreter)."""
    import subprocess
    import sys

    rng = np.random.default_rng(33)
    data = corpus.random_fastq(rng, 130_000, min_len=80, max_len=200, final_newline=False)  # ~40 MB
    path = tmp_path / "big.fq"
    path.write_bytes(data)
    want = O.fq_count_row(O.count(data, 0))
    code = ("import sys; sys.path.insert(0, %r); import seq_collection_b200 as fq\n"
            "with fq.FqGpu(chunk_bytes=(16 << 20) + 4096, flags=fq.F_CORE_ONLY) as c: print(fq.fq_count_row(c.count_file(%r)))\n") % (ROOT, str(path))
    for threads in ("1", "3", "8"):
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env={**os.environ, "FQGPU_READ_THREADS": threads})
        assert r.returncode == 0, r.stderr
        assert r.stdout.strip() == want, threads


@pytest.mark.parametrize("n", [1, 7, 100, 5000])
def test_meta_file_reads_only_the_sampled_head(tmp_path, n):
    """fqgpu_meta_file_as stops where the reference's sampling loop stops (src/fq_meta.nim:226): the quality fields
    equal the oracle's for the whole file, and the bytes scanned are those of the first 4 n lines."""
    rng = np.random.default_rng(40 + n)
    cases = {
        "lf": corpus.random_fastq(rng, 900, min_len=20, max_len=120),
        "crlf_open_end": corpus.random_fastq(rng, 900, min_len=20, max_len=120, crlf=True, final_newline=False),
        "short": corpus.random_fastq(rng, 3, min_len=5, max_len=9, final_newline=False),
        "empty": b"",
    }
    with fq.FqGpu(meta_records=n) as c:
        for name, data in cases.items():
            for gz in (False, True):
                p = tmp_path / (f"{name}_{n}.fq" + (".GZ" if gz else ""))
                p.write_bytes(gzip.compress(data) if gz else data)
                st = c.meta_file(str(p))
                want = O.count(data, n)
                assert fq.fq_meta_quality_fields(st) == O.fq_meta_quality_fields(want), (name, gz)
                head = b"".join(data.splitlines(keepends=True)[: 4 * n]) if b"\r\n" not in data else None
                if head is not None:
                    assert st.bytes == len(head), (name, gz, st.bytes, len(head))
                else:
                    assert st.bytes <= len(data)


def test_count_file_sharded_equals_oracle(tmp_path):
    """One plain file over several contexts in one process (devices may repeat: runs on a single GPU too); also the
    inputs that must fall back to a single context (small, .gz, a phase hypothesis that fails)."""
    import torch

    ndev = torch.cuda.device_count()
    rng = np.random.default_rng(55)
    big = corpus.random_fastq(rng, 150_000, min_len=80, max_len=200, final_newline=False)                 # ~47 MB
    crlf = corpus.random_fastq(rng, 120_000, min_len=50, max_len=150, crlf=True)
    # quality lines that look like headers and sequence lines that look like '+' lines: resync may guess wrong
    nasty = b"".join(b"@r%d\n+ACGT%d\n+\n@III%d\n" % (i, i, i) for i in range(700_000))
    small = corpus.random_fastq(rng, 100)
    cases = {"big": big, "crlf": crlf, "nasty": nasty, "small": small}
    for name, data in cases.items():
        p = tmp_path / (name + ".fq")
        p.write_bytes(data)
        want = O.count(data, 100)
        for devices in ([0, 0], [0, 0, 0], [g % ndev for g in range(5)]):
            st = fq.count_file_sharded(str(p), devices=devices, meta_records=100, chunk_bytes=8 << 20)
            assert_equal_stats(st.to_dict(), want, f"{name} devices={devices}")
    gz = tmp_path / "big.fq.gz"
    with gzip.open(gz, "wb", compresslevel=1) as f:
        f.write(big[: 6 << 20])
    assert_equal_stats(fq.count_file_sharded(str(gz), devices=[0, 0], meta_records=100).to_dict(), O.count(big[: 6 << 20], 100), "gz")
    st = fq.count_file_sharded(str(tmp_path / "big.fq"), world=0, flags=fq.F_CORE_ONLY)  # one shard per visible device
    assert fq.fq_count_row(st) == O.fq_count_row(O.count(big, 0))
    with pytest.raises(fq.FqGpuError) as ei:
        fq.count_file_sharded("/nonexistent/file.fq", devices=[0, 0])
    assert ei.value.code == fq.EIO


def test_paired_files_as_one_job(tmp_path):
    """R1 / R2: both mates scanned at the same time, one row per file as the reference's loop over the
    two files would print them (sc.nim:115-116), and the mate check."""
    import gzip

    rng = np.random.default_rng(77)
    r1 = corpus.random_fastq(rng, 5000, min_len=100, max_len=151)
    r2 = corpus.random_fastq(rng, 5000, min_len=100, max_len=151)
    p1, p2 = tmp_path / "lib_R1.fq", tmp_path / "lib_R2.fq.gz"
    p1.write_bytes(r1)
    p2.write_bytes(gzip.compress(r2))
    rc, a, b, paired = fq.count_pair(str(p1), str(p2), meta_records=100)
    assert rc == 0 and paired
    assert_equal_stats(a.to_dict(), O.count(r1, 100), "R1")
    assert_equal_stats(b.to_dict(), O.count(r2, 100), "R2")
    assert fq.fq_count_row(a) == O.fq_count_row(O.count(r1, 100)) and fq.fq_count_row(b) == O.fq_count_row(O.count(r2, 100))
    p3 = tmp_path / "short_R2.fq"
    p3.write_bytes(corpus.random_fastq(rng, 4999, min_len=100, max_len=151))
    rc, a, b, paired = fq.count_pair(str(p1), str(p3))
    assert rc == 0 and not paired and a.reads == 5000 and b.reads == 4999
    rc, a, b, paired = fq.count_pair(str(p1), str(tmp_path / "missing.fq"))
    assert rc == fq.EIO and not paired
