"""GPU parity of the single-launch scan kernel (csrc/fq_scan.cu) on the structures it treats specially: 32 KiB tiles
handed out by a ticket counter, the chained prefix over the tiles, groups with and without newlines, the item queue,
the byte walker for dense tiles, '\\r' bytes whose newline lies in another tile or launch.  Everything is compared with
the oracle bit for bit, in both modes (full statistics, FQGPU_F_CORE_ONLY)."""
import numpy as np
import pytest

import seq_collection_b200 as fq
from oracle import fq_oracle as O
from tests import corpus
from tests.test_gpu_parity import assert_equal_stats, QUAL_FIELDS

pytestmark = pytest.mark.gpu
TILE = 32768


@pytest.fixture(scope="module")
def full():
    with fq.FqGpu(meta_records=100) as c:
        yield c


@pytest.fixture(scope="module")
def core():
    with fq.FqGpu(meta_records=100, flags=fq.F_CORE_ONLY) as c:
        yield c


def check_both(full, core, data, msg):
    want = O.count(data, 100)
    assert_equal_stats(full.count_bytes(data).to_dict(), want, msg)
    got = core.count_bytes(data).to_dict()
    for k, v in want.items():
        if k in QUAL_FIELDS:
            continue
        assert np.array_equal(np.asarray(got[k]), np.asarray(v)), f"{msg} core-only field {k}: {got[k]} != {v}"


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


@pytest.mark.parametrize("seed", range(32))
def test_byte_soup(seed, full, core):
    """Differential fuzz against the oracle, whole and split in two launches at a random offset."""
    import torch

    rng = np.random.default_rng(1000 + seed)
    data = _soup(rng, int(rng.integers(1, 400_000)))
    want = O.count(data, 100)
    check_both(full, core, data, f"seed={seed} whole")
    buf = torch.frombuffer(bytearray(data) + bytearray(64), dtype=torch.uint8).cuda()
    cut = int(rng.integers(0, len(data) + 1))
    full.reset()
    full.scan_device(buf.data_ptr(), cut)
    full.scan_device(buf.data_ptr() + cut, len(data) - cut)
    assert_equal_stats(full.finish().to_dict(), want, f"seed={seed} cut={cut}")


@pytest.mark.parametrize("read_len", [20, 25, 36, 50, 75, 100, 151, 250])
def test_short_and_common_read_lengths(read_len, full, core):
    """Fixed-length reads from 20 bp (more newlines per tile than the item queue holds: byte walker) to 250 bp."""
    rng = np.random.default_rng(read_len)
    n = (6 * TILE) // (2 * read_len + 30) + 17
    recs = []
    for i in range(n):
        s = bytes(rng.choice(list(b"ACGTN"), size=read_len).astype(np.uint8))
        q = bytes(rng.integers(33, 75, size=read_len, dtype=np.uint8))
        recs.append(b"@r%d/1\n" % i + s + b"\n+\n" + q + b"\n")
    check_both(full, core, b"".join(recs), f"read_len={read_len}")


def test_crlf_pairs_across_tile_and_launch_edges(full, core):
    """A '\\r\\n' pair placed so that the '\\r' is the last byte of a tile (another CTA counts it, the newline's tile
    takes it back), the last byte of a launch (left to the next launch), or inside a tile; for every line class."""
    import torch

    rec = b"@hdr\r\n" + b"ACGTNACGTN" * 3 + b"\r\n+\r\n" + b"IIIIIFFFFF" * 3 + b"\r\n"
    assert len(rec) == 73
    body = rec * (3 * TILE // len(rec) + 3)
    for cls_off in (5, 37, 40, 72):          # offset of each '\n' inside a record
        k = (TILE - cls_off) // len(rec) - 1  # the k-th record's newline lands on the tile boundary after padding
        pad = TILE - (k * len(rec) + cls_off)
        # pad is prepended as a header-line prefix of an extra first record
        first = b"@" + b"p" * (pad - 1)
        data = first + body
        assert data[TILE] == 0x0A and data[TILE - 1] == 0x0D
        check_both(full, core, data, f"newline at tile start, class offset {cls_off}")
        want = O.count(data, 100)
        buf = torch.frombuffer(bytearray(data) + bytearray(64), dtype=torch.uint8).cuda()
        for cut in (TILE, TILE - 1, TILE + 1, 2 * TILE, len(data) - 1):
            full.reset()
            full.scan_device(buf.data_ptr(), cut)
            full.scan_device(buf.data_ptr() + cut, len(data) - cut)
            assert_equal_stats(full.finish().to_dict(), want, f"class offset {cls_off} cut={cut}")


def test_split_into_many_launches_at_odd_sizes(full):
    """One stream as dozens of launches of 8 MiB + 13 -like odd sizes (here scaled down): lines straddle every edge."""
    import torch

    rng = np.random.default_rng(5)
    data = corpus.random_fastq(rng, 20000, min_len=20, max_len=3000, crlf=False)
    want = O.count(data, 100)
    buf = torch.frombuffer(bytearray(data) + bytearray(64), dtype=torch.uint8).cuda()
    for step in (TILE + 13, 3 * TILE - 1, 100003, 7):
        if step == 7 and len(data) > 200000:
            lim = 20000  # tiny launches: only a prefix, the rest in one
        else:
            lim = len(data)
        full.reset()
        off = 0
        while off < lim:
            n = min(step, len(data) - off)
            full.scan_device(buf.data_ptr() + off, n)
            off += n
        if off < len(data):
            full.scan_device(buf.data_ptr() + off, len(data) - off)
        assert_equal_stats(full.finish().to_dict(), want, f"step={step}")


def test_dense_and_sparse_tiles_alternate(full, core):
    """Tiles of nothing but newlines / one-byte lines (walker) between ordinary tiles and tiles without any newline."""
    rng = np.random.default_rng(8)
    normal = corpus.random_fastq(rng, 400, min_len=80, max_len=160)
    data = normal + b"\n" * (TILE + 100) + normal + b"A\n" * TILE + b"G" * (2 * TILE + 5) + b"\n" + normal + b"\r\n" * 9000 + normal
    check_both(full, core, data, "dense/sparse")


def test_repeated_use_of_one_context_many_epochs(full):
    """More than 255 launches on one context: the look-back epochs wrap and the state words are cleared."""
    import torch

    rng = np.random.default_rng(13)
    data = corpus.random_fastq(rng, 1500, min_len=50, max_len=150)
    want = O.count(data, 100)
    buf = torch.frombuffer(bytearray(data) + bytearray(64), dtype=torch.uint8).cuda()
    for i in range(300):
        st = full.count_device(buf.data_ptr(), len(data))
        if i % 50 == 0 or i > 250:
            assert_equal_stats(st.to_dict(), want, f"iteration {i}")


@pytest.mark.parametrize("min_tiles", ["1", "3"])
def test_everything_again_with_tiny_spans(min_tiles):
    """FQGPU_SPAN_MIN_TILES (read once per process, hence the subprocess) cuts small inputs into spans of one / three
    tiles: span starts found from the content, spans that run past their range to finish a line, empty spans inside
    long lines, guessed phases verified by the last CTA and wrong ones (adversarial and malformed inputs) redone."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = {**os.environ, "FQGPU_SPAN_MIN_TILES": min_tiles, "FQGPU_TINY_SPANS_CHILD": "1"}
    r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_scan_tiles.py", "tests/test_gpu_parity.py", "tests/test_gpu_shards.py",
                        "-q", "-x", "-m", "gpu", "-k", "not tiny_spans and not two_gigabytes", "-p", "no:cacheprovider"],
                       cwd=root, env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1500:]
