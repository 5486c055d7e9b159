"""Record-offset index and header gather against the oracle."""
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


@pytest.fixture(params=["onepass", "onepass16", "onepass32", "onepass64", "2pass"], autouse=True)
def index_mode(request, monkeypatch):
    """Both index paths: the single launch with the chained tile prefix (default; tile size by input size, or each of
    the 64 / 128 / 256 KiB instantiations forced) and count -> prefix -> write."""
    monkeypatch.setenv("FQGPU_INDEX", request.param)
    if request.param[7:]:
        monkeypatch.setenv("FQGPU_INDEX_ROWS", request.param[7:])
    return request.param


def _check(c, torch, data: bytes, misalign: int = 0, n_headers: int = 25):
    want = O.record_offsets(data)
    buf = torch.zeros(len(data) + 64, dtype=torch.uint8, device="cuda")
    if data:
        buf[misalign:misalign + len(data)] = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
    offs = torch.full((len(want) + 3,), 2**62, dtype=torch.int64, device="cuda")
    n = c.index_device(buf.data_ptr() + misalign, len(data), offs.data_ptr(), offs.numel())
    assert n == len(want) == (O.count(data, 0)["reads"] if data else 0)
    got = offs.cpu().numpy().astype(np.uint64)
    assert (got[:n] == want).all() and (got[n:] == 2**62).all()
    # capacity smaller than the record count: only the first cap are written, the count is still exact
    if n > 2:
        small = torch.full((n,), 2**62, dtype=torch.int64, device="cuda")
        assert c.index_device(buf.data_ptr() + misalign, len(data), small.data_ptr(), 2) == n
        assert small.cpu().numpy().astype(np.uint64)[:2].tolist() == want[:2].tolist() and int(small[2]) == 2**62
    k = min(n, n_headers)
    for stride in (8, 64, 256):
        assert c.headers_device(buf.data_ptr() + misalign, len(data), offs.data_ptr(), k, stride) == O.header_lines(data, k, stride)


def test_index_edge_corpus_and_random():
    import torch
    rng = np.random.default_rng(5)
    cases = dict(corpus.edge_cases())
    cases["random"] = corpus.random_fastq(rng, 3000, min_len=0, max_len=300)
    cases["crlf"] = corpus.random_fastq(rng, 2000, min_len=1, max_len=200, crlf=True, final_newline=False)
    cases["dense"] = b"\n" * 70001 + b"x"
    cases["long"] = b"@a\n" + b"A" * 200000 + b"\n+\n" + b"I" * 200000 + b"\n@b\nAC\n+\nII"
    cases["nonl"] = b"ACGT" * 50000
    # several hundred 64 KiB tiles with ragged lines: the chained prefix walks back over many predecessors
    cases["many_tiles"] = corpus.random_fastq(rng, 60000, min_len=0, max_len=600, final_newline=False)
    with fq.FqGpu(meta_records=0) as c:
        for name, data in cases.items():
            for mis in (0, 5):
                _check(c, torch, data, mis)


def test_index_synthetic_large():
    """1.44 GB Illumina stream: every record starts at 360 k; the header gather returns the generator's headers."""
    import torch
    n_records = 4_000_000
    n = 360 * n_records
    buf = torch.empty(n, dtype=torch.uint8, device="cuda")
    with fq.FqGpu(meta_records=0) as c:
        c.synth_illumina(buf.data_ptr(), n, 0, n_records, 20240229)
        offs = torch.empty(n_records, dtype=torch.int64, device="cuda")
        assert c.index_device(buf.data_ptr(), n, offs.data_ptr(), n_records) == n_records
        assert bool((offs == torch.arange(n_records, device="cuda", dtype=torch.int64) * 360).all())
        heads = c.headers_device(buf.data_ptr(), n, offs.data_ptr(), 100, 128)
        ref = bytes(buf[:360 * 100].cpu().numpy())
        assert heads == [ref[360 * k:360 * k + 55] for k in range(100)]
        assert all(h.startswith(b"@A00156:217:HKJWGDSXX:") for h in heads)
