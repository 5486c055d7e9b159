"""The synthetic Illumina generator: the CPU twin (oracle/fq_synth_twin.c) against the oracle on its
own bytes (CPU), and the GPU generator / its tallies against the twin (GPU)."""
import numpy as np
import pytest

from oracle import fq_oracle as O
from tests.test_gpu_parity import assert_equal_stats

SEED = 20240229


@pytest.mark.parametrize("first,n,meta", [(0, 1, 1), (0, 257, 100), (12345, 300, 7), (99_999_700, 300, 1000)])
def test_twin_tally_equals_oracle_count_of_twin_bytes(first, n, meta):
    data = O.synth_illumina_bytes(first * 360, n * 360, SEED)
    assert data[0] == ord("@") and data[-1] == 10
    want = O.count(data, meta)
    got = O.synth_illumina_tally(first, n, SEED, meta)
    assert_equal_stats(got, want, f"first={first} n={n}")


def test_twin_byte_ranges_are_consistent():
    whole = O.synth_illumina_bytes(0, 360 * 40, SEED)
    for off, n in ((0, 1), (1, 359), (359, 2), (777, 5000), (360 * 39, 360)):
        assert np.array_equal(O.synth_illumina_bytes(off, n, SEED), whole[off:off + n])


@pytest.mark.gpu
def test_gpu_generator_equals_twin_and_tallies_agree():
    import torch
    import seq_collection_b200 as fq

    with fq.FqGpu(meta_records=100) as c, fq.FqGpu(meta_records=100, flags=fq.F_CORE_ONLY) as core:
        for first_byte, nbytes in ((0, 360 * 1000), (360 * 5 + 17, 100_003), (360 * 99_000_000 + 359, 70_001)):
            buf = torch.zeros(nbytes + 64, dtype=torch.uint8, device="cuda")
            c.synth_illumina_bytes(buf.data_ptr(), first_byte, nbytes, SEED)
            assert np.array_equal(buf[:nbytes].cpu().numpy(), O.synth_illumina_bytes(first_byte, nbytes, SEED)), (first_byte, nbytes)
        for first, n in ((0, 1), (3, 1000), (50_000_000, 123_457)):
            got = c.synth_illumina_tally(first, n, SEED).to_dict()
            assert_equal_stats(got, O.synth_illumina_tally(first, n, SEED, 100), f"tally first={first} n={n}")
            # ... and the scan of exactly these records reports them
            buf = torch.empty(n * 360 + 64, dtype=torch.uint8, device="cuda")
            c.synth_illumina(buf.data_ptr(), n * 360, first, n, SEED)
            assert_equal_stats(c.count_device(buf.data_ptr(), n * 360).to_dict(), got, f"scan first={first} n={n}")
            got_core = core.synth_illumina_tally(first, n, SEED).to_dict()
            assert_equal_stats(core.count_device(buf.data_ptr(), n * 360).to_dict(), got_core, f"core scan first={first} n={n}")
