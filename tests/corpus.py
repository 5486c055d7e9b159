"""Edge-case FASTQ corpus + small numpy synthetic generators shared by the CPU and GPU tests.

The reference's fixtures (tests/golden/fastq) contain no 'N', no lowercase base, no CR, no NUL and
at most 9 records, so everything the hot path must get right beyond them is
enumerated here and checked oracle-vs-oracle (C vs pure Python) on CPU and CUDA-vs-oracle on GPU.
"""
from __future__ import annotations

import numpy as np


def rec(h: bytes, s: bytes, q: bytes, plus: bytes = b"+", nl: bytes = b"\n") -> bytes:
    return b"@" + h + nl + s + nl + plus + nl + q + nl


def edge_cases() -> dict:
    c = {}
    c["empty"] = b""
    c["one_newline"] = b"\n"
    c["only_newlines"] = b"\n" * 37
    c["one_line_no_nl"] = b"@r1"
    c["one_line_nl"] = b"@r1\n"
    c["two_lines"] = b"@r1\nACGT"
    c["three_lines"] = b"@r1\nACGT\n+"
    c["four_lines_no_nl"] = b"@r1\nACGT\n+\nIIII"
    c["four_lines_nl"] = b"@r1\nACGT\n+\nIIII\n"
    c["header_only_partial_record"] = rec(b"a", b"GGCC", b"IIII") + b"@b\n"
    c["crlf"] = rec(b"a", b"ACGTN", b"IIIII", nl=b"\r\n") * 3
    c["crlf_no_final"] = rec(b"a", b"ACGTN", b"IIIII", nl=b"\r\n") + b"@b\r\nGG\r\n+\r\nII"
    c["cr_inside_line"] = b"@a\nAC\rGT\n+\nII\rII\n"
    c["trailing_cr_no_lf"] = b"@a\nACGT\r"
    c["double_cr_lf"] = b"@a\nACGT\r\r\n+\nIIII\r\r\n"
    c["lone_cr_line"] = b"@a\n\r\n+\n\r\n"
    c["blank_lines"] = b"@a\n\n+\n\n@b\nAC\n+\nII\n"
    c["blank_line_shifts_phase"] = b"\n@a\nACGT\n+\nIIII\n"
    c["lowercase"] = rec(b"a", b"acgtnACGTN", b"IIIIIIIIII")
    c["all_n"] = rec(b"a", b"N" * 50, b"#" * 50)
    c["seq_qual_len_mismatch"] = rec(b"a", b"A", b"F" * 101) * 5
    c["qual_starts_with_at"] = rec(b"a", b"ACGT", b"@III") + rec(b"b", b"GGGG", b"+@@@")
    c["plus_repeats_id"] = rec(b"id1", b"ACGT", b"IIII", plus=b"+id1")
    c["high_bytes"] = rec(b"a", bytes([200, 65, 255, 0x80]), bytes([33, 126, 127, 255, 32]))
    c["nul_bytes"] = rec(b"a", b"AC\x00GT", b"II\x00II")
    c["qual_invalid_then_valid"] = rec(b"a", b"ACGT", b"II I") + rec(b"b", b"ACGT", b"5555") + rec(b"c", b"AC", b"++")
    c["qual_invalid_last"] = rec(b"a", b"ACGT", b"5555") + rec(b"b", b"ACGT", b"I\x7fII")
    c["qual_invalid_twice"] = rec(b"a", b"AC", b"\x1fI") + rec(b"b", b"AC", b"\x1f#") + rec(b"c", b"AC", b"AB")
    c["empty_first_qual"] = b"@a\nACGT\n+\n\n" + rec(b"b", b"AC", b"II")
    c["empty_later_qual"] = rec(b"b", b"AC", b"I5") + b"@a\nACGT\n+\n\n"
    c["long_line_300k"] = b"@long\n" + b"ACGTN" * 60000 + b"\n+\n" + bytes((33 + (i * 7) % 60) for i in range(300000)) + b"\n"
    c["pos_bins_boundary"] = b"".join(rec(b"x", b"A" * L, b"I" * L) for L in (510, 511, 512, 513, 514, 1024))
    c["many_short"] = b"".join(rec(b"%d" % i, b"ACGTN"[: 1 + i % 5], b"!I5~#"[: 1 + i % 5]) for i in range(3000))
    c["single_byte_lines"] = b"A\n" * 1000 + b"G"
    c["no_newline_blob"] = b"ACGT" * 5000
    return c


# --- numpy twins of the device generators (seq-collection_b200/csrc/fq_synth.cuh); small sizes only ---

MASK64 = (1 << 64) - 1


def splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & MASK64
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
    return z ^ (z >> 31)


def random_fastq(rng: np.random.Generator, n_records: int, min_len=1, max_len=300, crlf=False,
                 final_newline=True, alphabet=b"ACGTN", qual_lo=33, qual_hi=74) -> bytes:
    out = []
    nl = b"\r\n" if crlf else b"\n"
    for i in range(n_records):
        L = int(rng.integers(min_len, max_len + 1))
        s = bytes(rng.choice(list(alphabet), size=L).astype(np.uint8))
        q = bytes(rng.integers(qual_lo, qual_hi + 1, size=L, dtype=np.uint8))
        out.append(b"@r%d len=%d" % (i, L) + nl + s + nl + b"+" + nl + q + nl)
    data = b"".join(out)
    if not final_newline and data:
        data = data[: -len(nl)]
    return data


def bgzf_bytes(data: bytes, block: int = 65280, level: int = 6, eof: bool = True, strategy: int = 0) -> bytes:
    """BGZF (blocked gzip, SAM spec 4.1): one gzip member per `block` input bytes, each with the 'BC' extra field
    holding its compressed size; optionally the 28-byte empty EOF member."""
    import struct
    import zlib

    out = bytearray()
    for i in range(0, max(len(data), 1), block):
        chunk = data[i:i + block]
        co = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strategy)
        comp = co.compress(chunk) + co.flush()
        out += struct.pack("<BBBBIBBH", 0x1F, 0x8B, 8, 4, 0, 0, 0xFF, 6) + b"BC" + struct.pack("<HH", 2, len(comp) + 25)
        out += comp + struct.pack("<II", zlib.crc32(chunk), len(chunk))
    if eof:
        out += bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
    return bytes(out)
