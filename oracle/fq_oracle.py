"""CPU oracle for the FASTQ scanning hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  Nothing under seq-collection_b200/ does.

Two independent restatements of the reference live here:
  * ctypes access to oracle/libfqoracle.so (fq_oracle.c, the C restatement used at size), and
  * `py_count`, a pure-Python twin written straight from the Nim sources
    (/root/reference/src/fq_count.nim:38-45, src/fq_meta.nim:94-102,226-248) for small inputs;
    the two are checked against each other and against the reference's golden tables
    (docs/fq-count.md:29-43, docs/fq-meta.md:34-37) in tests/test_oracle_golden.py.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
POS_BINS = 512
LEN_LOG2_BINS = 64
U64_MAX = (1 << 64) - 1


class Stats(C.Structure):
    """Mirror of fqgpu_stats (include/fqgpu.h) -- kept separate from the product binding on purpose."""

    _fields_ = [
        ("bytes", C.c_uint64),
        ("lines", C.c_uint64),
        ("reads", C.c_uint64),
        ("bases", C.c_uint64),
        ("gc_bases", C.c_uint64),
        ("n_bases", C.c_uint64),
        ("seq_lines", C.c_uint64),
        ("qual_lines", C.c_uint64),
        ("base_counts", C.c_uint64 * 256),
        ("qual_counts", C.c_uint64 * 256),
        ("seq_len_min", C.c_uint64),
        ("seq_len_max", C.c_uint64),
        ("qual_len_min", C.c_uint64),
        ("qual_len_max", C.c_uint64),
        ("seq_len_hist", C.c_uint64 * (POS_BINS + 1)),
        ("qual_len_hist", C.c_uint64 * (POS_BINS + 1)),
        ("seq_len_log2", C.c_uint64 * LEN_LOG2_BINS),
        ("qual_pos_sum", C.c_uint64 * (POS_BINS + 1)),
        ("qual_pos_cnt", C.c_uint64 * (POS_BINS + 1)),
        ("meta_qual_min", C.c_int64),
        ("meta_qual_max", C.c_int64),
        ("meta_lines", C.c_uint64),
        ("meta_status", C.c_uint32),
        ("reserved", C.c_uint32),
    ]


SCALARS = [
    "bytes", "lines", "reads", "bases", "gc_bases", "n_bases", "seq_lines", "qual_lines",
    "seq_len_min", "seq_len_max", "qual_len_min", "qual_len_max",
    "meta_qual_min", "meta_qual_max", "meta_lines", "meta_status",
]
ARRAYS = ["base_counts", "qual_counts", "seq_len_hist", "qual_len_hist", "seq_len_log2",
          "qual_pos_sum", "qual_pos_cnt"]


def stats_to_dict(st) -> dict:
    d = {k: int(getattr(st, k)) for k in SCALARS}
    for k in ARRAYS:
        d[k] = [int(v) for v in getattr(st, k)]
    return d


_lib = None


def build() -> str:
    """Compile oracle/libfqoracle.so with the committed Makefile (gcc only)."""
    subprocess.run(["make", "-s", "-C", HERE], check=True)
    return os.path.join(HERE, "libfqoracle.so")


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "libfqoracle.so")
        srcs = [os.path.join(HERE, f) for f in ("fq_oracle.c", "fq_synth_twin.c")]
        if not os.path.exists(path) or any(os.path.getmtime(path) < os.path.getmtime(f) for f in srcs):
            build()
        L = C.CDLL(path)
        L.fqo_stats_size.restype = C.c_size_t
        assert L.fqo_stats_size() == C.sizeof(Stats), "oracle Stats mirror out of sync with fqgpu.h"
        L.fqo_count_buffer.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64, C.POINTER(Stats)]
        L.fqo_count_buffer.restype = None
        L.fqo_count_buffer_chunked.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_uint64, C.POINTER(Stats)]
        L.fqo_count_buffer_chunked.restype = None
        L.fqo_count_file.argtypes = [C.c_char_p, C.c_uint64, C.POINTER(Stats)]
        L.fqo_count_file.restype = C.c_int
        L.fqo_format_float.argtypes = [C.c_double, C.c_char_p, C.c_size_t]
        L.fqo_format_float.restype = C.c_int
        L.fqo_format_fq_count_row.argtypes = [C.POINTER(Stats), C.c_char_p, C.c_size_t]
        L.fqo_format_fq_count_row.restype = C.c_int
        L.fqo_ref_fq_count_file.argtypes = [C.c_char_p, C.POINTER(C.c_uint64)]
        L.fqo_ref_fq_count_file.restype = C.c_int
        L.fqo_ref_fq_count_mem.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_uint64)]
        L.fqo_ref_fq_count_mem.restype = None
        L.fqo_synth_illumina_bytes.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64]
        L.fqo_synth_illumina_bytes.restype = None
        L.fqo_synth_illumina_tally.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(Stats)]
        L.fqo_synth_illumina_tally.restype = None
        _lib = L
    return _lib


def _as_buffer(data):
    """bytes / bytearray / numpy uint8 array -> (address, nbytes, keepalive)."""
    try:
        import numpy as np

        if isinstance(data, np.ndarray):
            a = np.ascontiguousarray(data, dtype=np.uint8)
            return a.ctypes.data, a.size, a
    except ImportError:  # pragma: no cover
        pass
    b = bytes(data)
    buf = C.create_string_buffer(b, len(b)) if len(b) else C.create_string_buffer(1)
    return C.addressof(buf), len(b), buf


def count(data, meta_records: int = 0, chunk: int = 0) -> dict:
    """C oracle over an in-memory byte string; `chunk` > 0 feeds it in pieces of that size."""
    addr, n, keep = _as_buffer(data)
    st = Stats()
    if chunk:
        lib().fqo_count_buffer_chunked(addr, n, chunk, meta_records, C.byref(st))
    else:
        lib().fqo_count_buffer(addr, n, meta_records, C.byref(st))
    del keep
    return stats_to_dict(st)


def synth_illumina_bytes(first_byte: int, nbytes: int, seed: int = 20240229):
    """CPU twin of the GPU generator (oracle/fq_synth_twin.c): bytes [first_byte, first_byte + nbytes) of the synthetic
    Illumina 2x150 stream as a numpy uint8 array."""
    import numpy as np

    out = np.empty(nbytes, dtype=np.uint8)
    lib().fqo_synth_illumina_bytes(out.ctypes.data, first_byte, nbytes, seed)
    return out


def synth_illumina_tally(first_record: int, n_records: int, seed: int = 20240229, meta_records: int = 0) -> dict:
    """Expected statistics of records [first_record, first_record + n_records) of that stream, derived from the random
    numbers alone (no bytes are read back): an independent check of the scan at any size."""
    st = Stats()
    lib().fqo_synth_illumina_tally(first_record, n_records, seed, meta_records, C.byref(st))
    return stats_to_dict(st)


def count_file(path: str, meta_records: int = 0) -> dict:
    st = Stats()
    rc = lib().fqo_count_file(os.fsencode(path), meta_records, C.byref(st))
    if rc != 0:
        raise OSError(f"Unable to open file: {path}")
    return stats_to_dict(st)


def format_float(v: float) -> str:
    buf = C.create_string_buffer(80)
    lib().fqo_format_float(v, buf, 80)
    return buf.value.decode()


def fq_count_row(d: dict) -> str:
    """src/fq_count.nim:47-51 -- the five tab-separated columns, from the integers."""
    den = d["bases"] - d["n_bases"]
    gc = float("nan") if den == 0 and d["gc_bases"] == 0 else (float(d["gc_bases"]) / float(den) if den else float("inf"))
    return "\t".join([str(d["reads"]), format_float(gc), str(d["gc_bases"]), str(d["n_bases"]), str(d["bases"])])


def ref_fq_count_mem(data) -> dict:
    """Reference work shape (line reader + 3 count passes), single core, for the cpu_baseline leg."""
    addr, n, keep = _as_buffer(data)
    out = (C.c_uint64 * 5)()
    lib().fqo_ref_fq_count_mem(addr, n, out)
    del keep
    return {"reads": out[0], "gc_bases": out[1], "n_bases": out[2], "bases": out[3], "lines": out[4]}


def ref_fq_count_file(path: str) -> dict:
    out = (C.c_uint64 * 5)()
    rc = lib().fqo_ref_fq_count_file(os.fsencode(path), out)
    if rc != 0:
        raise OSError(f"Unable to open file: {path}")
    return {"reads": out[0], "gc_bases": out[1], "n_bases": out[2], "bases": out[3], "lines": out[4]}


# --------------------------------------------------------------------------------------------
# Pure-Python twin (small inputs only), written from the Nim sources, not from fq_oracle.c.
# --------------------------------------------------------------------------------------------
QUAL = bytes(range(33, 127))  # src/fq_meta.nim:10


def nim_lines(data: bytes):
    """Nim 1.0.6 streams.lines over a plain file: split at LF, strip one CR before it, yield the
    unterminated tail only when non-empty."""
    start = 0
    n = len(data)
    while start < n:
        k = data.find(b"\n", start)
        if k < 0:
            yield data[start:]
            return
        line = data[start:k]
        if line.endswith(b"\r"):
            line = line[:-1]
        yield line
        start = k + 1


def py_count(data: bytes, meta_records: int = 0) -> dict:
    d = {k: 0 for k in SCALARS}
    for k in ARRAYS:
        d[k] = [0] * {"base_counts": 256, "qual_counts": 256, "seq_len_log2": LEN_LOG2_BINS}.get(k, POS_BINS + 1)
    d["seq_len_min"] = d["qual_len_min"] = U64_MAX
    d["bytes"] = len(data)
    qual_min = qual_max = -1
    i = 0
    for line in nim_lines(data):
        i += 1  # fq_count.nim:39
        if i % 4 == 1:
            d["reads"] += 1
        if i % 4 == 2:
            d["gc_bases"] += line.count(b"G") + line.count(b"C")
            d["n_bases"] += line.count(b"N")
            d["bases"] += len(line)
            d["seq_lines"] += 1
            for b in line:
                d["base_counts"][b] += 1
            L = len(line)
            d["seq_len_min"] = min(d["seq_len_min"], L)
            d["seq_len_max"] = max(d["seq_len_max"], L)
            d["seq_len_hist"][min(L, POS_BINS)] += 1
            d["seq_len_log2"][L.bit_length()] += 1
        if i % 4 == 0:
            d["qual_lines"] += 1
            L = len(line)
            for p, b in enumerate(line):
                d["qual_counts"][b] += 1
                d["qual_pos_sum"][min(p, POS_BINS)] += b
                d["qual_pos_cnt"][min(p, POS_BINS)] += 1
            d["qual_len_min"] = min(d["qual_len_min"], L)
            d["qual_len_max"] = max(d["qual_len_max"], L)
            d["qual_len_hist"][min(L, POS_BINS)] += 1
        j = i - 1  # fq_meta.nim's 0-based counter
        if meta_records and j < meta_records * 4:
            d["meta_lines"] = i
            if j % 4 == 3 and d["meta_status"] == 0:
                scores = [QUAL.find(bytes([c])) for c in line]  # qual_to_int, fq_meta.nim:94-95,99
                if qual_min >= 0:
                    scores += [qual_min, qual_max]  # :100-101
                if not scores:
                    d["meta_status"] = 1
                else:
                    qual_min, qual_max = min(scores), max(scores)  # :102
    d["lines"] = i
    d["meta_qual_min"], d["meta_qual_max"] = qual_min, qual_max
    return d


# src/fq_meta.nim:35-39 -- encoding table, bounds reproduced as written (not "fixed")
FASTQ_TYPES = [
    ("Sanger", "Phred+33", 0, 40),
    ("Solexa", "Solexa+64", 59, 104),
    ("Illumina 1.3+", "Phred+64", 64, 104),
    ("Illumina 1.5+", "Phred+64", 64, 104),
    ("Illumina 1.8+", "Phred+33", 0, 42),
]


def fq_meta_quality_fields(d: dict) -> list:
    """The six quality-related output columns of fq-meta (src/fq_meta.nim:255,259-260,272-277)."""
    qmin, qmax = d["meta_qual_min"], d["meta_qual_max"]
    hits = [t for t in FASTQ_TYPES if qmin >= t[2] and qmax <= t[3]]
    names = ";".join(t[0] for t in hits)
    phreds = []
    for t in hits:
        if t[1] not in phreds:
            phreds.append(t[1])
    return [names, ";".join(phreds), "true" if len(hits) > 1 else "false",
            str(qmin) if qmin >= 0 else "", str(qmax) if qmax >= 0 else "", str(d["meta_lines"] // 4)]


def record_offsets(data) -> "np.ndarray":
    """Byte offset of the first byte of every record (line 4k) under the reference's line rule
    (src/fq_count.nim:38-42): the checker of fqgpu_index_device."""
    import numpy as np
    arr = np.frombuffer(bytes(data), dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    if arr.size == 0:
        return np.zeros(0, dtype=np.uint64)
    nl = np.flatnonzero(arr == 10)
    starts = nl[3::4].astype(np.uint64) + 1
    starts = starts[starts < arr.size]
    return np.concatenate([np.zeros(1, dtype=np.uint64), starts])


def header_lines(data, n: int, stride: int = 256) -> list:
    """First n header lines (lines 4k) without the newline and without one '\\r' before it, truncated to stride."""
    b = bytes(data)
    offs = record_offsets(b)[:n]
    out = []
    for o in offs:
        o = int(o)
        e = b.find(b"\n", o)
        line = b[o:] if e < 0 else b[o:e]
        if e >= 0 and line.endswith(b"\r"):
            line = line[:-1]
        out.append(line[:stride])
    return out


def fq_dedup(data: bytes):
    """Restatement of fq_dedup* (src/fq_dedup.nim:14-84) without its Bloom filter (which only pre-selects the
    candidates; the second pass decides with an exact table): returns (stdout bytes, n_reads, n_dups, keep flags).
    The first record of every distinct header line (line 4k as `lines` yields it) is echoed, later ones are
    dropped with their following lines; `echo` terminates every line with '\\n'.  n_reads = lines div 4 (:50).
    The "false-positive" figures of the reference depend on its Bloom filter (nimble `bloom`, not vendored):
    parity unpinned, the mirrors print 0."""
    seen = set()
    out = []
    keep = []
    n_dups = 0
    write_ln = True
    i = 0
    for i, ln in enumerate(nim_lines(data), 1):
        if (i - 1) % 4 == 0:
            if ln in seen:
                write_ln = False
                n_dups += 1
                keep.append(0)
                continue
            seen.add(ln)
            keep.append(1)
            write_ln = True
            out.append(ln + b"\n")
        elif write_ln:
            out.append(ln + b"\n")
    return b"".join(out), i // 4, n_dups, keep
