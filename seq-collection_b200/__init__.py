"""seq-collection_b200 -- B200-native FASTQ scanning behind `sc fq-count` / `sc fq-meta`.

Python binding (ctypes) over the C ABI of libfqgpu.so (include/fqgpu.h).  The shared library is the
product; this module only mirrors the reference's two procs for this path:

    fq_count(fastq, basename, absolute)             /root/reference/src/fq_count.nim:14
    fq_meta_quality(fastq, sample_n, ...)           /root/reference/src/fq_meta.nim:197 (quality part)

There is no CPU fallback: if libfqgpu.so is missing or no CUDA device is usable, every compute call
raises FqGpuError.  (The package directory name has a hyphen, so import it with
`importlib.import_module("seq-collection_b200")` or via the `seq_collection_b200` alias module at
the repo root.)
"""
from __future__ import annotations

import ctypes as C
import math
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libfqgpu.so")
POS_BINS = 512
LEN_LOG2_BINS = 64

OK, ECUDA, ENCCL, EIO, EARG, ENOMEM = 0, -1, -2, -3, -4, -5
ERETRY = 1
F_CORE_ONLY = 1  # fqgpu_config.flags: only what `sc fq-count` prints; quality-line outputs are zero


class FqGpuError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libfqgpu error {code}: {msg}")
        self.code = code
        self.msg = msg


class Config(C.Structure):
    _fields_ = [
        ("device", C.c_int),
        ("chunk_bytes", C.c_size_t),
        ("n_buffers", C.c_int),
        ("meta_records", C.c_uint64),
        ("flags", C.c_uint32),
        ("reserved", C.c_uint32),
    ]


class Stats(C.Structure):
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

    SCALARS = ("bytes", "lines", "reads", "bases", "gc_bases", "n_bases", "seq_lines", "qual_lines",
               "seq_len_min", "seq_len_max", "qual_len_min", "qual_len_max",
               "meta_qual_min", "meta_qual_max", "meta_lines", "meta_status")
    ARRAYS = ("base_counts", "qual_counts", "seq_len_hist", "qual_len_hist", "seq_len_log2",
              "qual_pos_sum", "qual_pos_cnt")

    def to_dict(self) -> dict:
        d = {k: int(getattr(self, k)) for k in self.SCALARS}
        for k in self.ARRAYS:
            d[k] = [int(v) for v in getattr(self, k)]
        return d


_lib = None


def load_library(path: str | None = None):
    """dlopen libfqgpu.so and declare the C ABI.  Raises if the library has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("FQGPU_LIB") or LIB_PATH   # (FQGPU_LIB: tuning builds)
    if not os.path.exists(p):
        raise FqGpuError(ECUDA, f"{p} not found: build it with `python seq-collection_b200/build.py` "
                                "(there is no CPU fallback)")
    L = C.CDLL(p, mode=C.RTLD_GLOBAL)
    vp, sz, u64, i32 = C.c_void_p, C.c_size_t, C.c_uint64, C.c_int
    sig = {
        "fqgpu_abi_version": (i32, []),
        "fqgpu_stats_size": (sz, []),
        "fqgpu_build_info": (C.c_char_p, []),
        "fqgpu_device_count": (i32, []),
        "fqgpu_create": (i32, [C.POINTER(vp), C.POINTER(Config)]),
        "fqgpu_destroy": (None, [vp]),
        "fqgpu_last_error": (C.c_char_p, [vp]),
        "fqgpu_acquire": (vp, [vp, C.POINTER(sz)]),
        "fqgpu_submit": (i32, [vp, vp, sz]),
        "fqgpu_finish": (i32, [vp, C.POINTER(Stats)]),
        "fqgpu_reset": (i32, [vp]),
        "fqgpu_scan_host": (i32, [vp, vp, sz]),
        "fqgpu_count_host": (i32, [vp, vp, sz, C.POINTER(Stats)]),
        "fqgpu_count_file": (i32, [vp, C.c_char_p, C.POINTER(Stats)]),
        "fqgpu_count_file_as": (i32, [vp, C.c_char_p, i32, C.POINTER(Stats)]),
        "fqgpu_bgzf_members": (C.c_ulonglong, [vp]),
        "fqgpu_gzip_chunks": (C.c_ulonglong, [vp]),
        "fqgpu_gzip_false_starts": (C.c_ulonglong, [vp]),
        "fqgpu_gzip_second_passes": (C.c_ulonglong, [vp]),
        "fqgpu_meta_file_as": (i32, [vp, C.c_char_p, i32, C.POINTER(Stats)]),
        "fqgpu_count_file_sharded": (i32, [C.POINTER(Config), C.c_char_p, C.POINTER(i32), i32, C.POINTER(Stats)]),
        "fqgpu_count_files": (i32, [C.POINTER(Config), C.POINTER(C.c_char_p), C.POINTER(i32), i32, i32, C.POINTER(Stats), C.POINTER(i32)]),
        "fqgpu_scan_device": (i32, [vp, vp, sz]),
        "fqgpu_count_device": (i32, [vp, vp, sz, C.POINTER(Stats)]),
        "fqgpu_shard_block_words": (sz, []),
        "fqgpu_shard_begin": (i32, [vp, i32, i32]),
        "fqgpu_shard_export": (i32, [vp, vp]),
        "fqgpu_shard_combine": (i32, [vp, vp, C.POINTER(Stats)]),
        "fqgpu_shard_rescan": (i32, [vp, vp]),
        "fqgpu_shard_combine_host": (i32, [i32, vp, u64, C.POINTER(Stats)]),
        "fqgpu_count_pair": (i32, [C.POINTER(Config), C.c_char_p, C.c_char_p, C.POINTER(Stats), C.POINTER(Stats), C.POINTER(i32)]),
        "fqgpu_ipc_handle_bytes": (sz, []),
        "fqgpu_shard_xbuf_bytes": (sz, [i32]),
        "fqgpu_shard_exchange_create": (i32, [vp, i32, i32, vp]),
        "fqgpu_shard_exchange_open": (i32, [vp, vp]),
        "fqgpu_shard_xbuf": (vp, [vp]),
        "fqgpu_shard_exchange_set_peers": (i32, [vp, C.POINTER(vp)]),
        "fqgpu_shard_exchange_start": (i32, [vp]),
        "fqgpu_shard_exchange_finish": (i32, [vp, C.POINTER(Stats)]),
        "fqgpu_shard_exchange_combine": (i32, [vp, C.POINTER(Stats)]),
        "fqgpu_shard_gathered": (vp, [vp]),
        "fqgpu_shard_exchange_destroy": (None, [vp]),
        "fqgpu_last_timing": (i32, [vp, C.POINTER(C.c_double), C.POINTER(u64)]),
        "fqgpu_stream": (vp, [vp]),
        "fqgpu_synth_illumina": (i32, [vp, vp, sz, u64, u64, u64, C.POINTER(sz)]),
        "fqgpu_synth_illumina_bytes": (i32, [vp, vp, u64, u64, u64]),
        "fqgpu_synth_illumina_tally": (i32, [vp, u64, u64, u64, C.POINTER(Stats)]),
        "fqgpu_synth_ont": (i32, [vp, vp, sz, u64, u64, u64, C.POINTER(sz)]),
        "fqgpu_index_device": (i32, [vp, vp, sz, vp, u64, C.POINTER(u64)]),
        "fqgpu_headers_device": (i32, [vp, vp, sz, vp, u64, C.c_uint32, vp, vp]),
        "fqgpu_dedup_device": (i32, [vp, vp, sz, vp, u64, vp, C.POINTER(u64)]),
        "fqgpu_dedup_host": (i32, [vp, vp, sz, vp, u64, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)]),
        "fqgpu_index_lines": (u64, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    if L.fqgpu_stats_size() != C.sizeof(Stats):
        raise FqGpuError(EARG, "Stats mirror out of sync with include/fqgpu.h")
    if path is None:
        _lib = L
    return L


EXPORTED_SYMBOLS = [
    "fqgpu_abi_version", "fqgpu_stats_size", "fqgpu_build_info", "fqgpu_device_count", "fqgpu_create",
    "fqgpu_destroy", "fqgpu_last_error", "fqgpu_acquire", "fqgpu_submit", "fqgpu_finish", "fqgpu_reset",
    "fqgpu_scan_host", "fqgpu_count_host", "fqgpu_count_file", "fqgpu_count_file_as", "fqgpu_count_files", "fqgpu_bgzf_members", "fqgpu_gzip_chunks", "fqgpu_gzip_false_starts", "fqgpu_gzip_second_passes", "fqgpu_meta_file_as", "fqgpu_count_file_sharded", "fqgpu_scan_device", "fqgpu_count_device",
    "fqgpu_shard_block_words", "fqgpu_shard_begin", "fqgpu_shard_export", "fqgpu_shard_combine",
    "fqgpu_shard_rescan", "fqgpu_shard_combine_host", "fqgpu_count_pair", "fqgpu_ipc_handle_bytes", "fqgpu_shard_xbuf_bytes", "fqgpu_shard_exchange_create",
    "fqgpu_shard_exchange_open", "fqgpu_shard_xbuf", "fqgpu_shard_exchange_set_peers", "fqgpu_shard_exchange_start", "fqgpu_shard_exchange_finish",
    "fqgpu_shard_exchange_combine", "fqgpu_shard_gathered", "fqgpu_shard_exchange_destroy", "fqgpu_last_timing", "fqgpu_stream", "fqgpu_synth_illumina",
    "fqgpu_synth_illumina_bytes", "fqgpu_synth_illumina_tally", "fqgpu_synth_ont", "fqgpu_index_device", "fqgpu_headers_device", "fqgpu_dedup_device", "fqgpu_dedup_host", "fqgpu_index_lines",
]


class FqGpu:
    """One context = one GPU + its staging ring, device counters and stream carry (fqgpu_ctx)."""

    def __init__(self, device: int = -1, chunk_bytes: int = 0, n_buffers: int = 0, meta_records: int = 0,
                 flags: int = 0):
        self.lib = load_library()
        self._ctx = C.c_void_p()
        cfg = Config(device, chunk_bytes, n_buffers, meta_records, flags, 0)
        rc = self.lib.fqgpu_create(C.byref(self._ctx), C.byref(cfg))
        if rc != OK:
            msg = self.lib.fqgpu_last_error(None).decode()
            self._ctx = C.c_void_p()
            raise FqGpuError(rc, msg)

    def close(self):
        ctx = getattr(self, "_ctx", None)
        if ctx is not None and ctx.value:
            self._ctx = None
            try:
                self.lib.fqgpu_destroy(ctx)
            except Exception:  # interpreter shutdown
                pass

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc: int):
        if rc < 0:
            raise FqGpuError(rc, self.lib.fqgpu_last_error(self._ctx).decode())
        return rc

    # -- streaming -------------------------------------------------------------------------
    def reset(self):
        self._check(self.lib.fqgpu_reset(self._ctx))

    def acquire(self):
        cap = C.c_size_t()
        p = self.lib.fqgpu_acquire(self._ctx, C.byref(cap))
        if not p:
            raise FqGpuError(ECUDA, self.lib.fqgpu_last_error(self._ctx).decode())
        return p, cap.value

    def submit(self, chunk_ptr: int, nbytes: int):
        self._check(self.lib.fqgpu_submit(self._ctx, chunk_ptr, nbytes))

    def submit_bytes(self, data: bytes):
        """Feed `data` through the pinned ring (splits at the chunk capacity)."""
        off = 0
        n = len(data)
        while off < n:
            p, cap = self.acquire()
            k = min(cap, n - off)
            C.memmove(p, data[off:off + k], k)
            self.submit(p, k)
            off += k

    def finish(self) -> Stats:
        st = Stats()
        self._check(self.lib.fqgpu_finish(self._ctx, C.byref(st)))
        return st

    # -- whole buffers -----------------------------------------------------------------------
    def count_bytes(self, data) -> Stats:
        st = Stats()
        if isinstance(data, (bytes, bytearray)):
            buf = (C.c_char * max(1, len(data))).from_buffer_copy(bytes(data) or b"\0")
            self._check(self.lib.fqgpu_count_host(self._ctx, C.addressof(buf), len(data), C.byref(st)))
        else:  # numpy uint8 array
            self._check(self.lib.fqgpu_count_host(self._ctx, data.ctypes.data, data.size, C.byref(st)))
        return st

    def scan_host_ptr(self, ptr: int, nbytes: int):
        self._check(self.lib.fqgpu_scan_host(self._ctx, ptr, nbytes))

    def count_host_ptr(self, ptr: int, nbytes: int) -> Stats:
        st = Stats()
        self._check(self.lib.fqgpu_count_host(self._ctx, ptr, nbytes, C.byref(st)))
        return st

    def count_file(self, path: str, as_gz: bool | None = None) -> Stats:
        st = Stats()
        if as_gz is None:
            self._check(self.lib.fqgpu_count_file(self._ctx, os.fsencode(path), C.byref(st)))
        else:
            self._check(self.lib.fqgpu_count_file_as(self._ctx, os.fsencode(path), int(as_gz), C.byref(st)))
        return st

    def meta_file(self, path: str, as_gz: bool | None = None) -> Stats:
        """fqgpu_meta_file_as: the head of the file that the fq-meta sampling loop consumes (4 * meta_records lines)."""
        st = Stats()
        gz = path.lower().endswith(".gz") if as_gz is None else as_gz  # case-insensitive, src/fq_meta.nim:219
        self._check(self.lib.fqgpu_meta_file_as(self._ctx, os.fsencode(path), int(gz), C.byref(st)))
        return st

    def bgzf_members(self) -> int:
        """Members of the last count_file() input that were inflated on the device (0: plain gzip / host zlib)."""
        return int(self.lib.fqgpu_bgzf_members(self._ctx))

    def gzip_chunks(self) -> int:
        """Chunks of ordinary gzip input that the last count_file() inflated on the device (0: host zlib ran)."""
        return int(self.lib.fqgpu_gzip_chunks(self._ctx))

    def gzip_second_passes(self) -> int:
        """Batches of the last gzip input whose symbols did not fit the one-pass arena (they were decoded a second time)."""
        return int(self.lib.fqgpu_gzip_second_passes(self._ctx))

    def gzip_false_starts(self) -> int:
        """Block starts the device search proposed that turned out not to be block boundaries (they are skipped)."""
        return int(self.lib.fqgpu_gzip_false_starts(self._ctx))

    # -- HBM resident --------------------------------------------------------------------------
    def scan_device(self, dptr: int, nbytes: int):
        self._check(self.lib.fqgpu_scan_device(self._ctx, dptr, nbytes))

    def count_device(self, dptr: int, nbytes: int) -> Stats:
        st = Stats()
        self._check(self.lib.fqgpu_count_device(self._ctx, dptr, nbytes, C.byref(st)))
        return st

    def last_timing(self):
        ms = C.c_double()
        n = C.c_uint64()
        self._check(self.lib.fqgpu_last_timing(self._ctx, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    @property
    def stream(self) -> int:
        return self.lib.fqgpu_stream(self._ctx) or 0

    # -- multi-GPU shards ----------------------------------------------------------------------
    def shard_block_words(self) -> int:
        return self.lib.fqgpu_shard_block_words()

    def shard_begin(self, rank: int, world: int):
        self._check(self.lib.fqgpu_shard_begin(self._ctx, rank, world))

    def shard_export(self, d_blocks: int):
        self._check(self.lib.fqgpu_shard_export(self._ctx, d_blocks))

    def shard_combine(self, d_blocks: int):
        st = Stats()
        rc = self._check(self.lib.fqgpu_shard_combine(self._ctx, d_blocks, C.byref(st)))
        return rc, st

    def shard_rescan(self, d_blocks: int) -> int:
        """ERETRY (1): the exact carry was installed, scan this rank's range again; OK (0): just export again."""
        return self._check(self.lib.fqgpu_shard_rescan(self._ctx, d_blocks))

    # -- the collective inside the library (peer-memory all-gather + device combine; include/fqgpu.h) ----------
    def shard_exchange_create(self, rank: int, world: int) -> bytes:
        """Allocates this rank's exchange buffer; returns its CUDA IPC handle (bytes) for the other processes."""
        h = C.create_string_buffer(self.lib.fqgpu_ipc_handle_bytes())
        self._check(self.lib.fqgpu_shard_exchange_create(self._ctx, rank, world, h))
        return h.raw

    def shard_exchange_open(self, handles: bytes):
        """`handles`: the IPC handles of all ranks, concatenated in rank order."""
        buf = C.create_string_buffer(handles, len(handles))
        self._check(self.lib.fqgpu_shard_exchange_open(self._ctx, buf))

    def shard_xbuf(self) -> int:
        return int(self.lib.fqgpu_shard_xbuf(self._ctx) or 0)

    def shard_exchange_set_peers(self, xbufs):
        arr = (C.c_void_p * len(xbufs))(*xbufs)
        self._check(self.lib.fqgpu_shard_exchange_set_peers(self._ctx, arr))

    def shard_exchange_start(self):
        self._check(self.lib.fqgpu_shard_exchange_start(self._ctx))

    def shard_exchange_finish(self):
        """(rc, Stats): rc 0 = done, ERETRY = a rank's phase hypothesis was wrong (shard_rescan(shard_gathered()))."""
        st = Stats()
        rc = self._check(self.lib.fqgpu_shard_exchange_finish(self._ctx, C.byref(st)))
        return rc, st

    def shard_gathered(self) -> int:
        return int(self.lib.fqgpu_shard_gathered(self._ctx) or 0)

    # -- synthetic data --------------------------------------------------------------------------
    def synth_illumina(self, dptr: int, capacity: int, first_record: int, n_records: int, seed: int) -> int:
        w = C.c_size_t()
        self._check(self.lib.fqgpu_synth_illumina(self._ctx, dptr, capacity, first_record, n_records, seed, C.byref(w)))
        return w.value

    def synth_illumina_tally(self, first_record: int, n_records: int, seed: int) -> "Stats":
        """The generator's own tallies for records [first_record, first_record + n_records): see fqgpu.h."""
        st = Stats()
        self._check(self.lib.fqgpu_synth_illumina_tally(self._ctx, first_record, n_records, seed, C.byref(st)))
        return st

    def synth_illumina_bytes(self, dptr: int, first_byte: int, nbytes: int, seed: int):
        self._check(self.lib.fqgpu_synth_illumina_bytes(self._ctx, dptr, first_byte, nbytes, seed))

    def index_device(self, dptr: int, nbytes: int, offsets_ptr: int = 0, cap: int = 0) -> int:
        """Record-offset index of an HBM-resident buffer: writes up to `cap` uint64 offsets at `offsets_ptr`
        (device memory) and returns the number of records (fqgpu_index_device)."""
        n = C.c_uint64()
        self._check(self.lib.fqgpu_index_device(self._ctx, dptr, nbytes, offsets_ptr, cap, C.byref(n)))
        return n.value

    def headers_device(self, dptr: int, nbytes: int, offsets_ptr: int, n: int, stride: int = 256) -> list:
        """The first `n` header lines as bytes objects (truncated to `stride`), gathered on the device."""
        if n == 0:
            return []
        out = (C.c_uint8 * (n * stride))()
        lens = (C.c_uint32 * n)()
        self._check(self.lib.fqgpu_headers_device(self._ctx, dptr, nbytes, offsets_ptr, n, stride, out, lens))
        raw = bytes(out)
        return [raw[k * stride:k * stride + lens[k]] for k in range(n)]

    def dedup_device(self, dptr: int, nbytes: int, offsets_ptr: int, n_records: int, keep_ptr: int) -> int:
        """keep[k] = 1 for the first record of every distinct header line; returns the number of dropped records."""
        n = C.c_uint64()
        self._check(self.lib.fqgpu_dedup_device(self._ctx, dptr, nbytes, offsets_ptr, n_records, keep_ptr, C.byref(n)))
        return n.value

    def dedup_bytes(self, data: bytes):
        """(keep flags as bytes, n_records, n_lines, n_dups) of a FASTQ held in host memory (fqgpu_dedup_host)."""
        cap = len(data) // 2 + 1  # a record has at least one newline between two headers... generous bound
        keep = (C.c_uint8 * cap)()
        nrec, nlines, ndups = C.c_uint64(), C.c_uint64(), C.c_uint64()
        buf = (C.c_char * len(data)).from_buffer_copy(data) if data else None
        self._check(self.lib.fqgpu_dedup_host(self._ctx, buf, len(data), keep, cap, C.byref(nrec), C.byref(nlines), C.byref(ndups)))
        return bytes(keep[:nrec.value]), nrec.value, nlines.value, ndups.value

    def synth_ont(self, dptr: int, capacity: int, first_record: int, n_records: int, seed: int) -> int:
        w = C.c_size_t()
        self._check(self.lib.fqgpu_synth_ont(self._ctx, dptr, capacity, first_record, n_records, seed, C.byref(w)))
        return w.value


def shard_combine_host(world: int, blocks_ptr: int, meta_records: int = 0):
    """Combine step over gathered shard blocks in host memory (no GPU needed)."""
    st = Stats()
    rc = load_library().fqgpu_shard_combine_host(world, blocks_ptr, meta_records, C.byref(st))
    if rc < 0:
        raise FqGpuError(rc, "fqgpu_shard_combine_host failed")
    return rc, st


# ------------------------------------------------------------------------------------------------
# Host-side mirror of the reference's output code for this path (unchanged semantics).
# ------------------------------------------------------------------------------------------------
FQ_COUNT_HEADER = "\t".join(["reads", "gc_content", "gc_bases", "n_bases", "bases"])  # src/fq_count.nim:7-11


def nim_float_str(v: float) -> str:
    """Nim 1.0.6 `$`(float): "%.16g", ".0" appended when the text has no '.', ',' or letter."""
    if math.isnan(v):
        return "nan"
    if math.isinf(v):
        return "inf" if v > 0 else "-inf"
    s = "%.16g" % v
    if not any(ch == "." or ch.isalpha() for ch in s):
        s += ".0"
    return s


def output_header(header: str, basename: bool, absolute: bool) -> str:
    """src/utils/helpers.nim:200-208"""
    cols = [header, "basename" if basename else "", "absolute" if absolute else ""]
    return "\t".join(c for c in cols if c)


def output_w_fnames(row: str, path: str, basename: bool, absolute: bool) -> str:
    """src/utils/helpers.nim:210-224"""
    b = os.path.basename(path.rstrip("/")) if basename else ""
    a = ""
    if absolute:
        a = os.path.abspath(os.readlink(path)) if os.path.islink(path) else os.path.abspath(path)
    return "\t".join(c for c in [row, b, a] if c)


def fq_count_row(st) -> str:
    """src/fq_count.nim:47-51 from the integers returned by the GPU."""
    gc, n, total = int(st.gc_bases), int(st.n_bases), int(st.bases)
    den = float(total - n)
    if den == 0.0:
        gc_content = float("nan") if gc == 0 else float("inf")
    else:
        gc_content = float(gc) / den
    return "\t".join([str(int(st.reads)), nim_float_str(gc_content), str(gc), str(n), str(total)])


def fq_count(fastq: str, basename: bool = False, absolute: bool = False, ctx: FqGpu | None = None) -> str:
    """Mirror of fq_count* (src/fq_count.nim:14-53): returns the line the reference echoes.
    Raises FqGpuError(EIO) where the reference does quit_error("Unable to open file", 2)."""
    own = ctx is None
    ctx = ctx or FqGpu(meta_records=0, flags=F_CORE_ONLY)  # fq-count prints reads, GC, N and bases only
    try:
        st = ctx.count_file(fastq)
    finally:
        if own:
            ctx.close()
    return output_w_fnames(fq_count_row(st), fastq, basename, absolute)


DEVICE_ALL = -2  # fqgpu_config.device for count_files: spread the per-thread contexts over every visible GPU


def count_files(paths, n_threads: int = 0, device: int = -1, meta_records: int = 0, flags: int = 0,
                as_gz=None, chunk_bytes: int = 0):
    """fqgpu_count_files: the files are counted concurrently (one host thread + private context per file in
    flight).  Returns (first non-OK rc in file order, [(rc, Stats) per file])."""
    lib = load_library()
    n = len(paths)
    cfg = Config(device, chunk_bytes, 0, meta_records, flags, 0)
    arr = (C.c_char_p * max(n, 1))(*[os.fsencode(p) for p in paths])
    gz = None if as_gz is None else (C.c_int * max(n, 1))(*[int(bool(x)) for x in as_gz])
    out = (Stats * max(n, 1))()
    rcs = (C.c_int * max(n, 1))()
    rc = lib.fqgpu_count_files(C.byref(cfg), arr, gz, n, n_threads, out, rcs)
    return rc, [(rcs[i], out[i]) for i in range(n)]


def count_pair(r1: str, r2: str, device: int = -1, meta_records: int = 0, flags: int = 0):
    """fqgpu_count_pair: R1 / R2 of one library as one job (both mates scanned at the same time).
    Returns (rc, Stats of r1, Stats of r2, paired) -- paired: same number of records and lines."""
    lib = load_library()
    cfg = Config(device, 0, 0, meta_records, flags, 0)
    a, b, ok = Stats(), Stats(), C.c_int(0)
    rc = lib.fqgpu_count_pair(C.byref(cfg), os.fsencode(r1), os.fsencode(r2), C.byref(a), C.byref(b), C.byref(ok))
    return rc, a, b, bool(ok.value)


def count_file_sharded(path: str, devices=None, world: int = 0, meta_records: int = 0, flags: int = 0, chunk_bytes: int = 0) -> Stats:
    """fqgpu_count_file_sharded: one plain file, byte-range sharded over several contexts / GPUs in this process
    (devices = CUDA ordinal per shard, repeats allowed; None = round-robin over all visible devices)."""
    lib = load_library()
    cfg = Config(-1, chunk_bytes, 0, meta_records, flags, 0)
    st = Stats()
    if devices is not None:
        world = len(devices)
        arr = (C.c_int * world)(*devices)
    else:
        arr = None
    rc = lib.fqgpu_count_file_sharded(C.byref(cfg), os.fsencode(path), arr, world, C.byref(st))
    if rc < 0:
        raise FqGpuError(rc, lib.fqgpu_last_error(None).decode())
    return st


def fq_count_many(files, basename: bool = False, absolute: bool = False, n_threads: int = 0) -> list:
    """`sc fq-count a b c ...` (the loop of sc.nim:115-116) with the files counted concurrently: the rows in
    argument order.  Raises FqGpuError at the first file the sequential loop would have failed on."""
    rc, res = count_files(files, n_threads=n_threads, flags=F_CORE_ONLY)
    rows = []
    for f, (r, st) in zip(files, res):
        if r == EIO:
            raise FqGpuError(EIO, "Unable to open file: " + f)
        if r != OK:
            raise FqGpuError(r, load_library().fqgpu_last_error(None).decode())
        rows.append(output_w_fnames(fq_count_row(st), f, basename, absolute))
    return rows


def fq_dedup(fastq: str, ctx: FqGpu | None = None):
    """Mirror of fq_dedup* (src/fq_dedup.nim:14-84): returns (stdout bytes, stderr text).  The duplicate marks come
    from the GPU (index + hash + sort + byte compare); the host only writes the kept records.  The reference's
    "false-positive" figures describe its Bloom filter and are printed as 0 (parity unpinned)."""
    import gzip
    try:
        raw = open(fastq, "rb").read()
    except OSError:
        raise FqGpuError(EIO, "Unable to open file: " + fastq)
    data = gzip.decompress(raw) if fastq[-3:] == ".gz" else raw
    own = ctx is None
    ctx = ctx or FqGpu(meta_records=0)
    try:
        keep, nrec, nlines, ndups = ctx.dedup_bytes(data)
    finally:
        if own:
            ctx.close()
    out = []
    i = 0
    start, n = 0, len(data)
    write_ln = True
    while start < n:  # Nim `lines`: split at LF, drop one CR before it, echo adds LF
        k = data.find(b"\n", start)
        line = data[start:] if k < 0 else data[start:k]
        if k >= 0 and line.endswith(b"\r"):
            line = line[:-1]
        if i % 4 == 0:
            write_ln = bool(keep[i // 4])
        if write_ln:
            out.append(line + b"\n")
        i += 1
        start = n if k < 0 else k + 1
    err = []
    if ndups == 0:
        err += ["No Duplicates Found", "Copying fq to stdout"]
    err += ["total_reads: %d" % (nlines // 4), "duplicates %d" % ndups, "false-positive: 0",
            "false-positive-rate: " + nim_float_str(0.0 / ndups if ndups else float("nan"))]
    return b"".join(out), "\n".join(err) + "\n"


# src/fq_meta.nim:35-39 (bounds reproduced as written)
FASTQ_TYPES = [
    ("Sanger", "Phred+33", 0, 40),
    ("Solexa", "Solexa+64", 59, 104),
    ("Illumina 1.3+", "Phred+64", 64, 104),
    ("Illumina 1.5+", "Phred+64", 64, 104),
    ("Illumina 1.8+", "Phred+33", 0, 42),
]


def fq_meta_quality_fields(st) -> list:
    """qual_format, qual_phred, qual_multiple, min_qual, max_qual, n_lines (src/fq_meta.nim:255-277)."""
    qmin, qmax = int(st.meta_qual_min), int(st.meta_qual_max)
    hits = [t for t in FASTQ_TYPES if qmin >= t[2] and qmax <= t[3]]
    phreds = []
    for t in hits:
        if t[1] not in phreds:
            phreds.append(t[1])
    return [";".join(t[0] for t in hits), ";".join(phreds), "true" if len(hits) > 1 else "false",
            str(qmin) if qmin >= 0 else "", str(qmax) if qmax >= 0 else "", str(int(st.meta_lines) // 4)]


def fq_meta_quality(fastq: str, sample_n: int = 20) -> list:
    """Quality part of fq_meta* (src/fq_meta.nim:197, default sample_n = 20; the CLI passes 100)."""
    with FqGpu(meta_records=sample_n) as ctx:
        return fq_meta_quality_fields(ctx.meta_file(fastq))  # reads the sampled head only, like the reference
