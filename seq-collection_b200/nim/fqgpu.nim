## fqgpu.nim -- Nim binding of libfqgpu (include/fqgpu.h) for danielecook/seq-collection.
##
## NOT COMPILED IN THIS ENVIRONMENT: the build image has no Nim toolchain (nim / nimble absent, no
## network).  It follows the only Nim -> C pattern the reference itself uses (zlib through `importc`,
## gzip_stream.nim:1,14-23) and is what a maintainer drops next to src/fq_count.nim; see
## INTEGRATION.md for the replaced loops and the nim.cfg lines.
##
## Every proc maps 1:1 to an extern "C" entry point; only integers cross the boundary.

const
  FQGPU_POS_BINS* = 512
  FQGPU_LEN_LOG2_BINS* = 64
  FQGPU_OK* = 0
  FQGPU_EIO* = -3

const FQGPU_F_CORE_ONLY* = 1'u32   ## fqgpu_config.flags: only what `sc fq-count` prints

type
  FqgpuCtx* = pointer
  FqgpuConfig* {.bycopy.} = object
    device*: cint            ## CUDA ordinal, -1 = current
    chunk_bytes*: csize_t    ## pinned chunk size, 0 = 64 MiB
    n_buffers*: cint         ## ring depth, 0 = 3
    meta_records*: uint64    ## fq-meta sample_n (0 = skip the quality-range scan)
    flags*: uint32
    reserved*: uint32
  FqgpuStats* {.bycopy.} = object
    bytes*, lines*, reads*, bases*, gc_bases*, n_bases*, seq_lines*, qual_lines*: uint64
    base_counts*: array[256, uint64]
    qual_counts*: array[256, uint64]
    seq_len_min*, seq_len_max*, qual_len_min*, qual_len_max*: uint64
    seq_len_hist*: array[FQGPU_POS_BINS + 1, uint64]
    qual_len_hist*: array[FQGPU_POS_BINS + 1, uint64]
    seq_len_log2*: array[FQGPU_LEN_LOG2_BINS, uint64]
    qual_pos_sum*: array[FQGPU_POS_BINS + 1, uint64]
    qual_pos_cnt*: array[FQGPU_POS_BINS + 1, uint64]
    meta_qual_min*, meta_qual_max*: int64
    meta_lines*: uint64
    meta_status*: uint32
    reserved*: uint32

{.push importc, cdecl, dynlib: "libfqgpu.so".}
proc fqgpu_create*(ctx: ptr FqgpuCtx, cfg: ptr FqgpuConfig): cint
proc fqgpu_destroy*(ctx: FqgpuCtx)
proc fqgpu_last_error*(ctx: FqgpuCtx): cstring
proc fqgpu_acquire*(ctx: FqgpuCtx, capacity: ptr csize_t): pointer
proc fqgpu_submit*(ctx: FqgpuCtx, chunk: pointer, nbytes: csize_t): cint
proc fqgpu_finish*(ctx: FqgpuCtx, stats: ptr FqgpuStats): cint
proc fqgpu_reset*(ctx: FqgpuCtx): cint
proc fqgpu_count_file_as*(ctx: FqgpuCtx, path: cstring, as_gz: cint, stats: ptr FqgpuStats): cint
proc fqgpu_count_files*(cfg: ptr FqgpuConfig, paths: cstringArray, as_gz: ptr cint, n, n_threads: cint,
                        stats: ptr FqgpuStats, rc: ptr cint): cint
proc fqgpu_bgzf_members*(ctx: FqgpuCtx): culonglong
proc fqgpu_gzip_chunks*(ctx: FqgpuCtx): culonglong        # > 0: an ordinary .gz was inflated on the device
proc fqgpu_gzip_false_starts*(ctx: FqgpuCtx): culonglong
proc fqgpu_gzip_second_passes*(ctx: FqgpuCtx): culonglong
proc fqgpu_meta_file_as*(ctx: FqgpuCtx, path: cstring, as_gz: cint, stats: ptr FqgpuStats): cint
proc fqgpu_count_file_sharded*(cfg: ptr FqgpuConfig, path: cstring, devices: ptr cint, world: cint, stats: ptr FqgpuStats): cint
proc fqgpu_count_pair*(cfg: ptr FqgpuConfig, r1, r2: cstring, stats1, stats2: ptr FqgpuStats, paired: ptr cint): cint
{.pop.}

## ------------------------------------------------------------------------------------------------
## Replacement of the hot loop of src/fq_count.nim:38-45.  `stream` is the Stream the reference already
## opened at :30-36 (plain FileStream or GZFileStream); readData fills the pinned chunk exactly as
## gzip_stream.nim:16-17 fills a caller buffer.  Everything after the loop (:47-53) stays as it is.
## ------------------------------------------------------------------------------------------------
import streams

proc fq_count_gpu*(stream: Stream, gc_cnt, n_cnt, total_len: var int64, n_reads: var int) =
  var ctx: FqgpuCtx
  var cfg = FqgpuConfig(device: -1, flags: FQGPU_F_CORE_ONLY)   # fq-count prints reads, GC, N, bases: sequence lines only
  if fqgpu_create(addr ctx, addr cfg) != FQGPU_OK:
    raise newException(IOError, $fqgpu_last_error(nil))
  defer: fqgpu_destroy(ctx)
  while true:
    var cap: csize_t
    let chunk = fqgpu_acquire(ctx, addr cap)        # next free pinned chunk (blocks while all are in flight)
    let got = stream.readData(chunk, cap.int)       # plain read() or gzread() straight into pinned memory
    if got <= 0: break
    if fqgpu_submit(ctx, chunk, got.csize_t) != FQGPU_OK:   # async H2D + scan; a chunk may end mid-line
      raise newException(IOError, $fqgpu_last_error(ctx))
  var st: FqgpuStats
  if fqgpu_finish(ctx, addr st) != FQGPU_OK:
    raise newException(IOError, $fqgpu_last_error(ctx))
  n_reads = st.reads.int
  gc_cnt = st.gc_bases.int64
  n_cnt = st.n_bases.int64
  total_len = st.bases.int64

## Replacement of the fold at src/fq_meta.nim:245-246 (qual_min / qual_max / i over the first sample_n
## records); the header parsing of :229-242 keeps reading the first lines on the host.
proc fq_meta_quality_gpu*(fastq: string, as_gz: bool, sample_n: int): tuple[qual_min, qual_max, lines: int] =
  var ctx: FqgpuCtx
  var cfg = FqgpuConfig(device: -1, meta_records: sample_n.uint64)
  if fqgpu_create(addr ctx, addr cfg) != FQGPU_OK:
    raise newException(IOError, $fqgpu_last_error(nil))
  defer: fqgpu_destroy(ctx)
  var st: FqgpuStats
  let rc = fqgpu_meta_file_as(ctx, fastq.cstring, as_gz.cint, addr st)   # reads the sampled head only, like the loop at :226
  if rc == FQGPU_EIO: raise newException(IOError, "Unable to open file: " & fastq)   # quit_error(..., 2) upstream
  if rc != FQGPU_OK: raise newException(IOError, $fqgpu_last_error(ctx))
  if st.meta_status == 1: raise newException(IndexError, "index out of bounds, the container is empty")
  return (st.meta_qual_min.int, st.meta_qual_max.int, st.meta_lines.int)


## Replacement of the loop over the files, sc.nim:115-116 (`for fastq in opts.fastq: fq_count.fq_count(fastq, ...)`):
## the files are counted concurrently (one host thread and private context per file in flight) and the rows are
## echoed in argument order; the first failure in file order ends the run like the sequential loop would have.
proc fq_count_many_gpu*(files: seq[string]): seq[FqgpuStats] =
  var cfg = FqgpuConfig(device: -1, flags: FQGPU_F_CORE_ONLY)
  result = newSeq[FqgpuStats](files.len)
  var rcs = newSeq[cint](files.len)
  let paths = allocCStringArray(files)
  defer: deallocCStringArray(paths)
  discard fqgpu_count_files(addr cfg, paths, nil, files.len.cint, 0, addr result[0], addr rcs[0])
  for i, rc in rcs:
    if rc == FQGPU_EIO: raise newException(IOError, "Unable to open file: " & files[i])   # quit_error(..., 2) upstream
    if rc != FQGPU_OK: raise newException(IOError, $fqgpu_last_error(nil))
