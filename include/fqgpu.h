/*
 * fqgpu.h -- C ABI of libfqgpu: the B200 (sm_100a) FASTQ scanning hot path behind
 * `sc fq-count` and the quality-range scan of `sc fq-meta` (danielecook/seq-collection).
 *
 * The reference has no FFI for this path: the "interface" is two Nim procs,
 *   fq_count*(fastq, basename, absolute)            src/fq_count.nim:14  (called at sc.nim:116)
 *   qual_min_max*(quality_scores, prev_min, prev_max) src/fq_meta.nim:97 (called at src/fq_meta.nim:246)
 * whose bodies iterate text lines on one CPU core.  This header is what a Nim `importc` shim binds
 * (see INTEGRATION.md and seq-collection_b200/nim/fqgpu.nim): the loop body of
 * src/fq_count.nim:38-45 and the fold of src/fq_meta.nim:245-246 become
 *   "fill pinned chunk -> fqgpu_submit -> ... -> fqgpu_finish -> read integers",
 * and the unchanged output code (src/fq_count.nim:47-52, src/fq_meta.nim:255-278) runs on the
 * integers returned in fqgpu_stats.  Only integers cross the boundary; the single float
 * (gc_content, src/fq_count.nim:48) is derived by the caller exactly as the reference does.
 *
 * Plain C99: pointers and sizes only, no exceptions, no torch types.  All functions returning int
 * return FQGPU_OK (0) or a negative FQGPU_E* code; the message is available from fqgpu_last_error().
 * The library never writes to stdout and never calls exit().  There is NO CPU fallback: if no CUDA
 * device is usable every entry point that computes fails with FQGPU_ECUDA.
 */
#ifndef FQGPU_H
#define FQGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FQGPU_ABI_VERSION 1

/* Per-position tables hold positions 0..FQGPU_POS_BINS-1; index FQGPU_POS_BINS is the overflow bin
 * (all positions >= FQGPU_POS_BINS).  Exact length histograms use the same binning. */
#define FQGPU_POS_BINS 512
#define FQGPU_LEN_LOG2_BINS 64

enum {
  FQGPU_OK = 0,
  FQGPU_ECUDA = -1, /* CUDA runtime / no device / kernel failure              */
  FQGPU_ENCCL = -2, /* collective failure (multi-GPU): a peer buffer cannot be opened, a rank is missing */
  FQGPU_EIO = -3,   /* file could not be opened / read  (maps to exit code 2, src/fq_count.nim:36) */
  FQGPU_EARG = -4,  /* bad argument / call sequence                           */
  FQGPU_ENOMEM = -5
};

/* fqgpu_config.flags */
enum {
  FQGPU_F_CORE_ONLY = 1u << 0 /* compute only what `sc fq-count` prints (reads, bases, G/C/N):
                                 base_counts/len tables for sequence lines stay exact, all
                                 quality-line outputs are left zero.  Default: full statistics. */
};

typedef struct fqgpu_ctx fqgpu_ctx;

typedef struct {
  int device;            /* CUDA ordinal; -1 = current device                                   */
  size_t chunk_bytes;    /* pinned staging chunk size; 0 = default (64 MiB)                     */
  int n_buffers;         /* staging ring depth; 0 = default (3)                                 */
  uint64_t meta_records; /* fq-meta sample_n: quality range is folded over the first
                            meta_records records (sc.nim:70 default 100); 0 = skip the scan     */
  uint32_t flags;        /* FQGPU_F_*                                                            */
  uint32_t reserved;
} fqgpu_config;

/* meta_status values */
enum {
  FQGPU_META_OK = 0,
  FQGPU_META_EMPTY_QUAL = 1, /* the reference would raise here: an empty quality line reached
                                qual_min_max with no valid history (min() of an empty seq,
                                src/fq_meta.nim:102) */
  FQGPU_META_INCOMPLETE = 0x100 /* multi-GPU shards only (flag, OR-ed in): the sampled first meta_records
                                   records extend beyond shard 0, whose range is all that is reported */
};

/* Everything is an integer.  Line classes follow src/fq_count.nim:39-42: with the 1-based line
 * counter i, i mod 4 == 1 is a header line (reads++), i mod 4 == 2 a sequence line; we call
 * i mod 4 == 0 the quality line (src/fq_meta.nim:245 uses the 0-based i %% 4 == 3).  No '@'/'+'
 * validation is performed, exactly like the reference.  "Line" = Nim streams.lines semantics:
 * split at '\n', one '\r' directly before the '\n' is dropped, a trailing unterminated non-empty
 * line counts, blank lines count. */
typedef struct {
  uint64_t bytes;     /* bytes scanned                                                          */
  uint64_t lines;     /* lines yielded (== final value of i, src/fq_count.nim:39)               */
  uint64_t reads;     /* n_reads   src/fq_count.nim:40-41                                       */
  uint64_t bases;     /* total_len src/fq_count.nim:45                                          */
  uint64_t gc_bases;  /* gc_cnt    src/fq_count.nim:43  (uppercase 'G' + 'C' only)              */
  uint64_t n_bases;   /* n_cnt     src/fq_count.nim:44  (uppercase 'N' only)                    */
  uint64_t seq_lines; /* number of sequence lines (including empty ones)                        */
  uint64_t qual_lines;
  /* --- superset named by the north star; no reference oracle, defined by oracle/fq_oracle.c --- */
  uint64_t base_counts[256]; /* byte histogram over sequence lines (A/C/G/T/N/... fall out)     */
  uint64_t qual_counts[256]; /* byte histogram over quality lines                               */
  uint64_t seq_len_min, seq_len_max;   /* over sequence lines; min = UINT64_MAX, max = 0 if none */
  uint64_t qual_len_min, qual_len_max; /* over quality lines                                     */
  uint64_t seq_len_hist[FQGPU_POS_BINS + 1];  /* exact length histogram, last bin = len >= POS_BINS */
  uint64_t qual_len_hist[FQGPU_POS_BINS + 1];
  uint64_t seq_len_log2[FQGPU_LEN_LOG2_BINS]; /* bin 0: len 0; bin k: 2^(k-1) <= len < 2^k      */
  uint64_t qual_pos_sum[FQGPU_POS_BINS + 1];  /* sum of raw quality bytes at 0-based position p  */
  uint64_t qual_pos_cnt[FQGPU_POS_BINS + 1];  /* quality lines that have a byte at position p;
                                                 last bin = number of bytes at p >= POS_BINS     */
  /* --- fq-meta quality-range scan, src/fq_meta.nim:207-208,226-248,277 --- */
  int64_t meta_qual_min;  /* qual_min after the loop (ord-33 units, -1 = none/invalid)          */
  int64_t meta_qual_max;  /* qual_max                                                            */
  uint64_t meta_lines;    /* i after the loop: min(4*meta_records, lines); n_lines = meta_lines/4 */
  uint32_t meta_status;   /* FQGPU_META_*                                                        */
  uint32_t reserved;
} fqgpu_stats;

/* Version / capability probes (no GPU needed). */
int fqgpu_abi_version(void);
size_t fqgpu_stats_size(void);
const char* fqgpu_build_info(void);
int fqgpu_device_count(void); /* CUDA devices visible, <= 0 if none */

/* Context = one GPU + its staging ring + its device-resident counters and carry state. */
int fqgpu_create(fqgpu_ctx** out, const fqgpu_config* cfg); /* cfg NULL or zeroed = defaults */
void fqgpu_destroy(fqgpu_ctx* ctx);
const char* fqgpu_last_error(const fqgpu_ctx* ctx); /* ctx may be NULL: last create() failure */

/* Streaming interface (replaces the `for line in lines(stream)` loop, src/fq_count.nim:38).
 * acquire: next free PINNED host chunk (blocks while every chunk is in flight); the host reads
 *          file bytes or gzread()s (gzip_stream.nim:16-17) straight into it.
 * submit:  async H2D + scan kernel on the context's stream; a chunk may end anywhere (mid-line,
 *          between '\r' and '\n'); the carry is resolved on the device.
 * finish:  flush the carry (trailing unterminated line), reduce the per-CTA partial counters,
 *          copy the result (~20 KB) back.  reset: ready for the next file. */
void* fqgpu_acquire(fqgpu_ctx* ctx, size_t* capacity);
int fqgpu_submit(fqgpu_ctx* ctx, void* chunk, size_t nbytes);
int fqgpu_finish(fqgpu_ctx* ctx, fqgpu_stats* out);
int fqgpu_reset(fqgpu_ctx* ctx);

/* Continues the stream with `nbytes` of HOST memory (H2D straight from the caller's buffer, asynchronous
 * when it is pinned; the copy of chunk k+1 overlaps the scan of chunk k).  No reset, no finish. */
int fqgpu_scan_host(fqgpu_ctx* ctx, const void* buf, size_t nbytes);

/* Whole-buffer conveniences on top of the streaming interface. */
int fqgpu_count_host(fqgpu_ctx* ctx, const void* buf, size_t nbytes, fqgpu_stats* out); /* pageable host memory */
int fqgpu_count_file(fqgpu_ctx* ctx, const char* path, fqgpu_stats* out); /* plain or .gz (zlib inflate on host); FQGPU_EIO only when the
   file cannot be OPENED (exit 2 in the reference); a gz stream that turns out truncated / corrupt ends where zlib stops, like the reference's */
/* Same with the stream kind chosen by the caller: fq_count picks gz by a case-SENSITIVE ".gz" suffix
 * (src/fq_count.nim:31), fq_meta by a case-INsensitive one (src/fq_meta.nim:219). */
int fqgpu_count_file_as(fqgpu_ctx* ctx, const char* path, int as_gz, fqgpu_stats* out);

/* fq-meta reads only the head of the file: `while not stream.atEnd() and i < sample_n * 4: line = stream.readLine()`
 * (src/fq_meta.nim:226-248).  Same as fqgpu_count_file_as, but the stream ends with the line that completes
 * 4 * cfg.meta_records lines (the host counts newlines while it fills the pinned chunk), so the cost is that of the
 * sampled head, not of the file; every field of `out` describes that head.  meta_records == 0 reads nothing. */
int fqgpu_meta_file_as(fqgpu_ctx* ctx, const char* path, int as_gz, fqgpu_stats* out);

/* .gz input that is BGZF (blocked gzip: bgzip / htslib and many sequencer pipelines write it; every member is at
 * most 64 KiB and carries its sizes) is not inflated through zlib on the host (gzip_stream.nim:16-17) but on the
 * device, one thread per member, straight into the buffer the scan reads (csrc/fq_bgzf.cu).
 * Plain gzip, or anything malformed, takes the zlib path.  FQGPU_NO_BGZF=1 in the environment disables the device
 * path.  Returns the members the last fqgpu_count_file* call on this context inflated on the device. */
unsigned long long fqgpu_bgzf_members(const fqgpu_ctx* ctx);
/* Ordinary (single-member or concatenated) gzip that is not BGZF is inflated on the device too: chunks of the one
 * DEFLATE stream in parallel from guessed block starts, the guesses proven by the chunk before landing on them
 * (csrc/fq_gzip.cu).  What cannot be proven goes through gzread() as before (FQGPU_NO_GZIP_DEVICE=1 forces that).
 * Diagnostics since the last reset: chunks inflated on the device (0 = the host path ran), and block starts the
 * search proposed that were not block boundaries. */
unsigned long long fqgpu_gzip_chunks(const fqgpu_ctx* ctx);
unsigned long long fqgpu_gzip_false_starts(const fqgpu_ctx* ctx);
/* batches of the gzip path whose symbols did not fit the one-pass arena and were decoded a second time */
unsigned long long fqgpu_gzip_second_passes(const fqgpu_ctx* ctx);

/* Many files at once (replaces the sequential `for fastq in files` loop of sc.nim:115-116).
 * Files are independent streams, so up to n_threads host threads (0 = min(n, 8)) each own a private context --
 * staging ring, stream, device-resident counters -- and count one file at a time: the host-side work of one
 * file (read() or the zlib inflate of gzip_stream.nim:16-17, the ceiling of the .gz path) overlaps the others'
 * and the GPU scans of all of them interleave.  cfg->device == FQGPU_DEVICE_ALL spreads the contexts round-robin
 * over every visible GPU (one file per GPU instead of byte ranges).  as_gz: per-file stream kind, or NULL =
 * by a case-sensitive ".gz" suffix (src/fq_count.nim:31).  out[i] and rc[i] are filled for every file; the return
 * value is the first non-OK rc in FILE ORDER, which is where the sequential loop would have stopped. */
#define FQGPU_DEVICE_ALL (-2)
int fqgpu_count_files(const fqgpu_config* cfg, const char* const* paths, const int* as_gz, int n, int n_threads,
                      fqgpu_stats* out, int* rc);

/* Paired files (R1 / R2 of one library) as one job: both mates are scanned at the same time (two contexts), rows in
 * argument order exactly as two iterations of the reference's loop (sc.nim:115-116) would print them; *paired (may be
 * NULL) = the files have the same number of records and lines, i.e. can be mates.  Returns the first non-OK rc. */
int fqgpu_count_pair(const fqgpu_config* cfg, const char* r1, const char* r2, fqgpu_stats* out1, fqgpu_stats* out2, int* paired);

/* One uncompressed regular file, byte-range sharded over `world` contexts in ONE process (the "file mode" of SURVEY 8e;
 * the multi-process form is the fqgpu_shard_* protocol below, which this call drives itself).  devices[g] = CUDA
 * ordinal of shard g (an ordinal may repeat); NULL = round-robin over every visible device, world <= 0 = one shard
 * per visible device.  Shard g has its own host thread, pinned ring and context, reads bytes [g*N/world, (g+1)*N/world)
 * of the file, resyncs to the first record start and exports its block; the blocks are gathered in host memory and
 * combined exactly (fqgpu_shard_combine_host).  Gzip input, pipes, small files and input on which a shard's phase
 * hypothesis fails (malformed FASTQ) are counted by one context on devices[0] instead, so the result is always
 * the one fqgpu_count_file_as gives. */
int fqgpu_count_file_sharded(const fqgpu_config* cfg, const char* path, const int* devices, int world, fqgpu_stats* out);

/* HBM-resident interface (kernel-only measurements; data already on the context's device).
 * scan_device may be called repeatedly: each call continues the same stream of bytes (carry
 * state is kept on the device), exactly as if the buffers were concatenated. */
int fqgpu_scan_device(fqgpu_ctx* ctx, const void* dptr, size_t nbytes);
int fqgpu_count_device(fqgpu_ctx* ctx, const void* dptr, size_t nbytes, fqgpu_stats* out); /* reset+scan+finish */

/* Multi-GPU byte-range sharding: rank g of `world` scans bytes [g*N/world,(g+1)*N/world)
 * of ONE logical stream.  A rank > 0 does not know the line phase of its first byte, so it
 * resynchronises on the first true record start ('@' line whose line+2 starts with '+'), scans with
 * that guessed phase, keeps the statistics of its head fragment (bytes before its first newline)
 * and of its open tail detached, and exports everything as one block of uint64 words:
 *   fqgpu_shard_begin():   reset the context as shard `rank` of `world`
 *   [fqgpu_scan_device / fqgpu_submit ... as usual]
 *   fqgpu_shard_export():  write this rank's block into slot `rank` of a device buffer of
 *                          world*fqgpu_shard_block_words() uint64 that the caller zeroed
 *   [caller SUM-all-reduces the buffer: ncclAllReduce / torch.distributed -- the ONE collective]
 *   fqgpu_shard_combine(): every rank folds the world blocks into the final stats: stitches
 *                          tail(g-1)+head(g) lines, shifts head per-position sums, verifies every
 *                          guessed phase against the exact line counts.  Returns FQGPU_OK, or
 *                          FQGPU_ERETRY when some rank guessed wrong (malformed input): then each
 *                          rank calls fqgpu_shard_rescan() (the first wrong rank scans its range again
 *                          with the exact carry taken from the blocks), every rank exports again, and
 *                          all-reduce/combine repeat -- at most world-1 rounds. */
#define FQGPU_ERETRY 1
size_t fqgpu_shard_block_words(void);
int fqgpu_shard_begin(fqgpu_ctx* ctx, int rank, int world);
int fqgpu_shard_export(fqgpu_ctx* ctx, uint64_t* d_blocks);
int fqgpu_shard_combine(fqgpu_ctx* ctx, const uint64_t* d_blocks, fqgpu_stats* out);
/* After FQGPU_ERETRY: FQGPU_OK = this rank's block stands, just export it again; FQGPU_ERETRY = the exact
 * carry in front of this rank has been installed, scan the rank's byte range again, then export. */
int fqgpu_shard_rescan(fqgpu_ctx* ctx, const uint64_t* d_blocks);
/* The combine step alone, over gathered blocks in HOST memory (pure host arithmetic, needs no GPU). */
int fqgpu_shard_combine_host(int world, const uint64_t* h_blocks, uint64_t meta_records, fqgpu_stats* out);

/* The collective inside the library (no NCCL, no host arithmetic): every rank owns an EXCHANGE BUFFER in device memory
 * that the other ranks open (CUDA IPC between processes, plain pointers inside one process).  One launch per rank then
 * packs the rank's block, stores it into its slot of every rank's buffer through peer-mapped memory (NVLink / NVSwitch
 * P2P stores), publishes a flag, waits for the other ranks' flags in its own memory and combines the gathered blocks on
 * the device (same arithmetic as fqgpu_shard_combine_host); the host reads back one block: one kernel and one host
 * synchronisation per step.
 *   fqgpu_shard_exchange_create(): allocate this rank's buffer; ipc_handle_out (fqgpu_ipc_handle_bytes() bytes, may be
 *                                  NULL) receives the handle the other processes need
 *   fqgpu_shard_exchange_open():   handles of all ranks, rank-major (the caller all-gathers them once -- any transport)
 *   fqgpu_shard_exchange_set_peers(): the same inside one process: device pointers (fqgpu_shard_xbuf) of all ranks
 *   per step: fqgpu_shard_begin, scans, then fqgpu_shard_exchange_start() on EVERY rank (asynchronous: no rank can
 *             finish before all have started) and fqgpu_shard_exchange_finish() -> FQGPU_OK, FQGPU_ERETRY (a wrong
 *             hypothesis: fqgpu_shard_rescan(ctx, fqgpu_shard_gathered(ctx)) as above) or FQGPU_ENCCL (a rank is missing) */
size_t fqgpu_ipc_handle_bytes(void);
size_t fqgpu_shard_xbuf_bytes(int world);
int fqgpu_shard_exchange_create(fqgpu_ctx* ctx, int rank, int world, void* ipc_handle_out);
int fqgpu_shard_exchange_open(fqgpu_ctx* ctx, const void* ipc_handles);
void* fqgpu_shard_xbuf(fqgpu_ctx* ctx);
int fqgpu_shard_exchange_set_peers(fqgpu_ctx* ctx, void* const* xbufs);
int fqgpu_shard_exchange_start(fqgpu_ctx* ctx);
int fqgpu_shard_exchange_finish(fqgpu_ctx* ctx, fqgpu_stats* out);
int fqgpu_shard_exchange_combine(fqgpu_ctx* ctx, fqgpu_stats* out); /* start + finish */
const uint64_t* fqgpu_shard_gathered(fqgpu_ctx* ctx);
void fqgpu_shard_exchange_destroy(fqgpu_ctx* ctx);

/* Timing of the most recent scan launches on this context (CUDA events on the context's stream):
 * kernel_ms = sum of scan-kernel durations since the last reset, launches = how many kernels. */
int fqgpu_last_timing(fqgpu_ctx* ctx, double* kernel_ms, uint64_t* launches);

/* Raw CUDA stream of the context (cudaStream_t as void*), for callers that want to order their
 * own work (e.g. a collective) after the scan without a host sync. */
void* fqgpu_stream(fqgpu_ctx* ctx);

/* Record-offset index as a product (the stand-alone output of the boundary classification,
 * north-star kernel 1).  Record k is the line with index 4k under the reference's line rule
 * (src/fq_count.nim:38-42: '\n'-terminated lines, a non-empty unterminated last line counts).
 * d_offsets[k] (device memory, capacity `cap` entries) receives the byte offset of record k's first byte for
 * k < min(*n_records, cap); *n_records is the number of records of the buffer (== fqgpu_stats.reads). */
int fqgpu_index_device(fqgpu_ctx* ctx, const void* dptr, size_t nbytes, uint64_t* d_offsets, uint64_t cap,
                       uint64_t* n_records);
/* The first `n` header lines (records 0..n-1 of the index) gathered to HOST memory for fq-meta's
 * sequencer / barcode detection (src/fq_meta.nim:229-242 reads them with readLine): row k of h_out
 * (`stride` bytes per row) holds the line without its '\n' (and without one '\r' directly before it),
 * truncated to `stride` bytes; h_len[k] = bytes stored. */
/* Lines of the buffer of the most recent fqgpu_index_device call (src/fq_dedup.nim:50 needs `lines div 4`). */
uint64_t fqgpu_index_lines(fqgpu_ctx* ctx);
int fqgpu_headers_device(fqgpu_ctx* ctx, const void* dptr, size_t nbytes, const uint64_t* d_offsets, uint64_t n,
                         uint32_t stride, uint8_t* h_out, uint32_t* h_len);

/* Duplicate read IDs (src/fq_dedup.nim:14-84): d_keep[k] (device, one byte per record of the
 * index) = 1 when record k is the first one with its header line, 0 when an earlier record has the same header
 * (compared as the reference does: the line without '\n' and without one '\r' before it).  *n_dups = dropped
 * records (the reference's "duplicates" line).  Exact: hashes only group the candidates, bytes decide. */
int fqgpu_dedup_device(fqgpu_ctx* ctx, const void* dptr, size_t nbytes, const uint64_t* d_offsets, uint64_t n_records,
                       uint8_t* d_keep, uint64_t* n_dups);
/* The same for a FASTQ in HOST memory (H2D copy, index, marks, flags back): what `sc fq-dedup` calls.  h_keep
 * receives min(*n_records, cap) flags; *n_lines = lines of the input. */
int fqgpu_dedup_host(fqgpu_ctx* ctx, const void* host, size_t nbytes, uint8_t* h_keep, uint64_t cap,
                     uint64_t* n_records, uint64_t* n_lines, uint64_t* n_dups);

/* Synthetic FASTQ generators, counter-based so any byte range can be
 * produced independently on any GPU; used by bench.py and the parity tests.  `first_record` lets a
 * rank generate its own shard.  Both write whole records only and return the bytes written. */
int fqgpu_synth_illumina(fqgpu_ctx* ctx, void* dptr, size_t capacity, uint64_t first_record,
                         uint64_t n_records, uint64_t seed, size_t* bytes_written);
/* The generator's own tallies: what the scan must report for records [first_record, first_record + n_records) of the
 * Illumina stream as one whole stream (every field of fqgpu_stats, the fq-meta fields for the context's meta_records),
 * derived from the random numbers alone -- no byte is generated or read.  An independent check of the scan at any size
 * (bench.py asserts it on the 36 GB run); oracle/fq_synth_twin.c is its CPU twin. */
int fqgpu_synth_illumina_tally(fqgpu_ctx* ctx, uint64_t first_record, uint64_t n_records, uint64_t seed, fqgpu_stats* out);
int fqgpu_synth_ont(fqgpu_ctx* ctx, void* dptr, size_t capacity, uint64_t first_record,
                    uint64_t n_records, uint64_t seed, size_t* bytes_written);
/* Arbitrary byte range [first_byte, first_byte+nbytes) of the Illumina stream (a shard may start
 * mid-record). */
int fqgpu_synth_illumina_bytes(fqgpu_ctx* ctx, void* dptr, uint64_t first_byte, uint64_t nbytes, uint64_t seed);

#ifdef __cplusplus
}
#endif
#endif /* FQGPU_H */
