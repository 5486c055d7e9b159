// fq_scan.cu -- the FASTQ scanning hot path on sm_100a.
//
// Replaces the per-line loop of the reference (src/fq_count.nim:38-45: `for line in lines(stream)`,
// i mod 4 classing, count("G")+count("C"), count("N"), line.len) with one pass over the bytes; the quality fold
// of src/fq_meta.nim:245-246 is fq_meta.cu, beside it on a second stream.
//
// A launch cuts its byte range into SPANS, one CTA each (2 CTAs per SM resident; large launches get several spans
// per resident CTA).  A span is the run of whole lines that START inside its nominal byte range: the CTA finds the
// first line start at or after the range start itself (the first newline from there on) and runs past the range end
// to the end of its last line, so no line is ever shared between CTAs and no CTA ever waits for another.  What a
// span cannot know is the line PHASE (line number mod 4) of its first line: it is guessed from the content (the
// first line that starts with '@' and whose line+2 starts with '+' is a header) and the CTA that exits last verifies
// every guess against the exact newline counts of the spans before it; a span whose guess was wrong (malformed
// input: the reference classes lines purely by their number) is redone by the second pass, which first takes the
// wrong counts back.  Results are therefore exact on any input.
//
// Inside a span the CTA streams 32 KiB tiles through shared memory (1-D TMA, cp.async.bulk + mbarrier, two stages:
// the copy of the next tile runs under the work on this one), all warps moving through the phases together:
//
//   A   boundary classification: coalesced 16-byte groups -> 16-bit '\n' masks (SWAR compare, IDP.4A movemask)
//       -> the tile's newline bitmap in shared memory.
//   B   every thread owns 64 CONSECUTIVE bytes of the bitmap: popc, ONE block-wide prefix of the counts, then the
//       thread stores the offsets of its own newlines at their exact slots of the tile's newline index nl[].
//       The running (lines, open-line bytes) of the span is the tile-edge record carry.
//   L   line tasks, one thread per sequence / quality line with bytes in the tile (classes from the line number):
//       the line's first / last ragged 16-byte groups are counted right here under byte masks, its aligned full
//       groups are left to the workers as one record (first group, groups, line position); line-length tables and
//       the '\r' rule at the line's newline.
//   W   workers: slot x of a class maps to (line x / K, group x % K), K = the tile's largest group count, so every
//       lane takes one aligned 16-byte group (LDS.128) of some line: histogram by one IDP.4A (address) + one shared
//       atomic (ATOMS.POPC.INC, which merges equal addresses: un-striped 256-bin tables) per byte; per-position
//       quality sums as eight packed 16-bit-pair atomics into a bank-skewed table (even / odd position tables, no
//       alignment shifts).  Lines of more than 63 groups (long reads) are taken by all threads together.
//
// Counter tables live in shared memory for the life of the CTA and are added to the context's block by 64-bit
// atomics (K3); the CTA that exits last advances the stream carry (fq::Carry): chunk-edge carry on the device.
// Every input byte is read from HBM once (a span reads a few bytes of its successor's range to finish its last
// line).  Line semantics are Nim's streams.lines: split at '\n', drop one '\r' directly before it; the trailing
// unterminated line is accounted by the host from fq::Carry at finish().
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fq_dev.cuh"

namespace fq {

#ifndef FQ_THREADS
#define FQ_THREADS 512
#endif
#ifndef FQ_CTAS
#define FQ_CTAS 2
#endif
constexpr int THREADS = FQ_THREADS;
constexpr int CTAS_PER_SM = FQ_CTAS;
constexpr int NWARPS = THREADS / 32;
constexpr int GPT = 4;                        // 16-byte groups per thread and tile
constexpr int NG = THREADS * GPT;             // groups per tile
constexpr int TILE = NG * 16;                 // 32 KiB
constexpr int NSTAGE = 2;
constexpr int NL_CAP = THREADS * 4;            // newline index; tiles with more newlines take the byte walker
constexpr int REC_CAP = NL_CAP / 4 + 8;       // lines of one class with bytes in a tile
constexpr int KMAX = 63;                      // lines with more full groups are "long"
constexpr int LONG_CAP = TILE / (16 * (KMAX + 1)) + 2;
constexpr int OPEN_CLIP = 1 << 30;            // line positions saturate here (>= POS_BINS is the overflow bin anyway)
constexpr int MIN_SPAN_TILES = 32;            // spans are at least this long (1 MiB)
#ifndef FQ_WAVE_SPAN_TILES
#define FQ_WAVE_SPAN_TILES 512
#endif
constexpr int WAVE_SPAN_TILES = FQ_WAVE_SPAN_TILES;  // ... and 16 MiB before a launch gets more spans than resident CTAs
// Per-position quality sums, 16-bit pairs.  Q = position + 16.  Even Q: pair A = Q >> 1 of the EVEN table holds
// (Q, Q+1); odd Q: pair A = (Q+1) >> 1 of the ODD table holds (Q, Q+1).  Pair A lives in cell (r, c) with
// 8 c + r = A, r < 8, at word r * PT_STRIDE + c; a group adds its eight pairs at immediate offsets PT_STRIDE * i,
// so rows 8..15 alias (r - 8, c + 1) and consecutive groups of a line hit consecutive banks.  Two copies
// (alternating with the record parity) keep neighbouring records apart.
constexpr int PT_STRIDE = 33;
constexpr int PT_COPY = 16 * PT_STRIDE;       // 528 words
constexpr int PT_WORDS = 4 * PT_COPY;         // [odd][copy]
// The second round of the full-group slots starts behind the warps that ran the line tasks (about 3/8 of the CTA); the
// ragged ends of the sequence lines go to those warps, the ragged ends of the quality lines to the ones behind them.
constexpr uint32_t ROT_FULL = (THREADS * 3 / 8) & ~31u, ROT_SPART = 0, ROT_QPART = (THREADS * 3 / 8) & ~31u;
constexpr uint32_t FLUSH_BYTES = 4u << 20;    // 32-bit tables are added to the block at least this often

struct ScanArgs {
  const uint8_t* base;  // 16-byte aligned; the launch covers bytes [lo0, end) relative to base
  uint32_t lo0;
  u64 end;
  uint32_t ntiles, tps, nspans;  // tiles, tiles per span (nominal), spans
  Carry* carry;
  ShardInfo* shard;  // detached head of a multi-GPU shard (rank > 0)
  u64* acc;          // [BLOCK_WORDS] the context's counter block
  SpanDesc* desc;    // [nspans]
  u64* ctl;          // [CTL_WORDS]
  uint32_t unknown;  // shard with an unknown start (fqgpu_shard_begin, rank > 0, not rescanned)
  uint32_t pass;     // 0: every span under its guessed phase; 1: only the spans whose guess was wrong, exactly
};

struct Sel { uint32_t h0, h1, h2, h3; };  // IDP.4A selectors 4 << 8k: histogram address = byte_k * 4 + base

// The span's running state (shared memory; advanced by thread 0 at the end of every tile).
struct Run {
  u64 open;          // bytes of the open line before the tile
  u64 totalT;        // newlines of the span so far
  u64 proc_end;      // offset behind the last byte processed
  uint32_t cnt;      // lines before the tile (low bits; shards with an unknown start: under the hypothesis)
  uint32_t seen;     // a newline has been seen in the stream before the tile
  int prev_byte;     // byte before the tile's first valid byte (0x100: none)
  uint32_t prev_counted;  // ... and it was scanned by this launch (a '\r' there has been counted)
};
struct TileInfo {
  uint32_t T;        // newlines of the tile
  uint32_t first;    // extension tiles: offset of the first newline
  uint32_t K[2];     // most full groups of a (not long) line, per class
  uint32_t nlong;
  uint32_t last1;    // walker tiles: offset of the last newline + 1
};

struct __align__(128) Smem {
  uint8_t buf[NSTAGE][TILE];
  uint16_t bitmap[NG];               // bit b of entry g: byte 16 g + b is '\n'
  uint16_t nl[NL_CAP + 8];           // offsets of the tile's newlines, ascending
  u64 rec[2][REC_CAP];               // per class and line: first full group | full groups << 11 | fast << 17 | table cell << 18 | parity << 30 | position << 32
  uint32_t part[2][2 * REC_CAP];     // per class and line: its two ragged groups: group | lo << 11 | hi << 15 | parity << 20 | position of byte lo << 21
  u64 longl[LONG_CAP];               // lines of more than KMAX full groups: first | groups << 11 | class << 23 | parity << 24 | position << 32
  uint32_t inv[KMAX + 1];            // ceil(2^32 / K)
  uint32_t hist[2][256];             // [0] sequence, [1] quality
  uint32_t ptab[PT_WORDS];
  uint32_t pos32[POS_BINS + 2];      // per-position sums, 32-bit (second level of ptab; generic paths)
  uint32_t seq_len[POS_BINS + 2];
  uint32_t qual_len[POS_BINS + 2];
  uint32_t seq_log2[LOG2_BINS];
  uint4 masks[17];                   // masks[n]: 0xFF in the first n bytes of a group
  uint32_t wtot[NWARPS];             // newlines per warp, then their exclusive prefix
  int wlast[NWARPS];                 // walker tiles: offset of the last newline of the warps before, + 1
  u64 full_bar[NSTAGE];
  u64 big_min[2], big_max[2];        // line lengths >= 2^32
  u64 over;                          // quality bytes at positions >= POS_BINS
  uint32_t len_min[2], len_max[2];   // [0] seq, [1] qual
  uint32_t junk[2];                  // masked bytes counted in bin 0
  uint32_t ksel[4];
  uint32_t hiflag, flag;
  Run run;
  TileInfo ti;
};
static_assert(sizeof(Smem) <= (CTAS_PER_SM == 2 ? 115712 : 76800), "CTAs per SM");


// ---------------------------------------------------------------------------------------------
// 16 bytes into the histogram at hbase: one IDP.4A and one shared atomic per byte
__device__ __forceinline__ void hist16(const Sel& k, const uint4& v, uint32_t hbase) {
  red_inc(__dp4a(v.x, k.h0, hbase)); red_inc(__dp4a(v.x, k.h1, hbase)); red_inc(__dp4a(v.x, k.h2, hbase)); red_inc(__dp4a(v.x, k.h3, hbase));
  red_inc(__dp4a(v.y, k.h0, hbase)); red_inc(__dp4a(v.y, k.h1, hbase)); red_inc(__dp4a(v.y, k.h2, hbase)); red_inc(__dp4a(v.y, k.h3, hbase));
  red_inc(__dp4a(v.z, k.h0, hbase)); red_inc(__dp4a(v.z, k.h1, hbase)); red_inc(__dp4a(v.z, k.h2, hbase)); red_inc(__dp4a(v.z, k.h3, hbase));
  red_inc(__dp4a(v.w, k.h0, hbase)); red_inc(__dp4a(v.w, k.h1, hbase)); red_inc(__dp4a(v.w, k.h2, hbase)); red_inc(__dp4a(v.w, k.h3, hbase));
}

// Per-position sums of bytes [lo, hi) of a quality group (bytes outside are zero in v).  Q = 16 + line position of
// the group's byte 0 (>= 1).  Inside the table: eight packed pairs; beyond POS_BINS: the overflow bin; the group
// that straddles POS_BINS, and every group of a tile too dense for 16-bit halves: byte-wise into pos32.
__device__ __forceinline__ void pos16(Smem& sm, const uint4& v, uint32_t ga, uint32_t ptab_s, uint32_t Q, uint32_t lo, uint32_t hi,
                                      uint32_t par, bool dense, u64& over) {
  if (!dense && Q + hi <= (uint32_t)POS_BINS + 16u) {
    const uint32_t odd = Q & 1u;
    const uint32_t A = (Q + odd) >> 1;
    const uint32_t r0 = ptab_s + 4u * ((A & 7u) * PT_STRIDE + (A >> 3) + odd * (2u * PT_COPY) + par * PT_COPY);
    red_add_at<4 * PT_STRIDE * 0>(r0, __byte_perm(v.x, 0u, 0x4140)); red_add_at<4 * PT_STRIDE * 1>(r0, __byte_perm(v.x, 0u, 0x4342));
    red_add_at<4 * PT_STRIDE * 2>(r0, __byte_perm(v.y, 0u, 0x4140)); red_add_at<4 * PT_STRIDE * 3>(r0, __byte_perm(v.y, 0u, 0x4342));
    red_add_at<4 * PT_STRIDE * 4>(r0, __byte_perm(v.z, 0u, 0x4140)); red_add_at<4 * PT_STRIDE * 5>(r0, __byte_perm(v.z, 0u, 0x4342));
    red_add_at<4 * PT_STRIDE * 6>(r0, __byte_perm(v.w, 0u, 0x4140)); red_add_at<4 * PT_STRIDE * 7>(r0, __byte_perm(v.w, 0u, 0x4342));
  } else if (Q + lo >= (uint32_t)POS_BINS + 16u) {
    over += __dp4a(v.x, 0x01010101u, __dp4a(v.y, 0x01010101u, __dp4a(v.z, 0x01010101u, __dp4a(v.w, 0x01010101u, 0u))));
  } else {
    for (uint32_t x = lo; x < hi; x++) {
      const uint32_t b = lds8(ga + x), p = Q - 16u + x;
      if (p < (uint32_t)POS_BINS) atomicAdd(&sm.pos32[p], b);
      else over += b;
    }
  }
}

// ptab -> pos32 (position p = tid), then ptab is cleared.  Called by all threads; ends with the table zeroed but
// NOT yet synchronised (the next writer is behind a barrier).
__device__ __forceinline__ void flush_ptab(Smem& sm, int tid) {
  for (int p = tid; p < POS_BINS; p += THREADS) {
    const uint32_t Ae = (uint32_t)(p + 16) >> 1, Ao = (uint32_t)(p + 17) >> 1;
    const uint32_t ce = (Ae & 7u) * PT_STRIDE + (Ae >> 3), co = (Ao & 7u) * PT_STRIDE + (Ao >> 3);
    const uint32_t she = (p & 1) ? 16u : 0u, sho = (p & 1) ? 0u : 16u;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < 2; k++) {
      s += (sm.ptab[ce + k * PT_COPY] >> she) & 0xFFFFu;
      s += (sm.ptab[ce + 8 * PT_STRIDE - 1 + k * PT_COPY] >> she) & 0xFFFFu;
      s += (sm.ptab[2 * PT_COPY + co + k * PT_COPY] >> sho) & 0xFFFFu;
      s += (sm.ptab[2 * PT_COPY + co + 8 * PT_STRIDE - 1 + k * PT_COPY] >> sho) & 0xFFFFu;
    }
    if (s) atomicAdd(&sm.pos32[p], s);
  }
  __syncthreads();
  for (int i = tid; i < PT_WORDS; i += THREADS) sm.ptab[i] = 0;
}


// Everything the CTA holds in the 32-bit tables -> the context's block (64-bit atomics; sign < 0 takes the counts
// back: second pass), tables cleared.  The line-length extrema stay in shared memory.
__device__ __noinline__ void flush_tables(Smem& sm, u64* acc, int tid, u64& over, int sign) {
  __syncthreads();
  flush_ptab(sm, tid);
  if (over) { atomicAdd(&sm.over, over); over = 0; }
  __syncthreads();
  for (int i = tid; i < 512; i += THREADS) {
    uint32_t v = (&sm.hist[0][0])[i];
    if ((i & 255) == 0) v -= sm.junk[i >> 8];  // masked bytes were counted in bin 0
    if (v) atomicAdd(&acc[OFF_HIST_SEQ + i], sign > 0 ? (u64)v : 0ull - (u64)v);
    (&sm.hist[0][0])[i] = 0;
  }
  for (int i = tid; i <= POS_BINS; i += THREADS) {
    uint32_t v = sm.pos32[i];
    if (v) { atomicAdd(&acc[OFF_POS_SUM + i], sign > 0 ? (u64)v : 0ull - (u64)v); sm.pos32[i] = 0; }
    v = sm.seq_len[i];
    if (v) { atomicAdd(&acc[OFF_SEQ_LEN + i], sign > 0 ? (u64)v : 0ull - (u64)v); sm.seq_len[i] = 0; }
    v = sm.qual_len[i];
    if (v) { atomicAdd(&acc[OFF_QUAL_LEN + i], sign > 0 ? (u64)v : 0ull - (u64)v); sm.qual_len[i] = 0; }
  }
  if (tid < LOG2_BINS) {
    const uint32_t v = sm.seq_log2[tid];
    if (v) { atomicAdd(&acc[OFF_SEQ_LOG2 + tid], sign > 0 ? (u64)v : 0ull - (u64)v); sm.seq_log2[tid] = 0; }
  }
  __syncthreads();
  if (tid == 0) {
    if (sm.over) { atomicAdd(&acc[OFF_POS_SUM + POS_BINS], sign > 0 ? sm.over : 0ull - sm.over); sm.over = 0; }
    sm.junk[0] = sm.junk[1] = 0;
    sm.hiflag = 0;
  }
  __syncthreads();
}

// A counted line of raw length `rawlen` ends with the newline at tile offset `off`: the '\r' rule and the length
// tables.  ql: 0 sequence, 1 quality.  fn: the line started before the tile.
template <bool CORE>
__device__ __forceinline__ void line_end(Smem& sm, const ScanArgs& a, uint32_t buf_s, int vlo, u64 toff, int off, u64 rawlen, bool fn, uint32_t ql) {
  const int prev = off > vlo ? (int)lds8(buf_s + (uint32_t)off - 1u) : sm.run.prev_byte;
  const u64 cr = (rawlen > 0 && prev == '\r') ? 1 : 0;
  if (cr && (off > vlo || sm.run.prev_counted)) {  // the '\r' was counted as content where it stands: take it back
    // (not counted: the last byte of the previous launch, whose fate was left to this one)
    atomicSub(&sm.hist[ql]['\r'], 1u);
    if (!CORE && ql) {
      if (rawlen - 1 < (u64)POS_BINS) atomicSub(&sm.pos32[(uint32_t)(rawlen - 1)], (uint32_t)'\r');
      else atomicAdd(&sm.over, 0ull - '\r');
    }
  }
  const u64 len = rawlen - cr;
  if (a.unknown && fn && !sm.run.seen) {  // the shard's first line: its length is stitched by the combine step
    a.shard->head_len = rawlen;
    a.shard->head_cr = (unsigned)cr;
    a.ctl[CTL_FIRST_NL] = toff + (u64)off + 1;
    return;
  }
  const uint32_t bin = len < (u64)POS_BINS ? (uint32_t)len : (uint32_t)POS_BINS;
  if (ql) atomicAdd(&sm.qual_len[bin], 1u);
  else { atomicAdd(&sm.seq_len[bin], 1u); atomicAdd(&sm.seq_log2[log2_bin(len)], 1u); }
  if (len <= 0xFFFFFFFFull) { atomicMin(&sm.len_min[ql], (uint32_t)len); atomicMax(&sm.len_max[ql], (uint32_t)len); }
  else { atomicMin(&sm.big_min[ql], len); atomicMax(&sm.big_max[ql], len); }
}

// Phase A, rare forms (kept out of the hot code): this thread's groups again with the exact compare; an edge tile.
__device__ __noinline__ void phase_a_redo(uint32_t buf_s, uint32_t bm_s, int warp, int lane) {
  for (int k = 0; k < GPT; k++) {
    const uint32_t g = (uint32_t)(warp * (GPT * 32) + k * 32 + lane);
    sts16(bm_s + 2u * g, nl_mask16(lds128(buf_s + 16u * g)));
  }
}
__device__ __noinline__ uint32_t phase_a_edge(uint32_t buf_s, uint32_t bm_s, int lo, int hi, int warp, int lane) {
  uint32_t hib = 0;
  for (int k = 0; k < GPT; k++) {
    const int g = warp * (GPT * 32) + k * 32 + lane, off = g * 16;
    uint32_t m = 0;
    if (off < hi && off + 16 > lo) {
      const uint4 v = lds128(buf_s + 16u * (uint32_t)g);
      m = nl_mask16(v);
      int lo_k = lo - off; lo_k = lo_k < 0 ? 0 : lo_k;
      int hi_k = hi - off; hi_k = hi_k > 16 ? 16 : hi_k;
      m &= ((1u << hi_k) - 1u) & ~((1u << lo_k) - 1u);
      if (lo_k == 0 && hi_k == 16) hib |= (v.x | v.y) | (v.z | v.w);
      else {  // ragged group: only the valid bytes count (stale shared memory beside them)
        const uint32_t ww[4] = {v.x, v.y, v.z, v.w};
        for (int x = lo_k; x < hi_k; x++) hib |= (ww[x >> 2] >> (8 * (x & 3))) & 0x80u;
      }
    }
    sts16(bm_s + 2u * (uint32_t)g, m);
  }
  return hib;
}

// The byte walker: tiles with more newlines than the index holds.  Every thread takes its own 64 bytes one at a time
// (exact for any content); '\r' is counted where it stands and taken back by line_end like everywhere else.
// c / inc: newlines of this thread's run and their inclusive prefix inside the warp; wbase: newlines of the warps before.
template <bool CORE>
__device__ __noinline__ void phase_walk(Smem& sm, const ScanArgs& a, uint32_t buf_s, int vlo, int vhi, u64 toff, int tid, u64 w64, uint32_t c, uint32_t inc, uint32_t wbase) {
  const int lane = tid & 31, warp = tid >> 5;
  // offset of the last newline before this thread's run: inside the warp, else of the warps before, else none
  const uint32_t wl = (uint32_t)w64, wh = (uint32_t)(w64 >> 32);
  const int mylast = 64 * tid + (wh ? 63 - __clz(wh) : 31 - __clz(wl));
  const uint32_t hasb = __ballot_sync(0xffffffffu, c != 0);
  const uint32_t lower = hasb & lanemask_lt();
  const int lastW = __shfl_sync(0xffffffffu, mylast, (31 - __clz(lower)) & 31);
  const int warp_last = __shfl_sync(0xffffffffu, mylast, (31 - __clz(hasb)) & 31);
  if (lane == 0) sm.wlast[warp] = hasb ? warp_last + 1 : 0;
  __syncthreads();
  int last = -1;
  if (lower) last = lastW;
  else for (int w = warp - 1; w >= 0; w--) if (sm.wlast[w]) { last = sm.wlast[w] - 1; break; }
  uint32_t lc = wbase + inc - c;
  uint32_t cnt = sm.run.cnt + lc;
  u64 over = 0;
  for (int x = 0; x < 16 * GPT; x++) {
    const int off = 16 * GPT * tid + x;
    if (off < vlo || off >= vhi) continue;
    const uint32_t b = lds8(buf_s + (uint32_t)off);
    const uint32_t cls = cnt & 3u;
    const bool counted = CORE ? cls == 1u : (cls & 1u) != 0;
    const u64 pos = lc == 0 ? sm.run.open + (u64)(off - vlo) : (u64)(off - last - 1);
    if (b == '\n') {
      if (counted) line_end<CORE>(sm, a, buf_s, vlo, toff, off, pos, lc == 0, cls >> 1);
      cnt++; lc++;
      last = off;
    } else if (counted) {
      atomicAdd(&sm.hist[cls >> 1][b], 1u);
      if (!CORE && cls == 3u) {
        if (pos < (u64)POS_BINS) atomicAdd(&sm.pos32[(uint32_t)pos], b);
        else over += b;
      }
    }
  }
  if (over) atomicAdd(&sm.over, over);
  // the tile's last newline, for the running open-line length: kept where the index would have it
  __syncthreads();
  if (tid == 0) {
    int tl = 0;
    for (int w = NWARPS - 1; w >= 0; w--) if (sm.wlast[w]) { tl = sm.wlast[w]; break; }
    sm.ti.last1 = (uint32_t)tl;
  }
}

// Where a span starts and the line phase there.  The span owns the lines that start inside [N, Nend): its first byte
// is the successor of the first newline at offset >= N - 1.  Launch-relative: line j (j >= 1) starts after the j-th
// newline found; the first j whose line starts with '@' while line j+2 starts with '+' is a header, so the line at
// the span start has phase (1 - j) mod 4.  For valid 4-line FASTQ this is unambiguous (a quality line starting with
// '@' is followed two lines later by a sequence line, which cannot start with '+').  One warp.
// Returns the start offset (~0: no line starts inside the span); guess 0..3 or 4 = none.
constexpr int RESYNC_LINES = 40;
constexpr u64 RESYNC_BYTES = 4ull << 20;
__device__ __noinline__ u64 resync_span(const ScanArgs& a, u64 N, u64 Nend, int lane, uint32_t& guess_out) {
  uint32_t first[RESYNC_LINES + 1];
  int nlines = 0;
  uint32_t guess = 4;
  u64 start = ~0ull;
  const u64 o0 = (N - 1) & ~15ull;
  for (u64 o = o0; o < a.end && guess == 4 && nlines < RESYNC_LINES; o += 512) {
    if (start == ~0ull && o >= Nend) break;                     // no line starts inside the span
    if (start != ~0ull && o > start + RESYNC_BYTES) break;      // no guess within reach
    const u64 g = o + (u64)lane * 16;
    uint32_t m = 0;
    if (g < a.end) {
      const uint4 v = *reinterpret_cast<const uint4*>(a.base + g);
      m = nl_mask16(v);
      if (g + 16 > a.end) m &= (1u << (a.end - g)) - 1u;
      if (g < N - 1) m &= ~((1u << min((u64)16, N - 1 - g)) - 1u);
    }
    uint32_t any = __ballot_sync(0xffffffffu, m != 0);
    while (any && nlines < RESYNC_LINES) {
      const int src = __ffs(any) - 1;
      any &= any - 1;
      uint32_t mm = __shfl_sync(0xffffffffu, m, src);
      while (mm && nlines < RESYNC_LINES) {
        const int k = __ffs(mm) - 1;
        mm &= mm - 1;
        const u64 ls = o + (u64)src * 16 + (u64)k + 1;  // a line starts after this newline
        if (nlines == 0) {
          if (ls >= Nend || ls >= a.end) { guess_out = 4; return ~0ull; }
          start = ls;
        }
        nlines++;
        first[nlines] = ls < a.end ? (uint32_t)a.base[ls] : 0x100u;
        if (nlines >= 3 && first[nlines - 2] == '@' && first[nlines] == '+') {
          guess = (uint32_t)((1 - (nlines - 2)) & 3);
          break;
        }
      }
      if (guess != 4) break;
    }
  }
  guess_out = guess;
  return start;
}

// The guess at the start of a LAUNCH whose predecessor is on another GPU (shards with an unknown start): the phase of
// the line that contains the launch's first byte.
__device__ __noinline__ uint32_t resync_launch(const ScanArgs& a, int lane) {
  uint32_t g;
  const u64 start = resync_span(a, (u64)a.lo0 + 1, a.end, lane, g);  // first newline at or after the first byte
  (void)start;
  return g < 4 ? (g - 1u) & 3u : 4u;  // the line before the first line start
}

// The detached head of a shard with an unknown start (the bytes before the stream's first newline): the scan has
// added its per-position sums at positions relative to the shard start; they move from the block into
// ShardInfo::head_pos, which fqgpu_shard_combine shifts by the bytes the line had on the previous ranks.
// Run by the CTA that exits last.
__device__ __noinline__ void detach_head(const ScanArgs& a, int tid) {
  const u64 fl = ld_relaxed_gpu(a.ctl + CTL_HEAD);
  if (!(fl & HEAD_ACTIVE) || !(fl & HEAD_QUAL)) return;
  const u64 P0 = ld_relaxed_gpu(a.ctl + CTL_HEAD_P0);
  const u64 fnl = ld_relaxed_gpu(a.ctl + CTL_FIRST_NL);  // offset of the first newline + 1, 0 = none
  u64 ve = fnl ? fnl - 1 : a.end;                        // content end
  if (ve > (u64)a.lo0 && a.base[ve - 1] == '\r') ve--;   // dropped before the newline / left to the next launch at the end
  for (u64 o = (u64)a.lo0 + tid; o < ve; o += THREADS) {
    const u64 b = a.base[o], p = P0 + (o - a.lo0);
    const u64 bin = p < (u64)POS_BINS ? p : (u64)POS_BINS;
    atomicAdd(&a.shard->head_pos[bin], b);
    atomicAdd(&a.acc[OFF_POS_SUM + bin], 0ull - b);
  }
  if (tid == 0 && (fl & HEAD_PENDING_CR)) {  // the '\r' that ended the previous launch, counted by this one
    const u64 p = P0 - 1, bin = p < (u64)POS_BINS ? p : (u64)POS_BINS;
    atomicAdd(&a.shard->head_pos[bin], (u64)'\r');
    atomicAdd(&a.acc[OFF_POS_SUM + bin], 0ull - '\r');
  }
}

// Bytes [lo, hi) of group g belong to a counted line; posLo = line position of byte lo.
template <bool CORE>
__device__ __forceinline__ void masked_group(Smem& sm, const Sel& ksel, uint32_t buf_s, uint32_t hist_s, uint32_t ptab_s, uint32_t masks_s,
                                             uint32_t g, uint32_t lo, uint32_t hi, uint32_t posLo, uint32_t ql, uint32_t par, bool dense, u64& over) {
  const uint32_t ga = buf_s + 16u * g;
  uint4 v = lds128(ga);
  const uint4 ml = lds128(masks_s + 16u * lo), mh = lds128(masks_s + 16u * hi);
  v.x &= mh.x & ~ml.x; v.y &= mh.y & ~ml.y; v.z &= mh.z & ~ml.z; v.w &= mh.w & ~ml.w;
  hist16(ksel, v, hist_s + (ql << 10));
  atomicAdd(&sm.junk[ql], 16u - (hi - lo));
  if (!CORE && ql) pos16(sm, v, ga, ptab_s, posLo + 16u - lo, lo, hi, par, dense, over);
}
__device__ __forceinline__ uint32_t part_entry(uint32_t g, uint32_t lo, uint32_t hi, uint32_t posLo, uint32_t par) {
  return g | (lo << 11) | (hi << 15) | (par << 20) | ((posLo < 1023u ? posLo : 1023u) << 21);
}
// The eight packed pairs of a quality group whose table cell address is known (pos16 without the case analysis).
__device__ __forceinline__ void pos16_fast(const uint4& v, uint32_t r0) {
  red_add_at<4 * PT_STRIDE * 0>(r0, __byte_perm(v.x, 0u, 0x4140)); red_add_at<4 * PT_STRIDE * 1>(r0, __byte_perm(v.x, 0u, 0x4342));
  red_add_at<4 * PT_STRIDE * 2>(r0, __byte_perm(v.y, 0u, 0x4140)); red_add_at<4 * PT_STRIDE * 3>(r0, __byte_perm(v.y, 0u, 0x4342));
  red_add_at<4 * PT_STRIDE * 4>(r0, __byte_perm(v.z, 0u, 0x4140)); red_add_at<4 * PT_STRIDE * 5>(r0, __byte_perm(v.z, 0u, 0x4342));
  red_add_at<4 * PT_STRIDE * 6>(r0, __byte_perm(v.w, 0u, 0x4140)); red_add_at<4 * PT_STRIDE * 7>(r0, __byte_perm(v.w, 0u, 0x4342));
}

// One span under the line phase sm.run.cnt at its first byte `start`.  par_bits: mbarrier parities of the two stages
// (kept across calls).  Three CTA-wide barriers per tile; the copy of the next tile is started behind the first one
// (every thread is then done with the tile before this one, whose stage it overwrites).
template <bool CORE>
__device__ __forceinline__ void run_span(Smem& sm, const ScanArgs& a, const Sel& ksel, uint32_t sm0, int tid, u64 start, uint32_t t1,
                                         uint32_t& par_bits, u64& over, int sign) {
  const int lane = tid & 31, warp = tid >> 5;
  const uint32_t buf0_s = sm0 + (uint32_t)offsetof(Smem, buf);
  const uint32_t bar0_s = sm0 + (uint32_t)offsetof(Smem, full_bar);
  const uint32_t bm_s = sm0 + (uint32_t)offsetof(Smem, bitmap);
  const uint32_t nl_s = sm0 + (uint32_t)offsetof(Smem, nl);
  const uint32_t hist_s = sm0 + (uint32_t)offsetof(Smem, hist);
  const uint32_t ptab_s = sm0 + (uint32_t)offsetof(Smem, ptab);
  const uint32_t masks_s = sm0 + (uint32_t)offsetof(Smem, masks);
  const uint32_t rec_s = sm0 + (uint32_t)offsetof(Smem, rec);
  const uint32_t part_s = sm0 + (uint32_t)offsetof(Smem, part);
  const uint32_t t_first = (uint32_t)(start / TILE);
  uint32_t pt_lines = 0, since_flush = 0;  // (uniform) quality lines in ptab since its last flush; bytes since the tables' last flush
  int stage = 0;
  if (tid == 0) {
    const u64 toff = (u64)t_first * TILE;
    const uint32_t bytes = (uint32_t)((((a.end - toff) < (u64)TILE ? (a.end - toff) : (u64)TILE) + 15) & ~15ull);
    mbar_expect_tx(bar0_s, bytes);
    tma_load_1d(buf0_s, a.base + toff, bytes, bar0_s);
  }
  for (uint32_t tile = t_first;; tile++, stage ^= 1) {
    const u64 toff = (u64)tile * TILE;
    const int vlo = tile == t_first ? (int)(start - toff) : 0;
    int vhi = (a.end - toff) < (u64)TILE ? (int)(a.end - toff) : TILE;
    const uint32_t buf_s = buf0_s + (uint32_t)stage * TILE;
    const bool ext = tile >= t1;   // beyond the span's nominal range: only up to the first newline
    const bool have_next = tile + 1 < a.ntiles;
    mbar_wait(bar0_s + 8u * stage, (par_bits >> stage) & 1u);
    par_bits ^= 1u << stage;

    u64 w64;
    uint32_t c, inc, T, wbase;
#pragma unroll 1
    for (int round = 0;; round++) {
      // ---- A: newline masks -> bitmap.  Warp w classifies bytes [2048 w, 2048 w + 2048): coalesced 16-byte groups ----
      if (vlo == 0 && vhi == TILE) {
        const uint32_t g0 = (uint32_t)(warp * (GPT * 32) + lane);
        uint4 v[GPT];
#pragma unroll
        for (int k = 0; k < GPT; k++) v[k] = lds128(buf_s + 16u * (g0 + 32u * k));
        uint32_t hib = 0;
#pragma unroll
        for (int k = 0; k < GPT; k++) {
          hib |= (v[k].x | v[k].y) | (v[k].z | v[k].w);
          sts16(bm_s + 2u * (g0 + 32u * k), nl_mask16_ascii(v[k]));
        }
        if (hib & 0x80808080u) { phase_a_redo(buf_s, bm_s, warp, lane); sm.hiflag = 1; }  // the short compare is exact only for bytes < 0x80
      } else {
        if (phase_a_edge(buf_s, bm_s, vlo, vhi, warp, lane) & 0x80808080u) sm.hiflag = 1;
      }
      __syncwarp();
      // ---- B: this thread's 64 consecutive bytes of the bitmap (written by its own warp); prefix of the newline counts ----
      w64 = lds64(bm_s + 8u * (uint32_t)tid);
      c = (uint32_t)__popcll(w64);
      if (!__any_sync(0xffffffffu, c > 3u)) {  // counts of 0..3: the prefix from two ballots
        const uint32_t le = lanemask_lt() | (1u << lane);
        inc = (uint32_t)__popc(__ballot_sync(0xffffffffu, c & 1u) & le) + 2u * (uint32_t)__popc(__ballot_sync(0xffffffffu, c & 2u) & le);
      } else {
        inc = warp_incl_scan(c, lane);
      }
      if (lane == 31) sm.wtot[warp] = inc;
      __syncthreads();
      {
        const uint32_t e = lane < NWARPS ? sm.wtot[lane] : 0u;
        T = __reduce_add_sync(0xffffffffu, e);
        wbase = __reduce_add_sync(0xffffffffu, lane < warp ? e : 0u);
      }
      if (!ext || T == 0 || round == 1) break;
      // the span ends with the tile's first newline: classify the tile again up to it (rare: once per span)
      if (tid == 0) {
        uint32_t f = 0;
        for (int i = 0; i < NG / 4; i++) { const u64 w = reinterpret_cast<const u64*>(sm.bitmap)[i]; if (w) { f = 64u * i + (uint32_t)__ffsll((long long)w) - 1u; break; } }
        sm.ti.first = f;
      }
      __syncthreads();
      if ((int)sm.ti.first + 1 >= vhi) break;
      vhi = (int)sm.ti.first + 1;
      __syncthreads();
    }
    // the span ends behind the newline that closes the last line starting inside its nominal range
    const bool ends_nl = (lds16(bm_s + 2u * ((uint32_t)(vhi - 1) >> 4)) >> ((vhi - 1) & 15)) & 1u;
    const bool stop = (ext && T) || (!ext && tile + 1 == t1 && ends_nl) || !have_next;
    if (tid == 0 && have_next && !stop) {  // the copy of the next tile into the other stage: every thread has left the tile before this one
      const uint32_t bytes = (uint32_t)((((a.end - toff - TILE) < (u64)TILE ? (a.end - toff - TILE) : (u64)TILE) + 15) & ~15ull);
      mbar_expect_tx(bar0_s + 8u * (stage ^ 1), bytes);
      tma_load_1d(buf0_s + (uint32_t)(stage ^ 1) * TILE, a.base + toff + TILE, bytes, bar0_s + 8u * (stage ^ 1));
    }
    const uint32_t ex = wbase + inc - c;  // newlines of the tile before this thread's run
    const bool walker = T > (uint32_t)NL_CAP;
    const uint32_t line_bound = (T >> 2) + 2u;                      // quality lines with bytes in this tile, at most
    const uint32_t pt_limit = sm.hiflag ? 257u : 516u;              // 16-bit halves: lines a cell can take
    const bool dense = line_bound > pt_limit;
    const bool pt_flush = !CORE && !dense && pt_lines + line_bound > pt_limit;
    const uint32_t cnt_in = sm.run.cnt;
    if (pt_flush) { flush_ptab(sm, tid); pt_lines = 0; }
    if (!dense) pt_lines += line_bound;
    if (walker) {
      phase_walk<CORE>(sm, a, buf_s, vlo, vhi, toff, tid, w64, c, inc, wbase);
      __syncthreads();
    } else {
      // ---- B: the thread's newlines at their slots of the index ----
      if (c) {
        uint32_t slot = nl_s + 2u * ex, wl = (uint32_t)w64, wh = (uint32_t)(w64 >> 32);
        const uint32_t base = 64u * (uint32_t)tid;
        sts16(slot, base + (wl ? (uint32_t)__ffs(wl) - 1u : 31u + (uint32_t)__ffs(wh)));
        for (uint32_t k = 1; k < c; k++) {
          if (wl) wl &= wl - 1u; else wh &= wh - 1u;
          slot += 2u;
          sts16(slot, base + (wl ? (uint32_t)__ffs(wl) - 1u : 31u + (uint32_t)__ffs(wh)));
        }
      }
      const uint32_t step = CORE ? 4u : 2u;
      const uint32_t j0 = CORE ? (1u - cnt_in) & 3u : ((cnt_in & 1u) ? 0u : 1u);
      const uint32_t ntasks = j0 <= T ? (T - j0) / step + 1u : 0u;
      uint32_t nlines0, nlines1;  // lines per class with bytes in the tile
      if (CORE) { nlines0 = ntasks; nlines1 = 0; }
      else { const uint32_t qa = ((cnt_in + j0) & 3u) >> 1; nlines0 = qa ? ntasks >> 1 : (ntasks + 1u) >> 1; nlines1 = ntasks - nlines0; }
      if (tid == 0) { sm.ti.K[0] = sm.ti.K[1] = 0; sm.ti.nlong = 0; }
      __syncthreads();
      // ---- L: line tasks (lines j = 0..T of the tile; j = T is the open line behind the last newline) ----
      for (uint32_t i = (uint32_t)tid; i < ntasks; i += THREADS) {
        const uint32_t j = j0 + step * i;
        const int s = j > 0 ? (int)lds16(nl_s + 2u * (j - 1u)) + 1 : vlo;
        const int e = j < T ? (int)lds16(nl_s + 2u * j) : vhi;
        const uint32_t ln = cnt_in + j, ql = (ln >> 1) & 1u, par = (ln >> 2) & 1u;
        const uint32_t ord = CORE ? i : i >> 1;
        const bool fn = j == 0;
        uint32_t pos0 = 0;
        if (fn) { const u64 op = sm.run.open; pos0 = (uint32_t)(op < (u64)OPEN_CLIP ? op : (u64)OPEN_CLIP); }
        u64 rec = 0;
        uint32_t p0 = 0, p1 = 0;
        if (e > s) {
          const uint32_t gs = (uint32_t)s >> 4, ge = (uint32_t)e >> 4, ls = (uint32_t)s & 15u, le = (uint32_t)e & 15u;
          if (gs == ge) {
            p0 = part_entry(gs, ls, le, pos0, par);
          } else {
            uint32_t ff = gs;
            if (ls) { p0 = part_entry(gs, ls, 16u, pos0, par); ff++; }
            if (le) p1 = part_entry(ge, 0u, le, pos0 + (16u * ge - (uint32_t)s), par);
            const uint32_t nfull = ge - ff;
            const uint32_t qff = pos0 + (16u * ff - (uint32_t)s);
            if (nfull > (uint32_t)KMAX) {
              const uint32_t slot = atomicAdd(&sm.ti.nlong, 1u);
              if (slot < (uint32_t)LONG_CAP) sm.longl[slot] = (u64)(ff | (nfull << 11) | (ql << 23) | (par << 24)) | ((u64)qff << 32);
              else a.ctl[CTL_ERROR] = 2;
            } else if (nfull) {
              uint32_t lo32 = ff | (nfull << 11) | (par << 30);
              if (!CORE && ql && !dense && qff + 16u * nfull <= (uint32_t)POS_BINS) {  // every full group inside the packed table
                const uint32_t Q = qff + 16u, odd = Q & 1u, A = (Q + odd) >> 1;
                lo32 |= (1u << 17) | (((A & 7u) * PT_STRIDE + (A >> 3) + odd * (2u * PT_COPY) + par * PT_COPY) << 18);
              }
              rec = (u64)lo32 | ((u64)qff << 32);
              atomicMax(&sm.ti.K[ql], nfull);
            }
          }
        }
        if (ord < (uint32_t)REC_CAP) {
          sts64(rec_s + 8u * (ql * REC_CAP + ord), rec);
          sts64(part_s + 8u * (ql * REC_CAP + ord), (u64)p0 | ((u64)p1 << 32));
        } else a.ctl[CTL_ERROR] = 3;
        if (j < T) line_end<CORE>(sm, a, buf_s, vlo, toff, e, fn ? sm.run.open + (u64)(e - vlo) : (u64)(e - s), fn, ql);
      }
      __syncthreads();
      // ---- W: one lane per 16-byte group: the lines' full groups (slot x -> line x / K, group x % K), their ragged ends ----
#pragma unroll
      for (uint32_t ql = 0; ql < (CORE ? 1u : 2u); ql++) {
        const uint32_t nlines = ql ? nlines1 : nlines0;
        const uint32_t hb = hist_s + (ql << 10);
        uint32_t K = sm.ti.K[ql];
        if (K == 1) K = 2;  // (the reciprocal table starts at 2)
        if (K) {
          const uint32_t nslots = nlines * K, inv = sm.inv[K];
          // (the threads of the warps that ran the line tasks come last: the second round of slots goes to the others)
          for (uint32_t x = ((uint32_t)tid + THREADS - ROT_FULL) % THREADS; x < nslots; x += THREADS) {
            const uint32_t line = __umulhi(x, inv);
            const uint32_t k = x - line * K;
            const u64 r = lds64(rec_s + 8u * (ql * REC_CAP + line));
            const uint32_t r0 = (uint32_t)r;
            if (k < ((r0 >> 11) & 63u)) {
              const uint32_t ga = buf_s + 16u * ((r0 & 2047u) + k);
              const uint4 v = lds128(ga);
              hist16(ksel, v, hb);
              if (!CORE && ql) {
                if (r0 & (1u << 17)) pos16_fast(v, ptab_s + 4u * (((r0 >> 18) & 4095u) + k));
                else pos16(sm, v, ga, ptab_s, (uint32_t)(r >> 32) + 16u * k + 16u, 0u, 16u, (r0 >> 30) & 1u, dense, over);
              }
            }
          }
        }
        for (uint32_t x = ((uint32_t)tid + THREADS - (ql ? ROT_QPART : ROT_SPART)) % THREADS; x < 2u * nlines; x += THREADS) {
          const uint32_t ent = lds32(part_s + 4u * (2u * ql * REC_CAP + x));
          if (ent) masked_group<CORE>(sm, ksel, buf_s, hist_s, ptab_s, masks_s, ent & 2047u, (ent >> 11) & 15u, (ent >> 15) & 31u, ent >> 21, ql, (ent >> 20) & 1u, dense, over);
        }
      }
      {
        const uint32_t nlong = sm.ti.nlong < (uint32_t)LONG_CAP ? sm.ti.nlong : (uint32_t)LONG_CAP;
        for (uint32_t li = 0; li < nlong; li++) {
          const u64 e = sm.longl[li];
          const uint32_t e0 = (uint32_t)e, qff = (uint32_t)(e >> 32);
          const uint32_t ff = e0 & 2047u, nfull = (e0 >> 11) & 4095u, ql = (e0 >> 23) & 1u, par = (e0 >> 24) & 1u;
          for (uint32_t k = (uint32_t)tid; k < nfull; k += THREADS) {
            const uint32_t ga = buf_s + 16u * (ff + k);
            const uint4 v = lds128(ga);
            hist16(ksel, v, hist_s + (ql << 10));
            if (!CORE && ql) pos16(sm, v, ga, ptab_s, qff + 16u * k + 16u, 0u, 16u, par, dense, over);
          }
        }
      }
    }
    // ---- end of the tile: launch edges ('\r' whose successor lies in another launch), the running state (one thread of the
    // last warp, which has no line tasks) ----
    if (tid == THREADS - 32) {
      Run& r = sm.run;
      if (!r.prev_counted && tile == 0 && vlo == (int)a.lo0 && r.prev_byte == '\r' && r.open) {
        // the '\r' that ended the previous launch is content unless this launch starts with '\n'
        const uint32_t cls = cnt_in & 3u;
        const bool counted = CORE ? cls == 1u : (cls & 1u) != 0;
        if (counted && lds8(buf_s + (uint32_t)vlo) != '\n') {
          atomicAdd(&sm.hist[cls >> 1]['\r'], 1u);
          if (!CORE && cls == 3u) {
            const u64 p = r.open - 1;
            if (p < (u64)POS_BINS) atomicAdd(&sm.pos32[(uint32_t)p], (uint32_t)'\r'); else atomicAdd(&sm.over, (u64)'\r');
          }
        }
      }
      const int last1 = T ? (walker ? (int)sm.ti.last1 : (int)sm.nl[T - 1] + 1) : 0;  // offset of the last newline + 1
      const u64 open_out = T ? (u64)(vhi - last1) : r.open + (u64)(vhi - vlo);
      const u64 proc_end = toff + (u64)vhi;
      if (proc_end == a.end && lds8(buf_s + (uint32_t)vhi - 1u) == '\r') {  // counted where it stands: left to the next launch / finish()
        const uint32_t cls = (cnt_in + T) & 3u;
        const bool counted = CORE ? cls == 1u : (cls & 1u) != 0;
        if (counted) {
          atomicSub(&sm.hist[cls >> 1]['\r'], 1u);
          if (!CORE && cls == 3u) {
            const u64 p = open_out - 1;
            if (p < (u64)POS_BINS) atomicSub(&sm.pos32[(uint32_t)p], (uint32_t)'\r'); else atomicAdd(&sm.over, 0ull - '\r');
          }
        }
      }
      r.prev_byte = (int)lds8(buf_s + (uint32_t)vhi - 1u);
      r.prev_counted = 1;
      r.cnt = cnt_in + T;
      r.seen |= T != 0;
      r.open = open_out;
      r.totalT += T;
      r.proc_end = proc_end;
    }
    if (stop) break;
    since_flush += TILE;
    if (since_flush >= FLUSH_BYTES) { flush_tables(sm, a.acc, tid, over, sign); since_flush = 0; pt_lines = 0; }
  }
}

__device__ __forceinline__ u64 warp_incl_scan64(u64 v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const u64 n = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v += n;
  }
  return v;
}

// The CTA that exits last: exact line count in front of every span, verification of the guessed phases, the
// line-length extrema of the verified spans, the new stream carry.
__device__ __noinline__ void stitch(Smem& sm, const ScanArgs& a, int tid) {
  const int lane = tid & 31, warp = tid >> 5;
  u64* wsum = reinterpret_cast<u64*>(sm.buf[0]);        // scratch: the tile buffers are idle now
  u64* mm = wsum + 64;                                   // [0..1] min seq/qual, [2..3] max
  uint32_t* nredo = reinterpret_cast<uint32_t*>(mm + 8);
  const int per = ((int)a.nspans + THREADS - 1) / THREADS;
  const int s0 = tid * per, s1 = min((int)a.nspans, s0 + per);
  u64 mine = 0;
  for (int s = s0; s < s1; s++) mine += a.desc[s].T;
  const u64 inc = warp_incl_scan64(mine, lane);
  if (lane == 31) wsum[warp] = inc;
  if (tid == 0) { mm[0] = mm[1] = ~0ull; mm[2] = mm[3] = 0; *nredo = 0; }
  __syncthreads();
  u64 before = inc - mine, total = 0;
  for (int w = 0; w < NWARPS; w++) { const u64 x = wsum[w]; if (w < warp) before += x; total += x; }
  const Carry c = *a.carry;
  const unsigned fl = c.flags;
  const bool unknown = (fl & CARRY_UNKNOWN_START) != 0;
  const bool hyp_ok = !unknown || ((fl & CARRY_HYP_VALID) && !(fl & CARRY_HYP_FAILED));
  const uint32_t hyp = unknown ? (fl >> CARRY_HYP_SHIFT) & 3u : 0u;
  u64 G = c.lines + before;  // lines in front of span s0
  for (int s = s0; s < s1; s++) {
    SpanDesc& d = a.desc[s];
    const uint32_t exact = (uint32_t)((G + hyp) & 3u);
    d.exact = exact;
    // (a shard whose hypothesis failed is rescanned as a whole with the exact carry: nothing to verify)
    const bool ok = !d.nonempty || s == 0 || !hyp_ok || (d.guess_valid && d.guess == exact);
    d.state = ok ? SPAN_OK : SPAN_REDO;
    if (!ok) atomicAdd(nredo, 1u);
    else if (d.nonempty) {
      for (int q = 0; q < 2; q++) {
        if (d.len_min[q] != ~0ull) atomicMin(&mm[q], d.len_min[q]);
        if (d.len_max[q]) atomicMax(&mm[2 + q], d.len_max[q]);
      }
    }
    if (d.reached_end) mm[4] = d.end_open;  // exactly one span runs to the end of the launch
    G += d.T;
  }
  __syncthreads();
  if (tid == 0) {
    for (int q = 0; q < 2; q++) {
      if (mm[q] != ~0ull) atomicMin(&a.acc[q ? OFF_QUAL_LEN_MIN : OFF_SEQ_LEN_MIN], mm[q]);
      if (mm[2 + q]) atomicMax(&a.acc[q ? OFF_QUAL_LEN_MAX : OFF_SEQ_LEN_MAX], mm[2 + q]);
    }
    a.carry->lines = c.lines + total;
    a.carry->open_len = mm[4];
    a.carry->bytes = c.bytes + (a.end - (u64)a.lo0);
    a.carry->last_byte = a.base[a.end - 1];
    a.ctl[CTL_REDO] = *nredo;
    a.ctl[CTL_DONE] = 0; a.ctl[CTL_FIRST_NL] = 0; a.ctl[CTL_HEAD] = 0; a.ctl[CTL_HEAD_P0] = 0;
  }
}

// CORE: FQGPU_F_CORE_ONLY (sequence lines only: what `sc fq-count` prints).  A template parameter, not a runtime
// flag: each instantiation drops the other mode's code.
template <bool CORE>
__global__ void __launch_bounds__(THREADS, CTAS_PER_SM) fq_scan_kernel(const ScanArgs a) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31;
  const uint32_t span = blockIdx.x;
  SpanDesc& desc = a.desc[span];
  if (a.pass == 1 && desc.state != SPAN_REDO) return;
  const uint32_t sm0 = smem_u32(smem_raw);
  const uint32_t bar0_s = sm0 + (uint32_t)offsetof(Smem, full_bar);

  for (int i = tid; i < 512; i += THREADS) (&sm.hist[0][0])[i] = 0;
  for (int i = tid; i < POS_BINS + 2; i += THREADS) { sm.seq_len[i] = 0; sm.qual_len[i] = 0; sm.pos32[i] = 0; }
  for (int i = tid; i < PT_WORDS; i += THREADS) sm.ptab[i] = 0;
  if (tid < LOG2_BINS) sm.seq_log2[tid] = 0;
  if (tid < 17 * 4) {  // masks[n]: 0xFF in the first n bytes
    const int n = tid >> 2, w = tid & 3, k = n - 4 * w;
    (&sm.masks[0].x)[tid] = k >= 4 ? 0xFFFFFFFFu : (k <= 0 ? 0u : ((1u << (8 * k)) - 1u));
  }
  if (tid >= 128 && tid <= 128 + KMAX) {
    const uint32_t k = (uint32_t)(tid - 128);
    sm.inv[k] = k >= 2 ? (uint32_t)(((1ull << 32) + k - 1) / k) : 0u;
  }
  const uint32_t t1 = min(a.ntiles, (span + 1u) * a.tps);  // nominal end of the span (tiles)
  if (tid == 0) {
    sm.len_min[0] = sm.len_min[1] = 0xFFFFFFFFu; sm.len_max[0] = sm.len_max[1] = 0;
    sm.big_min[0] = sm.big_min[1] = ~0ull; sm.big_max[0] = sm.big_max[1] = 0;
    sm.over = 0; sm.junk[0] = sm.junk[1] = 0;
    sm.hiflag = 0; sm.flag = 0;
    for (int k = 0; k < 4; k++) sm.ksel[k] = 4u << (8 * k);
    for (int s = 0; s < NSTAGE; s++) mbar_init(bar0_s + 8u * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // ---- where the span starts and the line phase there ----
  u64 start = ~0ull;
  uint32_t guess = 0, guess_valid = 1;
  if (tid < 32) {
    Run r;
    r.totalT = 0; r.proc_end = 0;
    if (a.pass == 1) {
      start = desc.start;
      r.open = 0; r.cnt = desc.guess_valid ? desc.guess : 0u; r.seen = 1; r.prev_byte = '\n'; r.prev_counted = 1;
    } else if (span == 0) {
      const Carry c = *a.carry;
      unsigned fl = c.flags;
      uint32_t hyp = 0;
      if (fl & CARRY_UNKNOWN_START) {
        if (fl & CARRY_HYP_VALID) hyp = (fl >> CARRY_HYP_SHIFT) & 3u;
        else if (!(fl & CARRY_HYP_FAILED)) {
          const uint32_t g = resync_launch(a, lane);  // phase of the line that holds the launch's first byte
          if (g < 4) { hyp = (g - (uint32_t)c.lines) & 3u; fl |= CARRY_HYP_VALID | (hyp << CARRY_HYP_SHIFT); }
          else fl |= CARRY_HYP_FAILED;
          if (lane == 0) a.carry->flags = fl;
        }
      }
      start = (u64)a.lo0;
      r.open = c.open_len; r.cnt = (uint32_t)((c.lines + hyp) & 255u); r.seen = c.lines != 0;
      r.prev_byte = c.bytes ? (int)c.last_byte : 0x100; r.prev_counted = 0;
      guess = r.cnt & 3u;
      if (lane == 0) {
        if ((fl & CARRY_UNKNOWN_START) && c.bytes == 0) a.shard->first_byte = a.base[a.lo0];
        if (a.unknown && c.lines == 0) {
          const bool pending_cr = c.bytes && c.open_len && c.last_byte == '\r' && a.base[a.lo0] != '\n';
          a.ctl[CTL_HEAD] = HEAD_ACTIVE | ((r.cnt & 3u) == 3u ? HEAD_QUAL : 0u) | (pending_cr ? HEAD_PENDING_CR : 0u);
          a.ctl[CTL_HEAD_P0] = c.open_len;
        }
      }
    } else {
      uint32_t g;
      start = resync_span(a, (u64)span * a.tps * TILE, (u64)t1 * TILE, lane, g);
      guess_valid = g < 4;
      guess = guess_valid ? g : 0u;
      r.open = 0; r.cnt = guess; r.seen = 1; r.prev_byte = '\n'; r.prev_counted = 1;
    }
    if (lane == 0) { sm.run = r; sm.longl[0] = start; sm.longl[1] = ((u64)guess_valid << 32) | guess; }
  }
  __syncthreads();
  start = sm.longl[0];
  guess = (uint32_t)sm.longl[1]; guess_valid = (uint32_t)(sm.longl[1] >> 32);
  __syncthreads();
  Sel ksel;
  {
    const uint32_t ks = sm0 + (uint32_t)offsetof(Smem, ksel);
    ksel.h0 = lds32(ks); ksel.h1 = lds32(ks + 4); ksel.h2 = lds32(ks + 8); ksel.h3 = lds32(ks + 12);
  }
  u64 over = 0;
  uint32_t par_bits = 0;
  const bool nonempty = start != ~0ull;
  if (a.pass == 0) {
    if (nonempty) run_span<CORE>(sm, a, ksel, sm0, tid, start, t1, par_bits, over, 1);
    flush_tables(sm, a.acc, tid, over, 1);
    if (tid == 0) {
      const Run& r = sm.run;
      desc.T = r.totalT; desc.start = start; desc.end_open = r.open;
      for (int q = 0; q < 2; q++) {
        u64 mn = sm.len_min[q] != 0xFFFFFFFFu ? (u64)sm.len_min[q] : ~0ull, mx = sm.len_max[q];
        if (sm.big_min[q] < mn) mn = sm.big_min[q];
        if (sm.big_max[q] > mx) mx = sm.big_max[q];
        desc.len_min[q] = mn; desc.len_max[q] = mx;
      }
      desc.guess = guess; desc.guess_valid = guess_valid; desc.nonempty = nonempty ? 1u : 0u;
      desc.reached_end = nonempty && r.proc_end == a.end;
      // ---- the CTA that exits last verifies the guesses and advances the stream carry ----
      __threadfence();
      const u64 prev = atomicAdd(reinterpret_cast<unsigned long long*>(a.ctl + CTL_DONE), 1ull);
      sm.flag = prev + 1 == (u64)gridDim.x;
    }
    __syncthreads();
    if (!sm.flag) return;
    __threadfence();
    if (a.unknown && !CORE) detach_head(a, tid);
    stitch(sm, a, tid);
  } else {
    // second pass: take back what pass 0 counted under the wrong phase, then count under the exact one
    run_span<CORE>(sm, a, ksel, sm0, tid, start, t1, par_bits, over, -1);
    flush_tables(sm, a.acc, tid, over, -1);
    if (tid == 0) {
      Run& r = sm.run;
      r.totalT = 0; r.proc_end = 0; r.open = 0; r.cnt = desc.exact; r.seen = 1; r.prev_byte = '\n'; r.prev_counted = 1;
      sm.len_min[0] = sm.len_min[1] = 0xFFFFFFFFu; sm.len_max[0] = sm.len_max[1] = 0;
      sm.big_min[0] = sm.big_min[1] = ~0ull; sm.big_max[0] = sm.big_max[1] = 0;
    }
    __syncthreads();
    run_span<CORE>(sm, a, ksel, sm0, tid, start, t1, par_bits, over, 1);
    flush_tables(sm, a.acc, tid, over, 1);
    if (tid == 0) {
      for (int q = 0; q < 2; q++) {
        u64 mn = sm.len_min[q] != 0xFFFFFFFFu ? (u64)sm.len_min[q] : ~0ull, mx = sm.len_max[q];
        if (sm.big_min[q] < mn) mn = sm.big_min[q];
        if (sm.big_max[q] > mx) mx = sm.big_max[q];
        if (mn != ~0ull) atomicMin(&a.acc[q ? OFF_QUAL_LEN_MIN : OFF_SEQ_LEN_MIN], mn);
        if (mx) atomicMax(&a.acc[q ? OFF_QUAL_LEN_MAX : OFF_SEQ_LEN_MAX], mx);
      }
    }
  }
}

// Resets the counter block, the stream carry and the launch control words (a new file).
__global__ void fq_reset_kernel(u64* acc, Carry* carry, u64* ctl) {
  for (int w = threadIdx.x; w < BLOCK_WORDS; w += blockDim.x) acc[w] = (w == OFF_SEQ_LEN_MIN || w == OFF_QUAL_LEN_MIN) ? ~0ull : 0ull;
  if (threadIdx.x < CTL_WORDS) ctl[threadIdx.x] = 0;
  if (threadIdx.x == 0) {
    Carry c;
    c.lines = 0; c.open_len = 0; c.bytes = 0; c.last_byte = 0; c.flags = 0;
    c.meta_lines = 0; c.qual_min = -1; c.qual_max = -1; c.meta_status = 0; c.meta_pending_cr = 0;
    c.cur_has = 0; c.cur_min = 0; c.cur_max = 0; c.pad = 0;
    *carry = c;
  }
}

// ------------------------------------------------------------------------------------------
// host-callable launchers (used by fqgpu_api.cu)
// ------------------------------------------------------------------------------------------
size_t scan_smem_bytes() { return sizeof(Smem); }
int scan_tile_bytes() { return TILE; }
int scan_threads() { return THREADS; }
int scan_ctas_per_sm() { return CTAS_PER_SM; }

cudaError_t scan_configure() {
  cudaError_t e = cudaFuncSetAttribute(fq_scan_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(fq_scan_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
}

cudaError_t launch_reset(u64* acc, Carry* carry, u64* ctl, cudaStream_t st) {
  fq_reset_kernel<<<1, 512, 0, st>>>(acc, carry, ctl);
  return cudaGetLastError();
}

// Spans of a launch over `bytes` bytes (incl. the alignment slack in front) with `resident` CTAs resident at once.
// Large launches get several spans per resident CTA, handed out by the hardware as CTAs finish: of the two CTAs of an
// SM the one that started first runs faster, so with one span each the slower half finishes late and alone; with
// shorter spans the tail is one short span.  Spans stay >= MIN_SPAN_TILES tiles (FQGPU_SPAN_MIN_TILES: tests).
uint32_t scan_span_count(u64 bytes, int resident) {
  static const u64 env_min = getenv("FQGPU_SPAN_MIN_TILES") && atoi(getenv("FQGPU_SPAN_MIN_TILES")) > 0 ? (u64)atoi(getenv("FQGPU_SPAN_MIN_TILES")) : 0;
  static const u64 waves = getenv("FQGPU_SPAN_WAVES") && atoi(getenv("FQGPU_SPAN_WAVES")) > 0 ? (u64)atoi(getenv("FQGPU_SPAN_WAVES")) : (u64)SPAN_WAVES;
  const u64 ntiles = (bytes + TILE - 1) / TILE;
  const u64 max_waves = waves < (u64)SPAN_WAVES ? waves : (u64)SPAN_WAVES;
  if (env_min) {  // tests: spans of env_min tiles wherever the input allows
    u64 n = ntiles / env_min;
    const u64 cap = (u64)resident * max_waves;
    n = n < 1 ? 1 : (n > cap ? cap : n);
    if (n > (u64)resident) n = (n / (u64)resident) * (u64)resident;
    return (uint32_t)n;
  }
  // up to one span per resident CTA while spans stay >= MIN_SPAN_TILES; further waves only once every span of every
  // wave is >= WAVE_SPAN_TILES (a span costs its start-up -- the phase guess, the tables -- and its flush: on 4.5 GB
  // shards eight short spans per CTA cost more than the tail they save)
  u64 n = ntiles / (u64)MIN_SPAN_TILES;
  if (n < 1) n = 1;
  if (n <= (u64)resident) return (uint32_t)n;
  u64 w = ntiles / ((u64)resident * (u64)WAVE_SPAN_TILES);
  w = w < 1 ? 1 : (w > max_waves ? max_waves : w);
  return (uint32_t)((u64)resident * w);
}

// Scans `nbytes` at `ptr` (any alignment) as the continuation of the stream described by `carry`: pass 0 (every span
// under its guessed phase; its last CTA verifies and advances the carry) and pass 1 (the spans whose guess was wrong;
// exits at once otherwise).
cudaError_t launch_scan(const void* ptr, size_t nbytes, Carry* carry, ShardInfo* shard, u64* acc, SpanDesc* desc, u64* ctl,
                        bool unknown_start, int resident, bool core_only, cudaStream_t st) {
  if (nbytes == 0) return cudaSuccess;
  const uintptr_t addr = (uintptr_t)ptr;
  ScanArgs a;
  a.lo0 = (uint32_t)(addr & 15);
  a.base = (const uint8_t*)(addr - a.lo0);
  a.end = (u64)a.lo0 + nbytes;
  a.ntiles = (uint32_t)((a.end + TILE - 1) / TILE);
  const uint32_t nspans = scan_span_count(a.end, resident);
  a.tps = (a.ntiles + nspans - 1) / nspans;
  a.nspans = (a.ntiles + a.tps - 1) / a.tps;
  a.carry = carry; a.shard = shard; a.acc = acc; a.desc = desc; a.ctl = ctl;
  a.unknown = unknown_start ? 1u : 0u;
  for (uint32_t pass = 0; pass < 2; pass++) {
    a.pass = pass;
    if (core_only) fq_scan_kernel<true><<<a.nspans, THREADS, sizeof(Smem), st>>>(a);
    else fq_scan_kernel<false><<<a.nspans, THREADS, sizeof(Smem), st>>>(a);
  }
  return cudaGetLastError();
}

}  // namespace fq
