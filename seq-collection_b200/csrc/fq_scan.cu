// fq_scan.cu -- the FASTQ scanning hot path on sm_100a: ONE launch per buffer.
//
// Replaces the per-line loop of the reference (src/fq_count.nim:38-45: `for line in lines(stream)`,
// i mod 4 classing, count("G")+count("C"), count("N"), line.len) with one pass over the bytes; the quality fold
// of src/fq_meta.nim:245-246 is fq_meta.cu, beside it on a second stream.
//
// Persistent CTAs (2 per SM) take 32 KiB tiles from a ticket counter; a tile is brought into shared memory by a
// 1-D TMA copy (cp.async.bulk + mbarrier, two stages: the copy of the next tile runs under the work on this one)
// and goes through five phases, all warps together:
//
//   A   boundary classification: coalesced 16-byte groups -> 16-bit '\n' masks (SWAR compare, IDP.4A movemask)
//       -> the tile's newline bitmap in shared memory.
//   B1  every thread owns 64 CONSECUTIVE bytes of the bitmap: popc, one block-wide prefix (newline count and
//       position of the last newline) -> the tile total; warp 0 publishes it and walks back over the predecessor
//       tiles' words (chained "decoupled look-back" prefix; a tile only ever waits for tiles whose ticket was
//       taken earlier, i.e. that are running): exact line number (mod 256) and open-line length at the tile start.
//       This is the tile-edge / chunk-edge record carry, resolved on the device.
//   B2  every thread walks its four groups: a group without a newline gets a 16-bit descriptor (line class =
//       line number mod 4, record parity, line position of its first byte); a group with newlines is cut into
//       ITEMS (group, byte range, class, position, "ends its line") appended to a tile-wide queue.
//   C   one lane per group: LDS.128, histogram of sequence / quality bytes by one IDP.4A (address) + one shared
//       atomic (ATOMS.POPC.INC, which merges equal addresses: un-striped 256-bin tables) per byte; per-position
//       quality sums as eight packed 16-bit-pair atomics into a bank-skewed table (even / odd position tables,
//       no alignment shifts).
//   D   the item queue, one lane per item: masked histogram / per-position sums, line lengths, the '\r' rule.
//
// Counter tables live in shared memory for the life of the CTA and are added to the context's block by 64-bit
// atomics at CTA exit (K3); the CTA that exits last advances the stream carry (fq::Carry).  Every input byte is
// read from HBM once.  Line semantics are Nim's streams.lines: split at '\n', drop one '\r' directly before it;
// the trailing unterminated line is accounted by the host from fq::Carry at finish().
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fq_dev.cuh"

namespace fq {

constexpr int THREADS = 512;
constexpr int NWARPS = THREADS / 32;
constexpr int GPT = 4;                        // 16-byte groups per thread and tile
constexpr int NG = THREADS * GPT;             // groups per tile
constexpr int TILE = NG * 16;                 // 32 KiB
constexpr int NSTAGE = 2;
constexpr int QCAP = 2048;                    // item queue; tiles with more newlines take the byte walker
constexpr uint32_t QPOS_MAX = 1023;           // line positions saturate here (>= POS_BINS is the overflow bin anyway)
constexpr int OPEN_CLIP = 1 << 30;
// Per-position quality sums, 16-bit pairs.  Q = position + 16.  Even Q: pair A = Q >> 1 of the EVEN table holds
// (Q, Q+1); odd Q: pair A = (Q+1) >> 1 of the ODD table holds (Q, Q+1).  Pair A lives in cell (r, c) with
// 8 c + r = A, r < 8, at word r * PT_STRIDE + c; a group adds its eight pairs at immediate offsets PT_STRIDE * i,
// so rows 8..15 alias (r - 8, c + 1) and consecutive groups of a line hit consecutive banks.  Two copies
// (alternating with the record parity) keep neighbouring records apart.
constexpr int PT_STRIDE = 33;
constexpr int PT_COPY = 16 * PT_STRIDE;       // 528 words
constexpr int PT_WORDS = 4 * PT_COPY;         // [odd][copy]
constexpr uint32_t FLUSH_BYTES = 4u << 20;    // 32-bit tables are added to the block at least this often

struct ScanArgs {
  const uint8_t* base;  // 16-byte aligned; the launch covers bytes [lo0, end) relative to base
  uint32_t lo0;
  u64 end;
  uint32_t ntiles;
  Carry* carry;
  ShardInfo* shard;  // detached head of a multi-GPU shard (rank > 0)
  u64* acc;          // [BLOCK_WORDS] the context's counter block
  u64* state;        // [>= ntiles] look-back words
  u64* ctl;          // [CTL_WORDS]
  uint32_t epoch;    // 1..255, changes with every launch: stale look-back words read as "nothing yet"
  uint32_t unknown;  // shard with an unknown start (fqgpu_shard_begin, rank > 0, not rescanned)
  uint32_t dbg;      // 1: EXPERIMENT -- analytic prefix of the synthetic Illumina stream instead of the look-back
};

// look-back word: epoch << 56 | flag << 54 | payload
//   flag 1 (tile alone):  T << 18 | tail      T = newlines of the tile, tail = bytes after the last one (T == 0: tile length)
//   flag 2 (inclusive):   seen << 53 | cnt << 45 | open     cnt = lines so far mod 256, open = open-line bytes, seen = any newline so far
constexpr u64 ST_AGG = 1ull << 54, ST_INC = 2ull << 54;
constexpr u64 OPEN_MASK = (1ull << 45) - 1;

struct Sel { uint32_t h0, h1, h2, h3; };  // IDP.4A selectors 4 << 8k: histogram address = byte_k * 4 + base

struct TileIn {
  u64 open;          // bytes of the open line before the tile
  u64 open_out;      // ... after it
  uint32_t cnt;      // lines before the tile (mod 256; shards with an unknown start: under the hypothesis)
  uint32_t seen;     // a newline has been seen in the stream before the tile
  uint32_t T;        // newlines of the tile
  int prev_byte;     // byte before the tile's first valid byte (0x100: none)
};

struct __align__(128) Smem {
  uint8_t buf[NSTAGE][TILE];
  u64 queue[QCAP];
  uint16_t bitmap[NG];               // bit b of entry g: byte 16 g + b is '\n'
  uint16_t ginfo[NG];                // (line number & 7) | position << 3 for groups without a newline inside counted lines' reach; 0 = nothing to do
  uint32_t hist[2][256];             // [0] sequence, [1] quality
  uint32_t ptab[PT_WORDS];
  uint32_t pos32[POS_BINS + 2];      // per-position sums, 32-bit (second level of ptab; generic paths)
  uint32_t seq_len[POS_BINS + 2];
  uint32_t qual_len[POS_BINS + 2];
  uint32_t seq_log2[LOG2_BINS];
  uint4 masks[17];                   // masks[n]: 0xFF in the first n bytes of a group
  u64 wtot[NWARPS];                  // per warp: newlines << 32 | (offset of its last newline + 1)
  u64 full_bar[NSTAGE];
  u64 big_min[2], big_max[2];        // line lengths >= 2^32
  u64 over;                          // quality bytes at positions >= POS_BINS
  uint32_t len_min[2], len_max[2];   // [0] seq, [1] qual
  uint32_t junk[2];                  // masked bytes counted in bin 0
  uint32_t ksel[4];
  uint32_t tile[2];
  uint32_t nitems, pt_lines, hiflag, bytes_since_flush;
  TileIn in;
};
static_assert(sizeof(Smem) <= 115712, "two CTAs per SM");

struct TileCtx {
  uint32_t tile;
  int vlo, vhi;     // valid bytes of the tile
  u64 toff;         // offset of the tile relative to base
  uint32_t buf_s;
};

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

// Everything the CTA holds in shared memory -> the context's block (64-bit atomics), tables cleared.
__device__ __noinline__ void flush_all(Smem& sm, u64* acc, int tid, u64& over) {
  __syncthreads();
  flush_ptab(sm, tid);
  if (over) { atomicAdd(&sm.over, over); over = 0; }
  __syncthreads();
  for (int i = tid; i < 512; i += THREADS) {
    uint32_t v = (&sm.hist[0][0])[i];
    if ((i & 255) == 0) v -= sm.junk[i >> 8];  // masked bytes were counted in bin 0
    if (v) { atomicAdd(&acc[OFF_HIST_SEQ + i], (u64)v); }
    (&sm.hist[0][0])[i] = 0;
  }
  for (int i = tid; i <= POS_BINS; i += THREADS) {
    uint32_t v = sm.pos32[i];
    if (v) { atomicAdd(&acc[OFF_POS_SUM + i], (u64)v); sm.pos32[i] = 0; }
    v = sm.seq_len[i];
    if (v) { atomicAdd(&acc[OFF_SEQ_LEN + i], (u64)v); sm.seq_len[i] = 0; }
    v = sm.qual_len[i];
    if (v) { atomicAdd(&acc[OFF_QUAL_LEN + i], (u64)v); sm.qual_len[i] = 0; }
  }
  if (tid < LOG2_BINS) { const uint32_t v = sm.seq_log2[tid]; if (v) { atomicAdd(&acc[OFF_SEQ_LOG2 + tid], (u64)v); sm.seq_log2[tid] = 0; } }
  __syncthreads();
  if (tid == 0) {
    if (sm.over) { atomicAdd(&acc[OFF_POS_SUM + POS_BINS], sm.over); sm.over = 0; }
    sm.junk[0] = sm.junk[1] = 0;
    sm.bytes_since_flush = 0;
    sm.pt_lines = 0;
    sm.hiflag = 0;
    for (int q = 0; q < 2; q++) {
      u64 mn = sm.len_min[q] != 0xFFFFFFFFu ? (u64)sm.len_min[q] : ~0ull, mx = sm.len_max[q];
      if (sm.big_min[q] < mn) mn = sm.big_min[q];
      if (sm.big_max[q] > mx) mx = sm.big_max[q];
      if (mn != ~0ull) atomicMin(&acc[q ? OFF_QUAL_LEN_MIN : OFF_SEQ_LEN_MIN], mn);
      if (mx) atomicMax(&acc[q ? OFF_QUAL_LEN_MAX : OFF_SEQ_LEN_MAX], mx);
      sm.len_min[q] = 0xFFFFFFFFu; sm.len_max[q] = 0; sm.big_min[q] = ~0ull; sm.big_max[q] = 0;
    }
  }
  __syncthreads();
}

// A counted line of raw length `rawlen` ends with the newline at tile offset `off`: the '\r' rule and the length
// tables.  ql: 0 sequence, 1 quality.  fn: the line started before the tile.
template <bool CORE>
__device__ __forceinline__ void line_end(Smem& sm, const ScanArgs& a, const TileCtx& t, int off, u64 rawlen, bool fn, uint32_t ql) {
  const int prev = off > t.vlo ? (int)lds8(t.buf_s + (uint32_t)off - 1u) : sm.in.prev_byte;
  const u64 cr = (rawlen > 0 && prev == '\r') ? 1 : 0;
  if (cr) {  // the '\r' was counted as content where it stands: take it back
    const uint32_t p = (rawlen - 1) < (u64)POS_BINS ? (uint32_t)(rawlen - 1) : (uint32_t)POS_BINS;
    if (off > t.vlo) {
      atomicSub(&sm.hist[ql]['\r'], 1u);
      if (!CORE && ql) { if (p < (uint32_t)POS_BINS) atomicSub(&sm.pos32[p], (uint32_t)'\r'); else atomicAdd(&sm.over, (u64)(0ull - '\r')); }
    } else if (t.tile > 0) {  // it is the last byte of the previous tile: another CTA counted it
      atomicAdd(&a.acc[(ql ? OFF_HIST_QUAL : OFF_HIST_SEQ) + '\r'], ~0ull);
      if (!CORE && ql) atomicAdd(&a.acc[OFF_POS_SUM + p], 0ull - '\r');
    }  // else: the last byte of the previous launch, never counted (its fate was left to this launch)
  }
  const u64 len = rawlen - cr;
  if (a.unknown && fn && !sm.in.seen) {  // the shard's first line: its length is stitched by the combine step
    a.shard->head_len = rawlen;
    a.shard->head_cr = (unsigned)cr;
    a.ctl[CTL_FIRST_NL] = t.toff + (u64)off + 1;
    return;
  }
  const uint32_t bin = len < (u64)POS_BINS ? (uint32_t)len : (uint32_t)POS_BINS;
  if (ql) atomicAdd(&sm.qual_len[bin], 1u);
  else { atomicAdd(&sm.seq_len[bin], 1u); atomicAdd(&sm.seq_log2[log2_bin(len)], 1u); }
  if (len <= 0xFFFFFFFFull) { atomicMin(&sm.len_min[ql], (uint32_t)len); atomicMax(&sm.len_max[ql], (uint32_t)len); }
  else { atomicMin(&sm.big_min[ql], len); atomicMax(&sm.big_max[ql], len); }
}

// Phase A, rare forms (kept out of the hot code): this thread's groups again with the exact compare; an edge tile.
__device__ __noinline__ void phase_a_redo(uint32_t buf_s, uint32_t bm_s, int tid) {
  for (int k = 0; k < GPT; k++) {
    const uint32_t g = (uint32_t)(k * THREADS + tid);
    sts16(bm_s + 2u * g, nl_mask16(lds128(buf_s + 16u * g)));
  }
}
__device__ __noinline__ uint32_t phase_a_edge(uint32_t buf_s, uint32_t bm_s, int lo, int hi, int tid) {
  uint32_t hib = 0;
  for (int k = 0; k < GPT; k++) {
    const int g = k * THREADS + tid, off = g * 16;
    uint32_t m = 0;
    if (off < hi && off + 16 > lo) {
      const uint4 v = lds128(buf_s + 16u * (uint32_t)g);
      m = nl_mask16(v);
      int lo_k = lo - off; lo_k = lo_k < 0 ? 0 : lo_k;
      int hi_k = hi - off; hi_k = hi_k > 16 ? 16 : hi_k;
      m &= ((1u << hi_k) - 1u) & ~((1u << lo_k) - 1u);
      if (lo_k == 0 && hi_k == 16) hib |= (v.x | v.y) | (v.z | v.w);
      else hib |= 0x80u;  // ragged group: stale bytes beside the valid ones; be conservative
    }
    sts16(bm_s + 2u * (uint32_t)g, m);
  }
  return hib;
}

// Appends an item to the tile's queue.
//   lo32: group | lo << 11 | hi << 16 | quality << 21 | ends-its-line << 22 | started-before-the-tile << 23 | parity << 24
//   hi32: line position of byte lo (saturated at OPEN_CLIP)
__device__ __forceinline__ void push_item(Smem& sm, const ScanArgs& a, uint32_t e0, uint32_t pos) {
  const uint32_t slot = atomicAdd(&sm.nitems, 1u);
  if (slot < (uint32_t)QCAP) sm.queue[slot] = (u64)e0 | ((u64)pos << 32);
  else a.ctl[CTL_ERROR] = 1;  // cannot happen: the caller checked T + 4 <= QCAP
}

// Phase B2 for one thread: descriptors of its four groups, items for the groups with newlines.
// cnt = lines before the run (low bits), lc = newlines of the tile before the run, last = tile offset of the last
// newline before the run (negative: the line started before the tile).
template <bool CORE, bool EDGE>
__device__ __forceinline__ void phase_b2(Smem& sm, const ScanArgs& a, const TileCtx& t, uint32_t gi_s, int tid, uint32_t wl, uint32_t wh,
                                         uint32_t cnt, uint32_t lc, int last) {
  u64 gpack = 0;
#pragma unroll
  for (int k = 0; k < GPT; k++) {
    const uint32_t m = ((k < 2 ? wl : wh) >> (16 * (k & 1))) & 0xFFFFu;
    const int o = 64 * tid + 16 * k;
    int lo = 0, hi_end = 16;
    bool ragged = false;
    if (EDGE) {
      lo = t.vlo - o; lo = lo < 0 ? 0 : lo;
      hi_end = t.vhi - o; hi_end = hi_end > 16 ? 16 : hi_end;
      if (lo >= hi_end) continue;  // nothing valid in this group (its mask is zero)
      ragged = lo > 0 || hi_end < 16;
    }
    if (m == 0 && !ragged) {
      uint32_t q = (uint32_t)(o - last - 1);
      q = q < QPOS_MAX ? q : QPOS_MAX;
      gpack |= (u64)((cnt & 7u) | (q << 3)) << (16 * k);
    } else {
      uint32_t mm = m;
      int posi = o + lo - last - 1;
      uint32_t pos = (uint32_t)(posi < OPEN_CLIP ? posi : OPEN_CLIP);
      const uint32_t g = (uint32_t)(GPT * tid + k);
      for (;;) {
        const int hi = mm ? __ffs(mm) - 1 : hi_end;
        const uint32_t cls = cnt & 3u;
        const bool counted = CORE ? cls == 1u : (cls & 1u) != 0;
        if (counted && (mm || hi > lo))
          push_item(sm, a, g | ((uint32_t)lo << 11) | ((uint32_t)hi << 16) | ((cls >> 1) << 21) | (mm ? 1u << 22 : 0u) | (lc == 0 ? 1u << 23 : 0u) |
                               (((cnt >> 2) & 1u) << 24), pos);
        if (!mm) break;
        mm &= mm - 1;
        cnt++; lc++;
        last = o + hi;
        lo = hi + 1;
        pos = 0;
      }
    }
  }
  sts64(gi_s + 8u * (uint32_t)tid, gpack);
}
template <bool CORE>
__device__ __noinline__ void phase_b2_edge(Smem& sm, const ScanArgs& a, const TileCtx& t, uint32_t gi_s, int tid, uint32_t wl, uint32_t wh,
                                           uint32_t cnt, uint32_t lc, int last) {
  phase_b2<CORE, true>(sm, a, t, gi_s, tid, wl, wh, cnt, lc, last);
}

// The byte walker: tiles with too many newlines for the item queue.  Every thread takes its own 64 bytes one at
// a time (exact for any content); '\r' is counted where it stands and taken back by line_end like everywhere else.
template <bool CORE>
__device__ __noinline__ void phase_walk(Smem& sm, const ScanArgs& a, const TileCtx& t, int tid, uint32_t cnt, uint32_t lc, int last) {
  u64 over = 0;
  for (int x = 0; x < 16 * GPT; x++) {
    const int off = 16 * GPT * tid + x;
    if (off < t.vlo || off >= t.vhi) continue;
    const uint32_t b = lds8(t.buf_s + (uint32_t)off);
    const uint32_t cls = cnt & 3u;
    const bool counted = CORE ? cls == 1u : (cls & 1u) != 0;
    const u64 pos = lc == 0 ? sm.in.open + (u64)(off - t.vlo) : (u64)(off - last - 1);
    if (b == '\n') {
      if (counted) line_end<CORE>(sm, a, t, off, pos, lc == 0, cls >> 1);
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
}

// resync: guess the line phase at the start of a shard whose predecessor is on another GPU.  Launch-relative line j
// (j >= 1) starts after the launch's j-th newline; the first j whose line starts with '@' while line j+2 starts with
// '+' is a header, so (lines before the launch) = -j (mod 4).  For valid 4-line FASTQ this is unambiguous (a quality
// line starting with '@' is followed two lines later by a sequence line).  One warp; returns 0..3, or 4 = no guess.
constexpr int RESYNC_LINES = 40;
constexpr u64 RESYNC_BYTES = 4ull << 20;
__device__ __noinline__ uint32_t resync_guess(const ScanArgs& a, int lane) {
  const u64 stop = RESYNC_BYTES < a.end ? RESYNC_BYTES : a.end;
  uint32_t first[RESYNC_LINES + 1];
  int nlines = 0;
  uint32_t guess = 4;
  for (u64 o = 0; o < stop && guess == 4 && nlines < RESYNC_LINES; o += 512) {
    const u64 g = o + (u64)lane * 16;
    uint32_t m = 0;
    if (g < a.end) {
      const uint4 v = *reinterpret_cast<const uint4*>(a.base + g);
      m = nl_mask16(v);
      if (g + 16 > a.end) m &= (1u << (a.end - g)) - 1u;
      if (g < (u64)a.lo0) m &= ~((1u << min((u64)16, (u64)a.lo0 - g)) - 1u);
    }
    uint32_t any = __ballot_sync(0xffffffffu, m != 0);
    while (any && nlines < RESYNC_LINES) {
      const int src = __ffs(any) - 1;
      any &= any - 1;
      uint32_t mm = __shfl_sync(0xffffffffu, m, src);
      while (mm && nlines < RESYNC_LINES) {
        const int k = __ffs(mm) - 1;
        mm &= mm - 1;
        const u64 start = o + (u64)src * 16 + (u64)k + 1;  // line starts after this newline
        nlines++;
        first[nlines] = start < a.end ? (uint32_t)a.base[start] : 0x100u;
        if (nlines >= 3 && first[nlines - 2] == '@' && first[nlines] == '+') {
          guess = (uint32_t)((4 - ((nlines - 2) & 3)) & 3);
          break;
        }
      }
      if (guess != 4) break;
    }
  }
  return guess;
}

// The detached head of a shard with an unknown start (the bytes before the stream's first newline): the scan has
// added its per-position sums at positions relative to the shard start; they move from the block into
// ShardInfo::head_pos, which fqgpu_shard_combine shifts by the bytes the line had on the previous ranks.
// Run by the CTA that exits last.
__device__ __noinline__ void detach_head(const ScanArgs& a, int tid) {
  const u64 fl = a.ctl[CTL_HEAD];
  if (!(fl & HEAD_ACTIVE) || !(fl & HEAD_QUAL)) return;
  const u64 P0 = a.ctl[CTL_HEAD_P0];
  const u64 fnl = a.ctl[CTL_FIRST_NL];                  // offset of the first newline + 1, 0 = none
  u64 ve = fnl ? fnl - 1 : a.end;                       // content end
  if (ve > (u64)a.lo0 && a.base[ve - 1] == '\r') ve--;  // dropped before the newline / left to the next launch at the end
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

__device__ __forceinline__ u64 warp_sum_u64(u64 v) {
#pragma unroll
  for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

// Warp 0: the tile's place in the stream.  Tile 0 takes it from the carry (or the shard hypothesis); every other
// tile publishes its own totals and sums its predecessors' words, 32 at a time, back to the nearest inclusive one.
__device__ __forceinline__ void tile_prefix(Smem& sm, const ScanArgs& a, const TileCtx& t, int lane, uint32_t T, uint32_t tail) {
  const u64 ep = (u64)a.epoch << 56;
  uint32_t cnt_in = 0, seen_in = 0;
  u64 open_in = 0;
  int prev_byte = 0x100;
  if (t.tile == 0) {
    const Carry c = *a.carry;
    unsigned fl = c.flags;
    uint32_t hyp = 0;
    if (fl & CARRY_UNKNOWN_START) {
      if (fl & CARRY_HYP_VALID) hyp = (fl >> CARRY_HYP_SHIFT) & 3u;
      else if (!(fl & CARRY_HYP_FAILED)) {
        const uint32_t guess = resync_guess(a, lane);
        if (guess < 4) { hyp = (guess - (uint32_t)c.lines) & 3u; fl |= CARRY_HYP_VALID | (hyp << CARRY_HYP_SHIFT); }
        else fl |= CARRY_HYP_FAILED;
        if (lane == 0) a.carry->flags = fl;
      }
    }
    cnt_in = (uint32_t)((c.lines + hyp) & 255u);
    seen_in = c.lines != 0;
    open_in = c.open_len;
    prev_byte = c.bytes ? (int)c.last_byte : 0x100;
    if (lane == 0) {
      if ((fl & CARRY_UNKNOWN_START) && c.bytes == 0) a.shard->first_byte = a.base[a.lo0];
      const bool pending_cr = c.bytes && c.open_len && c.last_byte == '\r' && a.base[a.lo0] != '\n';
      if (a.unknown && c.lines == 0) {
        a.ctl[CTL_HEAD] = HEAD_ACTIVE | ((cnt_in & 3u) == 3u ? HEAD_QUAL : 0u) | (pending_cr ? HEAD_PENDING_CR : 0u);
        a.ctl[CTL_HEAD_P0] = c.open_len;
      }
    }
  } else if (a.dbg == 1) {
    const u64 X = t.toff - a.lo0, rec = X / 360, r = X % 360;
    const uint32_t line = r < 56 ? 0u : r < 207 ? 1u : r < 209 ? 2u : 3u;
    const uint32_t ls = line == 0 ? 0u : line == 1 ? 56u : line == 2 ? 207u : 209u;
    cnt_in = (uint32_t)((rec * 4 + line) & 255u); seen_in = 1; open_in = r - ls;
    prev_byte = (int)a.base[t.toff - 1];
  } else {
    if (lane == 0) {
      st_relaxed_gpu(a.state + t.tile, ep | ST_AGG | ((u64)T << 18) | (u64)tail);
      prev_byte = (int)a.base[t.toff - 1];
    }
    uint32_t cnt_acc = 0;
    u64 open_acc = 0;
    bool found = false;
    long long i = (long long)t.tile - 1;
    for (;;) {
      const long long p = i - lane;
      u64 st = 0;
      bool ready;
      do {
        if (p >= 0) st = ld_relaxed_gpu(a.state + p);
        ready = p < 0 || ((st >> 56) == (u64)a.epoch && ((st >> 54) & 3u) != 0);
      } while (!__all_sync(0xffffffffu, ready));
      const bool isinc = p >= 0 && ((st >> 54) & 3u) == 2u;
      const uint32_t pm = __ballot_sync(0xffffffffu, isinc);
      const bool act = pm ? lane <= __ffs(pm) - 1 : true;  // (tile 0 is inclusive from the start, so lanes with p < 0 lie behind it)
      const uint32_t Tl = (uint32_t)(st >> 18) & 0x3FFFFu, taill = (uint32_t)st & 0x3FFFFu;
      const uint32_t hb = __ballot_sync(0xffffffffu, act && !isinc && Tl != 0);
      const int nearest = hb ? __ffs(hb) - 1 : 32;
      uint32_t vc = 0;
      u64 vo = 0;
      if (act) {
        if (isinc) {
          vc = (uint32_t)(st >> 45) & 0xFFu;
          if (nearest == 32) vo = st & OPEN_MASK;
          seen_in |= (uint32_t)(st >> 53) & 1u;
        } else {
          vc = Tl;
          if (lane <= nearest) vo = taill;
        }
      }
      cnt_acc += __reduce_add_sync(0xffffffffu, vc);
      vo = warp_sum_u64(vo);
      if (!found) { open_acc += vo; found = nearest < 32; }
      if (pm) break;
      i -= 32;
    }
    seen_in = __any_sync(0xffffffffu, seen_in != 0) || found;
    cnt_in = cnt_acc & 255u;
    open_in = open_acc;
    prev_byte = __shfl_sync(0xffffffffu, prev_byte, 0);
  }
  const u64 open_out = T ? (u64)tail : open_in + (u64)tail;
  if (lane == 0) {
    st_relaxed_gpu(a.state + t.tile, ep | ST_INC | ((u64)((seen_in || T) ? 1 : 0) << 53) | ((u64)((cnt_in + T) & 255u) << 45) | (open_out & OPEN_MASK));
    sm.in.open = open_in; sm.in.open_out = open_out; sm.in.cnt = cnt_in; sm.in.seen = seen_in; sm.in.T = T; sm.in.prev_byte = prev_byte;
    sm.nitems = 0;
    if (t.tile + 1 == a.ntiles) a.ctl[CTL_OPEN_OUT] = open_out;
  }
}

// CORE: FQGPU_F_CORE_ONLY (sequence lines only: what `sc fq-count` prints).  A template parameter, not a runtime
// flag: each instantiation drops the other mode's code.
template <bool CORE>
__global__ void __launch_bounds__(THREADS, 2) fq_scan_kernel(const ScanArgs a) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t sm0 = smem_u32(smem_raw);
  const uint32_t buf0_s = sm0 + (uint32_t)offsetof(Smem, buf);
  const uint32_t bar0_s = sm0 + (uint32_t)offsetof(Smem, full_bar);
  const uint32_t bm_s = sm0 + (uint32_t)offsetof(Smem, bitmap);
  const uint32_t gi_s = sm0 + (uint32_t)offsetof(Smem, ginfo);
  const uint32_t hist_s = sm0 + (uint32_t)offsetof(Smem, hist);
  const uint32_t ptab_s = sm0 + (uint32_t)offsetof(Smem, ptab);
  const uint32_t masks_s = sm0 + (uint32_t)offsetof(Smem, masks);
  const uint32_t queue_s = sm0 + (uint32_t)offsetof(Smem, queue);
  unsigned long long* ticket = reinterpret_cast<unsigned long long*>(a.ctl + CTL_TICKET);

  for (int i = tid; i < 512; i += THREADS) (&sm.hist[0][0])[i] = 0;
  for (int i = tid; i < POS_BINS + 2; i += THREADS) { sm.seq_len[i] = 0; sm.qual_len[i] = 0; sm.pos32[i] = 0; }
  for (int i = tid; i < PT_WORDS; i += THREADS) sm.ptab[i] = 0;
  if (tid < LOG2_BINS) sm.seq_log2[tid] = 0;
  if (tid < 17 * 4) {  // masks[n]: 0xFF in the first n bytes
    const int n = tid >> 2, w = tid & 3, k = n - 4 * w;
    (&sm.masks[0].x)[tid] = k >= 4 ? 0xFFFFFFFFu : (k <= 0 ? 0u : ((1u << (8 * k)) - 1u));
  }
  uint32_t next_tile = 0;  // thread 0: the ticket of the tile after the one in flight
  if (tid == 0) {
    sm.len_min[0] = sm.len_min[1] = 0xFFFFFFFFu; sm.len_max[0] = sm.len_max[1] = 0;
    sm.big_min[0] = sm.big_min[1] = ~0ull; sm.big_max[0] = sm.big_max[1] = 0;
    sm.over = 0; sm.junk[0] = sm.junk[1] = 0;
    sm.nitems = 0; sm.pt_lines = 0; sm.hiflag = 0; sm.bytes_since_flush = 0;
    for (int k = 0; k < 4; k++) sm.ksel[k] = 4u << (8 * k);
    for (int s = 0; s < NSTAGE; s++) mbar_init(bar0_s + 8u * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const uint32_t t0 = (uint32_t)atomicAdd(ticket, 1ull);
    sm.tile[0] = t0;
    if (t0 < a.ntiles) {
      const u64 toff = (u64)t0 * TILE;
      const uint32_t bytes = (uint32_t)((((a.end - toff) < (u64)TILE ? (a.end - toff) : (u64)TILE) + 15) & ~15ull);
      mbar_expect_tx(bar0_s, bytes);
      tma_load_1d(buf0_s, a.base + toff, bytes, bar0_s);
      next_tile = (uint32_t)atomicAdd(ticket, 1ull);
    } else {
      next_tile = a.ntiles;
    }
  }
  u64 over = 0;     // quality bytes at positions >= POS_BINS
  u64 totalT = 0;   // newlines of this CTA's tiles (warp 0)
  __syncthreads();
  Sel ksel;
  {
    const uint32_t ks = sm0 + (uint32_t)offsetof(Smem, ksel);
    ksel.h0 = lds32(ks); ksel.h1 = lds32(ks + 4); ksel.h2 = lds32(ks + 8); ksel.h3 = lds32(ks + 12);
  }

  for (int it = 0;; it++) {
    TileCtx t;
    t.tile = sm.tile[it & 1];
    if (t.tile >= a.ntiles) break;
    const int stage = it & 1;
    t.toff = (u64)t.tile * TILE;
    t.vlo = t.tile == 0 ? (int)a.lo0 : 0;
    t.vhi = (a.end - t.toff) < (u64)TILE ? (int)(a.end - t.toff) : TILE;
    t.buf_s = buf0_s + (uint32_t)stage * TILE;
    const bool edge = t.vlo != 0 || t.vhi != TILE;
    if (tid == 0) {  // the copy of the next tile into the other stage (free since the end of the last iteration)
      sm.tile[(it + 1) & 1] = next_tile;
      if (next_tile < a.ntiles) {
        const u64 toff = (u64)next_tile * TILE;
        const uint32_t bytes = (uint32_t)((((a.end - toff) < (u64)TILE ? (a.end - toff) : (u64)TILE) + 15) & ~15ull);
        mbar_expect_tx(bar0_s + 8u * (stage ^ 1), bytes);
        tma_load_1d(buf0_s + (uint32_t)(stage ^ 1) * TILE, a.base + toff, bytes, bar0_s + 8u * (stage ^ 1));
        next_tile = (uint32_t)atomicAdd(ticket, 1ull);  // consumed in the next iteration: the round trip is off the critical path
      }
    }
    mbar_wait(bar0_s + 8u * stage, (uint32_t)(it >> 1) & 1u);

    // ---- A: newline masks of the tile's groups -> bitmap ----
    if (!edge) {
      uint4 v[GPT];
#pragma unroll
      for (int k = 0; k < GPT; k++) v[k] = lds128(t.buf_s + 16u * (uint32_t)(k * THREADS + tid));
      uint32_t hib = 0;
#pragma unroll
      for (int k = 0; k < GPT; k++) {
        hib |= (v[k].x | v[k].y) | (v[k].z | v[k].w);
        sts16(bm_s + 2u * (uint32_t)(k * THREADS + tid), nl_mask16_ascii(v[k]));
      }
      if (hib & 0x80808080u) { phase_a_redo(t.buf_s, bm_s, tid); sm.hiflag = 1; }  // the short compare is exact only for bytes < 0x80
    } else {
      if (phase_a_edge(t.buf_s, bm_s, t.vlo, t.vhi, tid) & 0x80808080u) sm.hiflag = 1;
    }
    __syncthreads();

    // ---- B1: this thread's 64 consecutive bytes of the bitmap; block-wide prefix ----
    const u64 w64 = lds64(bm_s + 8u * (uint32_t)tid);
    const uint32_t wl = (uint32_t)w64, wh = (uint32_t)(w64 >> 32);
    const uint32_t c = (uint32_t)__popc(wl) + (uint32_t)__popc(wh);
    const int mylast = 64 * tid + (wh ? 63 - __clz(wh) : 31 - __clz(wl));  // offset of the run's last newline (if c)
    const uint32_t inc = warp_incl_scan(c, lane);
    const uint32_t hasb = __ballot_sync(0xffffffffu, c != 0);
    const uint32_t lower = hasb & lanemask_lt();
    const int lastW = __shfl_sync(0xffffffffu, mylast, (31 - __clz(lower)) & 31);      // valid if lower
    const int warp_last = __shfl_sync(0xffffffffu, mylast, (31 - __clz(hasb)) & 31);   // valid if hasb
    if (lane == 31) sm.wtot[warp] = ((u64)inc << 32) | (u64)(uint32_t)(hasb ? warp_last + 1 : 0);
    __syncthreads();
    uint32_t wbase, T;
    int prevw_last1;  // (offset of the last newline of the earlier warps) + 1, 0 = none
    {
      const u64 e = lane < NWARPS ? sm.wtot[lane] : 0ull;
      const uint32_t ec = (uint32_t)(e >> 32), el = (uint32_t)e;
      const uint32_t einc = warp_incl_scan(ec, lane);
      T = __shfl_sync(0xffffffffu, einc, NWARPS - 1);
      wbase = __shfl_sync(0xffffffffu, einc - ec, warp);
      const uint32_t hw = __ballot_sync(0xffffffffu, el != 0);
      const uint32_t lw = hw & ((1u << warp) - 1u);
      prevw_last1 = lw ? (int)__shfl_sync(0xffffffffu, el, (31 - __clz(lw)) & 31) : 0;
      if (warp == 0) {
        const int tile_last1 = hw ? (int)__shfl_sync(0xffffffffu, el, (31 - __clz(hw)) & 31) : 0;
        const uint32_t tail = T ? (uint32_t)(t.vhi - tile_last1) : (uint32_t)(t.vhi - t.vlo);
        tile_prefix(sm, a, t, lane, T, tail);
        totalT += T;
      }
    }
    __syncthreads();

    // ---- B2: group descriptors and items ----
    const bool walker = T + 4u > (uint32_t)QCAP;
    const uint32_t line_bound = (T >> 2) + 2u;                      // quality lines with bytes in this tile, at most
    const uint32_t pt_limit = sm.hiflag ? 257u : 516u;              // 16-bit halves: lines a cell can take
    const bool dense = line_bound > pt_limit;
    const bool pt_flush = !CORE && !dense && sm.pt_lines + line_bound > pt_limit;
    const uint32_t cnt0 = sm.in.cnt + wbase + inc - c;
    const uint32_t lc0 = wbase + inc - c;
    int last0;
    if (lower) last0 = lastW;
    else if (prevw_last1) last0 = prevw_last1 - 1;
    else { const u64 op = sm.in.open; last0 = t.vlo - 1 - (int)(op < (u64)OPEN_CLIP ? op : (u64)OPEN_CLIP); }
    if (pt_flush) flush_ptab(sm, tid);
    if (walker) {
      phase_walk<CORE>(sm, a, t, tid, cnt0, lc0, last0);
    } else {
      if (!edge) phase_b2<CORE, false>(sm, a, t, gi_s, tid, wl, wh, cnt0, lc0, last0);
      else phase_b2_edge<CORE>(sm, a, t, gi_s, tid, wl, wh, cnt0, lc0, last0);
    }
    __syncthreads();
    if (tid == 0 && !CORE) sm.pt_lines = dense ? sm.pt_lines : (pt_flush ? line_bound : sm.pt_lines + line_bound);

    if (!walker) {
      // ---- C: one lane per group without a newline ----
      {
        const uint32_t g0 = (uint32_t)(warp * (GPT * 32) + lane);
        uint32_t gi[GPT];
#pragma unroll
        for (int j = 0; j < GPT; j++) gi[j] = lds16(gi_s + 2u * (g0 + 32u * j));
#pragma unroll
        for (int j = 0; j < GPT; j++) {
          const bool on = CORE ? (gi[j] & 3u) == 1u : (gi[j] & 1u) != 0;
          if (on) {
            const uint32_t ga = t.buf_s + 16u * (g0 + 32u * j);
            const uint4 v = lds128(ga);
            hist16(ksel, v, hist_s + ((gi[j] & 2u) << 9));
            if (!CORE && (gi[j] & 2u)) pos16(sm, v, ga, ptab_s, (gi[j] >> 3) + 16u, 0u, 16u, (gi[j] >> 2) & 1u, dense, over);
          }
        }
      }
      // ---- D: the items ----
      const uint32_t nitems = sm.nitems < (uint32_t)QCAP ? sm.nitems : (uint32_t)QCAP;
      for (uint32_t i = (uint32_t)tid; i < nitems; i += THREADS) {
        const u64 e = lds64(queue_s + 8u * i);
        const uint32_t e0 = (uint32_t)e, pos = (uint32_t)(e >> 32);
        const uint32_t g = e0 & 2047u, lo = (e0 >> 11) & 31u, hi = (e0 >> 16) & 31u, ql = (e0 >> 21) & 1u;
        const uint32_t ga = t.buf_s + 16u * g;
        if (hi > lo) {
          uint4 v = lds128(ga);
          const uint4 ml = lds128(masks_s + 16u * lo), mh = lds128(masks_s + 16u * hi);
          v.x &= mh.x & ~ml.x; v.y &= mh.y & ~ml.y; v.z &= mh.z & ~ml.z; v.w &= mh.w & ~ml.w;
          hist16(ksel, v, hist_s + (ql << 10));
          atomicAdd(&sm.junk[ql], 16u - (hi - lo));
          if (!CORE && ql) pos16(sm, v, ga, ptab_s, pos + 16u - lo, lo, hi, (e0 >> 24) & 1u, dense, over);
        }
        if (e0 & (1u << 22)) {
          const bool fn = (e0 >> 23) & 1u;
          const int off = (int)(16u * g + hi);
          const u64 rawlen = fn ? sm.in.open + (u64)(off - t.vlo) : (u64)(pos + (hi - lo));
          line_end<CORE>(sm, a, t, off, rawlen, fn, ql);
        }
      }
    }
    // ---- launch edges: a '\r' whose successor lies in another launch ----
    if (tid == 0) {
      if (t.tile == 0) {  // the '\r' that ended the previous launch is content unless this launch starts with '\n'
        const Carry& c = *a.carry;  // (bytes, open_len, last_byte are rewritten only by the CTA that exits last)
        const uint32_t cls = sm.in.cnt & 3u;
        const bool counted = CORE ? cls == 1u : (cls & 1u) != 0;
        if (counted && c.bytes && c.open_len && c.last_byte == '\r' && lds8(t.buf_s + (uint32_t)t.vlo) != '\n') {
          atomicAdd(&sm.hist[cls >> 1]['\r'], 1u);
          if (!CORE && cls == 3u) {
            const u64 p = c.open_len - 1;
            if (p < (u64)POS_BINS) atomicAdd(&sm.pos32[(uint32_t)p], (uint32_t)'\r'); else atomicAdd(&sm.over, (u64)'\r');
          }
        }
      }
      if (t.tile + 1 == a.ntiles && lds8(t.buf_s + (uint32_t)t.vhi - 1u) == '\r') {  // counted where it stands: left to the next launch / finish()
        const uint32_t cls = (sm.in.cnt + T) & 3u;
        const bool counted = CORE ? cls == 1u : (cls & 1u) != 0;
        if (counted) {
          atomicSub(&sm.hist[cls >> 1]['\r'], 1u);
          if (!CORE && cls == 3u) {
            const u64 p = sm.in.open_out - 1;
            if (p < (u64)POS_BINS) atomicSub(&sm.pos32[(uint32_t)p], (uint32_t)'\r'); else atomicAdd(&sm.over, 0ull - '\r');
          }
        }
      }
      sm.bytes_since_flush += TILE;
    }
    __syncthreads();
    if (sm.bytes_since_flush >= FLUSH_BYTES) flush_all(sm, a.acc, tid, over);
  }
  flush_all(sm, a.acc, tid, over);
  // ---- the CTA that exits last advances the stream carry and leaves the control words clean ----
  if (tid == 0) {
    if (totalT) atomicAdd(reinterpret_cast<unsigned long long*>(a.ctl + CTL_TOTAL_T), totalT);
    __threadfence();
    const u64 prev = atomicAdd(reinterpret_cast<unsigned long long*>(a.ctl + CTL_DONE), 1ull);
    sm.tile[0] = prev + 1 == (u64)gridDim.x;
  }
  __syncthreads();
  if (!sm.tile[0]) return;
  __threadfence();
  if (a.unknown && !CORE) detach_head(a, tid);
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    Carry* c = a.carry;
    c->lines += ld_relaxed_gpu(a.ctl + CTL_TOTAL_T);
    c->open_len = ld_relaxed_gpu(a.ctl + CTL_OPEN_OUT);
    c->bytes += a.end - (u64)a.lo0;
    c->last_byte = a.base[a.end - 1];
    a.ctl[CTL_TICKET] = 0; a.ctl[CTL_DONE] = 0; a.ctl[CTL_TOTAL_T] = 0; a.ctl[CTL_OPEN_OUT] = 0;
    a.ctl[CTL_FIRST_NL] = 0; a.ctl[CTL_HEAD] = 0; a.ctl[CTL_HEAD_P0] = 0;
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

cudaError_t scan_configure() {
  cudaError_t e = cudaFuncSetAttribute(fq_scan_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(fq_scan_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
}

cudaError_t launch_reset(u64* acc, Carry* carry, u64* ctl, cudaStream_t st) {
  fq_reset_kernel<<<1, 512, 0, st>>>(acc, carry, ctl);
  return cudaGetLastError();
}

u64 scan_tiles(const void* ptr, size_t nbytes) { return (((u64)((uintptr_t)ptr & 15) + nbytes) + TILE - 1) / TILE; }

// Scans `nbytes` at `ptr` (any alignment) as the continuation of the stream described by `carry`: ONE launch.
// `state` holds at least scan_tiles(ptr, nbytes) words; `epoch` (1..255) differs from that of every earlier launch
// since the words were last zeroed.
cudaError_t launch_scan(const void* ptr, size_t nbytes, Carry* carry, ShardInfo* shard, u64* acc, u64* state, u64* ctl,
                        uint32_t epoch, bool unknown_start, int resident, bool core_only, cudaStream_t st) {
  if (nbytes == 0) return cudaSuccess;
  const uintptr_t addr = (uintptr_t)ptr;
  ScanArgs a;
  a.lo0 = (uint32_t)(addr & 15);
  a.base = (const uint8_t*)(addr - a.lo0);
  a.end = (u64)a.lo0 + nbytes;
  a.ntiles = (uint32_t)((a.end + TILE - 1) / TILE);
  a.carry = carry; a.shard = shard; a.acc = acc; a.state = state; a.ctl = ctl;
  a.epoch = epoch; a.unknown = unknown_start ? 1u : 0u;
  a.dbg = getenv("FQGPU_DBG") ? (uint32_t)atoi(getenv("FQGPU_DBG")) : 0u;
  const unsigned grid = a.ntiles < (uint32_t)resident ? a.ntiles : (unsigned)resident;
  if (core_only) fq_scan_kernel<true><<<grid, THREADS, sizeof(Smem), st>>>(a);
  else fq_scan_kernel<false><<<grid, THREADS, sizeof(Smem), st>>>(a);
  return cudaGetLastError();
}

}  // namespace fq
