// fq_scan.cu -- the FASTQ scanning hot path on sm_100a.
//
// Replaces the per-line loop of the reference (src/fq_count.nim:38-45: `for line in lines(stream)`,
// i mod 4 classing, count("G")+count("C"), count("N"), line.len) and the quality fold of
// src/fq_meta.nim:245-246 with one pass over the bytes.
//
// A launch cuts its byte range into SPANS (one persistent CTA each, 2 CTAs per SM); a CTA streams its
// span tile by tile (16 KiB, 1-D TMA into a three-stage shared-memory ring) and never waits for
// another CTA:
//
//   resync  fq_resync_kernel guesses the line phase (line number mod 4) at every span start from the
//           content: the first line that starts with '@' and whose line+2 starts with '+' is a header.
//   K1a     boundary classification: every warp turns 16-byte groups into '\n' masks (SWAR compare,
//           IDP.4A movemask) -> tile bitmap.
//   K1b     four SCANNER warps popc / prefix-sum the bitmap into the tile's newline index and keep the
//           span's running line count / open-line length (the tile-edge record carry).
//   K1c     LINE tasks: one thread per sequence / quality line with bytes in the tile: line-length
//           tables, the '\r' rule, and the line's bytes FLATTENED into entries of 16-byte groups --
//           full groups (per class) and partial groups (first / last group of a line, byte range).
//   K2      WORKER warps consume the entry lists of the previous tile, 32 entries per warp step, every
//           lane one aligned 16-byte group (LDS.128): histogram addresses by IDP.4A (FMA pipe) into
//           lane-striped (conflict-free) shared-memory atomics; per-position quality sums by shared
//           atomics with immediate offsets into a bank-skewed table.  No per-line control flow, no
//           alignment shifts; lines that cross tile edges are simply two runs of entries.
//   stitch  fq_stitch_kernel prefix-sums the span descriptors, VERIFIES every guessed phase against
//           the exact line counts (a wrong guess -- malformed input -- marks the span for an exact
//           second pass, so results are exact on any input), commits the span blocks, accounts the
//           head fragment of every span (the bytes before its first newline, whose line started in
//           an earlier span) and advances the stream carry (fq::Carry): chunk-edge carry on device.
//   K3      fq_reduce_kernel folds the per-span counter blocks.
//
// Every input byte is read from HBM exactly once (twice only in rescanned spans).  Measured constants
// behind these choices are in profiles/microbench.  Line semantics are Nim's
// streams.lines: split at '\n', drop one '\r' directly before it; the trailing unterminated line is
// accounted by the host from fq::Carry at finish().
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fq_layout.h"

namespace fq {

typedef unsigned long long u64;

constexpr int TILE = 16384;
constexpr int THREADS = 512;
constexpr int NWARPS = THREADS / 32;
constexpr int SCAN_WARPS = 4;
constexpr int SCAN_THREADS = SCAN_WARPS * 32;
constexpr int LINE_WARPS = 4;                 // warps 0..LINE_WARPS-1 run the line tasks (the scanner warps first)
constexpr int LINE_THREADS = LINE_WARPS * 32;
constexpr int WORK_WARPS = NWARPS - LINE_WARPS;
constexpr int WORK_THREADS = WORK_WARPS * 32;
constexpr int BM_WORDS = TILE / 32;           // bitmap words per tile
constexpr int WPS = BM_WORDS / SCAN_THREADS;  // bitmap words per scanner thread (4)
constexpr int NL_CAP = 1024;                  // newline index capacity; denser tiles take the walker path
constexpr int PART_CAP = NL_CAP / 2 + 8;      // partial-group slots per class (two per line)
constexpr int REC_CAP = NL_CAP / 4 + 8;       // line records per class
constexpr int KMAX = 32;                      // lines with more full groups are "long": taken by all worker warps together
constexpr int LONG_CAP = 32;                  // long lines per tile (each has > KMAX * 16 bytes in the tile)
constexpr int PAD = 16;
constexpr int STAGE_BYTES = PAD + TILE + 16;
constexpr int NSTAGE = 3;
constexpr int HB = 128;                       // striped histogram bins (tiles with bytes >= 128 take the walker path)
// Per-position quality sums: 16-bit pairs.  q = position + 16; pair A = q >> 1 lives in cell (r, c) with
// 8 c + r = A, r < 16 (rows >= 8 alias (r - 8, c + 1)); a group adds its nine pairs at immediate offsets
// r0 + PT_STRIDE * i, consecutive groups of a line hit consecutive banks.  Two copies (alternating with the
// line number, 16 banks apart) keep lines of equal alignment from colliding.
constexpr int PT_STRIDE = 33;
constexpr int PT_COPY = 16 * PT_STRIDE;       // 528 words, = 16 (mod 32)
constexpr int PT_WORDS = 2 * PT_COPY;
constexpr int PT_MAX_LINES = 516;             // 16-bit halves: flush before more quality lines than this have been added
static_assert(WPS == 4, "scanner threads read their bitmap words with one LDS.128");
static_assert(LINE_WARPS >= SCAN_WARPS && WORK_WARPS > 0, "warp roles");

struct ScanArgs {
  const uint8_t* base;  // 16-byte aligned; the launch covers bytes [lo0, end) relative to base
  uint32_t lo0;
  u64 end;
  uint32_t ntiles, tps, nspans;  // tiles, tiles per span, spans
  SpanDesc* desc;
  LaunchHdr* hdr;
  Carry* carry;
  u64* pending;    // [nspans][BLOCK_WORDS] results of pass 0, committed by the stitch kernel
  u64* committed;  // [MAX_SPANS][BLOCK_WORDS] accumulated over launches
  ShardInfo* shard;  // detached head of a multi-GPU shard (rank > 0)
  uint32_t dbg;
  uint32_t core;     // FQGPU_F_CORE_ONLY: sequence lines only (what `sc fq-count` prints)
};

// IDP.4A byte selectors.  Loaded once from shared memory into registers: as immediates or kernel
// parameters the compiler re-materialises each of them (UMOV / LDCU) in front of every IDP.4A.
struct Sel {
  uint32_t h0, h1, h2, h3;  // 128 << 8k : histogram address = byte_k * 128 + base
  uint32_t p1, p2;          // 1 << 8k   : byte_k (k = 1, 2) extracted on the FMA pipe
};

struct TileMeta {
  u64 Lrel;     // newlines of the span before this tile
  u64 open;     // bytes of the open line (or of the head fragment) before this tile
  u64 toff;     // byte offset of the tile relative to base
  int T;        // newlines in the tile
  int lo, hi;   // valid byte range of the tile
  int walker;   // 1: dense or high-byte tile -> generic bitmap walker
  int R;        // sequence / quality lines with bytes in the tile (line tasks)
  int first_q;  // line task 0 is a quality line (the classes of the tasks alternate)
  uint32_t K;      // most full groups of a (not long) line of the tile: the workers map slot x -> (line x / K, group x % K)
  uint32_t nlong;  // long lines of the tile
};

// What a line task leaves for the workers (per class):
//   rec   first full group << 4 | full groups (0 for long lines) << 14 | q << 20 -- the line's aligned 16-byte groups
//   part  group | lo << 10 | hi << 14 | q << 19  (0 = empty)                      -- bytes [lo, hi) of its first / last group
//   longl first full group | full groups << 10 | q << 21 | class << 31           -- lines with more than KMAX full groups
// q = 16 + (line position of the group's byte 0), saturated at 1023 (positions >= POS_BINS all fall into
// the overflow bin).
struct __align__(128) Smem {
  uint8_t buf[NSTAGE][STAGE_BYTES];  // tile stages; data at buf[s] + PAD
  uint32_t hist[2][HB * 32];         // [0] sequence, [1] quality; word index = byte*32 + lane
  uint32_t ghist[2][256];            // un-striped tables of the generic paths
  uint32_t bitmap[2][BM_WORDS];      // bit b of word w: byte 32*w+b is '\n'
  uint16_t nl[NL_CAP];
  uint32_t rec[2][2][REC_CAP];       // [slot][class]
  uint32_t part[2][2][PART_CAP];     // [slot][class]; walker tiles: scratch for the per-word newline counts
  uint32_t longl[2][LONG_CAP];
  uint32_t inv[KMAX + 1];            // ceil(2^32 / K)
  uint32_t seq_len[POS_BINS + 2];
  uint32_t qual_len[POS_BINS + 2];
  uint32_t seq_log2[LOG2_BINS];
  uint32_t ptab[PT_WORDS];           // per-position quality sums (16-bit pairs, two copies)
  uint32_t gpos[POS_BINS + 2];       // the same, linear and 32-bit: generic paths (walker, groups that straddle POS_BINS)
  uint4 masks[17];                   // masks[n]: the first n bytes of a group
  TileMeta meta[2];
  uint32_t scan_tot[SCAN_WARPS];
  int last_nl;                       // position of the tile's last newline
  u64 full_bar[NSTAGE];              // mbarriers of the stages
  u64 len_min[2], len_max[2];        // [0] seq, [1] qual
  u64 run_L, run_open;               // span-running newline count / open-line bytes
  u64 head_len, junk[2], pos_over;
  uint32_t bytes_since_flush, hiflag[2], head_done, post[2], q_acc;
  uint32_t ksel[8];
};

static_assert(sizeof(Smem) <= 115712, "two CTAs per SM");

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar_s, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_s), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar_s, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar_s, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tFQ_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra FQ_DONE;\n\tbra FQ_WAIT;\n\tFQ_DONE:\n\t}" ::"r"(bar_s), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(uint32_t dst_s, const void* src, uint32_t bytes, uint32_t bar_s) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic reads of dst before the async write
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_s), "l"(src), "r"(bytes), "r"(bar_s) : "memory");
}
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds8(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts16(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((uint16_t)v) : "memory");
}
__device__ __forceinline__ void red_inc(uint32_t addr) {
  asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr) : "memory");
}
template <int OFF>
__device__ __forceinline__ void red_add_at(uint32_t addr, uint32_t v) {
  asm volatile("red.shared.add.u32 [%0+%2], %1;" ::"r"(addr), "r"(v), "n"(OFF) : "memory");
}

// 0x80 in every byte lane of w that equals '\n' (exact: no carries cross byte lanes)
__device__ __forceinline__ uint32_t nl_flags(uint32_t w) {
  uint32_t x = w ^ 0x0A0A0A0Au;
  return ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u;
}
// the same for words whose bytes are all < 0x80 (one operation less)
__device__ __forceinline__ uint32_t nl_flags_ascii(uint32_t w) {
  return ~((w ^ 0x0A0A0A0Au) + 0x7F7F7F7Fu) & 0x80808080u;
}
// 16-bit mask of the '\n' bytes of a 16-byte group; the movemask is two IDP.4A chains (FMA pipe)
__device__ __forceinline__ uint32_t nl_mask16(const uint4& v) {
  uint32_t lo = __dp4a(nl_flags(v.x), 0x08040201u, __dp4a(nl_flags(v.y), 0x80402010u, 0u));
  uint32_t hi = __dp4a(nl_flags(v.z), 0x08040201u, __dp4a(nl_flags(v.w), 0x80402010u, 0u));
  return (lo >> 7) | (hi << 1);  // the flags weigh 128
}
__device__ __forceinline__ uint32_t nl_mask16_ascii(const uint4& v) {
  uint32_t lo = __dp4a(nl_flags_ascii(v.x), 0x08040201u, __dp4a(nl_flags_ascii(v.y), 0x80402010u, 0u));
  uint32_t hi = __dp4a(nl_flags_ascii(v.z), 0x08040201u, __dp4a(nl_flags_ascii(v.w), 0x80402010u, 0u));
  return (lo >> 7) | (hi << 1);
}
__device__ __forceinline__ uint32_t bfind(uint32_t m) {  // index of the highest set bit, 0xFFFFFFFF for 0
  uint32_t k;
  asm("bfind.u32 %0, %1;" : "=r"(k) : "r"(m));
  return k;
}
// The (up to) two highest newlines of bitmap word m -> nl[] slots ending at shared address `end`
// (exclusive), positions relative to pos0; returns the bits that are left.  (Lines of 32 bytes and more put at
// most two newlines into a 32-byte word -- "...\n+\n" -- so the loop for the rest hardly ever runs.)
__device__ __forceinline__ uint32_t nl_extract2(uint32_t m, uint32_t end, uint32_t pos0) {
  {
    const uint32_t k = bfind(m);
    if (m) asm volatile("st.shared.u16 [%0+-2], %1;" ::"r"(end), "h"((uint16_t)(pos0 + k)) : "memory");
    m &= ~(1u << (k & 31u));
  }
  {
    const uint32_t k = bfind(m);
    if (m) asm volatile("st.shared.u16 [%0+-4], %1;" ::"r"(end), "h"((uint16_t)(pos0 + k)) : "memory");
    m &= ~(1u << (k & 31u));
  }
  return m;
}
__device__ __noinline__ void nl_extract_rest(uint32_t m, uint32_t end, uint32_t pos0) {  // (rare: kept out of the hot code)
  while (m) {
    const uint32_t k = bfind(m);
    end -= 2u;
    sts16(end, pos0 + k);
    m ^= 1u << k;
  }
}
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t n = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v += n;
  }
  return v;
}
__device__ __forceinline__ unsigned log2_bin(u64 len) { return len ? 64 - __clzll((long long)len) : 0; }

// One byte of a sequence (cls 1) / quality (cls 3) line -- stitch kernel, linear tables.
__device__ __forceinline__ void account_byte(uint32_t (*ghist)[256], uint32_t* pos_sum, int cls, uint32_t b, u64 pos) {
  atomicAdd(&ghist[cls == 3][b], 1u);
  if (cls == 3) {
    uint32_t p = pos < (u64)POS_BINS ? (uint32_t)pos : (uint32_t)POS_BINS;
    atomicAdd(&pos_sum[p], b);
  }
}
// The same byte through the scan kernel's tables (generic paths).
__device__ __forceinline__ void account_byte_tab(Smem& sm, int cls, uint32_t b, u64 pos, u64& over) {
  atomicAdd(&sm.ghist[cls == 3][b], 1u);
  if (cls == 3) {
    if (pos < (u64)POS_BINS) atomicAdd(&sm.gpos[(uint32_t)pos], b);
    else over += b;
  }
}
// A line length through the shared tables (generic paths; the line tasks keep their extrema in registers).
__device__ __noinline__ void account_line_len(Smem& sm, int cls, u64 len) {
  const int q = cls == 3;
  const uint32_t bin = len < (u64)POS_BINS ? (uint32_t)len : (uint32_t)POS_BINS;
  if (q) atomicAdd(&sm.qual_len[bin], 1u);
  else { atomicAdd(&sm.seq_len[bin], 1u); atomicAdd(&sm.seq_log2[log2_bin(len)], 1u); }
  atomicMin(&sm.len_min[q], len);
  atomicMax(&sm.len_max[q], len);
}

// 16 bytes into the lane-striped histogram at hbase (bytes < 128): one IDP.4A and one shared atomic per byte.
__device__ __forceinline__ void hist16(const Sel& k, const uint4& v, uint32_t hbase) {
  red_inc(__dp4a(v.x, k.h0, hbase)); red_inc(__dp4a(v.x, k.h1, hbase)); red_inc(__dp4a(v.x, k.h2, hbase)); red_inc(__dp4a(v.x, k.h3, hbase));
  red_inc(__dp4a(v.y, k.h0, hbase)); red_inc(__dp4a(v.y, k.h1, hbase)); red_inc(__dp4a(v.y, k.h2, hbase)); red_inc(__dp4a(v.y, k.h3, hbase));
  red_inc(__dp4a(v.z, k.h0, hbase)); red_inc(__dp4a(v.z, k.h1, hbase)); red_inc(__dp4a(v.z, k.h2, hbase)); red_inc(__dp4a(v.z, k.h3, hbase));
  red_inc(__dp4a(v.w, k.h0, hbase)); red_inc(__dp4a(v.w, k.h1, hbase)); red_inc(__dp4a(v.w, k.h2, hbase)); red_inc(__dp4a(v.w, k.h3, hbase));
}
// Per-position sums of the bytes [lo, hi) of a quality group (bytes outside are zero in v): q = 16 + position
// of the group's byte 0, pt_s = the table copy.  Inside the table: the group, shifted by one byte when q is odd,
// is nine 16-bit pairs -> nine shared atomics at immediate offsets; beyond POS_BINS: the overflow bin; the one
// group per long line that straddles POS_BINS: byte-wise into the linear table.
__device__ __forceinline__ void pos16(Smem& sm, const uint4& v, uint32_t ga, uint32_t pt_s, uint32_t q, uint32_t lo, uint32_t hi, u64& over) {
  if (q + hi <= (uint32_t)POS_BINS + 16u) {
    const uint32_t sh = (q & 1u) << 3;
    const uint32_t w0 = v.x << sh, w1 = __funnelshift_l(v.x, v.y, sh), w2 = __funnelshift_l(v.y, v.z, sh);
    const uint32_t w3 = __funnelshift_l(v.z, v.w, sh), w4 = __funnelshift_l(v.w, 0u, sh);
    const uint32_t A = q >> 1;
    const uint32_t r0 = pt_s + 4u * ((A & 7u) * PT_STRIDE + (A >> 3));
    red_add_at<4 * PT_STRIDE * 0>(r0, __byte_perm(w0, 0u, 0x4140)); red_add_at<4 * PT_STRIDE * 1>(r0, __byte_perm(w0, 0u, 0x4342));
    red_add_at<4 * PT_STRIDE * 2>(r0, __byte_perm(w1, 0u, 0x4140)); red_add_at<4 * PT_STRIDE * 3>(r0, __byte_perm(w1, 0u, 0x4342));
    red_add_at<4 * PT_STRIDE * 4>(r0, __byte_perm(w2, 0u, 0x4140)); red_add_at<4 * PT_STRIDE * 5>(r0, __byte_perm(w2, 0u, 0x4342));
    red_add_at<4 * PT_STRIDE * 6>(r0, __byte_perm(w3, 0u, 0x4140)); red_add_at<4 * PT_STRIDE * 7>(r0, __byte_perm(w3, 0u, 0x4342));
    if (sh) red_add_at<4 * PT_STRIDE * 8>(r0, w4);
  } else if (q + lo >= (uint32_t)POS_BINS + 16u) {
    over += __dp4a(v.x, 0x01010101u, __dp4a(v.y, 0x01010101u, __dp4a(v.z, 0x01010101u, __dp4a(v.w, 0x01010101u, 0u))));
  } else {
    for (uint32_t x = lo; x < hi; x++) {
      const uint32_t b = lds8(ga + x), p = q - 16u + x;
      if (p < (uint32_t)POS_BINS) atomicAdd(&sm.gpos[p], b);
      else over += b;
    }
  }
}

// Byte before / after a tile position, read from global memory when it lies outside the tile.
// before(): the launch's first byte is preceded by the stream's last byte so far (LaunchHdr).
// after(): -1 when the position is the end of the launch (the byte's fate is decided later).
__device__ __forceinline__ int byte_before(const ScanArgs& a, const uint8_t* buf, const TileMeta& m, int o) {
  if (o > m.lo) return buf[o - 1];
  const u64 g = m.toff + (u64)o;
  if (g > (u64)a.lo0) return a.base[g - 1];
  return a.hdr->bytes0 ? (int)a.hdr->last_byte0 : 0;
}
__device__ __forceinline__ int byte_after_tile(const ScanArgs& a, const TileMeta& m) {
  const u64 g = m.toff + (u64)m.hi;
  return g < a.end ? (int)a.base[g] : -1;
}

// Generic bitmap walker (dense tiles, tiles with bytes >= 128): one thread per 32-byte bitmap word,
// bytes taken one at a time.  Exact for any content; also keeps the line-length tables.
__device__ __noinline__ void tile_walker(Smem& sm, const ScanArgs& a, const uint8_t* buf, const TileMeta& m, uint32_t phase,
                                         const uint32_t* bitmap, const uint16_t* wordbase, int r, int nthr) {
  u64 over = 0;
  const int nwords = (m.hi + 31) >> 5;
  for (int w = r; w < nwords; w += nthr) {
    const int o0 = w * 32 > m.lo ? w * 32 : m.lo;
    const int o1 = (w * 32 + 32) < m.hi ? (w * 32 + 32) : m.hi;
    if (o0 >= o1) continue;
    u64 line = m.Lrel + wordbase[w];  // span-relative line index of byte o0
    int prev = -1;                    // offset of the newline preceding byte o0 inside the tile, or -1
    {
      const uint32_t below = bitmap[w] & ((o0 & 31) ? ((1u << (o0 & 31)) - 1u) : 0u);
      line += __popc(below);
      if (below) prev = w * 32 + 31 - __clz(below);
      else for (int x = w - 1; x >= 0; x--) { const uint32_t bwx = bitmap[x]; if (bwx) { prev = x * 32 + 31 - __clz(bwx); break; } }
    }
    u64 pos = prev >= 0 ? (u64)(o0 - prev - 1) : m.open + (u64)(o0 - m.lo);  // raw position of byte o0 in its line
    for (int o = o0; o < o1; o++) {
      const uint32_t b = buf[o];
      const int cls = (int)((phase + line) & 3);
      const bool counted = (cls & 1) && line != 0 && !(a.core && cls == 3);  // line 0 of the span is the head fragment (stitch kernel)
      if (b == '\n') {
        if (counted) {
          const int cr = (pos > 0 && byte_before(a, buf, m, o) == '\r') ? 1 : 0;
          account_line_len(sm, cls, pos - (u64)cr);
        }
        line++;
        pos = 0;
        continue;
      }
      if (counted) {
        bool content = true;
        if (b == '\r') {  // content unless the next byte is '\n'; at the end of the launch: decided later
          const int nx = (o + 1 < m.hi) ? (int)buf[o + 1] : byte_after_tile(a, m);
          content = nx != '\n' && nx >= 0;
        }
        if (content) account_byte_tab(sm, cls, b, pos, over);
      }
      pos++;
    }
  }
  if (over) atomicAdd(&sm.pos_over, over);
}

// Adds the packed per-position table (both copies, cell and alias cell of every pair) to the span block and clears it.
template <class S>
__device__ __forceinline__ void flush_pos_tab(S& sm, u64* block, int tid) {
  for (int t = tid; t < POS_BINS / 2; t += (int)blockDim.x) {
    const uint32_t A = (uint32_t)t + 8u;  // pair of q = 2A, 2A+1 -> positions 2t, 2t+1
    const uint32_t c0 = (A & 7u) * PT_STRIDE + (A >> 3), c1 = c0 + 8u * PT_STRIDE - 1u;
    u64 lo = 0, hi = 0;
#pragma unroll
    for (int k = 0; k < 2; k++) {
      const uint32_t x = sm.ptab[c0 + k * PT_COPY], y = sm.ptab[c1 + k * PT_COPY];
      sm.ptab[c0 + k * PT_COPY] = 0; sm.ptab[c1 + k * PT_COPY] = 0;
      lo += (x & 0xFFFFu) + (y & 0xFFFFu);
      hi += (x >> 16) + (y >> 16);
    }
    if (lo) block[OFF_POS_SUM + 2 * t] += lo;
    if (hi) block[OFF_POS_SUM + 2 * t + 1] += hi;
  }
}
// The linear table of the generic paths.
template <class S>
__device__ __forceinline__ void flush_gpos(S& sm, u64* block, int tid) {
  for (int p = tid; p < POS_BINS; p += (int)blockDim.x) {
    const uint32_t v = sm.gpos[p];
    if (v) { block[OFF_POS_SUM + p] += v; sm.gpos[p] = 0; }
  }
}

__device__ __noinline__ void flush_pos_tab_cold(Smem& sm, u64* block, int tid) { flush_pos_tab(sm, block, tid); }
__device__ __noinline__ void flush_gpos_cold(Smem& sm, u64* block, int tid) { flush_gpos(sm, block, tid); }

__device__ __forceinline__ uint32_t part_entry(int g, int lo, int hi, uint32_t q) {
  return (uint32_t)g | ((uint32_t)lo << 10) | ((uint32_t)hi << 14) | ((q < 1023u ? q : 1023u) << 19);
}

// Rare forms of K1a, kept out of the hot code: this thread's groups again with the exact compare; an edge tile
// (bytes [lo, hi) valid).
__device__ __noinline__ void k1a_redo(uint32_t buf_s, uint32_t bm_s, int r, int nthr) {
  for (int g = r; g < TILE / 16; g += nthr) sts16(bm_s + 2u * (uint32_t)g, nl_mask16(lds128(buf_s + 16u * (uint32_t)g)));
}
__device__ __noinline__ uint32_t k1a_edge(Smem& sm, uint32_t buf_s, uint32_t bm_s, int lo, int hi, int r, int nthr) {
  uint32_t hib = 0;
  for (int g = r; g < TILE / 16; g += nthr) {
    const int off = g * 16;
    uint32_t m = 0;
    if (off < hi && off + 16 > lo) {
      const uint4 v = lds128(buf_s + 16u * (uint32_t)g);
      m = nl_mask16(v);
      int lo_k = lo - off; lo_k = lo_k < 0 ? 0 : lo_k;
      int hi_k = hi - off; hi_k = hi_k > 16 ? 16 : hi_k;
      m &= ((1u << hi_k) - 1u) & ~((1u << lo_k) - 1u);
      // high bytes only matter inside the valid range (stale shared memory beyond it)
      const uint4 ml = sm.masks[lo_k], mh = sm.masks[hi_k];
      hib |= ((v.x & mh.x & ~ml.x) | (v.y & mh.y & ~ml.y)) | ((v.z & mh.z & ~ml.z) | (v.w & mh.w & ~ml.w));
    }
    sts16(bm_s + 2u * (uint32_t)g, m);
  }
  return hib;
}

// K1a: newline masks of a tile's 16-byte groups -> bitmap slot (threads r of nthr; bytes [lo, hi) are valid).
// Returns the OR of the valid bytes (bit 7 of any byte set: the tile takes the generic walker path).
template <int NTHR>
__device__ __forceinline__ uint32_t k1a_tile(Smem& sm, uint32_t buf_s, uint32_t bm_s, int lo, int hi, int r) {
  constexpr int nthr = NTHR;
  uint32_t hib = 0;
  if (lo == 0 && hi == TILE) {  // interior tile: no edge handling; unrolled (immediate offsets, loads first)
    constexpr int NIT = (TILE / 16 + NTHR - 1) / NTHR;
    const uint32_t a0 = buf_s + 16u * (uint32_t)r, b0 = bm_s + 2u * (uint32_t)r;
    uint4 v[NIT];
#pragma unroll
    for (int i = 0; i < NIT; i++) {
      v[i] = make_uint4(0u, 0u, 0u, 0u);
      if ((i + 1) * NTHR <= TILE / 16 || r + i * NTHR < TILE / 16)
        v[i] = lds128(a0 + (uint32_t)(16 * i * NTHR));
    }
#pragma unroll
    for (int i = 0; i < NIT; i++) {
      hib |= (v[i].x | v[i].y) | (v[i].z | v[i].w);
      if ((i + 1) * NTHR <= TILE / 16 || r + i * NTHR < TILE / 16)
        sts16(b0 + (uint32_t)(2 * i * NTHR), nl_mask16_ascii(v[i]));
    }
    if (hib & 0x80808080u) k1a_redo(buf_s, bm_s, r, nthr);  // the short compare is exact only for bytes < 0x80
  } else {
    hib = k1a_edge(sm, buf_s, bm_s, lo, hi, r, nthr);
  }
  return hib;
}

// One aligned 16-byte group of a quality / sequence line through the tables.
template <bool QL>
__device__ __forceinline__ void group_full(Smem& sm, const Sel& k, uint32_t ga, uint32_t hb, uint32_t ptab_s, uint32_t q, u64& over) {
  const uint4 v = lds128(ga);
  hist16(k, v, hb);
  if (QL) pos16(sm, v, ga, ptab_s, q, 0u, 16u, over);
}
// Full groups of the lines of one class: slot x = K * line + group (x16 = 16 x), 32 slots per warp step.
template <bool QL>
__device__ __forceinline__ void full_slots(Smem& sm, const Sel& k, uint32_t x, uint32_t nslots, uint32_t inv, uint32_t negK16,
                                           uint32_t rec_s, uint32_t buf_s, uint32_t hb, uint32_t ptab_s, u64& over) {
  for (; x < nslots; x += WORK_WARPS * 32) {
    const uint32_t line = __umulhi(x, inv);
    const uint32_t k16 = line * negK16 + (x << 4);
    const uint32_t r = lds32(rec_s + 4u * line);
    if (k16 < ((r >> 10) & 0x3F0u))
      group_full<QL>(sm, k, buf_s + (r & 0x3FF0u) + k16, hb, ptab_s + ((line & 2u) ? PT_COPY * 4u : 0u), (r >> 20) + k16, over);
  }
}
// First / last groups of the lines of one class: bytes [lo, hi) of the group, the others masked to zero
// (they land in bin 0; the caller keeps the count and subtracts it at the end).
template <bool QL>
__device__ __forceinline__ uint32_t part_slots(Smem& sm, const Sel& k, uint32_t x, uint32_t nslots, uint32_t part_s, uint32_t masks_s,
                                               uint32_t buf_s, uint32_t hb, uint32_t ptab_s, u64& over) {
  uint32_t junk = 0;
  for (; x < nslots; x += WORK_WARPS * 32) {
    const uint32_t ent = lds32(part_s + 4u * x);
    if (ent) {
      const uint32_t ga = buf_s + ((ent << 4) & 0x3FF0u);
      const uint32_t lo = (ent >> 10) & 15u, hi = (ent >> 14) & 31u;
      uint4 v = lds128(ga);
      const uint4 ml = lds128(masks_s + ((ent >> 6) & 0xF0u)), mh = lds128(masks_s + ((ent >> 10) & 0x1F0u));
      v.x &= mh.x & ~ml.x; v.y &= mh.y & ~ml.y; v.z &= mh.z & ~ml.z; v.w &= mh.w & ~ml.w;
      hist16(k, v, hb);
      if (QL) pos16(sm, v, ga, ptab_s + ((x & 4u) ? PT_COPY * 4u : 0u), ent >> 19, lo, hi, over);
      junk += 16u - (hi - lo);
    }
  }
  return junk;
}

// CORE: FQGPU_F_CORE_ONLY (sequence lines only).  A template parameter, not a runtime flag: the kernel is sensitive to
// code size, and each instantiation drops the other mode's code.
template <bool CORE>
__global__ void __launch_bounds__(THREADS, 2) fq_scan_kernel(const ScanArgs a, const int pass) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool scanner = warp < SCAN_WARPS;
  const bool liner = warp < LINE_WARPS;
  const int span = blockIdx.x;
  SpanDesc& desc = a.desc[span];
  if (pass == 1 && desc.state != SPAN_RESCAN) return;
  const uint32_t phase = pass == 0 ? desc.guess : desc.exact;
  const bool count_only = phase > 3;
  // pass 0 writes the span's pending block (committed by the stitch kernel); the exact second pass adds
  // straight into the committed block
  u64* block = (pass == 0 ? a.pending : a.committed) + (size_t)span * BLOCK_WORDS;
  const uint32_t t0 = (uint32_t)span * a.tps;
  const uint32_t t1 = min(a.ntiles, t0 + a.tps);
  const uint32_t sm0 = smem_u32(smem_raw);
  const uint32_t buf0_s = sm0 + (uint32_t)offsetof(Smem, buf) + PAD;
  const uint32_t bar0_s = sm0 + (uint32_t)offsetof(Smem, full_bar);

  for (int i = tid; i < 2 * HB * 32; i += THREADS) (&sm.hist[0][0])[i] = 0;
  for (int i = tid; i < 512; i += THREADS) (&sm.ghist[0][0])[i] = 0;
  for (int i = tid; i < POS_BINS + 2; i += THREADS) { sm.seq_len[i] = 0; sm.qual_len[i] = 0; sm.gpos[i] = 0; }
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
  if (pass == 0) {
    for (int i = tid; i < BLOCK_WORDS; i += THREADS) block[i] = (i == OFF_SEQ_LEN_MIN || i == OFF_QUAL_LEN_MIN) ? ~0ull : 0ull;
  }
  if (tid == 0) {
    sm.len_min[0] = sm.len_min[1] = ~0ull;
    sm.len_max[0] = sm.len_max[1] = 0;
    sm.bytes_since_flush = 0; sm.post[0] = sm.post[1] = 0; sm.q_acc = 0;
    sm.hiflag[0] = sm.hiflag[1] = 0;
    sm.run_L = 0; sm.run_open = 0; sm.head_len = 0; sm.head_done = 0;
    sm.junk[0] = sm.junk[1] = 0; sm.pos_over = 0;
    sm.meta[0].walker = sm.meta[1].walker = 0;
    sm.meta[0].R = sm.meta[1].R = 0;
    for (int k = 0; k < 4; k++) { sm.ksel[k] = 0x80u << (8 * k); sm.ksel[4 + k] = 1u << (8 * k); }
    for (int s = 0; s < NSTAGE; s++) mbar_init(bar0_s + 8u * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (t0 < t1) {  // prologue: the first tile into stage 0
      const u64 toff = (u64)t0 * TILE;
      const uint32_t bytes = (uint32_t)((((a.end - toff) < (u64)TILE ? (a.end - toff) : (u64)TILE) + 15) & ~15ull);
      mbar_expect_tx(bar0_s, bytes);
      tma_load_1d(buf0_s, a.base + toff, bytes, bar0_s);
    }
  }
  uint32_t mns = ~0u, mxs = 0, mnq = ~0u, mxq = 0;       // line-length extrema seen by this thread
  uint32_t junk_s = 0, junk_q = 0;                      // histogram slots of masked bytes (counted in bin 0)
  u64 over = 0;                                         // quality bytes at positions >= POS_BINS
  const uint32_t hb_seq = sm0 + (uint32_t)offsetof(Smem, hist) + 4u * lane;
  const uint32_t ptab_s = sm0 + (uint32_t)offsetof(Smem, ptab);
  const uint32_t masks_s = sm0 + (uint32_t)offsetof(Smem, masks);
  __syncthreads();
  Sel ksel;
  {
    const uint32_t ks = sm0 + (uint32_t)offsetof(Smem, ksel);
    ksel.h0 = lds32(ks); ksel.h1 = lds32(ks + 4); ksel.h2 = lds32(ks + 8); ksel.h3 = lds32(ks + 12);
    ksel.p1 = lds32(ks + 20); ksel.p2 = lds32(ks + 24);
  }

  // pipeline: iteration `it` classifies tile B = t0+it (stage it%3, slot it&1) while the workers
  // take tile C = B-1; the TMA of tile A = B+1 is started at the top.
  const int nt = (int)(t1 - t0);
  const int it_first = t0 == 0 ? 0 : -1;                 // iteration whose tile starts at lo0
  const int it_last = t1 == a.ntiles ? nt - 1 : -1;      // iteration whose tile ends at a.end
  const int hi_last = (int)(a.end - (u64)(a.ntiles - 1) * TILE);
  const uint8_t* next_src = a.base + ((u64)t0 + 1) * TILE;  // tile B+1 (thread 0 only)
  uint32_t par_bits = 0;          // mbarrier parity per stage (bit s)
  int stB = 0, stC = 2, stA = 1;  // stages of tiles B, C, A; rotated at the end of every iteration
#ifdef FQ_TRACE
  unsigned long long tr[9] = {0,0,0,0,0,0,0,0,0}; long long tprev = 0;
#define TRS() tprev = clock64()
#define TR(i) { long long tnow = clock64(); tr[i] += (unsigned long long)(tnow - tprev); tprev = tnow; }
#else
#define TRS()
#define TR(i)
#endif
  {  // prologue: all warps classify the span's first tile
    mbar_wait(bar0_s, 0u);
    par_bits = 1u;
    const uint32_t hib = k1a_tile<THREADS>(sm, buf0_s, sm0 + (uint32_t)offsetof(Smem, bitmap), 0 == it_first ? (int)a.lo0 : 0,
                                           0 == it_last ? hi_last : TILE, tid);
    if (hib & 0x80808080u) sm.hiflag[0] = 1;
    __syncthreads();
  }
  for (int it = 0; it <= nt; it++) {
    TRS();
    const int sb = it & 1, sc = sb ^ 1;
    const bool haveB = it < nt, haveC = it > 0;
    const uint32_t bufB_s = buf0_s + (uint32_t)stB * STAGE_BYTES;

    if (tid == 0 && it + 1 < nt) {  // stage stA held tile B-2, whose K2 finished last iteration
      const uint32_t bytes = (it + 1 == it_last) ? (uint32_t)((hi_last + 15) & ~15) : (uint32_t)TILE;
      mbar_expect_tx(bar0_s + 8u * stA, bytes);
      tma_load_1d(buf0_s + (uint32_t)stA * STAGE_BYTES, next_src, bytes, bar0_s + 8u * stA);
    }
    next_src += TILE;

    const int loB = it == it_first ? (int)a.lo0 : 0;
    const int hiB = it == it_last ? hi_last : TILE;
    const u64 toffB = (u64)(t0 + (uint32_t)it) * TILE;
    TR(0); TR(1); TR(2);

    if (liner) {
      // =====================================================================================
      // SCANNER warps: newline index of tile B and the span-running carry
      // =====================================================================================
      u64 carry_L = 0, carry_open = 0;  // span carry before tile B (read together with the bitmap: one round trip)
      int tileT = 0, tileR = 0;
      if (scanner) {
        TR(3);
        if (haveB) {
          carry_L = sm.run_L; carry_open = sm.run_open;
          const uint4 bw4 = *reinterpret_cast<const uint4*>(&sm.bitmap[sb][tid * WPS]);
          const uint32_t c0 = __popc(bw4.x), c1 = __popc(bw4.y), c2 = __popc(bw4.z), c3 = __popc(bw4.w);
          const uint32_t c = c0 + c1 + c2 + c3;
          const uint32_t inc = warp_incl_scan(c, lane);
          if (lane == 31) sm.scan_tot[warp] = inc;
          bar_sync(2, SCAN_THREADS);
          TR(4);
          uint32_t wbase = 0, T = 0;
#pragma unroll
          for (int w = 0; w < SCAN_WARPS; w++) {
            const uint32_t x = sm.scan_tot[w];
            if (w < warp) wbase += x;
            T += x;
          }
          const uint32_t first = wbase + inc - c;  // index of this thread's first newline
          if (c && first + c == T) {               // the thread that holds the tile's last newline
            const uint32_t lw = bw4.w ? bw4.w : (bw4.z ? bw4.z : (bw4.y ? bw4.y : bw4.x));
            const int lx = bw4.w ? 3 : (bw4.z ? 2 : (bw4.y ? 1 : 0));
            sm.last_nl = (tid * WPS + lx) * 32 + 31 - __clz(lw);
          }
          const bool walker = T > (uint32_t)NL_CAP || sm.hiflag[sb] != 0;
          if (walker) {  // the walker wants the newline count before every bitmap word (scratch: the partial-group slots)
            uint16_t* wb = reinterpret_cast<uint16_t*>(&sm.part[sb][0][0]) + tid * WPS;
            wb[0] = (uint16_t)first; wb[1] = (uint16_t)(first + c0); wb[2] = (uint16_t)(first + c0 + c1); wb[3] = (uint16_t)(first + c0 + c1 + c2);
          } else if (__any_sync(0xffffffffu, c != 0)) {  // (long reads: most warps see no newline at all)
            // newline index: the two highest newlines of every bitmap word without branches (one FLO
            // each, the four words are independent chains); denser words finish in a loop
            const uint32_t nl_s = sm0 + (uint32_t)offsetof(Smem, nl);
            const uint32_t e0 = nl_s + 2u * (first + c0), e1 = e0 + 2u * c1, e2 = e1 + 2u * c2, e3 = e2 + 2u * c3;
            const uint32_t pb = (uint32_t)tid * (WPS * 32u);
            const uint32_t r0 = nl_extract2(bw4.x, e0, pb), r1 = nl_extract2(bw4.y, e1, pb + 32u);
            const uint32_t r2 = nl_extract2(bw4.z, e2, pb + 64u), r3 = nl_extract2(bw4.w, e3, pb + 96u);
            if (r0 | r1 | r2 | r3) {
              nl_extract_rest(r0, e0 - 4u, pb); nl_extract_rest(r1, e1 - 4u, pb + 32u);
              nl_extract_rest(r2, e2 - 4u, pb + 64u); nl_extract_rest(r3, e3 - 4u, pb + 96u);
            }
          }
          {
            const uint32_t ph = (uint32_t)((phase + carry_L) & 3);  // class of the tile's line 0
            // first line of the tile that gets a task: odd class (core-only: class 1), then every 2nd (4th)
            const int jr0 = CORE ? (int)((1u - ph) & 3u) : ((ph & 1) ? 0 : 1);
            tileT = (int)T;
            tileR = (walker || count_only) ? 0 : (CORE ? ((int)T + 4 - jr0) >> 2 : ((int)T + 2 - jr0) >> 1);
          }
          if (tid == 0) {
            const u64 Lrel = carry_L, open = carry_open;  // before this tile
            TileMeta& m = sm.meta[sb];
            const uint32_t ph = (uint32_t)((phase + Lrel) & 3);
            const int jr0 = CORE ? (int)((1u - ph) & 3u) : ((ph & 1) ? 0 : 1);
            m.Lrel = Lrel; m.open = open; m.toff = toffB; m.T = (int)T; m.lo = loB; m.hi = hiB;
            m.walker = (walker && !count_only) ? 1 : 0;
            m.R = (walker || count_only) ? 0 : (CORE ? ((int)T + 4 - jr0) >> 2 : ((int)T + 2 - jr0) >> 1);
            m.first_q = ((ph + (uint32_t)jr0) & 3) == 3;
            m.K = 0; m.nlong = 0;
            sm.bytes_since_flush += (uint32_t)(hiB - loB);
            sm.run_L = Lrel + T;
          }
        }
      }
      TR(5);
      bar_sync(3, LINE_THREADS);  // newline index and meta of tile B are visible to all line warps
      TR(6);
      if (tid == 0) {
        sm.hiflag[sb] = 0;
        uint32_t post = 0;  // what the whole CTA has to do after this iteration's barrier
        if (haveB) {
          const TileMeta& m = sm.meta[sb];
          const int T = m.T;
          // the 16-bit halves of the packed per-position table hold at most PT_MAX_LINES quality lines; tile B's
          // lines are added in the next iteration
          const uint32_t nqB = CORE ? 0u : (uint32_t)(m.first_q ? (m.R + 1) >> 1 : m.R >> 1);
          uint32_t qa = sm.q_acc;
          if (qa + nqB > (uint32_t)PT_MAX_LINES) { post |= 2u; qa = 0; }
          sm.q_acc = qa + nqB;
          if (m.walker) post |= 1u;
          if (sm.bytes_since_flush > (1u << 24)) { post |= 4u; sm.bytes_since_flush = 0; }  // the 32-bit sums of the generic paths
          sm.run_open = T ? (u64)(hiB - (sm.last_nl + 1)) : m.open + (u64)(hiB - loB);
          if (!sm.head_done) {  // bytes of the span before its first newline (the stitch kernel's share)
            if (T) {
              int first_nl = 0;  // lowest set bit of the bitmap (the newline index is not built for walker tiles)
              for (int w = 0; w < BM_WORDS; w++) { const uint32_t x = sm.bitmap[sb][w]; if (x) { first_nl = w * 32 + __ffs(x) - 1; break; } }
              sm.head_len = m.open + (u64)(first_nl - loB);
              sm.head_done = 1;
            } else {
              sm.head_len = m.open + (u64)(hiB - loB);
            }
          }
        }
        sm.post[sb] = post;
      }

      // =====================================================================================
      // LINE tasks of tile B: one thread per sequence / quality line with bytes in the tile
      // =====================================================================================
      TileMeta& m = sm.meta[sb];
      const int R = haveB ? tileR : 0;
      // warp 0 issues the TMA and does the serial bookkeeping of thread 0: it takes no line tasks
      if (warp >= 1 && (warp - 1) * 32 < R) {
        const uint8_t* buf = &sm.buf[stB][PAD];
        const int T = tileT, lo = loB, hi = hiB;
        const u64 Lrel = carry_L, open = carry_open;
        const uint32_t ph = (uint32_t)((phase + Lrel) & 3);
        const int jr0 = CORE ? (int)((1u - ph) & 3u) : ((ph & 1) ? 0 : 1);
        const int jshift = CORE ? 2 : 1, cshift = CORE ? 0 : 1;  // line of task i, its index within its class
        for (int base = (warp - 1) * 32; base < R; base += LINE_THREADS - 32) {
          const int i = base + lane;
          uint32_t nd = 0;
          if (i < R) {
            const int j = jr0 + (i << jshift);
            const bool qual = ((ph + (uint32_t)j) & 3) == 3;
            const bool tail = j == T;  // the line still open at the tile end
            const int e = tail ? hi : (int)sm.nl[j];
            const int s = j ? (int)sm.nl[j - 1] + 1 : lo;
            const bool carried = j == 0 && open != 0;  // the line began in an earlier tile
            const bool live = j != 0 || Lrel != 0;      // line 0 of the span is the head fragment (stitch kernel)
            // a '\r' directly before the line end is dropped (e == 0 reads the pad byte in front of the tile)
            const bool cr_here = lds8(bufB_s + (uint32_t)max(e, 1) - 1u) == '\r' && e > s;
            uint32_t cr = cr_here, drop = cr_here;
            if (tail) {  // decided by the next byte; at the launch end: later
              if (cr_here) { const int nx = byte_after_tile(a, m); drop = (nx == '\n' || nx < 0); }
            } else if (live) {
              if (carried && e == s) cr = byte_before(a, buf, m, e) == '\r';  // the '\r' ended the previous tile and was dropped there
              if (carried && open > 0x7FFF0000ull) {  // (a line of more than 2 GB)
                account_line_len(sm, qual ? 3 : 1, open + (u64)(e - s) - (u64)cr);
              } else {
                const uint32_t len = (carried ? (uint32_t)open : 0u) + (uint32_t)(e - s) - cr;
                const uint32_t bin = len < (uint32_t)POS_BINS ? len : (uint32_t)POS_BINS;
                atomicAdd(qual ? &sm.qual_len[bin] : &sm.seq_len[bin], 1u);
                if (qual) { mnq = min(mnq, len); mxq = max(mxq, len); }
                else { atomicAdd(&sm.seq_log2[32 - __clz(len)], 1u); mns = min(mns, len); mxs = max(mxs, len); }
              }
            }
            // the line's bytes [s, xe) as 16-byte groups: head part, full groups, tail part
            const int xe = e - (int)drop;
            const bool has = live && xe > s;
            const uint32_t poff = carried ? (open < 2048ull ? (uint32_t)open : 2048u) : 0u;
            const int g0 = s >> 4, g1 = xe >> 4, a4 = s & 15, b4 = xe & 15;
            const bool same = g0 == g1;
            const uint32_t q0 = 16u + poff - (uint32_t)a4;  // q of group g0
            const int gf = g0 + (a4 != 0);
            const uint32_t pe0 = (has && (a4 != 0 || same)) ? part_entry(g0, a4, same ? b4 : 16, q0) : 0u;
            const uint32_t pe1 = (has && !same && b4 != 0) ? part_entry(g1, 0, b4, q0 + 16u * (uint32_t)(g1 - g0)) : 0u;
            const uint32_t nfull = (has && !same) ? (uint32_t)(g1 - gf) : 0u;
            const uint32_t gf0 = nfull ? (uint32_t)gf : 0u;  // (gf may be one past the tile)
            uint32_t qf0 = q0 + 16u * (uint32_t)(gf - g0);
            qf0 = qf0 < 1023u ? qf0 : 1023u;
            const int cl = qual ? 1 : 0;
            nd = nfull;
            if (nfull > (uint32_t)KMAX) {
              const uint32_t ix = atomicAdd(&m.nlong, 1u);
              sm.longl[sb][ix] = gf0 | (nfull << 10) | (qf0 << 21) | ((uint32_t)cl << 31);
              nd = 0;
            }
            sm.rec[sb][cl][i >> cshift] = (gf0 << 4) | (nd << 14) | (qf0 << 20);
            *reinterpret_cast<uint2*>(&sm.part[sb][cl][2 * (i >> cshift)]) = make_uint2(pe0, pe1);
          }
          const uint32_t kmax = __reduce_max_sync(0xffffffffu, nd);
          if (lane == 0 && kmax) atomicMax(&m.K, kmax);
        }
      }
    } else {
      // =====================================================================================
      // WORKER warps: byte statistics of tile C from its line records
      // =====================================================================================
      const TileMeta& m = sm.meta[sc];
      if (haveC && m.R > 0) {
        const uint32_t buf_s = buf0_s + (uint32_t)stC * STAGE_BYTES;
        const int R = m.R, n_a = (R + 1) >> 1, n_b = R >> 1;  // lines of the class of task 0 / of the other class
        const uint32_t nlq = CORE ? 0u : (uint32_t)(m.first_q ? n_a : n_b), nls = CORE ? (uint32_t)R : (uint32_t)(m.first_q ? n_b : n_a);
        uint32_t K = m.K;
        K = K == 1u ? 2u : K;
        const uint32_t inv = sm.inv[K], negK16 = 0u - 16u * K;
        const uint32_t rec_s = sm0 + (uint32_t)offsetof(Smem, rec) + (uint32_t)sc * (2u * REC_CAP * 4u);
        const uint32_t part_s = sm0 + (uint32_t)offsetof(Smem, part) + (uint32_t)sc * (2u * PART_CAP * 4u);
        // the four phases start at different warps (rotating from tile to tile) so that the ragged last rounds
        // of their 32-slot steps fall on different warps
        const int ww = warp - LINE_WARPS;
        int w0 = ww - (it % WORK_WARPS); w0 = w0 < 0 ? w0 + WORK_WARPS : w0;
        int w1 = w0 - WORK_WARPS / 4;     w1 = w1 < 0 ? w1 + WORK_WARPS : w1;
        int w2 = w0 - WORK_WARPS / 2;     w2 = w2 < 0 ? w2 + WORK_WARPS : w2;
        int w3 = w0 - 3 * WORK_WARPS / 4; w3 = w3 < 0 ? w3 + WORK_WARPS : w3;
        full_slots<true>(sm, ksel, (uint32_t)(w0 * 32 + lane), nlq * K, inv, negK16, rec_s + REC_CAP * 4u, buf_s, hb_seq + HB * 32 * 4, ptab_s, over);
        full_slots<false>(sm, ksel, (uint32_t)(w1 * 32 + lane), nls * K, inv, negK16, rec_s, buf_s, hb_seq, ptab_s, over);
        junk_q += part_slots<true>(sm, ksel, (uint32_t)(w2 * 32 + lane), 2u * nlq, part_s + PART_CAP * 4u, masks_s, buf_s, hb_seq + HB * 32 * 4, ptab_s, over);
        junk_s += part_slots<false>(sm, ksel, (uint32_t)(w3 * 32 + lane), 2u * nls, part_s, masks_s, buf_s, hb_seq, ptab_s, over);
        const uint32_t nlong = m.nlong;  // long lines: every worker warp takes every WORK_WARPS-th run of 32 groups
        for (uint32_t l = 0; l < nlong; l++) {
          const uint32_t r = sm.longl[sc][l];
          const uint32_t n = (r >> 10) & 2047u, g0 = r & 1023u, q0 = (r >> 21) & 1023u;
          if (r >> 31) {
            for (uint32_t t = (uint32_t)(ww * 32 + lane); t < n; t += WORK_WARPS * 32)
              group_full<true>(sm, ksel, buf_s + 16u * (g0 + t), hb_seq + HB * 32 * 4, ptab_s, q0 + 16u * t, over);
          } else {
            for (uint32_t t = (uint32_t)(ww * 32 + lane); t < n; t += WORK_WARPS * 32)
              group_full<false>(sm, ksel, buf_s + 16u * (g0 + t), hb_seq, ptab_s, 0u, over);
          }
        }
      }
      // ---- K1a of tile A = B+1 (its TMA was started at the top of this iteration) -> the other bitmap slot ----
      if (it + 1 < nt) {
        mbar_wait(bar0_s + 8u * stA, (par_bits >> stA) & 1u);
        par_bits ^= 1u << stA;
        const uint32_t hib = k1a_tile<WORK_THREADS>(sm, buf0_s + (uint32_t)stA * STAGE_BYTES, sm0 + (uint32_t)offsetof(Smem, bitmap) + (uint32_t)sc * (BM_WORDS * 4u),
                                                    it + 1 == it_first ? (int)a.lo0 : 0, it + 1 == it_last ? hi_last : TILE, tid - LINE_THREADS);
        if (hib & 0x80808080u) sm.hiflag[sc] = 1;
      }
    }

    TR(7);
    __syncthreads();
    TR(8);
    { const int t = stC; stC = stB; stB = stA; stA = t; }  // tile C is consumed, B becomes C
    const uint32_t post = sm.post[sb];  // (slot of this iteration: thread 0 rewrites it two iterations later)
    if (post) {
      if (post & 1u) {  // dense or high-byte tile B: the whole CTA walks it now (its bitmap is still in place)
        tile_walker(sm, a, &sm.buf[stC][PAD], sm.meta[sb], phase, sm.bitmap[sb], reinterpret_cast<const uint16_t*>(&sm.part[sb][0][0]),
                    tid, THREADS);
      }
      if (post & 2u) flush_pos_tab_cold(sm, block, tid);
      if (post & 4u) { __syncthreads(); flush_gpos_cold(sm, block, tid); }
      __syncthreads();
    }
  }

#ifdef FQ_TRACE
  if (blockIdx.x == 7 && (tid == 0 || tid == 33 || tid == 127 || tid == 200) && pass == 0) printf("tid %d nt %d: top->mbar %llu mbar %llu k1a %llu bar1 %llu scanA %llu nlx %llu bar3 %llu lines/work %llu sync %llu\n", tid, nt, tr[0]/nt, tr[1]/nt, tr[2]/nt, tr[3]/nt, tr[4]/nt, tr[5]/nt, tr[6]/nt, tr[7]/nt, tr[8]/nt);
#endif
  // ---- flush this span's counters into its block; pass 0 also records the span descriptor ----
  if (mns != ~0u) { atomicMin(&sm.len_min[0], (u64)mns); atomicMax(&sm.len_max[0], (u64)mxs); }
  if (mnq != ~0u) { atomicMin(&sm.len_min[1], (u64)mnq); atomicMax(&sm.len_max[1], (u64)mxq); }
  if (junk_s) atomicAdd(&sm.junk[0], (u64)junk_s);
  if (junk_q) atomicAdd(&sm.junk[1], (u64)junk_q);
  if (over) atomicAdd(&sm.pos_over, over);
  __syncthreads();
  if (pass == 0 && tid == 0) {
    desc.T = sm.run_L;
    desc.head_len = sm.head_len;
    desc.tail_len = sm.run_open;
  }
  if (pass == 1 && tid == 0) desc.state = SPAN_COMMITTED;
  if (count_only) return;
  for (int bin = tid; bin < 512; bin += THREADS) {  // fold the 32 lane copies (rotated: no bank conflicts)
    const int h = bin >> 8, b = bin & 255;
    u64 s = sm.ghist[h][b];
    if (b < HB) {
      const uint32_t* hp = &sm.hist[h][b << 5];
#pragma unroll 8
      for (int l = 0; l < 32; l++) s += hp[(l + bin) & 31];
    }
    if (b == 0) s -= sm.junk[h];  // masked bytes of partial groups were counted as byte 0
    if (s) block[OFF_HIST_SEQ + bin] += s;
  }
  for (int i = tid; i <= POS_BINS; i += THREADS) {
    if (sm.seq_len[i]) block[OFF_SEQ_LEN + i] += sm.seq_len[i];
    if (sm.qual_len[i]) block[OFF_QUAL_LEN + i] += sm.qual_len[i];
  }
  if (tid < LOG2_BINS && sm.seq_log2[tid]) block[OFF_SEQ_LOG2 + tid] += sm.seq_log2[tid];
  flush_pos_tab(sm, block, tid);
  __syncthreads();
  flush_gpos(sm, block, tid);
  if (tid == 0) {
    if (sm.pos_over) block[OFF_POS_SUM + POS_BINS] += sm.pos_over;
    if (sm.len_min[0] < block[OFF_SEQ_LEN_MIN]) block[OFF_SEQ_LEN_MIN] = sm.len_min[0];
    if (sm.len_max[0] > block[OFF_SEQ_LEN_MAX]) block[OFF_SEQ_LEN_MAX] = sm.len_max[0];
    if (sm.len_min[1] < block[OFF_QUAL_LEN_MIN]) block[OFF_QUAL_LEN_MIN] = sm.len_min[1];
    if (sm.len_max[1] > block[OFF_QUAL_LEN_MAX]) block[OFF_QUAL_LEN_MAX] = sm.len_max[1];
  }
}

#include "fq_scan_fast.cuh"

// ---------------------------------------------------------------------------------------------
// resync: guess the line phase at every span start.  Span-relative line j (j >= 1) starts after the
// span's j-th newline; the first j whose line starts with '@' while line j+2 starts with '+' is a
// header, so (lines before the span) = -j (mod 4).  For valid 4-line FASTQ this is unambiguous (a
// quality line starting with '@' is followed two lines later by a sequence line).  One warp per span.
// Block 0 also snapshots the stream carry into the launch header.
// ---------------------------------------------------------------------------------------------
constexpr int RESYNC_LINES = 40;
constexpr u64 RESYNC_BYTES = 4ull << 20;

__global__ void fq_resync_kernel(const ScanArgs a) {
  const int span = blockIdx.x, lane = threadIdx.x;
  const unsigned cflags = a.carry->flags;
  if (span == 0 && lane == 0) {
    LaunchHdr h;
    h.lines0 = a.carry->lines; h.open0 = a.carry->open_len; h.bytes0 = a.carry->bytes;
    h.last_byte0 = a.carry->last_byte; h.flags0 = cflags; h.mismatches = 0; h.pad = 0;
    *a.hdr = h;
    if ((cflags & CARRY_UNKNOWN_START) && a.carry->bytes == 0) a.shard->first_byte = a.base[a.lo0];
  }
  SpanDesc& d = a.desc[span];
  if (lane == 0) { d.T = 0; d.head_len = 0; d.tail_len = 0; d.state = SPAN_PENDING; d.exact = PHASE_UNKNOWN; d.G = 0; d.P0 = 0; d.pad = 0; }
  // The phase of the launch's first byte: exact from the stream carry, or (multi-GPU shard with an
  // unknown start) the shard's hypothesis plus the lines counted so far, or resynced like any span.
  const bool start_known = !(cflags & CARRY_UNKNOWN_START) || (cflags & CARRY_HYP_VALID);
  if (span == 0 && start_known) {
    const unsigned hyp = (cflags & CARRY_UNKNOWN_START) ? (cflags >> CARRY_HYP_SHIFT) & 3u : 0u;
    if (lane == 0) d.guess = (uint32_t)((hyp + a.carry->lines) & 3);
    return;
  }
  const u64 begin = (u64)span * a.tps * TILE;  // 16-byte aligned
  const u64 stop = begin + RESYNC_BYTES < a.end ? begin + RESYNC_BYTES : a.end;
  uint32_t first[RESYNC_LINES + 1];  // first byte of span-relative line j (0x100 = beyond the launch)
  int nlines = 0;                    // lines whose start has been seen: 1..nlines
  uint32_t guess = PHASE_UNKNOWN;
  for (u64 o = begin; o < stop && guess == PHASE_UNKNOWN && nlines < RESYNC_LINES; o += 512) {
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
      if (guess != PHASE_UNKNOWN) break;
    }
  }
  if (lane == 0) d.guess = guess;
}

// ---------------------------------------------------------------------------------------------
// stitch: exact prefix over the span descriptors, verification of the guessed phases, commit of the
// span blocks, the head fragment of every span, and the new stream carry.  One CTA per span.
// ---------------------------------------------------------------------------------------------
constexpr int STITCH_THREADS = 256;

__global__ void __launch_bounds__(STITCH_THREADS) fq_stitch_kernel(const ScanArgs a) {
  __shared__ uint32_t ghist[2][256];
  __shared__ uint32_t pos_sum[POS_BINS + 1];
  __shared__ u64 sP0;
  __shared__ int s_commit, s_detached;
  __shared__ uint32_t s_phase;
  const int span = blockIdx.x, tid = threadIdx.x;
  SpanDesc& d = a.desc[span];
  const LaunchHdr& h = *a.hdr;
  auto span_begin = [&](int c) -> u64 { return c == 0 ? (u64)a.lo0 : (u64)c * a.tps * TILE; };
  auto span_end = [&](int c) -> u64 { const u64 e = (u64)(c + 1) * a.tps * TILE; return e < a.end ? e : a.end; };
  // exact prefix over the earlier spans: G = lines before this span, P0 = open-line bytes before it
  // (tail of the last earlier span that has a newline, plus the lengths of the newline-free spans after it)
  __shared__ u64 s_sumT[STITCH_THREADS / 32], s_sumL[STITCH_THREADS / 32];
  __shared__ int s_last[STITCH_THREADS / 32];
  {
    u64 myT = 0;
    int mylast = -1;
    for (int c = tid; c < span; c += STITCH_THREADS) { const u64 T = a.desc[c].T; myT += T; if (T) mylast = c; }
    for (int d = 16; d > 0; d >>= 1) { myT += __shfl_xor_sync(0xffffffffu, myT, d); mylast = max(mylast, __shfl_xor_sync(0xffffffffu, mylast, d)); }
    if ((tid & 31) == 0) { s_sumT[tid >> 5] = myT; s_last[tid >> 5] = mylast; }
    __syncthreads();
    int last = -1;
    for (int w = 0; w < STITCH_THREADS / 32; w++) last = max(last, s_last[w]);
    u64 myL = 0;
    for (int c = last + 1 + tid; c < span; c += STITCH_THREADS) myL += span_end(c) - span_begin(c);
    for (int d = 16; d > 0; d >>= 1) myL += __shfl_xor_sync(0xffffffffu, myL, d);
    if ((tid & 31) == 0) s_sumL[tid >> 5] = myL;
    __syncthreads();
  }
  if (tid == 0) {
    u64 G = h.lines0, P0 = 0;
    int last = -1;
    for (int w = 0; w < STITCH_THREADS / 32; w++) { G += s_sumT[w]; P0 += s_sumL[w]; last = max(last, s_last[w]); }
    P0 += last >= 0 ? a.desc[last].tail_len : h.open0;
    // Multi-GPU shard with an unknown start: G counts from the shard start and the phase is the shard's
    // hypothesis -- carried over from an earlier launch, or fixed now by the first span whose resync
    // found a header (fqgpu_shard_combine verifies it against the exact counts of the other ranks).
    const bool unknown = (h.flags0 & CARRY_UNKNOWN_START) != 0;
    unsigned hyp = 0;
    bool hyp_ok = true;
    if (unknown) {
      if (h.flags0 & CARRY_HYP_VALID) hyp = (h.flags0 >> CARRY_HYP_SHIFT) & 3u;
      else {
        hyp_ok = false;
        u64 Gc = h.lines0;
        for (int c = 0; c < (int)a.nspans; c++) {
          if (a.desc[c].guess < 4) { hyp = (unsigned)((a.desc[c].guess - Gc) & 3); hyp_ok = true; break; }
          Gc += a.desc[c].T;
        }
      }
    }
    const uint32_t exact = hyp_ok ? (uint32_t)((hyp + G) & 3) : PHASE_UNKNOWN;
    sP0 = P0; s_phase = exact;
    s_detached = unknown && G == 0;  // still inside the shard's first line fragment
    d.G = G; d.P0 = P0; d.exact = exact;
    const int ok = hyp_ok && d.guess == exact && !d.pad;  // pad: the fast pass abandoned the span
    d.state = ok ? SPAN_COMMITTED : SPAN_RESCAN;
    if (!ok) atomicAdd(&a.hdr->mismatches, 1u);
    s_commit = ok;
    if (span == (int)a.nspans - 1) {  // the new stream carry
      const u64 len = span_end(span) - span_begin(span);
      a.carry->lines = G + d.T;
      a.carry->open_len = d.T ? d.tail_len : P0 + len;
      a.carry->bytes = h.bytes0 + (a.end - a.lo0);
      a.carry->last_byte = a.base[a.end - 1];
      if (unknown && hyp_ok) a.carry->flags = CARRY_UNKNOWN_START | CARRY_HYP_VALID | (hyp << CARRY_HYP_SHIFT);
    }
  }
  for (int i = tid; i < 512; i += STITCH_THREADS) (&ghist[0][0])[i] = 0;
  for (int i = tid; i <= POS_BINS; i += STITCH_THREADS) pos_sum[i] = 0;
  __syncthreads();
  u64* cblock = a.committed + (size_t)span * BLOCK_WORDS;
  if (s_commit) {
    const u64* pblock = a.pending + (size_t)span * BLOCK_WORDS;
    for (int w = tid; w < BLOCK_WORDS; w += STITCH_THREADS) {
      const u64 x = pblock[w];
      if (w == OFF_SEQ_LEN_MIN || w == OFF_QUAL_LEN_MIN) { if (x < cblock[w]) cblock[w] = x; }
      else if (w == OFF_SEQ_LEN_MAX || w == OFF_QUAL_LEN_MAX) { if (x > cblock[w]) cblock[w] = x; }
      else if (x) cblock[w] += x;
    }
  }
  __syncthreads();  // the commit is complete before the head fragment is added to the same block
  // ---- head fragment: bytes [b0, b0 + head_len) belong to line G, at line position P0 ----
  const u64 P0 = sP0;
  const bool detached = s_detached != 0;  // the fragment belongs to the shard's first line: P0 is relative
  if (s_phase > 3) return;
  const int cls = (int)s_phase;
  const u64 b0 = span_begin(span);
  const u64 hl = d.head_len;
  u64 ve = b0 + hl;  // content end
  int cr = 0;
  if (d.T) {  // terminated inside the span by the newline at b0 + hl
    if (hl > 0) { if (a.base[ve - 1] == '\r') { ve--; cr = 1; } }
    else if (P0 > 0) {  // the '\r' (if any) is the last byte before the span; it was already dropped there
      const int prev = b0 > (u64)a.lo0 ? (int)a.base[b0 - 1] : (h.bytes0 ? (int)h.last_byte0 : 0);
      if (prev == '\r') cr = 1;
    }
  } else if (hl > 0 && a.base[ve - 1] == '\r') {  // the fragment runs to the span end: look one byte ahead
    const int nx = ve < a.end ? (int)a.base[ve] : -1;
    if (nx == '\n' || nx < 0) ve--;
  }
  if (detached && tid == 0 && d.T) {  // the shard's first line ends here; its length is stitched by the combine step
    a.shard->head_len = P0 + hl;
    a.shard->head_cr = (unsigned)cr;
  }
  if (!(cls & 1) || (a.core && cls == 3)) return;
  for (u64 o = b0 + tid; o < ve; o += STITCH_THREADS) account_byte(ghist, pos_sum, cls, a.base[o], P0 + (o - b0));
  if (tid == 0) {
    // the '\r' that ended the previous launch is content unless this launch starts with '\n'
    if (span == 0 && h.bytes0 && h.open0 && h.last_byte0 == '\r' && a.base[a.lo0] != '\n')
      account_byte(ghist, pos_sum, cls, '\r', P0 - 1);
    if (d.T && !detached) {
      const u64 len = P0 + hl - (u64)cr;
      const uint32_t bin = len < (u64)POS_BINS ? (uint32_t)len : (uint32_t)POS_BINS;
      if (cls == 3) {
        cblock[OFF_QUAL_LEN + bin] += 1;
        if (len < cblock[OFF_QUAL_LEN_MIN]) cblock[OFF_QUAL_LEN_MIN] = len;
        if (len > cblock[OFF_QUAL_LEN_MAX]) cblock[OFF_QUAL_LEN_MAX] = len;
      } else {
        cblock[OFF_SEQ_LEN + bin] += 1;
        cblock[OFF_SEQ_LOG2 + log2_bin(len)] += 1;
        if (len < cblock[OFF_SEQ_LEN_MIN]) cblock[OFF_SEQ_LEN_MIN] = len;
        if (len > cblock[OFF_SEQ_LEN_MAX]) cblock[OFF_SEQ_LEN_MAX] = len;
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < 512; i += STITCH_THREADS) { const uint32_t v = (&ghist[0][0])[i]; if (v) cblock[OFF_HIST_SEQ + i] += v; }
  // per-position sums of a detached head are relative to the shard start: kept apart, shifted by the combine step
  u64* ptarget = detached ? a.shard->head_pos : cblock + OFF_POS_SUM;
  for (int i = tid; i <= POS_BINS; i += STITCH_THREADS) { const uint32_t v = pos_sum[i]; if (v) atomicAdd(&ptarget[i], (u64)v); }
}

// Resets the committed span blocks and the stream carry (a new file).
__global__ void fq_reset_kernel(u64* committed, int nblocks, Carry* carry) {
  const size_t n = (size_t)nblocks * BLOCK_WORDS;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int w = (int)(i % BLOCK_WORDS);
    committed[i] = (w == OFF_SEQ_LEN_MIN || w == OFF_QUAL_LEN_MIN) ? ~0ull : 0ull;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    Carry c;
    c.lines = 0; c.open_len = 0; c.bytes = 0; c.last_byte = 0; c.flags = 0;
    c.meta_lines = 0; c.qual_min = -1; c.qual_max = -1; c.meta_status = 0; c.meta_pending_cr = 0;
    c.cur_has = 0; c.cur_min = 0; c.cur_max = 0; c.pad = 0;
    *carry = c;
  }
}

// K3: fold the per-span counter blocks into one block (sum words, then the four min/max words).
__global__ void __launch_bounds__(256) fq_reduce_kernel(const u64* __restrict__ blocks, int nblocks, u64* __restrict__ out) {
  __shared__ u64 part[8][32];
  const int wl = threadIdx.x & 31, p = threadIdx.x >> 5;  // word within the CTA's 32 words, part 0..7
  const int w = blockIdx.x * 32 + wl;
  const bool is_min = (w == OFF_SEQ_LEN_MIN || w == OFF_QUAL_LEN_MIN);
  const bool is_max = (w == OFF_SEQ_LEN_MAX || w == OFF_QUAL_LEN_MAX);
  u64 acc = is_min ? ~0ull : 0ull;
  if (w < BLOCK_WORDS) {
    for (int b = p; b < nblocks; b += 8) {
      const u64 x = blocks[(size_t)b * BLOCK_WORDS + w];
      if (is_min) acc = x < acc ? x : acc;
      else if (is_max) acc = x > acc ? x : acc;
      else acc += x;
    }
  }
  part[p][wl] = acc;
  __syncthreads();
  if (p == 0 && w < BLOCK_WORDS) {
    for (int q = 1; q < 8; q++) {
      const u64 x = part[q][wl];
      if (is_min) acc = x < acc ? x : acc;
      else if (is_max) acc = x > acc ? x : acc;
      else acc += x;
    }
    out[w] = acc;
  }
}

// fq-meta quality-range fold over the first 4*meta_records lines (src/fq_meta.nim:226-248): one
// warp walks the stream prefix 32 bytes at a time; qual_to_int (src/fq_meta.nim:94-95) per lane,
// per-line min/max by warp reductions, the prev_min >= 0 rule (src/fq_meta.nim:100-102) per line.
__device__ __forceinline__ void meta_fold(long long& qmin, long long& qmax, unsigned& status, int has, int mn, int mx) {
  if (has) {
    long long a = mn, b = mx;
    if (qmin >= 0) { a = a < qmin ? a : qmin; b = b > qmax ? b : qmax; }
    qmin = a; qmax = b;
  } else if (qmin < 0) {
    status = FQGPU_META_EMPTY_QUAL;
  }
}

// One warp, 512 bytes per iteration (16 per lane): newline masks by SWAR, the segments between
// newlines are reduced with warp min/max; only quality lines (0-based index % 4 == 3) are examined.
__global__ void fq_meta_kernel(const uint8_t* __restrict__ base, uint32_t lo0, u64 end, Carry* __restrict__ carry, u64 meta_records) {
  const int lane = threadIdx.x;
  const u64 limit = meta_records * 4;
  u64 ml = carry->meta_lines;
  if (ml >= limit || end <= (u64)lo0) return;
  long long qmin = carry->qual_min, qmax = carry->qual_max;
  unsigned status = carry->meta_status;
  int cur_has = carry->cur_has, cur_min = carry->cur_min, cur_max = carry->cur_max;
  if (carry->meta_pending_cr && base[lo0] != '\n') {  // the '\r' that ended the previous chunk was content
    if (!cur_has) { cur_has = 1; cur_min = -1; cur_max = -1; } else { cur_min = -1; }
  }
  unsigned pending = 0;
  bool done = false;
  uint4 vnext = make_uint4(0, 0, 0, 0);
  if ((u64)lane * 16 < end) vnext = *reinterpret_cast<const uint4*>(base + (u64)lane * 16);
  for (u64 o = 0; o < end && !done; o += 512) {
    const u64 g = o + (u64)lane * 16;
    const uint4 v = vnext;
    if (g + 512 < end) vnext = *reinterpret_cast<const uint4*>(base + g + 512);  // prefetch the next window
    // valid byte range of this lane: [va, vb) within its 16 bytes
    int va = g >= (u64)lo0 ? 0 : (int)min((u64)16, (u64)lo0 - g);
    int vb = g + 16 <= end ? 16 : (g < end ? (int)(end - g) : 0);
    uint32_t nlm = nl_mask16(v) & ((1u << vb) - 1u) & ~((1u << va) - 1u);
    // byte following this lane's 16 (for the '\r' rule): next lane's first byte, or memory
    uint32_t nxt = __shfl_down_sync(0xffffffffu, v.x & 0xFFu, 1);
    if (lane == 31) nxt = (g + 16 < end) ? (uint32_t)base[g + 16] : 0u;
    uint32_t lanes_nl = __ballot_sync(0xffffffffu, nlm != 0);
    int seg_lane = 0, seg_k = 0;  // current segment starts at (lane, byte) = (seg_lane, seg_k)
    for (;;) {
      // next newline at or after the segment start
      int nl_lane = 32, nl_k = 0;
      if (lanes_nl) {
        nl_lane = __ffs(lanes_nl) - 1;
        const uint32_t m = __shfl_sync(0xffffffffu, nlm, nl_lane);
        nl_k = __ffs(m) - 1;
      }
      if ((ml & 3) == 3) {  // quality line: min/max of qual_to_int over [segment start, newline or window end)
        int a = lane < seg_lane ? 16 : (lane == seg_lane ? seg_k : 0);
        int b = lane > nl_lane ? 0 : (lane == nl_lane ? nl_k : 16);
        a = max(a, va); b = min(b, vb);
        int mn = 0x7fffffff, mx = -0x7fffffff;
#pragma unroll
        for (int k = 0; k < 16; k++) {
          const uint32_t wk = k < 4 ? v.x : (k < 8 ? v.y : (k < 12 ? v.z : v.w));
          const uint32_t c = (wk >> (8 * (k & 3))) & 0xFFu;
          bool use = k >= a && k < b;
          if (c == '\r' && use) {
            if (g + (u64)k + 1 >= end) { pending = 1; use = false; }  // last byte of the chunk: decided later
            else {
              const uint32_t wn = (k + 1) < 4 ? v.x : ((k + 1) < 8 ? v.y : ((k + 1) < 12 ? v.z : v.w));
              const uint32_t nx = k < 15 ? ((wn >> (8 * ((k + 1) & 3))) & 0xFFu) : nxt;
              if (nx == '\n') use = false;                            // dropped: directly before the newline
            }
          }
          if (use) {
            const int q = (c >= 33 && c <= 126) ? (int)c - 33 : -1;
            mn = min(mn, q); mx = max(mx, q);
          }
        }
        mn = __reduce_min_sync(0xffffffffu, mn);
        mx = __reduce_max_sync(0xffffffffu, mx);
        pending = __reduce_max_sync(0xffffffffu, pending);
        if (mn != 0x7fffffff) {
          if (!cur_has) { cur_has = 1; cur_min = mn; cur_max = mx; }
          else { cur_min = min(cur_min, mn); cur_max = max(cur_max, mx); }
        }
      }
      if (nl_lane == 32) break;
      if ((ml & 3) == 3 && status == FQGPU_META_OK) meta_fold(qmin, qmax, status, cur_has, cur_min, cur_max);
      ml++;
      cur_has = 0;
      pending = 0;
      if (ml >= limit) { done = true; break; }
      // consume this newline
      if (lane == nl_lane) nlm &= nlm - 1;
      lanes_nl = __ballot_sync(0xffffffffu, nlm != 0);
      seg_lane = nl_lane; seg_k = nl_k + 1;
      if (seg_k == 16) { seg_lane++; seg_k = 0; }
    }
  }
  if (lane == 0) {
    carry->meta_lines = ml;
    carry->qual_min = qmin; carry->qual_max = qmax;
    carry->meta_status = status;
    carry->meta_pending_cr = done ? 0u : pending;
    carry->cur_has = cur_has; carry->cur_min = cur_min; carry->cur_max = cur_max;
  }
}

// The same fold by one CTA of 1024 threads, 16 KiB per step (long-read prefixes are megabytes): every thread
// takes 16 bytes, a block-wide prefix sum of the newline counts gives the line index of every byte, the
// per-line (min, max) are shared-memory atomics on keys (0 = byte outside the table, b - 32 inside, so that
// key - 1 = qual_to_int), and thread 0 folds the lines in order at the end.  Holds up to META_CAP lines.
constexpr int META_THREADS = 1024;
constexpr int META_CAP = 4096;
constexpr uint32_t META_NONE = 0xFFFFFFFFu;

__global__ void __launch_bounds__(META_THREADS) fq_meta_par_kernel(const uint8_t* __restrict__ base, uint32_t lo0, u64 end,
                                                                 Carry* __restrict__ carry, u64 meta_records) {
  __shared__ uint32_t kmin[META_CAP + 1], kmax[META_CAP + 1];
  __shared__ uint32_t warp_cnt[META_THREADS / 32];
  __shared__ uint32_t s_pending;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const u64 limit = meta_records * 4;
  const u64 ml0 = carry->meta_lines;
  if (ml0 >= limit || end <= (u64)lo0) return;
  for (int i = tid; i <= META_CAP; i += META_THREADS) { kmin[i] = META_NONE; kmax[i] = 0; }
  if (tid == 0) s_pending = 0;
  __syncthreads();
  if (tid == 0) {  // the open line carried in from the previous chunk
    if (carry->cur_has) { kmin[0] = (uint32_t)(carry->cur_min + 1); kmax[0] = (uint32_t)(carry->cur_max + 1); }
    if (carry->meta_pending_cr && base[lo0] != '\n') kmin[0] = 0;  // the '\r' that ended the previous chunk was content
  }
  __syncthreads();
  u64 L0 = ml0;  // lines before the current window
  for (u64 o = 0; o < end && L0 < limit; o += (u64)META_THREADS * 16) {
    const u64 g = o + (u64)tid * 16;
    uint4 v = make_uint4(0, 0, 0, 0);
    int va = 16, vb = 0;  // valid bytes of this thread: [va, vb)
    if (g < end && g + 16 > (u64)lo0) {
      v = *reinterpret_cast<const uint4*>(base + g);
      va = g >= (u64)lo0 ? 0 : (int)((u64)lo0 - g);
      vb = g + 16 <= end ? 16 : (int)(end - g);
    }
    const uint32_t m = va < vb ? (nl_mask16(v) & ((1u << vb) - 1u) & ~((1u << va) - 1u)) : 0u;
    const uint32_t cnt = __popc(m);
    const uint32_t inc = warp_incl_scan(cnt, lane);
    if (lane == 31) warp_cnt[warp] = inc;
    // the byte after this thread's 16 (the '\r' rule): the next lane's first byte, or memory; 0x100 = end of the chunk
    uint32_t nxt = __shfl_down_sync(0xffffffffu, v.x & 0xFFu, 1);
    if (lane == 31) nxt = (g + 16 < end) ? (uint32_t)base[g + 16] : 0x100u;
    else if (g + 16 >= end) nxt = 0x100u;
    __syncthreads();
    uint32_t wbase = 0, total = 0;
    for (int w = 0; w < META_THREADS / 32; w++) { const uint32_t x = warp_cnt[w]; if (w < warp) wbase += x; total += x; }
    u64 L = L0 + wbase + inc - cnt;  // line index of this thread's first byte
    uint32_t mn = META_NONE, mx = 0;
    const uint32_t ww[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 16; k++) {
      if (k >= va && k < vb) {
        const uint32_t c = (ww[k >> 2] >> (8 * (k & 3))) & 0xFFu;
        if ((m >> k) & 1u) {
          if (mn != META_NONE && L < limit) { atomicMin(&kmin[L - ml0], mn); atomicMax(&kmax[L - ml0], mx); }
          mn = META_NONE; mx = 0;
          L++;
        } else if ((L & 3) == 3 && L < limit) {
          bool content = true;
          if (c == '\r') {
            const uint32_t nx = (k + 1 < vb) ? ((ww[(k + 1) >> 2] >> (8 * ((k + 1) & 3))) & 0xFFu) : (k + 1 < 16 ? 0x100u : nxt);
            if (nx == 0x100u) { s_pending = 1; content = false; }  // last byte of the chunk: decided by the next one
            else if (nx == '\n') content = false;                   // dropped: directly before the newline
          }
          if (content) {
            const uint32_t key = (c >= 33u && c <= 126u) ? c - 32u : 0u;
            mn = min(mn, key); mx = max(mx, key);
          }
        }
      }
    }
    if (mn != META_NONE && L < limit) { atomicMin(&kmin[L - ml0], mn); atomicMax(&kmax[L - ml0], mx); }
    L0 += total;
    __syncthreads();  // warp_cnt is rewritten by the next window
  }
  if (tid == 0) {
    const u64 ml_end = L0 < limit ? L0 : limit;
    long long qmin = carry->qual_min, qmax = carry->qual_max;
    unsigned status = carry->meta_status;
    for (u64 L = ml0; L < ml_end; L++) {
      if ((L & 3) != 3 || status != FQGPU_META_OK) continue;
      const uint32_t a = kmin[L - ml0];
      meta_fold(qmin, qmax, status, a != META_NONE, (int)a - 1, (int)kmax[L - ml0] - 1);
    }
    const bool done = ml_end >= limit;
    carry->meta_lines = ml_end;
    carry->qual_min = qmin; carry->qual_max = qmax;
    carry->meta_status = status;
    carry->meta_pending_cr = done ? 0u : s_pending;
    const uint32_t a = done ? META_NONE : kmin[ml_end - ml0];
    carry->cur_has = a != META_NONE;
    carry->cur_min = a != META_NONE ? (int)a - 1 : 0;
    carry->cur_max = a != META_NONE ? (int)kmax[ml_end - ml0] - 1 : 0;
  }
}

// ------------------------------------------------------------------------------------------
// host-callable launchers (used by fqgpu_api.cu)
// ------------------------------------------------------------------------------------------
size_t scan_smem_bytes() { return sizeof(Smem); }
int scan_tile_bytes() { return TILE; }
int scan_threads() { return THREADS; }

cudaError_t scan_configure() {
  cudaError_t e = cudaFuncSetAttribute(fq_scan_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FastSmem));
  if (e != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(fq_scan_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem))) != cudaSuccess) return e;
  return cudaFuncSetAttribute(fq_scan_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
}

cudaError_t launch_reset(u64* committed, int nblocks, Carry* carry, cudaStream_t st) {
  fq_reset_kernel<<<nblocks > 296 ? 592 : 64, 256, 0, st>>>(committed, nblocks, carry);
  return cudaGetLastError();
}

// Spans of a launch over `bytes` bytes (incl. the alignment slack in front) with `resident` CTAs resident at once.
// Large launches get several spans per resident CTA, handed out by the hardware as CTAs finish: of the two CTAs of an
// SM the one that started first runs ~1.2x faster, so with one span each the slower half
// finishes late and alone; with shorter spans the tail is one short span (+3 % at 36 GB).  Spans stay >= 16 MiB: below
// that the longer stitch / reduce over more span blocks costs more than the tail (measured on 4.5 GB shards).
uint32_t scan_span_count(u64 bytes, int resident) {
  static const uint32_t waves = getenv("FQGPU_SPAN_WAVES") ? (uint32_t)atoi(getenv("FQGPU_SPAN_WAVES")) : (uint32_t)SPAN_WAVES;
  const u64 ntiles = (bytes + TILE - 1) / TILE;
  if (ntiles < (u64)resident) return (uint32_t)ntiles;
  static const u64 min_tiles = getenv("FQGPU_SPAN_MIN_TILES") && atoi(getenv("FQGPU_SPAN_MIN_TILES")) > 0 ? (u64)atoi(getenv("FQGPU_SPAN_MIN_TILES")) : 1024u;  // (tests force waves on small inputs)
  u64 f = ntiles / ((u64)resident * min_tiles);
  f = f < 1 ? 1 : (f > waves ? waves : f);
  if (f > (u64)SPAN_WAVES) f = SPAN_WAVES;
  return (uint32_t)((u64)resident * f);
}

// Scans `nbytes` at `ptr` (any alignment) as the continuation of the stream described by `carry`:
// resync -> scan (guessed phases) -> stitch (verify, commit, head fragments, carry) -> scan pass 1
// (only spans whose guess was wrong or unknown; exits immediately otherwise).
cudaError_t launch_scan(const void* ptr, size_t nbytes, SpanDesc* desc, LaunchHdr* hdr, Carry* carry,
                        u64* pending, u64* committed, ShardInfo* shard, int max_spans, u64 meta_records, cudaStream_t st,
                        cudaStream_t meta_stream, cudaEvent_t ev_fork, cudaEvent_t ev_join, bool core_only) {
  if (nbytes == 0) return cudaSuccess;
  static const uint32_t dbg = getenv("FQGPU_DEBUG") ? (uint32_t)atoi(getenv("FQGPU_DEBUG")) : 0u;
  const uintptr_t addr = (uintptr_t)ptr;
  ScanArgs a;
  a.lo0 = (uint32_t)(addr & 15);
  a.base = (const uint8_t*)(addr - a.lo0);
  a.end = (u64)a.lo0 + nbytes;
  a.ntiles = (uint32_t)((a.end + TILE - 1) / TILE);
  const uint32_t nspans = scan_span_count((u64)a.lo0 + nbytes, max_spans);
  a.tps = (a.ntiles + nspans - 1) / nspans;
  a.nspans = (a.ntiles + a.tps - 1) / a.tps;
  a.desc = desc; a.hdr = hdr; a.carry = carry; a.pending = pending; a.committed = committed; a.shard = shard; a.dbg = dbg; a.core = core_only ? 1u : 0u;

  // The fq-meta prefix fold touches only its own fields of the carry: it runs beside the scan on its own stream.
  cudaError_t e;
  if (meta_records) {
    if ((e = cudaEventRecord(ev_fork, st)) != cudaSuccess) return e;
    if ((e = cudaStreamWaitEvent(meta_stream, ev_fork, 0)) != cudaSuccess) return e;
    if (meta_records * 4 <= (u64)META_CAP) fq_meta_par_kernel<<<1, META_THREADS, 0, meta_stream>>>(a.base, a.lo0, a.end, carry, meta_records);
    else fq_meta_kernel<<<1, 32, 0, meta_stream>>>(a.base, a.lo0, a.end, carry, meta_records);
    if ((e = cudaEventRecord(ev_join, meta_stream)) != cudaSuccess) return e;
  }
  fq_resync_kernel<<<a.nspans, 32, 0, st>>>(a);
  // pass 0: fq_scan_kernel.  FQGPU_SCAN=fast selects the experimental register-resident fq_scan_fast_kernel
  // instead (well-formed input only; it abandons a span otherwise and pass 1 redoes it).  Measured on B200 it is
  // slower than the tile kernel in every mode, so it is not the default.
  const char* scan_env = getenv("FQGPU_SCAN");
  if (scan_env && !strcmp(scan_env, "fast")) fq_scan_fast_kernel<<<a.nspans, F_THREADS, sizeof(FastSmem), st>>>(a);
  else if (core_only) fq_scan_kernel<true><<<a.nspans, THREADS, sizeof(Smem), st>>>(a, 0);
  else fq_scan_kernel<false><<<a.nspans, THREADS, sizeof(Smem), st>>>(a, 0);
  fq_stitch_kernel<<<a.nspans, STITCH_THREADS, 0, st>>>(a);
  if (core_only) fq_scan_kernel<true><<<a.nspans, THREADS, sizeof(Smem), st>>>(a, 1);
  else fq_scan_kernel<false><<<a.nspans, THREADS, sizeof(Smem), st>>>(a, 1);
  if (meta_records && (e = cudaStreamWaitEvent(st, ev_join, 0)) != cudaSuccess) return e;
  if (dbg & 16u) {  // diagnostics: spans the stitch kernel sent to the exact second pass
    LaunchHdr h;
    cudaStreamSynchronize(st);
    cudaMemcpy(&h, hdr, sizeof(h), cudaMemcpyDeviceToHost);
    fprintf(stderr, "fqgpu: launch of %zu bytes, %u spans, %u rescanned\n", nbytes, a.nspans, h.mismatches);
  }
  return cudaGetLastError();
}

cudaError_t launch_reduce(const u64* blocks, int nblocks, u64* out, cudaStream_t st) {
  fq_reduce_kernel<<<(BLOCK_WORDS + 31) / 32, 256, 0, st>>>(blocks, nblocks, out);
  return cudaGetLastError();
}

}  // namespace fq
