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
//   K1      boundary classification: every warp turns 16-byte groups into '\n' masks (SWAR compare,
//           IDP.4A movemask) -> tile bitmap; four SCANNER warps popc / prefix-sum the bitmap into the
//           tile's newline index, keep the span's running line count / open-line length (the
//           tile-edge record carry) and the line-length tables;  concurrently
//   K2      twelve WORKER warps walk the lines of the previous tile (a quarter-warp per line, 4 bytes
//           per lane and step, aligned to the line start; the four lines of a warp have one class):
//           histogram addresses and per-position sums are formed with IDP.4A (FMA pipe), histograms
//           are lane-striped (conflict-free) shared-memory atomics, per-position sums live in
//           registers.  Pieces that cross tile edges and long lines (ONT) are processed cooperatively
//           in aligned 16-byte groups.
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

#include "fq_layout.h"

namespace fq {

typedef unsigned long long u64;

constexpr int TILE = 16384;
constexpr int THREADS = 512;
constexpr int SCAN_WARPS = 4;
constexpr int SCAN_THREADS = SCAN_WARPS * 32;
constexpr int WORK_THREADS = THREADS - SCAN_THREADS;
constexpr int WORK_WARPS = WORK_THREADS / 32;
constexpr int GPT = TILE / 16 / THREADS;      // 16-byte groups per thread (2)
constexpr int BM_WORDS = TILE / 32;           // bitmap words per tile
constexpr int WPS = BM_WORDS / SCAN_THREADS;  // bitmap words per scanner thread (4)
constexpr int NL_CAP = 2048;                  // newline index capacity; denser tiles take the walker path
constexpr int PAD = 16;
constexpr int STAGE_BYTES = PAD + TILE + 16;
constexpr int NSTAGE = 3;
constexpr int LONG_SEG = 1024;                // lines longer than this are processed cooperatively
constexpr int PIECE_CAP = 24;
constexpr int REG_STEPS = 5;                  // per-position sums of positions < 32*REG_STEPS live in registers
constexpr int HB = 128;                       // striped histogram bins (tiles with bytes >= 128 take the walker path)
static_assert(WPS == 4, "scanner threads read their bitmap words with one LDS.128");

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
};

// IDP.4A byte selectors.  Loaded once from shared memory into registers: as immediates or kernel
// parameters the compiler re-materialises each of them (UMOV / LDCU) in front of every IDP.4A.
struct Sel {
  uint32_t h0, h1, h2, h3;  // 128 << 8k : histogram address = byte_k * 128 + base
  uint32_t p0, p1, p2, p3;  // 1 << 8k   : per-position sum += byte_k
};

struct Piece {  // a run of content bytes handled cooperatively: [vs, ve) of the tile
  int vs, ve;
  uint32_t line;  // tile-relative line index
  uint32_t pad;
  u64 vpos;       // position of byte vs inside its line
};

struct TileMeta {
  u64 Lrel;     // newlines of the span before this tile
  u64 open;     // bytes of the open line (or of the head fragment) before this tile
  u64 toff;     // byte offset of the tile relative to base
  int T;        // newlines in the tile
  int lo, hi;   // valid byte range of the tile
  int walker;   // 1: dense or high-byte tile -> generic bitmap walker
  int npieces;
  int nrec;     // relevant lines in rec[] (interior lines and, when it is one, the tail piece)
};

struct __align__(128) Smem {
  uint8_t buf[NSTAGE][STAGE_BYTES];  // tile stages; data at buf[s] + PAD
  uint32_t hist[2][HB * 32];         // [0] sequence, [1] quality; word index = byte*32 + lane
  uint32_t ghist[2][256];            // un-striped tables of the generic paths
  uint32_t bitmap[2][BM_WORDS];      // bit b of word w: byte 32*w+b is '\n'
  uint16_t wordbase[2][BM_WORDS];    // newlines before bitmap word w
  uint16_t nl[2][NL_CAP];
  uint32_t rec[2][NL_CAP / 2 + 2];   // relevant (sequence / quality) lines of a tile: start | length << 14
  uint32_t seq_len[POS_BINS + 1];
  uint32_t qual_len[POS_BINS + 1];
  uint32_t seq_log2[LOG2_BINS];
  uint32_t pos_sum[POS_BINS + 1];
  Piece pieces[2][PIECE_CAP];
  TileMeta meta[2];
  uint32_t scan_tot[SCAN_WARPS];
  int scan_last[SCAN_WARPS];
  u64 full_bar[NSTAGE];              // mbarriers of the stages
  u64 len_min[2], len_max[2];        // [0] seq, [1] qual
  u64 run_L, run_open;               // span-running newline count / open-line bytes
  u64 head_len, junk[2];
  uint32_t bytes_since_flush, hiflag[2], head_done;
  uint32_t ksel[8];
};

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(u64* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tFQ_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra FQ_DONE;\n\tbra FQ_WAIT;\n\tFQ_DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, u64* bar) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic reads of dst before the async write
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void red_inc(uint32_t addr) {
  asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr) : "memory");
}

// 0x80 in every byte lane of w that equals '\n' (exact: no carries cross byte lanes)
__device__ __forceinline__ uint32_t nl_flags(uint32_t w) {
  uint32_t x = w ^ 0x0A0A0A0Au;
  return ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u;
}
// 16-bit mask of the '\n' bytes of a 16-byte group; the movemask is two IDP.4A chains (FMA pipe)
__device__ __forceinline__ uint32_t nl_mask16(const uint4& v) {
  uint32_t lo = __dp4a(nl_flags(v.x), 0x08040201u, __dp4a(nl_flags(v.y), 0x80402010u, 0u));
  uint32_t hi = __dp4a(nl_flags(v.z), 0x08040201u, __dp4a(nl_flags(v.w), 0x80402010u, 0u));
  return (lo >> 7) | (hi << 1);  // the flags weigh 128
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

// One byte of a sequence (cls 1) / quality (cls 3) line -- generic path, un-striped tables.
__device__ __forceinline__ void account_byte(uint32_t (*ghist)[256], uint32_t* pos_sum, int cls, uint32_t b, u64 pos) {
  atomicAdd(&ghist[cls == 3][b], 1u);
  if (cls == 3) {
    uint32_t p = pos < (u64)POS_BINS ? (uint32_t)pos : (uint32_t)POS_BINS;
    atomicAdd(&pos_sum[p], b);
  }
}
__device__ __forceinline__ void account_line_len(Smem& sm, int cls, u64 len, u64* my_min, u64* my_max) {
  const int q = cls == 3;
  const uint32_t bin = len < (u64)POS_BINS ? (uint32_t)len : (uint32_t)POS_BINS;
  if (q) atomicAdd(&sm.qual_len[bin], 1u);
  else { atomicAdd(&sm.seq_len[bin], 1u); atomicAdd(&sm.seq_log2[log2_bin(len)], 1u); }
  if (len < my_min[q]) my_min[q] = len;
  if (len > my_max[q]) my_max[q] = len;
}

// Per-position sums held in registers: lane `sub` of a quarter-warp owns positions
// 32*st + 4*sub + k (k < 4) for st < REG_STEPS.
struct PosAcc {
  uint32_t a[REG_STEPS][4];
};
__device__ __forceinline__ void flush_pos_acc(Smem& sm, PosAcc& acc, int sub) {
#pragma unroll
  for (int st = 0; st < REG_STEPS; st++)
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (acc.a[st][k]) atomicAdd(&sm.pos_sum[32 * st + 4 * sub + k], acc.a[st][k]);
      acc.a[st][k] = 0;
    }
}

// One 32-byte step of a line for one quarter-warp lane: 4 line bytes at shared address al (+4 for the
// funnel), `rem` = line bytes left from this lane's word on.  Bytes past the line end are forced to
// 0 and counted in histogram bin 0 ("junk"); the caller keeps the junk total and subtracts it at the
// end, so the step has no branches and no predicates.
// One 32-byte step of a line for one quarter-warp lane: 4 line bytes at shared address al (+4 for the
// funnel).  MASKED steps force the bytes past the line end to 0; those are counted in histogram bin 0
// ("junk"), the caller keeps the junk total and subtracts it at the end -- no predicates in the step.
// msh = 32 - 8*(line bytes left from this lane's word on): <= 0 full word, >= 32 nothing left.
template <bool QUAL, bool MASKED>
__device__ __forceinline__ void line_step(const Sel& k, uint32_t al, uint32_t sh, int msh, uint32_t hbase, uint32_t* acc4) {
  uint32_t w = __funnelshift_r(lds32(al), lds32(al + 4), sh);
  if (MASKED) w &= __funnelshift_rc(0xFFFFFFFFu, 0u, (uint32_t)min(max(msh, 0), 32));
  red_inc(__dp4a(w, k.h0, hbase));
  red_inc(__dp4a(w, k.h1, hbase));
  red_inc(__dp4a(w, k.h2, hbase));
  red_inc(__dp4a(w, k.h3, hbase));
  if (QUAL) {
    acc4[0] = __dp4a(w, k.p0, acc4[0]);
    acc4[1] = __dp4a(w, k.p1, acc4[1]);
    acc4[2] = __dp4a(w, k.p2, acc4[2]);
    acc4[3] = __dp4a(w, k.p3, acc4[3]);
  }
}

// Four lines of one class per warp (a quarter-warp each): n content bytes at shared address a0
// (n == 0: this quarter-warp has no line).  Steps that are full for all four lines run unmasked.
// Returns the histogram slots touched by this lane.
template <bool QUAL>
__device__ __forceinline__ uint32_t lines_fast(Smem& sm, const Sel& k, uint32_t a0, int n, int sub, uint32_t hbase, PosAcc& acc) {
  const uint32_t sh = (a0 & 3u) * 8u;
  const uint32_t al = (a0 & ~3u) + 4u * sub;
  const int rem = n - 4 * sub;
  const int msh = 32 - 8 * rem;  // mask shift of step 0; each step adds 256
  // steps in the register window: nf full ones, then at most one partial one (ms - nf <= 1).  The
  // four lines of a warp usually have the same length, so this per-quarter-warp control flow rarely diverges.
  const int ms = min((n + 31) >> 5, REG_STEPS), nf = min(n >> 5, REG_STEPS);
  switch (nf) {
    case 5: line_step<QUAL, false>(k, al + 128, sh, 0, hbase, acc.a[4]);
    case 4: line_step<QUAL, false>(k, al + 96, sh, 0, hbase, acc.a[3]);
    case 3: line_step<QUAL, false>(k, al + 64, sh, 0, hbase, acc.a[2]);
    case 2: line_step<QUAL, false>(k, al + 32, sh, 0, hbase, acc.a[1]);
    case 1: line_step<QUAL, false>(k, al, sh, 0, hbase, acc.a[0]);
    default: break;
  }
  if (ms > nf) {
    switch (nf) {
      case 0: line_step<QUAL, true>(k, al, sh, msh, hbase, acc.a[0]); break;
      case 1: line_step<QUAL, true>(k, al + 32, sh, msh + 256, hbase, acc.a[1]); break;
      case 2: line_step<QUAL, true>(k, al + 64, sh, msh + 512, hbase, acc.a[2]); break;
      case 3: line_step<QUAL, true>(k, al + 96, sh, msh + 768, hbase, acc.a[3]); break;
      default: line_step<QUAL, true>(k, al + 128, sh, msh + 1024, hbase, acc.a[4]); break;
    }
  }
  if (n > 32 * REG_STEPS) {  // positions beyond the register window (reads longer than 160)
    uint32_t a2 = al + 32 * REG_STEPS;
    int r2 = rem - 32 * REG_STEPS;
    for (int base = 32 * REG_STEPS; base < n; base += 32) {
      const uint32_t w = __funnelshift_r(lds32(a2), lds32(a2 + 4), sh);
#pragma unroll
      for (int kk = 0; kk < 4; kk++) {
        if (r2 > kk) {
          const uint32_t b = (w >> (8 * kk)) & 0xFFu;
          red_inc(hbase + (b << 7));
          if (QUAL) {
            const int p = base + 4 * sub + kk;
            atomicAdd(&sm.pos_sum[p < POS_BINS ? p : POS_BINS], b);
          }
        }
      }
      a2 += 32;
      r2 -= 32;
    }
  }
  return 4u * (uint32_t)ms;
}

// A run of content bytes [vs, ve) processed by `nthr` threads (rank `r`): aligned 16-byte groups
// through the striped histogram, the ragged ends byte-wise.  Bytes < 128 guaranteed by the caller.
__device__ __forceinline__ void piece_coop(Smem& sm, const uint8_t* buf, int vs, int ve, u64 vpos, int cls,
                                           uint32_t hbase, int r, int nthr) {
  const int body0 = (vs + 15) & ~15, body1 = ve & ~15;
  if (body1 > body0) {
    for (int o = body0 + r * 16; o < body1; o += nthr * 16) {
      const uint4 v = *reinterpret_cast<const uint4*>(buf + o);
      const uint32_t ww[4] = {v.x, v.y, v.z, v.w};
      const u64 p = vpos + (u64)(o - vs);
      uint32_t over = 0;
#pragma unroll
      for (int x = 0; x < 4; x++) {
        red_inc(__dp4a(ww[x], 0x00000080u, hbase));
        red_inc(__dp4a(ww[x], 0x00008000u, hbase));
        red_inc(__dp4a(ww[x], 0x00800000u, hbase));
        red_inc(__dp4a(ww[x], 0x80000000u, hbase));
        over = __dp4a(ww[x], 0x01010101u, over);
      }
      if (cls == 3) {
        if (p >= (u64)POS_BINS) atomicAdd(&sm.pos_sum[POS_BINS], over);
        else {
#pragma unroll
          for (int x = 0; x < 16; x++) {
            const u64 pp = p + x;
            atomicAdd(&sm.pos_sum[pp < (u64)POS_BINS ? (uint32_t)pp : (uint32_t)POS_BINS], (ww[x >> 2] >> (8 * (x & 3))) & 0xFFu);
          }
        }
      }
    }
    for (int o = vs + r; o < body0; o += nthr) account_byte(sm.ghist, sm.pos_sum, cls, buf[o], vpos + (u64)(o - vs));
    for (int o = body1 + r; o < ve; o += nthr) account_byte(sm.ghist, sm.pos_sum, cls, buf[o], vpos + (u64)(o - vs));
  } else {
    for (int o = vs + r; o < ve; o += nthr) account_byte(sm.ghist, sm.pos_sum, cls, buf[o], vpos + (u64)(o - vs));
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
                                         const uint32_t* bitmap, const uint16_t* wordbase, int r, int nthr,
                                         u64* my_min, u64* my_max) {
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
      const bool counted = (cls & 1) && line != 0;  // line 0 of the span is the head fragment (stitch kernel)
      if (b == '\n') {
        if (counted) {
          const int cr = (pos > 0 && byte_before(a, buf, m, o) == '\r') ? 1 : 0;
          account_line_len(sm, cls, pos - (u64)cr, my_min, my_max);
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
        if (content) account_byte(sm.ghist, sm.pos_sum, cls, b, pos);
      }
      pos++;
    }
  }
}

__device__ __forceinline__ void flush_pos_sum_to(Smem& sm, u64* block, int tid) {
  for (int i = tid; i <= POS_BINS; i += THREADS) {
    uint32_t v = sm.pos_sum[i];
    if (v) { block[OFF_POS_SUM + i] += v; sm.pos_sum[i] = 0; }
  }
}

__global__ void __launch_bounds__(THREADS, 2) fq_scan_kernel(const ScanArgs a, const int pass) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, sub = tid & 7;
  const bool scanner = warp < SCAN_WARPS;
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

  for (int i = tid; i < 2 * HB * 32; i += THREADS) (&sm.hist[0][0])[i] = 0;
  for (int i = tid; i < 512; i += THREADS) (&sm.ghist[0][0])[i] = 0;
  for (int i = tid; i <= POS_BINS; i += THREADS) { sm.seq_len[i] = 0; sm.qual_len[i] = 0; sm.pos_sum[i] = 0; }
  if (tid < LOG2_BINS) sm.seq_log2[tid] = 0;
  if (pass == 0) {
    for (int i = tid; i < BLOCK_WORDS; i += THREADS) block[i] = (i == OFF_SEQ_LEN_MIN || i == OFF_QUAL_LEN_MIN) ? ~0ull : 0ull;
  }
  if (tid == 0) {
    sm.len_min[0] = sm.len_min[1] = ~0ull;
    sm.len_max[0] = sm.len_max[1] = 0;
    sm.bytes_since_flush = 0;
    sm.hiflag[0] = sm.hiflag[1] = 0;
    sm.run_L = 0; sm.run_open = 0; sm.head_len = 0; sm.head_done = 0;
    sm.junk[0] = sm.junk[1] = 0;
    for (int k = 0; k < 4; k++) { sm.ksel[k] = 0x80u << (8 * k); sm.ksel[4 + k] = 1u << (8 * k); }
    for (int s = 0; s < NSTAGE; s++) mbar_init(&sm.full_bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (t0 < t1) {  // prologue: the first tile into stage 0
      const u64 toff = (u64)t0 * TILE;
      const uint32_t bytes = (uint32_t)((((a.end - toff) < (u64)TILE ? (a.end - toff) : (u64)TILE) + 15) & ~15ull);
      mbar_expect_tx(&sm.full_bar[0], bytes);
      tma_load_1d(&sm.buf[0][PAD], a.base + toff, bytes, &sm.full_bar[0]);
    }
  }
  u64 my_min[2] = {~0ull, ~0ull}, my_max[2] = {0, 0};  // line-length extrema seen by this thread
  u64 slots[2] = {0, 0}, valid[2] = {0, 0};             // histogram slots touched / line bytes (junk = slots - valid)
  PosAcc acc;
#pragma unroll
  for (int st = 0; st < REG_STEPS; st++)
#pragma unroll
    for (int k = 0; k < 4; k++) acc.a[st][k] = 0;
  const uint32_t sm0 = smem_u32(smem_raw);
  const uint32_t hb_seq = sm0 + (uint32_t)offsetof(Smem, hist) + 4u * lane;
  const uint32_t hb_qual = hb_seq + HB * 32 * 4;
  __syncthreads();
  Sel ksel;
  {
    const uint32_t ks = sm0 + (uint32_t)offsetof(Smem, ksel);
    ksel.h0 = lds32(ks); ksel.h1 = lds32(ks + 4); ksel.h2 = lds32(ks + 8); ksel.h3 = lds32(ks + 12);
    ksel.p0 = lds32(ks + 16); ksel.p1 = lds32(ks + 20); ksel.p2 = lds32(ks + 24); ksel.p3 = lds32(ks + 28);
  }

  // pipeline: iteration `it` classifies tile B = t0+it (stage it%3, slot it&1) while the workers
  // take tile C = B-1; the TMA of tile A = B+1 is started at the top.
  const int nt = (int)(t1 - t0);
  uint32_t par_bits = 0;  // mbarrier parity per stage (bit s)
  int stB = 0, stC = 2, stA = 1;  // stages of tiles B, C, A; rotated at the end of every iteration
  for (int it = 0; it <= nt; it++) {
    const int sb = it & 1, sc = sb ^ 1;
    const bool haveB = it < nt, haveC = it > 0;
    const uint32_t tileB = t0 + (uint32_t)it;

    if (tid == 0 && it + 1 < nt) {  // stage stA held tile B-2, whose K2 finished last iteration
      const u64 noff = (u64)(tileB + 1) * TILE;
      const uint32_t bytes = (tileB + 2 < a.ntiles) ? (uint32_t)TILE : (uint32_t)(((a.end - noff) + 15) & ~15ull);
      mbar_expect_tx(&sm.full_bar[stA], bytes);
      tma_load_1d(&sm.buf[stA][PAD], a.base + noff, bytes, &sm.full_bar[stA]);
    }

    // ---- K1a (all warps): newline masks of tile B's 16-byte groups -> bitmap ----
    int loB = 0, hiB = 0;
    const u64 toffB = (u64)tileB * TILE;
    if (haveB) {
      loB = (tileB == 0) ? (int)a.lo0 : 0;
      hiB = (tileB + 1 < a.ntiles) ? TILE : (int)(a.end - toffB);
      mbar_wait(&sm.full_bar[stB], (par_bits >> stB) & 1u);
      par_bits ^= 1u << stB;
      const uint8_t* buf = &sm.buf[stB][PAD];
      uint16_t* bm16 = reinterpret_cast<uint16_t*>(sm.bitmap[sb]);
      uint32_t hib = 0;
      if (loB == 0 && hiB == TILE) {  // interior tile: no edge handling
#pragma unroll
        for (int j = 0; j < GPT; j++) {
          const int g = tid + j * THREADS;
          const uint4 v = *reinterpret_cast<const uint4*>(buf + g * 16);
          hib |= (v.x | v.y) | (v.z | v.w);
          bm16[g] = (uint16_t)nl_mask16(v);
        }
      } else {
#pragma unroll
        for (int j = 0; j < GPT; j++) {
          const int g = tid + j * THREADS;
          const int off = g * 16;
          uint32_t m = 0;
          if (off < hiB && off + 16 > loB) {
            const uint4 v = *reinterpret_cast<const uint4*>(buf + off);
            m = nl_mask16(v);
            hib |= (v.x | v.y) | (v.z | v.w);
            int lo_k = loB - off; lo_k = lo_k < 0 ? 0 : lo_k;
            int hi_k = hiB - off; hi_k = hi_k > 16 ? 16 : hi_k;
            m &= ((1u << hi_k) - 1u) & ~((1u << lo_k) - 1u);
          }
          bm16[g] = (uint16_t)m;
        }
      }
      if (hib & 0x80808080u) sm.hiflag[sb] = 1;
    }

    if (scanner) {
      // =====================================================================================
      // SCANNER warps: newline index of tile B, span-running carry, line-length tables, pieces
      // =====================================================================================
      bar_sync(1, THREADS);  // bitmap of tile B complete (workers only arrive)
      if (haveB) {
        const uint8_t* buf = &sm.buf[stB][PAD];
        const uint4 bw4 = *reinterpret_cast<const uint4*>(&sm.bitmap[sb][tid * WPS]);
        const uint32_t bw[WPS] = {bw4.x, bw4.y, bw4.z, bw4.w};
        const uint32_t c0 = __popc(bw[0]), c1 = __popc(bw[1]), c2 = __popc(bw[2]), c3 = __popc(bw[3]);
        const uint32_t c = c0 + c1 + c2 + c3;
        const uint32_t inc = warp_incl_scan(c, lane);
        int my_last = -1;
#pragma unroll
        for (int x = 0; x < WPS; x++) if (bw[x]) my_last = (tid * WPS + x) * 32 + 31 - __clz(bw[x]);
        my_last = __reduce_max_sync(0xffffffffu, my_last);
        if (lane == 31) { sm.scan_tot[warp] = inc; sm.scan_last[warp] = my_last; }
        bar_sync(2, SCAN_THREADS);
        uint32_t wbase = 0, T = 0;
        int last_nl = -1;
#pragma unroll
        for (int w = 0; w < SCAN_WARPS; w++) {
          const uint32_t x = sm.scan_tot[w];
          if (w < warp) wbase += x;
          T += x;
          last_nl = max(last_nl, sm.scan_last[w]);
        }
        uint32_t first = wbase + inc - c;  // index of this thread's first newline
        {
          uint16_t* wb = &sm.wordbase[sb][tid * WPS];
          wb[0] = (uint16_t)first; wb[1] = (uint16_t)(first + c0); wb[2] = (uint16_t)(first + c0 + c1); wb[3] = (uint16_t)(first + c0 + c1 + c2);
        }
        const bool walker = T > (uint32_t)NL_CAP || sm.hiflag[sb] != 0 || (a.dbg & 2);
        if (!walker) {
#pragma unroll
          for (int x = 0; x < WPS; x++) {
            uint32_t m = bw[x];
            while (m) {
              const int k = __ffs(m) - 1;
              m &= m - 1;
              sm.nl[sb][first] = (uint16_t)((tid * WPS + x) * 32 + k);
              first++;
            }
          }
        }
        const u64 Lrel = sm.run_L, open = sm.run_open;  // before this tile (written last iteration)
        TileMeta& m = sm.meta[sb];
        if (tid == 0) {
          m.Lrel = Lrel; m.open = open; m.toff = toffB; m.T = (int)T; m.lo = loB; m.hi = hiB;
          m.walker = walker ? 1 : 0; m.npieces = walker ? 0 : 2; m.nrec = 0;
          sm.hiflag[sb] = 0;
          sm.bytes_since_flush += (uint32_t)(hiB - loB);
        }
        bar_sync(2, SCAN_THREADS);  // nl index, meta and the old running carry are visible / consumed
        if (tid == 0) {
          sm.run_L = Lrel + T;
          sm.run_open = T ? (u64)(hiB - (last_nl + 1)) : open + (u64)(hiB - loB);
          if (!sm.head_done) {  // bytes of the span before its first newline (the stitch kernel's share)
            if (T) {
              int first_nl = 0;  // lowest set bit of the bitmap (the newline index is not built for walker tiles)
              for (int w = 0; w < BM_WORDS; w++) { const uint32_t x = sm.bitmap[sb][w]; if (x) { first_nl = w * 32 + __ffs(x) - 1; break; } }
              sm.head_len = open + (u64)(first_nl - loB);
              sm.head_done = 1;
            } else {
              sm.head_len = open + (u64)(hiB - loB);
            }
          }
        }
        if (walker) {
        } else if (!count_only) {
          const uint32_t ph = (uint32_t)((phase + Lrel) & 3);  // class of the tile's line 0
          const int jr0 = (ph & 1) ? 2 : 1;                    // first interior line with an odd class
          const int R = (int)T > jr0 ? ((int)T - jr0 + 1) >> 1 : 0;
          if (tid == 0) {
            // head piece: the line open at the tile start continues up to the first newline;
            // tail piece: the line open at the tile end.  A '\r' directly before '\n' is dropped.
            Piece hp; hp.vs = 0; hp.ve = 0; hp.line = 0; hp.pad = 0; hp.vpos = open;
            {
              const int e0 = T ? (int)sm.nl[sb][0] : hiB;
              int ve = e0;
              if (T) { if (ve > loB && buf[ve - 1] == '\r') ve--; }
              else if (ve > loB && buf[ve - 1] == '\r') { const int nx = byte_after_tile(a, m); if (nx == '\n' || nx < 0) ve--; }
              hp.vs = loB; hp.ve = ve;
            }
            sm.pieces[sb][0] = hp;
            Piece tp; tp.vs = 0; tp.ve = 0; tp.line = T; tp.pad = 0; tp.vpos = 0;
            int nrec = R;
            if (T) {
              const int vs = last_nl + 1;
              int ve = hiB;
              if (ve > vs && buf[ve - 1] == '\r') { const int nx = byte_after_tile(a, m); if (nx == '\n' || nx < 0) ve--; }
              // the tail piece starts a line in this tile: when it has an odd class it is the next relevant
              // line (line T), so it simply extends the record list; long ones stay cooperative pieces
              const bool relevant = (int)T >= jr0 && (((int)T - jr0) & 1) == 0;
              if (relevant && ve - vs <= LONG_SEG) { sm.rec[sb][R] = ve > vs ? ((uint32_t)vs | ((uint32_t)(ve - vs) << 14)) : 0u; nrec = R + 1; }
              else { tp.vs = vs; tp.ve = ve; }
            }
            sm.pieces[sb][1] = tp;
            m.nrec = nrec;
          }
          // line-length tables of the lines that end in tile B, records of the relevant interior lines;
          // long interior lines join the pieces
          for (int j = tid; j < (int)T; j += SCAN_THREADS) {
            const int e = (int)sm.nl[sb][j];
            const int s = j ? (int)sm.nl[sb][j - 1] + 1 : loB;
            const u64 lidx = Lrel + (u64)j;
            const int cls = (int)((ph + (uint32_t)j) & 3);
            if ((cls & 1) && lidx != 0) {
              const u64 raw = (j ? 0ull : open) + (u64)(e - s);
              const int cr = (raw > 0 && byte_before(a, buf, m, e) == '\r') ? 1 : 0;
              account_line_len(sm, cls, raw - (u64)cr, my_min, my_max);
              if (j) {
                uint32_t n = (uint32_t)(e - cr - s);
                if (e - s > LONG_SEG) {
                  const int q = atomicAdd(&m.npieces, 1);
                  Piece p; p.vs = s; p.ve = e - cr; p.line = (uint32_t)j; p.pad = 0; p.vpos = 0;
                  sm.pieces[sb][q] = p;
                  n = 0;
                }
                sm.rec[sb][(j - jr0) >> 1] = (uint32_t)s | (n << 14);
              }
            }
          }
        }
      }
    } else {
      // =====================================================================================
      // WORKER warps: byte statistics of tile C
      // =====================================================================================
      bar_arrive(1, THREADS);
      if (haveC && !count_only) {
        const TileMeta& m = sm.meta[sc];
        const uint8_t* buf = &sm.buf[stC][PAD];
        const uint32_t buf_s = sm0 + (uint32_t)offsetof(Smem, buf) + (uint32_t)stC * STAGE_BYTES + PAD;
        const int wr = tid - SCAN_THREADS;  // worker rank
        if (m.walker) {
          tile_walker(sm, a, buf, m, phase, sm.bitmap[sc], sm.wordbase[sc], wr, WORK_THREADS, my_min, my_max);
        } else {
          const uint32_t ph = (uint32_t)((phase + m.Lrel) & 3);  // class of the tile's line 0
          // relevant line r (record r) is tile line jr0 + 2r; its class alternates with r.  A warp takes
          // the four even (or the four odd) records of a group of eight: one class per warp.
          const int jr0 = (ph & 1) ? 2 : 1;
          const int R = m.nrec;
          const int ww = warp - SCAN_WARPS, qi = lane >> 3;
          static_assert((WORK_WARPS & 1) == 0, "a warp keeps its record parity (= class) for the whole tile");
          const bool qual = ((ph + (uint32_t)(jr0 + 2 * (ww & 1))) & 3) == 3;  // warp-uniform, same for all its tasks
          const uint32_t rec_s = sm0 + (uint32_t)offsetof(Smem, rec) + (uint32_t)sc * (uint32_t)sizeof(sm.rec[0]);
          // task: records 8g + 2*qi + (ww & 1), g = ww>>1, ww>>1 + WORK_WARPS/2, ...
          if (qual) {
            for (int r = 8 * (ww >> 1) + (ww & 1) + 2 * qi; r - 2 * qi < R; r += 4 * WORK_WARPS) {
              const uint32_t rc = r < R ? lds32(rec_s + 4u * (uint32_t)r) : 0u;
              const int n = (int)(rc >> 14);
              slots[1] += lines_fast<true>(sm, ksel, buf_s + (rc & 0x3FFFu), n, sub, hb_qual, acc);
              if (sub == 0) valid[1] += (u64)min(n, 32 * REG_STEPS);
            }
          } else {
            for (int r = 8 * (ww >> 1) + (ww & 1) + 2 * qi; r - 2 * qi < R; r += 4 * WORK_WARPS) {
              const uint32_t rc = r < R ? lds32(rec_s + 4u * (uint32_t)r) : 0u;
              const int n = (int)(rc >> 14);
              slots[0] += lines_fast<false>(sm, ksel, buf_s + (rc & 0x3FFFu), n, sub, hb_seq, acc);
              if (sub == 0) valid[0] += (u64)min(n, 32 * REG_STEPS);
            }
          }
          // pieces: a short one is taken by a single warp, a long one by all worker threads together
          const int np = m.npieces;
          for (int k = 0; k < np; k++) {
            const Piece p = sm.pieces[sc][k];
            const u64 lidx = m.Lrel + p.line;
            const int cls = (int)((phase + lidx) & 3);
            const int len = p.ve - p.vs;
            if (!(cls & 1) || lidx == 0 || len <= 0) continue;
            if (len <= 512) {
              if (ww == (k + (int)(m.toff >> 14)) % WORK_WARPS)
                for (int o = p.vs + lane; o < p.ve; o += 32) account_byte(sm.ghist, sm.pos_sum, cls, buf[o], p.vpos + (u64)(o - p.vs));
            } else {
              piece_coop(sm, buf, p.vs, p.ve, p.vpos, cls, cls == 3 ? hb_qual : hb_seq, wr, WORK_THREADS);
            }
          }
        }
      }
    }

    // ---- end of iteration: tile C is consumed, B becomes C ----
    { const int t = stC; stC = stB; stB = stA; stA = t; }
    __syncthreads();
    if (sm.bytes_since_flush > (1u << 24)) {  // keep the 32-bit per-position sums from overflowing
      flush_pos_acc(sm, acc, sub);
      __syncthreads();
      flush_pos_sum_to(sm, block, tid);
      if (tid == 0) sm.bytes_since_flush = 0;
      __syncthreads();
    }
  }

  // ---- flush this span's counters into its block; pass 0 also records the span descriptor ----
  if (!scanner) flush_pos_acc(sm, acc, sub);
  for (int q = 0; q < 2; q++) {
    if (my_min[q] != ~0ull) atomicMin(&sm.len_min[q], my_min[q]);
    if (my_max[q] != 0) atomicMax(&sm.len_max[q], my_max[q]);
    const u64 j = slots[q] - valid[q];  // may wrap per thread; the sum over the CTA is the junk total
    if (slots[q] | valid[q]) atomicAdd(&sm.junk[q], j);
  }
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
    if (b == 0) s -= sm.junk[h];  // histogram slots past line ends were counted as byte 0
    if (s) block[OFF_HIST_SEQ + bin] += s;
  }
  for (int i = tid; i <= POS_BINS; i += THREADS) {
    if (sm.seq_len[i]) block[OFF_SEQ_LEN + i] += sm.seq_len[i];
    if (sm.qual_len[i]) block[OFF_QUAL_LEN + i] += sm.qual_len[i];
  }
  if (tid < LOG2_BINS && sm.seq_log2[tid]) block[OFF_SEQ_LOG2 + tid] += sm.seq_log2[tid];
  flush_pos_sum_to(sm, block, tid);
  if (tid == 0) {
    if (sm.len_min[0] < block[OFF_SEQ_LEN_MIN]) block[OFF_SEQ_LEN_MIN] = sm.len_min[0];
    if (sm.len_max[0] > block[OFF_SEQ_LEN_MAX]) block[OFF_SEQ_LEN_MAX] = sm.len_max[0];
    if (sm.len_min[1] < block[OFF_QUAL_LEN_MIN]) block[OFF_QUAL_LEN_MIN] = sm.len_min[1];
    if (sm.len_max[1] > block[OFF_QUAL_LEN_MAX]) block[OFF_QUAL_LEN_MAX] = sm.len_max[1];
  }
}

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
  if (lane == 0) { d.T = 0; d.head_len = 0; d.tail_len = 0; d.state = SPAN_PENDING; d.exact = PHASE_UNKNOWN; d.G = 0; d.P0 = 0; }
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
    const int ok = hyp_ok && d.guess == exact;
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
  if (!(cls & 1)) return;
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

// ------------------------------------------------------------------------------------------
// host-callable launchers (used by fqgpu_api.cu)
// ------------------------------------------------------------------------------------------
size_t scan_smem_bytes() { return sizeof(Smem); }
int scan_tile_bytes() { return TILE; }
int scan_threads() { return THREADS; }

cudaError_t scan_configure() {
  return cudaFuncSetAttribute(fq_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
}

cudaError_t launch_reset(u64* committed, int nblocks, Carry* carry, cudaStream_t st) {
  fq_reset_kernel<<<64, 256, 0, st>>>(committed, nblocks, carry);
  return cudaGetLastError();
}

// Scans `nbytes` at `ptr` (any alignment) as the continuation of the stream described by `carry`:
// resync -> scan (guessed phases) -> stitch (verify, commit, head fragments, carry) -> scan pass 1
// (only spans whose guess was wrong or unknown; exits immediately otherwise).
cudaError_t launch_scan(const void* ptr, size_t nbytes, SpanDesc* desc, LaunchHdr* hdr, Carry* carry,
                        u64* pending, u64* committed, ShardInfo* shard, int max_spans, u64 meta_records, cudaStream_t st) {
  if (nbytes == 0) return cudaSuccess;
  static const uint32_t dbg = getenv("FQGPU_DEBUG") ? (uint32_t)atoi(getenv("FQGPU_DEBUG")) : 0u;
  const uintptr_t addr = (uintptr_t)ptr;
  ScanArgs a;
  a.lo0 = (uint32_t)(addr & 15);
  a.base = (const uint8_t*)(addr - a.lo0);
  a.end = (u64)a.lo0 + nbytes;
  a.ntiles = (uint32_t)((a.end + TILE - 1) / TILE);
  uint32_t nspans = a.ntiles < (uint32_t)max_spans ? a.ntiles : (uint32_t)max_spans;
  a.tps = (a.ntiles + nspans - 1) / nspans;
  a.nspans = (a.ntiles + a.tps - 1) / a.tps;
  a.desc = desc; a.hdr = hdr; a.carry = carry; a.pending = pending; a.committed = committed; a.shard = shard; a.dbg = dbg;

  if (meta_records) fq_meta_kernel<<<1, 32, 0, st>>>(a.base, a.lo0, a.end, carry, meta_records);
  fq_resync_kernel<<<a.nspans, 32, 0, st>>>(a);
  fq_scan_kernel<<<a.nspans, THREADS, sizeof(Smem), st>>>(a, 0);
  fq_stitch_kernel<<<a.nspans, STITCH_THREADS, 0, st>>>(a);
  fq_scan_kernel<<<a.nspans, THREADS, sizeof(Smem), st>>>(a, 1);
  return cudaGetLastError();
}

cudaError_t launch_reduce(const u64* blocks, int nblocks, u64* out, cudaStream_t st) {
  fq_reduce_kernel<<<(BLOCK_WORDS + 31) / 32, 256, 0, st>>>(blocks, nblocks, out);
  return cudaGetLastError();
}

}  // namespace fq
