// fq_gzip.cu -- on-device inflate of ordinary (single-member) gzip input, feeding the scan kernels (SURVEY 8f
// rank 3; BASELINE configs[4]).
//
// The reference inflates .gz input on the host, one byte stream through zlib's gzread (src/utils/gzip_stream.nim:
// 16-17, opened at src/fq_count.nim:31-32 and src/fq_meta.nim:219-220): one thread, 0.45 GB/s of FASTQ on this
// box, the GPU idle.  A gzip member is ONE DEFLATE stream -- a block's first bit is known only when the block
// before it has been decoded, and a match may reach 32 KiB back into bytes another block produced -- so there is
// nothing to hand to a second thread, unless both are guessed and the guesses are then proven:
//
//   1. gz_sync_kernel   the compressed batch is cut into chunks; one warp per chunk looks for the first bit
//                       position that holds a well-formed dynamic-Huffman block header (BFINAL = 0, BTYPE = 2,
//                       counts in range, a complete code-length code, a complete literal/length code with an
//                       end-of-block symbol, a usable distance code).  32 lanes test 32 positions at once on the
//                       17 fixed bits and the Kraft sum of the code-length code; the few survivors are parsed by
//                       the whole warp.
//   2. gz_both_kernel   every chunk that found a start decodes from it until a block ends exactly on the next
//                       chunk's start: that proves the next start (a start the decode runs over in mid-block was a
//                       false positive and is skipped).  The chunk records where it came to rest, how many bytes it
//                       produced and how far back its matches reached -- and writes its output, as 16-bit symbols, to
//                       its own slot of an over-sized arena: a byte, or -- for a match that reaches behind the
//                       chunk's first byte -- a MARKER holding the position in the 32 KiB window before the chunk.
//                       Matches copy symbols, so markers propagate.
//   3. gz_chain_kernel  follows the landings from the batch's first block (whose position IS known: the end of
//                       the gzip header, or where the batch before stopped), gives the chunks on that chain
//                       their output offsets, and checks every look-back against the bytes that exist.
//   4. (gz_count_kernel + gz_write_kernel: the same in two passes -- count first, then write into exact places --
//                       for a batch whose symbols did not fit the arena, or when no arena can be had.)
//   5. gz_rows_kernel   "the 32 KiB before chunk i+1" from "the 32 KiB before chunk i" is a serial recurrence; windows of
//                       symbols compose, so groups of chunks are walked in parallel relative to their own first
//                       window and only the groups are walked in series (shared memory, fed by TMA).
//   6. gz_resolve_kernel every marker is replaced through its chunk's window and its group's; bytes go to the
//                       buffer fqgpu_scan_device reads.
//   7. gz_crc_*         the CRC-32 of those bytes, which the host compares with the member's trailer like gzread.
//
// The decoder (fq_inflate.cuh) is the BGZF path's: a warp per stream in lockstep, lookup tables in shared memory,
// cooperative match copies.  Whatever is not proven -- a broken chain, a stream zlib would reject, a truncated
// file -- makes the caller fall back to gzread, so results and error behaviour stay the reference's.
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "fq_dev.cuh"
#include "fq_gzip.h"
#include "fq_inflate.cuh"

namespace fq {

#ifndef GZ_MINB
#define GZ_MINB 10  // resident CTAs per SM the decode kernels are compiled for (register cap)
#endif
constexpr int GZ_WARPS = 4;        // chunks per CTA
constexpr int GZ_MAX_PASSED = 6;   // false starts one chunk may run over before it gives up
constexpr uint32_t GZ_MAX_OUT = 0xF0000000u;

// The code lengths of a dynamic block (RFC 1951 3.2.7) into t.lens[0 .. nlen + ndist); the reader stands behind
// the three block-header bits.  Returns nonzero for what zlib's inflate() calls an invalid block.
__device__ __forceinline__ int gz_dyn_lengths(GzBits& b, WarpTables& t, int lane, int& nlen, int& ndist) {
  b.refill();
  nlen = (int)b.take(5) + 257; ndist = (int)b.take(5) + 1;
  const int ncode = (int)b.take(4) + 4;
  if (nlen > 286 || ndist > 30) return 1;
  for (int k = lane; k < 19; k += 32) t.lens[k] = 0;
  __syncwarp();
  for (int k = 0; k < ncode; k++) { b.refill(); const uint32_t v = b.take(3); if (lane == 0) t.lens[kClOrder[k]] = (uint8_t)v; }
  __syncwarp();
  if (huff_build<HUFF_PRECODE>(t.dcount, t.dsym, t.dist, PBITS, t.lens, 19, lane) != 0) return 1;  // the code-length code must be complete
  int idx = 0;
  while (idx < nlen + ndist) {
    const uint32_t pe = huff_peek<HUFF_PRECODE, PBITS>(b, t.dist, t.dcount, t.dsym);
    if (!(pe & 15u)) return 1;
    const int sym = (int)huff_take(b, pe);
    if (sym < 16) { if (lane == 0) t.lens[idx] = (uint8_t)sym; idx++; continue; }
    int prev = 0, rep;
    b.refill();
    __syncwarp();
    if (sym == 16) { if (idx == 0) return 1; prev = t.lens[idx - 1]; rep = 3 + (int)b.take(2); }
    else if (sym == 17) rep = 3 + (int)b.take(3);
    else rep = 11 + (int)b.take(7);
    if (idx + rep > nlen + ndist) return 1;
    if (lane == 0) for (int k = 0; k < rep; k++) t.lens[idx + k] = (uint8_t)prev;
    idx += rep;
    __syncwarp();
  }
  __syncwarp();
  if (t.lens[256] == 0) return 1;  // no end-of-block code
  return 0;
}

// ---- 1. block-start search ---------------------------------------------------------------------------------------
// Is bit t the first bit of a dynamic block the way a compressor writes one?  (Stricter than inflate(): both codes
// complete -- zlib, libdeflate, igzip and pigz always emit complete codes -- or a distance code of at most one
// symbol; no trailing zero lengths.  A real start that fails this is merely not used: the chunk before decodes through it.)
__device__ bool gz_header_plausible(const uint32_t* words, const uint32_t* wend, u64 t, WarpTables& tb, int lane) {
  GzBits b;
  b.init(words, wend, t + 3);
  int nlen, ndist;
  if (gz_dyn_lengths(b, tb, lane, nlen, ndist)) return false;
  // a compressor sends no more lengths than it has to: the last one of either alphabet is not zero
  if (tb.lens[nlen - 1] == 0 || (ndist > 1 && tb.lens[nlen + ndist - 1] == 0)) return false;
  uint32_t sl = 0, sd = 0, nd = 0;
  for (int s = lane; s < nlen; s += 32) { const int l = tb.lens[s]; if (l) sl += 32768u >> l; }
  for (int s = lane; s < ndist; s += 32) { const int l = tb.lens[nlen + s]; if (l) { sd += 32768u >> l; nd++; } }
  sl = __reduce_add_sync(0xffffffffu, sl);
  sd = __reduce_add_sync(0xffffffffu, sd);
  nd = __reduce_add_sync(0xffffffffu, nd);
  __syncwarp();
  if (sl != 32768u) return false;
  return sd == 32768u || nd == 0 || (nd == 1 && sd == 16384u);
}

__global__ void __launch_bounds__(32 * GZ_WARPS) gz_sync_kernel(const uint32_t* __restrict__ words, u64 nbytes, uint32_t chunk_bytes,
                                                                int nchunks, u64 start_bit, GzChunk* __restrict__ chunks, uint32_t* nfound) {
  __shared__ WarpTables tables[GZ_WARPS];
  __shared__ uint8_t kraft3[512];  // three 3-bit code lengths -> their Kraft terms (2^(7 - len), 0 for an unused symbol)
  for (int x = threadIdx.x; x < 512; x += 32 * GZ_WARPS)
    kraft3[x] = (uint8_t)(((128u >> (x & 7)) & 127u) + ((128u >> ((x >> 3) & 7)) & 127u) + ((128u >> (x >> 6)) & 127u));
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * GZ_WARPS + warp;
  if (c >= nchunks) return;
  const uint32_t* wend = words + ((nbytes + 3) >> 2);
  const u64 end_bits = nbytes * 8ull;
  u64 found = GZ_NONE;
  if (c == 0) {
    found = start_bit;
  } else {
    u64 lo = (u64)c * chunk_bytes * 8ull, hi = lo + (u64)chunk_bytes * 8ull;
    if (lo <= start_bit) lo = start_bit + 1;
    const u64 last = end_bits > 96 ? end_bits - 96 : 0;  // a block header and an end-of-block do not fit behind this
    if (hi > last) hi = last;
    // A round covers 32 words = 1024 positions; every lane takes the 32 positions that start in ITS word.  The
    // three header bits (BFINAL = 0, BTYPE = 2: bits 0, 0, 1 from the position on) leave about four of them; for each
    // the lane checks the rest of the fixed bits and the Kraft sum of the code-length code (the (HCLEN + 4) lengths of
    // 3 bits from bit 17, three per table lookup: a complete code sums to 128; the last one sent is not zero), and
    // keeps a bit mask of the survivors.  (Before: one position per lane and round -- four instructions per position
    // where this takes a tenth of that.)  The few survivors are parsed by the whole warp, in stream order.
    for (u64 r0 = lo & ~1023ull; r0 < hi && found == GZ_NONE; r0 += 1024) {
      const u64 t0 = r0 + 32ull * (u64)lane;  // the first position of this lane's word
      const uint32_t* q = words + (t0 >> 5);
      uint32_t x[5];
#pragma unroll
      for (int i = 0; i < 5; i++) x[i] = q + i < wend ? __ldg(q + i) : 0u;
      // candidates: bit p of the word and the two bits above it are 0, 0, 1
      uint32_t cand = ~x[0] & ~__funnelshift_r(x[0], x[1], 1) & __funnelshift_r(x[0], x[1], 2);
      if (t0 < lo) cand &= t0 + 32 <= lo ? 0u : ~((1u << (uint32_t)(lo - t0)) - 1u);
      if (t0 + 32 > hi) cand &= t0 >= hi ? 0u : (1u << (uint32_t)(hi - t0)) - 1u;
      uint32_t surv = 0;
      while (cand) {
        const uint32_t j = (uint32_t)__ffs((int)cand) - 1u;
        cand &= cand - 1u;
        const uint32_t a = __funnelshift_r(x[0], x[1], j), bq = __funnelshift_r(x[1], x[2], j), cq = __funnelshift_r(x[2], x[3], j);
        if (((a >> 3) & 31u) > 29u || ((a >> 8) & 31u) > 29u) continue;  // HLIT, HDIST
        const int ncode = (int)((a >> 13) & 15u) + 4;
        u64 pre = (((((u64)bq << 32) | a) >> 17) | ((u64)cq << 47)) & ((1ull << (3 * ncode)) - 1ull);
        if (ncode != 4 && (pre >> (3 * (ncode - 1))) == 0) continue;
        uint32_t kraft = 0;
#pragma unroll
        for (int k = 0; k < 7; k++) { kraft += kraft3[(uint32_t)pre & 511u]; pre >>= 9; }
        if (kraft == 128u) surv |= 1u << j;
      }
      uint32_t lanes = __ballot_sync(0xffffffffu, surv != 0u);
      while (lanes && found == GZ_NONE) {
        const int l = __ffs((int)lanes) - 1;
        lanes &= lanes - 1;
        uint32_t sv = __shfl_sync(0xffffffffu, surv, l);
        while (sv) {
          const uint32_t j = (uint32_t)__ffs((int)sv) - 1u;
          sv &= sv - 1u;
          const u64 t = r0 + 32ull * (u64)l + j;
          if (gz_header_plausible(words, wend, t, tables[warp], lane)) { found = t; break; }
        }
      }
    }
  }
  if (lane == 0) {
    GzChunk k;
    k.start_bit = found; k.end_bit = 0; k.out_off = 0; k.out_len = 0; k.flags = 0; k.need = 0; k.land = -1;
    chunks[c] = k;
    if (found != GZ_NONE) atomicAdd(nfound, 1u);
  }
}

// ---- 2. / 4. the chunk decoder -----------------------------------------------------------------------------------------
// GZ_COUNT: decode from the chunk's start until a block ends on a later chunk's start (or the stream / the batch
// ends); nothing is written.  GZ_WRITE: decode the same blocks again, up to the recorded end, as 16-bit symbols into
// m[]: a byte, or 0x8000 | (position in the 32 KiB window before the chunk).  With `win` the window is known (the
// batch's first chunk) and bytes are taken from it directly.  GZ_BOTH: count AND write in one pass -- the symbols go
// to the chunk's own slot of an over-sized arena (`cap` symbols: up to the slot of the next chunk that found a
// start); a chunk whose output does not fit there (a stretch that compresses better than the arena allows for, a
// false start run over) says so (GZC_REWRITE) and the batch takes the second pass after all.
enum { GZ_COUNT = 0, GZ_WRITE = 1, GZ_BOTH = 2 };
template <int MODE>
__device__ __noinline__ void gz_chunk(const uint32_t* __restrict__ words, u64 nbytes, u64 chunk_bits, GzChunk* chunks, int nchunks, int c,
                                      uint16_t* m, u64 stride, const uint8_t* __restrict__ win, uint32_t wvalid, WarpTables& t, int lane,
                                      uint32_t* err) {
  constexpr bool WRITE = MODE == GZ_WRITE;   // the second pass: where to stop is known
  constexpr bool STORE = MODE != GZ_COUNT;   // symbols are written
  const bool known = c == 0;  // the batch's first chunk: the window before it is known (wvalid bytes of it exist)
  const GzChunk ck = chunks[c];
  if (ck.start_bit == GZ_NONE) return;
  if (WRITE && !(ck.flags & GZC_CHAIN)) return;
  const uint32_t* wend = words + ((nbytes + 3) >> 2);
  const u64 end_bits = nbytes * 8ull;
  const uint8_t* bytes = reinterpret_cast<const uint8_t*>(words);
  GzBits b;
  b.init(words, wend, ck.start_bit);
  int nextc = c;
  u64 limit;
  if (WRITE) {
    limit = ck.end_bit;
  } else {
    nextc = c + 1;
    while (nextc < nchunks && chunks[nextc].start_bit == GZ_NONE) nextc++;
    limit = nextc < nchunks ? chunks[nextc].start_bit : GZ_NONE;
  }
  // (w - w0) <= pos / 32 + 2 always, so b.w > wstop proves pos > limit
  const uint32_t* wstop = limit == GZ_NONE ? wend + 2 : words + (limit >> 5) + 2;
  uint32_t produced = 0, need = 0, flags = 0;
  // (GZ_BOTH) room for the chunk's symbols: its own slot and those of the chunks behind it that found no start
  const u64 cap = MODE == GZ_BOTH ? (u64)(nextc - c) * stride : ~0ull;
  bool writing = STORE;
  u64 bpos = ck.start_bit;  // the last block boundary and the bytes produced up to it
  uint32_t bout = 0;
  int passed = 0, land = -1;

  // moves the limit past `pos` (COUNT): the starts in between were not block boundaries
  auto advance_limit = [&](u64 pos) -> bool {
    while (limit != GZ_NONE && pos > limit) {
      if (++passed > GZ_MAX_PASSED) return false;
      int k = (int)(pos / chunk_bits);
      if (k <= nextc) k = nextc + 1;
      while (k < nchunks && (chunks[k].start_bit == GZ_NONE || chunks[k].start_bit < pos)) k++;
      nextc = k;
      limit = k < nchunks ? chunks[k].start_bit : GZ_NONE;
    }
    wstop = limit == GZ_NONE ? wend + 2 : words + (limit >> 5) + 2;
    return true;
  };

  for (;;) {
    const u64 pos = b.bitpos();
    if (pos > end_bits) { produced = bout; flags |= GZC_INCOMPLETE; break; }  // the block "ended" in the padding
    bpos = pos; bout = produced;
    if (WRITE) {
      if (pos == limit) break;
      if (pos > limit) { flags |= GZC_ERROR; break; }
    } else {
      if (!advance_limit(pos)) { flags |= GZC_GIVEUP; break; }
      if (pos == limit) { land = nextc; break; }
      if (pos + 10 > end_bits) { flags |= GZC_INCOMPLETE; break; }  // (the shortest block: 3 + 7 bits)
    }
    b.refill();
    const int last = (int)b.take(1);
    const int type = (int)b.take(2);
    if (type == 3) { flags |= GZC_ERROR; break; }
    if (type == 0) {  // stored: LEN, ~LEN, LEN bytes from the next byte boundary
      b.take(b.cnt & 7);
      b.refill();
      const uint32_t len = b.take(16);
      b.refill();
      const uint32_t nlen = b.take(16);
      if ((len ^ 0xFFFFu) != nlen) { flags |= GZC_ERROR; break; }
      const u64 src = b.bitpos() >> 3;
      if (src + len > nbytes) {  // runs past the batch
        if (!WRITE && advance_limit(end_bits + 1) && limit == GZ_NONE) { produced = bout; flags |= GZC_INCOMPLETE; }
        else flags |= WRITE ? GZC_ERROR : GZC_GIVEUP;
        break;
      }
      if (produced > GZ_MAX_OUT - len) { flags |= GZC_GIVEUP; break; }
      if (STORE && writing && (u64)produced + len > cap) { writing = false; flags |= GZC_REWRITE; }
      if (STORE && writing) for (uint32_t k = (uint32_t)lane; k < len; k += 32u) m[produced + k] = bytes[src + k];
      produced += len;
      b.init(words, wend, (src + len) * 8ull);
    } else {
      __syncwarp();
      if (type == 1) {  // fixed codes
        for (int s = lane; s < 288; s += 32) t.lens[s] = s < 144 ? 8 : (s < 256 ? 9 : (s < 280 ? 7 : 8));
        __syncwarp();
        huff_build<HUFF_LITLEN>(t.lcount, t.lsym, t.lit, LBITS, t.lens, 288, lane);
        for (int s = lane; s < 30; s += 32) t.lens[s] = 5;
        __syncwarp();
        huff_build<HUFF_DIST>(t.dcount, t.dsym, t.dist, DBITS, t.lens, 30, lane);
      } else {
        int nl, ndist;
        if (gz_dyn_lengths(b, t, lane, nl, ndist)) { flags |= GZC_ERROR; break; }
        if (!huff_acceptable(huff_build<HUFF_DIST>(t.dcount, t.dsym, t.dist, DBITS, t.lens + nl, ndist, lane), t.dcount, ndist) ||
            !huff_acceptable(huff_build<HUFF_LITLEN>(t.lcount, t.lsym, t.lit, LBITS, t.lens, nl, lane), t.lcount, nl)) { flags |= GZC_ERROR; break; }
      }
      bool stop = false;
      for (;;) {  // literals and matches of this block
        if (b.w > wstop) {  // ran over the limit in mid-block
          if (WRITE) { flags |= GZC_ERROR; stop = true; break; }
          if (limit == GZ_NONE) { produced = bout; flags |= GZC_INCOMPLETE; stop = true; break; }
          if (!advance_limit(b.bitpos())) { flags |= GZC_GIVEUP; stop = true; break; }
          if (b.w > wstop) continue;
        }
        const uint32_t e = huff_peek<HUFF_LITLEN, LBITS>(b, t.lit, t.lcount, t.lsym);
        if (!(e & HK_MASK)) {  // a literal
          huff_take(b, e);
          if (STORE && writing) {
            if ((u64)produced < cap) { if (lane == 0) m[produced] = (uint16_t)(e >> 16); }
            else { writing = false; flags |= GZC_REWRITE; }
          }
          produced++;
        } else if ((e & HK_MASK) != HK_MATCH) {
          if ((e & HK_MASK) == HK_INVALID) { flags |= GZC_ERROR; stop = true; }
          else huff_take(b, e);  // end of block
          break;
        } else {
          const uint32_t len = huff_take(b, e);  // (>= 33 bits were there: 15 + 5 used)
          const uint32_t de = huff_peek<HUFF_DIST, DBITS>(b, t.dist, t.dcount, t.dsym);
          if ((de & HK_MASK) == HK_INVALID) { flags |= GZC_ERROR; stop = true; break; }
          const uint32_t dist = huff_take(b, de);
          if (dist > produced) {  // reaches behind the chunk's first byte
            const uint32_t back = dist - produced;
            if (known && back > wvalid) { flags |= GZC_ERROR; stop = true; break; }  // zlib: "invalid distance too far back"
            need = back > need ? back : need;
          }
          if (produced > GZ_MAX_OUT) { flags |= GZC_GIVEUP; stop = true; break; }
          if (STORE && writing && (u64)produced + len > cap) { writing = false; flags |= GZC_REWRITE; }
          if (STORE && writing) {
            __syncwarp();  // earlier symbols (lane 0's literals, other lanes' match symbols) are visible to every lane
            for (uint32_t k = (uint32_t)lane; k < len; k += 32u) {
              const uint32_t kk = dist >= len ? k : k % dist;  // a run shorter than its length repeats with period dist
              uint16_t v;
              if (produced + kk >= dist) {
                v = m[produced + kk - dist];
              } else {
                const uint32_t idx = GZ_WINDOW - (dist - produced - kk);
                v = known ? (uint16_t)win[idx] : (uint16_t)(0x8000u | idx);
              }
              m[produced + k] = v;
            }
          }
          produced += len;
        }
      }
      if (stop) break;
    }
    if (last) {
      const u64 e = b.bitpos();
      if (e > end_bits) { produced = bout; flags |= GZC_INCOMPLETE; }
      else { bpos = e; bout = produced; flags |= GZC_FINAL; }
      break;
    }
  }
  __syncwarp();
  if (WRITE) {
    if (lane == 0 && ((flags & (GZC_ERROR | GZC_GIVEUP)) || bout != ck.out_len || bpos != ck.end_bit)) atomicOr(err, 1u);
  } else if (lane == 0) {
    GzChunk& o = chunks[c];
    o.end_bit = bpos; o.out_len = bout; o.flags = flags; o.need = need; o.land = land;
  }
}

__global__ void __launch_bounds__(32 * GZ_WARPS, GZ_MINB) gz_count_kernel(const uint32_t* __restrict__ words, u64 nbytes, u64 chunk_bits, GzChunk* chunks,
                                                                 int nchunks, uint32_t wvalid) {
  __shared__ WarpTables tables[GZ_WARPS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * GZ_WARPS + warp;
  if (c >= nchunks) return;
  gz_chunk<GZ_COUNT>(words, nbytes, chunk_bits, chunks, nchunks, c, nullptr, 0, nullptr, wvalid, tables[warp], lane, nullptr);
}

__global__ void __launch_bounds__(32 * GZ_WARPS, GZ_MINB) gz_write_kernel(const uint32_t* __restrict__ words, u64 nbytes, u64 chunk_bits, GzChunk* chunks,
                                                                 int nchunks, uint16_t* markers, const uint8_t* __restrict__ window, uint32_t wvalid,
                                                                 uint32_t* err) {
  __shared__ WarpTables tables[GZ_WARPS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * GZ_WARPS + warp;
  if (c >= nchunks) return;
  gz_chunk<GZ_WRITE>(words, nbytes, chunk_bits, chunks, nchunks, c, markers + chunks[c].out_off, 0, window, wvalid, tables[warp], lane, err);
}

__global__ void __launch_bounds__(32 * GZ_WARPS, GZ_MINB) gz_both_kernel(const uint32_t* __restrict__ words, u64 nbytes, u64 chunk_bits, GzChunk* chunks,
                                                                        int nchunks, uint16_t* arena, u64 stride, const uint8_t* __restrict__ window,
                                                                        uint32_t wvalid) {
  __shared__ WarpTables tables[GZ_WARPS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * GZ_WARPS + warp;
  if (c >= nchunks) return;
  gz_chunk<GZ_BOTH>(words, nbytes, chunk_bits, chunks, nchunks, c, arena + (u64)c * stride, stride, window, wvalid, tables[warp], lane, nullptr);
}

// ---- 3. the chain ------------------------------------------------------------------------------------------------------
// One CTA.  Thread 0 follows the landings through shared memory; then all threads give the chain's chunks their
// output offsets (an exclusive scan) and check them.
constexpr int CHAIN_THREADS = 1024;
__global__ void __launch_bounds__(CHAIN_THREADS) gz_chain_kernel(GzChunk* chunks, int nchunks, u64 prior_out, GzResult* res, uint32_t* order, u64* coff,
                                                                 u64 stride, u64* sbase) {
  extern __shared__ int land_s[];  // nchunks landings
  __shared__ uint32_t nchain_s, bad_s, rewrite_s;
  __shared__ u64 wsum[CHAIN_THREADS / 32], carry_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int c = tid; c < nchunks; c += CHAIN_THREADS) land_s[c] = chunks[c].land;
  if (tid == 0) { bad_s = 0; carry_s = 0; rewrite_s = 0; }
  __syncthreads();
  if (tid == 0) {
    uint32_t n = 0;
    int c = 0;
    while (c >= 0 && n < (uint32_t)nchunks) { order[n++] = (uint32_t)c; const int nx = land_s[c]; c = nx > c ? nx : -1; }
    nchain_s = n;
  }
  __syncthreads();
  const uint32_t nchain = nchain_s;
  uint32_t passed = 0;
  for (uint32_t i0 = 0; i0 < nchain; i0 += CHAIN_THREADS) {
    const uint32_t i = i0 + (uint32_t)tid;
    u64 len = 0;
    uint32_t flags = 0, need = 0;
    int c = -1;
    if (i < nchain) {
      c = (int)order[i];
      len = chunks[c].out_len; flags = chunks[c].flags; need = chunks[c].need;
      // only the chain's last chunk may lack a landing
      if ((flags & (GZC_ERROR | GZC_GIVEUP)) || (i + 1 < nchain ? (flags & (GZC_FINAL | GZC_INCOMPLETE)) != 0 : !(flags & (GZC_FINAL | GZC_INCOMPLETE))))
        atomicOr(&bad_s, 1u);
      if (i + 1 < nchain) {  // chunks with a start between this one and its landing: starts that were run over
        for (int k = c + 1; k < (int)order[i + 1]; k++) passed += chunks[k].start_bit != GZ_NONE;
      }
    }
    u64 incl = len;
    for (int d = 1; d < 32; d <<= 1) { const u64 nb = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += nb; }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    u64 base = carry_s;
    for (int w = 0; w < warp; w++) base += wsum[w];
    const u64 off = base + incl - len;
    if (i < nchain) {
      chunks[c].out_off = off;
      chunks[c].flags = flags | GZC_CHAIN;
      coff[i] = off;
      if (stride) sbase[i] = (u64)c * stride - off;  // (wraps when off is the larger: added back modulo 2^64)
      if (flags & GZC_REWRITE) atomicOr(&rewrite_s, 1u);
      if (i > 0 && (u64)need > off + prior_out) atomicOr(&bad_s, 1u);  // a match reaches behind the first byte of the stream
      if (i + 1 == nchain) {
        coff[nchain] = off + len;
        res->total_out = off + len; res->end_bit = chunks[c].end_bit; res->final_block = (flags & GZC_FINAL) ? 1u : 0u;
      }
    }
    __syncthreads();
    if (tid == CHAIN_THREADS - 1) carry_s = base + incl;
    __syncthreads();
  }
  passed = __reduce_add_sync(0xffffffffu, passed);
  __shared__ uint32_t passed_s;
  if (tid == 0) passed_s = 0;
  __syncthreads();
  if (lane == 0 && passed) atomicAdd(&passed_s, passed);
  __syncthreads();
  if (tid == 0) { res->status = bad_s ? GZR_BROKEN : GZR_OK; res->nchain = nchain; res->passed = passed_s; res->rewrite = rewrite_s; res->pad = 0; }
}

// ---- 5. windows ----------------------------------------------------------------------------------------------------------
// The 32 KiB before chain chunk i+1 are the tail of (the 32 KiB before chunk i) ++ (chunk i's symbols resolved through
// those).  That recurrence is the one serial step of the whole scheme, and it does not thin out: in FASTQ a fifth of a
// chunk's last symbols are still markers (every header is a match on the header before it, back to the chunk's first
// record).  But a window whose entries are "a byte, or a position in some earlier window" composes: the chain is cut
// into GROUPS of K chunks, and
//   gz_rows_kernel<true>   one CTA per group, all groups at once, walks its K chunks starting from the identity window
//                          (entry j = "position j of the window before the group"): symrows[i] = the window before
//                          chunk i as 16-bit symbols relative to the group's window; grows[g] = the same for the window
//                          behind the group;
//   gz_rows_kernel<false>  one CTA walks the GROUPS: trows[g + 1] = grows[g] resolved through trows[g], in bytes.
// K steps in parallel plus nchain / K steps in series instead of nchain.  One step, for either kernel: the row before in
// shared memory, the symbols arriving by bulk copies (TMA) in halves of 16 Ki while the half before is resolved, the new
// row leaving by a bulk store.
constexpr int WIN_THREADS = 1024;
constexpr uint32_t WIN_HALF = GZ_WINDOW / 2;
constexpr uint32_t WIN_STAGE = WIN_HALF + 16;  // symbols per staging buffer (half a window, from a 16-byte boundary)
template <typename T>
constexpr size_t win_smem_bytes() { return 2 * GZ_WINDOW * sizeof(T) + 2 * WIN_STAGE * sizeof(uint16_t) + 16; }

// SYMBOLIC: CTA g walks chain chunks [g K, min((g + 1) K, nchain)); its symbols are the chunks' tails in `markers`.
// otherwise: one CTA walks the groups; the symbols of step g are the row grows[g].
template <bool SYMBOLIC>
__global__ void __launch_bounds__(WIN_THREADS) gz_rows_kernel(const u64* __restrict__ coff, const u64* __restrict__ sbase, uint32_t nchain, uint32_t K,
                                                              uint32_t ngroups, const uint16_t* __restrict__ markers, uint16_t* __restrict__ symrows,
                                                              uint16_t* __restrict__ grows, uint8_t* __restrict__ trows) {
  typedef typename std::conditional<SYMBOLIC, uint16_t, uint8_t>::type T;
  extern __shared__ __align__(128) uint8_t win_smem[];
  T* row_s = reinterpret_cast<T*>(win_smem);                                                   // 2 rows
  uint16_t* stage_s = reinterpret_cast<uint16_t*>(win_smem + 2 * GZ_WINDOW * sizeof(T));     // 2 staging buffers
  const uint32_t bar_s = smem_u32(win_smem + 2 * GZ_WINDOW * sizeof(T) + 2 * WIN_STAGE * sizeof(uint16_t));
  const int tid = threadIdx.x;
  const uint32_t first = SYMBOLIC ? blockIdx.x * K : 0u;
  const uint32_t nsteps = SYMBOLIC ? (first + K <= nchain ? K : nchain - first) : ngroups;

  // step k, half h: the symbols of window positions [h * 16Ki, (h + 1) * 16Ki) that come from the step's own output
  // (the others are still inside the row before).  `off`, `end` = the output range of the step's chunk.
  auto src_of = [&](uint32_t k, u64 off, u64 end, uint32_t h, u64& lo, uint32_t& keep) -> const uint16_t* {
    const u64 len = end - off;
    keep = len >= GZ_WINDOW ? 0u : GZ_WINDOW - (uint32_t)len;  // positions of the new window still inside the old one
    const uint32_t j0 = keep > h * WIN_HALF ? keep : h * WIN_HALF;  // the half's first position that is a symbol of this step
    lo = end - GZ_WINDOW + j0;                                      // ... and that symbol's index
    return SYMBOLIC ? markers : grows + (size_t)k * GZ_WINDOW;
  };
  auto fetch = [&](uint32_t hs, u64 off, u64 end) {  // half-step hs = 2 k + h
    u64 lo; uint32_t keep;
    const uint32_t k = hs >> 1, h = hs & 1u;
    const uint16_t* src = src_of(k, off, end, h, lo, keep);
    const u64 a = lo & ~7ull;
    const bool any = (keep > h * WIN_HALF ? keep : h * WIN_HALF) < (h + 1) * WIN_HALF;  // does any position of this half come from the step?
    const u64 hi = any ? end - GZ_WINDOW + (u64)(h + 1) * WIN_HALF : 0;                 // (one past the half's last symbol)
    // nothing from the step: a dummy copy all the same, so that every half-step has one and the phases stay in step
    const uint32_t bytes = any ? (uint32_t)(((hi - a) * 2 + 15) & ~15ull) : 16u;
    mbar_expect_tx(bar_s + 8u * (hs & 1u), bytes);
    tma_load_1d(smem_u32(stage_s + (size_t)(hs & 1u) * WIN_STAGE), src + (any ? a : 0), bytes, bar_s + 8u * (hs & 1u));
  };
  // the output range of step k: a chunk of the chain, or (group walk) a whole row that replaces the window
  auto range_of = [&](uint32_t k, u64& off, u64& end) {
    if (SYMBOLIC) {  // (as indices into `markers`: output positions, or -- one-pass mode -- positions in the arena)
      const u64 sh = sbase ? sbase[first + k] : 0ull;
      off = coff[first + k] + sh; end = coff[first + k + 1] + sh;
    } else { off = 0; end = GZ_WINDOW; }
  };

  if (SYMBOLIC) {
    for (uint32_t j = (uint32_t)tid; j < GZ_WINDOW; j += WIN_THREADS) row_s[j] = (T)(0x8000u | j);  // the identity window
  } else {
    for (int j = tid; j < (int)GZ_WINDOW / 16; j += WIN_THREADS)
      reinterpret_cast<uint4*>(row_s)[j] = reinterpret_cast<const uint4*>(trows)[j];  // row 0: the window before the batch
  }
  if (tid == 0) {
    mbar_init(bar_s, 1);
    mbar_init(bar_s + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (nsteps == 0) return;
  u64 off, end, off1 = 0, end1 = 0;  // this step's range and the next one's
  range_of(0, off, end);
  if (nsteps > 1) range_of(1, off1, end1);
  if (tid == 0) { fetch(0, off, end); fetch(1, off, end); }
  for (uint32_t k = 0; k < nsteps; k++) {
    u64 off2 = 0, end2 = 0;
    if (k + 2 < nsteps) range_of(k + 2, off2, end2);  // (two steps ahead: nothing in this step waits for it)
    const T* prev = row_s + (size_t)(k & 1u) * GZ_WINDOW;
    T* cur = row_s + (size_t)((k + 1) & 1u) * GZ_WINDOW;
    for (uint32_t h = 0; h < 2; h++) {
      const uint32_t hs = 2 * k + h;
      u64 lo; uint32_t keep;
      src_of(k, off, end, h, lo, keep);
      const uint16_t* sy = stage_s + (size_t)(hs & 1u) * WIN_STAGE + (lo & 7ull);  // sy[q] = symbol lo + q
      const uint32_t j0 = keep > h * WIN_HALF ? keep : h * WIN_HALF;
      mbar_wait(bar_s + 8u * (hs & 1u), (hs >> 1) & 1u);
      // position j of the new window: a symbol of this step, or -- in front of a chunk shorter than the window -- position
      // j + len of the old window, which reads like a marker
      uint32_t s2[WIN_HALF / WIN_THREADS];
#pragma unroll
      for (int r = 0; r < (int)(WIN_HALF / WIN_THREADS); r++) {
        const uint32_t j = h * WIN_HALF + (uint32_t)tid + (uint32_t)r * WIN_THREADS;
        s2[r] = j >= keep ? (uint32_t)sy[j - j0] : 0x8000u | (j + (GZ_WINDOW - keep));
      }
#pragma unroll
      for (int r = 0; r < (int)(WIN_HALF / WIN_THREADS); r++)
        if (s2[r] >= 256u) s2[r] = prev[s2[r] & 0x7FFFu];
#pragma unroll
      for (int r = 0; r < (int)(WIN_HALF / WIN_THREADS); r++) cur[h * WIN_HALF + (uint32_t)tid + (uint32_t)r * WIN_THREADS] = (T)s2[r];
      if (h == 1 && tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the row stored a step ago has left shared memory
      __syncthreads();
      if (tid == 0) {
        if (h == 1) {  // the row is complete: out it goes
          void* dst;  // the window before the next chunk of the group / behind the group / before the next group
          if (SYMBOLIC) dst = k + 1 < nsteps ? symrows + (size_t)(first + k + 1) * GZ_WINDOW : grows + (size_t)blockIdx.x * GZ_WINDOW;
          else dst = trows + (size_t)(k + 1) * GZ_WINDOW;
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(cur)),
                       "r"((uint32_t)(GZ_WINDOW * sizeof(T))) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        // the staging buffer just read is free: the half-step after the next one goes there
        if (k + 1 < nsteps) fetch(hs + 2, off1, end1);
      }
    }
    off = off1; end = end1; off1 = off2; end1 = end2;
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ---- 6. resolve ------------------------------------------------------------------------------------------------------------
constexpr int RES_THREADS = 256;
constexpr uint32_t RES_SLICE = RES_THREADS * 8 * 8;  // output positions per CTA step
template <bool ARENA>
__global__ void __launch_bounds__(RES_THREADS) gz_resolve_kernel(const uint16_t* __restrict__ markers, const u64* __restrict__ sbase,
                                                                 const uint16_t* __restrict__ symrows, const uint8_t* __restrict__ trows, uint32_t K,
                                                                 const u64* __restrict__ coff, uint32_t nchain, u64 total, uint8_t* __restrict__ out) {
  __shared__ uint32_t first_s;
  const u64 nslices = (total + RES_SLICE - 1) / RES_SLICE;
  for (u64 sl = blockIdx.x; sl < nslices; sl += gridDim.x) {
    const u64 s0 = sl * RES_SLICE;
    __syncthreads();
    if (threadIdx.x == 0) {  // the chain chunk holding the slice's first position: the last i with coff[i] <= s0
      uint32_t lo = 0, hi = nchain;
      while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (coff[mid] <= s0) lo = mid; else hi = mid; }
      first_s = lo;
    }
    __syncthreads();
    uint32_t i = first_s;
    u64 next = coff[i + 1];
    u64 sh = ARENA ? sbase[i] : 0ull;  // (one-pass mode: the chunk's symbols are in its slot of the arena)
#pragma unroll 2
    for (uint32_t r = 0; r < 8; r++) {
      const u64 p0 = s0 + ((u64)r * RES_THREADS + threadIdx.x) * 8;
      if (p0 >= total) break;
      uint32_t s[8];
      if (!ARENA) {
        const uint4 q = *reinterpret_cast<const uint4*>(markers + p0);  // (the buffer is padded to a multiple of 8 symbols)
        s[0] = q.x & 0xFFFFu; s[1] = q.x >> 16; s[2] = q.y & 0xFFFFu; s[3] = q.y >> 16;
        s[4] = q.z & 0xFFFFu; s[5] = q.z >> 16; s[6] = q.w & 0xFFFFu; s[7] = q.w >> 16;
      }
      u64 packed = 0;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const u64 p = p0 + (u64)k;
        while (p >= next && i + 1 < nchain) { i++; next = coff[i + 1]; if (ARENA) sh = sbase[i]; }
        uint32_t v = ARENA ? (p < total ? (uint32_t)markers[sh + p] : 0u) : s[k];
        if (v >= 256u && p < total) {
          // through the chunk's window relative to its group's, then through the group's window
          if (i % K) v = symrows[(size_t)i * GZ_WINDOW + (v & 0x7FFFu)];
          if (v >= 256u) v = trows[(size_t)(i / K) * GZ_WINDOW + (v & 0x7FFFu)];
        }
        packed |= (u64)(v & 0xFFu) << (8 * k);
      }
      *reinterpret_cast<u64*>(out + p0) = packed;  // (padded likewise)
    }
  }
}

// ---- 7. CRC-32 of the inflated bytes --------------------------------------------------------------------------------------
// gzread() checks the member's CRC-32; so does this path, which makes zlib's own criterion the last word on
// everything above (a wrong chain or a wrong window cannot produce the right CRC).  The register is linear in the
// data: for a register that starts at zero, raw(A ++ B) = raw(A) * x^(8|B|) + raw(B) over GF(2) modulo the CRC
// polynomial.  Slices of 4 KiB are summed by one thread each (table in shared memory), then folded in order:
// 1024 threads Horner-fold equal runs of slices, a tree joins the runs.  The host adds the tail, the bytes of the
// batches before, and the initial all-ones register (crc_* in fqgpu_api.cu).
constexpr uint32_t CRC_POLY = 0xEDB88320u;
constexpr uint32_t CRC_SLICE_BYTES = 4096;
constexpr int CRCA_THREADS = 128;
constexpr int CRCB_THREADS = 1024;

// a * b modulo the CRC polynomial; bit 31 is the coefficient of x^0 (the reflected form the register is in)
__device__ __forceinline__ uint32_t gf_mul(uint32_t a, uint32_t b) {
  uint32_t p = 0;
#pragma unroll
  for (int i = 0; i < 32; i++) {
    p ^= (a & (0x80000000u >> i)) ? b : 0u;
    b = (b >> 1) ^ ((b & 1u) ? CRC_POLY : 0u);
  }
  return p;
}

// raw[k] = the register after slice k when it starts at zero; raw[nfull] = the same for the bytes behind the last
// full slice
__global__ void __launch_bounds__(CRCA_THREADS) gz_crc_slices_kernel(const uint8_t* __restrict__ data, u64 total, uint32_t* __restrict__ raw) {
  __shared__ uint32_t tab[256];
  for (int t = threadIdx.x; t < 256; t += CRCA_THREADS) {
    uint32_t c = (uint32_t)t;
    for (int k = 0; k < 8; k++) c = (c & 1u) ? CRC_POLY ^ (c >> 1) : c >> 1;
    tab[t] = c;
  }
  __syncthreads();
  const u64 nfull = total / CRC_SLICE_BYTES;
  const u64 k = (u64)blockIdx.x * CRCA_THREADS + threadIdx.x;
  if (k > nfull) return;
  uint32_t v = 0;
  if (k < nfull) {
    const uint4* p = reinterpret_cast<const uint4*>(data + k * CRC_SLICE_BYTES);
    for (uint32_t q = 0; q < CRC_SLICE_BYTES / 16; q++) {
      const uint4 w = __ldg(p + q);
      const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int e = 0; e < 4; e++) {
        v ^= ws[e];  // four bytes at once into the low end of the register, then four table steps
#pragma unroll
        for (int b = 0; b < 4; b++) v = tab[v & 0xFFu] ^ (v >> 8);
      }
    }
  } else {
    for (u64 q = nfull * CRC_SLICE_BYTES; q < total; q++) v = tab[(v ^ data[q]) & 0xFFu] ^ (v >> 8);
  }
  raw[k] = v;
}

// the nfull slice registers folded in order into one: thread t folds slices [t q - pad, (t + 1) q - pad) (q = slices per
// thread, the slices before the first are zeros and change nothing), then the 1024 runs are joined pairwise.
// xs = x^(8 * 4096), xq = x^(8 * 4096 * q).
__global__ void __launch_bounds__(CRCB_THREADS) gz_crc_fold_kernel(const uint32_t* __restrict__ raw, u64 nfull, uint32_t q, uint32_t xs, uint32_t xq,
                                                                   uint32_t* out) {
  __shared__ uint32_t part[CRCB_THREADS];
  const int tid = threadIdx.x;
  const u64 pad = (u64)CRCB_THREADS * q - nfull;
  uint32_t v = 0;
  for (uint32_t j = 0; j < q; j++) {
    const u64 virt = (u64)tid * q + j;
    if (virt >= pad) v = gf_mul(xs, v) ^ raw[virt - pad];
  }
  part[tid] = v;
  uint32_t op = xq;  // x^(8 * bytes of the right-hand run)
  for (int d = 1; d < CRCB_THREADS; d <<= 1) {
    __syncthreads();
    uint32_t joined = 0;
    const bool mine = (tid & (2 * d - 1)) == 0;
    if (mine) joined = gf_mul(op, part[tid]) ^ part[tid + d];
    __syncthreads();
    if (mine) part[tid] = joined;
    op = gf_mul(op, op);
  }
  if (tid == 0) { out[0] = part[0]; out[1] = raw[nfull]; }
}

// ---- launchers ---------------------------------------------------------------------------------------------------------------
// once per device (fqgpu_create): the kernels that take more than 48 KiB of dynamic shared memory
cudaError_t gz_configure() {
  cudaError_t e = cudaFuncSetAttribute(gz_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GZ_MAX_CHUNKS * (int)sizeof(int));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(gz_rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)win_smem_bytes<uint16_t>());
  if (e == cudaSuccess) e = cudaFuncSetAttribute(gz_rows_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)win_smem_bytes<uint8_t>());
  return e;
}
cudaError_t launch_gz_sync(const uint8_t* d_comp, size_t nbytes, uint32_t chunk_bytes, int nchunks, unsigned long long start_bit,
                           GzChunk* chunks, uint32_t* nfound, cudaStream_t st) {
  gz_sync_kernel<<<(nchunks + GZ_WARPS - 1) / GZ_WARPS, 32 * GZ_WARPS, 0, st>>>(reinterpret_cast<const uint32_t*>(d_comp), (u64)nbytes, chunk_bytes,
                                                                                nchunks, start_bit, chunks, nfound);
  return cudaGetLastError();
}
cudaError_t launch_gz_count(const uint8_t* d_comp, size_t nbytes, uint32_t chunk_bytes, GzChunk* chunks, int nchunks, uint32_t wvalid,
                            cudaStream_t st) {
  gz_count_kernel<<<(nchunks + GZ_WARPS - 1) / GZ_WARPS, 32 * GZ_WARPS, 0, st>>>(reinterpret_cast<const uint32_t*>(d_comp), (u64)nbytes,
                                                                                 (u64)chunk_bytes * 8ull, chunks, nchunks, wvalid);
  return cudaGetLastError();
}
cudaError_t launch_gz_both(const uint8_t* d_comp, size_t nbytes, uint32_t chunk_bytes, GzChunk* chunks, int nchunks, uint16_t* arena,
                           unsigned long long stride, const uint8_t* window, uint32_t wvalid, cudaStream_t st) {
  gz_both_kernel<<<(nchunks + GZ_WARPS - 1) / GZ_WARPS, 32 * GZ_WARPS, 0, st>>>(reinterpret_cast<const uint32_t*>(d_comp), (u64)nbytes,
                                                                                (u64)chunk_bytes * 8ull, chunks, nchunks, arena, stride, window, wvalid);
  return cudaGetLastError();
}
cudaError_t launch_gz_chain(GzChunk* chunks, int nchunks, unsigned long long prior_out, GzResult* res, uint32_t* order,
                            unsigned long long* coff, unsigned long long stride, unsigned long long* sbase, cudaStream_t st) {
  if (nchunks > GZ_MAX_CHUNKS) return cudaErrorInvalidValue;
  gz_chain_kernel<<<1, CHAIN_THREADS, (size_t)nchunks * sizeof(int), st>>>(chunks, nchunks, prior_out, res, order, coff, stride, sbase);
  return cudaGetLastError();
}
cudaError_t launch_gz_write(const uint8_t* d_comp, size_t nbytes, uint32_t chunk_bytes, GzChunk* chunks, int nchunks, uint16_t* markers,
                            const uint8_t* window, uint32_t wvalid, uint32_t* err, cudaStream_t st) {
  gz_write_kernel<<<(nchunks + GZ_WARPS - 1) / GZ_WARPS, 32 * GZ_WARPS, 0, st>>>(reinterpret_cast<const uint32_t*>(d_comp), (u64)nbytes,
                                                                                 (u64)chunk_bytes * 8ull, chunks, nchunks, markers, window, wvalid, err);
  return cudaGetLastError();
}
uint32_t gz_group_chunks(uint32_t nchain, int sms) {
  const uint32_t k = (nchain + (uint32_t)sms - 1) / (uint32_t)sms;
  return k < 8 ? 8 : k;
}
cudaError_t launch_gz_windows(const unsigned long long* coff, const unsigned long long* sbase, uint32_t nchain, uint32_t K, const uint16_t* markers,
                              uint16_t* symrows, uint16_t* grows, uint8_t* trows, uint8_t* window, cudaStream_t st) {
  if (nchain == 0 || K == 0) return cudaErrorInvalidValue;
  const uint32_t ngroups = (nchain + K - 1) / K;
  cudaError_t e = cudaMemcpyAsync(trows, window, GZ_WINDOW, cudaMemcpyDeviceToDevice, st);  // row 0: the window before the batch
  if (e != cudaSuccess) return e;
  gz_rows_kernel<true><<<ngroups, WIN_THREADS, win_smem_bytes<uint16_t>(), st>>>(coff, sbase, nchain, K, ngroups, markers, symrows, grows, trows);
  gz_rows_kernel<false><<<1, WIN_THREADS, win_smem_bytes<uint8_t>(), st>>>(coff, sbase, nchain, K, ngroups, markers, symrows, grows, trows);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  return cudaMemcpyAsync(window, trows + (size_t)ngroups * GZ_WINDOW, GZ_WINDOW, cudaMemcpyDeviceToDevice, st);  // the window behind the batch
}
cudaError_t launch_gz_resolve(const uint16_t* markers, const unsigned long long* sbase, const uint16_t* symrows, const uint8_t* trows, uint32_t K,
                              const unsigned long long* coff, uint32_t nchain, unsigned long long total_out, uint8_t* out, int sms, cudaStream_t st) {
  if (total_out == 0) return cudaSuccess;
  const u64 want = (total_out + RES_SLICE - 1) / RES_SLICE;
  const int grid = (int)(want < (u64)sms * 8 ? (want ? want : 1) : (u64)sms * 8);
  if (sbase) gz_resolve_kernel<true><<<grid, RES_THREADS, 0, st>>>(markers, sbase, symrows, trows, K, coff, nchain, total_out, out);
  else gz_resolve_kernel<false><<<grid, RES_THREADS, 0, st>>>(markers, sbase, symrows, trows, K, coff, nchain, total_out, out);
  return cudaGetLastError();
}
cudaError_t launch_gz_crc(const uint8_t* d_out, unsigned long long total, uint32_t* d_raw, uint32_t xs, uint32_t xq, uint32_t q, uint32_t* d_crc2,
                          cudaStream_t st) {
  const u64 nfull = total / CRC_SLICE_BYTES;
  gz_crc_slices_kernel<<<(unsigned)((nfull + 1 + CRCA_THREADS - 1) / CRCA_THREADS), CRCA_THREADS, 0, st>>>(d_out, total, d_raw);
  gz_crc_fold_kernel<<<1, CRCB_THREADS, 0, st>>>(d_raw, nfull, q, xs, xq, d_crc2);
  return cudaGetLastError();
}

}  // namespace fq
