// fq_bgzf.cu -- on-device inflate of BGZF input (blocked gzip: bgzip, htslib, many sequencer pipelines) feeding
// the scan kernels.
//
// The reference inflates .gz input on the host, one byte stream through zlib (src/utils/gzip_stream.nim:16-17,
// src/fq_count.nim:32); a single gzip member cannot be inflated in parallel.  A BGZF file is a concatenation of
// gzip members of at most 64 KiB, each with its compressed size in the header's "BC" extra field and its
// uncompressed size in the trailer, so every member is an independent DEFLATE stream with a known output
// offset: the host only walks the member headers, the compressed bytes go to the GPU as they are, and ONE WARP
// PER MEMBER (lane 0 decodes: a DEFLATE stream is serial, and 32 members sharing a warp would run one after the
// other -- measured 1.01 active threads per instruction -- so each member gets its own warp and the machine hides
// the latency of ~10 k such warps behind each other) inflates its stream (RFC 1951: stored, fixed and dynamic
// Huffman blocks; canonical-code decoding without lookup tables, the code lengths kept in local memory) straight to
// its place in the output buffer, which fqgpu_scan_device then scans like any HBM-resident input.  Anything that is not well-formed BGZF makes the caller fall back to the zlib path.
#include <cuda_runtime.h>
#include <stdint.h>

#include "fq_bgzf.h"

namespace fq {

static __device__ const uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59,
                                                 67, 83, 99, 115, 131, 163, 195, 227, 258};
static __device__ const uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static __device__ const uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769,
                                                  1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static __device__ const uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
static __device__ const uint8_t kClOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

// LSB-first bit reader over [p, end).  cnt goes negative when the stream is read past its end (checked by the callers).
struct Bits {
  const uint8_t* p;
  const uint8_t* end;
  unsigned long long buf;
  int cnt;
  __device__ __forceinline__ void refill() {
    while (cnt <= 56 && p < end) { buf |= (unsigned long long)__ldg(p++) << cnt; cnt += 8; }
  }
  __device__ __forceinline__ uint32_t take(int n) {  // n <= 16; call refill() first
    const uint32_t v = (uint32_t)buf & ((1u << n) - 1u);
    buf >>= n; cnt -= n;
    return v;
  }
};

// Canonical Huffman code from code lengths: count[l] = codes of length l, symbol[] = symbols ordered by code.
// Returns < 0 for an over-subscribed set of lengths, > 0 for an incomplete one, 0 for a complete one.
__device__ int huff_build(uint16_t* count, uint16_t* symbol, const uint8_t* len, int n) {
  for (int l = 0; l <= 15; l++) count[l] = 0;
  for (int s = 0; s < n; s++) count[len[s]]++;
  if (count[0] == n) return 0;
  int left = 1;
  for (int l = 1; l <= 15; l++) { left <<= 1; left -= count[l]; if (left < 0) return left; }
  uint16_t offs[16];
  offs[1] = 0;
  for (int l = 1; l < 15; l++) offs[l + 1] = offs[l] + count[l];
  for (int s = 0; s < n; s++) if (len[s]) symbol[offs[len[s]]++] = (uint16_t)s;
  return left;
}
// Next symbol: walks the code one bit at a time on a copy of the bit buffer.
__device__ __forceinline__ int huff_decode(Bits& b, const uint16_t* count, const uint16_t* symbol) {
  b.refill();
  unsigned long long bb = b.buf;
  int code = 0, first = 0, index = 0;
  for (int l = 1; l <= 15; l++) {
    code |= (int)(bb & 1ull);
    bb >>= 1;
    const int c = count[l];
    if (code - c < first) { b.buf = bb; b.cnt -= l; return symbol[index + (code - first)]; }
    index += c; first += c;
    first <<= 1; code <<= 1;
  }
  return -1;
}

enum { BGZF_OK = 0, BGZF_EBLOCK = 1, BGZF_ECODE = 2, BGZF_EDIST = 3, BGZF_ESIZE = 4, BGZF_ETRUNC = 5, BGZF_ECRC = 6 };

// The serial part: one thread inflates member d.  Returns BGZF_*.
__device__ __noinline__ int bgzf_inflate_member(const uint8_t* __restrict__ comp, const BgzfMember& d, uint8_t* out) {
  Bits b;
  b.p = comp + d.in_off; b.end = b.p + d.in_len; b.buf = 0; b.cnt = 0;
  uint8_t* o = out + d.out_off;
  const uint32_t cap = d.out_len;
  uint32_t produced = 0;
  uint16_t lcount[16], lsym[288], dcount[16], dsym[32];
  uint8_t lens[320];
  int err = BGZF_OK, last = 0;
  do {
    b.refill();
    last = (int)b.take(1);
    const int type = (int)b.take(2);
    if (type == 0) {  // stored
      b.take(b.cnt & 7);
      b.refill();
      const uint32_t len = b.take(16), nlen = b.take(16);
      if (b.cnt < 0 || (len ^ 0xFFFFu) != nlen) { err = BGZF_EBLOCK; break; }
      if (produced + len > cap) { err = BGZF_ESIZE; break; }
      for (uint32_t k = 0; k < len; k++) { b.refill(); o[produced++] = (uint8_t)b.take(8); }
      if (b.cnt < 0) { err = BGZF_ETRUNC; break; }
      continue;
    }
    if (type == 3) { err = BGZF_EBLOCK; break; }
    if (type == 1) {  // fixed codes
      for (int s = 0; s < 144; s++) lens[s] = 8;
      for (int s = 144; s < 256; s++) lens[s] = 9;
      for (int s = 256; s < 280; s++) lens[s] = 7;
      for (int s = 280; s < 288; s++) lens[s] = 8;
      huff_build(lcount, lsym, lens, 288);
      for (int s = 0; s < 30; s++) lens[s] = 5;
      huff_build(dcount, dsym, lens, 30);
    } else {  // dynamic codes
      b.refill();
      const int nlen = (int)b.take(5) + 257, ndist = (int)b.take(5) + 1, ncode = (int)b.take(4) + 4;
      if (nlen > 286 || ndist > 30) { err = BGZF_ECODE; break; }
      for (int k = 0; k < 19; k++) lens[k] = 0;
      for (int k = 0; k < ncode; k++) { b.refill(); lens[kClOrder[k]] = (uint8_t)b.take(3); }
      if (huff_build(lcount, lsym, lens, 19) != 0) { err = BGZF_ECODE; break; }  // the code-length code must be complete
      int idx = 0;
      while (idx < nlen + ndist && !err) {
        int sym = huff_decode(b, lcount, lsym);
        if (sym < 0 || b.cnt < 0) { err = BGZF_ECODE; break; }
        if (sym < 16) { lens[idx++] = (uint8_t)sym; continue; }
        int prev = 0, rep;
        b.refill();
        if (sym == 16) { if (idx == 0) { err = BGZF_ECODE; break; } prev = lens[idx - 1]; rep = 3 + (int)b.take(2); }
        else if (sym == 17) rep = 3 + (int)b.take(3);
        else rep = 11 + (int)b.take(7);
        if (idx + rep > nlen + ndist) { err = BGZF_ECODE; break; }
        while (rep--) lens[idx++] = (uint8_t)prev;
      }
      if (err) break;
      if (lens[256] == 0) { err = BGZF_ECODE; break; }
      // (the distance lengths follow the literal/length lengths in lens[]; build the distance code first: the
      //  literal/length build overwrites nothing it needs)
      if (huff_build(dcount, dsym, lens + nlen, ndist) < 0) { err = BGZF_ECODE; break; }
      if (huff_build(lcount, lsym, lens, nlen) < 0) { err = BGZF_ECODE; break; }
    }
    for (;;) {  // literals and matches of this DEFLATE block
      int sym = huff_decode(b, lcount, lsym);
      if (sym < 0 || b.cnt < 0) { err = BGZF_ECODE; break; }
      if (sym < 256) {
        if (produced >= cap) { err = BGZF_ESIZE; break; }
        o[produced++] = (uint8_t)sym;
      } else if (sym == 256) {
        break;
      } else {
        sym -= 257;
        if (sym >= 29) { err = BGZF_ECODE; break; }
        b.refill();
        const uint32_t len = kLenBase[sym] + b.take(kLenExtra[sym]);
        const int ds = huff_decode(b, dcount, dsym);
        if (ds < 0 || ds >= 30 || b.cnt < 0) { err = BGZF_EDIST; break; }
        b.refill();
        const uint32_t dist = kDistBase[ds] + b.take(kDistExtra[ds]);
        if (dist > produced) { err = BGZF_EDIST; break; }
        if (produced + len > cap) { err = BGZF_ESIZE; break; }
        const uint8_t* src = o + produced - dist;
        for (uint32_t k = 0; k < len; k++) o[produced + k] = src[k];
        produced += len;
      }
    }
  } while (!last && !err);
  if (!err && b.cnt < 0) err = BGZF_ETRUNC;
  if (!err && produced != cap) err = BGZF_ESIZE;
  return err;
}

constexpr int BGZF_WARPS = 4;        // members per CTA
constexpr uint32_t CRC_SLICE = 2048;  // bytes per lane in the CRC pass (32 slices cover a 64 KiB member)

// crc_shift[k] = the CRC register after CRC_SLICE zero bytes when it starts as 1 << k (a linear map over GF(2)).
struct CrcShift { uint32_t col[32]; };

__device__ __forceinline__ uint32_t crc_bytes(const uint32_t* tab, const uint8_t* p, uint32_t n, uint32_t v) {
  for (uint32_t k = 0; k < n; k++) v = tab[(v ^ p[k]) & 0xFFu] ^ (v >> 8);
  return v;
}

__global__ void __launch_bounds__(32 * BGZF_WARPS) bgzf_inflate_kernel(const uint8_t* __restrict__ comp, const BgzfMember* __restrict__ members,
                                                                       int n, uint8_t* out, uint32_t* __restrict__ status, const CrcShift sh) {
  __shared__ uint32_t tab[256];
  for (int t = threadIdx.x; t < 256; t += 32 * BGZF_WARPS) {  // the reflected CRC-32 table (polynomial 0xEDB88320)
    uint32_t c = (uint32_t)t;
    for (int k = 0; k < 8; k++) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
    tab[t] = c;
  }
  __syncthreads();
  const int i = blockIdx.x * BGZF_WARPS + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= n) return;
  const BgzfMember d = members[i];
  int err = 0;
  if (lane == 0) err = bgzf_inflate_member(comp, d, out);
  err = __shfl_sync(0xffffffffu, err, 0);  // (also orders lane 0's writes before the reads below)
  if (!err) {
    // CRC-32 of the member's output: the first (len mod 2 KiB) bytes by lane 0 from the all-ones register, then the
    // full slices from zero registers in parallel, folded in order: v = shift(v) ^ raw(slice).
    const uint8_t* o = out + d.out_off;
    const uint32_t len = d.out_len, nfull = len / CRC_SLICE, rem = len % CRC_SLICE;
    const uint32_t raw = (uint32_t)lane < nfull ? crc_bytes(tab, o + rem + (uint32_t)lane * CRC_SLICE, CRC_SLICE, 0u) : 0u;
    uint32_t v = 0xFFFFFFFFu;
    if (lane == 0) v = crc_bytes(tab, o, rem, v);
    v = __shfl_sync(0xffffffffu, v, 0);
    for (uint32_t j = 0; j < nfull; j++) {   // every lane folds the same chain (uniform, 32 x nfull steps)
      uint32_t w = 0;
#pragma unroll 8
      for (int k = 0; k < 32; k++) w ^= ((v >> k) & 1u) ? sh.col[k] : 0u;
      v = w ^ __shfl_sync(0xffffffffu, raw, (int)j);
    }
    if ((v ^ 0xFFFFFFFFu) != d.crc) err = BGZF_ECRC;
  }
  if (lane == 0) status[i] = (uint32_t)err;
}

cudaError_t launch_bgzf_inflate(const uint8_t* d_comp, const BgzfMember* d_members, int n, uint8_t* d_out, uint32_t* d_status,
                                cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  static const CrcShift sh = [] {
    CrcShift s;
    for (int k = 0; k < 32; k++) {
      uint32_t v = 1u << k;
      for (uint32_t b = 0; b < CRC_SLICE * 8; b++) v = (v & 1u) ? 0xEDB88320u ^ (v >> 1) : v >> 1;  // one zero bit per step
      s.col[k] = v;
    }
    return s;
  }();
  bgzf_inflate_kernel<<<(n + BGZF_WARPS - 1) / BGZF_WARPS, 32 * BGZF_WARPS, 0, st>>>(d_comp, d_members, n, d_out, d_status, sh);
  return cudaGetLastError();
}

}  // namespace fq
