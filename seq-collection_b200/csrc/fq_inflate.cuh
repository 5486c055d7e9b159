// fq_inflate.cuh -- the DEFLATE (RFC 1951) pieces shared by the two on-device inflaters: BGZF members
// (fq_bgzf.cu) and ordinary single-member gzip (fq_gzip.cu).  The reference inflates on the host through zlib
// (src/utils/gzip_stream.nim:16-17); what zlib's inflate() rejects is rejected here too, so that a file the
// device path accepts is a file the reference reads to its end.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fq {

static __constant__ uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59,
                                              67, 83, 99, 115, 131, 163, 195, 227, 258};
static __constant__ uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static __constant__ uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769,
                                               1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static __constant__ uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
static __constant__ uint8_t kClOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

constexpr int LBITS = 10;             // literal/length codes up to this length decode with one table lookup
constexpr int DBITS = 8;              // distance codes

// Per-warp decoding state in shared memory.  Every lane of the warp runs the same decode on the same bits (the
// loads are broadcasts, there is no divergence, so 32 identical lanes cost what one costs) -- which makes the warp
// available, without any hand-over, for the parts that are parallel: filling the lookup tables, copying matches,
// the CRC.
struct WarpTables {
  uint16_t lit[1 << LBITS];   // symbol << 4 | code length; 0 = longer than LBITS (canonical walk)
  uint16_t dist[1 << DBITS];
  uint16_t lcount[16], lsym[288], dcount[16], dsym[32];
  uint8_t lens[320];
};

// LSB-first bit reader over aligned 32-bit words.  Reads at most into the member's own 8-byte trailer; beyond it
// the stream continues as zeros, so a truncated or corrupt stream stays inside the batch buffer and ends in a
// decoder error, a size mismatch or a CRC mismatch.
struct Bits {
  const uint32_t* w;
  const uint32_t* wend;
  unsigned long long buf;
  int cnt;
  __device__ __forceinline__ void init(const uint8_t* p, uint32_t nbytes) {
    const uint32_t mis = (uint32_t)((uintptr_t)p & 3u);
    w = reinterpret_cast<const uint32_t*>(p - mis);
    wend = reinterpret_cast<const uint32_t*>(p + ((nbytes + 8u) & ~3u));
    buf = (unsigned long long)(__ldg(w++) >> (8u * mis));
    cnt = 32 - 8 * (int)mis;
  }
  __device__ __forceinline__ void refill() {  // afterwards cnt >= 33
    if (cnt <= 32) { buf |= (unsigned long long)(w < wend ? __ldg(w) : 0u) << cnt; w++; cnt += 32; }
  }
  __device__ __forceinline__ uint32_t take(int n) {  // n <= 16
    const uint32_t v = (uint32_t)buf & ((1u << n) - 1u);
    buf >>= n; cnt -= n;
    return v;
  }
};

// Canonical Huffman code from code lengths (all lanes run it; the stores are the same values to the same places):
// count[l] = codes of length l, symbol[] = symbols ordered by code; and the lookup table `tab` of 2^bits entries
// (filled by the 32 lanes together).  Returns < 0 for an over-subscribed set of lengths, > 0 for an incomplete one.
static __device__ __noinline__ int huff_build(uint16_t* count, uint16_t* symbol, uint16_t* tab, int bits, const uint8_t* len, int n, int lane) {
  for (int l = 0; l <= 15; l++) count[l] = 0;
  __syncwarp();
  if (lane == 0) for (int s = 0; s < n; s++) count[len[s]]++;
  __syncwarp();
  for (int k = lane; k < (1 << bits); k += 32) tab[k] = 0;
  if (count[0] == n) return 0;
  int left = 1;
  uint32_t next[16];  // first code of every length
  uint32_t code = 0;
  next[0] = 0;
  for (int l = 1; l <= 15; l++) {
    left <<= 1; left -= count[l];
    if (left < 0) return left;
    code = (code + (l > 1 ? count[l - 1] : 0)) << 1;
    next[l] = code;
  }
  uint16_t offs[16];
  offs[1] = 0;
  for (int l = 1; l < 15; l++) offs[l + 1] = offs[l] + count[l];
  __syncwarp();
  for (int s = 0; s < n; s++) {
    const int l = len[s];
    if (!l) continue;
    if (lane == 0) symbol[offs[l]] = (uint16_t)s;
    offs[l]++;
    const uint32_t c = next[l]++;
    if (l <= bits) {  // every table slot whose low l bits are the reversed code
      const uint32_t r = __brev(c) >> (32 - l);
      const uint16_t e = (uint16_t)((s << 4) | l);
      for (uint32_t k = r + ((uint32_t)lane << l); k < (1u << bits); k += 32u << l) tab[k] = e;
    }
  }
  __syncwarp();
  return left;
}
// zlib's inflate_table(): an over-subscribed set of lengths is an error, an incomplete one too unless it is a single
// code of length 1 (`left` = huff_build's result for the n lengths counted in `count`).
__device__ __forceinline__ bool huff_acceptable(int left, const uint16_t* count, int n) {
  return left == 0 || (left > 0 && count[1] == 1 && count[0] == n - 1);
}
// Codes longer than the table: walk the canonical code one bit at a time.  Takes the bits BY VALUE and returns
// symbol | length << 16 (or -1): a reader passed by reference to a function that is not inlined would live in local
// memory for the whole decode loop (measured: a third of the loop's stalls were loads of the reader's own fields).
static __device__ __noinline__ int huff_walk(unsigned long long bb, const uint16_t* count, const uint16_t* symbol) {
  int code = 0, first = 0, index = 0;
  for (int l = 1; l <= 15; l++) {
    code |= (int)(bb & 1ull);
    bb >>= 1;
    const int c = count[l];
    if (code - c < first) return (int)symbol[index + (code - first)] | (l << 16);
    index += c; first += c;
    first <<= 1; code <<= 1;
  }
  return -1;
}
template <class B>
__device__ __forceinline__ int huff_decode(B& b, const uint16_t* tab, int bits, const uint16_t* count, const uint16_t* symbol) {
  b.refill();
  const uint32_t e = tab[(uint32_t)b.buf & ((1u << bits) - 1u)];
  int l = (int)(e & 15u);
  int sym = (int)(e >> 4);
  if (!l) {
    const int r = huff_walk(b.buf, count, symbol);
    if (r < 0) return -1;
    l = r >> 16; sym = r & 0xFFFF;
  }
  b.buf >>= l; b.cnt -= l;
  return sym;
}

}  // namespace fq
