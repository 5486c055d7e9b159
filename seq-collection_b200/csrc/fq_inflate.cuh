// fq_inflate.cuh -- the DEFLATE (RFC 1951) pieces shared by the two on-device inflaters: BGZF members
// (fq_bgzf.cu) and ordinary single-member gzip (fq_gzip.cu).  The reference inflates on the host through zlib
// (src/utils/gzip_stream.nim:16-17); what zlib's inflate() rejects is rejected here too, so that a file the
// device path accepts is a file the reference reads to its end.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fq {

static __constant__ uint8_t kClOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

constexpr int LBITS = 9;   // literal/length codes up to this length decode with one table lookup
constexpr int DBITS = 8;   // distance codes
constexpr int PBITS = 7;   // the code-length code of a dynamic block (its table lives in the distance table's slots)

// A table entry says everything the decode loop needs about a symbol, so that nothing else is looked up or computed
// on the critical path:  bits 0-3 the code's length (0: the code is longer than the table, walk the canonical code),
// bits 4-7 the number of extra bits behind it, bits 8-9 what it is, bits 16-31 the byte / the length's or
// distance's base / the symbol (code-length code).
enum : uint32_t { HK_LITERAL = 0u << 8, HK_MATCH = 1u << 8, HK_END = 2u << 8, HK_INVALID = 3u << 8, HK_MASK = 3u << 8 };
enum { HUFF_LITLEN = 0, HUFF_DIST = 1, HUFF_PRECODE = 2 };

// Per-warp decoding state in shared memory.  Every lane of the warp runs the same decode on the same bits (the
// loads are broadcasts, there is no divergence, so 32 identical lanes cost what one costs) -- which makes the warp
// available, without any hand-over, for the parts that are parallel: filling the lookup tables, copying matches,
// the CRC.
struct WarpTables {
  uint32_t lit[1 << LBITS];
  uint32_t dist[1 << DBITS];
  uint16_t lcount[16], lsym[288], dcount[16], dsym[32];
  uint8_t lens[320];
};

// LSB-first bit reader over a buffer of 32-bit words that knows its position.  Beyond `wend` the stream reads as
// zeros, so a truncated or corrupt stream stays inside the buffer and ends in a decoder error, a size mismatch or a
// CRC mismatch.
struct GzBits {
  const uint32_t* w0;
  const uint32_t* w;     // the next word to put into `buf`
  const uint32_t* wend;
  unsigned long long buf;
  uint32_t ahead;        // *w, loaded one refill early so that the load's latency is not waited for
  int cnt;
  __device__ __forceinline__ uint32_t word(const uint32_t* p) const { return p < wend ? __ldg(p) : 0u; }
  __device__ __forceinline__ void init(const uint32_t* base, const uint32_t* end, unsigned long long bit) {
    w0 = base; wend = end;
    w = base + (bit >> 5);
    const uint32_t sh = (uint32_t)bit & 31u;
    buf = (unsigned long long)(word(w) >> sh);
    w++;
    ahead = word(w);
    cnt = 32 - (int)sh;
  }
  __device__ __forceinline__ void refill() {  // afterwards cnt >= 33
    if (cnt <= 32) {
      buf |= (unsigned long long)ahead << cnt;
      w++;
      cnt += 32;
      ahead = word(w);
    }
  }
  __device__ __forceinline__ uint32_t take(int n) {  // n <= 16
    const uint32_t v = (uint32_t)buf & ((1u << n) - 1u);
    buf >>= n; cnt -= n;
    return v;
  }
  __device__ __forceinline__ unsigned long long bitpos() const { return (unsigned long long)(w - w0) * 32ull - (unsigned long long)cnt; }
};

// Base and extra bits of length symbol s (0..28) and distance symbol d (0..29), RFC 1951 3.2.5, computed rather than
// looked up: an indexed load from constant memory sits on the critical path of every match.
__device__ __forceinline__ void len_code(int s, uint32_t& base, int& extra) {
  extra = s < 8 ? 0 : (s - 4) >> 2;
  base = s < 8 ? 3u + (uint32_t)s : (((4u + ((uint32_t)s & 3u)) << extra) + 3u);
  if (s == 28) { base = 258; extra = 0; }
}
__device__ __forceinline__ void dist_code(int d, uint32_t& base, int& extra) {
  extra = d < 4 ? 0 : (d - 2) >> 1;
  base = d < 4 ? 1u + (uint32_t)d : (((2u + ((uint32_t)d & 1u)) << extra) + 1u);
}

// The table entry of symbol s of the given alphabet whose code is l bits long.
template <int KIND>
__device__ __forceinline__ uint32_t huff_entry(int s, int l) {
  if (KIND == HUFF_PRECODE) return ((uint32_t)s << 16) | (uint32_t)l;
  if (KIND == HUFF_DIST) {
    if (s >= 30) return HK_INVALID | (uint32_t)l;
    uint32_t base; int extra;
    dist_code(s, base, extra);
    return (base << 16) | HK_MATCH | ((uint32_t)extra << 4) | (uint32_t)l;
  }
  if (s < 256) return ((uint32_t)s << 16) | HK_LITERAL | (uint32_t)l;
  if (s == 256) return HK_END | (uint32_t)l;
  if (s >= 286) return HK_INVALID | (uint32_t)l;
  uint32_t base; int extra;
  len_code(s - 257, base, extra);
  return (base << 16) | HK_MATCH | ((uint32_t)extra << 4) | (uint32_t)l;
}

// Canonical Huffman code from code lengths (all lanes run it; the stores are the same values to the same places):
// count[l] = codes of length l, symbol[] = symbols ordered by code; and the lookup table `tab` of 2^bits entries
// (filled by the 32 lanes together).  Returns < 0 for an over-subscribed set of lengths, > 0 for an incomplete one.
template <int KIND>
static __device__ __noinline__ int huff_build(uint16_t* count, uint16_t* symbol, uint32_t* tab, int bits, const uint8_t* len, int n, int lane) {
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
      const uint32_t e = huff_entry<KIND>(s, l);
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
// The entry of the next symbol; the reader is refilled first (>= 33 bits) and NOT advanced.  HK_INVALID with
// length 0: no such code.
template <int KIND, int BITS, class B>
__device__ __forceinline__ uint32_t huff_peek(B& b, const uint32_t* tab, const uint16_t* count, const uint16_t* symbol) {
  b.refill();
  uint32_t e = tab[(uint32_t)b.buf & ((1u << BITS) - 1u)];
  if (!(e & 15u)) {
    const int r = huff_walk(b.buf, count, symbol);
    e = r < 0 ? (uint32_t)HK_INVALID : huff_entry<KIND>(r & 0xFFFF, r >> 16);
  }
  return e;
}
// value = base + extra bits; the reader moves past the code and its extra bits
template <class B>
__device__ __forceinline__ uint32_t huff_take(B& b, uint32_t e) {
  const int l = (int)(e & 15u), x = (int)((e >> 4) & 15u);
  const uint32_t v = (e >> 16) + ((uint32_t)(b.buf >> l) & ((1u << x) - 1u));
  b.buf >>= l + x; b.cnt -= l + x;
  return v;
}

}  // namespace fq
