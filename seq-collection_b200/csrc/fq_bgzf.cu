// fq_bgzf.cu -- on-device inflate of BGZF input (blocked gzip: bgzip, htslib, many sequencer pipelines) feeding
// the scan kernels.
//
// The reference inflates .gz input on the host, one byte stream through zlib (src/utils/gzip_stream.nim:16-17,
// src/fq_count.nim:32); a single gzip member has to be guessed at to be inflated in parallel (fq_gzip.cu does that).
// A BGZF file needs no guessing: it is a concatenation of
// gzip members of at most 64 KiB, each with its compressed size in the header's "BC" extra field and its
// uncompressed size in the trailer, so every member is an independent DEFLATE stream with a known output
// offset: the host only walks the member headers, the compressed bytes go to the GPU as they are, and ONE WARP
// PER MEMBER inflates its stream (RFC 1951: stored, fixed and dynamic Huffman blocks) straight to its place in the
// output buffer.  A DEFLATE stream is serial, and 32 members sharing a warp run one after the other (measured
// 1.01 active threads per instruction), so each member gets its own warp; all 32 lanes decode the same bits in
// lockstep (broadcast loads, no divergence: 32 identical lanes cost what one costs), which leaves the warp free for
// the parallel parts -- filling the per-warp lookup tables in shared memory (fq_inflate.cuh: 9-bit literal/length,
// 8-bit distance, one 32-bit entry per symbol with everything the loop needs; longer codes walk the canonical code),
// copying matches 32 bytes per step, the CRC.  The output is the buffer fqgpu_scan_device then scans like any
// HBM-resident input.  Anything that is not well-formed BGZF makes the caller take the next path (gzip on the device,
// then zlib).
#include <cuda_runtime.h>
#include <stdint.h>

#include "fq_bgzf.h"
#include "fq_inflate.cuh"

namespace fq {

#ifndef BGZF_MINB
#define BGZF_MINB 10  // resident CTAs per SM the kernel is compiled for (register cap)
#endif
constexpr int BGZF_WARPS = 4;         // members per CTA
constexpr uint32_t CRC_SLICE = 2048;  // bytes per lane in the CRC pass (32 slices cover a 64 KiB member)

enum { BGZF_OK = 0, BGZF_EBLOCK = 1, BGZF_ECODE = 2, BGZF_EDIST = 3, BGZF_ESIZE = 4, BGZF_ETRUNC = 5, BGZF_ECRC = 6 };

// Inflates member d (all 32 lanes in lockstep).  Returns BGZF_*.
__device__ __noinline__ int bgzf_inflate_member(const uint8_t* __restrict__ comp, const BgzfMember& d, uint8_t* out, WarpTables& t, int lane) {
  // (the batch buffer is 4-byte aligned; the reader may run into the member's own 8-byte trailer, not beyond)
  GzBits b;
  b.init(reinterpret_cast<const uint32_t*>(comp), reinterpret_cast<const uint32_t*>(comp) + ((d.in_off + d.in_len + 8u + 3u) >> 2), d.in_off * 8ull);
  uint8_t* o = out + d.out_off;
  const uint32_t cap = d.out_len;
  uint32_t produced = 0;
  int last = 0;
  do {
    b.refill();
    last = (int)b.take(1);
    const int type = (int)b.take(2);
    if (type == 0) {  // stored
      b.take(b.cnt & 7);
      b.refill();
      const uint32_t len = b.take(16);
      b.refill();
      const uint32_t nlen = b.take(16);
      if ((len ^ 0xFFFFu) != nlen) return BGZF_EBLOCK;
      if (produced + len > cap) return BGZF_ESIZE;
      for (uint32_t k = 0; k < len; k++) { b.refill(); const uint32_t v = b.take(8); if (lane == 0) o[produced + k] = (uint8_t)v; }
      produced += len;
      continue;
    }
    if (type == 3) return BGZF_EBLOCK;
    __syncwarp();
    if (type == 1) {  // fixed codes
      for (int s = lane; s < 288; s += 32) t.lens[s] = s < 144 ? 8 : (s < 256 ? 9 : (s < 280 ? 7 : 8));
      __syncwarp();
      huff_build<HUFF_LITLEN>(t.lcount, t.lsym, t.lit, LBITS, t.lens, 288, lane);
      for (int s = lane; s < 30; s += 32) t.lens[s] = 5;
      __syncwarp();
      huff_build<HUFF_DIST>(t.dcount, t.dsym, t.dist, DBITS, t.lens, 30, lane);
    } else {  // dynamic codes
      b.refill();
      const int nlen = (int)b.take(5) + 257, ndist = (int)b.take(5) + 1, ncode = (int)b.take(4) + 4;
      if (nlen > 286 || ndist > 30) return BGZF_ECODE;
      for (int k = lane; k < 19; k += 32) t.lens[k] = 0;
      __syncwarp();
      for (int k = 0; k < ncode; k++) { b.refill(); const uint32_t v = b.take(3); if (lane == 0) t.lens[kClOrder[k]] = (uint8_t)v; }
      __syncwarp();
      // the code-length code (19 symbols, at most 7 bits) goes through the distance table's slots
      if (huff_build<HUFF_PRECODE>(t.dcount, t.dsym, t.dist, PBITS, t.lens, 19, lane) != 0) return BGZF_ECODE;
      int idx = 0;
      while (idx < nlen + ndist) {
        const uint32_t pe = huff_peek<HUFF_PRECODE, PBITS>(b, t.dist, t.dcount, t.dsym);
        if (!(pe & 15u)) return BGZF_ECODE;
        const int sym = (int)huff_take(b, pe);
        if (sym < 16) { if (lane == 0) t.lens[idx] = (uint8_t)sym; idx++; continue; }
        int prev = 0, rep;
        b.refill();
        __syncwarp();
        if (sym == 16) { if (idx == 0) return BGZF_ECODE; prev = t.lens[idx - 1]; rep = 3 + (int)b.take(2); }
        else if (sym == 17) rep = 3 + (int)b.take(3);
        else rep = 11 + (int)b.take(7);
        if (idx + rep > nlen + ndist) return BGZF_ECODE;
        if (lane == 0) for (int k = 0; k < rep; k++) t.lens[idx + k] = (uint8_t)prev;
        idx += rep;
        __syncwarp();
      }
      __syncwarp();
      if (t.lens[256] == 0) return BGZF_ECODE;
      // (zlib rejects incomplete codes other than a single code of length 1: so does this, and the file goes to zlib)
      if (!huff_acceptable(huff_build<HUFF_DIST>(t.dcount, t.dsym, t.dist, DBITS, t.lens + nlen, ndist, lane), t.dcount, ndist)) return BGZF_ECODE;
      if (!huff_acceptable(huff_build<HUFF_LITLEN>(t.lcount, t.lsym, t.lit, LBITS, t.lens, nlen, lane), t.lcount, nlen)) return BGZF_ECODE;
    }
    for (;;) {  // literals and matches of this DEFLATE block
      const uint32_t e = huff_peek<HUFF_LITLEN, LBITS>(b, t.lit, t.lcount, t.lsym);
      if (!(e & HK_MASK)) {  // a literal
        huff_take(b, e);
        if (produced >= cap) return BGZF_ESIZE;
        if (lane == 0) o[produced] = (uint8_t)(e >> 16);
        produced++;
      } else if ((e & HK_MASK) != HK_MATCH) {
        if ((e & HK_MASK) == HK_INVALID) return BGZF_ECODE;
        huff_take(b, e);  // end of block
        break;
      } else {
        const uint32_t len = huff_take(b, e);  // (>= 33 bits were there: 15 + 5 used)
        const uint32_t de = huff_peek<HUFF_DIST, DBITS>(b, t.dist, t.dcount, t.dsym);
        if ((de & HK_MASK) == HK_INVALID) return BGZF_EDIST;
        const uint32_t dist = huff_take(b, de);
        if (dist > produced) return BGZF_EDIST;
        if (produced + len > cap) return BGZF_ESIZE;
        uint8_t* dst = o + produced;
        const uint8_t* src = dst - dist;
        __syncwarp();  // earlier bytes (lane 0's literals, other lanes' match bytes) are visible to every lane
        if (dist >= len) {
          for (uint32_t k = (uint32_t)lane; k < len; k += 32u) dst[k] = src[k];
        } else if (dist >= 32u) {  // overlapping, period >= one round: rounds read what earlier rounds wrote
          for (uint32_t k0 = 0; k0 < len; k0 += 32u) {
            const uint32_t k = k0 + (uint32_t)lane;
            if (k < len) dst[k] = src[k];
            __syncwarp();
          }
        } else {  // a run with a short period: every byte is one of the dist bytes before it
          for (uint32_t k = (uint32_t)lane; k < len; k += 32u) dst[k] = src[k % dist];
        }
        produced += len;
      }
    }
  } while (!last);
  if (produced != cap) return BGZF_ESIZE;
  return BGZF_OK;
}

// crc_shift[k] = the CRC register after CRC_SLICE zero bytes when it starts as 1 << k (a linear map over GF(2)).
struct CrcShift { uint32_t col[32]; };

__device__ __forceinline__ uint32_t crc_bytes(const uint32_t* tab, const uint8_t* p, uint32_t n, uint32_t v) {
  for (uint32_t k = 0; k < n; k++) v = tab[(v ^ p[k]) & 0xFFu] ^ (v >> 8);
  return v;
}

__global__ void __launch_bounds__(32 * BGZF_WARPS, BGZF_MINB) bgzf_inflate_kernel(const uint8_t* __restrict__ comp, const BgzfMember* __restrict__ members,
                                                                       int n, uint8_t* out, uint32_t* __restrict__ status, const CrcShift sh) {
  __shared__ uint32_t tab[256];
  __shared__ WarpTables tables[BGZF_WARPS];
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
  int err = bgzf_inflate_member(comp, d, out, tables[threadIdx.x >> 5], lane);
  __syncwarp();
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
