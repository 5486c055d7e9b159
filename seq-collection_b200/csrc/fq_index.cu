// fq_index.cu -- the record-offset index as a product (the stand-alone output of the
// boundary classification): byte offset of the first byte of every record (line 4k) of an HBM-resident
// FASTQ buffer, and a gather of the first n header lines for fq-meta's sequencer / barcode detection
// (src/fq_meta.nim:229-242 reads them with readLine; extract_read_info :151-178 stays on the host).
//
// Line semantics are those of the scan (Nim streams.lines, src/fq_count.nim:38): lines end at '\n'; the
// trailing unterminated line exists when it is non-empty; record k is the line with index 4k.
//
// One launch (fq_index_onepass_kernel): tiles of 64 / 128 / 256 KiB are handed out by a ticket counter; a tile becomes
// its newline bitmap in shared memory, one block-wide prefix gives its newline count, the count is published and the
// predecessors' counts are summed by a chained (decoupled look-back) prefix, and the successor of every newline whose
// index is 3 (mod 4) is stored.  The input is read once (1 byte of HBM traffic per input byte).
// FQGPU_INDEX=2pass keeps the earlier three launches: newline count per 64 KiB tile -> exclusive prefix over the
// tiles (one CTA) -> write pass (the same bitmap phase), two reads of the input.  Nothing depends on the statistics kernels.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "fqgpu_ctx.h"

namespace fq {

constexpr int IDX_THREADS = 256;
constexpr int IDX_ROWS = 16;                               // 16-byte groups per thread and tile
constexpr int IDX_TILE = IDX_THREADS * IDX_ROWS * 16;      // 64 KiB

__device__ __forceinline__ uint32_t idx_nl_flags(uint32_t w) {  // 0x80 in every byte lane that equals '\n' (exact)
  const uint32_t x = w ^ 0x0A0A0A0Au;
  return ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u;
}
__device__ __forceinline__ uint32_t idx_nl_mask16(const uint4& v) {
  const uint32_t lo = __dp4a(idx_nl_flags(v.x), 0x08040201u, __dp4a(idx_nl_flags(v.y), 0x80402010u, 0u));
  const uint32_t hi = __dp4a(idx_nl_flags(v.z), 0x08040201u, __dp4a(idx_nl_flags(v.w), 0x80402010u, 0u));
  return (lo >> 7) | (hi << 1);
}
// Newline mask of group g (16 bytes at base + 16 g), restricted to the valid bytes [lo0, end).
__device__ __forceinline__ uint32_t idx_group_mask(const uint8_t* __restrict__ base, u64 g, uint32_t lo0, u64 end) {
  const u64 off = g * 16;
  if (off >= end || off + 16 <= (u64)lo0) return 0u;
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(base + off));
  uint32_t m = idx_nl_mask16(v);
  if (off + 16 > end) m &= (1u << (end - off)) - 1u;
  if (off < (u64)lo0) m &= ~((1u << ((u64)lo0 - off)) - 1u);
  return m;
}

__global__ void __launch_bounds__(IDX_THREADS) fq_index_count_kernel(const uint8_t* __restrict__ base, uint32_t lo0, u64 end,
                                                                    uint32_t* __restrict__ tile_cnt) {
  __shared__ uint32_t wsum[IDX_THREADS / 32];
  const u64 g0 = (u64)blockIdx.x * (IDX_THREADS * IDX_ROWS);
  uint32_t c = 0;
  if (g0 * 16 >= (u64)lo0 && (g0 + IDX_THREADS * IDX_ROWS) * 16 <= end) {  // interior tile: no range checks; the flags
    // (0x80 per newline byte) are summed by IDP.4A, 128 per newline
    const uint4* p = reinterpret_cast<const uint4*>(base) + g0 + threadIdx.x;
    uint32_t acc = 0;
#pragma unroll 8
    for (int r = 0; r < IDX_ROWS; r++) {
      const uint4 v = __ldg(p + (size_t)r * IDX_THREADS);
      acc = __dp4a(idx_nl_flags(v.x), 0x01010101u, acc); acc = __dp4a(idx_nl_flags(v.y), 0x01010101u, acc);
      acc = __dp4a(idx_nl_flags(v.z), 0x01010101u, acc); acc = __dp4a(idx_nl_flags(v.w), 0x01010101u, acc);
    }
    c = acc >> 7;
  } else {
#pragma unroll 4
    for (int r = 0; r < IDX_ROWS; r++) c += __popc(idx_group_mask(base, g0 + (u64)r * IDX_THREADS + threadIdx.x, lo0, end));
  }
  c = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int w = 0; w < IDX_THREADS / 32; w++) t += wsum[w];
    tile_cnt[blockIdx.x] = t;
  }
}

// Exclusive prefix of the tile counts (one CTA); out[0] = lines (incl. a non-empty unterminated last one), out[1] = records.
__global__ void __launch_bounds__(1024) fq_index_scan_kernel(const uint32_t* __restrict__ tile_cnt, u64* __restrict__ tile_base, u64 ntiles,
                                                            const uint8_t* __restrict__ base, uint32_t lo0, u64 end, u64* __restrict__ out) {
  __shared__ u64 wtot[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const u64 per = (ntiles + 1023) / 1024, a = (u64)tid * per, b = a + per < ntiles ? a + per : ntiles;
  u64 s = 0;
#pragma unroll 8
  for (u64 t = a; t < b; t++) s += tile_cnt[t];
  u64 inc = s;  // inclusive prefix over the 1024 slices: inside the warp, then over the 32 warp totals
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { const u64 x = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += x; }
  if (lane == 31) wtot[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const u64 w = wtot[lane];
    u64 winc = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const u64 x = __shfl_up_sync(0xffffffffu, winc, d); if (lane >= d) winc += x; }
    wtot[lane] = winc - w;
    if (lane == 31) {
      const u64 n = end - (u64)lo0;
      const u64 lines = winc + ((n > 0 && base[end - 1] != '\n') ? 1 : 0);
      out[0] = lines;
      out[1] = (lines + 3) / 4;
    }
  }
  __syncthreads();
  u64 run = wtot[warp] + inc - s;
  for (u64 t = a; t < b; t++) { tile_base[t] = run; run += tile_cnt[t]; }
}

// Write pass.  Phase 1 turns the tile into its newline bitmap in shared memory (coalesced 16-byte loads, one 16-bit mask
// per group); phase 2 hands every thread 256 consecutive bytes of it (8 words), so the tile needs ONE block-wide prefix
// of the per-thread newline counts instead of one per row of 256 groups, and a thread walks its own few newlines.
__global__ void __launch_bounds__(IDX_THREADS) fq_index_write_kernel(const uint8_t* __restrict__ base, uint32_t lo0, u64 end,
                                                                    const u64* __restrict__ tile_base, u64* __restrict__ offsets, u64 cap) {
  __shared__ __align__(16) uint32_t bits[IDX_TILE / 32];  // bit b = byte b of the tile is '\n'
  __shared__ uint32_t wsum[IDX_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const u64 g0 = (u64)blockIdx.x * (IDX_THREADS * IDX_ROWS);
  const u64 n = end - (u64)lo0;
  if (blockIdx.x == 0 && tid == 0 && n > 0 && cap > 0) offsets[0] = 0;
  uint16_t* b16 = reinterpret_cast<uint16_t*>(bits);
  if (g0 * 16 >= (u64)lo0 && (g0 + IDX_THREADS * IDX_ROWS) * 16 <= end) {  // interior tile: no range checks
    const uint4* p = reinterpret_cast<const uint4*>(base) + g0 + tid;
#pragma unroll 8
    for (int r = 0; r < IDX_ROWS; r++) b16[r * IDX_THREADS + tid] = (uint16_t)idx_nl_mask16(__ldg(p + (size_t)r * IDX_THREADS));
  } else {
#pragma unroll 4
    for (int r = 0; r < IDX_ROWS; r++) b16[r * IDX_THREADS + tid] = (uint16_t)idx_group_mask(base, g0 + (u64)r * IDX_THREADS + tid, lo0, end);
  }
  __syncthreads();
  const uint4 lo4 = reinterpret_cast<const uint4*>(bits)[2 * tid], hi4 = reinterpret_cast<const uint4*>(bits)[2 * tid + 1];
  const uint32_t w[8] = {lo4.x, lo4.y, lo4.z, lo4.w, hi4.x, hi4.y, hi4.z, hi4.w};
  uint32_t c = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) c += __popc(w[i]);
  uint32_t inc = c;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { const uint32_t x = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += x; }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  uint32_t wb = 0;
#pragma unroll
  for (int i = 0; i < IDX_THREADS / 32; i++) { const uint32_t x = wsum[i]; if (i < warp) wb += x; }
  if (c == 0) return;
  u64 j = tile_base[blockIdx.x] + wb + inc - c;        // index of this thread's first newline
  const u64 byte0 = g0 * 16 + (u64)tid * 256 + 1 - (u64)lo0;  // successor of the thread's byte 0, relative to the data
#pragma unroll
  for (int i = 0; i < 8; i++) {
    uint32_t m = w[i];
    while (m) {
      const int k = __ffs(m) - 1;
      m &= m - 1;
      if (((uint32_t)j & 3u) == 3u) {
        const u64 start = byte0 + (u64)(32 * i + k), rec = (j + 1) >> 2;
        if (start < n && rec < cap) offsets[rec] = start;
      }
      j++;
    }
  }
}

// ---- single pass: the same tile work with a chained ("decoupled look-back") prefix, so the input is read once ------------
// state[t]: bits 63..62 = 0 nothing yet, 1 = newlines of tile t alone, 2 = newlines of tiles 0..t; tiles are handed out by a
// ticket counter (state[ntiles]), so a tile only ever waits for tiles that have already started.  state[ntiles + 1], [ntiles + 2]
// receive lines and records.  The host zeroes the whole array before the launch.
constexpr u64 IDX_AGG = 1ull << 62, IDX_INC = 2ull << 62, IDX_VAL = IDX_AGG - 1;
__device__ __forceinline__ u64 idx_ld_relaxed(const u64* p) {
  u64 v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void idx_st_relaxed(u64* p, u64 v) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }

template <int ROWS>  // 16-byte groups per thread: the tile is 4 KiB * ROWS
__global__ void __launch_bounds__(IDX_THREADS, ROWS == 64 ? 5 : ROWS == 32 ? 6 : 8) fq_index_onepass_kernel(const uint8_t* __restrict__ base, uint32_t lo0, u64 end, u64* state, u64 ntiles,
                                                                      u64* __restrict__ offsets, u64 cap) {
  __shared__ __align__(16) uint32_t bits[IDX_THREADS * ROWS / 2];  // bit b = byte b of the tile is '\n'
  __shared__ uint32_t wsum[IDX_THREADS / 32];
  __shared__ u64 s_tile, s_prefix;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_tile = atomicAdd(reinterpret_cast<unsigned long long*>(state + ntiles), 1ull);
  __syncthreads();
  const u64 tile = s_tile;
  const u64 g0 = tile * (IDX_THREADS * ROWS);
  const u64 n = end - (u64)lo0;
  if (tile == 0 && tid == 0 && n > 0 && cap > 0) offsets[0] = 0;
  uint16_t* b16 = reinterpret_cast<uint16_t*>(bits);
  if (g0 * 16 >= (u64)lo0 && (g0 + IDX_THREADS * ROWS) * 16 <= end) {  // interior tile: no range checks
    const uint4* p = reinterpret_cast<const uint4*>(base) + g0 + tid;
#pragma unroll 8
    for (int r = 0; r < ROWS; r++) b16[r * IDX_THREADS + tid] = (uint16_t)idx_nl_mask16(__ldg(p + (size_t)r * IDX_THREADS));
  } else {
#pragma unroll 4
    for (int r = 0; r < ROWS; r++) b16[r * IDX_THREADS + tid] = (uint16_t)idx_group_mask(base, g0 + (u64)r * IDX_THREADS + tid, lo0, end);
  }
  __syncthreads();
  constexpr int WPT = ROWS / 2;  // bitmap words per thread: 16 * ROWS consecutive bytes
  uint32_t w[WPT];
#pragma unroll
  for (int i = 0; i < WPT / 4; i++) {
    const uint4 q = reinterpret_cast<const uint4*>(bits)[(WPT / 4) * tid + i];
    w[4 * i] = q.x; w[4 * i + 1] = q.y; w[4 * i + 2] = q.z; w[4 * i + 3] = q.w;
  }
  uint32_t c = 0;
#pragma unroll
  for (int i = 0; i < WPT; i++) c += __popc(w[i]);
  uint32_t inc = c;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { const uint32_t x = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += x; }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  uint32_t wb = 0, tot = 0;
#pragma unroll
  for (int i = 0; i < IDX_THREADS / 32; i++) { const uint32_t x = wsum[i]; if (i < warp) wb += x; tot += x; }
  if (warp == 0) {  // publish this tile's count, then walk back over the predecessors, 32 at a time
    if (lane == 0 && tile > 0) idx_st_relaxed(state + tile, IDX_AGG | (u64)tot);
    u64 excl = 0;
    if (tile > 0) {
      long long i = (long long)tile - 1;
      for (;;) {
        const long long p = i - lane;
        u64 st;
        do { st = p >= 0 ? idx_ld_relaxed(state + p) : IDX_INC; } while (__any_sync(0xffffffffu, (st >> 62) == 0));
        const uint32_t pm = __ballot_sync(0xffffffffu, (st >> 62) == 2);
        const int first = pm ? __ffs(pm) - 1 : 31;  // the nearest predecessor that already knows its inclusive prefix
        u64 v = lane <= first ? (st & IDX_VAL) : 0;
#pragma unroll
        for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        excl += v;
        if (pm) break;
        i -= 32;
      }
    }
    if (lane == 0) {
      idx_st_relaxed(state + tile, IDX_INC | (excl + (u64)tot));
      s_prefix = excl;
      if (tile + 1 == ntiles) {
        const u64 lines = excl + (u64)tot + ((n > 0 && base[end - 1] != '\n') ? 1 : 0);
        state[ntiles + 1] = lines;
        state[ntiles + 2] = (lines + 3) / 4;
      }
    }
  }
  __syncthreads();
  if (c == 0) return;
  u64 j = s_prefix + wb + inc - c;                            // index of this thread's first newline
  const u64 byte0 = g0 * 16 + (u64)tid * (16 * ROWS) + 1 - (u64)lo0;  // successor of the thread's byte 0, relative to the data
#pragma unroll
  for (int i = 0; i < WPT; i++) {
    uint32_t m = w[i];
    while (m) {
      const int k = __ffs(m) - 1;
      m &= m - 1;
      if (((uint32_t)j & 3u) == 3u) {
        const u64 start = byte0 + (u64)(32 * i + k), rec = (j + 1) >> 2;
        if (start < n && rec < cap) offsets[rec] = start;
      }
      j++;
    }
  }
}

// First `n` header lines (records 0..n-1) into a row-major matrix: up to `stride` bytes each; len[k] = bytes
// copied (content up to '\n', a '\r' directly before it dropped, truncated at stride).  One warp per record.
__global__ void fq_headers_kernel(const uint8_t* __restrict__ data, u64 nbytes, const u64* __restrict__ offsets, u64 n,
                                  uint32_t stride, uint8_t* __restrict__ out, uint32_t* __restrict__ len) {
  const u64 k = (u64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (k >= n) return;
  const u64 s = offsets[k];
  uint32_t L = stride;  // content bytes to copy
  bool term = false;
  for (uint32_t o = 0; o < stride + 1 && !term; o += 32) {
    const u64 p = s + o + lane;
    const int c = p < nbytes ? (int)data[p] : -1;
    const uint32_t hit = __ballot_sync(0xffffffffu, c == '\n' || c < 0);
    if (hit) { const uint32_t e = o + (uint32_t)__ffs(hit) - 1; L = e < stride ? e : stride; term = e <= stride; }
  }
  if (term && L > 0 && L <= stride) {  // drop one '\r' directly before the newline
    const u64 p = s + L;
    if (p < nbytes && data[p] == '\n' && data[p - 1] == '\r') L--;
  }
  for (uint32_t o = lane; o < L; o += 32) out[k * stride + o] = data[s + o];
  if (lane == 0) len[k] = L;
}

}  // namespace fq

extern "C" {

int fqgpu_index_device(fqgpu_ctx* ctx, const void* dptr, size_t nbytes, uint64_t* d_offsets, uint64_t cap, uint64_t* n_records) {
  if (!ctx || !n_records || (cap && !d_offsets)) return FQGPU_EARG;
  *n_records = 0;
  if (nbytes == 0) return FQGPU_OK;
  if (!dptr) return fail(ctx, FQGPU_EARG, "fqgpu_index_device: NULL pointer");
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  const uintptr_t addr = (uintptr_t)dptr;
  const uint32_t lo0 = (uint32_t)(addr & 15);
  const uint8_t* base = (const uint8_t*)(addr - lo0);
  const fq::u64 end = (fq::u64)lo0 + nbytes;
  const fq::u64 ntiles = (end + fq::IDX_TILE - 1) / fq::IDX_TILE;
  if (ntiles > 0x7FFFFFFFull) return fail(ctx, FQGPU_EARG, "fqgpu_index_device: buffer too large for one call");
  const char* mode = getenv("FQGPU_INDEX");
  const bool two_pass = mode && strcmp(mode, "2pass") == 0;
  fq::u64 h[2] = {0, 0};
  cudaEvent_t e0 = fqgpu_get_event(ctx), e1 = fqgpu_get_event(ctx);
  if (!two_pass) {  // one launch: chained prefix over the tiles, the input is read once
    // per-tile costs (ticket, four barriers, the walk back) are spread over larger tiles once there are enough of them
    // to fill the machine: 256 KiB tiles from 320 MiB, 128 KiB from 160 MiB, else 64 KiB (measured 4.4 / 4.8 / 5.0 TB/s
    // with 64 / 128 / 256 KiB tiles at 8.6 GB); FQGPU_INDEX_ROWS = 16 | 32 | 64 forces one (tests)
    const char* re = getenv("FQGPU_INDEX_ROWS");
    int rows = re ? atoi(re) : nbytes >= ((size_t)320 << 20) ? 64 : nbytes >= ((size_t)160 << 20) ? 32 : 16;
    if (rows != 16 && rows != 32 && rows != 64) rows = 16;  // (only these three are instantiated)
    const fq::u64 tile_bytes = (fq::u64)fq::IDX_THREADS * 16 * (fq::u64)rows;
    const fq::u64 ntiles = (end + tile_bytes - 1) / tile_bytes;
    fq::u64* d_state = nullptr;  // [ntiles] tile states, then the ticket counter, lines, records
    CU_TRY(ctx, cudaMallocAsync((void**)&d_state, (ntiles + 3) * sizeof(fq::u64), ctx->stream));
    CU_TRY(ctx, cudaEventRecord(e0, ctx->stream));
    CU_TRY(ctx, cudaMemsetAsync(d_state, 0, (ntiles + 3) * sizeof(fq::u64), ctx->stream));
    if (rows == 32) fq::fq_index_onepass_kernel<32><<<(unsigned)ntiles, fq::IDX_THREADS, 0, ctx->stream>>>(base, lo0, end, d_state, ntiles, (fq::u64*)d_offsets, cap);
    else if (rows == 64) fq::fq_index_onepass_kernel<64><<<(unsigned)ntiles, fq::IDX_THREADS, 0, ctx->stream>>>(base, lo0, end, d_state, ntiles, (fq::u64*)d_offsets, cap);
    else fq::fq_index_onepass_kernel<16><<<(unsigned)ntiles, fq::IDX_THREADS, 0, ctx->stream>>>(base, lo0, end, d_state, ntiles, (fq::u64*)d_offsets, cap);
    CU_TRY(ctx, cudaGetLastError());
    CU_TRY(ctx, cudaEventRecord(e1, ctx->stream));
    ctx->timed.emplace_back(e0, e1);
    ctx->launches += 1;
    CU_TRY(ctx, cudaMemcpyAsync(h, d_state + ntiles + 1, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaFreeAsync(d_state, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  } else {  // count per tile -> prefix (one CTA) -> write pass: two reads of the input
    uint32_t* d_cnt = nullptr;
    fq::u64* d_base = nullptr;
    fq::u64* d_out = nullptr;
    CU_TRY(ctx, cudaMallocAsync((void**)&d_cnt, ntiles * sizeof(uint32_t), ctx->stream));
    CU_TRY(ctx, cudaMallocAsync((void**)&d_base, ntiles * sizeof(fq::u64), ctx->stream));
    CU_TRY(ctx, cudaMallocAsync((void**)&d_out, 2 * sizeof(fq::u64), ctx->stream));
    CU_TRY(ctx, cudaEventRecord(e0, ctx->stream));
    fq::fq_index_count_kernel<<<(unsigned)ntiles, fq::IDX_THREADS, 0, ctx->stream>>>(base, lo0, end, d_cnt);
    fq::fq_index_scan_kernel<<<1, 1024, 0, ctx->stream>>>(d_cnt, d_base, ntiles, base, lo0, end, d_out);
    fq::fq_index_write_kernel<<<(unsigned)ntiles, fq::IDX_THREADS, 0, ctx->stream>>>(base, lo0, end, d_base, (fq::u64*)d_offsets, cap);
    CU_TRY(ctx, cudaGetLastError());
    CU_TRY(ctx, cudaEventRecord(e1, ctx->stream));
    ctx->timed.emplace_back(e0, e1);
    ctx->launches += 3;
    CU_TRY(ctx, cudaMemcpyAsync(h, d_out, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaFreeAsync(d_cnt, ctx->stream));
    CU_TRY(ctx, cudaFreeAsync(d_base, ctx->stream));
    CU_TRY(ctx, cudaFreeAsync(d_out, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  }
  *n_records = h[1];
  ctx->index_lines = h[0];
  return FQGPU_OK;
}

uint64_t fqgpu_index_lines(fqgpu_ctx* ctx) { return ctx ? ctx->index_lines : 0; }

int fqgpu_headers_device(fqgpu_ctx* ctx, const void* dptr, size_t nbytes, const uint64_t* d_offsets, uint64_t n,
                         uint32_t stride, uint8_t* h_out, uint32_t* h_len) {
  if (!ctx || (n && (!dptr || !d_offsets || !h_out || !h_len)) || stride == 0) return FQGPU_EARG;
  if (n == 0) return FQGPU_OK;
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  uint8_t* d_out = nullptr;
  uint32_t* d_len = nullptr;
  CU_TRY(ctx, cudaMallocAsync((void**)&d_out, n * stride, ctx->stream));
  CU_TRY(ctx, cudaMallocAsync((void**)&d_len, n * sizeof(uint32_t), ctx->stream));
  CU_TRY(ctx, cudaMemsetAsync(d_out, 0, n * stride, ctx->stream));
  fq::fq_headers_kernel<<<(unsigned)((n + 7) / 8), 256, 0, ctx->stream>>>((const uint8_t*)dptr, nbytes, (const fq::u64*)d_offsets, n, stride, d_out, d_len);
  CU_TRY(ctx, cudaGetLastError());
  CU_TRY(ctx, cudaMemcpyAsync(h_out, d_out, n * stride, cudaMemcpyDeviceToHost, ctx->stream));
  CU_TRY(ctx, cudaMemcpyAsync(h_len, d_len, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  CU_TRY(ctx, cudaFreeAsync(d_out, ctx->stream));
  CU_TRY(ctx, cudaFreeAsync(d_len, ctx->stream));
  CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return FQGPU_OK;
}

}  // extern "C"
