// fq_dedup.cu -- duplicate read-ID detection (src/fq_dedup.nim:14-84).
//
// The reference keeps the first record of every distinct header line (line 4k, compared as the string Nim's
// `lines` yields: without '\n' and without one '\r' directly before it) and drops every later record with the
// same header; it finds them with a Bloom filter and two passes over the file.  Here: one warp per record hashes
// its header line (64-bit polynomial hash) and enters it into an open-addressing table in device memory (slot =
// smallest record index with these header BYTES: a slot is claimed by compare-and-swap, an occupant is accepted only
// after the warp has compared the two header lines byte for byte, and then lowered with an atomic minimum); a second
// launch looks every record up again and keeps it when it is the slot's representative.  The hash only chooses
// where to look, the bytes decide: exact whatever the hash does.  No sort, no library call.
// Input: the record-offset index of fq_index.cu.
#include <cuda_runtime.h>
#include <stdint.h>

#include "fqgpu_ctx.h"

namespace fq {

// Content length of the header line that starts at `s`: up to '\n' or the end of the data, one '\r' directly
// before the '\n' dropped (warp-cooperative; all lanes return the same value).
__device__ __forceinline__ u64 header_len(const uint8_t* __restrict__ d, u64 n, u64 s, int lane) {
  for (u64 o = s;; o += 32) {
    const u64 p = o + lane;
    const int c = p < n ? (int)d[p] : -1;
    const uint32_t hit = __ballot_sync(0xffffffffu, c == '\n' || c < 0);
    if (hit) {
      u64 e = o + (u64)__ffs(hit) - 1;
      if (e < n && e > s && d[e - 1] == '\r') e--;  // (e < n: terminated by '\n', not by the end of the data)
      return e - s;
    }
  }
}

__global__ void fq_dedup_hash_kernel(const uint8_t* __restrict__ d, u64 n, const u64* __restrict__ off, u64 nrec,
                                     u64* __restrict__ hash, uint32_t* __restrict__ hlen) {
  const u64 k = (u64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (k >= nrec) return;
  const u64 s = off[k];
  const u64 L = header_len(d, n, s, lane);
  const u64 P = 0x9E3779B97F4A7C15ull;  // odd multiplier; h = sum (byte + 1) * P^(i + 1)
  u64 pw = P, p32 = P;
  for (int i = 0; i < lane; i++) pw *= P;  // P^(lane + 1)
  for (int i = 0; i < 5; i++) p32 *= p32;   // P^32
  u64 h = 0;
  for (u64 i = lane; i < L; i += 32) { h += ((u64)d[s + i] + 1) * pw; pw *= p32; }
  for (int o = 16; o > 0; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o);
  h ^= L * 0xD6E8FEB86659FD93ull;
  if (lane == 0) { hash[k] = h; hlen[k] = L > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)L; }
}

// Do records k and k2 have the same header bytes?  (warp-cooperative, 32 bytes per step; all lanes return the same value)
__device__ __forceinline__ bool same_header(const uint8_t* __restrict__ d, const u64* __restrict__ off, const uint32_t* __restrict__ hlen, u64 k, u64 k2, int lane) {
  const uint32_t L = hlen[k];
  if (hlen[k2] != L) return false;
  const uint8_t* a = d + off[k];
  const uint8_t* b = d + off[k2];
  for (uint32_t o = 0; o < L; o += 32) {
    const uint32_t i = o + (uint32_t)lane;
    const bool ne = i < L && a[i] != b[i];
    if (__any_sync(0xffffffffu, ne)) return false;
  }
  return true;
}
__device__ __forceinline__ u64 ld_volatile_u64(const u64* p) { return *reinterpret_cast<const volatile u64*>(p); }

// table[slot] = 1 + the smallest record index among the records whose header bytes are the slot's (0 = empty).
__global__ void fq_dedup_insert_kernel(const uint8_t* __restrict__ d, const u64* __restrict__ off, const u64* __restrict__ hash,
                                       const uint32_t* __restrict__ hlen, u64 nrec, u64* table, u64 mask) {
  const u64 k = (u64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (k >= nrec) return;
  for (u64 slot = hash[k] & mask;; slot = (slot + 1) & mask) {
    u64 v = ld_volatile_u64(table + slot);
    if (v == 0) {  // empty: claim it (lane 0), or meet whoever was faster
      if (lane == 0) v = atomicCAS(reinterpret_cast<unsigned long long*>(table + slot), 0ull, (unsigned long long)(k + 1));
      v = __shfl_sync(0xffffffffu, v, 0);
      if (v == 0) return;
    }
    if (same_header(d, off, hlen, k, v - 1, lane)) {  // (the occupant may be lowered meanwhile -- only by records with these bytes)
      if (lane == 0) atomicMin(reinterpret_cast<unsigned long long*>(table + slot), (unsigned long long)(k + 1));
      return;
    }
  }
}

// keep[k] = 1 when record k is the representative (the first in file order) of its header bytes.
__global__ void fq_dedup_mark_kernel(const uint8_t* __restrict__ d, const u64* __restrict__ off, const u64* __restrict__ hash,
                                     const uint32_t* __restrict__ hlen, u64 nrec, const u64* __restrict__ table, u64 mask,
                                     uint8_t* __restrict__ keep, unsigned long long* __restrict__ ndups) {
  const u64 k = (u64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (k >= nrec) return;
  for (u64 slot = hash[k] & mask;; slot = (slot + 1) & mask) {
    const u64 v = table[slot];  // never empty before record k's own slot is met
    if (v == k + 1 || same_header(d, off, hlen, k, v - 1, lane)) {
      if (lane == 0) { keep[k] = v == k + 1 ? 1 : 0; if (v != k + 1) atomicAdd(ndups, 1ull); }
      return;
    }
  }
}

}  // namespace fq

namespace {
// device temporaries of one call, released on every path
struct Temps {
  cudaStream_t st;
  void* p[8];
  int n = 0;
  explicit Temps(cudaStream_t s) : st(s) {}
  template <class T>
  cudaError_t alloc(T** out, size_t bytes) {
    void* q = nullptr;
    cudaError_t e = cudaMallocAsync(&q, bytes ? bytes : 1, st);
    if (e == cudaSuccess) p[n++] = q;
    *out = (T*)q;
    return e;
  }
  ~Temps() { for (int i = 0; i < n; i++) cudaFreeAsync(p[i], st); }
};
}  // namespace

extern "C" int fqgpu_dedup_device(fqgpu_ctx* ctx, const void* dptr, size_t nbytes, const uint64_t* d_offsets, uint64_t n_records,
                                  uint8_t* d_keep, uint64_t* n_dups) {
  if (!ctx || !n_dups || (n_records && (!dptr || !d_offsets || !d_keep))) return FQGPU_EARG;
  *n_dups = 0;
  if (n_records == 0) return FQGPU_OK;
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  Temps tmp(st);
  fq::u64 *d_hash = nullptr, *d_table = nullptr;
  uint32_t* d_len = nullptr;
  unsigned long long* d_nd = nullptr;
  fq::u64 slots = 1024;
  while (slots < 2 * n_records) slots <<= 1;  // load factor <= 1/2
  CU_TRY(ctx, tmp.alloc(&d_hash, n_records * sizeof(fq::u64)));
  CU_TRY(ctx, tmp.alloc(&d_len, n_records * sizeof(uint32_t)));
  CU_TRY(ctx, tmp.alloc(&d_table, slots * sizeof(fq::u64)));
  CU_TRY(ctx, tmp.alloc(&d_nd, sizeof(unsigned long long)));
  CU_TRY(ctx, cudaMemsetAsync(d_nd, 0, sizeof(unsigned long long), st));
  cudaEvent_t e0 = fqgpu_get_event(ctx), e1 = fqgpu_get_event(ctx);
  CU_TRY(ctx, cudaEventRecord(e0, st));
  CU_TRY(ctx, cudaMemsetAsync(d_table, 0, slots * sizeof(fq::u64), st));
  const unsigned grid = (unsigned)((n_records + 7) / 8);
  fq::fq_dedup_hash_kernel<<<grid, 256, 0, st>>>((const uint8_t*)dptr, nbytes, (const fq::u64*)d_offsets, n_records, d_hash, d_len);
  fq::fq_dedup_insert_kernel<<<grid, 256, 0, st>>>((const uint8_t*)dptr, (const fq::u64*)d_offsets, d_hash, d_len, n_records, d_table, slots - 1);
  fq::fq_dedup_mark_kernel<<<grid, 256, 0, st>>>((const uint8_t*)dptr, (const fq::u64*)d_offsets, d_hash, d_len, n_records, d_table, slots - 1, d_keep, d_nd);
  CU_TRY(ctx, cudaGetLastError());
  CU_TRY(ctx, cudaEventRecord(e1, st));
  ctx->timed.emplace_back(e0, e1);
  ctx->launches += 3;
  unsigned long long h = 0;
  CU_TRY(ctx, cudaMemcpyAsync(&h, d_nd, sizeof h, cudaMemcpyDeviceToHost, st));
  CU_TRY(ctx, cudaStreamSynchronize(st));
  *n_dups = h;
  return FQGPU_OK;
}

// The whole pipeline for a FASTQ held in HOST memory (what the `sc fq-dedup` mirrors call): copy to the device,
// record-offset index (ONE pass: the record count is bounded by the bytes, a record has at least ... one byte per line is
// not guaranteed, so the bound is lines/4 <= (bytes + 1 + 3) / 4), duplicate marks, flags back.  h_keep receives
// min(*n_records, cap) flags.  The whole file is held in device memory: inputs beyond the free HBM fail with FQGPU_ENOMEM
// where the reference streams (documented limit).
extern "C" int fqgpu_dedup_host(fqgpu_ctx* ctx, const void* host, size_t nbytes, uint8_t* h_keep, uint64_t cap,
                                uint64_t* n_records, uint64_t* n_lines, uint64_t* n_dups) {
  if (!ctx || !n_records || !n_lines || !n_dups || (nbytes && !host)) return FQGPU_EARG;
  *n_records = *n_lines = *n_dups = 0;
  if (nbytes == 0) return FQGPU_OK;
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  Temps tmp(st);
  uint8_t* d_data = nullptr;
  uint64_t* d_off = nullptr;
  uint8_t* d_keep = nullptr;
  const uint64_t max_rec = (nbytes + 1 + 3) / 4 + 1;  // every line but the last ends with a byte of its own
  if (tmp.alloc(&d_data, nbytes + 16) != cudaSuccess || tmp.alloc(&d_off, max_rec * sizeof(uint64_t)) != cudaSuccess ||
      tmp.alloc(&d_keep, max_rec) != cudaSuccess) {
    cudaGetLastError();
    return fail(ctx, FQGPU_ENOMEM, "fqgpu_dedup_host: the file and its index do not fit into device memory");
  }
  CU_TRY(ctx, cudaMemcpyAsync(d_data, host, nbytes, cudaMemcpyHostToDevice, st));
  uint64_t nrec = 0;
  int rc = fqgpu_index_device(ctx, d_data, nbytes, d_off, max_rec, &nrec);
  if (rc == FQGPU_OK && nrec) {
    rc = fqgpu_dedup_device(ctx, d_data, nbytes, d_off, nrec, d_keep, n_dups);
    if (rc == FQGPU_OK && h_keep && cap) CU_TRY(ctx, cudaMemcpyAsync(h_keep, d_keep, nrec < cap ? nrec : cap, cudaMemcpyDeviceToHost, st));
  }
  CU_TRY(ctx, cudaStreamSynchronize(st));
  *n_records = nrec;
  *n_lines = fqgpu_index_lines(ctx);
  return rc;
}
