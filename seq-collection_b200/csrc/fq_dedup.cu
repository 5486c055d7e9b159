// fq_dedup.cu -- duplicate read-ID detection (src/fq_dedup.nim:14-84).
//
// The reference keeps the first record of every distinct header line (line 4k, compared as the string Nim's
// `lines` yields: without '\n' and without one '\r' directly before it) and drops every later record with the
// same header; it finds them with a Bloom filter and two passes over the file.  Here: one warp per record hashes
// its header line (64-bit polynomial hash), the (hash, record) pairs are sorted (stable: records of one hash stay
// in file order), and every record compares its header BYTES with the earlier records of its hash run -- so the
// result is exact whatever the hash does.  Input: the record-offset index of fq_index.cu.
#include <cuda_runtime.h>
#include <stdint.h>
#include <thrust/execution_policy.h>
#include <thrust/sequence.h>
#include <thrust/sort.h>

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

// keep[k] = 0 when an earlier record (in file order) has the same header bytes.
__global__ void fq_dedup_mark_kernel(const uint8_t* __restrict__ d, const u64* __restrict__ off, const u64* __restrict__ shash,
                                     const u64* __restrict__ sidx, const uint32_t* __restrict__ hlen, u64 nrec,
                                     uint8_t* __restrict__ keep, unsigned long long* __restrict__ ndups) {
  const u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nrec) return;
  const u64 k = sidx[j], h = shash[j];
  const uint32_t L = hlen[k];
  const uint8_t* a = d + off[k];
  bool dup = false;
  for (u64 x = j; x > 0 && shash[x - 1] == h && !dup; x--) {  // the earlier records of this hash run
    const u64 k2 = sidx[x - 1];
    if (hlen[k2] != L) continue;
    const uint8_t* b = d + off[k2];
    bool same = true;
    for (uint32_t i = 0; i < L && same; i++) same = a[i] == b[i];
    dup = same;
  }
  keep[k] = dup ? 0 : 1;
  if (dup) atomicAdd(ndups, 1ull);
}

}  // namespace fq

extern "C" int fqgpu_dedup_device(fqgpu_ctx* ctx, const void* dptr, size_t nbytes, const uint64_t* d_offsets, uint64_t n_records,
                                  uint8_t* d_keep, uint64_t* n_dups) {
  if (!ctx || !n_dups || (n_records && (!dptr || !d_offsets || !d_keep))) return FQGPU_EARG;
  *n_dups = 0;
  if (n_records == 0) return FQGPU_OK;
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  fq::u64 *d_hash = nullptr, *d_idx = nullptr;
  uint32_t* d_len = nullptr;
  unsigned long long* d_nd = nullptr;
  CU_TRY(ctx, cudaMallocAsync((void**)&d_hash, n_records * sizeof(fq::u64), st));
  CU_TRY(ctx, cudaMallocAsync((void**)&d_idx, n_records * sizeof(fq::u64), st));
  CU_TRY(ctx, cudaMallocAsync((void**)&d_len, n_records * sizeof(uint32_t), st));
  CU_TRY(ctx, cudaMallocAsync((void**)&d_nd, sizeof(unsigned long long), st));
  CU_TRY(ctx, cudaMemsetAsync(d_nd, 0, sizeof(unsigned long long), st));
  cudaEvent_t e0 = fqgpu_get_event(ctx), e1 = fqgpu_get_event(ctx);
  CU_TRY(ctx, cudaEventRecord(e0, st));
  fq::fq_dedup_hash_kernel<<<(unsigned)((n_records + 7) / 8), 256, 0, st>>>((const uint8_t*)dptr, nbytes, (const fq::u64*)d_offsets, n_records, d_hash, d_len);
  CU_TRY(ctx, cudaGetLastError());
  thrust::sequence(thrust::cuda::par.on(st), d_idx, d_idx + n_records);
  thrust::stable_sort_by_key(thrust::cuda::par.on(st), d_hash, d_hash + n_records, d_idx);
  fq::fq_dedup_mark_kernel<<<(unsigned)((n_records + 255) / 256), 256, 0, st>>>((const uint8_t*)dptr, (const fq::u64*)d_offsets, d_hash, d_idx, d_len,
                                                                             n_records, d_keep, d_nd);
  CU_TRY(ctx, cudaGetLastError());
  CU_TRY(ctx, cudaEventRecord(e1, st));
  ctx->timed.emplace_back(e0, e1);
  unsigned long long h = 0;
  CU_TRY(ctx, cudaMemcpyAsync(&h, d_nd, sizeof h, cudaMemcpyDeviceToHost, st));
  CU_TRY(ctx, cudaFreeAsync(d_hash, st));
  CU_TRY(ctx, cudaFreeAsync(d_idx, st));
  CU_TRY(ctx, cudaFreeAsync(d_len, st));
  CU_TRY(ctx, cudaFreeAsync(d_nd, st));
  CU_TRY(ctx, cudaStreamSynchronize(st));
  *n_dups = h;
  return FQGPU_OK;
}

// The whole pipeline for a FASTQ held in HOST memory (what the `sc fq-dedup` mirrors call): copy to the device,
// record-offset index, duplicate marks, flags back.  h_keep receives min(*n_records, cap) flags.
extern "C" int fqgpu_dedup_host(fqgpu_ctx* ctx, const void* host, size_t nbytes, uint8_t* h_keep, uint64_t cap,
                                uint64_t* n_records, uint64_t* n_lines, uint64_t* n_dups) {
  if (!ctx || !n_records || !n_lines || !n_dups || (nbytes && !host)) return FQGPU_EARG;
  *n_records = *n_lines = *n_dups = 0;
  if (nbytes == 0) return FQGPU_OK;
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  uint8_t* d_data = nullptr;
  CU_TRY(ctx, cudaMallocAsync((void**)&d_data, nbytes + 16, st));
  CU_TRY(ctx, cudaMemcpyAsync(d_data, host, nbytes, cudaMemcpyHostToDevice, st));
  uint64_t nrec = 0;
  int rc = fqgpu_index_device(ctx, d_data, nbytes, nullptr, 0, &nrec);  // count first: sizes the index
  uint64_t* d_off = nullptr;
  uint8_t* d_keep = nullptr;
  if (rc == FQGPU_OK && nrec) {
    CU_TRY(ctx, cudaMallocAsync((void**)&d_off, nrec * sizeof(uint64_t), st));
    CU_TRY(ctx, cudaMallocAsync((void**)&d_keep, nrec, st));
    rc = fqgpu_index_device(ctx, d_data, nbytes, d_off, nrec, &nrec);
    if (rc == FQGPU_OK) rc = fqgpu_dedup_device(ctx, d_data, nbytes, d_off, nrec, d_keep, n_dups);
    if (rc == FQGPU_OK && h_keep && cap) CU_TRY(ctx, cudaMemcpyAsync(h_keep, d_keep, nrec < cap ? nrec : cap, cudaMemcpyDeviceToHost, st));
    CU_TRY(ctx, cudaFreeAsync(d_off, st));
    CU_TRY(ctx, cudaFreeAsync(d_keep, st));
  }
  CU_TRY(ctx, cudaFreeAsync(d_data, st));
  CU_TRY(ctx, cudaStreamSynchronize(st));
  *n_records = nrec;
  *n_lines = fqgpu_index_lines(ctx);
  return rc;
}
