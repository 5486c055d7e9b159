// fq_gzip.h -- on-device inflate of ordinary (single-member) gzip input (fq_gzip.cu), used by the .gz branch of
// fqgpu_count_file_as when the file is not BGZF.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fq {

constexpr uint32_t GZ_WINDOW = 32768;            // DEFLATE's look-back
constexpr int GZ_MAX_CHUNKS = 16384;             // chunks (= decoding warps) per batch of compressed bytes
constexpr unsigned long long GZ_NONE = ~0ull;    // "no block start found in this chunk"

enum : uint32_t {
  GZC_FINAL = 1u,       // the chunk ended with the stream's last block
  GZC_INCOMPLETE = 2u,  // the compressed batch ended inside a block: the chunk stops at the last complete block
  GZC_ERROR = 4u,       // the bits are not DEFLATE as zlib accepts it
  GZC_GIVEUP = 8u,      // too many block starts passed without landing on one / output too large for one chunk
  GZC_CHAIN = 16u,      // the chunk is part of the verified chain that starts at the batch's known first block
  GZC_REWRITE = 32u,    // (one-pass mode) the chunk's symbols did not fit its slot of the arena: the batch needs the second pass
};

// One chunk of the compressed batch: the first DEFLATE block header found at or after the chunk's first bit, and
// where decoding from there came to rest.
struct GzChunk {
  unsigned long long start_bit;  // GZ_NONE: no start found (the chunk before decodes through)
  unsigned long long end_bit;    // the block boundary the decode stopped at
  unsigned long long out_off;    // first output byte of the chunk within the batch (chain chunks)
  uint32_t out_len;              // bytes the chunk's blocks inflate to
  uint32_t flags;                // GZC_*
  uint32_t need;                 // farthest reach of a match behind the chunk's first output byte
  int32_t land;                  // the chunk whose start_bit the decode landed on (-1: none)
};

enum : uint32_t { GZR_OK = 0, GZR_BROKEN = 1 };
struct GzResult {
  uint32_t status;               // GZR_*
  uint32_t nchain;               // chunks on the chain
  unsigned long long total_out;  // bytes the chain inflates to
  unsigned long long end_bit;    // the bit after the last complete block
  uint32_t final_block;          // the stream's last block is inside the batch
  uint32_t passed;               // starts found by the search that were not block boundaries
  uint32_t rewrite;              // (one-pass mode) some chunk of the chain could not keep its symbols: second pass needed
  uint32_t pad;
};

// Once per device, before the first launch.
cudaError_t gz_configure();

// All kernels run on `st`; positions are bits relative to the first byte of d_comp (4-byte aligned, the bytes
// behind nbytes up to the next word zeroed).  `window` = the 32 KiB before the batch's first block, right-aligned,
// of which the last `wvalid` bytes exist; launch_gz_windows replaces it with the window behind the batch.
cudaError_t launch_gz_sync(const uint8_t* d_comp, size_t nbytes, uint32_t chunk_bytes, int nchunks, unsigned long long start_bit,
                           GzChunk* chunks, uint32_t* nfound, cudaStream_t st);
cudaError_t launch_gz_count(const uint8_t* d_comp, size_t nbytes, uint32_t chunk_bytes, GzChunk* chunks, int nchunks, uint32_t wvalid,
                            cudaStream_t st);
// One-pass mode: count and write at once, chunk c's symbols to arena + c * stride (stride = symbols per chunk slot).
cudaError_t launch_gz_both(const uint8_t* d_comp, size_t nbytes, uint32_t chunk_bytes, GzChunk* chunks, int nchunks, uint16_t* arena,
                           unsigned long long stride, const uint8_t* window, uint32_t wvalid, cudaStream_t st);
// stride != 0 (one-pass mode): sbase[i] = where chain chunk i's symbols are in the arena, minus its output offset
cudaError_t launch_gz_chain(GzChunk* chunks, int nchunks, unsigned long long prior_out, GzResult* res, uint32_t* order,
                            unsigned long long* coff, unsigned long long stride, unsigned long long* sbase, cudaStream_t st);
cudaError_t launch_gz_write(const uint8_t* d_comp, size_t nbytes, uint32_t chunk_bytes, GzChunk* chunks, int nchunks, uint16_t* markers,
                            const uint8_t* window, uint32_t wvalid, uint32_t* err, cudaStream_t st);
// The windows: K = gz_group_chunks(nchain, SMs) chunks per group, ngroups = ceil(nchain / K).
// symrows: (nchain + 1) rows of 32768 symbols; grows: ngroups rows of 32768 symbols; trows: (ngroups + 1) rows of 32 KiB.
uint32_t gz_group_chunks(uint32_t nchain, int sms);
// sbase: nullptr = the symbols of output position p are at markers[p] (second pass); else at markers[sbase[i] + p] (arena)
cudaError_t launch_gz_windows(const unsigned long long* coff, const unsigned long long* sbase, uint32_t nchain, uint32_t K, const uint16_t* markers,
                              uint16_t* symrows, uint16_t* grows, uint8_t* trows, uint8_t* window, cudaStream_t st);
cudaError_t launch_gz_resolve(const uint16_t* markers, const unsigned long long* sbase, const uint16_t* symrows, const uint8_t* trows, uint32_t K,
                              const unsigned long long* coff, uint32_t nchain, unsigned long long total_out, uint8_t* out, int sms, cudaStream_t st);

// CRC-32 pieces of the batch's output: d_crc2[0] = the register (from zero) over the full 4 KiB slices, d_crc2[1] = over
// the bytes behind them.  d_raw: total / 4096 + 1 words; xs = x^(8 * 4096), xq = x^(8 * 4096 * q) modulo the CRC
// polynomial, q = ceil(slices / 1024).
constexpr uint32_t GZ_CRC_SLICE = 4096;
cudaError_t launch_gz_crc(const uint8_t* d_out, unsigned long long total, uint32_t* d_raw, uint32_t xs, uint32_t xq, uint32_t q, uint32_t* d_crc2,
                          cudaStream_t st);

}  // namespace fq
