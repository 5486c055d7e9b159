// fq_bgzf.h -- on-device BGZF inflate (fq_bgzf.cu), used by the .gz branch of fqgpu_count_file_as.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fq {

// One gzip member of a BGZF file: its DEFLATE payload inside the compressed batch and its place in the output.
struct BgzfMember {
  unsigned long long in_off;   // first payload byte, relative to the batch
  unsigned long long out_off;  // first output byte, relative to the batch's output buffer
  unsigned int in_len;         // payload bytes (without the gzip header and the CRC32 / ISIZE trailer)
  unsigned int out_len;        // ISIZE
  unsigned int crc;            // CRC32 of the output (trailer)
  unsigned int pad;
};

// One warp per member (all lanes in lockstep); status[i] = 0 or the reason member i could not be inflated.
cudaError_t launch_bgzf_inflate(const uint8_t* d_comp, const BgzfMember* d_members, int n, uint8_t* d_out, uint32_t* d_status,
                                cudaStream_t st);

}  // namespace fq
