// fqgpu_ctx.h -- internal definition of fqgpu_ctx shared by fqgpu_api.cu and fq_shard.cu.
#pragma once
#include <string.h>
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "fq_layout.h"

namespace fq {
typedef unsigned long long u64;
size_t scan_smem_bytes();
int scan_tile_bytes();
int scan_threads();
cudaError_t scan_configure();
cudaError_t launch_reset(u64* committed, int nblocks, Carry* carry, cudaStream_t st);
cudaError_t launch_scan(const void* ptr, size_t nbytes, SpanDesc* desc, LaunchHdr* hdr, Carry* carry,
                        u64* pending, u64* committed, ShardInfo* shard, int max_spans, u64 meta_records, cudaStream_t st,
                        cudaStream_t meta_stream, cudaEvent_t ev_fork, cudaEvent_t ev_join, bool core_only);
cudaError_t launch_reduce(const u64* blocks, int nblocks, u64* out, cudaStream_t st);
uint32_t scan_span_count(u64 bytes, int resident);
cudaError_t launch_synth_illumina(void* dptr, u64 first_byte, u64 nbytes, u64 seed, cudaStream_t st);
cudaError_t synth_ont(void* dptr, size_t capacity, u64 first_record, u64 n_records, u64 seed, size_t* bytes_written,
                      cudaStream_t st);
}  // namespace fq

using fq::u64;

// one launch = up to MAX_SPANS spans; a span stays below 2 GiB so its 32-bit shared-memory counters are exact
static const size_t kMaxLaunchBytes = (size_t)fq::BASE_SPANS * ((size_t)2 << 30) - ((size_t)64 << 20);
inline thread_local std::string g_create_error;

struct StageBuf {
  void* host = nullptr;
  cudaEvent_t copied = nullptr;  // H2D of this buffer finished -> host side reusable
  bool in_flight = false;
};

struct fqgpu_ctx {
  int device = 0;
  fqgpu_config cfg{};
  cudaStream_t stream = nullptr;
  cudaStream_t mstream = nullptr;                  // the fq-meta prefix kernel runs beside the scan
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int grid = 0;
  int span_hwm = fq::BASE_SPANS;  // most span blocks any launch of this context has used: reset / reduce cover [0, span_hwm)
  u64* d_pending = nullptr;    // [MAX_SPANS] blocks of the launch in flight
  u64* d_committed = nullptr;  // [MAX_SPANS] blocks accumulated since the last reset
  fq::Carry* d_carry = nullptr;
  fq::SpanDesc* d_desc = nullptr;
  fq::LaunchHdr* d_hdr = nullptr;
  u64* d_out = nullptr;
  u64* h_out = nullptr;  // pinned: reduced block followed by the carry
  std::vector<StageBuf> ring;
  // device landing buffers of the host paths: H2D of chunk k+1 (copy stream) overlaps the scan of
  // chunk k (compute stream)
  void* d_stage[2] = {nullptr, nullptr};
  cudaStream_t cstream = nullptr;
  cudaEvent_t ev_copied[2] = {nullptr, nullptr};   // H2D into d_stage[b] finished
  cudaEvent_t ev_scanned[2] = {nullptr, nullptr};  // scan of d_stage[b] finished -> buffer free
  u64 n_staged = 0;
  size_t chunk_bytes = 0;
  int next = 0;
  std::string err;
  // timing of scan launches since the last reset
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timed;
  std::vector<cudaEvent_t> event_pool;
  double kernel_ms_done = 0.0;
  unsigned long long index_lines = 0;  // lines of the buffer of the most recent fqgpu_index_device call
  u64 launches = 0;
  // shard mode
  int shard_rank = 0, shard_world = 1;
  bool shard_exact = false;   // this rank was rescanned with the exact carry
  fq::ShardInfo* d_shard = nullptr;
  u64* h_shard = nullptr;      // pinned: gathered shard blocks (up to 64 ranks)
  // on-device BGZF inflate (fq_bgzf.cu): compressed batch (pinned host + device), inflated bytes, member table
  uint8_t* h_comp = nullptr;
  uint8_t* d_comp = nullptr;
  size_t comp_cap = 0;
  uint8_t* d_inflated = nullptr;
  size_t inflated_cap = 0;
  void* d_members = nullptr;
  uint32_t* d_mstatus = nullptr;
  size_t members_cap = 0;
  u64 bgzf_members = 0;        // members inflated on the device since the last reset (diagnostics)
};


#define CU_TRY(ctx, call)                                                              \
  do {                                                                                 \
    cudaError_t e_ = (call);                                                           \
    if (e_ != cudaSuccess) {                                                           \
      (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                 \
      return FQGPU_ECUDA;                                                              \
    }                                                                                  \
  } while (0)

inline int fail(fqgpu_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->err = msg; else g_create_error = msg;
  return code;
}

// FQGPU_F_CORE_ONLY: quality lines are not examined; their outputs are defined as zero.
static inline void fqgpu_zero_quality(fqgpu_stats* st) {
  memset(st->qual_counts, 0, sizeof(st->qual_counts));
  memset(st->qual_len_hist, 0, sizeof(st->qual_len_hist));
  memset(st->qual_pos_sum, 0, sizeof(st->qual_pos_sum));
  memset(st->qual_pos_cnt, 0, sizeof(st->qual_pos_cnt));
  st->qual_lines = 0; st->qual_len_min = 0; st->qual_len_max = 0;
}
void fqgpu_assemble_stats(const fq::u64* blk, const fq::Carry& c, fq::u64 meta_records, fqgpu_stats* st, bool core_only = false);
cudaEvent_t fqgpu_get_event(fqgpu_ctx* ctx);
