// fqgpu_ctx.h -- internal definition of fqgpu_ctx shared by fqgpu_api.cu and fq_shard.cu.
#pragma once
#include <string.h>
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <string>
#include <vector>

#include "fq_layout.h"

namespace fq {
typedef unsigned long long u64;
size_t scan_smem_bytes();
int scan_tile_bytes();
int scan_threads();
int scan_ctas_per_sm();
cudaError_t scan_configure();
cudaError_t launch_reset(u64* acc, Carry* carry, u64* ctl, cudaStream_t st);
cudaError_t launch_scan(const void* ptr, size_t nbytes, Carry* carry, ShardInfo* shard, u64* acc, SpanDesc* desc, u64* ctl,
                        bool unknown_start, int resident, bool core_only, cudaStream_t st);
cudaError_t meta_configure();
size_t meta_seg_count(u64 end);
size_t meta_seg_bytes();
cudaError_t launch_meta(const uint8_t* base, uint32_t lo0, u64 end, Carry* carry, u64 meta_records, void* segs, u64* start, cudaStream_t st);
cudaError_t launch_synth_illumina(void* dptr, u64 first_byte, u64 nbytes, u64 seed, cudaStream_t st);
cudaError_t launch_synth_illumina_tally(u64 first_record, u64 n_records, u64 seed, u64* d_out, cudaStream_t st);
void synth_illumina_meta_range(u64 first_record, u64 m, u64 seed, long long* qmin, long long* qmax);
cudaError_t synth_ont(void* dptr, size_t capacity, u64 first_record, u64 n_records, u64 seed, size_t* bytes_written,
                      cudaStream_t st);
}  // namespace fq

using fq::u64;

// one launch: the tile index is 32-bit
static const size_t kMaxLaunchBytes = (size_t)64 << 30;
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
  int grid = 0;                // CTAs of a scan launch resident at once (2 per SM)
  u64* d_acc = nullptr;        // [BLOCK_WORDS] counter block accumulated since the last reset
  fq::SpanDesc* d_desc = nullptr;  // [MAX_SPANS] span descriptors of the launch in flight
  u64* d_ctl = nullptr;        // [CTL_WORDS] done counter and launch scratch
  fq::Carry* d_carry = nullptr;
  u64* h_out = nullptr;  // pinned: counter block followed by the carry
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
  // the collective inside the library (fq_shard.cu): exchange buffer, the peers' buffers, step counter
  void* x_buf = nullptr;
  void* x_peers[64] = {};
  bool x_peer_ipc[64] = {};
  int x_world = 0, x_rank = 0;
  u64 x_step = 0;
  u64* h_xres = nullptr;       // pinned: combined block + first wrong rank + end carry
  // on-device BGZF inflate (fq_bgzf.cu): compressed batch (pinned host + device), inflated bytes, member table
  // fq-meta fold (fq_meta.cu): per-segment results of a launch, and where the sequential walk takes over
  void* d_metaseg = nullptr;
  size_t metaseg_cap = 0;      // segments
  u64* d_metastart = nullptr;
  // two slots, so that the file read and the H2D copy of the next batch run under the kernels of this one
  struct CompSlot {
    uint8_t* h = nullptr;          // pinned: compressed bytes as read from the file
    uint8_t* d = nullptr;          // the same on the device
    void* h_members = nullptr;     // pinned: BGZF member table of the batch
    void* d_members = nullptr;
    uint32_t* h_status = nullptr;  // pinned: per-member outcome
    uint32_t* d_status = nullptr;
    cudaEvent_t h2d = nullptr;     // the slot's bytes are on the device
    cudaEvent_t done = nullptr;    // the kernels that read the slot have finished
  } comp[2];
  size_t comp_cap = 0;
  size_t members_cap = 0;
  uint8_t* d_inflated = nullptr;
  size_t inflated_cap = 0;
  u64 bgzf_members = 0;        // members inflated on the device since the last reset (diagnostics)
  // on-device inflate of ordinary gzip (fq_gzip.cu): chunk table, chain, 16-bit symbols, windows
  void* d_gzchunks = nullptr;
  uint32_t* d_gzorder = nullptr;
  u64* d_gzcoff = nullptr;
  u64* d_gzsb = nullptr;       // one-pass mode: where each chain chunk's symbols are in the arena
  void* d_gzres = nullptr;     // GzResult + the write pass's error word
  void* h_gzres = nullptr;     // pinned copy
  uint16_t* d_gzsym = nullptr;
  size_t gzsym_cap = 0;        // symbols
  uint8_t* d_gzwbuf = nullptr;
  size_t gzwbuf_cap = 0;       // rows of 32 KiB
  uint32_t* d_gzraw = nullptr;   // CRC-32 registers of the output's 4 KiB slices
  size_t gzraw_cap = 0;
  uint8_t* d_gzwindow = nullptr;
  u64 gzip_chunks = 0;         // chunks of single-member gzip inflated on the device since the last reset (diagnostics)
  u64 gzip_rewrites = 0;       // batches that needed the second decode pass after all
  u64 gzip_passed = 0;         // block starts the search found that turned out not to be block boundaries
};


#define CU_TRY(ctx, call)                                                              \
  do {                                                                                 \
    cudaError_t e_ = (call);                                                           \
    if (e_ != cudaSuccess) {                                                           \
      (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                 \
      return FQGPU_ECUDA;                                                              \
    }                                                                                  \
  } while (0)

// NVTX range over a C-ABI call (nsys / ncu timelines: fill, submit, scan, finish, exchange)
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

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
