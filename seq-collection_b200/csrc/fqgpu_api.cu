// fqgpu_api.cu -- host side of libfqgpu: contexts, the pinned staging ring, launch orchestration
// and the assembly of fqgpu_stats from the reduced counter block + stream carry.
// The C ABI is declared in include/fqgpu.h (each entry point cites the reference code it replaces).
#include <cuda_runtime.h>
#include <fcntl.h>
#include <stdio.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <atomic>
#include <chrono>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "fq_bgzf.h"
#include "fq_gzip.h"
#include "fq_layout.h"

#include "fqgpu_ctx.h"

extern "C" void fqgpu_shard_exchange_destroy(fqgpu_ctx* ctx);

extern "C" {

int fqgpu_abi_version(void) { return FQGPU_ABI_VERSION; }
size_t fqgpu_stats_size(void) { return sizeof(fqgpu_stats); }
const char* fqgpu_build_info(void) {
  return "libfqgpu sm_100a; scan tile " "32 KiB" "; built " __DATE__;
}
int fqgpu_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

const char* fqgpu_last_error(const fqgpu_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

void fqgpu_destroy(fqgpu_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  if (ctx->cstream) cudaStreamSynchronize(ctx->cstream);
  for (auto& b : ctx->ring) {
    if (b.host) cudaFreeHost(b.host);
    if (b.copied) cudaEventDestroy(b.copied);
  }
  for (auto& p : ctx->timed) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
  for (auto e : ctx->event_pool) cudaEventDestroy(e);
  for (int b = 0; b < 2; b++) {
    cudaFree(ctx->d_stage[b]);
    if (ctx->ev_copied[b]) cudaEventDestroy(ctx->ev_copied[b]);
    if (ctx->ev_scanned[b]) cudaEventDestroy(ctx->ev_scanned[b]);
  }
  if (ctx->cstream) cudaStreamDestroy(ctx->cstream);
  cudaFree(ctx->d_acc);
  cudaFree(ctx->d_desc);
  cudaFree(ctx->d_ctl);
  cudaFree(ctx->d_carry);
  cudaFree(ctx->d_shard);
  fqgpu_shard_exchange_destroy(ctx);
  if (ctx->h_xres) cudaFreeHost(ctx->h_xres);
  if (ctx->h_shard) cudaFreeHost(ctx->h_shard);
  for (auto& c : ctx->comp) {
    if (c.h) cudaFreeHost(c.h);
    if (c.h_members) cudaFreeHost(c.h_members);
    if (c.h_status) cudaFreeHost(c.h_status);
    cudaFree(c.d);
    cudaFree(c.d_members);
    cudaFree(c.d_status);
    if (c.h2d) cudaEventDestroy(c.h2d);
    if (c.done) cudaEventDestroy(c.done);
  }
  cudaFree(ctx->d_inflated);
  cudaFree(ctx->d_metaseg);
  cudaFree(ctx->d_metastart);
  cudaFree(ctx->d_gzchunks);
  cudaFree(ctx->d_gzorder);
  cudaFree(ctx->d_gzcoff);
  cudaFree(ctx->d_gzsb);
  cudaFree(ctx->d_gzres);
  if (ctx->h_gzres) cudaFreeHost(ctx->h_gzres);
  cudaFree(ctx->d_gzsym);
  cudaFree(ctx->d_gzwbuf);
  cudaFree(ctx->d_gzraw);
  cudaFree(ctx->d_gzwindow);
  if (ctx->h_out) cudaFreeHost(ctx->h_out);
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
  if (ctx->mstream) cudaStreamDestroy(ctx->mstream);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

int fqgpu_create(fqgpu_ctx** out, const fqgpu_config* cfg) {
  if (!out) return fail(nullptr, FQGPU_EARG, "fqgpu_create: out is NULL");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0) {
    cudaGetLastError();
    return fail(nullptr, FQGPU_ECUDA,
                std::string("no usable CUDA device (libfqgpu has no CPU fallback): ") + cudaGetErrorString(e));
  }
  fqgpu_ctx* ctx = new fqgpu_ctx();
  if (cfg) ctx->cfg = *cfg;
  else ctx->cfg.device = -1;
  int dev = ctx->cfg.device;
  if (dev < 0) { if (cudaGetDevice(&dev) != cudaSuccess) dev = 0; }
  if (dev >= ndev) { delete ctx; return fail(nullptr, FQGPU_EARG, "fqgpu_create: device ordinal out of range"); }
  ctx->device = dev;
  ctx->chunk_bytes = ctx->cfg.chunk_bytes ? ctx->cfg.chunk_bytes : ((size_t)64 << 20);
  ctx->chunk_bytes = (ctx->chunk_bytes + 4095) & ~(size_t)4095;
  int nbuf = ctx->cfg.n_buffers > 0 ? ctx->cfg.n_buffers : 3;
  auto bail = [&](const std::string& m) { g_create_error = m; fqgpu_destroy(ctx); return FQGPU_ECUDA; };
#define CU_NEW(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return bail(std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)
  CU_NEW(cudaSetDevice(dev));
  cudaDeviceProp prop;
  CU_NEW(cudaGetDeviceProperties(&prop, dev));
  if (prop.major < 10) return bail("libfqgpu is built for sm_100a (B200); device is sm_" + std::to_string(prop.major * 10 + prop.minor));
  CU_NEW(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  CU_NEW(cudaStreamCreateWithFlags(&ctx->mstream, cudaStreamNonBlocking));
  CU_NEW(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
  CU_NEW(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
  CU_NEW(fq::scan_configure());
  CU_NEW(fq::meta_configure());
  CU_NEW(fq::gz_configure());
  ctx->grid = prop.multiProcessorCount * fq::scan_ctas_per_sm();
  if (ctx->grid > fq::RESIDENT_CTAS) ctx->grid = fq::RESIDENT_CTAS;
  CU_NEW(cudaMalloc(&ctx->d_acc, fq::BLOCK_WORDS * sizeof(u64)));
  CU_NEW(cudaMalloc(&ctx->d_ctl, fq::CTL_WORDS * sizeof(u64)));
  CU_NEW(cudaMalloc(&ctx->d_desc, fq::MAX_SPANS * sizeof(fq::SpanDesc)));
  CU_NEW(cudaMalloc(&ctx->d_carry, sizeof(fq::Carry)));
  CU_NEW(cudaMalloc(&ctx->d_shard, sizeof(fq::ShardInfo)));
  CU_NEW(cudaMallocHost(&ctx->h_shard, (size_t)64 * fqgpu_shard_block_words() * sizeof(u64)));
  CU_NEW(cudaMallocHost(&ctx->h_out, fq::BLOCK_WORDS * sizeof(u64) + sizeof(fq::Carry) + sizeof(u64)));
  ctx->ring.resize(nbuf);  // pinned chunks are allocated lazily by the first acquire()
#undef CU_NEW
  *out = ctx;
  int rc = fqgpu_reset(ctx);
  if (rc != FQGPU_OK) { g_create_error = ctx->err; fqgpu_destroy(ctx); *out = nullptr; return rc; }
  return FQGPU_OK;
}

int fqgpu_reset(fqgpu_ctx* ctx) {
  if (!ctx) return FQGPU_EARG;
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  CU_TRY(ctx, fq::launch_reset(ctx->d_acc, ctx->d_carry, ctx->d_ctl, ctx->stream));
  for (auto& p : ctx->timed) { ctx->event_pool.push_back(p.first); ctx->event_pool.push_back(p.second); }
  ctx->timed.clear();
  ctx->kernel_ms_done = 0.0;
  ctx->launches = 0;
  ctx->bgzf_members = 0;
  ctx->gzip_chunks = 0;
  ctx->gzip_passed = 0;
  ctx->gzip_rewrites = 0;
  ctx->shard_rank = 0;
  ctx->shard_world = 1;
  ctx->shard_exact = false;
  return FQGPU_OK;
}

}  // extern "C"

cudaEvent_t fqgpu_get_event(fqgpu_ctx* ctx) {
  if (!ctx->event_pool.empty()) { cudaEvent_t e = ctx->event_pool.back(); ctx->event_pool.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}

extern "C" {

// The scan of [p, p + n) as the continuation of the context's stream (two launches: every span, then the spans whose
// guessed line phase was wrong -- none on well-formed input); the fq-meta prefix fold (which touches only its own fields
// of the carry) runs beside it on its own stream.
static int fqgpu_launch_scan(fqgpu_ctx* ctx, const uint8_t* p, size_t n, u64 meta_records) {
  if (meta_records) {
    const uintptr_t addr = (uintptr_t)p;
    const uint32_t lo0 = (uint32_t)(addr & 15);
    const size_t nseg = fq::meta_seg_count((u64)lo0 + n);
    if (ctx->metaseg_cap < nseg) {
      CU_TRY(ctx, cudaStreamSynchronize(ctx->mstream));
      cudaFree(ctx->d_metaseg);
      ctx->d_metaseg = nullptr; ctx->metaseg_cap = 0;
      const size_t cap = nseg + (nseg >> 2) + 64;
      CU_TRY(ctx, cudaMalloc(&ctx->d_metaseg, cap * fq::meta_seg_bytes()));
      ctx->metaseg_cap = cap;
    }
    if (!ctx->d_metastart) CU_TRY(ctx, cudaMalloc(&ctx->d_metastart, sizeof(u64)));
    CU_TRY(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
    CU_TRY(ctx, cudaStreamWaitEvent(ctx->mstream, ctx->ev_fork, 0));
    CU_TRY(ctx, fq::launch_meta((const uint8_t*)(addr - lo0), lo0, (u64)lo0 + n, ctx->d_carry, meta_records, ctx->d_metaseg, ctx->d_metastart, ctx->mstream));
    CU_TRY(ctx, cudaEventRecord(ctx->ev_join, ctx->mstream));
  }
  CU_TRY(ctx, fq::launch_scan(p, n, ctx->d_carry, ctx->d_shard, ctx->d_acc, ctx->d_desc, ctx->d_ctl,
                              ctx->shard_rank > 0 && !ctx->shard_exact, ctx->grid, (ctx->cfg.flags & FQGPU_F_CORE_ONLY) != 0,
                              ctx->stream));
  if (meta_records) CU_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
  ctx->launches += 2;
  return FQGPU_OK;
}

int fqgpu_scan_device(fqgpu_ctx* ctx, const void* dptr, size_t nbytes) {
  NvtxRange nvtx_("fqgpu_scan_device");
  if (!ctx) return FQGPU_EARG;
  if (nbytes == 0) return FQGPU_OK;
  if (!dptr) return fail(ctx, FQGPU_EARG, "fqgpu_scan_device: NULL pointer");
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  cudaEvent_t e0 = fqgpu_get_event(ctx), e1 = fqgpu_get_event(ctx);
  CU_TRY(ctx, cudaEventRecord(e0, ctx->stream));
  const uint8_t* p = (const uint8_t*)dptr;
  size_t left = nbytes;
  const u64 meta_records = ctx->shard_rank > 0 ? 0 : ctx->cfg.meta_records;
  while (left) {
    size_t n = left < kMaxLaunchBytes ? left : kMaxLaunchBytes;
    int rc = fqgpu_launch_scan(ctx, p, n, meta_records);
    if (rc != FQGPU_OK) return rc;
    p += n;
    left -= n;
  }
  CU_TRY(ctx, cudaEventRecord(e1, ctx->stream));
  ctx->timed.emplace_back(e0, e1);
  return FQGPU_OK;
}

void* fqgpu_stream(fqgpu_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int fqgpu_last_timing(fqgpu_ctx* ctx, double* kernel_ms, uint64_t* launches) {
  if (!ctx) return FQGPU_EARG;
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  for (auto& p : ctx->timed) {
    float ms = 0.f;
    CU_TRY(ctx, cudaEventElapsedTime(&ms, p.first, p.second));
    ctx->kernel_ms_done += ms;
    ctx->event_pool.push_back(p.first);
    ctx->event_pool.push_back(p.second);
  }
  ctx->timed.clear();
  if (kernel_ms) *kernel_ms = ctx->kernel_ms_done;
  if (launches) *launches = ctx->launches;
  return FQGPU_OK;
}

}  // extern "C"

// ---- stats assembly (host): the reduced block + the stream carry -> fqgpu_stats ----------------
static inline unsigned log2_bin_host(u64 len) { unsigned b = 0; while (len) { b++; len >>= 1; } return b; }

void fqgpu_assemble_stats(const u64* blk, const fq::Carry& c, u64 meta_records, fqgpu_stats* st, bool core_only) {
  memset(st, 0, sizeof(*st));
  for (int i = 0; i < 256; i++) { st->base_counts[i] = blk[fq::OFF_HIST_SEQ + i]; st->qual_counts[i] = blk[fq::OFF_HIST_QUAL + i]; }
  for (int i = 0; i <= fq::POS_BINS; i++) {
    st->seq_len_hist[i] = blk[fq::OFF_SEQ_LEN + i];
    st->qual_len_hist[i] = blk[fq::OFF_QUAL_LEN + i];
    st->qual_pos_sum[i] = blk[fq::OFF_POS_SUM + i];
  }
  for (int i = 0; i < fq::LOG2_BINS; i++) st->seq_len_log2[i] = blk[fq::OFF_SEQ_LOG2 + i];
  st->seq_len_min = blk[fq::OFF_SEQ_LEN_MIN]; st->seq_len_max = blk[fq::OFF_SEQ_LEN_MAX];
  st->qual_len_min = blk[fq::OFF_QUAL_LEN_MIN]; st->qual_len_max = blk[fq::OFF_QUAL_LEN_MAX];
  st->bytes = c.bytes;
  st->lines = c.lines;
  // trailing unterminated, non-empty line (Nim `lines` yields it as is; its '\r', if any, stays)
  const bool tail_cr = c.bytes && c.open_len && c.last_byte == '\r';
  if (c.open_len) {
    const int cls = (int)(c.lines & 3);
    const u64 len = c.open_len;
    st->lines++;
    if (cls == 1) {
      if (tail_cr) st->base_counts['\r']++;
      st->seq_len_hist[len < (u64)fq::POS_BINS ? len : (u64)fq::POS_BINS]++;
      st->seq_len_log2[log2_bin_host(len)]++;
      if (len < st->seq_len_min) st->seq_len_min = len;
      if (len > st->seq_len_max) st->seq_len_max = len;
    } else if (cls == 3) {
      if (tail_cr) {
        st->qual_counts['\r']++;
        st->qual_pos_sum[(len - 1) < (u64)fq::POS_BINS ? (len - 1) : (u64)fq::POS_BINS] += '\r';
      }
      st->qual_len_hist[len < (u64)fq::POS_BINS ? len : (u64)fq::POS_BINS]++;
      if (len < st->qual_len_min) st->qual_len_min = len;
      if (len > st->qual_len_max) st->qual_len_max = len;
    }
  }
  st->reads = (st->lines + 3) / 4;  // lines with i mod 4 == 1
  u64 bases = 0, quals = 0;
  for (int i = 0; i < 256; i++) { bases += st->base_counts[i]; quals += st->qual_counts[i]; }
  st->bases = bases;
  st->gc_bases = st->base_counts['G'] + st->base_counts['C'];
  st->n_bases = st->base_counts['N'];
  u64 sl = 0, ql = 0;
  for (int i = 0; i <= fq::POS_BINS; i++) { sl += st->seq_len_hist[i]; ql += st->qual_len_hist[i]; }
  st->seq_lines = sl;
  st->qual_lines = ql;
  // quality lines that have a byte at position p = lines longer than p
  u64 longer = ql, inbins = 0;
  for (int p = 0; p < fq::POS_BINS; p++) {
    longer -= st->qual_len_hist[p];
    st->qual_pos_cnt[p] = longer;
    inbins += longer;
  }
  st->qual_pos_cnt[fq::POS_BINS] = quals - inbins;
  // fq-meta fold: account the trailing line if the sampling loop would still read it
  long long qmin = c.qual_min, qmax = c.qual_max;
  u64 ml = c.meta_lines;
  unsigned status = c.meta_status;
  if (meta_records && c.open_len && ml < meta_records * 4) {
    int has = c.cur_has, mn = c.cur_min, mx = c.cur_max;
    if (c.meta_pending_cr) { if (!has) { has = 1; mn = mx = -1; } else mn = -1; }
    if ((ml & 3) == 3 && status == FQGPU_META_OK) {
      if (has) {
        long long a = mn, b = mx;
        if (qmin >= 0) { a = a < qmin ? a : qmin; b = b > qmax ? b : qmax; }
        qmin = a; qmax = b;
      } else if (qmin < 0) status = FQGPU_META_EMPTY_QUAL;
    }
    ml++;
  }
  st->meta_qual_min = qmin;
  st->meta_qual_max = qmax;
  st->meta_lines = ml;
  st->meta_status = status;
  if (core_only) fqgpu_zero_quality(st);
}

extern "C" {

int fqgpu_finish(fqgpu_ctx* ctx, fqgpu_stats* out) {
  NvtxRange nvtx_("fqgpu_finish");
  if (!ctx || !out) return FQGPU_EARG;
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  static_assert(sizeof(fq::Carry) % sizeof(u64) == 0, "h_out layout");
  u64* h_err = ctx->h_out + fq::BLOCK_WORDS + sizeof(fq::Carry) / sizeof(u64);
  CU_TRY(ctx, cudaMemcpyAsync(ctx->h_out, ctx->d_acc, fq::BLOCK_WORDS * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
  CU_TRY(ctx, cudaMemcpyAsync(ctx->h_out + fq::BLOCK_WORDS, ctx->d_carry, sizeof(fq::Carry), cudaMemcpyDeviceToHost, ctx->stream));
  CU_TRY(ctx, cudaMemcpyAsync(h_err, ctx->d_ctl + fq::CTL_ERROR, sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
  CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  if (*h_err) return fail(ctx, FQGPU_ECUDA, "fqgpu_finish: the scan kernel reported an internal consistency error");
  fq::Carry c;
  memcpy(&c, ctx->h_out + fq::BLOCK_WORDS, sizeof(c));
  fqgpu_assemble_stats(ctx->h_out, c, ctx->cfg.meta_records, out, (ctx->cfg.flags & FQGPU_F_CORE_ONLY) != 0);
  return FQGPU_OK;
}

int fqgpu_count_device(fqgpu_ctx* ctx, const void* dptr, size_t nbytes, fqgpu_stats* out) {
  int rc = fqgpu_reset(ctx);
  if (rc != FQGPU_OK) return rc;
  rc = fqgpu_scan_device(ctx, dptr, nbytes);
  if (rc != FQGPU_OK) return rc;
  return fqgpu_finish(ctx, out);
}

// ---- host -> device staging ------------------------------------------------------------------
static int ensure_stage(fqgpu_ctx* ctx) {
  if (ctx->d_stage[0]) return FQGPU_OK;
  if (!ctx->cstream) CU_TRY(ctx, cudaStreamCreateWithFlags(&ctx->cstream, cudaStreamNonBlocking));
  for (int b = 0; b < 2; b++) {
    CU_TRY(ctx, cudaMalloc(&ctx->d_stage[b], ctx->chunk_bytes));
    CU_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_copied[b], cudaEventDisableTiming));
    CU_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_scanned[b], cudaEventDisableTiming));
  }
  return FQGPU_OK;
}

// Copies `nbytes` (<= chunk_bytes) of host memory into the next landing buffer on the copy stream
// and scans it on the compute stream.  `copied` (optional) is recorded when the host bytes are no
// longer needed.
static int stage_and_scan(fqgpu_ctx* ctx, const void* host, size_t nbytes, cudaEvent_t copied) {
  int rc = ensure_stage(ctx);
  if (rc != FQGPU_OK) return rc;
  const int b = (int)(ctx->n_staged & 1);
  if (ctx->n_staged >= 2) CU_TRY(ctx, cudaStreamWaitEvent(ctx->cstream, ctx->ev_scanned[b], 0));
  CU_TRY(ctx, cudaMemcpyAsync(ctx->d_stage[b], host, nbytes, cudaMemcpyHostToDevice, ctx->cstream));
  CU_TRY(ctx, cudaEventRecord(ctx->ev_copied[b], ctx->cstream));
  if (copied) CU_TRY(ctx, cudaEventRecord(copied, ctx->cstream));
  CU_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_copied[b], 0));
  rc = fqgpu_scan_device(ctx, ctx->d_stage[b], nbytes);
  if (rc != FQGPU_OK) return rc;
  CU_TRY(ctx, cudaEventRecord(ctx->ev_scanned[b], ctx->stream));
  ctx->n_staged++;
  return FQGPU_OK;
}

// ---- staging ring -------------------------------------------------------------------------------
void* fqgpu_acquire(fqgpu_ctx* ctx, size_t* capacity) {
  NvtxRange nvtx_("fqgpu_acquire");
  if (!ctx) return nullptr;
  if (cudaSetDevice(ctx->device) != cudaSuccess) { ctx->err = "cudaSetDevice failed"; return nullptr; }
  StageBuf& b = ctx->ring[ctx->next];
  if (!b.host) {
    cudaError_t e = cudaMallocHost(&b.host, ctx->chunk_bytes);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&b.copied, cudaEventDisableTiming);
    if (e != cudaSuccess) { ctx->err = std::string("staging allocation failed: ") + cudaGetErrorString(e); return nullptr; }
  }
  if (b.in_flight) {
    cudaError_t e = cudaEventSynchronize(b.copied);
    if (e != cudaSuccess) { ctx->err = std::string("cudaEventSynchronize: ") + cudaGetErrorString(e); return nullptr; }
    b.in_flight = false;
  }
  if (capacity) *capacity = ctx->chunk_bytes;
  return b.host;
}

int fqgpu_submit(fqgpu_ctx* ctx, void* chunk, size_t nbytes) {
  NvtxRange nvtx_("fqgpu_submit");
  if (!ctx) return FQGPU_EARG;
  int slot = -1;
  for (size_t i = 0; i < ctx->ring.size(); i++) if (ctx->ring[i].host == chunk && chunk) slot = (int)i;
  if (slot < 0) return fail(ctx, FQGPU_EARG, "fqgpu_submit: chunk was not handed out by fqgpu_acquire");
  if (nbytes > ctx->chunk_bytes) return fail(ctx, FQGPU_EARG, "fqgpu_submit: nbytes exceeds the chunk capacity");
  if (nbytes == 0) return FQGPU_OK;
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  StageBuf& b = ctx->ring[slot];
  b.in_flight = true;
  ctx->next = (slot + 1) % (int)ctx->ring.size();
  return stage_and_scan(ctx, b.host, nbytes, b.copied);
}

int fqgpu_scan_host(fqgpu_ctx* ctx, const void* buf, size_t nbytes) {
  NvtxRange nvtx_("fqgpu_scan_host");
  if (!ctx || (!buf && nbytes)) return FQGPU_EARG;
  // H2D straight from the caller's buffer (asynchronous when it is pinned), double-buffered on
  // the device so the copy of chunk k+1 overlaps the scan of chunk k
  const uint8_t* p = (const uint8_t*)buf;
  size_t left = nbytes;
  while (left) {
    size_t n = left < ctx->chunk_bytes ? left : ctx->chunk_bytes;
    int rc = stage_and_scan(ctx, p, n, nullptr);
    if (rc != FQGPU_OK) return rc;
    p += n;
    left -= n;
  }
  return FQGPU_OK;
}

int fqgpu_count_host(fqgpu_ctx* ctx, const void* buf, size_t nbytes, fqgpu_stats* out) {
  if (!ctx || !out || (!buf && nbytes)) return FQGPU_EARG;
  int rc = fqgpu_reset(ctx);
  if (rc != FQGPU_OK) return rc;
  rc = fqgpu_scan_host(ctx, buf, nbytes);
  if (rc != FQGPU_OK) return rc;
  return fqgpu_finish(ctx, out);
}

// src/fq_count.nim:30-36: ".gz" (case-sensitive) selects the gz stream; gzread() inflates straight
// into the pinned chunk, the call shape of gzip_stream.nim:16-17.
int fqgpu_count_file(fqgpu_ctx* ctx, const char* path, fqgpu_stats* out) {
  if (!ctx || !path || !out) return FQGPU_EARG;
  const size_t L = strlen(path);
  return fqgpu_count_file_as(ctx, path, L >= 3 && strcmp(path + L - 3, ".gz") == 0, out);
}

}  // extern "C"

// ---- host ingest: a regular file is read by several threads at once -------------------------------------------
// One thread copies page-cache bytes into pinned memory at a few GB/s, an order of magnitude below what the H2D
// copy behind it moves; T threads pread() disjoint slices of the same chunk.  FQGPU_READ_THREADS overrides T
// (default: three quarters of the hardware threads, at most 16; shared out among the workers of fqgpu_count_files).  Returns the contiguous bytes read from `off` (short at EOF).
static thread_local int g_read_threads_share = 0;  // > 0: this thread is one of several ingesting at once (count_files)
static int read_threads() {
  static const int t = [] {
    if (const char* e = getenv("FQGPU_READ_THREADS")) return atoi(e) > 0 ? atoi(e) : 1;
    const unsigned hw = std::thread::hardware_concurrency();
    const int d = (int)(hw * 3 / 4);  // measured on a 16-thread host: 4 -> 24, 8 -> 34, 12 -> 43, 16 -> 44, 24 -> 39 GB/s
    return d < 1 ? 1 : (d > 16 ? 16 : d);
  }();
  if (g_read_threads_share > 0) return t / g_read_threads_share > 0 ? t / g_read_threads_share : 1;
  return t;
}

static size_t parallel_pread(int fd, uint8_t* dst, size_t n, size_t off, bool* io_error, int threads = 0) {
  const int T = threads > 0 ? threads : read_threads();
  std::atomic<bool> bad(false);
  auto slice_read = [&](size_t b, size_t e) {
    size_t g = 0;
    while (b + g < e) {
      const ssize_t r = pread(fd, dst + b + g, e - b - g, (off_t)(off + b + g));
      if (r < 0) bad = true;
      if (r <= 0) break;
      g += (size_t)r;
    }
    return g;
  };
  if (T <= 1 || n < ((size_t)8 << 20)) { const size_t g = slice_read(0, n); if (io_error) *io_error = bad; return g; }
  const size_t slice = (((n + (size_t)T - 1) / (size_t)T) + 4095) & ~(size_t)4095;
  std::vector<size_t> got((size_t)T, 0);
  std::vector<std::thread> pool;
  for (int k = 1; k < T; k++)
    pool.emplace_back([&, k] { const size_t b = (size_t)k * slice; if (b < n) got[(size_t)k] = slice_read(b, b + slice < n ? b + slice : n); });
  got[0] = slice_read(0, slice < n ? slice : n);
  for (auto& th : pool) th.join();
  size_t total = 0;
  for (int k = 0; k < T; k++) {
    const size_t b = (size_t)k * slice;
    if (b >= n) break;
    const size_t want = (b + slice < n ? b + slice : n) - b;
    total += got[(size_t)k];
    if (got[(size_t)k] < want) break;  // EOF (or an error) inside this slice: the rest is not contiguous
  }
  if (io_error) *io_error = bad;
  return total;
}

// ---- BGZF input: members walked on the host, inflated on the device (fq_bgzf.cu) ---------------------------
// Returns FQGPU_OK when the whole file went through the device path (the caller finishes the stream), 1 when the
// file is not well-formed BGZF (or cannot be opened): the caller resets and takes the zlib path, which also
// reports I/O errors the way the reference does.  < 0: CUDA failure.
static size_t env_size(const char* name, size_t dflt, size_t lo, size_t hi) {
  const char* e = getenv(name);
  if (!e) return dflt;
  const long long v = atoll(e);
  return v < (long long)lo ? lo : (v > (long long)hi ? hi : (size_t)v);
}

static const size_t kBgzfBatchBytes = (size_t)64 << 20;  // compressed bytes per batch (about 4000 members: one wave of warps)
static const size_t kBgzfBatchMembers = 32768;           // members (= inflating warps) per batch

// the two staging slots for compressed input (shared by the BGZF and the gzip path)
static cudaError_t ensure_comp_slots(fqgpu_ctx* ctx, size_t want, int nslots) {
  cudaError_t e = cudaSuccess;
  if (!ctx->cstream && (e = cudaStreamCreateWithFlags(&ctx->cstream, cudaStreamNonBlocking)) != cudaSuccess) return e;
  for (auto& c : ctx->comp) {
    if (!c.h2d && (e = cudaEventCreateWithFlags(&c.h2d, cudaEventDisableTiming)) != cudaSuccess) return e;
    if (!c.done && (e = cudaEventCreateWithFlags(&c.done, cudaEventDisableTiming)) != cudaSuccess) return e;
  }
  if (ctx->comp_cap < want) {  // (both slots have the same size)
    if ((e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) return e;
    for (auto& c : ctx->comp) {
      if (c.h) cudaFreeHost(c.h);
      cudaFree(c.d);
      c.h = nullptr; c.d = nullptr;
    }
    ctx->comp_cap = want;
  }
  // the second slot only when a second batch is read ahead: pinned memory is the most expensive thing to allocate here
  for (int k = 0; k < nslots && k < 2; k++) {
    fqgpu_ctx::CompSlot& c = ctx->comp[k];
    if (c.h && c.d) continue;
    if (!c.h && (e = cudaMallocHost(&c.h, ctx->comp_cap)) != cudaSuccess) { ctx->comp_cap = 0; return e; }
    if (!c.d && (e = cudaMalloc(&c.d, ctx->comp_cap + 64)) != cudaSuccess) { ctx->comp_cap = 0; return e; }
  }
  return cudaSuccess;
}

static int count_bgzf(fqgpu_ctx* ctx, const char* path) {
  const int fd = open(path, O_RDONLY);
  if (fd < 0) return 1;
  struct stat sb;
  uint8_t hd[18];
  if (fstat(fd, &sb) != 0 || !S_ISREG(sb.st_mode) || sb.st_size < 28 || pread(fd, hd, 18, 0) != 18 ||
      !(hd[0] == 0x1f && hd[1] == 0x8b && hd[2] == 8 && hd[3] == 4 && hd[12] == 'B' && hd[13] == 'C' && hd[14] == 2 && hd[15] == 0)) {
    close(fd);
    return 1;
  }
  const size_t fsize = (size_t)sb.st_size;
  // (on every way out the stream is drained: the slots may still be in use by copies and kernels in flight)
  auto bail = [&](int code) { cudaStreamSynchronize(ctx->cstream); cudaStreamSynchronize(ctx->stream); close(fd); return code; };
#define CU_B(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_); return bail(FQGPU_ECUDA); } } while (0)
  CU_B(cudaSetDevice(ctx->device));
  const size_t batch_cap = env_size("FQGPU_BGZF_BATCH_MB", kBgzfBatchBytes >> 20, 1, 1024) << 20;
  CU_B(ensure_comp_slots(ctx, fsize < batch_cap ? ((fsize + 4095) & ~(size_t)4095) : batch_cap, fsize > batch_cap ? 2 : 1));
  if (ctx->members_cap < kBgzfBatchMembers) {
    for (auto& c : ctx->comp) {
      CU_B(cudaMallocHost(&c.h_members, kBgzfBatchMembers * sizeof(fq::BgzfMember)));
      CU_B(cudaMalloc(&c.d_members, kBgzfBatchMembers * sizeof(fq::BgzfMember)));
      CU_B(cudaMallocHost(&c.h_status, kBgzfBatchMembers * sizeof(uint32_t)));
      CU_B(cudaMalloc(&c.d_status, kBgzfBatchMembers * sizeof(uint32_t)));
    }
    ctx->members_cap = kBgzfBatchMembers;
  }
  // Batch k: read and walked on the host while the device still inflates and scans batch k - 1; its H2D copy runs on
  // the copy stream.  A member that failed to inflate is noticed one or two batches later -- what was scanned in between
  // is thrown away with everything else when the caller resets for the zlib path.
  int pending[2] = {0, 0};  // members of the batch last sent through the slot, not yet checked
  auto settle = [&](int s) -> int {  // 1: some member of that batch did not inflate; < 0: CUDA failure
    if (!pending[s]) return 0;
    if (cudaEventSynchronize(ctx->comp[s].done) != cudaSuccess) return FQGPU_ECUDA;
    for (int i = 0; i < pending[s]; i++) if (ctx->comp[s].h_status[i]) return 1;
    pending[s] = 0;
    return 0;
  };
  size_t pos = 0;
  for (int k = 0; pos < fsize; k++) {
    const int s = k & 1;
    if (s && !ctx->comp[1].h) CU_B(ensure_comp_slots(ctx, ctx->comp_cap, 2));  // (a second batch after all: many tiny members)
    fqgpu_ctx::CompSlot& c = ctx->comp[s];
    if (int r = settle(s)) return bail(r);  // (also: the slot's buffers are free again)
    const size_t want_now = fsize - pos < ctx->comp_cap ? fsize - pos : ctx->comp_cap;
    bool io_bad = false;
    const size_t got = parallel_pread(fd, c.h, want_now, pos, &io_bad);
    if (io_bad) return bail(1);  // let the zlib path report it
    fq::BgzfMember* members = (fq::BgzfMember*)c.h_members;
    size_t n = 0, off = 0;
    u64 out_total = 0;
    while (off + 18 <= got && n < kBgzfBatchMembers) {
      const uint8_t* h = c.h + off;
      if (!(h[0] == 0x1f && h[1] == 0x8b && h[2] == 8 && h[3] == 4)) return bail(1);
      const size_t xlen = (size_t)h[10] | ((size_t)h[11] << 8);
      if (off + 12 + xlen > got) break;  // the header continues in the next batch
      long bsize = -1;
      for (size_t q = 12; q + 4 <= 12 + xlen;) {
        const size_t slen = (size_t)h[q + 2] | ((size_t)h[q + 3] << 8);
        if (h[q] == 'B' && h[q + 1] == 'C' && slen == 2 && q + 6 <= 12 + xlen) bsize = (long)h[q + 4] | ((long)h[q + 5] << 8);
        q += 4 + slen;
      }
      if (bsize < 0) return bail(1);
      const size_t total = (size_t)bsize + 1;
      if (total < 12 + xlen + 8) return bail(1);
      if (off + total > got) break;  // the member continues in the next batch
      const uint8_t* t = h + total - 4;
      const uint32_t isize = (uint32_t)t[0] | ((uint32_t)t[1] << 8) | ((uint32_t)t[2] << 16) | ((uint32_t)t[3] << 24);
      if (isize > 65536u) return bail(1);
      fq::BgzfMember& m = members[n++];
      m.in_off = off + 12 + xlen; m.out_off = out_total; m.in_len = (unsigned)(total - xlen - 12 - 8); m.out_len = isize;
      m.crc = (uint32_t)t[-4] | ((uint32_t)t[-3] << 8) | ((uint32_t)t[-2] << 16) | ((uint32_t)t[-1] << 24); m.pad = 0;
      out_total += isize;
      off += total;
    }
    if (off == 0) return bail(1);  // no complete member in a full batch / trailing garbage
    if (ctx->inflated_cap < out_total + 64) {
      CU_B(cudaStreamSynchronize(ctx->stream));  // (the scan of the batch before reads the old buffer)
      cudaFree(ctx->d_inflated);
      ctx->d_inflated = nullptr; ctx->inflated_cap = 0;
      const size_t cap = ((size_t)out_total + 64 + ((size_t)64 << 20)) & ~(size_t)4095;
      CU_B(cudaMalloc(&ctx->d_inflated, cap));
      ctx->inflated_cap = cap;
    }
    CU_B(cudaMemcpyAsync(c.d, c.h, off, cudaMemcpyHostToDevice, ctx->cstream));
    CU_B(cudaMemcpyAsync(c.d_members, c.h_members, n * sizeof(fq::BgzfMember), cudaMemcpyHostToDevice, ctx->cstream));
    CU_B(cudaEventRecord(c.h2d, ctx->cstream));
    CU_B(cudaStreamWaitEvent(ctx->stream, c.h2d, 0));
    CU_B(fq::launch_bgzf_inflate(c.d, (const fq::BgzfMember*)c.d_members, (int)n, ctx->d_inflated, c.d_status, ctx->stream));
    CU_B(cudaMemcpyAsync(c.h_status, c.d_status, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    const int rc = fqgpu_scan_device(ctx, ctx->d_inflated, (size_t)out_total);
    if (rc != FQGPU_OK) return bail(rc);
    CU_B(cudaEventRecord(c.done, ctx->stream));
    pending[s] = (int)n;
    ctx->launches += 1;
    ctx->bgzf_members += (u64)n;
    pos += off;
  }
  for (int s = 0; s < 2; s++) if (int r = settle(s)) return bail(r);
#undef CU_B
  close(fd);
  return FQGPU_OK;
}

// ---- ordinary gzip input: inflated on the device in parallel (fq_gzip.cu) ------------------------------------
// Same contract as count_bgzf: FQGPU_OK = the whole file went through the device path, 1 = anything the device
// path does not prove (not gzip, a chain that does not close, a stream zlib would reject, a truncated file, a
// wrong ISIZE): the caller resets and lets gzread decide, which keeps the reference's behaviour for broken input.
// The compressed file is taken in batches (FQGPU_GZ_BATCH_MB, default 512 MiB); a batch starts at the block
// boundary where the one before stopped, with the 32 KiB before it as its window.

// the first payload byte of the gzip member at `off` (RFC 1952), or 0 when there is no member header there
static size_t gzip_payload_offset(int fd, size_t off, size_t fsize) {
  uint8_t h[10];
  if (off + 18 > fsize || pread(fd, h, 10, (off_t)off) != 10) return 0;
  if (!(h[0] == 0x1f && h[1] == 0x8b && h[2] == 8) || (h[3] & 0xE0)) return 0;
  size_t p = off + 10;
  const int flg = h[3];
  if (flg & 4) {  // FEXTRA
    uint8_t x[2];
    if (pread(fd, x, 2, (off_t)p) != 2) return 0;
    p += 2 + ((size_t)x[0] | ((size_t)x[1] << 8));
  }
  for (int pass = 0; pass < 2; pass++) {  // FNAME, FCOMMENT: zero-terminated
    if (!(flg & (pass == 0 ? 8 : 16))) continue;
    for (;;) {
      uint8_t buf[256];
      const ssize_t r = p < fsize ? pread(fd, buf, sizeof buf, (off_t)p) : 0;
      if (r <= 0) return 0;
      const void* z = memchr(buf, 0, (size_t)r);
      if (z) { p += (size_t)((const uint8_t*)z - buf) + 1; break; }
      p += (size_t)r;
    }
  }
  if (flg & 2) p += 2;  // FHCRC
  return p + 8 < fsize ? p : 0;
}

// CRC-32 arithmetic for joining the pieces the device computes (fq_gzip.cu, step 7): polynomials over GF(2) modulo
// the CRC polynomial, bit 31 = the coefficient of x^0.
static uint32_t crc_mul(uint32_t a, uint32_t b) {
  uint32_t p = 0;
  for (int i = 0; i < 32; i++) {
    if (a & (0x80000000u >> i)) p ^= b;
    b = (b >> 1) ^ ((b & 1u) ? 0xEDB88320u : 0u);
  }
  return p;
}
static uint32_t crc_xpow8(u64 nbytes) {  // x^(8 nbytes)
  uint32_t r = 0x80000000u, sq = 0x00800000u;  // 1, x^8
  for (u64 n = nbytes; n; n >>= 1) {
    if (n & 1) r = crc_mul(sq, r);
    sq = crc_mul(sq, sq);
  }
  return r;
}

static int count_gzip(fqgpu_ctx* ctx, const char* path) {
  const int fd = open(path, O_RDONLY);
  if (fd < 0) return 1;
  struct stat sb;
  if (fstat(fd, &sb) != 0 || !S_ISREG(sb.st_mode)) { close(fd); return 1; }
  const size_t fsize = (size_t)sb.st_size;
  size_t payload = gzip_payload_offset(fd, 0, fsize);
  if (!payload) { close(fd); return 1; }
  // (on every way out the streams are drained: the slots may still be in use by copies in flight)
  auto bail = [&](int code) { if (ctx->cstream) cudaStreamSynchronize(ctx->cstream); cudaStreamSynchronize(ctx->stream); close(fd); return code; };
#define CU_B(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_); return bail(FQGPU_ECUDA); } } while (0)
  // buffers whose size follows the input: when they do not fit, the host path (which streams) takes the file
#define CU_ALLOC(call) do { cudaError_t e_ = (call); if (e_ == cudaErrorMemoryAllocation) { cudaGetLastError(); return bail(1); } if (e_ != cudaSuccess) { ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_); return bail(FQGPU_ECUDA); } } while (0)
  CU_B(cudaSetDevice(ctx->device));
  const size_t batch_cap = env_size("FQGPU_GZ_BATCH_MB", 512, 1, 1024) << 20;
  CU_ALLOC(ensure_comp_slots(ctx, fsize + 4096 < batch_cap ? ((fsize + 8191) & ~(size_t)4095) : batch_cap, 1));
  if (!ctx->d_gzchunks) {
    CU_B(cudaMalloc(&ctx->d_gzchunks, (size_t)fq::GZ_MAX_CHUNKS * sizeof(fq::GzChunk)));
    CU_B(cudaMalloc(&ctx->d_gzorder, (size_t)fq::GZ_MAX_CHUNKS * sizeof(uint32_t)));
    CU_B(cudaMalloc(&ctx->d_gzcoff, ((size_t)fq::GZ_MAX_CHUNKS + 1) * sizeof(u64)));
    CU_B(cudaMalloc(&ctx->d_gzsb, ((size_t)fq::GZ_MAX_CHUNKS + 1) * sizeof(u64)));
    CU_B(cudaMalloc(&ctx->d_gzres, sizeof(fq::GzResult) + 16));
    CU_B(cudaMallocHost(&ctx->h_gzres, sizeof(fq::GzResult) + 16));
    CU_B(cudaMalloc(&ctx->d_gzwindow, fq::GZ_WINDOW));
  }
  fq::GzChunk* chunks = (fq::GzChunk*)ctx->d_gzchunks;
  fq::GzResult* d_res = (fq::GzResult*)ctx->d_gzres;
  uint32_t* d_err = (uint32_t*)((uint8_t*)ctx->d_gzres + sizeof(fq::GzResult));
  const fq::GzResult* h_res = (const fq::GzResult*)ctx->h_gzres;
  const uint32_t* h_err = (const uint32_t*)((const uint8_t*)ctx->h_gzres + sizeof(fq::GzResult));
  const size_t chunk_min = env_size("FQGPU_GZ_CHUNK_KB", 16, 1, 1024) << 10;
  const bool two_pass = getenv("FQGPU_GZ_TWO_PASS") != nullptr;
  const size_t kArenaRatio = env_size("FQGPU_GZ_ARENA_RATIO", 8, 2, 64);

  const bool trace = getenv("FQGPU_GZ_TRACE") != nullptr;  // per-stage wall clock (adds stream syncs; diagnostics only)
  auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  double t_mark = now();
  auto mark = [&](const char* what) {
    if (!trace) return;
    cudaStreamSynchronize(ctx->stream);
    const double t = now();
    fprintf(stderr, "[gz] %-10s %8.3f ms\n", what, t - t_mark);
    t_mark = t;
  };
  // file bytes [at, at + cap) into a slot: read into its pinned buffer, copied on the copy stream once the kernels
  // that still read the slot's device buffer are through.  Returns the bytes read, SIZE_MAX on an I/O error.
  cudaError_t load_err = cudaSuccess;
  const size_t kPiece = (size_t)32 << 20;
  auto load = [&](int slot, size_t at) -> size_t {
    fqgpu_ctx::CompSlot& c = ctx->comp[slot];
    if ((load_err = cudaEventSynchronize(c.h2d)) != cudaSuccess) return 0;  // (the pinned buffer's last copy has left it)
    const size_t want_now = fsize - at < ctx->comp_cap ? fsize - at : ctx->comp_cap;
    if ((load_err = cudaStreamWaitEvent(ctx->cstream, c.done, 0)) != cudaSuccess) return 0;
    size_t n = 0;
    while (n < want_now) {  // in pieces: the copy of one piece runs under the read of the next
      const size_t piece = want_now - n < kPiece ? want_now - n : kPiece;
      bool io_bad = false;
      const size_t r = parallel_pread(fd, c.h + n, piece, at + n, &io_bad);
      if (io_bad) return SIZE_MAX;
      if (r && (load_err = cudaMemcpyAsync(c.d + n, c.h + n, r, cudaMemcpyHostToDevice, ctx->cstream)) != cudaSuccess) return 0;
      n += r;
      if (r < piece) break;
    }
    if ((load_err = cudaMemsetAsync(c.d + n, 0, 64, ctx->cstream)) != cudaSuccess) return 0;
    load_err = cudaEventRecord(c.h2d, ctx->cstream);
    return n;
  };
  u64 abs_bit = (u64)payload * 8ull;  // the next block's first bit, in the file
  u64 prior_out = 0;                  // bytes of this member already inflated
  uint32_t crc_reg = 0xFFFFFFFFu;     // the member's CRC-32 register so far
  size_t members = 0;                 // members finished
  // Batch k is decoded from slot s while the bytes of batch k + 1 are read and copied into the other slot.  Where
  // batch k + 1 has to start is known only when batch k has been counted (the block boundary it stopped at), so the
  // read is a guess: it starts a little before the end of batch k (kRewind); a last block longer than that is read again.
  const size_t kRewind = (size_t)4 << 20;
  int s = 0;
  size_t fpos = (size_t)(abs_bit >> 3) & ~(size_t)4095;
  size_t got = load(s, fpos);
  if (got == SIZE_MAX) return bail(1);
  CU_B(load_err);
  for (;;) {
    fqgpu_ctx::CompSlot& slot = ctx->comp[s];
    uint8_t* const d_comp = slot.d;
    CU_B(cudaStreamWaitEvent(ctx->stream, slot.h2d, 0));
    mark("pread+h2d");
    const u64 start_bit = abs_bit - (u64)fpos * 8ull;
    if (start_bit + 10 > (u64)got * 8ull) return bail(1);  // the stream ends without a last block: truncated
    size_t chunk_bytes = (got / 12288 + 4095) & ~(size_t)4095;
    if (chunk_bytes < chunk_min) chunk_bytes = chunk_min;
    if ((got + chunk_bytes - 1) / chunk_bytes > (size_t)fq::GZ_MAX_CHUNKS) chunk_bytes = ((got + fq::GZ_MAX_CHUNKS - 1) / fq::GZ_MAX_CHUNKS + 4095) & ~(size_t)4095;
    const int nchunks = (int)((got + chunk_bytes - 1) / chunk_bytes);
    const uint32_t wvalid = prior_out < fq::GZ_WINDOW ? (uint32_t)prior_out : fq::GZ_WINDOW;
    CU_B(cudaMemsetAsync(d_err, 0, 16, ctx->stream));
    CU_B(fq::launch_gz_sync(d_comp, got, (uint32_t)chunk_bytes, nchunks, start_bit, chunks, d_err + 1, ctx->stream));
    CU_B(cudaMemcpyAsync((uint8_t*)ctx->h_gzres + sizeof(fq::GzResult) + 4, d_err + 1, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU_B(cudaStreamSynchronize(ctx->stream));
    // blocks longer than 64 chunks on average (stored or fixed-code streams, no dynamic headers to find): one warp
    // would decode nearly everything alone -- zlib on the host is faster than that
    if (nchunks >= 64 && (size_t)h_err[1] * 64 < (size_t)nchunks) return bail(1);
    mark("sync");
    // One decode pass where the symbols can be kept where they fall: chunk c writes to its own slot of an arena with room
    // for `kArenaRatio` output bytes per compressed byte (FASTQ is near 4.4; the arena is 2 * kArenaRatio times the
    // batch).  A stretch that compresses better than that, a false start run over, or an arena that cannot be had
    // sends the batch through the second pass (gz_write_kernel) as before; FQGPU_GZ_TWO_PASS=1 forces that.
    const u64 stride = (u64)kArenaRatio * chunk_bytes;  // symbols per chunk slot
    bool one_pass = !two_pass;
    if (one_pass && ctx->gzsym_cap < (size_t)nchunks * stride + 16) {
      cudaFree(ctx->d_gzsym);
      ctx->d_gzsym = nullptr; ctx->gzsym_cap = 0;
      const size_t cap = (size_t)nchunks * stride + 4096;
      if (cudaMalloc(&ctx->d_gzsym, cap * sizeof(uint16_t)) == cudaSuccess) ctx->gzsym_cap = cap;
      else { cudaGetLastError(); ctx->d_gzsym = nullptr; one_pass = false; }
    }
    if (one_pass) {
      CU_B(fq::launch_gz_both(d_comp, got, (uint32_t)chunk_bytes, chunks, nchunks, ctx->d_gzsym, stride, ctx->d_gzwindow, wvalid, ctx->stream));
      CU_B(cudaEventRecord(slot.done, ctx->stream));  // (the last kernel that reads the slot, unless the second pass is needed)
    } else {
      CU_B(fq::launch_gz_count(d_comp, got, (uint32_t)chunk_bytes, chunks, nchunks, wvalid, ctx->stream));
    }
    CU_B(fq::launch_gz_chain(chunks, nchunks, prior_out, d_res, ctx->d_gzorder, ctx->d_gzcoff, one_pass ? stride : 0ull, ctx->d_gzsb, ctx->stream));
    CU_B(cudaMemcpyAsync(ctx->h_gzres, ctx->d_gzres, sizeof(fq::GzResult), cudaMemcpyDeviceToHost, ctx->stream));
    // while the device counts: the next batch's bytes
    size_t next_fpos = 0, next_got = 0;
    const bool ahead = fpos + got < fsize;
    if (ahead) {
      CU_ALLOC(ensure_comp_slots(ctx, ctx->comp_cap, 2));
      next_fpos = (fpos + got - (got / 2 < kRewind ? got / 2 : kRewind)) & ~(size_t)4095;
      next_got = load(s ^ 1, next_fpos);
      if (next_got == SIZE_MAX) return bail(1);
      CU_B(load_err);
    }
    CU_B(cudaStreamSynchronize(ctx->stream));
    ctx->launches += 3;
    mark("count");
    const fq::GzResult res = *h_res;
    if (res.status != fq::GZR_OK) return bail(1);
    if (res.total_out == 0 && !res.final_block) return bail(1);  // no complete block in a whole batch, or a truncated file
    const u64 total = res.total_out;
    const bool arena = one_pass && !res.rewrite;  // the symbols are where the one pass put them
    if (one_pass && res.rewrite) ctx->gzip_rewrites++;
    if (total) {
      if (!arena && ctx->gzsym_cap < total + 16) {
        cudaFree(ctx->d_gzsym);
        ctx->d_gzsym = nullptr; ctx->gzsym_cap = 0;
        const size_t cap = (size_t)total + ((size_t)total >> 3) + 4096;
        CU_ALLOC(cudaMalloc(&ctx->d_gzsym, cap * sizeof(uint16_t)));
        ctx->gzsym_cap = cap;
      }
      if (ctx->inflated_cap < total + 64) {
        cudaFree(ctx->d_inflated);
        ctx->d_inflated = nullptr; ctx->inflated_cap = 0;
        const size_t cap = ((size_t)total + ((size_t)total >> 3) + 8192) & ~(size_t)4095;
        CU_ALLOC(cudaMalloc(&ctx->d_inflated, cap));
        ctx->inflated_cap = cap;
      }
      const uint32_t K = fq::gz_group_chunks(res.nchain, ctx->grid / 2);
      const size_t ngroups = ((size_t)res.nchain + K - 1) / K;
      if (ctx->gzwbuf_cap < (size_t)res.nchain + 1) {  // rows of the window walk
        cudaFree(ctx->d_gzwbuf);
        ctx->d_gzwbuf = nullptr; ctx->gzwbuf_cap = 0;
        const size_t rows = (size_t)res.nchain + 1 + ((size_t)res.nchain >> 2);
        // per chunk a row of symbols; per group (at most one per 8 chunks) a row of symbols and a row of bytes
        CU_ALLOC(cudaMalloc(&ctx->d_gzwbuf, rows * fq::GZ_WINDOW * 2 + (rows / 8 + 2) * fq::GZ_WINDOW * 3));
        ctx->gzwbuf_cap = rows;
      }
      uint16_t* symrows = (uint16_t*)ctx->d_gzwbuf;
      uint16_t* grows = symrows + ctx->gzwbuf_cap * fq::GZ_WINDOW;
      uint8_t* trows = (uint8_t*)(grows + (ctx->gzwbuf_cap / 8 + 2) * fq::GZ_WINDOW);
      (void)ngroups;
      mark("alloc");
      if (!arena) {
        CU_B(fq::launch_gz_write(d_comp, got, (uint32_t)chunk_bytes, chunks, nchunks, ctx->d_gzsym, ctx->d_gzwindow, wvalid, d_err, ctx->stream));
        CU_B(cudaEventRecord(slot.done, ctx->stream));  // (the last kernel that reads the slot)
      }
      mark("write");
      const u64* sbase = arena ? ctx->d_gzsb : nullptr;
      CU_B(fq::launch_gz_windows(ctx->d_gzcoff, sbase, res.nchain, K, ctx->d_gzsym, symrows, grows, trows, ctx->d_gzwindow, ctx->stream));
      CU_B(fq::launch_gz_resolve(ctx->d_gzsym, sbase, symrows, trows, K, ctx->d_gzcoff, res.nchain, total, ctx->d_inflated, ctx->grid / 2, ctx->stream));
      mark("windows");
      const u64 nslices = total / fq::GZ_CRC_SLICE;
      const uint32_t q = (uint32_t)((nslices + 1023) / 1024);
      if (ctx->gzraw_cap < nslices + 1) {
        cudaFree(ctx->d_gzraw);
        ctx->d_gzraw = nullptr; ctx->gzraw_cap = 0;
        CU_ALLOC(cudaMalloc(&ctx->d_gzraw, (nslices + 1 + (nslices >> 3)) * sizeof(uint32_t)));
        ctx->gzraw_cap = nslices + 1 + (nslices >> 3);
      }
      CU_B(fq::launch_gz_crc(ctx->d_inflated, total, ctx->d_gzraw, crc_xpow8(fq::GZ_CRC_SLICE), crc_xpow8((u64)fq::GZ_CRC_SLICE * q), q, d_err + 2, ctx->stream));
      CU_B(cudaMemcpyAsync((uint8_t*)ctx->h_gzres + sizeof(fq::GzResult), d_err, 16, cudaMemcpyDeviceToHost, ctx->stream));
      ctx->launches += 7;
      const int rc = fqgpu_scan_device(ctx, ctx->d_inflated, (size_t)total);
      if (rc != FQGPU_OK) return bail(rc);
      CU_B(cudaStreamSynchronize(ctx->stream));  // the batch buffers are reused
      mark("crc+scan");
      if (*h_err) return bail(1);
      // the register over this batch's bytes (full slices, then the tail), behind the bytes before it
      const uint32_t batch_reg = crc_mul(crc_xpow8(total % fq::GZ_CRC_SLICE), h_err[2]) ^ h_err[3];
      crc_reg = crc_mul(crc_xpow8(total), crc_reg) ^ batch_reg;
    }
    ctx->gzip_chunks += res.nchain;
    ctx->gzip_passed += res.passed;
    prior_out += total;
    abs_bit = (u64)fpos * 8ull + res.end_bit;
    // where the stream goes on -- the next block, or, behind a member's last block, the next member's first
    bool more = true;
    if (res.final_block) {
      // the member's trailer: CRC-32 and ISIZE, both checked like gzread does
      const size_t tpos = (size_t)((abs_bit + 7) >> 3);
      uint8_t tr[8];
      if (tpos + 8 > fsize || pread(fd, tr, 8, (off_t)tpos) != 8) return bail(1);
      const uint32_t isize = (uint32_t)tr[4] | ((uint32_t)tr[5] << 8) | ((uint32_t)tr[6] << 16) | ((uint32_t)tr[7] << 24);
      const uint32_t crc = (uint32_t)tr[0] | ((uint32_t)tr[1] << 8) | ((uint32_t)tr[2] << 16) | ((uint32_t)tr[3] << 24);
      if (isize != (uint32_t)prior_out || crc != (crc_reg ^ 0xFFFFFFFFu)) return bail(1);
      // another member behind it (`cat a.gz b.gz`) continues the stream; anything else is ignored, as gzread does
      payload = gzip_payload_offset(fd, tpos + 8, fsize);
      if (!payload) {
        more = false;
      } else {
        // a long row of small members (each one costs a batch of its own): gzread is no slower on those
        if (++members >= 64 && (tpos + 8) / members < ((size_t)256 << 10)) return bail(1);
        abs_bit = (u64)payload * 8ull;
        prior_out = 0;
        crc_reg = 0xFFFFFFFFu;
      }
    }
    if (!more) break;
    // the bytes from there: already on their way into the other slot if the guess covered the position; if not, they
    // are read into this one again (its kernels are through: the sync above)
    if (ahead && (size_t)(abs_bit >> 3) >= next_fpos && (abs_bit >> 3) + 16 < (u64)next_fpos + next_got) {
      s ^= 1;
      fpos = next_fpos; got = next_got;
    } else {
      fpos = (size_t)(abs_bit >> 3) & ~(size_t)4095;
      if (fpos >= fsize) return bail(1);  // the stream ends without a last block: truncated
      got = load(s, fpos);
      if (got == SIZE_MAX) return bail(1);
      CU_B(load_err);
    }
  }
  if (ctx->cstream) cudaStreamSynchronize(ctx->cstream);  // (a read-ahead that was not needed)
#undef CU_B
#undef CU_ALLOC
  close(fd);
  return FQGPU_OK;
}

extern "C" {

unsigned long long fqgpu_bgzf_members(const fqgpu_ctx* ctx) { return ctx ? ctx->bgzf_members : 0; }
unsigned long long fqgpu_gzip_chunks(const fqgpu_ctx* ctx) { return ctx ? ctx->gzip_chunks : 0; }
unsigned long long fqgpu_gzip_false_starts(const fqgpu_ctx* ctx) { return ctx ? ctx->gzip_passed : 0; }
unsigned long long fqgpu_gzip_second_passes(const fqgpu_ctx* ctx) { return ctx ? ctx->gzip_rewrites : 0; }

int fqgpu_count_file_as(fqgpu_ctx* ctx, const char* path, int as_gz, fqgpu_stats* out) {
  NvtxRange nvtx_("fqgpu_count_file_as");
  if (!ctx || !path || !out) return FQGPU_EARG;
  const bool gz = as_gz != 0;
  int rc = fqgpu_reset(ctx);
  if (rc != FQGPU_OK) return rc;
  if (gz && !getenv("FQGPU_NO_BGZF")) {  // BGZF (blocked gzip): inflated on the device, member by member
    rc = count_bgzf(ctx, path);
    if (rc == FQGPU_OK) return fqgpu_finish(ctx, out);
    if (rc < 0) return rc;
    if ((rc = fqgpu_reset(ctx)) != FQGPU_OK) return rc;  // not BGZF
  }
  if (gz && !getenv("FQGPU_NO_GZIP_DEVICE")) {  // ordinary gzip: chunks of the one DEFLATE stream inflated in parallel
    rc = count_gzip(ctx, path);
    if (rc == FQGPU_OK) return fqgpu_finish(ctx, out);
    if (rc < 0) return rc;
    if ((rc = fqgpu_reset(ctx)) != FQGPU_OK) return rc;  // not proven: the host zlib path below
  }
  if (gz) {
    gzFile f = gzopen(path, "rb");
    if (!f) return fail(ctx, FQGPU_EIO, std::string("Unable to open file: ") + path);
    gzbuffer(f, 1 << 20);
    for (;;) {
      size_t cap = 0;
      uint8_t* chunk = (uint8_t*)fqgpu_acquire(ctx, &cap);
      if (!chunk) { gzclose(f); return FQGPU_ECUDA; }
      size_t got = 0;
      while (got < cap) {
        size_t want = cap - got;
        if (want > ((size_t)1 << 30)) want = (size_t)1 << 30;
        int r = gzread(f, chunk + got, (unsigned)want);
        // a truncated / corrupt stream simply ENDS, as in the reference (gzip_stream.nim:16-17 hands gzread's result to the
        // stream, whose atEnd() then holds): the lines inflated so far are counted and the row is printed
        if (r <= 0) break;
        got += (size_t)r;
      }
      if (got == 0) break;
      rc = fqgpu_submit(ctx, chunk, got);
      if (rc != FQGPU_OK) { gzclose(f); return rc; }
      if (got < cap) break;
    }
    gzclose(f);
  } else {
    int fd = open(path, O_RDONLY);
    if (fd < 0) return fail(ctx, FQGPU_EIO, std::string("Unable to open file: ") + path);
    struct stat sb;
    const bool regular = fstat(fd, &sb) == 0 && S_ISREG(sb.st_mode);
    size_t pos = 0;
    for (;;) {
      size_t cap = 0;
      uint8_t* chunk = (uint8_t*)fqgpu_acquire(ctx, &cap);
      if (!chunk) { close(fd); return FQGPU_ECUDA; }
      size_t got = 0;
      if (regular) {  // several threads fill the chunk (pread at explicit offsets)
        bool io_bad = false;
        got = parallel_pread(fd, chunk, cap, pos, &io_bad);
        if (io_bad) { close(fd); return fail(ctx, FQGPU_EIO, std::string("read failed: ") + path); }
        pos += got;
      } else {
        while (got < cap) {
          ssize_t r = read(fd, chunk + got, cap - got);
          if (r < 0) { close(fd); return fail(ctx, FQGPU_EIO, std::string("read failed: ") + path); }
          if (r == 0) break;
          got += (size_t)r;
        }
      }
      if (got == 0) break;
      rc = fqgpu_submit(ctx, chunk, got);
      if (rc != FQGPU_OK) { close(fd); return rc; }
      if (got < cap) break;
    }
    close(fd);
  }
  return fqgpu_finish(ctx, out);
}

// ---- the head of a file: what the fq-meta sampling loop consumes (src/fq_meta.nim:226-248) -----------------
int fqgpu_meta_file_as(fqgpu_ctx* ctx, const char* path, int as_gz, fqgpu_stats* out) {
  if (!ctx || !path || !out) return FQGPU_EARG;
  int rc = fqgpu_reset(ctx);
  if (rc != FQGPU_OK) return rc;
  const u64 want_lines = ctx->cfg.meta_records * 4ull;
  gzFile gf = nullptr;
  int fd = -1;
  if (as_gz) {
    gf = gzopen(path, "rb");
    if (!gf) return fail(ctx, FQGPU_EIO, std::string("Unable to open file: ") + path);
  } else {
    fd = open(path, O_RDONLY);
    if (fd < 0) return fail(ctx, FQGPU_EIO, std::string("Unable to open file: ") + path);
  }
  auto done = [&](int code) { if (gf) gzclose(gf); if (fd >= 0) close(fd); return code; };
  u64 lines = 0;
  bool eof = want_lines == 0;
  while (!eof && lines < want_lines) {
    size_t cap = 0;
    uint8_t* chunk = (uint8_t*)fqgpu_acquire(ctx, &cap);
    if (!chunk) return done(FQGPU_ECUDA);
    size_t got = 0;
    while (got < cap && lines < want_lines) {  // small reads: the head is a few KB to MB
      const size_t piece = cap - got < ((size_t)1 << 20) ? cap - got : ((size_t)1 << 20);
      long r = gf ? (long)gzread(gf, chunk + got, (unsigned)piece) : (long)read(fd, chunk + got, piece);
      if (r < 0 && !gf) return done(fail(ctx, FQGPU_EIO, std::string("read failed: ") + path));
      if (r <= 0) { eof = true; break; }  // (a corrupt gz stream ends here, like the reference's)
      const uint8_t* p = chunk + got;
      const uint8_t* end = p + r;
      while (p < end && lines < want_lines) {  // the stream ends with the newline that completes the last sampled line
        const uint8_t* nl = (const uint8_t*)memchr(p, '\n', (size_t)(end - p));
        if (!nl) { p = end; break; }
        lines++;
        p = nl + 1;
      }
      got = (size_t)(p - chunk);
    }
    if (got) {
      rc = fqgpu_submit(ctx, chunk, got);
      if (rc != FQGPU_OK) return done(rc);
    }
  }
  done(0);
  return fqgpu_finish(ctx, out);
}

// ---- one file over several contexts / GPUs in one process (file mode) -------------------------------
int fqgpu_count_file_sharded(const fqgpu_config* cfg, const char* path, const int* devices, int world, fqgpu_stats* out) {
  if (!path || !out) return FQGPU_EARG;
  fqgpu_config base;
  memset(&base, 0, sizeof base);
  base.device = -1;
  if (cfg) base = *cfg;
  const bool own_chunk = base.chunk_bytes != 0;
  const int ndev = fqgpu_device_count();
  if (world <= 0) world = ndev > 0 ? ndev : 1;
  if (world > 64) world = 64;
  auto dev_of = [&](int g) { return devices ? devices[g] : (ndev > 0 ? g % ndev : 0); };
  auto single = [&]() {  // everything the sharded path does not take
    fqgpu_config c = base;
    c.device = devices ? devices[0] : (base.device >= 0 ? base.device : -1);
    fqgpu_ctx* ctx = nullptr;
    int rc = fqgpu_create(&ctx, &c);
    if (rc != FQGPU_OK) return rc;
    rc = fqgpu_count_file(ctx, path, out);
    if (rc != FQGPU_OK) g_create_error = ctx->err;
    fqgpu_destroy(ctx);
    return rc;
  };
  const size_t L = strlen(path);
  const bool gz = L >= 3 && strcmp(path + L - 3, ".gz") == 0;
  struct stat sb;
  const int fd = gz ? -1 : open(path, O_RDONLY);
  if (fd < 0 || fstat(fd, &sb) != 0 || !S_ISREG(sb.st_mode) || world == 1 || (size_t)sb.st_size < (size_t)world * ((size_t)4 << 20)) {
    if (fd >= 0) close(fd);
    return single();
  }
  const size_t N = (size_t)sb.st_size;
  const size_t bw = fqgpu_shard_block_words();
  std::vector<uint64_t> blocks((size_t)world * bw, 0);
  std::vector<int> rcs((size_t)world, FQGPU_OK);
  std::vector<std::string> msgs((size_t)world);
  const int rthreads = read_threads() >= 2 * world ? read_threads() / world : 2;  // the reader threads are shared out
  auto shard = [&](int g) {
    fqgpu_config c = base;
    c.device = dev_of(g);
    if (!own_chunk) c.chunk_bytes = (size_t)16 << 20;  // several rings at once, each created for this call: smaller pinned chunks
    fqgpu_ctx* ctx = nullptr;
    int rc = fqgpu_create(&ctx, &c);
    if (rc != FQGPU_OK) { rcs[(size_t)g] = rc; msgs[(size_t)g] = g_create_error; return; }
    uint64_t* d_blocks = nullptr;
    auto finish = [&](int code) {
      if (code != FQGPU_OK) msgs[(size_t)g] = ctx->err;
      if (d_blocks) cudaFree(d_blocks);
      fqgpu_destroy(ctx);
      rcs[(size_t)g] = code;
    };
    if ((rc = fqgpu_shard_begin(ctx, g, world)) != FQGPU_OK) return finish(rc);
    size_t pos = g ? ((N * (size_t)g / (size_t)world) & ~(size_t)4095) : 0;
    const size_t end = g + 1 < world ? ((N * (size_t)(g + 1) / (size_t)world) & ~(size_t)4095) : N;
    while (pos < end) {
      size_t cap = 0;
      uint8_t* chunk = (uint8_t*)fqgpu_acquire(ctx, &cap);
      if (!chunk) return finish(FQGPU_ECUDA);
      const size_t want = end - pos < cap ? end - pos : cap;
      bool io_bad = false;
      const size_t got = parallel_pread(fd, chunk, want, pos, &io_bad, rthreads);
      if (io_bad || got != want) { ctx->err = std::string("read failed: ") + path; return finish(FQGPU_EIO); }
      if ((rc = fqgpu_submit(ctx, chunk, got)) != FQGPU_OK) return finish(rc);
      pos += got;
    }
    if (cudaSetDevice(ctx->device) != cudaSuccess || cudaMalloc(&d_blocks, (size_t)world * bw * sizeof(uint64_t)) != cudaSuccess) {
      ctx->err = "fqgpu_count_file_sharded: cudaMalloc failed";
      return finish(FQGPU_ECUDA);
    }
    cudaMemsetAsync(d_blocks, 0, (size_t)world * bw * sizeof(uint64_t), ctx->stream);
    if ((rc = fqgpu_shard_export(ctx, d_blocks)) != FQGPU_OK) return finish(rc);
    if (cudaMemcpyAsync(blocks.data() + (size_t)g * bw, d_blocks + (size_t)g * bw, bw * sizeof(uint64_t), cudaMemcpyDeviceToHost,
                        ctx->stream) != cudaSuccess || cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
      ctx->err = "fqgpu_count_file_sharded: copying the shard block failed";
      return finish(FQGPU_ECUDA);
    }
    finish(FQGPU_OK);
  };
  std::vector<std::thread> pool;
  for (int g = 1; g < world; g++) pool.emplace_back(shard, g);
  shard(0);
  for (auto& th : pool) th.join();
  close(fd);
  for (int g = 0; g < world; g++) if (rcs[(size_t)g] != FQGPU_OK) { g_create_error = msgs[(size_t)g]; return rcs[(size_t)g]; }
  const int rc = fqgpu_shard_combine_host(world, blocks.data(), base.meta_records, out);
  if (rc == FQGPU_ERETRY) return single();  // a shard resynced to a wrong phase (malformed input): one exact pass
  if (rc == FQGPU_OK && (base.flags & FQGPU_F_CORE_ONLY)) fqgpu_zero_quality(out);
  return rc;
}

// ---- many files at once (sc.nim:115-116) ------------------------------------------------------------
int fqgpu_count_files(const fqgpu_config* cfg, const char* const* paths, const int* as_gz, int n, int n_threads,
                      fqgpu_stats* out, int* rc) {
  if (n < 0 || (n > 0 && (!paths || !out || !rc))) return FQGPU_EARG;
  if (n == 0) return FQGPU_OK;
  fqgpu_config base;
  memset(&base, 0, sizeof base);
  base.device = -1;
  if (cfg) base = *cfg;
  if (base.chunk_bytes == 0) base.chunk_bytes = (size_t)16 << 20;  // several rings at once: smaller pinned chunks
  int ndev = 1, dev0 = base.device;
  if (base.device == FQGPU_DEVICE_ALL) {
    ndev = fqgpu_device_count();
    if (ndev <= 0) { g_create_error = "fqgpu_count_files: no CUDA device"; for (int i = 0; i < n; i++) rc[i] = FQGPU_ECUDA; return FQGPU_ECUDA; }
    dev0 = 0;
  }
  // default: up to 8 files in flight; `.gz` files, which are inflated on the device, keep the GPU busy one at a time and
  // pay for every context's buffers: two in flight (one is read while the other one's kernels run)
  bool any_gz = false;
  for (int i = 0; i < n && !any_gz; i++) {
    const size_t L = paths[i] ? strlen(paths[i]) : 0;
    any_gz = as_gz ? as_gz[i] != 0 : (L >= 3 && strcmp(paths[i] + L - 3, ".gz") == 0);
  }
  const int dflt = any_gz && !getenv("FQGPU_NO_GZIP_DEVICE") ? 2 * ndev : 8;
  int nthr = n_threads > 0 ? n_threads : (n < dflt ? n : dflt);
  if (nthr > n) nthr = n;
  for (int i = 0; i < n; i++) { rc[i] = FQGPU_ECUDA; memset(&out[i], 0, sizeof(fqgpu_stats)); }
  std::atomic<int> next(0);
  std::mutex mu;
  std::vector<std::string> msgs((size_t)n);  // what fqgpu_last_error(NULL) reports afterwards: the first failure in file order
  auto worker = [&](int t) {
    g_read_threads_share = nthr;  // the reader threads are shared out among the files in flight
    fqgpu_config c = base;
    if (base.device == FQGPU_DEVICE_ALL) c.device = dev0 + t % ndev;
    fqgpu_ctx* ctx = nullptr;
    const int crc = fqgpu_create(&ctx, &c);
    for (;;) {
      const int i = next.fetch_add(1);
      if (i >= n) break;
      if (crc != FQGPU_OK) { rc[i] = crc; std::lock_guard<std::mutex> g(mu); msgs[(size_t)i] = g_create_error; continue; }
      const char* p = paths[i];
      if (!p) { rc[i] = FQGPU_EARG; continue; }
      const size_t L = strlen(p);
      const int gz = as_gz ? as_gz[i] : (L >= 3 && strcmp(p + L - 3, ".gz") == 0);
      rc[i] = fqgpu_count_file_as(ctx, p, gz, &out[i]);
      if (rc[i] != FQGPU_OK) { std::lock_guard<std::mutex> g(mu); msgs[(size_t)i] = ctx->err; }
    }
    if (ctx) fqgpu_destroy(ctx);
    g_read_threads_share = 0;
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < nthr; t++) pool.emplace_back(worker, t);
  worker(0);
  for (auto& th : pool) th.join();
  for (int i = 0; i < n; i++) if (rc[i] != FQGPU_OK) { g_create_error = msgs[(size_t)i]; return rc[i]; }
  return FQGPU_OK;
}

// ---- paired files: R1 / R2 as ONE job -------------------------------------------------------
// The reference has no paired mode: `sc fq-count R1.fq.gz R2.fq.gz` is two passes of the sequential loop (sc.nim:115-116),
// one row per file.  Here both mates are scanned at the same time by two contexts (two host threads, two pinned rings,
// two streams on the same GPU, or one GPU each with FQGPU_DEVICE_ALL) and the rows come back in argument order; `paired`
// tells whether the two files can be mates at all (same number of records and of lines).
int fqgpu_count_pair(const fqgpu_config* cfg, const char* r1, const char* r2, fqgpu_stats* out1, fqgpu_stats* out2, int* paired) {
  if (!r1 || !r2 || !out1 || !out2) return FQGPU_EARG;
  const char* paths[2] = {r1, r2};
  fqgpu_stats st[2];
  int rc[2] = {FQGPU_ECUDA, FQGPU_ECUDA};
  const int r = fqgpu_count_files(cfg, paths, nullptr, 2, 2, st, rc);
  *out1 = st[0]; *out2 = st[1];
  if (paired) *paired = rc[0] == FQGPU_OK && rc[1] == FQGPU_OK && st[0].reads == st[1].reads && st[0].lines == st[1].lines;
  return r;
}

// ---- synthetic data -----------------------------------------------------------------------------
int fqgpu_synth_illumina(fqgpu_ctx* ctx, void* dptr, size_t capacity, uint64_t first_record,
                         uint64_t n_records, uint64_t seed, size_t* bytes_written) {
  if (!ctx || !dptr) return FQGPU_EARG;
  const u64 nbytes = n_records * 360ull;
  if (nbytes > capacity) return fail(ctx, FQGPU_EARG, "fqgpu_synth_illumina: capacity too small");
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  CU_TRY(ctx, fq::launch_synth_illumina(dptr, first_record * 360ull, nbytes, seed, ctx->stream));
  CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  if (bytes_written) *bytes_written = (size_t)nbytes;
  return FQGPU_OK;
}

int fqgpu_synth_illumina_bytes(fqgpu_ctx* ctx, void* dptr, uint64_t first_byte, uint64_t nbytes, uint64_t seed) {
  if (!ctx || !dptr) return FQGPU_EARG;
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  CU_TRY(ctx, fq::launch_synth_illumina(dptr, first_byte, nbytes, seed, ctx->stream));
  CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return FQGPU_OK;
}

int fqgpu_synth_illumina_tally(fqgpu_ctx* ctx, uint64_t first_record, uint64_t n_records, uint64_t seed, fqgpu_stats* out) {
  if (!ctx || !out) return FQGPU_EARG;
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  const size_t words = 16 + 152;
  u64* d_t = nullptr;
  CU_TRY(ctx, cudaMallocAsync((void**)&d_t, words * sizeof(u64), ctx->stream));
  std::vector<u64> h(words);
  cudaError_t e = cudaMemsetAsync(d_t, 0, words * sizeof(u64), ctx->stream);
  if (e == cudaSuccess) e = fq::launch_synth_illumina_tally(first_record, n_records, seed, d_t, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(h.data(), d_t, words * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream);
  cudaFreeAsync(d_t, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  CU_TRY(ctx, e);
  fqgpu_stats* st = out;
  memset(st, 0, sizeof(*st));
  static const char bases[5] = {'A', 'C', 'G', 'T', 'N'}, quals[4] = {'F', ':', ',', '#'};
  for (int k = 0; k < 5; k++) st->base_counts[(unsigned char)bases[k]] = h[k];
  for (int k = 0; k < 4; k++) st->qual_counts[(unsigned char)quals[k]] = h[5 + k];
  for (int p = 0; p < 150; p++) { st->qual_pos_sum[p] = h[16 + p]; st->qual_pos_cnt[p] = n_records; }
  st->bytes = n_records * 360ull; st->lines = 4 * n_records; st->reads = n_records; st->bases = 150ull * n_records;
  st->gc_bases = st->base_counts['G'] + st->base_counts['C']; st->n_bases = st->base_counts['N'];
  st->seq_lines = st->qual_lines = n_records;
  st->seq_len_min = st->qual_len_min = n_records ? 150 : UINT64_MAX;
  st->seq_len_max = st->qual_len_max = n_records ? 150 : 0;
  st->seq_len_hist[150] = st->qual_len_hist[150] = n_records;
  st->seq_len_log2[8] = n_records;
  const u64 m = ctx->cfg.meta_records < n_records ? ctx->cfg.meta_records : n_records;
  long long qmin, qmax;
  fq::synth_illumina_meta_range(first_record, m, seed, &qmin, &qmax);
  st->meta_qual_min = qmin; st->meta_qual_max = qmax; st->meta_lines = 4 * m;
  if (ctx->cfg.flags & FQGPU_F_CORE_ONLY) fqgpu_zero_quality(st);
  return FQGPU_OK;
}

int fqgpu_synth_ont(fqgpu_ctx* ctx, void* dptr, size_t capacity, uint64_t first_record,
                    uint64_t n_records, uint64_t seed, size_t* bytes_written) {
  if (!ctx || !dptr) return FQGPU_EARG;
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  size_t w = 0;
  cudaError_t e = fq::synth_ont(dptr, capacity, first_record, n_records, seed, &w, ctx->stream);
  if (e == cudaErrorInvalidValue) return fail(ctx, FQGPU_EARG, "fqgpu_synth_ont: capacity too small");
  CU_TRY(ctx, e);
  if (bytes_written) *bytes_written = w;
  return FQGPU_OK;
}


}  // extern "C"
