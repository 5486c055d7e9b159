// fq_shard.cu -- multi-GPU byte-range shard protocol (include/fqgpu.h).
//
// Rank g of `world` scans bytes [g*N/world, (g+1)*N/world) of one logical stream.  A rank > 0 does not
// know the line phase of its first byte: the CTA that takes the shard's first tile resynchronises on the content
// (resync_guess in fq_scan.cu), the shard is scanned under that hypothesis, and keeps what depends on the previous shards
// detached (fq::ShardInfo): the per-position sums and the length of its first line fragment.  Every
// rank exports one block of uint64 words into its slot of a buffer that the caller SUM-all-reduces
// (disjoint slots, so the sum is a gather) -- the one collective of the path.  The combine step then
// runs identically on every rank: exact line counts per shard verify each hypothesis, tail(g-1) and
// head(g) are stitched into one line, the head's per-position sums are shifted by the bytes the line
// had in earlier shards, a '\r' that ended a shard is attributed once its successor byte is known.
// A wrong hypothesis (malformed input) makes combine return FQGPU_ERETRY; fqgpu_shard_rescan() then
// installs the exact carry and the rank scans its range again.
#include <string.h>

#include "fqgpu_ctx.h"

using namespace fq;

static constexpr size_t kShardWords = (size_t)BLOCK_WORDS + SHARD_EXTRA_WORDS;

__global__ void fq_shard_pack_kernel(const u64* __restrict__ reduced, const Carry* __restrict__ carry,
                                     const ShardInfo* __restrict__ shard, int exact, u64* __restrict__ slot) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < BLOCK_WORDS) slot[i] = reduced[i];
  u64* ex = slot + BLOCK_WORDS;
  if (i <= POS_BINS) ex[SH_OFF_HEAD_POS + i] = shard->head_pos[i];
  if (i == 0) {
    u64* sc = ex + SH_OFF_SCALARS;
    const unsigned f = carry->flags;
    sc[SH_LINES] = carry->lines;
    sc[SH_BYTES] = carry->bytes;
    sc[SH_OPEN_LEN] = carry->open_len;
    sc[SH_LAST_BYTE] = carry->last_byte;
    sc[SH_FIRST_BYTE] = shard->first_byte;
    sc[SH_HEAD_LEN] = shard->head_len;
    sc[SH_HEAD_CR] = shard->head_cr;
    sc[SH_HYP] = (f >> CARRY_HYP_SHIFT) & 3u;
    sc[SH_HYP_VALID] = ((f & CARRY_HYP_VALID) && !(f & CARRY_HYP_FAILED)) ? 1 : 0;
    sc[SH_EXACT] = (exact || !(f & CARRY_UNKNOWN_START)) ? 1 : 0;
    sc[SH_META_LINES] = carry->meta_lines;
    sc[SH_META_QMIN] = (u64)carry->qual_min;
    sc[SH_META_QMAX] = (u64)carry->qual_max;
    sc[SH_META_STATUS] = carry->meta_status;
    sc[SH_META_PENDING_CR] = carry->meta_pending_cr;
    sc[SH_META_CUR_HAS] = (u64)(long long)carry->cur_has;
    sc[SH_META_CUR_MIN] = (u64)(long long)carry->cur_min;
    sc[SH_META_CUR_MAX] = (u64)(long long)carry->cur_max;
    sc[SH_PRESENT] = 1;
  }
}

__global__ void fq_shard_begin_kernel(Carry* carry, ShardInfo* shard, int unknown_start) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= POS_BINS) shard->head_pos[i] = 0;
  if (i == 0) {
    shard->head_len = 0; shard->head_cr = 0; shard->first_byte = 0;
    carry->flags = unknown_start ? CARRY_UNKNOWN_START : 0u;
  }
}

__global__ void fq_shard_set_carry_kernel(Carry* carry, u64 lines, u64 open_len, u64 bytes, unsigned last_byte) {
  carry->lines = lines; carry->open_len = open_len; carry->bytes = bytes; carry->last_byte = last_byte; carry->flags = 0;
}

static inline unsigned log2_bin_h(u64 len) { unsigned b = 0; while (len) { b++; len >>= 1; } return b; }

struct Chain {  // the stream state in front of a shard
  u64 lines, open, bytes;
  unsigned last_byte;
};

// Walks the shard slots in rank order.  Returns the index of the first shard whose hypothesis is wrong
// (or missing), or -1.  When `total` is given the exact stream totals are accumulated into it.
static int combine_walk(int world, const u64* blocks, u64* total, Carry* end_carry, Chain* before /*[world]*/) {
  Chain c{0, 0, 0, 0};
  int bad = -1;
  if (total) {
    memset(total, 0, BLOCK_WORDS * sizeof(u64));
    total[OFF_SEQ_LEN_MIN] = total[OFF_QUAL_LEN_MIN] = ~0ull;
  }
  for (int g = 0; g < world; g++) {
    const u64* S = blocks + (size_t)g * kShardWords;
    const u64* ex = S + BLOCK_WORDS;
    const u64* sc = ex + SH_OFF_SCALARS;
    if (before) before[g] = c;
    if (!sc[SH_PRESENT]) { if (bad < 0) bad = g; continue; }
    const bool exact = sc[SH_EXACT] != 0;
    if (!exact && (!sc[SH_HYP_VALID] || sc[SH_HYP] != (c.lines & 3))) { if (bad < 0) bad = g; }
    if (total) {
      for (int w = 0; w < BLOCK_WORDS; w++) {
        if (w == OFF_SEQ_LEN_MIN || w == OFF_QUAL_LEN_MIN) { if (S[w] < total[w]) total[w] = S[w]; }
        else if (w == OFF_SEQ_LEN_MAX || w == OFF_QUAL_LEN_MAX) { if (S[w] > total[w]) total[w] = S[w]; }
        else total[w] += S[w];
      }
      if (!exact && sc[SH_BYTES]) {  // stitch this shard's first line fragment to the stream so far
        const int cls = (int)(c.lines & 3);
        const u64 P0 = c.open;
        // the '\r' that ended the previous shard is content unless this shard starts with '\n'
        const bool prev_cr = c.bytes && c.open && c.last_byte == '\r';
        if (prev_cr && sc[SH_FIRST_BYTE] != '\n' && (cls & 1)) {
          total[(cls == 3 ? OFF_HIST_QUAL : OFF_HIST_SEQ) + '\r'] += 1;
          if (cls == 3) total[OFF_POS_SUM + ((P0 - 1) < (u64)POS_BINS ? (P0 - 1) : (u64)POS_BINS)] += '\r';
        }
        if (cls == 3) {
          for (int p = 0; p <= POS_BINS; p++) {
            const u64 v = ex[SH_OFF_HEAD_POS + p];
            if (!v) continue;
            const u64 q = (u64)p + P0;
            total[OFF_POS_SUM + (p < POS_BINS && q < (u64)POS_BINS ? q : (u64)POS_BINS)] += v;
          }
        }
        if (sc[SH_LINES] && (cls & 1)) {  // the line ends in this shard: its length
          const u64 raw = P0 + sc[SH_HEAD_LEN];
          const u64 cr = sc[SH_HEAD_LEN] ? sc[SH_HEAD_CR] : ((prev_cr && raw > 0) ? 1 : 0);
          const u64 len = raw - cr;
          const u64 bin = len < (u64)POS_BINS ? len : (u64)POS_BINS;
          if (cls == 3) {
            total[OFF_QUAL_LEN + bin] += 1;
            if (len < total[OFF_QUAL_LEN_MIN]) total[OFF_QUAL_LEN_MIN] = len;
            if (len > total[OFF_QUAL_LEN_MAX]) total[OFF_QUAL_LEN_MAX] = len;
          } else {
            total[OFF_SEQ_LEN + bin] += 1;
            total[OFF_SEQ_LOG2 + log2_bin_h(len)] += 1;
            if (len < total[OFF_SEQ_LEN_MIN]) total[OFF_SEQ_LEN_MIN] = len;
            if (len > total[OFF_SEQ_LEN_MAX]) total[OFF_SEQ_LEN_MAX] = len;
          }
        }
      }
    }
    // advance the chain.  A hypothesis shard counts lines / bytes from its own start; an exact (rescanned
    // or rank-0) shard was scanned with the true carry installed, so its carry is absolute.
    if (exact && g > 0) {
      c.lines = sc[SH_LINES]; c.open = sc[SH_OPEN_LEN]; c.bytes = sc[SH_BYTES];
    } else {
      if (sc[SH_LINES]) { c.lines += sc[SH_LINES]; c.open = sc[SH_OPEN_LEN]; }
      else c.open += sc[SH_BYTES];
      c.bytes += sc[SH_BYTES];
    }
    if (sc[SH_BYTES]) c.last_byte = (unsigned)sc[SH_LAST_BYTE];
  }
  if (end_carry) {
    memset(end_carry, 0, sizeof(*end_carry));
    end_carry->lines = c.lines; end_carry->open_len = c.open; end_carry->bytes = c.bytes; end_carry->last_byte = c.last_byte;
    // fq-meta fold: exact when the sampled prefix lies inside shard 0 (its state is complete)
    const u64* sc0 = blocks + BLOCK_WORDS + SH_OFF_SCALARS;
    end_carry->meta_lines = sc0[SH_META_LINES];
    end_carry->qual_min = (long long)sc0[SH_META_QMIN];
    end_carry->qual_max = (long long)sc0[SH_META_QMAX];
    end_carry->meta_status = (unsigned)sc0[SH_META_STATUS];
    end_carry->meta_pending_cr = (unsigned)sc0[SH_META_PENDING_CR];
    end_carry->cur_has = (int)(long long)sc0[SH_META_CUR_HAS];
    end_carry->cur_min = (int)(long long)sc0[SH_META_CUR_MIN];
    end_carry->cur_max = (int)(long long)sc0[SH_META_CUR_MAX];
  }
  return bad;
}

// ---------------------------------------------------------------------------------------------------------------
// The collective inside the library: ONE launch per rank packs the rank's block, stores it into slot `rank` of EVERY
// rank's exchange buffer through peer-mapped memory (NVLink / NVSwitch P2P stores: an all-gather with no library
// call), publishes a flag behind a system-scope fence, waits for the other ranks' flags in its own memory, and
// combines the gathered blocks right there (same arithmetic as combine_walk, spread over the CTA).  The host reads
// back one block + the carry: one kernel, one host synchronisation per step.
//
// Exchange buffer of a rank (device memory, opened by the peers through CUDA IPC or handed over as pointers):
//   slots  [2 parities][world][kShardWords] u64   blocks of the step, double-buffered by the step's parity: a peer can
//                                                 be one step ahead (it needs this rank's flag for the step after)
//   flags  [2][64] u64                            step number each rank has published for that parity
//   result [BLOCK_WORDS] u64 + Carry + 4 u64      the combined block, the end carry, first wrong rank + 1 (0: none)
// ---------------------------------------------------------------------------------------------------------------
static constexpr int kMaxWorld = 64;
struct XbufLayout {
  size_t slots, flags, result, bytes;
};
static inline XbufLayout xbuf_layout(int world) {
  XbufLayout L;
  L.slots = 0;
  L.flags = 2 * (size_t)world * kShardWords * sizeof(u64);
  L.result = L.flags + 2 * kMaxWorld * sizeof(u64);
  L.bytes = L.result + (BLOCK_WORDS + 4) * sizeof(u64) + sizeof(Carry);
  return L;
}
struct PeerPtrs { uint8_t* p[kMaxWorld]; };

__device__ __forceinline__ void st_release_sys(u64* p, u64 v) { asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ u64 ld_acquire_sys(const u64* p) {
  u64 v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned log2_bin_d(u64 len) { return len ? 64 - __clzll((long long)len) : 0; }

__global__ void __launch_bounds__(512) fq_shard_exchange_kernel(const u64* __restrict__ acc, const Carry* __restrict__ carry,
                                                              const ShardInfo* __restrict__ shard, int exact, int rank, int world, u64 step,
                                                              PeerPtrs peers, size_t off_flags, size_t off_result, u64 meta_records, u64 spin_limit) {
  __shared__ u64 sc_all[kMaxWorld][8];   // per rank: lines, bytes, open_len, last_byte, first_byte, head_len, head_cr, flags (exact | hyp_valid << 1 | hyp << 2 | present << 4)
  __shared__ u64 before[kMaxWorld][4];   // chain in front of each rank: lines, open, bytes, last_byte
  __shared__ int s_timeout;
  const int tid = threadIdx.x;
  const u64 par = step & 1;
  // ---- 1. this rank's block into slot `rank` of every rank's buffer ----
  for (int i = tid; i < (int)kShardWords; i += blockDim.x) {
    u64 v = 0;
    if (i < BLOCK_WORDS) v = acc[i];
    else {
      const int e = i - BLOCK_WORDS;
      if (e <= POS_BINS) v = shard->head_pos[e];
      else if (e >= SH_OFF_SCALARS && e < SH_OFF_SCALARS + SH_NSCALARS) {
        const unsigned f = carry->flags;
        switch (e - SH_OFF_SCALARS) {
          case SH_LINES: v = carry->lines; break;
          case SH_BYTES: v = carry->bytes; break;
          case SH_OPEN_LEN: v = carry->open_len; break;
          case SH_LAST_BYTE: v = carry->last_byte; break;
          case SH_FIRST_BYTE: v = shard->first_byte; break;
          case SH_HEAD_LEN: v = shard->head_len; break;
          case SH_HEAD_CR: v = shard->head_cr; break;
          case SH_HYP: v = (f >> CARRY_HYP_SHIFT) & 3u; break;
          case SH_HYP_VALID: v = ((f & CARRY_HYP_VALID) && !(f & CARRY_HYP_FAILED)) ? 1 : 0; break;
          case SH_EXACT: v = (exact || !(f & CARRY_UNKNOWN_START)) ? 1 : 0; break;
          case SH_META_LINES: v = carry->meta_lines; break;
          case SH_META_QMIN: v = (u64)carry->qual_min; break;
          case SH_META_QMAX: v = (u64)carry->qual_max; break;
          case SH_META_STATUS: v = carry->meta_status; break;
          case SH_META_PENDING_CR: v = carry->meta_pending_cr; break;
          case SH_META_CUR_HAS: v = (u64)(long long)carry->cur_has; break;
          case SH_META_CUR_MIN: v = (u64)(long long)carry->cur_min; break;
          case SH_META_CUR_MAX: v = (u64)(long long)carry->cur_max; break;
          case SH_PRESENT: v = 1; break;
        }
      }
    }
    for (int p = 0; p < world; p++)
      reinterpret_cast<u64*>(peers.p[p])[(par * world + rank) * kShardWords + i] = v;
  }
  __threadfence_system();
  __syncthreads();
  if (tid < world) st_release_sys(reinterpret_cast<u64*>(peers.p[tid] + off_flags) + par * kMaxWorld + rank, step);
  // ---- 2. wait for everybody's block of this step in this rank's own memory ----
  if (tid == 0) s_timeout = 0;
  __syncthreads();
  if (tid < world) {
    const u64* fl = reinterpret_cast<const u64*>(peers.p[rank] + off_flags) + par * kMaxWorld + tid;
    const long long t0 = clock64();
    while (ld_acquire_sys(fl) != step) {
      if ((u64)(clock64() - t0) > spin_limit) { s_timeout = 1; break; }
      __nanosleep(100);
    }
  }
  __syncthreads();
  const u64* slots = reinterpret_cast<const u64*>(peers.p[rank]) + par * world * kShardWords;
  u64* res = reinterpret_cast<u64*>(peers.p[rank] + off_result);
  // ---- 3. the chain in front of every rank (scalars, in rank order) ----
  if (tid < world) {
    const u64* sc = slots + (size_t)tid * kShardWords + BLOCK_WORDS + SH_OFF_SCALARS;
    sc_all[tid][0] = sc[SH_LINES]; sc_all[tid][1] = sc[SH_BYTES]; sc_all[tid][2] = sc[SH_OPEN_LEN]; sc_all[tid][3] = sc[SH_LAST_BYTE];
    sc_all[tid][4] = sc[SH_FIRST_BYTE]; sc_all[tid][5] = sc[SH_HEAD_LEN]; sc_all[tid][6] = sc[SH_HEAD_CR];
    sc_all[tid][7] = (sc[SH_EXACT] ? 1 : 0) | (sc[SH_HYP_VALID] ? 2 : 0) | ((sc[SH_HYP] & 3) << 2) | (sc[SH_PRESENT] ? 16 : 0);
  }
  __syncthreads();
  if (tid == 0) {
    u64 lines = 0, open = 0, bytes = 0, last_byte = 0;
    int bad = s_timeout ? 0 : -1;
    for (int g = 0; g < world; g++) {
      before[g][0] = lines; before[g][1] = open; before[g][2] = bytes; before[g][3] = last_byte;
      const u64* s = sc_all[g];
      const bool present = (s[7] & 16) != 0, exact_g = (s[7] & 1) != 0;
      if (!present) { if (bad < 0) bad = g; continue; }
      if (!exact_g && (!(s[7] & 2) || ((s[7] >> 2) & 3) != (lines & 3))) { if (bad < 0) bad = g; }
      if (exact_g && g > 0) { lines = s[0]; open = s[2]; bytes = s[1]; }
      else {
        if (s[0]) { lines += s[0]; open = s[2]; } else open += s[1];
        bytes += s[1];
      }
      if (s[1]) last_byte = s[3];
    }
    Carry c;
    memset(&c, 0, sizeof(c));
    c.lines = lines; c.open_len = open; c.bytes = bytes; c.last_byte = (unsigned)last_byte;
    const u64* sc0 = slots + BLOCK_WORDS + SH_OFF_SCALARS;  // fq-meta fold: shard 0's state
    c.meta_lines = sc0[SH_META_LINES]; c.qual_min = (long long)sc0[SH_META_QMIN]; c.qual_max = (long long)sc0[SH_META_QMAX];
    c.meta_status = (unsigned)sc0[SH_META_STATUS]; c.meta_pending_cr = (unsigned)sc0[SH_META_PENDING_CR];
    c.cur_has = (int)(long long)sc0[SH_META_CUR_HAS]; c.cur_min = (int)(long long)sc0[SH_META_CUR_MIN]; c.cur_max = (int)(long long)sc0[SH_META_CUR_MAX];
    // the sampled prefix extends beyond shard 0: flagged by the host AFTER the trailing line has been folded
    res[BLOCK_WORDS + 2] = (meta_records && c.meta_lines < meta_records * 4 && world > 1 && sc_all[0][1] < bytes) ? 1 : 0;
    *reinterpret_cast<Carry*>(res + BLOCK_WORDS + 4) = c;
    res[BLOCK_WORDS] = (u64)(bad + 1);
    res[BLOCK_WORDS + 1] = (u64)s_timeout;
  }
  __syncthreads();
  // ---- 4. the blocks, word by word; then the first line fragment of every hypothesis shard stitched to the stream ----
  for (int w = tid; w < BLOCK_WORDS; w += blockDim.x) {
    const bool is_min = w == OFF_SEQ_LEN_MIN || w == OFF_QUAL_LEN_MIN, is_max = w == OFF_SEQ_LEN_MAX || w == OFF_QUAL_LEN_MAX;
    u64 t = is_min ? ~0ull : 0ull;
    for (int g = 0; g < world; g++) {
      if (!(sc_all[g][7] & 16)) continue;
      const u64 x = slots[(size_t)g * kShardWords + w];
      if (is_min) t = x < t ? x : t; else if (is_max) t = x > t ? x : t; else t += x;
    }
    res[w] = t;
  }
  __syncthreads();
  for (int g = 0; g < world; g++) {
    const u64* s = sc_all[g];
    if (!(s[7] & 16) || (s[7] & 1) || !s[1]) continue;  // absent, exact, or empty
    const int cls = (int)(before[g][0] & 3);
    const u64 P0 = before[g][1];
    const bool prev_cr = before[g][2] && before[g][1] && before[g][3] == '\r';
    if (cls == 3) {  // the head's per-position sums move by the bytes the line had on the ranks before
      const u64* hp = slots + (size_t)g * kShardWords + BLOCK_WORDS + SH_OFF_HEAD_POS;
      for (int p = tid; p <= POS_BINS; p += blockDim.x) {
        const u64 v = hp[p];
        if (!v) continue;
        const u64 q = (u64)p + P0;
        atomicAdd(&res[OFF_POS_SUM + (p < POS_BINS && q < (u64)POS_BINS ? q : (u64)POS_BINS)], v);
      }
    }
    if (tid == 0) {
      if (prev_cr && s[4] != '\n' && (cls & 1)) {  // the '\r' that ended the previous shard is content unless this shard starts with '\n'
        atomicAdd(&res[(cls == 3 ? OFF_HIST_QUAL : OFF_HIST_SEQ) + '\r'], 1ull);
        if (cls == 3) atomicAdd(&res[OFF_POS_SUM + ((P0 - 1) < (u64)POS_BINS ? (P0 - 1) : (u64)POS_BINS)], (u64)'\r');
      }
      if (s[0] && (cls & 1)) {  // the line ends in this shard: its length
        const u64 raw = P0 + s[5];
        const u64 cr = s[5] ? s[6] : ((prev_cr && raw > 0) ? 1 : 0);
        const u64 len = raw - cr;
        const u64 bin = len < (u64)POS_BINS ? len : (u64)POS_BINS;
        if (cls == 3) {
          atomicAdd(&res[OFF_QUAL_LEN + bin], 1ull);
          atomicMin(&res[OFF_QUAL_LEN_MIN], len); atomicMax(&res[OFF_QUAL_LEN_MAX], len);
        } else {
          atomicAdd(&res[OFF_SEQ_LEN + bin], 1ull);
          atomicAdd(&res[OFF_SEQ_LOG2 + log2_bin_d(len)], 1ull);
          atomicMin(&res[OFF_SEQ_LEN_MIN], len); atomicMax(&res[OFF_SEQ_LEN_MAX], len);
        }
      }
    }
    __syncthreads();
  }
}

extern "C" {

size_t fqgpu_shard_block_words(void) { return kShardWords; }

int fqgpu_shard_begin(fqgpu_ctx* ctx, int rank, int world) {
  if (!ctx || world < 1 || world > 64 || rank < 0 || rank >= world) return fail(ctx, FQGPU_EARG, "fqgpu_shard_begin: bad rank/world");
  int rc = fqgpu_reset(ctx);
  if (rc != FQGPU_OK) return rc;
  ctx->shard_rank = rank;
  ctx->shard_world = world;
  ctx->shard_exact = false;
  fq_shard_begin_kernel<<<(POS_BINS + 256) / 256, 256, 0, ctx->stream>>>(ctx->d_carry, ctx->d_shard, rank > 0);
  CU_TRY(ctx, cudaGetLastError());
  return FQGPU_OK;
}

int fqgpu_shard_export(fqgpu_ctx* ctx, uint64_t* d_blocks) {
  if (!ctx || !d_blocks) return FQGPU_EARG;
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  cudaEvent_t e0 = fqgpu_get_event(ctx), e1 = fqgpu_get_event(ctx);
  CU_TRY(ctx, cudaEventRecord(e0, ctx->stream));
  u64* slot = (u64*)d_blocks + (size_t)ctx->shard_rank * kShardWords;
  fq_shard_pack_kernel<<<(BLOCK_WORDS + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_acc, ctx->d_carry, ctx->d_shard,
                                                                           ctx->shard_exact ? 1 : 0, slot);
  CU_TRY(ctx, cudaGetLastError());
  CU_TRY(ctx, cudaEventRecord(e1, ctx->stream));
  ctx->timed.emplace_back(e0, e1);
  return FQGPU_OK;
}

// Host-only combine over gathered blocks in HOST memory (also what the CPU tests drive).
int fqgpu_shard_combine_host(int world, const uint64_t* h_blocks, uint64_t meta_records, fqgpu_stats* out) {
  if (!h_blocks || !out || world < 1) return FQGPU_EARG;
  u64 total[BLOCK_WORDS];
  Carry endc;
  const int bad = combine_walk(world, (const u64*)h_blocks, total, &endc, nullptr);
  if (bad >= 0) return FQGPU_ERETRY;
  fqgpu_assemble_stats(total, endc, meta_records, out);
  if (meta_records && endc.meta_lines < meta_records * 4 && world > 1) {
    // the sampled prefix extends beyond shard 0: the range fields only cover what shard 0 saw
    const u64* sc0 = (const u64*)h_blocks + BLOCK_WORDS + SH_OFF_SCALARS;
    if (sc0[SH_BYTES] < endc.bytes) out->meta_status |= 0x100u;  // incomplete
  }
  return FQGPU_OK;
}

int fqgpu_shard_combine(fqgpu_ctx* ctx, const uint64_t* d_blocks, fqgpu_stats* out) {
  if (!ctx || !d_blocks || !out) return FQGPU_EARG;
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  const size_t bytes = (size_t)ctx->shard_world * kShardWords * sizeof(u64);
  CU_TRY(ctx, cudaMemcpyAsync(ctx->h_shard, d_blocks, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  const int rc = fqgpu_shard_combine_host(ctx->shard_world, (const uint64_t*)ctx->h_shard, ctx->cfg.meta_records, out);
  if (rc == FQGPU_OK && (ctx->cfg.flags & FQGPU_F_CORE_ONLY)) fqgpu_zero_quality(out);
  return rc;
}

// After a combine that returned FQGPU_ERETRY: returns FQGPU_OK when this rank's block stands (it only
// has to be exported again), or FQGPU_ERETRY after installing the exact carry in front of this rank --
// the caller then scans the rank's byte range again and exports.
int fqgpu_shard_rescan(fqgpu_ctx* ctx, const uint64_t* d_blocks) {
  if (!ctx || !d_blocks) return FQGPU_EARG;
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  const int world = ctx->shard_world, g = ctx->shard_rank;
  const size_t bytes = (size_t)world * kShardWords * sizeof(u64);
  CU_TRY(ctx, cudaMemcpyAsync(ctx->h_shard, d_blocks, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  Chain before[64];
  const int bad = combine_walk(world, ctx->h_shard, nullptr, nullptr, before);
  if (bad < 0 || g < bad) return FQGPU_OK;
  // every shard from the first wrong one on is rescanned exactly: the chain in front of a shard is only
  // known exactly up to the first wrong shard, so one shard is repaired per round
  if (g != bad) return FQGPU_OK;
  const Chain c = before[g];
  CU_TRY(ctx, launch_reset(ctx->d_acc, ctx->d_carry, ctx->d_ctl, ctx->stream));
  fq_shard_begin_kernel<<<(POS_BINS + 256) / 256, 256, 0, ctx->stream>>>(ctx->d_carry, ctx->d_shard, 0);
  fq_shard_set_carry_kernel<<<1, 1, 0, ctx->stream>>>(ctx->d_carry, c.lines, c.open, c.bytes, c.last_byte);
  CU_TRY(ctx, cudaGetLastError());
  ctx->shard_exact = true;
  return FQGPU_ERETRY;
}


// ---- the collective inside the library (see fq_shard_exchange_kernel) ---------------------------------------------
size_t fqgpu_shard_xbuf_bytes(int world) { return world >= 1 && world <= kMaxWorld ? xbuf_layout(world).bytes : 0; }
size_t fqgpu_ipc_handle_bytes(void) { return sizeof(cudaIpcMemHandle_t); }

static void xbuf_release(fqgpu_ctx* ctx) {
  for (int g = 0; g < kMaxWorld; g++) {
    if (ctx->x_peer_ipc[g]) cudaIpcCloseMemHandle(ctx->x_peers[g]);
    ctx->x_peer_ipc[g] = false;
    ctx->x_peers[g] = nullptr;
  }
  if (ctx->x_buf) cudaFree(ctx->x_buf);
  ctx->x_buf = nullptr;
  ctx->x_world = 0;
  ctx->x_step = 0;
}
extern "C" void fqgpu_shard_exchange_destroy(fqgpu_ctx* ctx) { if (ctx) { cudaSetDevice(ctx->device); xbuf_release(ctx); } }

int fqgpu_shard_exchange_create(fqgpu_ctx* ctx, int rank, int world, void* ipc_handle_out) {
  if (!ctx || world < 1 || world > kMaxWorld || rank < 0 || rank >= world) return fail(ctx, FQGPU_EARG, "fqgpu_shard_exchange_create: bad rank/world");
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  xbuf_release(ctx);
  const XbufLayout L = xbuf_layout(world);
  CU_TRY(ctx, cudaMalloc(&ctx->x_buf, L.bytes));
  CU_TRY(ctx, cudaMemset(ctx->x_buf, 0, L.bytes));
  ctx->x_world = world; ctx->x_rank = rank; ctx->x_step = 0;
  ctx->x_peers[rank] = ctx->x_buf;
  if (!ctx->h_xres) CU_TRY(ctx, cudaMallocHost(&ctx->h_xres, (BLOCK_WORDS + 4) * sizeof(u64) + sizeof(Carry)));
  if (ipc_handle_out) {
    cudaIpcMemHandle_t h;
    CU_TRY(ctx, cudaIpcGetMemHandle(&h, ctx->x_buf));
    memcpy(ipc_handle_out, &h, sizeof h);
  }
  return FQGPU_OK;
}

int fqgpu_shard_exchange_open(fqgpu_ctx* ctx, const void* ipc_handles) {
  if (!ctx || !ipc_handles || !ctx->x_buf) return fail(ctx, FQGPU_EARG, "fqgpu_shard_exchange_open: create the exchange buffer first");
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  for (int g = 0; g < ctx->x_world; g++) {
    if (g == ctx->x_rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, (const uint8_t*)ipc_handles + (size_t)g * sizeof h, sizeof h);
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { ctx->err = std::string("cudaIpcOpenMemHandle(rank ") + std::to_string(g) + "): " + cudaGetErrorString(e); cudaGetLastError(); return FQGPU_ENCCL; }
    ctx->x_peers[g] = p;
    ctx->x_peer_ipc[g] = true;
  }
  return FQGPU_OK;
}

void* fqgpu_shard_xbuf(fqgpu_ctx* ctx) { return ctx ? ctx->x_buf : nullptr; }

int fqgpu_shard_exchange_set_peers(fqgpu_ctx* ctx, void* const* xbufs) {
  if (!ctx || !xbufs || !ctx->x_buf) return fail(ctx, FQGPU_EARG, "fqgpu_shard_exchange_set_peers: create the exchange buffer first");
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  for (int g = 0; g < ctx->x_world; g++) {
    if (g == ctx->x_rank) continue;
    if (!xbufs[g]) return fail(ctx, FQGPU_EARG, "fqgpu_shard_exchange_set_peers: NULL peer buffer");
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, xbufs[g]) == cudaSuccess && at.device != ctx->device) {
      cudaError_t e = cudaDeviceEnablePeerAccess(at.device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { ctx->err = std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e); cudaGetLastError(); return FQGPU_ENCCL; }
      cudaGetLastError();
    }
    ctx->x_peers[g] = xbufs[g];
  }
  return FQGPU_OK;
}

// Starts this rank's part of the collective (asynchronous: every rank must start it before any can finish).
int fqgpu_shard_exchange_start(fqgpu_ctx* ctx) {
  NvtxRange nvtx_("fqgpu_shard_exchange_start");
  if (!ctx || !ctx->x_buf) return fail(ctx, FQGPU_EARG, "fqgpu_shard_exchange: no exchange buffer");
  for (int g = 0; g < ctx->x_world; g++) if (!ctx->x_peers[g]) return fail(ctx, FQGPU_EARG, "fqgpu_shard_exchange: peers not opened");
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  const XbufLayout L = xbuf_layout(ctx->x_world);
  PeerPtrs pp;
  memset(&pp, 0, sizeof pp);
  for (int g = 0; g < ctx->x_world; g++) pp.p[g] = (uint8_t*)ctx->x_peers[g];
  ctx->x_step++;
  // (not part of fqgpu_last_timing: the kernel's duration is mostly the wait for the slowest rank, not work)
  fq_shard_exchange_kernel<<<1, 512, 0, ctx->stream>>>(ctx->d_acc, ctx->d_carry, ctx->d_shard, ctx->shard_exact ? 1 : 0, ctx->x_rank, ctx->x_world,
                                                    ctx->x_step, pp, L.flags, L.result, ctx->cfg.meta_records, 20ull * 2000000000ull);
  CU_TRY(ctx, cudaGetLastError());
  CU_TRY(ctx, cudaMemcpyAsync(ctx->h_xres, (uint8_t*)ctx->x_buf + L.result, (BLOCK_WORDS + 4) * sizeof(u64) + sizeof(Carry), cudaMemcpyDeviceToHost, ctx->stream));
  return FQGPU_OK;
}
// Waits for it and assembles the statistics: FQGPU_OK, or FQGPU_ERETRY when a rank's hypothesis was wrong (then
// fqgpu_shard_rescan(ctx, fqgpu_shard_gathered(ctx)) as after fqgpu_shard_combine), or FQGPU_ENCCL when a rank never showed up.
int fqgpu_shard_exchange_finish(fqgpu_ctx* ctx, fqgpu_stats* out) {
  NvtxRange nvtx_("fqgpu_shard_exchange_finish");
  if (!ctx || !out || !ctx->x_buf) return FQGPU_EARG;
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  const u64* r = ctx->h_xres;
  if (r[BLOCK_WORDS + 1]) return fail(ctx, FQGPU_ENCCL, "fqgpu_shard_exchange: a rank did not publish its block (20 s)");
  if (r[BLOCK_WORDS]) return FQGPU_ERETRY;  // rank r[BLOCK_WORDS] - 1 scanned under a wrong phase hypothesis
  Carry c;
  memcpy(&c, r + BLOCK_WORDS + 4, sizeof c);
  fqgpu_assemble_stats(r, c, ctx->cfg.meta_records, out);
  if (r[BLOCK_WORDS + 2]) out->meta_status |= 0x100u;  // incomplete fq-meta range (see fqgpu_shard_combine_host)
  if (ctx->cfg.flags & FQGPU_F_CORE_ONLY) fqgpu_zero_quality(out);
  return FQGPU_OK;
}
int fqgpu_shard_exchange_combine(fqgpu_ctx* ctx, fqgpu_stats* out) {
  int rc = fqgpu_shard_exchange_start(ctx);
  return rc == FQGPU_OK ? fqgpu_shard_exchange_finish(ctx, out) : rc;
}
// The blocks gathered by the most recent exchange (device memory of this rank; what fqgpu_shard_rescan takes).
const uint64_t* fqgpu_shard_gathered(fqgpu_ctx* ctx) {
  if (!ctx || !ctx->x_buf) return nullptr;
  return (const uint64_t*)ctx->x_buf + (ctx->x_step & 1) * (size_t)ctx->x_world * kShardWords;
}

}  // extern "C"
