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

}  // extern "C"
