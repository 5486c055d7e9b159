// fq_scan.cu -- the FASTQ scanning hot path on sm_100a.
//
// Replaces the per-line loop of the reference (src/fq_count.nim:38-45: `for line in lines(stream)`,
// i mod 4 classing, count("G")+count("C"), count("N"), line.len) and the quality fold of
// src/fq_meta.nim:245-246 with one pass over the bytes:
//
//   K1  boundary classification: 16-byte vector loads, SWAR '\n' detection, warp prefix sums
//       -> newline index of the tile; the absolute line number and the bytes of the open line at
//       the tile start come from a single-pass decoupled look-back over tile descriptors, so the
//       chunk-edge / tile-edge record carry is resolved on the device (fq::Carry, fq::TileState).
//   K2  per-line statistics from the shared-memory resident tile (fused with K1 in one kernel so
//       every input byte is read from HBM exactly once): lane-striped (bank-conflict-free)
//       shared-memory histograms of sequence / quality bytes, length tables, per-position sums.
//   K3  reduction of the per-CTA partial counter blocks (fq_reduce_kernel).
//
// Line semantics are Nim's streams.lines: split at '\n', drop one '\r' directly before it; the
// trailing unterminated line is accounted by the host from fq::Carry at finish().
#include <cuda_runtime.h>
#include <stdint.h>

#include "fq_layout.h"

namespace fq {

typedef unsigned long long u64;

constexpr int TILE = 16384;
constexpr int THREADS = 512;
constexpr int NWARPS = THREADS / 32;
constexpr int GPT = TILE / 16 / THREADS;  // 16-byte groups per thread
constexpr int NROWS = GPT * NWARPS;       // 512-byte rows per tile
constexpr int NL_CAP = 2048;              // newline-index window (tiles with more are done in passes)
constexpr int PAD = 16;                   // bytes kept in front of the tile (look-behind byte at PAD-1)
constexpr int LONG_SEG = 2048;            // pieces longer than this are processed by the whole CTA
constexpr int LONG_CAP = TILE / LONG_SEG + 2;
static_assert(NROWS == 32, "row prefix is one warp scan");

constexpr u64 ST_MASK = 3ull << 62;
constexpr u64 ST_AGG = 1ull << 62;
constexpr u64 ST_PFX = 2ull << 62;
constexpr u64 HAS_NL = 1ull << 61;
constexpr u64 VAL_MASK = (1ull << 61) - 1;

struct LongSeg {
  int vs, ve, cls, pad;
  u64 vpos;
};

struct __align__(16) Smem {
  uint8_t buf[PAD + TILE];
  uint16_t nl[NL_CAP];
  uint32_t hist[2][256 * 32];  // [0] sequence, [1] quality; index = byte*32 + lane
  uint32_t seq_len[POS_BINS + 1];
  uint32_t qual_len[POS_BINS + 1];
  uint32_t seq_log2[LOG2_BINS];
  uint32_t pos_sum[POS_BINS + 1];
  uint32_t rowcnt[NROWS];
  uint32_t rowbase[NROWS];
  LongSeg longs[LONG_CAP];
  u64 L0, P0;
  u64 len_min[2], len_max[2];  // [0] seq, [1] qual
  uint32_t tile_id, T, nlong, bytes_since_flush;
  int last_nl, prev_nl;
};

__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ u64 ld_relaxed(const u64* p) {
  u64 v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed(u64* p, u64 v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// 4-bit mask of the bytes of w equal to '\n' (exact, no carries between byte lanes)
__device__ __forceinline__ uint32_t nl_bits(uint32_t w) {
  uint32_t x = w ^ 0x0A0A0A0Au;
  uint32_t m = ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u;
  return (((m >> 7) * 0x00204081u) >> 21) & 0xFu;
}
__device__ __forceinline__ uint32_t nl_mask16(const uint4& v) {
  return nl_bits(v.x) | (nl_bits(v.y) << 4) | (nl_bits(v.z) << 8) | (nl_bits(v.w) << 12);
}
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t n = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v += n;
  }
  return v;
}
__device__ __forceinline__ u64 warp_sum64(u64 v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}
__device__ __forceinline__ unsigned log2_bin(u64 len) { return len ? 64 - __clzll((long long)len) : 0; }

// One byte of a sequence (cls 1) / quality (cls 3) line.
__device__ __forceinline__ void account_byte(Smem& sm, int cls, uint32_t b, u64 pos, int lane) {
  atomicAdd(&sm.hist[cls == 3][(b << 5) + lane], 1u);
  if (cls == 3) {
    uint32_t p = pos < (u64)POS_BINS ? (uint32_t)pos : (uint32_t)POS_BINS;
    atomicAdd(&sm.pos_sum[p], b);
  }
}

__device__ __forceinline__ void flush_pos_sum(Smem& sm, u64* block, int tid) {
  for (int i = tid; i <= POS_BINS; i += THREADS) {
    uint32_t v = sm.pos_sum[i];
    if (v) { block[OFF_POS_SUM + i] += v; sm.pos_sum[i] = 0; }
  }
}

__global__ void __launch_bounds__(THREADS, 2)
fq_scan_kernel(const uint8_t* __restrict__ base, uint32_t lo0, u64 end, uint32_t ntiles,
               TileState* __restrict__ ts, LaunchInfo* __restrict__ info, Carry* __restrict__ carry,
               u64* __restrict__ partials) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  u64* block = partials + (size_t)blockIdx.x * BLOCK_WORDS;

  for (int i = tid; i < 2 * 256 * 32; i += THREADS) (&sm.hist[0][0])[i] = 0;
  for (int i = tid; i <= POS_BINS; i += THREADS) { sm.seq_len[i] = 0; sm.qual_len[i] = 0; sm.pos_sum[i] = 0; }
  if (tid < LOG2_BINS) sm.seq_log2[tid] = 0;
  if (tid == 0) {
    sm.len_min[0] = sm.len_min[1] = ~0ull;
    sm.len_max[0] = sm.len_max[1] = 0;
    sm.bytes_since_flush = 0;
  }
  u64 my_min[2] = {~0ull, ~0ull}, my_max[2] = {0, 0};  // lane 0 of each warp: line-length extrema

  for (;;) {
    if (tid == 0) {
      sm.tile_id = atomicAdd(&info->tile_counter, 1u);
      sm.last_nl = -1;
      sm.nlong = 0;
    }
    __syncthreads();
    const uint32_t t = sm.tile_id;
    if (t >= ntiles) break;
    const u64 toff = (u64)t * TILE;
    const int lo = (t == 0) ? (int)lo0 : 0;
    const int hi = (int)((end - toff) < (u64)TILE ? (end - toff) : (u64)TILE);

    // ---- load: global -> registers -> shared (every HBM byte is read once) ----
    uint4 v[GPT];
#pragma unroll
    for (int j = 0; j < GPT; j++) {
      int off = (tid + j * THREADS) * 16;
      v[j] = (off < hi) ? ldg_stream(base + toff + off) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int j = 0; j < GPT; j++) {
      int off = (tid + j * THREADS) * 16;
      *reinterpret_cast<uint4*>(&sm.buf[PAD + off]) = v[j];
    }
    if (tid == 0) {  // look-behind byte (decides the fate of a '\r' that ended the previous tile)
      uint32_t lb = (t == 0) ? info->lookbehind : (uint32_t)base[toff - 1];
      sm.buf[PAD + lo - 1] = (uint8_t)lb;
    }

    // ---- K1: newline masks, per-row prefix ----
    uint32_t mask[GPT], ex[GPT];
    int my_last = -1;
#pragma unroll
    for (int j = 0; j < GPT; j++) {
      int off = (tid + j * THREADS) * 16;
      uint32_t m = nl_mask16(v[j]);
      if (off < lo || off + 16 > hi) {  // partial group at a stream edge
        int a = lo - off; a = a < 0 ? 0 : (a > 16 ? 16 : a);
        int b = hi - off; b = b < 0 ? 0 : (b > 16 ? 16 : b);
        m &= ((1u << b) - 1u) & ~((1u << a) - 1u);
      }
      mask[j] = m;
      uint32_t c = __popc(m);
      uint32_t inc = warp_incl_scan(c, lane);
      ex[j] = inc - c;
      if (lane == 31) sm.rowcnt[j * NWARPS + warp] = inc;
      if (m) my_last = off + 31 - __clz(m);
    }
    my_last = __reduce_max_sync(0xffffffffu, my_last);
    if (lane == 0 && my_last >= 0) atomicMax(&sm.last_nl, my_last);
    __syncthreads();

    // ---- warp 0: row prefix, publish aggregate, decoupled look-back, publish inclusive prefix ----
    if (warp == 0) {
      uint32_t c = sm.rowcnt[lane];
      uint32_t inc = warp_incl_scan(c, lane);
      sm.rowbase[lane] = inc - c;
      const uint32_t T = __shfl_sync(0xffffffffu, inc, 31);
      const int last_nl = sm.last_nl;
      const u64 tile_len = (u64)(hi - lo);
      const u64 own_b = T ? (HAS_NL | (u64)(hi - (last_nl + 1))) : tile_len;
      if (lane == 0) {
        st_relaxed(&ts[t + 1].a, ST_AGG | (u64)T);
        st_relaxed(&ts[t + 1].b, ST_AGG | own_b);
      }
      u64 cnt_acc = 0, open_acc = 0;
      bool open_done = false;
      long long j = (long long)t;  // slot of the nearest predecessor (slot 0 = carry, always a prefix)
      for (;;) {
        long long idx = j - lane;
        bool valid = idx >= 0;
        u64 a = ST_PFX, b = ST_PFX;
        if (valid) {
          for (;;) {
            a = ld_relaxed(&ts[idx].a);
            b = ld_relaxed(&ts[idx].b);
            if ((a & ST_MASK) != 0 && (a & ST_MASK) == (b & ST_MASK)) break;
          }
        }
        bool is_pfx = valid && (a & ST_MASK) == ST_PFX;
        uint32_t pm = __ballot_sync(0xffffffffu, is_pfx);
        int cut = pm ? (__ffs(pm) - 1) : 31;  // lanes 0..cut take part (lane = distance back)
        bool in = valid && lane <= cut;
        cnt_acc += warp_sum64(in ? (a & VAL_MASK) : 0ull);
        if (!open_done) {
          bool stop = in && (is_pfx || (b & HAS_NL));
          uint32_t sm_ = __ballot_sync(0xffffffffu, stop);
          int first = sm_ ? (__ffs(sm_) - 1) : 32;
          open_acc += warp_sum64((in && lane <= first) ? (b & VAL_MASK) : 0ull);
          if (sm_) open_done = true;
        }
        if (pm) break;
        j -= 32;
      }
      const u64 L0 = cnt_acc, P0 = open_acc;
      const u64 incl_open = T ? (u64)(hi - (last_nl + 1)) : P0 + tile_len;
      if (lane == 0) {
        st_relaxed(&ts[t + 1].a, ST_PFX | (L0 + T));
        st_relaxed(&ts[t + 1].b, ST_PFX | incl_open);
        sm.L0 = L0; sm.P0 = P0; sm.T = T;
        sm.prev_nl = lo - 1;
        sm.bytes_since_flush += (uint32_t)tile_len;
        if (t == ntiles - 1) {  // the stream carry for the next launch / finish()
          carry->lines = L0 + T;
          carry->open_len = incl_open;
          carry->bytes += end - lo0;
          carry->last_byte = sm.buf[PAD + hi - 1];
        }
      }
    }
    __syncthreads();
    const u64 L0 = sm.L0, P0 = sm.P0;
    const int T = (int)sm.T;

    // ---- K2: per-line statistics, in windows of NL_CAP newlines ----
    const int nwin = T / NL_CAP + 1;
    for (int w = 0; w < nwin; w++) {
      const int wb = w * NL_CAP;
#pragma unroll
      for (int j = 0; j < GPT; j++) {
        uint32_t m = mask[j];
        if (m) {
          int off = (tid + j * THREADS) * 16;
          int idx = (int)(sm.rowbase[j * NWARPS + warp] + ex[j]) - wb;
          while (m) {
            int k = __ffs(m) - 1;
            m &= m - 1;
            if (idx >= 0 && idx < NL_CAP) sm.nl[idx] = (uint16_t)(off + k);
            else if (idx == -1) sm.prev_nl = off + k;
            idx++;
          }
        }
      }
      __syncthreads();
      const int seg_end = (wb + NL_CAP < T + 1) ? wb + NL_CAP : T + 1;
      for (int i = wb + warp; i < seg_end; i += NWARPS) {
        const int cls = (int)((L0 + (u64)i) & 3);
        if (!(cls & 1)) continue;  // header / '+' lines carry no statistics
        const int rel = i - wb;
        const int s = ((rel == 0) ? sm.prev_nl : (int)sm.nl[rel - 1]) + 1;
        const bool term = i < T;
        const int e = term ? (int)sm.nl[rel] : hi;
        const u64 pos0 = (i == 0) ? P0 : 0ull;
        const u64 raw_len = pos0 + (u64)(e - s);
        const int cr = (term && raw_len > 0 && sm.buf[PAD + e - 1] == '\r') ? 1 : 0;
        int vs = s, ve = e - cr;
        u64 vpos = pos0;
        if (i == 0 && pos0 > 0 && sm.buf[PAD + lo - 1] == '\r') { vs = lo - 1; vpos = pos0 - 1; }
        if (!term && e > vs && sm.buf[PAD + e - 1] == '\r') ve = e - 1;  // fate decided by the next tile
        if (ve - vs > LONG_SEG) {
          if (lane == 0) {
            uint32_t q = atomicAdd(&sm.nlong, 1u);
            sm.longs[q].vs = vs; sm.longs[q].ve = ve; sm.longs[q].cls = cls; sm.longs[q].vpos = vpos;
          }
        } else {
          for (int o = vs + lane; o < ve; o += 32) account_byte(sm, cls, sm.buf[PAD + o], vpos + (u64)(o - vs), lane);
        }
        if (term && lane == 0) {
          const u64 len = raw_len - (u64)cr;
          const int q = cls == 3;
          const uint32_t bin = len < (u64)POS_BINS ? (uint32_t)len : (uint32_t)POS_BINS;
          if (q) atomicAdd(&sm.qual_len[bin], 1u);
          else { atomicAdd(&sm.seq_len[bin], 1u); atomicAdd(&sm.seq_log2[log2_bin(len)], 1u); }
          if (len < my_min[q]) my_min[q] = len;
          if (len > my_max[q]) my_max[q] = len;
        }
      }
      __syncthreads();
      const int nlong = (int)sm.nlong;
      for (int q = 0; q < nlong; q++) {
        const LongSeg ls = sm.longs[q];
        for (int o = ls.vs + tid; o < ls.ve; o += THREADS)
          account_byte(sm, ls.cls, sm.buf[PAD + o], ls.vpos + (u64)(o - ls.vs), lane);
      }
      if (nlong) {
        __syncthreads();
        if (tid == 0) sm.nlong = 0;
      }
    }
    if (sm.bytes_since_flush > (1u << 24)) {  // keep the 32-bit per-position sums from overflowing
      __syncthreads();
      flush_pos_sum(sm, block, tid);
      if (tid == 0) sm.bytes_since_flush = 0;
    }
  }

  // ---- flush this CTA's counters into its partial block (K3 folds the blocks) ----
  if (lane == 0) {
    for (int q = 0; q < 2; q++) {
      if (my_min[q] != ~0ull) atomicMin(&sm.len_min[q], my_min[q]);
      if (my_max[q] != 0) atomicMax(&sm.len_max[q], my_max[q]);
    }
  }
  __syncthreads();
  for (int bin = tid; bin < 512; bin += THREADS) {  // fold the 32 lane copies (rotated: no bank conflicts)
    const uint32_t* h = &sm.hist[bin >> 8][(bin & 255) << 5];
    u64 s = 0;
#pragma unroll 8
    for (int l = 0; l < 32; l++) s += h[(l + bin) & 31];
    if (s) block[OFF_HIST_SEQ + bin] += s;
  }
  for (int i = tid; i <= POS_BINS; i += THREADS) {
    if (sm.seq_len[i]) block[OFF_SEQ_LEN + i] += sm.seq_len[i];
    if (sm.qual_len[i]) block[OFF_QUAL_LEN + i] += sm.qual_len[i];
  }
  if (tid < LOG2_BINS && sm.seq_log2[tid]) block[OFF_SEQ_LOG2 + tid] += sm.seq_log2[tid];
  flush_pos_sum(sm, block, tid);
  if (tid == 0) {
    if (sm.len_min[0] < block[OFF_SEQ_LEN_MIN]) block[OFF_SEQ_LEN_MIN] = sm.len_min[0];
    if (sm.len_max[0] > block[OFF_SEQ_LEN_MAX]) block[OFF_SEQ_LEN_MAX] = sm.len_max[0];
    if (sm.len_min[1] < block[OFF_QUAL_LEN_MIN]) block[OFF_QUAL_LEN_MIN] = sm.len_min[1];
    if (sm.len_max[1] > block[OFF_QUAL_LEN_MAX]) block[OFF_QUAL_LEN_MAX] = sm.len_max[1];
  }
}

// Prepares a launch: slot 0 of the tile descriptors is the stream carry (an inclusive prefix).
__global__ void fq_begin_launch_kernel(TileState* ts, LaunchInfo* info, const Carry* carry) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    ts[0].a = ST_PFX | carry->lines;
    ts[0].b = ST_PFX | carry->open_len;
    info->tile_counter = 0;
    info->lookbehind = carry->bytes ? carry->last_byte : 0u;
  }
}

// Resets the per-CTA partial blocks and the stream carry (a new file).
__global__ void fq_reset_kernel(u64* partials, int nblocks, Carry* carry) {
  const size_t n = (size_t)nblocks * BLOCK_WORDS;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int w = (int)(i % BLOCK_WORDS);
    partials[i] = (w == OFF_SEQ_LEN_MIN || w == OFF_QUAL_LEN_MIN) ? ~0ull : 0ull;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    Carry c;
    c.lines = 0; c.open_len = 0; c.bytes = 0; c.last_byte = 0; c.flags = 0;
    c.meta_lines = 0; c.qual_min = -1; c.qual_max = -1; c.meta_status = 0; c.meta_pending_cr = 0;
    c.cur_has = 0; c.cur_min = 0; c.cur_max = 0; c.pad = 0;
    *carry = c;
  }
}

// K3: fold the per-CTA partial blocks into one block (sum words, then the four min/max words).
__global__ void fq_reduce_kernel(const u64* __restrict__ partials, int nblocks, u64* __restrict__ out) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= BLOCK_WORDS) return;
  const bool is_min = (w == OFF_SEQ_LEN_MIN || w == OFF_QUAL_LEN_MIN);
  const bool is_max = (w == OFF_SEQ_LEN_MAX || w == OFF_QUAL_LEN_MAX);
  u64 acc = is_min ? ~0ull : 0ull;
  for (int b = 0; b < nblocks; b++) {
    const u64 x = partials[(size_t)b * BLOCK_WORDS + w];
    if (is_min) acc = x < acc ? x : acc;
    else if (is_max) acc = x > acc ? x : acc;
    else acc += x;
  }
  out[w] = acc;
}

// fq-meta quality-range fold over the first 4*meta_records lines (src/fq_meta.nim:226-248): one
// warp walks the stream prefix 32 bytes at a time; qual_to_int (src/fq_meta.nim:94-95) per lane,
// per-line min/max by warp reductions, the prev_min >= 0 rule (src/fq_meta.nim:100-102) per line.
__device__ __forceinline__ void meta_fold(long long& qmin, long long& qmax, unsigned& status, int has, int mn, int mx) {
  if (has) {
    long long a = mn, b = mx;
    if (qmin >= 0) { a = a < qmin ? a : qmin; b = b > qmax ? b : qmax; }
    qmin = a; qmax = b;
  } else if (qmin < 0) {
    status = FQGPU_META_EMPTY_QUAL;
  }
}

__global__ void fq_meta_kernel(const uint8_t* __restrict__ p, u64 n, Carry* __restrict__ carry, u64 meta_records) {
  const int lane = threadIdx.x;
  const u64 limit = meta_records * 4;
  u64 ml = carry->meta_lines;
  if (ml >= limit || n == 0) return;
  long long qmin = carry->qual_min, qmax = carry->qual_max;
  unsigned status = carry->meta_status;
  int cur_has = carry->cur_has, cur_min = carry->cur_min, cur_max = carry->cur_max;
  if (carry->meta_pending_cr && p[0] != '\n') {  // the '\r' that ended the previous chunk was content
    if (!cur_has) { cur_has = 1; cur_min = -1; cur_max = -1; } else { cur_min = -1; }
  }
  unsigned pending = 0;
  bool done = false;
  for (u64 o = 0; o < n && !done; o += 32) {
    const u64 idx = o + lane;
    const bool valid = idx < n;
    const uint32_t b = valid ? p[idx] : 0u;
    const uint32_t nxt = (idx + 1 < n) ? p[idx + 1] : 0x100u;
    const bool is_nl = valid && b == '\n';
    const bool pend = valid && b == '\r' && idx + 1 == n;
    const bool contributes = valid && !is_nl && !(b == '\r' && nxt == '\n') && !pend;
    const int q = (b >= 33 && b <= 126) ? (int)b - 33 : -1;
    uint32_t nlmask = __ballot_sync(0xffffffffu, is_nl);
    if (__ballot_sync(0xffffffffu, pend)) pending = 1;
    int start = 0;
    for (;;) {
      const int endl = nlmask ? (__ffs(nlmask) - 1) : 32;
      const bool inseg = contributes && lane >= start && lane < endl;
      if (__ballot_sync(0xffffffffu, inseg)) {
        const int mn = __reduce_min_sync(0xffffffffu, inseg ? q : 0x7fffffff);
        const int mx = __reduce_max_sync(0xffffffffu, inseg ? q : -0x7fffffff);
        if (!cur_has) { cur_has = 1; cur_min = mn; cur_max = mx; }
        else { cur_min = mn < cur_min ? mn : cur_min; cur_max = mx > cur_max ? mx : cur_max; }
      }
      if (endl == 32) break;
      if ((ml & 3) == 3 && status == FQGPU_META_OK) meta_fold(qmin, qmax, status, cur_has, cur_min, cur_max);
      ml++;
      cur_has = 0;
      if (ml >= limit) { done = true; break; }
      nlmask &= nlmask - 1;
      start = endl + 1;
    }
  }
  if (lane == 0) {
    carry->meta_lines = ml;
    carry->qual_min = qmin; carry->qual_max = qmax;
    carry->meta_status = status;
    carry->meta_pending_cr = done ? 0u : pending;
    carry->cur_has = cur_has; carry->cur_min = cur_min; carry->cur_max = cur_max;
  }
}

// ------------------------------------------------------------------------------------------
// host-callable launchers (used by fqgpu_api.cu)
// ------------------------------------------------------------------------------------------
size_t scan_smem_bytes() { return sizeof(Smem); }
int scan_tile_bytes() { return TILE; }
int scan_threads() { return THREADS; }

cudaError_t scan_configure() {
  return cudaFuncSetAttribute(fq_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
}

cudaError_t launch_reset(u64* partials, int nblocks, Carry* carry, cudaStream_t st) {
  fq_reset_kernel<<<64, 256, 0, st>>>(partials, nblocks, carry);
  return cudaGetLastError();
}

// Scans `nbytes` at `ptr` (any alignment) as the continuation of the stream described by `carry`.
cudaError_t launch_scan(const void* ptr, size_t nbytes, TileState* ts, LaunchInfo* info, Carry* carry,
                        u64* partials, int grid, u64 meta_records, cudaStream_t st) {
  if (nbytes == 0) return cudaSuccess;
  const uintptr_t addr = (uintptr_t)ptr;
  const uint32_t lo0 = (uint32_t)(addr & 15);
  const uint8_t* base = (const uint8_t*)(addr - lo0);
  const u64 end = (u64)lo0 + nbytes;
  const uint32_t ntiles = (uint32_t)((end + TILE - 1) / TILE);
  cudaError_t e = cudaMemsetAsync(ts + 1, 0, (size_t)ntiles * sizeof(TileState), st);
  if (e != cudaSuccess) return e;
  if (meta_records) fq_meta_kernel<<<1, 32, 0, st>>>((const uint8_t*)ptr, nbytes, carry, meta_records);
  fq_begin_launch_kernel<<<1, 32, 0, st>>>(ts, info, carry);
  fq_scan_kernel<<<grid, THREADS, sizeof(Smem), st>>>(base, lo0, end, ntiles, ts, info, carry, partials);
  return cudaGetLastError();
}

cudaError_t launch_reduce(const u64* partials, int nblocks, u64* out, cudaStream_t st) {
  fq_reduce_kernel<<<(BLOCK_WORDS + 255) / 256, 256, 0, st>>>(partials, nblocks, out);
  return cudaGetLastError();
}

}  // namespace fq
