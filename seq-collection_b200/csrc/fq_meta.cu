// fq_meta.cu -- the fq-meta quality-range fold (src/fq_meta.nim:94-102,226-248) over the first 4*meta_records lines
// of the stream.  Runs on its own stream beside the scan and touches only the meta_* / cur_* fields of fq::Carry.
#include <cuda_runtime.h>
#include <stdint.h>

#include "fq_dev.cuh"

namespace fq {

// fq-meta quality-range fold over the first 4*meta_records lines (src/fq_meta.nim:226-248): qual_to_int
// (src/fq_meta.nim:94-95) per byte, per-line min/max, the prev_min >= 0 rule (src/fq_meta.nim:100-102) per line in order.
__device__ __forceinline__ void meta_fold(long long& qmin, long long& qmax, unsigned& status, int has, int mn, int mx) {
  if (has) {
    long long a = mn, b = mx;
    if (qmin >= 0) { a = a < qmin ? a : qmin; b = b > qmax ? b : qmax; }
    qmin = a; qmax = b;
  } else if (qmin < 0) {
    status = FQGPU_META_EMPTY_QUAL;
  }
}

// One CTA of 256 threads walks the stream prefix in windows of 4 KiB (long-read prefixes are megabytes, `-n <all reads>`
// gigabytes): every thread takes 16 bytes, a block-wide prefix sum of the newline counts gives the line index of every
// byte, the per-line (min, max) of the window are shared-memory atomics on keys (0 = byte outside the table, b - 32
// inside, so that key - 1 = qual_to_int), and after every window thread 0 folds the window's finished lines in order
// (the prev_min >= 0 rule needs the order) while the open line's entry moves to the front.
constexpr int META_THREADS = 256;
constexpr int META_CAP = META_THREADS * 16;   // lines a window can end
constexpr uint32_t META_NONE = 0xFFFFFFFFu;

__global__ void __launch_bounds__(META_THREADS) fq_meta_par_kernel(const uint8_t* __restrict__ base, uint32_t first, const u64* __restrict__ start, u64 end,
                                                                 Carry* __restrict__ carry, u64 meta_records) {
  // where the segment fold (below) stopped: the launch's first byte, or a segment boundary inside it
  const u64 lo0 = start ? *start : (u64)first;
  __shared__ uint32_t kmin[META_CAP + 1], kmax[META_CAP + 1];
  __shared__ uint32_t warp_cnt[META_THREADS / 32];
  __shared__ uint32_t s_pending;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const u64 limit = meta_records * 4;
  const u64 ml0 = carry->meta_lines;
  if (ml0 >= limit || end <= lo0) return;
  for (int i = tid; i <= META_CAP; i += META_THREADS) { kmin[i] = META_NONE; kmax[i] = 0; }
  if (tid == 0) s_pending = 0;
  __syncthreads();
  long long qmin = 0, qmax = 0;  // (thread 0) the running fold
  unsigned status = 0;
  if (tid == 0) {  // the open line carried in from the previous chunk
    qmin = carry->qual_min; qmax = carry->qual_max; status = carry->meta_status;
    if (carry->cur_has) { kmin[0] = (uint32_t)(carry->cur_min + 1); kmax[0] = (uint32_t)(carry->cur_max + 1); }
    if (carry->meta_pending_cr && base[lo0] != '\n') kmin[0] = 0;  // the '\r' that ended the previous chunk was content
  }
  __syncthreads();
  u64 L0 = ml0;  // lines before the current window
  for (u64 o = lo0 & ~((u64)META_THREADS * 16 - 1); o < end && L0 < limit; o += (u64)META_THREADS * 16) {
    const u64 g = o + (u64)tid * 16;
    uint4 v = make_uint4(0, 0, 0, 0);
    int va = 16, vb = 0;  // valid bytes of this thread: [va, vb)
    if (g < end && g + 16 > lo0) {
      v = *reinterpret_cast<const uint4*>(base + g);
      va = g >= lo0 ? 0 : (int)(lo0 - g);
      vb = g + 16 <= end ? 16 : (int)(end - g);
    }
    const uint32_t m = va < vb ? (nl_mask16(v) & ((1u << vb) - 1u) & ~((1u << va) - 1u)) : 0u;
    const uint32_t cnt = __popc(m);
    const uint32_t inc = warp_incl_scan(cnt, lane);
    if (lane == 31) warp_cnt[warp] = inc;
    // the byte after this thread's 16 (the '\r' rule): the next lane's first byte, or memory; 0x100 = end of the chunk
    uint32_t nxt = __shfl_down_sync(0xffffffffu, v.x & 0xFFu, 1);
    if (lane == 31) nxt = (g + 16 < end) ? (uint32_t)base[g + 16] : 0x100u;
    else if (g + 16 >= end) nxt = 0x100u;
    __syncthreads();
    uint32_t wbase = 0, total = 0;
    for (int w = 0; w < META_THREADS / 32; w++) { const uint32_t x = warp_cnt[w]; if (w < warp) wbase += x; total += x; }
    uint32_t rel = wbase + inc - cnt;  // window-relative line index of this thread's first byte
    uint32_t mn = META_NONE, mx = 0;
    const uint32_t ww[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 16; k++) {
      if (k >= va && k < vb) {
        const uint32_t c = (ww[k >> 2] >> (8 * (k & 3))) & 0xFFu;
        const u64 L = L0 + rel;
        if ((m >> k) & 1u) {
          if (mn != META_NONE && L < limit) { atomicMin(&kmin[rel], mn); atomicMax(&kmax[rel], mx); }
          mn = META_NONE; mx = 0;
          rel++;
        } else if ((L & 3) == 3 && L < limit) {
          bool content = true;
          if (c == '\r') {
            const uint32_t nx = (k + 1 < vb) ? ((ww[(k + 1) >> 2] >> (8 * ((k + 1) & 3))) & 0xFFu) : (k + 1 < 16 ? 0x100u : nxt);
            if (nx == 0x100u) { s_pending = 1; content = false; }  // last byte of the chunk: decided by the next one
            else if (nx == '\n') content = false;                   // dropped: directly before the newline
          }
          if (content) {
            const uint32_t key = (c >= 33u && c <= 126u) ? c - 32u : 0u;
            mn = min(mn, key); mx = max(mx, key);
          }
        }
      }
    }
    if (mn != META_NONE && L0 + rel < limit) { atomicMin(&kmin[rel], mn); atomicMax(&kmax[rel], mx); }
    __syncthreads();
    // fold the lines this window ended (in order), keep the open line's entry
    uint32_t open_min = META_NONE, open_max = 0;
    if (tid == 0) {
      const u64 nfold = L0 + total < limit ? (u64)total : limit - L0;
      for (u64 r = 0; r < nfold; r++) {
        if (((L0 + r) & 3) != 3 || status != FQGPU_META_OK) continue;
        const uint32_t a = kmin[r];
        meta_fold(qmin, qmax, status, a != META_NONE, (int)a - 1, (int)kmax[r] - 1);
      }
      open_min = kmin[total]; open_max = kmax[total];
    }
    __syncthreads();
    for (uint32_t i = (uint32_t)tid; i <= total; i += META_THREADS) { kmin[i] = META_NONE; kmax[i] = 0; }
    __syncthreads();
    if (tid == 0) { kmin[0] = open_min; kmax[0] = open_max; }
    L0 += total;
    __syncthreads();  // warp_cnt is rewritten by the next window
  }
  if (tid == 0) {
    const u64 ml_end = L0 < limit ? L0 : limit;
    const bool done = ml_end >= limit;
    carry->meta_lines = ml_end;
    carry->qual_min = qmin; carry->qual_max = qmax;
    carry->meta_status = status;
    carry->meta_pending_cr = done ? 0u : s_pending;
    const uint32_t a = done ? META_NONE : kmin[0];
    carry->cur_has = a != META_NONE;
    carry->cur_min = a != META_NONE ? (int)a - 1 : 0;
    carry->cur_max = a != META_NONE ? (int)kmax[0] - 1 : 0;
  }
}

// ---- the same fold, many CTAs wide -------------------------------------------------------------------------------------
// The fold above walks the stream on ONE CTA because src/fq_meta.nim:100-102 is sequential in the line order (a line whose
// minimum is negative makes the next line REPLACE the running range).  With `-n <all reads>` on a large file that walk is
// what the whole command waits for (0.9 GB/s).  But when every sampled quality line is non-empty and has only bytes of
// the table -- every well-formed file -- the fold is a plain min/max over the quality lines, and a quality line is
// "line number = 3 mod 4": so the launch is cut into segments of 64 KiB, one CTA per segment keeps min / max / flags
// for each of the four residues of the line number RELATIVE to its own first line, and one warp then walks the
// segments in order -- it knows the true number of every segment's first line -- and takes the residue that is the
// quality line.  It stops at the first segment it cannot vouch for (an empty quality line, a byte outside the table, a
// lone '\r', the segment in which the sample ends) and hands the rest of the launch to the sequential kernel, which starts
// there with the carry exactly as the sequential walk would have left it.
constexpr uint32_t SEG_BYTES = 65536;
constexpr int SEG_THREADS = 256;
constexpr uint32_t SEG_PER_THREAD = SEG_BYTES / SEG_THREADS;  // 256 consecutive bytes per thread
struct MetaSeg {
  uint32_t lines;                 // '\n' in the segment
  uint8_t mn[4], mx[4];           // per residue of (line number - number of the segment's first line): min / max key (byte - 32)
  uint8_t has, invalid, empty;    // bit r: the residue has content / a byte outside the table (or a lone '\r') / an empty line
  uint8_t open_has, open_mn, open_mx;  // the content behind the segment's last '\n' (of the whole segment if it has none)
  uint8_t ends_cr;                // the segment's last byte is a '\r' at the very end of the launch (not attributed yet)
  uint8_t pad[5];
};
static_assert(sizeof(MetaSeg) == 24, "MetaSeg layout");

struct LineAcc {  // min / max of one line's content bytes so far, in two lanes of 16 bits (VIMNMX3.U16x2 takes two words per step)
  uint32_t mn2, mx2, has;
  __device__ __forceinline__ void clear() { mn2 = 0x00FF00FFu; mx2 = 0; has = 0; }
  // the bytes of the 16-byte group v whose bit is set in `mask`
  __device__ __forceinline__ void add_masked(const uint4& v, uint32_t mask) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const uint32_t bm = ((((mask >> (4 * j)) & 15u) * 0x00204081u) & 0x01010101u) * 0xFFu;  // four bits -> four bytes
      const uint32_t lo = w[j] | ~bm, hi = w[j] & bm;                                          // left out: 0xFF for the min, 0 for the max
      mn2 = __vminu2(mn2, __vminu2(lo & 0x00FF00FFu, (lo >> 8) & 0x00FF00FFu));
      mx2 = __vmaxu2(mx2, __vmaxu2(hi & 0x00FF00FFu, (hi >> 8) & 0x00FF00FFu));
    }
    has |= mask != 0u;
  }
  __device__ __forceinline__ uint32_t min_byte() const { return min(mn2 & 0xFFFFu, mn2 >> 16); }
  __device__ __forceinline__ uint32_t max_byte() const { return max(mx2 & 0xFFFFu, mx2 >> 16); }
  // a byte outside '!'..'~' (which is what makes the fold give up on the line); as keys: byte - 32, 0 = outside the table
  __device__ __forceinline__ uint32_t invalid() const { return min_byte() < 33u || max_byte() > 126u; }
  __device__ __forceinline__ uint32_t min_key() const { return invalid() ? 0u : min_byte() - 32u; }
  __device__ __forceinline__ uint32_t max_key() const { const uint32_t m = max_byte(); return m >= 33u && m <= 126u ? m - 32u : 0u; }
};

__global__ void __launch_bounds__(SEG_THREADS) fq_meta_seg_kernel(const uint8_t* __restrict__ base, uint32_t lo0, u64 end,
                                                                 const Carry* __restrict__ carry, u64 meta_records, uint32_t seg_first,
                                                                 const u64* __restrict__ start, MetaSeg* __restrict__ segs) {
  // (a slab of segments; the slabs before it were joined: nothing to do if the join stopped, or the sample is complete)
  if (carry->meta_lines >= meta_records * 4 || (seg_first && *start != (u64)seg_first * SEG_BYTES)) return;
  // the segment in shared memory, one row of 256 bytes per thread (+ 16: rows start in different banks)
  constexpr uint32_t ROW = SEG_PER_THREAD + 16;
  extern __shared__ __align__(16) uint8_t seg_s[];
  __shared__ uint32_t warp_cnt[SEG_THREADS / 32];
  __shared__ uint32_t t_cnt[SEG_THREADS], t_tail[SEG_THREADS];  // per thread: newlines; has << 16 | min key << 8 | max key of the content behind its last one
  __shared__ uint32_t red[SEG_THREADS / 32][12];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const u64 seg0 = (u64)(seg_first + blockIdx.x) * SEG_BYTES;
#pragma unroll 4
  for (uint32_t j = 0; j < SEG_BYTES / 16 / SEG_THREADS; j++) {  // coalesced in, row-wise out
    const uint32_t idx = j * SEG_THREADS + (uint32_t)tid;
    const u64 g = seg0 + (u64)idx * 16;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (g < end && g + 16 > (u64)lo0) v = *reinterpret_cast<const uint4*>(base + g);
    *reinterpret_cast<uint4*>(seg_s + (idx >> 4) * ROW + (idx & 15u) * 16u) = v;
  }
  __syncthreads();
  const uint8_t* row = seg_s + (uint32_t)tid * ROW;
  const u64 g0 = seg0 + (u64)tid * SEG_PER_THREAD;
  u64 a = g0, b = g0 + SEG_PER_THREAD;  // this thread's bytes [a, b)
  if (a < lo0) a = lo0;
  if (b > end) b = end;
  // the walk, with line numbers relative to the THREAD's first line (what that line's number is in the segment comes out
  // of the prefix sum below): min / max / flags per residue, and the line open at the thread's end
  uint32_t mn[4] = {255, 255, 255, 255}, mx[4] = {0, 0, 0, 0}, has = 0, inv = 0, empty = 0, ends_cr = 0, rel = 0;
  LineAcc cur;
  cur.clear();
  auto close_line = [&](uint32_t r) {  // the content gathered so far belongs to residue r
    if (cur.has) {
      const uint32_t lo = cur.min_key(), hi = cur.max_key(), bad = cur.invalid();
#pragma unroll
      for (uint32_t q = 0; q < 4; q++)
        if (q == r) { mn[q] = min(mn[q], lo); mx[q] = max(mx[q], hi); has |= 1u << q; inv |= bad << q; }
    }
    cur.clear();
  };
  if (a < b) {
    // Every group of 16 bytes goes through the same steps, whatever it holds (a warp's lanes meet their newlines in
    // different rounds: a branch per case ran with 6 of 32 lanes active, measured): masks of the newlines and of the
    // '\r' that are dropped, the content before each newline to the line that is open, the rest to the next one.
    // What the two bytes before a newline were decides whether the line was empty; for the group's first bytes those
    // are the last bytes of the group before, kept as bits: '\n'?, '\r'?, before the launch's first byte?
    auto before = [&](uint32_t d, uint32_t& is_nl, uint32_t& is_cr, uint32_t& is_st) {  // the byte d before the thread's range
      is_nl = is_cr = is_st = 0;
      if (a < (u64)lo0 + d) { is_st = 1; return; }
      const u64 pos = a - d;
      const uint32_t c = pos >= seg0 ? seg_s[(uint32_t)((pos - seg0) >> 8) * ROW + (uint32_t)((pos - seg0) & 255u)] : base[pos];
      is_nl = c == '\n'; is_cr = c == '\r';
    };
    uint32_t pnl1, pnl2, pcr1, pcr2, pst1, pst2;
    before(1, pnl1, pcr1, pst1);
    before(2, pnl2, pcr2, pst2);
    const bool carried_content = carry->cur_has != 0;  // (the line open at the launch's first byte)
    for (uint32_t q = 0; q < SEG_PER_THREAD / 16; q++) {
      const u64 g = g0 + q * 16;
      if (g >= b || g + 16 <= a) continue;
      const uint4 v = *reinterpret_cast<const uint4*>(row + q * 16);
      // the byte after the group: the row's next, the next thread's first, or memory (the segment's last group)
      uint32_t after = 0x100u;
      if (g + 16 < end) after = q + 1 < SEG_PER_THREAD / 16 ? row[q * 16 + 16] : (tid + 1 < SEG_THREADS ? row[ROW] : (uint32_t)base[g + 16]);
      const uint32_t ka = g >= a ? 0u : (uint32_t)(a - g), kb = g + 16 <= b ? 16u : (uint32_t)(b - g);
      const uint32_t kend = end - g > 16 ? 17u : (uint32_t)(end - g);  // bytes of the launch from g on (capped)
      const uint32_t inrange = ((1u << kb) - 1u) & ~((1u << ka) - 1u);
      const uint32_t st_all = g < (u64)lo0 ? (1u << (uint32_t)((u64)lo0 - g)) - 1u : 0u;       // bytes before the launch's first
      const uint32_t data = (kend >= 16 ? 0xFFFFu : (1u << kend) - 1u) & ~st_all;             // bytes of the launch
      const uint32_t nl_all = nl_mask16(v) & data, cr_all = cr_mask16(v) & data;
      // a '\r' directly before a newline is dropped; a '\r' that is the launch's last byte is left to the next launch
      uint32_t drop = cr_all & ((nl_all >> 1) | (after == '\n' ? 0x8000u : 0u));
      const uint32_t last_cr = kend <= 16 ? cr_all & inrange & (1u << (kend - 1)) : 0u;
      if (last_cr) ends_cr = 1;
      drop |= last_cr;
      uint32_t nls = nl_all & inrange, rem = inrange & ~nl_all & ~drop;
      // bit k: the byte before byte k is a newline, or a '\r' with a newline before it (positions shifted by 2 to take
      // the two carried bytes in); the same for "before the launch's first byte"
      const uint32_t e_nl = (nl_all << 2) | (pnl1 << 1) | pnl2, e_cr = (cr_all << 2) | (pcr1 << 1) | pcr2, e_st = (st_all << 2) | (pst1 << 1) | pst2;
      const uint32_t bare = ((e_nl << 1) | ((e_cr << 1) & (e_nl << 2))) >> 2, atst = ((e_st << 1) | ((e_cr << 1) & (e_st << 2))) >> 2;
      const uint32_t empties = nls & (bare | (carried_content ? 0u : atst));
      while (nls) {
        const uint32_t pos = (uint32_t)__ffs((int)nls) - 1u, below = (1u << pos) - 1u;
        cur.add_masked(v, rem & below);
        if ((empties >> pos) & 1u) empty |= 1u << (rel & 3u);
        close_line(rel & 3u);
        rel++;
        rem &= ~below;
        nls &= nls - 1u;
      }
      cur.add_masked(v, rem);
      pnl1 = (nl_all >> 15) & 1u; pnl2 = (nl_all >> 14) & 1u; pcr1 = (cr_all >> 15) & 1u; pcr2 = (cr_all >> 14) & 1u;
      pst1 = (st_all >> 15) & 1u; pst2 = (st_all >> 14) & 1u;
    }
  }
  // the open part at the thread's end also counts for its residue; and it is what the segment's open line is made of
  const uint32_t cnt = rel;
  t_cnt[tid] = cnt;
  t_tail[tid] = cur.has ? (1u << 16) | (cur.min_key() << 8) | cur.max_key() : 0u;
  close_line(rel & 3u);
  // the thread's first line is line `first` of the segment: turn the residues by that much
  const uint32_t inc = warp_incl_scan(cnt, lane);
  if (lane == 31) warp_cnt[warp] = inc;
  __syncthreads();
  uint32_t first = inc - cnt, total = 0;
  for (int w = 0; w < SEG_THREADS / 32; w++) { const uint32_t x = warp_cnt[w]; if (w < warp) first += x; total += x; }
  {
    const uint32_t sh = first & 3u;
    uint32_t rmn[4], rmx[4];
#pragma unroll
    for (uint32_t q = 0; q < 4; q++) {  // residue q of the segment = residue (q - sh) of the thread
      const uint32_t src = (q - sh) & 3u;
      rmn[q] = src == 0 ? mn[0] : (src == 1 ? mn[1] : (src == 2 ? mn[2] : mn[3]));
      rmx[q] = src == 0 ? mx[0] : (src == 1 ? mx[1] : (src == 2 ? mx[2] : mx[3]));
    }
#pragma unroll
    for (uint32_t q = 0; q < 4; q++) { mn[q] = rmn[q]; mx[q] = rmx[q]; }
    auto turn = [sh](uint32_t m) { return ((m << sh) | (m >> (4u - sh))) & 15u; };
    has = turn(has); inv = turn(inv); empty = turn(empty);
  }
  // reduce over the CTA
  uint32_t vals[12] = {mn[0], mn[1], mn[2], mn[3], mx[0], mx[1], mx[2], mx[3], has, inv, empty, ends_cr};
#pragma unroll
  for (int k = 0; k < 12; k++) {
    uint32_t v = vals[k];
    if (k < 4) v = __reduce_min_sync(0xffffffffu, v);
    else if (k < 8) v = __reduce_max_sync(0xffffffffu, v);
    else v = __reduce_or_sync(0xffffffffu, v);
    if (lane == 0) red[warp][k] = v;
  }
  __syncthreads();
  if (warp == 0) {
    // the segment's open line: the tails of the last thread that saw a newline and of every thread behind it
    uint32_t last = 0;  // 1 + the last thread with a newline (0: none)
    for (int t0 = 0; t0 < SEG_THREADS; t0 += 32) {
      const uint32_t m = __ballot_sync(0xffffffffu, t_cnt[t0 + lane] != 0);
      if (m) last = (uint32_t)t0 + 32u - (uint32_t)__clz((int)m);
    }
    uint32_t omn = 255, omx = 0, ohas = 0;
    for (uint32_t t = (last ? last - 1 : 0) + (uint32_t)lane; t < SEG_THREADS; t += 32) {
      const uint32_t x = t_tail[t];
      if (x >> 16) { ohas = 1; omn = min(omn, (x >> 8) & 255u); omx = max(omx, x & 255u); }
    }
    omn = __reduce_min_sync(0xffffffffu, omn); omx = __reduce_max_sync(0xffffffffu, omx); ohas = __reduce_or_sync(0xffffffffu, ohas);
    if (lane == 0) {
      MetaSeg o;
      uint32_t r[12];
      for (int k = 0; k < 12; k++) {
        uint32_t v = red[0][k];
        for (int w = 1; w < SEG_THREADS / 32; w++) v = k < 4 ? min(v, red[w][k]) : (k < 8 ? max(v, red[w][k]) : (v | red[w][k]));
        r[k] = v;
      }
      o.lines = total;
      for (int q = 0; q < 4; q++) { o.mn[q] = (uint8_t)r[q]; o.mx[q] = (uint8_t)r[4 + q]; }
      o.has = (uint8_t)r[8]; o.invalid = (uint8_t)r[9]; o.empty = (uint8_t)r[10]; o.ends_cr = (uint8_t)r[11];
      o.open_has = (uint8_t)ohas; o.open_mn = (uint8_t)omn; o.open_mx = (uint8_t)omx;
      for (int k = 0; k < 5; k++) o.pad[k] = 0;
      segs[seg_first + blockIdx.x] = o;
    }
  }
}

// One CTA walks the segments of a slab in order, 1024 at a time (a prefix sum gives every segment the number of its
// first line, and with it the residue that is the quality line); *start = where the sequential kernel has to take
// over (`end`: nowhere).
constexpr int JOIN_THREADS = 1024;
__device__ __forceinline__ uint32_t block_reduce(uint32_t v, int op, uint32_t* scratch) {  // op 0: min, 1: max, 2: add; every thread gets the result
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = op == 0 ? __reduce_min_sync(0xffffffffu, v) : (op == 1 ? __reduce_max_sync(0xffffffffu, v) : __reduce_add_sync(0xffffffffu, v));
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  uint32_t r = scratch[0];
  for (int w = 1; w < JOIN_THREADS / 32; w++) r = op == 0 ? min(r, scratch[w]) : (op == 1 ? max(r, scratch[w]) : r + scratch[w]);
  return r;
}
__global__ void __launch_bounds__(JOIN_THREADS) fq_meta_join_kernel(const uint8_t* __restrict__ base, uint32_t lo0, u64 end,
                                                                    const MetaSeg* __restrict__ segs, uint32_t seg_first, uint32_t seg_end,
                                                                    uint32_t nseg, Carry* __restrict__ carry, u64 meta_records, u64* __restrict__ start) {
  __shared__ uint32_t scratch[JOIN_THREADS / 32], wsum[JOIN_THREADS / 32];
  __shared__ int go;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const u64 limit = meta_records * 4;
  if (tid == 0) {
    go = 1;
    if (seg_first == 0) *start = lo0;
    else if (*start != (u64)seg_first * SEG_BYTES) go = 0;  // an earlier slab could not be vouched for to its end
    if (carry->meta_lines >= limit || end <= (u64)lo0) go = 0;
    // the '\r' that ended the launch before was content, or the open line already holds a byte outside the table: the sequential walk
    if (seg_first == 0 && carry->meta_pending_cr && base[lo0] != '\n') go = 0;
    if (carry->cur_has && carry->cur_min < 0) go = 0;
  }
  __syncthreads();
  if (!go) return;
  u64 L = carry->meta_lines;                                    // lines before the current batch
  long long qmin = carry->qual_min, qmax = carry->qual_max;
  const unsigned status = carry->meta_status;
  // the line open where the slab begins, as keys
  uint32_t ohas = carry->cur_has != 0, omn = ohas ? (uint32_t)(carry->cur_min + 1) : 255u, omx = ohas ? (uint32_t)(carry->cur_max + 1) : 0u;
  uint32_t fmn = 255, fmx = 0;                                  // what the slab adds to the fold
  if (ohas && (L & 3) == 3) { fmn = omn; fmx = omx; }          // (the open line is a quality line and goes on in the slab)
  uint32_t stop = seg_end;
  bool ends_cr = false;
  for (uint32_t batch = seg_first; batch < seg_end; batch += JOIN_THREADS) {
    const uint32_t s = batch + (uint32_t)tid;
    const bool valid = s < seg_end;
    MetaSeg g;
    g.lines = 0; g.has = 0; g.invalid = 0; g.empty = 0; g.open_has = 0; g.open_mn = 255; g.open_mx = 0; g.ends_cr = 0;
    if (valid) g = segs[s];
    // lines before this segment
    const uint32_t inc = warp_incl_scan(g.lines, lane);
    __syncthreads();
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    u64 Ls = L + inc - g.lines;
    for (int w = 0; w < warp; w++) Ls += wsum[w];
    const uint32_t q = (uint32_t)(3u - (uint32_t)(Ls & 3u)) & 3u;  // the residue that is a quality line
    // the first segment that cannot be vouched for: the sample ends in it, or its quality lines are not plain
    const bool fail = valid && (Ls + g.lines >= limit || (((g.invalid | g.empty) >> q) & 1u));
    const uint32_t first_fail = block_reduce(fail ? (uint32_t)tid : (uint32_t)JOIN_THREADS, 0, scratch);
    const bool in = valid && (uint32_t)tid < first_fail;
    const bool hasq = in && ((g.has >> q) & 1u);
    uint32_t a = 255, b = 0;
#pragma unroll
    for (uint32_t r = 0; r < 4; r++) if (hasq && r == q) { a = g.mn[r]; b = g.mx[r]; }
    fmn = min(fmn, block_reduce(a, 0, scratch));
    fmx = max(fmx, block_reduce(b, 1, scratch));
    // the open line behind the batch: the open part of the last segment with a newline and everything behind it
    const uint32_t last_nl = block_reduce(in && g.lines ? (uint32_t)tid + 1u : 0u, 1, scratch);  // 1 + its index, 0: none
    const bool tail = in && (uint32_t)tid + 1u >= last_nl && g.open_has;
    const uint32_t tmn = block_reduce(tail ? (uint32_t)g.open_mn : 255u, 0, scratch);
    const uint32_t tmx = block_reduce(tail ? (uint32_t)g.open_mx : 0u, 1, scratch);
    const uint32_t tany = block_reduce(tail ? 1u : 0u, 1, scratch);
    if (last_nl) { ohas = tany; omn = tany ? tmn : 255u; omx = tany ? tmx : 0u; }
    else if (tany) { ohas = 1; omn = min(omn, tmn); omx = max(omx, tmx); }
    const uint32_t folded = first_fail < (uint32_t)JOIN_THREADS ? first_fail : (seg_end - batch < (uint32_t)JOIN_THREADS ? seg_end - batch : (uint32_t)JOIN_THREADS);
    if (folded) ends_cr = block_reduce(valid && (uint32_t)tid + 1u == folded ? (uint32_t)g.ends_cr : 0u, 1, scratch) != 0;
    // lines of the segments folded
    {
      const uint32_t lo_lines = block_reduce(in ? g.lines & 0xFFFFu : 0u, 2, scratch);  // (sums of up to 1024 * 65536: in two halves)
      const uint32_t hi_lines = block_reduce(in ? g.lines >> 16 : 0u, 2, scratch);
      L += (u64)lo_lines + ((u64)hi_lines << 16);
    }
    if (first_fail < (uint32_t)JOIN_THREADS) { stop = batch + first_fail; break; }
  }
  if (stop == seg_first) return;  // nothing vouched for: the carry stays as it is
  if (tid == 0) {
    if (fmn != 255u && status == FQGPU_META_OK) {
      long long x = (long long)fmn - 1, y = (long long)fmx - 1;
      if (qmin >= 0) { x = x < qmin ? x : qmin; y = y > qmax ? y : qmax; }
      qmin = x; qmax = y;
    }
    const u64 at = stop < nseg ? (u64)stop * SEG_BYTES : end;
    *start = at;
    carry->meta_lines = L;
    carry->qual_min = qmin; carry->qual_max = qmax;
    carry->meta_pending_cr = at >= end && ends_cr ? 1u : 0u;
    carry->cur_has = (int)ohas;
    carry->cur_min = ohas ? (int)omn - 1 : 0;
    carry->cur_max = ohas ? (int)omx - 1 : 0;
  }
}

// The prefix fold of a launch over bytes [lo0, end) of `base` (16-byte aligned).  `segs`: room for meta_seg_count(end)
// segments; `start`: one word.
// once per device (fqgpu_create)
cudaError_t meta_configure() {
  return cudaFuncSetAttribute(fq_meta_seg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SEG_THREADS * (SEG_PER_THREAD + 16));
}
size_t meta_seg_count(u64 end) { return (size_t)((end + SEG_BYTES - 1) / SEG_BYTES); }
size_t meta_seg_bytes() { return sizeof(MetaSeg); }
cudaError_t launch_meta(const uint8_t* base, uint32_t lo0, u64 end, Carry* carry, u64 meta_records, void* segs, u64* start, cudaStream_t st) {
  const uint32_t nseg = (uint32_t)meta_seg_count(end);
  if (nseg == 0) return cudaSuccess;
  if (meta_records * 4 <= 4096) {  // a small sample (the default is 100 reads): the sequential walk alone
    fq_meta_par_kernel<<<1, META_THREADS, 0, st>>>(base, lo0, nullptr, end, carry, meta_records);
    return cudaGetLastError();
  }
  // slabs of segments, growing: a sample that ends early costs a few small launches, not a pass over the whole input
  for (uint32_t first = 0, slab = 4; first < nseg; first += slab, slab = slab < 16384 ? slab * 4 : slab) {
    const uint32_t n = nseg - first < slab ? nseg - first : slab;
    fq_meta_seg_kernel<<<n, SEG_THREADS, SEG_THREADS * (SEG_PER_THREAD + 16), st>>>(base, lo0, end, carry, meta_records, first, start, (MetaSeg*)segs);
    fq_meta_join_kernel<<<1, JOIN_THREADS, 0, st>>>(base, lo0, end, (const MetaSeg*)segs, first, first + n, nseg, carry, meta_records, start);
  }
  fq_meta_par_kernel<<<1, META_THREADS, 0, st>>>(base, lo0, start, end, carry, meta_records);
  return cudaGetLastError();
}

}  // namespace fq
