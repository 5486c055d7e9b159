// fq_meta.cu -- the fq-meta quality-range fold (src/fq_meta.nim:94-102,226-248) over the first 4*meta_records lines
// of the stream.  Runs on its own stream beside the scan and touches only the meta_* / cur_* fields of fq::Carry.
#include <cuda_runtime.h>
#include <stdint.h>

#include "fq_dev.cuh"

namespace fq {

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

// One warp, 512 bytes per iteration (16 per lane): newline masks by SWAR, the segments between
// newlines are reduced with warp min/max; only quality lines (0-based index % 4 == 3) are examined.
__global__ void fq_meta_kernel(const uint8_t* __restrict__ base, uint32_t lo0, u64 end, Carry* __restrict__ carry, u64 meta_records) {
  const int lane = threadIdx.x;
  const u64 limit = meta_records * 4;
  u64 ml = carry->meta_lines;
  if (ml >= limit || end <= (u64)lo0) return;
  long long qmin = carry->qual_min, qmax = carry->qual_max;
  unsigned status = carry->meta_status;
  int cur_has = carry->cur_has, cur_min = carry->cur_min, cur_max = carry->cur_max;
  if (carry->meta_pending_cr && base[lo0] != '\n') {  // the '\r' that ended the previous chunk was content
    if (!cur_has) { cur_has = 1; cur_min = -1; cur_max = -1; } else { cur_min = -1; }
  }
  unsigned pending = 0;
  bool done = false;
  uint4 vnext = make_uint4(0, 0, 0, 0);
  if ((u64)lane * 16 < end) vnext = *reinterpret_cast<const uint4*>(base + (u64)lane * 16);
  for (u64 o = 0; o < end && !done; o += 512) {
    const u64 g = o + (u64)lane * 16;
    const uint4 v = vnext;
    if (g + 512 < end) vnext = *reinterpret_cast<const uint4*>(base + g + 512);  // prefetch the next window
    // valid byte range of this lane: [va, vb) within its 16 bytes
    int va = g >= (u64)lo0 ? 0 : (int)min((u64)16, (u64)lo0 - g);
    int vb = g + 16 <= end ? 16 : (g < end ? (int)(end - g) : 0);
    uint32_t nlm = nl_mask16(v) & ((1u << vb) - 1u) & ~((1u << va) - 1u);
    // byte following this lane's 16 (for the '\r' rule): next lane's first byte, or memory
    uint32_t nxt = __shfl_down_sync(0xffffffffu, v.x & 0xFFu, 1);
    if (lane == 31) nxt = (g + 16 < end) ? (uint32_t)base[g + 16] : 0u;
    uint32_t lanes_nl = __ballot_sync(0xffffffffu, nlm != 0);
    int seg_lane = 0, seg_k = 0;  // current segment starts at (lane, byte) = (seg_lane, seg_k)
    for (;;) {
      // next newline at or after the segment start
      int nl_lane = 32, nl_k = 0;
      if (lanes_nl) {
        nl_lane = __ffs(lanes_nl) - 1;
        const uint32_t m = __shfl_sync(0xffffffffu, nlm, nl_lane);
        nl_k = __ffs(m) - 1;
      }
      if ((ml & 3) == 3) {  // quality line: min/max of qual_to_int over [segment start, newline or window end)
        int a = lane < seg_lane ? 16 : (lane == seg_lane ? seg_k : 0);
        int b = lane > nl_lane ? 0 : (lane == nl_lane ? nl_k : 16);
        a = max(a, va); b = min(b, vb);
        int mn = 0x7fffffff, mx = -0x7fffffff;
#pragma unroll
        for (int k = 0; k < 16; k++) {
          const uint32_t wk = k < 4 ? v.x : (k < 8 ? v.y : (k < 12 ? v.z : v.w));
          const uint32_t c = (wk >> (8 * (k & 3))) & 0xFFu;
          bool use = k >= a && k < b;
          if (c == '\r' && use) {
            if (g + (u64)k + 1 >= end) { pending = 1; use = false; }  // last byte of the chunk: decided later
            else {
              const uint32_t wn = (k + 1) < 4 ? v.x : ((k + 1) < 8 ? v.y : ((k + 1) < 12 ? v.z : v.w));
              const uint32_t nx = k < 15 ? ((wn >> (8 * ((k + 1) & 3))) & 0xFFu) : nxt;
              if (nx == '\n') use = false;                            // dropped: directly before the newline
            }
          }
          if (use) {
            const int q = (c >= 33 && c <= 126) ? (int)c - 33 : -1;
            mn = min(mn, q); mx = max(mx, q);
          }
        }
        mn = __reduce_min_sync(0xffffffffu, mn);
        mx = __reduce_max_sync(0xffffffffu, mx);
        pending = __reduce_max_sync(0xffffffffu, pending);
        if (mn != 0x7fffffff) {
          if (!cur_has) { cur_has = 1; cur_min = mn; cur_max = mx; }
          else { cur_min = min(cur_min, mn); cur_max = max(cur_max, mx); }
        }
      }
      if (nl_lane == 32) break;
      if ((ml & 3) == 3 && status == FQGPU_META_OK) meta_fold(qmin, qmax, status, cur_has, cur_min, cur_max);
      ml++;
      cur_has = 0;
      pending = 0;
      if (ml >= limit) { done = true; break; }
      // consume this newline
      if (lane == nl_lane) nlm &= nlm - 1;
      lanes_nl = __ballot_sync(0xffffffffu, nlm != 0);
      seg_lane = nl_lane; seg_k = nl_k + 1;
      if (seg_k == 16) { seg_lane++; seg_k = 0; }
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

// The same fold by one CTA of 1024 threads, 16 KiB per step (long-read prefixes are megabytes): every thread
// takes 16 bytes, a block-wide prefix sum of the newline counts gives the line index of every byte, the
// per-line (min, max) are shared-memory atomics on keys (0 = byte outside the table, b - 32 inside, so that
// key - 1 = qual_to_int), and thread 0 folds the lines in order at the end.  Holds up to META_CAP lines.
constexpr int META_THREADS = 1024;
constexpr int META_CAP = 4096;
constexpr uint32_t META_NONE = 0xFFFFFFFFu;

__global__ void __launch_bounds__(META_THREADS) fq_meta_par_kernel(const uint8_t* __restrict__ base, uint32_t lo0, u64 end,
                                                                 Carry* __restrict__ carry, u64 meta_records) {
  __shared__ uint32_t kmin[META_CAP + 1], kmax[META_CAP + 1];
  __shared__ uint32_t warp_cnt[META_THREADS / 32];
  __shared__ uint32_t s_pending;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const u64 limit = meta_records * 4;
  const u64 ml0 = carry->meta_lines;
  if (ml0 >= limit || end <= (u64)lo0) return;
  for (int i = tid; i <= META_CAP; i += META_THREADS) { kmin[i] = META_NONE; kmax[i] = 0; }
  if (tid == 0) s_pending = 0;
  __syncthreads();
  if (tid == 0) {  // the open line carried in from the previous chunk
    if (carry->cur_has) { kmin[0] = (uint32_t)(carry->cur_min + 1); kmax[0] = (uint32_t)(carry->cur_max + 1); }
    if (carry->meta_pending_cr && base[lo0] != '\n') kmin[0] = 0;  // the '\r' that ended the previous chunk was content
  }
  __syncthreads();
  u64 L0 = ml0;  // lines before the current window
  for (u64 o = 0; o < end && L0 < limit; o += (u64)META_THREADS * 16) {
    const u64 g = o + (u64)tid * 16;
    uint4 v = make_uint4(0, 0, 0, 0);
    int va = 16, vb = 0;  // valid bytes of this thread: [va, vb)
    if (g < end && g + 16 > (u64)lo0) {
      v = *reinterpret_cast<const uint4*>(base + g);
      va = g >= (u64)lo0 ? 0 : (int)((u64)lo0 - g);
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
    u64 L = L0 + wbase + inc - cnt;  // line index of this thread's first byte
    uint32_t mn = META_NONE, mx = 0;
    const uint32_t ww[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 16; k++) {
      if (k >= va && k < vb) {
        const uint32_t c = (ww[k >> 2] >> (8 * (k & 3))) & 0xFFu;
        if ((m >> k) & 1u) {
          if (mn != META_NONE && L < limit) { atomicMin(&kmin[L - ml0], mn); atomicMax(&kmax[L - ml0], mx); }
          mn = META_NONE; mx = 0;
          L++;
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
    if (mn != META_NONE && L < limit) { atomicMin(&kmin[L - ml0], mn); atomicMax(&kmax[L - ml0], mx); }
    L0 += total;
    __syncthreads();  // warp_cnt is rewritten by the next window
  }
  if (tid == 0) {
    const u64 ml_end = L0 < limit ? L0 : limit;
    long long qmin = carry->qual_min, qmax = carry->qual_max;
    unsigned status = carry->meta_status;
    for (u64 L = ml0; L < ml_end; L++) {
      if ((L & 3) != 3 || status != FQGPU_META_OK) continue;
      const uint32_t a = kmin[L - ml0];
      meta_fold(qmin, qmax, status, a != META_NONE, (int)a - 1, (int)kmax[L - ml0] - 1);
    }
    const bool done = ml_end >= limit;
    carry->meta_lines = ml_end;
    carry->qual_min = qmin; carry->qual_max = qmax;
    carry->meta_status = status;
    carry->meta_pending_cr = done ? 0u : s_pending;
    const uint32_t a = done ? META_NONE : kmin[ml_end - ml0];
    carry->cur_has = a != META_NONE;
    carry->cur_min = a != META_NONE ? (int)a - 1 : 0;
    carry->cur_max = a != META_NONE ? (int)kmax[ml_end - ml0] - 1 : 0;
  }
}

// The prefix fold of a launch over bytes [lo0, end) of `base` (16-byte aligned).
cudaError_t launch_meta(const uint8_t* base, uint32_t lo0, u64 end, Carry* carry, u64 meta_records, cudaStream_t st) {
  if (meta_records * 4 <= (u64)META_CAP) fq_meta_par_kernel<<<1, META_THREADS, 0, st>>>(base, lo0, end, carry, meta_records);
  else fq_meta_kernel<<<1, 32, 0, st>>>(base, lo0, end, carry, meta_records);
  return cudaGetLastError();
}

}  // namespace fq
