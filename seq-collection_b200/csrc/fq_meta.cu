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
  long long qmin = 0, qmax = 0;  // (thread 0) the running fold
  unsigned status = 0;
  if (tid == 0) {  // the open line carried in from the previous chunk
    qmin = carry->qual_min; qmax = carry->qual_max; status = carry->meta_status;
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

// The prefix fold of a launch over bytes [lo0, end) of `base` (16-byte aligned).
cudaError_t launch_meta(const uint8_t* base, uint32_t lo0, u64 end, Carry* carry, u64 meta_records, cudaStream_t st) {
  fq_meta_par_kernel<<<1, META_THREADS, 0, st>>>(base, lo0, end, carry, meta_records);
  return cudaGetLastError();
}

}  // namespace fq
