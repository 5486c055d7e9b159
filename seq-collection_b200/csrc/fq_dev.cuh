// fq_dev.cuh -- small device helpers shared by the scan and fq-meta kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fq_layout.h"

namespace fq {

typedef unsigned long long u64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar_s, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_s), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar_s, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar_s, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tFQ_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra FQ_DONE;\n\tbra FQ_WAIT;\n\tFQ_DONE:\n\t}" ::"r"(bar_s), "r"(parity) : "memory");
}
// 1-D bulk copy global -> shared (TMA), completion on an mbarrier; 16-byte aligned, size a multiple of 16
__device__ __forceinline__ void tma_load_1d(uint32_t dst_s, const void* src, uint32_t bytes, uint32_t bar_s) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic reads of dst before the async write
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_s), "l"(src), "r"(bytes), "r"(bar_s) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds16(uint32_t addr) {
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds8(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ u64 lds64(uint32_t addr) {
  u64 v;
  asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts16(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((uint16_t)v) : "memory");
}
__device__ __forceinline__ void sts64(uint32_t addr, u64 v) { asm volatile("st.shared.u64 [%0], %1;" ::"r"(addr), "l"(v) : "memory"); }
__device__ __forceinline__ void red_inc(uint32_t addr) { asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr) : "memory"); }
template <int OFF>
__device__ __forceinline__ void red_add_at(uint32_t addr, uint32_t v) {
  asm volatile("red.shared.add.u32 [%0+%2], %1;" ::"r"(addr), "r"(v), "n"(OFF) : "memory");
}
__device__ __forceinline__ u64 ld_relaxed_gpu(const u64* p) {
  u64 v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_gpu(u64* p, u64 v) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ uint32_t lanemask_lt() {
  uint32_t m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

// 0x80 in every byte lane of w that equals '\n' (exact: no carries cross byte lanes)
__device__ __forceinline__ uint32_t nl_flags(uint32_t w) {
  uint32_t x = w ^ 0x0A0A0A0Au;
  return ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u;
}
// the same for words whose bytes are all < 0x80 (one operation less)
__device__ __forceinline__ uint32_t nl_flags_ascii(uint32_t w) { return ~((w ^ 0x0A0A0A0Au) + 0x7F7F7F7Fu) & 0x80808080u; }
// 16-bit mask of the '\n' bytes of a 16-byte group; the movemask is two IDP.4A chains (FMA pipe)
__device__ __forceinline__ uint32_t nl_mask16(const uint4& v) {
  uint32_t lo = __dp4a(nl_flags(v.x), 0x08040201u, __dp4a(nl_flags(v.y), 0x80402010u, 0u));
  uint32_t hi = __dp4a(nl_flags(v.z), 0x08040201u, __dp4a(nl_flags(v.w), 0x80402010u, 0u));
  return (lo >> 7) | (hi << 1);  // the flags weigh 128
}
__device__ __forceinline__ uint32_t nl_mask16_ascii(const uint4& v) {
  uint32_t lo = __dp4a(nl_flags_ascii(v.x), 0x08040201u, __dp4a(nl_flags_ascii(v.y), 0x80402010u, 0u));
  uint32_t hi = __dp4a(nl_flags_ascii(v.z), 0x08040201u, __dp4a(nl_flags_ascii(v.w), 0x80402010u, 0u));
  return (lo >> 7) | (hi << 1);
}
// the same for '\r'
__device__ __forceinline__ uint32_t cr_mask16(const uint4& v) {
  auto flags = [](uint32_t w) { const uint32_t x = w ^ 0x0D0D0D0Du; return ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u; };
  uint32_t lo = __dp4a(flags(v.x), 0x08040201u, __dp4a(flags(v.y), 0x80402010u, 0u));
  uint32_t hi = __dp4a(flags(v.z), 0x08040201u, __dp4a(flags(v.w), 0x80402010u, 0u));
  return (lo >> 7) | (hi << 1);
}
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t n = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v += n;
  }
  return v;
}
__device__ __forceinline__ unsigned log2_bin(u64 len) { return len ? 64 - __clzll((long long)len) : 0; }

}  // namespace fq
