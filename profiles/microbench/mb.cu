// Microbenchmarks that ground the kernel design (DESIGN.md "measured constants"):
//  1. read-only HBM stream rate (LDG.128) -- the practical ceiling for a scan kernel
//  2. 1-D TMA bulk copy (cp.async.bulk) tile ring -> smem -> LDS.128 consume
//  3. shared-memory atomic (ATOMS/RED) throughput: lane-striped conflict-free, partial warps, conflicts
//  4. LDS.32 throughput
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o mb mb.cu
// Run under gpurun; prints one JSON object per line.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

// ---------------------------------------------------------------- 1. LDG stream
template <int UNROLL>
__global__ void __launch_bounds__(512) k_read_ldg(const uint4* __restrict__ p, size_t n16, unsigned long long* out) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t acc = 0;
  for (; i + (UNROLL - 1) * stride < n16; i += UNROLL * stride) {
    uint4 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      const uint4* q = p + i + u * stride;
      asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                   : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(q));
    }
#pragma unroll
    for (int u = 0; u < UNROLL; u++) acc += v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
  }
  for (; i < n16; i += stride) { uint4 v = p[i]; acc += v.x ^ v.y ^ v.z ^ v.w; }
  if (acc == 0x12345678u) atomicAdd(out, 1ull);
}

// ---------------------------------------------------------------- 2. TMA bulk ring
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// Persistent CTA: static round-robin tiles, STAGES-deep ring, all threads consume with LDS.128.
template <int TILE, int STAGES>
__global__ void __launch_bounds__(512) k_read_tma(const uint8_t* __restrict__ p, size_t ntiles, unsigned long long* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t full[STAGES];
  uint8_t* buf = smem;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  size_t t0 = blockIdx.x, step = gridDim.x;
  // prologue
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) {
      size_t t = t0 + (size_t)s * step;
      if (t < ntiles) { mbar_expect_tx(&full[s], TILE); tma_bulk_g2s(buf + (size_t)s * TILE, p + t * TILE, TILE, &full[s]); }
    }
  }
  uint32_t acc = 0; int stage = 0; uint32_t phase = 0;
  for (size_t t = t0; t < ntiles; t += step) {
    mbar_wait(&full[stage], phase);
    const uint4* q = reinterpret_cast<const uint4*>(buf + (size_t)stage * TILE);
#pragma unroll
    for (int i = 0; i < TILE / 16 / 512; i++) { uint4 v = q[threadIdx.x + i * 512]; acc += v.x ^ v.y ^ v.z ^ v.w; }
    __syncthreads();  // everyone done with this stage
    if (threadIdx.x == 0) {
      size_t tn = t + (size_t)STAGES * step;
      if (tn < ntiles) { mbar_expect_tx(&full[stage], TILE); tma_bulk_g2s(buf + (size_t)stage * TILE, p + tn * TILE, TILE, &full[stage]); }
    }
    if (++stage == STAGES) { stage = 0; phase ^= 1; }
  }
  if (acc == 0x12345678u) atomicAdd(out, 1ull);
}

// ---------------------------------------------------------------- 3. shared atomics
// MODE 0: lane-striped conflict-free (bin*32+lane), random bins; MODE 1: un-striped random bins (bank conflicts);
// MODE 2: all lanes same address; MODE 3: striped, but only ACTIVE lanes participate (predicated)
template <int MODE>
__global__ void __launch_bounds__(512) k_atoms(int iters, int active, unsigned long long* cyc_out, uint32_t* sink) {
  extern __shared__ uint32_t hist[];  // 256*32 words = 32 KB
  for (int i = threadIdx.x; i < 256 * 32; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  uint32_t lane = threadIdx.x & 31;
  uint32_t x = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 12345u;
  bool on = (int)lane < active;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      x = x * 1664525u + 1013904223u;
      uint32_t bin = (x >> 24);
      uint32_t idx;
      if (MODE == 0 || MODE == 3) idx = bin * 32 + lane;
      else if (MODE == 1) idx = (x >> 19);         // 13 bits -> 8192 words, random bank
      else idx = 7;
      if (MODE == 3) { if (on) atomicAdd(&hist[idx], 1u); }
      else atomicAdd(&hist[idx], 1u);
    }
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cyc_out[blockIdx.x] = (unsigned long long)(t1 - t0);
  uint32_t s = 0;
  for (int i = threadIdx.x; i < 256 * 32; i += blockDim.x) s += hist[i];
  if (s == 0xdeadbeefu) sink[0] = s;
}

// ALU-only version of the same loop (to subtract the LCG cost)
__global__ void __launch_bounds__(512) k_alu(int iters, unsigned long long* cyc_out, uint32_t* sink) {
  uint32_t x = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 12345u;
  uint32_t acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int k = 0; k < 8; k++) { x = x * 1664525u + 1013904223u; acc += (x >> 24) * 32 + (threadIdx.x & 31); }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc_out[blockIdx.x] = (unsigned long long)(t1 - t0);
  if (acc == 0xdeadbeefu) sink[0] = acc;
}

// ---------------------------------------------------------------- 4. LDS.32
__global__ void __launch_bounds__(512) k_lds(int iters, unsigned long long* cyc_out, uint32_t* sink) {
  extern __shared__ uint32_t sm[];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = i * 7u;
  __syncthreads();
  uint32_t acc = 0, idx = threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int k = 0; k < 8; k++) { acc += sm[(idx + k * 512) & 8191]; }
    idx += acc & 1;
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc_out[blockIdx.x] = (unsigned long long)(t1 - t0);
  if (acc == 0xdeadbeefu) sink[0] = acc;
}

template <typename F>
static float time_ms(F f, int reps) {
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    CK(cudaEventRecord(a)); f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int nsm = prop.multiProcessorCount;
  printf("{\"device\": \"%s\", \"sms\": %d, \"smem_optin\": %zu, \"clock_khz\": %d}\n", prop.name, nsm,
         prop.sharedMemPerBlockOptin, prop.clockRate);
  size_t nbytes = (size_t)8 << 30;
  uint8_t* d; CK(cudaMalloc(&d, nbytes)); CK(cudaMemset(d, 0x41, nbytes));
  unsigned long long* dout; CK(cudaMalloc(&dout, 4096 * sizeof(unsigned long long)));
  uint32_t* sink; CK(cudaMalloc(&sink, 64));

  for (int cps = 1; cps <= 4; cps *= 2) {
    float ms = time_ms([&] { k_read_ldg<4><<<nsm * cps, 512>>>((const uint4*)d, nbytes / 16, dout); }, 5);
    printf("{\"bench\": \"read_ldg_u4\", \"ctas_per_sm\": %d, \"ms\": %.3f, \"GBps\": %.1f}\n", cps, ms, nbytes / ms / 1e6);
    ms = time_ms([&] { k_read_ldg<8><<<nsm * cps, 512>>>((const uint4*)d, nbytes / 16, dout); }, 5);
    printf("{\"bench\": \"read_ldg_u8\", \"ctas_per_sm\": %d, \"ms\": %.3f, \"GBps\": %.1f}\n", cps, ms, nbytes / ms / 1e6);
  }
  {
    constexpr int TILE = 32768;
    CK(cudaFuncSetAttribute(k_read_tma<TILE, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE * 3));
    CK(cudaFuncSetAttribute(k_read_tma<TILE, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE * 6));
    CK(cudaFuncSetAttribute(k_read_tma<16384, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 4));
    float ms = time_ms([&] { k_read_tma<TILE, 3><<<nsm * 2, 512, TILE * 3>>>(d, nbytes / TILE, dout); }, 5);
    printf("{\"bench\": \"read_tma_32k_s3_2cta\", \"ms\": %.3f, \"GBps\": %.1f}\n", ms, nbytes / ms / 1e6);
    ms = time_ms([&] { k_read_tma<TILE, 6><<<nsm, 512, TILE * 6>>>(d, nbytes / TILE, dout); }, 5);
    printf("{\"bench\": \"read_tma_32k_s6_1cta\", \"ms\": %.3f, \"GBps\": %.1f}\n", ms, nbytes / ms / 1e6);
    ms = time_ms([&] { k_read_tma<TILE, 3><<<nsm, 512, TILE * 3>>>(d, nbytes / TILE, dout); }, 5);
    printf("{\"bench\": \"read_tma_32k_s3_1cta\", \"ms\": %.3f, \"GBps\": %.1f}\n", ms, nbytes / ms / 1e6);
    ms = time_ms([&] { k_read_tma<16384, 4><<<nsm * 3, 512, 16384 * 4>>>(d, nbytes / 16384, dout); }, 5);
    printf("{\"bench\": \"read_tma_16k_s4_3cta\", \"ms\": %.3f, \"GBps\": %.1f}\n", ms, nbytes / ms / 1e6);
  }
  // smem atomics: 1 CTA of 512 threads per SM (16 warps), iters*8 atomics per thread
  {
    int iters = 2000;
    size_t sh = 256 * 32 * 4;
    auto report = [&](const char* name, int active) {
      unsigned long long h[4096]; CK(cudaMemcpy(h, dout, nsm * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
      double avg = 0; for (int i = 0; i < nsm; i++) avg += (double)h[i]; avg /= nsm;
      double warp_instr = 16.0 * iters * 8;  // per SM
      printf("{\"bench\": \"%s\", \"active_lanes\": %d, \"cycles\": %.0f, \"cyc_per_warp_instr\": %.3f, \"lane_ops_per_clk_per_sm\": %.2f}\n",
             name, active, avg, avg / warp_instr, warp_instr * active / avg);
    };
    k_alu<<<nsm, 512>>>(iters, dout, sink); CK(cudaDeviceSynchronize()); report("alu_only", 32);
    k_atoms<0><<<nsm, 512, sh>>>(iters, 32, dout, sink); CK(cudaDeviceSynchronize()); report("atoms_striped", 32);
    k_atoms<1><<<nsm, 512, sh>>>(iters, 32, dout, sink); CK(cudaDeviceSynchronize()); report("atoms_random_bank", 32);
    k_atoms<2><<<nsm, 512, sh>>>(iters, 32, dout, sink); CK(cudaDeviceSynchronize()); report("atoms_same_addr", 32);
    for (int a : {32, 24, 16, 8, 4, 1}) {
      k_atoms<3><<<nsm, 512, sh>>>(iters, a, dout, sink); CK(cudaDeviceSynchronize()); report("atoms_striped_pred", a);
    }
    k_lds<<<nsm, 512, 32768>>>(iters, dout, sink); CK(cudaDeviceSynchronize()); report("lds32", 32);
  }
  // atomics with HBM streaming running concurrently is covered by the real kernel; done.
  CK(cudaFree(d));
  return 0;
}
