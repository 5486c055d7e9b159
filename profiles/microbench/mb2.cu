// Instruction-throughput microbenchmarks (per SM, 16 resident warps): which pipe each candidate
// SWAR primitive uses and whether two pipes overlap.  Prints warp-instr/clk/SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

template <int OP>
__global__ void __launch_bounds__(512) k_op(int iters, unsigned long long* cyc, uint32_t* sink, uint32_t seed) {
  uint32_t a[8];
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 2654435761u + i * 40503u + seed;
  uint32_t c = seed | 1, d = seed * 3 + 7;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (OP == 0) a[i] = __dp4a(a[i], 0x80000000u >> (8 * (i & 3)), a[i]);          // IDP.4A
      else if (OP == 1) a[i] = __byte_perm(a[i], c, 0x4140 + i);                          // PRMT
      else if (OP == 2) a[i] = (a[i] & c) ^ d;                                            // LOP3
      else if (OP == 3) a[i] = a[i] * c + d;                                              // IMAD
      else if (OP == 4) a[i] = __funnelshift_r(a[i], c, 8 + i);                           // SHF
      else if (OP == 5) a[i] = __popc(a[i]) + c;                                          // POPC (+IADD)
      else if (OP == 6) { a[i] = (a[i] & c) ^ d; a[i] = a[i] * c + d; }                   // LOP3 + IMAD alternating
      else if (OP == 7) { a[i] = (a[i] & c) ^ d; a[i] = __dp4a(a[i], 0x00800000u, a[i]); } // LOP3 + IDP
      else if (OP == 8) { a[i] = __byte_perm(a[i], c, 0x4140 + i); a[i] = (a[i] & c) ^ d; } // PRMT + LOP3 (same pipe?)
      else if (OP == 9) a[i] = a[i] + c + d;                                              // IADD3
      else if (OP == 10) { a[i] = __dp4a(a[i], 0x00800000u, a[i]); a[i] = a[i] * c + d; }  // IDP + IMAD (same pipe?)
      else if (OP == 11) a[i] = __vabsdiffu4(a[i], c) + d;                                // VABSDIFF4 (+IADD)
    }
  }
  long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s ^= a[i];
  if (threadIdx.x == 0) cyc[blockIdx.x] = (unsigned long long)(t1 - t0);
  if (s == 0xdeadbeefu) sink[0] = s;
}

// shared atomics without ALU in the loop: 8 fixed lane-striped addresses per thread
template <int MODE>
__global__ void __launch_bounds__(512) k_atoms2(int iters, unsigned long long* cyc, uint32_t* sink, uint32_t seed) {
  extern __shared__ uint32_t hist[];
  for (int i = threadIdx.x; i < 256 * 32; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  uint32_t lane = threadIdx.x & 31;
  uint32_t idx[8];
#pragma unroll
  for (int k = 0; k < 8; k++) {
    uint32_t x = (threadIdx.x * 2654435761u + k * 40503u + seed);
    uint32_t bin = (MODE == 2) ? ((x >> 13) & 3) * 9 + 35 : (x >> 24);     // MODE 2: 4 hot bins
    idx[k] = (MODE == 1 || MODE == 2) ? bin : bin * 32 + lane;                // MODE 1/2: un-striped (1 KB table)
  }
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      if (MODE == 3) atomicAdd(&hist[idx[k]], idx[k] | 2u);   // ATOMS.ADD with a data operand
      else atomicAdd(&hist[idx[k]], 1u);
    }
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cyc[blockIdx.x] = (unsigned long long)(t1 - t0);
  uint32_t s = 0;
  for (int i = threadIdx.x; i < 256 * 32; i += blockDim.x) s += hist[i];
  if (s == 0xdeadbeefu) sink[0] = s;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int nsm = prop.multiProcessorCount;
  unsigned long long* dc; CK(cudaMalloc(&dc, 4096 * 8)); uint32_t* sink; CK(cudaMalloc(&sink, 64));
  unsigned long long h[4096];
  const int iters = 4000;
  const char* names[] = {"IDP.4A", "PRMT", "LOP3", "IMAD", "SHF", "POPC+IADD", "LOP3+IMAD", "LOP3+IDP", "PRMT+LOP3", "IADD3", "IDP+IMAD", "VABSDIFF4+IADD"};
  int nops[] = {1, 1, 1, 1, 1, 2, 2, 2, 2, 1, 2, 2};
#define RUN(OP) { k_op<OP><<<nsm, 512>>>(iters, dc, sink, 12345u); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(h, dc, nsm * 8, cudaMemcpyDeviceToHost)); \
    double avg = 0; for (int i = 0; i < nsm; i++) avg += h[i]; avg /= nsm; double wi = 16.0 * iters * 8 * nops[OP]; \
    printf("{\"bench\": \"op_%s\", \"warp_instr_per_clk_per_sm\": %.3f, \"cycles\": %.0f}\n", names[OP], wi / avg, avg); }
  RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10) RUN(11)
#define RUNA(MODE, NAME) { k_atoms2<MODE><<<nsm, 512, 32768>>>(iters, dc, sink, 777u); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(h, dc, nsm * 8, cudaMemcpyDeviceToHost)); \
    double avg = 0; for (int i = 0; i < nsm; i++) avg += h[i]; avg /= nsm; double wi = 16.0 * iters * 8; \
    printf("{\"bench\": \"%s\", \"cyc_per_warp_instr\": %.3f, \"lane_ops_per_clk_per_sm\": %.2f}\n", NAME, avg / wi, wi * 32 / avg); }
  RUNA(0, "atoms_striped_noalu") RUNA(1, "atoms_unstriped_256bins") RUNA(2, "atoms_unstriped_4hotbins") RUNA(3, "atoms_add_striped")
  return 0;
}
