// Dual-pipe test: independent instruction streams written in PTX so nothing is folded.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)
#define LOP(x, c, d)  asm volatile("lop3.b32 %0, %0, %1, %2, 0x6A;" : "+r"(x) : "r"(c), "r"(d))
#define IMAD(x, c, d) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(c), "r"(d))
#define IDP(x, c)     asm volatile("dp4a.u32.u32 %0, %0, %1, %0;" : "+r"(x) : "r"(c))
#define PRMT(x, c)    asm volatile("prmt.b32 %0, %0, %1, 0x4321;" : "+r"(x) : "r"(c))
#define SHF(x, c)     asm volatile("shf.r.wrap.b32 %0, %0, %1, 9;" : "+r"(x) : "r"(c))
#define IADD(x, c, d) asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(c))
#define SHL(x)        asm volatile("shl.b32 %0, %0, 1;" : "+r"(x))
#define POPC(x)       asm volatile("popc.b32 %0, %0;" : "+r"(x))
#define FFMA(f, g)    asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f) : "f"(g))

template <int OP>
__global__ void __launch_bounds__(512) k(int iters, unsigned long long* cyc, uint32_t* sink, uint32_t seed) {
  uint32_t a[8]; float f[4];
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 2654435761u + i * 40503u + seed;
#pragma unroll
  for (int i = 0; i < 4; i++) f[i] = (float)a[i];
  uint32_t c = seed | 1, d = seed * 3 + 7; float g = (float)seed;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      if (OP == 0) { LOP(a[i], c, d); LOP(a[i + 4], c, d); }
      if (OP == 1) { IMAD(a[i], c, d); IMAD(a[i + 4], c, d); }
      if (OP == 2) { LOP(a[i], c, d); IMAD(a[i + 4], c, d); }
      if (OP == 3) { LOP(a[i], c, d); IDP(a[i + 4], c); }
      if (OP == 4) { PRMT(a[i], c); IDP(a[i + 4], c); }
      if (OP == 5) { PRMT(a[i], c); LOP(a[i + 4], c, d); }
      if (OP == 6) { SHF(a[i], c); IMAD(a[i + 4], c, d); }
      if (OP == 7) { IADD(a[i], c, d); IADD(a[i + 4], d, c); }
      if (OP == 8) { IADD(a[i], c, d); IMAD(a[i + 4], c, d); }
      if (OP == 9) { LOP(a[i], c, d); FFMA(f[i], g); }
      if (OP == 10) { IMAD(a[i], c, d); FFMA(f[i], g); }
      if (OP == 11) { POPC(a[i]); POPC(a[i + 4]); }
      if (OP == 12) { POPC(a[i]); IMAD(a[i + 4], c, d); }
      if (OP == 13) { SHL(a[i]); SHL(a[i + 4]); }
      if (OP == 14) { IDP(a[i], c); IDP(a[i + 4], c); }
      if (OP == 15) { LOP(a[i], c, d); LOP(a[i + 4], c, d); IMAD(a[(i + 1) & 3], c, d); }   // 2 alu : 1 fma
      if (OP == 16) { LOP(a[i], c, d); IMAD(a[i + 4], c, d); IDP(a[(i + 1) & 3], c); }      // 1 alu : 2 fma
    }
  }
  long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s ^= a[i];
#pragma unroll
  for (int i = 0; i < 4; i++) s ^= (uint32_t)f[i];
  if (threadIdx.x == 0) cyc[blockIdx.x] = (unsigned long long)(t1 - t0);
  if (s == 0xdeadbeefu) sink[0] = s;
}
int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int nsm = prop.multiProcessorCount;
  unsigned long long* dc; CK(cudaMalloc(&dc, 4096 * 8)); uint32_t* sink; CK(cudaMalloc(&sink, 64));
  unsigned long long h[4096];
  const int iters = 4000;
  const char* names[] = {"LOP|LOP", "IMAD|IMAD", "LOP|IMAD", "LOP|IDP", "PRMT|IDP", "PRMT|LOP", "SHF|IMAD", "IADD|IADD", "IADD|IMAD", "LOP|FFMA", "IMAD|FFMA", "POPC|POPC", "POPC|IMAD", "SHL|SHL", "IDP|IDP", "2LOP|IMAD", "LOP|IMAD|IDP"};
  int nops[] = {2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 3};
#define RUN(OP) { k<OP><<<nsm, 512>>>(iters, dc, sink, 12345u); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(h, dc, nsm * 8, cudaMemcpyDeviceToHost)); \
    double avg = 0; for (int i = 0; i < nsm; i++) avg += h[i]; avg /= nsm; double wi = 16.0 * iters * 4 * nops[OP]; \
    printf("{\"bench\": \"pipe_%s\", \"warp_instr_per_clk_per_sm\": %.3f}\n", names[OP], wi / avg); }
  RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10) RUN(11) RUN(12) RUN(13) RUN(14) RUN(15) RUN(16)
  return 0;
}
