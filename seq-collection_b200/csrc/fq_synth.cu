// fq_synth.cu -- counter-based synthetic FASTQ generators.
// Every byte is a pure function of (seed, absolute position), so any byte range of the Illumina
// stream can be produced independently on any GPU (multi-GPU shards are generated in place), and
// the same bytes can be copied back for the CPU oracle.  Not part of the timed path.
#include <cuda_runtime.h>
#include <stdint.h>
#include <thrust/execution_policy.h>
#include <thrust/scan.h>

namespace fq {
typedef unsigned long long u64;

__host__ __device__ __forceinline__ u64 splitmix64(u64 x) {
  x += 0x9E3779B97F4A7C15ull;
  u64 z = x;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// ---------------------------------------------------------------------------------------------
// Illumina 2x150, interleaved pairs.  Record = 360 bytes:
//   "@A00156:217:HKJWGDSXX:L:TTTT:XXXXX:YYYYY R:N:0:AACGCTTA\n"   56  (header of tests/fastq/novaseq.fq:1, fixed width)
//   150 bases '\n'                                                151  P(A,C,G,T,N) = .2945 .205 .205 .2945 .001
//   "+\n"                                                           2
//   150 quals '\n'                                                151  NovaSeq 4-bin set F : , #  P = .90 .06 .035 .005
// ---------------------------------------------------------------------------------------------
constexpr int ILL_REC = 360;
constexpr int ILL_HDR = 56;
constexpr int ILL_LEN = 150;

__device__ __forceinline__ uint8_t ill_byte(u64 rec, int k, u64 seed) {
  if (k < ILL_HDR) {
    const char* tmpl = "@A00156:217:HKJWGDSXX:L:TTTT:XXXXX:YYYYY R:N:0:AACGCTTA\n";
    if (k < 22 || k == 23 || k == 28 || k == 34 || k >= 40) {
      if (k == 41) return (uint8_t)('1' + (rec & 1));
      return (uint8_t)tmpl[k];
    }
    const u64 h = splitmix64(seed ^ ((rec >> 1) * 0xD1B54A32D192ED03ull + 0x1234567ull));
    if (k == 22) return (uint8_t)('1' + (h & 3));
    uint32_t val; int digit;  // digit 0 = most significant
    if (k < 28) { val = 1101u + (uint32_t)((h >> 8) % 1578u); digit = k - 24; val = (val / (digit == 0 ? 1000u : digit == 1 ? 100u : digit == 2 ? 10u : 1u)) % 10u; }
    else if (k < 34) { val = 1000u + (uint32_t)((h >> 24) % 31000u); digit = k - 29; val = (val / (digit == 0 ? 10000u : digit == 1 ? 1000u : digit == 2 ? 100u : digit == 3 ? 10u : 1u)) % 10u; }
    else { val = 1000u + (uint32_t)((h >> 44) % 35000u); digit = k - 35; val = (val / (digit == 0 ? 10000u : digit == 1 ? 1000u : digit == 2 ? 100u : digit == 3 ? 10u : 1u)) % 10u; }
    return (uint8_t)('0' + val);
  }
  k -= ILL_HDR;
  if (k < ILL_LEN) {
    const u64 r = splitmix64(seed ^ (rec * 128ull + (u64)(k >> 2)) * 0x9E3779B97F4A7C15ull);
    const uint32_t u = (uint32_t)(r >> (16 * (k & 3))) & 0xFFFFu;
    return u < 19300u ? 'A' : u < 32735u ? 'C' : u < 46170u ? 'G' : u < 65470u ? 'T' : 'N';
  }
  if (k == ILL_LEN) return '\n';
  if (k == ILL_LEN + 1) return '+';
  if (k == ILL_LEN + 2) return '\n';
  k -= ILL_LEN + 3;
  if (k < ILL_LEN) {
    const u64 r = splitmix64(seed ^ (rec * 128ull + 64ull + (u64)(k >> 2)) * 0x9E3779B97F4A7C15ull);
    const uint32_t u = (uint32_t)(r >> (16 * (k & 3))) & 0xFFFFu;
    return u < 58982u ? 'F' : u < 62914u ? ':' : u < 65208u ? ',' : '#';
  }
  return '\n';
}

__global__ void synth_illumina_kernel(uint8_t* __restrict__ out, u64 first_byte, u64 nbytes, u64 seed) {
  const u64 nvec = (nbytes + 15) / 16;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (u64)gridDim.x * blockDim.x) {
    const u64 o = i * 16;
    u64 pos = first_byte + o;
    u64 rec = pos / ILL_REC;
    int k = (int)(pos - rec * ILL_REC);
    uint32_t w[4] = {0, 0, 0, 0};
    const int n = (nbytes - o) < 16 ? (int)(nbytes - o) : 16;
    for (int b = 0; b < n; b++) {
      w[b >> 2] |= (uint32_t)ill_byte(rec, k, seed) << (8 * (b & 3));
      if (++k == ILL_REC) { k = 0; rec++; }
    }
    if (n == 16 && (((uintptr_t)(out + o)) & 15) == 0) *reinterpret_cast<uint4*>(out + o) = make_uint4(w[0], w[1], w[2], w[3]);
    else for (int b = 0; b < n; b++) out[o + b] = (uint8_t)(w[b >> 2] >> (8 * (b & 3)));
  }
}

cudaError_t launch_synth_illumina(void* dptr, u64 first_byte, u64 nbytes, u64 seed, cudaStream_t st) {
  if (nbytes == 0) return cudaSuccess;
  synth_illumina_kernel<<<148 * 16, 256, 0, st>>>((uint8_t*)dptr, first_byte, nbytes, seed);
  return cudaGetLastError();
}

// The generator's own tallies ("emitting their own expected tallies as a third, independent
// check"): the statistics of records [first, first + n) from the random numbers alone -- no byte is written or read,
// nothing of the scan is involved.  out[0..4] = A C G T N, out[5..8] = F : , #, out[16 + p] = sum of the quality bytes
// at position p.  Thread (r, q) takes quad q of records r, r + R, ...
constexpr int ILL_QUADS = (ILL_LEN + 3) / 4;
__global__ void synth_illumina_tally_kernel(u64 first_record, u64 n_records, u64 seed, u64* __restrict__ out) {
  const u64 gt = (u64)blockIdx.x * blockDim.x + threadIdx.x, total = (u64)gridDim.x * blockDim.x;
  const u64 R = total / ILL_QUADS;
  if (gt >= R * ILL_QUADS) return;
  const int q = (int)(gt % ILL_QUADS);
  unsigned long long cb[5] = {0, 0, 0, 0, 0}, cq[4] = {0, 0, 0, 0}, ps[4] = {0, 0, 0, 0};
  for (u64 i = gt / ILL_QUADS; i < n_records; i += R) {
    const u64 rec = first_record + i;
    const u64 rb = splitmix64(seed ^ (rec * 128ull + (u64)q) * 0x9E3779B97F4A7C15ull);
    const u64 rq = splitmix64(seed ^ (rec * 128ull + 64ull + (u64)q) * 0x9E3779B97F4A7C15ull);
#pragma unroll
    for (int j = 0; j < 4; j++) {
      if (4 * q + j >= ILL_LEN) break;
      const uint32_t ub = (uint32_t)(rb >> (16 * j)) & 0xFFFFu, uq = (uint32_t)(rq >> (16 * j)) & 0xFFFFu;
      cb[ub < 19300u ? 0 : ub < 32735u ? 1 : ub < 46170u ? 2 : ub < 65470u ? 3 : 4]++;
      const int qi = uq < 58982u ? 0 : uq < 62914u ? 1 : uq < 65208u ? 2 : 3;
      cq[qi]++;
      ps[j] += qi == 0 ? 'F' : qi == 1 ? ':' : qi == 2 ? ',' : '#';
    }
  }
  for (int k = 0; k < 5; k++) if (cb[k]) atomicAdd(&out[k], cb[k]);
  for (int k = 0; k < 4; k++) if (cq[k]) atomicAdd(&out[5 + k], cq[k]);
  for (int j = 0; j < 4; j++) if (ps[j]) atomicAdd(&out[16 + 4 * q + j], ps[j]);
}
// quality range of the first m records (the fq-meta fold), on the host
void synth_illumina_meta_range(u64 first_record, u64 m, u64 seed, long long* qmin, long long* qmax) {
  int mn = 1000, mx = -1;
  for (u64 i = 0; i < m; i++) {
    const u64 rec = first_record + i;
    for (int q = 0; q < ILL_QUADS; q++) {
      const u64 rq = splitmix64(seed ^ (rec * 128ull + 64ull + (u64)q) * 0x9E3779B97F4A7C15ull);
      for (int j = 0; j < 4 && 4 * q + j < ILL_LEN; j++) {
        const uint32_t uq = (uint32_t)(rq >> (16 * j)) & 0xFFFFu;
        const int v = (uq < 58982u ? 'F' : uq < 62914u ? ':' : uq < 65208u ? ',' : '#') - 33;
        mn = v < mn ? v : mn; mx = v > mx ? v : mx;
      }
    }
  }
  *qmin = m ? mn : -1; *qmax = m ? mx : -1;
}
cudaError_t launch_synth_illumina_tally(u64 first_record, u64 n_records, u64 seed, u64* d_out /* [16 + 152] zeroed */, cudaStream_t st) {
  synth_illumina_tally_kernel<<<148 * 8, 256, 0, st>>>(first_record, n_records, seed, d_out);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// ONT-style long reads: L = clip(round(exp(N(ln 9000, 0.9))), 1000, 100000); header
//   "@<32 hex> runid=<40 hex> read=<n> ch=<1..512> start_time=2026-01-01T00:00:00Z"
// bases iid uniform ACGT; quals phred+33 with Q = clip(round(N(18,7)), 1, 50).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float u01(u64 r) { return ((float)(uint32_t)(r >> 40) + 0.5f) * (1.0f / 16777216.0f); }

__device__ uint32_t ont_len(u64 rec, u64 seed) {
  const u64 r1 = splitmix64(seed ^ (rec * 4ull + 1ull) * 0xA24BAED4963EE407ull);
  const u64 r2 = splitmix64(seed ^ (rec * 4ull + 2ull) * 0xA24BAED4963EE407ull);
  const float z = sqrtf(-2.0f * logf(u01(r1))) * cospif(2.0f * u01(r2));
  float L = rintf(expf(9.104979856f + 0.9f * z));  // ln 9000
  L = fminf(fmaxf(L, 1000.0f), 100000.0f);
  return (uint32_t)L;
}
__device__ int dec_digits(u64 v) { int d = 1; while (v >= 10) { v /= 10; d++; } return d; }
__device__ int ont_hdr_len(u64 rec, u64 seed) {
  const uint32_t ch = 1u + (uint32_t)(splitmix64(seed ^ (rec * 4ull + 3ull)) % 512u);
  // '@' 32 " runid=" 40 " read=" n " ch=" c " start_time=2026-01-01T00:00:00Z"
  return 1 + 32 + 7 + 40 + 6 + dec_digits(rec) + 4 + dec_digits(ch) + 12 + 20;
}
__global__ void ont_sizes_kernel(u64 first_record, u64 n, u64 seed, u64* sizes) {
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const u64 rec = first_record + i;
  sizes[i] = (u64)ont_hdr_len(rec, seed) + 1 + 2ull * ont_len(rec, seed) + 1 + 2 + 1;
}
__device__ __forceinline__ char hexd(uint32_t v) { return (char)(v < 10 ? '0' + v : 'a' + (v - 10)); }

__global__ void ont_write_kernel(uint8_t* __restrict__ out, u64 first_record, u64 n, u64 seed, const u64* __restrict__ offs) {
  for (u64 i = blockIdx.x; i < n; i += gridDim.x) {
    const u64 rec = first_record + i;
    uint8_t* p = out + offs[i];
    const uint32_t L = ont_len(rec, seed);
    const int H = ont_hdr_len(rec, seed);
    if (threadIdx.x == 0) {  // header, sequentially
      int k = 0;
      p[k++] = '@';
      u64 a = splitmix64(seed ^ (rec * 8ull + 5ull)), b = splitmix64(seed ^ (rec * 8ull + 6ull)), c = splitmix64(seed ^ 0x5eedull);
      for (int j = 0; j < 16; j++) p[k++] = hexd((uint32_t)(a >> (4 * j)) & 15);
      for (int j = 0; j < 16; j++) p[k++] = hexd((uint32_t)(b >> (4 * j)) & 15);
      const char* s1 = " runid=";
      for (int j = 0; s1[j]; j++) p[k++] = s1[j];
      for (int j = 0; j < 40; j++) p[k++] = hexd((uint32_t)(splitmix64(c + (u64)(j >> 4)) >> (4 * (j & 15))) & 15);
      const char* s2 = " read=";
      for (int j = 0; s2[j]; j++) p[k++] = s2[j];
      { int d = dec_digits(rec); u64 v = rec; for (int j = d - 1; j >= 0; j--) { p[k + j] = (uint8_t)('0' + v % 10); v /= 10; } k += d; }
      const char* s3 = " ch=";
      for (int j = 0; s3[j]; j++) p[k++] = s3[j];
      { uint32_t ch = 1u + (uint32_t)(splitmix64(seed ^ (rec * 4ull + 3ull)) % 512u); int d = dec_digits(ch); u64 v = ch; for (int j = d - 1; j >= 0; j--) { p[k + j] = (uint8_t)('0' + v % 10); v /= 10; } k += d; }
      const char* s4 = " start_time=2026-01-01T00:00:00Z";
      for (int j = 0; s4[j]; j++) p[k++] = s4[j];
      p[k++] = '\n';
      p[H + 1 + L] = '\n';
      p[H + 1 + L + 1] = '+';
      p[H + 1 + L + 2] = '\n';
      p[H + 1 + L + 3 + L] = '\n';
    }
    uint8_t* sq = p + H + 1;
    uint8_t* ql = p + H + 1 + L + 3;
    for (uint32_t k = threadIdx.x; k < L; k += blockDim.x) {
      const u64 r = splitmix64(seed ^ (rec * 0x100000ull + (u64)(k >> 3)) * 0x9E3779B97F4A7C15ull);
      sq[k] = "ACGT"[(r >> (8 * (k & 7))) & 3];
      const u64 q1 = splitmix64(seed ^ (rec * 0x100000ull + 0x80000ull + (u64)k) * 0xD6E8FEB86659FD93ull);
      const float z = sqrtf(-2.0f * logf(u01(q1))) * cospif(2.0f * u01(q1 * 0x2545F4914F6CDD1Dull));
      float q = rintf(18.0f + 7.0f * z);
      q = fminf(fmaxf(q, 1.0f), 50.0f);
      ql[k] = (uint8_t)(33 + (int)q);
    }
  }
}

cudaError_t synth_ont(void* dptr, size_t capacity, u64 first_record, u64 n_records, u64 seed, size_t* bytes_written,
                      cudaStream_t st) {
  *bytes_written = 0;
  if (n_records == 0) return cudaSuccess;
  u64* sizes = nullptr;
  cudaError_t e = cudaMalloc(&sizes, (n_records + 1) * sizeof(u64));
  if (e != cudaSuccess) return e;
  cudaMemsetAsync(sizes + n_records, 0, sizeof(u64), st);
  ont_sizes_kernel<<<(unsigned)((n_records + 255) / 256), 256, 0, st>>>(first_record, n_records, seed, sizes);
  thrust::exclusive_scan(thrust::cuda::par.on(st), sizes, sizes + n_records + 1, sizes);
  u64 total = 0;
  e = cudaMemcpyAsync(&total, sizes + n_records, sizeof(u64), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e == cudaSuccess && total > capacity) e = cudaErrorInvalidValue;
  if (e == cudaSuccess) {
    unsigned grid = (unsigned)(n_records < 148 * 8 ? n_records : 148 * 8);
    ont_write_kernel<<<grid, 256, 0, st>>>((uint8_t*)dptr, first_record, n_records, seed, sizes);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) *bytes_written = (size_t)total;
  }
  cudaFree(sizes);
  return e;
}

}  // namespace fq
