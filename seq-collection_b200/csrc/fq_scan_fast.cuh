// fq_scan_fast.cuh -- EXPERIMENTAL pass 0 of the FASTQ scan for well-formed input (included by fq_scan.cu inside
// namespace fq; selected by FQGPU_SCAN=fast only: measured slower than fq_scan_kernel).
//
// Same contract as fq_scan_kernel pass 0 (src/fq_count.nim:38-45 over one span: counter block into `pending`,
// SpanDesc T / head_len / tail_len), but built for the common case only and ABANDONING the span -- SpanDesc.pad = 1,
// the stitch kernel then marks it SPAN_RESCAN and fq_scan_kernel pass 1 redoes it exactly -- whenever the input
// leaves that case: bytes >= 0x80, a '\r' inside a sequence / quality line, more than two newlines in one
// 32-byte group, a sequence / quality line that lies wholly inside a group between two newlines, or an open line
// of 2 GB and more.  The newline counts (T, head_len, tail_len) are exact on any input, so the stitch kernel's
// prefix and the exact phases of the rescanned spans do not depend on the abandoned statistics.
//
// Layout of the work: no shared-memory tile and no warp roles.  A tile is 12 KiB (12 warps); warp w owns bytes
// [1024 w, 1024 w + 1024) of it and every lane one aligned 32-byte GROUP, loaded straight from global memory
// into registers one tile ahead.
//   phase A  newline mask of the group (SWAR compare + IDP.4A movemask), warp ballots of "has a newline" /
//            "has two", the chunk's newline count and the bytes after its last newline -> entry[w].
//   barrier  (the only one per tile)
//   phase B  every warp folds the 12 entries into its own start (lines before the chunk, bytes of the open
//            line), every lane derives the line class and line position of its group from the two ballots,
//            and the group goes through the statistics with ONE byte mask: a group holds at most one counted
//            segment besides the "\n+\n" case -- bytes before the first newline when the class of byte 0 is
//            sequence / quality, bytes after it otherwise -- so all lanes run the same code: masked-out bytes
//            become '\n', whose bin can never hold content and is cleared at the end.
//            "\n+\n" groups (sequence tail, quality head) push their second segment onto a per-warp stack that is
//            worked off 32 entries at a time by the same routine.
// Per-position sums, histograms and length tables are the ones of fq_scan_kernel (lane-striped histograms,
// bank-skewed 16-bit pairs).
#pragma once

constexpr int F_THREADS = 384;
constexpr int F_NW = F_THREADS / 32;     // warps per CTA
constexpr int F_GROUP = 32;              // bytes per lane and tile
constexpr int F_CHUNK = 32 * F_GROUP;    // bytes per warp and tile
constexpr int F_TILE = F_NW * F_CHUNK;   // 12 KiB (spans are cut by the host in units of TILE = 16 KiB: the last tile of a span is partial)
constexpr int F_QCAP = 64;               // stack slots per warp (popped at 32)
constexpr int M_ABOVE = 33;              // masks[k]: bytes < k (k = 0..32); masks[33 + k]: bytes > k
constexpr int M_NONE = M_ABOVE + 32;     // no byte
constexpr int M_COUNT = M_NONE + 1;

struct __align__(128) FastSmem {
  uint32_t hist[2][HB * 32];             // [0] sequence, [1] quality; word index = byte*32 + lane
  uint4 qdata[F_NW][F_QCAP][2];          // deferred groups (second segment of "\n+\n" groups)
  uint32_t qmeta[F_NW][F_QCAP];          // last newline | quality << 5 | table copy << 6 | valid bytes << 8
  uint4 masks[M_COUNT][2];
  uint32_t ptab[PT_WORDS];
  uint32_t gpos[POS_BINS + 2];
  uint32_t seq_len[POS_BINS + 2];
  uint32_t qual_len[POS_BINS + 2];
  uint32_t seq_log2[LOG2_BINS];
  uint2 entry[2][32];                    // (newlines of the chunk, bytes after its last newline / valid bytes when none); slots >= F_NW stay 0
  uint32_t first[2][F_NW];               // offset of the chunk's first newline (while the span's head fragment is open)
  u64 len_min[2], len_max[2], pos_over;
  uint32_t abandon_a[2];                 // set before the tile barrier (phase A, slot = tile parity) and read after it: the span is only counted from there on
  uint32_t abandon_b;                    // set anywhere; read at the end
  uint32_t ksel[8];
};
static_assert(sizeof(FastSmem) <= 115712, "two CTAs per SM");
#define FS_OFF(f) ((uint32_t)offsetof(FastSmem, f))

__device__ __forceinline__ uint4 ldg128(const uint8_t* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t x) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(x) : "memory"); }
__device__ __forceinline__ void sts64(uint32_t addr, uint32_t x, uint32_t y) {
  asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t sel_nl(uint32_t v, uint32_t m) { return (v & m) | (0x0A0A0A0Au & ~m); }
__device__ __forceinline__ uint4 sel_nl4(const uint4& v, const uint4& m) {
  return make_uint4(sel_nl(v.x, m.x), sel_nl(v.y, m.y), sel_nl(v.z, m.z), sel_nl(v.w, m.w));
}
__device__ __forceinline__ uint4 and4(const uint4& v, const uint4& m) { return make_uint4(v.x & m.x, v.y & m.y, v.z & m.z, v.w & m.w); }

// Per-position sums of one 16-byte half (bytes outside the counted segment are zero): q = 16 + line position of byte 0.
__device__ __forceinline__ void fast_pos16(uint32_t sm0, const uint4& v, uint32_t pt_s, uint32_t q, u64& over) {
  if (q <= (uint32_t)POS_BINS) {
    const uint32_t sh = (q & 1u) << 3;
    const uint32_t w0 = v.x << sh, w1 = __funnelshift_l(v.x, v.y, sh), w2 = __funnelshift_l(v.y, v.z, sh);
    const uint32_t w3 = __funnelshift_l(v.z, v.w, sh), w4 = __funnelshift_l(v.w, 0u, sh);
    const uint32_t A = q >> 1;
    const uint32_t r0 = pt_s + 4u * ((A & 7u) * PT_STRIDE + (A >> 3));
    red_add_at<4 * PT_STRIDE * 0>(r0, __byte_perm(w0, 0u, 0x4140)); red_add_at<4 * PT_STRIDE * 1>(r0, __byte_perm(w0, 0u, 0x4342));
    red_add_at<4 * PT_STRIDE * 2>(r0, __byte_perm(w1, 0u, 0x4140)); red_add_at<4 * PT_STRIDE * 3>(r0, __byte_perm(w1, 0u, 0x4342));
    red_add_at<4 * PT_STRIDE * 4>(r0, __byte_perm(w2, 0u, 0x4140)); red_add_at<4 * PT_STRIDE * 5>(r0, __byte_perm(w2, 0u, 0x4342));
    red_add_at<4 * PT_STRIDE * 6>(r0, __byte_perm(w3, 0u, 0x4140)); red_add_at<4 * PT_STRIDE * 7>(r0, __byte_perm(w3, 0u, 0x4342));
    if (sh) red_add_at<4 * PT_STRIDE * 8>(r0, w4);
  } else if (q >= (uint32_t)POS_BINS + 16u) {
    over += __dp4a(v.x, 0x01010101u, __dp4a(v.y, 0x01010101u, __dp4a(v.z, 0x01010101u, __dp4a(v.w, 0x01010101u, 0u))));
  } else {  // the one half per long line that straddles POS_BINS
    uint32_t w0 = v.x, w1 = v.y, w2 = v.z, w3 = v.w;
#pragma unroll 1
    for (uint32_t p = q - 16u; p < q; p++) {
      const uint32_t b = w0 & 0xFFu;
      if (p < (uint32_t)POS_BINS) asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(sm0 + FS_OFF(gpos) + 4u * p), "r"(b) : "memory");
      else over += b;
      w0 = __funnelshift_r(w0, w1, 8); w1 = __funnelshift_r(w1, w2, 8); w2 = __funnelshift_r(w2, w3, 8); w3 >>= 8;
    }
  }
}

// Works off `n` stack entries of this warp starting at `first`: bytes after the group's last newline (and below
// its valid length) are the head of a sequence / quality line.
__device__ __noinline__ void fast_side_pass(uint32_t sm0, int warp, int lane, uint32_t first, uint32_t n, bool core, u64& over) {
  if ((uint32_t)lane < n) {
    Sel ksel;
    ksel.h0 = 0x80u; ksel.h1 = 0x8000u; ksel.h2 = 0x800000u; ksel.h3 = 0x80000000u; ksel.p1 = ksel.p2 = 0;
    const uint32_t e = (uint32_t)warp * F_QCAP + first + (uint32_t)lane;
    const uint4 lo = lds128(sm0 + FS_OFF(qdata) + 32u * e), hi = lds128(sm0 + FS_OFF(qdata) + 32u * e + 16u);
    const uint32_t meta = lds32(sm0 + FS_OFF(qmeta) + 4u * e);
    const uint32_t k2 = meta & 31u, vb = meta >> 8;
    const bool ql = (meta >> 5) & 1u;
    const uint32_t ma = sm0 + FS_OFF(masks) + 32u * ((uint32_t)M_ABOVE + k2), mb = sm0 + FS_OFF(masks) + 32u * vb;
    const uint4 ml = and4(lds128(ma), lds128(mb)), mh = and4(lds128(ma + 16u), lds128(mb + 16u));
    const uint32_t tb = sm0 + FS_OFF(hist) + 4u * (uint32_t)lane + (ql ? HB * 32u * 4u : 0u);
    hist16(ksel, sel_nl4(lo, ml), tb);
    hist16(ksel, sel_nl4(hi, mh), tb);
    if (ql && !core) {
      const uint32_t pt = sm0 + FS_OFF(ptab) + (((meta >> 6) & 1u) ? PT_COPY * 4u : 0u);
      const uint32_t qe = 15u - k2;  // 16 + line position of byte 0 (negative when the newline is in the upper half)
      if (k2 < 16u) fast_pos16(sm0, and4(lo, ml), pt, qe, over);
      fast_pos16(sm0, and4(hi, mh), pt, qe + 16u, over);
    }
  }
  __syncwarp();
}

__global__ void __launch_bounds__(F_THREADS, 2) fq_scan_fast_kernel(const ScanArgs a) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  FastSmem& sm = *reinterpret_cast<FastSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int span = blockIdx.x;
  SpanDesc& desc = a.desc[span];
  const uint32_t phase = desc.guess;
  const bool count_only = phase > 3;
  const bool core = a.core != 0;
  u64* block = a.pending + (size_t)span * BLOCK_WORDS;
  const uint32_t sm0 = smem_u32(smem_raw);
  const u64 span_off = (u64)span * a.tps * TILE;
  const u64 lim = min(a.end, span_off + (u64)a.tps * TILE);   // the span is [span_off, lim) (span 0 starts at lo0)
  const uint32_t span_bytes = (uint32_t)(lim - span_off);     // < 2 GiB (host)
  const int nt = (int)((span_bytes + F_TILE - 1) / F_TILE);
  const int it_partial = (span_bytes % F_TILE) ? nt - 1 : -1;

  for (int i = tid; i < 2 * HB * 32; i += F_THREADS) (&sm.hist[0][0])[i] = 0;
  for (int i = tid; i < POS_BINS + 2; i += F_THREADS) { sm.seq_len[i] = 0; sm.qual_len[i] = 0; sm.gpos[i] = 0; }
  for (int i = tid; i < PT_WORDS; i += F_THREADS) sm.ptab[i] = 0;
  if (tid < LOG2_BINS) sm.seq_log2[tid] = 0;
  if (tid < 64) (&sm.entry[0][0])[tid] = make_uint2(0u, 0u);
  for (int i = tid; i < M_COUNT * 8; i += F_THREADS) {
    const int n = i >> 3, w = i & 7;
    const int below = n < M_ABOVE ? n : n - M_ABOVE + 1;  // bytes < below ...
    const int k = below - 4 * w;
    const uint32_t bw = k >= 4 ? 0xFFFFFFFFu : (k <= 0 ? 0u : ((1u << (8 * k)) - 1u));
    (&sm.masks[0][0].x)[i] = n < M_ABOVE ? bw : ~bw;        // ... or their complement (bytes > n - 33)
  }
  for (int i = tid; i < BLOCK_WORDS; i += F_THREADS) block[i] = (i == OFF_SEQ_LEN_MIN || i == OFF_QUAL_LEN_MIN) ? ~0ull : 0ull;
  if (tid == 0) {
    sm.len_min[0] = sm.len_min[1] = ~0ull;
    sm.len_max[0] = sm.len_max[1] = 0;
    sm.pos_over = 0; sm.abandon_a[0] = sm.abandon_a[1] = 0; sm.abandon_b = 0;
    for (int k = 0; k < 4; k++) { sm.ksel[k] = 0x80u << (8 * k); sm.ksel[4 + k] = 1u << (8 * k); }
  }
  __syncthreads();
  Sel ksel;
  {
    const uint32_t ks = sm0 + FS_OFF(ksel);
    ksel.h0 = lds32(ks); ksel.h1 = lds32(ks + 4); ksel.h2 = lds32(ks + 8); ksel.h3 = lds32(ks + 12);
    ksel.p1 = ksel.p2 = 0;
  }

  const uint32_t lt = (1u << lane) - 1u;
  const uint32_t hb_lane = sm0 + FS_OFF(hist) + 4u * (uint32_t)lane;
  const uint32_t masks_s = sm0 + FS_OFF(masks);
  const uint32_t ptab_s = sm0 + FS_OFF(ptab);
  const uint32_t my_off = (uint32_t)warp * F_CHUNK + (uint32_t)lane * F_GROUP;
  uint32_t open_tile = span == 0 ? 0u - a.lo0 : 0u;  // bytes of the open line before the tile (span-relative, mod 2^32)
  uint32_t L_tile = 0;                               // newlines of the span before the tile
  uint32_t head_len = 0;
  uint32_t qacc = 0, qcount = 0;
  bool abandoned = false;                            // phase A left the well-formed case (sticky, the same in every warp)
  uint32_t mns = ~0u, mxs = 0, mnq = ~0u, mxq = 0;
  u64 over = 0;
  uint32_t par = 0, ent_s = sm0 + FS_OFF(entry), first_s = sm0 + FS_OFF(first);  // slots of the current tile (toggle per tile)

  const uint8_t* gp = a.base + span_off + my_off;  // this lane's group of the NEXT tile to load
  uint4 nlo = make_uint4(0, 0, 0, 0), nhi = nlo;
  auto load_tile = [&](int it) {
    if (it != it_partial) {
      nlo = ldg128(gp); nhi = ldg128(gp + 16);
    } else {  // the span's last tile
      const uint32_t g = (uint32_t)it * F_TILE + my_off;
      nlo = make_uint4(0, 0, 0, 0); nhi = nlo;
      if (g < span_bytes) nlo = ldg128(gp);
      if (g + 16u < span_bytes) nhi = ldg128(gp + 16);
    }
    gp += F_TILE;
  };
  if (nt > 0) {
    load_tile(0);
    if (span == 0 && tid == 0 && a.lo0) {  // bytes before the launch's first byte
      const uint4 z = lds128(masks_s + 32u * a.lo0);
      nlo.x &= ~z.x; nlo.y &= ~z.y; nlo.z &= ~z.z; nlo.w &= ~z.w;
    }
  }

#pragma unroll 1
  for (int it = 0; it < nt; it++) {
    uint4 lo = nlo, hi = nhi;
    if (it + 1 < nt) load_tile(it + 1);
    const bool partial = it == it_partial;
    uint32_t vb = 32u, nvalid = F_CHUNK;
    if (partial) {
      const uint32_t cs = (uint32_t)it * F_TILE + (uint32_t)warp * F_CHUNK, g = cs + (uint32_t)lane * F_GROUP;
      nvalid = cs >= span_bytes ? 0u : min((uint32_t)F_CHUNK, span_bytes - cs);
      vb = g >= span_bytes ? 0u : min((uint32_t)F_GROUP, span_bytes - g);
      lo = and4(lo, lds128(masks_s + 32u * vb)); hi = and4(hi, lds128(masks_s + 32u * vb + 16u));
    }

    // ---- phase A: newline structure of this warp's chunk ----
    const uint32_t hib = ((lo.x | lo.y) | (lo.z | lo.w)) | ((hi.x | hi.y) | (hi.z | hi.w));
    uint32_t m = nl_mask16_ascii(lo) | (nl_mask16_ascii(hi) << 16);
    if (__any_sync(0xffffffffu, (hib & 0x80808080u) != 0)) {  // the short compare is exact only for bytes < 0x80
      m = nl_mask16(lo) | (nl_mask16(hi) << 16);
      if (lane == 0) sts32(sm0 + FS_OFF(abandon_a) + 4u * par, 1u);
    }
    const uint32_t b1 = __ballot_sync(0xffffffffu, m != 0);
    const uint32_t b2 = __ballot_sync(0xffffffffu, (m & (m - 1u)) != 0);
    const uint32_t Tw = __reduce_add_sync(0xffffffffu, __popc(m));
    const uint32_t hp = bfind(m);
    {
      const uint32_t wl = b1 ? 31u - (uint32_t)__clz(b1) : 0u;  // the lane that holds the chunk's last newline writes the entry
      const uint32_t tail = b1 ? nvalid - (32u * (uint32_t)lane + hp + 1u) : nvalid;
      if ((uint32_t)lane == wl) sts64(ent_s + 8u * (uint32_t)warp, Tw, tail);
      if (Tw != (uint32_t)(__popc(b1) + __popc(b2)) && lane == 0) sts32(sm0 + FS_OFF(abandon_a) + 4u * par, 1u);  // three newlines and more in a group
    }
    if (L_tile == 0 && b1) {  // the span's head fragment may end in this chunk
      const uint32_t pf = (uint32_t)__ffs(b1) - 1u;
      if ((uint32_t)lane == pf) sts32(first_s + 4u * (uint32_t)warp, 32u * pf + (uint32_t)__ffs(m) - 1u);
    }
    __syncthreads();

    // ---- phase B: this warp's start from the entries ----
    const uint2 e = lds64(ent_s + 8u * (uint32_t)lane);
    const uint32_t Ttot = __reduce_add_sync(0xffffffffu, e.x);
    const uint32_t Lw = __reduce_add_sync(0xffffffffu, lane < warp ? e.x : 0u);
    const uint32_t nz = __ballot_sync(0xffffffffu, e.x != 0);
    const uint32_t nzb = nz & ((1u << warp) - 1u);
    const int jl = 31 - __clz(nzb);  // last chunk before this one that has a newline (-1: none)
    uint32_t openw = __reduce_add_sync(0xffffffffu, (lane > jl && lane < warp) ? e.y : 0u);
    openw += nzb ? __shfl_sync(0xffffffffu, e.y, jl & 31) : open_tile;
    const int ja = 31 - __clz(nz);
    const uint32_t rest = __reduce_add_sync(0xffffffffu, lane > ja ? e.y : 0u);
    if (L_tile == 0 && nz) {
      const int jf = __ffs(nz) - 1;
      head_len = open_tile + __reduce_add_sync(0xffffffffu, lane < jf ? e.y : 0u) + lds32(first_s + 4u * (uint32_t)jf);
    }
    const uint32_t open_next = (nz ? __shfl_sync(0xffffffffu, e.y, ja & 31) : open_tile) + rest;
    const bool too_long = (int)open_tile > 0x70000000;
    if (too_long && tid == 0) sts32(sm0 + FS_OFF(abandon_b), 1u);
    abandoned |= lds32(sm0 + FS_OFF(abandon_a) + 4u * par) != 0;  // (this slot is written again two tiles later, after two barriers)
    const bool skip = count_only || too_long || abandoned;

    // the 16-bit halves of the packed per-position table take PT_MAX_LINES quality lines between two flushes
    const uint32_t qneed = (Ttot >> 2) + 2u;
    if (qacc + qneed > (uint32_t)PT_MAX_LINES) {
      if (qcount) { fast_side_pass(sm0, warp, lane, 0u, qcount, core, over); qcount = 0; }
      __syncthreads();
      flush_pos_tab(sm, block, tid);
      __syncthreads();
      qacc = 0;
    }
    qacc += qneed;

    if (!skip) {
      const uint32_t prev = b1 & lt;
      const uint32_t lrel = L_tile + Lw + __popc(prev) + __popc(b2 & lt);  // span-relative line of the group's byte 0
      const uint32_t cls = (lrel + phase) & 3u;
      const uint32_t p = bfind(prev);
      const uint32_t hpv = __shfl_sync(0xffffffffu, hp, p & 31u);
      // 16 + line position of byte 0
      const uint32_t q = prev ? 32u * ((uint32_t)lane - p) + 15u - hpv : openw + 32u * (uint32_t)lane + 16u;
      const bool odd = (cls & 1u) != 0, head = lrel == 0;  // line 0 of the span is the head fragment (stitch kernel)
      const bool c1 = m != 0, c2 = (m & (m - 1u)) != 0;
      const uint32_t k1 = c1 ? (uint32_t)__ffs(m) - 1u : 32u;
      // the one counted segment: bytes before the first newline (class of byte 0 is sequence / quality), else after it
      uint32_t mi = odd ? k1 : (uint32_t)M_ABOVE + k1;
      if (odd && head) mi = (uint32_t)M_NONE;
      uint4 ml = lds128(masks_s + 32u * mi), mh = lds128(masks_s + 32u * mi + 16u);
      if (partial) { ml = and4(ml, lds128(masks_s + 32u * vb)); mh = and4(mh, lds128(masks_s + 32u * vb + 16u)); }
      const uint32_t tb = hb_lane + ((cls & 2u) << 13);  // classes 2 (-> 3) and 3: the quality table
      hist16(ksel, sel_nl4(lo, ml), tb);
      hist16(ksel, sel_nl4(hi, mh), tb);
      if (!core && ((cls == 3u && !head) || (cls == 2u && c1 && !c2))) {
        const uint32_t lineq = lrel + (odd ? 0u : 1u);
        const uint32_t pt = ptab_s + ((lineq & 4u) ? PT_COPY * 4u : 0u);
        const uint32_t qe = odd ? q : 15u - k1;
        if (odd || k1 < 16u) fast_pos16(sm0, and4(lo, ml), pt, qe, over);
        fast_pos16(sm0, and4(hi, mh), pt, qe + 16u, over);
      }
      if (c1 && odd && !head) {  // a sequence / quality line ends at k1
        const uint32_t len = q + k1 - 16u;
        const uint32_t bin = len < (uint32_t)POS_BINS ? len : (uint32_t)POS_BINS;
        if (cls == 3u) {
          if (!core) { red_inc(sm0 + FS_OFF(qual_len) + 4u * bin); mnq = min(mnq, len); mxq = max(mxq, len); }
        } else {
          red_inc(sm0 + FS_OFF(seq_len) + 4u * bin); red_inc(sm0 + FS_OFF(seq_log2) + 4u * (32u - (uint32_t)__clz(len)));
          mns = min(mns, len); mxs = max(mxs, len);
        }
      }
      if (c2 && !odd) sts32(sm0 + FS_OFF(abandon_b), 1u);  // a sequence / quality line between two newlines of one group
      const bool push = c2 && odd && !(core && cls == 1u);
      const uint32_t pm = __ballot_sync(0xffffffffu, push);
      if (pm) {
        if (push) {
          const uint32_t slot = (uint32_t)warp * F_QCAP + qcount + __popc(pm & lt);
          sts128(sm0 + FS_OFF(qdata) + 32u * slot, lo); sts128(sm0 + FS_OFF(qdata) + 32u * slot + 16u, hi);
          sts32(sm0 + FS_OFF(qmeta) + 4u * slot, hp | (cls == 1u ? 32u : 0u) | (((lrel + 2u) & 4u) ? 64u : 0u) | (vb << 8));
        }
        qcount += __popc(pm);
        __syncwarp();
        if (qcount >= 32u) { qcount -= 32u; fast_side_pass(sm0, warp, lane, qcount, 32u, core, over); }
      }
    }
    L_tile += Ttot;
    open_tile = open_next;
    par ^= 1u;
    ent_s = sm0 + FS_OFF(entry) + par * (uint32_t)sizeof(sm.entry[0]);
    first_s = sm0 + FS_OFF(first) + par * (uint32_t)sizeof(sm.first[0]);
  }

  if (qcount) fast_side_pass(sm0, warp, lane, 0u, qcount, core, over);
  if (mns != ~0u) { atomicMin(&sm.len_min[0], (u64)mns); atomicMax(&sm.len_max[0], (u64)mxs); }
  if (mnq != ~0u) { atomicMin(&sm.len_min[1], (u64)mnq); atomicMax(&sm.len_max[1], (u64)mxq); }
  if (over) atomicAdd(&sm.pos_over, over);
  __syncthreads();
  if (tid == 0) {
    desc.T = L_tile;
    desc.head_len = L_tile ? head_len : open_tile;
    desc.tail_len = open_tile;
  }
  if (count_only) return;
  for (int bin = tid; bin < 2 * HB; bin += F_THREADS) {  // fold the 32 lane copies (rotated: no bank conflicts)
    const int h = bin / HB, b = bin % HB;
    u64 s = 0;
    const uint32_t* hp32 = &sm.hist[h][b << 5];
#pragma unroll 8
    for (int l = 0; l < 32; l++) s += hp32[(l + bin) & 31];
    if (b == '\n') s = 0;                                   // masked-out bytes
    if (b == '\r' && s && !(core && h == 1)) sm.abandon_b = 1;  // a '\r' inside a sequence / quality line: the exact pass decides
    if (s && !(core && h == 1)) block[OFF_HIST_SEQ + 256 * h + b] += s;
  }
  if (!core) {
    for (int i = tid; i <= POS_BINS; i += F_THREADS) if (sm.qual_len[i]) block[OFF_QUAL_LEN + i] += sm.qual_len[i];
    flush_pos_tab(sm, block, tid);
  }
  for (int i = tid; i <= POS_BINS; i += F_THREADS) if (sm.seq_len[i]) block[OFF_SEQ_LEN + i] += sm.seq_len[i];
  if (tid < LOG2_BINS && sm.seq_log2[tid]) block[OFF_SEQ_LOG2 + tid] += sm.seq_log2[tid];
  __syncthreads();
  if (!core) flush_gpos(sm, block, tid);
  if (tid == 0) {
    if (sm.pos_over && !core) block[OFF_POS_SUM + POS_BINS] += sm.pos_over;
    if (sm.len_min[0] < block[OFF_SEQ_LEN_MIN]) block[OFF_SEQ_LEN_MIN] = sm.len_min[0];
    if (sm.len_max[0] > block[OFF_SEQ_LEN_MAX]) block[OFF_SEQ_LEN_MAX] = sm.len_max[0];
    if (sm.len_min[1] < block[OFF_QUAL_LEN_MIN]) block[OFF_QUAL_LEN_MIN] = sm.len_min[1];
    if (sm.len_max[1] > block[OFF_QUAL_LEN_MAX]) block[OFF_QUAL_LEN_MAX] = sm.len_max[1];
    if (abandoned || sm.abandon_b) desc.pad = 1;
  }
}
