// fq_layout.h -- constants shared by the device kernels and the host side of libfqgpu.
// The counter block is an array of uint64 words; every persistent CTA of a scan launch adds its shared-memory
// tables into the context's one block (K3 is this reduction by 64-bit atomics at CTA exit).
#pragma once
#include <stdint.h>

#include "../../include/fqgpu.h"

namespace fq {

constexpr int POS_BINS = FQGPU_POS_BINS;     // 512 (+1 overflow bin)
constexpr int LOG2_BINS = FQGPU_LEN_LOG2_BINS;

// ---- counter block layout (uint64 words) -------------------------------------------------
constexpr int OFF_HIST_SEQ = 0;                          // [256] byte histogram, sequence lines
constexpr int OFF_HIST_QUAL = OFF_HIST_SEQ + 256;        // [256] byte histogram, quality lines
constexpr int OFF_SEQ_LEN = OFF_HIST_QUAL + 256;         // [POS_BINS+1] exact length histogram
constexpr int OFF_QUAL_LEN = OFF_SEQ_LEN + POS_BINS + 1; // [POS_BINS+1]
constexpr int OFF_SEQ_LOG2 = OFF_QUAL_LEN + POS_BINS + 1;  // [64]
constexpr int OFF_POS_SUM = OFF_SEQ_LOG2 + LOG2_BINS;      // [POS_BINS+1] per-position quality byte sums
constexpr int OFF_SEQ_LEN_MIN = OFF_POS_SUM + POS_BINS + 1;  // min-reduced (UINT64_MAX when none)
constexpr int OFF_SEQ_LEN_MAX = OFF_SEQ_LEN_MIN + 1;         // max-reduced
constexpr int OFF_QUAL_LEN_MIN = OFF_SEQ_LEN_MAX + 1;
constexpr int OFF_QUAL_LEN_MAX = OFF_QUAL_LEN_MIN + 1;
constexpr int BLOCK_WORDS = ((OFF_QUAL_LEN_MAX + 1 + 31) / 32) * 32;
constexpr int N_SUM_WORDS = OFF_SEQ_LEN_MIN;  // words [0, N_SUM_WORDS) are sum-reduced

constexpr int RESIDENT_CTAS = 592;         // most CTAs resident at once (up to 4 per SM on a 148-SM B200)
constexpr int SPAN_WAVES = 8;              // large launches are cut into up to this many spans per resident CTA
constexpr int MAX_SPANS = RESIDENT_CTAS * SPAN_WAVES;

// ---- launch control words (device, u64[CTL_WORDS]); zeroed by the reset kernel, left clean by every launch --------
enum {
  CTL_DONE = 0,     // CTAs of the launch that have finished
  CTL_REDO,         // spans whose guessed line phase was wrong (malformed input): the second pass redoes them
  CTL_FIRST_NL,     // shards with an unknown start: offset of the stream's first newline + 1 (0 = none in this launch)
  CTL_HEAD,         // shards: HEAD_* flags of this launch
  CTL_HEAD_P0,      // shards: bytes of the detached head before this launch
  CTL_ERROR,        // sticky: an internal consistency check failed
  CTL_WORDS = 16
};
enum { HEAD_ACTIVE = 1u, HEAD_QUAL = 2u, HEAD_PENDING_CR = 4u };

// One span = the run of whole lines one CTA scans in a launch: the lines that START inside its nominal byte range
// (span 0 also owns the line that is open at the launch start).  The line phase at a span start is GUESSED from the
// content (first '@' line whose line+2 starts with '+') so that no CTA ever waits for another; the CTA that exits
// last verifies every guess against the exact line counts.
enum { SPAN_OK = 0, SPAN_REDO = 1 };
struct SpanDesc {
  unsigned long long T;            // newlines of the span
  unsigned long long start;        // offset of its first byte (relative to the launch base)
  unsigned long long end_open;     // open-line bytes where the span stopped (meaningful when reached_end)
  unsigned long long len_min[2], len_max[2];  // line-length extrema of the span ([0] seq, [1] qual), committed once the guess is verified
  unsigned int guess;              // (lines before the span) mod 4 used by pass 0
  unsigned int guess_valid;        // 0: the content gave no guess (the span is redone with the exact phase)
  unsigned int nonempty;           // a line starts inside the span
  unsigned int reached_end;        // the span ran to the end of the launch
  unsigned int exact;              // exact (lines before the span) mod 4 (written by the last CTA)
  unsigned int state;              // SPAN_*
  unsigned int pad[2];
};

// ---- stream carry: device-resident state that makes consecutive scans one logical stream ----
struct Carry {
  unsigned long long lines;     // '\n' seen so far == terminated lines
  unsigned long long open_len;  // raw bytes since the last '\n' (the open line)
  unsigned long long bytes;     // bytes scanned so far
  unsigned int last_byte;       // last byte of the stream so far (0 when bytes == 0)
  unsigned int flags;           // CARRY_* (multi-GPU shards)
  // fq-meta fold (src/fq_meta.nim:207-208,226-248)
  unsigned long long meta_lines;  // lines consumed by the sampling loop (<= 4*meta_records)
  long long qual_min, qual_max;   // running qual_min/qual_max, -1 initially
  unsigned int meta_status;
  unsigned int meta_pending_cr;   // stream so far ends with a '\r' not yet attributed
  int cur_has;                    // open line: any attributed byte so far
  int cur_min, cur_max;           // open line: min/max of qual_to_int so far
  unsigned int pad;
};

// Multi-GPU shards: a rank > 0 does not know the line phase of its first byte.
//   CARRY_UNKNOWN_START  the shard started with an unknown phase; `lines` counts from the shard start
//   CARRY_HYP_VALID      flags >> 8 & 3 is the phase hypothesis (resynced from the content) the shard
//                        was scanned with; verified by fqgpu_shard_combine against the exact counts
//   CARRY_HYP_FAILED     no record start was found where the hypothesis was needed: the shard's block is not usable
//                        (fqgpu_shard_combine reports FQGPU_ERETRY and the rank is rescanned with the exact carry)
enum { CARRY_UNKNOWN_START = 1u, CARRY_HYP_VALID = 2u, CARRY_HYP_FAILED = 4u, CARRY_HYP_SHIFT = 8 };

// What a shard exports besides its counter block (one slot of the all-reduced buffer).
constexpr int SH_OFF_HEAD_POS = 0;                         // [POS_BINS+1] per-position sums of the detached head fragment
constexpr int SH_OFF_SCALARS = SH_OFF_HEAD_POS + POS_BINS + 1;
enum {
  SH_LINES = 0, SH_BYTES, SH_OPEN_LEN, SH_LAST_BYTE, SH_FIRST_BYTE, SH_HEAD_LEN, SH_HEAD_CR, SH_HYP, SH_HYP_VALID,
  SH_EXACT, SH_META_LINES, SH_META_QMIN, SH_META_QMAX, SH_META_STATUS, SH_META_PENDING_CR, SH_META_CUR_HAS,
  SH_META_CUR_MIN, SH_META_CUR_MAX, SH_PRESENT, SH_NSCALARS
};
constexpr int SHARD_EXTRA_WORDS = ((SH_OFF_SCALARS + SH_NSCALARS + 31) / 32) * 32;

// Device-resident description of the shard's detached head (first line fragment of the shard).
struct ShardInfo {
  unsigned long long head_pos[POS_BINS + 1];  // per-position sums relative to the shard start
  unsigned long long head_len;                // bytes before the shard's first newline
  unsigned int head_cr;                       // the byte before that newline is '\r' (and inside the shard)
  unsigned int first_byte;                    // first byte of the shard
  unsigned int pad[2];
};

}  // namespace fq
