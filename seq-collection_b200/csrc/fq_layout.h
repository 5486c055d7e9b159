// fq_layout.h -- constants shared by the device kernels and the host side of libfqgpu.
// Counter blocks are arrays of uint64 words; every span (one persistent CTA per span and launch)
// owns one block, the reduction kernel (K3) folds them into one block of the same layout.
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

constexpr int BASE_SPANS = 296;             // CTAs resident at once: 2 per SM on a 148-SM B200
constexpr int SPAN_WAVES = 8;               // large launches are cut into up to this many spans per resident CTA (see launch_scan)
constexpr int MAX_SPANS = BASE_SPANS * SPAN_WAVES;

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
enum { CARRY_UNKNOWN_START = 1u, CARRY_HYP_VALID = 2u, CARRY_HYP_SHIFT = 8 };

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

// One span = the contiguous run of tiles one CTA scans in a launch.  The line phase at a span start
// is GUESSED from the content (first '@' line whose line+2 starts with '+') so that no CTA ever
// waits for another; the stitch kernel verifies every guess against the exact line counts.
enum { SPAN_PENDING = 0, SPAN_COMMITTED = 1, SPAN_RESCAN = 2 };
constexpr unsigned int PHASE_UNKNOWN = 4;
struct SpanDesc {
  unsigned long long T;         // newlines in the span
  unsigned long long head_len;  // bytes before the first newline (span length when T == 0)
  unsigned long long tail_len;  // bytes after the last newline (T > 0)
  unsigned long long G, P0;     // exact lines / open-line bytes before the span (stitch kernel)
  unsigned int guess;           // (lines before the span) mod 4 used by pass 0; PHASE_UNKNOWN = count only
  unsigned int state;           // SPAN_*
  unsigned int exact;           // exact phase (stitch kernel)
  unsigned int pad;
  unsigned long long pad2[2];
};

// Snapshot of the carry taken at the start of a launch (read by every stitch CTA while the last
// one writes the new carry).
struct LaunchHdr {
  unsigned long long lines0, open0, bytes0;
  unsigned int last_byte0, flags0;   // flags0 = Carry.flags at the start of the launch
  unsigned int mismatches, pad;
};

}  // namespace fq
