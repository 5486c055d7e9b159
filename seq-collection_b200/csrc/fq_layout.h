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

constexpr int MAX_SPANS = 296;  // 2 CTAs per SM on a 148-SM B200

// ---- stream carry: device-resident state that makes consecutive scans one logical stream ----
struct Carry {
  unsigned long long lines;     // '\n' seen so far == terminated lines
  unsigned long long open_len;  // raw bytes since the last '\n' (the open line)
  unsigned long long bytes;     // bytes scanned so far
  unsigned int last_byte;       // last byte of the stream so far (0 when bytes == 0)
  unsigned int flags;
  // fq-meta fold (src/fq_meta.nim:207-208,226-248)
  unsigned long long meta_lines;  // lines consumed by the sampling loop (<= 4*meta_records)
  long long qual_min, qual_max;   // running qual_min/qual_max, -1 initially
  unsigned int meta_status;
  unsigned int meta_pending_cr;   // stream so far ends with a '\r' not yet attributed
  int cur_has;                    // open line: any attributed byte so far
  int cur_min, cur_max;           // open line: min/max of qual_to_int so far
  unsigned int pad;
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
  unsigned int last_byte0, phase_known;
  unsigned int mismatches, pad;
};

}  // namespace fq
