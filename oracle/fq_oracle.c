/*
 * fq_oracle.c -- CPU restatement of the reference's FASTQ scanning path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (libfqgpu, the `sc` mirror CLI, the Python
 * binding) links, imports or executes this file; only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may.
 *
 * What is restated (reference = /root/reference, danielecook/seq-collection):
 *   src/fq_count.nim:22-45   state + `for line in lines(stream)` loop (i mod 4 classing,
 *                            count("G")+count("C"), count("N"), line.len)
 *   src/fq_count.nim:47-52   output order and `$` formatting (fqo_format_fq_count_row)
 *   src/fq_meta.nim:10,94-102  qual table, qual_to_int, qual_min_max (prev_min >= 0 rule)
 *   src/fq_meta.nim:226-227,245-248,277  sampling loop bounds, which line, n_lines
 *
 * Arithmetic that lives in un-vendored dependencies (absent from /root/reference) and is restated
 * from the published behaviour of the pinned versions (Nim 1.0.6, .github/workflows/build.yml:45):
 *   streams.lines / io.readLine(File): split at '\n', drop ONE '\r' directly before it, yield a
 *       trailing unterminated line only if non-empty, yield blank lines.
 *   strutils.count(s, sub) for 1-char subs == byte count, case-sensitive.
 *   `$`(float) == "%.16g" plus ".0" when the text has no '.', letter; NaN prints "nan".
 * Pinned by: the 15 golden rows of docs/fq-count.md:29-43 (11 fixtures have no trailing newline, so
 * "unterminated last line counts" is pinned; dup.fq.gz pins the gz path) and the 4 min/max rows of
 * docs/fq-meta.md:34-37 -- see tests/test_oracle_golden.py.
 * PARITY UNPINNED (no reference fixture contains them; behaviour rests on reading the sources):
 *   CRLF stripping, lone '\r' and NUL bytes (treated here as ordinary bytes, which is what
 *   io.readLine(File) does for the plain-file path of fq_count), the exact `$float` string,
 *   an empty quality line with no history (reference raises; we report FQGPU_META_EMPTY_QUAL),
 *   and the whole extended counter set (base/qual histograms, length tables, per-position sums),
 *   which the reference does not compute at all and which this file DEFINES.
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include "../include/fqgpu.h"

typedef struct fqo_state {
  fqgpu_stats st;
  uint64_t i;            /* 1-based line counter, src/fq_count.nim:23,39 */
  uint64_t meta_records; /* sample_n, src/fq_meta.nim:197 */
  int64_t qual_min, qual_max; /* src/fq_meta.nim:207-208 */
  uint32_t meta_status;
  uint8_t* cur;          /* bytes of the current, not yet terminated line */
  size_t cur_len, cur_cap;
} fqo_state;

/* src/fq_meta.nim:10,94-95: index of the char in "!".."~", -1 when absent */
static inline int qual_to_int(uint8_t c) { return (c >= 33 && c <= 126) ? (int)c - 33 : -1; }

/* src/fq_meta.nim:97-102 */
static void qual_min_max(fqo_state* s, const uint8_t* line, size_t len) {
  int64_t mn = 0, mx = 0;
  int have = 0;
  for (size_t k = 0; k < len; k++) {
    int q = qual_to_int(line[k]);
    if (!have) { mn = mx = q; have = 1; }
    else { if (q < mn) mn = q; if (q > mx) mx = q; }
  }
  if (s->qual_min >= 0) { /* :100-101 concat(@[prev_min, prev_max]) */
    if (!have) { mn = s->qual_min; mx = s->qual_max; have = 1; }
    else {
      if (s->qual_min < mn) mn = s->qual_min;
      if (s->qual_min > mx) mx = s->qual_min;
      if (s->qual_max < mn) mn = s->qual_max;
      if (s->qual_max > mx) mx = s->qual_max;
    }
  }
  if (!have) { /* min() of an empty seq: the reference raises (unpinned) */
    s->meta_status = FQGPU_META_EMPTY_QUAL;
    return;
  }
  s->qual_min = mn; /* :102 */
  s->qual_max = mx;
}

static inline unsigned log2_bin(uint64_t len) {
  unsigned b = 0;
  while (len) { b++; len >>= 1; }
  return b; /* 0 -> 0, 1 -> 1, 2..3 -> 2, 4..7 -> 3 ... */
}

static void process_line(fqo_state* s, const uint8_t* line, size_t len) {
  fqgpu_stats* st = &s->st;
  s->i++; /* src/fq_count.nim:39 */
  uint64_t i = s->i;
  if ((i % 4) == 1) st->reads++; /* :40-41 */
  if ((i % 4) == 2) {            /* :42-45 */
    uint64_t g = 0, c = 0, n = 0;
    for (size_t k = 0; k < len; k++) { /* three count() passes in the reference; same totals */
      uint8_t b = line[k];
      g += (b == 'G'); c += (b == 'C'); n += (b == 'N');
      st->base_counts[b]++;
    }
    st->gc_bases += g + c;
    st->n_bases += n;
    st->bases += len;
    st->seq_lines++;
    if (len < st->seq_len_min) st->seq_len_min = len;
    if (len > st->seq_len_max) st->seq_len_max = len;
    st->seq_len_hist[len < FQGPU_POS_BINS ? len : FQGPU_POS_BINS]++;
    st->seq_len_log2[log2_bin(len)]++;
  }
  if ((i % 4) == 0) { /* quality line (0-based i %% 4 == 3, src/fq_meta.nim:245) */
    st->qual_lines++;
    for (size_t k = 0; k < len; k++) {
      uint8_t b = line[k];
      st->qual_counts[b]++;
      size_t p = k < FQGPU_POS_BINS ? k : FQGPU_POS_BINS;
      st->qual_pos_sum[p] += b;
      st->qual_pos_cnt[p]++;
    }
    if (len < st->qual_len_min) st->qual_len_min = len;
    if (len > st->qual_len_max) st->qual_len_max = len;
    st->qual_len_hist[len < FQGPU_POS_BINS ? len : FQGPU_POS_BINS]++;
  }
  /* fq-meta loop: `while not atEnd and i < sample_n*4` with 0-based i = s->i - 1, :226,245-248 */
  if (s->meta_records && (i - 1) < s->meta_records * 4) {
    st->meta_lines = i;
    if (((i - 1) % 4) == 3 && s->meta_status == FQGPU_META_OK) qual_min_max(s, line, len);
  }
}

static void cur_append(fqo_state* s, const uint8_t* p, size_t n) {
  if (s->cur_len + n > s->cur_cap) {
    size_t cap = s->cur_cap ? s->cur_cap : 256;
    while (cap < s->cur_len + n) cap *= 2;
    s->cur = (uint8_t*)realloc(s->cur, cap);
    s->cur_cap = cap;
  }
  memcpy(s->cur + s->cur_len, p, n);
  s->cur_len += n;
}

fqo_state* fqo_new(uint64_t meta_records) {
  fqo_state* s = (fqo_state*)calloc(1, sizeof(fqo_state));
  s->meta_records = meta_records;
  s->qual_min = -1;
  s->qual_max = -1;
  s->st.seq_len_min = UINT64_MAX;
  s->st.qual_len_min = UINT64_MAX;
  return s;
}

void fqo_free(fqo_state* s) {
  if (!s) return;
  free(s->cur);
  free(s);
}

/* Nim streams.lines over an arbitrary chunking of the byte stream. */
void fqo_feed(fqo_state* s, const uint8_t* buf, size_t n) {
  s->st.bytes += n;
  size_t pos = 0;
  while (pos < n) {
    const uint8_t* nl = (const uint8_t*)memchr(buf + pos, '\n', n - pos);
    if (!nl) { cur_append(s, buf + pos, n - pos); return; }
    size_t seg = (size_t)(nl - (buf + pos));
    const uint8_t* line;
    size_t len;
    if (s->cur_len) {
      cur_append(s, buf + pos, seg);
      line = s->cur; len = s->cur_len;
    } else {
      line = buf + pos; len = seg;
    }
    if (len > 0 && line[len - 1] == '\r') len--; /* one CR directly before the LF */
    process_line(s, line, len);
    s->cur_len = 0;
    pos += seg + 1;
  }
}

void fqo_finish(fqo_state* s, fqgpu_stats* out) {
  if (s->cur_len) { /* trailing unterminated, non-empty line is yielded as is */
    process_line(s, s->cur, s->cur_len);
    s->cur_len = 0;
  }
  s->st.lines = s->i;
  s->st.meta_qual_min = s->qual_min;
  s->st.meta_qual_max = s->qual_max;
  s->st.meta_status = s->meta_status;
  *out = s->st;
}

void fqo_count_buffer(const uint8_t* buf, size_t n, uint64_t meta_records, fqgpu_stats* out) {
  fqo_state* s = fqo_new(meta_records);
  fqo_feed(s, buf, n);
  fqo_finish(s, out);
  fqo_free(s);
}

/* Same, but feeding `chunk`-byte pieces (exercises the straddling-line code above). */
void fqo_count_buffer_chunked(const uint8_t* buf, size_t n, size_t chunk, uint64_t meta_records,
                              fqgpu_stats* out) {
  fqo_state* s = fqo_new(meta_records);
  if (chunk == 0) chunk = n ? n : 1;
  for (size_t off = 0; off < n; off += chunk) fqo_feed(s, buf + off, (n - off < chunk) ? n - off : chunk);
  fqo_finish(s, out);
  fqo_free(s);
}

/* src/fq_count.nim:30-36: ".gz" (case-sensitive, last three chars) selects the gz stream. zlib's
 * gzread is the mechanism under zip/gzipfiles (sc.nimble:10) and gzip_stream.nim:16-17.
 * Returns 0, or -1 when the file cannot be opened (reference: quit_error(..., 2)). */
int fqo_count_file(const char* path, uint64_t meta_records, fqgpu_stats* out) {
  size_t L = strlen(path);
  int gz = (L >= 3 && strcmp(path + L - 3, ".gz") == 0);
  fqo_state* s = fqo_new(meta_records);
  uint8_t* buf = (uint8_t*)malloc(1 << 20);
  if (gz) {
    gzFile f = gzopen(path, "rb");
    if (!f) { free(buf); fqo_free(s); return -1; }
    int r;
    while ((r = gzread(f, buf, 1 << 20)) > 0) fqo_feed(s, buf, (size_t)r);
    gzclose(f);
  } else {
    FILE* f = fopen(path, "rb");
    if (!f) { free(buf); fqo_free(s); return -1; }
    size_t r;
    while ((r = fread(buf, 1, 1 << 20, f)) > 0) fqo_feed(s, buf, r);
    fclose(f);
  }
  fqo_finish(s, out);
  free(buf);
  fqo_free(s);
  return 0;
}

/* Nim 1.0.6 `$`(float): system/formatfloat.nim writeFloatToBuffer -- "%.16g", append ".0" when no
 * '.', ',' or letter is present, NaN -> "nan", infinities -> "inf"/"-inf". */
int fqo_format_float(double v, char* buf, size_t cap) {
  char tmp[80];
  int n = snprintf(tmp, sizeof tmp, "%.16g", v);
  int has_dot = 0;
  for (int k = 0; k < n; k++) {
    if (tmp[k] == ',') { tmp[k] = '.'; has_dot = 1; }
    else if ((tmp[k] >= 'a' && tmp[k] <= 'z') || (tmp[k] >= 'A' && tmp[k] <= 'Z') || tmp[k] == '.') has_dot = 1;
  }
  if (!has_dot) { tmp[n++] = '.'; tmp[n++] = '0'; tmp[n] = 0; }
  if (tmp[n - 1] == 'n' || tmp[n - 1] == 'N') { strcpy(tmp, "nan"); n = 3; }
  else if (tmp[n - 1] == 'f' || tmp[n - 1] == 'F') { strcpy(tmp, tmp[0] == '-' ? "-inf" : "inf"); n = (int)strlen(tmp); }
  if ((size_t)n + 1 > cap) return -1;
  memcpy(buf, tmp, (size_t)n + 1);
  return n;
}

/* src/fq_count.nim:47-51: [$n_reads, $(gc/(total_len-n_cnt)), $gc_cnt, $n_cnt, $total_len].join("\t") */
int fqo_format_fq_count_row(const fqgpu_stats* st, char* buf, size_t cap) {
  char f[80];
  double gc_content = (double)(int64_t)st->gc_bases / (double)((int64_t)st->bases - (int64_t)st->n_bases);
  fqo_format_float(gc_content, f, sizeof f);
  int n = snprintf(buf, cap, "%lld\t%s\t%lld\t%lld\t%lld", (long long)st->reads, f, (long long)st->gc_bases,
                   (long long)st->n_bases, (long long)st->bases);
  return (n < 0 || (size_t)n >= cap) ? -1 : n;
}

/* ------------------------------------------------------------------------------------------
 * CPU-baseline leg: the reference's own WORK SHAPE for fq-count, one core, as in
 * src/fq_count.nim:38-45 over a plain file: io.readLine(File) = fgets into a '\n'-prefilled,
 * growing buffer + memchr for the terminator, then THREE separate strutils.count passes
 * (each a byte loop over the line) plus len.  Only the five fq-count outputs are produced.
 * out5 = {reads, gc_bases, n_bases, bases, lines}.  Returns 0 / -1.
 * ------------------------------------------------------------------------------------------ */
static size_t count_char(const char* s, size_t len, char c) { /* strutils.count(s, sub) for 1-char sub */
  size_t k = 0;
  for (size_t i = 0; i < len; i++) k += (s[i] == c);
  return k;
}

static int ref_readline(FILE* f, char** line, size_t* cap, size_t* out_len) {
  size_t pos = 0, sp = 80;
  if (*cap < sp) { *line = (char*)realloc(*line, sp); *cap = sp; }
  for (;;) {
    if (*cap < pos + sp) { *line = (char*)realloc(*line, pos + sp); *cap = pos + sp; }
    memset(*line + pos, '\n', sp);
    int ok = fgets(*line + pos, (int)sp, f) != NULL;
    char* m = (char*)memchr(*line + pos, '\n', sp);
    if (m) {
      size_t last = (size_t)(m - *line);
      if (last > 0 && (*line)[last - 1] == '\r') { *out_len = last - 1; return last > 1 || ok; }
      if (last > 0 && (*line)[last - 1] == '\0') {
        if (last < pos + sp - 1 && (*line)[last + 1] != '\0') last--;
      }
      *out_len = last;
      return last > 0 || ok;
    }
    sp--; /* fgets wrote a NUL at the end */
    pos += sp;
    sp = 128;
  }
}

int fqo_ref_fq_count_file(const char* path, uint64_t* out5) {
  FILE* f = fopen(path, "rb");
  if (!f) return -1;
  char* line = NULL;
  size_t cap = 0, len = 0;
  uint64_t i = 0, reads = 0, gc = 0, n = 0, total = 0;
  while (ref_readline(f, &line, &cap, &len)) {
    i++;
    if ((i % 4) == 1) reads++;
    if ((i % 4) == 2) {
      gc += count_char(line, len, 'G') + count_char(line, len, 'C');
      n += count_char(line, len, 'N');
      total += len;
    }
  }
  fclose(f);
  free(line);
  out5[0] = reads; out5[1] = gc; out5[2] = n; out5[3] = total; out5[4] = i;
  return 0;
}

/* Same work shape over a memory-resident byte range (memchr finds the terminator as in readLine;
 * the line is copied into a private string as `lines` does, then the three passes run). */
void fqo_ref_fq_count_mem(const uint8_t* buf, size_t nbytes, uint64_t* out5) {
  char* line = NULL;
  size_t cap = 0;
  uint64_t i = 0, reads = 0, gc = 0, n = 0, total = 0;
  size_t pos = 0;
  while (pos < nbytes) {
    const uint8_t* nl = (const uint8_t*)memchr(buf + pos, '\n', nbytes - pos);
    size_t len = nl ? (size_t)(nl - (buf + pos)) : nbytes - pos;
    size_t adv = nl ? len + 1 : len;
    if (len + 1 > cap) { cap = (len + 1) * 2; line = (char*)realloc(line, cap); }
    memcpy(line, buf + pos, len);
    if (nl && len > 0 && line[len - 1] == '\r') len--;
    pos += adv;
    i++;
    if ((i % 4) == 1) reads++;
    if ((i % 4) == 2) {
      gc += count_char(line, len, 'G') + count_char(line, len, 'C');
      n += count_char(line, len, 'N');
      total += len;
    }
  }
  free(line);
  out5[0] = reads; out5[1] = gc; out5[2] = n; out5[3] = total; out5[4] = i;
}

size_t fqo_stats_size(void) { return sizeof(fqgpu_stats); }
