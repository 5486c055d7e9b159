// sc_main.cpp -- host-side mirror of the reference CLI for the two commands on the accelerated path:
//   sc fq-count [-t|--header] [-b|--basename] [-a|--absolute] [fastq ...]        sc.nim:103-116
//   sc fq-meta  [-n N|--lines N] [-t] [-b] [-a] [fastq ...]                      sc.nim:67-79
// The reference is a Nim binary; no Nim toolchain exists in this image, so the host side above the C ABI
// is restated in C++ with the same flags, output columns, error texts and exit codes.  The per-line
// loop of src/fq_count.nim:38-45 and the quality fold of src/fq_meta.nim:245-246 run on the GPU through
// libfqgpu (include/fqgpu.h); everything else here is the reference's unchanged host logic:
//   output formatting            src/fq_count.nim:47-53, src/fq_meta.nim:255-278, src/utils/helpers.nim:200-224
//   error convention             src/utils/helpers.nim:29-34 (red "Error N: msg" on stderr, exit N)
//   header parsing / sequencer   src/fq_meta.nim:41-92,104-195,229-242,251-253 (string/regex work on <= n headers)
#include <fqgpu.h>
#include <limits.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <map>
#include <regex>
#include <string>
#include <vector>

using std::string;
using std::vector;

static const char* kVersion = "0.0.2";  // sc.nim:48

// helpers.nim:29-34
[[noreturn]] static void quit_error(const string& msg, int code = 1) {
  fprintf(stderr, "\x1b[31mError %d: %s\x1b[0m\n", code, msg.c_str());
  exit(code);
}

static string join(const vector<string>& v, const char* sep) {
  string out;
  for (size_t i = 0; i < v.size(); i++) { if (i) out += sep; out += v[i]; }
  return out;
}

// helpers.nim:200-208
static string output_header(const string& header, bool basename, bool absolute) {
  vector<string> c;
  for (const string& x : {header, string(basename ? "basename" : ""), string(absolute ? "absolute" : "")}) if (!x.empty()) c.push_back(x);
  return join(c, "\t");
}

static string last_path_part(const string& p) {  // os.lastPathPart
  string s = p;
  while (s.size() > 1 && s.back() == '/') s.pop_back();
  size_t k = s.find_last_of('/');
  return k == string::npos ? s : s.substr(k + 1);
}

static string absolute_path(const string& p) {  // os.absolutePath (no symlink resolution, no normalisation)
  if (!p.empty() && p[0] == '/') return p;
  char cwd[PATH_MAX];
  if (!getcwd(cwd, sizeof cwd)) return p;
  return string(cwd) + "/" + p;
}

// helpers.nim:210-224
static string output_w_fnames(const string& row, const string& path, bool basename, bool absolute) {
  string b = basename ? last_path_part(path) : "";
  string a;
  if (absolute) {
    struct stat st;
    char buf[PATH_MAX];
    ssize_t n;
    if (lstat(path.c_str(), &st) == 0 && S_ISLNK(st.st_mode) && (n = readlink(path.c_str(), buf, sizeof buf - 1)) > 0) {
      buf[n] = 0;
      a = absolute_path(buf);
    } else {
      a = absolute_path(path);
    }
  }
  vector<string> c;
  for (const string& x : {row, b, a}) if (!x.empty()) c.push_back(x);
  return join(c, "\t");
}

// Nim 1.0.6 `$`(float): "%.16g", ".0" appended when no '.', ',' or letter is present; NaN -> "nan".
static string nim_float(double v) {
  if (std::isnan(v)) return "nan";
  if (std::isinf(v)) return v > 0 ? "inf" : "-inf";
  char buf[64];
  snprintf(buf, sizeof buf, "%.16g", v);
  bool dot = false;
  for (char* p = buf; *p; p++) if (*p == '.' || isalpha((unsigned char)*p)) dot = true;
  string s(buf);
  if (!dot) s += ".0";
  return s;
}

static fqgpu_ctx* g_ctx = nullptr;

static fqgpu_ctx* context(uint64_t meta_records, uint32_t flags = 0) {
  if (g_ctx) { fqgpu_destroy(g_ctx); g_ctx = nullptr; }
  fqgpu_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.device = -1;
  cfg.meta_records = meta_records;
  cfg.flags = flags;
  if (const char* e = getenv("FQGPU_CHUNK_MB")) cfg.chunk_bytes = (size_t)atol(e) << 20;
  if (fqgpu_create(&g_ctx, &cfg) != FQGPU_OK) quit_error(string("GPU unavailable: ") + fqgpu_last_error(nullptr), 1);
  return g_ctx;
}

// ------------------------------------------------------------------------------------------------
// fq-count  (src/fq_count.nim:14-53)
// ------------------------------------------------------------------------------------------------
static const char* kFqCountHeader = "reads\tgc_content\tgc_bases\tn_bases\tbases";

static void fq_count_row(const string& fastq, int rc, const fqgpu_stats& st, const char* err, bool basename, bool absolute) {
  if (rc == FQGPU_EIO) quit_error("Unable to open file: " + fastq, 2);  // fq_count.nim:35-36
  if (rc != FQGPU_OK) quit_error(err, 1);
  const double gc_content = (double)(int64_t)st.gc_bases / (double)((int64_t)st.bases - (int64_t)st.n_bases);  // :48
  vector<string> out = {std::to_string(st.reads), nim_float(gc_content), std::to_string(st.gc_bases),
                        std::to_string(st.n_bases), std::to_string(st.bases)};
  puts(output_w_fnames(join(out, "\t"), fastq, basename, absolute).c_str());
}

// ------------------------------------------------------------------------------------------------
// fq-meta  (src/fq_meta.nim)
// ------------------------------------------------------------------------------------------------
static const char* kFqMetaHeader =
    "machine\tsequencer\tprob_sequencer\tflowcell\tflowcell_description\trun\tlane\tsequence_id\tindex1\tindex2\t"
    "qual_format\tqual_phred\tqual_multiple\tmin_qual\tmax_qual\tn_lines";

struct Instrument { const char* pattern; vector<string> sequencer; };
struct Flowcell { const char* pattern; vector<string> sequencer; const char* description; };

static const vector<Instrument> kInstrumentIDs = {  // fq_meta.nim:46-59
    {"HWI-M[0-9]{4}$", {"MiSeq"}}, {"HWUSI", {"GenomeAnalyzerIIx"}}, {"M[0-9]{5}$", {"MiSeq"}}, {"A[0-9]{5}$", {"NovaSeq"}},
    {"HWI-C[0-9]{5}$", {"HiSeq1500"}}, {"C[0-9]{5}$", {"HiSeq1500"}}, {"HWI-D[0-9]{5}$", {"HiSeq2500"}}, {"D[0-9]{5}$", {"HiSeq2500"}},
    {"J[0-9]{5}$", {"HiSeq3000"}}, {"K[0-9]{5}$", {"HiSeq3000", "HiSeq4000"}}, {"E[0-9]{5}$", {"HiSeqX"}},
    {"NB[0-9]{6}$", {"NextSeq"}}, {"NS[0-9]{6}$", {"NextSeq"}}, {"MN[0-9]{5}$", {"MiniSeq"}}};

static const vector<Flowcell> kFCIDs = {  // fq_meta.nim:69-92
    {"AAXX$", {"GenomeAnalyzer"}, ""},
    {"C[A-Z,0-9]{4}ANXX$", {"HiSeq1500", "HiSeq2000", "HiSeq2500"}, "High Output (8-lane) v4 flow cell"},
    {"C[A-Z,0-9]{4}ACXX$", {"HiSeq1000", "HiSeq1500", "HiSeq2000", "HiSeq2500"}, "High Output (8-lane) v3 flow cell"},
    {"H[A-Z,0-9]{4}ADXX$", {"HiSeq1500", "HiSeq2500"}, "Rapid Run (2-lane) v1 flow cell"},
    {"H[A-Z,0-9]{4}BCXX$", {"HiSeq1500", "HiSeq2500"}, "Rapid Run (2-lane) v2 flow cell"},
    {"H[A-Z,0-9]{4}BCXY$", {"HiSeq1500", "HiSeq2500"}, "Rapid Run (2-lane) v2 flow cell"},
    {"H[A-Z,0-9]{4}BBXX$", {"HiSeq4000"}, "(8-lane) v1 flow cell"},
    {"H[A-Z,0-9]{4}BBXY$", {"HiSeq4000"}, "(8-lane) v1 flow cell"},
    {"H[A-Z,0-9]{4}CCXX$", {"HiSeqX"}, "(8-lane) flow cell"},
    {"H[A-Z,0-9]{4}CCXY$", {"HiSeqX"}, "(8-lane) flow cell"},
    {"H[A-Z,0-9]{4}ALXX$", {"HiSeqX"}, "(8-lane) flow cell"},
    {"H[A-Z,0-9]{4}AGXX$", {"NextSeq"}, "High output flow cell"},
    {"H[A-Z,0-9]{4}BGXX$", {"NextSeq"}, "High output flow cell"},
    {"H[A-Z,0-9]{4}BGXY$", {"NextSeq"}, "High output flow cell"},
    {"H[A-Z,0-9]{4}BGX2$", {"NextSeq"}, "High output flow cell"},
    {"H[A-Z,0-9]{4}AFXX$", {"NextSeq"}, "Mid output flow cell"},
    {"H[A-Z,0-9]{4}DMXX$", {"NovaSeq"}, "S2 flow cell"},
    {"H[A-Z,0-9]{4}DSXX$", {"NovaSeq"}, "S2 flow cell"},
    {"^A[A-Z,0-9]{4}$", {"MiSeq"}, "MiSeq flow cell"},
    {"^B[A-Z,0-9]{4}$", {"MiSeq"}, "MiSeq flow cell"},
    {"^D[A-Z,0-9]{4}$", {"MiSeq"}, "MiSeq nano flow cell"},
    {"^G[A-Z,0-9]{4}$", {"MiSeq"}, "MiSeq micro flow cell"}};

static vector<string> dedup(const vector<string>& v) {
  vector<string> o;
  for (auto& x : v) if (std::find(o.begin(), o.end(), x) == o.end()) o.push_back(x);
  return o;
}

struct Detected { vector<string> sequencers; string prob, description; };

static Detected detect_sequencer(const string& machine, const string& flowcell) {  // fq_meta.nim:118-149
  vector<string> by_iid, by_fcid;
  string desc;
  for (auto& k : kInstrumentIDs) if (std::regex_search(machine, std::regex(k.pattern))) for (auto& s : k.sequencer) by_iid.push_back(s);
  for (auto& k : kFCIDs) if (std::regex_search(flowcell, std::regex(k.pattern))) { desc = k.description; for (auto& s : k.sequencer) by_fcid.push_back(s); }
  if (by_iid.empty() && by_fcid.empty()) return {{}, "", ""};
  if (by_iid.empty()) return {by_fcid, "likely:flowcell", desc};
  if (by_fcid.empty()) return {by_iid, "likely:machine", desc};
  vector<string> both;
  for (auto& i : by_iid) for (auto& j : by_fcid) if (i == j) both.push_back(i);
  both = dedup(both);
  if (!both.empty()) return {both, "high:machine+flowcell", desc};
  vector<string> u = by_iid;
  u.insert(u.end(), by_fcid.begin(), by_fcid.end());
  return {dedup(u), "uncertain", ""};
}

static vector<string> split_any(const string& s, const char* seps) {  // strutils.split(set[char])
  vector<string> o;
  string cur;
  for (char c : s) { if (strchr(seps, c)) { o.push_back(cur); cur.clear(); } else cur += c; }
  o.push_back(cur);
  return o;
}

static string strip_chars(const string& s, char ch) {
  size_t a = 0, b = s.size();
  while (a < b && s[a] == ch) a++;
  while (b > a && s[b - 1] == ch) b--;
  return s.substr(a, b - a);
}

struct ReadInfo { string sequence_id, machine, run, lane, flowcell; };

static ReadInfo extract_read_info(const string& line) {  // fq_meta.nim:152-178
  ReadInfo r;
  vector<string> q = split_any(line, ":/#");
  if (q.size() == 1) {
    r.sequence_id = strip_chars(q[0], '@');
  } else {
    r.machine = strip_chars(q[0], '@');
    if (line.find('/') != string::npos) {
      r.lane = q[1];
    } else {
      if (q.size() < 4) quit_error("index out of bounds", 1);  // qual_line[2] / [3] raise IndexError -> sc.nim:299-305
      r.run = q[1];
      r.flowcell = q[2];
      if (r.flowcell.find('_') != string::npos) r.flowcell = r.flowcell.substr(r.flowcell.find_last_of('_') + 1);
      r.lane = q[3];
    }
  }
  return r;
}

static string get_sequencer_name(const vector<string>& s) {  // fq_meta.nim:180-195
  auto has = [&](const char* x) { return std::find(s.begin(), s.end(), x) != s.end(); };
  if (has("HiSeq2000") || has("HiSeq2500")) return "HiSeq2000/2500";
  if (has("HiSeq1500") || has("HiSeq2500")) return "HiSeq1500/2500";
  if (has("HiSeq3000") || has("HiSeq4000")) return "HiSeq3000/4000";
  return s.empty() ? "" : s.back();
}

// First `max_lines` lines of the file, read on the host exactly as the reference does for the header
// columns (plain or gz stream).  The quality range itself comes from the GPU.
static bool head_lines(const string& path, bool gz, size_t max_lines, vector<string>* out) {
  string cur;
  auto feed = [&](const char* buf, int n) {
    for (int i = 0; i < n && out->size() < max_lines; i++) {
      if (buf[i] == '\n') { if (!cur.empty() && cur.back() == '\r') cur.pop_back(); out->push_back(cur); cur.clear(); }
      else cur += buf[i];
    }
  };
  char buf[65536];
  if (gz) {
    gzFile f = gzopen(path.c_str(), "rb");
    if (!f) return false;
    int n;
    while (out->size() < max_lines && (n = gzread(f, buf, sizeof buf)) > 0) feed(buf, n);
    gzclose(f);
  } else {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    size_t n;
    while (out->size() < max_lines && (n = fread(buf, 1, sizeof buf, f)) > 0) feed(buf, (int)n);
    fclose(f);
  }
  if (out->size() < max_lines && !cur.empty()) out->push_back(cur);
  return true;
}

struct FastqType { const char* name; const char* phred; int minimum, maximum; };
static const FastqType kFastqTypes[] = {  // fq_meta.nim:35-39 (bounds as written)
    {"Sanger", "Phred+33", 0, 40}, {"Solexa", "Solexa+64", 59, 104}, {"Illumina 1.3+", "Phred+64", 64, 104},
    {"Illumina 1.5+", "Phred+64", 64, 104}, {"Illumina 1.8+", "Phred+33", 0, 42}};

static void fq_meta(const string& fastq, long sample_n, bool basename, bool absolute) {
  if (fastq.size() < 3) quit_error("index out of bounds, the container is empty", 1);
  string low = fastq;
  std::transform(low.begin(), low.end(), low.begin(), ::tolower);
  const bool gz = low.compare(low.size() - 3, 3, ".gz") == 0;  // case-INsensitive, fq_meta.nim:219
  vector<string> lines;
  if (!head_lines(fastq, gz, (size_t)std::max(0L, sample_n) * 4, &lines)) quit_error("Unable to open file: " + fastq, 2);  // :223-224

  // ---- header-derived columns: unchanged host logic (fq_meta.nim:229-242, 251-258) ----
  ReadInfo info;
  vector<string> barcodes;
  // re"[ATCGN\+\-]{3,12}+" (fq_meta.nim:210): nim-regex reads the trailing '+' as one more repetition of the bounded group, so
  // dual-index barcodes such as AACGCTTA+GGTTCAGT (17 characters) match as a whole
  static const std::regex barcode_re("([ATCGN+\\-]{3,12})+");
  for (size_t i = 0; i < lines.size(); i += 4) {
    if (i == 0) info = extract_read_info(lines[0]);
    vector<string> q = split_any(lines[i], ":/#");
    if (q.size() > 2) {
      const string& bc = lines[i].find('/') != string::npos ? q[q.size() - 2] : q[q.size() - 1];
      if (std::regex_match(bc, barcode_re)) barcodes.push_back(bc);
    }
  }
  string sequencer, sequencer_prob, flowcell_description;
  if (!info.machine.empty() || !info.flowcell.empty()) {
    Detected d = detect_sequencer(info.machine, info.flowcell);
    sequencer = get_sequencer_name(d.sequencers);
    sequencer_prob = d.prob;
    flowcell_description = d.description;
  }
  string most_comm_barcode;
  if (!barcodes.empty()) {  // CountTable.largest(); ties keep the first barcode seen (unpinned: Nim iterates in hash order)
    std::map<string, int> cnt;
    int best = 0;
    for (auto& b : barcodes) { int c = ++cnt[b]; if (c > best) best = c; }
    for (auto& b : barcodes) if (cnt[b] == best) { most_comm_barcode = b; break; }
  }

  // ---- quality range: the GPU scan (replaces fq_meta.nim:245-246) ----
  fqgpu_ctx* ctx = context((uint64_t)std::max(0L, sample_n));
  fqgpu_stats st;
  int rc = fqgpu_meta_file_as(ctx, fastq.c_str(), gz, &st);  // only the sampled head is read, like the loop at :226
  if (rc == FQGPU_EIO) quit_error("Unable to open file: " + fastq, 2);
  if (rc != FQGPU_OK) quit_error(fqgpu_last_error(ctx), 1);
  if (st.meta_status == FQGPU_META_EMPTY_QUAL) quit_error("index out of bounds, the container is empty", 1);  // min() of an empty seq
  const long qual_min = (long)st.meta_qual_min, qual_max = (long)st.meta_qual_max;
  vector<string> names, phreds;
  for (auto& t : kFastqTypes) if (qual_min >= t.minimum && qual_max <= t.maximum) {  // :255
    names.push_back(t.name);
    if (std::find(phreds.begin(), phreds.end(), t.phred) == phreds.end()) phreds.push_back(t.phred);
  }
  vector<string> out = {info.machine, sequencer, sequencer_prob, info.flowcell, flowcell_description, info.run, info.lane,
                        info.sequence_id, most_comm_barcode, "", join(names, ";"), join(phreds, ";"),
                        names.size() > 1 ? "true" : "false", qual_min >= 0 ? std::to_string(qual_min) : "",
                        qual_max >= 0 ? std::to_string(qual_max) : "", std::to_string(st.meta_lines / 4)};  // :262-277
  puts(output_w_fnames(join(out, "\t"), fastq, basename, absolute).c_str());
}

// ------------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------------
// fq-dedup  (src/fq_dedup.nim:14-84)
// ------------------------------------------------------------------------------------------------
// The file (plain or gz, chosen like fq_count.nim:31) is read into host memory, the GPU marks the duplicate
// records (index + header hash + sort + byte compare), the kept records are echoed line by line as the
// reference does (Nim `lines` + echo: one '\r' before '\n' dropped, every line ends with '\n').
static void fq_dedup(const string& fastq) {
  if (fastq.size() < 3) quit_error("index out of bounds, the container is empty", 1);
  const bool gz = fastq.compare(fastq.size() - 3, 3, ".gz") == 0;
  (void)gz;  // the suffix selects GZFileStream vs FileStream in the reference (:31-34); zlib reads both transparently
  string data;
  {
    char buf[1 << 16];
    gzFile f = gzopen(fastq.c_str(), "rb");
    if (!f) quit_error("Unable to open file: " + fastq, 2);  // fq_dedup.nim:36-37
    int n;
    while ((n = gzread(f, buf, sizeof buf)) > 0) data.append(buf, (size_t)n);
    gzclose(f);
  }
  fqgpu_ctx* ctx = context(0);
  vector<uint8_t> keep(data.size() / 2 + 1);
  uint64_t nrec = 0, nlines = 0, ndups = 0;
  if (fqgpu_dedup_host(ctx, data.data(), data.size(), keep.data(), keep.size(), &nrec, &nlines, &ndups) != FQGPU_OK)
    quit_error(fqgpu_last_error(ctx), 1);
  if (ndups == 0) fprintf(stderr, "No Duplicates Found\nCopying fq to stdout\n");  // :52-54 (with a Bloom filter free of false positives)
  size_t start = 0, i = 0;
  bool write_ln = true;
  const size_t n = data.size();
  while (start < n) {
    const void* q = memchr(data.data() + start, '\n', n - start);
    const size_t e = q ? (size_t)((const char*)q - data.data()) : n;
    size_t len = e - start;
    if (q && len && data[e - 1] == '\r') len--;
    if (i % 4 == 0) write_ln = keep[i / 4] != 0;
    if (write_ln) { fwrite(data.data() + start, 1, len, stdout); fputc('\n', stdout); }
    i++;
    start = q ? e + 1 : n;
  }
  const double fpr = (double)0 / (double)ndups;  // :82: fp.float / n_dups.float with fp = 0
  fprintf(stderr, "total_reads: %llu\nduplicates %llu\nfalse-positive: 0\nfalse-positive-rate: %s\n", (unsigned long long)(nlines / 4),
          (unsigned long long)ndups, nim_float(fpr).c_str());
}

static void usage() {
  printf("Sequence data utilities (Version %s)\n\nUsage:\n  sc [options] COMMAND\n\nCommands:\n\n"
         "  fq-meta          Output metadata for FASTQ\n  fq-count         Counts lines in a FASTQ\n  fq-dedup         Removes exact duplicates from FASTQ Files\n\n"
         "Options:\n  --debug                    Debug\n  -h, --help                 Show this help\n"
         "\n(B200 build: only the FASTQ scanning commands are provided; see DESIGN.md)\n", kVersion);
}

int main(int argc, char** argv) {
  signal(SIGPIPE, SIG_IGN);  // sc.nim:45-46
  vector<string> args(argv + 1, argv + argc);
  // sc.nim:274-284: "-" becomes "STDIN" when stdin is a FIFO (and then fails to open, as in the reference)
  struct stat sst;
  if (fstat(0, &sst) == 0 && S_ISFIFO(sst.st_mode)) for (auto& a : args) if (a == "-") { a = "STDIN"; break; }
  if (args.size() <= 1) { usage(); return 0; }  // sc.nim:288-290 adds -h
  const string cmd = args[0];
  bool header = false, basename = false, absolute = false, help = false;
  string lines_opt = "100";  // sc.nim:70
  vector<string> files;
  for (size_t i = 1; i < args.size(); i++) {
    const string& a = args[i];
    if (a == "-t" || a == "--header") header = true;
    else if (a == "-b" || a == "--basename") basename = true;
    else if (a == "-a" || a == "--absolute") absolute = true;
    else if (a == "-h" || a == "--help") help = true;
    else if (a == "--debug") {}
    else if ((a == "-n" || a == "--lines") && cmd == "fq-meta") { if (i + 1 >= args.size()) quit_error("Error: option -n needs a value"); lines_opt = args[++i]; }
    else if (a.rfind("--lines=", 0) == 0 && cmd == "fq-meta") lines_opt = a.substr(8);
    else if (a.size() > 1 && a[0] == '-' && a != "-") quit_error("Error: unknown option " + a);
    else files.push_back(a);
  }
  if (cmd == "fq-count") {
    if (help) { printf("Counts lines in a FASTQ\n\nUsage:\n  fq-count [options] [fastq ...]\n\nArguments:\n  [fastq ...]      Input FASTQ\n\nOptions:\n  -t, --header               Output the header\n  -b, --basename             Add basename column\n  -a, --absolute             Add column for absolute path\n  -h, --help                 Show this help\n"); return 0; }
    if (header) puts(output_header(kFqCountHeader, basename, absolute).c_str());  // sc.nim:110-111
    else if (files.empty()) quit_error("No FASTQ specified", 3);                  // :112-113
    if (!files.empty()) {
      // sc.nim:115-116 loops over the files one after the other; here they are counted concurrently (one host
      // thread + private context each, fqgpu_count_files) and the rows are printed in argument order, stopping
      // where the sequential loop would have stopped.  FQGPU_THREADS=1 restores the sequential order of work.
      const int n = (int)files.size();
      for (int i = 0; i < n; i++) if (files[i].size() < 3) { files.resize(i + 1); break; }  // fastq[^3 .. ^1] raises there
      const int nok = files.back().size() < 3 ? (int)files.size() - 1 : (int)files.size();
      fqgpu_config cfg;
      memset(&cfg, 0, sizeof cfg);
      cfg.device = -1;
      cfg.flags = FQGPU_F_CORE_ONLY;  // fq-count prints reads, GC, N and bases only
      if (const char* e = getenv("FQGPU_CHUNK_MB")) cfg.chunk_bytes = (size_t)atol(e) << 20;
      if (const char* e = getenv("FQGPU_DEVICES")) if (!strcmp(e, "all")) cfg.device = FQGPU_DEVICE_ALL;
      vector<const char*> paths;
      for (int i = 0; i < nok; i++) paths.push_back(files[i].c_str());
      vector<fqgpu_stats> st((size_t)nok + 1);
      vector<int> rcs((size_t)nok + 1, FQGPU_OK);
      const int threads = getenv("FQGPU_THREADS") ? atoi(getenv("FQGPU_THREADS")) : 0;
      if (nok == 1 && cfg.device == FQGPU_DEVICE_ALL) {  // one file, all GPUs: byte-range shards in this process
        cfg.device = -1;
        rcs[0] = fqgpu_count_file_sharded(&cfg, paths[0], nullptr, 0, &st[0]);
      } else if (nok) {
        fqgpu_count_files(&cfg, paths.data(), nullptr, nok, threads, st.data(), rcs.data());
      }
      for (int i = 0; i < nok; i++) {
        if (rcs[i] == FQGPU_ECUDA && i == 0) quit_error(string("GPU unavailable: ") + fqgpu_last_error(nullptr), 1);
        fq_count_row(files[i], rcs[i], st[i], fqgpu_last_error(nullptr), basename, absolute);
      }
      if (nok < (int)files.size()) quit_error("index out of bounds, the container is empty", 1);  // (unpinned text)
    }
  } else if (cmd == "fq-meta") {
    if (help) { printf("Output metadata for FASTQ\n\nUsage:\n  fq-meta [options] [fastq ...]\n\nArguments:\n  [fastq ...]      List of FASTQ files\n\nOptions:\n  -n, --lines=LINES          Number of sequences to sample (n_lines) for qual and index/barcode determination (default: 100)\n  -t, --header               Output the header\n  -b, --basename             Add basename column\n  -a, --absolute             Add column for absolute path\n  -h, --help                 Show this help\n"); return 0; }
    if (header) puts(output_header(kFqMetaHeader, basename, absolute).c_str());   // sc.nim:75-76
    char* endp = nullptr;
    const long n = strtol(lines_opt.c_str(), &endp, 10);                            // parseInt(opts.lines), sc.nim:79
    if (!files.empty() && (endp == lines_opt.c_str() || *endp)) quit_error("invalid integer: " + lines_opt, 1);
    for (auto& f : files) fq_meta(f, n, basename, absolute);
  } else if (cmd == "fq-dedup") {
    if (help) { printf("Removes exact duplicates from FASTQ Files\n\nUsage:\n  fq-dedup [options] fastq\n\nArguments:\n  fastq            Input FASTQ\n\nOptions:\n  -h, --help                 Show this help\n"); return 0; }
    if (files.size() != 1) quit_error("Error: fq-dedup takes one FASTQ", 1);  // sc.nim:120 nargs = 1
    fq_dedup(files[0]);
  } else if (cmd == "-h" || cmd == "--help") {
    usage();
  } else {
    quit_error("Error: Unknown command '" + cmd + "' (this build provides fq-count, fq-meta and fq-dedup)");
  }
  if (g_ctx) fqgpu_destroy(g_ctx);
  return 0;
}
