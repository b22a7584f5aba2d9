// gt_host.cpp -- see gt_host.h
#include "gt_host.h"
#include <limits.h>
#include <stdlib.h>
#include <string.h>
#include <iostream>

namespace gt {

void die(const std::string &msg) {
  fprintf(stderr, "\n");
  fprintf(stderr, "Error: %s\n", msg.c_str());
  exit(1);
}

void die_line(long line, const std::string &msg) {
  fprintf(stderr, "\n");
  fprintf(stderr, "Error: Line %ld: %s\n", line, msg.c_str());
  exit(1);
}

// ---------------------------------------------------------------------------------------------
// CmdLine
// ---------------------------------------------------------------------------------------------
void CmdLine::AddOperation(const std::string &op, const std::string &usage, const std::string &description, const std::string &details) {
  ops_[op] = Operation{usage, description, details};
}

void CmdLine::AddOption(const char *opt, bool *ptr, bool def, const char *d) { Option o{opt, d, 'b', ptr}; o.def_b = def; options_.push_back(o); }
void CmdLine::AddOption(const char *opt, char *ptr, char def, const char *d) { Option o{opt, d, 'c', ptr}; o.def_c = def; options_.push_back(o); }
void CmdLine::AddOption(const char *opt, long *ptr, long def, const char *d) { Option o{opt, d, 'l', ptr}; o.def_l = def; options_.push_back(o); }
void CmdLine::AddOption(const char *opt, unsigned long *ptr, unsigned long def, const char *d) { Option o{opt, d, 'u', ptr}; o.def_u = def; options_.push_back(o); }
void CmdLine::AddOption(const char *opt, double *ptr, double def, const char *d) { Option o{opt, d, 'd', ptr}; o.def_d = def; options_.push_back(o); }
void CmdLine::AddOption(const char *opt, const char **ptr, const char *def, const char *d) { Option o{opt, d, 's', ptr}; o.def_s = def; options_.push_back(o); }

int CmdLine::Read(char **argv, int argc) {
  // defaults first (CmdLine::Init, core.cpp:2438)
  for (auto &o : options_) {
    switch (o.type) {
      case 'b': *(bool *)o.ptr = o.def_b; break;
      case 'c': *(char *)o.ptr = o.def_c; break;
      case 'l': *(long *)o.ptr = o.def_l; break;
      case 'u': *(unsigned long *)o.ptr = o.def_u; break;
      case 'd': *(double *)o.ptr = o.def_d; break;
      case 's': o.cur_s = o.def_s; *(const char **)o.ptr = o.cur_s.c_str(); break;
    }
  }
  int i = 1;
  while (i < argc) {
    if (argv[i][0] != '-') return i;                                   // first non-option ends the options (core.cpp:2428)
    Option *found = nullptr;
    for (auto &o : options_) if (o.opt == argv[i]) found = &o;        // a later registration of the same name wins, like the map
    if (!found) { fprintf(stderr, "Error: unknown option '%s'!\n", argv[i]); exit(1); }
    if (found->type == 'b') { *(bool *)found->ptr = !*(bool *)found->ptr; i++; continue; }   // flags TOGGLE (core.cpp:2212)
    if (i + 1 >= argc) { fprintf(stderr, "Error: could not set option '%s'!\n", found->opt.c_str()); exit(1); }
    const char *v = argv[i + 1];
    switch (found->type) {
      case 'c': *(char *)found->ptr = v[0]; break;
      case 'l': *(long *)found->ptr = atol(v); break;
      case 'u': *(unsigned long *)found->ptr = (unsigned long)atol(v); break;
      case 'd': *(double *)found->ptr = atof(v); break;
      case 's': found->cur_s = v; *(const char **)found->ptr = found->cur_s.c_str(); break;
    }
    i += 2;
  }
  return argc;
}

void CmdLine::Print() {
  for (auto &o : options_) {
    printf("  %-25s %-80s ", o.opt.c_str(), o.description.c_str());
    switch (o.type) {
      case 'b': printf("[%s]", *(bool *)o.ptr ? "true" : "false"); break;
      case 'c': printf("[%c]", *(char *)o.ptr); break;
      case 'l': printf("[%ld]", *(long *)o.ptr); break;
      case 'u': printf("[%lu]", *(unsigned long *)o.ptr); break;
      case 'd': printf("[%.6e]", *(double *)o.ptr); break;
      case 's': printf("[%s]", *(const char **)o.ptr); break;
    }
    printf("\n");
  }
}

void CmdLine::OperationSummary(const std::string &usage, const std::string &description) {
  std::cout << '\n' << "USAGE: \n" << "  " << program_ << " " << usage << '\n' << '\n';
  if (version_ != "") std::cout << "VERSION: \n" << "  " << version_ << '\n' << '\n';
  std::cout << "DESCRIPTION: \n" << "  " << description << '\n' << '\n' << "OPERATION: \n";
  std::cout.flush();
  for (auto &kv : ops_) printf("  %-15s %s\n", kv.first.c_str(), kv.second.description.c_str());
  fflush(stdout);
  std::cout << '\n';
}

void CmdLine::OperationUsage() {
  auto it = ops_.find(current_);
  if (it == ops_.end()) { fprintf(stderr, "Error: [CmdLine::AddOperation] operation not found!\n"); exit(1); }
  std::cout << '\n' << "USAGE: \n" << "  " << program_ << " " << it->first << " " << it->second.usage << '\n' << '\n';
  std::cout << "DESCRIPTION: \n" << "  " << it->second.description << '\n' << '\n';
  if (it->second.details != "") std::cout << "DETAILS: \n" << "  " << it->second.details << '\n' << '\n';
  std::cout << "OPTIONS: \n";
  std::cout.flush();
  Print();
  fflush(stdout);
  std::cout << '\n';
}

// ---------------------------------------------------------------------------------------------
// LineReader
// ---------------------------------------------------------------------------------------------
LineReader::LineReader(const char *path) {
  buf_.resize(1 << 22);
  if (path == nullptr) { fp_ = stdin; return; }
  FILE *probe = fopen(path, "r");
  if (probe == nullptr) { fprintf(stderr, "[CreateFileBuffer] Error: cannot open file '%s'!\n", path); exit(1); }
  fclose(probe);
  gz_ = gzopen(path, "rb");                          // zlib passes plain text through unchanged
  if (gz_ == nullptr) { fprintf(stderr, "[CreateFileBuffer] Error: cannot open file '%s'!\n", path); exit(1); }
  gzbuffer(gz_, 1 << 20);
}

LineReader::~LineReader() {
  if (gz_) gzclose(gz_);
}

bool LineReader::Fill() {
  if (eof_) return false;
  if (begin_ > 0) { memmove(buf_.data(), buf_.data() + begin_, end_ - begin_); end_ -= begin_; begin_ = 0; }
  if (end_ == buf_.size()) buf_.resize(buf_.size() * 2);             // a line longer than the buffer
  size_t want = buf_.size() - end_;
  long got = gz_ ? (long)gzread(gz_, buf_.data() + end_, (unsigned)std::min<size_t>(want, 1u << 30))
                 : (long)fread(buf_.data() + end_, 1, want, fp_);
  if (got <= 0) { eof_ = true; return false; }
  end_ += (size_t)got;
  return true;
}

char *LineReader::Next() {
  size_t scan = begin_;
  for (;;) {
    char *nl = (char *)memchr(buf_.data() + scan, '\n', end_ - scan);
    if (nl) {
      char *line = buf_.data() + begin_;
      *nl = 0;
      begin_ = (size_t)(nl - buf_.data()) + 1;
      line_no_++;
      return line;
    }
    const size_t had = end_ - begin_;
    if (!Fill()) return nullptr;                                       // an unterminated last line is dropped (core.cpp:243)
    scan = begin_ + had;
  }
}

// ---------------------------------------------------------------------------------------------
// tokenizer helpers with the reference's semantics (core.cpp:577-625): leading blanks are skipped,
// a token ends at the delimiter; numbers go through atol.
// ---------------------------------------------------------------------------------------------
static int CountTokens(const char *s, char delim) {
  if (s == nullptr) return 0;
  int k = 0, n = 0;
  while (s[k] == ' ') k++;
  for (;;) {
    if (s[k] == 0) return n;
    while (s[k] != 0 && s[k] != delim) k++;
    if (s[k] == delim) k++;
    n++;
    while (s[k] == ' ') k++;
    if (s[k] == 0) return n;
  }
}

static char *NextToken(char **p, char delim) {
  char *b = *p;
  while (*b == ' ') b++;
  int k = 0;
  while (b[k] != 0 && b[k] != delim) k++;
  if (b[k] == 0) *p = b + k;
  else { b[k] = 0; *p = b + k + 1; }
  return b;
}

char ProcessStrand(const char *t) {
  if (!strcmp(t, "1") || !strcmp(t, "+")) return '+';
  if (!strcmp(t, "-1") || !strcmp(t, "-")) return '-';
  if (!strcmp(t, ".")) return '+';
  std::cerr << "Error: invalid strand '" << t << "'!\n";
  exit(1);
}

int32_t ChromTable::Get(const char *chrom) {
  auto it = id.find(chrom);
  if (it != id.end()) return it->second;
  int32_t v = (int32_t)name.size();
  id.emplace(chrom, v);
  name.push_back(chrom);
  return v;
}

void RegionBatch::Clear() {
  chrom.clear(); start.clear(); stop.clear(); strand.clear(); weight.clear(); label.clear(); line.clear();
  offset.assign(1, 0);
  multi = false;
}

bool RegionWellFormed(const RegionBatch &b, int64_t k) {
  for (int64_t i = b.offset[k] + 1; i < b.offset[k + 1]; i++) {
    if (b.chrom[i] != b.chrom[b.offset[k]] || b.strand[i] != b.strand[b.offset[k]]) return false;
    if (b.start[i] < b.start[i - 1] || b.start[i] <= b.stop[i - 1]) return false;
  }
  return true;
}

bool SortChecker::Accept(const std::string &c, char s, long st) {
  bool ok = true;
  if (have) {
    const int cmp = strcmp(c.c_str(), chrom.c_str());                 // IsBefore, genomic_intervals.cpp:396-401
    if (cmp != 0) ok = !(cmp < 0);
    else if (by_strand && s != strand) ok = !(s < strand);
    else ok = !(st < start);
  }
  have = true; chrom = c; strand = s; start = st;
  return ok;
}

// ---------------------------------------------------------------------------------------------
// RegionReader
// ---------------------------------------------------------------------------------------------
RegionReader::RegionReader(const char *path, ChromTable *chroms, bool keep_labels, long max_label_value)
    : reader_(path), chroms_(chroms), keep_labels_(keep_labels), max_label_value_(max_label_value) {
  // header skipping and format sniffing (genomic_intervals.cpp:3713-3759)
  char *next = reader_.Next();
  auto is_track = [](const char *s) { return strncmp(s, "browser ", 8) == 0 || strncmp(s, "track ", 6) == 0; };
  if (next == nullptr) { format_ = "EMPTY"; return; }
  if (is_track(next)) { while (next && is_track(next)) next = reader_.Next(); }
  else if (next[0] == '@') { format_ = "SAM"; while (next && next[0] == '@') next = reader_.Next(); }
  else if (next[0] == '#' && next[1] == '#') { format_ = "GFF"; while (next && next[0] == '#' && next[1] == '#') next = reader_.Next(); }
  if (next == nullptr) { format_ = "EMPTY"; return; }
  pending_ = next;
  if (format_ != "") return;
  if (next[0] == '>') { format_ = "SEQ"; return; }
  const long n_tokens = CountTokens(next, '\t');
  if (n_tokens == 1) format_ = "BED";                                 // space-separated BED lands here
  if (n_tokens == 2) format_ = "REG";
  if (n_tokens >= 3 && n_tokens <= 6) format_ = "BED";
  else if (n_tokens >= 6) {
    std::string copy(next);
    char *p = &copy[0], *t = nullptr;
    for (int i = 1; i <= 6; i++) t = NextToken(&p, '\t');
    if (strchr(t, '+') != nullptr || strchr(t, '-') != nullptr) format_ = "BED";
    else if (n_tokens >= 11) format_ = "SAM";
    else if (n_tokens >= 8 && n_tokens <= 10) format_ = "GFF";
  }
}

void RegionReader::Push(RegionBatch *out, const char *chrom, char strand, long start, long stop, long line_no) {
  if (start < INT32_MIN || start > INT32_MAX || stop < INT32_MIN || stop > INT32_MAX)
    die_line(line_no, "coordinate does not fit in 32 bits (not supported by the GPU engine)!");
  out->chrom.push_back(chroms_->Get(chrom));
  out->strand.push_back((int8_t)strand);
  out->start.push_back((int32_t)start);
  out->stop.push_back((int32_t)stop);
}

void RegionReader::ParseLine(char *inp, long line_no, RegionBatch *out) {
  std::string label;
  const size_t first_interval = out->chrom.size();
  if (format_ == "BED") {                                            // GenomicRegionBED::Read, genomic_intervals.cpp:2157-2182
    const char sep = strchr(inp, '\t') == nullptr ? ' ' : '\t';
    const int n_tokens = CountTokens(inp, sep);
    if (n_tokens < 3) die_line(line_no, "number of tokens should be at least 3 for BED format!");
    const char *chromosome = NextToken(&inp, sep);
    const long start = atol(NextToken(&inp, sep)) + 1;
    const long stop = atol(NextToken(&inp, sep));
    char strand = '+';
    label = n_tokens == 3 ? "_" : NextToken(&inp, sep);
    if (n_tokens >= 5) NextToken(&inp, sep);                          // score
    if (n_tokens >= 6) strand = ProcessStrand(NextToken(&inp, sep));
    if (n_tokens >= 8) { NextToken(&inp, sep); NextToken(&inp, sep); }  // thickStart, thickEnd
    if (n_tokens >= 9) NextToken(&inp, sep);                          // itemRgb
    if (n_tokens != 12) Push(out, chromosome, strand, start, stop, line_no);
    else {
      const long n_intervals = atol(NextToken(&inp, sep));
      char *sizes = NextToken(&inp, sep);
      char *starts = NextToken(&inp, sep);
      std::vector<long> bsize((size_t)std::max(n_intervals, 0L)), bstart((size_t)std::max(n_intervals, 0L));
      for (long k = 0; k < n_intervals; k++) bsize[k] = atol(NextToken(&sizes, ','));
      for (long k = 0; k < n_intervals; k++) bstart[k] = atol(NextToken(&starts, ','));
      for (long k = 0; k < n_intervals; k++) {
        const long s = start + bstart[k];
        Push(out, chromosome, strand, s, bsize[k] + s - 1, line_no);
      }
    }
  } else if (format_ == "REG") {                                     // GenomicRegion::Read, genomic_intervals.cpp:805-838
    label = NextToken(&inp, '\t');
    if (strchr(inp, ',') == nullptr) {
      const int n_tokens = CountTokens(inp, ' ');
      if (n_tokens < 4 || n_tokens % 4 != 0) die_line(line_no, "invalid number of tokens!");
      for (int k = 0; k < n_tokens / 4; k++) {
        const char *chromosome = NextToken(&inp, ' ');
        const char strand = ProcessStrand(NextToken(&inp, ' '));
        const long start = atol(NextToken(&inp, ' '));
        const long stop = atol(NextToken(&inp, ' '));
        Push(out, chromosome, strand, start, stop, line_no);
      }
    } else {
      if (CountTokens(inp, ' ') != 4) die_line(line_no, "invalid number of tokens in compact format!");
      const char *chromosome = NextToken(&inp, ' ');
      const char strand = ProcessStrand(NextToken(&inp, ' '));
      char *starts = NextToken(&inp, ' ');
      char *stops = NextToken(&inp, ' ');
      const int n_intervals = CountTokens(starts, ',');
      if (CountTokens(stops, ',') != n_intervals) die_line(line_no, "number of starts/stops should be equal");
      for (int k = 0; k < n_intervals; k++) {
        const long start = atol(NextToken(&starts, ','));
        const long stop = atol(NextToken(&stops, ','));
        Push(out, chromosome, strand, start, stop, line_no);
      }
    }
  } else if (format_ == "GFF") {                                     // GenomicRegionGFF::Read, genomic_intervals.cpp:3501-3517
    const int n_tokens = CountTokens(inp, '\t');
    if (n_tokens < 8 || n_tokens > 10) die_line(line_no, "wrong number of tokens for GFF format!");
    const char *seqname = NextToken(&inp, '\t');
    NextToken(&inp, '\t'); NextToken(&inp, '\t');                      // source, feature
    const long start = atol(NextToken(&inp, '\t'));
    const long end = atol(NextToken(&inp, '\t'));
    NextToken(&inp, '\t');                                             // score
    const char strand = NextToken(&inp, '\t')[0];                      // raw character, NOT normalised (:3512)
    NextToken(&inp, '\t');                                             // frame
    label = n_tokens == 8 ? "_" : NextToken(&inp, '\t');
    Push(out, seqname, strand, start, end, line_no);
  } else if (format_ == "SAM" || format_ == "SEQ") {
    die("input format " + format_ + " is not supported by this build (BED, REG and GFF are)!\n");
  } else {
    die("unsupported input format!\n");
  }
  if (out->chrom.size() - first_interval != 1) out->multi = true;
  out->offset.push_back((int64_t)out->chrom.size());
  out->line.push_back(line_no);
  if (keep_labels_) out->label.push_back(label);
  if (max_label_value_ > 1) {                                          // GetLabelValue, genomic_intervals.cpp:1081-1085
    const long w = std::min(max_label_value_, atol(label.c_str()));
    if (w < INT32_MIN || w > INT32_MAX) die_line(line_no, "label value does not fit in 32 bits (not supported by the GPU engine)!");
    out->weight.push_back((int32_t)w);
  }
}

int64_t RegionReader::Read(RegionBatch *out, int64_t max_regions) {
  out->Clear();
  if (format_ == "EMPTY") return 0;
  int64_t n = 0;
  while (n < max_regions) {
    char *line = pending_ ? pending_ : reader_.Next();
    pending_ = nullptr;
    if (line == nullptr) break;
    ParseLine(line, reader_.line_no(), out);
    n++;
  }
  return n;
}

}  // namespace gt
