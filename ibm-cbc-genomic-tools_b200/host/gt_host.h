// gt_host.h -- host side of the drop-in drivers: command line, line reader (text / gzip / stdin),
// region-file parsers that emit the packed SoA arrays of include/gtb200.h.
//
// Mirrors, for the overlap/count path only, the reference's
//   CmdLineWithOperations          gtools/core.h:772-932, core.cpp:2204-2646
//   FileBufferText / FileBufferGZ  gtools/core.cpp:130-349 (incl. the "last line without newline is dropped" quirk, :243)
//   GenomicRegionSet format sniffing and header skipping   gtools/genomic_intervals.cpp:3713-3759
//   GenomicRegionBED::Read / GenomicRegion::Read (REG) / GenomicRegionGFF::Read / GenomicRegionSAM::Read
//                                  :2157-2182, :805-838, :3501-3517, :2771-2813
// Written from their behaviour; no reference code is reused.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <sys/types.h>
#include <zlib.h>
#include "gt_inflate.h"
#include <map>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace gt {

[[noreturn]] void die(const std::string &msg);                     // "\nError: msg\n" on stderr, exit(1)   (GenomicRegionSet::PrintError)
[[noreturn]] void die_line(long line, const std::string &msg);     // "\nError: Line N: msg\n", exit(1)      (GenomicRegion::PrintError)

// ---------------------------------------------------------------------------------------------
// command line: first argument = operation, then options until the first token not starting with '-'
// ---------------------------------------------------------------------------------------------
class CmdLine {
 public:
  CmdLine(const std::string &program, const std::string &version) : program_(program), version_(version) {}
  void AddOperation(const std::string &op, const std::string &usage, const std::string &description, const std::string &details);
  bool HasOperation(const std::string &op) const { return ops_.count(op) != 0; }
  void SetCurrentOperation(const std::string &op) { current_ = op; }
  const std::string &current() const { return current_; }
  void AddOption(const char *opt, bool *ptr, bool def, const char *description);
  void AddOption(const char *opt, char *ptr, char def, const char *description);
  void AddOption(const char *opt, long *ptr, long def, const char *description);
  void AddOption(const char *opt, unsigned long *ptr, unsigned long def, const char *description);
  void AddOption(const char *opt, double *ptr, double def, const char *description);
  void AddOption(const char *opt, const char **ptr, const char *def, const char *description);
  // argv[0] is skipped; returns the index (relative to argv) of the first non-option argument
  int Read(char **argv, int argc);
  void OperationSummary(const std::string &usage, const std::string &description);
  void OperationUsage();

 private:
  struct Option {
    std::string opt, description;
    char type;            // b c l u d s
    void *ptr;
    bool def_b; char def_c; long def_l; unsigned long def_u; double def_d; std::string def_s, cur_s;
  };
  struct Operation { std::string usage, description, details; };
  void Print();
  std::string program_, version_, current_;
  std::map<std::string, Operation> ops_;
  std::vector<Option> options_;
};

// ---------------------------------------------------------------------------------------------
// line reader
// ---------------------------------------------------------------------------------------------
class LineReader {
 public:
  explicit LineReader(const char *path);            // nullptr = stdin
  ~LineReader();
  // next complete line (without its '\n'), or nullptr at end.  A trailing line that is not terminated by
  // '\n' is dropped, exactly as the reference does.  The pointer is valid until the next call.
  char *Next();
  long line_no() const { return line_no_; }         // 1-based number of the line last returned
  // The rest of the current block, or the next one: complete lines, each still terminated by '\n', in [*begin, *end);
  // false at end.  The pointers are valid until the next call.  The caller counts the lines and reports them with Advance.
  bool NextRun(char **begin, char **end);
  void Advance(long n_lines) { line_no_ += n_lines; }

 private:
  // A producer thread reads (and inflates) the file into a ring of blocks that end on a line boundary while the consumer
  // parses the previous block.
  struct Block { char *data = nullptr; size_t cap = 0, len = 0; };
  static const int kBlocks = 3;
  void Produce();
  long ReadSome(char *dst, size_t want);
  bool Acquire();                                     // make the next block current; false at end
  gzFile gz_ = nullptr;
  int fd_ = -1;
  struct BamDecoder;                                  // BAM input: records re-spelt as SAM text lines, as the reference's FileBufferBAM does
  BamDecoder *bam_ = nullptr;
  struct Bgzf;                                        // blocked gzip (bgzip, BAM): the members of a stretch of the file are inflated side by side
  Bgzf *bgzf_ = nullptr;
  GzipStream *gzs_ = nullptr;                         // any other gzip file: one thread, a decoder faster than zlib's (gt_inflate.h)
  bool fault_warned_ = false;
  long ReadInflated(void *dst, size_t want);          // from the gzip stream, whichever way it is read
  char prefix_[4];                                    // the first inflated bytes of a gzip file (read to tell BAM from text)
  int prefix_len_ = 0, prefix_pos_ = 0;
  bool regular_ = false;                              // a regular file, read with pread at offset_ (by several threads when the request is large)
  off_t offset_ = 0;
  Block block_[kBlocks];
  std::mutex mu_;
  std::condition_variable cv_;
  int ready_ = 0, head_ = 0;                          // filled blocks not yet released (the current one included); next block to hand out
  bool done_ = false, quit_ = false;
  std::thread producer_;
  int cur_ = -1;                                      // block the consumer holds
  size_t pos_ = 0;                                    // first unread byte of it
  long line_no_ = 0;
};

// ---------------------------------------------------------------------------------------------
// region sets
// ---------------------------------------------------------------------------------------------
struct ChromTable {
  std::unordered_map<std::string, int32_t> id;
  std::vector<std::string> name;
  int32_t Get(const char *chrom);                     // thread-safe
 private:
  std::mutex mu_;
};

// Growable array of a plain type that does not initialise what it grows by (the parser fills it from several threads).
template <typename T>
class PodVec {
 public:
  PodVec() = default;
  PodVec(const PodVec &) = delete;
  PodVec &operator=(const PodVec &) = delete;
  ~PodVec() { free(p_); }
  T *data() { return p_; }
  const T *data() const { return p_; }
  size_t size() const { return n_; }
  bool empty() const { return n_ == 0; }
  T &operator[](size_t i) { return p_[i]; }
  const T &operator[](size_t i) const { return p_[i]; }
  T &back() { return p_[n_ - 1]; }
  const T &back() const { return p_[n_ - 1]; }
  void clear() { n_ = 0; }
  void reserve(size_t n) { if (n > cap_) Grow(n); }
  void resize(size_t n) { reserve(n); n_ = n; }      // new elements are NOT initialised
  void push_back(T v) { if (n_ == cap_) Grow(cap_ ? cap_ * 2 : 1024); p_[n_++] = v; }
  void assign(size_t n, T v) { resize(n); for (size_t i = 0; i < n; i++) p_[i] = v; }
 private:
  void Grow(size_t cap) {
    p_ = (T *)realloc(p_, cap * sizeof(T));
    if (p_ == nullptr) { fprintf(stderr, "Error: out of memory!\n"); exit(1); }
    cap_ = cap;
  }
  T *p_ = nullptr;
  size_t n_ = 0, cap_ = 0;
};

struct RegionBatch {                                  // packed SoA, regions in file order
  PodVec<int32_t> chrom, start, stop;
  PodVec<int8_t> strand;
  PodVec<int32_t> weight;                             // per region (only if weights are in use)
  PodVec<int64_t> offset;                             // per region + 1; maintained always, passed on only if some region is multi-interval
  std::vector<std::string> label;                     // per region (only if keep_labels)
  long first_line = 0;                                // source line of region 0; every data line is one region, so region k is on first_line + k
  bool multi = false;
  int64_t n_regions() const { return (int64_t)offset.size() - 1; }
  long line(int64_t k) const { return first_line + (long)k; }
  void Clear();
};

// A fatal condition found while parsing.  Lines are parsed by several threads at once; the one that comes first in the file
// is the one reported, exactly as the reference's sequential reader would.
struct ParseError {
  bool with_line = false;
  long line = 0;
  std::string message;                                // printed as die_line / die would
  bool raw = false;                                   // message goes to stderr as is (ProcessStrand's wording)
};

class RegionReader {
 public:
  RegionReader(const char *path, ChromTable *chroms, bool keep_labels, long max_label_value);
  const std::string &format() const { return format_; }
  // the header lines skipped in front of the data (UCSC browser / track lines, SAM '@' lines, GFF '##' lines), each with its
  // newline: the reference echoes them to stdout for sets it opens with hide_header == false (ProcessFileHeader, :3713-3731)
  const std::string &header() const { return header_; }
  // Parses regions into `out` (cleared first) until it holds at least max_regions of them or the input ends; returns the
  // number parsed (0 at end).  Lines are parsed by several threads.  A malformed line ends the stream: the regions before it
  // are returned, failed() turns true and Fail() reports it -- so that a caller that streams can first report what an
  // earlier line did wrong further down the pipeline, as the reference's line-by-line loop would.
  int64_t Read(RegionBatch *out, int64_t max_regions);
  bool failed() const { return failed_; }
  [[noreturn]] void Fail() const;                     // prints the parse error the way the reference does, exit(1)
  // For the operations that print regions (subset, overlap, gsort): like Read, one thread, and every region's input line is
  // kept in `raw` (appended; raw[k] belongs to region k of out).  Needs keep_labels.
  int64_t ReadKeep(RegionBatch *out, std::vector<std::string> *raw, int64_t max_regions);
  // Read, then Fail() at once if a line was malformed (sets that are loaded whole)
  int64_t ReadAll(RegionBatch *out) { const int64_t n = Read(out, INT64_MAX); if (failed_) Fail(); return n; }
 private:
  enum Format { F_NONE, F_BED, F_REG, F_GFF, F_SAM, F_SEQ, F_EMPTY };
  struct ChromCache;                                  // per-thread view of the chromosome table
  struct Piece;                                       // one thread's share of a run
  bool ParseLine(char *line, RegionBatch *out, ChromCache *cache, ParseError *err) const;
  bool Push(RegionBatch *out, ChromCache *cache, const char *chrom, char strand, long start, long stop, ParseError *err) const;
  // lines of [begin, end), over the parsing threads, appended to out; false on a malformed line (error_ set)
  bool ParseRun(char *begin, char *end, RegionBatch *out);
  char *ParseBedLine(char *line, RegionBatch *out, ChromCache *cache) const;   // clean BED3-6 only; the line's '\n', or nullptr = not handled
  char *ParseSamLine(char *line, RegionBatch *out, ChromCache *cache) const;   // clean SAM lines likewise
  LineReader reader_;
  ChromTable *chroms_;
  bool keep_labels_;
  long max_label_value_;
  std::string format_;                                // BED REG GFF SAM SEQ EMPTY ""
  std::string header_;
  Format fmt_ = F_NONE;
  int threads_ = 1;                                   // GT_PARSE_THREADS, default: the host's cores (at most 32)
  char *pending_ = nullptr;                           // first data line, already read during format detection
  bool failed_ = false;
  ParseError error_;
  std::vector<Piece *> piece_;                        // kept between runs: their batches' memory is reused
 public:
  ~RegionReader();
};

// What GenomicRegion{,BED,GFF,SAM}::Print writes for region k of `b` (genomic_intervals.cpp:843-849, :2188-2219, :3523-3530,
// :2819-2825), appended to `out` with its newline: the parsed coordinates and strand, the label, and the fields of the input line
// the reference keeps as they are (BED score / thick / rgb, GFF source / feature / score / frame / comment, the SAM columns).
// `format` is RegionReader::format(); `raw` the region's input line.
void PrintRegion(const std::string &format, const std::string &raw, const RegionBatch &b, int64_t k, const ChromTable &chroms, std::string *out);
// What GenomicRegion{,BED,GFF,SAM}::PrintModified(label, start, stop) writes for region k (genomic_intervals.cpp:909-913, :2310-2314,
// :2854-2860, :3555-3562): the region's first interval moved to [start, stop] under a new label, the format's other columns
// from the input line.  The region's intervals must share chromosome and strand (fatal otherwise, as in the reference).
void PrintModified(const std::string &format, const std::string &raw, const RegionBatch &b, int64_t k, const ChromTable &chroms,
                   const char *label, long start, long stop, std::string *out);

char ProcessStrand(const char *token);                // '1','+' -> '+'; '-1','-' -> '-'; '.' -> '+'; else fatal  (genomic_intervals.cpp:5956-5962)
bool RegionWellFormed(const RegionBatch &b, int64_t k);

// Wall-clock marks of a driver's phases, printed to stderr at exit when GT_TIMING is set (stdout is never touched).
class PhaseTimer {
 public:
  PhaseTimer();
  void Mark(const char *phase);                       // the time since the previous mark belongs to `phase`
  ~PhaseTimer();
 private:
  bool on_;
  double last_, t0_;
  std::vector<std::pair<std::string, double>> marks_;
};

// -S sortedness checker (GenomicRegion::IsBefore on consecutive regions)
struct SortChecker {
  bool by_strand = false;
  bool have = false;
  std::string chrom; char strand = 0; long start = 0;
  // returns false if region (chrom,strand,start) sorts before the previous one
  bool Accept(const std::string &c, char s, long st);
};

// The drivers create the CUDA context on a helper thread while the main thread reads the reference set (context creation takes
// 0.4 - 2 s on a cold box; serialised after the loading it was most of a small run's wall time).  Every fatal path of the host
// code ends in exit(); leaving the process while the helper thread is inside the CUDA runtime's initialisation is asking for a
// crash in its atexit handlers, so exit() is routed through Exit(), which first runs `exit_hook` (the driver's "wait for the
// helper thread").
extern void (*exit_hook)();
[[noreturn]] void Exit(int code);

}  // namespace gt
#define exit(code) ::gt::Exit(code)
