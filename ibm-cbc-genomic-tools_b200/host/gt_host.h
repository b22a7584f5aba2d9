// gt_host.h -- host side of the drop-in drivers: command line, line reader (text / gzip / stdin),
// region-file parsers that emit the packed SoA arrays of include/gtb200.h.
//
// Mirrors, for the overlap/count path only, the reference's
//   CmdLineWithOperations          gtools/core.h:772-932, core.cpp:2204-2646
//   FileBufferText / FileBufferGZ  gtools/core.cpp:130-349 (incl. the "last line without newline is dropped" quirk, :243)
//   GenomicRegionSet format sniffing and header skipping   gtools/genomic_intervals.cpp:3713-3759
//   GenomicRegionBED::Read / GenomicRegion::Read (REG) / GenomicRegionGFF::Read   :2157-2182, :805-838, :3501-3517
// Written from their behaviour; no reference code is reused.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <zlib.h>
#include <map>
#include <string>
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

 private:
  bool Fill();
  gzFile gz_ = nullptr;
  FILE *fp_ = nullptr;
  std::vector<char> buf_;
  size_t begin_ = 0, end_ = 0;
  bool eof_ = false;
  long line_no_ = 0;
};

// ---------------------------------------------------------------------------------------------
// region sets
// ---------------------------------------------------------------------------------------------
struct ChromTable {
  std::unordered_map<std::string, int32_t> id;
  std::vector<std::string> name;
  int32_t Get(const char *chrom);
};

struct RegionBatch {                                  // packed SoA, regions in file order
  std::vector<int32_t> chrom, start, stop;
  std::vector<int8_t> strand;
  std::vector<int32_t> weight;                        // per region (only if weights are in use)
  std::vector<int64_t> offset;                        // per region + 1; maintained always, passed on only if some region is multi-interval
  std::vector<std::string> label;                     // per region (only if keep_labels)
  std::vector<long> line;                             // per region: source line number
  bool multi = false;
  int64_t n_regions() const { return (int64_t)offset.size() - 1; }
  void Clear();
};

class RegionReader {
 public:
  RegionReader(const char *path, ChromTable *chroms, bool keep_labels, long max_label_value);
  const std::string &format() const { return format_; }
  // parses up to max_regions regions into `out` (cleared first); returns the number parsed (0 at end)
  int64_t Read(RegionBatch *out, int64_t max_regions);
  // sortedness check state for -S (IsBefore, genomic_intervals.cpp:396-401): call per region in order
 private:
  void ParseLine(char *line, long line_no, RegionBatch *out);
  void Push(RegionBatch *out, const char *chrom, char strand, long start, long stop, long line_no);
  LineReader reader_;
  ChromTable *chroms_;
  bool keep_labels_;
  long max_label_value_;
  std::string format_;                                // BED REG GFF SAM EMPTY ""
  char *pending_ = nullptr;                           // first data line, already read during format detection
};

char ProcessStrand(const char *token);                // '1','+' -> '+'; '-1','-' -> '-'; '.' -> '+'; else fatal  (genomic_intervals.cpp:5956-5962)
bool RegionWellFormed(const RegionBatch &b, int64_t k);

// -S sortedness checker (GenomicRegion::IsBefore on consecutive regions)
struct SortChecker {
  bool by_strand = false;
  bool have = false;
  std::string chrom; char strand = 0; long start = 0;
  // returns false if region (chrom,strand,start) sorts before the previous one
  bool Accept(const std::string &c, char s, long st);
};

}  // namespace gt
