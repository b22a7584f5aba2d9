// genomic_regions -- drop-in driver for `genomic_regions gsort` on the B200 engine.
//
// The reference's global sort (gtools/genomic_regions.cpp:421-427, :529-532, :742; GenomicRegionSet::RunGlobalSort,
// genomic_intervals.cpp:4547-4570) loads the whole set, sorts the intervals inside every region (r->Sort(), :6055-6060), bins the
// regions by chromosome / strand / start and list::sort()s every bin by (start ascending, stop descending).  Here the host reads
// and keeps the lines, the device radix-sorts the keys (gtb_sort_regions) and the host prints the regions in that order, each
// the way GenomicRegion*::Print writes its format.  -b (the reference's bucket width) is accepted and has no effect on the order.
// The other operations of genomic_regions are outside the accelerated path and refuse to run.
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <iostream>
#include <numeric>
#include <thread>
#include "gt_host.h"
#include "gtb200.h"

static const char *PROGRAM = "genomic_regions";
static const char *VERSION = "genomic_tools 2.8.1a";
static bool HELP, VERBOSE, SORTED_BY_STRAND;
static long BIN_BITS;

int main(int argc, char *argv[]) {
  gt::CmdLine cmd(PROGRAM, VERSION);
  cmd.AddOperation("gsort", "[OPTIONS] <REGION-SET>", "Global sort: sorts the entire region set.",
                   "* Input formats: REG, GFF, BED, SAM\n  * Operand: region-set\n  * Region requirements: none\n  * Region-set requirements: none");
  if (argc < 2) { cmd.OperationSummary("OPERATION [OPTIONS] <REGION-SET>", "Performs operations on genomic regions (this build: gsort)."); exit(1); }
  std::string op = argv[1];
  if (op[0] == '-') op = op.substr(1);
  if (op != "gsort") {
    std::cerr << "Operation '" << op << "' is outside the GPU-accelerated path of this build (gsort is available)!\n";
    exit(1);
  }
  cmd.SetCurrentOperation(op);
  cmd.AddOption("--help", &HELP, false, "help");
  cmd.AddOption("-h", &HELP, false, "help");
  cmd.AddOption("-v", &VERBOSE, false, "verbose mode");
  cmd.AddOption("-s", &SORTED_BY_STRAND, false, "sort by strand in addition to chromosome and start position");
  cmd.AddOption("-b", &BIN_BITS, 12L, "bucket size (in bits) used in bucket sort");
  const int next_arg = cmd.Read(argv + 1, argc - 1) + 1;
  if (HELP) { cmd.OperationUsage(); exit(1); }
  const char *file = next_arg == argc ? nullptr : argv[next_arg];

  gt::PhaseTimer timer;
  // the CUDA context comes up while the file is read
  setenv("CUDA_VISIBLE_DEVICES", "0", 0);
  static std::thread ctx_thread;
  static gtb_ctx *ctx = nullptr;
  static int ctx_rc = GTB_OK;
  ctx_thread = std::thread([] { ctx_rc = gtb_ctx_create(0, &ctx); });
  gt::exit_hook = [] { if (ctx_thread.joinable()) ctx_thread.join(); };

  gt::ChromTable chroms;
  gt::RegionReader rr(file, &chroms, true, 1);
  fwrite(rr.header().data(), 1, rr.header().size(), stdout);           // the set is opened with hide_header == false
  gt::RegionBatch b;
  std::vector<std::string> raw;
  rr.ReadKeep(&b, &raw, INT64_MAX);
  if (rr.failed()) rr.Fail();
  const int64_t n = b.n_regions();
  if (VERBOSE) std::cerr << "Reading from '" << (file ? file : "<standard input>") << "'; number of regions = " << n << "; format = " << rr.format() << "\n";
  timer.Mark("load");

  // r->Sort(): the intervals of a region by chromosome name, strand byte, start (CompareGenomicIntervals, :6055-6060)
  for (int64_t k = 0; k < n; k++) {
    const int64_t lo = b.offset[k], hi = b.offset[k + 1];
    if (hi - lo < 2) continue;
    std::vector<int64_t> ord((size_t)(hi - lo));
    std::iota(ord.begin(), ord.end(), lo);
    std::sort(ord.begin(), ord.end(), [&](int64_t x, int64_t y) {
      const int c = b.chrom[x] == b.chrom[y] ? 0 : strcmp(chroms.name[b.chrom[x]].c_str(), chroms.name[b.chrom[y]].c_str());
      if (c != 0) return c < 0;
      if (b.strand[x] != b.strand[y]) return (char)b.strand[x] < (char)b.strand[y];
      return b.start[x] < b.start[y];
    });
    std::vector<int32_t> c2, s2, e2; std::vector<int8_t> t2;
    for (int64_t i : ord) { c2.push_back(b.chrom[i]); s2.push_back(b.start[i]); e2.push_back(b.stop[i]); t2.push_back(b.strand[i]); }
    for (int64_t i = lo; i < hi; i++) { b.chrom[i] = c2[(size_t)(i - lo)]; b.start[i] = s2[(size_t)(i - lo)]; b.stop[i] = e2[(size_t)(i - lo)]; b.strand[i] = t2[(size_t)(i - lo)]; }
  }
  // chromosome ranks: strcmp order of the names = the order the reference's std::map iterates in
  std::vector<int32_t> by_name((size_t)chroms.name.size());
  std::iota(by_name.begin(), by_name.end(), 0);
  std::sort(by_name.begin(), by_name.end(), [&](int32_t x, int32_t y) { return strcmp(chroms.name[x].c_str(), chroms.name[y].c_str()) < 0; });
  std::vector<int32_t> rank((size_t)chroms.name.size());
  for (size_t r = 0; r < by_name.size(); r++) rank[(size_t)by_name[r]] = (int32_t)r;
  std::vector<int32_t> k_chrom((size_t)n), k_start((size_t)n), k_stop((size_t)n);
  std::vector<int8_t> k_strand((size_t)n);
  for (int64_t k = 0; k < n; k++) {
    const int64_t lo = b.offset[k], hi = b.offset[k + 1];
    k_chrom[(size_t)k] = rank[(size_t)b.chrom[lo]]; k_start[(size_t)k] = b.start[lo]; k_stop[(size_t)k] = b.stop[hi - 1]; k_strand[(size_t)k] = b.strand[lo];
  }
  timer.Mark("keys");
  if (ctx_thread.joinable()) ctx_thread.join();
  gt::exit_hook = nullptr;
  if (ctx_rc != GTB_OK) { fprintf(stderr, "\nError: no CUDA device available (status %d); this build has no CPU fallback\n", ctx_rc); exit(1); }
  timer.Mark("cuda_context");
  std::vector<int64_t> perm((size_t)std::max<int64_t>(n, 1));
  const int rc = gtb_sort_regions(ctx, n, k_chrom.data(), k_start.data(), k_stop.data(), k_strand.data(), SORTED_BY_STRAND ? 1 : 0, perm.data());
  if (rc != GTB_OK) { fprintf(stderr, "\nError: [gtb_sort_regions] %s (status %d)\n", gtb_ctx_last_error(ctx), rc); exit(1); }
  timer.Mark("sort");
  std::string text;
  for (int64_t k = 0; k < n; k++) {
    gt::PrintRegion(rr.format(), raw[(size_t)perm[(size_t)k]], b, perm[(size_t)k], chroms, &text);
    if (text.size() > (1u << 24)) { fwrite(text.data(), 1, text.size(), stdout); text.clear(); }
  }
  fwrite(text.data(), 1, text.size(), stdout);
  fflush(stdout);
  timer.Mark("print");
  return 0;
}
