// genomic_regions -- drop-in driver for `genomic_regions gsort`, `link`, `inv` and `union` on the B200 engine.
//
// union (GenomicRegionSet::RunUnion, genomic_intervals.cpp:4397-4403; GenomicRegion::Union, :1649-1671): every region's intervals
// sorted by start and merged where they overlap, the region printed in its own format -- a per-line operation, on the host.
// link (GenomicRegionSet::RunGlobalLink, genomic_intervals.cpp:4607-4644): consecutive regions of a sorted stream that are
// compatible and lie within -d of the stop reached so far become one region.  The host reads, checks (single-interval, sorted)
// and prints; which regions begin a linked region and where each one ends comes from the device (gtb_link_regions: a
// prefix-maximum scan; the sequential loop only for a negative -d, which has no scan form).
// inv (RunGlobalInvert, :4576-4601): the gaps between consecutive regions of a (chromosome, strand) run and to the chromosome's
// ends -- a function of adjacent pairs, computed while printing.
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
#include <map>
#include <numeric>
#include <sstream>
#include <thread>
#include "gt_host.h"
#include "gtb200.h"

static const char *PROGRAM = "genomic_regions";
static const char *VERSION = "genomic_tools 2.8.1a";
static bool HELP, VERBOSE, SORTED_BY_STRAND;
static long BIN_BITS, LINK_MAX_DIFFERENCE;
static const char *LINK_LABEL_FUNC, *GENOME_REG_FILE;

static std::thread ctx_thread;
static gtb_ctx *ctx = nullptr;
static int ctx_rc = GTB_OK;
static void start_context() {
  setenv("CUDA_VISIBLE_DEVICES", "0", 0);
  ctx_thread = std::thread([] { ctx_rc = gtb_ctx_create(0, &ctx); });
  gt::exit_hook = [] { if (ctx_thread.joinable()) ctx_thread.join(); };
}
static void wait_context() {
  if (ctx_thread.joinable()) ctx_thread.join();
  gt::exit_hook = nullptr;
  if (ctx_rc != GTB_OK) { fprintf(stderr, "\nError: no CUDA device available (status %d); this build has no CPU fallback\n", ctx_rc); exit(1); }
}
static void region_error(long line, const std::string &msg) {          // GenomicRegion::PrintError, genomic_intervals.cpp:1001-1006
  fflush(stdout);
  fprintf(stderr, "\nError: Line %ld: %s\n", line, msg.c_str());
  exit(1);
}

// ReadBounds, genomic_intervals.cpp:5997-6015
static std::map<std::string, long> ReadBounds(const char *genome_reg_file) {
  if (genome_reg_file == nullptr || strlen(genome_reg_file) == 0) { std::cerr << "Error: genome region file is necessary for this operation!\n"; exit(1); }
  std::map<std::string, long> bounds;
  gt::ChromTable chroms;
  gt::RegionBatch b;
  gt::RegionReader rr(genome_reg_file, &chroms, true, 1);
  rr.ReadAll(&b);
  for (int64_t k = 0; k < b.n_regions(); k++) {
    if (b.offset[k + 1] - b.offset[k] != 1) { std::cerr << "label = " << b.label[k] << '\n'; gt::die_line(b.line(k), "genome regions should be single-interval regions!\n"); }
    const std::string &chr = chroms.name[b.chrom[b.offset[k]]];
    const long stop = b.stop[b.offset[k]];
    auto it = bounds.find(chr);
    if (it == bounds.end()) bounds[chr] = stop;
    else if (it->second != stop) { std::cerr << "Error: chromosome " << chr << " has multiple lengths in genome file '" << genome_reg_file << "' line " << (k + 1) << "!\n"; exit(1); }
  }
  return bounds;
}

// The first region of the set that the reference's streaming loop would die on -- a region that is not single-interval, or one
// that sorts before its predecessor (GenomicRegionSet::Next(sorted_by_strand, ..), :3874-3882; in that order of checks for a
// region) -- or n if there is none.  `within_group_only`: RunGlobalInvert checks the order inside a (chromosome, strand) run only.
static int64_t first_bad_region(const gt::RegionBatch &b, const gt::ChromTable &chroms, bool by_strand, bool within_group_only, std::string *msg) {
  const int64_t n = b.n_regions();
  for (int64_t k = 0; k < n; k++) {
    const int64_t i = b.offset[k];
    // (RunGlobalInvert looks at the number of intervals first, RunGlobalLink's Next() at the order first)
    if (within_group_only && b.offset[k + 1] - i != 1) { *msg = "not a single-interval region!"; return k; }
    if (k > 0 && b.offset[k] - b.offset[k - 1] == 1) {
      const int64_t p = b.offset[k - 1];
      const bool same_group = b.chrom[i] == b.chrom[p] && (!by_strand || b.strand[i] == b.strand[p]);
      if (!within_group_only || same_group) {
        const int c = b.chrom[i] == b.chrom[p] ? 0 : strcmp(chroms.name[b.chrom[i]].c_str(), chroms.name[b.chrom[p]].c_str());   // IsBefore, :396-401
        bool before = c < 0;
        if (c == 0) before = by_strand && b.strand[i] != b.strand[p] ? (char)b.strand[i] < (char)b.strand[p] : b.start[i] < b.start[p];
        if (before) { *msg = std::string("input regions are not sorted (sorted-by-strand = ") + (by_strand ? "true" : "false") + ")!"; return k; }
      }
    }
    if (b.offset[k + 1] - i != 1) { *msg = "not a single-interval region!"; return k; }
  }
  return n;
}

static int run_link(const char *file);
static int run_inv(const char *file);
static int run_union(const char *file);

static int driver_main(int argc, char *argv[]) {
  gt::CmdLine cmd(PROGRAM, VERSION);
  cmd.AddOperation("gsort", "[OPTIONS] <REGION-SET>", "Global sort: sorts the entire region set.",
                   "* Input formats: REG, GFF, BED, SAM\n  * Operand: region-set\n  * Region requirements: none\n  * Region-set requirements: none");
  cmd.AddOperation("inv", "[OPTIONS] <REGION-SET>", "Inverts regions given the genome chromosomal boundaries.",
                   "* Input formats: REG, GFF, BED, SAM\n  * Operand: region-set\n  * Region requirements: single-interval\n  * Region-set requirements: sorted by chromosome/strand/start");
  cmd.AddOperation("link", "[OPTIONS] <REGION-SET>", "Links consecutive regions to produce a non-overlapping set.",
                   "* Input formats: REG, GFF, BED, SAM\n  * Operand: region-set\n  * Region requirements: single-interval\n  * Region-set requirements: sorted by chromosome/(strand)/start");
  cmd.AddOperation("union", "[OPTIONS] <REGION-SET>", "Computes union of region intervals.",
                   "* Input formats: REG, GFF, BED, SAM\n  * Operand: region\n  * Region requirements: chromosome/strand-compatible\n  * Region-set requirements: none");
  if (argc < 2) { cmd.OperationSummary("OPERATION [OPTIONS] <REGION-SET>", "Performs operations on genomic regions (this build: gsort, inv, link, union)."); exit(1); }
  std::string op = argv[1];
  if (op[0] == '-') op = op.substr(1);
  if (op != "gsort" && op != "link" && op != "inv" && op != "union") {
    std::cerr << "Operation '" << op << "' is outside the GPU-accelerated path of this build (gsort, inv, link, union are available)!\n";
    exit(1);
  }
  cmd.SetCurrentOperation(op);
  cmd.AddOption("--help", &HELP, false, "help");
  cmd.AddOption("-h", &HELP, false, "help");
  cmd.AddOption("-v", &VERBOSE, false, "verbose mode");
  if (op == "gsort") {
    cmd.AddOption("-s", &SORTED_BY_STRAND, false, "sort by strand in addition to chromosome and start position");
    cmd.AddOption("-b", &BIN_BITS, 12L, "bucket size (in bits) used in bucket sort");
  } else if (op == "inv") {
    cmd.AddOption("-g", &GENOME_REG_FILE, "", "genome region-set file");
  } else if (op == "link") {
    cmd.AddOption("-s", &SORTED_BY_STRAND, false, "input regions are sorted by strand");
    cmd.AddOption("-d", &LINK_MAX_DIFFERENCE, 0L, "maximum difference between successive regions");
    cmd.AddOption("--label-func", &LINK_LABEL_FUNC, "", "label function = {min,max,sum,%c}, where %c is used as delimiter");
  }
  const int next_arg = cmd.Read(argv + 1, argc - 1) + 1;
  if (HELP) { cmd.OperationUsage(); exit(1); }
  const char *file = next_arg == argc ? nullptr : argv[next_arg];
  if (op == "link") return run_link(file);
  if (op == "inv") return run_inv(file);
  if (op == "union") return run_union(file);

  gt::PhaseTimer timer;
  // the CUDA context comes up while the file is read
  start_context();

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
  wait_context();
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

// genomic_regions link
static int run_link(const char *file) {
  start_context();
  gt::ChromTable chroms;
  gt::RegionReader rr(file, &chroms, true, 1);
  fwrite(rr.header().data(), 1, rr.header().size(), stdout);
  gt::RegionBatch b;
  std::vector<std::string> raw;
  rr.ReadKeep(&b, &raw, INT64_MAX);
  std::string msg;
  const int64_t n_all = b.n_regions();
  const int64_t n = first_bad_region(b, chroms, SORTED_BY_STRAND, false, &msg);     // the regions the reference gets through
  const bool parse_failed = rr.failed();
  const bool is_func = strcmp(LINK_LABEL_FUNC, "min") == 0 || strcmp(LINK_LABEL_FUNC, "max") == 0 || strcmp(LINK_LABEL_FUNC, "sum") == 0;
  const bool use_labels = strlen(LINK_LABEL_FUNC) > 0;
  // which regions begin a linked region, and where each linked region ends
  std::vector<int64_t> head((size_t)std::max<int64_t>(n, 1));
  std::vector<int32_t> lstop((size_t)std::max<int64_t>(n, 1));
  int64_t n_linked = 0;
  if (n > 0 && LINK_MAX_DIFFERENCE >= 0) {
    std::vector<int32_t> group((size_t)n), start((size_t)n), stop((size_t)n);
    int32_t g = 0;
    for (int64_t k = 0; k < n; k++) {
      const int64_t i = b.offset[k];
      if (k > 0) { const int64_t p = b.offset[k - 1]; if (b.chrom[i] != b.chrom[p] || (SORTED_BY_STRAND && b.strand[i] != b.strand[p])) g++; }
      group[(size_t)k] = g; start[(size_t)k] = b.start[i]; stop[(size_t)k] = b.stop[i];
    }
    wait_context();
    const int rc = gtb_link_regions(ctx, n, group.data(), start.data(), stop.data(), LINK_MAX_DIFFERENCE, &n_linked, head.data(), lstop.data());
    if (rc != GTB_OK) { fprintf(stderr, "\nError: [gtb_link_regions] %s (status %d)\n", gtb_ctx_last_error(ctx), rc); exit(1); }
  } else if (n > 0) {
    // a negative -d (successive regions must overlap by that much): the stop reached so far, not the prefix maximum, decides -- the loop as it stands
    long new_stop = 0;
    for (int64_t k = 0; k < n; k++) {
      const int64_t i = b.offset[k], h = n_linked ? b.offset[head[(size_t)n_linked - 1]] : 0;
      const bool joins = n_linked > 0 && b.chrom[i] == b.chrom[h] && (!SORTED_BY_STRAND || b.strand[i] == b.strand[h]) && (long)b.start[i] - new_stop <= LINK_MAX_DIFFERENCE;
      if (joins) new_stop = std::max(new_stop, (long)b.stop[i]);
      else { head[(size_t)n_linked++] = k; new_stop = b.stop[i]; }
      lstop[(size_t)n_linked - 1] = (int32_t)new_stop;
    }
    wait_context();
  } else {
    wait_context();
  }
  // the linked region in progress when the reference dies (or when a malformed line ends the input) is never printed
  const bool cut_short = n < n_all || parse_failed;
  const int64_t n_print = cut_short && n_linked > 0 ? n_linked - 1 : n_linked;
  std::string text;
  for (int64_t j = 0; j < n_print; j++) {
    const int64_t k0 = head[(size_t)j], k1 = j + 1 < n_linked ? head[(size_t)j + 1] : n;
    std::string label = use_labels ? b.label[(size_t)k0] : "_";
    if (use_labels) {
      double val = atof(b.label[(size_t)k0].c_str());
      for (int64_t k = k0 + 1; k < k1; k++) {
        if (!is_func) { label += LINK_LABEL_FUNC; label += b.label[(size_t)k]; }
        else if (strcmp(LINK_LABEL_FUNC, "min") == 0) val = std::min(val, atof(b.label[(size_t)k].c_str()));
        else if (strcmp(LINK_LABEL_FUNC, "max") == 0) val = std::max(val, atof(b.label[(size_t)k].c_str()));
        else val = val + atof(b.label[(size_t)k].c_str());
      }
      if (is_func) { std::ostringstream os; os << val; label = os.str(); }
    }
    gt::PrintModified(rr.format(), raw[(size_t)k0], b, k0, chroms, label.c_str(), (long)b.start[b.offset[k0]], (long)lstop[(size_t)j], &text);
    if (text.size() > (1u << 24)) { fwrite(text.data(), 1, text.size(), stdout); text.clear(); }
  }
  fwrite(text.data(), 1, text.size(), stdout);
  fflush(stdout);
  if (n < n_all) region_error(b.line(n), msg);
  if (parse_failed) rr.Fail();
  return 0;
}

// genomic_regions inv
static int run_inv(const char *file) {
  gt::ChromTable chroms;
  gt::RegionReader rr(file, &chroms, true, 1);
  fwrite(rr.header().data(), 1, rr.header().size(), stdout);
  gt::RegionBatch b;
  std::vector<std::string> raw;
  rr.ReadKeep(&b, &raw, INT64_MAX);
  if (b.n_regions() == 0 && !rr.failed()) return 0;
  std::map<std::string, long> bounds = ReadBounds(GENOME_REG_FILE);
  std::string msg;
  const int64_t n_all = b.n_regions();
  const int64_t n = first_bad_region(b, chroms, true, true, &msg);
  std::string text;
  auto flush = [&](bool force) { if (force || text.size() > (1u << 24)) { fwrite(text.data(), 1, text.size(), stdout); text.clear(); } };
  for (int64_t k = 0; k < n;) {
    const int64_t i0 = b.offset[k];
    const std::string &chr = chroms.name[b.chrom[i0]];
    auto it = bounds.find(chr);
    if (it == bounds.end()) { flush(true); fflush(stdout); fprintf(stderr, "Line %ld: chromosome %s not found!\n", b.line(k) + 1, chr.c_str()); exit(1); }
    const long chrom_size = it->second;
    if (b.start[i0] > 1) gt::PrintModified(rr.format(), raw[(size_t)k], b, k, chroms, "_", 1L, (long)b.start[i0] - 1, &text);
    int64_t prev = k;
    for (k++; k < n; k++) {
      const int64_t i = b.offset[k], p = b.offset[prev];
      if (b.chrom[i] != b.chrom[p] || b.strand[i] != b.strand[p]) break;                   // IsCompatibleWith(r0, ignore_strand = false)
      if ((long)b.start[i] > (long)b.stop[p] + 1) gt::PrintModified(rr.format(), raw[(size_t)k], b, k, chroms, "_", (long)b.stop[p] + 1, (long)b.start[i] - 1, &text);
      prev = k;
      flush(false);
    }
    // the reference dies on region n: if that region continues this run, the run's last gap is never printed
    if (k == n && n < n_all) {
      const int64_t i = b.offset[n], p = b.offset[prev];
      if (b.chrom[i] == b.chrom[p] && b.strand[i] == b.strand[p]) break;
    }
    const long stop = b.stop[b.offset[prev]];
    if (stop + 1 < chrom_size) gt::PrintModified(rr.format(), raw[(size_t)prev], b, prev, chroms, "_", stop + 1, chrom_size, &text);
  }
  flush(true);
  fflush(stdout);
  if (n < n_all) region_error(b.line(n), msg);
  if (rr.failed()) rr.Fail();
  return 0;
}

// genomic_regions union
static int run_union(const char *file) {
  gt::ChromTable chroms;
  gt::RegionReader rr(file, &chroms, true, 1);
  fwrite(rr.header().data(), 1, rr.header().size(), stdout);
  gt::RegionBatch b;
  std::vector<std::string> raw;
  rr.ReadKeep(&b, &raw, INT64_MAX);
  const int64_t n = b.n_regions();
  std::string text;
  std::vector<std::pair<int32_t, int32_t>> iv;
  for (int64_t k = 0; k < n; k++) {
    const int64_t lo = b.offset[k], hi = b.offset[k + 1];
    int64_t m = hi - lo;
    if (m > 1) {                                                         // "if (I.size()<=1) return;", :1651
      for (int64_t i = lo + 1; i < hi; i++)                               // IsCompatible(false), :1116-1122
        if (b.chrom[i] != b.chrom[lo] || b.strand[i] != b.strand[lo]) {
          fwrite(text.data(), 1, text.size(), stdout);
          region_error(b.line(k), "region intervals must have the same chromosome/strand for this operation!");
        }
      iv.clear();
      for (int64_t i = lo; i < hi; i++) iv.emplace_back(b.start[i], b.stop[i]);
      // Sort(): by start (CompareGenomicIntervals, :6055-6060; chromosome and strand are equal here).  What the merge leaves
      // does not depend on the order of intervals with equal starts.
      std::sort(iv.begin(), iv.end(), [](const std::pair<int32_t, int32_t> &x, const std::pair<int32_t, int32_t> &y) { return x.first < y.first; });
      m = 0;
      int32_t start = iv[0].first, stop = iv[0].second;
      for (size_t i = 1; i < iv.size(); i++) {
        if (stop < iv[i].first) { b.start[(size_t)(lo + m)] = start; b.stop[(size_t)(lo + m)] = stop; m++; start = iv[i].first; stop = iv[i].second; }
        else stop = std::max(stop, iv[i].second);
      }
      b.start[(size_t)(lo + m)] = start; b.stop[(size_t)(lo + m)] = stop; m++;
    }
    const int64_t next = b.offset[(size_t)k + 1];
    b.offset[(size_t)k + 1] = lo + m;                                     // (the region as PrintRegion is to see it)
    gt::PrintRegion(rr.format(), raw[(size_t)k], b, k, chroms, &text);
    b.offset[(size_t)k + 1] = next;
    if (text.size() > (1u << 24)) { fwrite(text.data(), 1, text.size(), stdout); text.clear(); }
  }
  fwrite(text.data(), 1, text.size(), stdout);
  fflush(stdout);
  if (rr.failed()) rr.Fail();
  return 0;
}

// the driver's work is done and its output written when driver_main returns: the process leaves through gt::Exit (gt_host.h)
int main(int argc, char *argv[]) {
  exit(driver_main(argc, argv));
}
