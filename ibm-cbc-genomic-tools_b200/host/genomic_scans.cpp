// genomic_scans -- drop-in driver for `genomic_scans counts` on the B200 engine.
//
// Same operation, flags, defaults, usage text, stdout format and exit codes as the reference driver
// (gtools/genomic_scans.cpp:73-150 flags, :399-436 RunCounts, :449-462 main).  The scanner objects
// (UnsortedGenomicRegionSetScanner / SortedGenomicRegionSetScanner, genomic_intervals.cpp:4877-5141) are
// replaced by gtb_scan_* of include/gtb200.h; the `-r` window filter (GenomicRegionSetIndex::GetOverlap,
// :5528-5538, or the sorted merge of Scanner::Next(GenomicRegionSet*), :5146-5165) runs on the host over the
// windows that survived the device-side `-min` compaction -- the two filters commute.
// `peaks` (PeakFinder, genomic_scans.cpp:215-380): two scanners walked in step, tail probabilities per window on the device
// (gtb_scan_peaks; the reference calls GSL, so the printed p-values agree to the printed digits but are not bit-exact), the
// q-value pass of ComputeQValues (:164-209) on the host.  The third input (a uniqueness track) is not supported.
#include <stdlib.h>
#include <string.h>
#include <limits.h>
#include <time.h>
#include <unistd.h>
#include <algorithm>
#include <iostream>
#include <thread>
#include "gt_host.h"
#include "gtb200.h"

static const char *PROGRAM = "genomic_scans";
static const char *VERSION = "genomic_tools 2.8.1a";

static bool HELP, VERBOSE, SORTED, REF_SORTED, IGNORE_STRAND;
static const char *GENOME_REG_FILE, *REF_REG_FILE;
static char PREPROCESS;
static long MAX_LABEL_VALUE, WIN_SIZE, WIN_DIST, MIN_READS;
static bool NORM, COMPARE, PRINT_DETAILS;
static const char *METHOD;
static double PVAL_CUTOFF, QVAL_CUTOFF;

static const char *DETAILS =
    "* Input formats: REG, GFF, BED, SAM\n"
    "  * Operand: interval\n"
    "  * Region requirements: single-interval if -S option is set\n"
    "  * Region-set requirements: sorted by chromosome/strand/start if -S option is set";

static void check(gtb_ctx *ctx, int rc, const char *what) {
  if (rc == GTB_OK) return;
  fprintf(stderr, "\nError: [%s] %s (status %d)\n", what, ctx ? gtb_ctx_last_error(ctx) : "", rc);
  exit(1);
}

// ReadBounds, genomic_intervals.cpp:5997-6015: chromosome -> STOP of its single interval
static std::map<std::string, long> ReadBounds(const char *genome_reg_file) {
  if (genome_reg_file == nullptr || strlen(genome_reg_file) == 0) {
    std::cerr << "Error: genome region file is necessary for this operation!\n";
    exit(1);
  }
  std::map<std::string, long> bounds;
  gt::ChromTable chroms;
  gt::RegionBatch b;
  gt::RegionReader rr(genome_reg_file, &chroms, true, 1);
  rr.ReadAll(&b);
  for (int64_t k = 0; k < b.n_regions(); k++) {
    if (b.offset[k + 1] - b.offset[k] != 1) {
      std::cerr << "label = " << b.label[k] << '\n';
      gt::die_line(b.line(k), "genome regions should be single-interval regions!\n");
    }
    const std::string &chr = chroms.name[b.chrom[b.offset[k]]];
    const long stop = b.stop[b.offset[k]];
    auto it = bounds.find(chr);
    if (it == bounds.end()) bounds[chr] = stop;
    else if (it->second != stop) {
      std::cerr << "Error: chromosome " << chr << " has multiple lengths in genome file '" << genome_reg_file << "' line " << (k + 1) << "!\n";
      exit(1);
    }
  }
  return bounds;
}

// The `-r` filter: does window [a,b] on (chrom, strand) overlap a block of any indexable reference region?
// (GenomicRegionSetIndex: regions whose span has start > stop or stop <= 0 are not indexed, :5363, :5412;
//  GetOverlap(w, match_gaps=false, ignore_strand): some block overlaps w and, unless -i, the region's first
//  strand equals the window's, :5528-5538.)
struct RefFilter {
  struct Track { std::vector<long> start, max_stop; };
  std::map<std::pair<int32_t, char>, Track> tracks;
  bool ignore_strand = false;
  void Build(const gt::RegionBatch &ref, bool ignore) {
    ignore_strand = ignore;
    std::map<std::pair<int32_t, char>, std::vector<std::pair<long, long>>> tmp;
    for (int64_t k = 0; k < ref.n_regions(); k++) {
      const int64_t lo = ref.offset[k], hi = ref.offset[k + 1];
      if (!gt::RegionWellFormed(ref, k)) gt::die_line(ref.line(k), "index regions should be compatible, sorted and non-overlapping!");
      const long s = ref.start[lo], e = ref.stop[hi - 1];
      if (s > e || e <= 0) continue;
      const char strand = ignore ? '+' : (char)ref.strand[lo];
      for (int64_t i = lo; i < hi; i++) tmp[{ref.chrom[i], strand}].push_back({(long)ref.start[i], (long)ref.stop[i]});
    }
    for (auto &kv : tmp) {
      std::sort(kv.second.begin(), kv.second.end());
      Track &t = tracks[kv.first];
      long mx = LONG_MIN;
      for (auto &p : kv.second) { mx = std::max(mx, p.second); t.start.push_back(p.first); t.max_stop.push_back(mx); }
    }
  }
  bool Overlaps(int32_t chrom, char strand, long a, long b) const {
    auto it = tracks.find({chrom, ignore_strand ? '+' : strand});
    if (it == tracks.end()) return false;
    const Track &t = it->second;
    const size_t n = (size_t)(std::upper_bound(t.start.begin(), t.start.end(), b) - t.start.begin());   // blocks with start <= b
    return n > 0 && t.max_stop[n - 1] >= a;
  }
};

// CUDA context creation off the main thread (see gt_host.h: exit_hook)
static std::thread g_ctx_thread;
static gtb_ctx *g_ctx = nullptr;
static int g_ctx_rc = GTB_OK;
static void start_context() {
  // The driver uses one GPU.  On a multi-GPU host the CUDA runtime would initialise every visible device first (seconds);
  // unless the user has chosen devices, only the first one is made visible.
  setenv("CUDA_VISIBLE_DEVICES", "0", 0);
  g_ctx_thread = std::thread([] { g_ctx_rc = gtb_ctx_create(0, &g_ctx); });
  gt::exit_hook = [] { if (g_ctx_thread.joinable()) g_ctx_thread.join(); };
}
static gtb_ctx *wait_context() {
  if (g_ctx_thread.joinable()) g_ctx_thread.join();
  gt::exit_hook = nullptr;
  if (g_ctx_rc != GTB_OK) { fprintf(stderr, "\nError: no CUDA device available (status %d); this build has no CPU fallback\n", g_ctx_rc); exit(1); }
  return g_ctx;
}

// Streams a read file into a scanner: chunks of 4 Mi regions, the parse of chunk k+1 overlapping the device work of chunk k.
// Under -S the reads must be single-interval and sorted (SortedGenomicRegionSetScanner).  Returns the sum of the regions' label
// values (CountGenomicRegions, genomic_intervals.cpp:6206-6214: the number of regions unless --max-label-value is in use).
static long stream_reads(gtb_ctx *ctx, gtb_scan *scan, gt::RegionReader &reads, gt::ChromTable &chroms) {
  const int64_t CHUNK = 4 << 20;
  gt::RegionBatch chunk[2];
  gt::SortChecker sc; sc.by_strand = !IGNORE_STRAND;
  long total = 0;
  for (int which = 0;; which ^= 1) {
    gt::RegionBatch &b = chunk[which];
    check(ctx, gtb_ctx_synchronize(ctx), "gtb_ctx_synchronize");
    if (reads.Read(&b, CHUNK) == 0) break;
    if (SORTED)
      for (int64_t k = 0; k < b.n_regions(); k++) {
        const int64_t i = b.offset[k];
        if (b.offset[k + 1] - i != 1) gt::die_line(b.line(k), "single-interval regions expected for this operation!\n");
        if (!sc.Accept(chroms.name[b.chrom[i]], (char)b.strand[i], b.start[i]))
          gt::die_line(b.line(k), std::string("input regions are not sorted (sorted-by-strand = ") + (IGNORE_STRAND ? "false" : "true") + ")!");
      }
    if (b.weight.empty()) total += (long)b.n_regions();
    else for (int64_t k = 0; k < b.n_regions(); k++) total += b.weight[k];
    gtb_set s;
    s.n_regions = b.n_regions(); s.n_intervals = (int64_t)b.chrom.size();
    s.chrom = b.chrom.data(); s.start = b.start.data(); s.stop = b.stop.data(); s.strand = b.strand.data();
    s.weight = b.weight.empty() ? nullptr : b.weight.data();
    s.region_offset = b.multi ? b.offset.data() : nullptr;
    check(ctx, gtb_scan_add_reads(scan, &s, GTB_MEM_HOST), "gtb_scan_add_reads");
  }
  if (reads.failed()) reads.Fail();
  return total;
}

// ComputeQValues, genomic_scans.cpp:164-209 (one permutation): the p-value below which the q-value stays within the cutoff, or -1
static double ComputeQValues(std::vector<double> pval, std::vector<double> pval_rnd, long n_permutations, double qval_cutoff) {
  std::sort(pval.begin(), pval.end());
  std::sort(pval_rnd.begin(), pval_rnd.end());
  const long n = (long)pval.size();
  if (n == 0) return -1.0;
  std::vector<unsigned long> counts((size_t)n, 0);
  {
    long k = 0;
    size_t i = 0;
    for (size_t j = 0; i != pval.size() && j != pval_rnd.size(); j++) {
      while (i != pval.size() && pval_rnd[j] > pval[i]) { i++; k++; }
      if (k < n - 1) counts[(size_t)k]++;
    }
  }
  std::vector<double> q((size_t)n);
  for (long k = 0; k < n; k++) {
    q[(size_t)k] = (float)counts[(size_t)k] / n_permutations / (k + 1);
    if (k + 1 == n) break;
    counts[(size_t)k + 1] += counts[(size_t)k];
  }
  float min_q = (float)q[(size_t)n - 1];
  double pval_cutoff = -1.0;
  for (long k = n - 2; k >= 0; k--) {
    if (min_q <= qval_cutoff) { pval_cutoff = pval[(size_t)k + 1]; break; }
    if (q[(size_t)k] > min_q) q[(size_t)k] = min_q; else min_q = (float)q[(size_t)k];
  }
  return pval_cutoff;
}

static int run_peaks(const char *signal_file, const char *control_file, const char *uniq_file);

static int driver_main(int argc, char *argv[]) {
  gt::CmdLine cmd(PROGRAM, VERSION);
  cmd.AddOperation("counts", "[OPTIONS] <REG-FILE>", "Determines input read counts in sliding windows of reference regions.", DETAILS);
  cmd.AddOperation("peaks", "[OPTIONS] SIGNAL-REG-FILE [CONTROL-REG-FILE [GENOME-UNIQ-REG-FILE]]", "Scans input reads to identify peaks.", DETAILS);
  if (argc < 2) { cmd.OperationSummary("OPERATION [OPTIONS] INPUT-FILES", "Performs whole-genome scanning operations."); exit(1); }
  std::string op = argv[1];
  if (op[0] == '-') op = op.substr(1);
  cmd.SetCurrentOperation(op);
  cmd.AddOption("--help", &HELP, false, "help");
  cmd.AddOption("-h", &HELP, false, "help");
  cmd.AddOption("-v", &VERBOSE, false, "verbose mode");
  if (op == "counts") {
    cmd.AddOption("-S", &SORTED, false, "input regions are sorted");
    cmd.AddOption("-Sref", &REF_SORTED, false, "reference regions (option -r) are sorted");
    cmd.AddOption("-g", &GENOME_REG_FILE, "", "genome region file");
    cmd.AddOption("-r", &REF_REG_FILE, "", "reference region file");
    cmd.AddOption("-i", &IGNORE_STRAND, false, "ignore strand information");
    cmd.AddOption("-op", &PREPROCESS, '1', "preprocess operator (1=start, c=center, p=all points)");
    cmd.AddOption("--max-label-value", &MAX_LABEL_VALUE, 1L, "maximum region label value to be used");
    cmd.AddOption("-w", &WIN_SIZE, 500L, "window size (must be a multiple of window distance)");
    cmd.AddOption("-d", &WIN_DIST, 25L, "window distance");
    cmd.AddOption("-min", &MIN_READS, 10L, "minimum reads in window");
  } else if (op == "peaks") {
    cmd.AddOption("-S", &SORTED, false, "input regions are sorted");
    cmd.AddOption("-g", &GENOME_REG_FILE, "genome.reg+", "genome region file");
    cmd.AddOption("-i", &IGNORE_STRAND, false, "ignore strand information");
    cmd.AddOption("--max-label-value", &MAX_LABEL_VALUE, 1L, "maximum region label value to be used");
    cmd.AddOption("-w", &WIN_SIZE, 500L, "window size (must be a multiple of window distance)");
    cmd.AddOption("-d", &WIN_DIST, 25L, "window distance");
    cmd.AddOption("-min", &MIN_READS, 10L, "minimum reads in window");
    cmd.AddOption("-M", &METHOD, "binomial", "method (binomial, poisson)");
    cmd.AddOption("-norm", &NORM, false, "equalize background probabilities");
    cmd.AddOption("-cmp", &COMPARE, false, "compare signal to control window");
    cmd.AddOption("-pval", &PVAL_CUTOFF, 1.0, "pvalue cutoff");
    cmd.AddOption("-qval", &QVAL_CUTOFF, 0.05, "qvalue cutoff");
    cmd.AddOption("-D", &PRINT_DETAILS, false, "print details");
  } else {
    std::cerr << "Unknown operation '" << op << "'!\n";
    exit(1);
  }
  const int next_arg = cmd.Read(argv + 1, argc - 1) + 1;
  if (HELP || (op == "peaks" && argc - next_arg < 1)) { cmd.OperationUsage(); exit(1); }
  if (op == "peaks")
    return run_peaks(argv[next_arg], next_arg + 1 < argc ? argv[next_arg + 1] : nullptr, next_arg + 2 < argc ? argv[next_arg + 2] : nullptr);
  const char *input_file = next_arg == argc ? nullptr : argv[next_arg];

  gt::PhaseTimer timer;
  start_context();                                                      // comes up while the genome file and the first reads are read
  // ---- genome bounds; chromosome ids in strcmp order of the names = the reference's std::map order
  std::map<std::string, long> bounds = ReadBounds(GENOME_REG_FILE);
  gt::ChromTable chroms;
  std::vector<int64_t> bound;
  for (auto &kv : bounds) { chroms.Get(kv.first.c_str()); bound.push_back(kv.second); }
  const int32_t n_genome = (int32_t)bound.size();

  gt::RegionReader reads(input_file, &chroms, false, MAX_LABEL_VALUE);
  if (VERBOSE) std::cerr << "Reading from '" << (input_file ? input_file : "<standard input>") << "'; format = " << reads.format() << "\n";
  if (reads.format() == "SEQ") { std::cerr << "Error: this operation does not accept SEQ format!\n"; exit(1); }
  if (reads.format() == "EMPTY") exit(0);                               // scanner ctor returns early, Next() yields nothing
  if (WIN_DIST <= 0 || WIN_SIZE % WIN_DIST != 0) {
    std::cerr << "Error: window size must be a multiple of window step in 'GenomicRegionSetScanner'!\n";
    exit(1);
  }
  if (SORTED && PREPROCESS != '1') {
    // the sorted scanner knows '1' and 'p' only (genomic_intervals.cpp:4939-4946); its 'p' branch skips every other read (a
    // reference defect) and is not reproduced
    fprintf(stderr, "Error: [SortedGenomicRegionSetScanner] preprocess operator '%c' not supported!\n", PREPROCESS);
    exit(1);
  }
  if (!SORTED && PREPROCESS != '1' && PREPROCESS != 'c') {
    std::cerr << "Error: preprocess operator '" << PREPROCESS << "' not supported!\n";      // :5046
    exit(1);
  }

  timer.Mark("setup");
  gtb_ctx *ctx = wait_context();
  timer.Mark("cuda_context");
  gtb_scan_params prm;
  memset(&prm, 0, sizeof prm);
  prm.win_step = WIN_DIST; prm.win_size = WIN_SIZE; prm.min_reads = MIN_READS;
  prm.op = PREPROCESS; prm.ignore_strand = IGNORE_STRAND ? 1 : 0; prm.emulate_sorted = SORTED ? 1 : 0;
  gtb_scan *scan = nullptr;
  check(ctx, gtb_scan_create(ctx, n_genome, bound.data(), &prm, &scan), "gtb_scan_create");

  timer.Mark("scan_create");
  // ---- reads: streamed in chunks, parse of chunk k+1 overlaps the device work of chunk k
  stream_reads(ctx, scan, reads, chroms);
  timer.Mark("stream_reads");
  int64_t n_windows = 0;
  check(ctx, gtb_scan_finish(scan, &n_windows), "gtb_scan_finish");
  timer.Mark("finish");

  // ---- optional reference filter
  RefFilter filter;
  const bool use_filter = REF_REG_FILE != nullptr && strlen(REF_REG_FILE) > 0;
  if (use_filter) {
    gt::RegionBatch ref;
    gt::RegionReader rr(REF_REG_FILE, &chroms, false, 1);
    rr.ReadAll(&ref);
    if (REF_SORTED) {
      gt::SortChecker rs; rs.by_strand = !IGNORE_STRAND;
      for (int64_t k = 0; k < ref.n_regions(); k++) {
        const int64_t i = ref.offset[k];
        if (!rs.Accept(chroms.name[ref.chrom[i]], (char)ref.strand[i], ref.start[i]))
          gt::die_line(ref.line(k), std::string("input regions are not sorted (sorted-by-strand = ") + (IGNORE_STRAND ? "false" : "true") + ")!");
      }
    }
    filter.Build(ref, IGNORE_STRAND);
  }

  // ---- output (genomic_scans.cpp:421-428; PrintInterval genomic_intervals.cpp:5109-5112)
  const int64_t FETCH = 1 << 22;
  std::vector<int32_t> o_chrom((size_t)std::min(n_windows, FETCH) + 1);
  std::vector<int8_t> o_strand(o_chrom.size());
  std::vector<int64_t> o_win(o_chrom.size()), o_val(o_chrom.size());
  std::vector<char> text;
  text.reserve(1 << 24);
  for (int64_t first = 0; first < n_windows; first += FETCH) {
    const int64_t cnt = std::min(FETCH, n_windows - first);
    check(ctx, gtb_scan_fetch(scan, first, cnt, o_chrom.data(), o_strand.data(), o_win.data(), o_val.data()), "gtb_scan_fetch");
    for (int64_t j = 0; j < cnt; j++) {
      const long a = WIN_DIST * (o_win[j] - 1) + 1, b = WIN_DIST * (o_win[j] - 1) + WIN_SIZE;
      if (use_filter && !filter.Overlaps(o_chrom[j], (char)o_strand[j], a, b)) continue;
      // value TAB chromosome SPACE strand SPACE start SPACE stop: the numbers go through a fixed buffer, the name (any length) is appended as it is
      const std::string &name = chroms.name[o_chrom[j]];
      char num[96];
      int len = snprintf(num, sizeof num, "%ld\t", (long)o_val[j]);
      text.insert(text.end(), num, num + len);
      text.insert(text.end(), name.begin(), name.end());
      len = snprintf(num, sizeof num, " %c %ld %ld\n", (char)o_strand[j], a, b);
      text.insert(text.end(), num, num + len);
      if (text.size() > (1u << 24) - 512) { fwrite(text.data(), 1, text.size(), stdout); text.clear(); }
    }
  }
  if (!text.empty()) fwrite(text.data(), 1, text.size(), stdout);
  fflush(stdout);
  timer.Mark("print");
  // the process is about to end: the driver reclaims the window table (1 GB for hg19) faster than cudaFree would
  (void)scan;
  return 0;
}

// PeakFinder (genomic_scans.cpp:215-380)
static int run_peaks(const char *signal_file, const char *control_file, const char *uniq_file) {
  if (uniq_file != nullptr) { std::cerr << "Error: a genome uniqueness file (third input of 'peaks') is not supported by this build!\n"; exit(1); }
  int method;
  if (strcmp(METHOD, "binomial") == 0) method = GTB_PEAKS_BINOMIAL;
  else if (strcmp(METHOD, "poisson") == 0) method = GTB_PEAKS_POISSON;
  else if (COMPARE && strcmp(METHOD, "binomial2") == 0) method = GTB_PEAKS_BINOMIAL2;
  else if (COMPARE && strcmp(METHOD, "cbinomial") == 0) method = GTB_PEAKS_CBINOMIAL;
  else if (COMPARE && strcmp(METHOD, "normal") == 0) method = GTB_PEAKS_NORMAL;
  else method = -1;                                                      // fatal at the first window that qualifies (:337, :349)
  start_context();
  const char preprocess = SORTED ? '1' : 'c';                             // :232
  std::map<std::string, long> bounds = ReadBounds(GENOME_REG_FILE);
  gt::ChromTable chroms;
  std::vector<int64_t> bound;
  unsigned long effective_genome_size = 0;
  for (auto &kv : bounds) { chroms.Get(kv.first.c_str()); bound.push_back(kv.second); effective_genome_size += (unsigned long)kv.second; }
  fprintf(stderr, "* Effective genome size = %lu\n", effective_genome_size);
  if (WIN_DIST <= 0 || WIN_SIZE % WIN_DIST != 0) {
    std::cerr << "Error: window size must be a multiple of window step in 'GenomicRegionSetScanner'!\n";
    exit(1);
  }
  gtb_ctx *ctx = wait_context();
  gtb_scan_params prm;
  memset(&prm, 0, sizeof prm);
  prm.win_step = WIN_DIST; prm.win_size = WIN_SIZE; prm.min_reads = MIN_READS;
  prm.op = preprocess; prm.ignore_strand = IGNORE_STRAND ? 1 : 0; prm.emulate_sorted = SORTED ? 1 : 0;
  gtb_scan *signal = nullptr, *control = nullptr;
  long n_signal_reads = 0, n_control_reads = 0;
  {
    gt::RegionReader reads(signal_file, &chroms, false, MAX_LABEL_VALUE);
    if (reads.format() == "SEQ") { std::cerr << "Error: this operation does not accept SEQ format!\n"; exit(1); }
    check(ctx, gtb_scan_create(ctx, (int32_t)bound.size(), bound.data(), &prm, &signal), "gtb_scan_create");
    if (reads.format() != "EMPTY") n_signal_reads = stream_reads(ctx, signal, reads, chroms);
  }
  const double p_signal = (double)n_signal_reads / effective_genome_size;
  double p_control = p_signal;
  if (control_file != nullptr) {
    gt::RegionReader reads(control_file, &chroms, false, MAX_LABEL_VALUE);
    if (reads.format() == "SEQ") { std::cerr << "Error: this operation does not accept SEQ format!\n"; exit(1); }
    check(ctx, gtb_scan_create(ctx, (int32_t)bound.size(), bound.data(), &prm, &control), "gtb_scan_create");
    if (reads.format() != "EMPTY") n_control_reads = stream_reads(ctx, control, reads, chroms);
    p_control = (double)n_control_reads / effective_genome_size;
  } else {
    n_control_reads = n_signal_reads;
  }
  const double p_ratio = p_signal / p_control;
  fprintf(stderr, "* Signal input file = %s (reads = %lu; background probability = %.2e)\n", signal_file, (unsigned long)n_signal_reads, p_signal);
  fprintf(stderr, "* Control input file = %s (reads = %lu; background probability = %.2e)\n", control_file ? control_file : "(null)", (unsigned long)n_control_reads, p_control);
  fprintf(stderr, "* Signal/Control background probability = %f\n", p_ratio);

  gtb_peaks_params pp;
  memset(&pp, 0, sizeof pp);
  pp.method = method < 0 ? GTB_PEAKS_BINOMIAL : method; pp.compare = COMPARE ? 1 : 0; pp.norm = NORM ? 1 : 0; pp.min_reads = MIN_READS;
  pp.n_signal_reads = n_signal_reads; pp.n_control_reads = n_control_reads; pp.p_signal = p_signal; pp.p_control = p_control;
  pp.pval_cutoff = method < 0 ? 2.0 : PVAL_CUTOFF;
  const char *seed_env = getenv("GT_SEED");
  pp.seed = seed_env ? strtoull(seed_env, nullptr, 10) : (uint64_t)getpid() + (uint64_t)time(nullptr);   // InitRandomGenerator(getpid()+time(NULL)), :457
  int64_t n = 0;
  check(ctx, gtb_scan_peaks(signal, control, &pp, &n), "gtb_scan_peaks");
  if (method < 0 && n > 0) { std::cerr << "Error: unknown probability distribution!\n"; exit(1); }
  std::vector<int32_t> o_chrom((size_t)n + 1);
  std::vector<int8_t> o_strand((size_t)n + 1);
  std::vector<int64_t> o_win((size_t)n + 1);
  std::vector<double> p1((size_t)n), p2((size_t)n);
  if (n > 0) check(ctx, gtb_scan_peaks_fetch(signal, 0, n, o_chrom.data(), o_strand.data(), o_win.data(), p1.data(), p2.data()), "gtb_scan_peaks_fetch");
  const double pval_cutoff = ComputeQValues(p1, p2, 1, QVAL_CUTOFF);
  std::string text;
  char num[160];
  for (int64_t k = 0; k < n; k++) {
    if (!(p1[(size_t)k] <= pval_cutoff)) continue;
    const long a = WIN_DIST * (o_win[(size_t)k] - 1) + 1, b = WIN_DIST * (o_win[(size_t)k] - 1) + WIN_SIZE;
    int len = snprintf(num, sizeof num, "%.4e\t", p1[(size_t)k]);
    text.append(num, (size_t)len);
    text += chroms.name[o_chrom[(size_t)k]];
    len = snprintf(num, sizeof num, " %c %ld %ld\n", (char)o_strand[(size_t)k], a, b);
    text.append(num, (size_t)len);
    if (text.size() > (1u << 24)) { fwrite(text.data(), 1, text.size(), stdout); text.clear(); }
  }
  fwrite(text.data(), 1, text.size(), stdout);
  fflush(stdout);
  return 0;
}

// the driver's work is done and its output written when driver_main returns: the process leaves through gt::Exit (gt_host.h)
int main(int argc, char *argv[]) {
  exit(driver_main(argc, argv));
}
