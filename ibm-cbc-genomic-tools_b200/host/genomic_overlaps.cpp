// genomic_overlaps -- drop-in driver for `genomic_overlaps count | coverage | density | rpkm` on the B200 engine.
//
// Same operations, flags, defaults, usage text, stdout format and exit codes as the reference driver
// (gtools/genomic_overlaps.cpp:73-261 flags, :408-490 count/coverage/density, :746-775 rpkm); the engine calls
// GenomicRegionSetOverlaps::CountIndexOverlaps / CalcIndexCoverage are replaced by the C ABI of include/gtb200.h.
// `subset` and `overlap` (:782-800, :706-739) are the per-query dual: gtb_index_query_counts + the reference's Print formats.
// The other operations of the reference (annotate, bin, dist, intersect, offset) enumerate pairs and are outside the
// accelerated path: they are listed for the usage text and refuse to run.
#include <stdlib.h>
#include <string.h>
#include <iostream>
#include <thread>
#include "gt_host.h"
#include "gtb200.h"

static const char *PROGRAM = "genomic_overlaps";
static const char *VERSION = "genomic_tools 2.8.1a";

static bool HELP, VERBOSE, IS_SORTED, SORTED_BY_STRAND, IGNORE_STRAND, MATCH_GAPS, MERGE_LABELS, SUBSET_NONOVERLAPS;
static const char *BIN_BITS;
static long MAX_LABEL_VALUE;
static unsigned long MIN_COUNT;
static double MIN_RPKM, MIN_DENSITY;

static const char *DETAILS_REGION =
    "* Input formats: REG, GFF, BED, SAM\n"
    "  * Operands: region, region-set\n"
    "  * Region requirements: chromosome/strand-compatible, sorted, non-overlapping\n"
    "  * Region-set requirements: sorted if -S option is used";

static void check(gtb_ctx *ctx, int rc, const char *what) {
  if (rc == GTB_OK) return;
  fprintf(stderr, "\nError: [%s] %s (status %d)\n", what, ctx ? gtb_ctx_last_error(ctx) : "", rc);
  exit(1);
}

[[noreturn]] static void die_query(int rc, long line) {
  gt::die_line(line, rc == GTB_ERR_QUERY_STOP_NONPOSITIVE ? "stop position must be positive!"
                     : rc == GTB_ERR_QUERY_START_GT_STOP ? "start position cannot be greater than stop position!"
                                                          : "query regions should be compatible, sorted and non-overlapping!");
}

static gtb_set as_set(const gt::RegionBatch &b) {
  gtb_set s;
  s.n_regions = b.n_regions();
  s.n_intervals = (int64_t)b.chrom.size();
  s.chrom = b.chrom.data(); s.start = b.start.data(); s.stop = b.stop.data(); s.strand = b.strand.data();
  s.weight = b.weight.empty() ? nullptr : b.weight.data();
  s.region_offset = b.multi ? b.offset.data() : nullptr;
  return s;
}

// a streamed query batch: regions that all have the same number of intervals (read pairs) go without their offsets, which
// spares the engine the pass that reads them (include/gtb200.h, gtb_set)
static gtb_set as_query_set(const gt::RegionBatch &b) {
  gtb_set s = as_set(b);
  if (s.region_offset && s.n_regions > 0 && s.n_intervals % s.n_regions == 0) {
    const int64_t k = s.n_intervals / s.n_regions;
    bool uniform = k >= 2;
    for (int64_t r = 1; uniform && r < s.n_regions; r++) uniform = b.offset[r] == r * k;
    if (uniform) s.region_offset = nullptr;
  }
  return s;
}

// CUDA context creation off the main thread (see gt_host.h: exit_hook)
static std::thread g_ctx_thread;
static gtb_ctx *g_ctx = nullptr;
static gtb_mgpu *g_mgpu = nullptr;
static int g_gpus = 1;
static std::vector<int> g_gpu_list;
static int g_ctx_rc = GTB_OK;
static void start_context() {
  // The driver uses one GPU.  On a multi-GPU host the CUDA runtime would initialise every visible device first (seconds);
  // unless the user has chosen devices, only the first one is made visible.
  // GTB_GPUS=N in the environment: devices 0..N-1 behind the index (gtb_mgpu_*, query batches cut into one slice per device);
  // GTB_GPUS=a,b,c names the devices (a device may repeat: several contexts on it)
  if (const char *env = getenv("GTB_GPUS")) {
    if (strchr(env, ',')) {
      for (const char *p = env; *p;) { g_gpu_list.push_back(atoi(p)); p = strchr(p, ','); if (!p) break; p++; }
      g_gpus = (int)g_gpu_list.size();
    } else {
      g_gpus = std::max(1, std::min(64, atoi(env)));
    }
  }
  if (g_gpus > 1) {
    g_ctx_thread = std::thread([] {
      g_ctx_rc = gtb_mgpu_create(g_gpus, g_gpu_list.empty() ? nullptr : g_gpu_list.data(), &g_mgpu);
      if (g_ctx_rc == GTB_OK) g_ctx = gtb_mgpu_ctx(g_mgpu, 0);
    });
  } else {
    setenv("CUDA_VISIBLE_DEVICES", "0", 0);
    g_ctx_thread = std::thread([] { g_ctx_rc = gtb_ctx_create(0, &g_ctx); });
  }
  gt::exit_hook = [] { if (g_ctx_thread.joinable()) g_ctx_thread.join(); };
}
static gtb_ctx *wait_context() {
  if (g_ctx_thread.joinable()) g_ctx_thread.join();
  gt::exit_hook = nullptr;
  if (g_ctx_rc != GTB_OK) { fprintf(stderr, "\nError: no CUDA device available (status %d); this build has no CPU fallback\n", g_ctx_rc); exit(1); }
  return g_ctx;
}

static int driver_main(int argc, char *argv[]) {
  gt::CmdLine cmd(PROGRAM, VERSION);
  const char *USAGE = "[OPTIONS] REFERENCE-REGION-FILE <TEST-REGION-FILE>";
  cmd.AddOperation("annotate", USAGE, "Annotates test regions according to reference regions.", "");
  cmd.AddOperation("bin", USAGE, "Finds overlaps of interval pairs with reference regions.", "");
  cmd.AddOperation("count", USAGE, "Counts the number of overlapping test regions per reference region.", DETAILS_REGION);
  cmd.AddOperation("coverage", USAGE, "Calculates the depth coverage (i.e. the total number of overlapping nucleotides) per reference region.", DETAILS_REGION);
  cmd.AddOperation("density", USAGE, "Computes the density (i.e. the coverage divided by the size of the reference region) of overlaps per reference region.", DETAILS_REGION);
  cmd.AddOperation("dist", USAGE, "Computes the distance between a pair of intervals given breakpoints in reference file (e.g. restriction enzyme sites) [UNDER DEVELOPMENT].", "");
  cmd.AddOperation("intersect", USAGE, "Computes the intersection between all pairs of test and reference regions. Results are grouped by test region.", "");
  cmd.AddOperation("offset", USAGE, "Computes the distances of test regions from their overlapping reference regions.", "");
  cmd.AddOperation("overlap", USAGE, "Finds the overlaps between all pairs of test and reference regions. Results are grouped by test region.", DETAILS_REGION);
  cmd.AddOperation("rpkm", USAGE, "Computing reference region RPKM values.", DETAILS_REGION);
  cmd.AddOperation("subset", USAGE, "Picks a subset of test regions depending on their overlap with reference regions. Results are grouped by test region.", DETAILS_REGION);
  if (argc < 2) {
    cmd.OperationSummary("OPERATION [OPTIONS] REFERENCE-REGION-FILE <TEST-REGION-FILE>",
                         "Performs overlap operations between a test and a reference set of genomic regions.");
    exit(1);
  }
  std::string op = argv[1];
  if (op[0] == '-') op = op.substr(1);                                // compatibility with the previous version (genomic_overlaps.cpp:181)
  cmd.SetCurrentOperation(op);
  cmd.AddOption("--help", &HELP, false, "help");
  cmd.AddOption("-h", &HELP, false, "help");
  cmd.AddOption("-v", &VERBOSE, false, "verbose mode");
  cmd.AddOption("-B", &BIN_BITS, "17,20,23,26", "number of shift-bits for each bin level");
  cmd.AddOption("-S", &IS_SORTED, false, "test and reference regions are sorted by chromosome and start position");
  cmd.AddOption("-s", &SORTED_BY_STRAND, false, "test and reference regions are also sorted by strand (-S must be set)");
  cmd.AddOption("-i", &IGNORE_STRAND, false, "ignore strand while finding overlaps");
  if (op == "count") {
    cmd.AddOption("-gaps", &MATCH_GAPS, false, "matching gaps between intervals are considered overlaps");
    cmd.AddOption("--max-label-value", &MAX_LABEL_VALUE, 1L, "maximum region label value to be used");
    cmd.AddOption("-min", &MIN_COUNT, 0UL, "minimum count");
  } else if (op == "coverage") {
    cmd.AddOption("-gaps", &MATCH_GAPS, false, "matching gaps between intervals are considered overlaps");
    cmd.AddOption("--max-label-value", &MAX_LABEL_VALUE, 1L, "maximum region label value to be used");
    cmd.AddOption("-min", &MIN_COUNT, 0UL, "minimum coverage");
  } else if (op == "density") {
    cmd.AddOption("-gaps", &MATCH_GAPS, false, "matching gaps between intervals are considered overlaps");
    cmd.AddOption("--max-label-value", &MAX_LABEL_VALUE, 1L, "maximum region label value to be used");
    cmd.AddOption("-min", &MIN_DENSITY, 0.0, "minimum density");
  } else if (op == "rpkm") {
    cmd.AddOption("-gaps", &MATCH_GAPS, false, "matching gaps between intervals are considered overlaps");
    cmd.AddOption("--max-label-value", &MAX_LABEL_VALUE, 1L, "maximum region label value to be used");
    cmd.AddOption("-min", &MIN_RPKM, 0.0, "minimum RPKM");
  } else if (op == "overlap") {
    cmd.AddOption("-gaps", &MATCH_GAPS, false, "matching gaps between intervals are considered overlaps");
    cmd.AddOption("-label", &MERGE_LABELS, false, "print query label for each match");
  } else if (op == "subset") {
    cmd.AddOption("-gaps", &MATCH_GAPS, false, "matching gaps between intervals are considered overlaps");
    cmd.AddOption("-inv", &SUBSET_NONOVERLAPS, false, "print test regions that do *not* overlap with reference regions");
  } else if (cmd.HasOperation(op)) {
    std::cerr << "Operation '" << op << "' is outside the GPU-accelerated path of this build (count, coverage, density, rpkm, subset, overlap are available)!\n";
    exit(1);
  } else {
    std::cerr << "Unknown operation '" << op << "'!\n";
    exit(1);
  }
  const int next_arg = cmd.Read(argv + 1, argc - 1) + 1;
  if (HELP || argc - next_arg < 1) { cmd.OperationUsage(); exit(1); }
  if (IS_SORTED && SORTED_BY_STRAND && IGNORE_STRAND) {
    fprintf(stderr, "[Error]: the input is sorted by chromosome/strand/start (i.e. -S and -s are set), therefore the overlap algorithm can only report strand-specific results (i.e. -i cannot be set)!\n");
    exit(1);
  }
  // -B: the shift-bits of the bin levels (UnsortedGenomicRegionSetOverlaps, genomic_intervals.cpp:5628-5637).  Only `overlap -label`
  // can tell one bin layout from another: the matches of a query come out in the order the bins are walked.
  std::vector<int> bin_bits;
  for (const char *p = BIN_BITS; p && *p;) {
    bin_bits.push_back(atoi(p));
    const char *c = strchr(p, ',');
    p = c ? c + 1 : nullptr;
  }
  const char *ref_file = argv[next_arg];
  const char *test_file = next_arg + 1 == argc ? nullptr : argv[next_arg + 1];

  // ---- reference (index) set: loaded in memory, labels kept for the output (the CUDA context comes up meanwhile)
  gt::PhaseTimer timer;
  start_context();
  gt::ChromTable chroms;
  gt::RegionBatch ref;
  {
    gt::RegionReader rr(ref_file, &chroms, true, 1);
    rr.ReadAll(&ref);
    if (VERBOSE) std::cerr << "Reading from '" << ref_file << "'; number of regions = " << ref.n_regions() << "; format = " << rr.format() << "\n";
  }
  // -S: the reference's SortedGenomicRegionSetOverlaps reads the index set lazily, while the queries advance
  // (LoadIndexBuffer, genomic_intervals.cpp:5840-5875): an index region is checked for well-formedness when it becomes the
  // current one and for sortedness when it is fetched, so index lines behind the last query's reach are never looked at.
  // `index_pos` is that fetch position; advance_index() repeats the loop for one query (first interval's chromosome/strand/start,
  // last interval's stop) and dies where the reference would -- naming the region by its 0-based place in the file, because
  // CountIndexOverlaps / CalcIndexCoverage have overwritten the index regions' line numbers with it by then (:5309, :5274).
  int64_t index_pos = 0;
  auto chrom_cmp = [&](int32_t a, int32_t b) { return a == b ? 0 : strcmp(chroms.name[a].c_str(), chroms.name[b].c_str()); };
  auto advance_index = [&](int32_t qc, char qstrand, long qs, long qe) {
    while (index_pos < ref.n_regions()) {
      const int64_t k = index_pos, lo = ref.offset[k], hi = ref.offset[k + 1];
      if (!gt::RegionWellFormed(ref, k)) gt::die_line((long)k, "index regions should be compatible, sorted and non-overlapping!");
      int d = chrom_cmp(qc, ref.chrom[lo]);                              // GenomicRegion::CalcDirection, :1225-1237
      if (d == 0 && SORTED_BY_STRAND) d = (int)qstrand - (int)(char)ref.strand[lo];
      if (d == 0) d = (long)ref.stop[hi - 1] < qs ? 1 : qe < (long)ref.start[lo] ? -1 : 0;
      if (d < 0) break;
      index_pos++;
      if (index_pos < ref.n_regions()) {                                 // IsBefore(previous), :396-401
        const int64_t a = ref.offset[index_pos];
        int c = chrom_cmp(ref.chrom[a], ref.chrom[lo]);
        bool before = c < 0;
        if (c == 0) before = SORTED_BY_STRAND && ref.strand[a] != ref.strand[lo] ? (char)ref.strand[a] < (char)ref.strand[lo] : ref.start[a] < ref.start[lo];
        if (before) gt::die_line((long)index_pos, std::string("index regions are not sorted (sorted-by-strand = ") + (SORTED_BY_STRAND ? "true" : "false") + ")!");
      }
    }
  };

  // subset / overlap open the test set with hide_header == false: its header lines are echoed before anything else happens
  // (genomic_overlaps.cpp:711, :787; ProcessFileHeader)
  gt::RegionReader *qr_keep = nullptr;
  if (op == "subset" || op == "overlap") {
    qr_keep = new gt::RegionReader(test_file, &chroms, true, 1);
    fwrite(qr_keep->header().data(), 1, qr_keep->header().size(), stdout);
  }
  timer.Mark("load_reference");
  gtb_ctx *ctx = wait_context();
  int rc = GTB_OK;
  timer.Mark("cuda_context");
  const bool want_coverage = op == "coverage" || op == "density";
  // -S: the Sorted class's admission (no fatal checks on zero-length or non-positive intervals, no index region skipped)
  const unsigned flags = (MATCH_GAPS ? GTB_MATCH_GAPS : 0u) | (IGNORE_STRAND ? GTB_IGNORE_STRAND : 0u) | (IS_SORTED ? GTB_SORTED_RULES : 0u);
  gtb_index *index = nullptr;
  int64_t err_index = -1;
  gtb_set ref_set = as_set(ref);
  ref_set.weight = nullptr;
  gtb_mgpu_index *mindex = nullptr;                                     // GTB_GPUS > 1 (count / coverage / density / rpkm)
  const bool multi_gpu = g_mgpu != nullptr && op != "subset" && op != "overlap";
  if (multi_gpu) rc = gtb_mgpu_index_create(g_mgpu, &ref_set, want_coverage ? GTB_OP_COVERAGE : GTB_OP_COUNT, flags, &mindex, &err_index);
  else rc = gtb_index_create(ctx, &ref_set, want_coverage ? GTB_OP_COVERAGE : GTB_OP_COUNT, flags, &index, &err_index);
  if (rc == GTB_ERR_INDEX_REGION) gt::die_line(ref.line(err_index), "index regions should be compatible, sorted and non-overlapping!");
  if (multi_gpu && rc != GTB_OK) { fprintf(stderr, "\nError: [gtb_mgpu_index_create] %s (status %d)\n", gtb_mgpu_last_error(g_mgpu), rc); exit(1); }
  check(ctx, rc, "gtb_index_create");
  auto mcheck = [&](int status, const char *what) {
    if (status != GTB_OK) { fprintf(stderr, "\nError: [%s] %s (status %d)\n", what, gtb_mgpu_last_error(g_mgpu), status); exit(1); }
  };
  auto finish_values = [&](uint64_t *out, int64_t *ei) {
    return multi_gpu ? gtb_mgpu_index_finish(mindex, out, ei) : gtb_index_finish(index, out, GTB_MEM_HOST, ei);
  };

  timer.Mark("index");
  if (op == "subset" || op == "overlap") {
    // ---- the per-QUERY operations (genomic_overlaps.cpp:782-800, :706-739): a query region is printed if it has an overlap (or has
    // none: -inv), resp. once per overlapping reference region.  The engine gives the number of overlaps per query
    // (gtb_index_query_counts); the regions are printed from their input lines the way GenomicRegion*::Print would.
    gt::RegionReader &qr = *qr_keep;
    if (VERBOSE) std::cerr << "Reading from '" << (test_file ? test_file : "<standard input>") << "'; format = " << qr.format() << "\n";
    gt::SortChecker sc; sc.by_strand = SORTED_BY_STRAND;
    gt::RegionBatch b;
    std::vector<std::string> raw;
    std::vector<uint32_t> n_overlaps;
    std::string text;
    for (;;) {
      raw.clear();
      const int64_t nq = qr.ReadKeep(&b, &raw, 1 << 20);
      if (nq > 0) {
        if (IS_SORTED)
          for (int64_t k = 0; k < nq; k++) {
            const int64_t i = b.offset[k];
            if (!gt::RegionWellFormed(b, k)) gt::die_line(b.line(k), "query regions should be compatible, sorted and non-overlapping!");
            if (!sc.Accept(chroms.name[b.chrom[i]], (char)b.strand[i], b.start[i]))
              gt::die_line(b.line(k), std::string("query regions are not sorted (sorted-by-strand = ") + (SORTED_BY_STRAND ? "true" : "false") + ")!");
            advance_index(b.chrom[i], (char)b.strand[i], b.start[i], b.stop[b.offset[k + 1] - 1]);
          }
        gtb_set qs = as_set(b);
        qs.weight = nullptr;
        n_overlaps.assign((size_t)nq, 0u);
        rc = gtb_index_query_counts(index, &qs, GTB_MEM_HOST, n_overlaps.data(), GTB_MEM_HOST, &err_index);
        const bool fatal = rc == GTB_ERR_QUERY_STOP_NONPOSITIVE || rc == GTB_ERR_QUERY_START_GT_STOP || rc == GTB_ERR_QUERY_REGION;
        if (!fatal) check(ctx, rc, "gtb_index_query_counts");
        const int64_t upto = fatal ? err_index : nq;                     // the reference has printed the queries before the fatal one
        text.clear();
        if (MERGE_LABELS && upto > 0) {
          // overlap -label: one line per matching reference region, labelled "query:reference", in the order the reference's
          // engine walks its matches (gtb_index_query_matches) -- for the queries in front of a fatal one, as above
          std::vector<int64_t> m_off((size_t)upto + 1, 0);
          for (int64_t k = 0; k < upto; k++) m_off[(size_t)k + 1] = m_off[(size_t)k] + n_overlaps[(size_t)k];
          std::vector<int32_t> m_id((size_t)std::max<int64_t>(m_off[(size_t)upto], 1));
          gtb_set head = qs;
          head.n_regions = upto; head.n_intervals = b.offset[upto];
          check(ctx, gtb_index_query_matches(index, &head, GTB_MEM_HOST, bin_bits.data(), (int)bin_bits.size(), m_off.data(), m_id.data(), &err_index), "gtb_index_query_matches");
          for (int64_t k = 0; k < upto; k++) {
            if (m_off[(size_t)k + 1] == m_off[(size_t)k]) continue;
            std::string own;
            own.swap(b.label[(size_t)k]);
            for (int64_t t = m_off[(size_t)k]; t < m_off[(size_t)k + 1]; t++) {
              b.label[(size_t)k] = own + ":" + ref.label[(size_t)m_id[(size_t)t]];         // genomic_overlaps.cpp:721-727
              gt::PrintRegion(qr.format(), raw[(size_t)k], b, k, chroms, &text);
            }
            b.label[(size_t)k].swap(own);
            if (text.size() > (1u << 24)) { fwrite(text.data(), 1, text.size(), stdout); text.clear(); }
          }
        }
        for (int64_t k = 0; k < (MERGE_LABELS ? 0 : upto); k++) {
          const uint32_t times = op == "overlap" ? n_overlaps[(size_t)k] : ((n_overlaps[(size_t)k] == 0) == SUBSET_NONOVERLAPS ? 1u : 0u);
          for (uint32_t t = 0; t < times; t++) gt::PrintRegion(qr.format(), raw[(size_t)k], b, k, chroms, &text);
          if (text.size() > (1u << 24)) { fwrite(text.data(), 1, text.size(), stdout); text.clear(); }
        }
        fwrite(text.data(), 1, text.size(), stdout);
        if (fatal) { fflush(stdout); die_query(rc, b.line(err_index)); }
      }
      if (qr.failed()) { fflush(stdout); qr.Fail(); }
      if (nq == 0) break;
    }
    fflush(stdout);
    timer.Mark("stream_queries");
    return 0;
  }
  // ---- test (query) set: streamed in chunks; parsing of chunk k+1 overlaps the device work of chunk k
  const int64_t CHUNK = 4 << 20;
  gt::RegionBatch chunk[2];
  long first_query_line = 0;                                           // every data line is one query: query k is on first_query_line + k
  {
    gt::RegionReader qr(test_file, &chroms, false, MAX_LABEL_VALUE);
    if (VERBOSE) std::cerr << "Reading from '" << (test_file ? test_file : "<standard input>") << "'; format = " << qr.format() << "\n";
    gt::SortChecker sc; sc.by_strand = SORTED_BY_STRAND;
    int which = 0;
    int64_t seen = 0;
    for (;;) {
      gt::RegionBatch &b = chunk[which];
      if (!multi_gpu) check(ctx, gtb_ctx_synchronize(ctx), "gtb_ctx_synchronize");     // the buffer about to be overwritten has been consumed
      if (qr.Read(&b, CHUNK) == 0) break;
      if (IS_SORTED)
        for (int64_t k = 0; k < b.n_regions(); k++) {
          const int64_t i = b.offset[k];
          if (!gt::RegionWellFormed(b, k)) gt::die_line(b.line(k), "query regions should be compatible, sorted and non-overlapping!");
          if (!sc.Accept(chroms.name[b.chrom[i]], (char)b.strand[i], b.start[i]))
            gt::die_line(b.line(k), std::string("query regions are not sorted (sorted-by-strand = ") + (SORTED_BY_STRAND ? "true" : "false") + ")!");
          advance_index(b.chrom[i], (char)b.strand[i], b.start[i], b.stop[b.offset[k + 1] - 1]);
        }
      gtb_set qs = as_query_set(b);
      if (multi_gpu) mcheck(gtb_mgpu_index_add_queries(mindex, &qs), "gtb_mgpu_index_add_queries");   // (the copies have left the buffer on return)
      else check(ctx, gtb_index_add_queries(index, &qs, GTB_MEM_HOST), "gtb_index_add_queries");
      if (seen == 0) first_query_line = b.first_line;
      seen += b.n_regions();
      which ^= 1;
    }
    if (!multi_gpu) check(ctx, gtb_ctx_synchronize(ctx), "gtb_ctx_synchronize");
    if (qr.failed()) {
      // a malformed line ended the stream; a query before it that the engine refuses comes first in the file and is the one
      // the reference would have stopped at
      std::vector<uint64_t> scratch((size_t)std::max<int64_t>(ref.n_regions(), 1));
      rc = finish_values(scratch.data(), &err_index);
      if (rc != GTB_ERR_QUERY_STOP_NONPOSITIVE && rc != GTB_ERR_QUERY_START_GT_STOP && rc != GTB_ERR_QUERY_REGION) qr.Fail();
      die_query(rc, first_query_line + err_index);
    }
  }
  timer.Mark("stream_queries");
  std::vector<uint64_t> values((size_t)std::max<int64_t>(ref.n_regions(), 1));
  rc = finish_values(values.data(), &err_index);
  if (rc == GTB_ERR_QUERY_STOP_NONPOSITIVE || rc == GTB_ERR_QUERY_START_GT_STOP || rc == GTB_ERR_QUERY_REGION)
    die_query(rc, first_query_line + err_index);
  if (multi_gpu) mcheck(rc, "gtb_mgpu_index_finish");
  check(ctx, rc, "gtb_index_finish");

  timer.Mark("finish");
  // ---- output, reference-file order (genomic_overlaps.cpp:420-427, :449-455, :476-486, :763-772)
  auto region_size = [&](int64_t k, bool skip_gaps) -> long {        // GenomicRegion::GetSize, genomic_intervals.cpp:1047-1055
    const int64_t lo = ref.offset[k], hi = ref.offset[k + 1];
    if (!skip_gaps) return (long)ref.stop[hi - 1] - (long)ref.start[lo] + 1;
    size_t size = 0;
    for (int64_t i = lo; i < hi; i++) size += ref.start[i] > ref.stop[i] ? 0 : (size_t)((long)ref.stop[i] - (long)ref.start[i] + 1);
    return (long)size;
  };
  if (op == "count" || op == "coverage") {
    for (int64_t k = 0; k < ref.n_regions(); k++)
      if (values[k] >= MIN_COUNT) printf("%s\t%lu\n", ref.label[k].c_str(), (unsigned long)values[k]);
  } else if (op == "density") {
    for (int64_t k = 0; k < ref.n_regions(); k++) {
      const long size = region_size(k, !MATCH_GAPS);
      const double density = (double)values[k] / size;
      if (density >= MIN_DENSITY) printf("%s\t%.4e\n", ref.label[k].c_str(), density);
    }
  } else {  // rpkm
    unsigned long nreads = 0;
    for (int64_t k = 0; k < ref.n_regions(); k++) nreads += values[k];
    if (VERBOSE) fprintf(stderr, "* %lu reads overlap reference regions.\n", nreads);
    const double mreads = (double)nreads / 1000000;
    for (int64_t k = 0; k < ref.n_regions(); k++) {
      // the reference filters on MIN_COUNT here, which its rpkm branch never sets (0): every region is printed
      const long eff_len = region_size(k, !MATCH_GAPS);
      const double rpkm = eff_len <= 0 ? 0.0 / 0.0 : (double)1000 * values[k] / eff_len / mreads;
      printf("%s\t%.4e\n", ref.label[k].c_str(), rpkm);
    }
  }
  fflush(stdout);
  timer.Mark("print");
  // the process is about to end: the driver reclaims device memory faster than freeing it buffer by buffer would
  (void)index; (void)mindex;
  return 0;
}

// the driver's work is done and its output written when driver_main returns: the process leaves through gt::Exit (gt_host.h)
int main(int argc, char *argv[]) {
  exit(driver_main(argc, argv));
}
