// oracle/shim/genomic_intervals.h -- TEST INFRASTRUCTURE: proves the drop-in of INTEGRATION.md section 2.
//
// The reference's own driver, gtools/genomic_overlaps.cpp, is compiled UNMODIFIED from where it lies (oracle/Makefile feeds it
// to the compiler on standard input, so that its `#include "genomic_intervals.h"` finds THIS file first); this file pulls in
// the reference's header of the same name, adds the one class a maintainer would add -- GPUGenomicRegionSetOverlaps, which
// forwards CountIndexOverlaps / CalcIndexCoverage to the C ABI of include/gtb200.h -- and points the driver's engine
// types at it.  tests/test_dropin_shim.py runs the resulting binary (oracle/_ref/genomic_overlaps_gpu) against the stock one.
//
// Why macros instead of a subclass passed through the base pointer: CountIndexOverlaps and CalcIndexCoverage are not virtual in
// the reference (genomic_intervals.h:2453, :2471), so the driver's `GenomicRegionSetOverlaps *overlaps` would call the base
// versions.  A maintainer would write `virtual` in front of the two declarations; not editing the reference, the shim renames the
// static type instead.  Every other operation of the driver (annotate, overlap, subset, ...) walks GetQuery / GetMatch, which this
// class hands to the reference's own engine, built on first use.
#include_next "genomic_intervals.h"
#include <stdint.h>
#include "gtb200.h"

class GPUGenomicRegionSetOverlaps : public GenomicRegionSetOverlaps {
 public:
  // the signatures the driver uses for the Sorted and the Unsorted engine (genomic_overlaps.cpp:416-417)
  GPUGenomicRegionSetOverlaps(GenomicRegionSet *QuerySet, GenomicRegionSet *IndexSet, bool sorted_by_strand)
    : GenomicRegionSetOverlaps(QuerySet, IndexSet), sorted(true), by_strand(sorted_by_strand), bin_bits(NULL), inner(NULL), ctx(NULL) {}
  GPUGenomicRegionSetOverlaps(GenomicRegionSet *QuerySet, GenomicRegionSet *IndexSet, const char *bits = NULL)
    : GenomicRegionSetOverlaps(QuerySet, IndexSet), sorted(false), by_strand(false), bin_bits(bits), inner(NULL), ctx(NULL) {}
  ~GPUGenomicRegionSetOverlaps() { if (inner) delete inner; if (ctx) gtb_ctx_destroy(ctx); }

  // replaces genomic_intervals.cpp:5304-5317 and :5269-5285
  unsigned long int *CountIndexOverlaps(bool match_gaps, bool ignore_strand, long int max_label_value) { return Run(GTB_OP_COUNT, match_gaps, ignore_strand, max_label_value); }
  unsigned long int *CalcIndexCoverage(bool match_gaps, bool ignore_strand, long int max_label_value) { return Run(GTB_OP_COVERAGE, match_gaps, ignore_strand, max_label_value); }

  // everything else stays the reference's
  GenomicRegion *GetQuery() { return Inner()->GetQuery(); }
  GenomicRegion *NextQuery() { return Inner()->NextQuery(); }
  GenomicRegion *GetMatch() { return Inner()->GetMatch(); }
  GenomicRegion *NextMatch() { return Inner()->NextMatch(); }
  bool Done() { return Inner()->Done(); }
  GenomicRegion *GetOverlap(bool match_gaps, bool ignore_strand) { return Inner()->GetOverlap(match_gaps, ignore_strand); }
  GenomicRegion *NextOverlap(bool match_gaps, bool ignore_strand) { return Inner()->NextOverlap(match_gaps, ignore_strand); }
  unsigned long int CalcQueryCoverage(bool g, bool i, long int m) { return Inner()->CalcQueryCoverage(g, i, m); }
  unsigned long int CountQueryOverlaps(bool g, bool i, long int m) { return Inner()->CountQueryOverlaps(g, i, m); }

 private:
  struct SoA { std::vector<int32_t> chrom, start, stop, weight; std::vector<int8_t> strand; std::vector<int64_t> off; SoA() : off(1, 0) {} };
  bool sorted, by_strand;
  const char *bin_bits;
  GenomicRegionSetOverlaps *inner;
  gtb_ctx *ctx;
  std::map<std::string, int32_t> chrom_id;                            // any numbering works for count / coverage

  GenomicRegionSetOverlaps *Inner() {
    if (!inner) inner = sorted ? (GenomicRegionSetOverlaps *)new SortedGenomicRegionSetOverlaps(QuerySet, IndexSet, by_strand)
                               : (GenomicRegionSetOverlaps *)new UnsortedGenomicRegionSetOverlaps(QuerySet, IndexSet, bin_bits);
    return inner;
  }
  void Push(SoA &s, GenomicRegion *r, long int max_label_value) {
    for (GenomicIntervalSet::iterator it = r->I.begin(); it != r->I.end(); it++) {
      GenomicInterval *i = *it;
      std::map<std::string, int32_t>::iterator f = chrom_id.find(i->CHROMOSOME);
      if (f == chrom_id.end()) f = chrom_id.insert(std::make_pair(std::string(i->CHROMOSOME), (int32_t)chrom_id.size())).first;
      s.chrom.push_back(f->second);
      s.start.push_back((int32_t)i->START); s.stop.push_back((int32_t)i->STOP); s.strand.push_back((int8_t)i->STRAND);
    }
    s.off.push_back((int64_t)s.chrom.size());
    s.weight.push_back((int32_t)r->GetLabelValue(max_label_value));
  }
  static gtb_set View(const SoA &s) {
    gtb_set v;
    v.n_regions = (int64_t)s.off.size() - 1; v.n_intervals = (int64_t)s.chrom.size();
    v.chrom = s.chrom.data(); v.start = s.start.data(); v.stop = s.stop.data(); v.strand = s.strand.data();
    v.weight = s.weight.data(); v.region_offset = s.off.data();
    return v;
  }
  static void Die(long line, const char *msg) { fprintf(stderr, "\nError: Line %ld: %s\n", line, msg); exit(1); }
  void Check(int rc, const char *what) {
    if (rc != GTB_OK) { fprintf(stderr, "\nError: [%s] %s (status %d)\n", what, ctx ? gtb_ctx_last_error(ctx) : "", rc); exit(1); }
  }
  void Flush(gtb_index *index, SoA &batch) {
    gtb_set b = View(batch);
    Check(gtb_index_add_queries(index, &b, GTB_MEM_HOST), "gtb_index_add_queries");          // copied inside the call
    batch = SoA();
  }
  unsigned long int *Run(int op, bool match_gaps, bool ignore_strand, long int max_label_value) {
    if (IndexSet->load_in_memory == false) { fprintf(stderr, "[GPUGenomicRegionSetOverlaps]: index set must be loaded in memory for this operation!\n"); exit(1); }
    if (!ctx && gtb_ctx_create(0, &ctx) != GTB_OK) { fprintf(stderr, "\nError: no CUDA device (libgtb200 has no CPU fallback)\n"); exit(1); }
    SoA idx;
    for (long int k = 0; k < IndexSet->n_regions; k++) Push(idx, IndexSet->R[k], 1);
    gtb_set iv = View(idx);
    iv.weight = NULL;
    const unsigned flags = (match_gaps ? GTB_MATCH_GAPS : 0u) | (ignore_strand ? GTB_IGNORE_STRAND : 0u) | (sorted ? GTB_SORTED_RULES : 0u);
    gtb_index *index = NULL;
    int64_t bad = -1;
    int rc = gtb_index_create(ctx, &iv, op, flags, &index, &bad);
    if (rc == GTB_ERR_INDEX_REGION) IndexSet->R[bad]->PrintError("index regions should be compatible, sorted and non-overlapping!");
    Check(rc, "gtb_index_create");
    std::vector<long> line_of;                                         // query stream index -> input line, for the error messages
    SoA batch;
    for (GenomicRegion *q = QuerySet->Get(); q != NULL; q = QuerySet->Next()) {              // the reference's streaming reader
      Push(batch, q, max_label_value);
      line_of.push_back((long)q->n_line);
      if (batch.off.size() > ((size_t)4 << 20)) Flush(index, batch);
    }
    Flush(index, batch);
    unsigned long int *out = new unsigned long int[IndexSet->n_regions > 0 ? IndexSet->n_regions : 1];   // the caller frees it, as before
    rc = gtb_index_finish(index, (uint64_t *)out, GTB_MEM_HOST, &bad);
    if (rc == GTB_ERR_QUERY_STOP_NONPOSITIVE) Die(line_of[bad], "stop position must be positive!");
    if (rc == GTB_ERR_QUERY_START_GT_STOP) Die(line_of[bad], "start position cannot be greater than stop position!");
    if (rc == GTB_ERR_QUERY_REGION) Die(line_of[bad], "query regions should be compatible, sorted and non-overlapping!");
    Check(rc, "gtb_index_finish");
    gtb_index_destroy(index);
    return out;
  }
};

// from here on the driver's engine types are the class above
#define GenomicRegionSetOverlaps GPUGenomicRegionSetOverlaps
#define SortedGenomicRegionSetOverlaps GPUGenomicRegionSetOverlaps
#define UnsortedGenomicRegionSetOverlaps GPUGenomicRegionSetOverlaps
