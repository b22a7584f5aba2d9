// gtb_overlap.cuh -- data structures of the overlap count / coverage engines.
//
// The device index replaces the reference's UCSC-style bin index + linked lists
// (genomic_intervals.cpp:5593-5675) and its sorted sweep (5807-5937) with three things:
//
//  (1) RANK structures.  For index "targets" t = [ts,te] (region spans for count / -gaps, region
//      blocks for coverage) grouped by g = (chromosome, strand class), the sorted distinct
//      evaluation points P_g = {te} U {ts-1} plus one +inf sentinel slot.  A query [qs,qe] of weight
//      w falls into slots jS = lower_bound(P_g,qs), jE = lower_bound(P_g,qe) and is accumulated
//      into per-slot histograms; prefix sums of those give, exactly and in integers,
//          count[t]    = #{qs <= te} - #{qe <= ts-1}
//          coverage[t] = F(te) - F(ts-1),  F(x) = sum_q w*|[qs,qe] /\ (-inf,x]|
//      (SURVEY.md section 7.1).  Queries with jS == jE -- almost all short reads -- cost a single
//      atomic ("both" histogram); stragglers that straddle an evaluation point use the S/E pair.
//  (2) ENUMERATE structures: indexable regions keyed by (chromosome, level, bin) in CSR order,
//      for the cases ranks cannot express (count without -gaps when either side has
//      multi-interval regions: "any pair of blocks overlaps", genomic_intervals.cpp:1167-1172).
#pragma once
#include "gtb_internal.cuh"

typedef unsigned long long ull;

// histogram planes (each n_slots long) inside gtb_index::d_hist
enum { H_BOTH = 0, H_SCNT = 1, H_ECNT = 2, H_SSUM = 3, H_ESUM = 4, H_PLANES_COUNT = 3, H_PLANES_COVERAGE = 5 };

struct QueryView {                 // device pointers
  int64_t n_regions;
  const int32_t *chrom, *start, *stop;
  const int8_t *strand;
  const int32_t *weight;           // per region or null
  const int64_t *region_offset;    // [n_regions+1] or null
  int64_t interval_base;           // value to subtract from region_offset entries
  int64_t index_base;              // stream-order index of region 0 of this batch (for errors)
};

struct RankView {                  // device pointers
  int32_t n_chrom, n_class;
  const int8_t *class_of;          // [256] strand byte -> class or -1
  const uint8_t *chrom_present;    // [n_chrom]
  const int32_t *goff;             // [n_groups+1] slot offsets
  const int32_t *points;           // [n_slots]
  int64_t n_slots;
  ull *hist;                       // planes of n_slots
  ull *err;                        // min over (index<<8 | code)
  int32_t admission;               // 0: the Unsorted class's fatal checks; 1: GTB_SORTED_RULES (zero-length and stop <= 0 queries pass);
                                   // 2: the batch holds BLOCKS of regions that were checked as a whole (no errors; a block with start > stop counts nothing)
};

struct EnumView {
  int64_t n_entries;
  const ull *keys;                 // sorted (chrom<<35 | level<<32 | bin)
  const int32_t *rid;              // region id per entry
  const int32_t *r_chrom, *r_start, *r_stop;   // index intervals
  const int8_t *r_strand;
  const int64_t *r_off;            // [n_regions+1]
  ull *direct;                     // [n_regions]
};

struct gtb_index {
  gtb_ctx *ctx = nullptr;
  int op = GTB_OP_COUNT;
  bool match_gaps = false, ignore_strand = false, sorted_rules = false;
  bool flat_blocks = false;        // transient: the batch being accumulated is the flattened block list of multi-interval queries
  bool leftovers = false;          // transient: the batch being accumulated is what the one-pass engine has left for the enumeration engine
  int admission() const { return flat_blocks ? 2 : sorted_rules ? 1 : 0; }
  unsigned engine = GTB_ENGINE_AUTO;
  int64_t n_regions = 0, n_intervals = 0;
  bool index_multi = false;        // some index region has more than one interval
  int planes = H_PLANES_COUNT;

  // host copy of the index set (small) for lazy construction of secondary structures
  std::vector<int32_t> h_chrom, h_start, h_stop;
  std::vector<int8_t> h_strand;
  std::vector<int64_t> h_off;
  std::vector<uint8_t> h_malformed; // GTB_SORTED_RULES only: index regions that are not well-formed (never indexed; empty = none)

  // rank
  int32_t n_chrom = 0, n_class = 0, n_groups = 0;
  int64_t n_slots = 0, n_targets = 0;
  std::vector<int32_t> h_points, h_goff;       // kept for the builders of the bucket and direct engines
  std::vector<int8_t> h_class_of;
  std::vector<uint8_t> h_present;
  dbuf<int8_t> d_class_of;
  dbuf<uint8_t> d_present;
  dbuf<int32_t> d_goff, d_points, d_t_hi, d_t_lo, d_t_base;
  dbuf<int64_t> d_t_off;                       // [n_regions+1] targets per region
  dbuf<ull> d_hist, d_hist_scan, d_scan_scratch;

  // enumerate (lazy)
  bool enum_ready = false;
  int64_t n_entries = 0;
  dbuf<ull> d_keys;
  bool match_ready = false;       // the reference's own bin index as an entry list (gtb_index_query_matches)
  std::vector<int> match_bits;
  int64_t n_match_entries = 0;
  dbuf<ull> d_mkeys;
  dbuf<int32_t> d_mrid;
  dbuf<int32_t> d_rid, d_r_chrom, d_r_start, d_r_stop;
  dbuf<int8_t> d_r_strand;
  dbuf<int64_t> d_r_off;
  dbuf<ull> d_direct;
  // per-query counts (lazy): exclusive prefix pairs per slot, see query_counts_kernel
  bool qpre_ready = false;
  dbuf<uint2> d_qpre;

  // bucket engine state (gtb_bucket.cu)
  struct gtb_bucket_state *bucket = nullptr;
  // direct engine state (gtb_direct.cu)
  struct gtb_direct_state *direct = nullptr;

  // results / errors
  dbuf<ull> d_err, d_out;
  // spans of multi-interval query regions (-gaps): one single-interval query per region, written by the prepass
  dbuf<int32_t> sp_chrom, sp_start, sp_stop;
  dbuf<int8_t> sp_strand;
  dbuf<int64_t> uni_off;            // offsets written out for a batch of uniform regions (see accumulate_device)
  int64_t queries_seen = 0;

  // double-buffered staging for host-resident query batches
  struct stage {
    dbuf<int32_t> chrom, start, stop, weight;
    dbuf<int8_t> strand;
    dbuf<int64_t> off;
    dbuf<uint32_t> meta;                       // packed form of a chunk (gtb_ingest.cpp), expanded by unpack_kernel
    cudaEvent_t copied = nullptr, consumed = nullptr;
    bool in_flight = false;
  } stages[2];
  int next_stage = 0;
};

// ---- bucket engine (gtb_bucket.cu) ---------------------------------------------------------------
int gtb_bucket_prepare(gtb_index *ix);
int gtb_bucket_accumulate(gtb_index *ix, const QueryView &q);
void gtb_bucket_destroy(gtb_index *ix);
bool gtb_bucket_supported(gtb_index *ix, const QueryView &q, bool batch_multi);

// ---- direct engine (gtb_direct.cu) ---------------------------------------------------------------
int gtb_direct_prepare(gtb_index *ix);
// pair_check: the batch is the flattened intervals of two-interval regions.  1: the intervals are the queries (coverage), every pair
// is checked as the reference checks a region.  2: the pairs' spans are the queries (-gaps)
//   3: count without -gaps -- spans that hold no evaluation point are counted, the other pairs are left on a list
//   (gtb_direct_exceptions) for the enumeration engine.  GTB_ERR_UNSUPPORTED from modes 2 and 3: nothing of the batch has been
//   counted, the caller takes its general path
//   4: count without -gaps, the queries are the spans of regions of any shape (region_offsets: the batch's offsets, rebased to 0):
//   multi-interval regions whose span holds an evaluation point are left on a list of region numbers (gtb_direct_exception_list)
int gtb_direct_accumulate(gtb_index *ix, const QueryView &q, int pair_check = 0, const int64_t *region_offsets = nullptr);
int64_t gtb_direct_exceptions(gtb_index *ix, QueryView *out);
int64_t gtb_direct_exception_list(gtb_index *ix, const int32_t **list);
void gtb_direct_destroy(gtb_index *ix);
void gtb_direct_reset(gtb_index *ix);     // a new query stream: the watchdog's verdict on the previous one no longer holds
bool gtb_direct_supported(gtb_index *ix, const QueryView &q, bool batch_multi);
bool gtb_direct_usable(gtb_index *ix);
