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
//  (3) CELL structures for the fast path (gtb_cell.cu): the groups' coordinate axes are cut into
//      uniform cells of 2^k bp, laid end to end.  A bitmap of the "hot" cells (those containing an
//      evaluation point) lives in shared memory; a query whose start and stop fall into one cold
//      cell costs a single fire-and-forget red.global on an L2-resident per-cell table.  Only
//      queries touching a hot cell compare against that cell's (<= 13) points, fetched as one
//      32-byte record, and bump per-point correction counters.
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
  bool match_gaps = false, ignore_strand = false;
  unsigned engine = GTB_ENGINE_AUTO;
  int64_t n_regions = 0, n_intervals = 0;
  bool index_multi = false;        // some index region has more than one interval
  int planes = H_PLANES_COUNT;

  // host copy of the index set (small) for lazy construction of secondary structures
  std::vector<int32_t> h_chrom, h_start, h_stop;
  std::vector<int8_t> h_strand;
  std::vector<int64_t> h_off;

  // rank
  int32_t n_chrom = 0, n_class = 0, n_groups = 0;
  int64_t n_slots = 0, n_targets = 0;
  std::vector<int32_t> h_points, h_goff;       // kept for the cell engine's builder
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
  dbuf<int32_t> d_rid, d_r_chrom, d_r_start, d_r_stop;
  dbuf<int8_t> d_r_strand;
  dbuf<int64_t> d_r_off;
  dbuf<ull> d_direct;

  // cell engine state (gtb_cell.cu)
  struct gtb_cell_state *cell = nullptr;
  // bucket engine state (gtb_bucket.cu)
  struct gtb_bucket_state *bucket = nullptr;
  // direct engine state (gtb_direct.cu)
  struct gtb_direct_state *direct = nullptr;

  // results / errors
  dbuf<ull> d_err, d_out;
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

// ---- cell engine (gtb_cell.cu) -----------------------------------------------------------------
// planes of the per-cell tables (n_cells each) and per-slot correction tables (n_slots each)
enum { C_BOTH = 0, C_SCNT = 1, C_ECNT = 2, C_SSUM = 3, C_ESUM = 4 };      // cell planes (count uses 0..2)
enum { X_SCNT = 0, X_ECNT = 1, X_SSUM = 2, X_ESUM = 3 };                  // correction planes (count uses 0..1)

struct __align__(16) HotRec {      // one 32-byte sector per hot cell
  uint32_t slot_base;              // slot of the first evaluation point inside the cell
  uint16_t n;                      // points in the cell; > HOT_MAX means "too many, use the general path"
  uint16_t off[13];                // point & (cell width - 1), ascending
};
constexpr int HOT_MAX = 13;

struct gtb_cell_state {
  bool ready = false;
  int k = 0;                       // cell width = 2^k
  uint32_t n_cells = 0, n_words = 0, n_hot = 0;
  int cell_planes = 3, corr_planes = 2;
  size_t smem_bytes = 0;
  dbuf<int32_t> d_gsize;           // [n_groups] largest evaluation point of the group (0 = empty group)
  dbuf<uint32_t> d_gbase;          // [n_groups] first cell of the group
  dbuf<int2> d_gtab;               // [n_groups] (gsize, gbase) interleaved
  dbuf<uint32_t> d_bitmap, d_wrank;   // [n_words]
  dbuf<HotRec> d_hot;              // [n_hot]
  dbuf<uint32_t> d_slot_cell;      // [n_slots] cell of each slot's point (0xFFFFFFFF: point <= 0 or sentinel)
  dbuf<ull> d_cells, d_cells_scan; // cell_planes * n_cells
  dbuf<ull> d_corr;                // corr_planes * n_slots
  bool dirty = false;              // something was accumulated since the last reset
};

struct CellFinalView {             // what finalize needs from the cell engine (all null/0 when unused)
  const ull *cells_scan;           // inclusive prefix sums of the cell planes
  const ull *corr;
  const uint32_t *slot_cell;
  const uint32_t *gbase;           // per group
  const int32_t *t_group;          // per target
  int64_t n_cells;
};

int gtb_cell_prepare(gtb_index *ix);
int gtb_cell_accumulate(gtb_index *ix, const QueryView &q);
int gtb_cell_reset(gtb_index *ix);
int gtb_cell_scan_for_finish(gtb_index *ix, CellFinalView *out);
void gtb_cell_destroy(gtb_index *ix);
bool gtb_cell_supported(gtb_index *ix, const QueryView &q, bool batch_multi);

// ---- bucket engine (gtb_bucket.cu) ---------------------------------------------------------------
int gtb_bucket_prepare(gtb_index *ix);
int gtb_bucket_accumulate(gtb_index *ix, const QueryView &q);
void gtb_bucket_destroy(gtb_index *ix);
bool gtb_bucket_supported(gtb_index *ix, const QueryView &q, bool batch_multi);

// ---- direct engine (gtb_direct.cu) ---------------------------------------------------------------
int gtb_direct_prepare(gtb_index *ix);
int gtb_direct_accumulate(gtb_index *ix, const QueryView &q);
void gtb_direct_destroy(gtb_index *ix);
bool gtb_direct_supported(gtb_index *ix, const QueryView &q, bool batch_multi);
