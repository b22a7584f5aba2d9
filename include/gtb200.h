/* gtb200.h -- C ABI of the B200-native interval overlap / coverage / window-count engine.
 *
 * This is the drop-in boundary for the hot path of GenomicTools 2.8.1a (tsirigos/ibm-cbc-genomic-tools).
 * The reference has no FFI; the seam this ABI replaces is the C++ class API its drivers call
 * (SURVEY.md section 8b).  Each entry point below cites the reference interface it stands in for
 * (paths relative to the reference's gtools/ directory).  INTEGRATION.md shows the binding a
 * maintainer of the reference would add.
 *
 * Conventions
 *  - Plain C: pointers and sizes only.  Every function returns an int status (GTB_OK == 0) and
 *    never calls exit(); the host driver maps a status to the reference's message and exit code.
 *  - Coordinates are the reference's internal ones: 1-based, closed (genomic_intervals.h:333-337),
 *    i.e. BED start+1 (genomic_intervals.cpp:2163), REG/GFF as written.
 *  - Chromosome ids are small non-negative integers chosen by the caller.  For gtb_scan_* they must
 *    be assigned in strcmp order of the names, because ascending id is the output order
 *    (the reference iterates a std::map<string,...>, genomic_intervals.cpp:5099, :5125-5141).
 *  - Strand is the byte the reference keeps in GenomicInterval::STRAND: '+', '-', or for GFF the
 *    raw first character of column 7 (genomic_intervals.cpp:3512).
 *  - Weights are GenomicRegion::GetLabelValue(max_label_value) (genomic_intervals.cpp:1081-1085),
 *    computed by the caller while parsing: 1 when max_label_value <= 1, else min(max, atol(label)).
 *  - The library owns all device memory.  One host thread per context.  There is no CPU fallback:
 *    without a CUDA device gtb_ctx_create fails with GTB_ERR_NO_DEVICE.
 */
#ifndef GTB200_H
#define GTB200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define GTB200_ABI_VERSION 1

typedef struct gtb_ctx gtb_ctx;         /* device, streams, scratch                                */
typedef struct gtb_index gtb_index;     /* an index (reference) region set resident on the device  */
typedef struct gtb_scan gtb_scan;       /* a genome-wide micro-window histogram on the device      */

/* A region set in struct-of-arrays form.  Replaces the reference's
 * GenomicRegion{LABEL, vector<GenomicInterval*>} objects (genomic_intervals.h:962-964, :333-337).
 * Region k owns intervals [region_offset[k], region_offset[k+1]); region_offset == NULL means one
 * interval per region (n_intervals == n_regions), which is what BED3-BED11, GFF and single-interval
 * REG lines produce -- or, for QUERY sets with n_intervals == k * n_regions, k intervals per region
 * (region r = intervals [r * k, (r + 1) * k): read pairs are k == 2), which saves the offsets' 8 bytes per
 * region and, for coverage over read pairs, the pass that checks the regions. */
typedef struct {
  int64_t n_regions;
  int64_t n_intervals;
  const int32_t *chrom;          /* [n_intervals] */
  const int32_t *start;          /* [n_intervals] */
  const int32_t *stop;           /* [n_intervals] */
  const int8_t  *strand;         /* [n_intervals] */
  const int32_t *weight;         /* [n_regions] or NULL (= 1) */
  const int64_t *region_offset;  /* [n_regions+1] or NULL */
} gtb_set;

/* option bits (the reference's -gaps and -i, genomic_overlaps.cpp:206-221, :191) */
enum {
  GTB_MATCH_GAPS       = 1u << 0,
  GTB_IGNORE_STRAND    = 1u << 1,
  /* the admission rules of SortedGenomicRegionSetOverlaps (what -S selects, genomic_intervals.cpp:5807-5937) instead of the
   * Unsorted class's: no fatal check on queries with stop <= 0 or start == stop + 1 (a zero-length BED line), no index region
   * skipped for those reasons; both are matched with the raw predicate of CalcDirection (:1225-1237).  Sortedness itself is
   * the caller's to check (the drivers do).  Intervals with start > stop + 1 remain GTB_ERR_QUERY_START_GT_STOP / skipped. */
  GTB_SORTED_RULES     = 1u << 2,
  /* where the arrays of a query/read gtb_set live */
  GTB_MEM_HOST         = 0u,        /* host memory (pinned or pageable); copied inside the call */
  GTB_MEM_DEVICE       = 1u << 8,   /* device pointers on the context's device                   */
  /* engine selection (testing / benchmarking; results are identical) */
  GTB_ENGINE_AUTO      = 0u,
  GTB_ENGINE_ENUMERATE = 1u << 16,  /* candidate enumeration (general; any region shape)         */
  GTB_ENGINE_RANK      = 1u << 17,  /* rank/rank-sum with global binary search                    */
  GTB_ENGINE_BUCKET    = 1u << 19,  /* two passes: partition by genome bucket, rank in shared memory (any index size)        */
  GTB_ENGINE_DIRECT    = 1u << 20   /* one pass: slot lookup in an L2-resident cell table, byte counters in shared memory
                                       (count, up to ~200 k evaluation points; the default there)                           */
};

enum { GTB_OP_COUNT = 0, GTB_OP_COVERAGE = 1 };

enum {
  GTB_OK = 0,
  GTB_ERR_ARG = 1,
  /* fatal input conditions of the reference; *err_index = 0-based index of the offending region */
  GTB_ERR_QUERY_STOP_NONPOSITIVE = 2,   /* genomic_intervals.cpp:5740 "stop position must be positive!" */
  GTB_ERR_QUERY_START_GT_STOP    = 3,   /* genomic_intervals.cpp:5741 "start position cannot be greater than stop position!" */
  GTB_ERR_QUERY_REGION           = 4,   /* genomic_intervals.cpp:5698,:5709 "query regions should be compatible, sorted and non-overlapping!" */
  GTB_ERR_INDEX_REGION           = 5,   /* genomic_intervals.cpp:5607 "index regions should be compatible, sorted and non-overlapping!" */
  GTB_ERR_WINDOW                 = 6,   /* genomic_intervals.cpp:4845 window size not a multiple of window step */
  GTB_ERR_UNSUPPORTED            = 7,
  GTB_ERR_NO_DEVICE              = 100,
  GTB_ERR_CUDA                   = 101,
  GTB_ERR_NOMEM                  = 102
};

/* ---- context ------------------------------------------------------------------------------ */
int  gtb_abi_version(void);
int  gtb_ctx_create(int device, gtb_ctx **out);
void gtb_ctx_destroy(gtb_ctx *ctx);
/* Launch all work of this context on an existing CUDA stream (cudaStream_t passed as void*), e.g.
 * so that a caller can bracket it with its own events.  NULL restores the context's own stream. */
int  gtb_ctx_set_stream(gtb_ctx *ctx, void *cuda_stream);
/* The CUDA stream (cudaStream_t) the context's kernels run on: its own non-blocking stream unless gtb_ctx_set_stream gave it
 * another.  For callers that order their own device work against the library's with events instead of host waits. */
void *gtb_ctx_get_stream(const gtb_ctx *ctx);
int  gtb_ctx_synchronize(gtb_ctx *ctx);
const char *gtb_ctx_last_error(const gtb_ctx *ctx);
/* Kernel accounting: number of kernels launched by this context since creation / last reset, and
 * (when profiling is on) per-kernel CUDA-event times written as a JSON object into buf. */
int64_t gtb_ctx_launch_count(const gtb_ctx *ctx);
/* Bytes this context actually moved over the host link since creation, and how the host-resident query chunks
 * travelled: re-encoded to 8 B/interval by the host packing pool ("packed") or in the plain 13 B/interval layout
 * ("raw").  Any pointer may be NULL.  GTB_INGEST_THREADS=0 in the environment switches the packing pool off. */
int  gtb_ctx_transfer_stats(const gtb_ctx *ctx, int64_t *h2d_bytes, int64_t *d2h_bytes, int64_t *packed_chunks, int64_t *raw_chunks);
int  gtb_ctx_profile(gtb_ctx *ctx, int enable);
int  gtb_ctx_profile_report(gtb_ctx *ctx, char *buf, size_t buf_size);

/* ---- overlap count / coverage ------------------------------------------------------------- */
/* Builds the device index for an in-memory reference region set.
 * Stands in for the UnsortedGenomicRegionSetOverlaps / SortedGenomicRegionSetOverlaps constructors
 * (genomic_intervals.cpp:5593-5675, :5807-5816) over a GenomicRegionSet loaded with
 * load_in_memory=true (genomic_overlaps.cpp:412).  `regions` are host arrays.  op is GTB_OP_COUNT or
 * GTB_OP_COVERAGE; flags carries GTB_MATCH_GAPS / GTB_IGNORE_STRAND and optionally an engine bit.
 * GTB_ERR_INDEX_REGION: *err_index is the first malformed region (the reference is fatal, :5607). */
int  gtb_index_create(gtb_ctx *ctx, const gtb_set *regions, int op, unsigned flags,
                      gtb_index **out, int64_t *err_index);
void gtb_index_destroy(gtb_index *index);
/* Forget all queries seen so far (values back to zero). */
int  gtb_index_reset(gtb_index *index);
/* Streams one batch of query regions through the index and accumulates; asynchronous on the
 * context's stream.  Stands in for one stretch of the query loop
 * `for (qreg=GetQuery(); qreg; qreg=NextQuery())` of CountIndexOverlaps / CalcIndexCoverage
 * (genomic_intervals.cpp:5310-5314, :5275-5282).  mem is GTB_MEM_HOST or GTB_MEM_DEVICE.
 * Batches may be any size; region indices reported in errors count across batches. */
int  gtb_index_add_queries(gtb_index *index, const gtb_set *queries, unsigned mem);
/* The same for a batch of single-interval, unweighted reads of ONE length in the compact form a parser can write directly:
 * 5 bytes per read instead of the 13 of a gtb_set (read k = chromosome meta[k] & 0x7F, strand '-' if meta[k] & 0x80 else '+',
 * [start[k], start[k] + read_len - 1]).  It is the form the library itself re-encodes host-resident gtb_set chunks into before
 * they cross the host link; a caller that has it already saves the library the pass over 13 bytes per read.  Reads it cannot
 * express (chromosome ids >= 128, other strands, other lengths) go through gtb_index_add_queries. */
typedef struct {
  int64_t n;
  const int32_t *start;          /* [n] */
  const uint8_t *meta;           /* [n] chromosome id | 0x80 for the '-' strand */
  int32_t read_len;              /* >= 1 */
} gtb_packed_reads;
int  gtb_index_add_packed(gtb_index *index, const gtb_packed_reads *reads, unsigned mem);
/* Completes the accumulation and writes one value per index region, in index-file order:
 * the `unsigned long *hits` / `*coverage` array CountIndexOverlaps / CalcIndexCoverage return
 * (genomic_intervals.cpp:5308, :5273).  out is host memory unless mem == GTB_MEM_DEVICE.
 * A fatal query condition (GTB_ERR_QUERY_*) is reported here with *err_index = the first offending
 * query region in stream order; out is then unspecified (the reference prints nothing). */
int  gtb_index_finish(gtb_index *index, uint64_t *out, unsigned mem, int64_t *err_index);
/* The same in two halves, for callers that have more device work to queue behind the values (the multi-GPU driver's
 * all-gather): gtb_index_finish_async enqueues everything and returns at once (out must be device memory and is valid in
 * stream order); gtb_index_status then waits for the stream and reports what gtb_index_finish would have. */
int  gtb_index_finish_async(gtb_index *index, uint64_t *out, unsigned mem);
int  gtb_index_status(gtb_index *index, int64_t *err_index);

/* The per-QUERY dual: for every query region the number of index regions it overlaps -- the length of the
 * GetOverlap / NextOverlap walk that `genomic_overlaps subset` (is it empty?) and `overlap` (print the query once per step)
 * make for it (genomic_overlaps.cpp:782-800, :706-739; GenomicRegionSetOverlaps::GetOverlap, genomic_intervals.cpp:5224-5250).
 * `index` must have been created with GTB_OP_COUNT (its flags say -gaps / -i / -S rules); it is not modified, and queries
 * streamed with gtb_index_add_queries are unaffected.  out[k] belongs to query region k (host or device memory: out_mem).
 * Synchronous.  Fatal query conditions are reported as by gtb_index_finish. */
int  gtb_index_query_counts(gtb_index *index, const gtb_set *queries, unsigned mem, uint32_t *out, unsigned out_mem, int64_t *err_index);
/* The walk itself: the index regions (0-based numbers in index-file order) that GetOverlap / NextOverlap hand out for every
 * query region, in the order the reference's engine hands them out -- what `genomic_overlaps overlap -label` prints as
 * "query-label:index-label", one line per step (genomic_overlaps.cpp:718-731).  The Unsorted class walks its bin levels in
 * turn, the bins of a level from the query's first to its last, every bin's chain from the region inserted last to the first
 * (genomic_intervals.cpp:5665-5669, :5729-5764): bin_bits[0..n_bin_bits) are the shift-bits of the levels (the -B option; NULL =
 * the default 17,20,23,26; the closing level of 60 bits is added as the reference does, :5637).  Under GTB_SORTED_RULES the
 * matches of a query come in index-file order (the Sorted class's buffer, :5902-5927).
 * match_offset[k] .. match_offset[k + 1] says where query k's matches go in `matches`, relative to match_offset[0]: the
 * running sum of gtb_index_query_counts' result for the same queries (n_regions + 1 entries; GTB_ERR_ARG if it disagrees).
 * Queries and outputs are host memory.  Synchronous.  Fatal query conditions are reported as by gtb_index_finish. */
int  gtb_index_query_matches(gtb_index *index, const gtb_set *queries, unsigned mem, const int *bin_bits, int n_bin_bits,
                             const int64_t *match_offset, int32_t *matches, int64_t *err_index);

/* One-shot conveniences == create + add + finish + destroy.
 * gtb_overlap_count    <-> GenomicRegionSetOverlaps::CountIndexOverlaps(match_gaps, ignore_strand, max_label_value)  genomic_intervals.h:2471, .cpp:5304-5317
 * gtb_overlap_coverage <-> GenomicRegionSetOverlaps::CalcIndexCoverage(match_gaps, ignore_strand, max_label_value)   genomic_intervals.h:2453, .cpp:5269-5285 */
int  gtb_overlap_count(gtb_ctx *ctx, const gtb_set *queries, unsigned queries_mem, const gtb_set *regions,
                       unsigned flags, uint64_t *out, int64_t *err_index);
int  gtb_overlap_coverage(gtb_ctx *ctx, const gtb_set *queries, unsigned queries_mem, const gtb_set *regions,
                          unsigned flags, uint64_t *out, int64_t *err_index);

/* ---- several GPUs behind one index (one process, one host thread per device) --------------------------------------------- */
/* Results are sums over queries (genomic_intervals.cpp:5310-5314): every device holds the whole index, a batch of
 * host-resident queries is cut into one contiguous slice per device -- no routing, each slice crosses its own host link --
 * and gtb_mgpu_index_finish adds the devices' values up (the query-sharded decomposition; the genome-sharded one,
 * for queries that are already resident on the devices, is the torch.distributed driver of python/gtb200/sharded.py).
 * Errors carry the stream-order index of the first offending query region over all slices.  devices[k] may repeat a device
 * (several contexts on one GPU: what the single-GPU tests do). */
typedef struct gtb_mgpu gtb_mgpu;
typedef struct gtb_mgpu_index gtb_mgpu_index;
int  gtb_mgpu_create(int n_devices, const int *devices, gtb_mgpu **out);
void gtb_mgpu_destroy(gtb_mgpu *mg);
int  gtb_mgpu_device_count(const gtb_mgpu *mg);
gtb_ctx *gtb_mgpu_ctx(gtb_mgpu *mg, int k);
const char *gtb_mgpu_last_error(const gtb_mgpu *mg);
int  gtb_mgpu_index_create(gtb_mgpu *mg, const gtb_set *regions, int op, unsigned flags, gtb_mgpu_index **out, int64_t *err_index);
void gtb_mgpu_index_destroy(gtb_mgpu_index *index);
int  gtb_mgpu_index_reset(gtb_mgpu_index *index);
int  gtb_mgpu_index_add_queries(gtb_mgpu_index *index, const gtb_set *queries);          /* host arrays */
int  gtb_mgpu_index_add_packed(gtb_mgpu_index *index, const gtb_packed_reads *reads);    /* host arrays */
int  gtb_mgpu_index_finish(gtb_mgpu_index *index, uint64_t *out, int64_t *err_index);    /* host array */

/* ---- sliding-window read counts ----------------------------------------------------------- */
typedef struct {
  int64_t win_step;        /* -d, genomic_scans.cpp:120 */
  int64_t win_size;        /* -w, genomic_scans.cpp:119; must be a multiple of win_step */
  int64_t min_reads;       /* -min, genomic_scans.cpp:121 */
  int32_t op;              /* '1' = interval start, 'c' = interval centre (-op, genomic_scans.cpp:117) */
  int32_t ignore_strand;   /* -i */
  int32_t emulate_sorted;  /* 0: UnsortedGenomicRegionSetScanner semantics incl. its spurious window on
                              chromosomes shorter than a window; 1: SortedGenomicRegionSetScanner values */
  int32_t reserved;
} gtb_scan_params;

/* Allocates the per-chromosome, per-strand micro-window histograms on the device.
 * Stands in for the UnsortedGenomicRegionSetScanner constructor's allocation
 * (genomic_intervals.cpp:5024-5033) given ReadBounds() output (genomic_intervals.cpp:5997-6015):
 * bound[c] = STOP of chromosome c in the genome file, < 0 if c is not in the genome file. */
int  gtb_scan_create(gtb_ctx *ctx, int32_t n_chrom, const int64_t *bound, const gtb_scan_params *params,
                     gtb_scan **out);
void gtb_scan_destroy(gtb_scan *scan);
int  gtb_scan_reset(gtb_scan *scan);
/* Histogram pass over a batch of reads (every interval of every region counts once,
 * genomic_intervals.cpp:5038-5053).  Asynchronous on the context's stream. */
int  gtb_scan_add_reads(gtb_scan *scan, const gtb_set *reads, unsigned mem);
/* Sliding sum of win_size/win_step micro-windows (genomic_intervals.cpp:5058-5075) followed by the
 * `v >= MIN_READS` filter of RunCounts (genomic_scans.cpp:421-428), compacted on the device in the
 * reference's output order.  *n_windows = number of qualifying windows. */
int  gtb_scan_finish(gtb_scan *scan, int64_t *n_windows);
/* Copies qualifying windows [first, first+count) to host arrays: chromosome id, strand ('+'/'-'),
 * 1-based window number k (interval = [win_step*(k-1)+1, win_step*(k-1)+win_size],
 * genomic_intervals.cpp:5109-5112) and value -- what successive Scanner::Next() calls return
 * (genomic_intervals.h:2213-2217).  Any output pointer may be NULL. */
int  gtb_scan_fetch(gtb_scan *scan, int64_t first, int64_t count, int32_t *chrom, int8_t *strand,
                    int64_t *win, int64_t *value);

/* ---- peak statistics on two window scanners ------------------------------------------------ */
/* PeakFinder::Run (genomic_scans.cpp:298-368): the signal scanner and the control scanner (same genome, same window
 * parameters) are walked window by window; v1 / v2 = the two window values clamped to the window size (:303-305), optionally
 * equalised (-norm, :307); a window with v1 >= min_reads gets the tail probabilities pval1 (signal against background or
 * control) and pval2 (the other way round) of the chosen method (:310-352, GSL's gsl_cdf_binomial_Q / _poisson_Q /
 * _ugaussian_Q, computed here in double precision on the device) and is kept if pval1 <= pval_cutoff (:354).
 * control == NULL: v2 is a Poisson variate of mean win_size * p_signal (:301; the reference seeds its generator with the time,
 * here `seed`).  The kept windows stay on the device in scan order; gtb_scan_peaks_fetch copies a range of them out. */
enum { GTB_PEAKS_BINOMIAL = 0, GTB_PEAKS_POISSON = 1, GTB_PEAKS_BINOMIAL2 = 2, GTB_PEAKS_CBINOMIAL = 3, GTB_PEAKS_NORMAL = 4 };
typedef struct {
  int32_t method;          /* -M */
  int32_t compare;         /* -cmp */
  int32_t norm;            /* -norm */
  int32_t reserved;
  int64_t min_reads;       /* -min */
  int64_t n_signal_reads, n_control_reads;   /* CountGenomicRegions of the two files (:252, :259) */
  double p_signal, p_control;                /* reads / effective genome size (:254, :261) */
  double pval_cutoff;      /* -pval */
  uint64_t seed;
} gtb_peaks_params;
int  gtb_scan_peaks(gtb_scan *signal, gtb_scan *control, const gtb_peaks_params *params, int64_t *n_windows);
int  gtb_scan_peaks_fetch(gtb_scan *signal, int64_t first, int64_t count, int32_t *chrom, int8_t *strand, int64_t *win,
                          double *pval1, double *pval2);

/* ---- global sort of a region set ----------------------------------------------------------- */
/* The order GenomicRegionSet::RunGlobalSort prints a region set in (genomic_regions gsort, genomic_intervals.cpp:4547-4570 with
 * BinGenomicRegions :6095-6150 and CompareBinnedGenomicRegions :6045-6049): chromosome rank ascending (the caller ranks the
 * names in strcmp order, the reference's std::map order), with by_strand the '+' regions of a chromosome before all others,
 * START ascending, STOP descending, equal keys in input order.  Per region: the rank of its chromosome, the START of its first
 * interval and the STOP of its last one after the reference's r->Sort() (intervals by start), and its strand byte -- host arrays.
 * perm[k] = input index of the region printed k-th (host array).  Device LSD radix sort, 8-bit digits, stable. */
int  gtb_sort_regions(gtb_ctx *ctx, int64_t n_regions, const int32_t *chrom_rank, const int32_t *start, const int32_t *stop,
                      const int8_t *strand, int by_strand, int64_t *perm);

/* ---- multi-GPU merge ---------------------------------------------------------------------- */
/* out[k] = table[index[k]], k < n, all device pointers: puts the per-shard value vectors an all-gather has laid side by side
 * (every region is owned by one shard) into index-file order -- the `hits[ireg->n_line]` indexing of
 * genomic_intervals.cpp:5312 across shards.  Runs on cuda_stream (a cudaStream_t, e.g. the one the collective was queued
 * on) or, if NULL, on the context's stream. */
int  gtb_gather_u64(gtb_ctx *ctx, const uint64_t *table, const int64_t *index, int64_t n, uint64_t *out, void *cuda_stream);

/* ---- synthetic inputs (bench / tests) ------------------------------------------------------ */
/* Fills DEVICE arrays with reads [first, first+n) of the counter-based generator documented in
 * DESIGN.md ("Synthetic inputs"); identical to tests/support.py:synth_reads. */
int  gtb_synth_reads(gtb_ctx *ctx, uint64_t seed, int64_t first, int64_t n, int32_t read_len,
                     int32_t n_chrom, const int64_t *chrom_len /*host*/, int32_t *d_chrom, int32_t *d_start,
                     int32_t *d_stop, int8_t *d_strand);

/* The same generator restricted to effective positions p in [p_lo, p_hi) of the concatenated genome
 * (p counts valid read starts: chromosome c contributes chrom_len[c] - read_len + 1 of them) -- the reads
 * of one genome shard of the multi-GPU driver.  p_hi is clamped to the genome end. */
int  gtb_synth_reads_range(gtb_ctx *ctx, uint64_t seed, int64_t first, int64_t n, int32_t read_len,
                           int32_t n_chrom, const int64_t *chrom_len /*host*/, uint64_t p_lo, uint64_t p_hi,
                           int32_t *d_chrom, int32_t *d_start, int32_t *d_stop, int8_t *d_strand);

/* ---- linking consecutive regions of a sorted stream ------------------------------------------ */
/* GenomicRegionSet::RunGlobalLink (genomic_regions link, genomic_intervals.cpp:4607-4644) for max_difference >= 0: the n
 * single-interval regions of a stream sorted by chromosome / (strand) / start -- group_rank[k] = the rank of region k's
 * chromosome (and strand, if the stream is sorted by strand) in stream order, non-decreasing -- fall into *n_linked linked
 * regions; linked region j begins at region head_index[j] and reaches linked_stop[j].  Host arrays; head_index and
 * linked_stop have room for n entries.  Prefix-maximum scan + prefix sum on the device (csrc/gtb_link.cu). */
int  gtb_link_regions(gtb_ctx *ctx, int64_t n, const int32_t *group_rank, const int32_t *start, const int32_t *stop,
                      int64_t max_difference, int64_t *n_linked, int64_t *head_index, int32_t *linked_stop);

#ifdef __cplusplus
}
#endif
#endif
