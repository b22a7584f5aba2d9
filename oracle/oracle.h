/* oracle.h -- CPU restatement of the GenomicTools 2.8.1a overlap / coverage / window-count path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product library (libgtb200.so), the CLI drivers or
 * the python binding may include, link or call this.  It exists so that tests/, the smoke test
 * and bench.py's cpu_baseline leg have an independent, scalar, single-threaded statement of what
 * the reference computes, on the same packed SoA arrays the C-ABI takes.
 *
 * Pinning: the reference ships no tests or golden vectors for this path, so
 * this restatement is pinned against the reference BINARIES built by oracle/Makefile into
 * oracle/_ref/ (tests/test_oracle_vs_reference.py, tests/golden/).
 *
 * All coordinates are the reference's internal ones: 1-based, closed (genomic_intervals.h:333-337).
 */
#ifndef GTB200_ORACLE_H
#define GTB200_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* A set of regions in struct-of-arrays form.  Region k owns intervals
 * [region_offset[k], region_offset[k+1]) ; region_offset == NULL means every region has exactly
 * one interval (n_intervals == n_regions). */
typedef struct {
  int64_t n_regions;
  int64_t n_intervals;
  const int32_t *chrom;         /* per interval: chromosome id                                   */
  const int32_t *start;         /* per interval: START (1-based, closed)                         */
  const int32_t *stop;          /* per interval: STOP                                            */
  const int8_t  *strand;        /* per interval: '+', '-', or the raw GFF character              */
  const int32_t *weight;        /* per region: GetLabelValue() result; NULL => 1                 */
  const int64_t *region_offset; /* n_regions+1 entries or NULL                                   */
} orc_set;

enum { ORC_MATCH_GAPS = 1u, ORC_IGNORE_STRAND = 2u };

enum {
  ORC_OK = 0,
  ORC_ERR_ARG = 1,
  ORC_ERR_QUERY_STOP_NONPOSITIVE = 2, /* "stop position must be positive!"                    genomic_intervals.cpp:5740 */
  ORC_ERR_QUERY_START_GT_STOP = 3,    /* "start position cannot be greater than stop position!" genomic_intervals.cpp:5741 */
  ORC_ERR_QUERY_REGION = 4,           /* "query regions should be compatible, sorted and non-overlapping!" :5698,:5709 */
  ORC_ERR_INDEX_REGION = 5,           /* "index regions should be compatible, sorted and non-overlapping!" :5607 */
  ORC_ERR_WINDOW = 6                  /* win_size % win_step != 0                              genomic_intervals.cpp:4845 */
};

/* hits[k] per index region, index-file order.  Follows
 * GenomicRegionSetOverlaps::CountIndexOverlaps (genomic_intervals.cpp:5304-5317) driven by
 * UnsortedGenomicRegionSetOverlaps (genomic_intervals.cpp:5593-5764).
 * On a fatal condition returns the code above and *err_index = 0-based index of the offending
 * region (query or index set); out[] is then unspecified (the reference prints nothing). */
int orc_overlap_count(const orc_set *queries, const orc_set *index, unsigned flags,
                      uint64_t *out, int64_t *err_index);

/* coverage[k] per index region (total overlapping nucleotides), following
 * GenomicRegionSetOverlaps::CalcIndexCoverage (genomic_intervals.cpp:5269-5285). */
int orc_overlap_coverage(const orc_set *queries, const orc_set *index, unsigned flags,
                         uint64_t *out, int64_t *err_index);

/* Sliding-window read counts, following UnsortedGenomicRegionSetScanner
 * (ctor genomic_intervals.cpp:5019-5080, Next :5125-5141) and the printing loop of RunCounts
 * (genomic_scans.cpp:421-428).
 *   bound[c]  : STOP of chromosome c in the genome file, or < 0 if c is not in the genome file.
 *               Chromosome ids are assumed to be in strcmp order of their names, so ascending id
 *               is the reference's std::map iteration order.
 *   op        : '1' (interval start) or 'c' (interval centre)
 *   emulate_sorted: 0 = default (unsorted) scanner including its one-spurious-window quirk for
 *               chromosomes shorter than a window; 1 = values as SortedGenomicRegionSetScanner
 *               emits them (no spurious windows).
 * Emits, in reference output order, every window with value >= min_reads:
 *   out_chrom[i], out_strand[i] ('+'/'-'), out_win[i] (k, 1-based; interval is
 *   [win_step*(k-1)+1, win_step*(k-1)+win_size]), out_value[i].
 * Returns the number of windows that qualify (may exceed cap; only the first cap are stored), or
 * a negative ORC_ERR_* code. */
int64_t orc_scan_counts(const orc_set *reads, int32_t n_chrom, const int64_t *bound,
                        int64_t win_step, int64_t win_size, char op, int ignore_strand,
                        int64_t min_reads, int emulate_sorted, int64_t cap,
                        int32_t *out_chrom, int8_t *out_strand, int64_t *out_win, int64_t *out_value);

#ifdef __cplusplus
}
#endif
#endif
