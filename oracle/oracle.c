/* oracle.c -- scalar CPU restatement of the reference's overlap/coverage/window-count path.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  Written from the behaviour of GenomicTools 2.8.1a,
 * file:line citations are into /root/reference/gtools/.  The product never links this file.
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* region helpers                                                                              */
/* ------------------------------------------------------------------------------------------ */

static inline int64_t reg_lo(const orc_set *s, int64_t k) { return s->region_offset ? s->region_offset[k] : k; }
static inline int64_t reg_hi(const orc_set *s, int64_t k) { return s->region_offset ? s->region_offset[k + 1] : k + 1; }
static inline int64_t reg_weight(const orc_set *s, int64_t k) { return s->weight ? (int64_t)s->weight[k] : 1; }

/* GenomicRegion::IsCompatibleSortedAndNonoverlapping (genomic_intervals.cpp:1153-1161, with
 * IsCompatible :1116-1122 and IsCompatibleSorted :1140-1147): every interval shares the first
 * interval's chromosome and strand, starts never decrease, and each interval begins after the
 * previous one ends. */
static int region_is_well_formed(const orc_set *s, int64_t k) {
  int64_t lo = reg_lo(s, k), hi = reg_hi(s, k);
  for (int64_t i = lo + 1; i < hi; i++) {
    if (s->chrom[i] != s->chrom[lo] || s->strand[i] != s->strand[lo]) return 0;
    if (s->start[i] < s->start[i - 1]) return 0;
    if (s->start[i] <= s->stop[i - 1]) return 0;
  }
  return 1;
}

/* GenomicInterval::OverlapsWith (genomic_intervals.cpp:624-630) lifted to regions
 * (GenomicRegion::OverlapsWith :1167-1172): any pair of intervals intersects. */
static int regions_overlap(const orc_set *a, int64_t ka, const orc_set *b, int64_t kb, int ignore_strand) {
  for (int64_t i = reg_lo(a, ka); i < reg_hi(a, ka); i++)
    for (int64_t j = reg_lo(b, kb); j < reg_hi(b, kb); j++) {
      if (a->chrom[i] != b->chrom[j]) continue;
      if (!ignore_strand && a->strand[i] != b->strand[j]) continue;
      if (a->start[i] > b->stop[j] || a->stop[i] < b->start[j]) continue;
      return 1;
    }
  return 0;
}

/* GenomicInterval::CalcOverlap (genomic_intervals.cpp:427-432) summed over all interval pairs
 * (GenomicRegion::CalcOverlap :1196-1202). */
static int64_t regions_overlap_length(const orc_set *a, int64_t ka, const orc_set *b, int64_t kb, int ignore_strand) {
  int64_t total = 0;
  for (int64_t i = reg_lo(a, ka); i < reg_hi(a, ka); i++)
    for (int64_t j = reg_lo(b, kb); j < reg_hi(b, kb); j++) {
      if (a->chrom[i] != b->chrom[j]) continue;
      if (!ignore_strand && a->strand[i] != b->strand[j]) continue;
      int64_t hi = a->stop[i] < b->stop[j] ? a->stop[i] : b->stop[j];
      int64_t lo = a->start[i] > b->start[j] ? a->start[i] : b->start[j];
      if (hi - lo + 1 > 0) total += hi - lo + 1;
    }
  return total;
}

/* ------------------------------------------------------------------------------------------ */
/* multi-level bin index over the index set (UnsortedGenomicRegionSetOverlaps ctor,            */
/* genomic_intervals.cpp:5593-5675): levels shift 17,20,23,26,60; a region goes to the first   */
/* level where start and stop share a bin; regions in a bin form a LIFO chain.                 */
/* ------------------------------------------------------------------------------------------ */

#define N_LEVELS 5
static const int LEVEL_BITS[N_LEVELS] = {17, 20, 23, 26, 60};

typedef struct {
  int      present;             /* chromosome has at least one indexable region */
  int64_t  n_bins[N_LEVELS];
  int64_t *head[N_LEVELS];      /* newest region in the bin, -1 if empty */
} chrom_bins;

typedef struct {
  int32_t     n_chrom;
  chrom_bins *chr;
  int64_t    *next;             /* chain link per region */
} bin_index;

static void bin_index_free(bin_index *ix) {
  if (ix->chr) {
    for (int32_t c = 0; c < ix->n_chrom; c++)
      for (int l = 0; l < N_LEVELS; l++) free(ix->chr[c].head[l]);
    free(ix->chr);
  }
  free(ix->next);
}

static int region_is_indexable(const orc_set *s, int64_t k) {
  int64_t start = s->start[reg_lo(s, k)], stop = s->stop[reg_hi(s, k) - 1];
  return !(start > stop || stop <= 0);                                    /* :5610, :5659 */
}

static int bin_index_build(bin_index *ix, const orc_set *s, int64_t *err_index) {
  memset(ix, 0, sizeof(*ix));
  int32_t max_chrom = -1;
  for (int64_t i = 0; i < s->n_intervals; i++) if (s->chrom[i] > max_chrom) max_chrom = s->chrom[i];
  ix->n_chrom = max_chrom + 1;
  ix->chr = (chrom_bins *)calloc((size_t)(ix->n_chrom > 0 ? ix->n_chrom : 1), sizeof(chrom_bins));
  ix->next = (int64_t *)malloc(sizeof(int64_t) * (size_t)(s->n_regions > 0 ? s->n_regions : 1));
  int64_t *chrom_size = (int64_t *)calloc((size_t)(ix->n_chrom > 0 ? ix->n_chrom : 1), sizeof(int64_t));

  for (int64_t k = 0; k < s->n_regions; k++) {
    ix->next[k] = -1;
    if (!region_is_well_formed(s, k)) { *err_index = k; free(chrom_size); return ORC_ERR_INDEX_REGION; }  /* :5607 */
    if (!region_is_indexable(s, k)) continue;
    int32_t c = s->chrom[reg_lo(s, k)];
    int64_t stop = s->stop[reg_hi(s, k) - 1];
    if (!ix->chr[c].present) { ix->chr[c].present = 1; chrom_size[c] = stop; }
    else if (stop > chrom_size[c]) chrom_size[c] = stop;                   /* :5611-5613 */
  }
  for (int32_t c = 0; c < ix->n_chrom; c++) {
    if (!ix->chr[c].present) continue;
    for (int l = 0; l < N_LEVELS; l++) {
      int64_t nb = (chrom_size[c] >> LEVEL_BITS[l]) + 1;                    /* :5645 */
      ix->chr[c].n_bins[l] = nb;
      ix->chr[c].head[l] = (int64_t *)malloc(sizeof(int64_t) * (size_t)nb);
      for (int64_t b = 0; b < nb; b++) ix->chr[c].head[l][b] = -1;
    }
  }
  for (int64_t k = 0; k < s->n_regions; k++) {
    if (!region_is_indexable(s, k)) continue;
    int64_t start = s->start[reg_lo(s, k)], stop = s->stop[reg_hi(s, k) - 1];
    if (start <= 0) start = 1;                                              /* :5660 */
    chrom_bins *cb = &ix->chr[s->chrom[reg_lo(s, k)]];
    for (int l = 0; l < N_LEVELS; l++) {
      int64_t b0 = start >> LEVEL_BITS[l], b1 = stop >> LEVEL_BITS[l];
      if (b0 == b1) {                                                       /* :5665-5669 */
        if (cb->head[l][b0] != -1) ix->next[k] = cb->head[l][b0];
        cb->head[l][b0] = k;
        break;
      }
    }
  }
  free(chrom_size);
  return ORC_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* the query loop shared by count and coverage                                                 */
/* ------------------------------------------------------------------------------------------ */

static int overlap_engine(const orc_set *q, const orc_set *idx, unsigned flags, int want_coverage,
                          uint64_t *out, int64_t *err_index) {
  if (!q || !idx || !out || !err_index) return ORC_ERR_ARG;
  const int match_gaps = (flags & ORC_MATCH_GAPS) != 0;
  const int ignore_strand = (flags & ORC_IGNORE_STRAND) != 0;
  *err_index = -1;
  bin_index ix;
  int rc = bin_index_build(&ix, idx, err_index);
  if (rc != ORC_OK) { bin_index_free(&ix); return rc; }
  for (int64_t k = 0; k < idx->n_regions; k++) out[k] = 0;                  /* :5309 / :5274 */

  for (int64_t n = 0; n < q->n_regions; n++) {                              /* :5310 / :5275 */
    if (!region_is_well_formed(q, n)) { *err_index = n; rc = ORC_ERR_QUERY_REGION; break; }  /* :5698,:5709 */
    int64_t qlo = reg_lo(q, n), qhi = reg_hi(q, n);
    int32_t c = q->chrom[qlo];
    if (c < 0 || c >= ix.n_chrom || !ix.chr[c].present) continue;           /* :5719-5720, :5731 */
    const chrom_bins *cb = &ix.chr[c];
    int64_t start = q->start[qlo], stop = q->stop[qhi - 1];
    if (stop <= 0) { *err_index = n; rc = ORC_ERR_QUERY_STOP_NONPOSITIVE; break; }           /* :5740 */
    if (start > stop) { *err_index = n; rc = ORC_ERR_QUERY_START_GT_STOP; break; }           /* :5741 */
    if (start <= 0) start = 1;                                              /* :5742 */
    if ((start >> LEVEL_BITS[0]) >= cb->n_bins[0]) continue;                /* :5745 */
    int64_t w = reg_weight(q, n);
    for (int l = 0; l < N_LEVELS; l++) {                                    /* :5748-5762 */
      int64_t b = start >> LEVEL_BITS[l];
      int64_t b_last = stop >> LEVEL_BITS[l];
      if (b_last > cb->n_bins[l] - 1) b_last = cb->n_bins[l] - 1;
      for (; b <= b_last; b++) {
        for (int64_t k = cb->head[l][b]; k != -1; k = ix.next[k]) {
          int64_t ilo = reg_lo(idx, k), ihi = reg_hi(idx, k);
          if (!(start <= idx->stop[ihi - 1] && stop >= idx->start[ilo])) continue;          /* :5752 */
          if (!(match_gaps || regions_overlap(q, n, idx, k, ignore_strand))) continue;      /* :5227 */
          if (!ignore_strand && q->strand[qlo] != idx->strand[ilo]) continue;               /* :5229 */
          if (!want_coverage) {
            out[k] += (uint64_t)w;                                          /* :5312 */
          } else {
            int64_t cc;
            if (match_gaps) {                                               /* :5277 */
              int64_t hi = q->stop[qhi - 1] < idx->stop[ihi - 1] ? q->stop[qhi - 1] : idx->stop[ihi - 1];
              int64_t lo = q->start[qlo] > idx->start[ilo] ? q->start[qlo] : idx->start[ilo];
              cc = hi - lo + 1;
            } else {
              cc = regions_overlap_length(idx, k, q, n, ignore_strand);
            }
            out[k] += (uint64_t)(cc * w);                                   /* :5278-5279 */
          }
        }
      }
    }
  }
  bin_index_free(&ix);
  return rc;
}

int orc_overlap_count(const orc_set *queries, const orc_set *index, unsigned flags,
                      uint64_t *out, int64_t *err_index) {
  return overlap_engine(queries, index, flags, 0, out, err_index);
}

int orc_overlap_coverage(const orc_set *queries, const orc_set *index, unsigned flags,
                         uint64_t *out, int64_t *err_index) {
  return overlap_engine(queries, index, flags, 1, out, err_index);
}

/* ------------------------------------------------------------------------------------------ */
/* sliding-window counts                                                                       */
/* ------------------------------------------------------------------------------------------ */

int64_t orc_scan_counts(const orc_set *reads, int32_t n_chrom, const int64_t *bound,
                        int64_t win_step, int64_t win_size, char op, int ignore_strand,
                        int64_t min_reads, int emulate_sorted, int64_t cap,
                        int32_t *out_chrom, int8_t *out_strand, int64_t *out_win, int64_t *out_value) {
  if (!reads || !bound || win_step <= 0 || win_size <= 0) return -ORC_ERR_ARG;
  if (win_size % win_step != 0) return -ORC_ERR_WINDOW;                     /* :4845 */
  if (op != '1' && op != 'c') return -ORC_ERR_ARG;                           /* :5046 */
  const int64_t combine = win_size / win_step;                              /* :4846 */
  const int n_strands = ignore_strand ? 1 : 2;

  /* micro-window histograms, 1-based like the reference's v[1..n]            :5024-5033 */
  uint64_t ***v = (uint64_t ***)calloc((size_t)(n_chrom > 0 ? n_chrom : 1), sizeof(uint64_t **));
  int64_t *n_micro = (int64_t *)calloc((size_t)(n_chrom > 0 ? n_chrom : 1), sizeof(int64_t));
  for (int32_t c = 0; c < n_chrom; c++) {
    if (bound[c] < 0) continue;
    n_micro[c] = bound[c] / win_step;
    v[c] = (uint64_t **)calloc(2, sizeof(uint64_t *));
    for (int z = 0; z < n_strands; z++) v[c][z] = (uint64_t *)calloc((size_t)n_micro[c] + 2, sizeof(uint64_t));
  }
  for (int64_t k = 0; k < reads->n_regions; k++) {                          /* :5038-5053 */
    int64_t w = reg_weight(reads, k);
    for (int64_t i = reg_lo(reads, k); i < reg_hi(reads, k); i++) {
      int64_t s = reads->start[i], e = reads->stop[i];
      if (s > e || e <= 0) continue;                                        /* :5040 */
      int32_t c = reads->chrom[i];
      if (c < 0 || c >= n_chrom || bound[c] < 0) continue;                  /* :5041-5042 */
      int64_t pos = (op == '1') ? s : s + (e - s) / 2;                      /* :5044-5045 */
      int64_t win = (pos - 1) / win_step + 1;                               /* :5047 */
      int z = (ignore_strand || reads->strand[i] == '+') ? 0 : 1;           /* :5048 */
      if (pos >= 1 && win <= n_micro[c]) v[c][z][win] += (uint64_t)w;       /* :5049 */
    }
  }
  /* emit in map order: chromosome ascending, '+' then '-', window ascending  :5125-5141 */
  int64_t n_out = 0;
  int first_slot = 1;
  for (int32_t c = 0; c < n_chrom; c++) {
    if (bound[c] < 0) continue;
    for (int z = 0; z < n_strands; z++) {
      const uint64_t *h = v[c][z];
      int64_t n_win = n_micro[c] < combine ? 0 : n_micro[c] - combine + 1;  /* :5061-5064 */
      if (n_win == 0) {
        /* The unsorted scanner's Next() steps onto the new chromosome/strand with
         * current_win = 1 and returns current_v[1] without re-checking current_n (:5128-5140):
         * one spurious window carrying the un-summed first micro-window.  The very first slot
         * is entered through Init() with current_win = 0 and is skipped properly.  With n == 0
         * the reference reads one element past its allocation; 0 is what it yields in practice. */
        if (!emulate_sorted && !first_slot) {
          int64_t val = n_micro[c] >= 1 ? (int64_t)h[1] : 0;
          if (val >= min_reads) {
            if (n_out < cap) { out_chrom[n_out] = c; out_strand[n_out] = z ? '-' : '+'; out_win[n_out] = 1; out_value[n_out] = val; }
            n_out++;
          }
        }
      } else {
        uint64_t sum = 0;
        for (int64_t j = 1; j <= combine - 1; j++) sum += h[j];             /* :5066 */
        for (int64_t k = 1; k <= n_win; k++) {                              /* :5068-5073 */
          sum += h[k + combine - 1];
          int64_t val = (int64_t)sum;
          if (val >= min_reads) {
            if (n_out < cap) { out_chrom[n_out] = c; out_strand[n_out] = z ? '-' : '+'; out_win[n_out] = k; out_value[n_out] = val; }
            n_out++;
          }
          sum -= h[k];
        }
      }
      first_slot = 0;
    }
  }
  for (int32_t c = 0; c < n_chrom; c++) {
    if (!v[c]) continue;
    for (int z = 0; z < 2; z++) free(v[c][z]);
    free(v[c]);
  }
  free(v); free(n_micro);
  return n_out;
}
