// gtb_overlap.cu -- overlap count / coverage: index construction, the general RANK and ENUMERATE
// engines, finalisation, and the streaming C-ABI around them (include/gtb200.h).
#include "gtb_rank_device.cuh"
#include <chrono>
#include <algorithm>
#include <limits.h>

// =================================================================================================
// device helpers
// =================================================================================================
namespace {

__device__ __forceinline__ int64_t lower_bound_u64(const ull *__restrict__ p, int64_t n, ull x) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (__ldg(p + mid) < x) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// Shared front half of both engines: the per-query-region checks the reference performs, in its
// order.  Returns false if the region contributes nothing (or was reported as fatal).
//   well-formedness: GenomicRegion::IsCompatibleSortedAndNonoverlapping, genomic_intervals.cpp:1153-1161,
//                    applied to every query in GetQuery/NextQuery (:5698, :5709)
//   chromosome unknown to the index -> no matches, no checks (:5719-5720, :5731)
//   stop <= 0, start > stop -> fatal (:5740-5741)
__device__ __forceinline__ bool admit_query(const QueryView &q, int64_t r, const RankView &rv, int64_t &lo, int64_t &hi) {
  const uint8_t *__restrict__ present = rv.chrom_present;
  const int32_t n_chrom = rv.n_chrom;
  ull *err = rv.err;
  if (q.region_offset) { lo = q.region_offset[r] - q.interval_base; hi = q.region_offset[r + 1] - q.interval_base; }
  else { lo = r; hi = r + 1; }
  if (hi <= lo) return false;
  const int32_t c = q.chrom[lo];
  const int8_t sb = q.strand[lo];
  for (int64_t i = lo + 1; i < hi; i++) {
    if (q.chrom[i] != c || q.strand[i] != sb || q.start[i] < q.start[i - 1] || q.start[i] <= q.stop[i - 1]) {
      report_error(err, q.index_base + r, GTB_ERR_QUERY_REGION);
      return false;
    }
  }
  if (c < 0 || c >= n_chrom || !present[c]) return false;
  const int32_t qs = q.start[lo], qe = q.stop[hi - 1];
  return admit_interval(rv, qs, qe, q.index_base + r);
}

// -------------------------------------------------------------------------------------------------
// RANK engine, general form: binary search in global memory, 64-bit atomics on the slot histograms.
// One thread per query region.  Handles weights, any interval length, negative coordinates,
// multi-interval queries (spans under -gaps, blocks for coverage).
// -------------------------------------------------------------------------------------------------
template <bool COVERAGE, bool BLOCKS>
__global__ void __launch_bounds__(256) rank_accumulate_kernel(QueryView q, RankView ix) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < q.n_regions; r += stride) {
    int64_t lo, hi;
    if (!admit_query(q, r, ix, lo, hi)) continue;
    const int cls = ix.class_of[(uint8_t)q.strand[lo]];            // front-interval strand, :5229
    if (cls < 0) continue;
    const int g = q.chrom[lo] * ix.n_class + cls;
    const int gb = ix.goff[g], ge = ix.goff[g + 1];
    if (ge == gb) continue;
    const int64_t w = q.weight ? (int64_t)q.weight[r] : 1;
    if (!BLOCKS) {
      rank_item<COVERAGE>(ix, gb, ge, q.start[lo], q.stop[hi - 1], w);   // span: count, or -gaps (:5227, :5277)
    } else {
      for (int64_t i = lo; i < hi; i++) {                           // coverage = sum over block pairs (:1196-1202)
        const int32_t bs = q.start[i], be = q.stop[i];
        if (bs > be) continue;                                      // CalcOverlap clamps such blocks to 0 (:427-432)
        rank_item<COVERAGE>(ix, gb, ge, bs, be, w);
      }
    }
  }
}

// -------------------------------------------------------------------------------------------------
// ENUMERATE engine: candidates from the (chromosome, level, bin) CSR, exact predicate per pair.
// -------------------------------------------------------------------------------------------------
constexpr int ENUM_LEVELS = 6;
__constant__ int c_enum_bits[ENUM_LEVELS] = {14, 17, 20, 23, 26, 62};

__device__ __forceinline__ ull enum_key(int32_t c, int level, int64_t bin) {
  return ((ull)(uint32_t)c << 35) | ((ull)level << 32) | (ull)(uint32_t)bin;
}

// region_list (with n_list entries): only those query regions, by number (what the one-pass engine has left over, see
// accumulate_multi_fast); null: all of them
template <bool COVERAGE>
__global__ void __launch_bounds__(128) enumerate_kernel(QueryView q, RankView ix, EnumView ev, bool match_gaps, bool ignore_strand,
                                                        const int32_t *__restrict__ region_list, int64_t n_list) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n_work = region_list ? n_list : q.n_regions;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_work; t += stride) {
    const int64_t r = region_list ? (int64_t)region_list[t] : t;
    int64_t lo, hi;
    if (!admit_query(q, r, ix, lo, hi)) continue;
    const int32_t c = q.chrom[lo];
    const int8_t qstrand = q.strand[lo];
    const int64_t qs_true = q.start[lo], qe = q.stop[hi - 1];
    const int64_t qs = qs_true <= 0 ? 1 : qs_true;                  // :5742
    const int64_t w = q.weight ? (int64_t)q.weight[r] : 1;
    for (int l = 0; l < ENUM_LEVELS; l++) {
      const ull k_lo = enum_key(c, l, qs >> c_enum_bits[l]);
      const ull k_hi = enum_key(c, l, qe >> c_enum_bits[l]);
      for (int64_t e = lower_bound_u64(ev.keys, ev.n_entries, k_lo); e < ev.n_entries && ev.keys[e] <= k_hi; e++) {
        const int32_t k = ev.rid[e];
        const int64_t ilo = ev.r_off[k], ihi = ev.r_off[k + 1];
        if (!(qs <= ev.r_stop[ihi - 1] && qe >= ev.r_start[ilo])) continue;             // span test, :5752
        if (!ignore_strand && qstrand != ev.r_strand[ilo]) continue;                   // :5229
        int64_t cc = 0;
        bool any = false;
        if (match_gaps) {                                                              // :5227, :5277
          any = true;
          const int64_t a = qe < ev.r_stop[ihi - 1] ? qe : ev.r_stop[ihi - 1];
          const int64_t b = qs_true > ev.r_start[ilo] ? qs_true : ev.r_start[ilo];
          cc = a - b + 1;
        } else {
          for (int64_t i = lo; i < hi; i++)
            for (int64_t j = ilo; j < ihi; j++) {
              // chromosome and (unless -i) strand are region-wide after the well-formedness checks
              const int64_t a = q.stop[i] < ev.r_stop[j] ? q.stop[i] : ev.r_stop[j];
              const int64_t b = q.start[i] > ev.r_start[j] ? q.start[i] : ev.r_start[j];
              if (!(q.start[i] > ev.r_stop[j] || q.stop[i] < ev.r_start[j])) any = true;   // :624-630
              if (a - b + 1 > 0) cc += a - b + 1;                                      // :427-432
            }
        }
        if (!any) continue;
        atomicAdd(ev.direct + k, COVERAGE ? (ull)(cc * w) : (ull)w);
      }
    }
  }
}

// -------------------------------------------------------------------------------------------------
// Per-QUERY overlap counts: the dual of the engines above, for `genomic_overlaps subset / overlap`
// (genomic_overlaps.cpp:706-739, :782-800: the GetOverlap / NextOverlap walk of one query region).
//   n(q) = #{targets: ts <= qe} - #{targets: te < qs}                   (targets valid: ts <= te)
//        = #{(ts - 1) points below qe} - #{te points below qs}
// i.e. two prefix counts over the group's sorted evaluation points, looked up at the slots of qs and qe.
// `pre[j]` = (x: targets whose ts - 1 lies in a slot before j, y: targets whose te does), group-relative.
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) query_counts_kernel(QueryView q, RankView ix, const uint2 *__restrict__ pre, uint32_t *__restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < q.n_regions; r += stride) {
    uint32_t n = 0;
    int64_t lo, hi;
    if (admit_query(q, r, ix, lo, hi)) {
      const int cls = ix.class_of[(uint8_t)q.strand[lo]];
      if (cls >= 0) {
        const int g = q.chrom[lo] * ix.n_class + cls;
        const int gb = ix.goff[g], ge = ix.goff[g + 1];
        if (ge > gb) {
          const int32_t qs = q.start[lo], qe = q.stop[hi - 1];
          const int jS = lower_bound_i32(ix.points, gb, ge - 1, qs);
          const int jE = lower_bound_i32(ix.points, qe < qs ? gb : jS, ge - 1, qe);
          GTB_ASSERT(jS >= gb && jS < ge && jE >= gb && jE < ge);
          n = pre[jE].x - pre[jS].y;
        }
      }
    }
    out[r] = n;
  }
}

// the same where ranks cannot speak (multi-interval regions on either side without -gaps): candidates, exact predicate per pair
__global__ void __launch_bounds__(128) query_enumerate_kernel(QueryView q, RankView ix, EnumView ev, bool ignore_strand, uint32_t *__restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < q.n_regions; r += stride) {
    uint32_t n = 0;
    int64_t lo, hi;
    if (admit_query(q, r, ix, lo, hi)) {
      const int32_t c = q.chrom[lo];
      const int8_t qstrand = q.strand[lo];
      const int64_t qs_true = q.start[lo], qe = q.stop[hi - 1];
      const int64_t qs = qs_true <= 0 ? 1 : qs_true;                  // :5742
      for (int l = 0; l < ENUM_LEVELS; l++) {
        const ull k_lo = enum_key(c, l, qs >> c_enum_bits[l]);
        const ull k_hi = enum_key(c, l, qe >> c_enum_bits[l]);
        for (int64_t e = lower_bound_u64(ev.keys, ev.n_entries, k_lo); e < ev.n_entries && ev.keys[e] <= k_hi; e++) {
          const int32_t k = ev.rid[e];
          const int64_t ilo = ev.r_off[k], ihi = ev.r_off[k + 1];
          if (!(qs <= ev.r_stop[ihi - 1] && qe >= ev.r_start[ilo])) continue;             // span test, :5752
          if (!ignore_strand && qstrand != ev.r_strand[ilo]) continue;                   // :5229
          bool any = false;
          for (int64_t i = lo; i < hi && !any; i++)
            for (int64_t j = ilo; j < ihi; j++)
              if (!(q.start[i] > ev.r_stop[j] || q.stop[i] < ev.r_start[j])) { any = true; break; }   // :624-630
          n += any ? 1u : 0u;
        }
      }
    }
    out[r] = n;
  }
}

// The walk itself: for every query region the index regions GetOverlap / NextOverlap hand out, IN THE ORDER the reference's
// engine hands them out.  The Unsorted class (:5729-5764) walks its bin levels in turn, the bins of a level from the query's
// first to its last, and every bin's chain from the region inserted last to the one inserted first (:5665-5669) -- the entries
// of `ev` are sorted that way (key ascending, region number descending; levels = the reference's -B list, see build_match_structures),
// so the walk below meets the matches in that order.  The Sorted class (:5902-5927) steps through its buffer of index regions,
// which holds them in file order: the matches of a query are put into ascending order after the walk (by_id).
// match_off[r] - match_base .. match_off[r + 1] - match_base: where query r's matches go (from gtb_index_query_counts).
struct MatchLevels { int n; int bits[8]; };

__global__ void __launch_bounds__(128) query_matches_kernel(QueryView q, RankView ix, EnumView ev, MatchLevels lv, bool match_gaps, bool ignore_strand, bool by_id,
                                                            const int64_t *__restrict__ match_off, int64_t match_base, int32_t *__restrict__ match_out,
                                                            uint32_t *__restrict__ mismatch) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < q.n_regions; r += stride) {
    const int64_t room = match_off[r + 1] - match_off[r];
    int32_t *dst = match_out + (match_off[r] - match_base);
    GTB_ASSERT(room >= 0 && match_off[r] >= match_base);
    int64_t n = 0;
    int64_t lo, hi;
    if (admit_query(q, r, ix, lo, hi)) {
      const int32_t c = q.chrom[lo];
      const int8_t qstrand = q.strand[lo];
      const int64_t qs = q.start[lo], qe = q.stop[hi - 1];
      const int64_t b_first = max(min(qs, qe), (int64_t)1), b_last = max(max(qs, qe), (int64_t)1);   // :5742; the Sorted class admits qs == qe + 1
      for (int l = 0; l < lv.n; l++) {
        const ull k_lo = enum_key(c, l, b_first >> lv.bits[l]);
        const ull k_hi = enum_key(c, l, b_last >> lv.bits[l]);
        for (int64_t e = lower_bound_u64(ev.keys, ev.n_entries, k_lo); e < ev.n_entries && ev.keys[e] <= k_hi; e++) {
          const int32_t k = ev.rid[e];
          const int64_t ilo = ev.r_off[k], ihi = ev.r_off[k + 1];
          if (!(qs <= ev.r_stop[ihi - 1] && qe >= ev.r_start[ilo])) continue;             // span test, :5752
          if (!ignore_strand && qstrand != ev.r_strand[ilo]) continue;                   // :5229
          bool any = match_gaps;                                                          // :5227
          for (int64_t i = lo; i < hi && !any; i++)
            for (int64_t j = ilo; j < ihi; j++)
              if (!(q.start[i] > ev.r_stop[j] || q.stop[i] < ev.r_start[j])) { any = true; break; }   // :624-630
          if (!any) continue;
          if (n < room) dst[n] = k;
          n++;
        }
      }
    }
    if (n != room) { atomicOr(mismatch, 1u); continue; }
    if (by_id)
      for (int64_t a = 1; a < n; a++) {
        const int32_t v = dst[a];
        int64_t b = a;
        for (; b > 0 && dst[b - 1] > v; b--) dst[b] = dst[b - 1];
        dst[b] = v;
      }
  }
}

// -------------------------------------------------------------------------------------------------
// finalisation: after the slot histograms have been prefix-summed, evaluate every target and sum
// the targets of each region.  One thread per region.
// -------------------------------------------------------------------------------------------------
template <bool COVERAGE>
__global__ void __launch_bounds__(256) finalize_kernel(int64_t n_regions, const int64_t *__restrict__ t_off,
                                                       const int32_t *__restrict__ t_hi, const int32_t *__restrict__ t_lo,
                                                       const int32_t *__restrict__ t_group, const int32_t *__restrict__ goff,
                                                       const int32_t *__restrict__ points, const ull *__restrict__ scan, int64_t K,
                                                       const ull *__restrict__ direct, ull *__restrict__ out) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_regions) return;
  ull total = direct ? direct[k] : 0ull;
  for (int64_t t = t_off[k]; t < t_off[k + 1]; t++) {
    const int hi = t_hi[t], lo = t_lo[t], g = t_group[t], base = goff[g];
    GTB_ASSERT(hi >= base && hi < K && lo >= base && lo < K);
    // group-relative inclusive prefix of slot plane p at slot j
    auto pre = [&](int p, int j) -> ull {
      const ull *a = scan + (int64_t)p * K;
      return a[j] - (base > 0 ? a[base - 1] : 0ull);
    };
    if (!COVERAGE) {
      const ull starts_le_te = pre(H_BOTH, hi) + pre(H_SCNT, hi);   // #{qs <= te}
      const ull stops_lt_ts = pre(H_BOTH, lo) + pre(H_ECNT, lo);    // #{qe <= ts-1}
      total += starts_le_te - stops_lt_ts;
    } else {
      auto F = [&](int j) -> ull {
        const ull x = (ull)(int64_t)points[j];
        const ull L = pre(H_BOTH, j);
        const ull cs = pre(H_SCNT, j), ss = pre(H_SSUM, j);
        const ull ce = pre(H_ECNT, j), se = pre(H_ESUM, j);
        return L + (x + 1ull) * cs - ss - x * ce + se;
      };
      total += F(hi) - F(lo);
    }
  }
  out[k] = total;
}

// -------------------------------------------------------------------------------------------------
// Prepass for batches of multi-interval query regions, so that the single-interval engines can serve them:
//   SPANS (count / coverage with -gaps, :5227, :5277): checks each region as the reference does and writes its span
//     [first start, last stop] as one single-interval query; the engines then apply the usual admission to the spans.
//   !SPANS (coverage without -gaps): coverage is additive over the pairs of blocks (:1196-1202), so the engines take the
//     batch's intervals as they lie, one query per block, under admission mode 2 (no errors).  What the reference decides per
//     REGION -- well-formedness (:5698, :5709) and the fatal span conditions on chromosomes the index knows (:5740-5741) -- is
//     decided here.  A region on a chromosome the index has never seen contributes nothing either way.
// -------------------------------------------------------------------------------------------------
template <bool SPANS>
__global__ void __launch_bounds__(256) region_prepass_kernel(QueryView q, RankView rv, int32_t *__restrict__ o_chrom, int32_t *__restrict__ o_start,
                                                             int32_t *__restrict__ o_stop, int8_t *__restrict__ o_strand) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < q.n_regions; r += stride) {
    const int64_t lo = q.region_offset[r] - q.interval_base, hi = q.region_offset[r + 1] - q.interval_base;
    bool ok = hi > lo;
    int32_t c = -1, qs = 1, qe = 0;
    int8_t sb = 0;
    if (ok) {
      c = q.chrom[lo]; sb = q.strand[lo]; qs = q.start[lo]; qe = q.stop[hi - 1];
      for (int64_t i = lo + 1; i < hi; i++)
        if (q.chrom[i] != c || q.strand[i] != sb || q.start[i] < q.start[i - 1] || q.start[i] <= q.stop[i - 1]) { ok = false; break; }
      if (!ok) report_error(rv.err, q.index_base + r, GTB_ERR_QUERY_REGION);
    }
    if (SPANS) {
      // a region that is empty or malformed leaves a query that no engine counts and none objects to: an unknown chromosome
      o_chrom[r] = ok ? c : -1; o_start[r] = qs; o_stop[r] = qe; o_strand[r] = sb;
    } else if (ok && (uint32_t)c < (uint32_t)rv.n_chrom && rv.chrom_present[c]) {
      admit_interval(rv, qs, qe, q.index_base + r);                    // reports; the blocks themselves are never an error
    }
  }
}

// region_offset of a batch whose regions all have k intervals (the caller said so instead of sending offsets): written out only
// for the paths that read offsets
__global__ void __launch_bounds__(256) uniform_offsets_kernel(int64_t n_regions, int64_t k, int64_t base, int64_t *__restrict__ off) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r <= n_regions; r += stride) off[r] = base + r * k;
}

// Expands a packed chunk (start + meta of W bytes, see gtb_ingest.cpp) into the SoA layout the engines read.
template <int W>
__device__ __forceinline__ void unpack_one(uint32_t m, uint32_t len0, int32_t s, int32_t &c, int32_t &e, uint32_t &sb) {
  if (W == 4) { c = (int32_t)((m >> 16) & 0x3FFFu); e = s + (int32_t)(m & 0xFFFFu); sb = (m >> 30) ? 45u : 43u; }
  if (W == 2) { c = (int32_t)((m >> 8) & 0x7Fu); e = s + (int32_t)(m & 0xFFu); sb = (m >> 15) ? 45u : 43u; }
  if (W == 1) { c = (int32_t)(m & 0x7Fu); e = s + (int32_t)len0; sb = (m >> 7) ? 45u : 43u; }
}

template <int W>
__global__ void __launch_bounds__(256) unpack_kernel(int64_t n, const void *__restrict__ meta_v, uint32_t len0, const int32_t *__restrict__ start,
                                                     int32_t *__restrict__ chrom, int32_t *__restrict__ stop, int8_t *__restrict__ strand) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    uint32_t m[4];
    const int lim = (int)min((int64_t)4, n - i);
    if (lim == 4) {
      if (W == 4) { const uint4 v = *reinterpret_cast<const uint4 *>((const uint32_t *)meta_v + i); m[0] = v.x; m[1] = v.y; m[2] = v.z; m[3] = v.w; }
      if (W == 2) { const uint2 v = *reinterpret_cast<const uint2 *>((const uint16_t *)meta_v + i); m[0] = v.x & 0xFFFFu; m[1] = v.x >> 16; m[2] = v.y & 0xFFFFu; m[3] = v.y >> 16; }
      if (W == 1) { const uint32_t v = *reinterpret_cast<const uint32_t *>((const uint8_t *)meta_v + i); m[0] = v & 0xFFu; m[1] = (v >> 8) & 0xFFu; m[2] = (v >> 16) & 0xFFu; m[3] = v >> 24; }
      const int4 s = *reinterpret_cast<const int4 *>(start + i);
      int32_t c[4], e[4];
      uint32_t sb[4];
      unpack_one<W>(m[0], len0, s.x, c[0], e[0], sb[0]); unpack_one<W>(m[1], len0, s.y, c[1], e[1], sb[1]);
      unpack_one<W>(m[2], len0, s.z, c[2], e[2], sb[2]); unpack_one<W>(m[3], len0, s.w, c[3], e[3], sb[3]);
      *reinterpret_cast<int4 *>(chrom + i) = make_int4(c[0], c[1], c[2], c[3]);
      *reinterpret_cast<int4 *>(stop + i) = make_int4(e[0], e[1], e[2], e[3]);
      *reinterpret_cast<uint32_t *>(strand + i) = sb[0] | (sb[1] << 8) | (sb[2] << 16) | (sb[3] << 24);
    } else {
      for (int64_t j = i; j < n; j++) {
        const uint32_t mj = W == 4 ? ((const uint32_t *)meta_v)[j] : W == 2 ? (uint32_t)((const uint16_t *)meta_v)[j] : (uint32_t)((const uint8_t *)meta_v)[j];
        int32_t c, e;
        uint32_t sb;
        unpack_one<W>(mj, len0, start[j], c, e, sb);
        chrom[j] = c; stop[j] = e; strand[j] = (int8_t)sb;
      }
    }
  }
}

}  // namespace

// =================================================================================================
// host side: index construction
// =================================================================================================
static bool region_well_formed(const gtb_set *s, int64_t k) {
  const int64_t lo = s->region_offset ? s->region_offset[k] : k, hi = s->region_offset ? s->region_offset[k + 1] : k + 1;
  for (int64_t i = lo + 1; i < hi; i++) {
    if (s->chrom[i] != s->chrom[lo] || s->strand[i] != s->strand[lo]) return false;
    if (s->start[i] < s->start[i - 1] || s->start[i] <= s->stop[i - 1]) return false;
  }
  return true;
}

template <typename T>
static int upload(gtb_ctx *ctx, dbuf<T> &d, const std::vector<T> &h) {
  GTB_TRY(d.reserve(ctx, h.size() ? h.size() : 1));
  if (h.size()) GTB_CUDA_OK(ctx, cudaMemcpyAsync(d.p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  return GTB_OK;
}

static int build_rank_structures(gtb_index *ix) {
  gtb_ctx *ctx = ix->ctx;
  const int64_t M = ix->n_regions;
  const bool use_blocks = ix->op == GTB_OP_COVERAGE && !ix->match_gaps;
  auto lo_of = [&](int64_t k) { return ix->h_off[k]; };
  auto hi_of = [&](int64_t k) { return ix->h_off[k + 1]; };
  auto indexable = [&](int64_t k) {
    if (hi_of(k) <= lo_of(k)) return false;
    if (!ix->h_malformed.empty() && ix->h_malformed[(size_t)k]) return false;
    int64_t s = ix->h_start[lo_of(k)], e = ix->h_stop[hi_of(k) - 1];
    // the Sorted class skips nothing.  Regions with stop <= 0 are exact in ranks against any admitted query; regions with
    // start > stop are not (a zero-length region against a zero-length query at the same place) and keep the value 0:
    // a documented divergence ("Known divergences")
    if (ix->sorted_rules) return s <= e;
    return !(s > e || e <= 0);                                        // :5610, :5659
  };
  // chromosome table and strand classes
  int32_t max_chrom = -1;
  for (int64_t k = 0; k < M; k++) if (indexable(k)) max_chrom = std::max(max_chrom, ix->h_chrom[lo_of(k)]);
  ix->n_chrom = max_chrom + 1;
  ix->h_present.assign((size_t)std::max(ix->n_chrom, 1), 0);
  ix->h_class_of.assign(256, (int8_t)(ix->ignore_strand ? 0 : -1));
  ix->n_class = ix->ignore_strand ? 1 : 0;
  for (int64_t k = 0; k < M; k++) {
    if (!indexable(k)) continue;
    if (ix->h_chrom[lo_of(k)] < 0) return gtb_fail(ctx, GTB_ERR_ARG, "negative chromosome id in index set");
    ix->h_present[ix->h_chrom[lo_of(k)]] = 1;
    uint8_t sb = (uint8_t)ix->h_strand[lo_of(k)];
    if (!ix->ignore_strand && ix->h_class_of[sb] < 0) {
      if (ix->n_class >= 120) return gtb_fail(ctx, GTB_ERR_UNSUPPORTED, "too many distinct strand characters");
      ix->h_class_of[sb] = (int8_t)ix->n_class++;
    }
  }
  if (ix->n_class == 0) ix->n_class = 1;
  if ((int64_t)ix->n_chrom * ix->n_class > (int64_t)INT_MAX / 4) return gtb_fail(ctx, GTB_ERR_UNSUPPORTED, "too many chromosome groups");
  ix->n_groups = ix->n_chrom * ix->n_class;

  // targets
  struct target { int32_t g; int32_t ts, te; };
  std::vector<target> targets;
  std::vector<int64_t> t_off((size_t)M + 1, 0);
  for (int64_t k = 0; k < M; k++) {
    t_off[k] = (int64_t)targets.size();
    if (!indexable(k)) continue;
    const int32_t g = ix->h_chrom[lo_of(k)] * ix->n_class + ix->h_class_of[(uint8_t)ix->h_strand[lo_of(k)]];
    if (!use_blocks) targets.push_back({g, ix->h_start[lo_of(k)], ix->h_stop[hi_of(k) - 1]});
    else for (int64_t i = lo_of(k); i < hi_of(k); i++)
      if (ix->h_start[i] <= ix->h_stop[i]) targets.push_back({g, ix->h_start[i], ix->h_stop[i]});
  }
  t_off[M] = (int64_t)targets.size();
  ix->n_targets = (int64_t)targets.size();

  // evaluation points per group: {te} U {ts-1}, sorted, distinct, + sentinel
  std::vector<std::pair<int32_t, int32_t>> pts;     // (group, x)
  pts.reserve(targets.size() * 2);
  for (auto &t : targets) {
    pts.push_back({t.g, t.te});
    pts.push_back({t.g, t.ts == INT32_MIN ? INT32_MIN : t.ts - 1});
  }
  std::sort(pts.begin(), pts.end());
  pts.erase(std::unique(pts.begin(), pts.end()), pts.end());
  ix->h_goff.assign((size_t)ix->n_groups + 1, 0);
  ix->h_points.clear();
  ix->h_points.reserve(pts.size() + 64);
  {
    size_t i = 0;
    for (int32_t g = 0; g < ix->n_groups; g++) {
      ix->h_goff[g] = (int32_t)ix->h_points.size();
      const size_t begin = i;
      while (i < pts.size() && pts[i].first == g) ix->h_points.push_back(pts[i++].second);
      if (i > begin) ix->h_points.push_back(INT32_MAX);              // sentinel slot
    }
    ix->h_goff[ix->n_groups] = (int32_t)ix->h_points.size();
  }
  ix->n_slots = (int64_t)ix->h_points.size();
  if (ix->n_slots > (int64_t)INT_MAX - 8) return gtb_fail(ctx, GTB_ERR_UNSUPPORTED, "too many evaluation points");

  std::vector<int32_t> t_hi(targets.size()), t_lo(targets.size()), t_group(targets.size());
  for (size_t t = 0; t < targets.size(); t++) {
    const int32_t gb = ix->h_goff[targets[t].g], ge = ix->h_goff[targets[t].g + 1];
    const int32_t *b = ix->h_points.data() + gb, *e = ix->h_points.data() + ge - 1;
    const int32_t lo_x = targets[t].ts == INT32_MIN ? INT32_MIN : targets[t].ts - 1;
    t_hi[t] = (int32_t)(std::lower_bound(b, e, targets[t].te) - ix->h_points.data());
    t_lo[t] = (int32_t)(std::lower_bound(b, e, lo_x) - ix->h_points.data());
    t_group[t] = targets[t].g;
  }

  GTB_TRY(upload(ctx, ix->d_class_of, ix->h_class_of));
  GTB_TRY(upload(ctx, ix->d_present, ix->h_present));
  GTB_TRY(upload(ctx, ix->d_goff, ix->h_goff));
  GTB_TRY(upload(ctx, ix->d_points, ix->h_points));
  GTB_TRY(upload(ctx, ix->d_t_hi, t_hi));
  GTB_TRY(upload(ctx, ix->d_t_lo, t_lo));
  GTB_TRY(upload(ctx, ix->d_t_base, t_group));
  GTB_TRY(upload(ctx, ix->d_t_off, t_off));
  ix->planes = ix->op == GTB_OP_COVERAGE ? H_PLANES_COVERAGE : H_PLANES_COUNT;
  const size_t hist_elems = (size_t)ix->planes * (size_t)std::max<int64_t>(ix->n_slots, 1);
  GTB_TRY(ix->d_hist.reserve(ctx, hist_elems));
  GTB_TRY(ix->d_hist_scan.reserve(ctx, hist_elems));
  GTB_TRY(ix->d_err.reserve(ctx, 1));
  GTB_TRY(ix->d_out.reserve(ctx, (size_t)std::max<int64_t>(M, 1)));
  GTB_TRY(ix->d_direct.reserve(ctx, (size_t)std::max<int64_t>(M, 1)));
  GTB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));               // host vectors above go out of scope
  return GTB_OK;
}

static int build_enum_structures(gtb_index *ix) {
  if (ix->enum_ready) return GTB_OK;
  gtb_ctx *ctx = ix->ctx;
  static const int bits[6] = {14, 17, 20, 23, 26, 62};
  std::vector<std::pair<ull, int32_t>> ent;
  ent.reserve((size_t)ix->n_regions);
  for (int64_t k = 0; k < ix->n_regions; k++) {
    const int64_t lo = ix->h_off[k], hi = ix->h_off[k + 1];
    if (hi <= lo) continue;
    if (!ix->h_malformed.empty() && ix->h_malformed[(size_t)k]) continue;
    int64_t s = ix->h_start[lo], e = ix->h_stop[hi - 1];
    if (s > e || e <= 0) continue;
    if (s <= 0) s = 1;                                                // :5660
    for (int l = 0; l < 6; l++)
      if ((s >> bits[l]) == (e >> bits[l])) {
        ent.push_back({((ull)(uint32_t)ix->h_chrom[lo] << 35) | ((ull)l << 32) | (ull)(uint32_t)(s >> bits[l]), (int32_t)k});
        break;
      }
  }
  std::sort(ent.begin(), ent.end());
  std::vector<ull> keys(ent.size());
  std::vector<int32_t> rid(ent.size());
  for (size_t i = 0; i < ent.size(); i++) { keys[i] = ent[i].first; rid[i] = ent[i].second; }
  ix->n_entries = (int64_t)ent.size();
  GTB_TRY(upload(ctx, ix->d_keys, keys));
  GTB_TRY(upload(ctx, ix->d_rid, rid));
  GTB_TRY(upload(ctx, ix->d_r_chrom, ix->h_chrom));
  GTB_TRY(upload(ctx, ix->d_r_start, ix->h_start));
  GTB_TRY(upload(ctx, ix->d_r_stop, ix->h_stop));
  GTB_TRY(upload(ctx, ix->d_r_strand, ix->h_strand));
  GTB_TRY(upload(ctx, ix->d_r_off, ix->h_off));
  GTB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  ix->enum_ready = true;
  return GTB_OK;
}

extern "C" int gtb_index_reset(gtb_index *ix) {
  if (!ix) return GTB_ERR_ARG;
  gtb_ctx *ctx = ix->ctx;
  GTB_CUDA_OK(ctx, cudaMemsetAsync(ix->d_hist.p, 0, sizeof(ull) * (size_t)ix->planes * (size_t)std::max<int64_t>(ix->n_slots, 1), ctx->stream));
  GTB_CUDA_OK(ctx, cudaMemsetAsync(ix->d_direct.p, 0, sizeof(ull) * (size_t)std::max<int64_t>(ix->n_regions, 1), ctx->stream));
  GTB_CUDA_OK(ctx, cudaMemsetAsync(ix->d_err.p, 0xFF, sizeof(ull), ctx->stream));
  gtb_direct_reset(ix);
  ix->queries_seen = 0;
  return GTB_OK;
}

extern "C" int gtb_index_create(gtb_ctx *ctx, const gtb_set *regions, int op, unsigned flags, gtb_index **out, int64_t *err_index) {
  if (!ctx || !regions || !out) return GTB_ERR_ARG;
  *out = nullptr;
  if (err_index) *err_index = -1;
  if (op != GTB_OP_COUNT && op != GTB_OP_COVERAGE) return gtb_fail(ctx, GTB_ERR_ARG, "op must be GTB_OP_COUNT or GTB_OP_COVERAGE");
  if (regions->n_regions < 0 || regions->n_intervals < 0) return gtb_fail(ctx, GTB_ERR_ARG, "negative sizes");
  if (!regions->region_offset && regions->n_regions != regions->n_intervals)
    return gtb_fail(ctx, GTB_ERR_ARG, "region_offset is NULL but n_regions != n_intervals");
  if (regions->n_intervals > 0 && (!regions->chrom || !regions->start || !regions->stop || !regions->strand))
    return gtb_fail(ctx, GTB_ERR_ARG, "null interval arrays");
  if (regions->region_offset) {
    if (regions->region_offset[0] != 0 || regions->region_offset[regions->n_regions] != regions->n_intervals)
      return gtb_fail(ctx, GTB_ERR_ARG, "region_offset must run from 0 to n_intervals");
    for (int64_t k = 0; k < regions->n_regions; k++)
      if (regions->region_offset[k + 1] < regions->region_offset[k]) return gtb_fail(ctx, GTB_ERR_ARG, "region_offset must be non-decreasing");
  }
  GTB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  // The Unsorted class checks every index region in its constructor (fatal, :5607).  The Sorted class checks a region when the
  // queries reach it (:5845), so under GTB_SORTED_RULES that is the caller's business (the drivers repeat the reference's walk): a
  // malformed region the caller let through was never reached by a query and simply has no value.
  const bool sorted_rules = (flags & GTB_SORTED_RULES) != 0;
  std::vector<uint8_t> malformed;
  for (int64_t k = 0; k < regions->n_regions; k++)
    if (!region_well_formed(regions, k)) {
      if (sorted_rules) { if (malformed.empty()) malformed.assign((size_t)regions->n_regions, 0); malformed[(size_t)k] = 1; continue; }
      if (err_index) *err_index = k;
      return gtb_fail(ctx, GTB_ERR_INDEX_REGION, "index regions should be compatible, sorted and non-overlapping!");
    }
  gtb_index *ix = new gtb_index();
  ix->h_malformed.swap(malformed);
  ix->ctx = ctx; ix->op = op;
  ix->match_gaps = (flags & GTB_MATCH_GAPS) != 0;
  ix->ignore_strand = (flags & GTB_IGNORE_STRAND) != 0;
  ix->sorted_rules = (flags & GTB_SORTED_RULES) != 0;
  ix->engine = flags & (GTB_ENGINE_ENUMERATE | GTB_ENGINE_RANK | GTB_ENGINE_BUCKET | GTB_ENGINE_DIRECT);
  ix->n_regions = regions->n_regions; ix->n_intervals = regions->n_intervals;
  const size_t ni = (size_t)regions->n_intervals;
  ix->h_chrom.assign(regions->chrom, regions->chrom + ni);
  ix->h_start.assign(regions->start, regions->start + ni);
  ix->h_stop.assign(regions->stop, regions->stop + ni);
  ix->h_strand.assign(regions->strand, regions->strand + ni);
  ix->h_off.resize((size_t)regions->n_regions + 1);
  for (int64_t k = 0; k <= regions->n_regions; k++) ix->h_off[k] = regions->region_offset ? regions->region_offset[k] : k;
  for (int64_t k = 0; k < regions->n_regions; k++) if (ix->h_off[k + 1] - ix->h_off[k] > 1) ix->index_multi = true;
  for (auto &st : ix->stages) {
    cudaEventCreateWithFlags(&st.copied, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&st.consumed, cudaEventDisableTiming);
  }
  int rc = build_rank_structures(ix);
  if (rc == GTB_OK) rc = gtb_index_reset(ix);
  if (rc != GTB_OK) { gtb_index_destroy(ix); return rc; }
  *out = ix;
  return GTB_OK;
}

extern "C" void gtb_index_destroy(gtb_index *ix) {
  if (!ix) return;
  cudaSetDevice(ix->ctx->device);
  gtb_ctx_synchronize(ix->ctx);
  gtb_bucket_destroy(ix);
  gtb_direct_destroy(ix);
  ix->d_class_of.release(); ix->d_present.release(); ix->d_goff.release(); ix->d_points.release();
  ix->d_t_hi.release(); ix->d_t_lo.release(); ix->d_t_base.release(); ix->d_t_off.release();
  ix->d_hist.release(); ix->d_hist_scan.release(); ix->d_scan_scratch.release();
  ix->d_keys.release(); ix->d_rid.release(); ix->d_mkeys.release(); ix->d_mrid.release(); ix->d_r_chrom.release(); ix->d_r_start.release();
  ix->d_r_stop.release(); ix->d_r_strand.release(); ix->d_r_off.release(); ix->d_direct.release();
  ix->d_err.release(); ix->d_out.release();
  ix->sp_chrom.release(); ix->sp_start.release(); ix->sp_stop.release(); ix->sp_strand.release(); ix->d_qpre.release();
  for (auto &st : ix->stages) {
    st.chrom.release(); st.start.release(); st.stop.release(); st.weight.release(); st.strand.release(); st.off.release(); st.meta.release();
    if (st.copied) cudaEventDestroy(st.copied);
    if (st.consumed) cudaEventDestroy(st.consumed);
  }
  delete ix;
}

// =================================================================================================
// accumulate
// =================================================================================================
static RankView rank_view(gtb_index *ix) {
  RankView v;
  v.n_chrom = ix->n_chrom; v.n_class = ix->n_class;
  v.class_of = ix->d_class_of.p; v.chrom_present = ix->d_present.p;
  v.goff = ix->d_goff.p; v.points = ix->d_points.p; v.n_slots = ix->n_slots;
  v.hist = ix->d_hist.p; v.err = ix->d_err.p; v.admission = ix->admission();
  return v;
}

// which engine may serve this (index, batch) pair
static unsigned choose_engine(gtb_index *ix, const QueryView &q, bool batch_multi) {
  const bool rank_valid = ix->op == GTB_OP_COVERAGE || ix->match_gaps || (!ix->index_multi && !batch_multi);
  if (ix->engine & GTB_ENGINE_ENUMERATE) return GTB_ENGINE_ENUMERATE;
  if (!rank_valid) return GTB_ENGINE_ENUMERATE;
  if (ix->engine & GTB_ENGINE_RANK) return GTB_ENGINE_RANK;
  // one pass where the index is small enough for byte counters in shared memory and the batch large enough to pay for the
  // per-batch dump of the counters; GTB_ENGINE_DIRECT asks for it whatever the batch size
  if (!(ix->engine & GTB_ENGINE_BUCKET) && ((ix->engine & GTB_ENGINE_DIRECT) || q.n_regions >= (1 << 18)) &&
      gtb_direct_supported(ix, q, batch_multi))
    return GTB_ENGINE_DIRECT;
  if (gtb_bucket_supported(ix, q, batch_multi)) return GTB_ENGINE_BUCKET;
  return GTB_ENGINE_RANK;
}

static int accumulate_device(gtb_index *ix, const QueryView &q, bool batch_multi, int64_t n_intervals, int uniform_k = 0);

// Multi-interval query regions in front of the single-interval engines (see region_prepass_kernel).  Returns GTB_ERR_UNSUPPORTED
// if this batch is not of that kind.
static EnumView enum_view(gtb_index *ix) {
  EnumView ev;
  ev.n_entries = ix->n_entries; ev.keys = ix->d_keys.p; ev.rid = ix->d_rid.p;
  ev.r_chrom = ix->d_r_chrom.p; ev.r_start = ix->d_r_start.p; ev.r_stop = ix->d_r_stop.p;
  ev.r_strand = ix->d_r_strand.p; ev.r_off = ix->d_r_off.p; ev.direct = ix->d_direct.p;
  return ev;
}

static int accumulate_multi_fast(gtb_index *ix, const QueryView &q, int64_t n_intervals) {
  gtb_ctx *ctx = ix->ctx;
  const bool spans = ix->match_gaps;
  if (q.weight || ix->flat_blocks || ix->leftovers) return GTB_ERR_UNSUPPORTED;
  if (!(spans || ix->op == GTB_OP_COVERAGE)) {
    // count without -gaps: a region counts a query once if any of its blocks overlaps it.  Against single-interval index regions
    // that is the count of the query's SPAN unless a region lies strictly inside a gap between blocks -- impossible if the span
    // holds no evaluation point.  So: spans written by the prepass, counted by the one-pass engine, which leaves the
    // multi-interval regions whose span does hold one (a few percent of spliced reads) on a list for the enumeration engine.
    if (ix->index_multi || (ix->engine & (GTB_ENGINE_ENUMERATE | GTB_ENGINE_RANK | GTB_ENGINE_BUCKET)) || q.n_regions < (1 << 18) || !gtb_direct_usable(ix) ||
        getenv("GTB_NO_MULTI_FAST"))
      return GTB_ERR_UNSUPPORTED;
    RankView rv = rank_view(ix);
    const size_t nr = (size_t)q.n_regions;
    GTB_TRY(ix->sp_chrom.reserve(ctx, nr)); GTB_TRY(ix->sp_start.reserve(ctx, nr)); GTB_TRY(ix->sp_stop.reserve(ctx, nr)); GTB_TRY(ix->sp_strand.reserve(ctx, nr));
    GTB_LAUNCH(ctx, "region_spans", region_prepass_kernel<true>, gtb_grid_for(q.n_regions, 256, (int64_t)ctx->sm_count * 16), 256, 0, q, rv,
               ix->sp_chrom.p, ix->sp_start.p, ix->sp_stop.p, ix->sp_strand.p);
    GTB_TRY(gtb_check_launch(ctx));
    QueryView flat = q;
    flat.weight = nullptr; flat.region_offset = nullptr; flat.interval_base = 0;
    flat.chrom = ix->sp_chrom.p; flat.start = ix->sp_start.p; flat.stop = ix->sp_stop.p; flat.strand = ix->sp_strand.p;
    if (!gtb_direct_supported(ix, flat, false)) return GTB_ERR_UNSUPPORTED;
    const int rc = gtb_direct_accumulate(ix, flat, 4, q.region_offset);
    if (rc != GTB_OK) return rc;                                         // (GTB_ERR_UNSUPPORTED: nothing counted, the whole batch by enumeration)
    const int32_t *list = nullptr;
    const int64_t n_ex = gtb_direct_exception_list(ix, &list);
    if (n_ex > 0) {
      GTB_TRY(build_enum_structures(ix));
      const EnumView ev = enum_view(ix);
      GTB_LAUNCH(ctx, "enumerate_count", enumerate_kernel<false>, gtb_grid_for(n_ex, 128, (int64_t)ctx->sm_count * 16), 128, 0, q, rv, ev, false, ix->ignore_strand, list, n_ex);
      GTB_TRY(gtb_check_launch(ctx));
    }
    return GTB_OK;
  }
  if (ix->engine & (GTB_ENGINE_ENUMERATE | GTB_ENGINE_RANK)) return GTB_ERR_UNSUPPORTED;
  // coverage of spans: spans are long and of every length.  DIRECT takes them at a reduction each (its byte counters count reads of
  // one common length); BUCKET's elements have 8 bits for a length, so where DIRECT cannot serve the index the general rank step does
  if (spans && ix->op == GTB_OP_COVERAGE && ((ix->engine & GTB_ENGINE_BUCKET) || q.n_regions < (1 << 18) || !gtb_direct_usable(ix))) return GTB_ERR_UNSUPPORTED;
  if (getenv("GTB_NO_MULTI_FAST")) return GTB_ERR_UNSUPPORTED;          // (tests: the general path on the same input)
  RankView rv = rank_view(ix);
  const unsigned grid = gtb_grid_for(q.n_regions, 256, (int64_t)ctx->sm_count * 16);
  QueryView flat = q;
  flat.weight = nullptr; flat.region_offset = nullptr; flat.interval_base = 0;
  if (spans) {
    const size_t nr = (size_t)q.n_regions;
    GTB_TRY(ix->sp_chrom.reserve(ctx, nr)); GTB_TRY(ix->sp_start.reserve(ctx, nr)); GTB_TRY(ix->sp_stop.reserve(ctx, nr)); GTB_TRY(ix->sp_strand.reserve(ctx, nr));
    GTB_LAUNCH(ctx, "region_spans", region_prepass_kernel<true>, grid, 256, 0, q, rv, ix->sp_chrom.p, ix->sp_start.p, ix->sp_stop.p, ix->sp_strand.p);
    GTB_TRY(gtb_check_launch(ctx));
    flat.chrom = ix->sp_chrom.p; flat.start = ix->sp_start.p; flat.stop = ix->sp_stop.p; flat.strand = ix->sp_strand.p;
    return accumulate_device(ix, flat, false, q.n_regions);              // spans are ordinary queries: the usual admission, errors by region index
  }
  GTB_LAUNCH(ctx, "region_check", region_prepass_kernel<false>, grid, 256, 0, q, rv, (int32_t *)nullptr, (int32_t *)nullptr, (int32_t *)nullptr, (int8_t *)nullptr);
  GTB_TRY(gtb_check_launch(ctx));
  flat.n_regions = n_intervals;
  ix->flat_blocks = true;
  const int rc = accumulate_device(ix, flat, false, n_intervals);
  ix->flat_blocks = false;
  return rc;
}

static int accumulate_device(gtb_index *ix, const QueryView &q_in, bool batch_multi, int64_t n_intervals, int uniform_k) {
  gtb_ctx *ctx = ix->ctx;
  if (q_in.n_regions <= 0) return GTB_OK;
  QueryView q = q_in;
  if (uniform_k > 1) {
    // Regions of k intervals each, no offsets sent.  Read pairs under coverage (k == 2, no -gaps, no weights, an index the DIRECT
    // engine serves): the engine takes the intervals as they lie and checks every pair in its registers -- what the reference
    // decides per region (well-formedness, :5698, :5709; the fatal span conditions, :5740-5741) costs no pass of its own.
    // Under -gaps (count or coverage) the engine takes each pair's span as the query, formed in its registers.  Count without
    // -gaps against single-interval index regions: a region counts a pair once if either mate overlaps it, which differs from
    // the span's count only for regions strictly inside the gap -- a pair whose span holds no evaluation point has none, the
    // other pairs (a percent or two) go to the candidate-enumeration engine afterwards.
    const bool count_no_gaps = ix->op == GTB_OP_COUNT && !ix->match_gaps;
    const bool pairs = uniform_k == 2 && (ix->match_gaps || ix->op == GTB_OP_COVERAGE || !ix->index_multi) && !q.weight && !ix->flat_blocks &&
                       !(ix->engine & (GTB_ENGINE_ENUMERATE | GTB_ENGINE_RANK | GTB_ENGINE_BUCKET)) && n_intervals >= (1 << 18) && !getenv("GTB_NO_MULTI_FAST");
    if (pairs) {
      QueryView flat = q;
      flat.n_regions = n_intervals; flat.weight = nullptr; flat.region_offset = nullptr; flat.interval_base = 0;
      if (gtb_direct_supported(ix, flat, false)) {
        const int mode = ix->match_gaps ? 2 : count_no_gaps ? 3 : 1;
        ix->flat_blocks = mode == 1;
        const int rc = gtb_direct_accumulate(ix, flat, mode);
        ix->flat_blocks = false;
        if (rc == GTB_OK && mode == 3) {
          QueryView ex = q;
          const int64_t n_ex = gtb_direct_exceptions(ix, &ex);
          if (n_ex > 0) {
            ex.index_base = q.index_base;                                // (these pairs have passed every check: nothing to report)
            GTB_TRY(ix->uni_off.reserve(ctx, (size_t)n_ex + 1));
            GTB_LAUNCH(ctx, "uniform_offsets", uniform_offsets_kernel, gtb_grid_for(n_ex + 1, 256, (int64_t)ctx->sm_count * 8), 256, 0, n_ex, (int64_t)2, (int64_t)0, ix->uni_off.p);
            GTB_TRY(gtb_check_launch(ctx));
            ex.region_offset = ix->uni_off.p;
            ix->leftovers = true;                                        // (the engine's exception arrays are this batch: no second round there)
            const int rc_ex = accumulate_device(ix, ex, true, 2 * n_ex, 0);
            ix->leftovers = false;
            return rc_ex;
          }
        }
        if (rc != GTB_ERR_UNSUPPORTED) return rc;
      }
    }
    // every other shape: the offsets are written out and the batch is an ordinary CSR one
    GTB_TRY(ix->uni_off.reserve(ctx, (size_t)q.n_regions + 1));
    GTB_LAUNCH(ctx, "uniform_offsets", uniform_offsets_kernel, gtb_grid_for(q.n_regions + 1, 256, (int64_t)ctx->sm_count * 8), 256, 0, q.n_regions, (int64_t)uniform_k,
               q.interval_base, ix->uni_off.p);
    GTB_TRY(gtb_check_launch(ctx));
    q.region_offset = ix->uni_off.p;
  }
  if (batch_multi) {
    const int rc = accumulate_multi_fast(ix, q, n_intervals);
    if (rc != GTB_ERR_UNSUPPORTED) return rc;
  }
  const unsigned engine = choose_engine(ix, q, batch_multi);
  if (engine == GTB_ENGINE_BUCKET) return gtb_bucket_accumulate(ix, q);
  if (engine == GTB_ENGINE_DIRECT) {
    const int rc = gtb_direct_accumulate(ix, q);
    if (rc != GTB_ERR_UNSUPPORTED || !q.weight) return rc;              // (a weighted batch the engine has just left: the general step below)
  }
  RankView rv = rank_view(ix);
  if (engine == GTB_ENGINE_ENUMERATE) {
    GTB_TRY(build_enum_structures(ix));
    const EnumView ev = enum_view(ix);
    const unsigned grid = gtb_grid_for(q.n_regions, 128, (int64_t)ctx->sm_count * 16);
    if (ix->op == GTB_OP_COVERAGE)
      GTB_LAUNCH(ctx, "enumerate_coverage", enumerate_kernel<true>, grid, 128, 0, q, rv, ev, ix->match_gaps, ix->ignore_strand, (const int32_t *)nullptr, (int64_t)0);
    else
      GTB_LAUNCH(ctx, "enumerate_count", enumerate_kernel<false>, grid, 128, 0, q, rv, ev, ix->match_gaps, ix->ignore_strand, (const int32_t *)nullptr, (int64_t)0);
    return gtb_check_launch(ctx);
  }
  const unsigned grid = gtb_grid_for(q.n_regions, 256, (int64_t)ctx->sm_count * 8);
  const bool blocks = ix->op == GTB_OP_COVERAGE && !ix->match_gaps && batch_multi;
  if (ix->op == GTB_OP_COVERAGE) {
    if (blocks) GTB_LAUNCH(ctx, "rank_coverage_blocks", (rank_accumulate_kernel<true, true>), grid, 256, 0, q, rv);
    else GTB_LAUNCH(ctx, "rank_coverage", (rank_accumulate_kernel<true, false>), grid, 256, 0, q, rv);
  } else {
    GTB_LAUNCH(ctx, "rank_count", (rank_accumulate_kernel<false, false>), grid, 256, 0, q, rv);
  }
  return gtb_check_launch(ctx);
}

extern "C" int gtb_index_add_queries(gtb_index *ix, const gtb_set *queries, unsigned mem) {
  if (!ix || !queries) return GTB_ERR_ARG;
  gtb_ctx *ctx = ix->ctx;
  if (queries->n_regions < 0 || queries->n_intervals < 0) return gtb_fail(ctx, GTB_ERR_ARG, "negative sizes");
  if (queries->n_regions == 0) return GTB_OK;
  // regions of k >= 2 intervals each need no offsets: region r is intervals [r * k, (r + 1) * k)
  int uniform_k = 0;
  if (!queries->region_offset && queries->n_regions != queries->n_intervals) {
    if (queries->n_intervals % queries->n_regions != 0 || queries->n_intervals / queries->n_regions < 2 || queries->n_intervals / queries->n_regions > (1 << 20))
      return gtb_fail(ctx, GTB_ERR_ARG, "region_offset is NULL but n_intervals is not a multiple (2 or more) of n_regions");
    uniform_k = (int)(queries->n_intervals / queries->n_regions);
  }
  if (!queries->chrom || !queries->start || !queries->stop || !queries->strand) return gtb_fail(ctx, GTB_ERR_ARG, "null interval arrays");
  GTB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  // a CSR is judged by its offsets, not by its totals (regions of 0 and 2 intervals would add up to "one each"); one whose
  // regions all have one interval is passed on as the plain single-interval layout.  Device-resident offsets are taken as they are.
  bool batch_multi = false;
  if (queries->region_offset && !(mem & GTB_MEM_DEVICE)) {
    if (queries->region_offset[0] != 0 || queries->region_offset[queries->n_regions] != queries->n_intervals)
      return gtb_fail(ctx, GTB_ERR_ARG, "region_offset must run from 0 to n_intervals");
    for (int64_t k = 0; k < queries->n_regions; k++) {
      const int64_t d = queries->region_offset[k + 1] - queries->region_offset[k];
      if (d < 1) return gtb_fail(ctx, GTB_ERR_ARG, "region_offset must be increasing: every region has at least one interval");
      batch_multi = batch_multi || d != 1;
    }
  } else if (queries->region_offset) {
    batch_multi = true;
  }
  const bool pass_offsets = batch_multi;                                // (offsets exist and matter)
  if (uniform_k > 1) batch_multi = true;

  if (mem & GTB_MEM_DEVICE) {
    QueryView q;
    q.n_regions = queries->n_regions; q.chrom = queries->chrom; q.start = queries->start; q.stop = queries->stop;
    q.strand = queries->strand; q.weight = queries->weight; q.region_offset = pass_offsets ? queries->region_offset : nullptr;
    q.interval_base = 0; q.index_base = ix->queries_seen;
    GTB_TRY(accumulate_device(ix, q, batch_multi, queries->n_intervals, uniform_k));
    ix->queries_seen += queries->n_regions;
    return GTB_OK;
  }

  // host-resident batch: chunk, stage through two device buffers, copies on copy_stream overlap
  // the kernels of the previous chunk on the compute stream.
  const int64_t CHUNK = (int64_t)4 << 20;
  for (int64_t r0 = 0; r0 < queries->n_regions; r0 += CHUNK) {
    const int64_t r1 = std::min(queries->n_regions, r0 + CHUNK);
    const int64_t i0 = pass_offsets ? queries->region_offset[r0] : uniform_k > 1 ? r0 * uniform_k : r0;
    const int64_t i1 = pass_offsets ? queries->region_offset[r1] : uniform_k > 1 ? r1 * uniform_k : r1;
    const size_t nr = (size_t)(r1 - r0), ni = (size_t)(i1 - i0);
    gtb_index::stage &st = ix->stages[ix->next_stage];
    ix->next_stage ^= 1;
    if (st.in_flight) GTB_CUDA_OK(ctx, cudaStreamWaitEvent(ctx->copy_stream, st.consumed, 0));
    GTB_TRY(st.chrom.reserve(ctx, ni)); GTB_TRY(st.start.reserve(ctx, ni)); GTB_TRY(st.stop.reserve(ctx, ni));
    GTB_TRY(st.strand.reserve(ctx, ni));
    cudaStream_t cs = ctx->copy_stream;
    // Packed path (single-interval, unweighted chunks): host threads re-encode 13 B/interval into 8 B/interval in pinned
    // staging while the previous chunk is on the wire; unpack_kernel expands it next to the engine.
    bool packed = false;
    int pack_w = 4;
    uint32_t pack_len0 = 0;
    gtb_pinned_slot *slot = nullptr;
    // Packing pays only while the pool re-encodes faster than the link would move the 5 bytes it saves: 13 B/interval at
    // ~52 GB/s is 4 G intervals/s.  With several ranks per host (few threads each, shared memory bandwidth) it does not, and
    // the chunk goes raw; the rate is re-measured every 64 chunks.
    const bool pack_worthwhile = ctx->pack_rate == 0.0 || ctx->pack_rate > 4.2e9 || (++ctx->pack_skipped % 64) == 0;
    if (!pass_offsets && !queries->weight && ni >= 65536 && pack_worthwhile && gtb_ctx_ingest_ready(ctx, ni, &slot) == GTB_OK) {
      cudaPointerAttributes attr;
      const bool start_is_pinned = cudaPointerGetAttributes(&attr, queries->start + i0) == cudaSuccess && attr.type == cudaMemoryTypeHost;
      cudaGetLastError();
      // narrowest form first: 1 byte of meta per interval (one read length, < 128 chromosomes), then 2, then 4; a chunk that
      // does not fit widens the form for the chunks after it, and a narrower one is tried again every 64 chunks
      if (ctx->pack_width > 1 && ++ctx->pack_wide_chunks >= 64) { ctx->pack_width = 1; ctx->pack_wide_chunks = 0; }
      const auto t0 = std::chrono::steady_clock::now();
      for (int width = ctx->pack_width; width <= 4 && !packed; width *= 2) {
        packed = gtb_ingest_pack_width(ctx->ingest, width, queries->chrom + i0, queries->start + i0, queries->stop + i0, queries->strand + i0,
                                       (int64_t)ni, slot->meta, start_is_pinned ? nullptr : slot->start, &pack_len0) != 0;
        if (packed) pack_w = width;
        else if (width < 4) { ctx->pack_width = width * 2; ctx->pack_wide_chunks = 0; }
      }
      const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      // (the pool's first chunk pays for thread start-up and first-touch of the pinned slots: not a rate to judge it by)
      if (dt > 0 && ctx->pack_timed++ > 0) ctx->pack_rate = ctx->pack_rate == 0.0 ? (double)ni / dt : 0.5 * ctx->pack_rate + 0.5 * (double)ni / dt;
      if (packed) {
        GTB_TRY(st.meta.reserve(ctx, ni));
        GTB_CUDA_OK(ctx, cudaMemcpyAsync(st.start.p, start_is_pinned ? queries->start + i0 : slot->start, ni * 4, cudaMemcpyHostToDevice, cs));
        GTB_CUDA_OK(ctx, cudaMemcpyAsync(st.meta.p, slot->meta, ni * (size_t)pack_w, cudaMemcpyHostToDevice, cs));
        GTB_CUDA_OK(ctx, cudaEventRecord(slot->h2d_done, cs));
        slot->in_flight = true;
        ctx->packed_chunks++;
        ctx->h2d_bytes += (int64_t)ni * (4 + pack_w);
      }
    }
    if (!packed) {
      ctx->raw_chunks++;
      ctx->h2d_bytes += (int64_t)ni * 13;
      GTB_CUDA_OK(ctx, cudaMemcpyAsync(st.chrom.p, queries->chrom + i0, ni * 4, cudaMemcpyHostToDevice, cs));
      GTB_CUDA_OK(ctx, cudaMemcpyAsync(st.start.p, queries->start + i0, ni * 4, cudaMemcpyHostToDevice, cs));
      GTB_CUDA_OK(ctx, cudaMemcpyAsync(st.stop.p, queries->stop + i0, ni * 4, cudaMemcpyHostToDevice, cs));
      GTB_CUDA_OK(ctx, cudaMemcpyAsync(st.strand.p, queries->strand + i0, ni, cudaMemcpyHostToDevice, cs));
    }
    if (queries->weight) {
      GTB_TRY(st.weight.reserve(ctx, nr));
      GTB_CUDA_OK(ctx, cudaMemcpyAsync(st.weight.p, queries->weight + r0, nr * 4, cudaMemcpyHostToDevice, cs));
      ctx->h2d_bytes += (int64_t)nr * 4;
    }
    if (pass_offsets) {
      GTB_TRY(st.off.reserve(ctx, nr + 1));
      GTB_CUDA_OK(ctx, cudaMemcpyAsync(st.off.p, queries->region_offset + r0, (nr + 1) * 8, cudaMemcpyHostToDevice, cs));
      ctx->h2d_bytes += (int64_t)(nr + 1) * 8;
    }
    GTB_CUDA_OK(ctx, cudaEventRecord(st.copied, cs));
    GTB_CUDA_OK(ctx, cudaStreamWaitEvent(ctx->stream, st.copied, 0));
    if (packed) {
      const unsigned ugrid = gtb_grid_for((int64_t)(ni + 3) / 4, 256, (int64_t)ctx->sm_count * 8);
      if (pack_w == 1) GTB_LAUNCH(ctx, "unpack", unpack_kernel<1>, ugrid, 256, 0, (int64_t)ni, (const void *)st.meta.p, pack_len0, st.start.p, st.chrom.p, st.stop.p, st.strand.p);
      else if (pack_w == 2) GTB_LAUNCH(ctx, "unpack", unpack_kernel<2>, ugrid, 256, 0, (int64_t)ni, (const void *)st.meta.p, pack_len0, st.start.p, st.chrom.p, st.stop.p, st.strand.p);
      else GTB_LAUNCH(ctx, "unpack", unpack_kernel<4>, ugrid, 256, 0, (int64_t)ni, (const void *)st.meta.p, pack_len0, st.start.p, st.chrom.p, st.stop.p, st.strand.p);
      GTB_TRY(gtb_check_launch(ctx));
    }
    QueryView q;
    q.n_regions = (int64_t)nr; q.chrom = st.chrom.p; q.start = st.start.p; q.stop = st.stop.p; q.strand = st.strand.p;
    q.weight = queries->weight ? st.weight.p : nullptr; q.region_offset = pass_offsets ? st.off.p : nullptr;
    q.interval_base = i0; q.index_base = ix->queries_seen + r0;
    GTB_TRY(accumulate_device(ix, q, batch_multi, (int64_t)ni, uniform_k));
    GTB_CUDA_OK(ctx, cudaEventRecord(st.consumed, ctx->stream));
    st.in_flight = true;
  }
  ix->queries_seen += queries->n_regions;
  // "copied inside the call" (gtb200.h): on return the caller's arrays are its own again.  Pageable memory has been staged by
  // the runtime by now; pinned memory is read by the DMA engine after cudaMemcpyAsync returns, so the last chunk's copies are
  // waited for (the earlier chunks' were ordered before them on the copy stream).  The kernels stay asynchronous.
  GTB_CUDA_OK(ctx, cudaEventSynchronize(ix->stages[ix->next_stage ^ 1].copied));
  return GTB_OK;
}

// reads in the packed host form (gtb200.h): start + one byte; expanded next to the engine by unpack_kernel<1>
extern "C" int gtb_index_add_packed(gtb_index *ix, const gtb_packed_reads *reads, unsigned mem) {
  if (!ix || !reads) return GTB_ERR_ARG;
  gtb_ctx *ctx = ix->ctx;
  if (reads->n < 0 || reads->read_len < 1) return gtb_fail(ctx, GTB_ERR_ARG, "negative size or read length below 1");
  if (reads->n == 0) return GTB_OK;
  if (!reads->start || !reads->meta) return gtb_fail(ctx, GTB_ERR_ARG, "null arrays");
  GTB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  const uint32_t len0 = (uint32_t)reads->read_len - 1u;
  const int64_t CHUNK = (int64_t)4 << 20;
  for (int64_t r0 = 0; r0 < reads->n; r0 += CHUNK) {
    const size_t ni = (size_t)std::min(reads->n - r0, CHUNK);
    gtb_index::stage &st = ix->stages[ix->next_stage];
    ix->next_stage ^= 1;
    cudaStream_t cs = ctx->copy_stream;
    if (st.in_flight) GTB_CUDA_OK(ctx, cudaStreamWaitEvent(cs, st.consumed, 0));
    GTB_TRY(st.chrom.reserve(ctx, ni)); GTB_TRY(st.start.reserve(ctx, ni)); GTB_TRY(st.stop.reserve(ctx, ni)); GTB_TRY(st.strand.reserve(ctx, ni));
    const int32_t *d_start = reads->start + r0;
    const void *d_meta = reads->meta + r0;
    if (!(mem & GTB_MEM_DEVICE)) {
      GTB_TRY(st.meta.reserve(ctx, (ni + 3) / 4));
      GTB_CUDA_OK(ctx, cudaMemcpyAsync(st.start.p, reads->start + r0, ni * 4, cudaMemcpyHostToDevice, cs));
      GTB_CUDA_OK(ctx, cudaMemcpyAsync(st.meta.p, reads->meta + r0, ni, cudaMemcpyHostToDevice, cs));
      ctx->h2d_bytes += (int64_t)ni * 5;
      ctx->packed_chunks++;
      d_start = st.start.p; d_meta = st.meta.p;
    } else if (((uintptr_t)d_start & 15) != 0 || ((uintptr_t)d_meta & 3) != 0) {
      return gtb_fail(ctx, GTB_ERR_ARG, "device-resident packed reads must be 16-byte (start) and 4-byte (meta) aligned");
    }
    GTB_CUDA_OK(ctx, cudaEventRecord(st.copied, cs));
    GTB_CUDA_OK(ctx, cudaStreamWaitEvent(ctx->stream, st.copied, 0));
    const unsigned ugrid = gtb_grid_for((int64_t)(ni + 3) / 4, 256, (int64_t)ctx->sm_count * 8);
    int32_t *o_start = st.start.p;
    if (mem & GTB_MEM_DEVICE) {                                         // the engine reads `start` where it lies
      o_start = const_cast<int32_t *>(d_start);
    }
    GTB_LAUNCH(ctx, "unpack", unpack_kernel<1>, ugrid, 256, 0, (int64_t)ni, d_meta, len0, d_start, st.chrom.p, st.stop.p, st.strand.p);
    GTB_TRY(gtb_check_launch(ctx));
    QueryView q;
    q.n_regions = (int64_t)ni; q.chrom = st.chrom.p; q.start = o_start; q.stop = st.stop.p; q.strand = st.strand.p;
    q.weight = nullptr; q.region_offset = nullptr; q.interval_base = 0; q.index_base = ix->queries_seen + r0;
    GTB_TRY(accumulate_device(ix, q, false, (int64_t)ni));
    GTB_CUDA_OK(ctx, cudaEventRecord(st.consumed, ctx->stream));
    st.in_flight = true;
  }
  ix->queries_seen += reads->n;
  if (!(mem & GTB_MEM_DEVICE)) GTB_CUDA_OK(ctx, cudaEventSynchronize(ix->stages[ix->next_stage ^ 1].copied));   // "copied inside the call"
  return GTB_OK;
}

// =================================================================================================
// per-query overlap counts (subset / overlap)
// =================================================================================================
static int build_query_prefix(gtb_index *ix) {
  if (ix->qpre_ready) return GTB_OK;
  gtb_ctx *ctx = ix->ctx;
  // the targets' slots are on the device (d_t_hi / d_t_lo): count them per slot there is no need -- rebuild on the host from
  // the same region arrays the rank structures were built from
  std::vector<uint2> pre((size_t)std::max<int64_t>(ix->n_slots, 1), make_uint2(0u, 0u));
  auto indexable = [&](int64_t k) {
    const int64_t lo = ix->h_off[k], hi = ix->h_off[k + 1];
    if (hi <= lo) return false;
    if (!ix->h_malformed.empty() && ix->h_malformed[(size_t)k]) return false;
    const int64_t s = ix->h_start[lo], e = ix->h_stop[hi - 1];
    if (ix->sorted_rules) return s <= e;
    return !(s > e || e <= 0);
  };
  for (int64_t k = 0; k < ix->n_regions; k++) {
    if (!indexable(k)) continue;
    const int64_t lo = ix->h_off[k], hi = ix->h_off[k + 1];
    const int32_t g = ix->h_chrom[lo] * ix->n_class + ix->h_class_of[(uint8_t)ix->h_strand[lo]];
    const int32_t gb = ix->h_goff[g], ge = ix->h_goff[g + 1];
    const int32_t *b = ix->h_points.data() + gb, *e = ix->h_points.data() + ge - 1;
    const int32_t ts = ix->h_start[lo], te = ix->h_stop[hi - 1];
    pre[(size_t)(std::lower_bound(b, e, ts == INT32_MIN ? INT32_MIN : ts - 1) - ix->h_points.data())].x++;
    pre[(size_t)(std::lower_bound(b, e, te) - ix->h_points.data())].y++;
  }
  for (int32_t g = 0; g < ix->n_groups; g++) {                          // exclusive prefix sums inside every group
    uint32_t ax = 0, ay = 0;
    for (int32_t j = ix->h_goff[g]; j < ix->h_goff[g + 1]; j++) {
      const uint2 v = pre[(size_t)j];
      pre[(size_t)j] = make_uint2(ax, ay);
      ax += v.x; ay += v.y;
    }
  }
  GTB_TRY(upload(ctx, ix->d_qpre, pre));
  GTB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  ix->qpre_ready = true;
  return GTB_OK;
}

extern "C" int gtb_index_query_counts(gtb_index *ix, const gtb_set *queries, unsigned mem, uint32_t *out, unsigned out_mem, int64_t *err_index) {
  if (err_index) *err_index = -1;
  if (!ix || !queries) return GTB_ERR_ARG;
  gtb_ctx *ctx = ix->ctx;
  if (ix->op != GTB_OP_COUNT) return gtb_fail(ctx, GTB_ERR_ARG, "per-query counts need an index created with GTB_OP_COUNT");
  if (queries->n_regions < 0 || queries->n_intervals < 0) return gtb_fail(ctx, GTB_ERR_ARG, "negative sizes");
  if (queries->n_regions == 0) return GTB_OK;
  if (!out || !queries->chrom || !queries->start || !queries->stop || !queries->strand) return gtb_fail(ctx, GTB_ERR_ARG, "null arrays");
  if (!queries->region_offset && queries->n_regions != queries->n_intervals)
    return gtb_fail(ctx, GTB_ERR_ARG, "region_offset is NULL but n_regions != n_intervals");
  GTB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  const bool dev_in = (mem & GTB_MEM_DEVICE) != 0, dev_out = (out_mem & GTB_MEM_DEVICE) != 0;
  bool batch_multi = queries->region_offset != nullptr;
  if (queries->region_offset && !dev_in) {
    batch_multi = false;
    for (int64_t k = 0; k < queries->n_regions; k++) {
      const int64_t d = queries->region_offset[k + 1] - queries->region_offset[k];
      if (d < 1) return gtb_fail(ctx, GTB_ERR_ARG, "region_offset must be increasing: every region has at least one interval");
      batch_multi = batch_multi || d != 1;
    }
  }
  const bool ranks = ix->match_gaps || (!ix->index_multi && !batch_multi);   // spans decide (-gaps), or nothing but single intervals
  if (ranks) GTB_TRY(build_query_prefix(ix)); else GTB_TRY(build_enum_structures(ix));
  GTB_CUDA_OK(ctx, cudaMemsetAsync(ix->d_err.p, 0xFF, sizeof(ull), ctx->stream));
  RankView rv = rank_view(ix);
  const int64_t CHUNK = (int64_t)4 << 20;
  dbuf<uint32_t> d_out;
  for (int64_t r0 = 0; r0 < queries->n_regions; r0 += CHUNK) {
    const int64_t r1 = std::min(queries->n_regions, r0 + CHUNK);
    const int64_t i0 = batch_multi && !dev_in ? queries->region_offset[r0] : r0;
    const int64_t i1 = batch_multi && !dev_in ? queries->region_offset[r1] : r1;
    const size_t nr = (size_t)(r1 - r0), ni = (size_t)(i1 - i0);
    QueryView q;
    q.weight = nullptr; q.index_base = r0; q.interval_base = 0; q.region_offset = nullptr; q.n_regions = (int64_t)nr;
    if (dev_in) {
      // device-resident: offsets (if any) index the arrays as passed; a chunk of regions keeps the arrays' origin
      q.chrom = queries->chrom; q.start = queries->start; q.stop = queries->stop; q.strand = queries->strand;
      if (batch_multi) q.region_offset = queries->region_offset + r0;
      else { q.chrom += r0; q.start += r0; q.stop += r0; q.strand += r0; }
    } else {
      gtb_index::stage &st = ix->stages[0];
      GTB_TRY(st.chrom.reserve(ctx, ni)); GTB_TRY(st.start.reserve(ctx, ni)); GTB_TRY(st.stop.reserve(ctx, ni)); GTB_TRY(st.strand.reserve(ctx, ni));
      GTB_CUDA_OK(ctx, cudaMemcpyAsync(st.chrom.p, queries->chrom + i0, ni * 4, cudaMemcpyHostToDevice, ctx->stream));
      GTB_CUDA_OK(ctx, cudaMemcpyAsync(st.start.p, queries->start + i0, ni * 4, cudaMemcpyHostToDevice, ctx->stream));
      GTB_CUDA_OK(ctx, cudaMemcpyAsync(st.stop.p, queries->stop + i0, ni * 4, cudaMemcpyHostToDevice, ctx->stream));
      GTB_CUDA_OK(ctx, cudaMemcpyAsync(st.strand.p, queries->strand + i0, ni, cudaMemcpyHostToDevice, ctx->stream));
      ctx->h2d_bytes += (int64_t)ni * 13;
      q.chrom = st.chrom.p; q.start = st.start.p; q.stop = st.stop.p; q.strand = st.strand.p;
      if (batch_multi) {
        GTB_TRY(st.off.reserve(ctx, nr + 1));
        GTB_CUDA_OK(ctx, cudaMemcpyAsync(st.off.p, queries->region_offset + r0, (nr + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
        q.region_offset = st.off.p; q.interval_base = i0;
      }
    }
    uint32_t *o = dev_out ? out + r0 : nullptr;
    if (!dev_out) { GTB_TRY(d_out.reserve(ctx, nr)); o = d_out.p; }
    if (ranks) {
      const unsigned grid = gtb_grid_for((int64_t)nr, 256, (int64_t)ctx->sm_count * 8);
      GTB_LAUNCH(ctx, "query_counts", query_counts_kernel, grid, 256, 0, q, rv, (const uint2 *)ix->d_qpre.p, o);
    } else {
      EnumView ev;
      ev.n_entries = ix->n_entries; ev.keys = ix->d_keys.p; ev.rid = ix->d_rid.p;
      ev.r_chrom = ix->d_r_chrom.p; ev.r_start = ix->d_r_start.p; ev.r_stop = ix->d_r_stop.p;
      ev.r_strand = ix->d_r_strand.p; ev.r_off = ix->d_r_off.p; ev.direct = ix->d_direct.p;
      const unsigned grid = gtb_grid_for((int64_t)nr, 128, (int64_t)ctx->sm_count * 16);
      GTB_LAUNCH(ctx, "query_enumerate", query_enumerate_kernel, grid, 128, 0, q, rv, ev, ix->ignore_strand, o);
    }
    GTB_TRY(gtb_check_launch(ctx));
    if (!dev_out) {
      GTB_CUDA_OK(ctx, cudaMemcpyAsync(out + r0, d_out.p, nr * 4, cudaMemcpyDeviceToHost, ctx->stream));
      ctx->d2h_bytes += (int64_t)nr * 4;
    }
    GTB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));               // the staging buffers are reused by the next chunk
  }
  d_out.release();
  return gtb_index_status(ix, err_index);
}

// The bin index of UnsortedGenomicRegionSetOverlaps (:5600-5675) as a sorted entry list: levels of `bits` shift-bits (the -B list
// plus the closing level of 60 bits, :5620-5637), an admitted region in the first level whose bin holds both its ends, the
// chain of a bin from the region inserted last to the first.
static int build_match_structures(gtb_index *ix, const int *bits, int n_levels) {
  gtb_ctx *ctx = ix->ctx;
  GTB_TRY(build_enum_structures(ix));                                   // (the interval arrays of the index regions)
  if (ix->match_ready && ix->match_bits == std::vector<int>(bits, bits + n_levels)) return GTB_OK;
  std::vector<std::pair<ull, int32_t>> ent;
  ent.reserve((size_t)ix->n_regions);
  for (int64_t k = 0; k < ix->n_regions; k++) {
    const int64_t lo = ix->h_off[k], hi = ix->h_off[k + 1];
    if (hi <= lo) continue;
    if (!ix->h_malformed.empty() && ix->h_malformed[(size_t)k]) continue;
    int64_t s = ix->h_start[lo], e = ix->h_stop[hi - 1];
    if (s > e || e <= 0) continue;                                      // :5659
    if (s <= 0) s = 1;                                                  // :5660
    for (int l = 0; l < n_levels; l++)
      if ((s >> bits[l]) == (e >> bits[l])) {
        ent.push_back({((ull)(uint32_t)ix->h_chrom[lo] << 35) | ((ull)l << 32) | (ull)(uint32_t)(s >> bits[l]), (int32_t)k});
        break;
      }
  }
  std::sort(ent.begin(), ent.end(), [](const std::pair<ull, int32_t> &a, const std::pair<ull, int32_t> &b) {
    return a.first != b.first ? a.first < b.first : a.second > b.second;
  });
  std::vector<ull> keys(ent.size());
  std::vector<int32_t> rid(ent.size());
  for (size_t i = 0; i < ent.size(); i++) { keys[i] = ent[i].first; rid[i] = ent[i].second; }
  ix->n_match_entries = (int64_t)ent.size();
  GTB_TRY(upload(ctx, ix->d_mkeys, keys));
  GTB_TRY(upload(ctx, ix->d_mrid, rid));
  GTB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  ix->match_bits.assign(bits, bits + n_levels);
  ix->match_ready = true;
  return GTB_OK;
}

extern "C" int gtb_index_query_matches(gtb_index *ix, const gtb_set *queries, unsigned mem, const int *bin_bits, int n_bin_bits,
                                       const int64_t *match_offset, int32_t *matches, int64_t *err_index) {
  if (err_index) *err_index = -1;
  if (!ix || !queries) return GTB_ERR_ARG;
  gtb_ctx *ctx = ix->ctx;
  if (ix->op != GTB_OP_COUNT) return gtb_fail(ctx, GTB_ERR_ARG, "per-query matches need an index created with GTB_OP_COUNT");
  if (queries->n_regions < 0 || queries->n_intervals < 0) return gtb_fail(ctx, GTB_ERR_ARG, "negative sizes");
  if (queries->n_regions == 0) return GTB_OK;
  if (mem & GTB_MEM_DEVICE) return gtb_fail(ctx, GTB_ERR_UNSUPPORTED, "gtb_index_query_matches takes host-resident queries");
  if (!match_offset || !queries->chrom || !queries->start || !queries->stop || !queries->strand) return gtb_fail(ctx, GTB_ERR_ARG, "null arrays");
  if (!queries->region_offset && queries->n_regions != queries->n_intervals)
    return gtb_fail(ctx, GTB_ERR_ARG, "region_offset is NULL but n_regions != n_intervals");
  if (match_offset[queries->n_regions] > match_offset[0] && !matches) return gtb_fail(ctx, GTB_ERR_ARG, "null match array");
  static const int default_bits[4] = {17, 20, 23, 26};
  if (!bin_bits) { bin_bits = default_bits; n_bin_bits = 4; }
  MatchLevels lv;
  if (n_bin_bits < 0 || n_bin_bits > 7) return gtb_fail(ctx, GTB_ERR_UNSUPPORTED, "at most 7 bin levels (plus the closing one)");
  for (int l = 0; l < n_bin_bits; l++) {
    if (bin_bits[l] < 0 || bin_bits[l] > 62) return gtb_fail(ctx, GTB_ERR_ARG, "bin bits must lie in 0..62");
    lv.bits[l] = bin_bits[l];
  }
  lv.bits[n_bin_bits] = 60; lv.n = n_bin_bits + 1;                      // :5637
  GTB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  bool batch_multi = false;
  for (int64_t k = 0; k < queries->n_regions; k++) {
    if (queries->region_offset) {
      const int64_t d = queries->region_offset[k + 1] - queries->region_offset[k];
      if (d < 1) return gtb_fail(ctx, GTB_ERR_ARG, "region_offset must be increasing: every region has at least one interval");
      batch_multi = batch_multi || d != 1;
    }
    if (match_offset[k + 1] < match_offset[k]) return gtb_fail(ctx, GTB_ERR_ARG, "match_offset must not decrease");
  }
  if (ix->sorted_rules)                                                 // the Sorted class skips no index region; the bin lists hold the Unsorted class's
    for (int64_t k = 0; k < ix->n_regions; k++) {
      const int64_t lo = ix->h_off[k], hi = ix->h_off[k + 1];
      if (hi > lo && (ix->h_start[lo] > ix->h_stop[hi - 1] || ix->h_stop[hi - 1] <= 0))
        return gtb_fail(ctx, GTB_ERR_UNSUPPORTED, "index regions with start > stop or stop <= 0 under GTB_SORTED_RULES: no match lists");
    }
  GTB_TRY(build_match_structures(ix, lv.bits, lv.n));
  GTB_CUDA_OK(ctx, cudaMemsetAsync(ix->d_err.p, 0xFF, sizeof(ull), ctx->stream));
  RankView rv = rank_view(ix);
  EnumView ev = enum_view(ix);
  ev.n_entries = ix->n_match_entries; ev.keys = ix->d_mkeys.p; ev.rid = ix->d_mrid.p;
  dbuf<int32_t> d_match;
  dbuf<int64_t> d_moff;
  dbuf<uint32_t> d_flag;
  GTB_TRY(d_flag.reserve(ctx, 1));
  GTB_CUDA_OK(ctx, cudaMemsetAsync(d_flag.p, 0, sizeof(uint32_t), ctx->stream));
  const int64_t CHUNK = (int64_t)1 << 20;
  for (int64_t r0 = 0; r0 < queries->n_regions; r0 += CHUNK) {
    const int64_t r1 = std::min(queries->n_regions, r0 + CHUNK);
    const int64_t i0 = batch_multi ? queries->region_offset[r0] : r0, i1 = batch_multi ? queries->region_offset[r1] : r1;
    const size_t nr = (size_t)(r1 - r0), ni = (size_t)(i1 - i0), nm = (size_t)(match_offset[r1] - match_offset[r0]);
    gtb_index::stage &st = ix->stages[0];
    GTB_TRY(st.chrom.reserve(ctx, ni)); GTB_TRY(st.start.reserve(ctx, ni)); GTB_TRY(st.stop.reserve(ctx, ni)); GTB_TRY(st.strand.reserve(ctx, ni));
    GTB_CUDA_OK(ctx, cudaMemcpyAsync(st.chrom.p, queries->chrom + i0, ni * 4, cudaMemcpyHostToDevice, ctx->stream));
    GTB_CUDA_OK(ctx, cudaMemcpyAsync(st.start.p, queries->start + i0, ni * 4, cudaMemcpyHostToDevice, ctx->stream));
    GTB_CUDA_OK(ctx, cudaMemcpyAsync(st.stop.p, queries->stop + i0, ni * 4, cudaMemcpyHostToDevice, ctx->stream));
    GTB_CUDA_OK(ctx, cudaMemcpyAsync(st.strand.p, queries->strand + i0, ni, cudaMemcpyHostToDevice, ctx->stream));
    ctx->h2d_bytes += (int64_t)ni * 13 + (int64_t)(nr + 1) * 8;
    QueryView q;
    q.weight = nullptr; q.index_base = r0; q.interval_base = 0; q.region_offset = nullptr; q.n_regions = (int64_t)nr;
    q.chrom = st.chrom.p; q.start = st.start.p; q.stop = st.stop.p; q.strand = st.strand.p;
    if (batch_multi) {
      GTB_TRY(st.off.reserve(ctx, nr + 1));
      GTB_CUDA_OK(ctx, cudaMemcpyAsync(st.off.p, queries->region_offset + r0, (nr + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
      q.region_offset = st.off.p; q.interval_base = i0;
    }
    GTB_TRY(d_moff.reserve(ctx, nr + 1));
    GTB_TRY(d_match.reserve(ctx, nm ? nm : 1));
    GTB_CUDA_OK(ctx, cudaMemcpyAsync(d_moff.p, match_offset + r0, (nr + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    const unsigned grid = gtb_grid_for((int64_t)nr, 128, (int64_t)ctx->sm_count * 16);
    GTB_LAUNCH(ctx, "query_matches", query_matches_kernel, grid, 128, 0, q, rv, ev, lv, ix->match_gaps, ix->ignore_strand, ix->sorted_rules,
               (const int64_t *)d_moff.p, match_offset[r0], d_match.p, d_flag.p);
    GTB_TRY(gtb_check_launch(ctx));
    if (nm) {
      GTB_CUDA_OK(ctx, cudaMemcpyAsync(matches + (match_offset[r0] - match_offset[0]), d_match.p, nm * 4, cudaMemcpyDeviceToHost, ctx->stream));
      ctx->d2h_bytes += (int64_t)nm * 4;
    }
    GTB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));               // the staging buffers are reused by the next chunk
  }
  uint32_t flag = 0;
  GTB_CUDA_OK(ctx, cudaMemcpyAsync(&flag, d_flag.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  GTB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  d_match.release(); d_moff.release(); d_flag.release();
  const int rc = gtb_index_status(ix, err_index);
  if (rc != GTB_OK) return rc;
  if (flag) return gtb_fail(ctx, GTB_ERR_ARG, "match_offset does not agree with the counts of gtb_index_query_counts for these queries");
  return GTB_OK;
}

// =================================================================================================
// finish
// =================================================================================================
// the device work of a finish: scans, finalisation, the copy of the values to `out` -- all enqueued, nothing waited for
static int finish_enqueue(gtb_index *ix, uint64_t *out, unsigned mem) {
  if (!ix || (!out && ix->n_regions > 0)) return GTB_ERR_ARG;
  gtb_ctx *ctx = ix->ctx;
  GTB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  const int64_t K = std::max<int64_t>(ix->n_slots, 1);
  // scan into a second buffer so that more batches may still be added after a finish
  if (ix->n_slots > 0)
    GTB_TRY(gtb_inclusive_scan_planes_u64(ctx, ix->d_hist.p, ix->d_hist_scan.p, ix->n_slots, ix->planes, K, ix->d_scan_scratch));
  if (ix->n_regions > 0) {
    const unsigned grid = (unsigned)((ix->n_regions + 255) / 256);
    if (ix->op == GTB_OP_COVERAGE)
      GTB_LAUNCH(ctx, "finalize_coverage", finalize_kernel<true>, grid, 256, 0, ix->n_regions, ix->d_t_off.p, ix->d_t_hi.p,
                 ix->d_t_lo.p, ix->d_t_base.p, ix->d_goff.p, ix->d_points.p, ix->d_hist_scan.p, K, ix->d_direct.p, ix->d_out.p);
    else
      GTB_LAUNCH(ctx, "finalize_count", finalize_kernel<false>, grid, 256, 0, ix->n_regions, ix->d_t_off.p, ix->d_t_hi.p,
                 ix->d_t_lo.p, ix->d_t_base.p, ix->d_goff.p, ix->d_points.p, ix->d_hist_scan.p, K, ix->d_direct.p, ix->d_out.p);
    GTB_TRY(gtb_check_launch(ctx));
    GTB_CUDA_OK(ctx, cudaMemcpyAsync(out, ix->d_out.p, sizeof(ull) * (size_t)ix->n_regions,
                                     (mem & GTB_MEM_DEVICE) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, ctx->stream));
    if (!(mem & GTB_MEM_DEVICE)) ctx->d2h_bytes += (int64_t)sizeof(ull) * ix->n_regions;
  }
  return GTB_OK;
}

extern "C" int gtb_index_finish_async(gtb_index *ix, uint64_t *out, unsigned mem) {
  if (!(mem & GTB_MEM_DEVICE)) return ix ? gtb_fail(ix->ctx, GTB_ERR_ARG, "gtb_index_finish_async writes to device memory") : GTB_ERR_ARG;
  return finish_enqueue(ix, out, mem);
}

extern "C" int gtb_index_status(gtb_index *ix, int64_t *err_index) {
  if (!ix) return GTB_ERR_ARG;
  gtb_ctx *ctx = ix->ctx;
  if (err_index) *err_index = -1;
  GTB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  ull err = ~0ull;
  GTB_CUDA_OK(ctx, cudaMemcpyAsync(&err, ix->d_err.p, sizeof(ull), cudaMemcpyDeviceToHost, ctx->stream));
  GTB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  if (err != ~0ull) {
    if (err_index) *err_index = (int64_t)(err >> 8);
    const int code = (int)(err & 0xFF);
    return gtb_fail(ctx, code, code == GTB_ERR_QUERY_STOP_NONPOSITIVE ? "stop position must be positive!"
                               : code == GTB_ERR_QUERY_START_GT_STOP ? "start position cannot be greater than stop position!"
                               : "query regions should be compatible, sorted and non-overlapping!");
  }
  return GTB_OK;
}

extern "C" int gtb_index_finish(gtb_index *ix, uint64_t *out, unsigned mem, int64_t *err_index) {
  if (err_index) *err_index = -1;
  GTB_TRY(finish_enqueue(ix, out, mem));
  return gtb_index_status(ix, err_index);
}

static int one_shot(gtb_ctx *ctx, int op, const gtb_set *queries, unsigned queries_mem, const gtb_set *regions,
                    unsigned flags, uint64_t *out, int64_t *err_index) {
  gtb_index *ix = nullptr;
  int rc = gtb_index_create(ctx, regions, op, flags, &ix, err_index);
  if (rc != GTB_OK) return rc;
  rc = gtb_index_add_queries(ix, queries, queries_mem);
  if (rc == GTB_OK) rc = gtb_index_finish(ix, out, GTB_MEM_HOST, err_index);
  gtb_index_destroy(ix);
  return rc;
}

extern "C" int gtb_overlap_count(gtb_ctx *ctx, const gtb_set *queries, unsigned queries_mem, const gtb_set *regions,
                                 unsigned flags, uint64_t *out, int64_t *err_index) {
  return one_shot(ctx, GTB_OP_COUNT, queries, queries_mem, regions, flags, out, err_index);
}

extern "C" int gtb_overlap_coverage(gtb_ctx *ctx, const gtb_set *queries, unsigned queries_mem, const gtb_set *regions,
                                    unsigned flags, uint64_t *out, int64_t *err_index) {
  return one_shot(ctx, GTB_OP_COVERAGE, queries, queries_mem, regions, flags, out, err_index);
}
