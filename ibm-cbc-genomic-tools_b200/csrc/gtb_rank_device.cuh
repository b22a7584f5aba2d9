// gtb_rank_device.cuh -- device code shared by the RANK engine (gtb_overlap.cu) and the BUCKET / DIRECT engines
// (which fall back to the general rank step for the few queries their fast paths cannot place).
#pragma once
#include "gtb_overlap.cuh"

__device__ __forceinline__ void report_error(ull *err, int64_t index, int code) {
  atomicMin(err, ((ull)index << 8) | (ull)code);
}

// The fatal conditions of a query on a chromosome the index knows: UnsortedGenomicRegionSetOverlaps::GetQuery/NextQuery,
// genomic_intervals.cpp:5740-5741.  SortedGenomicRegionSetOverlaps (:5807-5937, what -S selects) has no such checks and matches
// with the raw predicate (CalcDirection, :1225-1237): under its rules zero-length intervals (start == stop + 1, what BED lines
// "chr 5 5" become) and stops <= 0 pass and are counted by that predicate, which the rank step evaluates for them exactly.
// Intervals inverted by more than that stay an error here (documented divergence: the reference's -S engine goes on).
__device__ __forceinline__ bool admit_interval(const RankView &rv, int32_t qs, int32_t qe, int64_t index) {
  if (rv.admission == 2) return qs <= qe;            // a block of a region the prepass has admitted (CalcOverlap clamps such blocks to 0, :427-432)
  if (rv.admission == 1) {
    if ((int64_t)qs > (int64_t)qe + 1) { report_error(rv.err, index, GTB_ERR_QUERY_START_GT_STOP); return false; }
    return true;
  }
  if (qe <= 0) { report_error(rv.err, index, GTB_ERR_QUERY_STOP_NONPOSITIVE); return false; }
  if (qs > qe) { report_error(rv.err, index, GTB_ERR_QUERY_START_GT_STOP); return false; }
  return true;
}

// first slot j in [lo,hi) with points[j] >= x
__device__ __forceinline__ int lower_bound_i32(const int32_t *__restrict__ p, int lo, int hi, int32_t x) {
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (__ldg(p + mid) < x) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// One query interval [qs,qe] of weight w against the evaluation points of its group [gb,ge):
// binary search in global memory, 64-bit reductions on the slot histograms.
template <bool COVERAGE>
__device__ __forceinline__ void rank_item(const RankView &ix, int gb, int ge, int32_t qs, int32_t qe, int64_t w) {
  const int jS = lower_bound_i32(ix.points, gb, ge - 1, qs);      // ge-1 is the +inf sentinel: result <= ge-1
  const int jE = lower_bound_i32(ix.points, qe < qs ? gb : jS, ge - 1, qe);   // (a zero-length interval under the Sorted class's rules: stop == start - 1)
  ull *h = ix.hist;
  const int64_t K = ix.n_slots;
  GTB_ASSERT(gb >= 0 && jS >= gb && jS < ge && jE >= gb && jE < ge && ge <= K);
  if (jS == jE) {
    if (COVERAGE) atomicAdd(h + H_BOTH * K + jS, (ull)(w * ((int64_t)qe - qs + 1)));
    else atomicAdd(h + H_BOTH * K + jS, (ull)w);
  } else {
    atomicAdd(h + H_SCNT * K + jS, (ull)w);
    atomicAdd(h + H_ECNT * K + jE, (ull)w);
    if (COVERAGE) {
      atomicAdd(h + H_SSUM * K + jS, (ull)(w * (int64_t)qs));
      atomicAdd(h + H_ESUM * K + jE, (ull)(w * (int64_t)qe));
    }
  }
}
