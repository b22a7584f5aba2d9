// gtb_rank_device.cuh -- device code shared by the RANK engine (gtb_overlap.cu) and the CELL engine
// (gtb_cell.cu, which falls back to the general rank step for the few queries it cannot place).
#pragma once
#include "gtb_overlap.cuh"

__device__ __forceinline__ void report_error(ull *err, int64_t index, int code) {
  atomicMin(err, ((ull)index << 8) | (ull)code);
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
  const int jE = lower_bound_i32(ix.points, jS, ge - 1, qe);
  ull *h = ix.hist;
  const int64_t K = ix.n_slots;
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
