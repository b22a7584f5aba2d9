// gtb_direct.cu -- DIRECT engine: overlap count / coverage in ONE pass over the queries, for indices whose evaluation points
// fit a byte counter each in one SM's shared memory (up to ~190 k slots: ~95 k regions).
//
// Why (measured on B200, profiles/microbench/gather_rate_b200.txt and red_rate_b200.txt): a random 8-byte load from a
// table that lives in L2 (up to 16 MB) costs 1.05 SM-cycles per element, a random shared-memory atomic 0.17, a random
// global reduction 1.45.  The BUCKET engine pays 13 + 4 + 4 bytes of HBM traffic and two kernels' worth of
// shared-memory work per query to avoid global reductions; this engine pays 13 bytes, one L2 gather and one shared atomic:
//
//   * Every (chromosome, strand) owns a block of `stride` cells of 2^c bp (c chosen so that the cells queries can reach stay
//     within 13 MB: hg19 x 2 strands -> c = 12), so the cell of a coordinate is arithmetic.  One 8-byte entry per cell: the
//     slot of the first evaluation point at or after the cell's start (24 bits), the offsets of up to two points inside the
//     cell (16 bits each, 0xFFFF = none) and three flags: nothing to count here (no such group / beyond its last point), a
//     dense cell (more than two points: walk them from the cell's first slot, four loads per round trip), general path
//     (a group whose points are all <= 0).  For a query [s, e]
//         jS = slot0(cell(s)) + #{p < s & (2^c - 1)},   jE likewise for e  (a second gather only if e is in another cell),
//     which is lower_bound(points, s) / lower_bound(points, e) of the RANK engine (gtb_rank_device.cuh) without a search.
//   * jS == jE (almost every short read): one shared-memory atomicAdd on a BYTE counter (four slots per 32-bit word).  The
//     add that finds its byte at 127 moves 128 to the global plane (one atomicSub, one reduction), so a byte stays below 256
//     unless more than 127 further adds land before that correction does; the add that finds 255 raises a flag.  The word is
//     arithmetically exact throughout (transient carries between bytes cancel), so without the flag the bytes are the
//     exact counts.  With the flag the whole batch is discarded and replayed through the general rank step by the commit kernel
//     (which otherwise only sums) -- correct whatever the skew, and the host switches the engine off for this index.
//   * jS != jE (a read straddling an evaluation point, ~0.1 %): reductions into per-batch delta planes.  Strands other than
//     +/- and invalid intervals: the general path inline (admission rules, binary search), into the same delta planes, so that
//     a discarded batch leaves no trace.
//   * Position-sorted input: a warp whose 128 queries share one slot sends one reduction of 128.  Clustered input (neighbouring
//     queries sharing slots): slots shared by eight or more lanes of a warp leave as one reduction each.
//   * Coverage: the counters count the queries that have the batch's common length (the first query's) and the commit
//     multiplies; a query of another length costs a reduction (and is counted: many of them send the index to BUCKET).
//   * At the end each CTA writes its counters to its row of a [CTAs x slots] byte matrix (coalesced); the commit kernel sums the
//     rows and the delta planes into the index's histogram planes.  Finalisation is the RANK engine's.
//
// Reference semantics reproduced: admission rules of UnsortedGenomicRegionSetOverlaps::GetQuery/NextQuery
// (genomic_intervals.cpp:5719-5745), the overlap predicate (:624-630, :5227-5229) and CalcOverlap (:427-432) through the rank
// formulation of SURVEY.md section 7.1 -- see gtb_overlap.cuh.
#include "gtb_rank_device.cuh"
#include <algorithm>

namespace {

constexpr int DR_THREADS = 1024;
constexpr int DR_ITEMS = 4;                                  // (the clustered-input test in the kernel spells the four items out)
constexpr int DR_TILE = DR_THREADS * DR_ITEMS;
constexpr uint32_t DR_NOPOINT = 0xFFFFu;
constexpr size_t DR_TABLE_BYTES_TARGET = (size_t)13 << 20;     // cells that queries can touch: inside L2's fast range (<= 16 MB measured)

struct DirectView {
  int cbits;                        // cell = 2^cbits bp (<= 14)
  int32_t n_chrom;
  uint32_t nsig;                    // 2: '+' and '-' queries have cell blocks of their own, 1: they share one (-i)
  uint32_t stride;                  // cells per block; the last cell of every block says "nothing to count"
  const uint2 *cells;               // [(n_chrom + 1) * nsig * stride]  x = slot0 | general << 24 | nothing << 25 | scan << 26,  y = p0 | p1 << 16 (0xFFFF: none)
  uint32_t n_words;                 // 32-bit words of byte counters per CTA (n_slots / 4, rounded up to a multiple of 4)
  uint32_t *cta_counts;             // [grid][n_words]
  ull *delta;                       // [3][n_slots] this batch's contributions that did not go through the byte counters
  uint32_t *flag;                   // [0] generation of the last batch in which a byte counter overflowed (that batch is discarded and replayed)
                                    // [1] coverage: queries whose length was not the batch's common one (a global reduction each)
                                    // [2] ... those of them 256 bp or longer (which the BUCKET engine's elements cannot describe)
  uint32_t gen;                     // this batch's generation (1, 2, ...)
  uint32_t n_cells;                 // entries of `cells` (bounds checks)
  int pair_check;                   // queries 2p and 2p + 1 are the two intervals of region p.  1 (COVERAGE): counted as they lie, the pair checked as
                                    // a region.  2 (-gaps): the pair's span [first start, second stop] is the query
                                    // 3 (count without -gaps): as 2 for a pair whose span holds no evaluation point -- every region that overlaps the
                                    // span then overlaps both mates -- and the pair itself onto the exception list otherwise
                                    // 4 (count without -gaps, regions of any shape): the queries ARE spans already (written by region_prepass_kernel,
                                    // one per region); a multi-interval region whose span holds an evaluation point goes onto the exception list by its
                                    // number (in ex_chrom), everything else is counted as the span it is
  int region_admission;             // ... under these admission rules (0: the default engine's, 1: the Sorted class's)
  const int64_t *q_off;             // pair_check == 4: the batch's region offsets (region r has q_off[r + 1] - q_off[r] intervals)
  uint32_t *ex_count;               // pair_check == 3: pairs the candidate-enumeration engine has to look at (two entries each in the arrays)
  int32_t *ex_chrom, *ex_start, *ex_stop;
  int8_t *ex_strand;
};
constexpr uint32_t DR_GENERAL = 1u << 24, DR_NOTHING = 1u << 25, DR_SCAN = 1u << 26;

__device__ __forceinline__ void dr_red64(ull *p, ull v) { asm volatile("red.global.add.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ uint4 dr_ldg128(const void *p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint32_t dr_ldg32(const void *p) {
  uint32_t r;
  asm volatile("ld.global.nc.L1::no_allocate.b32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}
// one cell entry; allocating in L1 (measured: 0.585 ms per 100 M queries against 0.682 ms with L1::no_allocate)
__device__ __forceinline__ uint2 dr_gather(const uint2 *p) {
  uint2 r;
  asm volatile("ld.global.nc.v2.b32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}

// The general path for one query: the reference's admission rules, then the rank step (gtb_rank_device.cuh).
// Same decisions as admit_query + rank_accumulate_kernel of gtb_overlap.cu for a single-interval, unweighted query.
template <bool COVERAGE>
__device__ __noinline__ void dr_general(const RankView &rv, int32_t c, int32_t qs, int32_t qe, int sbyte, int64_t index, int64_t w = 1) {
  if ((uint32_t)c >= (uint32_t)rv.n_chrom || !rv.chrom_present[c]) return;           // :5719-5720
  if (!admit_interval(rv, qs, qe, index)) return;                                      // :5740-5741
  const int cls = rv.class_of[(uint8_t)sbyte];
  if (cls < 0) return;                                                              // no index region carries this strand, :5229
  const int g = c * rv.n_class + cls;
  const int gb = rv.goff[g], ge = rv.goff[g + 1];
  if (ge > gb) rank_item<COVERAGE>(rv, gb, ge, qs, qe, w);
}

// COVERAGE: the batch's common read length - 1, as every thread of both kernels computes it: the majority of three samples
// (first, middle, last query), so that one atypical read at the head of a batch does not make every other read the odd one
__device__ __forceinline__ uint32_t dr_common_len(const QueryView &q) {
  if (q.n_regions <= 0) return 0u;
  const int64_t m = q.n_regions / 2, z = q.n_regions - 1;
  const uint32_t a = (uint32_t)(q.stop[0] - q.start[0]), b = (uint32_t)(q.stop[m] - q.start[m]), c = (uint32_t)(q.stop[z] - q.start[z]);
  return b == c ? b : a;
}

// number of the cell's (at most two) points that lie below offset `off` (an absent point is 0xFFFF and never does)
__device__ __forceinline__ uint32_t dr_below(uint32_t y, uint32_t off) {
  return ((y & 0xFFFFu) < off ? 1u : 0u) + ((y >> 16) < off ? 1u : 0u);
}

// first slot at or after j whose point is >= x: the way through a cell with more than two points (four loads per round trip; a
// group ends with a +inf sentinel, and the array is followed by slack, so the look-ahead is harmless)
__device__ __forceinline__ uint32_t dr_scan(const int32_t *__restrict__ pts, uint32_t j, int32_t x) {
  for (;;) {
    const int32_t a = __ldg(pts + j), b = __ldg(pts + j + 1), c = __ldg(pts + j + 2), d = __ldg(pts + j + 3);
    if (!(a < x)) return j;
    if (!(b < x)) return j + 1;
    if (!(c < x)) return j + 2;
    if (!(d < x)) return j + 3;
    j += 4;
  }
}

// What the reference decides per REGION for a read pair whose two intervals the engine takes as two queries (see
// region_prepass_kernel in gtb_overlap.cu, which does the same for regions of any shape in a pass of its own): well-formedness
// (same chromosome and strand, sorted, non-overlapping: :5698, :5709) and the fatal conditions of the span on a chromosome the
// index knows (:5740-5741).  Nothing is written unless something is wrong.
__device__ __forceinline__ void dr_check_pair(const RankView &rv, int admission, uint32_t c0, uint32_t c1, uint32_t sb0, uint32_t sb1,
                                              int32_t s0, int32_t e0, int32_t s1, int32_t e1, int64_t region) {
  if (!(c0 == c1 && sb0 == sb1 && s1 >= s0 && s1 > e0)) { report_error(rv.err, region, GTB_ERR_QUERY_REGION); return; }
  const bool fatal = admission == 1 ? (int64_t)s0 > (int64_t)e1 + 1 : (e1 <= 0 || s0 > e1);
  if (fatal && c0 < (uint32_t)rv.n_chrom && rv.chrom_present[c0])
    report_error(rv.err, region, admission == 1 || e1 > 0 ? GTB_ERR_QUERY_START_GT_STOP : GTB_ERR_QUERY_STOP_NONPOSITIVE);
}

// pair_check == 3: a read pair for the exception list
__device__ __forceinline__ void dr_append_pair(const DirectView &dv, int32_t c, int32_t s1, int32_t e1, int32_t s2, int32_t e2, int sbyte) {
  const uint32_t k = atomicAdd(dv.ex_count, 1u);                        // (room for every pair of the batch: cannot overflow)
  dv.ex_chrom[2 * k] = c; dv.ex_chrom[2 * k + 1] = c;
  dv.ex_start[2 * k] = s1; dv.ex_stop[2 * k] = e1; dv.ex_start[2 * k + 1] = s2; dv.ex_stop[2 * k + 1] = e2;
  dv.ex_strand[2 * k] = (int8_t)sbyte; dv.ex_strand[2 * k + 1] = (int8_t)sbyte;
}
// ... after the checks the general path makes for the pair's span (dr_general above, without its rank step)
__device__ __noinline__ void dr_general_pair(const RankView &rv, const DirectView &dv, int32_t c, int32_t s1, int32_t e1, int32_t s2, int32_t e2, int sbyte, int64_t index) {
  if ((uint32_t)c >= (uint32_t)rv.n_chrom || !rv.chrom_present[c]) return;
  if (!admit_interval(rv, s1, e2, index)) return;
  if (rv.class_of[(uint8_t)sbyte] < 0) return;
  dr_append_pair(dv, c, s1, e1, s2, e2, sbyte);
}

// pair_check == 4: the general path for the span of region r -- a multi-interval region is left to the exact engine once it
// has passed the checks, a single-interval one is its span
template <bool COVERAGE>
__device__ __noinline__ void dr_general_span(const RankView &rv, const DirectView &dv, int32_t c, int32_t qs, int32_t qe, int sbyte, int64_t index, int64_t r) {
  if (dv.q_off[r + 1] - dv.q_off[r] <= 1) { dr_general<COVERAGE>(rv, c, qs, qe, sbyte, index); return; }
  if ((uint32_t)c >= (uint32_t)rv.n_chrom || !rv.chrom_present[c]) return;
  if (!admit_interval(rv, qs, qe, index)) return;
  if (rv.class_of[(uint8_t)sbyte] < 0) return;
  dv.ex_chrom[atomicAdd(dv.ex_count, 1u)] = (int32_t)r;
}

// The 13 bytes per query come as 128-bit loads straight into registers, one tile ahead of the tile being counted, so that
// 4 096 gathers per SM are in flight.  Staging the input through a TMA ring in shared memory instead was measured and lost:
// next to 120 KB of counters the ring leaves the L1 too small to track the gathers' misses (1.85 SM-cycles per gather with
// 224 KB carved out), and tiles small enough to fit leave too few gathers in flight (1.2 - 1.6 ms against 0.59 ms per 100 M
// queries; profiles/r1_experiments.md).
//
// The hot loop holds no table but the cells: every (chromosome, strand) has a block of `stride` cells, so the cell of a
// coordinate is pure arithmetic, and what used to be per-group knowledge (no such group, beyond the last point, a dense
// cell, a group left to the general path) is three flag bits of the cell entry.
//
// COVERAGE: the "both" plane takes the query's length instead of 1.  Reads of a sequencing run mostly share one length, so the
// byte counters count the queries whose length equals the batch's first query's and the commit kernel multiplies; a query of
// another length costs a global reduction (and is counted, so that the host can leave the engine if they are many).
//
// WEIGHTED (--max-label-value: a query counts min(max, atol(label)) times, genomic_intervals.cpp:1081-1085): a weight of 1..127
// is what the query adds to its byte counter -- the add that takes a byte from below 128 to 128 or more moves 128 out, the add
// that would take it past 255 raises the flag -- any other weight (0, negative, larger) leaves as a reduction of its own.
template <bool COVERAGE, bool WEIGHTED>
__global__ void __launch_bounds__(DR_THREADS, 1) direct_count_kernel(const __grid_constant__ QueryView q, const __grid_constant__ RankView rv,
                                                                       const __grid_constant__ DirectView dv) {
  extern __shared__ __align__(16) uint32_t s_cnt[];                   // [n_words] four byte counters per word, then one dummy word per lane
  for (uint32_t i = threadIdx.x; i < dv.n_words + 32u; i += DR_THREADS) s_cnt[i] = 0;
  __syncthreads();
  const uint32_t cbits = (uint32_t)dv.cbits, cmask = (1u << cbits) - 1u;
  const uint32_t stride = dv.stride, last = stride - 1u, sigmask = dv.nsig - 1u, n_chrom = (uint32_t)dv.n_chrom;
  const int64_t n = q.n_regions;
  const int64_t n_full = n / DR_TILE;
  const int lane = threadIdx.x & 31;
  const uint32_t len0m1 = COVERAGE ? dr_common_len(q) : 0u;                            // the common length - 1
  const ull unit = COVERAGE ? (ull)(int64_t)(int32_t)len0m1 + 1ull : 1ull;                // what one counted query is worth in the "both" plane
  uint32_t sorted_wait = 0;                                            // tiles until this warp tries the sorted-input test again
  uint32_t odd = 0, odd_long = 0;                                      // COVERAGE: queries of another length; those of 256 bp and more
  bool overflowed = false;

  uint4 nc, ns, ne, nw = make_uint4(1u, 1u, 1u, 1u);
  uint32_t nst;
  int64_t tile = blockIdx.x;
  if (tile < n_full) {
    const int64_t first = tile * DR_TILE + (int64_t)threadIdx.x * DR_ITEMS;
    nc = dr_ldg128(q.chrom + first); ns = dr_ldg128(q.start + first); ne = dr_ldg128(q.stop + first); nst = dr_ldg32(q.strand + first);
    if (WEIGHTED) nw = dr_ldg128(q.weight + first);
  }
  for (; tile < n_full; tile += gridDim.x) {
    const uint4 cc = nc, cs = ns, ce = ne, cw = nw;
    const uint32_t stw = nst;
    const int64_t next = tile + gridDim.x;
    if (next < n_full) {                                               // the next tile's 13 bytes per query are on their way while this one is counted
      const int64_t first = next * DR_TILE + (int64_t)threadIdx.x * DR_ITEMS;
      nc = dr_ldg128(q.chrom + first); ns = dr_ldg128(q.start + first); ne = dr_ldg128(q.stop + first); nst = dr_ldg32(q.strand + first);
      if (WEIGHTED) nw = dr_ldg128(q.weight + first);
    }
    const int32_t wt[DR_ITEMS] = {(int)cw.x, (int)cw.y, (int)cw.z, (int)cw.w};          // (all 1 unless WEIGHTED)
    uint32_t c[DR_ITEMS] = {cc.x, cc.y, cc.z, cc.w};
    int32_t s[DR_ITEMS] = {(int)cs.x, (int)cs.y, (int)cs.z, (int)cs.w};
    int32_t e[DR_ITEMS] = {(int)ce.x, (int)ce.y, (int)ce.z, (int)ce.w};
    // strands: '+' = 0x2B, '-' = 0x2D.  xw has 0x00 / 0x06 in the bytes of '+' / '-' queries
    const uint32_t xw = stw ^ 0x2B2B2B2Bu;
    uint32_t general = 0;                                              // bit i: item i takes the general path
    if (xw & 0xF9F9F9F9u) {                                            // some strand other than '+'/'-'
#pragma unroll
      for (int i = 0; i < DR_ITEMS; i++) general |= ((xw >> (8 * i)) & 0xF9u) ? (1u << i) : 0u;
    }
#pragma unroll
    for (int i = 0; i < DR_ITEMS; i++)
      if (!(s[i] >= 1 && s[i] <= e[i])) general |= 1u << i;            // the reference's fatal cases (or nothing, on an unknown chromosome)
    if (dv.pair_check == 1) {
      if (COVERAGE) {
        const int64_t region = q.index_base + (tile * DR_TILE + (int64_t)threadIdx.x * DR_ITEMS) / 2;
        dr_check_pair(rv, dv.region_admission, c[0], c[1], stw & 0xFFu, (stw >> 8) & 0xFFu, s[0], e[0], s[1], e[1], region);
        dr_check_pair(rv, dv.region_admission, c[2], c[3], (stw >> 16) & 0xFFu, stw >> 24, s[2], e[2], s[3], e[3], region + 1);
      }
    } else if (dv.pair_check == 2 || dv.pair_check == 3) {
      // -gaps (and count without it, see DirectView): items 0 and 2 become the spans of their pairs, items 1 and 3 queries on a chromosome nobody has (nothing to count,
      // nothing to object to); a malformed pair is reported and counts nothing either
#pragma unroll
      for (int p = 0; p < 2; p++) {
        const int a = 2 * p, b = 2 * p + 1;
        const bool ok = c[a] == c[b] && ((xw >> (8 * a)) & 0xFFu) == ((xw >> (8 * b)) & 0xFFu) && s[b] >= s[a] && s[b] > e[a];
        if (!ok) { report_error(rv.err, q.index_base + (tile * DR_TILE + (int64_t)threadIdx.x * DR_ITEMS) / 2 + p, GTB_ERR_QUERY_REGION); c[a] = n_chrom; s[a] = 1; e[a] = 1; }
        else e[a] = e[b];
        c[b] = n_chrom; s[b] = 1; e[b] = 1;
      }
      general = 0;
      if (xw & 0x00F900F9u) general = ((xw & 0xF9u) ? 1u : 0u) | ((xw & 0xF90000u) ? 4u : 0u);     // the strands of items 0 and 2 only
#pragma unroll
      for (int i = 0; i < DR_ITEMS; i += 2)
        if (!(s[i] >= 1 && s[i] <= e[i])) general |= 1u << i;
    }
    // Position-sorted input, before any lookup: if the warp's 128 queries lie on one chromosome between two cells that share a
    // first slot, the later one free of evaluation points, no point lies anywhere among them -- per strand one reduction (or
    // nothing to count).  Two uniform lookups per strand instead of 128 divergent ones and everything that follows.  A warp
    // that fails the test (random input always does) does not try again for 16 tiles.
    if (!WEIGHTED && sorted_wait == 0u) {
      bool mine = general == 0u && c[0] == c[1] && c[1] == c[2] && c[2] == c[3] && c[0] < n_chrom;
      if (COVERAGE) mine = mine && (uint32_t)(e[0] - s[0]) == len0m1 && (uint32_t)(e[1] - s[1]) == len0m1 && (uint32_t)(e[2] - s[2]) == len0m1 && (uint32_t)(e[3] - s[3]) == len0m1;
      const uint32_t c_lo = __reduce_min_sync(0xffffffffu, c[0]), c_hi = __reduce_max_sync(0xffffffffu, c[0]);
      bool done_fast = false;
      if (__all_sync(0xffffffffu, mine) && c_lo == c_hi) {
        const uint32_t lo = __reduce_min_sync(0xffffffffu, (uint32_t)min(min(s[0], s[1]), min(s[2], s[3])));
        const uint32_t hi = __reduce_max_sync(0xffffffffu, (uint32_t)max(max(e[0], e[1]), max(e[2], e[3])));
        const uint32_t n_minus = dv.nsig == 2u ? __reduce_add_sync(0xffffffffu, (uint32_t)__popc(xw & 0x02020202u)) : 0u;
        const uint32_t cnt[2] = {32u * DR_ITEMS - n_minus, n_minus};
        uint32_t slot[2] = {0u, 0u};
        bool good = true, count_it[2] = {false, false};
#pragma unroll
        for (uint32_t sg = 0; sg < 2u; sg++) {
          if (cnt[sg] == 0u) continue;                                  // (with one block for both strands, everything is in cnt[0])
          const uint32_t b0 = (c_lo * dv.nsig + sg) * stride;
          const uint2 ea = dr_gather(dv.cells + b0 + min(lo >> cbits, last)), eb = dr_gather(dv.cells + b0 + min(hi >> cbits, last));
          if (ea.x & DR_GENERAL) { good = false; continue; }
          if (ea.x & DR_NOTHING) continue;                              // every start is beyond the group's last point (or there is no group): nothing to count
          good = good && !((ea.x | eb.x) & (DR_SCAN | DR_GENERAL | DR_NOTHING)) && ((ea.x ^ eb.x) & 0xFFFFFFu) == 0u && eb.y == 0xFFFFFFFFu;
          slot[sg] = ea.x & 0xFFFFFFu; count_it[sg] = true;
        }
        if (good) {
          if (lane == 0) {
            if (count_it[0]) dr_red64(dv.delta + slot[0], (ull)cnt[0] * unit);
            if (count_it[1]) dr_red64(dv.delta + slot[1], (ull)cnt[1] * unit);
          }
          done_fast = true;
        }
      }
      if (done_fast) continue;
      sorted_wait = 16u;
    } else if (!WEIGHTED) {
      sorted_wait--;
    }
    uint32_t base[DR_ITEMS];
    uint2 ent[DR_ITEMS];
#pragma unroll
    for (int i = 0; i < DR_ITEMS; i++) {
      base[i] = (min(c[i], n_chrom) * dv.nsig + ((xw >> (8 * i + 1)) & sigmask)) * stride;   // unknown chromosome -> the all-"nothing" block
      GTB_ASSERT(base[i] + last < dv.n_cells);
      ent[i] = dr_gather(dv.cells + base[i] + min((uint32_t)s[i] >> cbits, last));
    }
    uint32_t jS[DR_ITEMS], jE[DR_ITEMS], j0E[DR_ITEMS];
    uint32_t scan = 0;
    uint32_t skip = general;                                           // bit i: item i is not counted by the fast path
#pragma unroll
    for (int i = 0; i < DR_ITEMS; i++) {
      uint2 ee = ent[i];
      if (((uint32_t)s[i] ^ (uint32_t)e[i]) >> cbits) {                // a read that crosses a cell boundary (~1 %)
        ee = dr_gather(dv.cells + base[i] + min((uint32_t)e[i] >> cbits, last));
      }
      const uint32_t fl = ent[i].x | (ee.x & (DR_GENERAL | DR_SCAN));   // "nothing" is about the START: a stop beyond the last point lands in the sentinel slot
      general |= (fl & DR_GENERAL) ? (1u << i) : 0u;                   // a group whose points are all <= 0
      skip |= (fl & (DR_GENERAL | DR_NOTHING)) ? (1u << i) : 0u;
      jS[i] = (ent[i].x & 0xFFFFFFu) + dr_below(ent[i].y, (uint32_t)s[i] & cmask);
      jE[i] = (ee.x & 0xFFFFFFu) + dr_below(ee.y, (uint32_t)e[i] & cmask);
      scan |= (fl & DR_SCAN) ? (1u << i) : 0u;
      j0E[i] = ee.x & 0xFFFFFFu;
    }
    scan &= ~skip;
    if (scan) {                                                        // more than two points in one of the cells: walk them
#pragma unroll
      for (int i = 0; i < DR_ITEMS; i++)
        if ((scan >> i) & 1u) {
          jS[i] = dr_scan(rv.points, ent[i].x & 0xFFFFFFu, s[i]);
          jE[i] = dr_scan(rv.points, max(jS[i], j0E[i]), e[i]);
        }
    }
    // position-sorted input: the warp's 128 queries in one slot leave as one reduction
    const uint32_t lead = __shfl_sync(0xffffffffu, jS[0], 0);
    bool same = skip == 0u && !WEIGHTED;
#pragma unroll
    for (int i = 0; i < DR_ITEMS; i++) same = same && jS[i] == lead && jE[i] == lead && (!COVERAGE || (uint32_t)(e[i] - s[i]) == len0m1);
    if (__all_sync(0xffffffffu, same)) {
      if (lane == 0) dr_red64(dv.delta + lead, (ull)(32 * DR_ITEMS) * unit);
      continue;
    }
    // Clustered input (reads piled onto a few loci, e.g. the boundary reads a neighbouring genome shard hands over in one run):
    // hundreds of adds would reach one byte before its first spill lands.  Neighbouring queries sharing a slot give such input
    // away; the warp then looks for slots shared by eight or more of its lanes and sends ONE reduction for each of those.
    uint32_t done = skip;                                              // bit i: item i needs no shared atomic
    // (three neighbours, not two: the mates of a read pair sit next to each other in the stream and share a slot as a rule; any
    // run of five or more queries in one slot still has three of them in one thread)
    if (!WEIGHTED && __any_sync(0xffffffffu, (!(skip & 0x7u) && jS[0] == jS[1] && jS[1] == jS[2]) || (!(skip & 0xEu) && jS[1] == jS[2] && jS[2] == jS[3]))) {
#pragma unroll
      for (int i = 0; i < DR_ITEMS; i++) {
        const bool both = !((skip >> i) & 1u) && jS[i] == jE[i] && (!COVERAGE || (uint32_t)(e[i] - s[i]) == len0m1);
        const uint32_t peers = __match_any_sync(0xffffffffu, both ? jS[i] : 0x80000000u | (uint32_t)lane);
        if (both && __popc(peers) >= 8) {
          if ((peers & ((1u << lane) - 1u)) == 0u) dr_red64(dv.delta + jS[i], (ull)__popc(peers) * unit);
          done |= 1u << i;
        }
      }
    }
    // the four shared atomics go out back to back (an item with nothing for the "both" plane adds 0 to a word of its lane's own);
    // only then are the returned bytes looked at
    uint32_t old[DR_ITEMS];
    uint32_t counted = 0;                                              // bit i: item i went to its byte counter
#pragma unroll
    for (int i = 0; i < DR_ITEMS; i++) {
      const bool both = !((done >> i) & 1u) && jS[i] == jE[i] && (!COVERAGE || (uint32_t)(e[i] - s[i]) == len0m1) &&
                        (!WEIGHTED || (uint32_t)(wt[i] - 1) < 127u);
      counted |= both ? 1u << i : 0u;
      GTB_ASSERT(!both || (ull)jS[i] < (ull)rv.n_slots);
      GTB_ASSERT(((skip >> i) & 1u) || ((ull)jS[i] < (ull)rv.n_slots && (ull)jE[i] < (ull)rv.n_slots));
      old[i] = atomicAdd(&s_cnt[both ? (jS[i] >> 2) : dv.n_words + (uint32_t)lane], both ? (uint32_t)wt[i] << ((jS[i] & 3u) * 8u) : 0u);
    }
#pragma unroll
    for (int i = 0; i < DR_ITEMS; i++) {
      if (!((done >> i) & 1u)) {
        const ull w = (ull)(int64_t)wt[i];
        if ((counted >> i) & 1u) {                                     // the query's weight in slot jS of the "both" plane
          const uint32_t sh = (jS[i] & 3u) * 8u;
          const uint32_t ob = (old[i] >> sh) & 0xFFu, nb = ob + (uint32_t)wt[i];
          if (nb >= 128u) {
            if (ob < 128u) { atomicSub(&s_cnt[jS[i] >> 2], 128u << sh); dr_red64(dv.delta + jS[i], 128ull * unit); }
            overflowed |= nb > 255u;
          }
        } else if (jS[i] == jE[i]) {                                   // a query of another length (COVERAGE) or of a weight the bytes cannot take
          dr_red64(dv.delta + jS[i], COVERAGE ? w * (ull)((int64_t)e[i] - (int64_t)s[i] + 1) : w);
          if (COVERAGE && (uint32_t)(e[i] - s[i]) != len0m1) { odd++; odd_long += (uint32_t)(e[i] - s[i]) >= 255u ? 1u : 0u; }
        } else if (!COVERAGE && !WEIGHTED && dv.pair_check == 3) {      // an evaluation point inside the pair's span: the exact engine decides
          dr_append_pair(dv, (int32_t)c[i], s[i], i == 0 ? (int)ce.x : (int)ce.z, i == 0 ? (int)cs.y : (int)cs.w, e[i], (int)(int8_t)((stw >> (8 * i)) & 0xFFu));
        } else if (!COVERAGE && !WEIGHTED && dv.pair_check == 4 &&
                   dv.q_off[tile * DR_TILE + (int64_t)threadIdx.x * DR_ITEMS + i + 1] - dv.q_off[tile * DR_TILE + (int64_t)threadIdx.x * DR_ITEMS + i] > 1) {
          dv.ex_chrom[atomicAdd(dv.ex_count, 1u)] = (int32_t)(tile * DR_TILE + (int64_t)threadIdx.x * DR_ITEMS + i);     // ... for a region of several blocks
        } else {
          dr_red64(dv.delta + (ull)H_SCNT * (ull)rv.n_slots + jS[i], w);
          dr_red64(dv.delta + (ull)H_ECNT * (ull)rv.n_slots + jE[i], w);
          if (COVERAGE) {
            dr_red64(dv.delta + (ull)H_SSUM * (ull)rv.n_slots + jS[i], w * (ull)(int64_t)s[i]);
            dr_red64(dv.delta + (ull)H_ESUM * (ull)rv.n_slots + jE[i], w * (ull)(int64_t)e[i]);
          }
        }
      }
    }
    if (general) {
#pragma unroll
      for (int i = 0; i < DR_ITEMS; i++)
        if ((general >> i) & 1u) {
          const int64_t item = tile * DR_TILE + (int64_t)threadIdx.x * DR_ITEMS + i;
          const int64_t index = q.index_base + (dv.pair_check == 2 || dv.pair_check == 3 ? item / 2 : item);
          if (!COVERAGE && !WEIGHTED && dv.pair_check == 4)
            dr_general_span<COVERAGE>(rv, dv, (int32_t)c[i], s[i], e[i], (int)(int8_t)((stw >> (8 * i)) & 0xFFu), index, item);
          else if (!COVERAGE && !WEIGHTED && dv.pair_check == 3)
            dr_general_pair(rv, dv, (int32_t)c[i], s[i], i == 0 ? (int)ce.x : (int)ce.z, i == 0 ? (int)cs.y : (int)cs.w, e[i], (int)(int8_t)((stw >> (8 * i)) & 0xFFu), index);
          else
            dr_general<COVERAGE>(rv, (int32_t)c[i], s[i], e[i], (int)(int8_t)((stw >> (8 * i)) & 0xFFu), index, wt[i]);
        }
    }
  }
  // the last, partial tile: general path
  if ((int64_t)blockIdx.x == n_full % gridDim.x) {
    if (dv.pair_check < 2)
      for (int64_t r = n_full * DR_TILE + threadIdx.x; r < n; r += DR_THREADS)
        dr_general<COVERAGE>(rv, q.chrom[r], q.start[r], q.stop[r], (int)q.strand[r], q.index_base + r, WEIGHTED ? (int64_t)q.weight[r] : 1);
    if (dv.pair_check == 4)
      for (int64_t r = n_full * DR_TILE + threadIdx.x; r < n; r += DR_THREADS)
        dr_general_span<COVERAGE>(rv, dv, q.chrom[r], q.start[r], q.stop[r], (int)q.strand[r], q.index_base + r, r);
    if (dv.pair_check >= 1 && dv.pair_check <= 3)
      for (int64_t p = n_full * (DR_TILE / 2) + threadIdx.x; p < n / 2; p += DR_THREADS) {
        const int32_t c0 = q.chrom[2 * p], c1 = q.chrom[2 * p + 1], s0 = q.start[2 * p], e0 = q.stop[2 * p], s1 = q.start[2 * p + 1], e1 = q.stop[2 * p + 1];
        const uint32_t b0 = (uint32_t)(uint8_t)q.strand[2 * p], b1 = (uint32_t)(uint8_t)q.strand[2 * p + 1];
        if (dv.pair_check == 1) { if (COVERAGE) dr_check_pair(rv, dv.region_admission, (uint32_t)c0, (uint32_t)c1, b0, b1, s0, e0, s1, e1, q.index_base + p); }
        else if (!(c0 == c1 && b0 == b1 && s1 >= s0 && s1 > e0)) report_error(rv.err, q.index_base + p, GTB_ERR_QUERY_REGION);
        else if (dv.pair_check == 3) dr_general_pair(rv, dv, c0, s0, e0, s1, e1, (int)(int8_t)b0, q.index_base + p);
        else dr_general<COVERAGE>(rv, c0, s0, e1, (int)(int8_t)b0, q.index_base + p);
      }
  }
  if (overflowed) atomicMax(dv.flag, dv.gen);
  if (COVERAGE && odd) { atomicAdd(dv.flag + 1, odd); if (odd_long) atomicAdd(dv.flag + 2, odd_long); }
  __syncthreads();
  uint4 *row = reinterpret_cast<uint4 *>(dv.cta_counts + (size_t)blockIdx.x * dv.n_words);
  for (uint32_t i = threadIdx.x; i < dv.n_words / 4; i += DR_THREADS) row[i] = reinterpret_cast<const uint4 *>(s_cnt)[i];
}

// Sums the CTAs' byte counters and the delta planes into the index's histogram planes and clears the delta planes -- or, if a
// byte counter overflowed somewhere in this batch (flag == the batch's generation), only clears them and then sends every
// query through the general path straight into the histogram planes (nothing else touches those in that case).
template <bool COVERAGE>
__global__ void __launch_bounds__(256) direct_commit_kernel(DirectView dv, QueryView q, RankView rv_hist, int64_t n_slots, unsigned rows) {
  // (pair_check >= 2: the queries are spans of pairs, not the intervals of q -- a discarded batch is not replayed here, the host
  // sends it down the general path)
  // a block = 32 counter words x 8 groups of rows: each thread sums its share of the rows of one word (coalesced across the warp),
  // shared memory joins the eight partial sums, the first warp updates the planes
  __shared__ uint32_t s_part[8][32][4];
  ull *hist = rv_hist.hist;
  const bool discard = dv.flag[0] == dv.gen;
  const uint32_t lane = threadIdx.x & 31u, grp = threadIdx.x >> 5;
  const uint32_t w = blockIdx.x * 32u + lane;
  uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  if (!discard && w < dv.n_words)
    for (unsigned r = grp; r < rows; r += 8) {
      const uint32_t v = dv.cta_counts[(size_t)r * dv.n_words + w];
      a0 += v & 0xFFu; a1 += (v >> 8) & 0xFFu; a2 += (v >> 16) & 0xFFu; a3 += v >> 24;
    }
  s_part[grp][lane][0] = a0; s_part[grp][lane][1] = a1; s_part[grp][lane][2] = a2; s_part[grp][lane][3] = a3;
  __syncthreads();
  if (threadIdx.x < 128) {                                             // thread t: slot (t & 3) of word (t >> 2) of this block
    const uint32_t wl = threadIdx.x >> 2, b = threadIdx.x & 3u;
    const int64_t j = ((int64_t)blockIdx.x * 32 + wl) * 4 + b;
    if (j < n_slots) {
      uint32_t acc = 0;
      for (int g = 0; g < 8; g++) acc += s_part[g][wl][b];
      // COVERAGE: a counted query is worth the batch's common length (direct_count_kernel)
      const ull unit = COVERAGE ? (ull)(int64_t)(int32_t)dr_common_len(q) + 1ull : 1ull;
      for (int p = 0; p < (COVERAGE ? H_PLANES_COVERAGE : H_PLANES_COUNT); p++) {
        const ull d = dv.delta[(int64_t)p * n_slots + j];
        const ull add = discard ? 0ull : d + (p == H_BOTH ? (ull)acc * unit : 0ull);
        if (add) hist[(int64_t)p * n_slots + j] += add;
        if (d) dv.delta[(int64_t)p * n_slots + j] = 0;
      }
    }
  }
  if (!discard || dv.pair_check >= 2) return;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < q.n_regions; r += stride)
    dr_general<COVERAGE>(rv_hist, q.chrom[r], q.start[r], q.stop[r], (int)q.strand[r], q.index_base + r, q.weight ? (int64_t)q.weight[r] : 1);
}

template <typename T>
int upload_d(gtb_ctx *ctx, dbuf<T> &d, const std::vector<T> &h) {
  GTB_TRY(d.reserve(ctx, std::max<size_t>(h.size(), 1)));
  if (!h.empty()) GTB_CUDA_OK(ctx, cudaMemcpyAsync(d.p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  return GTB_OK;
}

}  // namespace

struct gtb_direct_state {
  bool ready = false, failed = false, off = false;
  bool off_lengths = false;                                              // off because the reads have no common length (coverage): survives a reset
  int cbits = 0;
  uint32_t n_cells = 0, n_words = 0, nsig = 2, stride = 0;
  unsigned grid = 0;
  size_t smem = 0;
  dbuf<uint2> d_cells;
  dbuf<uint32_t> d_cta_counts, d_flag;
  dbuf<ull> d_delta;
  int64_t queries_since_check = 0, queries_in_check = 0;            // queries since the last flag copy was issued / covered by the copy in flight
  uint32_t gen = 0, odd_seen = 0, long_seen = 0;
  dbuf<int32_t> ex_chrom, ex_start, ex_stop;                           // pair_check == 3: the exception list (gtb_direct_exceptions)
  dbuf<int8_t> ex_strand;
  dbuf<uint32_t> ex_count;
  uint32_t *h_sync = nullptr;                                          // pinned: [0] flag word, [1] exception count of a pair batch
  int64_t n_exceptions = 0;
  uint32_t *h_flag = nullptr;                                          // pinned: where the flag words land
  cudaEvent_t flag_copied = nullptr;
  bool check_pending = false;
};

int gtb_direct_prepare(gtb_index *ix) {
  if (ix->direct && (ix->direct->ready || ix->direct->failed)) return ix->direct->ready ? GTB_OK : GTB_ERR_UNSUPPORTED;
  gtb_ctx *ctx = ix->ctx;
  if (!ix->direct) ix->direct = new gtb_direct_state();
  gtb_direct_state *ds = ix->direct;
  ds->failed = true;                                                    // until proven otherwise
  if (ix->n_slots == 0 || ix->n_slots >= (1 << 24)) return GTB_ERR_UNSUPPORTED;
  const int G = ix->n_groups;
  const uint32_t n_words = (uint32_t)(((ix->n_slots + 3) / 4 + 3) & ~(int64_t)3);
  const size_t smem = (size_t)n_words * 4 + 32 * 4;
  // The gathers need L1 to track their misses in: measured (profiles/microbench/gather_rate_b200.txt), a gather costs 1.02
  // SM-cycles with up to 192 KB of the SM's 256 KB carved out as shared memory and 1.85 with 224 KB.
  if (smem > std::min<size_t>(ctx->smem_optin, (size_t)192 * 1024)) return GTB_ERR_UNSUPPORTED;
  // largest evaluation point per group (0: none, -1: only points <= 0)
  std::vector<int32_t> gsize((size_t)std::max(G, 1), 0);
  int32_t gmax = 0;
  uint64_t span = 0;
  for (int g = 0; g < G; g++) {
    const int32_t gb = ix->h_goff[g], ge = ix->h_goff[g + 1];
    if (ge - gb >= 2) { const int32_t mx = ix->h_points[ge - 2]; gsize[g] = mx >= 1 ? mx : -1; }
    if (gsize[g] > 0) { span += (uint64_t)gsize[g] + 1; gmax = std::max(gmax, gsize[g]); }
  }
  if (span == 0) return GTB_ERR_UNSUPPORTED;
  // cell width: the cells that queries can touch (those up to each group's last point) should stay well inside L2
  int cbits = 10;
  while (cbits < 14 && (span >> cbits) * sizeof(uint2) > DR_TABLE_BYTES_TARGET) cbits++;
  if (const char *env = getenv("GTB_DIRECT_CELL_BITS")) cbits = std::max(4, std::min(14, atoi(env)));
  const int cp = ix->h_class_of[(uint8_t)'+'], cm = ix->h_class_of[(uint8_t)'-'];
  const uint32_t nsig = cp == cm ? 1u : 2u;
  const uint64_t stride = ((uint64_t)gmax >> cbits) + 2;                // cells up to the largest last point, and one "nothing" cell to clamp to
  const uint64_t blocks = ((uint64_t)ix->n_chrom + 1) * nsig;
  const uint64_t cells = blocks * stride;
  if (cells * sizeof(uint2) > ((size_t)1 << 30) || cells >= ((uint64_t)1 << 31)) return GTB_ERR_UNSUPPORTED;
  std::vector<uint2> tab((size_t)cells, make_uint2(DR_NOTHING, 0xFFFFFFFFu));      // no such group: nothing to count
  size_t many = 0;
  for (int c = 0; c < ix->n_chrom; c++)
    for (uint32_t sg = 0; sg < nsig; sg++) {
      const int cls = sg ? cm : cp;
      if (cls < 0) continue;
      const int g = c * ix->n_class + cls;
      uint2 *blk = tab.data() + ((size_t)c * nsig + sg) * stride;
      const int32_t gb = ix->h_goff[g], ge = ix->h_goff[g + 1] - 1;     // [gb, ge): the group's points without the sentinel
      if (ix->h_goff[g + 1] == gb || gsize[g] == 0) continue;           // no points: nothing to count
      if (gsize[g] < 0) {                                               // only points <= 0: left to the general path
        for (uint64_t x = 0; x < stride; x++) blk[x] = make_uint2(DR_GENERAL, 0xFFFFFFFFu);
        continue;
      }
      const uint64_t nc = ((uint64_t)gsize[g] >> cbits) + 1;            // cells that hold a point or precede one
      int32_t j = gb;
      for (uint64_t x = 0; x < stride; x++) {
        if (x >= nc) { blk[x] = make_uint2((uint32_t)ge | DR_NOTHING, 0xFFFFFFFFu); continue; }   // wholly beyond the last point: the sentinel slot
        const int64_t lo = (int64_t)x << cbits, hi = lo + ((int64_t)1 << cbits);
        while (j < ge && ix->h_points[j] < lo) j++;
        uint32_t p[2] = {DR_NOPOINT, DR_NOPOINT};
        int32_t t = j, cnt = 0;
        while (t < ge && ix->h_points[t] < hi) { if (cnt < 2) p[cnt] = (uint32_t)(ix->h_points[t] - lo); cnt++; t++; }
        many += cnt > 2 ? 1 : 0;
        blk[x] = make_uint2((uint32_t)j | (cnt > 2 ? DR_SCAN : 0u), p[0] | (p[1] << 16));
      }
    }
  ds->cbits = cbits; ds->n_cells = (uint32_t)cells; ds->n_words = n_words; ds->smem = smem; ds->nsig = nsig; ds->stride = (uint32_t)stride;
  ds->grid = (unsigned)ctx->sm_count;
  GTB_TRY(upload_d(ctx, ds->d_cells, tab));
  GTB_TRY(ds->d_cta_counts.reserve(ctx, (size_t)ds->grid * n_words));
  GTB_TRY(ds->d_delta.reserve(ctx, (size_t)ix->planes * ix->n_slots));
  GTB_TRY(ds->d_flag.reserve(ctx, 3));
  if (!ds->h_flag) GTB_CUDA_OK(ctx, cudaHostAlloc((void **)&ds->h_flag, 3 * sizeof(uint32_t), cudaHostAllocDefault));
  if (!ds->flag_copied) GTB_CUDA_OK(ctx, cudaEventCreateWithFlags(&ds->flag_copied, cudaEventDisableTiming));
  GTB_CUDA_OK(ctx, cudaMemsetAsync(ds->d_delta.p, 0, (size_t)ix->planes * ix->n_slots * sizeof(ull), ctx->stream));
  GTB_CUDA_OK(ctx, cudaMemsetAsync(ds->d_flag.p, 0, 3 * sizeof(uint32_t), ctx->stream));
  GTB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  if (getenv("GTB_DEBUG_DIRECT"))
    fprintf(stderr, "[gtb direct] slots %lld cells %llu (2^%d bp, %u per block, %.1f MB allocated, %.1f MB within reach) general cells %zu smem %zu\n",
            (long long)ix->n_slots, (unsigned long long)cells, cbits, (unsigned)stride, cells * 8 / 1e6, (double)(span >> cbits) * 8 / 1e6, many, smem);

  ds->ready = true; ds->failed = false;
  return GTB_OK;
}

// the index fits the engine and the watchdog has not sent it away
bool gtb_direct_usable(gtb_index *ix) { return gtb_direct_prepare(ix) == GTB_OK && !ix->direct->off; }

bool gtb_direct_supported(gtb_index *ix, const QueryView &q, bool batch_multi) {
  if (batch_multi || q.region_offset) return false;                    // single-interval batches
  if (q.n_regions >= ((int64_t)1 << 40)) return false;
  if ((((uintptr_t)q.chrom | (uintptr_t)q.start | (uintptr_t)q.stop | (uintptr_t)q.weight) & 15) != 0 || ((uintptr_t)q.strand & 3) != 0) return false;   // 128-bit loads
  if (gtb_direct_prepare(ix) != GTB_OK) return false;
  return !ix->direct->off;
}

int gtb_direct_accumulate(gtb_index *ix, const QueryView &q, int pair_check, const int64_t *region_offsets) {
  gtb_ctx *ctx = ix->ctx;
  if (gtb_direct_prepare(ix) != GTB_OK) return gtb_fail(ctx, GTB_ERR_UNSUPPORTED, "direct engine cannot serve this index");
  gtb_direct_state *ds = ix->direct;
  if (q.region_offset) return gtb_fail(ctx, GTB_ERR_UNSUPPORTED, "direct engine takes single-interval batches");
  // Watchdog, without ever making the host wait: after a batch the two flag words travel to pinned host memory behind it; the
  // next batch that finds that copy complete looks at them.  A replayed batch means byte counters overflow on this input, many
  // odd-length reads under coverage mean a reduction each -- either way the BUCKET engine serves this index from then on.
  if (ds->check_pending && cudaEventQuery(ds->flag_copied) == cudaSuccess) {
    ds->check_pending = false;
    if (ds->h_flag[0] != 0) ds->off = true;
    // reads of every length: BUCKET counts them in shared memory as long as its elements can hold the length (< 256 bp);
    // longer ones (spans of read pairs under -gaps) it would hand to the general step one by one -- those stay here, a
    // reduction each
    const uint32_t odd_new = ds->h_flag[1] - ds->odd_seen, long_new = ds->h_flag[2] - ds->long_seen;
    if ((int64_t)odd_new > ds->queries_in_check / 16 && (int64_t)long_new * 2 < (int64_t)odd_new) ds->off = ds->off_lengths = true;
    ds->odd_seen = ds->h_flag[1]; ds->long_seen = ds->h_flag[2];
    if (ds->off && pair_check >= 2) return GTB_ERR_UNSUPPORTED;                                 // spans of pairs: the caller's general path
    if (ds->off && gtb_bucket_supported(ix, q, false)) return gtb_bucket_accumulate(ix, q);   // (else this batch still goes here: slow, not wrong)
    if (ds->off && q.weight) return GTB_ERR_UNSUPPORTED;                                       // weighted: the caller's general rank step
  }
  cudaGetLastError();                                                   // cudaErrorNotReady of the query above is not an error
  RankView rv;
  rv.n_chrom = ix->n_chrom; rv.n_class = ix->n_class; rv.class_of = ix->d_class_of.p; rv.chrom_present = ix->d_present.p;
  rv.goff = ix->d_goff.p; rv.points = ix->d_points.p; rv.n_slots = ix->n_slots; rv.hist = ds->d_delta.p; rv.err = ix->d_err.p; rv.admission = ix->admission();
  DirectView dv;
  dv.cbits = ds->cbits; dv.n_chrom = ix->n_chrom; dv.nsig = ds->nsig; dv.stride = ds->stride; dv.cells = ds->d_cells.p;
  dv.n_cells = ds->n_cells; dv.n_words = ds->n_words; dv.cta_counts = ds->d_cta_counts.p; dv.delta = ds->d_delta.p; dv.flag = ds->d_flag.p;
  if (++ds->gen == 0) ds->gen = 1;                                      // (a wrap after 2^32 batches could only cost a spurious replay)
  dv.gen = ds->gen;
  dv.pair_check = pair_check; dv.region_admission = ix->sorted_rules ? 1 : 0;
  dv.ex_count = nullptr; dv.ex_chrom = dv.ex_start = dv.ex_stop = nullptr; dv.ex_strand = nullptr; dv.q_off = region_offsets;
  ds->n_exceptions = 0;
  if (pair_check >= 2 && !ds->h_sync) GTB_CUDA_OK(ctx, cudaHostAlloc((void **)&ds->h_sync, 2 * sizeof(uint32_t), cudaHostAllocDefault));
  if (pair_check == 4 && !region_offsets) return gtb_fail(ctx, GTB_ERR_ARG, "direct engine: region offsets missing");
  if (pair_check == 4) {
    GTB_TRY(ds->ex_chrom.reserve(ctx, (size_t)q.n_regions + 2));        // region numbers, room for every region
    GTB_TRY(ds->ex_count.reserve(ctx, 1));
    GTB_CUDA_OK(ctx, cudaMemsetAsync(ds->ex_count.p, 0, sizeof(uint32_t), ctx->stream));
    dv.ex_count = ds->ex_count.p; dv.ex_chrom = ds->ex_chrom.p;
  }
  if (pair_check == 3) {
    const size_t cap = (size_t)q.n_regions + 2;                         // two entries per pair, room for every pair
    GTB_TRY(ds->ex_chrom.reserve(ctx, cap)); GTB_TRY(ds->ex_start.reserve(ctx, cap)); GTB_TRY(ds->ex_stop.reserve(ctx, cap)); GTB_TRY(ds->ex_strand.reserve(ctx, cap));
    GTB_TRY(ds->ex_count.reserve(ctx, 1));
    GTB_CUDA_OK(ctx, cudaMemsetAsync(ds->ex_count.p, 0, sizeof(uint32_t), ctx->stream));
    dv.ex_count = ds->ex_count.p; dv.ex_chrom = ds->ex_chrom.p; dv.ex_start = ds->ex_start.p; dv.ex_stop = ds->ex_stop.p; dv.ex_strand = ds->ex_strand.p;
  }
  RankView rv_hist = rv;
  rv_hist.hist = ix->d_hist.p;
  const int64_t tiles = q.n_regions / DR_TILE;
  const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>((int64_t)ds->grid, tiles));
#define GTB_DIRECT_LAUNCH(COV, WGT, NAME)                                                                                                   \
  do {                                                                                                                                     \
    GTB_CUDA_OK(ctx, cudaFuncSetAttribute(direct_count_kernel<COV, WGT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ds->smem));   \
    GTB_LAUNCH(ctx, NAME, (direct_count_kernel<COV, WGT>), grid, DR_THREADS, ds->smem, q, rv, dv);                                        \
    GTB_TRY(gtb_check_launch(ctx));                                                                                                        \
    GTB_LAUNCH(ctx, "direct_commit", direct_commit_kernel<COV>, (ds->n_words + 31) / 32, 256, 0, dv, q, rv_hist, ix->n_slots, grid);       \
  } while (0)
  if (ix->op == GTB_OP_COVERAGE) {
    if (q.weight) GTB_DIRECT_LAUNCH(true, true, "direct_coverage_weighted"); else GTB_DIRECT_LAUNCH(true, false, "direct_coverage");
  } else {
    if (q.weight) GTB_DIRECT_LAUNCH(false, true, "direct_count_weighted"); else GTB_DIRECT_LAUNCH(false, false, "direct_count");
  }
#undef GTB_DIRECT_LAUNCH
  if (pair_check >= 2) {
    // The queries were spans formed in the kernel, so a batch whose byte counters overflowed could not be replayed there (the
    // commit kernel has dropped it): the host has to know now.  One wait per batch; the exception count comes with it.
    GTB_CUDA_OK(ctx, cudaMemcpyAsync(ds->h_sync, ds->d_flag.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (pair_check >= 3) GTB_CUDA_OK(ctx, cudaMemcpyAsync(ds->h_sync + 1, ds->ex_count.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    GTB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    if (ds->h_sync[0] == ds->gen) { ds->off = true; return GTB_ERR_UNSUPPORTED; }       // nothing of this batch has reached the planes
    if (pair_check >= 3) ds->n_exceptions = (int64_t)ds->h_sync[1];
  }
  ds->queries_since_check += q.n_regions;
  if (!ds->check_pending) {
    GTB_CUDA_OK(ctx, cudaMemcpyAsync(ds->h_flag, ds->d_flag.p, 3 * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    GTB_CUDA_OK(ctx, cudaEventRecord(ds->flag_copied, ctx->stream));
    ds->check_pending = true;
    ds->queries_in_check = ds->queries_since_check;
    ds->queries_since_check = 0;
  }
  return gtb_check_launch(ctx);
}

// gtb_index_reset: the values go back to zero and a new query stream begins, so what the watchdog concluded about the last
// one's skew (byte counters overflowing) is forgotten -- the engine is tried again.
// The device-side odd-length counter keeps running; the host's copy of it is what the next check subtracts.
void gtb_direct_reset(gtb_index *ix) {
  gtb_direct_state *ds = ix->direct;
  if (!ds) return;
  ds->off = ds->off_lengths;                                            // reads of every length: a property of the data, expected to last
  ds->queries_since_check = 0;
}

// pair_check == 4: the numbers of the regions of the last batch that the candidate-enumeration engine has to count
int64_t gtb_direct_exception_list(gtb_index *ix, const int32_t **list) {
  gtb_direct_state *ds = ix->direct;
  if (!ds || ds->n_exceptions == 0) return 0;
  *list = ds->ex_chrom.p;
  return ds->n_exceptions;
}

// pair_check == 3: the pairs of the last batch that the candidate-enumeration engine has to count (two intervals each, no offsets)
int64_t gtb_direct_exceptions(gtb_index *ix, QueryView *out) {
  gtb_direct_state *ds = ix->direct;
  if (!ds || ds->n_exceptions == 0) return 0;
  out->n_regions = ds->n_exceptions; out->chrom = ds->ex_chrom.p; out->start = ds->ex_start.p; out->stop = ds->ex_stop.p; out->strand = ds->ex_strand.p;
  out->weight = nullptr; out->region_offset = nullptr; out->interval_base = 0;
  return ds->n_exceptions;
}

void gtb_direct_destroy(gtb_index *ix) {
  gtb_direct_state *ds = ix->direct;
  if (!ds) return;
  ds->ex_chrom.release(); ds->ex_start.release(); ds->ex_stop.release(); ds->ex_strand.release(); ds->ex_count.release();
  if (ds->h_sync) cudaFreeHost(ds->h_sync);
  ds->d_cells.release(); ds->d_cta_counts.release(); ds->d_flag.release(); ds->d_delta.release();
  if (ds->h_flag) cudaFreeHost(ds->h_flag);
  if (ds->flag_copied) cudaEventDestroy(ds->flag_copied);
  delete ds;
  ix->direct = nullptr;
}
