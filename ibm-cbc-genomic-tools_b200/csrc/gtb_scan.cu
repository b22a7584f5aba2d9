// gtb_scan.cu -- genome-wide sliding-window read counts (genomic_scans counts).
//
// Replaces UnsortedGenomicRegionSetScanner (genomic_intervals.cpp:5019-5141): a dense micro-window
// histogram per (chromosome, strand) "slot", a sliding sum of win_size/win_step micro-windows, and
// the `v >= MIN_READS` filter of RunCounts (genomic_scans.cpp:421-428) as a device stream
// compaction that preserves the reference's output order (chromosome id ascending, '+' then '-',
// window ascending).
//
// Two ways to build the histogram:
//   * scan_histogram_kernel: one global reduction per read.  The table of a genome (hg19, step 50: 0.5 GB) is far larger than
//     L2, so every reduction is a 32-byte read-modify-write in HBM.  Kept for small batches, weighted reads and tiny genomes.
//   * bucketed (default for large unweighted batches): the write-combining partition of gtb_wc_partition.cuh (this file's
//     instance: 1 024 threads, one CTA per SM, up to 1 024 buckets) sorts the reads' micro-window numbers into genome buckets
//     (13 B in, 4 B out per read); scan_bucket_hist_kernel then counts each bucket in shared memory -- as many 16-bit counters
//     per CTA as an SM holds, a bucket wider than that is walked by several CTAs, each keeping its own sub-range (the bucket's
//     elements come from L2 after the first of them) -- and adds the non-zero counters to the table with plain coalesced
//     read-modify-writes: a bucket range belongs to exactly one CTA.
// The table is a uint32 plane plus a carry plane that stays untouched until an entry passes 2^32 (ScanTable below); the
// qualifying windows stay on the device in a compact form and are widened by gtb_scan_fetch.
// genomic_scans peaks (PeakFinder::Run, genomic_scans.cpp:298-354) walks two such tables in step: scan_peaks_kernel.
#include "gtb_internal.cuh"
#ifndef GTB_SCAN_WC_THREADS
#define GTB_SCAN_WC_THREADS 1024
#endif
#define GTB_WC_THREADS GTB_SCAN_WC_THREADS            // this file's instance of the partition (1 024: up to 1 024 buckets, one CTA per SM)
#include "gtb_wc_partition.cuh"
#include <algorithm>

typedef unsigned long long ull;

namespace {

struct SlotTable {                 // device pointers, one entry per slot (+1 for offsets)
  int32_t n_slots;
  const int64_t *hist_off;         // [n_slots+1] offset of the slot's micro-windows in hist
  const int64_t *win_off;          // [n_slots+1] offset of the slot's windows in the window space
  const int32_t *spurious;         // [n_slots] 1 if the slot emits the reference's single spurious window
  const int32_t *chrom;            // [n_slots]
  const int8_t *strand;            // [n_slots] '+' / '-'
};

struct ReadView {
  int64_t n_regions, n_intervals;
  const int32_t *chrom, *start, *stop;
  const int8_t *strand;
  const int32_t *weight;
  const int64_t *region_offset;
  int64_t interval_base;
};

// The micro-window table.  The reference's counters are 64-bit (`long`); here a value is a uint32 plane plus a carry plane that
// is neither written nor read until some entry passes 2^32 (flags[0]), which halves the table's traffic for every real input
// and keeps the arithmetic exact (mod 2^64, like the reference's) for the others.
struct ScanTable {
  uint32_t *lo, *hi;
  uint32_t *flags;                 // [0] the carry plane holds something  [1] some qualifying window's value needs more than 32 bits
};
__device__ __forceinline__ void table_carry(const ScanTable &t, int64_t m, uint32_t c) { atomicAdd(t.hi + m, c); t.flags[0] = 1u; }
__device__ __forceinline__ void table_add(const ScanTable &t, int64_t m, ull w) {          // entry m += w, atomically
  const uint32_t wlo = (uint32_t)w;
  uint32_t whi = (uint32_t)(w >> 32);
  if (wlo) { const uint32_t old = atomicAdd(t.lo + m, wlo); whi += old + wlo < old ? 1u : 0u; }
  if (whi) table_carry(t, m, whi);
}
__device__ __forceinline__ ull table_get(const ScanTable &t, bool wide, int64_t m) {
  return (ull)t.lo[m] | (wide ? (ull)t.hi[m] << 32 : 0ull);
}

// one thread per read region; every interval of the region counts once (genomic_intervals.cpp:5039)
__global__ void __launch_bounds__(256) scan_histogram_kernel(ReadView q, int32_t n_chrom, const int32_t *__restrict__ slot_of_chrom,
                                                             const int64_t *__restrict__ hist_off, long long win_step, int op,
                                                             int ignore_strand, ScanTable hist) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < q.n_regions; r += stride) {
    int64_t lo = r, hi = r + 1;
    if (q.region_offset) { lo = q.region_offset[r] - q.interval_base; hi = q.region_offset[r + 1] - q.interval_base; }
    const long long w = q.weight ? (long long)q.weight[r] : 1;
    for (int64_t i = lo; i < hi; i++) {
      const long long s = q.start[i], e = q.stop[i];
      if (s > e || e <= 0) continue;                                 // :5040
      const int32_t c = q.chrom[i];
      if (c < 0 || c >= n_chrom) continue;
      int slot = slot_of_chrom[c];
      if (slot < 0) continue;                                        // chromosome not in the genome file, :5041-5042
      const long long pos = op == '1' ? s : s + (e - s) / 2;         // :5044-5045
      if (pos < 1) continue;                                         // :5049
      const long long win = (pos - 1) / win_step;                    // 0-based micro-window
      if (!(ignore_strand || q.strand[i] == '+')) slot += 1;         // :5048
      const long long n_micro = hist_off[slot + 1] - hist_off[slot];
      if (win >= n_micro) continue;                                  // :5049  (w <= n)
      table_add(hist, hist_off[slot] + win, (ull)w);
    }
  }
}

// ---- bucketed histogram ----------------------------------------------------------------------------------------------------
// Front of the write-combining partition for reads.  Table: one 16-byte entry per (chromosome, strand selector):
//   x = last position that counts (n_micro * win_step; 0: chromosome not in the genome file), y = first micro-window of the slot.
// Element = micro-window number inside the bucket; bucket = global micro-window number >> mb.
struct ScanFront {
  static constexpr bool FAIL_IS_SLOW = false;         // a read that counts nowhere is simply dropped (:5040-5049)
  const uint4 *tab;                                   // [2 * n_chrom + 2], the last two entries are empty
  uint32_t n_chrom;
  uint32_t magic; int shift;                          // (pos - 1) / win_step == ((pos - 1) * magic) >> shift for pos - 1 < 2^31
  uint32_t mb;
  int centre, ignore_strand;
  ScanTable hist;
  __device__ __forceinline__ uint32_t table_size() const { return 2u * n_chrom + 2u; }
  __device__ __forceinline__ uint4 table_entry(uint32_t i) const { return tab[i]; }
  __device__ __forceinline__ uint32_t table_index(int32_t c, uint32_t xw, int i) const {
    const uint32_t sel = ignore_strand ? 0u : (((xw >> (8 * i)) & 0xFFu) != 0u ? 1u : 0u);       // '+' -> forward, anything else -> reverse (:5048)
    return 2u * min((uint32_t)c, n_chrom) + sel;
  }
  __device__ __forceinline__ uint32_t slow_mask(uint32_t) const { return 0u; }
  __device__ __forceinline__ bool micro(uint4 g, int32_t s, int32_t e, uint32_t &m) const {
    const uint32_t d = (uint32_t)e - (uint32_t)s;                                                // exact when s <= e
    const int32_t pos = centre ? s + (int32_t)(d >> 1) : s;                                      // :5044-5045
    m = g.y + (uint32_t)(((uint64_t)(uint32_t)(pos - 1) * magic) >> shift);
    return s <= e && (uint32_t)(pos - 1) < g.x;                                                  // pos >= 1 (hence stop > 0) and inside the micro-windows
  }
  __device__ __forceinline__ uint32_t classify(uint4 g, int32_t s, int32_t e, uint32_t &elem) const {
    uint32_t m;
    const bool ok = micro(g, s, e, m);
    elem = m & ((1u << mb) - 1u);
    return ok ? m >> mb : WC_NONE;
  }
  __device__ __forceinline__ uint32_t resolve(uint32_t bucket, int32_t, int32_t, int32_t, uint32_t, int64_t) const { return bucket; }
  __device__ __forceinline__ void divert(int32_t c, int32_t s, int32_t e, uint32_t sbyte, int64_t) const {
    const uint32_t sel = ignore_strand ? 0u : (sbyte != (uint32_t)'+' ? 1u : 0u);
    uint32_t m;
    if (micro(tab[2u * min((uint32_t)c, n_chrom) + sel], s, e, m)) table_add(hist, m, 1ull);
  }
};

// CTA (bucket b, sub-range z): counts the elements of b that fall into [z * span, (z + 1) * span) and adds them to the table;
// span = as many 16-bit counters as one SM's shared memory holds (hg19, step 50: buckets of 2^18 micro-windows, three CTAs each).
// Counters are 16 bits, two to a word.  The thread whose add takes a counter from below 2^15 to 2^15 or more moves 2^15 to the
// table; a thread that finds a counter at 3 * 2^14 or more (that correction still pending) takes its own add back and sends it to
// the table instead.  Every thread has at most one add of <= 4 in flight, so a counter stays below 3 * 2^14 + 4 096 < 2^16.
template <int SB_THREADS, int SB_UNROLL_>
__global__ void __launch_bounds__(SB_THREADS, 1) scan_bucket_hist_kernel(WcView wv, uint32_t mb, uint32_t n_sub, uint32_t words, ScanTable hist) {
  extern __shared__ __align__(16) uint32_t s_cnt[];   // [words] counters + [32] one dummy word per lane
  const uint32_t b = blockIdx.x / n_sub, z = blockIdx.x % n_sub;
  const uint32_t first = wv.line_off[b], last = wv.line_off[b + 1];
  if (first == last) return;
  for (uint32_t i = threadIdx.x * 4; i < words + 32; i += SB_THREADS * 4) *reinterpret_cast<uint4 *>(s_cnt + i) = make_uint4(0, 0, 0, 0);
  __syncthreads();
  const uint32_t span = 2u * words, lo = z * span;
  const ull m0 = ((ull)b << mb) + lo;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // the rare part of an add of v that found the counter at `old`
  auto fixup = [&](uint32_t loc, uint32_t v, uint32_t old) {
    const uint32_t sh = (loc & 1u) * 16u;
    if (old < 0x8000u) {                               // this add crossed 2^15: move 2^15 to the table
      atomicSub(&s_cnt[loc >> 1], 0x8000u << sh);
      table_add(hist, (int64_t)(m0 + loc), 0x8000ull);
    } else if (old + v >= 0xC000u) {                   // the correction above is still pending and the counter keeps climbing: count elsewhere
      atomicSub(&s_cnt[loc >> 1], v << sh);
      table_add(hist, (int64_t)(m0 + loc), (ull)v);
    }
    __threadfence();                                   // ordered before this CTA's plain read-modify-write of the same entry below
  };
  auto add = [&](uint32_t loc, uint32_t v) {
    GTB_ASSERT(loc < span);
    const uint32_t sh = (loc & 1u) * 16u;
    const uint32_t old = (atomicAdd(&s_cnt[loc >> 1], v << sh) >> sh) & 0xFFFFu;
    if (old + v >= 0x8000u) fixup(loc, v, old);
  };
  // a warp takes SB_UNROLL blocks (512 bytes each) per trip, and the block ids of the next trip are fetched while this trip's
  // data is on its way: one memory latency per trip instead of two dependent ones, 128 KB in flight per SM
  constexpr uint32_t SB_UNROLL = SB_UNROLL_, SB_WARPS = SB_THREADS / 32;
  uint32_t blk0 = first + warp * SB_UNROLL;
  uint32_t e_next = blk0 + lane < last && lane < SB_UNROLL ? __ldg(wv.sorted_lines + blk0 + lane) : 0u;
  for (; blk0 < last; blk0 += SB_WARPS * SB_UNROLL) {
    const uint32_t e_mine = e_next;
    uint32_t fill[SB_UNROLL];
    uint4 d[SB_UNROLL];
#pragma unroll
    for (uint32_t r = 0; r < SB_UNROLL; r++) {
      const uint32_t entry = __shfl_sync(0xffffffffu, e_mine, r);
      fill[r] = blk0 + r < last ? (entry >> 25) + 1u : 0u;
      d[r] = make_uint4(0, 0, 0, 0);
      if (fill[r] > lane * 4u) d[r] = ldg_stream128(reinterpret_cast<const uint4 *>(wv.pool + (size_t)(entry & 0x01FFFFFFu) * WC_BLOCK_ELEMS) + lane);
    }
    {
      const uint32_t nb0 = blk0 + SB_WARPS * SB_UNROLL;
      e_next = nb0 + lane < last && lane < SB_UNROLL ? __ldg(wv.sorted_lines + nb0 + lane) : 0u;
    }
#pragma unroll
    for (uint32_t r = 0; r < SB_UNROLL; r++) {
      const uint32_t q4 = lane * 4u;
      if (fill[r] <= q4) continue;
      const uint32_t el[4] = {d[r].x, d[r].y, d[r].z, d[r].w};
      const uint32_t nv = min(fill[r] - q4, 4u);
      // this CTA's sub-range only (a bucket wider than one CTA's counters is read by several CTAs): keep the test lean
      const uint32_t l0 = el[0] - lo, l1 = el[1] - lo, l2 = el[2] - lo, l3 = el[3] - lo;
      const bool in0 = l0 < span, in1 = l1 < span && nv > 1u, in2 = l2 < span && nv > 2u, in3 = l3 < span && nv > 3u;
      if (!(in0 | in1 | in2 | in3)) continue;
      if (nv == 4u && el[0] == el[3] && el[0] == el[1] && el[0] == el[2]) { add(l0, 4u); continue; }   // position-sorted reads: one add
      // the four adds go out back to back; their results are only looked at afterwards (and almost never matter)
      const uint32_t s0 = (l0 & 1u) * 16u, s1 = (l1 & 1u) * 16u, s2 = (l2 & 1u) * 16u, s3 = (l3 & 1u) * 16u;
      // (an element outside the sub-range adds 0 to the lane's own dummy word: no branch, no reconvergence barrier around the atomic)
      const uint32_t dummy = words + lane;
      GTB_ASSERT((!in0 || l0 < span) && (!in1 || l1 < span) && (!in2 || l2 < span) && (!in3 || l3 < span) && dummy < words + 32u);
      const uint32_t o0 = atomicAdd(&s_cnt[in0 ? l0 >> 1 : dummy], in0 ? 1u << s0 : 0u) >> s0;
      const uint32_t o1 = atomicAdd(&s_cnt[in1 ? l1 >> 1 : dummy], in1 ? 1u << s1 : 0u) >> s1;
      const uint32_t o2 = atomicAdd(&s_cnt[in2 ? l2 >> 1 : dummy], in2 ? 1u << s2 : 0u) >> s2;
      const uint32_t o3 = atomicAdd(&s_cnt[in3 ? l3 >> 1 : dummy], in3 ? 1u << s3 : 0u) >> s3;
      if (((o0 + 1u) | (o1 + 1u) | (o2 + 1u) | (o3 + 1u)) & 0x8000u) {     // some counter was at 2^15 - 1 or beyond
        if (in0 && (o0 & 0xFFFFu) >= 0x7FFFu) fixup(l0, 1u, o0 & 0xFFFFu);
        if (in1 && (o1 & 0xFFFFu) >= 0x7FFFu) fixup(l1, 1u, o1 & 0xFFFFu);
        if (in2 && (o2 & 0xFFFFu) >= 0x7FFFu) fixup(l2, 1u, o2 & 0xFFFFu);
        if (in3 && (o3 & 0xFFFFu) >= 0x7FFFu) fixup(l3, 1u, o3 & 0xFFFFu);
      }
    }
  }
  __syncthreads();
  // two counters (one word) per thread and trip: 8-byte reads and writes of the table's uint32 plane, all loads of a trip issued
  // before its stores; an entry that passes 2^32 here sends its carry to the other plane
  constexpr uint32_t FL_UNROLL = 4;
  uint2 *hv = reinterpret_cast<uint2 *>(hist.lo + m0);                        // m0 is a multiple of 8
  for (uint32_t i0 = threadIdx.x; i0 < words; i0 += SB_THREADS * FL_UNROLL) {
    uint32_t w[FL_UNROLL];
    uint2 h[FL_UNROLL];
#pragma unroll
    for (uint32_t r = 0; r < FL_UNROLL; r++) {
      const uint32_t i = i0 + r * SB_THREADS;
      w[r] = i < words ? s_cnt[i] : 0u;
      if (w[r]) h[r] = hv[i];
    }
#pragma unroll
    for (uint32_t r = 0; r < FL_UNROLL; r++) {
      const uint32_t i = i0 + r * SB_THREADS;
      if (w[r]) {
        const uint2 n = make_uint2(h[r].x + (w[r] & 0xFFFFu), h[r].y + (w[r] >> 16));
        hv[i] = n;
        if (n.x < h[r].x) table_carry(hist, (int64_t)(m0 + 2u * i), 1u);
        if (n.y < h[r].y) table_carry(hist, (int64_t)(m0 + 2u * i + 1u), 1u);
      }
    }
  }
}

constexpr int WIN_THREADS = 256;
constexpr int WIN_ITEMS = 8;
constexpr int WIN_TILE = WIN_THREADS * WIN_ITEMS;
constexpr int WIN_MAX_COMBINE = 64;                    // micro-windows per window the staged form handles

__device__ __forceinline__ int find_slot(const int64_t *__restrict__ win_off, int n_slots, int64_t g) {
  int lo = 0, hi = n_slots;            // last slot with win_off[slot] <= g
  while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (win_off[mid] <= g) lo = mid; else hi = mid; }
  return lo;
}

__device__ __forceinline__ long long window_value(const SlotTable &t, const ScanTable &hist, bool wide, int slot, int64_t k0, int combine) {
  const int64_t base = t.hist_off[slot];
  if (t.spurious[slot]) {
    // Next() hands back current_v[1] of a slot that has no complete window (:5128-5140); with no
    // micro-window at all the reference reads past its allocation -- 0 in practice.
    return t.hist_off[slot + 1] - base >= 1 ? (long long)table_get(hist, wide, base) : 0;
  }
  ull sum = 0;
  for (int j = 0; j < combine; j++) sum += table_get(hist, wide, base + k0 + j);     // :5066-5073
  return (long long)sum;
}

// What the window pass leaves on the device: 10 bytes per qualifying window (the slot says chromosome and strand); gtb_scan_fetch
// widens a range of them to the interface's arrays.
struct WindowsOut {
  uint32_t *win;                   // 1-based window number inside its slot, :5109-5112
  uint32_t *val, *val_hi;          // value; the upper half only exists if some value needs it (flags[1])
  uint16_t *slot16; uint32_t *slot32;   // one of the two
};

// MODE 0: count qualifying windows per tile.  MODE 1: write them at tile_offset + local rank.
template <int MODE>
__global__ void __launch_bounds__(WIN_THREADS) scan_windows_kernel(SlotTable t, ScanTable hist, int64_t total_windows,
                                                                    int combine, long long min_reads, ull *__restrict__ tile_counts, WindowsOut out) {
  __shared__ int warp_counts[WIN_THREADS / 32];
  __shared__ ull tile_base;
  // the tile's micro-windows, staged with one pad entry per 8 so that thread t's run (8 t ...) starts in its own bank pair
  __shared__ ull s_h[(WIN_TILE + WIN_MAX_COMBINE) + (WIN_TILE + WIN_MAX_COMBINE) / 8 + 1];
  const bool wide = *reinterpret_cast<volatile const uint32_t *>(hist.flags) != 0u;
  const int64_t tile_first = (int64_t)blockIdx.x * WIN_TILE;
  const int64_t tile_last = min(total_windows, tile_first + WIN_TILE) - 1;
  const int64_t first = tile_first + (int64_t)threadIdx.x * WIN_ITEMS;
  long long val[WIN_ITEMS];
  int slot_of[WIN_ITEMS];
  unsigned keep = 0;
  const int slot_a = find_slot(t.win_off, t.n_slots, tile_first);
  if (tile_last < t.win_off[slot_a + 1] && !t.spurious[slot_a] && combine <= WIN_MAX_COMBINE) {
    // the whole tile lies in one (chromosome, strand) slot -- all but a few dozen tiles: coalesced load, sliding sums from shared memory
    const int n_w = (int)(tile_last - tile_first + 1), n_h = n_w + combine - 1;
    const int64_t src = t.hist_off[slot_a] + (tile_first - t.win_off[slot_a]);
    for (int i = threadIdx.x; i < n_h; i += WIN_THREADS) s_h[i + (i >> 3)] = table_get(hist, wide, src + i);
    __syncthreads();
    const int w0 = threadIdx.x * WIN_ITEMS;
    auto at = [&](int i) { return s_h[i + (i >> 3)]; };
    ull sum = 0;
    if (w0 < n_w)
      for (int j = 0; j < combine; j++) sum += at(w0 + j);                   // :5066-5073
#pragma unroll
    for (int i = 0; i < WIN_ITEMS; i++) {
      val[i] = 0; slot_of[i] = slot_a;
      if (w0 + i < n_w) {
        val[i] = (long long)sum;
        if (val[i] >= min_reads) keep |= 1u << i;
        if (w0 + i + 1 < n_w) sum += at(w0 + i + combine) - at(w0 + i);
      }
    }
  } else {
    int slot = first < total_windows ? find_slot(t.win_off, t.n_slots, first) : 0;
#pragma unroll
    for (int i = 0; i < WIN_ITEMS; i++) {
      const int64_t g = first + i;
      val[i] = 0; slot_of[i] = slot;
      if (g < total_windows) {
        while (g >= t.win_off[slot + 1]) slot++;
        slot_of[i] = slot;
        val[i] = window_value(t, hist, wide, slot, g - t.win_off[slot], combine);
        if (val[i] >= min_reads) keep |= 1u << i;
      }
    }
  }
  const int mine = __popc(keep);
  // block-wide exclusive scan of `mine`
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { int v = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += v; }
  if (lane == 31) warp_counts[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = lane < WIN_THREADS / 32 ? warp_counts[lane] : 0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int v = __shfl_up_sync(0xffffffffu, w, d); if (lane >= d) w += v; }
    if (lane < WIN_THREADS / 32) warp_counts[lane] = w;
  }
  __syncthreads();
  const int block_total = warp_counts[WIN_THREADS / 32 - 1];
  if (MODE == 0) {
    if (threadIdx.x == 0) tile_counts[blockIdx.x] = (ull)block_total;
    bool big = false;                                                   // a qualifying value beyond 32 bits: the emit pass then writes upper halves
#pragma unroll
    for (int i = 0; i < WIN_ITEMS; i++) big = big || ((keep >> i) & 1u && ((ull)val[i] >> 32) != 0ull);
    if (big) hist.flags[1] = 1u;
    return;
  }
  if (threadIdx.x == 0) tile_base = blockIdx.x == 0 ? 0ull : tile_counts[blockIdx.x - 1];   // inclusive-scanned counts
  __syncthreads();                                                     // also: everybody is done reading s_h
  // the kept windows go to shared memory in output order and leave as coalesced runs of the output arrays
  __shared__ uint16_t s_idx[WIN_TILE], s_slot[WIN_TILE];
  ull *s_val = s_h;
  int r = (warp > 0 ? warp_counts[warp - 1] : 0) + inc - mine;
#pragma unroll
  for (int i = 0; i < WIN_ITEMS; i++) {
    if (keep & (1u << i)) {
      s_val[r] = (ull)val[i];
      s_idx[r] = (uint16_t)(threadIdx.x * WIN_ITEMS + i);
      s_slot[r] = (uint16_t)(slot_of[i] - slot_a);                     // a tile spans at most WIN_TILE slots
      r++;
    }
  }
  __syncthreads();
  const int64_t o0 = (int64_t)tile_base;
  for (int j = threadIdx.x; j < block_total; j += WIN_THREADS) {
    const int s = slot_a + s_slot[j];
    if (out.slot16) out.slot16[o0 + j] = (uint16_t)s; else out.slot32[o0 + j] = (uint32_t)s;
    out.win[o0 + j] = (uint32_t)(tile_first + s_idx[j] - t.win_off[s] + 1);
    out.val[o0 + j] = (uint32_t)s_val[j];
    if (out.val_hi) out.val_hi[o0 + j] = (uint32_t)(s_val[j] >> 32);
  }
}

// a range of the compact windows -> the arrays of gtb_scan_fetch (device staging; null pointers are skipped)
__global__ void __launch_bounds__(256) scan_expand_kernel(SlotTable t, WindowsOut w, int64_t first, int64_t count, int32_t *__restrict__ o_chrom,
                                                           int8_t *__restrict__ o_strand, int64_t *__restrict__ o_win, int64_t *__restrict__ o_value) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
    const int64_t g = first + i;
    const int s = w.slot16 ? (int)w.slot16[g] : (int)w.slot32[g];
    if (o_chrom) o_chrom[i] = t.chrom[s];
    if (o_strand) o_strand[i] = t.strand[s];
    if (o_win) o_win[i] = (int64_t)w.win[g];
    if (o_value) o_value[i] = (int64_t)((ull)w.val[g] | (w.val_hi ? (ull)w.val_hi[g] << 32 : 0ull));
  }
}

// gtb_scan_reset: the carry plane is cleared only if it was ever written
__global__ void __launch_bounds__(256) scan_clear_carry_kernel(ScanTable hist, int64_t n) {
  if (*reinterpret_cast<volatile const uint32_t *>(hist.flags) == 0u) return;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) hist.hi[i] = 0u;
}


// ---- peak statistics (PeakFinder::Run, genomic_scans.cpp:298-368) ---------------------------------------------------------------
// Tail probabilities in double precision.  The reference calls GSL; these are the textbook forms of the same functions:
// P(X > k), X ~ Bin(n, p) = I_p(k + 1, n - k) (regularised incomplete beta, continued fraction by the modified Lentz method);
// P(X > k), X ~ Poisson(mu) = P(k + 1, mu) (regularised lower incomplete gamma: series below a + 1, continued fraction above).
__device__ double pk_beta_cf(double a, double b, double x) {
  const double TINY = 1e-300, EPS = 1e-15;
  const double qab = a + b, qap = a + 1.0, qam = a - 1.0;
  double c = 1.0, d = 1.0 - qab * x / qap;
  if (fabs(d) < TINY) d = TINY;
  d = 1.0 / d;
  double h = d;
  for (int m = 1; m <= 20000; m++) {
    const double m2 = 2.0 * m;
    double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
    d = 1.0 + aa * d; if (fabs(d) < TINY) d = TINY;
    c = 1.0 + aa / c; if (fabs(c) < TINY) c = TINY;
    d = 1.0 / d; h *= d * c;
    aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
    d = 1.0 + aa * d; if (fabs(d) < TINY) d = TINY;
    c = 1.0 + aa / c; if (fabs(c) < TINY) c = TINY;
    d = 1.0 / d;
    const double del = d * c;
    h *= del;
    if (fabs(del - 1.0) < EPS) break;
  }
  return h;
}
__device__ double pk_beta_inc(double a, double b, double x) {
  if (x <= 0.0) return 0.0;
  if (x >= 1.0) return 1.0;
  const double ln_bt = lgamma(a + b) - lgamma(a) - lgamma(b) + a * log(x) + b * log1p(-x);
  if (x < (a + 1.0) / (a + b + 2.0)) return exp(ln_bt) * pk_beta_cf(a, b, x) / a;
  return 1.0 - exp(ln_bt) * pk_beta_cf(b, a, 1.0 - x) / b;
}
// gsl_cdf_binomial_Q(unsigned k, double p, unsigned n)
__device__ double pk_binomial_Q(long long k, double p, long long n) {
  const unsigned uk = (unsigned)k, un = (unsigned)n;
  if (uk >= un) return 0.0;
  return pk_beta_inc((double)uk + 1.0, (double)un - (double)uk, p);
}
// gsl_cdf_poisson_Q(unsigned k, double mu) = gsl_cdf_gamma_P(mu, k + 1, 1)
__device__ double pk_poisson_Q(long long k, double mu) {
  const double a = (double)(unsigned)k + 1.0, x = mu;
  if (x <= 0.0) return 0.0;
  const double ln_pre = -x + a * log(x) - lgamma(a);
  if (x < a + 1.0) {
    double ap = a, sum = 1.0 / a, del = sum;
    for (int n = 0; n < 100000; n++) { ap += 1.0; del *= x / ap; sum += del; if (fabs(del) < fabs(sum) * 1e-16) break; }
    return sum * exp(ln_pre);
  }
  const double TINY = 1e-300;
  double b = x + 1.0 - a, c = 1.0 / TINY, d = 1.0 / b, h = d;
  for (int i = 1; i < 100000; i++) {
    const double an = -(double)i * ((double)i - a);
    b += 2.0;
    d = an * d + b; if (fabs(d) < TINY) d = TINY;
    c = b + an / c; if (fabs(c) < TINY) c = TINY;
    d = 1.0 / d;
    const double del = d * c;
    h *= del;
    if (fabs(del - 1.0) < 1e-15) break;
  }
  return 1.0 - exp(ln_pre) * h;
}
__device__ double pk_ugaussian_Q(double x) { return 0.5 * erfc(x * 0.70710678118654752440); }

struct PeaksView {
  gtb_peaks_params prm;
  ScanTable control;               // lo == nullptr: no control scanner
  long long win_size;
};

// a Poisson variate of mean mu for window g (inversion on a hashed uniform; the reference: gsl_ran_poisson on a time-seeded generator)
__device__ long long pk_poisson_variate(uint64_t seed, int64_t g, double mu) {
  uint64_t z = seed + (uint64_t)g * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
  const double u = (double)(z >> 11) * (1.0 / 9007199254740992.0);
  double p = exp(-mu), s = p;
  long long k = 0;
  while (u > s && k < 100000) { k++; p *= mu / (double)k; s += p; }
  return k;
}

// the statistics of one window (genomic_scans.cpp:303-354); false: the window is dropped
__device__ bool pk_window(const PeaksView &pv, long long v1, long long v2, double &pval1, double &pval2) {
  const gtb_peaks_params &P = pv.prm;
  const long long v0 = pv.win_size;                                      // (no uniqueness scanner)
  v1 = min(v1, v0); v2 = min(v2, v0);
  if (P.norm) {
    const double ratio = P.p_signal / P.p_control;
    if (ratio < 1.0) v2 = (long long)floor((double)(float)v2 * ratio); else v1 = (long long)floor((double)(float)v1 / ratio);
  }
  if (v1 < P.min_reads) return false;
  if (P.method == GTB_PEAKS_POISSON) {
    pval1 = pk_poisson_Q(v1 + 5, (double)(v2 + 5));
    pval2 = pk_poisson_Q(v2 + 5, (double)(v1 + 5));
  } else if (!P.compare) {
    pval1 = pk_binomial_Q(v1, P.p_signal, v0 + 1);
    pval2 = pk_binomial_Q(v2, P.p_control, v0 + 1);
  } else if (P.method == GTB_PEAKS_BINOMIAL) {
    const float pp_control = (float)(((double)(float)v2 + 1.0) / ((double)v0 + 1.0));
    pval1 = pk_binomial_Q(v1, fmax((double)pp_control, P.p_signal), v0 + 1);
    const float pp_signal = (float)(((double)(float)v1 + 1.0) / ((double)v0 + 1.0));
    pval2 = pk_binomial_Q(v2, fmax((double)pp_signal, P.p_control), v0 + 1);
  } else if (P.method == GTB_PEAKS_BINOMIAL2) {
    const double pp_control = (double)(v2 + 1) / (double)P.n_control_reads, pp_signal = (double)(v1 + 1) / (double)P.n_signal_reads;
    pval1 = pk_binomial_Q(v1 + 1, pp_control, P.n_signal_reads);
    pval2 = pk_binomial_Q(v2 + 1, pp_signal, P.n_control_reads);
  } else if (P.method == GTB_PEAKS_CBINOMIAL) {
    pval1 = pk_binomial_Q(v1 + 1, 0.5, v1 + v2 + 2);
    pval2 = pk_binomial_Q(v2 + 1, 0.5, v1 + v2 + 2);
  } else {
    const double pp_control = (double)(v2 + 1) / (double)P.n_control_reads, pp_signal = (double)(v1 + 1) / (double)P.n_signal_reads;
    const double ns = (double)P.n_signal_reads, nc = (double)P.n_control_reads;
    pval1 = pk_ugaussian_Q(((double)(v1 + 1) - ns * pp_control) / sqrt(ns * pp_control));
    pval2 = pk_ugaussian_Q(((double)(v2 + 1) - nc * pp_signal) / sqrt(nc * pp_signal));
  }
  return pval1 <= P.pval_cutoff;
}

constexpr int PK_THREADS = 256, PK_ITEMS = 4, PK_TILE = PK_THREADS * PK_ITEMS;
// MODE 0: kept windows per tile.  MODE 1: the kept windows, at the tile's offset (tile_counts inclusive-scanned), in scan order.
template <int MODE>
__global__ void __launch_bounds__(PK_THREADS) scan_peaks_kernel(SlotTable t, ScanTable sig, PeaksView pv, int64_t total_windows, int combine,
                                                                 ull *__restrict__ tile_counts, WindowsOut out, double *__restrict__ o_p1, double *__restrict__ o_p2) {
  __shared__ int warp_counts[PK_THREADS / 32];
  const bool wide_s = *reinterpret_cast<volatile const uint32_t *>(sig.flags) != 0u;
  const bool has_c = pv.control.lo != nullptr;
  const bool wide_c = has_c && *reinterpret_cast<volatile const uint32_t *>(pv.control.flags) != 0u;
  const int64_t first = (int64_t)blockIdx.x * PK_TILE + (int64_t)threadIdx.x * PK_ITEMS;
  int slot = first < total_windows ? find_slot(t.win_off, t.n_slots, first) : 0;
  double p1[PK_ITEMS], p2[PK_ITEMS];
  int slot_of[PK_ITEMS];
  unsigned keep = 0;
  long long s1 = 0, s2 = 0;
  bool sliding = false;                                                  // s1 / s2 hold the previous window's sums of the same slot
#pragma unroll
  for (int i = 0; i < PK_ITEMS; i++) {
    const int64_t g = first + i;
    slot_of[i] = slot;
    if (g >= total_windows) continue;
    if (g >= t.win_off[slot + 1]) { while (g >= t.win_off[slot + 1]) slot++; sliding = false; }
    slot_of[i] = slot;
    const int64_t k0 = g - t.win_off[slot];
    if (sliding && !t.spurious[slot]) {                                 // :5066-5073, one step to the right
      const int64_t base = t.hist_off[slot] + k0;
      s1 += (long long)table_get(sig, wide_s, base + combine - 1) - (long long)table_get(sig, wide_s, base - 1);
      if (has_c) s2 += (long long)table_get(pv.control, wide_c, base + combine - 1) - (long long)table_get(pv.control, wide_c, base - 1);
    } else {
      s1 = window_value(t, sig, wide_s, slot, k0, combine);
      if (has_c) s2 = window_value(t, pv.control, wide_c, slot, k0, combine);
      sliding = true;
    }
    const long long v2 = has_c ? s2 : pk_poisson_variate(pv.prm.seed, g, (double)pv.win_size * pv.prm.p_signal);
    if (pk_window(pv, s1, v2, p1[i], p2[i])) keep |= 1u << i;
  }
  const int mine = __popc(keep);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { int v = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += v; }
  if (lane == 31) warp_counts[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = lane < PK_THREADS / 32 ? warp_counts[lane] : 0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int v = __shfl_up_sync(0xffffffffu, w, d); if (lane >= d) w += v; }
    if (lane < PK_THREADS / 32) warp_counts[lane] = w;
  }
  __syncthreads();
  if (MODE == 0) {
    if (threadIdx.x == 0) tile_counts[blockIdx.x] = (ull)warp_counts[PK_THREADS / 32 - 1];
    return;
  }
  int64_t r = (int64_t)(blockIdx.x == 0 ? 0ull : tile_counts[blockIdx.x - 1]) + (warp > 0 ? warp_counts[warp - 1] : 0) + inc - mine;
#pragma unroll
  for (int i = 0; i < PK_ITEMS; i++)
    if (keep & (1u << i)) {
      const int s = slot_of[i];
      if (out.slot16) out.slot16[r] = (uint16_t)s; else out.slot32[r] = (uint32_t)s;
      out.win[r] = (uint32_t)(first + i - t.win_off[s] + 1);
      o_p1[r] = p1[i]; o_p2[r] = p2[i];
      r++;
    }
}

__global__ void __launch_bounds__(256) scan_peaks_expand_kernel(SlotTable t, WindowsOut w, int64_t first, int64_t count, int32_t *__restrict__ o_chrom,
                                                                 int8_t *__restrict__ o_strand, int64_t *__restrict__ o_win) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
    const int64_t g = first + i;
    const int s = w.slot16 ? (int)w.slot16[g] : (int)w.slot32[g];
    if (o_chrom) o_chrom[i] = t.chrom[s];
    if (o_strand) o_strand[i] = t.strand[s];
    if (o_win) o_win[i] = (int64_t)w.win[g];
  }
}

}  // namespace

struct gtb_scan {
  gtb_ctx *ctx = nullptr;
  gtb_scan_params prm{};
  int32_t n_chrom = 0, n_slots = 0;
  int combine = 1;
  int64_t total_micro = 0, total_windows = 0, n_out = 0;
  dbuf<int32_t> d_slot_of_chrom, d_spurious, d_slot_chrom;
  dbuf<int8_t> d_slot_strand;
  dbuf<int64_t> d_hist_off, d_win_off;
  dbuf<uint32_t> d_lo, d_hi, d_flags;                  // the table (ScanTable)
  dbuf<ull> d_tile_counts, d_scan_scratch;
  dbuf<uint32_t> c_win, c_val, c_val_hi, c_slot32; dbuf<uint16_t> c_slot16;   // the qualifying windows, compact (WindowsOut)
  bool wide_values = false;                            // c_val_hi is in use
  dbuf<double> pk_p1, pk_p2;                           // gtb_scan_peaks: the kept windows' tail probabilities (their slot and number in c_slot*, c_win)
  int64_t n_peaks = 0;
  dbuf<int32_t> o_chrom; dbuf<int8_t> o_strand; dbuf<int64_t> o_win, o_value;   // staging of gtb_scan_fetch
  uint32_t *h_flags = nullptr;                         // pinned: [0..1] the table's flags, [2..3] the window total
  ScanTable table() const { return ScanTable{d_lo.p, d_hi.p, d_flags.p}; }
  WindowsOut windows() const {
    return WindowsOut{c_win.p, c_val.p, wide_values ? c_val_hi.p : nullptr, n_slots <= 65536 ? c_slot16.p : nullptr, n_slots <= 65536 ? nullptr : c_slot32.p};
  }
  // bucketed histogram (unweighted batches of at least bucket_min intervals)
  bool bucket_ok = false;
  int64_t bucket_min = (int64_t)2 << 20;
  uint32_t mb = 0, n_buckets = 0, n_sub = 1, sb_words = 0;   // a bucket = 2^mb micro-windows, counted by n_sub CTAs of 2 * sb_words counters
  uint32_t magic = 0; int shift = 0;
  dbuf<uint4> d_front_tab;
  WcBuffers wc;
  struct stage {
    dbuf<int32_t> chrom, start, stop, weight; dbuf<int8_t> strand; dbuf<int64_t> off;
    cudaEvent_t copied = nullptr, consumed = nullptr; bool in_flight = false;
  } stages[2];
  int next_stage = 0;
};

template <typename T>
static int upload_vec(gtb_ctx *ctx, dbuf<T> &d, const std::vector<T> &h) {
  GTB_TRY(d.reserve(ctx, h.size() ? h.size() : 1));
  if (h.size()) GTB_CUDA_OK(ctx, cudaMemcpyAsync(d.p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  return GTB_OK;
}

extern "C" int gtb_scan_reset(gtb_scan *sc) {
  if (!sc) return GTB_ERR_ARG;
  gtb_ctx *ctx = sc->ctx;
  const int64_t n = std::max<int64_t>(sc->total_micro, 1);
  GTB_CUDA_OK(ctx, cudaMemsetAsync(sc->d_lo.p, 0, sizeof(uint32_t) * (size_t)n, ctx->stream));
  GTB_LAUNCH(ctx, "scan_clear_carry", scan_clear_carry_kernel, (unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 8), 256, 0, sc->table(), n);
  GTB_CUDA_OK(ctx, cudaMemsetAsync(sc->d_flags.p, 0, 2 * sizeof(uint32_t), ctx->stream));
  sc->n_out = 0;
  return gtb_check_launch(ctx);
}

extern "C" int gtb_scan_create(gtb_ctx *ctx, int32_t n_chrom, const int64_t *bound, const gtb_scan_params *params, gtb_scan **out) {
  if (!ctx || !bound || !params || !out || n_chrom < 0) return GTB_ERR_ARG;
  *out = nullptr;
  if (params->win_step <= 0 || params->win_size <= 0) return gtb_fail(ctx, GTB_ERR_ARG, "window size/step must be positive");
  if (params->win_size % params->win_step != 0)                      // :4845
    return gtb_fail(ctx, GTB_ERR_WINDOW, "window size must be a multiple of window step");
  if (params->op != '1' && params->op != 'c')                        // :5046
    return gtb_fail(ctx, GTB_ERR_UNSUPPORTED, "preprocess operator not supported (use '1' or 'c')");
  if (params->win_size / params->win_step > (1 << 20)) return gtb_fail(ctx, GTB_ERR_UNSUPPORTED, "window/step ratio too large");
  GTB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  gtb_scan *sc = new gtb_scan();
  sc->ctx = ctx; sc->prm = *params; sc->n_chrom = n_chrom;
  sc->combine = (int)(params->win_size / params->win_step);
  const int n_strands = params->ignore_strand ? 1 : 2;
  std::vector<int32_t> slot_of_chrom((size_t)std::max(n_chrom, 1), -1), spurious, slot_chrom;
  std::vector<int8_t> slot_strand;
  std::vector<int64_t> hist_off(1, 0), win_off(1, 0);
  for (int32_t c = 0; c < n_chrom; c++) {
    if (bound[c] < 0) continue;
    const int64_t n_micro = bound[c] / params->win_step;             // :5026
    slot_of_chrom[c] = (int32_t)slot_chrom.size();
    for (int z = 0; z < n_strands; z++) {
      const bool first_slot = slot_chrom.empty();
      int64_t n_win = n_micro < sc->combine ? 0 : n_micro - sc->combine + 1;   // :5061-5064
      int sp = 0;
      if (n_win == 0 && !params->emulate_sorted && !first_slot) { n_win = 1; sp = 1; }
      slot_chrom.push_back(c); slot_strand.push_back(z ? '-' : '+'); spurious.push_back(sp);
      hist_off.push_back(hist_off.back() + n_micro);
      win_off.push_back(win_off.back() + n_win);
    }
  }
  sc->n_slots = (int32_t)slot_chrom.size();
  sc->total_micro = hist_off.back(); sc->total_windows = win_off.back();
  int rc = upload_vec(ctx, sc->d_slot_of_chrom, slot_of_chrom);
  // bucketed histogram: <= WC_MAX_BUCKETS buckets of 2^mb micro-windows, at least 256 of them (fewer overfill the partition's
  // rings); a bucket wider than the 16-bit counters one SM's shared memory holds is walked by several CTAs, at most 16
  {
    uint32_t mb = 10;
    while (mb < 24 && ((sc->total_micro + (((int64_t)1 << mb) - 1)) >> mb) > WC_MAX_BUCKETS) mb++;
    const int64_t nb = (sc->total_micro + (((int64_t)1 << mb) - 1)) >> mb;
    const size_t tab_entries = (size_t)2 * std::max(n_chrom, 1) + 2;
    const size_t smem = wc_smem_bytes((uint32_t)std::max<int64_t>(nb, 1), tab_entries);
    const uint32_t max_counters = (uint32_t)(((ctx->smem_optin - 32 * 4 - 1024) / 2) & ~(size_t)7);     // 16-bit counters per CTA
    const uint32_t n_sub = (uint32_t)((((uint64_t)1 << mb) + max_counters - 1) / max_counters);
    if (n_sub <= 16 && nb >= 256 && params->win_step < ((int64_t)1 << 31) && smem <= ctx->smem_optin && !getenv("GTB_SCAN_DIRECT")) {
      sc->bucket_ok = true; sc->mb = mb; sc->n_buckets = (uint32_t)nb;
      sc->n_sub = n_sub;
      sc->sb_words = (uint32_t)((((((uint64_t)1 << mb) + n_sub - 1) / n_sub + 7) & ~(uint64_t)7) / 2);
      const uint64_t d = (uint64_t)params->win_step;
      int l = 0;
      while (((uint64_t)1 << l) < d) l++;
      sc->shift = 31 + l;
      sc->magic = (uint32_t)((((uint64_t)1 << sc->shift) + d - 1) / d);      // ceil(2^(31 + l) / d): exact quotients for dividends < 2^31
      std::vector<uint4> tab(tab_entries, make_uint4(0, 0, 0, 0));
      for (int32_t c = 0; c < n_chrom; c++) {
        const int32_t s0 = slot_of_chrom[c];
        if (s0 < 0) continue;
        for (int z = 0; z < 2; z++) {
          const int32_t slot = s0 + (z && n_strands == 2 ? 1 : 0);
          const int64_t n_micro = hist_off[slot + 1] - hist_off[slot];
          const int64_t last_pos = std::min<int64_t>(n_micro * params->win_step, 0x7FFFFFFF);
          tab[2 * c + z] = make_uint4((uint32_t)last_pos, (uint32_t)hist_off[slot], 0, 0);
        }
      }
      if (rc == GTB_OK) rc = upload_vec(ctx, sc->d_front_tab, tab);
      if (const char *env = getenv("GTB_SCAN_BUCKET_MIN")) sc->bucket_min = std::max<long long>(1, atoll(env));
    }
  }
  if (rc == GTB_OK) rc = upload_vec(ctx, sc->d_spurious, spurious);
  if (rc == GTB_OK) rc = upload_vec(ctx, sc->d_slot_chrom, slot_chrom);
  if (rc == GTB_OK) rc = upload_vec(ctx, sc->d_slot_strand, slot_strand);
  if (rc == GTB_OK) rc = upload_vec(ctx, sc->d_hist_off, hist_off);
  if (rc == GTB_OK) rc = upload_vec(ctx, sc->d_win_off, win_off);
  if (rc == GTB_OK) rc = sc->d_lo.reserve(ctx, (size_t)std::max<int64_t>(sc->total_micro, 1) + 2);   // + 2: the bucketed flush moves entries in pairs
  if (rc == GTB_OK) rc = sc->d_hi.reserve(ctx, (size_t)std::max<int64_t>(sc->total_micro, 1) + 2);
  if (rc == GTB_OK) rc = sc->d_flags.reserve(ctx, 2);
  if (rc == GTB_OK && cudaHostAlloc((void **)&sc->h_flags, 4 * sizeof(uint32_t), cudaHostAllocDefault) != cudaSuccess) rc = GTB_ERR_CUDA;
  if (rc == GTB_OK && cudaMemsetAsync(sc->d_hi.p, 0, sizeof(uint32_t) * ((size_t)std::max<int64_t>(sc->total_micro, 1) + 2), ctx->stream) != cudaSuccess) rc = GTB_ERR_CUDA;
  if (rc == GTB_OK && cudaMemsetAsync(sc->d_flags.p, 0, 2 * sizeof(uint32_t), ctx->stream) != cudaSuccess) rc = GTB_ERR_CUDA;
  if (rc == GTB_OK) rc = gtb_scan_reset(sc);
  if (rc == GTB_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = GTB_ERR_CUDA;
  for (auto &st : sc->stages) {
    cudaEventCreateWithFlags(&st.copied, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&st.consumed, cudaEventDisableTiming);
  }
  if (rc != GTB_OK) { gtb_scan_destroy(sc); return rc; }
  *out = sc;
  return GTB_OK;
}

extern "C" void gtb_scan_destroy(gtb_scan *sc) {
  if (!sc) return;
  cudaSetDevice(sc->ctx->device);
  gtb_ctx_synchronize(sc->ctx);
  sc->d_slot_of_chrom.release(); sc->d_spurious.release(); sc->d_slot_chrom.release(); sc->d_slot_strand.release();
  sc->d_hist_off.release(); sc->d_win_off.release(); sc->d_lo.release(); sc->d_hi.release(); sc->d_flags.release();
  sc->d_tile_counts.release(); sc->d_scan_scratch.release();
  sc->c_win.release(); sc->c_val.release(); sc->c_val_hi.release(); sc->c_slot16.release(); sc->c_slot32.release();
  sc->pk_p1.release(); sc->pk_p2.release();
  sc->o_chrom.release(); sc->o_strand.release(); sc->o_win.release(); sc->o_value.release();
  if (sc->h_flags) cudaFreeHost(sc->h_flags);
  sc->d_front_tab.release(); sc->wc.release();
  for (auto &st : sc->stages) {
    st.chrom.release(); st.start.release(); st.stop.release(); st.weight.release(); st.strand.release(); st.off.release();
    if (st.copied) cudaEventDestroy(st.copied);
    if (st.consumed) cudaEventDestroy(st.consumed);
  }
  delete sc;
}

static int scan_accumulate_device(gtb_scan *sc, const ReadView &q) {
  gtb_ctx *ctx = sc->ctx;
  if (q.n_regions <= 0 || sc->n_slots == 0) return GTB_OK;
  if (sc->bucket_ok && !q.weight && q.n_intervals >= sc->bucket_min && q.n_intervals < ((int64_t)1 << 31)) {
    // every interval of every region counts once (:5039): without weights the region structure does not matter
    WcView wv;
    unsigned gridw = 0;
    if (sc->wc.plan(ctx, q.n_intervals, sc->n_buckets, &wv, &gridw) == GTB_OK) {
      const ScanFront front{sc->d_front_tab.p, (uint32_t)sc->n_chrom, sc->magic, sc->shift, sc->mb, sc->prm.op == 'c' ? 1 : 0,
                            sc->prm.ignore_strand ? 1 : 0, sc->table()};
      const WcQueries wq{q.n_intervals, q.chrom, q.start, q.stop, q.strand, 0};
      GTB_TRY(wc_partition_launch(ctx, "scan_partition", wq, front, wv, gridw, wc_smem_bytes(sc->n_buckets, (size_t)2 * std::max(sc->n_chrom, 1) + 2)));
      const size_t smem = ((size_t)sc->sb_words + 32) * 4;
      static const int variant = getenv("GTB_SB_VARIANT") ? atoi(getenv("GTB_SB_VARIANT")) : 0;    // tuning knob
#define GTB_SB_LAUNCH(T, U)                                                                                                          \
  do {                                                                                                                               \
    GTB_CUDA_OK(ctx, cudaFuncSetAttribute(scan_bucket_hist_kernel<T, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  \
    GTB_LAUNCH(ctx, "scan_bucket_hist", (scan_bucket_hist_kernel<T, U>), sc->n_buckets * sc->n_sub, T, smem, wv, sc->mb, sc->n_sub,  \
               sc->sb_words, sc->table());                                                                                           \
  } while (0)
      if (variant == 1) GTB_SB_LAUNCH(1024, 2);
      else if (variant == 2) GTB_SB_LAUNCH(512, 8);
      else GTB_SB_LAUNCH(1024, 4);
#undef GTB_SB_LAUNCH
      return gtb_check_launch(ctx);
    }
  }
  const unsigned grid = gtb_grid_for(q.n_regions, 256, (int64_t)ctx->sm_count * 8);
  GTB_LAUNCH(ctx, "scan_histogram", scan_histogram_kernel, grid, 256, 0, q, sc->n_chrom, sc->d_slot_of_chrom.p, sc->d_hist_off.p,
             (long long)sc->prm.win_step, (int)sc->prm.op, (int)sc->prm.ignore_strand, sc->table());
  return gtb_check_launch(ctx);
}

extern "C" int gtb_scan_add_reads(gtb_scan *sc, const gtb_set *reads, unsigned mem) {
  if (!sc || !reads) return GTB_ERR_ARG;
  gtb_ctx *ctx = sc->ctx;
  if (reads->n_regions < 0 || reads->n_intervals < 0) return gtb_fail(ctx, GTB_ERR_ARG, "negative sizes");
  if (reads->n_regions == 0) return GTB_OK;
  if (!reads->region_offset && reads->n_regions != reads->n_intervals)
    return gtb_fail(ctx, GTB_ERR_ARG, "region_offset is NULL but n_regions != n_intervals");
  if (!reads->chrom || !reads->start || !reads->stop || !reads->strand) return gtb_fail(ctx, GTB_ERR_ARG, "null interval arrays");
  GTB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  bool multi = false;
  if (reads->region_offset && !(mem & GTB_MEM_DEVICE)) {                 // a CSR is judged by its offsets, not by its totals
    if (reads->region_offset[0] != 0 || reads->region_offset[reads->n_regions] != reads->n_intervals)
      return gtb_fail(ctx, GTB_ERR_ARG, "region_offset must run from 0 to n_intervals");
    for (int64_t k = 0; k < reads->n_regions; k++) {
      const int64_t d = reads->region_offset[k + 1] - reads->region_offset[k];
      if (d < 1) return gtb_fail(ctx, GTB_ERR_ARG, "region_offset must be increasing: every region has at least one interval");
      multi = multi || d != 1;
    }
  } else if (reads->region_offset) {
    multi = true;                                                        // device-resident offsets are taken as they are
  }
  if (mem & GTB_MEM_DEVICE) {
    ReadView q{reads->n_regions, reads->n_intervals, reads->chrom, reads->start, reads->stop, reads->strand, reads->weight,
               multi ? reads->region_offset : nullptr, 0};
    return scan_accumulate_device(sc, q);
  }
  const int64_t CHUNK = (int64_t)8 << 20;
  for (int64_t r0 = 0; r0 < reads->n_regions; r0 += CHUNK) {
    const int64_t r1 = std::min(reads->n_regions, r0 + CHUNK);
    const int64_t i0 = multi ? reads->region_offset[r0] : r0, i1 = multi ? reads->region_offset[r1] : r1;
    const size_t nr = (size_t)(r1 - r0), ni = (size_t)(i1 - i0);
    gtb_scan::stage &st = sc->stages[sc->next_stage];
    sc->next_stage ^= 1;
    cudaStream_t cs = ctx->copy_stream;
    if (st.in_flight) GTB_CUDA_OK(ctx, cudaStreamWaitEvent(cs, st.consumed, 0));
    GTB_TRY(st.chrom.reserve(ctx, ni)); GTB_TRY(st.start.reserve(ctx, ni)); GTB_TRY(st.stop.reserve(ctx, ni)); GTB_TRY(st.strand.reserve(ctx, ni));
    GTB_CUDA_OK(ctx, cudaMemcpyAsync(st.chrom.p, reads->chrom + i0, ni * 4, cudaMemcpyHostToDevice, cs));
    GTB_CUDA_OK(ctx, cudaMemcpyAsync(st.start.p, reads->start + i0, ni * 4, cudaMemcpyHostToDevice, cs));
    GTB_CUDA_OK(ctx, cudaMemcpyAsync(st.stop.p, reads->stop + i0, ni * 4, cudaMemcpyHostToDevice, cs));
    GTB_CUDA_OK(ctx, cudaMemcpyAsync(st.strand.p, reads->strand + i0, ni, cudaMemcpyHostToDevice, cs));
    if (reads->weight) {
      GTB_TRY(st.weight.reserve(ctx, nr));
      GTB_CUDA_OK(ctx, cudaMemcpyAsync(st.weight.p, reads->weight + r0, nr * 4, cudaMemcpyHostToDevice, cs));
    }
    if (multi) {
      GTB_TRY(st.off.reserve(ctx, nr + 1));
      GTB_CUDA_OK(ctx, cudaMemcpyAsync(st.off.p, reads->region_offset + r0, (nr + 1) * 8, cudaMemcpyHostToDevice, cs));
    }
    GTB_CUDA_OK(ctx, cudaEventRecord(st.copied, cs));
    GTB_CUDA_OK(ctx, cudaStreamWaitEvent(ctx->stream, st.copied, 0));
    ReadView q{(int64_t)nr, (int64_t)ni, st.chrom.p, st.start.p, st.stop.p, st.strand.p, reads->weight ? st.weight.p : nullptr,
               multi ? st.off.p : nullptr, i0};
    GTB_TRY(scan_accumulate_device(sc, q));
    GTB_CUDA_OK(ctx, cudaEventRecord(st.consumed, ctx->stream));
    st.in_flight = true;
  }
  // "copied inside the call" (gtb200.h): the caller may reuse its arrays on return, so the last chunk's copies have to have
  // left them (pinned memory is read by the DMA engine after cudaMemcpyAsync returns); the kernels stay asynchronous
  GTB_CUDA_OK(ctx, cudaEventSynchronize(sc->stages[sc->next_stage ^ 1].copied));
  return GTB_OK;
}

extern "C" int gtb_scan_finish(gtb_scan *sc, int64_t *n_windows) {
  if (!sc || !n_windows) return GTB_ERR_ARG;
  gtb_ctx *ctx = sc->ctx;
  *n_windows = 0; sc->n_out = 0;
  if (sc->total_windows == 0) return GTB_OK;
  GTB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  SlotTable t{sc->n_slots, sc->d_hist_off.p, sc->d_win_off.p, sc->d_spurious.p, sc->d_slot_chrom.p, sc->d_slot_strand.p};
  const int64_t n_tiles = (sc->total_windows + WIN_TILE - 1) / WIN_TILE;
  GTB_TRY(sc->d_tile_counts.reserve(ctx, (size_t)n_tiles));
  GTB_CUDA_OK(ctx, cudaMemsetAsync(sc->d_flags.p + 1, 0, sizeof(uint32_t), ctx->stream));      // "wide values" is about this pass
  GTB_LAUNCH(ctx, "scan_windows_count", scan_windows_kernel<0>, (unsigned)n_tiles, WIN_THREADS, 0, t, sc->table(), sc->total_windows,
             sc->combine, (long long)sc->prm.min_reads, sc->d_tile_counts.p, WindowsOut{nullptr, nullptr, nullptr, nullptr, nullptr});
  GTB_TRY(gtb_check_launch(ctx));
  GTB_TRY(gtb_inclusive_scan_u64(ctx, sc->d_tile_counts.p, n_tiles, sc->d_scan_scratch));
  GTB_CUDA_OK(ctx, cudaMemcpyAsync(sc->h_flags + 2, sc->d_tile_counts.p + (n_tiles - 1), sizeof(ull), cudaMemcpyDeviceToHost, ctx->stream));
  GTB_CUDA_OK(ctx, cudaMemcpyAsync(sc->h_flags, sc->d_flags.p, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  GTB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  ull total;
  memcpy(&total, sc->h_flags + 2, sizeof total);
  sc->n_out = (int64_t)total;
  sc->wide_values = sc->h_flags[1] != 0u;
  *n_windows = sc->n_out;
  if (sc->n_out == 0) return GTB_OK;
  GTB_TRY(sc->c_win.reserve(ctx, (size_t)sc->n_out)); GTB_TRY(sc->c_val.reserve(ctx, (size_t)sc->n_out));
  if (sc->wide_values) GTB_TRY(sc->c_val_hi.reserve(ctx, (size_t)sc->n_out));
  if (sc->n_slots <= 65536) GTB_TRY(sc->c_slot16.reserve(ctx, (size_t)sc->n_out)); else GTB_TRY(sc->c_slot32.reserve(ctx, (size_t)sc->n_out));
  GTB_LAUNCH(ctx, "scan_windows_emit", scan_windows_kernel<1>, (unsigned)n_tiles, WIN_THREADS, 0, t, sc->table(), sc->total_windows,
             sc->combine, (long long)sc->prm.min_reads, sc->d_tile_counts.p, sc->windows());
  GTB_TRY(gtb_check_launch(ctx));
  GTB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  return GTB_OK;
}

extern "C" int gtb_scan_fetch(gtb_scan *sc, int64_t first, int64_t count, int32_t *chrom, int8_t *strand, int64_t *win, int64_t *value) {
  if (!sc || first < 0 || count < 0 || first + count > sc->n_out) return GTB_ERR_ARG;
  gtb_ctx *ctx = sc->ctx;
  if (count == 0) return GTB_OK;
  GTB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  SlotTable t{sc->n_slots, sc->d_hist_off.p, sc->d_win_off.p, sc->d_spurious.p, sc->d_slot_chrom.p, sc->d_slot_strand.p};
  // the windows are widened piece by piece into a bounded device staging area and copied out from there
  const int64_t PIECE = (int64_t)4 << 20;
  const size_t np = (size_t)std::min(count, PIECE);
  if (chrom) GTB_TRY(sc->o_chrom.reserve(ctx, np));
  if (strand) GTB_TRY(sc->o_strand.reserve(ctx, np));
  if (win) GTB_TRY(sc->o_win.reserve(ctx, np));
  if (value) GTB_TRY(sc->o_value.reserve(ctx, np));
  for (int64_t p0 = 0; p0 < count; p0 += PIECE) {
    const int64_t n = std::min(PIECE, count - p0);
    GTB_LAUNCH(ctx, "scan_expand", scan_expand_kernel, (unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 8), 256, 0, t, sc->windows(),
               first + p0, n, chrom ? sc->o_chrom.p : nullptr, strand ? sc->o_strand.p : nullptr, win ? sc->o_win.p : nullptr, value ? sc->o_value.p : nullptr);
    GTB_TRY(gtb_check_launch(ctx));
    if (chrom) GTB_CUDA_OK(ctx, cudaMemcpyAsync(chrom + p0, sc->o_chrom.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (strand) GTB_CUDA_OK(ctx, cudaMemcpyAsync(strand + p0, sc->o_strand.p, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    if (win) GTB_CUDA_OK(ctx, cudaMemcpyAsync(win + p0, sc->o_win.p, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (value) GTB_CUDA_OK(ctx, cudaMemcpyAsync(value + p0, sc->o_value.p, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  }
  GTB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  return GTB_OK;
}

extern "C" int gtb_scan_peaks(gtb_scan *sc, gtb_scan *control, const gtb_peaks_params *params, int64_t *n_windows) {
  if (!sc || !params || !n_windows) return GTB_ERR_ARG;
  gtb_ctx *ctx = sc->ctx;
  *n_windows = 0; sc->n_peaks = 0; sc->n_out = 0;                        // (the window arrays are shared with gtb_scan_finish)
  if (params->method < GTB_PEAKS_BINOMIAL || params->method > GTB_PEAKS_NORMAL) return gtb_fail(ctx, GTB_ERR_ARG, "unknown probability distribution");
  if (!params->compare && params->method != GTB_PEAKS_BINOMIAL && params->method != GTB_PEAKS_POISSON)
    return gtb_fail(ctx, GTB_ERR_ARG, "unknown probability distribution");                 // genomic_scans.cpp:349
  if (control) {
    if (control->ctx != ctx || control->total_micro != sc->total_micro || control->total_windows != sc->total_windows || control->n_slots != sc->n_slots ||
        control->combine != sc->combine || control->prm.win_step != sc->prm.win_step)
      return gtb_fail(ctx, GTB_ERR_ARG, "signal and control scanners differ in genome or window parameters");
  }
  if (sc->total_windows == 0) return GTB_OK;
  GTB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  SlotTable t{sc->n_slots, sc->d_hist_off.p, sc->d_win_off.p, sc->d_spurious.p, sc->d_slot_chrom.p, sc->d_slot_strand.p};
  PeaksView pv;
  pv.prm = *params; pv.win_size = (long long)sc->prm.win_size;
  pv.control = control ? control->table() : ScanTable{nullptr, nullptr, nullptr};
  const int64_t n_tiles = (sc->total_windows + PK_TILE - 1) / PK_TILE;
  GTB_TRY(sc->d_tile_counts.reserve(ctx, (size_t)n_tiles));
  const WindowsOut none{nullptr, nullptr, nullptr, nullptr, nullptr};
  GTB_LAUNCH(ctx, "scan_peaks_count", scan_peaks_kernel<0>, (unsigned)n_tiles, PK_THREADS, 0, t, sc->table(), pv, sc->total_windows, sc->combine,
             sc->d_tile_counts.p, none, (double *)nullptr, (double *)nullptr);
  GTB_TRY(gtb_check_launch(ctx));
  GTB_TRY(gtb_inclusive_scan_u64(ctx, sc->d_tile_counts.p, n_tiles, sc->d_scan_scratch));
  GTB_CUDA_OK(ctx, cudaMemcpyAsync(sc->h_flags + 2, sc->d_tile_counts.p + (n_tiles - 1), sizeof(ull), cudaMemcpyDeviceToHost, ctx->stream));
  GTB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  ull total;
  memcpy(&total, sc->h_flags + 2, sizeof total);
  sc->n_peaks = (int64_t)total;
  *n_windows = sc->n_peaks;
  if (total == 0) return GTB_OK;
  GTB_TRY(sc->c_win.reserve(ctx, (size_t)total));
  if (sc->n_slots <= 65536) GTB_TRY(sc->c_slot16.reserve(ctx, (size_t)total)); else GTB_TRY(sc->c_slot32.reserve(ctx, (size_t)total));
  GTB_TRY(sc->pk_p1.reserve(ctx, (size_t)total)); GTB_TRY(sc->pk_p2.reserve(ctx, (size_t)total));
  GTB_LAUNCH(ctx, "scan_peaks_emit", scan_peaks_kernel<1>, (unsigned)n_tiles, PK_THREADS, 0, t, sc->table(), pv, sc->total_windows, sc->combine,
             sc->d_tile_counts.p, sc->windows(), sc->pk_p1.p, sc->pk_p2.p);
  GTB_TRY(gtb_check_launch(ctx));
  GTB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  return GTB_OK;
}

extern "C" int gtb_scan_peaks_fetch(gtb_scan *sc, int64_t first, int64_t count, int32_t *chrom, int8_t *strand, int64_t *win, double *pval1, double *pval2) {
  if (!sc || first < 0 || count < 0 || first + count > sc->n_peaks) return GTB_ERR_ARG;
  gtb_ctx *ctx = sc->ctx;
  if (count == 0) return GTB_OK;
  GTB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  SlotTable t{sc->n_slots, sc->d_hist_off.p, sc->d_win_off.p, sc->d_spurious.p, sc->d_slot_chrom.p, sc->d_slot_strand.p};
  const int64_t PIECE = (int64_t)4 << 20;
  const size_t np = (size_t)std::min(count, PIECE);
  if (chrom) GTB_TRY(sc->o_chrom.reserve(ctx, np));
  if (strand) GTB_TRY(sc->o_strand.reserve(ctx, np));
  if (win) GTB_TRY(sc->o_win.reserve(ctx, np));
  for (int64_t p0 = 0; p0 < count; p0 += PIECE) {
    const int64_t n = std::min(PIECE, count - p0);
    if (chrom || strand || win) {
      GTB_LAUNCH(ctx, "scan_peaks_expand", scan_peaks_expand_kernel, (unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 8), 256, 0, t, sc->windows(),
                 first + p0, n, chrom ? sc->o_chrom.p : nullptr, strand ? sc->o_strand.p : nullptr, win ? sc->o_win.p : nullptr);
      GTB_TRY(gtb_check_launch(ctx));
    }
    if (chrom) GTB_CUDA_OK(ctx, cudaMemcpyAsync(chrom + p0, sc->o_chrom.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (strand) GTB_CUDA_OK(ctx, cudaMemcpyAsync(strand + p0, sc->o_strand.p, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    if (win) GTB_CUDA_OK(ctx, cudaMemcpyAsync(win + p0, sc->o_win.p, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (pval1) GTB_CUDA_OK(ctx, cudaMemcpyAsync(pval1 + p0, sc->pk_p1.p + first + p0, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (pval2) GTB_CUDA_OK(ctx, cudaMemcpyAsync(pval2 + p0, sc->pk_p2.p + first + p0, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  }
  GTB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  return GTB_OK;
}
