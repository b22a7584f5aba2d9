// gtb_bucket.cu -- BUCKET engine: partition the queries by genome bucket, then rank them against the
// bucket's evaluation points entirely in shared memory.
//
// Why (measured on B200, profiles/microbench/red_rate_b200.txt): a random-address red.global costs
// ~1.47 SM-cycles per element (1.9e11/s chip-wide, the same for u32 and u64) whereas a random
// shared-memory atomicAdd costs ~0.18.  One global reduction per query therefore caps any
// single-pass design at ~2e11 queries/s before any other work; counting in shared memory does not.
// Shared memory cannot hold the evaluation points of a whole genome, so the queries are first
// bucketed by position:
//
//   pass 1  bucket_partition_kernel   13 B/query in, 4 B/query out
//       The groups' coordinate axes are laid end to end into one axis u (cells of 2^k bp only serve
//       as alignment / directory granularity).  A bucket is 2^ub consecutive u (ub <= 24).  A query
//       becomes one 32-bit element  (u_start - bucket_start) | length << ub.  Per tile of 4 096
//       queries: rank inside the bucket with ONE shared-memory atomicAdd per query, exclusive scan
//       of the bucket counts, scatter into a shared staging buffer, copy out in bucket order so the
//       global stores are runs.  Bucket storage is paged (4 096-element pages from a pool, page
//       table per bucket) so memory is exact whatever the skew; a tile reserves its slice of each
//       bucket with one global atomicAdd per non-empty bucket.
//   pass 2  bucket_count_kernel       4 B/query in
//       Work units of 65 536 elements; a CTA walks a contiguous range of units.  Per bucket it
//       loads the bucket's evaluation points (bucket-local u) and a per-cell directory into shared
//       memory, then for every element: directory lookup + a short forward scan give the slots of
//       start and stop, and the slot histograms (the RANK engine's both / S / E planes) take
//       shared-memory atomics.  Histograms are flushed to the global planes with a reduction per
//       non-zero slot when the CTA moves to another bucket.
//
// What cannot be expressed in an element (start <= 0, length >= 2^(32-ub), a query that crosses the
// end of its bucket, strand bytes other than '+'/'-', invalid intervals) takes the general rank
// step inline in pass 1.  Finalisation is the RANK engine's.
#include "gtb_rank_device.cuh"
#include "gtb_wc_partition.cuh"
#include <algorithm>

namespace {

#ifndef GTB_PART_THREADS
#define GTB_PART_THREADS 512
#endif
#ifndef GTB_PART_CTAS
#define GTB_PART_CTAS 2
#endif
constexpr int PART_THREADS = GTB_PART_THREADS;
constexpr int PART_CTAS = GTB_PART_CTAS;                  // resident CTAs per SM the kernel is sized for
constexpr int PART_ITEMS = 8;
constexpr int PART_TILE = PART_THREADS * PART_ITEMS;      // 4 096 queries
constexpr int PAGE_SHIFT = 12;
constexpr uint32_t PAGE = 1u << PAGE_SHIFT;               // elements per page (== PART_TILE: a tile's slice spans <= 2 pages)
constexpr int UNIT_PAGES = 16;                            // pass-2 work unit = 65 536 elements
constexpr int MAX_BUCKETS = 2048;
constexpr size_t COUNT_SMEM_BUDGET = 200 * 1024;

struct BucketView {
  int k, ub;                        // cell = 2^k bp, bucket = 2^ub u
  int32_t n_chrom, n_class, n_groups;
  int cls_plus, cls_minus;
  const int8_t *class_of;
  const uint8_t *chrom_present;
  const int2 *gtab;                 // per group: (largest point, first cell)
  const int4 *pm_tab;               // [2 * n_chrom] per (chromosome, '+'/'-'): (largest point, u0 & (2^ub - 1), u0 >> ub, 0)
  uint32_t n_buckets;
  // paged bucket storage
  uint32_t *pool;                   // pages of PAGE elements
  uint32_t *page_table;             // [n_buckets * pt_stride]  gen << 24 | (page id + 1); an entry counts only if its tag is the batch's
  uint32_t gen;                     // generation tag of this batch (1..255) -- saves clearing the table per batch
  uint32_t pt_stride;
  uint32_t *cursor;                 // [n_buckets] elements appended so far
  uint32_t *next_page;
  // pass 2
  const int32_t *j0;                // [n_buckets + 1] first slot of each bucket
  const uint32_t *slot_lu;          // [n_slots] bucket-local u of each slot
  const ull *slot_u0;               // [n_slots] u of coordinate 0 of the slot's group
  const uint16_t *dir;              // [n_buckets << (ub - k)] first local slot at or after the cell start
  uint32_t *unit_off;               // [n_buckets + 1]
  int max_local;                    // largest number of slots in a bucket (excluding the catch-all)
  // write-combining form of pass 1 (gtb_wc_partition.cuh): what pass 2 reads of it
  const uint32_t *line_off;         // [n_buckets + 1] first block of each bucket in sorted_lines
  const uint32_t *sorted_lines;     // [total blocks] block slot | (elements - 1) << 25, grouped by bucket
};

__device__ __forceinline__ void red_add64(ull *p, ull v) { asm volatile("red.global.add.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}


// The rare queries an element cannot describe: admission checks of the reference, then the general
// rank step (binary search in global memory).
template <bool COVERAGE>
__device__ __noinline__ void special_query(const BucketView &bv, const RankView &rv, int32_t c, int32_t qs, int32_t qe, int strand,
                                           int64_t w, int64_t index) {
  if (qe <= 0 || qs > qe) {                                                // fatal only on indexed chromosomes, :5731-5741
    if (!bv.chrom_present[c] || !admit_interval(rv, qs, qe, index)) return;
  }
  const int cls = bv.class_of[(uint8_t)strand];
  if (cls < 0) return;                                                     // no index region carries this strand, :5229
  const int g = c * bv.n_class + cls;
  const int gb = rv.goff[g], ge = rv.goff[g + 1];
  if (ge > gb) rank_item<COVERAGE>(rv, gb, ge, qs, qe, w);
}

// The overlap engines' front for the write-combining partition (gtb_wc_partition.cuh).  Table: one 16-byte entry per
// (chromosome, '+'/'-') group, x = max(largest evaluation point, 0) so that one unsigned compare covers 1 <= start <= x,
// y = (u of coordinate 0 of the group) & (bucket size - 1), z = (u of coordinate 0) >> ub, w = the signed largest point;
// the last two entries mean "no such group".  For a start coordinate s:  t = y + s,  bucket = z + (t >> ub),  bucket-local
// u = t & (2^ub - 1);  element = local u | length << ub.
template <bool COVERAGE>
struct OverlapFront {
  static constexpr bool FAIL_IS_SLOW = true;
  RankView rv;
  BucketView bv;
  __device__ __forceinline__ uint32_t table_size() const { return 2u * (uint32_t)bv.n_chrom + 2u; }
  __device__ __forceinline__ uint4 table_entry(uint32_t i) const {
    int4 g = i < 2u * (uint32_t)bv.n_chrom ? bv.pm_tab[i] : make_int4(0, 0, 0, 0);
    g.w = g.x; g.x = max(g.x, 0);
    return make_uint4((uint32_t)g.x, (uint32_t)g.y, (uint32_t)g.z, (uint32_t)g.w);
  }
  __device__ __forceinline__ uint32_t table_index(int32_t c, uint32_t xw, int i) const {
    return 2u * min((uint32_t)c, (uint32_t)bv.n_chrom) + ((xw >> (8 * i + 1)) & 1u);          // unknown chromosome -> an empty entry
  }
  __device__ __forceinline__ uint32_t slow_mask(uint32_t xw) const { return (xw & 0xF9F9F9F9u) ? 0xFu : 0u; }   // a strand other than '+'/'-'
  __device__ __forceinline__ uint32_t classify(uint4 g, int32_t s, int32_t e, uint32_t &elem) const {
    const uint32_t ub = (uint32_t)bv.ub, ubmask = (1u << ub) - 1u, len_max = 0xFFFFFFFFu >> ub;
    const uint32_t t = g.y + (uint32_t)s;
    const uint32_t lu = t & ubmask;
    const uint32_t len = (uint32_t)(min(e, (int)g.x + 1) - s);                                // e < s wraps to a huge value and fails the next test
    const bool ok = (uint32_t)(s - 1) < g.x && len <= len_max && lu + len <= ubmask;
    elem = lu | (len << ub);
    return ok ? g.z + (t >> ub) : WC_NONE;
  }
  // everything the fast classification cannot decide (rare): the reference's admission rules, then the general rank step
  __device__ __forceinline__ uint32_t resolve(uint32_t bucket, int32_t c, int32_t s, int32_t e, uint32_t sbyte, int64_t index) const {
    const uint32_t d = sbyte - (uint32_t)'+';
    const bool addressable = (d & ~2u) == 0;
    if (bucket != WC_NONE && addressable) return bucket;                                       // only flagged because a sibling has an odd strand
    if ((uint32_t)c >= (uint32_t)bv.n_chrom) return WC_NONE;                                   // chromosome the index has never seen
    const int gx = addressable ? bv.pm_tab[2 * c + (int)(d >> 1)].x : 0;
    const bool valid = s >= 1 && s <= e;
    const bool nothing = addressable && valid && gx >= 0 && (gx == 0 || s > gx);              // no points in the group / start beyond the last one
    if (!nothing) special_query<COVERAGE>(bv, rv, c, s, e, (int)(int8_t)sbyte, 1, index);
    return WC_NONE;
  }
  __device__ __forceinline__ void divert(int32_t c, int32_t s, int32_t e, uint32_t sbyte, int64_t index) const {
    special_query<COVERAGE>(bv, rv, c, s, e, (int)(int8_t)sbyte, 1, index);
  }
};

// ------------------------------------------------------------------------------------------------
// pass 1
// ------------------------------------------------------------------------------------------------
// Per (chromosome, '+'/'-') group, staged in shared memory as one 16-byte entry:
//   x = largest evaluation point (0: no such group / no points, < 0: only points <= 0)
//   y = (u of coordinate 0 of the group) & (bucket size - 1)      z = (u of coordinate 0) >> ub
// so that for a start coordinate s:  t = y + s,  bucket = z + (t >> ub),  bucket-local u = t & (2^ub - 1).
template <bool COVERAGE, int VEC, int RES>
__global__ void __launch_bounds__(PART_THREADS, PART_CTAS) bucket_partition_kernel(const __grid_constant__ QueryView q, const __grid_constant__ RankView rv,
                                                                                    const __grid_constant__ BucketView bv) {
  extern __shared__ __align__(128) uint32_t smem[];
  // raw tile, filled by TMA bulk copies: chrom | start | stop (PART_TILE ints each) | strand (PART_TILE bytes)
  int32_t *s_chrom = reinterpret_cast<int32_t *>(smem);
  int32_t *s_start = s_chrom + PART_TILE;
  int32_t *s_stop = s_start + PART_TILE;
  int8_t *s_strand = reinterpret_cast<int8_t *>(s_stop + PART_TILE);
  uint2 *s_stage = reinterpret_cast<uint2 *>(smem + 3 * PART_TILE + PART_TILE / 4);   // [PART_TILE] (element, bucket) in bucket order
  uint32_t *s_cnt = smem + 3 * PART_TILE + PART_TILE / 4 + 2 * PART_TILE + 4;         // [nb4 + 4] per-bucket counts of this tile (+ dummy)
  const uint32_t nb4 = (bv.n_buckets + 3) & ~3u;
  uint32_t *s_off = s_cnt + nb4 + 4;                                  // [nb4 + 4] exclusive scan of s_cnt
  uint2 *s_dl = reinterpret_cast<uint2 *>(s_off + nb4 + 4);           // [nb4] (global address - staged position, staged position where the slice spills into the next page)
  uint32_t *s_g1 = s_off + 3 * nb4 + 4;                               // [nb4] global address of the spilled part
  int4 *s_pm = reinterpret_cast<int4 *>(s_off + 4 * nb4 + 4);         // [2 * n_chrom] group table
  __shared__ uint32_t s_warp_tot[PART_THREADS / 32];
  __shared__ __align__(8) uint64_t s_bar;

  for (int i = threadIdx.x; i < 2 * bv.n_chrom; i += blockDim.x) s_pm[i] = bv.pm_tab[i];
  for (uint32_t i = threadIdx.x; i < nb4 + 4; i += blockDim.x) s_cnt[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&s_bar, 1); fence_proxy_async(); }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t n_tiles = (q.n_regions + PART_TILE - 1) / PART_TILE;
  const int64_t n_full = VEC == 8 ? q.n_regions / PART_TILE : 0;      // tiles that TMA can fetch (complete, aligned)
  const uint32_t ubmask = (1u << bv.ub) - 1u;
  const uint32_t len_max = 0xFFFFFFFFu >> bv.ub;
  const uint32_t n_chrom = (uint32_t)bv.n_chrom;
  constexpr uint32_t TILE_BYTES = PART_TILE * 13;

  auto issue = [&](int64_t tile) {                                    // one thread: 4 bulk copies, 53 248 bytes
    const int64_t first = tile * PART_TILE;
    mbar_expect_tx(&s_bar, TILE_BYTES);
    tma_bulk_g2s(s_chrom, q.chrom + first, PART_TILE * 4, &s_bar);
    tma_bulk_g2s(s_start, q.start + first, PART_TILE * 4, &s_bar);
    tma_bulk_g2s(s_stop, q.stop + first, PART_TILE * 4, &s_bar);
    tma_bulk_g2s(s_strand, q.strand + first, PART_TILE, &s_bar);
  };
  if (threadIdx.x == 0 && (int64_t)blockIdx.x < n_full) issue(blockIdx.x);
  uint32_t parity = 0;

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    // ---- classify, build the element, rank inside the bucket (one shared atomic per query)
    uint32_t elem[PART_ITEMS], br[PART_ITEMS];                         // br = bucket << 16 | rank  (bucket nb4: no element)
    {
      int32_t c[PART_ITEMS], s[PART_ITEMS], e[PART_ITEMS];
      uint2 stw;
      if (tile < n_full) {
        mbar_wait(&s_bar, parity);
        parity ^= 1u;
        const int4 c0 = reinterpret_cast<const int4 *>(s_chrom)[threadIdx.x * 2], c1 = reinterpret_cast<const int4 *>(s_chrom)[threadIdx.x * 2 + 1];
        const int4 s0 = reinterpret_cast<const int4 *>(s_start)[threadIdx.x * 2], s1 = reinterpret_cast<const int4 *>(s_start)[threadIdx.x * 2 + 1];
        const int4 e0 = reinterpret_cast<const int4 *>(s_stop)[threadIdx.x * 2], e1 = reinterpret_cast<const int4 *>(s_stop)[threadIdx.x * 2 + 1];
        stw = reinterpret_cast<const uint2 *>(s_strand)[threadIdx.x];
        c[0] = c0.x; c[1] = c0.y; c[2] = c0.z; c[3] = c0.w; c[4] = c1.x; c[5] = c1.y; c[6] = c1.z; c[7] = c1.w;
        s[0] = s0.x; s[1] = s0.y; s[2] = s0.z; s[3] = s0.w; s[4] = s1.x; s[5] = s1.y; s[6] = s1.z; s[7] = s1.w;
        e[0] = e0.x; e[1] = e0.y; e[2] = e0.z; e[3] = e0.w; e[4] = e1.x; e[5] = e1.y; e[6] = e1.z; e[7] = e1.w;
      } else {
        const int64_t first = tile * PART_TILE + (int64_t)threadIdx.x * PART_ITEMS;
        unsigned sx = 0, sy = 0;
#pragma unroll
        for (int i = 0; i < PART_ITEMS; i++) {
          const int64_t r = first + i;
          const bool ok = r < q.n_regions;
          c[i] = ok ? q.chrom[r] : -1; s[i] = ok ? q.start[r] : 1; e[i] = ok ? q.stop[r] : 1;
          const unsigned sb = ok ? (unsigned)(uint8_t)q.strand[r] : (unsigned)'+';
          if (i < 4) sx |= sb << (i * 8); else sy |= sb << ((i & 3) * 8);
        }
        stw = make_uint2(sx, sy);
      }
#pragma unroll
      for (int i = 0; i < PART_ITEMS; i++) {
        const uint32_t sbyte = ((i < 4 ? stw.x : stw.y) >> ((i & 3) * 8)) & 0xFFu;
        const uint32_t d = sbyte - (uint32_t)'+';                                   // '+' -> 0, '-' -> 2
        const bool addressable = (uint32_t)c[i] < n_chrom && (d & ~2u) == 0;        // known chromosome, '+'/'-' strand
        const int4 gt = s_pm[addressable ? 2 * c[i] + (int)(d >> 1) : 0];
        const uint32_t len = (uint32_t)(min(e[i], gt.x + 1) - s[i]);
        const uint32_t t = (uint32_t)gt.y + (uint32_t)s[i];
        const uint32_t lu = t & ubmask;
        // the common case: valid interval starting inside the cells, before the group's last point, short
        // enough for the length field, not crossing the end of its bucket
        const bool normal = addressable && s[i] >= 1 && s[i] <= e[i] && s[i] <= gt.x && len <= len_max && lu + len <= ubmask;
        elem[i] = lu | (len << bv.ub);
        br[i] = nb4 << 16;                                                          // bucket nb4: no element
        if (normal) {
          const uint32_t b = (uint32_t)gt.z + (t >> bv.ub);
          br[i] = (b << 16) | atomicAdd(&s_cnt[b], 1u);
        } else if ((uint32_t)c[i] < n_chrom) {
          // rare: decide between "nothing to count" and the general path.  Nothing: a valid '+'/'-' query that
          // starts at >= 1 in a group without points, or beyond the group's last point.
          const bool nothing = addressable && s[i] >= 1 && s[i] <= e[i] && gt.x >= 0 && (gt.x == 0 || s[i] > gt.x);
          if (!nothing)
            special_query<COVERAGE>(bv, rv, c[i], s[i], e[i], (int)(int8_t)sbyte, 1,
                                    q.index_base + tile * PART_TILE + (int64_t)threadIdx.x * PART_ITEMS + i);
        }
      }
    }
    __syncthreads();
    // the raw tile has been consumed by every thread: fetch the next one while the rest of this tile is processed
    if (threadIdx.x == 0 && tile + gridDim.x < n_full) { fence_proxy_async(); issue(tile + gridDim.x); }

    // ---- reserve global space: one global atomic per non-empty bucket, one bucket per thread, all
    // issued before anything waits on them; the exclusive scan of the counts runs in their shadow
    uint32_t r_cnt[RES], r_old[RES], r_g0[RES], r_g1[RES], r_split[RES];
#pragma unroll
    for (int j = 0; j < RES; j++) {
      const uint32_t b = threadIdx.x + j * PART_THREADS;
      r_cnt[j] = b < bv.n_buckets ? s_cnt[b] : 0u;
      r_old[j] = 0; r_g0[j] = 0; r_g1[j] = 0; r_split[j] = 0;
      if (r_cnt[j]) r_old[j] = atomicAdd(bv.cursor + b, r_cnt[j]);
    }
    {
      uint32_t v[4] = {0, 0, 0, 0}, sum = 0;
      const uint32_t b0 = threadIdx.x * 4;
      if (b0 < nb4) {
        const uint4 t = *reinterpret_cast<const uint4 *>(s_cnt + b0);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; sum = t.x + t.y + t.z + t.w;
      }
      uint32_t inc = sum;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
      if (lane == 31) s_warp_tot[warp] = inc;
      __syncthreads();
      uint32_t base = inc - sum;
      for (int w = 0; w < warp; w++) base += s_warp_tot[w];
      if (b0 < nb4) {
        uint4 o;
        o.x = base; o.y = base + v[0]; o.z = o.y + v[1]; o.w = o.z + v[2];
        *reinterpret_cast<uint4 *>(s_off + b0) = o;
      }
    }
    // pages: allocate the ones whose first element is ours, wait for the ones somebody else allocates
#pragma unroll
    for (int j = 0; j < RES; j++) {
      if (r_cnt[j]) {
        const uint32_t b = threadIdx.x + j * PART_THREADS, old = r_old[j];
        const uint32_t p_first = old >> PAGE_SHIFT, p_last = (old + r_cnt[j] - 1) >> PAGE_SHIFT;
        uint32_t *pt = bv.page_table + (size_t)b * bv.pt_stride;
        const uint32_t tag = bv.gen << 24;
        if ((old & (PAGE - 1)) == 0) { const uint32_t pid = atomicAdd(bv.next_page, 1u); atomicExch(pt + p_first, tag | (pid + 1)); }
        if (p_last != p_first) { const uint32_t pid = atomicAdd(bv.next_page, 1u); atomicExch(pt + p_last, tag | (pid + 1)); }
        uint32_t id0, id1;
        while (((id0 = ld_volatile_u32(pt + p_first)) >> 24) != bv.gen) {}
        id1 = id0;
        if (p_last != p_first) while (((id1 = ld_volatile_u32(pt + p_last)) >> 24) != bv.gen) {}
        id0 &= 0xFFFFFFu; id1 &= 0xFFFFFFu;
        r_g0[j] = ((id0 - 1) << PAGE_SHIFT) + (old & (PAGE - 1));
        r_g1[j] = (id1 - 1) << PAGE_SHIFT;
        r_split[j] = min(r_cnt[j], PAGE - (old & (PAGE - 1)));
      }
    }
    __syncthreads();
    // per bucket: (address delta of the first part, staged position where the second part begins) and the second part's base
#pragma unroll
    for (int j = 0; j < RES; j++) {
      const uint32_t b = threadIdx.x + j * PART_THREADS;
      if (b < bv.n_buckets) {
        const uint32_t off = s_off[b];
        s_dl[b] = make_uint2(r_g0[j] - off, off + r_split[j]);
        s_g1[b] = r_g1[j];
      }
    }

    // ---- scatter into the staging buffer in bucket order
#pragma unroll
    for (int i = 0; i < PART_ITEMS; i++) {
      const uint32_t b = br[i] >> 16;
      const uint32_t pos = b == nb4 ? (uint32_t)PART_TILE : s_off[b] + (br[i] & 0xFFFFu);    // PART_TILE = trash slot
      s_stage[pos] = make_uint2(elem[i], b);
    }
    const uint32_t total = s_off[nb4 - 1] + s_cnt[nb4 - 1];
    __syncthreads();

    // ---- copy out: consecutive staged elements of a bucket go to consecutive addresses
#pragma unroll 4
    for (uint32_t p = threadIdx.x; p < total; p += PART_THREADS) {
      const uint2 eb = s_stage[p];
      const uint2 dl = s_dl[eb.y];
      const uint32_t addr = p < dl.y ? p + dl.x : s_g1[eb.y] + (p - dl.y);
      bv.pool[addr] = eb.x;
    }
    for (uint32_t i = threadIdx.x; i < nb4 + 4; i += blockDim.x) s_cnt[i] = 0;
    __syncthreads();
  }
}


// ------------------------------------------------------------------------------------------------
// between the passes: work units per bucket (exclusive scan of ceil(count / unit size))
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) bucket_units_kernel(BucketView bv) {
  __shared__ uint32_t s_tot[32];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint32_t base = 0; base < bv.n_buckets; base += blockDim.x) {
    const uint32_t b = base + threadIdx.x;
    const uint32_t u = b < bv.n_buckets ? (bv.cursor[b] + (PAGE * UNIT_PAGES) - 1) / (PAGE * UNIT_PAGES) : 0;
    uint32_t inc = u;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
    if (lane == 31) s_tot[warp] = inc;
    __syncthreads();
    uint32_t pre = carry;
    for (int w = 0; w < warp; w++) pre += s_tot[w];
    if (b < bv.n_buckets) bv.unit_off[b] = pre + inc - u;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = pre + inc;
    __syncthreads();
  }
  if (threadIdx.x == 0) bv.unit_off[bv.n_buckets] = carry;
}

// ------------------------------------------------------------------------------------------------
// pass 2
// ------------------------------------------------------------------------------------------------
// COUNT_THREADS x resident CTAs = 1 024 threads per SM: an index dense enough to need most of an SM's shared memory for one
// bucket (1 M regions) runs one CTA of 1 024 threads, a sparse one (60 k regions) four of 256
template <bool COVERAGE, bool LINES, int COUNT_THREADS>
__global__ void __launch_bounds__(COUNT_THREADS, 1024 / COUNT_THREADS) bucket_count_kernel(BucketView bv, RankView rv) {
  extern __shared__ __align__(128) uint32_t smem[];
  const int cb = bv.ub - bv.k;
  const uint32_t n_dir = 1u << cb;
  uint16_t *s_dir = reinterpret_cast<uint16_t *>(smem);                       // [n_dir]
  uint32_t *s_pts = smem + ((n_dir / 2 + 1) & ~1u);                           // [max_local + 1], last = +inf (even word offset)
  const uint32_t cap = (uint32_t)bv.max_local + 1;
  // histogram planes: count -> 3 planes of u32 ; coverage -> 5 planes of u64
  uint32_t *s_h32 = s_pts + ((cap + 1) & ~1u);
  ull *s_h64 = reinterpret_cast<ull *>(s_h32);

  // work items: 65 536-element units of the paged buckets, or 16-element lines of the write-combined ones
  const uint32_t *w_off = LINES ? bv.line_off : bv.unit_off;
  const uint32_t total_units = w_off[bv.n_buckets];
  const uint32_t u_begin = (uint32_t)(((ull)total_units * blockIdx.x) / gridDim.x);
  const uint32_t u_end = (uint32_t)(((ull)total_units * (blockIdx.x + 1)) / gridDim.x);
  if (u_begin >= u_end) return;
  const uint32_t kmaskless = bv.k;
  const uint32_t ubmask = (1u << bv.ub) - 1u;
  const int64_t K = rv.n_slots;

  // bucket of the first unit: last b with unit_off[b] <= u_begin
  uint32_t b;
  {
    uint32_t lo = 0, hi = bv.n_buckets;
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (w_off[mid] <= u_begin) lo = mid; else hi = mid; }
    b = lo;
    while (b + 1 < bv.n_buckets && w_off[b + 1] <= u_begin) b++;            // skip empty buckets sharing the offset
  }
  int loaded = -1;
  uint32_t n_local = 0;
  int32_t slot0 = 0;

  auto flush = [&]() {
    if (loaded < 0) return;
    const ull bucket_u = (ull)(uint32_t)loaded << bv.ub;
    for (uint32_t j = threadIdx.x; j <= n_local; j += blockDim.x) {
      const int64_t J = (int64_t)slot0 + j;
      if (J >= K) continue;
      if (!COVERAGE) {
        const uint32_t vb = s_h32[j], vs = s_h32[cap + j], ve = s_h32[2 * cap + j];
        if (vb) red_add64(rv.hist + H_BOTH * K + J, vb);
        if (vs) red_add64(rv.hist + H_SCNT * K + J, vs);
        if (ve) red_add64(rv.hist + H_ECNT * K + J, ve);
      } else {
        const ull vb = s_h64[j], vs = s_h64[cap + j], ve = s_h64[2 * cap + j], ss = s_h64[3 * cap + j], se = s_h64[4 * cap + j];
        const ull shift = bucket_u - bv.slot_u0[J];                     // bucket-local u -> coordinate inside the slot's group
        if (vb) red_add64(rv.hist + H_BOTH * K + J, vb);
        if (vs) { red_add64(rv.hist + H_SCNT * K + J, vs); red_add64(rv.hist + H_SSUM * K + J, ss + vs * shift); }
        if (ve) { red_add64(rv.hist + H_ECNT * K + J, ve); red_add64(rv.hist + H_ESUM * K + J, se + ve * shift); }
      }
    }
  };
  auto load_bucket = [&](uint32_t nb) {
    __syncthreads();
    flush();
    __syncthreads();
    loaded = (int)nb;
    slot0 = bv.j0[nb];
    n_local = (uint32_t)(bv.j0[nb + 1] - slot0);
    const uint16_t *gdir = bv.dir + ((size_t)nb << cb);
    for (uint32_t i = threadIdx.x; i < n_dir; i += blockDim.x) s_dir[i] = gdir[i];
    for (uint32_t i = threadIdx.x; i < n_local; i += blockDim.x) s_pts[i] = bv.slot_lu[slot0 + i];
    if (threadIdx.x == 0) s_pts[n_local] = 0xFFFFFFFFu;
    const uint32_t words = COVERAGE ? 5 * cap * 2 : 3 * cap;
    for (uint32_t i = threadIdx.x; i < words; i += blockDim.x) s_h32[i] = 0;
    __syncthreads();
  };

  for (uint32_t u = u_begin; u < u_end;) {
    while (b + 1 < bv.n_buckets && w_off[b + 1] <= u) b++;
    if ((int)b != loaded) load_bucket(b);
    uint32_t n_el, n_v4, e_begin = 0;
    const uint32_t *pt = nullptr;
    if (LINES) {                                                          // all lines of this bucket inside the CTA's range
      const uint32_t seg_end = min(u_end, w_off[b + 1]);
      n_v4 = (seg_end - u) * (WC_BLOCK_ELEMS / 4); n_el = n_v4 * 4;
    } else {
      const uint32_t part = u - bv.unit_off[b];
      const uint32_t cnt = bv.cursor[b];
      e_begin = part * (PAGE * UNIT_PAGES);
      const uint32_t e_end = min(cnt, e_begin + PAGE * UNIT_PAGES);
      pt = bv.page_table + (size_t)b * bv.pt_stride;
      n_el = e_end - e_begin;                                             // elements of this unit (e_begin is page aligned)
      n_v4 = (n_el + 3) >> 2;
    }
    constexpr int UNROLL = 4;
    // block ids of a trip (LINES): fetched one trip ahead, so that a trip waits for one memory latency, not two dependent ones
    uint32_t ent[UNROLL];
    auto fetch_entries = [&](uint32_t base) {
#pragma unroll
      for (int r = 0; r < UNROLL; r++) {
        const uint32_t v4 = base + r * COUNT_THREADS + threadIdx.x;
        ent[r] = v4 < n_v4 ? __ldg(bv.sorted_lines + u + (v4 >> 5)) : 0u;                        // 32 loads of 128 bits per block
      }
    };
    if (LINES) fetch_entries(0);
    for (uint32_t base = 0; base < n_v4; base += COUNT_THREADS * UNROLL) {
      uint4 d[UNROLL];
      uint32_t nv[UNROLL];                                                // valid elements of each 128-bit load
#pragma unroll
      for (int r = 0; r < UNROLL; r++) {
        const uint32_t v4 = base + r * COUNT_THREADS + threadIdx.x;
        d[r] = make_uint4(0, 0, 0, 0);
        nv[r] = 0;
        if (v4 < n_v4) {
          if (LINES) {
            const uint32_t entry = ent[r];
            const uint32_t fill = (entry >> 25) + 1u, q4 = (v4 & 31u) * 4u;
            nv[r] = fill > q4 ? min(fill - q4, 4u) : 0u;
            if (nv[r]) d[r] = ldg_stream128(reinterpret_cast<const uint4 *>(bv.pool + (size_t)(entry & 0x01FFFFFFu) * WC_BLOCK_ELEMS) + (v4 & 31u));
          } else {
            const uint32_t pg = (e_begin >> PAGE_SHIFT) + (v4 >> (PAGE_SHIFT - 2));
            const uint32_t page_id = (__ldg(pt + pg) & 0xFFFFFFu) - 1;
            nv[r] = min(n_el - v4 * 4, 4u);
            d[r] = ldg_stream128(reinterpret_cast<const uint4 *>(bv.pool + ((size_t)page_id << PAGE_SHIFT)) + (v4 & ((PAGE >> 2) - 1)));
          }
        }
      }
      if (LINES) fetch_entries(base + COUNT_THREADS * UNROLL);
#pragma unroll
      for (int r = 0; r < UNROLL; r++) {
        const uint32_t el[4] = {d[r].x, d[r].y, d[r].z, d[r].w};
        uint32_t us[4], ue[4], jS[4], pS[4];
#pragma unroll
        for (int i = 0; i < 4; i++) { us[i] = el[i] & ubmask; ue[i] = us[i] + (el[i] >> bv.ub); jS[i] = s_dir[us[i] >> kmaskless]; }
#pragma unroll
        for (int i = 0; i < 4; i++) pS[i] = s_pts[jS[i]];
        uint32_t jEv[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
          while (pS[i] < us[i]) pS[i] = s_pts[++jS[i]];                  // first slot whose point is >= start
          uint32_t jE = jS[i], pE = pS[i];
          while (pE < ue[i]) pE = s_pts[++jE];                           // ... >= stop
          jEv[i] = jE;
        }
        // Sorted input: the whole warp (4 x 32 consecutive elements) sits in one slot; one atomic instead of 128 on one address
        if (!COVERAGE) {
          const uint32_t j0s = __shfl_sync(0xffffffffu, jS[0], 0);
          const bool same = nv[r] == 4u && jS[0] == j0s && jS[1] == j0s && jS[2] == j0s && jS[3] == j0s && jEv[0] == j0s && jEv[1] == j0s &&
                            jEv[2] == j0s && jEv[3] == j0s;
          if (__all_sync(0xffffffffu, same)) {
            if ((threadIdx.x & 31) == 0) atomicAdd(&s_h32[j0s], 128u);
            continue;
          }
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const uint32_t jE = jEv[i];
          if ((uint32_t)i < nv[r]) {
            if (!COVERAGE) {
              if (jS[i] == jE) atomicAdd(&s_h32[jS[i]], 1u);
              else { atomicAdd(&s_h32[cap + jS[i]], 1u); atomicAdd(&s_h32[2 * cap + jE], 1u); }
            } else {
              if (jS[i] == jE) {                                          // 64-bit sum as a 32-bit atomic + a rare carry: far cheaper than atom.shared.add.u64
                const uint32_t v = ue[i] - us[i] + 1u;
                const uint32_t old = atomicAdd(&s_h32[2 * jS[i]], v);
                if (old + v < old) atomicAdd(&s_h32[2 * jS[i] + 1], 1u);
              }
              else {
                atomicAdd(&s_h64[cap + jS[i]], 1ull); atomicAdd(&s_h64[2 * cap + jE], 1ull);
                atomicAdd(&s_h64[3 * cap + jS[i]], (ull)us[i]); atomicAdd(&s_h64[4 * cap + jE], (ull)ue[i]);
              }
            }
          }
        }
      }
    }
    u = LINES ? min(u_end, w_off[b + 1]) : u + 1;
  }
  __syncthreads();
  flush();
}

// pass 2 with as many threads per CTA as keeps 1 024 threads on an SM
template <bool COVERAGE, bool LINES>
int launch_count_pass(gtb_ctx *ctx, const char *name, unsigned ctas_per_sm, size_t smem, const BucketView &bv, const RankView &rv) {
#define GTB_COUNT_LAUNCH(T)                                                                                                       \
  do {                                                                                                                            \
    GTB_CUDA_OK(ctx, cudaFuncSetAttribute(bucket_count_kernel<COVERAGE, LINES, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    GTB_LAUNCH(ctx, name, (bucket_count_kernel<COVERAGE, LINES, T>), (unsigned)ctx->sm_count * (1024 / T), T, smem, bv, rv);       \
  } while (0)
  if (ctas_per_sm >= 4) GTB_COUNT_LAUNCH(256);
  else if (ctas_per_sm >= 2) GTB_COUNT_LAUNCH(512);
  else GTB_COUNT_LAUNCH(1024);
#undef GTB_COUNT_LAUNCH
  return gtb_check_launch(ctx);
}

template <typename T>
int upload_b(gtb_ctx *ctx, dbuf<T> &d, const std::vector<T> &h) {
  GTB_TRY(d.reserve(ctx, h.size() ? h.size() : 1));
  if (h.size()) GTB_CUDA_OK(ctx, cudaMemcpyAsync(d.p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  return GTB_OK;
}

}  // namespace

struct gtb_bucket_state {
  bool ready = false, failed = false;
  uint32_t gen = 0;                 // page-table generation tag
  int k = 0, ub = 0;
  uint32_t n_buckets = 0;
  int max_local = 0;
  size_t count_smem = 0, part_smem = 0;
  dbuf<int2> d_gtab;
  dbuf<int4> d_pm;
  dbuf<int32_t> d_j0;
  dbuf<uint32_t> d_slot_lu;
  dbuf<ull> d_slot_u0;
  dbuf<uint16_t> d_dir;
  dbuf<uint32_t> d_pool, d_page_table, d_cursor, d_next_page, d_unit_off;
  // write-combining form
  bool wc_ok = false;                                   // the bucket count fits (one owner thread per bucket)
  bool wc_off = false;                                  // switched off after a batch with many diverted queries
  size_t wc_smem = 0;
  WcBuffers wc;
  ull diverted_seen = 0;
  int64_t wc_queries = 0;                               // queries sent through the write-combining form since the last check
};

static size_t count_smem_bytes(int ub, int k, int max_local, bool coverage) {
  const size_t n_dir = (size_t)1 << (ub - k);
  const size_t cap = (size_t)max_local + 1;
  return n_dir * 2 + 8 + ((cap + 1) & ~(size_t)1) * 4 + (coverage ? 5 * cap * 8 : 3 * cap * 4) + 16;
}

int gtb_bucket_prepare(gtb_index *ix) {
  if (ix->bucket && (ix->bucket->ready || ix->bucket->failed)) return ix->bucket->ready ? GTB_OK : GTB_ERR_UNSUPPORTED;
  gtb_ctx *ctx = ix->ctx;
  if (!ix->bucket) ix->bucket = new gtb_bucket_state();
  gtb_bucket_state *bs = ix->bucket;
  bs->failed = true;                                                    // until proven otherwise
  const int G = ix->n_groups;
  if (ix->n_slots == 0) return GTB_ERR_UNSUPPORTED;
  std::vector<int32_t> gsize((size_t)std::max(G, 1), 0);
  uint64_t span = 0, n_points = 0;
  for (int g = 0; g < G; g++) {
    const int32_t gb = ix->h_goff[g], ge = ix->h_goff[g + 1];
    if (ge - gb >= 2) { const int32_t mx = ix->h_points[ge - 2]; gsize[g] = mx >= 1 ? mx : -1; n_points += (uint64_t)(ge - gb - 1); }
    if (gsize[g] > 0) span += (uint64_t)gsize[g] + 2;
  }
  if (span == 0) return GTB_ERR_UNSUPPORTED;
  // cell width: ~1/64 evaluation point per cell, so that a warp of pass 2 rarely needs even one step of the forward scan
  // (measured on B200: k = 9 instead of 14 for hg19 x 60 k regions takes bucket_count from 0.283 to 0.221 ms)
  int k = 4;
  while (k < 14 && (span >> (k + 1)) >= 64 * n_points) k++;
  // ... but no finer than 2^15 cells per bucket (a 64 KB directory): a dense index (1 M regions) would otherwise be pushed to
  // buckets so narrow that there are thousands of them
  {
    int ub0 = 24;
    while (ub0 > 20 && (span >> ub0) < 256) ub0--;
    k = std::max(k, ub0 - 15);
  }
  if (const char *env = getenv("GTB_BUCKET_K")) k = std::max(0, std::min(16, atoi(env)));
  std::vector<uint32_t> gbase((size_t)std::max(G, 1), 0);
  uint64_t cells = 0;
  for (int g = 0; g < G; g++) {
    gbase[g] = (uint32_t)cells;
    if (gsize[g] > 0) cells += (((uint64_t)gsize[g] + 1) >> k) + 1;
    if (cells >= ((uint64_t)1 << 31)) return GTB_ERR_UNSUPPORTED;
  }
  // slot coordinates on the concatenated axis
  std::vector<ull> slot_u((size_t)ix->n_slots), slot_u0((size_t)ix->n_slots);
  for (int g = 0; g < G; g++) {
    const int32_t gb = ix->h_goff[g], ge = ix->h_goff[g + 1];
    if (ge == gb) continue;
    const ull u0 = (ull)gbase[g] << k;
    const ull n_cells_g = gsize[g] > 0 ? ((((ull)gsize[g] + 1) >> k) + 1) : 0;
    for (int32_t j = gb; j < ge; j++) {
      const int32_t p = ix->h_points[j];
      slot_u0[j] = u0;
      if (j == ge - 1) slot_u[j] = n_cells_g ? u0 + (n_cells_g << k) - 1 : u0;   // sentinel: last u of the group
      else slot_u[j] = p >= 1 ? u0 + (ull)p : u0;
    }
  }
  // groups without a positive point occupy no cells: their slots share u with the next group's start,
  // which is harmless because no element is ever produced for them.
  for (size_t j = 1; j < slot_u.size(); j++) if (slot_u[j] < slot_u[j - 1]) slot_u[j] = slot_u[j - 1];
  const ull total_u = std::max<ull>(cells << k, 1);
  // bucket width: 2^24 u at most (the element keeps 32 - ub bits for the length), narrower while that leaves fewer than 256
  // buckets -- the write-combining pass wants <= ~8 queries per bucket and round (-i halves the axis: 2^23 there)
  int ub = 24;
  while (ub > 20 && ub > k + 1 && ((total_u + ((ull)1 << ub) - 1) >> ub) < 256) ub--;
  if (const char *env = getenv("GTB_BUCKET_BITS")) ub = std::max(k + 1, std::min(27, atoi(env)));
  if (ub < k + 1) ub = k + 1;
  std::vector<int32_t> j0;
  int max_local = 0;
  const bool cov = ix->op == GTB_OP_COVERAGE;
  for (;; ub--) {
    if (ub < k + 1 || ub < 8) return GTB_ERR_UNSUPPORTED;
    const ull nb = (total_u + ((ull)1 << ub) - 1) >> ub;
    if (nb > MAX_BUCKETS) return GTB_ERR_UNSUPPORTED;
    j0.assign((size_t)nb + 1, 0);
    max_local = 0;
    for (ull b = 0; b <= nb; b++)
      j0[b] = (int32_t)(std::lower_bound(slot_u.begin(), slot_u.end(), b << ub) - slot_u.begin());
    for (ull b = 0; b < nb; b++) max_local = std::max(max_local, j0[b + 1] - j0[b]);
    if (max_local < 65000 && count_smem_bytes(ub, k, max_local, cov) <= std::min(COUNT_SMEM_BUDGET, ctx->smem_optin)) break;
  }
  const uint32_t nb = (uint32_t)(j0.size() - 1);
  bs->k = k; bs->ub = ub; bs->n_buckets = nb; bs->max_local = max_local;
  bs->count_smem = count_smem_bytes(ub, k, max_local, cov);
  const uint32_t nb4 = (nb + 3) & ~3u;
  bs->part_smem = (size_t)PART_TILE * 13 + (size_t)PART_TILE * 8 + 64 + (size_t)nb4 * 20 + (size_t)std::max(ix->n_chrom, 1) * 32;
  // with fewer than 256 buckets a 2 048-query round overfills the 32-element rings too often; GTB_BUCKET_WC=1 forces it (tests)
  bs->wc_ok = nb <= (uint32_t)WC_MAX_BUCKETS && (nb >= 256 || getenv("GTB_BUCKET_WC") != nullptr);
  bs->wc_smem = wc_smem_bytes(nb, (size_t)2 * std::max(ix->n_chrom, 1) + 2);
  // directory and bucket-local slot coordinates
  const int cb = ub - k;
  std::vector<uint16_t> dir((size_t)nb << cb);
  std::vector<uint32_t> slot_lu((size_t)ix->n_slots);
  for (uint32_t b = 0; b < nb; b++) {
    const ull bu = (ull)b << ub;
    int32_t j = j0[b];
    for (uint32_t cidx = 0; cidx < (1u << cb); cidx++) {
      const ull cu = bu + ((ull)cidx << k);
      while (j < j0[b + 1] && slot_u[j] < cu) j++;
      dir[((size_t)b << cb) + cidx] = (uint16_t)(j - j0[b]);
    }
    for (int32_t jj = j0[b]; jj < j0[b + 1]; jj++) slot_lu[jj] = (uint32_t)(slot_u[jj] - bu);
  }
  std::vector<int2> gtab((size_t)std::max(G, 1));
  for (int g = 0; g < G; g++) gtab[g] = make_int2(gsize[g], (int)gbase[g]);
  GTB_TRY(upload_b(ctx, bs->d_gtab, gtab));
  {
    std::vector<int4> pm((size_t)std::max(ix->n_chrom, 1) * 2, make_int4(0, 0, 0, 0));
    const int cp = ix->h_class_of[(uint8_t)'+'], cm = ix->h_class_of[(uint8_t)'-'];
    auto entry = [&](int g) {
      const ull u0 = (ull)gbase[g] << k;
      return make_int4(gsize[g], (int)(u0 & (((ull)1 << ub) - 1)), (int)(u0 >> ub), 0);
    };
    for (int c = 0; c < ix->n_chrom; c++) {
      if (cp >= 0) pm[2 * c] = entry(c * ix->n_class + cp);
      if (cm >= 0) pm[2 * c + 1] = entry(c * ix->n_class + cm);
    }
    GTB_TRY(upload_b(ctx, bs->d_pm, pm));
  }
  GTB_TRY(upload_b(ctx, bs->d_j0, j0));
  GTB_TRY(upload_b(ctx, bs->d_slot_lu, slot_lu));
  GTB_TRY(upload_b(ctx, bs->d_slot_u0, slot_u0));
  GTB_TRY(upload_b(ctx, bs->d_dir, dir));
  GTB_TRY(bs->d_cursor.reserve(ctx, nb));
  GTB_TRY(bs->d_next_page.reserve(ctx, 1));
  GTB_TRY(bs->d_unit_off.reserve(ctx, (size_t)nb + 1));
  GTB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  bs->ready = true; bs->failed = false;
  return GTB_OK;
}

bool gtb_bucket_supported(gtb_index *ix, const QueryView &q, bool batch_multi) {
  if (batch_multi || q.region_offset || q.weight) return false;        // single-interval, unweighted batches
  if (q.n_regions >= ((int64_t)1 << 31)) return false;
  if (gtb_bucket_prepare(ix) != GTB_OK) return false;
  return ix->bucket->part_smem <= ix->ctx->smem_optin && ix->bucket->count_smem <= ix->ctx->smem_optin;
}

int gtb_bucket_accumulate(gtb_index *ix, const QueryView &q) {
  gtb_ctx *ctx = ix->ctx;
  if (gtb_bucket_prepare(ix) != GTB_OK) return gtb_fail(ctx, GTB_ERR_UNSUPPORTED, "bucket engine cannot serve this index");
  gtb_bucket_state *bs = ix->bucket;
  if (q.region_offset || q.weight) return gtb_fail(ctx, GTB_ERR_UNSUPPORTED, "bucket engine takes single-interval, unweighted batches");
  const uint32_t nb = bs->n_buckets;
  const uint64_t n = (uint64_t)q.n_regions;
  BucketView bv;
  bv.k = bs->k; bv.ub = bs->ub; bv.n_chrom = ix->n_chrom; bv.n_class = ix->n_class; bv.n_groups = ix->n_groups;
  bv.cls_plus = ix->h_class_of[(uint8_t)'+']; bv.cls_minus = ix->h_class_of[(uint8_t)'-'];
  bv.class_of = ix->d_class_of.p; bv.chrom_present = ix->d_present.p; bv.gtab = bs->d_gtab.p; bv.pm_tab = bs->d_pm.p;
  bv.n_buckets = nb; bv.pool = nullptr; bv.page_table = nullptr; bv.pt_stride = 0;
  bv.cursor = bs->d_cursor.p; bv.next_page = bs->d_next_page.p; bv.gen = 0;
  bv.line_off = nullptr; bv.sorted_lines = nullptr;
  bv.j0 = bs->d_j0.p; bv.slot_lu = bs->d_slot_lu.p; bv.slot_u0 = bs->d_slot_u0.p; bv.dir = bs->d_dir.p;
  bv.unit_off = bs->d_unit_off.p; bv.max_local = bs->max_local;
  RankView rv;
  rv.n_chrom = ix->n_chrom; rv.n_class = ix->n_class; rv.class_of = ix->d_class_of.p; rv.chrom_present = ix->d_present.p;
  rv.goff = ix->d_goff.p; rv.points = ix->d_points.p; rv.n_slots = ix->n_slots; rv.hist = ix->d_hist.p; rv.err = ix->d_err.p; rv.admission = ix->admission();

  const bool cov = ix->op == GTB_OP_COVERAGE;
  const size_t per_sm = 227 * 1024;
  const unsigned ctas_per_sm = (unsigned)std::max<size_t>(1, std::min<size_t>(4, per_sm / (bs->count_smem + 1024)));

  // ---- write-combining form of pass 1 (default)
  if (bs->wc_ok && !bs->wc_off && bs->wc_smem <= ctx->smem_optin && !getenv("GTB_BUCKET_PAGED")) {
    WcView wv;
    unsigned gridw = 0;
    const bool planned = bs->wc.plan(ctx, q.n_regions, nb, &wv, &gridw) == GTB_OK;   // not OK: more block slots than their ids can name
    if (planned) {
      // skew watchdog: many queries diverted to the general path => go back to the paged form from now on.  Checked at the
      // first batch after a reset (the previous finish has synchronised) and every 64 M queries of a long stream.
      if (bs->wc_queries > 0 && (q.index_base == 0 || bs->wc_queries >= ((int64_t)64 << 20))) {
        ull host_div = 0;
        GTB_CUDA_OK(ctx, cudaMemcpyAsync(&host_div, wv.diverted, sizeof(ull), cudaMemcpyDeviceToHost, ctx->stream));
        GTB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
        if ((int64_t)(host_div - bs->diverted_seen) > bs->wc_queries / 16) bs->wc_off = true;
        bs->diverted_seen = host_div; bs->wc_queries = 0;
      }
    }
    if (planned && !bs->wc_off) {
      bv.pool = wv.pool; bv.line_off = wv.line_off; bv.sorted_lines = wv.sorted_lines;
      const WcQueries wq{q.n_regions, q.chrom, q.start, q.stop, q.strand, q.index_base};
      if (cov) {
        GTB_TRY(wc_partition_launch(ctx, "bucket_partition", wq, OverlapFront<true>{rv, bv}, wv, gridw, bs->wc_smem));
        GTB_TRY((launch_count_pass<true, true>)(ctx, "bucket_coverage", ctas_per_sm, bs->count_smem, bv, rv));
      } else {
        GTB_TRY(wc_partition_launch(ctx, "bucket_partition", wq, OverlapFront<false>{rv, bv}, wv, gridw, bs->wc_smem));
        GTB_TRY((launch_count_pass<false, true>)(ctx, "bucket_count", ctas_per_sm, bs->count_smem, bv, rv));
      }
      bs->wc_queries += q.n_regions;
      if (getenv("GTB_DEBUG_WC")) {                                      // diagnostics: how the batch travelled
        ull host_div = 0; uint32_t blocks = 0;
        cudaStreamSynchronize(ctx->stream);
        cudaMemcpy(&host_div, wv.diverted, sizeof(ull), cudaMemcpyDeviceToHost);
        cudaMemcpy(&blocks, wv.line_off + nb, sizeof(uint32_t), cudaMemcpyDeviceToHost);
        fprintf(stderr, "[gtb wc] queries %lld buckets %u blocks %u (%.1f elements/block) diverted so far %llu\n", (long long)q.n_regions, nb, blocks,
                blocks ? (double)q.n_regions / blocks : 0.0, host_div);
      }
      return gtb_check_launch(ctx);
    }
  }

  // ---- paged form
  const uint32_t pt_stride = (uint32_t)((n + PAGE - 1) / PAGE + 1);
  const uint64_t n_pages = (n + PAGE - 1) / PAGE + nb + 1;
  GTB_TRY(bs->d_pool.reserve(ctx, (size_t)n_pages * PAGE));
  if (n_pages >= (1u << 24) - 2) return gtb_fail(ctx, GTB_ERR_UNSUPPORTED, "batch too large for the bucket engine's page table");
  if (bs->d_page_table.cap < (size_t)nb * pt_stride) { GTB_TRY(bs->d_page_table.reserve(ctx, (size_t)nb * pt_stride)); bs->gen = 255; }
  if (++bs->gen >= 256) {                                               // new table or tag wrap-around: clear once
    GTB_CUDA_OK(ctx, cudaMemsetAsync(bs->d_page_table.p, 0, bs->d_page_table.cap * 4, ctx->stream));
    bs->gen = 1;
  }
  GTB_CUDA_OK(ctx, cudaMemsetAsync(bs->d_cursor.p, 0, (size_t)nb * 4, ctx->stream));
  GTB_CUDA_OK(ctx, cudaMemsetAsync(bs->d_next_page.p, 0, 4, ctx->stream));
  bv.pool = bs->d_pool.p; bv.page_table = bs->d_page_table.p; bv.pt_stride = pt_stride; bv.gen = bs->gen;
  const bool aligned = ((uintptr_t)q.chrom % 16 == 0) && ((uintptr_t)q.start % 16 == 0) && ((uintptr_t)q.stop % 16 == 0) &&
                       ((uintptr_t)q.strand % 16 == 0);          // TMA bulk copies need 16-byte aligned sources
  const int64_t n_tiles = (q.n_regions + PART_TILE - 1) / PART_TILE;
  const unsigned grid1 = (unsigned)std::max<int64_t>(1, std::min<int64_t>((int64_t)ctx->sm_count * PART_CTAS, n_tiles));
#define GTB_PART_LAUNCH(COV, VEC)                                                                                          \
  do {                                                                                                                     \
    auto kern = nb <= PART_THREADS ? bucket_partition_kernel<COV, VEC, 1>                                                  \
                : nb <= 2 * PART_THREADS ? bucket_partition_kernel<COV, VEC, 2>                                            \
                                         : bucket_partition_kernel<COV, VEC, MAX_BUCKETS / PART_THREADS>;                  \
    GTB_CUDA_OK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bs->part_smem));         \
    GTB_LAUNCH(ctx, "bucket_partition", kern, grid1, PART_THREADS, bs->part_smem, q, rv, bv);                              \
  } while (0)
  if (cov) { if (aligned) GTB_PART_LAUNCH(true, 8); else GTB_PART_LAUNCH(true, 1); }
  else { if (aligned) GTB_PART_LAUNCH(false, 8); else GTB_PART_LAUNCH(false, 1); }
#undef GTB_PART_LAUNCH
  GTB_TRY(gtb_check_launch(ctx));
  GTB_LAUNCH(ctx, "bucket_units", bucket_units_kernel, 1, 1024, 0, bv);
  if (cov) {
    GTB_TRY((launch_count_pass<true, false>)(ctx, "bucket_coverage", ctas_per_sm, bs->count_smem, bv, rv));
  } else {
    GTB_TRY((launch_count_pass<false, false>)(ctx, "bucket_count", ctas_per_sm, bs->count_smem, bv, rv));
  }
  return gtb_check_launch(ctx);
}

void gtb_bucket_destroy(gtb_index *ix) {
  gtb_bucket_state *bs = ix->bucket;
  if (!bs) return;
  bs->d_gtab.release(); bs->d_pm.release(); bs->d_j0.release(); bs->d_slot_lu.release(); bs->d_slot_u0.release(); bs->d_dir.release();
  bs->d_pool.release(); bs->d_page_table.release(); bs->d_cursor.release(); bs->d_next_page.release(); bs->d_unit_off.release();
  bs->wc.release();
  delete bs;
  ix->bucket = nullptr;
}
