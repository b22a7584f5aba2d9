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
constexpr int COUNT_THREADS = 256;
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
  // write-combining form of pass 1 (wc_partition_kernel): elements leave shared memory as full 64-byte lines, eight lines of one
  // bucket to a 512-byte BLOCK, into block slots the CTA owns outright (CTA c owns slots [c * lines_per_cta, (c + 1) *
  // lines_per_cta)) -- no global reservation at all.  ("line_*" below counts blocks.)
  uint32_t lines_per_cta;
  uint32_t *line_info;              // [grid * lines_per_cta] bucket | (elements in the block - 1) << 16
  uint32_t *cta_lines;              // [grid] block slots used by each CTA
  uint32_t *n_lines;                // [n_buckets] blocks per bucket (from pass 1)
  uint32_t *line_off;               // [n_buckets + 1] exclusive scan of n_lines
  uint32_t *line_cursor;            // [n_buckets] scatter cursors
  uint32_t *sorted_lines;           // [total blocks] block slot | (elements - 1) << 25, grouped by bucket
  unsigned long long *diverted;     // queries that found their bucket's ring full and took the general path
};

__device__ __forceinline__ uint4 ldg_stream128(const uint4 *p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void red_add64(ull *p, ull v) { asm volatile("red.global.add.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// ---- TMA bulk copy (global -> shared) with mbarrier completion, sm_90+/sm_100 PTX -----------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}"
      :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// The rare queries an element cannot describe: admission checks of the reference, then the general
// rank step (binary search in global memory).
template <bool COVERAGE>
__device__ __noinline__ void special_query(const BucketView &bv, const RankView &rv, int32_t c, int32_t qs, int32_t qe, int strand,
                                           int64_t w, int64_t index) {
  if (qe <= 0 || qs > qe) {                                                // fatal only on indexed chromosomes, :5731-5741
    if (bv.chrom_present[c]) report_error(rv.err, index, qe <= 0 ? GTB_ERR_QUERY_STOP_NONPOSITIVE : GTB_ERR_QUERY_START_GT_STOP);
    return;
  }
  const int cls = bv.class_of[(uint8_t)strand];
  if (cls < 0) return;                                                     // no index region carries this strand, :5229
  const int g = c * bv.n_class + cls;
  const int gb = rv.goff[g], ge = rv.goff[g + 1];
  if (ge > gb) rank_item<COVERAGE>(rv, gb, ge, qs, qe, w);
}

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
// pass 1, write-combining form (default when the bucket count allows it)
// ------------------------------------------------------------------------------------------------
// The paged form above spends most of its time between barriers: rank -> scan -> reserve (global atomics) -> stage ->
// copy out, five barriers and ~8 dependent shared-memory accesses per query (profiles/r1_experiments.md).  Here every
// bucket has a 32-element ring in shared memory.  A query costs one shared atomic (its position in the ring, from a word that
// also carries the ring's free space and write position) and one store.  After the round's barrier the thread that owns a bucket
// moves every complete 16-element line of its ring to global memory with four 128-bit stores, into the next line slot of a
// range only this CTA writes, and notes the line's bucket; a tiny kernel then groups the line slots by bucket for pass 2.
// No scan, no staging buffer, no global reservation, two barriers per round.
// A query that finds its ring full (more than ~16-32 queries of one 2 048-query round in one bucket, i.e. heavily skewed input)
// takes the general rank step instead; the host watches the diverted count and goes back to the paged form if it is large.
#ifndef GTB_WC_THREADS
#define GTB_WC_THREADS 512
#endif
constexpr int WC_THREADS = GTB_WC_THREADS;
constexpr int WC_ITEMS = 4;
constexpr int WC_TILE = WC_THREADS * WC_ITEMS;            // 2 048 queries per round
constexpr int WC_LINE = 16;                               // elements per line (64 bytes)
constexpr int WC_BLOCK = 8;                               // lines per block: the unit pass 2 looks up (512 bytes of one bucket)
constexpr int WC_BLOCK_ELEMS = WC_LINE * WC_BLOCK;
constexpr int WC_CAP = 32;                                // ring capacity per bucket
constexpr int WC_STRIDE = 36;                             // words between rings: 144 B keeps 16-byte alignment, spreads owners over all banks
constexpr int WC_MAX_BUCKETS = 512;                       // one owner thread per bucket
#ifndef GTB_WC_STAGES
#define GTB_WC_STAGES 1
#endif
constexpr int WC_STAGES = GTB_WC_STAGES;                  // raw tiles in flight per CTA

// shared-memory accessors on 32-bit shared-window addresses: no generic-address arithmetic in the hot loop
__device__ __forceinline__ uint32_t lds32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t atoms_add32(uint32_t a, uint32_t v) {
  uint32_t r;
  asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(r) : "r"(a), "r"(v) : "memory");
  return r;
}

template <bool COVERAGE>
__global__ void __launch_bounds__(WC_THREADS, 2) wc_partition_kernel(const __grid_constant__ QueryView q, const __grid_constant__ RankView rv,
                                                                             const __grid_constant__ BucketView bv) {
  extern __shared__ __align__(128) uint32_t smem[];
  // raw tiles: WC_STAGES buffers of chrom | start | stop (WC_TILE ints each) | strand (WC_TILE bytes), filled by TMA bulk copies
  constexpr int RAW_WORDS = 3 * WC_TILE + WC_TILE / 4;
  // (ring and word n_buckets are a sink: queries with nothing to insert go there, which keeps the insert step free of branches)
  uint32_t *s_ring = smem + WC_STAGES * RAW_WORDS;                    // [n_buckets + 1][WC_STRIDE]
  uint32_t *s_word = s_ring + (size_t)(bv.n_buckets + 1) * WC_STRIDE; // [n_buckets + 1] count of this round | free << 12 | write position << 18
  uint32_t *s_direct = s_word + ((bv.n_buckets + 4) & ~3u);           // [n_buckets] blocks written straight from registers (below)
  int4 *s_pm = reinterpret_cast<int4 *>(s_direct + ((bv.n_buckets + 3) & ~3u)); // [2 * n_chrom + 2] group table, last two = "no such group"
  __shared__ __align__(8) uint64_t s_bar[WC_STAGES];
  __shared__ uint32_t s_next_line;

  // group entries for the hot path: x = max(largest point, 0) so that one unsigned compare covers 1 <= start <= x;
  // w keeps the signed value for the rare general path
  for (int i = threadIdx.x; i < 2 * bv.n_chrom + 2; i += blockDim.x) {
    int4 g = i < 2 * bv.n_chrom ? bv.pm_tab[i] : make_int4(0, 0, 0, 0);
    g.w = g.x; g.x = max(g.x, 0);
    s_pm[i] = g;
  }
  for (uint32_t i = threadIdx.x; i <= bv.n_buckets; i += blockDim.x) s_word[i] = i < bv.n_buckets ? (uint32_t)WC_CAP << 12 : 0u;   // the sink has no room
  for (uint32_t i = threadIdx.x; i < bv.n_buckets; i += blockDim.x) s_direct[i] = 0;
  if (threadIdx.x == 0) {
    s_next_line = 0;
    for (int st = 0; st < WC_STAGES; st++) mbar_init(&s_bar[st], 1);
    fence_proxy_async();
  }
  __syncthreads();
  const uint32_t a_raw = smem_u32(smem), a_ring = smem_u32(s_ring), a_word = smem_u32(s_word), a_pm = smem_u32(s_pm);
  const int64_t n_tiles = (q.n_regions + WC_TILE - 1) / WC_TILE;
  const bool aligned = ((reinterpret_cast<uintptr_t>(q.chrom) | reinterpret_cast<uintptr_t>(q.start) | reinterpret_cast<uintptr_t>(q.stop) |
                         reinterpret_cast<uintptr_t>(q.strand)) & 15) == 0;
  const int64_t n_full = aligned ? q.n_regions / WC_TILE : 0;         // tiles that TMA can fetch (complete, aligned)
  const uint32_t ub = (uint32_t)bv.ub, ubmask = (1u << ub) - 1u;
  const uint32_t len_max = 0xFFFFFFFFu >> ub;
  const uint32_t n_chrom = (uint32_t)bv.n_chrom;
  constexpr uint32_t TILE_BYTES = WC_TILE * 13;
  const size_t line_base = (size_t)blockIdx.x * bv.lines_per_cta;
  const int lane = threadIdx.x & 31;
  uint32_t diverted = 0;

  auto issue = [&](int64_t tile, int st) {                             // one thread: 4 bulk copies into stage st
    const int64_t first = tile * WC_TILE;
    uint32_t *raw = smem + st * RAW_WORDS;
    mbar_expect_tx(&s_bar[st], TILE_BYTES);
    tma_bulk_g2s(raw, q.chrom + first, WC_TILE * 4, &s_bar[st]);
    tma_bulk_g2s(raw + WC_TILE, q.start + first, WC_TILE * 4, &s_bar[st]);
    tma_bulk_g2s(raw + 2 * WC_TILE, q.stop + first, WC_TILE * 4, &s_bar[st]);
    tma_bulk_g2s(raw + 3 * WC_TILE, q.strand + first, WC_TILE, &s_bar[st]);
  };
  // ---- owner state (thread b owns bucket b): ring occupancy carried over (< WC_LINE), its head (0 or WC_LINE), and the open block
  uint32_t occ = 0, head = 0;
  uint32_t blk = 0xFFFFFFFFu, used = 0, last_fill = WC_LINE, my_blocks = 0;
  const uint32_t a_myring = a_ring + threadIdx.x * (uint32_t)(WC_STRIDE * 4), a_myword = a_word + threadIdx.x * 4u;
  auto close_block = [&]() {
    if (blk != 0xFFFFFFFFu) bv.line_info[line_base + blk] = threadIdx.x | (((used - 1u) * WC_LINE + last_fill - 1u) << 16);
  };
  // one 64-byte line of the owner's ring -> the next line of the bucket's open block (`fresh` replaces a full block)
  auto flush_line = [&](uint32_t pos, uint32_t fresh, uint32_t fill) {
    if (blk == 0xFFFFFFFFu || used == (uint32_t)WC_BLOCK) { close_block(); blk = fresh; used = 0; }
    const uint32_t src = a_myring + pos * 4u;
    const uint4 a0 = lds128(src), a1 = lds128(src + 16), a2 = lds128(src + 32), a3 = lds128(src + 48);
    uint4 *dst = reinterpret_cast<uint4 *>(bv.pool) + ((line_base + blk) * WC_BLOCK + used) * (WC_LINE / 4);
    dst[0] = a0; dst[1] = a1; dst[2] = a2; dst[3] = a3;
    used++; last_fill = fill;
  };
  // block slots for a whole warp of owners with ONE shared atomic (a same-address atomic per owner costs far more than the copy)
  auto warp_slots = [&](bool mine) -> uint32_t {
    const uint32_t mask = __ballot_sync(0xffffffffu, mine);
    uint32_t base = 0;
    if (lane == 0 && mask) base = atomicAdd(&s_next_line, (uint32_t)__popc(mask));
    base = __shfl_sync(0xffffffffu, base, 0);
    return base + (uint32_t)__popc(mask & ((1u << lane) - 1u));
  };
  // everything the fast classification cannot decide (rare): the reference's admission rules, then the general rank step
  auto general = [&](int32_t c, int32_t s, int32_t e, uint32_t sbyte, int64_t index) {
    if ((uint32_t)c >= n_chrom) return;                                            // chromosome the index has never seen
    const uint32_t d = sbyte - (uint32_t)'+';
    const bool addressable = (d & ~2u) == 0;
    const int gx = addressable ? bv.pm_tab[2 * c + (int)(d >> 1)].x : 0;
    const bool valid = s >= 1 && s <= e;
    const bool nothing = addressable && valid && gx >= 0 && (gx == 0 || s > gx);  // no points in the group / start beyond the last one
    if (!nothing) special_query<COVERAGE>(bv, rv, c, s, e, (int)(int8_t)sbyte, 1, index);
  };

  if (threadIdx.x == 0)
    for (int st = 0; st < WC_STAGES; st++)
      if ((int64_t)blockIdx.x + (int64_t)st * gridDim.x < n_full) issue(blockIdx.x + (int64_t)st * gridDim.x, st);
  uint32_t round = 0;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, round++) {
    // ---- load + classify: touches no shared structure that the previous round's owners may still be updating
    int32_t c[WC_ITEMS], s[WC_ITEMS], e[WC_ITEMS];
    uint32_t stw;
    const int stage = (int)(round % WC_STAGES);
    if (tile < n_full) {
      mbar_wait(&s_bar[stage], (round / WC_STAGES) & 1u);
      const uint32_t a = a_raw + (uint32_t)stage * (RAW_WORDS * 4) + threadIdx.x * 16u;
      const uint4 c0 = lds128(a), s0 = lds128(a + WC_TILE * 4), e0 = lds128(a + 2 * WC_TILE * 4);
      stw = lds32(a_raw + (uint32_t)stage * (RAW_WORDS * 4) + 3 * WC_TILE * 4 + threadIdx.x * 4u);
      c[0] = (int)c0.x; c[1] = (int)c0.y; c[2] = (int)c0.z; c[3] = (int)c0.w; s[0] = (int)s0.x; s[1] = (int)s0.y; s[2] = (int)s0.z; s[3] = (int)s0.w;
      e[0] = (int)e0.x; e[1] = (int)e0.y; e[2] = (int)e0.z; e[3] = (int)e0.w;
    } else {
      const int64_t first = tile * WC_TILE + (int64_t)threadIdx.x * WC_ITEMS;
      stw = 0;
#pragma unroll
      for (int i = 0; i < WC_ITEMS; i++) {
        const int64_t r = first + i;
        const bool ok = r < q.n_regions;
        c[i] = ok ? q.chrom[r] : -1; s[i] = ok ? q.start[r] : 1; e[i] = ok ? q.stop[r] : 1;
        stw |= (ok ? (unsigned)(uint8_t)q.strand[r] : (unsigned)'+') << (i * 8);
      }
    }
    // strands: '+' = 0x2B, '-' = 0x2D.  xw has 0x00 / 0x06 in the bytes of '+' / '-' queries; any other bit: a strand the fast path does not take
    const uint32_t xw = stw ^ 0x2B2B2B2Bu;
    uint32_t elem[WC_ITEMS], bk[WC_ITEMS];                             // bk = bucket, or 0xFFFFFFFF: nothing to insert
    uint4 g[WC_ITEMS];
#pragma unroll
    for (int i = 0; i < WC_ITEMS; i++) {
      const uint32_t idx = 2u * min((uint32_t)c[i], n_chrom) + ((xw >> (8 * i + 1)) & 1u);      // unknown chromosome -> an empty entry
      g[i] = lds128(a_pm + idx * 16u);
    }
    uint32_t slow = (xw & 0xF9F9F9F9u) ? 0xFu : 0u;                    // bit i: item i goes through the exact general decision
#pragma unroll
    for (int i = 0; i < WC_ITEMS; i++) {
      const uint32_t t = g[i].y + (uint32_t)s[i];
      const uint32_t lu = t & ubmask;
      const uint32_t len = (uint32_t)(min(e[i], (int)g[i].x + 1) - s[i]);           // e < s wraps to a huge value and fails the next test
      const bool ok = (uint32_t)(s[i] - 1) < g[i].x && len <= len_max && lu + len <= ubmask;
      elem[i] = lu | (len << ub);
      bk[i] = ok ? g[i].z + (t >> ub) : 0xFFFFFFFFu;
      slow |= ok ? 0u : (1u << i);
    }
    if (slow) {
#pragma unroll
      for (int i = 0; i < WC_ITEMS; i++)
        if ((slow >> i) & 1u) {
          const uint32_t sbyte = (stw >> (8 * i)) & 0xFFu;
          const bool fast_ok = bk[i] != 0xFFFFFFFFu && ((sbyte - (uint32_t)'+') & ~2u) == 0;   // only flagged because a sibling has an odd strand
          if (!fast_ok) {
            bk[i] = 0xFFFFFFFFu;
            general(c[i], s[i], e[i], sbyte, q.index_base + tile * WC_TILE + (int64_t)threadIdx.x * WC_ITEMS + i);
          }
        }
    }
    __syncthreads();                    // B1: raw tile consumed by everybody; ring words of the previous round are final
    if (threadIdx.x == 0 && tile + (int64_t)WC_STAGES * gridDim.x < n_full) { fence_proxy_async(); issue(tile + (int64_t)WC_STAGES * gridDim.x, stage); }

    // ---- insert: one shared atomic and one store per query; the four atomics are issued back to back
#ifdef GTB_WC_FRONT_ONLY
#pragma unroll
    for (int i = 0; i < WC_ITEMS; i++) diverted += (elem[i] ^ bk[i]) & 1u;      // timing experiment: front end only
    continue;
#endif
    // Position-sorted input (what -S promises, and what aligners emit): the 128 queries of a warp fall into one bucket.  They
    // would overfill its ring at once, and they need no combining either: the warp writes them as one full 512-byte block.
    {
      const uint32_t lead = __shfl_sync(0xffffffffu, bk[0], 0);
      const bool same = lead != 0xFFFFFFFFu && bk[0] == lead && bk[1] == lead && bk[2] == lead && bk[3] == lead;
      if (__all_sync(0xffffffffu, same)) {
        uint32_t slot = 0;
        if (lane == 0) {
          slot = atomicAdd(&s_next_line, 1u);
          atomicAdd(&s_direct[lead], 1u);
          bv.line_info[line_base + slot] = lead | ((uint32_t)(WC_BLOCK_ELEMS - 1) << 16);
        }
        slot = __shfl_sync(0xffffffffu, slot, 0);
        reinterpret_cast<uint4 *>(bv.pool)[(line_base + slot) * (WC_BLOCK_ELEMS / 4) + lane] = make_uint4(elem[0], elem[1], elem[2], elem[3]);
#pragma unroll
        for (int i = 0; i < WC_ITEMS; i++) bk[i] = 0xFFFFFFFFu;        // nothing left for the rings
      }
    }
    uint32_t w[WC_ITEMS], bx[WC_ITEMS];
#pragma unroll
    for (int i = 0; i < WC_ITEMS; i++) {
      bx[i] = min(bk[i], bv.n_buckets);                                // nothing to insert -> the sink
      w[i] = atoms_add32(a_word + bx[i] * 4u, 1u);
    }
    slow = 0;
#pragma unroll
    for (int i = 0; i < WC_ITEMS; i++) {
      const uint32_t cnt = w[i] & 0xFFFu, free_ = (w[i] >> 12) & 0x3Fu, wp = (w[i] >> 18) & (uint32_t)(WC_CAP - 1);
      const bool fits = cnt < free_;
      // a query that does not fit its ring stores into the sink's ring instead (and then takes the general step below)
      sts32(a_ring + ((fits ? bx[i] : bv.n_buckets) * (uint32_t)WC_STRIDE + ((wp + cnt) & (uint32_t)(WC_CAP - 1))) * 4u, elem[i]);
      slow |= (bk[i] != 0xFFFFFFFFu && !fits) ? (1u << i) : 0u;
    }
    if (slow) {                         // ring full: general step (exact, slow; skewed input only)
#pragma unroll
      for (int i = 0; i < WC_ITEMS; i++)
        if ((slow >> i) & 1u) {
          diverted++;
          special_query<COVERAGE>(bv, rv, c[i], s[i], e[i], (int)(int8_t)((stw >> (i * 8)) & 0xFFu), 1,
                                  q.index_base + tile * WC_TILE + (int64_t)threadIdx.x * WC_ITEMS + i);
        }
    }
    __syncthreads();                    // B2: all elements of the round are in the rings

    // ---- owners: move complete lines out, publish the ring state for the next round
    if ((threadIdx.x & ~31u) < bv.n_buckets) {                        // warp-uniform: the warps that hold owners
      const bool owner = threadIdx.x < bv.n_buckets;
      const uint32_t cnt = owner ? lds32(a_myword) & 0xFFFu : 0u;
      occ += min(cnt, (uint32_t)WC_CAP - occ);                        // what the inserters were allowed to store
      const uint32_t nl = occ / WC_LINE;                              // 0, 1 or 2 complete lines
      const bool want_block = nl && (blk == 0xFFFFFFFFu || used + nl > (uint32_t)WC_BLOCK);
      const uint32_t fresh = warp_slots(want_block);
      my_blocks += want_block ? 1u : 0u;
      for (uint32_t l = 0; l < nl; l++) { flush_line(head, fresh, WC_LINE); head ^= (uint32_t)WC_LINE; }
      occ &= (uint32_t)(WC_LINE - 1);
      if (cnt) sts32(a_myword, (((uint32_t)WC_CAP - occ) << 12) | (((head + occ) & (uint32_t)(WC_CAP - 1)) << 18));
    }
  }
  __syncthreads();
  // ---- the rings' remainders leave as partial lines
  if ((threadIdx.x & ~31u) < bv.n_buckets) {
    const bool owner = threadIdx.x < bv.n_buckets;
    const bool rest = owner && occ != 0;
    const bool want_block = rest && (blk == 0xFFFFFFFFu || used == (uint32_t)WC_BLOCK);
    const uint32_t fresh = warp_slots(want_block);
    my_blocks += want_block ? 1u : 0u;
    if (rest) flush_line(head, fresh, occ);
    if (owner) { close_block(); my_blocks += s_direct[threadIdx.x]; }
    if (my_blocks) atomicAdd(bv.n_lines + threadIdx.x, my_blocks);
  }
#ifdef GTB_WC_FRONT_ONLY
  if (diverted == 0x7FFFFFF1u) atomicAdd(bv.diverted, 1ull);
#else
  if (diverted) atomicAdd(bv.diverted, (unsigned long long)diverted);
#endif
  __syncthreads();
  if (threadIdx.x == 0) bv.cta_lines[blockIdx.x] = s_next_line;
}

// exclusive scan of the per-bucket line counts (one CTA) and reset of the scatter cursors
__global__ void __launch_bounds__(1024) wc_line_offsets_kernel(BucketView bv) {
  __shared__ uint32_t s_tot[32];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint32_t base = 0; base < bv.n_buckets; base += blockDim.x) {
    const uint32_t b = base + threadIdx.x;
    const uint32_t u = b < bv.n_buckets ? bv.n_lines[b] : 0;
    uint32_t inc = u;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
    if (lane == 31) s_tot[warp] = inc;
    __syncthreads();
    uint32_t pre = carry;
    for (int w = 0; w < warp; w++) pre += s_tot[w];
    if (b < bv.n_buckets) { bv.line_off[b] = pre + inc - u; bv.line_cursor[b] = 0; }
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = pre + inc;
    __syncthreads();
  }
  if (threadIdx.x == 0) bv.line_off[bv.n_buckets] = carry;
}

// groups the line slots by bucket: CTA c walks the slots CTA c of pass 1 filled (same grid)
__global__ void __launch_bounds__(512) wc_line_scatter_kernel(BucketView bv) {
  __shared__ uint32_t s_cnt[WC_MAX_BUCKETS], s_base[WC_MAX_BUCKETS];
  for (uint32_t i = threadIdx.x; i < bv.n_buckets; i += blockDim.x) s_cnt[i] = 0;
  __syncthreads();
  const uint32_t n = bv.cta_lines[blockIdx.x];
  const size_t first = (size_t)blockIdx.x * bv.lines_per_cta;
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) atomicAdd(&s_cnt[bv.line_info[first + i] & 0xFFFFu], 1u);
  __syncthreads();
  for (uint32_t b = threadIdx.x; b < bv.n_buckets; b += blockDim.x) {
    const uint32_t c = s_cnt[b];
    s_base[b] = bv.line_off[b] + (c ? atomicAdd(bv.line_cursor + b, c) : 0u);
    s_cnt[b] = 0;
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
    const uint32_t info = bv.line_info[first + i];
    const uint32_t b = info & 0xFFFFu;
    const uint32_t pos = s_base[b] + atomicAdd(&s_cnt[b], 1u);
    bv.sorted_lines[pos] = (uint32_t)(first + i) | ((info >> 16) << 25);
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
template <bool COVERAGE, bool LINES>
__global__ void __launch_bounds__(COUNT_THREADS, 4) bucket_count_kernel(BucketView bv, RankView rv) {
  extern __shared__ __align__(16) uint32_t smem[];
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
            const uint32_t entry = __ldg(bv.sorted_lines + u + (v4 >> 5));                       // 32 loads of 128 bits per block
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
#pragma unroll
      for (int r = 0; r < UNROLL; r++) {
        const uint32_t el[4] = {d[r].x, d[r].y, d[r].z, d[r].w};
        uint32_t us[4], ue[4], jS[4], pS[4];
#pragma unroll
        for (int i = 0; i < 4; i++) { us[i] = el[i] & ubmask; ue[i] = us[i] + (el[i] >> bv.ub); jS[i] = s_dir[us[i] >> kmaskless]; }
#pragma unroll
        for (int i = 0; i < 4; i++) pS[i] = s_pts[jS[i]];
#pragma unroll
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
              if (jS[i] == jE) atomicAdd(&s_h64[jS[i]], (ull)(ue[i] - us[i]) + 1ull);
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
  dbuf<uint32_t> d_line_info, d_cta_lines, d_n_lines, d_line_off, d_line_cursor, d_sorted_lines;
  dbuf<ull> d_diverted;
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
  bs->wc_smem = (size_t)WC_STAGES * WC_TILE * 13 + (size_t)(nb + 1) * WC_STRIDE * 4 + (size_t)(nb4 + 4) * 8 + (size_t)std::max(ix->n_chrom, 1) * 32 + 32 + 64;
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
  GTB_TRY(bs->d_n_lines.reserve(ctx, (size_t)nb + 1));
  GTB_TRY(bs->d_line_off.reserve(ctx, (size_t)nb + 1));
  GTB_TRY(bs->d_line_cursor.reserve(ctx, (size_t)nb + 1));
  GTB_TRY(bs->d_diverted.reserve(ctx, 1));
  GTB_CUDA_OK(ctx, cudaMemsetAsync(bs->d_diverted.p, 0, sizeof(ull), ctx->stream));
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
  bv.lines_per_cta = 0; bv.line_info = nullptr; bv.cta_lines = nullptr; bv.n_lines = nullptr; bv.line_off = nullptr;
  bv.line_cursor = nullptr; bv.sorted_lines = nullptr; bv.diverted = bs->d_diverted.p;
  bv.j0 = bs->d_j0.p; bv.slot_lu = bs->d_slot_lu.p; bv.slot_u0 = bs->d_slot_u0.p; bv.dir = bs->d_dir.p;
  bv.unit_off = bs->d_unit_off.p; bv.max_local = bs->max_local;
  RankView rv;
  rv.n_chrom = ix->n_chrom; rv.n_class = ix->n_class; rv.class_of = ix->d_class_of.p; rv.chrom_present = ix->d_present.p;
  rv.goff = ix->d_goff.p; rv.points = ix->d_points.p; rv.n_slots = ix->n_slots; rv.hist = ix->d_hist.p; rv.err = ix->d_err.p;

  const bool cov = ix->op == GTB_OP_COVERAGE;
  const size_t per_sm = 227 * 1024;
  const unsigned ctas_per_sm = (unsigned)std::max<size_t>(1, std::min<size_t>(4, per_sm / (bs->count_smem + 1024)));
  const unsigned grid2 = (unsigned)ctx->sm_count * ctas_per_sm;

  // ---- write-combining form of pass 1 (default)
  if (bs->wc_ok && !bs->wc_off && bs->wc_smem <= ctx->smem_optin && !getenv("GTB_BUCKET_PAGED")) {
    // skew watchdog: many queries diverted to the general path => go back to the paged form from now on.  Checked at the
    // first batch after a reset (the previous finish has synchronised) and every 64 M queries of a long stream.
    if (bs->wc_queries > 0 && (q.index_base == 0 || bs->wc_queries >= ((int64_t)64 << 20))) {
      ull host_div = 0;
      GTB_CUDA_OK(ctx, cudaMemcpyAsync(&host_div, bs->d_diverted.p, sizeof(ull), cudaMemcpyDeviceToHost, ctx->stream));
      GTB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
      if ((int64_t)(host_div - bs->diverted_seen) > bs->wc_queries / 16) bs->wc_off = true;
      bs->diverted_seen = host_div; bs->wc_queries = 0;
    }
    const int64_t wc_tiles = (q.n_regions + WC_TILE - 1) / WC_TILE;
    const unsigned gridw = (unsigned)std::max<int64_t>(1, std::min<int64_t>((int64_t)ctx->sm_count * 2, wc_tiles));
    const uint64_t tiles_cta = ((uint64_t)wc_tiles + gridw - 1) / gridw;
    const uint64_t lines_per_cta = tiles_cta * (WC_TILE / WC_BLOCK_ELEMS) + nb + 1;          // block slots: full blocks + one open block per bucket
    const uint64_t total_slots = lines_per_cta * gridw;
    if (!bs->wc_off && total_slots < ((uint64_t)1 << 25)) {
      GTB_TRY(bs->d_pool.reserve(ctx, (size_t)total_slots * WC_BLOCK_ELEMS));
      GTB_TRY(bs->d_line_info.reserve(ctx, (size_t)total_slots));
      GTB_TRY(bs->d_sorted_lines.reserve(ctx, (size_t)total_slots));
      GTB_TRY(bs->d_cta_lines.reserve(ctx, (size_t)gridw));
      GTB_CUDA_OK(ctx, cudaMemsetAsync(bs->d_n_lines.p, 0, (size_t)nb * 4, ctx->stream));
      bv.lines_per_cta = (uint32_t)lines_per_cta; bv.line_info = bs->d_line_info.p; bv.cta_lines = bs->d_cta_lines.p;
      bv.n_lines = bs->d_n_lines.p; bv.line_off = bs->d_line_off.p; bv.line_cursor = bs->d_line_cursor.p;
      bv.sorted_lines = bs->d_sorted_lines.p; bv.diverted = bs->d_diverted.p; bv.pool = bs->d_pool.p;
      if (cov) {
        GTB_CUDA_OK(ctx, cudaFuncSetAttribute(wc_partition_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bs->wc_smem));
        GTB_LAUNCH(ctx, "bucket_partition", wc_partition_kernel<true>, gridw, WC_THREADS, bs->wc_smem, q, rv, bv);
      } else {
        GTB_CUDA_OK(ctx, cudaFuncSetAttribute(wc_partition_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bs->wc_smem));
        GTB_LAUNCH(ctx, "bucket_partition", wc_partition_kernel<false>, gridw, WC_THREADS, bs->wc_smem, q, rv, bv);
      }
      GTB_TRY(gtb_check_launch(ctx));
      GTB_LAUNCH(ctx, "bucket_line_offsets", wc_line_offsets_kernel, 1, 1024, 0, bv);
      GTB_LAUNCH(ctx, "bucket_line_scatter", wc_line_scatter_kernel, gridw, 512, 0, bv);
      if (cov) {
        GTB_CUDA_OK(ctx, cudaFuncSetAttribute(bucket_count_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bs->count_smem));
        GTB_LAUNCH(ctx, "bucket_coverage", (bucket_count_kernel<true, true>), grid2, COUNT_THREADS, bs->count_smem, bv, rv);
      } else {
        GTB_CUDA_OK(ctx, cudaFuncSetAttribute(bucket_count_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bs->count_smem));
        GTB_LAUNCH(ctx, "bucket_count", (bucket_count_kernel<false, true>), grid2, COUNT_THREADS, bs->count_smem, bv, rv);
      }
      bs->wc_queries += q.n_regions;
      if (getenv("GTB_DEBUG_WC")) {                                      // diagnostics: how the batch travelled
        ull host_div = 0; uint32_t blocks = 0;
        cudaStreamSynchronize(ctx->stream);
        cudaMemcpy(&host_div, bs->d_diverted.p, sizeof(ull), cudaMemcpyDeviceToHost);
        cudaMemcpy(&blocks, bs->d_line_off.p + nb, sizeof(uint32_t), cudaMemcpyDeviceToHost);
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
    GTB_CUDA_OK(ctx, cudaFuncSetAttribute(bucket_count_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bs->count_smem));
    GTB_LAUNCH(ctx, "bucket_coverage", (bucket_count_kernel<true, false>), grid2, COUNT_THREADS, bs->count_smem, bv, rv);
  } else {
    GTB_CUDA_OK(ctx, cudaFuncSetAttribute(bucket_count_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bs->count_smem));
    GTB_LAUNCH(ctx, "bucket_count", (bucket_count_kernel<false, false>), grid2, COUNT_THREADS, bs->count_smem, bv, rv);
  }
  return gtb_check_launch(ctx);
}

void gtb_bucket_destroy(gtb_index *ix) {
  gtb_bucket_state *bs = ix->bucket;
  if (!bs) return;
  bs->d_gtab.release(); bs->d_pm.release(); bs->d_j0.release(); bs->d_slot_lu.release(); bs->d_slot_u0.release(); bs->d_dir.release();
  bs->d_pool.release(); bs->d_page_table.release(); bs->d_cursor.release(); bs->d_next_page.release(); bs->d_unit_off.release();
  bs->d_line_info.release(); bs->d_cta_lines.release(); bs->d_n_lines.release(); bs->d_line_off.release(); bs->d_line_cursor.release();
  bs->d_sorted_lines.release(); bs->d_diverted.release();
  delete bs;
  ix->bucket = nullptr;
}
