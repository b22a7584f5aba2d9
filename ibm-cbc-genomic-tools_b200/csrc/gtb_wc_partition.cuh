// gtb_wc_partition.cuh -- write-combining partition of a query batch into genome buckets (pass 1 of the BUCKET overlap
// engine, gtb_bucket.cu, and of the bucketed window-count engine, gtb_scan.cu).
//
// Every bucket has a 32-element ring in shared memory.  A query costs one shared atomic (its position in the ring, from a word
// that also carries the ring's free space and write position) and one store.  After the round's barrier the thread that owns a
// bucket moves every complete 16-element line of its ring to global memory with four 128-bit stores, eight lines of one bucket
// to a 512-byte BLOCK, into block slots the CTA owns outright (CTA c owns slots [c * lines_per_cta, (c + 1) * lines_per_cta))
// -- no scan, no staging buffer, no global reservation, two barriers per round.  A tiny kernel pair then groups the block slots
// by bucket for pass 2.  The raw SoA tile (13 bytes per query) arrives by TMA bulk copies (cp.async.bulk + mbarrier).
//
// What a query means is the FRONT's business (a policy struct passed by value):
//   table_size() / table_entry(i)        16-byte entries staged in shared memory, looked up once per query
//   table_index(c, xw, i)                entry of item i (xw = the four strand bytes ^ 0x2B2B2B2B: '+' -> 0x00, '-' -> 0x06)
//   slow_mask(xw)                        items that must take the exact decision whatever classify() says
//   classify(g, s, e, elem)              fast decision: bucket (or WC_NONE) and the 32-bit element that goes there
//   FAIL_IS_SLOW                         whether a query classify() rejects needs resolve()
//   resolve(bucket, c, s, e, sbyte, index)   exact decision for flagged items: the bucket that stands, or WC_NONE after doing
//                                        whatever the query needs (errors, general path)
//   divert(c, s, e, sbyte, index)        a query that found its bucket's ring full (heavily skewed input): count it some other way
#pragma once
#include "gtb_internal.cuh"
#include <algorithm>

namespace {

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
constexpr int WC_MAX_BUCKETS = WC_THREADS;                // one owner thread per bucket
constexpr int WC_CTAS_PER_SM = WC_THREADS > 512 ? 1 : 2;  // 1 024 threads per SM either way
#ifndef GTB_WC_STAGES
#define GTB_WC_STAGES 1
#endif
constexpr int WC_STAGES = GTB_WC_STAGES;                  // raw tiles in flight per CTA
constexpr uint32_t WC_NONE = 0xFFFFFFFFu;

struct WcQueries {                 // the raw SoA batch (device pointers)
  int64_t n;
  const int32_t *chrom, *start, *stop;
  const int8_t *strand;
  int64_t index_base;              // stream-order index of query 0 (for error reports)
};

struct WcView {
  uint32_t n_buckets;
  uint32_t *pool;                   // block slots of WC_BLOCK_ELEMS elements
  uint32_t lines_per_cta;           // block slots per CTA of pass 1
  uint32_t *line_info;              // [grid * lines_per_cta] bucket | (elements in the block - 1) << 16
  uint32_t *cta_lines;              // [grid] block slots used by each CTA
  uint32_t *n_lines;                // [n_buckets] blocks per bucket (from pass 1)
  uint32_t *line_off;               // [n_buckets + 1] exclusive scan of n_lines
  uint32_t *line_cursor;            // [n_buckets] scatter cursors
  uint32_t *sorted_lines;           // [total blocks] block slot | (elements - 1) << 25, grouped by bucket
  unsigned long long *diverted;     // queries that found their bucket's ring full
};

__device__ __forceinline__ uint4 ldg_stream128(const uint4 *p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
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

// shared-memory accessors on 32-bit shared-window addresses: no generic-address arithmetic in the hot loop
__device__ __forceinline__ uint32_t lds32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts128(uint32_t a, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t atoms_add32(uint32_t a, uint32_t v) {
  uint32_t r;
  asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(r) : "r"(a), "r"(v) : "memory");
  return r;
}

// dynamic shared memory the kernel needs for n_buckets buckets and a front table of table_entries 16-byte entries
inline size_t wc_smem_bytes(uint32_t n_buckets, size_t table_entries) {
  const uint32_t nb4 = (n_buckets + 3) & ~3u;
  return (size_t)WC_STAGES * WC_TILE * 13 + (size_t)(n_buckets + 1) * WC_STRIDE * 4 + (size_t)(nb4 + 4) * 8 + table_entries * 16 + 64;
}

template <class Front>
__global__ void __launch_bounds__(WC_THREADS, WC_CTAS_PER_SM) wc_partition_kernel(const __grid_constant__ WcQueries q, const __grid_constant__ Front front,
                                                                     const __grid_constant__ WcView wv) {
  extern __shared__ __align__(128) uint32_t smem[];
  // raw tiles: WC_STAGES buffers of chrom | start | stop (WC_TILE ints each) | strand (WC_TILE bytes), filled by TMA bulk copies
  constexpr int RAW_WORDS = 3 * WC_TILE + WC_TILE / 4;
  // (ring and word n_buckets are a sink: queries with nothing to insert go there, which keeps the insert step free of branches)
  uint32_t *s_ring = smem + WC_STAGES * RAW_WORDS;                    // [n_buckets + 1][WC_STRIDE]
  uint32_t *s_word = s_ring + (size_t)(wv.n_buckets + 1) * WC_STRIDE; // [n_buckets + 1] count of this round | free << 12 | write position << 18
  uint32_t *s_direct = s_word + ((wv.n_buckets + 4) & ~3u);           // [n_buckets] blocks written straight from registers (below)
  uint4 *s_tab = reinterpret_cast<uint4 *>(s_direct + ((wv.n_buckets + 3) & ~3u));   // the front's table
  __shared__ __align__(8) uint64_t s_bar[WC_STAGES];
  __shared__ uint32_t s_next_line;

  for (uint32_t i = threadIdx.x; i < front.table_size(); i += blockDim.x) s_tab[i] = front.table_entry(i);
  for (uint32_t i = threadIdx.x; i <= wv.n_buckets; i += blockDim.x) s_word[i] = i < wv.n_buckets ? (uint32_t)WC_CAP << 12 : 0u;   // the sink has no room
  for (uint32_t i = threadIdx.x; i < wv.n_buckets; i += blockDim.x) s_direct[i] = 0;
  if (threadIdx.x == 0) {
    s_next_line = 0;
    for (int st = 0; st < WC_STAGES; st++) mbar_init(&s_bar[st], 1);
    fence_proxy_async();
  }
  __syncthreads();
  const uint32_t a_raw = smem_u32(smem), a_ring = smem_u32(s_ring), a_word = smem_u32(s_word), a_tab = smem_u32(s_tab);
  const int64_t n_tiles = (q.n + WC_TILE - 1) / WC_TILE;
  const bool aligned = ((reinterpret_cast<uintptr_t>(q.chrom) | reinterpret_cast<uintptr_t>(q.start) | reinterpret_cast<uintptr_t>(q.stop) |
                         reinterpret_cast<uintptr_t>(q.strand)) & 15) == 0;
  const int64_t n_full = aligned ? q.n / WC_TILE : 0;                 // tiles that TMA can fetch (complete, aligned)
  constexpr uint32_t TILE_BYTES = WC_TILE * 13;
  const size_t line_base = (size_t)blockIdx.x * wv.lines_per_cta;
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
  uint32_t blk = WC_NONE, used = 0, last_fill = WC_LINE, my_blocks = 0;
  const uint32_t a_myring = a_ring + threadIdx.x * (uint32_t)(WC_STRIDE * 4), a_myword = a_word + threadIdx.x * 4u;
  auto close_block = [&]() {
    if (blk != WC_NONE) wv.line_info[line_base + blk] = threadIdx.x | (((used - 1u) * WC_LINE + last_fill - 1u) << 16);
  };
  // one 64-byte line of the owner's ring -> the next line of the bucket's open block (`fresh` replaces a full block)
  auto flush_line = [&](uint32_t pos, uint32_t fresh, uint32_t fill) {
    if (blk == WC_NONE || used == (uint32_t)WC_BLOCK) { close_block(); blk = fresh; used = 0; }
    const uint32_t src = a_myring + pos * 4u;
    const uint4 a0 = lds128(src), a1 = lds128(src + 16), a2 = lds128(src + 32), a3 = lds128(src + 48);
    uint4 *dst = reinterpret_cast<uint4 *>(wv.pool) + ((line_base + blk) * WC_BLOCK + used) * (WC_LINE / 4);
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
        const bool ok = r < q.n;
        c[i] = ok ? q.chrom[r] : -1; s[i] = ok ? q.start[r] : 1; e[i] = ok ? q.stop[r] : 1;
        stw |= (ok ? (unsigned)(uint8_t)q.strand[r] : (unsigned)'+') << (i * 8);
      }
    }
    // strands: '+' = 0x2B, '-' = 0x2D.  xw has 0x00 / 0x06 in the bytes of '+' / '-' queries
    const uint32_t xw = stw ^ 0x2B2B2B2Bu;
    uint32_t elem[WC_ITEMS], bk[WC_ITEMS];                             // bk = bucket, or WC_NONE: nothing to insert
    uint4 g[WC_ITEMS];
#pragma unroll
    for (int i = 0; i < WC_ITEMS; i++) g[i] = lds128(a_tab + front.table_index(c[i], xw, i) * 16u);
    uint32_t slow = front.slow_mask(xw);                               // bit i: item i goes through the exact decision
#pragma unroll
    for (int i = 0; i < WC_ITEMS; i++) {
      bk[i] = front.classify(g[i], s[i], e[i], elem[i]);
      if (Front::FAIL_IS_SLOW) slow |= bk[i] != WC_NONE ? 0u : (1u << i);
    }
    if (slow) {
#pragma unroll
      for (int i = 0; i < WC_ITEMS; i++)
        if ((slow >> i) & 1u)
          bk[i] = front.resolve(bk[i], c[i], s[i], e[i], (stw >> (8 * i)) & 0xFFu, q.index_base + tile * WC_TILE + (int64_t)threadIdx.x * WC_ITEMS + i);
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
    const uint32_t lead = __shfl_sync(0xffffffffu, bk[0], 0);
    const bool same = lead != WC_NONE && bk[0] == lead && bk[1] == lead && bk[2] == lead && bk[3] == lead;
    if (__all_sync(0xffffffffu, same)) {
      uint32_t slot = 0;
      if (lane == 0) {
        slot = atomicAdd(&s_next_line, 1u);
        atomicAdd(&s_direct[lead], 1u);
        wv.line_info[line_base + slot] = lead | ((uint32_t)(WC_BLOCK_ELEMS - 1) << 16);
      }
      slot = __shfl_sync(0xffffffffu, slot, 0);
      reinterpret_cast<uint4 *>(wv.pool)[(line_base + slot) * (WC_BLOCK_ELEMS / 4) + lane] = make_uint4(elem[0], elem[1], elem[2], elem[3]);
    } else {
      uint32_t w[WC_ITEMS], bx[WC_ITEMS];
#pragma unroll
      for (int i = 0; i < WC_ITEMS; i++) {
        bx[i] = min(bk[i], wv.n_buckets);                              // nothing to insert -> the sink
        w[i] = atoms_add32(a_word + bx[i] * 4u, 1u);
      }
      slow = 0;
#pragma unroll
      for (int i = 0; i < WC_ITEMS; i++) {
        const uint32_t cnt = w[i] & 0xFFFu, free_ = (w[i] >> 12) & 0x3Fu, wp = (w[i] >> 18) & (uint32_t)(WC_CAP - 1);
        const bool fits = cnt < free_;
        // a query that does not fit its ring stores into the sink's ring instead (and is then counted some other way below)
        sts32(a_ring + ((fits ? bx[i] : wv.n_buckets) * (uint32_t)WC_STRIDE + ((wp + cnt) & (uint32_t)(WC_CAP - 1))) * 4u, elem[i]);
        slow |= (bk[i] != WC_NONE && !fits) ? (1u << i) : 0u;
      }
      if (slow) {                       // ring full (skewed input only)
#pragma unroll
        for (int i = 0; i < WC_ITEMS; i++)
          if ((slow >> i) & 1u) {
            diverted++;
            front.divert(c[i], s[i], e[i], (stw >> (i * 8)) & 0xFFu, q.index_base + tile * WC_TILE + (int64_t)threadIdx.x * WC_ITEMS + i);
          }
      }
    }
    __syncthreads();                    // B2: all elements of the round are in the rings

    // ---- owners: move complete lines out, publish the ring state for the next round
    if ((threadIdx.x & ~31u) < wv.n_buckets) {                        // warp-uniform: the warps that hold owners
      const bool owner = threadIdx.x < wv.n_buckets;
      const uint32_t cnt = owner ? lds32(a_myword) & 0xFFFu : 0u;
      occ += min(cnt, (uint32_t)WC_CAP - occ);                        // what the inserters were allowed to store
      const uint32_t nl = occ / WC_LINE;                              // 0, 1 or 2 complete lines
      const bool want_block = nl && (blk == WC_NONE || used + nl > (uint32_t)WC_BLOCK);
      const uint32_t fresh = warp_slots(want_block);
      my_blocks += want_block ? 1u : 0u;
      for (uint32_t l = 0; l < nl; l++) { flush_line(head, fresh, WC_LINE); head ^= (uint32_t)WC_LINE; }
      occ &= (uint32_t)(WC_LINE - 1);
      if (cnt) sts32(a_myword, (((uint32_t)WC_CAP - occ) << 12) | (((head + occ) & (uint32_t)(WC_CAP - 1)) << 18));
    }
  }
  __syncthreads();
  // ---- the rings' remainders leave as partial lines
  if ((threadIdx.x & ~31u) < wv.n_buckets) {
    const bool owner = threadIdx.x < wv.n_buckets;
    const bool rest = owner && occ != 0;
    const bool want_block = rest && (blk == WC_NONE || used == (uint32_t)WC_BLOCK);
    const uint32_t fresh = warp_slots(want_block);
    my_blocks += want_block ? 1u : 0u;
    if (rest) flush_line(head, fresh, occ);
    if (owner) { close_block(); my_blocks += s_direct[threadIdx.x]; }
    if (my_blocks) atomicAdd(wv.n_lines + threadIdx.x, my_blocks);
  }
#ifdef GTB_WC_FRONT_ONLY
  if (diverted == 0x7FFFFFF1u) atomicAdd(wv.diverted, 1ull);
#else
  if (diverted) atomicAdd(wv.diverted, (unsigned long long)diverted);
#endif
  __syncthreads();
  if (threadIdx.x == 0) wv.cta_lines[blockIdx.x] = s_next_line;
}

// exclusive scan of the per-bucket block counts (one CTA) and reset of the scatter cursors
__global__ void __launch_bounds__(1024) wc_line_offsets_kernel(WcView wv) {
  __shared__ uint32_t s_tot[32];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint32_t base = 0; base < wv.n_buckets; base += blockDim.x) {
    const uint32_t b = base + threadIdx.x;
    const uint32_t u = b < wv.n_buckets ? wv.n_lines[b] : 0;
    uint32_t inc = u;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
    if (lane == 31) s_tot[warp] = inc;
    __syncthreads();
    uint32_t pre = carry;
    for (int w = 0; w < warp; w++) pre += s_tot[w];
    if (b < wv.n_buckets) { wv.line_off[b] = pre + inc - u; wv.line_cursor[b] = 0; }
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = pre + inc;
    __syncthreads();
  }
  if (threadIdx.x == 0) wv.line_off[wv.n_buckets] = carry;
}

// groups the block slots by bucket: CTA c walks the slots CTA c of pass 1 filled (same grid)
__global__ void __launch_bounds__(512) wc_line_scatter_kernel(WcView wv) {
  __shared__ uint32_t s_cnt[WC_MAX_BUCKETS], s_base[WC_MAX_BUCKETS];
  for (uint32_t i = threadIdx.x; i < wv.n_buckets; i += blockDim.x) s_cnt[i] = 0;
  __syncthreads();
  const uint32_t n = wv.cta_lines[blockIdx.x];
  const size_t first = (size_t)blockIdx.x * wv.lines_per_cta;
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) atomicAdd(&s_cnt[wv.line_info[first + i] & 0xFFFFu], 1u);
  __syncthreads();
  for (uint32_t b = threadIdx.x; b < wv.n_buckets; b += blockDim.x) {
    const uint32_t c = s_cnt[b];
    s_base[b] = wv.line_off[b] + (c ? atomicAdd(wv.line_cursor + b, c) : 0u);
    s_cnt[b] = 0;
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
    const uint32_t info = wv.line_info[first + i];
    const uint32_t b = info & 0xFFFFu;
    const uint32_t pos = s_base[b] + atomicAdd(&s_cnt[b], 1u);
    wv.sorted_lines[pos] = (uint32_t)(first + i) | ((info >> 16) << 25);
  }
}

// ---- host side: buffers of one partition target and the three launches --------------------------------------------------------
struct WcBuffers {
  dbuf<uint32_t> pool, line_info, cta_lines, n_lines, line_off, line_cursor, sorted_lines;
  dbuf<unsigned long long> diverted;
  bool diverted_ready = false;
  void release() {
    pool.release(); line_info.release(); cta_lines.release(); n_lines.release(); line_off.release(); line_cursor.release();
    sorted_lines.release(); diverted.release(); diverted_ready = false;
  }
  // Sizes the buffers for n queries into nb buckets.  Returns GTB_ERR_UNSUPPORTED if the batch needs more block slots than the
  // 25-bit slot ids of sorted_lines can name.
  int plan(gtb_ctx *ctx, int64_t n, uint32_t nb, WcView *wv, unsigned *grid) {
    const int64_t tiles = (n + WC_TILE - 1) / WC_TILE;
    const unsigned gridw = (unsigned)std::max<int64_t>(1, std::min<int64_t>((int64_t)ctx->sm_count * WC_CTAS_PER_SM, tiles));
    const uint64_t tiles_cta = ((uint64_t)tiles + gridw - 1) / gridw;
    const uint64_t lines_per_cta = tiles_cta * (WC_TILE / WC_BLOCK_ELEMS) + nb + 1;      // full blocks + one open block per bucket
    const uint64_t total_slots = lines_per_cta * gridw;
    if (total_slots >= ((uint64_t)1 << 25)) return GTB_ERR_UNSUPPORTED;
    GTB_TRY(pool.reserve(ctx, (size_t)total_slots * WC_BLOCK_ELEMS));
    GTB_TRY(line_info.reserve(ctx, (size_t)total_slots));
    GTB_TRY(sorted_lines.reserve(ctx, (size_t)total_slots));
    GTB_TRY(cta_lines.reserve(ctx, (size_t)gridw));
    GTB_TRY(n_lines.reserve(ctx, (size_t)nb + 1));
    GTB_TRY(line_off.reserve(ctx, (size_t)nb + 1));
    GTB_TRY(line_cursor.reserve(ctx, (size_t)nb + 1));
    if (!diverted_ready) {
      GTB_TRY(diverted.reserve(ctx, 1));
      GTB_CUDA_OK(ctx, cudaMemsetAsync(diverted.p, 0, sizeof(unsigned long long), ctx->stream));
      diverted_ready = true;
    }
    wv->n_buckets = nb; wv->pool = pool.p; wv->lines_per_cta = (uint32_t)lines_per_cta; wv->line_info = line_info.p;
    wv->cta_lines = cta_lines.p; wv->n_lines = n_lines.p; wv->line_off = line_off.p; wv->line_cursor = line_cursor.p;
    wv->sorted_lines = sorted_lines.p; wv->diverted = diverted.p;
    *grid = gridw;
    return GTB_OK;
  }
};

// pass 1 + the grouping of its blocks by bucket; afterwards line_off / sorted_lines / pool describe every bucket's elements
template <class Front>
int wc_partition_launch(gtb_ctx *ctx, const char *name, const WcQueries &q, const Front &front, const WcView &wv, unsigned grid, size_t smem) {
  GTB_CUDA_OK(ctx, cudaMemsetAsync(wv.n_lines, 0, (size_t)wv.n_buckets * 4, ctx->stream));
  GTB_CUDA_OK(ctx, cudaFuncSetAttribute(wc_partition_kernel<Front>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  GTB_LAUNCH(ctx, name, wc_partition_kernel<Front>, grid, WC_THREADS, smem, q, front, wv);
  GTB_TRY(gtb_check_launch(ctx));
  GTB_LAUNCH(ctx, "bucket_line_offsets", wc_line_offsets_kernel, 1, 1024, 0, wv);
  GTB_LAUNCH(ctx, "bucket_line_scatter", wc_line_scatter_kernel, grid, 512, 0, wv);
  return gtb_check_launch(ctx);
}

}  // namespace
