// gtb_cell.cu -- CELL engine: the single-pass fast path for single-interval queries.
//
// Why this shape (DESIGN.md "Kernels"): with unsorted queries every per-query table access is a
// random access.  Shared-memory atomics cost ~2 cycles/lane and a divergent global load costs one
// L1 wavefront per lane, so at the ~1 query/SM-clock the HBM roofline asks for there is room for
// about one cheap shared-memory LOOKUP and one fire-and-forget global REDUCTION per query, nothing
// more.  The evaluation points (2 per region) do not fit in shared memory, but one BIT per genome
// cell does:
//   * every group's coordinate axis [0, size_g+1] is cut into cells of 2^k bp, laid end to end
//     (n_cells <= 1 572 864, so the bitmap fits in 192 KB of shared memory; hg19 x 2 strands: k = 12);
//   * a cell is HOT if it contains an evaluation point.  A query whose start and stop fall into the
//     same COLD cell lies inside one inter-point segment: a single red.global.add on the cell's
//     "both" counter (table is a few MB: L2 resident, no HBM traffic);
//   * otherwise start and stop are handled separately: a red on the cell's start / stop counter,
//     and if that cell is hot its <=13 points arrive as ONE 32-byte record and each point at or
//     beyond the coordinate gets a red on its correction counter;
//   * finalisation turns the cell tables into prefix sums, so for an evaluation point p
//       #{qs <= p} = sum(cells of the group before cell(p)) + correction[p]            (exactly).
// Divergence: ~8 % of queries touch a hot cell, so nearly every warp would have a lane on the slow
// branch.  The kernel therefore works tile by tile (4 096 queries per CTA): a branch-free front
// half retires cold queries on the spot; the rest are compacted into a shared-memory queue (warp
// scan + one shared atomic per warp) and processed with all lanes busy while the NEXT tile's front
// half runs (two queue buffers, one barrier per tile).
// Queries the scheme cannot place (start <= 0, a cell holding more than 13 points) take the
// general rank step; results add up because every table is a sum over queries.
#include "gtb_rank_device.cuh"
#include <algorithm>

namespace {

constexpr uint32_t CELL_MAX_CELLS = 1572864;     // 49 152 bitmap words = 192 KB of shared memory
constexpr int CELL_QCAP = 768;                   // deferred-query queue entries per buffer (2 buffers x 20 B)
constexpr int CELL_MIN_K = 3;
constexpr int CELL_MAX_K = 16;                   // point offsets inside a cell are stored in 16 bits
constexpr int CELL_THREADS = 512;
constexpr int CELL_SMEM_GROUPS = 512;            // group tables are staged in shared memory up to this many groups

struct CellView {
  int k;
  uint32_t cell_mask;
  int32_t n_chrom, n_class, n_groups;
  int cls_plus, cls_minus;                       // class_of['+'], class_of['-'] hoisted out of the table
  const int8_t *class_of;
  const uint8_t *chrom_present;
  const int32_t *gsize;
  const uint32_t *gbase;
  const int2 *gtab;                              // (gsize, gbase) interleaved, for genomes with many groups
  uint32_t n_cells, n_words;
  const uint32_t *bitmap, *wrank;
  const HotRec *hot;
  ull *cells;                                    // planes of n_cells
  ull *corr;                                     // planes of n_slots
};

// 256-bit streaming load (sm_100a LDG.E.256): read-only path, no L1 allocation, first to leave L2
struct int8v { int v[8]; };
__device__ __forceinline__ int8v ldg_stream256(const int *p) {
  int8v r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
               : "l"(p));
  return r;
}
__device__ __forceinline__ int2 ldg_stream64(const int *p) {
  int2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.b32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}

// fire-and-forget 64-bit reduction (no return value -> RED, not ATOM)
__device__ __forceinline__ void red_add(ull *p, ull v) {
  asm volatile("red.global.add.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

// One endpoint (start or stop) of a deferred query: bump the cell's start/stop counter and, if the
// cell is hot, the correction counter of every point at or beyond the coordinate.  The cell's
// points are sorted, so "at or beyond" is a suffix [f, n) of the record.
template <bool COVERAGE>
__device__ __forceinline__ void endpoint_update(const CellView &cv, int64_t K, int cnt_plane, int sum_plane, int xcnt_plane,
                                                int xsum_plane, uint32_t cell, uint32_t off, const uint4 &ra, const uint4 &rb,
                                                uint32_t n, ull w, ull wx) {
  red_add(cv.cells + (int64_t)cnt_plane * cv.n_cells + cell, w);
  if (COVERAGE) red_add(cv.cells + (int64_t)sum_plane * cv.n_cells + cell, wx);
  if (n == 0) return;
  const uint32_t words[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
  uint32_t f = 0;
#pragma unroll
  for (int i = 0; i < HOT_MAX; i++) {                     // fully unrolled: the record stays in registers
    const int h = i + 3;
    const uint32_t po = (words[h >> 1] >> ((h & 1) * 16)) & 0xFFFFu;
    f += ((uint32_t)i < n && po < off) ? 1u : 0u;
  }
  ull *xc = cv.corr + (int64_t)xcnt_plane * K + ra.x, *xs = cv.corr + (int64_t)xsum_plane * K + ra.x;
  for (uint32_t i = f; i < n; i++) {
    red_add(xc + i, w);
    if (COVERAGE) red_add(xs + i, wx);
  }
}

// A deferred query (everything the branch-free front half did not retire): the reference's
// admission checks, the general rank step for what cells cannot express, else the hot-cell path.
template <bool COVERAGE>
__device__ __noinline__ void deferred_query(const CellView &cv, const RankView &rv, const uint32_t *s_bitmap, const int2 *gtab,
                                            int32_t c, int32_t qs, int32_t qe, int strand, int64_t w, int64_t index) {
  // (the front half only defers queries on chromosome ids the index knows)
  if (qe <= 0 || qs > qe) {                                                // fatal only on indexed chromosomes, :5731-5741
    if (cv.chrom_present[c]) report_error(rv.err, index, qe <= 0 ? GTB_ERR_QUERY_STOP_NONPOSITIVE : GTB_ERR_QUERY_START_GT_STOP);
    return;
  }
  const int cls = cv.class_of[(uint8_t)strand];
  if (cls < 0) return;                                                     // no index region carries this strand, :5229
  const int g = c * cv.n_class + cls;
  const int2 gt = gtab[g];                                                 // (largest point, first cell)
  const int32_t gs = gt.x;
  if (gs <= 0 || qs < 1) {                                                 // no point > 0 in the group, or start before the cells
    const int gb = rv.goff[g], ge = rv.goff[g + 1];
    if (ge > gb && (qs < 1 || gs < 0)) rank_item<COVERAGE>(rv, gb, ge, qs, qe, w);
    return;
  }
  if (qs > gs) return;
  const int64_t K = rv.n_slots;
  const int32_t qe_c = qe < gs + 1 ? qe : gs + 1;
  const uint32_t cs = (uint32_t)gt.y + ((uint32_t)qs >> cv.k), ce = (uint32_t)gt.y + ((uint32_t)qe_c >> cv.k);
  const uint32_t word_s = s_bitmap[cs >> 5], word_e = s_bitmap[ce >> 5];
  const bool hot_s = (word_s >> (cs & 31)) & 1u, hot_e = (word_e >> (ce & 31)) & 1u;
  uint4 sa = make_uint4(0, 0, 0, 0), sb = sa, ea = sa, eb = sa;
  uint32_t hs = 0, he = 0;
  if (hot_s) hs = __ldg(cv.wrank + (cs >> 5)) + __popc(word_s & ((1u << (cs & 31)) - 1u));
  if (hot_e) he = __ldg(cv.wrank + (ce >> 5)) + __popc(word_e & ((1u << (ce & 31)) - 1u));
  if (hot_s) { sa = __ldg(reinterpret_cast<const uint4 *>(cv.hot + hs)); sb = __ldg(reinterpret_cast<const uint4 *>(cv.hot + hs) + 1); }
  if (hot_e) { ea = __ldg(reinterpret_cast<const uint4 *>(cv.hot + he)); eb = __ldg(reinterpret_cast<const uint4 *>(cv.hot + he) + 1); }
  const uint32_t ns = hot_s ? (sa.y & 0xFFFFu) : 0u, ne = hot_e ? (ea.y & 0xFFFFu) : 0u;
  if (ns > (uint32_t)HOT_MAX || ne > (uint32_t)HOT_MAX) {                  // overfull cell: general step
    rank_item<COVERAGE>(rv, rv.goff[g], rv.goff[g + 1], qs, qe, w);
    return;
  }
  endpoint_update<COVERAGE>(cv, K, C_SCNT, C_SSUM, X_SCNT, X_SSUM, cs, (uint32_t)qs & cv.cell_mask, sa, sb, ns, (ull)w, (ull)(w * (int64_t)qs));
  endpoint_update<COVERAGE>(cv, K, C_ECNT, C_ESUM, X_ECNT, X_ESUM, ce, (uint32_t)qe_c & cv.cell_mask, ea, eb, ne, (ull)w, (ull)(w * (int64_t)qe));
}

// VEC: 8 = 256-bit loads of chrom/start/stop (+ 64 bits of strand) per thread, 1 = scalar (unaligned batches)
template <bool COVERAGE, bool WEIGHTED, int VEC>
__global__ void __launch_bounds__(CELL_THREADS, 1) cell_accumulate_kernel(QueryView q, RankView rv, CellView cv) {
  extern __shared__ __align__(16) uint32_t smem[];
  int4 *s_queue = reinterpret_cast<int4 *>(smem);                                    // [2][CELL_QCAP] (qs, qe, chrom | strand << 24, weight)
  uint32_t *s_qidx = smem + 8 * CELL_QCAP;                                           // [2][CELL_QCAP] index inside the batch
  int2 *s_gtab = reinterpret_cast<int2 *>(smem + 10 * CELL_QCAP);                    // [CELL_SMEM_GROUPS] (largest point, first cell)
  uint32_t *s_bitmap = smem + 10 * CELL_QCAP + 2 * CELL_SMEM_GROUPS;                 // [n_words]
  __shared__ unsigned s_qcount[3];
  for (uint32_t i = threadIdx.x; i < cv.n_words; i += blockDim.x) s_bitmap[i] = cv.bitmap[i];
  const bool groups_in_smem = cv.n_groups <= CELL_SMEM_GROUPS;
  if (groups_in_smem)
    for (int i = threadIdx.x; i < cv.n_groups; i += blockDim.x) s_gtab[i] = cv.gtab[i];
  if (threadIdx.x < 3) s_qcount[threadIdx.x] = 0;
  __syncthreads();
  const int2 *gtab = groups_in_smem ? s_gtab : cv.gtab;
  const int lane = threadIdx.x & 31;
  constexpr int PER_THREAD = 8;
  const int64_t tile_items = (int64_t)CELL_THREADS * PER_THREAD;
  const int64_t n_tiles = (q.n_regions + tile_items - 1) / tile_items;
  const uint32_t last_cell = cv.n_cells - 1;

  // tile data lives in registers; the next tile's loads are issued before this tile is processed
  int32_t nc[PER_THREAD], ns[PER_THREAD], ne[PER_THREAD], nw[PER_THREAD];
  int2 nst = make_int2(0, 0);
  auto fetch = [&](int64_t tile) {
    const int64_t first = tile * tile_items + (int64_t)threadIdx.x * PER_THREAD;
    if (VEC == 8 && first + PER_THREAD <= q.n_regions) {
      const int8v cc = ldg_stream256(q.chrom + first), ss = ldg_stream256(q.start + first), ee = ldg_stream256(q.stop + first);
      nst = ldg_stream64(reinterpret_cast<const int *>(q.strand + first));
#pragma unroll
      for (int i = 0; i < PER_THREAD; i++) { nc[i] = cc.v[i]; ns[i] = ss.v[i]; ne[i] = ee.v[i]; }
      if (WEIGHTED) {
        const int8v ww = ldg_stream256(q.weight + first);
#pragma unroll
        for (int i = 0; i < PER_THREAD; i++) nw[i] = ww.v[i];
      }
    } else {
      unsigned sx = 0, sy = 0;
#pragma unroll
      for (int i = 0; i < PER_THREAD; i++) {
        const int64_t r = first + i;
        const bool ok = r < q.n_regions;
        nc[i] = ok ? q.chrom[r] : -1; ns[i] = ok ? q.start[r] : 1; ne[i] = ok ? q.stop[r] : 1;
        const unsigned sb = ok ? (unsigned)(uint8_t)q.strand[r] : (unsigned)'+';
        if (i < 4) sx |= sb << (i * 8); else sy |= sb << ((i & 3) * 8);
        if (WEIGHTED) nw[i] = ok ? q.weight[r] : 1;
      }
      nst = make_int2((int)sx, (int)sy);
    }
  };
  if ((int64_t)blockIdx.x < n_tiles) fetch(blockIdx.x);

  // drains queue buffer `buf` whose fill count is in counter `cnt`
  auto drain = [&](int buf, int cnt) {
    const unsigned n_q = min(s_qcount[cnt], (unsigned)CELL_QCAP);
    for (unsigned j = threadIdx.x; j < n_q; j += blockDim.x) {
      const int4 d = s_queue[buf * CELL_QCAP + j];
      deferred_query<COVERAGE>(cv, rv, s_bitmap, gtab, d.z & 0xFFFFFF, d.x, d.y, (int)(int8_t)((unsigned)d.z >> 24), (int64_t)d.w,
                               q.index_base + s_qidx[buf * CELL_QCAP + j]);
    }
  };

  int it = 0;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++) {
    const int64_t first = tile * tile_items + (int64_t)threadIdx.x * PER_THREAD;
    int32_t c[PER_THREAD], s[PER_THREAD], e[PER_THREAD], wt[PER_THREAD];
    const int2 stw = nst;
#pragma unroll
    for (int i = 0; i < PER_THREAD; i++) { c[i] = nc[i]; s[i] = ns[i]; e[i] = ne[i]; wt[i] = WEIGHTED ? nw[i] : 1; }
    if (tile + gridDim.x < n_tiles) fetch(tile + gridDim.x);

    // ---- front half: straight-line, predicated.  Retires queries that sit in one cold cell (one
    // reduction) and those that provably contribute nothing; everything else is deferred.
    unsigned pending = 0;
#pragma unroll
    for (int i = 0; i < PER_THREAD; i++) {
      const int strand = ((i < 4 ? stw.x : stw.y) >> ((i & 3) * 8)) & 0xFF;
      const bool known = (uint32_t)c[i] < (uint32_t)cv.n_chrom && (uint32_t)c[i] < 0x1000000u;   // else: no match, no checks (:5719)
      const int cls = strand == '+' ? cv.cls_plus : (strand == '-' ? cv.cls_minus : -1);
      const bool plain = known && cls >= 0 && s[i] >= 1 && s[i] <= e[i];
      const int2 gt = gtab[plain ? c[i] * cv.n_class + cls : 0];
      const bool in_cells = plain && gt.x > 0;
      const bool beyond = in_cells && s[i] > gt.x;                        // past every evaluation point: contributes nothing
      const int32_t qe_c = min(e[i], gt.x + 1);
      const uint32_t cs = min((uint32_t)gt.y + ((uint32_t)s[i] >> cv.k), last_cell);
      const uint32_t ce = (uint32_t)gt.y + ((uint32_t)qe_c >> cv.k);
      const bool hot_s = (s_bitmap[cs >> 5] >> (cs & 31)) & 1u;
      const bool cold = in_cells && !beyond && cs == ce && !hot_s;
      if (cold) red_add(cv.cells + cs, COVERAGE ? (ull)((int64_t)wt[i] * ((int64_t)e[i] - s[i] + 1)) : (ull)(int64_t)wt[i]);   // plane C_BOTH
      // defer: hot or straddling cells, odd strand bytes, start <= 0, invalid intervals, point-free groups
      const bool skip = !known || cold || beyond || (plain && gt.x == 0) ||                  // empty group: nothing to count
                        (known && (strand == '+' || strand == '-') && cls < 0 && s[i] <= e[i] && e[i] > 0);
      if (!skip) pending |= 1u << i;
    }
    // ---- compact this tile's deferred queries into queue buffer it&1
    const int mine = __popc(pending);
    int incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += v; }
    const int warp_total = __shfl_sync(0xffffffffu, incl, 31);
    unsigned warp_base = 0;
    if (warp_total > 0) {
      if (lane == 31) warp_base = atomicAdd(&s_qcount[it % 3], (unsigned)warp_total);
      warp_base = __shfl_sync(0xffffffffu, warp_base, 31);
    }
    unsigned slot = warp_base + (unsigned)(incl - mine);
    const int buf = it & 1;
#pragma unroll
    for (int i = 0; i < PER_THREAD; i++) {
      if (pending & (1u << i)) {
        const int strand = ((i < 4 ? stw.x : stw.y) >> ((i & 3) * 8)) & 0xFF;
        if (slot < (unsigned)CELL_QCAP) {
          s_queue[buf * CELL_QCAP + slot] = make_int4(s[i], e[i], c[i] | (strand << 24), wt[i]);
          s_qidx[buf * CELL_QCAP + slot] = (uint32_t)(first + i);
        } else {                                                          // queue full: do it now
          deferred_query<COVERAGE>(cv, rv, s_bitmap, gtab, c[i], s[i], e[i], (int)(int8_t)strand, (int64_t)wt[i], q.index_base + first + i);
        }
        slot++;
      }
    }
    // ---- back half of the PREVIOUS tile (its queue was completed by the barrier below), all lanes busy
    if (it > 0) drain(buf ^ 1, (it + 2) % 3);
    if (threadIdx.x == 0) s_qcount[(it + 1) % 3] = 0;      // drained during the previous iteration; next tile appends here
    __syncthreads();
  }
  if (it > 0) drain((it - 1) & 1, (it + 2) % 3);
}

template <typename T>
int upload_v(gtb_ctx *ctx, dbuf<T> &d, const std::vector<T> &h) {
  GTB_TRY(d.reserve(ctx, h.size() ? h.size() : 1));
  if (h.size()) GTB_CUDA_OK(ctx, cudaMemcpyAsync(d.p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  return GTB_OK;
}

}  // namespace

// -------------------------------------------------------------------------------------------------
// host: build the cell structures from the rank structures (points per group)
// -------------------------------------------------------------------------------------------------
int gtb_cell_prepare(gtb_index *ix) {
  if (ix->cell && ix->cell->ready) return GTB_OK;
  gtb_ctx *ctx = ix->ctx;
  if (!ix->cell) ix->cell = new gtb_cell_state();
  gtb_cell_state *cs = ix->cell;
  const int G = ix->n_groups;
  // group sizes: largest real point (sentinel excluded); -1 marks "only points <= 0"
  std::vector<int32_t> gsize((size_t)std::max(G, 1), 0);
  for (int g = 0; g < G; g++) {
    const int32_t gb = ix->h_goff[g], ge = ix->h_goff[g + 1];
    if (ge - gb >= 2) { const int32_t mx = ix->h_points[ge - 2]; gsize[g] = mx >= 1 ? mx : -1; }
  }
  // smallest k whose cell count fits the shared-memory budget
  int k = CELL_MIN_K;
  if (const char *env = getenv("GTB_CELL_K")) k = std::max(0, std::min(CELL_MAX_K, atoi(env)));   // tests: force a width
  uint64_t total = 0;
  for (; k <= CELL_MAX_K; k++) {
    total = 0;
    for (int g = 0; g < G; g++) if (gsize[g] > 0) total += (((uint64_t)gsize[g] + 1) >> k) + 1;
    if (total <= CELL_MAX_CELLS) break;
  }
  if (k > CELL_MAX_K) { cs->ready = false; return GTB_ERR_UNSUPPORTED; }
  cs->k = k;
  std::vector<uint32_t> gbase((size_t)std::max(G, 1), 0);
  uint32_t next = 0;
  for (int g = 0; g < G; g++) { gbase[g] = next; if (gsize[g] > 0) next += (uint32_t)((((uint64_t)gsize[g] + 1) >> k) + 1); }
  cs->n_cells = std::max<uint32_t>(next, 1);
  cs->n_words = (cs->n_cells + 31) / 32;
  // hot cells and their records
  std::vector<uint32_t> bitmap(cs->n_words, 0), wrank(cs->n_words, 0), slot_cell((size_t)std::max<int64_t>(ix->n_slots, 1), 0xFFFFFFFFu);
  for (int g = 0; g < G; g++) {
    const int32_t gb = ix->h_goff[g], ge = ix->h_goff[g + 1];
    for (int32_t j = gb; j < ge - 1; j++) {                 // ge-1 = sentinel
      const int32_t p = ix->h_points[j];
      if (p < 1) continue;                                  // no fast-path query is <= such a point
      const uint32_t c = gbase[g] + ((uint32_t)p >> k);
      slot_cell[j] = c;
      bitmap[c >> 5] |= 1u << (c & 31);
    }
  }
  uint32_t n_hot = 0;
  for (uint32_t w = 0; w < cs->n_words; w++) { wrank[w] = n_hot; n_hot += (uint32_t)__builtin_popcount(bitmap[w]); }
  cs->n_hot = n_hot;
  std::vector<HotRec> hot((size_t)std::max<uint32_t>(n_hot, 1));
  memset(hot.data(), 0, hot.size() * sizeof(HotRec));
  {
    std::vector<uint32_t> count((size_t)std::max<uint32_t>(n_hot, 1), 0);
    for (int64_t j = 0; j < ix->n_slots; j++) {
      const uint32_t c = slot_cell[j];
      if (c == 0xFFFFFFFFu) continue;
      const uint32_t h = wrank[c >> 5] + (uint32_t)__builtin_popcount(bitmap[c >> 5] & ((1u << (c & 31)) - 1u));
      HotRec &r = hot[h];
      if (count[h] == 0) r.slot_base = (uint32_t)j;         // slots of one cell are consecutive (sorted points, one group)
      if (count[h] < (uint32_t)HOT_MAX) r.off[count[h]] = (uint16_t)((uint32_t)ix->h_points[j] & ((1u << k) - 1u));
      count[h]++;
      r.n = (uint16_t)std::min<uint32_t>(count[h], 0xFFFFu);
    }
  }
  cs->cell_planes = ix->op == GTB_OP_COVERAGE ? 5 : 3;
  cs->corr_planes = ix->op == GTB_OP_COVERAGE ? 4 : 2;
  cs->smem_bytes = (size_t)CELL_QCAP * 40 + (size_t)CELL_SMEM_GROUPS * 8 + (size_t)cs->n_words * 4;
  GTB_TRY(upload_v(ctx, cs->d_gsize, gsize));
  GTB_TRY(upload_v(ctx, cs->d_gbase, gbase));
  {
    std::vector<int2> gtab((size_t)std::max(G, 1));
    for (int g = 0; g < G; g++) gtab[g] = make_int2(gsize[g], (int)gbase[g]);
    GTB_TRY(upload_v(ctx, cs->d_gtab, gtab));
  }
  GTB_TRY(upload_v(ctx, cs->d_bitmap, bitmap));
  GTB_TRY(upload_v(ctx, cs->d_wrank, wrank));
  GTB_TRY(upload_v(ctx, cs->d_hot, hot));
  GTB_TRY(upload_v(ctx, cs->d_slot_cell, slot_cell));
  GTB_TRY(cs->d_cells.reserve(ctx, (size_t)cs->cell_planes * cs->n_cells));
  GTB_TRY(cs->d_cells_scan.reserve(ctx, (size_t)cs->cell_planes * cs->n_cells));
  GTB_TRY(cs->d_corr.reserve(ctx, (size_t)cs->corr_planes * (size_t)std::max<int64_t>(ix->n_slots, 1)));
  GTB_CUDA_OK(ctx, cudaMemsetAsync(cs->d_cells.p, 0, sizeof(ull) * (size_t)cs->cell_planes * cs->n_cells, ctx->stream));
  GTB_CUDA_OK(ctx, cudaMemsetAsync(cs->d_corr.p, 0, sizeof(ull) * (size_t)cs->corr_planes * (size_t)std::max<int64_t>(ix->n_slots, 1), ctx->stream));
  GTB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  cs->ready = true;
  cs->dirty = false;
  return GTB_OK;
}

bool gtb_cell_supported(gtb_index *ix, const QueryView &q, bool batch_multi) {
  if (batch_multi || q.region_offset) return false;                  // single-interval batches only
  if (ix->n_slots == 0 || ix->n_chrom > (1 << 24) || q.n_regions >= ((int64_t)1 << 32)) return false;
  if (ix->cell && !ix->cell->ready) return false;                    // a previous prepare said "unsupported"
  if (!ix->cell) {
    int rc = gtb_cell_prepare(ix);
    if (rc != GTB_OK) { if (!ix->cell) ix->cell = new gtb_cell_state(); ix->cell->ready = false; return false; }
  }
  return ix->cell->smem_bytes <= ix->ctx->smem_optin;
}

int gtb_cell_reset(gtb_index *ix) {
  gtb_cell_state *cs = ix->cell;
  if (!cs || !cs->ready || !cs->dirty) return GTB_OK;
  gtb_ctx *ctx = ix->ctx;
  GTB_CUDA_OK(ctx, cudaMemsetAsync(cs->d_cells.p, 0, sizeof(ull) * (size_t)cs->cell_planes * cs->n_cells, ctx->stream));
  GTB_CUDA_OK(ctx, cudaMemsetAsync(cs->d_corr.p, 0, sizeof(ull) * (size_t)cs->corr_planes * (size_t)std::max<int64_t>(ix->n_slots, 1), ctx->stream));
  cs->dirty = false;
  return GTB_OK;
}

int gtb_cell_accumulate(gtb_index *ix, const QueryView &q) {
  gtb_ctx *ctx = ix->ctx;
  GTB_TRY(gtb_cell_prepare(ix));
  gtb_cell_state *cs = ix->cell;
  if (!cs->ready) return gtb_fail(ctx, GTB_ERR_UNSUPPORTED, "cell engine cannot serve this index (genome too large for 16-bit cell offsets)");
  if (q.region_offset) return gtb_fail(ctx, GTB_ERR_UNSUPPORTED, "cell engine takes single-interval batches");
  CellView cv;
  cv.k = cs->k; cv.cell_mask = (1u << cs->k) - 1u;
  cv.n_chrom = ix->n_chrom; cv.n_class = ix->n_class; cv.n_groups = ix->n_groups;
  cv.cls_plus = ix->h_class_of[(uint8_t)'+']; cv.cls_minus = ix->h_class_of[(uint8_t)'-'];
  cv.class_of = ix->d_class_of.p; cv.chrom_present = ix->d_present.p;
  cv.gsize = cs->d_gsize.p; cv.gbase = cs->d_gbase.p; cv.gtab = cs->d_gtab.p;
  cv.n_cells = cs->n_cells; cv.n_words = cs->n_words;
  cv.bitmap = cs->d_bitmap.p; cv.wrank = cs->d_wrank.p; cv.hot = cs->d_hot.p;
  cv.cells = cs->d_cells.p; cv.corr = cs->d_corr.p;
  RankView rv;
  rv.n_chrom = ix->n_chrom; rv.n_class = ix->n_class; rv.class_of = ix->d_class_of.p; rv.chrom_present = ix->d_present.p;
  rv.goff = ix->d_goff.p; rv.points = ix->d_points.p; rv.n_slots = ix->n_slots; rv.hist = ix->d_hist.p; rv.err = ix->d_err.p;

  const bool aligned = ((uintptr_t)q.chrom % 32 == 0) && ((uintptr_t)q.start % 32 == 0) && ((uintptr_t)q.stop % 32 == 0) &&
                       ((uintptr_t)q.strand % 8 == 0) && (!q.weight || (uintptr_t)q.weight % 32 == 0);
  const bool cov = ix->op == GTB_OP_COVERAGE, wt = q.weight != nullptr;
  const size_t smem = cs->smem_bytes;
  const int64_t per_block = (int64_t)CELL_THREADS * 8;
  const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ctx->sm_count, (q.n_regions + per_block - 1) / per_block));

#define GTB_CELL_LAUNCH(COV, WT, VEC)                                                                                   \
  do {                                                                                                                  \
    auto kern = cell_accumulate_kernel<COV, WT, VEC>;                                                                   \
    GTB_CUDA_OK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));               \
    GTB_LAUNCH(ctx, COV ? "cell_coverage" : "cell_count", kern, grid, CELL_THREADS, smem, q, rv, cv);                   \
  } while (0)
  if (aligned) {
    if (cov) { if (wt) GTB_CELL_LAUNCH(true, true, 8); else GTB_CELL_LAUNCH(true, false, 8); }
    else { if (wt) GTB_CELL_LAUNCH(false, true, 8); else GTB_CELL_LAUNCH(false, false, 8); }
  } else {
    if (cov) { if (wt) GTB_CELL_LAUNCH(true, true, 1); else GTB_CELL_LAUNCH(true, false, 1); }
    else { if (wt) GTB_CELL_LAUNCH(false, true, 1); else GTB_CELL_LAUNCH(false, false, 1); }
  }
#undef GTB_CELL_LAUNCH
  cs->dirty = true;
  return gtb_check_launch(ctx);
}

int gtb_cell_scan_for_finish(gtb_index *ix, CellFinalView *out) {
  memset(out, 0, sizeof(*out));
  gtb_cell_state *cs = ix->cell;
  if (!cs || !cs->ready || !cs->dirty) return GTB_OK;
  gtb_ctx *ctx = ix->ctx;
  GTB_TRY(gtb_inclusive_scan_planes_u64(ctx, cs->d_cells.p, cs->d_cells_scan.p, cs->n_cells, cs->cell_planes, cs->n_cells, ix->d_scan_scratch));
  out->cells_scan = cs->d_cells_scan.p;
  out->corr = cs->d_corr.p;
  out->slot_cell = cs->d_slot_cell.p;
  out->gbase = cs->d_gbase.p;
  out->n_cells = cs->n_cells;
  return GTB_OK;
}

void gtb_cell_destroy(gtb_index *ix) {
  gtb_cell_state *cs = ix->cell;
  if (!cs) return;
  cs->d_gsize.release(); cs->d_gbase.release(); cs->d_gtab.release(); cs->d_bitmap.release(); cs->d_wrank.release(); cs->d_hot.release();
  cs->d_slot_cell.release(); cs->d_cells.release(); cs->d_cells_scan.release(); cs->d_corr.release();
  delete cs;
  ix->cell = nullptr;
}
