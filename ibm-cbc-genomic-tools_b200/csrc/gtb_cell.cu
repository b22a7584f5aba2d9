// gtb_cell.cu -- CELL engine: the single-pass fast path for single-interval queries.
//
// Why this shape (DESIGN.md "Kernels"): with unsorted queries every per-query table access is a
// random access.  Shared-memory atomics cost ~2 cycles/lane and a divergent global load costs one
// L1 wavefront per lane, so at the ~1 query/SM-clock the HBM roofline asks for there is room for
// about one cheap shared-memory LOOKUP and one fire-and-forget global REDUCTION per query, nothing
// more.  The evaluation points (2 per region) do not fit in shared memory, but one BIT per genome
// cell does:
//   * every group's coordinate axis [0, size_g+1] is cut into cells of 2^k bp, laid end to end
//     (n_cells <= 851 968, so bitmap + per-word hot-rank fit in 208 KB of shared memory);
//   * a cell is HOT if it contains an evaluation point.  A query whose start and stop fall into the
//     same COLD cell lies inside one inter-point segment: a single red.global.add on the cell's
//     "both" counter (table is a few MB: L2 resident, no HBM traffic);
//   * otherwise start and stop are handled separately: a red on the cell's start / stop counter,
//     and if that cell is hot its <=13 points arrive as ONE 32-byte record and each point at or
//     beyond the coordinate gets a red on its correction counter;
//   * finalisation turns the cell tables into prefix sums, so for an evaluation point p
//       #{qs <= p} = sum(cells of the group before cell(p)) + correction[p]            (exactly).
// Queries the scheme cannot place (start <= 0, a cell holding more than 13 points) take the
// general rank step inline; results add up because every table is a sum over queries.
#include "gtb_rank_device.cuh"
#include <algorithm>

namespace {

constexpr uint32_t CELL_MAX_CELLS = 851968;      // 26 624 words * 8 B = 208 KB of shared memory
constexpr int CELL_MIN_K = 3;
constexpr int CELL_MAX_K = 16;                   // point offsets inside a cell are stored in 16 bits
constexpr int CELL_THREADS = 1024;
constexpr int CELL_SMEM_GROUPS = 1024;           // group tables are staged in shared memory up to this many groups

struct CellView {
  int k;
  uint32_t cell_mask;
  int32_t n_chrom, n_class, n_groups;
  int cls_plus, cls_minus;                       // class_of['+'], class_of['-'] hoisted out of the table
  const int8_t *class_of;
  const uint8_t *chrom_present;
  const int32_t *gsize;
  const uint32_t *gbase;
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

__device__ __forceinline__ void load_hot(const HotRec *__restrict__ hot, uint32_t h, uint32_t (&w)[8]) {
  const uint4 a = __ldg(reinterpret_cast<const uint4 *>(hot + h));
  const uint4 b = __ldg(reinterpret_cast<const uint4 *>(hot + h) + 1);
  w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
}
__device__ __forceinline__ uint32_t hot_n(const uint32_t (&w)[8]) { return w[1] & 0xFFFFu; }
__device__ __forceinline__ uint32_t hot_off(const uint32_t (&w)[8], int i) {      // off[i], i in [0,13)
  const int h = i + 3;                                                            // halfword index inside the record
  return (w[h >> 1] >> ((h & 1) * 16)) & 0xFFFFu;
}

template <bool COVERAGE>
__device__ __forceinline__ void process_query(const CellView &cv, const RankView &rv, const uint32_t *__restrict__ s_bitmap,
                                              const uint32_t *__restrict__ s_wrank, const int32_t *__restrict__ gsize,
                                              const uint32_t *__restrict__ gbase, int32_t c, int32_t qs, int32_t qe, int strand,
                                              int64_t w, int64_t index, int64_t K) {
  if ((uint32_t)c >= (uint32_t)cv.n_chrom) return;                         // chromosome unknown to the index, :5719-5720
  if (qe <= 0 || qs > qe) {                                                // fatal only on indexed chromosomes, :5731-5741
    if (cv.chrom_present[c]) report_error(rv.err, index, qe <= 0 ? GTB_ERR_QUERY_STOP_NONPOSITIVE : GTB_ERR_QUERY_START_GT_STOP);
    return;
  }
  const int cls = strand == '+' ? cv.cls_plus : (strand == '-' ? cv.cls_minus : (int)cv.class_of[(uint8_t)strand]);
  if (cls < 0) return;                                                     // no index region carries this strand, :5229
  const int g = c * cv.n_class + cls;
  const int32_t gs = gsize[g];
  if (gs <= 0) {                                                           // no point > 0 in the group
    if (qs < 1 || gs < 0) {                                                // ... but there may be points <= 0: general step
      const int gb = rv.goff[g], ge = rv.goff[g + 1];
      if (ge > gb) rank_item<COVERAGE>(rv, gb, ge, qs, qe, w);
    }
    return;
  }
  if (qs > gs) return;                                                     // beyond every evaluation point of the group
  if (qs < 1) {                                                            // cells start at coordinate 0
    rank_item<COVERAGE>(rv, rv.goff[g], rv.goff[g + 1], qs, qe, w);
    return;
  }
  const int32_t qe_c = qe < gs + 1 ? qe : gs + 1;                          // everything past the last point is one segment
  const uint32_t base = gbase[g];
  const uint32_t cs = base + ((uint32_t)qs >> cv.k), ce = base + ((uint32_t)qe_c >> cv.k);
  const uint32_t word_s = s_bitmap[cs >> 5];
  const bool hot_s = (word_s >> (cs & 31)) & 1u;
  ull *cells = cv.cells;
  const int64_t NC = cv.n_cells;
  if (cs == ce && !hot_s) {                                                // the common case: one cold cell
    red_add(cells + C_BOTH * NC + cs, COVERAGE ? (ull)(w * ((int64_t)qe - qs + 1)) : (ull)w);
    return;
  }
  const uint32_t word_e = s_bitmap[ce >> 5];
  const bool hot_e = (word_e >> (ce & 31)) & 1u;
  uint32_t rec_s[8], rec_e[8];
  if (hot_s) load_hot(cv.hot, s_wrank[cs >> 5] + __popc(word_s & ((1u << (cs & 31)) - 1u)), rec_s);
  if (hot_e) load_hot(cv.hot, s_wrank[ce >> 5] + __popc(word_e & ((1u << (ce & 31)) - 1u)), rec_e);
  if ((hot_s && hot_n(rec_s) > (uint32_t)HOT_MAX) || (hot_e && hot_n(rec_e) > (uint32_t)HOT_MAX)) {
    rank_item<COVERAGE>(rv, rv.goff[g], rv.goff[g + 1], qs, qe, w);        // overfull cell: general step
    return;
  }
  // start coordinate: counted for every point >= qs
  red_add(cells + C_SCNT * NC + cs, (ull)w);
  if (COVERAGE) red_add(cells + C_SSUM * NC + cs, (ull)(w * (int64_t)qs));
  if (hot_s) {
    const uint32_t o = (uint32_t)qs & cv.cell_mask, n = hot_n(rec_s), sb = rec_s[0];
    for (uint32_t i = 0; i < n; i++)
      if (o <= hot_off(rec_s, i)) {
        red_add(cv.corr + X_SCNT * K + sb + i, (ull)w);
        if (COVERAGE) red_add(cv.corr + X_SSUM * K + sb + i, (ull)(w * (int64_t)qs));
      }
  }
  // stop coordinate: counted for every point >= qe
  red_add(cells + C_ECNT * NC + ce, (ull)w);
  if (COVERAGE) red_add(cells + C_ESUM * NC + ce, (ull)(w * (int64_t)qe));
  if (hot_e) {
    const uint32_t o = (uint32_t)qe_c & cv.cell_mask, n = hot_n(rec_e), sb = rec_e[0];
    for (uint32_t i = 0; i < n; i++)
      if (o <= hot_off(rec_e, i)) {
        red_add(cv.corr + X_ECNT * K + sb + i, (ull)w);
        if (COVERAGE) red_add(cv.corr + X_ESUM * K + sb + i, (ull)(w * (int64_t)qe));
      }
  }
}

// VEC: 8 = 256-bit loads of chrom/start/stop (+ 64 bits of strand) per thread, 1 = scalar (unaligned batches)
template <bool COVERAGE, bool WEIGHTED, int VEC>
__global__ void __launch_bounds__(CELL_THREADS, 1) cell_accumulate_kernel(QueryView q, RankView rv, CellView cv) {
  extern __shared__ __align__(16) uint32_t smem[];
  uint32_t *s_bitmap = smem;
  uint32_t *s_wrank = smem + cv.n_words;
  int32_t *s_gsize = reinterpret_cast<int32_t *>(smem + 2 * cv.n_words);
  uint32_t *s_gbase = reinterpret_cast<uint32_t *>(s_gsize + CELL_SMEM_GROUPS);
  for (uint32_t i = threadIdx.x; i < cv.n_words; i += blockDim.x) { s_bitmap[i] = cv.bitmap[i]; s_wrank[i] = cv.wrank[i]; }
  const bool groups_in_smem = cv.n_groups <= CELL_SMEM_GROUPS;
  if (groups_in_smem)
    for (int i = threadIdx.x; i < cv.n_groups; i += blockDim.x) { s_gsize[i] = cv.gsize[i]; s_gbase[i] = cv.gbase[i]; }
  __syncthreads();
  const int32_t *gsize = groups_in_smem ? s_gsize : cv.gsize;
  const uint32_t *gbase = groups_in_smem ? s_gbase : cv.gbase;
  const int64_t K = rv.n_slots;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;

  if (VEC == 8) {
    const int64_t n8 = q.n_regions >> 3;
    for (int64_t v = tid; v < n8; v += stride) {
      const int8v c = ldg_stream256(q.chrom + (v << 3)), s = ldg_stream256(q.start + (v << 3)), e = ldg_stream256(q.stop + (v << 3));
      const int2 st = ldg_stream64(reinterpret_cast<const int *>(q.strand + (v << 3)));
      int8v w;
      if (WEIGHTED) w = ldg_stream256(q.weight + (v << 3));
      const int64_t idx = q.index_base + (v << 3);
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const int sw = i < 4 ? st.x : st.y;
        process_query<COVERAGE>(cv, rv, s_bitmap, s_wrank, gsize, gbase, c.v[i], s.v[i], e.v[i],
                                (int)(int8_t)((sw >> ((i & 3) * 8)) & 0xFF), WEIGHTED ? (int64_t)w.v[i] : 1, idx + i, K);
      }
    }
    for (int64_t r = (n8 << 3) + tid; r < q.n_regions; r += stride)
      process_query<COVERAGE>(cv, rv, s_bitmap, s_wrank, gsize, gbase, q.chrom[r], q.start[r], q.stop[r], (int)q.strand[r],
                              WEIGHTED ? (int64_t)q.weight[r] : 1, q.index_base + r, K);
  } else {
    for (int64_t r = tid; r < q.n_regions; r += stride)
      process_query<COVERAGE>(cv, rv, s_bitmap, s_wrank, gsize, gbase, q.chrom[r], q.start[r], q.stop[r], (int)q.strand[r],
                              WEIGHTED ? (int64_t)q.weight[r] : 1, q.index_base + r, K);
  }
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
  cs->smem_bytes = (size_t)cs->n_words * 8 + (size_t)CELL_SMEM_GROUPS * 8;
  GTB_TRY(upload_v(ctx, cs->d_gsize, gsize));
  GTB_TRY(upload_v(ctx, cs->d_gbase, gbase));
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
  if (ix->n_slots == 0) return false;
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
  cv.gsize = cs->d_gsize.p; cv.gbase = cs->d_gbase.p;
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
  const int64_t per_block = (int64_t)CELL_THREADS * (aligned ? 8 : 1);
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
  const size_t n = (size_t)cs->cell_planes * cs->n_cells;
  GTB_CUDA_OK(ctx, cudaMemcpyAsync(cs->d_cells_scan.p, cs->d_cells.p, n * sizeof(ull), cudaMemcpyDeviceToDevice, ctx->stream));
  for (int p = 0; p < cs->cell_planes; p++)
    GTB_TRY(gtb_inclusive_scan_u64(ctx, cs->d_cells_scan.p + (size_t)p * cs->n_cells, cs->n_cells, ix->d_scan_scratch));
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
  cs->d_gsize.release(); cs->d_gbase.release(); cs->d_bitmap.release(); cs->d_wrank.release(); cs->d_hot.release();
  cs->d_slot_cell.release(); cs->d_cells.release(); cs->d_cells_scan.release(); cs->d_corr.release();
  delete cs;
  ix->cell = nullptr;
}
