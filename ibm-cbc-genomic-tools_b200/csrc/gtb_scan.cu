// gtb_scan.cu -- genome-wide sliding-window read counts (genomic_scans counts).
//
// Replaces UnsortedGenomicRegionSetScanner (genomic_intervals.cpp:5019-5141): a dense micro-window
// histogram per (chromosome, strand) "slot", a sliding sum of win_size/win_step micro-windows, and
// the `v >= MIN_READS` filter of RunCounts (genomic_scans.cpp:421-428) as a device stream
// compaction that preserves the reference's output order (chromosome id ascending, '+' then '-',
// window ascending).
#include "gtb_internal.cuh"
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
  int64_t n_regions;
  const int32_t *chrom, *start, *stop;
  const int8_t *strand;
  const int32_t *weight;
  const int64_t *region_offset;
  int64_t interval_base;
};

// one thread per read region; every interval of the region counts once (genomic_intervals.cpp:5039)
__global__ void __launch_bounds__(256) scan_histogram_kernel(ReadView q, int32_t n_chrom, const int32_t *__restrict__ slot_of_chrom,
                                                             const int64_t *__restrict__ hist_off, long long win_step, int op,
                                                             int ignore_strand, ull *__restrict__ hist) {
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
      atomicAdd(hist + hist_off[slot] + win, (ull)w);
    }
  }
}

constexpr int WIN_THREADS = 256;
constexpr int WIN_ITEMS = 8;
constexpr int WIN_TILE = WIN_THREADS * WIN_ITEMS;

__device__ __forceinline__ int find_slot(const int64_t *__restrict__ win_off, int n_slots, int64_t g) {
  int lo = 0, hi = n_slots;            // last slot with win_off[slot] <= g
  while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (win_off[mid] <= g) lo = mid; else hi = mid; }
  return lo;
}

__device__ __forceinline__ long long window_value(const SlotTable &t, const ull *__restrict__ hist, int slot, int64_t k0, int combine) {
  const int64_t base = t.hist_off[slot];
  if (t.spurious[slot]) {
    // Next() hands back current_v[1] of a slot that has no complete window (:5128-5140); with no
    // micro-window at all the reference reads past its allocation -- 0 in practice.
    return t.hist_off[slot + 1] - base >= 1 ? (long long)hist[base] : 0;
  }
  ull sum = 0;
  for (int j = 0; j < combine; j++) sum += hist[base + k0 + j];     // :5066-5073
  return (long long)sum;
}

// MODE 0: count qualifying windows per tile.  MODE 1: write them at tile_offset + local rank.
template <int MODE>
__global__ void __launch_bounds__(WIN_THREADS) scan_windows_kernel(SlotTable t, const ull *__restrict__ hist, int64_t total_windows,
                                                                    int combine, long long min_reads, ull *__restrict__ tile_counts,
                                                                    int32_t *__restrict__ o_chrom, int8_t *__restrict__ o_strand,
                                                                    int64_t *__restrict__ o_win, int64_t *__restrict__ o_value) {
  __shared__ int warp_counts[WIN_THREADS / 32];
  __shared__ ull tile_base;
  const int64_t first = (int64_t)blockIdx.x * WIN_TILE + (int64_t)threadIdx.x * WIN_ITEMS;
  long long val[WIN_ITEMS];
  int slot_of[WIN_ITEMS];
  unsigned keep = 0;
  int slot = first < total_windows ? find_slot(t.win_off, t.n_slots, first) : 0;
#pragma unroll
  for (int i = 0; i < WIN_ITEMS; i++) {
    const int64_t g = first + i;
    val[i] = 0; slot_of[i] = slot;
    if (g < total_windows) {
      while (g >= t.win_off[slot + 1]) slot++;
      slot_of[i] = slot;
      val[i] = window_value(t, hist, slot, g - t.win_off[slot], combine);
      if (val[i] >= min_reads) keep |= 1u << i;
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
    return;
  }
  if (threadIdx.x == 0) tile_base = blockIdx.x == 0 ? 0ull : tile_counts[blockIdx.x - 1];   // inclusive-scanned counts
  __syncthreads();
  int64_t o = (int64_t)tile_base + (warp > 0 ? warp_counts[warp - 1] : 0) + inc - mine;
#pragma unroll
  for (int i = 0; i < WIN_ITEMS; i++) {
    if (keep & (1u << i)) {
      const int s = slot_of[i];
      o_chrom[o] = t.chrom[s]; o_strand[o] = t.strand[s];
      o_win[o] = first + i - t.win_off[s] + 1;                     // 1-based window number, :5109-5112
      o_value[o] = val[i];
      o++;
    }
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
  dbuf<ull> d_hist, d_tile_counts, d_scan_scratch;
  dbuf<int32_t> o_chrom; dbuf<int8_t> o_strand; dbuf<int64_t> o_win, o_value;
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
  GTB_CUDA_OK(sc->ctx, cudaMemsetAsync(sc->d_hist.p, 0, sizeof(ull) * (size_t)std::max<int64_t>(sc->total_micro, 1), sc->ctx->stream));
  sc->n_out = 0;
  return GTB_OK;
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
  if (rc == GTB_OK) rc = upload_vec(ctx, sc->d_spurious, spurious);
  if (rc == GTB_OK) rc = upload_vec(ctx, sc->d_slot_chrom, slot_chrom);
  if (rc == GTB_OK) rc = upload_vec(ctx, sc->d_slot_strand, slot_strand);
  if (rc == GTB_OK) rc = upload_vec(ctx, sc->d_hist_off, hist_off);
  if (rc == GTB_OK) rc = upload_vec(ctx, sc->d_win_off, win_off);
  if (rc == GTB_OK) rc = sc->d_hist.reserve(ctx, (size_t)std::max<int64_t>(sc->total_micro, 1));
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
  sc->d_hist_off.release(); sc->d_win_off.release(); sc->d_hist.release(); sc->d_tile_counts.release(); sc->d_scan_scratch.release();
  sc->o_chrom.release(); sc->o_strand.release(); sc->o_win.release(); sc->o_value.release();
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
  const unsigned grid = gtb_grid_for(q.n_regions, 256, (int64_t)ctx->sm_count * 8);
  GTB_LAUNCH(ctx, "scan_histogram", scan_histogram_kernel, grid, 256, 0, q, sc->n_chrom, sc->d_slot_of_chrom.p, sc->d_hist_off.p,
             (long long)sc->prm.win_step, (int)sc->prm.op, (int)sc->prm.ignore_strand, sc->d_hist.p);
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
  const bool multi = reads->region_offset != nullptr && reads->n_intervals != reads->n_regions;
  if (mem & GTB_MEM_DEVICE) {
    ReadView q{reads->n_regions, reads->chrom, reads->start, reads->stop, reads->strand, reads->weight,
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
    ReadView q{(int64_t)nr, st.chrom.p, st.start.p, st.stop.p, st.strand.p, reads->weight ? st.weight.p : nullptr,
               multi ? st.off.p : nullptr, i0};
    GTB_TRY(scan_accumulate_device(sc, q));
    GTB_CUDA_OK(ctx, cudaEventRecord(st.consumed, ctx->stream));
    st.in_flight = true;
  }
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
  GTB_LAUNCH(ctx, "scan_windows_count", scan_windows_kernel<0>, (unsigned)n_tiles, WIN_THREADS, 0, t, sc->d_hist.p, sc->total_windows,
             sc->combine, (long long)sc->prm.min_reads, sc->d_tile_counts.p, nullptr, nullptr, nullptr, nullptr);
  GTB_TRY(gtb_check_launch(ctx));
  GTB_TRY(gtb_inclusive_scan_u64(ctx, sc->d_tile_counts.p, n_tiles, sc->d_scan_scratch));
  ull total = 0;
  GTB_CUDA_OK(ctx, cudaMemcpyAsync(&total, sc->d_tile_counts.p + (n_tiles - 1), sizeof(ull), cudaMemcpyDeviceToHost, ctx->stream));
  GTB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  sc->n_out = (int64_t)total;
  *n_windows = sc->n_out;
  if (sc->n_out == 0) return GTB_OK;
  GTB_TRY(sc->o_chrom.reserve(ctx, (size_t)sc->n_out)); GTB_TRY(sc->o_strand.reserve(ctx, (size_t)sc->n_out));
  GTB_TRY(sc->o_win.reserve(ctx, (size_t)sc->n_out)); GTB_TRY(sc->o_value.reserve(ctx, (size_t)sc->n_out));
  GTB_LAUNCH(ctx, "scan_windows_emit", scan_windows_kernel<1>, (unsigned)n_tiles, WIN_THREADS, 0, t, sc->d_hist.p, sc->total_windows,
             sc->combine, (long long)sc->prm.min_reads, sc->d_tile_counts.p, sc->o_chrom.p, sc->o_strand.p, sc->o_win.p, sc->o_value.p);
  GTB_TRY(gtb_check_launch(ctx));
  GTB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  return GTB_OK;
}

extern "C" int gtb_scan_fetch(gtb_scan *sc, int64_t first, int64_t count, int32_t *chrom, int8_t *strand, int64_t *win, int64_t *value) {
  if (!sc || first < 0 || count < 0 || first + count > sc->n_out) return GTB_ERR_ARG;
  gtb_ctx *ctx = sc->ctx;
  if (count == 0) return GTB_OK;
  GTB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  if (chrom) GTB_CUDA_OK(ctx, cudaMemcpyAsync(chrom, sc->o_chrom.p + first, (size_t)count * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (strand) GTB_CUDA_OK(ctx, cudaMemcpyAsync(strand, sc->o_strand.p + first, (size_t)count, cudaMemcpyDeviceToHost, ctx->stream));
  if (win) GTB_CUDA_OK(ctx, cudaMemcpyAsync(win, sc->o_win.p + first, (size_t)count * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (value) GTB_CUDA_OK(ctx, cudaMemcpyAsync(value, sc->o_value.p + first, (size_t)count * 8, cudaMemcpyDeviceToHost, ctx->stream));
  GTB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  return GTB_OK;
}
