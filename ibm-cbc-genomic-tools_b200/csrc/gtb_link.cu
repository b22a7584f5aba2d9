// gtb_link.cu -- linking consecutive regions of a sorted stream (genomic_regions link).
//
// GenomicRegionSet::RunGlobalLink (genomic_intervals.cpp:4607-4644) walks single-interval regions sorted by chromosome /
// (strand) / start: a region joins the linked region in progress if it is compatible with its head (same chromosome, and
// strand if the stream is sorted by strand) and starts no further than max_difference beyond the stop reached so far; the
// linked region is printed as [head's start, largest stop].  That loop is sequential, but for max_difference >= 0 it has a
// closed form: with P[k] = the largest stop among regions 0..k of k's group (groups are contiguous in a sorted stream),
//     region k starts a linked region  <=>  k is the first of its group  or  start[k] - P[k-1] > max_difference
// (an earlier linked region of the group ended below start - max_difference of the head that followed it, hence below every
// later start - max_difference: the prefix maximum and the maximum of the region in progress decide alike), and the stop of a
// linked region is P at its last member.  So: one inclusive prefix-maximum scan over keys (group rank << 32 | biased stop -- a
// later group outranks every earlier one, which makes the plain scan a segmented one), head flags, an inclusive prefix sum of
// the flags, and a scatter of (head index, linked stop) per linked region.  Negative max_difference (regions must overlap by
// that much) breaks the equivalence; the caller keeps the sequential loop for it.
#include "gtb_internal.cuh"
#include <algorithm>

typedef unsigned long long ull;

namespace {
constexpr int LK_THREADS = 256, LK_ITEMS = 8, LK_TILE = LK_THREADS * LK_ITEMS;

__device__ __forceinline__ ull lk_key(int32_t group, int32_t stop) { return ((ull)(uint32_t)group << 32) | (ull)((uint32_t)stop ^ 0x80000000u); }
__device__ __forceinline__ int32_t lk_stop(ull key) { return (int32_t)((uint32_t)key ^ 0x80000000u); }

// block-wide EXCLUSIVE scan of one value per thread under `max` (0 for thread 0); *total = the block's maximum
__device__ __forceinline__ ull lk_block_exclusive_max(ull v, ull *total) {
  __shared__ ull warp_max[LK_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  ull incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { const ull t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl = max(incl, t); }
  if (lane == 31) warp_max[warp] = incl;
  __syncthreads();
  ull before = 0;
  for (int w = 0; w < warp; w++) before = max(before, warp_max[w]);
  if (total) { ull t = 0; for (int w = 0; w < LK_THREADS / 32; w++) t = max(t, warp_max[w]); *total = t; }
  ull prev = __shfl_up_sync(0xffffffffu, incl, 1);
  if (lane == 0) prev = 0;
  __syncthreads();
  return max(before, prev);
}

// phase 1: per tile, the inclusive prefix maximum of the keys inside the tile and the tile's maximum
__global__ void __launch_bounds__(LK_THREADS) link_scan_tiles_kernel(const int32_t *__restrict__ group, const int32_t *__restrict__ stop, int64_t n,
                                                                      ull *__restrict__ pmax, ull *__restrict__ tile_max) {
  const int64_t base = (int64_t)blockIdx.x * LK_TILE + (int64_t)threadIdx.x * LK_ITEMS;
  ull v[LK_ITEMS], run = 0;
#pragma unroll
  for (int i = 0; i < LK_ITEMS; i++) { v[i] = base + i < n ? lk_key(group[base + i], stop[base + i]) : 0ull; run = max(run, v[i]); v[i] = run; }
  ull total;
  const ull before = lk_block_exclusive_max(run, &total);
#pragma unroll
  for (int i = 0; i < LK_ITEMS; i++) if (base + i < n) pmax[base + i] = max(v[i], before);
  if (threadIdx.x == 0) tile_max[blockIdx.x] = total;
}

// phase 2: one block turns the tiles' maxima into "maximum of all tiles before this one"
__global__ void __launch_bounds__(LK_THREADS) link_scan_carry_kernel(ull *tile_max, int64_t n_tiles) {
  __shared__ ull carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int64_t base = 0; base < n_tiles; base += LK_THREADS) {
    const int64_t i = base + threadIdx.x;
    const ull v = i < n_tiles ? tile_max[i] : 0ull;
    ull total;
    const ull before = lk_block_exclusive_max(v, &total);
    if (i < n_tiles) tile_max[i] = max(before, carry);
    __syncthreads();
    if (threadIdx.x == 0) carry = max(carry, total);
    __syncthreads();
  }
}

// phase 3: head flags (as 0/1 in a u64 array for the prefix sum that follows); pmax becomes the global prefix maximum
__global__ void __launch_bounds__(LK_THREADS) link_heads_kernel(const int32_t *__restrict__ group, const int32_t *__restrict__ start, int64_t n, long long max_difference,
                                                                 ull *__restrict__ pmax, const ull *__restrict__ tile_before, ull *__restrict__ head) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
    // the prefix maximum up to k - 1 (global): the tile-local value joined with what the tiles before reached
    bool is_head = true;
    if (k > 0) {
      const ull p = max(pmax[k - 1], tile_before[(k - 1) / LK_TILE]);
      is_head = (uint32_t)(p >> 32) != (uint32_t)group[k] || (long long)start[k] - (long long)lk_stop(p) > max_difference;
    }
    head[k] = is_head ? 1ull : 0ull;
  }
}

// phase 4 (after the inclusive prefix sum of the flags): every head writes its index, every last member of a linked region its stop
__global__ void __launch_bounds__(LK_THREADS) link_emit_kernel(const ull *__restrict__ run_of, const ull *__restrict__ pmax, const ull *__restrict__ tile_before, int64_t n,
                                                                int64_t *__restrict__ run_head, int32_t *__restrict__ run_stop) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
    const ull r = run_of[k];                                             // 1-based number of k's linked region
    if (k == 0 || run_of[k - 1] != r) run_head[r - 1] = k;
    if (k == n - 1 || run_of[k + 1] != r) run_stop[r - 1] = lk_stop(max(pmax[k], tile_before[k / LK_TILE]));
  }
}
}  // namespace

extern "C" int gtb_link_regions(gtb_ctx *ctx, int64_t n, const int32_t *group_rank, const int32_t *start, const int32_t *stop, int64_t max_difference,
                                int64_t *n_linked, int64_t *head_index, int32_t *linked_stop) {
  if (!ctx || n < 0 || !n_linked || (n > 0 && (!group_rank || !start || !stop || !head_index || !linked_stop))) return GTB_ERR_ARG;
  if (max_difference < 0) return gtb_fail(ctx, GTB_ERR_UNSUPPORTED, "negative max_difference has no scan form");
  *n_linked = 0;
  if (n == 0) return GTB_OK;
  GTB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  dbuf<int32_t> d_group, d_start, d_stop, d_run_stop;
  dbuf<ull> d_pmax, d_tile, d_head, d_scratch;
  dbuf<int64_t> d_run_head;
  const int64_t n_tiles = (n + LK_TILE - 1) / LK_TILE;
  int rc = GTB_OK;
  auto body = [&]() -> int {
    GTB_TRY(d_group.reserve(ctx, (size_t)n)); GTB_TRY(d_start.reserve(ctx, (size_t)n)); GTB_TRY(d_stop.reserve(ctx, (size_t)n));
    GTB_TRY(d_pmax.reserve(ctx, (size_t)n)); GTB_TRY(d_head.reserve(ctx, (size_t)n)); GTB_TRY(d_tile.reserve(ctx, (size_t)n_tiles));
    GTB_CUDA_OK(ctx, cudaMemcpyAsync(d_group.p, group_rank, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    GTB_CUDA_OK(ctx, cudaMemcpyAsync(d_start.p, start, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    GTB_CUDA_OK(ctx, cudaMemcpyAsync(d_stop.p, stop, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    GTB_LAUNCH(ctx, "link_scan_tiles", link_scan_tiles_kernel, (unsigned)n_tiles, LK_THREADS, 0, d_group.p, d_stop.p, n, d_pmax.p, d_tile.p);
    GTB_LAUNCH(ctx, "link_scan_carry", link_scan_carry_kernel, 1, LK_THREADS, 0, d_tile.p, n_tiles);
    const unsigned grid = gtb_grid_for(n, LK_THREADS, (int64_t)ctx->sm_count * 8);
    GTB_LAUNCH(ctx, "link_heads", link_heads_kernel, grid, LK_THREADS, 0, d_group.p, d_start.p, n, (long long)max_difference, d_pmax.p, d_tile.p, d_head.p);
    GTB_TRY(gtb_check_launch(ctx));
    GTB_TRY(gtb_inclusive_scan_u64(ctx, d_head.p, n, d_scratch));
    ull total = 0;
    GTB_CUDA_OK(ctx, cudaMemcpyAsync(&total, d_head.p + (n - 1), sizeof(ull), cudaMemcpyDeviceToHost, ctx->stream));
    GTB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    *n_linked = (int64_t)total;
    GTB_TRY(d_run_head.reserve(ctx, (size_t)total)); GTB_TRY(d_run_stop.reserve(ctx, (size_t)total));
    GTB_LAUNCH(ctx, "link_emit", link_emit_kernel, grid, LK_THREADS, 0, d_head.p, d_pmax.p, d_tile.p, n, d_run_head.p, d_run_stop.p);
    GTB_TRY(gtb_check_launch(ctx));
    GTB_CUDA_OK(ctx, cudaMemcpyAsync(head_index, d_run_head.p, (size_t)total * 8, cudaMemcpyDeviceToHost, ctx->stream));
    GTB_CUDA_OK(ctx, cudaMemcpyAsync(linked_stop, d_run_stop.p, (size_t)total * 4, cudaMemcpyDeviceToHost, ctx->stream));
    GTB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    return GTB_OK;
  };
  rc = body();
  d_group.release(); d_start.release(); d_stop.release(); d_run_stop.release(); d_pmax.release(); d_tile.release(); d_head.release();
  d_scratch.release(); d_run_head.release();
  return rc;
}
