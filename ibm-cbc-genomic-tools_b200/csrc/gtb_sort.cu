// gtb_sort.cu -- device LSD radix sort of region keys: the global sort of a region set.
//
// Stands in for GenomicRegionSet::RunGlobalSort (genomic_intervals.cpp:4547-4570): the reference bins the regions by chromosome
// (std::map order = strcmp order of the names), optionally by strand ('+' first, every other strand after it, :6140), by START
// >> bin_bits, and list::sort()s every bin with CompareBinnedGenomicRegions (:6045-6049: START ascending, then STOP of the last
// interval DESCENDING); list::sort is stable, so regions equal in all of that keep their input order.  Bins ascend with START,
// so the printed order is the stable sort by the 96-bit key
//      [ chromosome rank : 31 | strand class : 1 ] [ START (sign bit flipped) : 32 ] [ ~STOP (sign bit flipped) : 32 ]
// whatever bin_bits is.  The device produces the permutation; the region objects (text) stay with the caller.
//
// One pass per 8-bit digit, least significant first, each pass stable:
//   sort_hist_kernel      a CTA counts the digits of its tile: a warp takes 32 keys at a time, __match_any_sync groups the lanes
//                         that hold the same digit, the lowest lane of a group adds the group's size to the warp's own
//                         256-bin histogram in shared memory (no atomics: one writer per bin and round)
//   (inclusive scan)      over the [digit][tile] table -> where every (digit, tile) run starts in the output
//   sort_scatter_kernel   the same walk again: a key's place = start of its (digit, tile) run + the keys of that digit in the
//                         warps before it + those in the earlier rounds of its own warp + its rank inside its match group
// Digits on which all keys agree (the high bits of START, most of the chromosome word) are skipped: the host knows the keys.
#include "gtb_internal.cuh"
#include <algorithm>

typedef unsigned long long ull;

namespace {

constexpr int SORT_THREADS = 256;
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int SORT_ROUNDS = 16;                                   // keys per lane and tile
constexpr int SORT_TILE = SORT_THREADS * SORT_ROUNDS;            // 4 096 keys: a warp owns 512 consecutive ones

struct SortArrays { uint32_t *k0, *k1, *k2, *idx; };

__global__ void __launch_bounds__(256) sort_keys_kernel(int64_t n, const int32_t *__restrict__ chrom_rank, const int32_t *__restrict__ start,
                                                        const int32_t *__restrict__ stop, const int8_t *__restrict__ strand, int by_strand, SortArrays a) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    a.k0[i] = ((uint32_t)chrom_rank[i] << 1) | (by_strand && strand[i] != '+' ? 1u : 0u);
    a.k1[i] = (uint32_t)start[i] ^ 0x80000000u;
    a.k2[i] = ~((uint32_t)stop[i] ^ 0x80000000u);
    a.idx[i] = (uint32_t)i;
  }
}

// digit of key i in the pass's word, 256 (a bin of its own, after all real ones) for the padding beyond n
__device__ __forceinline__ uint32_t sort_digit(const uint32_t *__restrict__ word, int64_t i, int64_t n, int shift) {
  return i < n ? (word[i] >> shift) & 0xFFu : 256u;
}

__global__ void __launch_bounds__(SORT_THREADS) sort_hist_kernel(int64_t n, const uint32_t *__restrict__ word, int shift, int64_t n_tiles, ull *__restrict__ table) {
  __shared__ uint32_t s_hist[SORT_WARPS][257];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < SORT_WARPS * 257; i += SORT_THREADS) (&s_hist[0][0])[i] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * SORT_TILE + (int64_t)warp * (32 * SORT_ROUNDS);
#pragma unroll 4
  for (int r = 0; r < SORT_ROUNDS; r++) {
    const uint32_t d = sort_digit(word, base + r * 32 + lane, n, shift);
    const uint32_t peers = __match_any_sync(0xffffffffu, d);
    if ((peers & ((1u << lane) - 1u)) == 0u) s_hist[warp][d] += __popc(peers);
    __syncwarp();
  }
  __syncthreads();
  for (int d = threadIdx.x; d < 256; d += SORT_THREADS) {
    uint32_t c = 0;
#pragma unroll
    for (int w = 0; w < SORT_WARPS; w++) c += s_hist[w][d];
    table[(int64_t)d * n_tiles + blockIdx.x] = c;                      // digit-major: the scan of the table is the output order
  }
}

__global__ void __launch_bounds__(SORT_THREADS) sort_scatter_kernel(int64_t n, SortArrays in, SortArrays out, int which, int shift, int64_t n_tiles,
                                                                      const ull *__restrict__ table_scan) {
  __shared__ uint32_t s_hist[SORT_WARPS][257];
  __shared__ ull s_base[SORT_WARPS][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t *word = which == 0 ? in.k0 : which == 1 ? in.k1 : in.k2;
  for (int i = threadIdx.x; i < SORT_WARPS * 257; i += SORT_THREADS) (&s_hist[0][0])[i] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * SORT_TILE + (int64_t)warp * (32 * SORT_ROUNDS);
  uint32_t dig[SORT_ROUNDS];
#pragma unroll
  for (int r = 0; r < SORT_ROUNDS; r++) {
    dig[r] = sort_digit(word, base + r * 32 + lane, n, shift);
    const uint32_t peers = __match_any_sync(0xffffffffu, dig[r]);
    if ((peers & ((1u << lane) - 1u)) == 0u) s_hist[warp][dig[r]] += __popc(peers);
    __syncwarp();
  }
  __syncthreads();
  // where warp w's keys of digit d begin: the (d, tile) run's start (exclusive scan = inclusive scan of the entry before) + the warps before w
  for (int d = threadIdx.x; d < 256; d += SORT_THREADS) {
    const int64_t e = (int64_t)d * n_tiles + blockIdx.x;
    ull run = e > 0 ? table_scan[e - 1] : 0ull;
#pragma unroll
    for (int w = 0; w < SORT_WARPS; w++) { s_base[w][d] = run; run += s_hist[w][d]; }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < SORT_ROUNDS; r++) {
    const int64_t i = base + r * 32 + lane;
    const uint32_t d = dig[r];
    const uint32_t peers = __match_any_sync(0xffffffffu, d);
    const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
    if (d < 256u) {
      const ull pos = s_base[warp][d] + rank;
      GTB_ASSERT(pos < (ull)n);
      out.k0[pos] = in.k0[i]; out.k1[pos] = in.k1[i]; out.k2[pos] = in.k2[i]; out.idx[pos] = in.idx[i];
    }
    __syncwarp();
    if (d < 256u && rank == 0u) s_base[warp][d] += __popc(peers);      // the next round's keys of this digit come after these
    __syncwarp();
  }
}

__global__ void __launch_bounds__(256) sort_perm_kernel(int64_t n, const uint32_t *__restrict__ idx, int64_t *__restrict__ perm) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) perm[i] = (int64_t)idx[i];
}

}  // namespace

extern "C" int gtb_sort_regions(gtb_ctx *ctx, int64_t n, const int32_t *chrom_rank, const int32_t *start, const int32_t *stop, const int8_t *strand,
                                int by_strand, int64_t *perm) {
  if (!ctx || n < 0 || (n > 0 && (!chrom_rank || !start || !stop || !strand || !perm))) return GTB_ERR_ARG;
  if (n == 0) return GTB_OK;
  if (n >= ((int64_t)1 << 31)) return gtb_fail(ctx, GTB_ERR_UNSUPPORTED, "more than 2^31 regions in one sort");
  GTB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  // which bits differ between keys: digits on which all keys agree are not sorted on
  uint32_t vary[3] = {0, 0, 0};
  {
    const uint32_t f0 = ((uint32_t)chrom_rank[0] << 1) | (by_strand && strand[0] != '+' ? 1u : 0u), f1 = (uint32_t)start[0], f2 = (uint32_t)stop[0];
    for (int64_t i = 0; i < n; i++) {
      if (chrom_rank[i] < 0 || chrom_rank[i] >= (1 << 30)) return gtb_fail(ctx, GTB_ERR_ARG, "chromosome rank out of range");
      vary[0] |= (((uint32_t)chrom_rank[i] << 1) | (by_strand && strand[i] != '+' ? 1u : 0u)) ^ f0;
      vary[1] |= (uint32_t)start[i] ^ f1;
      vary[2] |= (uint32_t)stop[i] ^ f2;
    }
  }
  dbuf<int32_t> d_chrom, d_start, d_stop;
  dbuf<int8_t> d_strand;
  dbuf<uint32_t> buf[2][4];
  dbuf<ull> d_table, d_scratch;
  dbuf<int64_t> d_perm;
  const size_t nn = (size_t)n;
  const int64_t n_tiles = (n + SORT_TILE - 1) / SORT_TILE;
  auto cleanup = [&]() {
    d_chrom.release(); d_start.release(); d_stop.release(); d_strand.release(); d_table.release(); d_scratch.release(); d_perm.release();
    for (auto &b : buf) for (auto &x : b) x.release();
  };
  int rc = GTB_OK;
  auto run = [&]() -> int {
    GTB_TRY(d_chrom.reserve(ctx, nn)); GTB_TRY(d_start.reserve(ctx, nn)); GTB_TRY(d_stop.reserve(ctx, nn)); GTB_TRY(d_strand.reserve(ctx, nn));
    for (auto &b : buf) for (auto &x : b) GTB_TRY(x.reserve(ctx, nn));
    GTB_TRY(d_table.reserve(ctx, (size_t)256 * (size_t)n_tiles)); GTB_TRY(d_perm.reserve(ctx, nn));
    GTB_CUDA_OK(ctx, cudaMemcpyAsync(d_chrom.p, chrom_rank, nn * 4, cudaMemcpyHostToDevice, ctx->stream));
    GTB_CUDA_OK(ctx, cudaMemcpyAsync(d_start.p, start, nn * 4, cudaMemcpyHostToDevice, ctx->stream));
    GTB_CUDA_OK(ctx, cudaMemcpyAsync(d_stop.p, stop, nn * 4, cudaMemcpyHostToDevice, ctx->stream));
    GTB_CUDA_OK(ctx, cudaMemcpyAsync(d_strand.p, strand, nn, cudaMemcpyHostToDevice, ctx->stream));
    ctx->h2d_bytes += (int64_t)nn * 13;
    SortArrays a[2] = {{buf[0][0].p, buf[0][1].p, buf[0][2].p, buf[0][3].p}, {buf[1][0].p, buf[1][1].p, buf[1][2].p, buf[1][3].p}};
    const unsigned g = gtb_grid_for(n, 256, (int64_t)ctx->sm_count * 8);
    GTB_LAUNCH(ctx, "sort_keys", sort_keys_kernel, g, 256, 0, n, d_chrom.p, d_start.p, d_stop.p, d_strand.p, by_strand, a[0]);
    GTB_TRY(gtb_check_launch(ctx));
    int cur = 0;
    for (int which = 2; which >= 0; which--)                           // least significant word first: ~STOP, START, chromosome | strand
      for (int shift = 0; shift < 32; shift += 8) {
        if (((vary[which] >> shift) & 0xFFu) == 0u) continue;
        const uint32_t *word = which == 0 ? a[cur].k0 : which == 1 ? a[cur].k1 : a[cur].k2;
        GTB_LAUNCH(ctx, "sort_hist", sort_hist_kernel, (unsigned)n_tiles, SORT_THREADS, 0, n, word, shift, n_tiles, d_table.p);
        GTB_TRY(gtb_check_launch(ctx));
        GTB_TRY(gtb_inclusive_scan_u64(ctx, d_table.p, (int64_t)256 * n_tiles, d_scratch));
        GTB_LAUNCH(ctx, "sort_scatter", sort_scatter_kernel, (unsigned)n_tiles, SORT_THREADS, 0, n, a[cur], a[cur ^ 1], which, shift, n_tiles, (const ull *)d_table.p);
        GTB_TRY(gtb_check_launch(ctx));
        cur ^= 1;
      }
    GTB_LAUNCH(ctx, "sort_perm", sort_perm_kernel, g, 256, 0, n, (const uint32_t *)a[cur].idx, d_perm.p);
    GTB_TRY(gtb_check_launch(ctx));
    GTB_CUDA_OK(ctx, cudaMemcpyAsync(perm, d_perm.p, nn * 8, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->d2h_bytes += (int64_t)nn * 8;
    GTB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    return GTB_OK;
  };
  rc = run();
  if (rc != GTB_OK) cudaStreamSynchronize(ctx->stream);
  cleanup();
  return rc;
}
