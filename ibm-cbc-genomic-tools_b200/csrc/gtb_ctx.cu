// gtb_ctx.cu -- context life cycle, stream selection, launch accounting / CUDA-event profiling,
// and the multi-block inclusive scan used by the finalisation kernels.
#include "gtb_internal.cuh"
#include <thread>

extern "C" int gtb_abi_version(void) { return GTB200_ABI_VERSION; }

extern "C" int gtb_ctx_create(int device, gtb_ctx **out) {
  if (!out) return GTB_ERR_ARG;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) return GTB_ERR_NO_DEVICE;   // no CPU fallback, by design
  if (device < 0 || device >= n) return GTB_ERR_ARG;
  gtb_ctx *ctx = new gtb_ctx();
  ctx->device = device;
  if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return GTB_ERR_CUDA; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) {
    ctx->sm_count = prop.multiProcessorCount;
    ctx->smem_optin = prop.sharedMemPerBlockOptin;
  }
  if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete ctx; return GTB_ERR_CUDA;
  }
  ctx->stream = ctx->own_stream;
  *out = ctx;
  return GTB_OK;
}

extern "C" void gtb_ctx_destroy(gtb_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (auto &p : ctx->pending) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  gtb_ingest_destroy(ctx->ingest);
  for (auto &sl : ctx->slots) {
    if (sl.meta) cudaFreeHost(sl.meta);
    if (sl.start) cudaFreeHost(sl.start);
    if (sl.h2d_done) cudaEventDestroy(sl.h2d_done);
  }
  delete ctx;
}

// Lazily starts the packing pool and hands out the next pinned staging slot, sized for n_intervals and no longer
// referenced by a copy in flight.  Returns GTB_ERR_UNSUPPORTED when packing is off (GTB_INGEST_THREADS=0, one core).
int gtb_ctx_ingest_ready(gtb_ctx *ctx, size_t n_intervals, gtb_pinned_slot **slot) {
  if (!ctx->ingest_tried) {
    ctx->ingest_tried = true;
    int threads = (int)std::thread::hardware_concurrency();
    if (const char *lw = getenv("LOCAL_WORLD_SIZE")) { const int w = atoi(lw); if (w > 1) threads /= w; }   // torchrun: ranks share the host
    threads = threads * 3 / 4;                         // leave cores to the caller's thread and the driver (measured: 12 of 16 is best)
    if (threads > 32) threads = 32;
    if (const char *env = getenv("GTB_INGEST_THREADS")) threads = atoi(env);
    if (threads >= 2) ctx->ingest = gtb_ingest_create(threads);
  }
  if (!ctx->ingest) return GTB_ERR_UNSUPPORTED;
  gtb_pinned_slot &sl = ctx->slots[ctx->next_slot];
  ctx->next_slot = (ctx->next_slot + 1) % 3;
  if (sl.in_flight) { GTB_CUDA_OK(ctx, cudaEventSynchronize(sl.h2d_done)); sl.in_flight = false; }
  if (sl.cap < n_intervals) {
    if (sl.meta) cudaFreeHost(sl.meta);
    if (sl.start) cudaFreeHost(sl.start);
    sl.meta = nullptr; sl.start = nullptr; sl.cap = 0;
    GTB_CUDA_OK(ctx, cudaHostAlloc((void **)&sl.meta, n_intervals * sizeof(uint32_t), cudaHostAllocDefault));
    GTB_CUDA_OK(ctx, cudaHostAlloc((void **)&sl.start, n_intervals * sizeof(int32_t), cudaHostAllocDefault));
    sl.cap = n_intervals;
  }
  if (!sl.h2d_done) GTB_CUDA_OK(ctx, cudaEventCreateWithFlags(&sl.h2d_done, cudaEventDisableTiming));
  *slot = &sl;
  return GTB_OK;
}

extern "C" int gtb_ctx_set_stream(gtb_ctx *ctx, void *cuda_stream) {
  if (!ctx) return GTB_ERR_ARG;
  GTB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
  return GTB_OK;
}

extern "C" void *gtb_ctx_get_stream(const gtb_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

extern "C" int gtb_ctx_synchronize(gtb_ctx *ctx) {
  if (!ctx) return GTB_ERR_ARG;
  GTB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->copy_stream));
  GTB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  return GTB_OK;
}

extern "C" const char *gtb_ctx_last_error(const gtb_ctx *ctx) { return ctx ? ctx->last_error.c_str() : "null context"; }

extern "C" int gtb_ctx_transfer_stats(const gtb_ctx *ctx, int64_t *h2d_bytes, int64_t *d2h_bytes, int64_t *packed_chunks, int64_t *raw_chunks) {
  if (!ctx) return GTB_ERR_ARG;
  if (h2d_bytes) *h2d_bytes = ctx->h2d_bytes;
  if (d2h_bytes) *d2h_bytes = ctx->d2h_bytes;
  if (packed_chunks) *packed_chunks = ctx->packed_chunks;
  if (raw_chunks) *raw_chunks = ctx->raw_chunks;
  return GTB_OK;
}

extern "C" int64_t gtb_ctx_launch_count(const gtb_ctx *ctx) { return ctx ? ctx->launches : 0; }

static void drain_pending(gtb_ctx *ctx) {
  for (auto &p : ctx->pending) {
    float ms = 0.f;
    cudaEventSynchronize(p.b);
    if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
      auto &s = ctx->stats[p.name];
      s.launches++; s.total_ms += ms;
    }
    cudaEventDestroy(p.a); cudaEventDestroy(p.b);
  }
  ctx->pending.clear();
}

extern "C" int gtb_ctx_profile(gtb_ctx *ctx, int enable) {
  if (!ctx) return GTB_ERR_ARG;
  drain_pending(ctx);
  ctx->profiling = enable != 0;
  if (enable) { ctx->stats.clear(); ctx->launches = 0; }
  return GTB_OK;
}

extern "C" int gtb_ctx_profile_report(gtb_ctx *ctx, char *buf, size_t buf_size) {
  if (!ctx || !buf || buf_size < 3) return GTB_ERR_ARG;
  drain_pending(ctx);
  std::string s = "{";
  bool first = true;
  for (auto &kv : ctx->stats) {
    char line[256];
    snprintf(line, sizeof line, "%s\"%s\": {\"launches\": %lld, \"total_ms\": %.6f}", first ? "" : ", ",
             kv.first.c_str(), (long long)kv.second.launches, kv.second.total_ms);
    s += line; first = false;
  }
  s += "}";
  if (s.size() + 1 > buf_size) return GTB_ERR_ARG;
  memcpy(buf, s.c_str(), s.size() + 1);
  return GTB_OK;
}

// ------------------------------------------------------------------------------------------------
// inclusive scan (uint64, wrapping).  Three launches: per-tile scan + tile totals, scan of totals
// by one block, add-back.  Used on arrays of at most a few million evaluation points, so it is
// sized for simplicity; traffic is 3 reads + 2 writes per element.
// ------------------------------------------------------------------------------------------------
namespace {
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ unsigned long long block_exclusive_scan(unsigned long long v, unsigned long long *total) {
  __shared__ unsigned long long warp_sums[SCAN_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned long long inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    unsigned long long t = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += t;
  }
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    unsigned long long w = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0ull;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      unsigned long long t = __shfl_up_sync(0xffffffffu, w, d);
      if (lane >= d) w += t;
    }
    if (lane < SCAN_THREADS / 32) warp_sums[lane] = w;
  }
  __syncthreads();
  unsigned long long base = warp > 0 ? warp_sums[warp - 1] : 0ull;
  if (total) *total = warp_sums[SCAN_THREADS / 32 - 1];
  unsigned long long r = base + inc - v;
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_tiles_kernel(const unsigned long long *src, unsigned long long *d, int64_t n, int64_t plane_stride,
                                                                   unsigned long long *tile_totals) {
  src += (int64_t)blockIdx.y * plane_stride; d += (int64_t)blockIdx.y * plane_stride; tile_totals += (int64_t)blockIdx.y * gridDim.x;
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  unsigned long long v[SCAN_ITEMS], sum = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++) { v[i] = base + i < n ? src[base + i] : 0ull; sum += v[i]; }
  unsigned long long total;
  unsigned long long ex = block_exclusive_scan(sum, &total);
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++) { ex += v[i]; if (base + i < n) d[base + i] = ex; }
  if (threadIdx.x == 0) tile_totals[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_totals_kernel(unsigned long long *totals, int64_t n_tiles) {
  totals += (int64_t)blockIdx.y * n_tiles;
  // one block walks the tile totals in chunks, carrying the running sum (exclusive result in place)
  __shared__ unsigned long long carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int64_t base = 0; base < n_tiles; base += SCAN_THREADS) {
    int64_t i = base + threadIdx.x;
    unsigned long long v = i < n_tiles ? totals[i] : 0ull, total;
    unsigned long long ex = block_exclusive_scan(v, &total);
    if (i < n_tiles) totals[i] = carry + ex;
    __syncthreads();
    if (threadIdx.x == 0) carry += total;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_addback_kernel(unsigned long long *d, int64_t n, int64_t plane_stride, const unsigned long long *tile_offsets) {
  d += (int64_t)blockIdx.y * plane_stride; tile_offsets += (int64_t)blockIdx.y * gridDim.x;
  const unsigned long long off = tile_offsets[blockIdx.x];
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
  for (int i = threadIdx.x; i < SCAN_TILE; i += SCAN_THREADS)
    if (base + i < n) d[base + i] += off;
}
}  // namespace

int gtb_inclusive_scan_planes_u64(gtb_ctx *ctx, const unsigned long long *src, unsigned long long *dst, int64_t n, int planes,
                                    int64_t plane_stride, dbuf<unsigned long long> &scratch) {
  if (n <= 0 || planes <= 0) return GTB_OK;
  const int64_t n_tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  GTB_TRY(scratch.reserve(ctx, (size_t)n_tiles * planes));
  GTB_LAUNCH(ctx, "scan_tiles", scan_tiles_kernel, dim3((unsigned)n_tiles, (unsigned)planes), SCAN_THREADS, 0, src, dst, n, plane_stride, scratch.p);
  if (n_tiles > 1) {
    GTB_LAUNCH(ctx, "scan_totals", scan_totals_kernel, dim3(1, (unsigned)planes), SCAN_THREADS, 0, scratch.p, n_tiles);
    GTB_LAUNCH(ctx, "scan_addback", scan_addback_kernel, dim3((unsigned)n_tiles, (unsigned)planes), SCAN_THREADS, 0, dst, n, plane_stride, scratch.p);
  }
  return gtb_check_launch(ctx);
}

int gtb_inclusive_scan_u64(gtb_ctx *ctx, unsigned long long *d, int64_t n, dbuf<unsigned long long> &scratch) {
  return gtb_inclusive_scan_planes_u64(ctx, d, d, n, 1, n, scratch);
}

// ---- file-order scatter of gathered per-shard values (the multi-GPU driver's last step) ------------------------------------------
namespace {
__global__ void __launch_bounds__(256) gather_u64_kernel(const unsigned long long *__restrict__ table, const int64_t *__restrict__ index,
                                                         int64_t n, unsigned long long *__restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) out[k] = table[index[k]];
}
}  // namespace

extern "C" int gtb_gather_u64(gtb_ctx *ctx, const uint64_t *table, const int64_t *index, int64_t n, uint64_t *out, void *cuda_stream) {
  if (!ctx || n < 0 || (n > 0 && (!table || !index || !out))) return GTB_ERR_ARG;
  if (n == 0) return GTB_OK;
  GTB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream;
  ctx->launches++;
  gather_u64_kernel<<<gtb_grid_for(n, 256, (int64_t)ctx->sm_count * 8), 256, 0, st>>>((const unsigned long long *)table, index, n, (unsigned long long *)out);
  return gtb_check_launch(ctx);
}
