// gtb_internal.cuh -- shared internals of libgtb200 (context, device buffers, launch accounting).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>
#include <map>
#include "../../include/gtb200.h"

#define GTB_SM_COUNT_FALLBACK 148

// Bounds checks on the data-dependent indices of the kernels.  compute-sanitizer is closed on the pool this was developed
// on (profiles/r2j_sanitizer_closed.txt), so `make checked` builds lib/libgtb200_checked.so with device-side assert() at
// every such index; the GPU tests run against it once (GTB200_LIB=...; profiles/scripts/r2k_checked.sh).  In the product
// build the macro is empty.
#ifdef GTB_BOUNDS_CHECK
#include <assert.h>
#define GTB_ASSERT(cond) assert(cond)
#else
#define GTB_ASSERT(cond) ((void)0)
#endif

struct gtb_kernel_stat {
  int64_t launches = 0;
  double total_ms = 0.0;
};

// host-side re-encoding of query chunks before they cross PCIe (gtb_ingest.cpp)
struct gtb_ingest;
gtb_ingest *gtb_ingest_create(int threads);
void gtb_ingest_destroy(gtb_ingest *p);
int gtb_ingest_threads(const gtb_ingest *p);
int gtb_ingest_pack(gtb_ingest *p, const int32_t *chrom, const int32_t *start, const int32_t *stop, const int8_t *strand, int64_t n,
                    uint32_t *meta, int32_t *start_out);

int gtb_ingest_pack_width(gtb_ingest *p, int width, const int32_t *chrom, const int32_t *start, const int32_t *stop, const int8_t *strand,
                          int64_t n, void *meta, int32_t *start_out, uint32_t *len0);

struct gtb_pinned_slot {                   // pinned host staging for one packed chunk
  uint32_t *meta = nullptr;
  int32_t *start = nullptr;
  size_t cap = 0;                          // intervals
  cudaEvent_t h2d_done = nullptr;
  bool in_flight = false;
};

struct gtb_ctx {
  int device = 0;
  int sm_count = GTB_SM_COUNT_FALLBACK;
  size_t smem_optin = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t copy_stream = nullptr;      // H2D staging overlaps kernels on `stream`
  cudaStream_t stream = nullptr;           // where kernels go (own_stream or caller's)
  std::string last_error;
  int64_t launches = 0;
  bool profiling = false;
  struct pending_event { const char *name; cudaEvent_t a, b; };
  std::vector<pending_event> pending;
  std::map<std::string, gtb_kernel_stat> stats;
  // packed host->device path
  gtb_ingest *ingest = nullptr;
  bool ingest_tried = false;
  gtb_pinned_slot slots[3];
  int next_slot = 0;
  int64_t packed_chunks = 0, raw_chunks = 0;   // how the host-resident chunks travelled (diagnostics)
  int64_t h2d_bytes = 0, d2h_bytes = 0;        // bytes of query batches / results that crossed the host link
  double pack_rate = 0.0;                      // measured packing throughput of this context's pool, intervals/s (0: not yet known)
  int64_t pack_timed = 0;                      // chunks the pool has been timed on (the first one does not count)
  int64_t pack_skipped = 0;                    // chunks sent raw because packing would have been the slower leg
  int pack_width = 1;                          // bytes of meta per interval the next chunk tries first (1, 2, 4); widened when a chunk does not fit
  int64_t pack_wide_chunks = 0;                // chunks since the width was last widened (a narrower form is retried now and then)
};
int gtb_ctx_ingest_ready(gtb_ctx *ctx, size_t n_intervals, gtb_pinned_slot **slot);   // gtb_ctx.cu

// ---- error plumbing ---------------------------------------------------------------------------
#define GTB_CUDA_OK(ctx, call)                                                              \
  do {                                                                                      \
    cudaError_t e__ = (call);                                                               \
    if (e__ != cudaSuccess) {                                                               \
      char b__[512];                                                                        \
      snprintf(b__, sizeof b__, "%s:%d: %s -> %s", __FILE__, __LINE__, #call,               \
               cudaGetErrorString(e__));                                                    \
      (ctx)->last_error = b__;                                                              \
      return e__ == cudaErrorMemoryAllocation ? GTB_ERR_NOMEM : GTB_ERR_CUDA;               \
    }                                                                                       \
  } while (0)

#define GTB_TRY(expr)                  \
  do {                                 \
    int rc__ = (expr);                 \
    if (rc__ != GTB_OK) return rc__;   \
  } while (0)

static inline int gtb_fail(gtb_ctx *ctx, int code, const char *msg) {
  if (ctx) ctx->last_error = msg;
  return code;
}

// ---- growable device buffer -------------------------------------------------------------------
template <typename T>
struct dbuf {
  T *p = nullptr;
  size_t cap = 0;   // elements
  int reserve(gtb_ctx *ctx, size_t n) {
    if (n <= cap) return GTB_OK;
    if (p) { cudaFree(p); p = nullptr; cap = 0; }
    size_t want = n + n / 8 + 256;
    GTB_CUDA_OK(ctx, cudaMalloc((void **)&p, want * sizeof(T)));
    cap = want;
    return GTB_OK;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// ---- kernel launch accounting -----------------------------------------------------------------
// Every kernel goes through GTB_LAUNCH so that gtb_ctx_launch_count() is exact and, with profiling
// on, each launch is bracketed by CUDA events on the launching stream.
struct gtb_launch_scope {
  gtb_ctx *ctx; const char *name; cudaEvent_t a = nullptr, b = nullptr;
  gtb_launch_scope(gtb_ctx *c, const char *n) : ctx(c), name(n) {
    ctx->launches++;
    if (ctx->profiling) {
      cudaEventCreate(&a); cudaEventCreate(&b);
      cudaEventRecord(a, ctx->stream);
    }
  }
  ~gtb_launch_scope() {
    if (ctx->profiling) {
      cudaEventRecord(b, ctx->stream);
      ctx->pending.push_back({name, a, b});
    }
  }
};
#define GTB_LAUNCH(ctx, name, kernel, grid, block, smem, ...)                       \
  do {                                                                              \
    gtb_launch_scope scope__((ctx), (name));                                        \
    kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);                \
  } while (0)

static inline int gtb_check_launch(gtb_ctx *ctx) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    ctx->last_error = std::string("kernel launch failed: ") + cudaGetErrorString(e);
    return GTB_ERR_CUDA;
  }
  return GTB_OK;
}

static inline unsigned gtb_grid_for(int64_t n_items, int per_block, int64_t max_blocks) {
  int64_t g = (n_items + per_block - 1) / per_block;
  if (g < 1) g = 1;
  if (g > max_blocks) g = max_blocks;
  return (unsigned)g;
}

// in-place inclusive prefix sum over n uint64 values on ctx->stream (gtb_scan_util.cu)
int gtb_inclusive_scan_u64(gtb_ctx *ctx, unsigned long long *d, int64_t n, dbuf<unsigned long long> &scratch);
// the same for `planes` arrays of n values spaced plane_stride apart, reading src and writing dst (may alias)
int gtb_inclusive_scan_planes_u64(gtb_ctx *ctx, const unsigned long long *src, unsigned long long *dst, int64_t n, int planes,
                                    int64_t plane_stride, dbuf<unsigned long long> &scratch);
