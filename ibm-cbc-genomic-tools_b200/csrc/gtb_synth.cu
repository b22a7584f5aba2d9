// gtb_synth.cu -- counter-based synthetic read generator on the device (bench / tests utility).
// Same stream as tests/support.py:synth_reads:
//   a = splitmix64(seed*0x9E3779B97F4A7C15 + i); b = splitmix64(a)
//   p = a mod sum_c(len_c - read_len + 1); chromosome = interval containing p; start = offset + 1
//   stop = start + read_len - 1; strand = (b & 1) ? '-' : '+'
// The ranged form draws p from [p_lo, p_hi) instead (p = p_lo + a mod (p_hi - p_lo)): reads of one genome shard.
#include "gtb_internal.cuh"

namespace {
constexpr int SYNTH_MAX_CHROM = 1024;
__constant__ unsigned long long c_cum[SYNTH_MAX_CHROM + 1];

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

__global__ void __launch_bounds__(256) synth_reads_kernel(unsigned long long seed, long long first, long long n, int read_len,
                                                          int n_chrom, unsigned long long p_lo, unsigned long long p_span,
                                                          int32_t *chrom, int32_t *start, int32_t *stop, int8_t *strand) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
    const unsigned long long a = splitmix64(seed * 0x9E3779B97F4A7C15ull + (unsigned long long)(first + k));
    const unsigned long long b = splitmix64(a);
    const unsigned long long p = p_lo + a % p_span;
    int lo = 0, hi = n_chrom;                 // last c with cum[c] <= p
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (c_cum[mid] <= p) lo = mid; else hi = mid; }
    const int32_t s = (int32_t)(p - c_cum[lo]) + 1;
    chrom[k] = lo; start[k] = s; stop[k] = s + read_len - 1; strand[k] = (b & 1ull) ? '-' : '+';
  }
}
}  // namespace

extern "C" int gtb_synth_reads_range(gtb_ctx *ctx, uint64_t seed, int64_t first, int64_t n, int32_t read_len, int32_t n_chrom,
                                     const int64_t *chrom_len, uint64_t p_lo, uint64_t p_hi, int32_t *d_chrom, int32_t *d_start,
                                     int32_t *d_stop, int8_t *d_strand) {
  if (!ctx || !chrom_len || n < 0 || n_chrom <= 0 || n_chrom > SYNTH_MAX_CHROM || read_len <= 0) return GTB_ERR_ARG;
  if (n == 0) return GTB_OK;
  GTB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  unsigned long long cum[SYNTH_MAX_CHROM + 1];
  cum[0] = 0;
  for (int c = 0; c < n_chrom; c++) {
    long long eff = chrom_len[c] - read_len + 1;
    cum[c + 1] = cum[c] + (unsigned long long)(eff > 0 ? eff : 0);
  }
  if (cum[n_chrom] == 0) return gtb_fail(ctx, GTB_ERR_ARG, "no chromosome is long enough for the read length");
  if (p_hi > cum[n_chrom]) p_hi = cum[n_chrom];
  if (p_lo >= p_hi) return gtb_fail(ctx, GTB_ERR_ARG, "empty position range");
  GTB_CUDA_OK(ctx, cudaMemcpyToSymbolAsync(c_cum, cum, sizeof(unsigned long long) * (n_chrom + 1), 0, cudaMemcpyHostToDevice, ctx->stream));
  const unsigned grid = gtb_grid_for(n, 256, (int64_t)ctx->sm_count * 16);
  GTB_LAUNCH(ctx, "synth_reads", synth_reads_kernel, grid, 256, 0, (unsigned long long)seed, (long long)first, (long long)n,
             (int)read_len, (int)n_chrom, (unsigned long long)p_lo, (unsigned long long)(p_hi - p_lo), d_chrom, d_start, d_stop, d_strand);
  GTB_TRY(gtb_check_launch(ctx));
  GTB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));   // cum[] lives on this stack frame
  return GTB_OK;
}

extern "C" int gtb_synth_reads(gtb_ctx *ctx, uint64_t seed, int64_t first, int64_t n, int32_t read_len, int32_t n_chrom,
                               const int64_t *chrom_len, int32_t *d_chrom, int32_t *d_start, int32_t *d_stop, int8_t *d_strand) {
  return gtb_synth_reads_range(ctx, seed, first, n, read_len, n_chrom, chrom_len, 0, ~0ull, d_chrom, d_start, d_stop, d_strand);
}
