// gather_rate.cu -- what does one random-address LOAD per element from an L2-resident table cost on this GPU?
// (Companion of red_rate.cu: the question behind a single-pass design that looks a query's slot up in a cell table kept in L2
//  and counts in shared memory.)  N pseudo-random indices computed in registers, four independent loads per trip.
//   mode 0: ld.global.nc.u32           mode 1: ld.global.nc.v2.u32 (8 bytes)
//   mode 2: ld.global.nc.u32, then a shared-memory atomicAdd addressed by the loaded value (byte counters packed in words)
//   mode 3: as 2 with ld.global.nc.L1::no_allocate
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_rate gather_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
__device__ __forceinline__ uint32_t ldnc(const uint32_t *p) { uint32_t v; asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p)); return v; }
__device__ __forceinline__ uint32_t ldnc_na(const uint32_t *p) { uint32_t v; asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p)); return v; }
__device__ __forceinline__ uint2 ldnc64(const uint2 *p) { uint2 v; asm volatile("ld.global.nc.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p)); return v; }

template <int MODE>
__global__ void __launch_bounds__(512, 2) k(const uint32_t *tab, uint32_t mask, int trips, unsigned *sink, uint32_t nw) {
  extern __shared__ unsigned sm[];
  for (int i = threadIdx.x; i < (int)nw; i += blockDim.x) sm[i] = 0;
  __syncthreads();
  uint32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
  unsigned acc = 0;
  for (int t = 0; t < trips; t++) {
    uint32_t idx[4], v[4];
#pragma unroll
    for (int i = 0; i < 4; i++) { x = mix(x + (uint32_t)(t * 4 + i)); idx[i] = x & mask; }
#pragma unroll
    for (int i = 0; i < 4; i++) {
      if (MODE == 0 || MODE == 2) v[i] = ldnc(tab + idx[i]);
      if (MODE == 3) v[i] = ldnc_na(tab + idx[i]);
      if (MODE == 1) { const uint2 w = ldnc64(reinterpret_cast<const uint2 *>(tab) + (idx[i] >> 1)); v[i] = w.x + w.y; }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
      if (MODE >= 2) atomicAdd(&sm[(v[i] >> 2) % nw], 1u << (8 * (v[i] & 3u)));
      else acc += v[i];
    }
  }
  if (acc == 0xdeadbeef) *sink = acc;
  if (MODE >= 2 && threadIdx.x == 0) *sink = sm[5];
}

int main() {
  const int trips = 85;                                   // x 4 loads x 296 x 512 threads = 51.5 M ... scaled below
  const int grid = 296, block = 512;
  uint32_t *tab; unsigned *sink;
  const size_t max_entries = (size_t)1 << 26;
  cudaMalloc(&tab, max_entries * 4); cudaMalloc(&sink, 4);
  uint32_t *h = (uint32_t *)malloc(max_entries * 4);
  for (size_t i = 0; i < max_entries; i++) h[i] = (uint32_t)(i * 2654435761u >> 7);
  cudaMemcpy(tab, h, max_entries * 4, cudaMemcpyHostToDevice);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  size_t smem_bytes = 98304;
  auto run = [&](const char *name, auto kern, uint32_t bits) {
    const uint32_t mask = (1u << bits) - 1;
    const int tr = trips * 2;
    const double total = (double)grid * block * tr * 4;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    kern<<<grid, block, smem_bytes>>>(tab, mask, 8, sink, (uint32_t)(smem_bytes / 4));
    cudaEventRecord(a);
    for (int r = 0; r < 5; r++) kern<<<grid, block, smem_bytes>>>(tab, mask, tr, sink, (uint32_t)(smem_bytes / 4));
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 5;
    printf("%-34s table=%6.1f MB  %.3f ms for %.0f M  %.3e ops/s  %.2f SM-cycles/op(@1.965GHz)  100M would take %.3f ms\n", name,
           (double)(mask + 1) * 4 / 1e6, ms, total / 1e6, total / (ms * 1e-3), ms * 1e-3 * 1.965e9 * 148 / total, ms * 1e8 / total);
  };
  for (uint32_t bits : {16u, 18u, 20u, 22u, 24u, 26u}) {
    run("ld.nc.u32 random", k<0>, bits);
    run("ld.nc.v2.u32 random", k<1>, bits);
    run("ld.nc.u32 + smem byte atomic", k<2>, bits);
    run("ld.nc.na.u32 + smem byte atomic", k<3>, bits);
  }
  // the same gathers with more and more of the SM's 256 KB carved out as shared memory (less L1 to track misses in); one CTA per SM
  // from 128 KB up, so the grid halves and the per-thread trip count doubles
  for (size_t kb : {32, 64, 96, 112}) {
    smem_bytes = kb * 1024;
    printf("2 CTAs x %zu KB shared memory per SM: ", kb);
    run("ld.nc.na.u32 + smem byte atomic", k<3>, 21);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
