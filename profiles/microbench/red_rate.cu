// red_rate.cu -- what does one random-address reduction per element cost on this GPU?
// Measures, for N pseudo-random table indices computed in registers (no input traffic):
//   red.global.add.u32 / .u64 into tables of several sizes, atomicAdd on shared memory,
//   and a plain shared-memory load (bitmap-style lookup) for reference.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o red_rate red_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

template <int MODE>
__global__ void __launch_bounds__(512) k(unsigned long long *t64, unsigned *t32, uint32_t mask, long long n_per_thread, unsigned *sink) {
  extern __shared__ unsigned sm[];
  for (int i = threadIdx.x; i < 32768; i += blockDim.x) sm[i] = i;
  __syncthreads();
  uint32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
  unsigned acc = 0;
  for (long long i = 0; i < n_per_thread; i++) {
    x = mix(x + (uint32_t)i);
    const uint32_t idx = x & mask;
    if (MODE == 0) asm volatile("red.global.add.u32 [%0], %1;" :: "l"(t32 + idx), "r"(1u) : "memory");
    if (MODE == 1) asm volatile("red.global.add.u64 [%0], %1;" :: "l"(t64 + idx), "l"(1ull) : "memory");
    if (MODE == 2) atomicAdd(&sm[idx & 32767], 1u);
    if (MODE == 3) acc += sm[idx & 32767];
    if (MODE == 4) acc += idx;
    if (MODE == 5) { const unsigned m = __match_any_sync(0xffffffffu, idx >> 5); acc += m; }
  }
  if (acc == 0xdeadbeef) *sink = acc;
  if (MODE == 2 && threadIdx.x == 0) *sink = sm[5];
}

int main() {
  const long long n_per_thread = 2048;
  const int grid = 148, block = 512;
  const double total = (double)grid * block * n_per_thread;
  unsigned long long *t64; unsigned *t32, *sink;
  cudaMalloc(&t64, (size_t)1 << 30); cudaMalloc(&t32, (size_t)1 << 29); cudaMalloc(&sink, 4);
  cudaMemset(t64, 0, (size_t)1 << 30); cudaMemset(t32, 0, (size_t)1 << 29);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  auto run = [&](const char *name, auto kern, uint32_t mask) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
    kern<<<grid, block, 131072>>>(t64, t32, mask, 64, sink);
    cudaEventRecord(a);
    kern<<<grid, block, 131072>>>(t64, t32, mask, n_per_thread, sink);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("%-28s mask=%08x  %.3f ms  %.3e ops/s  %.2f SM-cycles/op(@1.9GHz)\n", name, mask, ms, total / (ms * 1e-3),
           ms * 1e-3 * 1.9e9 * 148 / total);
  };
  for (uint32_t bits : {16u, 20u, 21u, 24u, 27u}) {
    run("red.u32 random", k<0>, (1u << bits) - 1);
    run("red.u64 random", k<1>, (1u << bits) - 1);
  }
  run("atomicAdd smem random", k<2>, 0xffffffffu);
  run("lds random", k<3>, 0xffffffffu);
  run("alu only", k<4>, 0xffffffffu);
  run("match_any", k<5>, 0xffffffffu);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
