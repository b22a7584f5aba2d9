// dsmem_rate.cu -- what does a random shared-memory atomic / load cost when the word lives in ANOTHER CTA of the cluster?
// Decides whether a counter array wider than one SM's shared memory can be spread over a thread-block cluster
// (window counts: a bucket of 2^18 16-bit counters over four CTAs; overlap counts: byte counters of a 1 M-region index over 16).
//   MODE 0  atom.shared::cta.add.u32 with result      (local, the reference point)
//   MODE 1  atom.shared::cluster.add.u32 with result  (uniformly random CTA of the cluster: (C - 1) / C of them remote)
//   MODE 2  red.shared::cluster.add.u32               (same addresses, no result)
//   MODE 3  ld.shared::cluster.u32                    (same addresses)
//   MODE 4  address arithmetic only
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dsmem_rate dsmem_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

constexpr int WORDS = 32768;                              // 128 KB of counters per CTA
constexpr int THREADS = 1024;

__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
  return r;
}

template <int MODE>
__global__ void __launch_bounds__(THREADS, 1) k(int csize_log2, int iters, unsigned *sink) {
  extern __shared__ uint32_t sm[];
  for (int i = threadIdx.x; i < WORDS; i += THREADS) sm[i] = 0;
  cg::cluster_group cluster = cg::this_cluster();
  cluster.sync();
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm);
  const uint32_t cmask = (1u << csize_log2) - 1u;
  uint32_t x = (blockIdx.x * THREADS + threadIdx.x) * 2654435761u + 12345u;
  uint32_t acc = 0;
  for (int it = 0; it < iters; it++) {
    uint32_t a[4], o[4] = {0, 0, 0, 0};
#pragma unroll
    for (int j = 0; j < 4; j++) {
      x = x * 1664525u + 1013904223u;
      const uint32_t word = (x >> 8) & (WORDS - 1), rank = (x >> 28) & cmask;
      a[j] = MODE == 0 ? base + word * 4u : mapa(base + word * 4u, rank);
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
      if (MODE == 0) asm volatile("atom.shared::cta.add.u32 %0, [%1], %2;" : "=r"(o[j]) : "r"(a[j]), "r"(1u) : "memory");
      if (MODE == 1) asm volatile("atom.shared::cluster.add.u32 %0, [%1], %2;" : "=r"(o[j]) : "r"(a[j]), "r"(1u) : "memory");
      if (MODE == 2) asm volatile("red.shared::cluster.add.u32 [%0], %1;" :: "r"(a[j]), "r"(1u) : "memory");
      if (MODE == 3) asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(o[j]) : "r"(a[j]) : "memory");
      if (MODE == 4) o[j] = a[j];
    }
    acc += o[0] + o[1] + o[2] + o[3];
  }
  cluster.sync();
  if (acc == 0xdeadbeefu) *sink = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *sink = sm[5];
}

template <int MODE>
void run(const char *name, int csize_log2, unsigned *sink) {
  const int iters = 2048;
  const int csize = 1 << csize_log2;
  const int grid = (148 / csize) * csize;                 // whole clusters only
  auto kern = k<MODE>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, WORDS * 4);
  if (csize > 8) cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = WORDS * 4;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int max_clusters = 0;
  cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaLaunchKernelEx(&cfg, kern, csize_log2, 16, sink);
  cudaEventRecord(a);
  cudaLaunchKernelEx(&cfg, kern, csize_log2, iters, sink);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms = 0; cudaEventElapsedTime(&ms, a, b);
  const double ops = (double)grid * THREADS * iters * 4;
  const int resident = max_clusters * csize < grid ? max_clusters * csize : grid;      // CTAs that run at once
  printf("%-34s cluster %2d  grid %3d (resident %3d)  %.3f ms  %.3e ops/s  %.3f SM-cycles/op per busy SM (@1.965 GHz)  %s\n", name, csize, grid, resident, ms,
         ops / (ms * 1e-3), ms * 1e-3 * 1.965e9 * resident / ops, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  unsigned *sink; cudaMalloc(&sink, 4);
  for (int cl = 0; cl <= 4; cl++) {
    run<0>("atom.shared::cta (local)", cl, sink);
    run<1>("atom.shared::cluster (random CTA)", cl, sink);
    run<2>("red.shared::cluster (random CTA)", cl, sink);
    run<3>("ld.shared::cluster (random CTA)", cl, sink);
    run<4>("address arithmetic only", cl, sink);
  }
  return 0;
}
