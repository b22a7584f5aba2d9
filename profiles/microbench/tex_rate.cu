// tex_rate.cu -- does the TEXTURE path look random cell-table entries up faster than the LSU path?
// gather_rate.cu says a fully divergent ld.global costs one L1TEX tag cycle per lane (1.05-1.18 SM-cycles per element whatever
// the table size up to 16 MB).  Same indices here, fetched with tex1Dfetch from a linear texture object (u32 / uint2 / uint4
// texels), and with plain loads again for reference.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tex_rate tex_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

template <int MODE>
__global__ void __launch_bounds__(512, 2) k(cudaTextureObject_t tex, const uint32_t *tab, uint32_t mask, int trips, unsigned *sink) {
  uint32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
  unsigned acc = 0;
  for (int t = 0; t < trips; t++) {
    uint32_t idx[4], v[4];
#pragma unroll
    for (int i = 0; i < 4; i++) { x = mix(x + (uint32_t)(t * 4 + i)); idx[i] = x & mask; }
#pragma unroll
    for (int i = 0; i < 4; i++) {
      if (MODE == 0) v[i] = __ldg(tab + idx[i]);
      if (MODE == 1) v[i] = tex1Dfetch<unsigned>(tex, (int)idx[i]);
      if (MODE == 2) { const uint2 w = tex1Dfetch<uint2>(tex, (int)(idx[i] >> 1)); v[i] = w.x + w.y; }
      if (MODE == 3) { const uint4 w = tex1Dfetch<uint4>(tex, (int)(idx[i] >> 2)); v[i] = w.x + w.y + w.z + w.w; }
      if (MODE == 4) v[i] = idx[i];
    }
#pragma unroll
    for (int i = 0; i < 4; i++) acc += v[i];
  }
  if (acc == 0xdeadbeef) *sink = acc;
}

int main() {
  const int grid = 296, block = 512, trips = 170;
  uint32_t *tab; unsigned *sink;
  const size_t max_entries = (size_t)1 << 24;                       // 64 MB of u32
  cudaMalloc(&tab, max_entries * 4); cudaMalloc(&sink, 4);
  cudaMemset(tab, 1, max_entries * 4);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  auto mktex = [&](cudaChannelFormatDesc d, size_t bytes) {
    cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = tab; rd.res.linear.desc = d; rd.res.linear.sizeInBytes = bytes;
    cudaTextureDesc td = {}; td.readMode = cudaReadModeElementType;
    cudaTextureObject_t t = 0; cudaCreateTextureObject(&t, &rd, &td, nullptr); return t;
  };
  for (uint32_t bits : {16u, 20u, 22u}) {
    const uint32_t mask = (1u << bits) - 1;
    const size_t bytes = ((size_t)mask + 1) * 4;
    cudaTextureObject_t t1 = mktex(cudaCreateChannelDesc<unsigned>(), bytes), t2 = mktex(cudaCreateChannelDesc<uint2>(), bytes), t4 = mktex(cudaCreateChannelDesc<uint4>(), bytes);
    auto run = [&](const char *name, auto kern, cudaTextureObject_t t) {
      const double total = (double)grid * block * trips * 4;
      kern<<<grid, block>>>(t, tab, mask, 8, sink);
      cudaEventRecord(a);
      for (int r = 0; r < 5; r++) kern<<<grid, block>>>(t, tab, mask, trips, sink);
      cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b); ms /= 5;
      printf("%-28s table=%6.1f MB  %.3f ms for %.0f M  %.2f SM-cycles/op(@1.965GHz)  %s\n", name, bytes / 1e6, ms, total / 1e6,
             ms * 1e-3 * 1.965e9 * 148 / total, cudaGetErrorString(cudaGetLastError()));
    };
    run("ld.global.nc u32", k<0>, t1);
    run("tex1Dfetch u32", k<1>, t1);
    run("tex1Dfetch uint2", k<2>, t2);
    run("tex1Dfetch uint4", k<3>, t4);
    run("index arithmetic only", k<4>, t1);
  }
  return 0;
}
