// Microbenchmark: TMEM -> register (tcgen05.ld 32x32b.x32) throughput per SM for 4/8/16 reading warps.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../moco_flow_b200/csrc/ptx.cuh"
using namespace mcf;
__global__ void __launch_bounds__(512, 1) k(int nwarps, int iters, unsigned long long* out, unsigned* sink) {
  __shared__ uint32_t tbase;
  int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&tbase, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  uint32_t base = tbase + ((uint32_t)((warp & 3) * 32) << 16);
  unsigned acc = 0;
  __syncthreads();
  long long t0 = clock64();
  if (warp < nwarps) {
    for (int i = 0; i < iters; ++i) {
      uint32_t v[32];
      tmem_ld32(base + ((i * 32 + (warp >> 2) * 128) & 511), v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) acc ^= v[j];
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
  if (acc == 0x12345678u) sink[0] = acc;
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 512);
}
int main() {
  unsigned long long* out; unsigned* sink;
  cudaMalloc(&out, 148 * 8); cudaMalloc(&sink, 4);
  for (int nw : {1, 4, 8, 16}) {
    int iters = 4096;
    k<<<148, 512>>>(nw, iters, out, sink);
    cudaDeviceSynchronize();
    unsigned long long h[148]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    double cyc = (double)h[0];
    double bytes = (double)nw * iters * 32 * 32 * 4;
    printf("warps=%2d  cycles=%.0f  bytes/clk/SM=%.1f  cycles per x32 load per warp=%.1f  err=%s\n", nw, cyc, bytes / cyc,
           cyc / iters, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
