// Microbenchmarks that bound the chain kernel's design (run on a B200):
//  (1) tcgen05.mma issue/execute rate per SM for M=128, N in {128, 256}, K=16 (bf16 -> fp32), cta_group::1
//  (2) how far the issuing thread runs ahead of the tensor core (issue time vs completion time)
//  (3) 1-D bulk copy (UBLKCP) L2 -> shared memory: latency of one 16 KB copy, and per-SM throughput with
//      1/2/4/8 copies in flight while all SMs stream the same 1.2 MB weight set
#include <cstdio>
#include <cuda_runtime.h>
#include "../../moco_flow_b200/csrc/ptx.cuh"
using namespace mcf;

__global__ void __launch_bounds__(128, 1) k_mma(int n_mma, int ncols, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tbase;
  __shared__ uint64_t bar;
  int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&tbase, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (threadIdx.x == 32) {
    const uint32_t a = smem_u32(smem), b = smem_u32(smem) + 16384;
    const uint32_t idesc = make_idesc((uint32_t)ncols);
    long long t0 = clock64();
    for (int i = 0; i < n_mma; ++i) {
      const uint32_t k = i & 3;
      umma_bf16(tbase + ((i >> 2) & 1) * 256, make_sdesc(a + k * 32, 0, 1024), make_sdesc(b + k * 32, 0, 1024), idesc,
                (i > 7) ? 1u : 0u);
    }
    umma_commit(&bar);
    long long t1 = clock64();
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    out[blockIdx.x * 2 + 0] = (unsigned long long)(t1 - t0);
    out[blockIdx.x * 2 + 1] = (unsigned long long)(t2 - t0);
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 512);
}

// depth copies of `bytes` in flight, each SM walks the whole buffer `iters` times
__global__ void __launch_bounds__(128, 1) k_copy(const uint8_t* src, int total_bytes, int bytes, int depth, int iters,
                                                 unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[8];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int n = total_bytes / bytes * iters;
    uint32_t phase[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long t0 = clock64();
    long long lat = 0;
    for (int i = 0; i < n + depth; ++i) {
      const int s = i % depth;
      if (i >= depth) { mbar_wait(&bar[s], phase[s]); phase[s] ^= 1u; }
      if (i == depth) lat = clock64() - t0;
      if (i < n) {
        mbar_arrive_expect_tx(&bar[s], bytes);
        bulk_g2s(smem + s * bytes, src + (size_t)((i * (long long)bytes) % total_bytes), bytes, &bar[s]);
      }
    }
    long long t1 = clock64();
    out[blockIdx.x * 2 + 0] = (unsigned long long)(t1 - t0);
    out[blockIdx.x * 2 + 1] = (unsigned long long)lat;
  }
}


// `nthr` issuing threads (one per warp), each keeping `depth` copies of `bytes` in flight into its own smem region
__device__ __forceinline__ void bulk_g2s_cta(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(256, 1) k_copy_mt(const uint8_t* src, int total_bytes, int bytes, int depth, int iters,
                                                    int nthr, unsigned long long* out, int cta_form = 0) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[8][8];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 64; ++i) mbar_init(&bar[i / 8][i % 8], 1);
    fence_mbar_init();
  }
  __syncthreads();
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0 && w < nthr) {
    const int n = total_bytes / bytes * iters / nthr;
    uint32_t phase[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long t0 = clock64();
    long long t_issue = 0;
    for (int i = 0; i < n + depth; ++i) {
      const int s = i % depth;
      if (i >= depth) { mbar_wait(&bar[w][s], phase[s]); phase[s] ^= 1u; }
      if (i < n) {
        mbar_arrive_expect_tx(&bar[w][s], bytes);
        long long a = clock64();
        if (cta_form)
          bulk_g2s_cta(smem + (size_t)(w * depth + s) * bytes, src + (size_t)(((long long)(i * nthr + w) * bytes) % total_bytes),
                       bytes, &bar[w][s]);
        else
          bulk_g2s(smem + (size_t)(w * depth + s) * bytes, src + (size_t)(((long long)(i * nthr + w) * bytes) % total_bytes), bytes,
                   &bar[w][s]);
        t_issue += clock64() - a;
      }
    }
    long long t1 = clock64();
    if (w == 0) {
      out[blockIdx.x * 2 + 0] = (unsigned long long)(t1 - t0);
      out[blockIdx.x * 2 + 1] = (unsigned long long)(t_issue / n);
    }
  }
}

// n_mma MMAs (N=256) with a tcgen05.commit to a rotating mbarrier after every `group` of them: does a commit
// put a bubble into the tensor pipe?
__global__ void __launch_bounds__(128, 1) k_commit(int n_mma, int group, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tbase;
  __shared__ uint64_t bar[8], done;
  int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1); mbar_init(&done, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&tbase, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (threadIdx.x == 32) {
    const uint32_t a = smem_u32(smem), b = smem_u32(smem) + 16384;
    const uint32_t idesc = make_idesc(256u);
    long long t0 = clock64();
    int nb = 0;
    for (int i = 0; i < n_mma; ++i) {
      const uint32_t k = i & 3;
      umma_bf16(tbase + ((i >> 2) & 1) * 256, make_sdesc(a + k * 32, 0, 1024), make_sdesc(b + k * 32, 0, 1024), idesc,
                (i > 7) ? 1u : 0u);
      if ((i + 1) % group == 0) { umma_commit(&bar[nb & 7]); ++nb; }
    }
    umma_commit(&done);
    mbar_wait(&done, 0);
    long long t2 = clock64();
    out[blockIdx.x * 2 + 1] = (unsigned long long)(t2 - t0);
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 512);
}

// raw cta_group::2 rate: the leader of a CTA pair issues n M=256 N=256 K=16 MMAs on resident operands
__global__ void __launch_bounds__(128, 1) k_mma_pair(int n_mma, unsigned long long* out, int commit_every = 0, int mask = 3) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tbase;
  __shared__ uint64_t bar;
  __shared__ uint64_t rot[8];
  int warp = threadIdx.x >> 5;
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0) for (int i = 0; i < 8; ++i) mbar_init(&rot[i], 1);
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc_pair(&tbase, 512); tmem_relinquish_pair(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); cluster_sync_all(); tc_fence_after();
  if (threadIdx.x == 32 && rank == 0) {
    const uint32_t a = smem_u32(smem), b = smem_u32(smem) + 16384;
    const uint32_t idesc = make_idesc(256u, false, false, 256u);
    const uint64_t ad = make_sdesc(a, 0, 1024), bd = make_sdesc(b, 0, 1024);
    long long t0 = clock64();
    for (int i = 0; i < n_mma; ++i) {
      const uint32_t k = i & 3;
      umma_bf16_pair(tbase + ((i >> 2) & 1) * 256, ad + 2 * k, bd + 2 * k, idesc, (i > 7) ? 1u : 0u);
      if (commit_every && k == 3 && ((i >> 2) % commit_every) == commit_every - 1) umma_commit_pair(&rot[(i >> 2) & 7], (uint16_t)mask);
    }
    umma_commit_pair(&bar, 1);
    long long t1 = clock64();
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    out[blockIdx.x * 2 + 0] = (unsigned long long)(t1 - t0);
    out[blockIdx.x * 2 + 1] = (unsigned long long)(t2 - t0);
  }
  tc_fence_before(); __syncthreads(); cluster_sync_all();
  if (warp == 0) tmem_dealloc_pair(tbase, 512);
}

int main() {
  unsigned long long* out;
  cudaMalloc(&out, 148 * 16);
  unsigned long long h[296];
  cudaFuncSetAttribute(k_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaFuncSetAttribute(k_copy, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  for (int ncols : {128, 256}) {
    for (int n : {4, 16, 64, 1024}) {
      for (int grid : {1, 148}) {
        k_mma<<<grid, 128, 64 * 1024>>>(n, ncols, out);
        cudaDeviceSynchronize();
        cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        printf("mma N=%3d n=%4d grid=%3d: issue %6llu clk (%.1f/mma), complete %6llu clk (%.1f/mma)  %s\n", ncols, n, grid,
               h[0], (double)h[0] / n, h[1], (double)h[1] / n, cudaGetErrorString(cudaGetLastError()));
      }
    }
  }
  cudaFuncSetAttribute(k_mma_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int n : {16, 256, 4096}) {
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(148); lc.blockDim = dim3(128); lc.dynamicSmemBytes = 64 * 1024; lc.stream = 0;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    lc.attrs = at; lc.numAttrs = 1;
    cudaLaunchKernelEx(&lc, k_mma_pair, n, out, 0, 3);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("pair mma M=256 N=256 n=%4d: issue %.1f clk/mma, complete %.1f clk/mma (ideal 128)  %s\n", n, (double)h[0] / n,
           (double)h[1] / n, cudaGetErrorString(e));
    if (n == 4096) {
      for (int mask : {1, 3}) for (int every : {1, 2}) {
        cudaLaunchKernelEx(&lc, k_mma_pair, n, out, every, mask);
        e = cudaDeviceSynchronize();
        cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        printf("pair mma + commit(mask %d) every %d x 4 MMAs: complete %.1f clk/mma (ideal 128)  %s\n", mask, every,
               (double)h[1] / n, cudaGetErrorString(e));
      }
    }
  }
  cudaFuncSetAttribute(k_commit, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int group : {1, 2, 4, 8, 16, 1024}) {
    k_commit<<<148, 128, 64 * 1024>>>(1024, group, out);
    cudaDeviceSynchronize();
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("commit every %4d MMAs (N=256): %.1f clk/mma (ideal 128)  %s\n", group, (double)h[1] / 1024, cudaGetErrorString(cudaGetLastError()));
  }
  uint8_t* src;
  const int total = 1216 * 1024;
  cudaMalloc(&src, total);
  cudaMemset(src, 1, total);
  for (int grid : {1, 148}) {
    for (int bytes : {8192, 16384}) {
      for (int depth : {1, 2, 4, 8}) {
        k_copy<<<grid, 128, 160 * 1024>>>(src, total, bytes, depth, 4, out);
        cudaDeviceSynchronize();
        cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        double n = (double)(total / bytes * 4);
        printf("copy grid=%3d bytes=%5d depth=%d: %.1f B/clk/SM, first-batch latency %llu clk  %s\n", grid, bytes, depth,
               n * bytes / (double)h[0], h[1], cudaGetErrorString(cudaGetLastError()));
      }
    }
  }
  cudaFuncSetAttribute(k_copy_mt, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int form : {0, 1}) for (int nthr : {1, 4}) {
    k_copy_mt<<<148, 256, 200 * 1024>>>(src, total, 16384, 2, 8, nthr, out, form);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    double n = (double)(total / 16384 * 8 / nthr) * nthr;
    printf("copy form=%s threads=%d: %.1f B/clk/SM, issue instr %llu clk  %s\n", form ? "shared::cta" : "shared::cluster", nthr,
           n * 16384 / (double)h[0], h[1], cudaGetErrorString(e));
  }
  for (int grid : {1, 148}) {
    for (int bytes : {4096, 16384, 32768, 65536}) {
      for (int nthr : {1, 2, 4}) {
        for (int depth : {1, 2}) {
          if ((long long)bytes * nthr * depth > 196608) continue;
          k_copy_mt<<<grid, 256, 200 * 1024>>>(src, total, bytes, depth, 8, nthr, out);
          cudaDeviceSynchronize();
          cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
          double n = (double)(total / bytes * 8 / nthr) * nthr;
          printf("copy_mt grid=%3d bytes=%5d threads=%d depth=%d: %.1f B/clk/SM, issue instr %llu clk  %s\n", grid, bytes, nthr,
                 depth, n * bytes / (double)h[0], h[1], cudaGetErrorString(cudaGetLastError()));
        }
      }
    }
  }
  return 0;
}
