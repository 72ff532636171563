// Microbenchmark: the chain kernel's weight ring in isolation (producer warps -> 16 KB stages -> tcgen05.mma), no
// epilogue.  What is the best tensor-pipe utilisation this loop structure can reach, and what does each piece cost?
#include <cstdio>
#include <cuda_runtime.h>
#include "../../moco_flow_b200/csrc/ptx.cuh"
using namespace mcf;

struct Cfg { int stages, fuse, fence, producers, chunks, timing; };

__global__ void __launch_bounds__(128, 1) k_ring(const uint8_t* src, int total_bytes, Cfg cfg, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tbase;
  __shared__ uint64_t w_full[8], w_empty[8], done;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* abuf = smem;                 // 64 KB activation operand (4 K-blocks)
  uint8_t* ring = smem + 65536;
  for (int i = threadIdx.x; i < 65536 / 4; i += 128) reinterpret_cast<uint32_t*>(abuf)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    mbar_init(&done, 1);
    fence_mbar_init();
  }
  if (warp == 2) { tmem_alloc(&tbase, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const int S = cfg.stages;
  if (warp != 1) {
    if (lane == 0) {
      const int me = warp == 0 ? 0 : warp - 1;
      if (me < cfg.producers) {
        uint32_t stage = 0, phase = 0; int turn = 0;
        for (int c = 0; c < cfg.chunks; ++c) {
          if (turn == me) {
            mbar_wait(&w_empty[stage], phase ^ 1u);
            mbar_arrive_expect_tx(&w_full[stage], 16384);
            bulk_g2s(ring + stage * 16384, src + ((size_t)c * 16384) % total_bytes, 16384, &w_full[stage]);
          }
          if (++turn == cfg.producers) turn = 0;
          if (++stage == S) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (lane == 0) {
    uint32_t stage = 0, phase = 0;
    long long t_wait = 0, t_issue = 0;
    long long t0 = clock64(), t = t0;
    const uint32_t a_addr = smem_u32(abuf), r_addr = smem_u32(ring);
    for (int c = 0; c < cfg.chunks; ++c) {
      mbar_wait(&w_full[stage], phase);
      const bool fuse = cfg.fuse != 0;
      if (fuse) mbar_wait(&w_full[stage + 1], phase);
      if (cfg.fence) tc_fence_after();
      if (cfg.timing) { long long n = clock64(); t_wait += n - t; t = n; }
      const uint32_t idesc = make_idesc(fuse ? 256u : 128u);
      const uint32_t a_base = a_addr + (c & 3) * 16384, b_base = r_addr + stage * 16384;
      const uint32_t d = tbase + ((c >> 3) & 1) * 256;
      for (uint32_t k = 0; k < 4; ++k)
        umma_bf16(d, make_sdesc(a_base + k * 32, 0, 1024), make_sdesc(b_base + k * 32, 0, 1024), idesc, (c | k) ? 1u : 0u);
      umma_commit(&w_empty[stage]);
      if (++stage == S) { stage = 0; phase ^= 1u; }
      if (fuse) {
        umma_commit(&w_empty[stage]);
        if (++stage == S) { stage = 0; phase ^= 1u; }
        ++c;
      }
      if (cfg.timing) { long long n = clock64(); t_issue += n - t; t = n; }
    }
    umma_commit(&done);
    mbar_wait(&done, 0);
    long long t1 = clock64();
    out[blockIdx.x * 4 + 0] = (unsigned long long)(t1 - t0);
    out[blockIdx.x * 4 + 1] = (unsigned long long)t_wait;
    out[blockIdx.x * 4 + 2] = (unsigned long long)t_issue;
  }
  __syncthreads();
  if (warp == 2) tmem_dealloc(tbase, 512);
}

// CTA-pair variant: cluster of 2, tcgen05 cta_group::2, M=256 N=256; each CTA streams its half of B (16 KB per K=64
// block) into its own ring; the peer relays "landed" to the leader, the leader's commits free both rings.
__global__ void __launch_bounds__(128, 1) k_ring_pair(const uint8_t* src, int total_bytes, Cfg cfg, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tbase;
  __shared__ uint64_t w_full[8], w_empty[8], w_peer[8], done;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  uint8_t* abuf = smem;
  uint8_t* ring = smem + 65536;
  for (int i = threadIdx.x; i < 65536 / 4; i += 128) reinterpret_cast<uint32_t*>(abuf)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); mbar_init(&w_peer[i], 1); }
    mbar_init(&done, 1);
    fence_mbar_init();
  }
  if (warp == 2) { tmem_alloc_pair(&tbase, 512); tmem_relinquish_pair(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); cluster_sync_all(); tc_fence_after();
  const int S = cfg.stages;
  const int n_st = cfg.chunks / 2;   // stages per CTA (each covers two of the single-CTA kernel's chunks)
  if (rank != 0 && cfg.fence == 4 && warp >= 1) {
    // relay mode 4: three relay threads (lane 0 of warps 1..3 of the peer), stage sequence split round-robin
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int c = 0; c < n_st; ++c) {
        if (c % 3 == warp - 1) {
          mbar_wait(&w_full[stage], phase);
          mbar_arrive_cluster_relaxed(map_to_cta(smem_u32(&w_peer[stage]), 0));
        }
        if (++stage == S) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp != 1) {
    if (lane == 0) {
      const int me = warp == 0 ? 0 : warp - 1;
      if (me < cfg.producers) {
        uint32_t stage = 0, phase = 0; int turn = 0;
        for (int c = 0; c < n_st; ++c) {
          if (turn == me) {
            if (cfg.timing == 2) mbar_wait(&w_empty[stage], phase ^ 1u); else mbar_wait_cluster(&w_empty[stage], phase ^ 1u);
            const uint8_t* g = src + ((size_t)(2 * c + rank) * 16384) % total_bytes;
            if (cfg.fence == 3) {
              // direct signalling: both halves complete_tx on the LEADER's barrier
              if (rank == 0) mbar_arrive_expect_tx(&w_full[stage], 32768);
              const uint32_t bar_addr = map_to_cta(smem_u32(&w_full[stage]), 0);
              const uint32_t dst_addr = map_to_cta(smem_u32(ring + stage * 16384), rank);
              asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst_addr),
                           "l"(g), "r"(16384u), "r"(bar_addr) : "memory");
            } else {
              mbar_arrive_expect_tx(&w_full[stage], 16384);
              bulk_g2s(ring + stage * 16384, g, 16384, &w_full[stage]);
            }
          }
          if (++turn == cfg.producers) turn = 0;
          if (++stage == S) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (rank != 0 && cfg.fence == 3) {
    // nothing to relay
  } else if (rank != 0) {
    // relay modes (cfg.fence): 0 = one thread, release.cluster arrive; 1 = one thread, relaxed arrive;
    // 2 = one lane per stage, release arrive
    if (cfg.fence == 2) {
      if (lane < S) {
        const uint32_t remote = map_to_cta(smem_u32(&w_peer[lane]), 0);
        uint32_t phase = 0;
        for (int c = lane; c < n_st; c += S) {
          mbar_wait(&w_full[lane], phase);
          mbar_arrive_cluster(remote);
          phase ^= 1u;
        }
      }
    } else if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int c = 0; c < n_st; ++c) {
        mbar_wait(&w_full[stage], phase);
        const uint32_t remote = map_to_cta(smem_u32(&w_peer[stage]), 0);
        if (cfg.fence == 1) asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
        else mbar_arrive_cluster(remote);
        if (++stage == S) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (lane == 0) {
    uint32_t stage = 0, phase = 0;
    long long t_wait = 0, t_issue = 0;
    long long t0 = clock64(), t = t0;
    const uint32_t a_addr = smem_u32(abuf), r_addr = smem_u32(ring);
    const uint32_t idesc = make_idesc(256u, false, false, 256u);
    for (int c = 0; c < n_st; ++c) {
      mbar_wait(&w_full[stage], phase);
      if (cfg.fence != 3) { if (cfg.timing == 2) mbar_wait(&w_peer[stage], phase); else mbar_wait_cluster(&w_peer[stage], phase); }
      tc_fence_after();
      if (cfg.timing) { long long n = clock64(); t_wait += n - t; t = n; }
      const uint32_t a_base = a_addr + (c & 3) * 16384, b_base = r_addr + stage * 16384;
      const uint32_t d = tbase + ((c >> 2) & 1) * 256;
      for (uint32_t k = 0; k < 4; ++k)
        umma_bf16_pair(d, make_sdesc(a_base + k * 32, 0, 1024), make_sdesc(b_base + k * 32, 0, 1024), idesc, (c | k) ? 1u : 0u);
      umma_commit_pair(&w_empty[stage], 3);
      if (++stage == S) { stage = 0; phase ^= 1u; }
      if (cfg.timing) { long long n = clock64(); t_issue += n - t; t = n; }
    }
    umma_commit_pair(&done, 1);
    mbar_wait(&done, 0);
    long long t1 = clock64();
    out[blockIdx.x * 4 + 0] = (unsigned long long)(t1 - t0);
    out[blockIdx.x * 4 + 1] = (unsigned long long)t_wait;
    out[blockIdx.x * 4 + 2] = (unsigned long long)t_issue;
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc_pair(tbase, 512);
}

int main() {
  unsigned long long* out;
  cudaMalloc(&out, 148 * 32);
  unsigned long long h[592];
  uint8_t* src;
  const int total = 1216 * 1024;
  cudaMalloc(&src, total);
  cudaMemset(src, 0x3c, total);
  cudaFuncSetAttribute(k_ring, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int chunks = 4096;
  for (int stages : {4, 8}) for (int fuse : {0, 1}) for (int fence : {1, 0}) for (int prod : {1, 3}) for (int timing : {1, 0}) {
    Cfg cfg{stages, fuse, fence, prod, chunks, timing};
    k_ring<<<148, 128, 200 * 1024>>>(src, total, cfg, out);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("stages=%d fuse=%d fence=%d producers=%d timing=%d: %.1f clk/chunk (ideal 256), wait %.1f issue %.1f  %s\n", stages, fuse,
           fence, prod, timing, (double)h[0] / chunks, (double)h[1] / chunks, (double)h[2] / chunks, cudaGetErrorString(e));
  }
  cudaFuncSetAttribute(k_ring_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int stages : {4, 8}) for (int prod : {1}) for (int relay : {1, 4}) for (int timing : {2}) {  // relay 3 (direct remote signalling) faults
    Cfg cfg{stages, 1, relay, prod, chunks, timing};
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(148); lc.blockDim = dim3(128); lc.dynamicSmemBytes = 200 * 1024; lc.stream = 0;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    lc.attrs = at; lc.numAttrs = 1;
    cudaLaunchKernelEx(&lc, k_ring_pair, (const uint8_t*)src, total, cfg, out);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    // same FLOPs per SM as the single-CTA runs: `chunks` K=64 x N=128 blocks of tensor work per CTA -> ideal 256 clk/chunk
    printf("PAIR stages=%d producers=%d relay=%d waits=%s: %.1f clk/chunk-equivalent (ideal 256), wait %.1f issue %.1f  %s\n", stages,
           prod, relay, timing == 2 ? "cta" : "cluster", (double)h[0] / chunks, (double)h[1] / chunks, (double)h[2] / chunks, cudaGetErrorString(e));
  }
  return 0;
}
