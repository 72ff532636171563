// Weight-gradient GEMM of the MLP backward on tcgen05:  C[i][j] += sum_rows P[row][i] * Q[row][j].
//
// P (= dY of a layer) and Q (= the layer's input) are the bf16 128B-swizzled tile images the chain
// kernels saved ([128 rows][64 cols] blocks).  The contraction runs over rows, so both operands are
// consumed MN-major straight from those images -- no transpose pass.  A CTA owns one 128-wide i-block
// and all j (<=256) and a slice of the row tiles (split-K); accumulators stay in TMEM for the whole
// slice and are reduced into C with vector fp32 atomics.  Optionally an extra N=64 MMA against a
// ones-tile produces colsum_p[i] = sum_rows P[row][i] (the bias gradient).
// Reference semantics: autograd of nn.Linear (models/nerf.py:30-58, models/nof.py:42-53).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/moco_flow_b200.h"
#include "ptx.cuh"

namespace mcf {

constexpr int kDwThreads = 256;
constexpr int kDwStages = 4;
// bulk-copy issuing threads (lane 0 of warps 0, 2, 3).  Three were measured against one: no difference, the kernel
// is not issue-bound.
constexpr int kDwProducers = 1;
constexpr uint32_t kBlkD = MCF_BLOCK_BYTES;
constexpr uint32_t kHalf = kBlkD / 2;                    // 64 rows of one 64-column block
// one stage = 64 rows: P i-block (2 blocks) + Q (<= 4 blocks), 8 KB each -> 48 KB; 4 stages keep ~150 KB of
// HBM reads in flight per SM, enough to cover the DRAM latency at full bandwidth
constexpr uint32_t kStageBytes = 2 * kHalf + 4 * kHalf;

struct DwCtl {
  uint64_t full[kDwStages];
  uint64_t empty[kDwStages];
  uint64_t done;
  uint32_t tmem_base;
  uint32_t pad;
};
constexpr uint32_t kDwOffOnes = kDwStages * kStageBytes;
constexpr uint32_t kDwOffCtl = kDwOffOnes + kBlkD;
constexpr uint32_t kDwSmem = kDwOffCtl + sizeof(DwCtl);

struct DwBatchArgs {
  const mcf_dw_job_t* jobs;
  const uint8_t* base[3];   // forward save record, backward save record, per-ray feature images
  long long tile_bytes[3];
  float* staging;
  long long n_tiles;
};

// second source of Q blocks (blocks >= split come from base + tile*tile_bytes + off)
struct DwQ2 {
  const uint8_t* base;
  long long tile_bytes;
  uint32_t off;
  int split;
};

__device__ __forceinline__ void dw_body(const mcf_dw_params_t& p, const DwQ2& q2, int nib, int cta, int n_cta);

__global__ void __launch_bounds__(kDwThreads, 1) k_dw(const __grid_constant__ mcf_dw_params_t p, int nib) {
  DwQ2 q2 = {nullptr, 0, 0u, 1 << 20};
  dw_body(p, q2, nib, blockIdx.x, gridDim.x);
}

// one launch for a list of jobs: blockIdx.y selects the job, blockIdx.x the (i-block, split) inside it
__global__ void __launch_bounds__(kDwThreads, 1) k_dw_batch(const __grid_constant__ DwBatchArgs a) {
  const mcf_dw_job_t j = a.jobs[blockIdx.y];
  if (!j.enabled) return;
  mcf_dw_params_t p;
  if (j.p_src < 0 || j.p_src > 2 || j.q_src < 0 || j.q_src > 2 || a.base[j.p_src] == nullptr || a.base[j.q_src] == nullptr) {
    if (threadIdx.x == 0 && blockIdx.x == 0) atomicExch(&g_mcf_device_error, 0xBAD50000u | blockIdx.y);
    return;
  }
  p.p_base = a.base[j.p_src]; p.p_tile_bytes = a.tile_bytes[j.p_src]; p.p_off = j.p_off; p.p_cols = j.p_cols;
  p.q_base = a.base[j.q_src]; p.q_tile_bytes = a.tile_bytes[j.q_src]; p.q_off = j.q_off; p.q_cols = j.q_cols;
  p.out = a.staging + j.st_off; p.ld_out = j.ld; p.n_i = j.n_i; p.n_j = j.n_j;
  p.colsum_p = j.colsum_off >= 0 ? a.staging + j.colsum_off : nullptr;
  p.n_tiles = a.n_tiles; p.max_ctas = 0;
  const int nib = (j.n_i + 127) / 128;
  int n_cta = (int)gridDim.x - ((int)gridDim.x % nib);   // CTAs of this job that take part
  if ((long long)(n_cta / nib) > a.n_tiles) n_cta = (int)a.n_tiles * nib;
  if ((int)blockIdx.x >= n_cta) return;
  DwQ2 q2 = {a.base[2], a.tile_bytes[2], j.q2_off, (j.q_split >= 0 && a.base[2] != nullptr) ? j.q_split : (1 << 20)};
  dw_body(p, q2, nib, blockIdx.x, n_cta);
}

__device__ __forceinline__ void dw_body(const mcf_dw_params_t& p, const DwQ2& q2, int nib, int cta, int n_cta) {
  extern __shared__ __align__(1024) uint8_t smem[];
  DwCtl& ctl = *reinterpret_cast<DwCtl*>(smem + kDwOffCtl);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if ((smem_u32(smem) & 1023u) != 0u) {
    if (threadIdx.x == 0) atomicExch(&g_mcf_device_error, 0xA11C0001u);
    return;
  }
  const int ib = cta % nib;
  const int split = cta / nib, nsplit = n_cta / nib;
  const uint32_t q_blocks = (uint32_t)(p.q_cols + 63) / 64;
  const uint32_t n_mma = q_blocks * 64u;  // whole 64-column atoms only (canonical MN-major SW128 shapes)
  const bool want_colsum = p.colsum_p != nullptr;

  // ones tile: column 0 of every row = 1.0 (bf16), rest 0
  for (int i = threadIdx.x; i < (int)(kBlkD / 16); i += kDwThreads)
    reinterpret_cast<uint4*>(smem + kDwOffOnes)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  if (threadIdx.x < 128)
    *reinterpret_cast<uint16_t*>(smem + kDwOffOnes + sw128_off(threadIdx.x, 0)) = 0x3F80u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kDwStages; ++s) {
      mbar_init(&ctl.full[s], 1);
      mbar_init(&ctl.empty[s], 1);
    }
    mbar_init(&ctl.done, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(&ctl.tmem_base, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl.tmem_base;

  long long my_tiles = 0;
  for (long long t = split; t < p.n_tiles; t += nsplit) ++my_tiles;

  if (warp == 0 || warp == 2 || warp == 3) {
    const uint32_t me = warp == 0 ? 0u : (uint32_t)warp - 1u;
    if (lane == 0 && me < (uint32_t)kDwProducers) {
      uint32_t stage = 0, phase = 0;
      const uint8_t* pb = reinterpret_cast<const uint8_t*>(p.p_base);
      const uint8_t* qb = reinterpret_cast<const uint8_t*>(p.q_base);
      for (long long t = split; t < p.n_tiles; t += nsplit) {
        const uint8_t* psrc = pb + t * p.p_tile_bytes + p.p_off + (uint32_t)ib * 2u * kBlkD;
        const uint8_t* qsrc = qb + t * p.q_tile_bytes + p.q_off;
        const uint8_t* q2src = q2.base ? q2.base + t * q2.tile_bytes + q2.off : nullptr;
        for (uint32_t half = 0; half < 2; ++half) {
          mbar_wait(&ctl.empty[stage], phase ^ 1u, 0x500u | stage);
          // producer 0 posts the stage's byte count; the others' complete_tx may land first (the phase cannot
          // complete before producer 0's arrival)
          if (me == 0) mbar_arrive_expect_tx(&ctl.full[stage], (2 + q_blocks) * kHalf);
          uint8_t* dst = smem + stage * kStageBytes;
          for (uint32_t b = me; b < 2 + q_blocks; b += kDwProducers) {
            const uint8_t* src = b < 2 ? psrc + b * kBlkD
                                 : ((int)(b - 2) < q2.split ? qsrc + (b - 2) * kBlkD
                                                            : q2src + (b - 2 - (uint32_t)q2.split) * kBlkD);
            bulk_g2s(dst + b * kHalf, src + half * kHalf, kHalf, &ctl.full[stage]);
          }
          if (++stage == kDwStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      const uint32_t idesc = make_idesc(n_mma, true, true);
      const uint32_t idesc1 = make_idesc(64, true, true);
      const uint32_t ones_addr = smem_u32(smem + kDwOffOnes);
      bool first = true;
      for (long long t = split; t < p.n_tiles; t += nsplit) {
        for (uint32_t half = 0; half < 2; ++half) {
          mbar_wait(&ctl.full[stage], phase, 0x600u | stage);
          tc_fence_after();
          const uint32_t a_base = smem_u32(smem + stage * kStageBytes);
          const uint32_t b_base = a_base + 2 * kHalf;
          const uint32_t acc0 = first ? 0u : 1u;
#pragma unroll 1
          for (uint32_t k = 0; k < 4; ++k) {  // 4 x 16 rows
            const uint64_t ad = make_sdesc(a_base + k * 2048u, kHalf, 1024u);
            const uint64_t bd = make_sdesc(b_base + k * 2048u, kHalf, 1024u);
            umma_bf16(tmem_base, ad, bd, idesc, k == 0 ? acc0 : 1u);
          }
          if (want_colsum) {
#pragma unroll 1
            for (uint32_t k = 0; k < 4; ++k) {
              const uint64_t ad = make_sdesc(a_base + k * 2048u, kHalf, 1024u);
              const uint64_t od = make_sdesc(ones_addr + k * 2048u, kBlkD, 1024u);
              umma_bf16(tmem_base + 256, ad, od, idesc1, k == 0 ? acc0 : 1u);
            }
          }
          first = false;
          umma_commit(&ctl.empty[stage]);
          if (++stage == kDwStages) { stage = 0; phase ^= 1u; }
        }
      }
      umma_commit(&ctl.done);
    }
  } else if (warp >= 4 && my_tiles > 0) {
    const int qtr = warp & 3;
    const uint32_t row = qtr * 32 + lane;
    const int i = ib * 128 + (int)row;
    mbar_wait(&ctl.done, 0u, 0x700u);
    tc_fence_after();
    const uint32_t t_row = tmem_base + ((uint32_t)(qtr * 32) << 16);
    for (uint32_t c0 = 0; c0 < n_mma; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(t_row + c0, v);
      tmem_ld_wait();
      if (i < p.n_i) {
        float* dst = p.out + (long long)i * p.ld_out + c0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          if ((int)c0 + q * 4 < p.n_j) {  // n_j is padded to a multiple of 4 by the host plan
            float4 a = make_float4(__uint_as_float(v[q * 4 + 0]), __uint_as_float(v[q * 4 + 1]),
                                   __uint_as_float(v[q * 4 + 2]), __uint_as_float(v[q * 4 + 3]));
            atomicAdd(reinterpret_cast<float4*>(dst + q * 4), a);
          }
        }
      }
    }
    if (want_colsum) {
      uint32_t v[16];
      tmem_ld16(t_row + 256, v);
      tmem_ld_wait();
      if (i < p.n_i) atomicAdd(p.colsum_p + i, __uint_as_float(v[0]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// per-ray feature columns -> bf16 tile images (thread = row, 8 x 16 B per row)
__global__ void k_rayfeat_image(const float* __restrict__ rayfeat, int stride, int dim, long long n_rows,
                                int rows_per_ray, uint8_t* __restrict__ out) {
  const long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x;   // padded row index (tile * 128 + row)
  const long long n_pad = (n_rows + MCF_TILE_ROWS - 1) / MCF_TILE_ROWS * MCF_TILE_ROWS;
  if (m >= n_pad) return;
  const bool valid = m < n_rows;
  const float* rf = rayfeat + (valid ? m / rows_per_ray : 0) * stride;
  uint8_t* blk = out + (m / MCF_TILE_ROWS) * (long long)MCF_BLOCK_BYTES;
  const uint32_t row = (uint32_t)(m % MCF_TILE_ROWS);
  for (int c8 = 0; c8 < 8; ++c8) {
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = c8 * 8 + j;
      f[j] = (valid && c < dim) ? __ldg(rf + c) : 0.f;
    }
    uint4 v;
    v.x = pack_bf16x2(f[0], f[1]); v.y = pack_bf16x2(f[2], f[3]);
    v.z = pack_bf16x2(f[4], f[5]); v.w = pack_bf16x2(f[6], f[7]);
    *reinterpret_cast<uint4*>(blk + sw128_off(row, c8)) = v;
  }
}

}  // namespace mcf

extern "C" int mcf_rayfeat_image(const float* rayfeat, int rayfeat_stride, int rayfeat_dim, long long n_rows,
                                 int rows_per_ray, void* out, cudaStream_t stream) {
  if (n_rows <= 0) return 0;
  if (!rayfeat || !out || rows_per_ray <= 0 || rayfeat_dim < 0 || rayfeat_dim > 64) return MCF_ERR_BAD_ARG;
  const long long n_pad = (n_rows + MCF_TILE_ROWS - 1) / MCF_TILE_ROWS * MCF_TILE_ROWS;
  mcf::k_rayfeat_image<<<(unsigned)(n_pad / 128), 128, 0, stream>>>(rayfeat, rayfeat_stride, rayfeat_dim, n_rows,
                                                                   rows_per_ray, reinterpret_cast<uint8_t*>(out));
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

extern "C" int mcf_dw_gemm_batch(const mcf_dw_job_t* jobs_dev, int n_jobs, const void* fwd_save,
                                 long long fwd_tile_bytes, const void* bwd_save, long long bwd_tile_bytes,
                                 const void* aux, long long aux_tile_bytes, float* staging, long long n_tiles,
                                 int ctas_per_job, cudaStream_t stream) {
  if (n_jobs <= 0 || n_tiles <= 0) return 0;
  if (ctas_per_job < 2) return MCF_ERR_BAD_ARG;
  mcf::DwBatchArgs a;
  a.jobs = jobs_dev;
  a.base[0] = reinterpret_cast<const uint8_t*>(fwd_save);
  a.base[1] = reinterpret_cast<const uint8_t*>(bwd_save);
  a.tile_bytes[0] = fwd_tile_bytes;
  a.tile_bytes[1] = bwd_tile_bytes;
  a.base[2] = reinterpret_cast<const uint8_t*>(aux);
  a.tile_bytes[2] = aux_tile_bytes;
  a.staging = staging;
  a.n_tiles = n_tiles;
  cudaError_t e = cudaFuncSetAttribute(mcf::k_dw_batch, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mcf::kDwSmem);
  if (e != cudaSuccess) return (int)e;
  mcf::k_dw_batch<<<dim3((unsigned)ctas_per_job, (unsigned)n_jobs), mcf::kDwThreads, mcf::kDwSmem, stream>>>(a);
  e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

extern "C" int mcf_dw_gemm(const mcf_dw_params_t* pp, cudaStream_t stream) {
  if (!pp) return MCF_ERR_BAD_ARG;
  const mcf_dw_params_t& p = *pp;
  if (p.n_tiles <= 0) return 0;
  if (p.n_i <= 0 || p.n_j <= 0 || p.n_j > 256 || (p.n_j & 3) || (p.ld_out & 3) || p.p_cols % 128 != 0 ||
      p.n_i > p.p_cols || p.q_cols > 256 || p.n_j > ((p.q_cols + 63) / 64) * 64)
    return MCF_ERR_BAD_ARG;
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    e = cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return (int)e;
  }
  const int nib = (p.n_i + 127) / 128;
  int cap = p.max_ctas > 0 ? p.max_ctas : n_sm;
  long long nsplit = cap / nib;
  if (nsplit < 1) nsplit = 1;
  if (nsplit > p.n_tiles) nsplit = p.n_tiles;
  cudaError_t e = cudaFuncSetAttribute(mcf::k_dw, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mcf::kDwSmem);
  if (e != cudaSuccess) return (int)e;
  mcf::k_dw<<<(unsigned)(nsplit * nib), mcf::kDwThreads, mcf::kDwSmem, stream>>>(p, nib);
  e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}
