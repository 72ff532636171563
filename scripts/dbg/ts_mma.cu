// tcgen05.mma with the A operand in tensor memory ("TS" form): layout check + issue rate (run on a B200).
//  (1) D[128 x 64] = A[128 x 64] * B[64 x 64]^T with A written to TMEM by tcgen05.st.32x32b (thread = row = lane),
//      two bf16 per 32-bit column; which half holds the even k is probed (variant 0: even k in the low half).
//      The same product with A in shared memory (SS form) validates the harness.
//  (2) rate of M=128 N=128 K=16 TS-form MMAs (the NoF layer shape) against the SS form.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../../moco_flow_b200/csrc/ptx.cuh"
using namespace mcf;

__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

__host__ __device__ inline float a_val(int m, int k) { return (float)((m * 7 + k * 3) % 17 - 8) / 8.0f; }
__host__ __device__ inline float b_val(int n, int k) { return (float)((n * 5 + k * 11) % 13 - 6) / 4.0f; }

// mode 0: SS (A in smem), 1: TS even-k-low, 2: TS odd-k-low
__global__ void __launch_bounds__(128, 1) k_check(int mode, float* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tbase;
  __shared__ uint64_t bar;
  const int warp = threadIdx.x >> 5, row = threadIdx.x;
  uint8_t* sA = smem;            // [128][64] bf16, SW128 K-major (16 KB)
  uint8_t* sB = smem + 16384;    // [64][64] bf16, SW128 K-major (8 KB)
  for (int c8 = 0; c8 < 8; ++c8) {
    uint32_t w[4];
    for (int j = 0; j < 4; ++j) w[j] = pack_bf16x2(a_val(row, c8 * 8 + 2 * j), a_val(row, c8 * 8 + 2 * j + 1));
    *reinterpret_cast<uint4*>(sA + sw128_off(row, c8)) = make_uint4(w[0], w[1], w[2], w[3]);
    if (row < 64) {
      for (int j = 0; j < 4; ++j) w[j] = pack_bf16x2(b_val(row, c8 * 8 + 2 * j), b_val(row, c8 * 8 + 2 * j + 1));
      *reinterpret_cast<uint4*>(sB + sw128_off(row, c8)) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&tbase, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t t_row = tbase + ((uint32_t)(warp * 32) << 16);
  const uint32_t a_col = 256;
  if (mode != 0) {
    uint32_t w[32];
    for (int j = 0; j < 32; ++j) {
      float e = a_val(row, 2 * j), o = a_val(row, 2 * j + 1);
      w[j] = mode == 1 ? pack_bf16x2(e, o) : pack_bf16x2(o, e);
    }
    tmem_st32(t_row + a_col, w);
    tmem_st_wait();
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (threadIdx.x == 32) {
    const uint32_t idesc = make_idesc(64u);
    const uint64_t ad = make_sdesc(smem_u32(sA), 0, 1024), bd = make_sdesc(smem_u32(sB), 0, 1024);
    for (uint32_t k = 0; k < 4; ++k) {
      if (mode == 0) umma_bf16(tbase, ad + 2u * k, bd + 2u * k, idesc, k ? 1u : 0u);
      else umma_bf16_ts(tbase, tbase + a_col + 8u * k, bd + 2u * k, idesc, k ? 1u : 0u);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  uint32_t v[32];
  for (int c0 = 0; c0 < 64; c0 += 32) {
    tmem_ld32(t_row + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[row * 64 + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 512);
}

// n MMAs M=128 N=128 K=16 on resident operands, SS or TS form
// dmode: how consecutive MMAs map to accumulators -- 0: four K-steps into one accumulator, then the other (what a layer
// does); 2 / 4: round-robin over 2 / 4 independent accumulators (consecutive MMAs never depend on each other)
__global__ void __launch_bounds__(128, 1) k_rate(int n_mma, int ts, unsigned long long* out, int dmode = 0, int ncols = 128) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tbase;
  __shared__ uint64_t bar;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&tbase, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  {
    uint32_t w[32];
    for (int j = 0; j < 32; ++j) w[j] = 0x3c003c00u;
    tmem_st32(tbase + ((uint32_t)(warp * 32) << 16) + 128, w);
    tmem_st32(tbase + ((uint32_t)(warp * 32) << 16) + 160, w);
    tmem_st_wait();
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (threadIdx.x == 32) {
    const uint32_t a = smem_u32(smem), b = smem_u32(smem) + 16384;
    const uint32_t idesc = make_idesc((uint32_t)ncols);
    const uint64_t ad = make_sdesc(a, 0, 1024), bd = make_sdesc(b, 0, 1024);
    long long t0 = clock64();
    for (int i = 0; i < n_mma; ++i) {
      const uint32_t k = i & 3;
      // dmode 0: accumulators 256 columns apart (one per slot), 4 K-steps each; dmode 1: the same, 128 columns apart;
      // dmode 2: alternate between the two 256-apart accumulators on every MMA
      const uint32_t d = tbase + (dmode == 0 ? ((i >> 2) & 1) * 256 : (dmode == 1 ? ((i >> 2) & 1) * 128 + 256 : (uint32_t)(i & 1) * 256));
      if (ts) umma_bf16_ts(d, tbase + 128 + 8u * k, bd + 2u * k, idesc, (i > 7) ? 1u : 0u);
      else umma_bf16(d, ad + 2u * k, bd + 2u * k, idesc, (i > 7) ? 1u : 0u);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    out[blockIdx.x] = (unsigned long long)(t2 - t0);
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 512);
}

int main() {
  float* out;
  cudaMalloc(&out, 128 * 64 * 4);
  std::vector<float> h(128 * 64), ref(128 * 64);
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < 64; ++n) {
      double s = 0;
      for (int k = 0; k < 64; ++k) s += (double)a_val(m, k) * (double)b_val(n, k);
      ref[m * 64 + n] = (float)s;
    }
  cudaFuncSetAttribute(k_check, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const char* names[3] = {"SS (A in smem)", "TS, even k in the low half", "TS, odd k in the low half"};
  for (int mode = 0; mode < 3; ++mode) {
    cudaMemset(out, 0, 128 * 64 * 4);
    k_check<<<1, 128, 64 * 1024>>>(mode, out);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h.data(), out, 128 * 64 * 4, cudaMemcpyDeviceToHost);
    double mx = 0;
    for (int i = 0; i < 128 * 64; ++i) mx = fmax(mx, fabs((double)h[i] - ref[i]));
    printf("check %-30s max abs err %.6f   (%s)  D[0][0..3] = %g %g %g %g  ref %g %g %g %g\n", names[mode], mx,
           cudaGetErrorString(e), h[0], h[1], h[2], h[3], ref[0], ref[1], ref[2], ref[3]);
    if (e != cudaSuccess) return 1;
  }
  unsigned long long* t;
  cudaMalloc(&t, 148 * 8);
  unsigned long long ht[148];
  cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int ncols : {128, 64})
    for (int ts = 0; ts < 2; ++ts)
      for (int dmode : {0, 1, 2}) {
        k_rate<<<148, 128, 64 * 1024>>>(1024, ts, t, dmode, ncols);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(ht, t, sizeof(ht), cudaMemcpyDeviceToHost);
        printf("rate %s M=128 N=%3d K=16 n=1024 grid=148 accumulators %s: %.1f clk/mma  (%s)\n", ts ? "TS" : "SS", ncols,
               dmode == 0 ? "0/256, 4 K-steps each" : (dmode == 1 ? "256/384, 4 K-steps each" : "0/256, alternating every MMA"),
               (double)ht[0] / 1024, cudaGetErrorString(e));
      }
  return 0;
}
