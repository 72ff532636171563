// Microbenchmark: the ReLU epilogue loop (TMEM ld -> +bias (FADD2) -> bf16x2 relu cvt -> swizzled st.shared) in isolation.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../moco_flow_b200/csrc/ptx.cuh"
using namespace mcf;
__device__ __forceinline__ void store_h8(uint8_t* hbuf, uint32_t row, uint32_t col0, uint4 v) {
  uint32_t block = col0 >> 6, c16 = (col0 & 63u) >> 3;
  *reinterpret_cast<uint4*>(hbuf + block * 16384 + sw128_off(row, c16)) = v;
}
__device__ __forceinline__ void load32f(const float* __restrict__ p, float (&b)[32]) {
#pragma unroll
  for (int q = 0; q < 8; ++q) { float4 t = __ldg(reinterpret_cast<const float4*>(p) + q); b[q*4]=t.x; b[q*4+1]=t.y; b[q*4+2]=t.z; b[q*4+3]=t.w; }
}
__device__ unsigned g_sink;
__device__ __forceinline__ uint32_t pack_int_relu(uint32_t lo, uint32_t hi) {
  // relu via signed max, round-half-up via +0x8000, pack the two upper halves
  int a = max((int)lo, 0), b = max((int)hi, 0);
  uint32_t ra = (uint32_t)a + 0x8000u, rb = (uint32_t)b + 0x8000u;
  return __byte_perm(ra, rb, 0x7632);
}
template <int VAR>
__device__ __forceinline__ void chunk_var(uint8_t* hbuf, uint32_t row, uint32_t col0, uint32_t (&v)[32], const float (&b)[32], unsigned& acc) {
  if (VAR != 3 && VAR != 4 && VAR != 7) {
#pragma unroll
    for (int j = 0; j < 16; ++j) add_f32x2(v[2*j], v[2*j+1], b[2*j], b[2*j+1]);
  }
  if (VAR == 4) {
#pragma unroll
    for (int j = 0; j < 32; ++j) acc ^= v[j];
    return;
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint4 o;
    if (VAR == 3 || VAR == 6) { o.x = v[q*8] ^ v[q*8+1]; o.y = v[q*8+2] ^ v[q*8+3]; o.z = v[q*8+4] ^ v[q*8+5]; o.w = v[q*8+6] ^ v[q*8+7]; }
    else if (VAR == 5) { o.x = pack_int_relu(v[q*8], v[q*8+1]); o.y = pack_int_relu(v[q*8+2], v[q*8+3]); o.z = pack_int_relu(v[q*8+4], v[q*8+5]); o.w = pack_int_relu(v[q*8+6], v[q*8+7]); }
    else {
      o.x = cvt_bf16x2_relu_bits(v[q*8+0], v[q*8+1]); o.y = cvt_bf16x2_relu_bits(v[q*8+2], v[q*8+3]);
      o.z = cvt_bf16x2_relu_bits(v[q*8+4], v[q*8+5]); o.w = cvt_bf16x2_relu_bits(v[q*8+6], v[q*8+7]);
    }
    if (VAR == 2) acc ^= o.x ^ o.y ^ o.z ^ o.w; else store_h8(hbuf, row, col0 + q*8, o);
  }
}
__device__ __forceinline__ void chunk(uint8_t* hbuf, uint32_t row, uint32_t col0, uint32_t (&v)[32], const float (&b)[32]) {
#pragma unroll
  for (int j = 0; j < 16; ++j) add_f32x2(v[2*j], v[2*j+1], b[2*j], b[2*j+1]);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint4 o;
    o.x = cvt_bf16x2_relu_bits(v[q*8+0], v[q*8+1]); o.y = cvt_bf16x2_relu_bits(v[q*8+2], v[q*8+3]);
    o.z = cvt_bf16x2_relu_bits(v[q*8+4], v[q*8+5]); o.w = cvt_bf16x2_relu_bits(v[q*8+6], v[q*8+7]);
    store_h8(hbuf, row, col0 + q*8, o);
  }
}
// mode 0: serial (ld, wait, process); mode 1: software pipelined (as in chain.cu)
__global__ void __launch_bounds__(384, 1) k(int nslots, int iters, int mode, const float* bias, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tbase;
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) { tmem_alloc(&tbase, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  __syncthreads();
  long long t0 = clock64();
  if (warp >= 4 && warp < 4 + 4 * nslots) {
    int s = (warp - 4) >> 2, qtr = warp & 3;
    uint32_t row = qtr * 32 + lane;
    uint8_t* hbuf = smem + s * 65536;
    uint32_t t_acc = tbase + ((uint32_t)(qtr * 32) << 16) + s * 256;
    for (int it = 0; it < iters; ++it) {
      if (mode >= 2) {
        unsigned acc = 0;
        for (int c0 = 0; c0 < 256; c0 += 32) {
          uint32_t v[32]; float b[32];
          tmem_ld32(t_acc + c0, v); if (mode != 3 && mode != 4 && mode != 7) load32f(bias + c0, b); tmem_ld_wait();
          if (mode == 2) chunk_var<2>(hbuf, row, c0, v, b, acc);
          else if (mode == 3) chunk_var<3>(hbuf, row, c0, v, b, acc);
          else if (mode == 5) chunk_var<5>(hbuf, row, c0, v, b, acc);
          else if (mode == 6) chunk_var<6>(hbuf, row, c0, v, b, acc);
          else if (mode == 7) chunk_var<7>(hbuf, row, c0, v, b, acc);
          else chunk_var<4>(hbuf, row, c0, v, b, acc);
        }
        if (acc == 0x1234567u) g_sink = acc;
      } else if (mode == 0) {
        for (int c0 = 0; c0 < 256; c0 += 32) {
          uint32_t v[32]; float b[32];
          tmem_ld32(t_acc + c0, v); load32f(bias + c0, b); tmem_ld_wait();
          chunk(hbuf, row, c0, v, b);
        }
      } else {
        uint32_t va[32], vb[32]; float b0[32], b1[32];
        load32f(bias, b0);
        tmem_ld32(t_acc, va);
        for (int c0 = 0; c0 < 256; c0 += 64) {
          load32f(bias + c0 + 32, b1);
          tmem_ld_wait();
          tmem_ld32(t_acc + c0 + 32, vb);
          chunk(hbuf, row, c0, va, b0);
          bool more = c0 + 64 < 256;
          if (more) load32f(bias + c0 + 64, b0);
          tmem_ld_wait();
          if (more) tmem_ld32(t_acc + c0 + 64, va);
          chunk(hbuf, row, c0 + 32, vb, b1);
        }
      }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 128) out[blockIdx.x] = (unsigned long long)(t1 - t0);
  __syncthreads();
  // keep the shared-memory stores alive
  unsigned acc = 0;
  for (int i = threadIdx.x; i < 131072 / 4; i += blockDim.x) acc ^= reinterpret_cast<uint32_t*>(smem)[i];
  if (acc == 0x12345u) out[147] = acc;
  if (warp == 0) tmem_dealloc(tbase, 512);
}
int main() {
  unsigned long long* out; float* bias;
  cudaMalloc(&out, 148 * 8); cudaMalloc(&bias, 4096); cudaMemset(bias, 0, 4096);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
  for (int mode : {0, 5, 6, 7, 4}) for (int ns : {2}) {
    int iters = 512;
    k<<<148, 384, 131072>>>(ns, iters, mode, bias, out);
    cudaError_t e1 = cudaGetLastError();
    cudaError_t e2 = cudaDeviceSynchronize();
    if (e1 != cudaSuccess || e2 != cudaSuccess) printf("launch %s sync %s\n", cudaGetErrorString(e1), cudaGetErrorString(e2));
    unsigned long long h[148]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("raw h0=%llu h1=%llu ", h[0], h[1]);
    printf("mode=%d slots=%d: cycles per 256-col round = %.0f (per 32-col chunk %.0f)  err=%s\n", mode, ns, (double)h[0] / iters,
           (double)h[0] / iters / 8, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
