// NoF flow MLP (hidden width 128) as a TMEM-resident chain on tcgen05 (sm_100a only).
//
// The generic chain kernel (chain.cu) keeps activations in shared memory: at M = 128, N = 128 the tensor core then
// reads 64 KB of operands per layer tile, the epilogue writes 32 KB back and (training) a bulk store reads them
// again -- 128 B/clk of shared-memory traffic at the MMA's peak rate, which is the SM's whole shared-memory
// bandwidth, and every tile re-streams 132 KB of weights through a ring.  This kernel removes all three:
//   * weights: the whole packed stream of the program (<= 144 KB) is copied into shared memory ONCE per CTA;
//   * activations: the A operand of every layer lives in TENSOR MEMORY (tcgen05.mma with A in TMEM): the epilogue
//     writes relu(acc + b) as packed bf16 straight from registers with tcgen05.st -- thread = row = TMEM lane, two K
//     elements per 32-bit column, so no swizzle arithmetic, no st.shared, no operand re-read from shared memory;
//   * epilogue latency: 8 epilogue warps per tile (two column halves), 16 per CTA, instead of 4 / 8.
// Shared memory is only touched by the B operand (32 KB per layer tile) and, in training, by the staging copy of the
// operand images that the weight-gradient GEMM needs in HBM (st.shared + one bulk store per layer).
//
// TMEM map of slot s (256 columns each, accumulators 256 columns apart -- see scripts/dbg/ts_mma.cu):
//   [  0,128) fp32 accumulator      [128,192) h / dY operand (128 bf16)      [192,224) x0 operand (64 bf16)
// Warp roles: warp 0 loads the weights, warp 1 lane 0 issues the MMAs, warp 2 owns the TMEM allocation,
// warps 4-11 / 12-19 are the epilogue of slot 0 / 1 (warp & 3 = TMEM lane quarter, (warp - 4) >> 2 & 1 = column half).
// The layer programs are the same tables as chain.cu's (plans.py builds them with resident = 2).
//
// Reference semantics: models/nof.py:55-85 (kornia 0.6.5 quaternion helpers), models/embedding.py:42-46.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/moco_flow_b200.h"
#include "ptx.cuh"
#include "nof_math.cuh"

namespace mcf {

namespace {

constexpr int kNofThreads = 640;
constexpr int kEpiThreads = 256;          // per slot
constexpr uint32_t kBlkN = MCF_BLOCK_BYTES;
constexpr uint32_t kNofResBytes = 147456;
constexpr uint32_t kNofNone = 0xFFFFFFFFu;
constexpr uint32_t kSlotStride = 256, kColH = 128, kColX0 = 192;
constexpr int kNofMaxChunks = 32, kNofMaxRounds = 16;

struct NofTab {
  mcf_chunk_t chunks[kNofMaxChunks];
  mcf_round_t rounds[kNofMaxRounds];
  uint64_t act_ready[2];
  uint64_t acc_full[2];
  uint64_t w_res;
  uint32_t tmem_base;
  uint32_t pad;
  float pe_freq[12];
  float pe_weight[12];
};

template <bool kStage>
struct NofSmem {
  static constexpr uint32_t off_w = 0;
  static constexpr uint32_t off_stage = kNofResBytes;                       // 2 x 32 KB operand images (training)
  static constexpr uint32_t off_tab = off_stage + (kStage ? 2u * 2u * kBlkN : 0u);
  static constexpr uint32_t total = off_tab + sizeof(NofTab);
};
static_assert(NofSmem<true>::total <= 232448, "shared memory budget exceeded");

// How the training saves (operand images for the weight-gradient GEMM) reach HBM:
//   0 = staged in shared memory, one bulk store of the whole image by the epilogue group's first thread (group barriers)
//   1 = staged, every epilogue warp bulk-stores its own 32 rows of the full-width rounds (no group barrier per round)
//   2 = no staging: every thread writes its row's 16-byte pieces straight to the image in HBM (no barrier, no wait)
#ifndef MCF_NOF_WARP_STORE
#define MCF_NOF_WARP_STORE 1
#endif
constexpr bool kWarpStore = MCF_NOF_WARP_STORE == 1;
constexpr bool kDirectSave = MCF_NOF_WARP_STORE == 2;

__device__ __forceinline__ void ld32f(const float* __restrict__ p, float (&b)[32]) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    float4 t;
    asm volatile("ld.global.nc.L1::evict_last.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w)
                 : "l"(reinterpret_cast<const float4*>(p) + q));
    b[q * 4 + 0] = t.x; b[q * 4 + 1] = t.y; b[q * 4 + 2] = t.z; b[q * 4 + 3] = t.w;
  }
}

// 16 packed bf16x2 words (32 columns starting at col0) -> the 128B-swizzled operand image (staging copy or HBM record)
__device__ __forceinline__ void stage32(uint8_t* img, uint32_t row, uint32_t col0, const uint32_t (&w)[16]) {
  const uint32_t block = col0 >> 6, c16 = (col0 & 63u) >> 3;
#pragma unroll
  for (int q = 0; q < 4; ++q)
    *reinterpret_cast<uint4*>(img + block * kBlkN + sw128_off(row, c16 + q)) =
        make_uint4(w[q * 4 + 0], w[q * 4 + 1], w[q * 4 + 2], w[q * 4 + 3]);
}

}  // namespace

// kBwd: backward dX program (head-gradient prologue, mask / PE-Jacobian epilogues); kSave: the launch writes operand
// images / masks / head values for the backward pass and the weight-gradient GEMM (always true for kBwd).
template <bool kBwd, bool kSave>
__global__ void __launch_bounds__(kNofThreads, 1) k_nof(const __grid_constant__ mcf_chain_params_t p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  using L = NofSmem<kSave>;
  NofTab& tab = *reinterpret_cast<NofTab*>(smem + L::off_tab);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if ((smem_u32(smem) & 1023u) != 0u) {
    if (threadIdx.x == 0) atomicExch(&g_mcf_device_error, 0xA11C0002u);
    return;
  }

  // ---- one-time setup ----
  {
    const uint32_t* src_c = reinterpret_cast<const uint32_t*>(p.chunks);
    uint32_t* dst_c = reinterpret_cast<uint32_t*>(tab.chunks);
    for (int i = threadIdx.x; i < p.n_chunks * 4; i += kNofThreads) dst_c[i] = src_c[i];
    const uint32_t* src_r = reinterpret_cast<const uint32_t*>(p.rounds);
    uint32_t* dst_r = reinterpret_cast<uint32_t*>(tab.rounds);
    for (int i = threadIdx.x; i < p.n_rounds * 8; i += kNofThreads) dst_r[i] = src_r[i];
    if (threadIdx.x < 10) {
      const int k = threadIdx.x;
      tab.pe_freq[k] = p.pe_table ? p.pe_table[k] : p.pe_freq[k];
      tab.pe_weight[k] = p.pe_table ? p.pe_table[MCF_MAX_FREQS + k] : p.pe_weight[k];
    }
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tab.act_ready[s], 8);   // one arrival per epilogue warp of the slot
      mbar_init(&tab.acc_full[s], 1);
    }
    mbar_init(&tab.w_res, 1);
    fence_mbar_init();
    const uint32_t total = p.wpack_bytes;
    mbar_arrive_expect_tx(&tab.w_res, total);
    const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.wpack);
    for (uint32_t off = 0; off < total; off += 32768u) {
      const uint32_t n = total - off < 32768u ? total - off : 32768u;
      bulk_g2s(smem + L::off_w + off, wsrc + off, n, &tab.w_res);
    }
  }
  if (warp == 2) {
    tmem_alloc(&tab.tmem_base, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tab.tmem_base;

  const long long n_tiles = (p.n_rows + MCF_TILE_ROWS - 1) / MCF_TILE_ROWS;
  const long long n_units = (n_tiles + 1) / 2;

  // register budget: the launch allocates 640 x 96; the control warpgroup gives back (96 - 32) x 128 = 8192 registers,
  // exactly what the 512 epilogue threads need to go from 96 to 112 (an unbalanced pair would block forever)
  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;\n");
    if (warp == 1 && lane == 0) {
      // =========================== MMA issuer ===========================
      uint32_t ar_phase[2] = {0u, 0u};
      const uint32_t w_addr = smem_u32(smem + L::off_w);
      mbar_wait(&tab.w_res, 0u, 0x700u);
      tc_fence_after();
      for (long long unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        for (int r = 0; r < p.n_rounds; ++r) {
          const int cb = tab.rounds[r].chunk_begin, ce = tab.rounds[r].chunk_end;
          for (int s = 0; s < 2; ++s) {
            if (2 * unit + s >= n_tiles) continue;
            mbar_wait(&tab.act_ready[s], ar_phase[s], 0x200u | s);
            ar_phase[s] ^= 1u;
            tc_fence_after();
            const uint32_t slot = tmem_base + s * kSlotStride;
            for (int c = cb; c < ce; ++c) {
              const mcf_chunk_t ch = tab.chunks[c];
              // A operand: 64 K elements (one "k block") = 32 TMEM columns, 8 columns per K = 16 step
              uint32_t a_tmem = slot + (ch.a_buf ? kColH : kColX0) + ch.a_kblock * 32u;
              uint64_t bd = make_sdesc(w_addr + ch.src_off, 0u, 1024u);
              const uint32_t idesc = make_idesc((uint32_t)ch.n);
              const uint32_t d_tmem = slot + ch.acc_col;
              uint32_t acc = (ch.flags & 1u) ? 0u : 1u;
              for (uint32_t k = 0; k < ch.ksteps; ++k) {
                umma_bf16_ts(d_tmem, a_tmem, bd, idesc, acc);
                a_tmem += 8u; bd += 2u; acc = 1u;
              }
            }
            umma_commit(&tab.acc_full[s]);
          }
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;\n");
    // =========================== epilogue groups ===========================
    const int s = (warp - 4) >> 3;            // slot
    const int half = ((warp - 4) >> 2) & 1;   // column half of the slot's rounds
    const int qtr = warp & 3;                 // TMEM lane quarter this warp may access
    const uint32_t row = qtr * 32 + lane;
    const int gtid = threadIdx.x - 128 - s * kEpiThreads;
    uint8_t* stage = smem + L::off_stage + (kSave ? s * 2u * kBlkN : 0u);
    const uint32_t t_row = tmem_base + ((uint32_t)(qtr * 32) << 16) + s * kSlotStride;
    uint32_t af_phase = 0;
    bool store_pending = false;   // a store of the whole staging image issued by the group's first thread
    bool warp_pending = false;    // stores of this warp's own 32 rows issued by its lane 0
    auto stage_free = [&]() {      // earlier bulk stores must have finished reading what is about to be overwritten
      if (kSave && kWarpStore && warp_pending) {
        if (lane == 0) bulk_wait_read_all();
        __syncwarp();
        warp_pending = false;
      }
      if (kSave && store_pending) {
        if (gtid == 0) bulk_wait_read_all();
        named_bar_sync(1 + s, kEpiThreads);
        store_pending = false;
      }
    };
    // A warp's 32 rows of one 64-column block are 4 KB contiguous in the swizzled image (a row is 128 B): in the
    // full-width rounds each warp owns such a piece outright and pushes it to HBM itself -- no group-wide barrier.
    auto warp_store = [&](uint8_t* dst, uint32_t block) {
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        const uint32_t off = block * kBlkN + (uint32_t)qtr * 4096u;
        bulk_s2g(dst + off, stage + off, 4096u);
        bulk_commit();
      }
      warp_pending = true;
    };
    // the x0 image is written by both column halves of a row (two warps): once the warps run unsynchronised through
    // the rounds, the first write of a tile has to wait for the other half's last store as well
    auto stage_free_shared = [&]() {
      stage_free();
      if (kSave && kWarpStore) named_bar_sync(1 + s, kEpiThreads);
    };
    auto stage_store = [&](uint8_t* dst, uint32_t nbytes) {   // staging image -> save record
      fence_proxy_async_smem();
      named_bar_sync(1 + s, kEpiThreads);
      if (gtid == 0) {
        bulk_s2g(dst, stage, nbytes);
        bulk_commit();
      }
      store_pending = true;
    };
    auto arrive = [&]() {
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tab.act_ready[s]);
    };

    for (long long unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
      const long long tile = 2 * unit + s;
      if (tile >= n_tiles) break;
      const bool saving = kSave && p.save != nullptr;
      const long long m = tile * MCF_TILE_ROWS + row;
      const bool valid = m < p.n_rows;
      const long long mc = valid ? m : (p.n_rows - 1);
      const long long ray = mc / p.rows_per_ray;
      uint8_t* save_tile = saving ? reinterpret_cast<uint8_t*>(p.save) + tile * p.save_tile_bytes : nullptr;
      float dx[3] = {0.f, 0.f, 0.f};   // backward: accumulated d_xyz (half 0 threads)

      // ------------------------- prologue: the first operand, into TMEM -------------------------
      if (!kBwd) {
        float mine[32];   // this thread's 32 of the 64 x0 columns: [32 * half, 32 * half + 32)
#pragma unroll
        for (int c = 0; c < 32; ++c) mine[c] = 0.f;
        if (p.prologue == MCF_PRO_PE_XYZ) {
          float x[3] = {0.f, 0.f, 0.f};
          if (valid) { x[0] = p.xyz[m * 3 + 0]; x[1] = p.xyz[m * 3 + 1]; x[2] = p.xyz[m * 3 + 2]; }
          // channel order of models/embedding.py:42-46: [x | w0 sin(f0 x) | w0 cos(f0 x) | w1 sin(f1 x) | ...]
          if (half == 0) { mine[0] = x[0]; mine[1] = x[1]; mine[2] = x[2]; }
          float sn[3] = {0.f, 0.f, 0.f}, cs[3] = {1.f, 1.f, 1.f};
          bool have_prev = false;
#pragma unroll
          for (int k = 0; k < 10; ++k) {
            // frequency k owns channels [3 + 6k, 9 + 6k): skipped by the half that holds none of them
            const bool needed = (half == 0) ? (3 + 6 * k < 32) : (8 + 6 * k >= 32);
            if (k < p.pe_n_freqs && needed) {
              const float f = tab.pe_freq[k], w = tab.pe_weight[k];
              // full-range sincosf at every 4th octave; in between sin/cos(2a) from sin/cos(a) (<= 3 doublings,
              // error <= ~8 ulp, far below the bf16 rounding of the operand); non-octave tables stay exact
              const bool exact = !have_prev || (k & 3) == 0 || f != 2.0f * tab.pe_freq[k > 0 ? k - 1 : 0];
#pragma unroll
              for (int c = 0; c < 3; ++c) {
                if (exact) {
                  sincosf(f * x[c], &sn[c], &cs[c]);
                } else {
                  const float s2 = 2.0f * sn[c] * cs[c];
                  const float c2 = fmaf(-2.0f * sn[c], sn[c], 1.0f);
                  sn[c] = s2;
                  cs[c] = c2;
                }
                const int is = 3 + 6 * k + c, ic = is + 3;   // compile-time channel numbers
                if ((is >> 5) == half) mine[is & 31] = w * sn[c];
                if ((ic >> 5) == half) mine[ic & 31] = w * cs[c];
              }
              have_prev = true;
            }
          }
        } else if (p.prologue == MCF_PRO_DENSE) {
          const float* src = p.dense + mc * p.dense_stride + 32 * half;
#pragma unroll
          for (int c = 0; c < 32; ++c) mine[c] = (valid && 32 * half + c < p.dense_cols) ? src[c] : 0.f;
        } else {
          if (gtid == 0) atomicExch(&g_mcf_device_error, 0xBADF0000u | (uint32_t)p.prologue);
        }
        uint32_t w16[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) w16[j] = pack_bf16x2(mine[2 * j], mine[2 * j + 1]);
        tmem_st16(t_row + kColX0 + 16u * half, w16);
        arrive();   // the tensor core only needs the TMEM operand: the HBM copy below overlaps its first layer
        if (saving && p.x0_save_off != kNofNone) {
          if (kDirectSave) {
            stage32(save_tile + p.x0_save_off, row, 32u * half, w16);
          } else {
            stage_free_shared();
            stage32(stage, row, 32u * half, w16);
            stage_store(save_tile + p.x0_save_off, kBlkN);
          }
        }
      } else {
        // backward of the flow head (models/nof.py:75-82): d{v,s,t} as a K = 16 operand, dL/dx into dx
        if (half == 0) {
          float g[3] = {0.f, 0.f, 0.f}, hs[12];
#pragma unroll
          for (int j = 0; j < 12; ++j) hs[j] = 0.f;
          if (valid) {
            g[0] = p.g_out[m * 3 + 0]; g[1] = p.g_out[m * 3 + 1]; g[2] = p.g_out[m * 3 + 2];
            const float4* hp = reinterpret_cast<const float4*>(p.head_save + m * 12);
            float4 a = hp[0], b = hp[1], c = hp[2];
            hs[0] = a.x; hs[1] = a.y; hs[2] = a.z; hs[3] = a.w; hs[4] = b.x; hs[5] = b.y;
            hs[6] = b.z; hs[7] = b.w; hs[8] = c.x; hs[9] = c.y; hs[10] = c.z; hs[11] = c.w;
          }
          float d16[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) d16[j] = 0.f;
          if (p.use_quat) {
            nof_quat_backward(hs, hs + 9, g, d16, dx);
          } else {  // out = head3 + x   (models/nof.py:82)
            d16[0] = g[0]; d16[1] = g[1]; d16[2] = g[2];
            dx[0] = g[0]; dx[1] = g[1]; dx[2] = g[2];
          }
          if (valid && p.d_head) {
            float4* dh = reinterpret_cast<float4*>(p.d_head + m * 12);
            dh[0] = make_float4(d16[0], d16[1], d16[2], d16[3]);
            dh[1] = make_float4(d16[4], d16[5], d16[6], d16[7]);
            dh[2] = make_float4(d16[8], 0.f, 0.f, 0.f);
          }
          uint32_t w8[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) w8[j] = pack_bf16x2(d16[2 * j], d16[2 * j + 1]);
          tmem_st8(t_row + kColH, w8);
          if (saving && p.x0_save_off != kNofNone) {
            if (!kDirectSave) stage_free_shared();
            uint32_t w16[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) w16[j] = j < 8 ? w8[j] : 0u;
            stage32(kDirectSave ? save_tile + p.x0_save_off : stage, row, 0u, w16);
          }
        } else if (saving && p.x0_save_off != kNofNone) {
          if (!kDirectSave) stage_free_shared();
          uint32_t w16[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) w16[j] = 0u;
          stage32(kDirectSave ? save_tile + p.x0_save_off : stage, row, 32u, w16);
        }
        if (p.prologue != MCF_PRO_B_NOF && gtid == 0) atomicExch(&g_mcf_device_error, 0xBADF0000u | (uint32_t)p.prologue);
        arrive();
        if (!kDirectSave && saving && p.x0_save_off != kNofNone) stage_store(save_tile + p.x0_save_off, kBlkN);
      }

      // ------------------------------- rounds -------------------------------
      for (int r = 0; r < p.n_rounds; ++r) {
        const mcf_round_t rd = tab.rounds[r];
        const int n_half = rd.n_out >> 1;                  // columns of this thread in a full-width round
        const int c_lo = half * n_half;
        const float* bias_p = (rd.raybias >= 0) ? (p.raybias[rd.raybias] + ray * rd.n_out) : (p.consts + rd.const_off);
        const bool relu_round = !kBwd && rd.epi == MCF_EPI_RELU;
        const bool mask_round = kBwd && rd.epi == MCF_EPI_B_MASK;
        const bool writes_h = relu_round || mask_round;
        float b[32];
        uint32_t mw[2] = {0xFFFFFFFFu, 0xFFFFFFFFu};
        if (relu_round) ld32f(bias_p + c_lo, b);
        if (mask_round && rd.mask_off != kNofNone) {
          const uint32_t* mkp = p.fwd_masks + tile * p.fwd_mask_tile_words + rd.mask_off + row;
          mw[0] = __ldg(mkp + ((c_lo >> 5) + 0) * 128);
          mw[1] = __ldg(mkp + ((c_lo >> 5) + 1) * 128);
        }
        mbar_wait(&tab.acc_full[s], af_phase, 0x400u | s);
        af_phase ^= 1u;
        tc_fence_after();
        const uint32_t t_acc = t_row + rd.acc_col;
        if (!kDirectSave && writes_h && saving && rd.save_off != kNofNone) stage_free();

        if (relu_round) {
          const bool want_mask = kSave && p.masks != nullptr && rd.mask_off != kNofNone;
          uint32_t* mk = want_mask ? (p.masks + tile * p.mask_tile_words + rd.mask_off + row) : nullptr;
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int c0 = c_lo + 32 * i;
            uint32_t v[32];
            tmem_ld32(t_acc + c0, v);
            if (i == 1) ld32f(bias_p + c0, b);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) add_f32x2(v[2 * j], v[2 * j + 1], b[2 * j], b[2 * j + 1]);
            if (want_mask) {
              uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0;
#pragma unroll
              for (int j = 7; j >= 0; --j) {
                w0 = __funnelshift_l(v[j], w0, 1);
                w1 = __funnelshift_l(v[8 + j], w1, 1);
                w2 = __funnelshift_l(v[16 + j], w2, 1);
                w3 = __funnelshift_l(v[24 + j], w3, 1);
              }
              mk[(c0 >> 5) * 128] = ~(w0 | (w1 << 8) | (w2 << 16) | (w3 << 24));
            }
            uint32_t w16[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) w16[j] = cvt_bf16x2_relu_bits(v[2 * j], v[2 * j + 1]);
            tmem_st16(t_row + kColH + (uint32_t)(c0 >> 1), w16);
            if (saving && rd.save_off != kNofNone)
              stage32(kDirectSave ? save_tile + rd.save_off : stage, row, (uint32_t)c0, w16);
          }
        } else if (!kBwd && rd.epi == MCF_EPI_NOF_HEAD) {
          if (half == 0) {
            uint32_t v[16];
            tmem_ld16(t_acc, v);
            tmem_ld_wait();
            float h9[9], x[3] = {0.f, 0.f, 0.f}, o[3];
#pragma unroll
            for (int j = 0; j < 9; ++j) h9[j] = __uint_as_float(v[j]) + __ldg(bias_p + j);
            if (valid) { x[0] = p.xyz[m * 3 + 0]; x[1] = p.xyz[m * 3 + 1]; x[2] = p.xyz[m * 3 + 2]; }
            if (p.use_quat) {
              nof_quat_apply(h9, x, o);
            } else {
              o[0] = h9[0] + x[0]; o[1] = h9[1] + x[1]; o[2] = h9[2] + x[2];
            }
            if (valid) {
              p.out[m * 3 + 0] = o[0]; p.out[m * 3 + 1] = o[1]; p.out[m * 3 + 2] = o[2];
              if (kSave && p.head_save) {
                float4* hp = reinterpret_cast<float4*>(p.head_save + m * 12);
                hp[0] = make_float4(h9[0], h9[1], h9[2], h9[3]);
                hp[1] = make_float4(h9[4], h9[5], h9[6], h9[7]);
                hp[2] = make_float4(h9[8], x[0], x[1], x[2]);
              }
            }
          }
        } else if (mask_round) {
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int c0 = c_lo + 32 * i;
            uint32_t v[32];
            tmem_ld32(t_acc + c0, v);
            const uint32_t word = mw[i];
            tmem_ld_wait();
            uint32_t w16[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const uint32_t lo = ((word >> (2 * j)) & 1u) ? v[2 * j] : 0u;
              const uint32_t hi = ((word >> (2 * j + 1)) & 1u) ? v[2 * j + 1] : 0u;
              w16[j] = cvt_bf16x2_bits(lo, hi);
            }
            tmem_st16(t_row + kColH + (uint32_t)(c0 >> 1), w16);
            if (saving && rd.save_off != kNofNone)
              stage32(kDirectSave ? save_tile + rd.save_off : stage, row, (uint32_t)c0, w16);
          }
        } else if (kBwd && rd.epi == MCF_EPI_B_DPE) {
          // d_xyz += J_PE(x)^T dPE, with sin/cos taken from the saved first-layer operand image (half 0 owns dx)
          if (half == 0) {
            const uint8_t* x0img = reinterpret_cast<const uint8_t*>(p.fwd_save) + tile * p.fwd_save_tile_bytes + p.fwd_x0_off;
            uint32_t pw[32];   // the row's 64 saved bf16 channels, two per word (channel c: word c >> 1, even c low)
#pragma unroll
            for (int c8 = 0; c8 < 8; ++c8) {
              const uint4 t = *reinterpret_cast<const uint4*>(x0img + sw128_off(row, c8));
              pw[c8 * 4 + 0] = t.x; pw[c8 * 4 + 1] = t.y; pw[c8 * 4 + 2] = t.z; pw[c8 * 4 + 3] = t.w;
            }
#define MCF_PE_AT(c_) (((c_) & 1) ? bf16_hi(pw[(c_) >> 1]) : bf16_lo(pw[(c_) >> 1]))
            // d/dx [w sin(f x)] = f (w cos(f x)),  d/dx [w cos(f x)] = -f (w sin(f x)):
            //   dx_c += f_k (pe[cos] dPE[sin] - pe[sin] dPE[cos]),  sin channel 3 + 6k + c, cos channel +3.
            // The 64 dPE columns are read in two halves; only the pair (29, 32) straddles them: its sine-side
            // gradient is carried over in `carry`.
            float carry = 0.f;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              uint32_t v[32];
              tmem_ld32(t_acc + 32 * i, v);
              tmem_ld_wait();
              if (i == 0) {
#pragma unroll
                for (int c = 0; c < 3; ++c) dx[c] += __uint_as_float(v[c]);
              }
#pragma unroll
              for (int k = 0; k < 10; ++k) {
                if (k < p.pe_n_freqs) {
                  const float f = tab.pe_freq[k];
#pragma unroll
                  for (int c = 0; c < 3; ++c) {
                    const int is = 3 + 6 * k + c, ic = is + 3;   // compile-time
                    if ((is >> 5) == i && (ic >> 5) == i) {
                      dx[c] += f * (MCF_PE_AT(ic) * __uint_as_float(v[is & 31]) - MCF_PE_AT(is) * __uint_as_float(v[ic & 31]));
                    } else if ((is >> 5) == 0 && (ic >> 5) == 1) {
                      if (i == 0) carry = __uint_as_float(v[is & 31]);
                      else dx[c] += f * (MCF_PE_AT(ic) * carry - MCF_PE_AT(is) * __uint_as_float(v[ic & 31]));
                    }
                  }
                }
              }
            }
#undef MCF_PE_AT
            if (rd.aux_off == 1u && valid && p.d_xyz) {
              p.d_xyz[m * 3 + 0] = dx[0]; p.d_xyz[m * 3 + 1] = dx[1]; p.d_xyz[m * 3 + 2] = dx[2];
            }
          }
        } else {
          if (gtid == 0) atomicExch(&g_mcf_device_error, 0xBADE0000u | (uint32_t)rd.epi);
        }

        // signal the tensor core first (it needs the TMEM operand only), then push the staged image to HBM
        if (r + 1 < p.n_rounds) arrive();
        else tc_fence_before();
        if (!kDirectSave && writes_h && saving && rd.save_off != kNofNone) {
          if (kWarpStore && rd.n_out == 128) warp_store(save_tile + rd.save_off, (uint32_t)half);
          else stage_store(save_tile + rd.save_off, ((uint32_t)rd.n_out + 63u) / 64u * kBlkN);
        }
      }
    }
    if (lane == 0) bulk_wait_all();
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

int launch_nof_ts(const mcf_chain_params_t& p, cudaStream_t stream) {
  if (p.width != 128 || p.program_kind != 1 || p.wpack_bytes == 0 || p.wpack_bytes > kNofResBytes || (p.wpack_bytes & 15u))
    return MCF_ERR_BAD_ARG;
  if (p.n_chunks > kNofMaxChunks || p.n_rounds > kNofMaxRounds) return MCF_ERR_BAD_ARG;
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    e = cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return (int)e;
  }
  const long long n_tiles = (p.n_rows + MCF_TILE_ROWS - 1) / MCF_TILE_ROWS;
  const long long n_units = (n_tiles + 1) / 2;
  int cap = p.max_ctas > 0 ? p.max_ctas : n_sm;
  const int grid = (int)(n_units < cap ? n_units : cap);
  const bool bwd = p.prologue == MCF_PRO_B_NOF;
  const bool save = bwd || p.save != nullptr || p.masks != nullptr || p.head_save != nullptr;
  const void* fn = bwd ? (const void*)k_nof<true, true> : (save ? (const void*)k_nof<false, true> : (const void*)k_nof<false, false>);
  const int smem = save ? (int)NofSmem<true>::total : (int)NofSmem<false>::total;
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return (int)e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kNofThreads);
  cfg.dynamicSmemBytes = (size_t)smem;
  cfg.stream = stream;
  void* args[1] = {const_cast<mcf_chain_params_t*>(&p)};
  e = cudaLaunchKernelExC(&cfg, fn, args);
  if (e != cudaSuccess) return (int)e;
  e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

}  // namespace mcf
