// Fused MLP chains on tcgen05 / TMEM (sm_100a only).
//
// One persistent CTA per SM works on two 128-sample tiles ("slots") at a time.  Warp roles:
//   warp 0      : weight producer -- streams pre-swizzled bf16 weight chunk images (16 KB each) from
//                 L2/HBM into a 4-stage shared-memory ring with 1-D bulk async copies (UBLKCP)
//   warp 1      : MMA issuer -- one thread issues tcgen05.mma (M=128, N<=128 per chunk, K=16 steps);
//                 accumulators live in TMEM (256 columns per slot)
//   warp 2      : TMEM allocator
//   warps 4-7   : epilogue group of slot 0, warps 8-11: epilogue group of slot 1 (thread == tile row):
//                 build the first-layer operand (positional encoding fused here), then per layer
//                 TMEM -> registers -> bias/activation -> bf16 -> 128B-swizzled smem operand of the
//                 next layer, so activations never leave the SM.  The two slots ping-pong: while the
//                 tensor core runs layer l of slot B, slot A's epilogue of layer l is in flight.
// The layer program (which chunks feed which accumulator columns, which epilogue follows) is a table
// built by the host shim, so the same kernel runs NeRF / NoF forward and their backward dX chains.
//
// Width 256 runs on CTA pairs (template C == 2, cluster of two, tcgen05 cta_group::2): the leader CTA's MMA thread
// issues M=256 instructions over both CTAs' tiles, each CTA streams only its 128-row half of every [256 x 64] weight
// tile, the peer relays its ring's "full" barriers to the leader, the leader's commits are multicast to both CTAs.
// Cross-CTA arrivals are relaxed on purpose (they carry control only).
//
// Reference semantics: models/nerf.py:61-102, models/nof.py:55-85, models/embedding.py:42-46.
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "../../include/moco_flow_b200.h"
#include "ptx.cuh"
#include "nof_math.cuh"

namespace mcf {

constexpr int kThreads = 384;
constexpr int kMaxStages = 4;
constexpr int kProducers = 1;  // producer warps (0, then 2, 3).  Three were measured against one in both ring modes: no gain
constexpr uint32_t kBlk = MCF_BLOCK_BYTES;
constexpr int kMaxChunks = 128;
constexpr int kMaxRounds = 24;
constexpr uint32_t kNone = 0xFFFFFFFFu;
constexpr uint32_t kSlotCols = 256;
// training saves: 1 = every epilogue warp stores its own rows of the operand image, 0 = one store per image by the
// epilogue group's first thread behind a group barrier
#ifndef MCF_CHAIN_WARP_STORE
#define MCF_CHAIN_WARP_STORE 1
#endif
constexpr bool kWarpStore = MCF_CHAIN_WARP_STORE != 0;

struct Tables {
  mcf_chunk_t chunks[kMaxChunks];  // 2048 B
  mcf_round_t rounds[kMaxRounds];  // 768 B
  uint64_t w_full[kMaxStages];
  uint64_t w_empty[kMaxStages];
  uint64_t w_peer[kMaxStages];  // CTA pairs: the peer's half of a ring stage has landed (leader CTA only)
  uint64_t act_ready[2];
  uint64_t acc_full[2];
  uint64_t w_res;               // resident variant: the weight stream has landed
  uint32_t tmem_base;
  uint32_t pad[1];
  float pe_freq[12];    // encoder tables: from the launch parameters, or from p.pe_table (device) when given
  float pe_weight[12];
};

template <int W>
struct Smem {
  static constexpr int kStages = 4;   // 16 KB weight-ring stages.  8 at W=128 measured slower: the larger carve-out shrinks L1
  static constexpr uint32_t kHBlocks = W / 64;
  static constexpr uint32_t kHBytes = kHBlocks * kBlk;
  static constexpr uint32_t off_h = 0;
  static constexpr uint32_t off_x0 = off_h + 2 * kHBytes;
  static constexpr uint32_t off_ring = off_x0 + 2 * kBlk;
  static constexpr uint32_t off_tab = off_ring + kStages * kBlk;
  static constexpr uint32_t total = off_tab + sizeof(Tables);
};
static_assert(Smem<256>::total <= 232448 && Smem<128>::total <= 232448, "shared memory budget exceeded");

// Resident-weight variant (NoF, W = 128): the whole packed weight stream of the program (<= 144 KB) is copied into
// shared memory once per CTA; there is no ring, no producer and no per-tile weight traffic.  The first-layer operand
// (x0) lives in block 0 of the slot's activation buffer (the round-0 epilogue overwrites it once the tensor core has
// consumed it), which is what makes two slots + the weights fit.
constexpr uint32_t kResBytes = 147456;
template <int W>
struct SmemRes {
  static constexpr int kStages = 1;
  static constexpr uint32_t kHBlocks = W / 64;
  static constexpr uint32_t kHBytes = kHBlocks * kBlk;
  static constexpr uint32_t off_h = 0;
  static constexpr uint32_t off_x0 = 0;                       // aliased: slot s -> off_h + s * kHBytes
  static constexpr uint32_t off_ring = off_h + 2 * kHBytes;   // the resident weights
  static constexpr uint32_t off_tab = off_ring + kResBytes;
  static constexpr uint32_t total = off_tab + sizeof(Tables);
};
static_assert(SmemRes<128>::total <= 232448, "shared memory budget exceeded");

// ---------------------------------------------------------------------------------------------
// epilogue helpers (thread == row)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void store_h8(uint8_t* hbuf, uint32_t row, uint32_t col0, uint4 v) {
  uint32_t block = col0 >> 6, c16 = (col0 & 63u) >> 3;
#ifdef MCF_EXP_NOSTS   // timing experiment only (wrong results): drop the activation stores unless a value is "magic"
  if (v.x != 0x7fc12345u) return;
#endif
  *reinterpret_cast<uint4*>(hbuf + block * kBlk + sw128_off(row, c16)) = v;
}

// pack 32 fp32 -> 32 bf16 and store as 4 x 16 B into the swizzled activation buffer
__device__ __forceinline__ void store_h32(uint8_t* hbuf, uint32_t row, uint32_t col0, const float (&f)[32]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint4 v;
    v.x = pack_bf16x2(f[q * 8 + 0], f[q * 8 + 1]);
    v.y = pack_bf16x2(f[q * 8 + 2], f[q * 8 + 3]);
    v.z = pack_bf16x2(f[q * 8 + 4], f[q * 8 + 5]);
    v.w = pack_bf16x2(f[q * 8 + 6], f[q * 8 + 7]);
    store_h8(hbuf, row, col0 + q * 8, v);
  }
}

__device__ __forceinline__ void load32f(const float* __restrict__ p, float (&b)[32]) {
#ifdef MCF_EXP_NOBIAS
#pragma unroll
  for (int q = 0; q < 32; ++q) b[q] = 0.25f;
  return;
#endif
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    // biases / head weights are re-read by every tile: ask L1 (only ~28 KB next to the 227 KB carve-out) to keep them
    float4 t;
    asm volatile("ld.global.nc.L1::evict_last.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w)
                 : "l"(reinterpret_cast<const float4*>(p) + q));
    b[q * 4 + 0] = t.x;
    b[q * 4 + 1] = t.y;
    b[q * 4 + 2] = t.z;
    b[q * 4 + 3] = t.w;
  }
}

// Bias (+ReLU) epilogue for 32 accumulator columns: packed fp32x2 adds, ReLU fused into the bf16x2 convert,
// sign bits funnel-shifted into the ReLU bit mask (bit j set <=> column j is active).
template <bool kRelu, bool kMask>
__device__ __forceinline__ uint32_t bias_act_store32(uint8_t* hbuf, uint32_t row, uint32_t col0, uint32_t (&v)[32],
                                                     const float (&b)[32]) {
#pragma unroll
  for (int j = 0; j < 16; ++j) add_f32x2(v[2 * j], v[2 * j + 1], b[2 * j], b[2 * j + 1]);
  uint32_t word = 0;
  if (kMask) {
    // sign bits -> bit j of the word; four independent 8-deep funnel-shift chains instead of one 32-deep one
    uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0;
#pragma unroll
    for (int j = 7; j >= 0; --j) {
      w0 = __funnelshift_l(v[j], w0, 1);
      w1 = __funnelshift_l(v[8 + j], w1, 1);
      w2 = __funnelshift_l(v[16 + j], w2, 1);
      w3 = __funnelshift_l(v[24 + j], w3, 1);
    }
    word = ~(w0 | (w1 << 8) | (w2 << 16) | (w3 << 24));
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint4 o;
    if (kRelu) {
      o.x = cvt_bf16x2_relu_bits(v[q * 8 + 0], v[q * 8 + 1]);
      o.y = cvt_bf16x2_relu_bits(v[q * 8 + 2], v[q * 8 + 3]);
      o.z = cvt_bf16x2_relu_bits(v[q * 8 + 4], v[q * 8 + 5]);
      o.w = cvt_bf16x2_relu_bits(v[q * 8 + 6], v[q * 8 + 7]);
    } else {
      o.x = cvt_bf16x2_bits(v[q * 8 + 0], v[q * 8 + 1]);
      o.y = cvt_bf16x2_bits(v[q * 8 + 2], v[q * 8 + 3]);
      o.z = cvt_bf16x2_bits(v[q * 8 + 4], v[q * 8 + 5]);
      o.w = cvt_bf16x2_bits(v[q * 8 + 6], v[q * 8 + 7]);
    }
    store_h8(hbuf, row, col0 + q * 8, o);
  }
  return word;
}

struct RowState {   // per-thread state that lives across the rounds of one tile
  float sigma;      // fwd: sigma head value; bwd: d_sigma
  float dx[3];      // bwd: accumulated d_xyz
  float aux[4];
};

// ---------------------------------------------------------------------------------------------
// the chain kernel
// ---------------------------------------------------------------------------------------------
#define MCF_T0(var) long long var = timing ? clock64() : 0
#define MCF_TACC(slot, var) \
  do { if (timing) { long long _n = clock64(); tacc[slot] += (unsigned long long)(_n - var); var = _n; } } while (0)

// C = CTAs per cluster.  C == 2: the two CTAs of a pair run tcgen05 cta_group::2 -- the leader's MMA thread issues
// M=256 instructions over both CTAs' tiles, every CTA streams only its half of each weight chunk (the N split of
// the B operand), which halves the L2 -> shared-memory weight traffic per tile and doubles the ring's depth in time.
// kBwd selects the program family the instantiation can run (forward: PE / dense prologue, bias-ReLU / head
// epilogues; backward: head-gradient prologues, mask / PE-Jacobian epilogues): each kernel carries half the code.
// kSave: the launch writes training saves (operand images, ReLU masks, head values); inference instantiations carry
// none of that code.  The in-kernel cycle counters exist only in builds with -DMCF_TIMING (scripts/chain_timing.py).
// kNoF: NoF program (flow head, quaternion transform) vs NeRF program (sigma / rgb heads).
// kRes: resident-weight variant (SmemRes): W = 128, C = 1 only.
template <int W, int C, bool kBwd, bool kSave, bool kNoF, bool kRes = false>
__global__ void __launch_bounds__(kThreads, 1) k_chain(const __grid_constant__ mcf_chain_params_t p) {
  static_assert(!kRes || (W == 128 && C == 1), "resident weights: W = 128, single CTA");
  extern __shared__ __align__(1024) uint8_t smem[];
#ifdef MCF_TIMING
  const bool timing = p.timing != nullptr;
#else
  constexpr bool timing = false;
#endif
  unsigned long long tacc[4] = {0ull, 0ull, 0ull, 0ull};
  const long long t_kernel0 = timing ? clock64() : 0;
  using L = typename std::conditional<kRes, SmemRes<W>, Smem<W>>::type;
  constexpr int kStages = L::kStages;
  Tables& tab = *reinterpret_cast<Tables*>(smem + L::off_tab);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = (C == 2) ? cluster_ctarank() : 0u;

  if ((smem_u32(smem) & 1023u) != 0u) {
    if (threadIdx.x == 0) atomicExch(&g_mcf_device_error, 0xA11C0000u);
    return;
  }

  // ---- one-time setup ----
  {
    const uint32_t* src_c = reinterpret_cast<const uint32_t*>(p.chunks);
    uint32_t* dst_c = reinterpret_cast<uint32_t*>(tab.chunks);
    for (int i = threadIdx.x; i < p.n_chunks * 4; i += kThreads) dst_c[i] = src_c[i];
    const uint32_t* src_r = reinterpret_cast<const uint32_t*>(p.rounds);
    uint32_t* dst_r = reinterpret_cast<uint32_t*>(tab.rounds);
    for (int i = threadIdx.x; i < p.n_rounds * 8; i += kThreads) dst_r[i] = src_r[i];
    if (threadIdx.x < 10) {
      const int k = threadIdx.x;
      tab.pe_freq[k] = p.pe_table ? p.pe_table[k] : p.pe_freq[k];
      tab.pe_weight[k] = p.pe_table ? p.pe_table[MCF_MAX_FREQS + k] : p.pe_weight[k];
    }
    if (!kRes) {
      uint4* x0z = reinterpret_cast<uint4*>(smem + L::off_x0);
      for (int i = threadIdx.x; i < (int)(2 * kBlk / 16); i += kThreads) x0z[i] = make_uint4(0, 0, 0, 0);
    }
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&tab.w_full[s], 1);
      mbar_init(&tab.w_empty[s], 1);
      mbar_init(&tab.w_peer[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tab.act_ready[s], C == 2 ? 8 : 128);  // pairs: one arrival per epilogue warp of both CTAs, on the leader's barrier
      mbar_init(&tab.acc_full[s], 1);
    }
    mbar_init(&tab.w_res, 1);
    fence_mbar_init();
    if (kRes) {   // the whole weight stream, once (32 KB pieces; the mbarrier counts the bytes)
      const uint32_t total = p.wpack_bytes;
      mbar_arrive_expect_tx(&tab.w_res, total);
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.wpack);
      for (uint32_t off = 0; off < total; off += 32768u) {
        const uint32_t n = total - off < 32768u ? total - off : 32768u;
        bulk_g2s(smem + L::off_ring + off, wsrc + off, n, &tab.w_res);
      }
    }
  }
  if (warp == 2) {
    if (C == 2) {
      tmem_alloc_pair(&tab.tmem_base, 512);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(&tab.tmem_base, 512);
      tmem_relinquish();
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  if (C == 2) cluster_sync_all();  // the peer's barriers are initialised before anything arrives on them
  tc_fence_after();
  const uint32_t tmem_base = tab.tmem_base;

  // work units: a unit is 2 slots x C CTAs = 2C consecutive tiles; tile(u, s) = (2u + s) * C + cta_rank.
  // C == 1 skips a slot whose tile does not exist; a pair runs missing tiles as all-invalid rows so that both
  // CTAs walk identical weight / barrier sequences.
  const long long n_tiles = (p.n_rows + MCF_TILE_ROWS - 1) / MCF_TILE_ROWS;
  const long long n_pairs = (n_tiles + 2 * C - 1) / (2 * C);
  const long long unit0 = blockIdx.x / C, unit_step = gridDim.x / C;
  auto tile_of = [&](long long u, int s) -> long long { return (2 * u + s) * C + cta_rank; };
  auto wait_x = [&](uint64_t* bar, uint32_t parity, uint32_t tag) {
    // cta-scope acquire also for barriers signalled from the peer CTA (as CUTLASS' ClusterBarrier::wait does): what
    // they guard is either async-proxy data (weights) or this SM's own shared memory / TMEM
    mbar_wait(bar, parity, tag);
  };

  // register re-distribution between the warpgroups: the producer/MMA/allocator warpgroup needs few registers,
  // the two epilogue warpgroups hold 64 accumulator + 64 bias values in flight (128*72 + 256*208 <= 64K)
  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 72;\n");
  if (warp != 1) {
    // =========================== weight producers (warps 0, 2, 3) ===========================
    // Every producer walks the whole copy sequence (to track stage / phase) and issues every kProducers-th copy.
    const int me = (warp == 0) ? 0 : warp - 1;
    if (!kRes && lane == 0 && me < kProducers) {
      int turn = 0;
      uint32_t stage = 0, phase = 0;
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.wpack);
      for (long long pair = unit0; pair < n_pairs; pair += unit_step) {
        for (int r = 0; r < p.n_rounds; ++r) {
          const int cb = tab.rounds[r].chunk_begin, ce = tab.rounds[r].chunk_end;
          for (int s = 0; s < 2; ++s) {
            if (C == 1 && tile_of(pair, s) >= n_tiles) continue;
            for (int c = cb; c < ce; ++c) {
              uint32_t bytes = tab.chunks[c].bytes, src_off = tab.chunks[c].src_off;
              bool merge = false;
              if (C == 2) {
                const uint32_t fl = tab.chunks[c].flags;
                if (fl & 2u) {   // [256 x 64] weight tile: this CTA's 128-row half is one chunk
                  bytes = tab.chunks[c + cta_rank].bytes;
                  src_off = tab.chunks[c + cta_rank].src_off;
                  ++c;
                  // flag bit 3: this CTA's halves of this k-block and of the next one are adjacent in the packed
                  // stream -> ONE 32 KB copy into two consecutive ring stages (not across the ring's wrap).  A bulk-copy issue costs the
                  // thread ~330 clk whatever its size: 16 KB copies sustain 24 B/clk, a pair tile needs 32 B/clk.
                  merge = (fl & 8u) != 0u && stage != (uint32_t)(kStages - 1);
                  if (merge) { bytes += kBlk; c += 2; }
                } else {                          // single chunk: this CTA's half of its rows
                  bytes >>= 1;
                  src_off += cta_rank * bytes;
                }
              }
              if (turn == me) {
                MCF_T0(tw);
                wait_x(&tab.w_empty[stage], phase ^ 1u, 0x100u | stage);
                if (merge) wait_x(&tab.w_empty[stage + 1], phase ^ 1u, 0x100u | (stage + 1));
                MCF_TACC(0, tw);
                mbar_arrive_expect_tx(&tab.w_full[stage], bytes);
                bulk_g2s(smem + L::off_ring + stage * kBlk, wsrc + src_off, bytes, &tab.w_full[stage]);
              }
              if (++turn == kProducers) turn = 0;
              if (++stage == kStages) { stage = 0; phase ^= 1u; }
              if (merge) { if (++stage == kStages) { stage = 0; phase ^= 1u; } }
            }
          }
        }
      }
      if (timing && me == 0) p.timing[blockIdx.x * 16 + 11] = tacc[0];
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (lane == 0 && C == 2 && cta_rank != 0u) {
      // ---- peer CTA of a pair: relay "my half of the stage has landed" to the leader's MMA thread ----
      uint32_t stage = 0, fph = 0;   // fph: per-stage phase bits of w_full (a merged copy skips the odd stage's barrier)
      uint32_t peer_bar[kStages];
      for (int i = 0; i < kStages; ++i) peer_bar[i] = map_to_cta(smem_u32(&tab.w_peer[i]), 0u);
      for (long long pair = unit0; pair < n_pairs; pair += unit_step) {
        for (int r = 0; r < p.n_rounds; ++r) {
          const int cb = tab.rounds[r].chunk_begin, ce = tab.rounds[r].chunk_end;
          for (int s = 0; s < 2; ++s) {
            for (int c = cb; c < ce; ++c) {
              const uint32_t fl = tab.chunks[c].flags;
              bool merge = false;
              if (fl & 2u) {
                ++c;
                merge = (fl & 8u) != 0u && stage != (uint32_t)(kStages - 1);
                if (merge) c += 2;
              }
              mbar_wait(&tab.w_full[stage], (fph >> stage) & 1u, 0x500u | stage);
              fph ^= 1u << stage;
              mbar_arrive_cluster_relaxed(peer_bar[stage]);
              stage = (stage + (merge ? 2u : 1u)) & (kStages - 1);
            }
          }
        }
      }
    } else if (lane == 0) {
      uint32_t stage = 0, fph = 0, pph = 0;   // per-stage phase bits of w_full / w_peer
      uint32_t ar_phase[2] = {0u, 0u};
      const uint32_t h_addr = smem_u32(smem + L::off_h), x0_addr = smem_u32(smem + L::off_x0);
      const uint32_t ring_addr = smem_u32(smem + L::off_ring);
      if (kRes) {
        mbar_wait(&tab.w_res, 0u, 0x700u);
        tc_fence_after();
      }
      for (long long pair = unit0; pair < n_pairs; pair += unit_step) {
        for (int r = 0; r < p.n_rounds; ++r) {
          const int cb = tab.rounds[r].chunk_begin, ce = tab.rounds[r].chunk_end;
          for (int s = 0; s < 2; ++s) {
            if (C == 1 && tile_of(pair, s) >= n_tiles) continue;
            MCF_T0(tm);
            wait_x(&tab.act_ready[s], ar_phase[s], 0x200u | s);
            ar_phase[s] ^= 1u;
            tc_fence_after();
            MCF_TACC(0, tm);
            for (int c = cb; c < ce; ++c) {
              if (!kRes) {
                mbar_wait(&tab.w_full[stage], (fph >> stage) & 1u, 0x300u | stage);
                fph ^= 1u << stage;
              }
              const mcf_chunk_t ch = tab.chunks[c];
              // flag bit1: this chunk and the next one are the two 128-row halves of one [256 x 64] weight tile
              // sitting in consecutive (even, odd) ring stages -> one N=256 instruction per K step
              const bool fuse = (ch.flags & 2u) != 0u;   // resident variant: the two images are adjacent in the stream
              // CTA pairs: one 32 KB copy per CTA may carry this k-block and the next one (see the producer)
              const bool merge = C == 2 && fuse && (ch.flags & 8u) != 0u && stage != (uint32_t)(kStages - 1);
              if (C == 2) {
                mbar_wait(&tab.w_peer[stage], (pph >> stage) & 1u, 0x600u | stage);
                pph ^= 1u << stage;
              } else if (fuse && !kRes) {
                mbar_wait(&tab.w_full[stage + 1], (fph >> (stage + 1)) & 1u, 0x300u | (stage + 1));
                fph ^= 1u << (stage + 1);
              }
              if (!kRes) tc_fence_after();
              MCF_TACC(1, tm);
              // resident variant: x0 is block 0 of the slot's activation buffer, weights sit at their stream offset
              const uint32_t a_base = ((ch.a_buf || kRes) ? (h_addr + s * L::kHBytes) : (x0_addr + s * kBlk)) + ch.a_kblock * kBlk;
              const uint32_t b_base = kRes ? ring_addr + ch.src_off : ring_addr + stage * kBlk;
              const uint32_t idesc = make_idesc(fuse ? 2u * ch.n : (uint32_t)ch.n, false, false, 128u * C);
              const uint32_t d_tmem = tmem_base + s * kSlotCols + ch.acc_col;
              // one thread feeds the tensor core: keep the per-instruction work to two 64-bit adds (a K step of 16
              // bf16 = 32 B moves the 16-byte-unit start-address field of both descriptors by 2)
              uint64_t ad = make_sdesc(a_base, 0u, 1024u), bd = make_sdesc(b_base, 0u, 1024u);
              uint32_t acc = (ch.flags & 1u) ? 0u : 1u;
              if (ch.ksteps == 4) {
#pragma unroll
                for (uint32_t k = 0; k < 4; ++k) {
                  if (C == 2) umma_bf16_pair(d_tmem, ad + 2u * k, bd + 2u * k, idesc, k ? 1u : acc);
                  else umma_bf16(d_tmem, ad + 2u * k, bd + 2u * k, idesc, k ? 1u : acc);
                }
              } else {
                for (uint32_t k = 0; k < ch.ksteps; ++k) {
                  if (C == 2) umma_bf16_pair(d_tmem, ad, bd, idesc, acc);
                  else umma_bf16(d_tmem, ad, bd, idesc, acc);
                  ad += 2u; bd += 2u; acc = 1u;
                }
              }
              if (kRes) {
                if (fuse) ++c;   // nothing to release
              } else if (C == 2) {
                umma_commit_pair(&tab.w_empty[stage], 3);   // frees the stage in both CTAs
                stage = (stage + 1u) & (kStages - 1);
                if (fuse) ++c;
                if (merge) {   // the next k-block of the same layer sits in the following (odd) stage
                  const mcf_chunk_t c2 = tab.chunks[c + 1];
                  const uint32_t a2 = ((c2.a_buf) ? (h_addr + s * L::kHBytes) : (x0_addr + s * kBlk)) + c2.a_kblock * kBlk;
                  const uint64_t ad2 = make_sdesc(a2, 0u, 1024u), bd2 = make_sdesc(ring_addr + stage * kBlk, 0u, 1024u);
                  const uint32_t d2 = tmem_base + s * kSlotCols + c2.acc_col;
#pragma unroll
                  for (uint32_t k = 0; k < 4; ++k)
                    umma_bf16_pair(d2, ad2 + 2u * k, bd2 + 2u * k, idesc, (k || !(c2.flags & 1u)) ? 1u : 0u);
                  umma_commit_pair(&tab.w_empty[stage], 3);
                  stage = (stage + 1u) & (kStages - 1);
                  c += 2;
                }
              } else {
                umma_commit(&tab.w_empty[stage]);
                stage = (stage + 1u) & (kStages - 1);
                if (fuse) {
                  umma_commit(&tab.w_empty[stage]);
                  stage = (stage + 1u) & (kStages - 1);
                  ++c;
                }
              }
              MCF_TACC(2, tm);
            }
            if (C == 2) umma_commit_pair(&tab.acc_full[s], 3);
            else umma_commit(&tab.acc_full[s]);
          }
        }
      }
      if (timing) {
        p.timing[blockIdx.x * 16 + 8] = tacc[0];
        p.timing[blockIdx.x * 16 + 9] = tacc[1];
        p.timing[blockIdx.x * 16 + 10] = tacc[2];
      }
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 208;\n");
    // =========================== epilogue groups ===========================
    const int s = (warp - 4) >> 2;          // slot
    const int qtr = warp & 3;               // TMEM lane quarter this warp may access
    const uint32_t row = qtr * 32 + lane;   // tile row owned by this thread
    const int gtid = threadIdx.x - 128 - s * 128;
    uint8_t* hbuf = smem + L::off_h + s * L::kHBytes;
    uint8_t* x0buf = kRes ? hbuf : smem + L::off_x0 + s * kBlk;
    const uint32_t t_row = tmem_base + ((uint32_t)(qtr * 32) << 16) + s * kSlotCols;
    const uint32_t act_ready_leader = (C == 2) ? map_to_cta(smem_u32(&tab.act_ready[s]), 0u) : 0u;
    auto arrive_act_ready = [&]() {
      if (C == 2) {
        // Every lane has already executed fence.proxy.async after its st.shared (the operand is read by this SM's own
        // tensor core), so the remote arrival only carries control: a release at cluster scope here also drains the
        // warp's global stores and was measured at ~1300 clk per round.
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster_relaxed(act_ready_leader);
      } else {
        mbar_arrive(&tab.act_ready[s]);
      }
    };
    uint32_t af_phase = 0;
    bool store_pending = false;
    // Training saves.  A thread owns a tile row, so a warp's 32 rows of every 64-column block are 4 KB contiguous in
    // the swizzled image: with kWarpStore each warp pushes its own pieces to HBM (lane 0 issues and later waits for
    // them) and the epilogue group needs no barrier around the stores; otherwise the group's first thread stores
    // the whole image after a group barrier.
    auto store_image = [&](uint8_t* dst, const uint8_t* src, uint32_t n_blocks) {
      if (kWarpStore) {
        __syncwarp();   // every lane has executed fence.proxy.async after its st.shared
        if (lane == 0) {
          for (uint32_t b = 0; b < n_blocks; ++b) {
            const uint32_t off = b * kBlk + (uint32_t)qtr * 4096u;
            bulk_s2g(dst + off, src + off, 4096u);
          }
          bulk_commit();
        }
      } else {
        named_bar_sync(1 + s, 128);
        if (gtid == 0) {
          bulk_s2g(dst, src, n_blocks * kBlk);
          bulk_commit();
        }
      }
      store_pending = true;
    };
    auto store_read_done = [&]() {   // an earlier store no longer reads the buffers about to be overwritten
      if (store_pending) {
        if (kWarpStore) {
          if (lane == 0) bulk_wait_read_all();
          __syncwarp();
        } else {
          if (gtid == 0) bulk_wait_read_all();
          named_bar_sync(1 + s, 128);
        }
        store_pending = false;
      }
    };

    for (long long pair = unit0; pair < n_pairs; pair += unit_step) {
      const long long tile = tile_of(pair, s);
      if (C == 1 && tile >= n_tiles) break;
      const bool tile_ok = tile < n_tiles;                     // false: a pair's padding tile (no loads/stores)
      const long long tile_r = tile_ok ? tile : n_tiles - 1;   // tile index for reads
      const bool saving = kSave && p.save != nullptr && tile_ok;
      const long long m = tile * MCF_TILE_ROWS + row;
      const bool valid = m < p.n_rows;
      const long long mc = valid ? m : (p.n_rows - 1);
      const long long ray = mc / p.rows_per_ray;
      uint8_t* save_tile = saving ? reinterpret_cast<uint8_t*>(p.save) + tile * p.save_tile_bytes : nullptr;
      RowState st;
      st.sigma = 0.f; st.dx[0] = st.dx[1] = st.dx[2] = 0.f;
      st.aux[0] = st.aux[1] = st.aux[2] = st.aux[3] = 0.f;

      MCF_T0(te);
      // make sure an earlier bulk store no longer reads the buffers we are about to overwrite
      store_read_done();

      // ------------------------- prologue: build the first operand -------------------------
      if (!kBwd && p.prologue == MCF_PRO_PE_XYZ) {
        float x[3] = {0.f, 0.f, 0.f};
        if (valid) { x[0] = p.xyz[m * 3 + 0]; x[1] = p.xyz[m * 3 + 1]; x[2] = p.xyz[m * 3 + 2]; }
        // channel order of models/embedding.py:42-46: [x | w0 sin(f0 x) | w0 cos(f0 x) | w1 sin(f1 x) | ...]
        float ch[64];
#pragma unroll
        for (int c = 0; c < 64; ++c) ch[c] = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) ch[c] = x[c];
        float sn[3] = {0.f, 0.f, 0.f}, cs[3] = {1.f, 1.f, 1.f};
#pragma unroll
        for (int k = 0; k < 10; ++k) {
          if (k < p.pe_n_freqs) {
            const float f = tab.pe_freq[k], w = tab.pe_weight[k];
            // full-range sincosf at every 4th octave; in between sin/cos(2a) from sin/cos(a) (<= 3 doublings,
            // error <= ~8 ulp, far below the bf16 rounding of the operand); non-octave tables stay exact
            const bool exact = (k & 3) == 0 || f != 2.0f * tab.pe_freq[k > 0 ? k - 1 : 0];
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
              ch[3 + 6 * k + c] = w * sn[c];
              ch[3 + 6 * k + 3 + c] = w * cs[c];
            }
          }
        }
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) {
          uint4 v;
          v.x = pack_bf16x2(ch[c8 * 8 + 0], ch[c8 * 8 + 1]); v.y = pack_bf16x2(ch[c8 * 8 + 2], ch[c8 * 8 + 3]);
          v.z = pack_bf16x2(ch[c8 * 8 + 4], ch[c8 * 8 + 5]); v.w = pack_bf16x2(ch[c8 * 8 + 6], ch[c8 * 8 + 7]);
          *reinterpret_cast<uint4*>(x0buf + sw128_off(row, c8)) = v;
        }
      } else if (!kBwd && p.prologue == MCF_PRO_DENSE) {
        const float* src = p.dense + mc * p.dense_stride;
        for (int c8 = 0; c8 < 8; ++c8) {
          float f[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            int c = c8 * 8 + j;
            f[j] = (valid && c < p.dense_cols) ? src[c] : 0.f;
          }
          uint4 v;
          v.x = pack_bf16x2(f[0], f[1]); v.y = pack_bf16x2(f[2], f[3]);
          v.z = pack_bf16x2(f[4], f[5]); v.w = pack_bf16x2(f[6], f[7]);
          *reinterpret_cast<uint4*>(x0buf + sw128_off(row, c8)) = v;
        }
      } else if (kBwd && !kNoF && p.prologue == MCF_PRO_B_NERF) {
        // backward through rgb = sigmoid(W_rgb he + b) and he = relu(.)   (models/nerf.py:98-99)
        float4 g = valid ? *reinterpret_cast<const float4*>(p.g_out + m * 4) : make_float4(0, 0, 0, 0);
        float4 o = valid ? *reinterpret_cast<const float4*>(p.fwd_out + m * 4) : make_float4(0, 0, 0, 0);
        float d0 = g.x * o.x * (1.f - o.x), d1 = g.y * o.y * (1.f - o.y), d2 = g.z * o.z * (1.f - o.z);
        st.sigma = g.w;
        if (valid && p.d_head) *reinterpret_cast<float4*>(p.d_head + m * 4) = make_float4(d0, d1, d2, g.w);
        if (saving && p.dhead_save_off != kNone) {
          uint8_t* blk = save_tile + p.dhead_save_off;
          uint4 v = make_uint4(pack_bf16x2(d0, d1), pack_bf16x2(d2, g.w), 0u, 0u);
          *reinterpret_cast<uint4*>(blk + sw128_off(row, 0)) = v;
#pragma unroll
          for (int c8 = 1; c8 < 8; ++c8) *reinterpret_cast<uint4*>(blk + sw128_off(row, c8)) = make_uint4(0, 0, 0, 0);
        }
        const mcf_round_t& r0 = tab.rounds[0];
        const int nhe = W / 2;
        const float* wrgb = p.consts + r0.aux_off;  // [3][W/2]
        const uint32_t* mk = p.fwd_masks + tile_r * p.fwd_mask_tile_words + r0.mask_off;
        for (int c0 = 0; c0 < nhe; c0 += 32) {
          float w0[32], w1[32], w2[32], f[32];
          load32f(wrgb + c0, w0);
          load32f(wrgb + nhe + c0, w1);
          load32f(wrgb + 2 * nhe + c0, w2);
          const uint32_t word = mk[(c0 >> 5) * 128 + row];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float v = d0 * w0[j] + d1 * w1[j] + d2 * w2[j];
            f[j] = ((word >> j) & 1u) ? v : 0.f;
          }
          store_h32(hbuf, row, c0, f);
        }
      } else if (kBwd && kNoF && p.prologue == MCF_PRO_B_NOF) {
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
          nof_quat_backward(hs, hs + 9, g, d16, st.dx);
        } else {  // out = head3 + x   (models/nof.py:82)
          d16[0] = g[0]; d16[1] = g[1]; d16[2] = g[2];
          st.dx[0] = g[0]; st.dx[1] = g[1]; st.dx[2] = g[2];
        }
        if (valid && p.d_head) {
          float4* dh = reinterpret_cast<float4*>(p.d_head + m * 12);
          dh[0] = make_float4(d16[0], d16[1], d16[2], d16[3]);
          dh[1] = make_float4(d16[4], d16[5], d16[6], d16[7]);
          dh[2] = make_float4(d16[8], 0.f, 0.f, 0.f);
        }
        uint4 v;
        v.x = pack_bf16x2(d16[0], d16[1]); v.y = pack_bf16x2(d16[2], d16[3]);
        v.z = pack_bf16x2(d16[4], d16[5]); v.w = pack_bf16x2(d16[6], d16[7]);
        store_h8(hbuf, row, 0, v);
        v.x = pack_bf16x2(d16[8], d16[9]); v.y = pack_bf16x2(d16[10], d16[11]);
        v.z = pack_bf16x2(d16[12], d16[13]); v.w = pack_bf16x2(d16[14], d16[15]);
        store_h8(hbuf, row, 8, v);
#pragma unroll
        for (int c8 = 2; c8 < 8; ++c8) store_h8(hbuf, row, c8 * 8, make_uint4(0, 0, 0, 0));
      } else {
        if (gtid == 0) atomicExch(&g_mcf_device_error, 0xBADF0000u | (uint32_t)p.prologue);
      }
      fence_proxy_async_smem();
      if (saving && p.x0_save_off != kNone) {
        // the prologue's operand is itself needed by the weight-gradient GEMM: store its image
        const bool fwd = !kBwd;
        const uint32_t n_blocks = fwd ? 1u : (p.prologue == MCF_PRO_B_NERF ? (uint32_t)(W / 2 / 64) : 1u);
        store_image(save_tile + p.x0_save_off, fwd ? x0buf : hbuf, n_blocks);
      }
      arrive_act_ready();
      MCF_TACC(0, te);

      // ------------------------------- rounds -------------------------------
      for (int r = 0; r < p.n_rounds; ++r) {
        const mcf_round_t rd = tab.rounds[r];
        const float* bias_p = (rd.raybias >= 0) ? (p.raybias[rd.raybias] + ray * rd.n_out) : (p.consts + rd.const_off);
        const bool is_bias_epi = !kBwd && (rd.epi == MCF_EPI_RELU || rd.epi == MCF_EPI_RELU_SIGMA || rd.epi == MCF_EPI_LINEAR);
        const bool is_mask_epi = kBwd && (rd.epi == MCF_EPI_B_MASK || rd.epi == MCF_EPI_B_MASK_SIGMA);
        // operands that do not depend on the accumulator are fetched before waiting for the tensor core
        float b0[32];
        uint32_t mwords[8];
        if (is_bias_epi) load32f(bias_p, b0);
        if (is_mask_epi && rd.mask_off != kNone) {
          const uint32_t* mkp = p.fwd_masks + tile_r * p.fwd_mask_tile_words + rd.mask_off + row;
#pragma unroll
          for (int j = 0; j < 8; ++j) mwords[j] = (j * 32 < rd.n_out) ? __ldg(mkp + j * 128) : 0u;
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) mwords[j] = 0xFFFFFFFFu;
        }
        MCF_TACC(2, te);
        wait_x(&tab.acc_full[s], af_phase, 0x400u | s);
        af_phase ^= 1u;
        tc_fence_after();
        MCF_TACC(1, te);
        const bool writes_h = rd.epi != MCF_EPI_NOF_HEAD && rd.epi != MCF_EPI_B_DPE &&
                              !(rd.epi == MCF_EPI_NERF_RGB && !saving);
        if (writes_h) store_read_done();
        const uint32_t t_acc = t_row + rd.acc_col;

        if (!kBwd && (rd.epi == MCF_EPI_RELU || rd.epi == MCF_EPI_RELU_SIGMA || rd.epi == MCF_EPI_LINEAR)) {
          const bool want_mask = kSave && p.masks != nullptr && rd.mask_off != kNone && tile_ok;
          uint32_t* mk = want_mask ? (p.masks + tile * p.mask_tile_words + rd.mask_off + row) : nullptr;
          float sig = 0.f;
          float b1[32];
          auto do_chunk = [&](int c0, uint32_t (&v)[32], const float (&b)[32]) {
            if (rd.epi == MCF_EPI_LINEAR) {
              bias_act_store32<false, false>(hbuf, row, c0, v, b);
            } else if (!kNoF && rd.epi == MCF_EPI_RELU_SIGMA) {
              float ws[32];
              load32f(p.consts + rd.aux_off + c0, ws);
#pragma unroll
              for (int j = 0; j < 32; ++j) sig = fmaf(fmaxf(__uint_as_float(v[j]) + b[j], 0.f), ws[j], sig);
              const uint32_t word = bias_act_store32<true, true>(hbuf, row, c0, v, b);
              if (want_mask) mk[(c0 >> 5) * 128] = word;
            } else if (want_mask) {
              mk[(c0 >> 5) * 128] = bias_act_store32<true, true>(hbuf, row, c0, v, b);
            } else {
              bias_act_store32<true, false>(hbuf, row, c0, v, b);
            }
          };
          // TMEM loads run one 32-column chunk ahead of the math (n_out is a multiple of 64 for these rounds)
          uint32_t va[32], vb[32];
          tmem_ld32(t_acc, va);
          for (int c0 = 0; c0 < rd.n_out; c0 += 64) {
            load32f(bias_p + c0 + 32, b1);
            tmem_ld_wait();
            tmem_ld32(t_acc + c0 + 32, vb);
            do_chunk(c0, va, b0);
            const bool more = c0 + 64 < rd.n_out;
            if (more) load32f(bias_p + c0 + 64, b0);
            tmem_ld_wait();
            if (more) tmem_ld32(t_acc + c0 + 64, va);
            do_chunk(c0 + 32, vb, b1);
          }
          if (!kNoF && rd.epi == MCF_EPI_RELU_SIGMA) {
            st.sigma = sig + __ldg(p.consts + rd.aux_off + rd.n_out);
            if (p.sigma_col == 0 && valid) p.out[m * p.out_stride] = st.sigma;  // sigma-only program
          }
        } else if (!kBwd && !kNoF && rd.epi == MCF_EPI_NERF_RGB) {
          const int nhe = rd.n_out;
          const float* wrgb = p.consts + rd.aux_off;  // [3][nhe] then b_rgb[3]
          float a0 = __ldg(wrgb + 3 * nhe + 0), a1 = __ldg(wrgb + 3 * nhe + 1), a2 = __ldg(wrgb + 3 * nhe + 2);
          for (int c0 = 0; c0 < nhe; c0 += 32) {
            uint32_t v[32];
            float f[32], b[32], w0[32];
            tmem_ld32(t_acc + c0, v);
            load32f(bias_p + c0, b);
            tmem_ld_wait();
            uint32_t word = 0;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float t = fmaxf(__uint_as_float(v[j]) + b[j], 0.f);
              word |= (t > 0.f ? 1u : 0u) << j;
              f[j] = t;
            }
            load32f(wrgb + c0, w0);
#pragma unroll
            for (int j = 0; j < 32; ++j) a0 = fmaf(f[j], w0[j], a0);
            load32f(wrgb + nhe + c0, w0);
#pragma unroll
            for (int j = 0; j < 32; ++j) a1 = fmaf(f[j], w0[j], a1);
            load32f(wrgb + 2 * nhe + c0, w0);
#pragma unroll
            for (int j = 0; j < 32; ++j) a2 = fmaf(f[j], w0[j], a2);
            if (saving) store_h32(hbuf, row, c0, f);
            if (kSave && p.masks && rd.mask_off != kNone && tile_ok)
              p.masks[tile * p.mask_tile_words + rd.mask_off + (c0 >> 5) * 128 + row] = word;
          }
          if (valid) {
            float4 o;
            o.x = 1.f / (1.f + __expf(-a0));
            o.y = 1.f / (1.f + __expf(-a1));
            o.z = 1.f / (1.f + __expf(-a2));
            o.w = st.sigma;
            *reinterpret_cast<float4*>(p.out + m * 4) = o;
          }
        } else if (!kBwd && kNoF && rd.epi == MCF_EPI_NOF_HEAD) {
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
        } else if (kBwd && (rd.epi == MCF_EPI_B_MASK || rd.epi == MCF_EPI_B_MASK_SIGMA || rd.epi == MCF_EPI_B_LINEAR)) {
          for (int c0 = 0; c0 < rd.n_out; c0 += 32) {
            uint32_t v[32];
            float f[32];
            tmem_ld32(t_acc + c0, v);
            // select tree over registers (an indexed read would put the eight words in local memory)
            const int ci = c0 >> 5;
            const uint32_t s01 = (ci & 1) ? mwords[1] : mwords[0], s23 = (ci & 1) ? mwords[3] : mwords[2];
            const uint32_t s45 = (ci & 1) ? mwords[5] : mwords[4], s67 = (ci & 1) ? mwords[7] : mwords[6];
            const uint32_t s03 = (ci & 2) ? s23 : s01, s47 = (ci & 2) ? s67 : s45;
            const uint32_t word = (ci & 4) ? s47 : s03;
            tmem_ld_wait();
            if (!kNoF && rd.epi == MCF_EPI_B_MASK_SIGMA) {
              float ws[32];
              load32f(p.consts + rd.aux_off + c0, ws);
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = fmaf(st.sigma, ws[j], __uint_as_float(v[j]));
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = ((word >> j) & 1u) ? f[j] : 0.f;
            store_h32(hbuf, row, c0, f);
          }
        } else if (kBwd && !kNoF && rd.epi == MCF_EPI_B_DPE && p.d_dense != nullptr) {
          // gradient w.r.t. already-embedded input rows: the raw dX of the first / skip layer, summed over the rounds
          // (st.aux[0] counts the dX rounds of this tile: the first one writes, later ones add)
          float* dst = p.d_dense + m * p.d_dense_stride;
          for (int c0 = 0; c0 < 64; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(t_acc + c0, v);
            tmem_ld_wait();
            if (valid) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const int c = c0 + j;
                if (c < p.dense_cols) dst[c] = (st.aux[0] != 0.f ? dst[c] : 0.f) + __uint_as_float(v[j]);
              }
            }
          }
          st.aux[0] = 1.f;
        } else if (kBwd && rd.epi == MCF_EPI_B_DPE) {
          // d_xyz += J_PE(x)^T dPE, with sin/cos taken from the saved first-layer operand image
          const uint8_t* x0img = reinterpret_cast<const uint8_t*>(p.fwd_save) + tile_r * p.fwd_save_tile_bytes + p.fwd_x0_off;
          float pe[64];
#pragma unroll
          for (int c8 = 0; c8 < 8; ++c8) {
            uint4 t = *reinterpret_cast<const uint4*>(x0img + sw128_off(row, c8));
            pe[c8 * 8 + 0] = bf16_lo(t.x); pe[c8 * 8 + 1] = bf16_hi(t.x);
            pe[c8 * 8 + 2] = bf16_lo(t.y); pe[c8 * 8 + 3] = bf16_hi(t.y);
            pe[c8 * 8 + 4] = bf16_lo(t.z); pe[c8 * 8 + 5] = bf16_hi(t.z);
            pe[c8 * 8 + 6] = bf16_lo(t.w); pe[c8 * 8 + 7] = bf16_hi(t.w);
          }
          uint32_t v0[32], v1[32];
          tmem_ld32(t_acc, v0);
          tmem_ld32(t_acc + 32, v1);
          tmem_ld_wait();
          float dpe[64];
#pragma unroll
          for (int j = 0; j < 32; ++j) { dpe[j] = __uint_as_float(v0[j]); dpe[32 + j] = __uint_as_float(v1[j]); }
#pragma unroll
          for (int c = 0; c < 3; ++c) st.dx[c] += dpe[c];
#pragma unroll
          for (int k = 0; k < 10; ++k) {
            if (k < p.pe_n_freqs) {
              const float f = tab.pe_freq[k];
#pragma unroll
              for (int c = 0; c < 3; ++c) {
                const int is = 3 + 6 * k + c, ic = is + 3;
                st.dx[c] += f * (pe[ic] * dpe[is] - pe[is] * dpe[ic]);
              }
            }
          }
          if (rd.aux_off == 1u && valid && p.d_xyz) {
            p.d_xyz[m * 3 + 0] = st.dx[0]; p.d_xyz[m * 3 + 1] = st.dx[1]; p.d_xyz[m * 3 + 2] = st.dx[2];
          }
        } else {
          // an epilogue this instantiation was not compiled for: refuse loudly instead of skipping it
          if (gtid == 0) atomicExch(&g_mcf_device_error, 0xBADE0000u | (uint32_t)rd.epi);
        }

        tc_fence_before();
        MCF_TACC(2, te);
        // p.reserved0 != 0: signal the tensor core BEFORE pushing the operand image to HBM (the next layer's MMA and
        // the bulk store both only read the buffer; the barrier + bulk-copy issue then overlap the MMA)
        const bool early = p.reserved0 != 0;
        if (writes_h) fence_proxy_async_smem();
        if (early && r + 1 < p.n_rounds) arrive_act_ready();
        if (writes_h) {
          if (saving && rd.save_off != kNone) store_image(save_tile + rd.save_off, hbuf, ((uint32_t)rd.n_out + 63u) / 64u);
        }
        if (!early && r + 1 < p.n_rounds) arrive_act_ready();
        MCF_TACC(3, te);
      }
    }
    if (kWarpStore ? lane == 0 : gtid == 0) bulk_wait_all();
    if (timing && gtid == 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) p.timing[blockIdx.x * 16 + s * 4 + j] = tacc[j];
    }
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (C == 2) cluster_sync_all();  // the peer may still be signalling this CTA's barriers / reading its operands
  if (timing && threadIdx.x == 0) p.timing[blockIdx.x * 16 + 12] = (unsigned long long)(clock64() - t_kernel0);
  if (warp == 2) {
    if (C == 2) tmem_dealloc_pair(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------
// weight packer: fp32 nn.Linear tensors -> bf16 128B-swizzled chunk images / fp32 constants
// ---------------------------------------------------------------------------------------------
struct PackPtrs {
  const float* t[MCF_MAX_PACK_TENSORS];
};

__global__ void k_pack(const mcf_pack_t* __restrict__ table, PackPtrs ptrs, uint8_t* __restrict__ wpack,
                       float* __restrict__ consts) {
  const mcf_pack_t e = table[blockIdx.x];
  const float* src = ptrs.t[e.tensor];
  if (e.kind == 1) {
    for (uint32_t i = threadIdx.x; i < e.bytes; i += blockDim.x) {
      int r = (int)(i / (uint32_t)max(e.ncols, 1)), c = (int)(i % (uint32_t)max(e.ncols, 1));
      float v = 0.f;
      if (r < e.nrows && c < e.ncols) v = src[(long long)(e.row0 + r) * e.ld + e.col0 + c];
      consts[e.dst_off + i] = v;
    }
    return;
  }
  const uint32_t rows = e.bytes / 128u;
  for (uint32_t i = threadIdx.x; i < rows * 8u; i += blockDim.x) {
    const uint32_t r = i >> 3, c16 = i & 7u;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = (int)c16 * 8 + j;
      float v = 0.f;
      if ((int)r < e.nrows && c < e.ncols) {
        v = e.transposed ? src[(long long)(e.col0 + c) * e.ld + e.row0 + (int)r]
                         : src[(long long)(e.row0 + (int)r) * e.ld + e.col0 + c];
      }
      f[j] = v;
    }
    uint4 v;
    v.x = pack_bf16x2(f[0], f[1]); v.y = pack_bf16x2(f[2], f[3]);
    v.z = pack_bf16x2(f[4], f[5]); v.w = pack_bf16x2(f[6], f[7]);
    *reinterpret_cast<uint4*>(wpack + e.dst_off + sw128_off(r, c16)) = v;
  }
}

__global__ void k_unpack(const mcf_unpack_t* __restrict__ table, const float* __restrict__ staging,
                         float* __restrict__ grads) {
  const mcf_unpack_t e = table[blockIdx.x];
  const int n = e.nrows * e.ncols;
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < n; i += gridDim.y * blockDim.x) {
    const int r = i / e.ncols, c = i - r * e.ncols;
    const float v = e.transposed ? staging[e.src_off + (long long)c * e.src_ld + r]
                                 : staging[e.src_off + (long long)r * e.src_ld + c];
    grads[e.dst_off + (long long)r * e.dst_ld + c] = v;
  }
}

struct UnpackPtrs {
  float* d[MCF_MAX_UNPACK_PTRS];
};
__global__ void k_unpack_acc(const mcf_unpack_t* __restrict__ table, const float* __restrict__ staging, UnpackPtrs ptrs) {
  const mcf_unpack_t e = table[blockIdx.x];
  float* dst = ptrs.d[blockIdx.x];
  if (dst == nullptr) return;
  const int n = e.nrows * e.ncols;
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < n; i += gridDim.y * blockDim.x) {
    const int r = i / e.ncols, c = i - r * e.ncols;
    const float v = e.transposed ? staging[e.src_off + (long long)c * e.src_ld + r]
                                 : staging[e.src_off + (long long)r * e.src_ld + c];
    dst[(long long)r * e.dst_ld + c] += v;
  }
}

// Column sums of a row-major [n_rows][stride] fp32 array (first ncols columns).  The array is walked as a flat,
// fully coalesced stream; each thread keeps the running sum of the single column its flat indices map to
// (blockDim * gridDim is a multiple of stride, so that column never changes), then shared-memory + global atomics.
__global__ void k_colsum(const float* __restrict__ src, long long n_elems, int stride, int ncols,
                         float* __restrict__ out) {
  __shared__ float bins[16];
  if (threadIdx.x < 16) bins[threadIdx.x] = 0.f;
  __syncthreads();
  const long long step = (long long)gridDim.x * blockDim.x;
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int col = (int)(i % stride);
  float acc = 0.f;
  for (; i < n_elems; i += step) acc += src[i];
  if (col < ncols) atomicAdd(&bins[col], acc);
  __syncthreads();
  if (threadIdx.x < ncols) atomicAdd(out + threadIdx.x, bins[threadIdx.x]);
}

int launch_nof_ts(const mcf_chain_params_t& p, cudaStream_t stream);   // nof_chain.cu

}  // namespace mcf

extern "C" {

int mcf_unpack(const mcf_unpack_t* table_dev, int n_entries, const float* staging, float* grads,
               cudaStream_t stream) {
  if (n_entries <= 0) return 0;
  mcf::k_unpack<<<dim3(n_entries, 16), 256, 0, stream>>>(table_dev, staging, grads);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

int mcf_unpack_accumulate(const mcf_unpack_t* table_dev, int n_entries, const float* staging,
                          float* const* dst_ptrs_host, cudaStream_t stream) {
  if (n_entries <= 0) return 0;
  if (n_entries > MCF_MAX_UNPACK_PTRS) return MCF_ERR_BAD_ARG;
  mcf::UnpackPtrs ptrs;
  for (int i = 0; i < MCF_MAX_UNPACK_PTRS; ++i) ptrs.d[i] = i < n_entries ? dst_ptrs_host[i] : nullptr;
  mcf::k_unpack_acc<<<dim3(n_entries, 16), 256, 0, stream>>>(table_dev, staging, ptrs);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

int mcf_colsum(const float* src, long long n_rows, int stride, int ncols, float* out, cudaStream_t stream) {
  if (n_rows <= 0) return 0;
  if (ncols < 1 || ncols > 16 || stride < ncols || stride > 16) return MCF_ERR_BAD_ARG;
  // 192 threads = lcm-friendly for strides 4 and 12; grid*block must be a multiple of stride
  const int threads = 192;
  if ((threads % stride) != 0) return MCF_ERR_UNSUPPORTED;
  long long n_elems = n_rows * stride;
  long long blocks = (n_elems + threads * 8 - 1) / (threads * 8);
  if (blocks > 1184) blocks = 1184;
  if (blocks < 1) blocks = 1;
  mcf::k_colsum<<<(unsigned)blocks, threads, 0, stream>>>(src, n_elems, stride, ncols, out);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

int mcf_abi_version(void) { return MCF_ABI_VERSION; }

int mcf_device_error_flag(unsigned int* flag_host) {
  cudaError_t e = cudaDeviceSynchronize();
  unsigned int v = 0, zero = 0;
  cudaError_t e2 = cudaMemcpyFromSymbol(&v, mcf::g_mcf_device_error, sizeof(v));
  if (e2 == cudaSuccess) cudaMemcpyToSymbol(mcf::g_mcf_device_error, &zero, sizeof(zero));
  if (flag_host) *flag_host = v;
  if (e != cudaSuccess) return (int)e;
  if (e2 != cudaSuccess) return (int)e2;
  return v ? MCF_ERR_DEVICE_FLAG : 0;
}

int mcf_pack(const mcf_pack_t* table_dev, int n_entries, const float* const* tensors_host, int n_tensors, void* wpack,
             float* consts, cudaStream_t stream) {
  if (n_entries <= 0) return 0;
  if (n_tensors > MCF_MAX_PACK_TENSORS || n_tensors < 0) return MCF_ERR_BAD_ARG;
  mcf::PackPtrs ptrs;
  for (int i = 0; i < MCF_MAX_PACK_TENSORS; ++i) ptrs.t[i] = i < n_tensors ? tensors_host[i] : nullptr;
  mcf::k_pack<<<n_entries, 256, 0, stream>>>(table_dev, ptrs, reinterpret_cast<uint8_t*>(wpack), consts);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

int mcf_chain_launch(const mcf_chain_params_t* pp, cudaStream_t stream) {
  if (!pp) return MCF_ERR_BAD_ARG;
  const mcf_chain_params_t& p = *pp;
  if (p.n_rows <= 0) return 0;
  if (p.n_chunks <= 0 || p.n_chunks > mcf::kMaxChunks || p.n_rounds <= 0 || p.n_rounds > mcf::kMaxRounds)
    return MCF_ERR_BAD_ARG;
  if (p.rows_per_ray <= 0 || p.pe_n_freqs > 10 || p.pe_n_freqs < 0) return MCF_ERR_BAD_ARG;
  if (p.resident == 2) {   // NoF program built for the TMEM-resident kernel (nof_chain.cu)
#ifndef MCF_TIMING
    if (p.timing != nullptr) return MCF_ERR_UNSUPPORTED;
#endif
    return mcf::launch_nof_ts(p, stream);
  }
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    e = cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return (int)e;
  }
  long long n_tiles = (p.n_rows + MCF_TILE_ROWS - 1) / MCF_TILE_ROWS;
  const int ctas = (p.cta_pair != 0 && p.width == 256) ? 2 : 1;
  long long n_units = (n_tiles + 2 * ctas - 1) / (2 * ctas);
  int cap = (p.max_ctas > 0 ? p.max_ctas : n_sm) / ctas;
  if (cap < 1) cap = 1;
  int grid = (int)(n_units < cap ? n_units : cap) * ctas;
  const bool bwd = p.prologue == MCF_PRO_B_NERF || p.prologue == MCF_PRO_B_NOF;
  cudaError_t e;
  const void* fn = nullptr;
  int smem = 0;
  const bool save = bwd || p.save != nullptr || p.masks != nullptr || p.head_save != nullptr;
#define MCF_PICK2(W_, C_, N_)                                                                             \
  (bwd ? (const void*)mcf::k_chain<W_, C_, true, true, N_>                                                \
       : (save ? (const void*)mcf::k_chain<W_, C_, false, true, N_> : (const void*)mcf::k_chain<W_, C_, false, false, N_>))
#define MCF_PICK_RES()                                                                                    \
  (bwd ? (const void*)mcf::k_chain<128, 1, true, true, true, true>                                        \
       : (save ? (const void*)mcf::k_chain<128, 1, false, true, true, true>                               \
               : (const void*)mcf::k_chain<128, 1, false, false, true, true>))
#define MCF_PICK(W_, C_) (nof ? MCF_PICK2(W_, C_, true) : MCF_PICK2(W_, C_, false))
  const bool nof = p.program_kind == 1;
  if (p.program_kind != 0 && p.program_kind != 1) return MCF_ERR_BAD_ARG;
  if ((p.prologue == MCF_PRO_B_NOF) != (bwd && nof)) return MCF_ERR_BAD_ARG;
  if (p.width == 256 && ctas == 2) {
    fn = MCF_PICK(256, 2);
    smem = (int)mcf::Smem<256>::total;
  } else if (p.width == 256) {
    fn = MCF_PICK(256, 1);
    smem = (int)mcf::Smem<256>::total;
  } else if (p.width == 128 && p.resident) {
    // the program was built for the resident-weight kernel (x0 aliased into the activation buffer): NoF only
    if (!nof || p.wpack_bytes == 0 || p.wpack_bytes > mcf::kResBytes || (p.wpack_bytes & 15u)) return MCF_ERR_BAD_ARG;
    fn = MCF_PICK_RES();
    smem = (int)mcf::SmemRes<128>::total;
  } else if (p.width == 128) {
    fn = MCF_PICK(128, 1);
    smem = (int)mcf::Smem<128>::total;
  } else {
    return MCF_ERR_UNSUPPORTED;
  }
#undef MCF_PICK2
#undef MCF_PICK
#undef MCF_PICK_RES
#ifndef MCF_TIMING
  if (p.timing != nullptr) return MCF_ERR_UNSUPPORTED;   // cycle counters need a -DMCF_TIMING build
#endif
  e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return (int)e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(mcf::kThreads);
  cfg.dynamicSmemBytes = (size_t)smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)ctas;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  void* args[1] = {const_cast<mcf_chain_params_t*>(&p)};
  e = cudaLaunchKernelExC(&cfg, fn, args);
  if (e != cudaSuccess) return (int)e;
  e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

}  // extern "C"
