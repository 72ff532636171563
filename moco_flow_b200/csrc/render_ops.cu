// fp32 HBM-bound kernels of the ray-rendering path (compiled with -fmad=false so that the
// arithmetic is the same sequence of individually rounded fp32 operations the reference's
// unfused PyTorch ops perform):
//   * stratified depth sampling + point generation      (models/rendering.py:245-263)
//   * standalone positional encoding fwd/bwd            (models/embedding.py:42-46)
//   * per-ray bias folding of per-ray features          (replaces rendering.py:73-75,133-142 cat)
//   * alpha compositing fwd + analytic bwd              (models/rendering.py:158-190)
//   * sample_pdf inverse-CDF + sort-merge               (models/rendering.py:5-46, :326)
//   * masked flow-consistency residual fwd/bwd          (models/rendering.py:304-314,363-373)
// One warp per ray everywhere; lane-strided (coalesced) sample access; warp-shuffle scans.
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "../../include/moco_flow_b200.h"

namespace {

constexpr int kWarpsPerBlock = 4;
constexpr unsigned kFull = 0xffffffffu;

inline int grid_for_warps(long long n_warps, int warps_per_block = kWarpsPerBlock) {
  long long b = (n_warps + warps_per_block - 1) / warps_per_block;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

// -------------------------------------------------------------------------------------------------
// depth sampling + points
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ float z_at(float near, float far, float t, int use_disp) {
  float omt = 1.0f - t;
  if (!use_disp) return near * omt + far * t;                     // rendering.py:247
  return 1.0f / ((1.0f / near) * omt + (1.0f / far) * t);         // rendering.py:249
}

__global__ void k_coarse_samples(const float* __restrict__ rays, int ray_stride, const float* __restrict__ t_steps,
                                 const float* __restrict__ perturb_rand, float perturb, int use_disp, int R, int S,
                                 float* __restrict__ z_out, float* __restrict__ xyz_out) {
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= (long long)R * S) return;
  int r = static_cast<int>(idx / S), i = static_cast<int>(idx - (long long)r * S);
  const float* ray = rays + (long long)r * ray_stride;
  float near = ray[6], far = ray[7];
  float z = z_at(near, far, t_steps[i], use_disp);
  if (perturb > 0.0f) {  // rendering.py:253-260
    float lo = z, hi = z;
    if (i > 0) lo = 0.5f * (z_at(near, far, t_steps[i - 1], use_disp) + z);
    if (i < S - 1) hi = 0.5f * (z + z_at(near, far, t_steps[i + 1], use_disp));
    float pr = perturb * perturb_rand[idx];
    z = lo + (hi - lo) * pr;
  }
  z_out[idx] = z;
  if (xyz_out) {  // rendering.py:262-263
    xyz_out[idx * 3 + 0] = ray[0] + ray[3] * z;
    xyz_out[idx * 3 + 1] = ray[1] + ray[4] * z;
    xyz_out[idx * 3 + 2] = ray[2] + ray[5] * z;
  }
}

__global__ void k_points(const float* __restrict__ rays, int ray_stride, const float* __restrict__ z, int R, int S,
                         float* __restrict__ xyz_out) {  // rendering.py:329-330
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= (long long)R * S) return;
  int r = static_cast<int>(idx / S);
  const float* ray = rays + (long long)r * ray_stride;
  float zz = z[idx];
  xyz_out[idx * 3 + 0] = ray[0] + ray[3] * zz;
  xyz_out[idx * 3 + 1] = ray[1] + ray[4] * zz;
  xyz_out[idx * 3 + 2] = ray[2] + ray[5] * zz;
}

// -------------------------------------------------------------------------------------------------
// standalone positional encoding
// -------------------------------------------------------------------------------------------------
struct PEArgs {
  int n_freqs;
  float freq[MCF_MAX_FREQS];
  float weight[MCF_MAX_FREQS];
};

// tab (optional): device {freq[MCF_MAX_FREQS], weight[MCF_MAX_FREQS]} overriding the by-value tables
__global__ void k_pe_fwd(const float* __restrict__ x, long long B, int C, PEArgs pe, const float* __restrict__ tab,
                         int out_stride, float* __restrict__ out) {
  int OC = C * (2 * pe.n_freqs + 1);
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= B * OC) return;
  long long m = idx / OC;
  int oc = static_cast<int>(idx - m * OC);
  float v;
  if (oc < C) {
    v = x[m * C + oc];
  } else {
    int q = (oc - C) / C, c = (oc - C) - q * C;
    int k = q >> 1;
    float arg = (tab ? tab[k] : pe.freq[k]) * x[m * C + c];
    v = (tab ? tab[MCF_MAX_FREQS + k] : pe.weight[k]) * ((q & 1) ? cosf(arg) : sinf(arg));
  }
  out[m * out_stride + oc] = v;
}

// dx[m,c] = dy[m,c] + sum_k w_k f_k (cos(f_k x) dy_sin - sin(f_k x) dy_cos)
__global__ void k_pe_bwd(const float* __restrict__ x, const float* __restrict__ dy, long long B, int C, PEArgs pe,
                         const float* __restrict__ tab, int dy_stride, float* __restrict__ dx) {
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= B * C) return;
  long long m = idx / C;
  int c = static_cast<int>(idx - m * C);
  float xv = x[idx];
  const float* g = dy + m * dy_stride;
  float acc = g[c];
  for (int k = 0; k < pe.n_freqs; ++k) {
    float s, co;
    const float fk = tab ? tab[k] : pe.freq[k];
    sincosf(fk * xv, &s, &co);
    float wf = (tab ? tab[MCF_MAX_FREQS + k] : pe.weight[k]) * fk;
    acc += wf * (co * g[C + (2 * k) * C + c] - s * g[C + (2 * k + 1) * C + c]);
  }
  dx[idx] = acc;
}

// -------------------------------------------------------------------------------------------------
// per-ray bias:  out[r, n] = bias[n] + sum_j W[n, col_off + j] * feat[r, j]      (fp32, exact fold of
// the per-ray constant input columns -- index / direction embeddings -- of a Linear layer)
// -------------------------------------------------------------------------------------------------
// One block = 32 rays x all N outputs; the [N x E] weight slice is staged (transposed) in shared memory so that
// both the weight reads and the output writes are coalesced.
constexpr int kRbRays = 32;
__global__ void __launch_bounds__(256)
k_ray_bias(const float* __restrict__ W, int w_stride, int col_off, const float* __restrict__ bias,
           const float* __restrict__ feat, int feat_stride, int E, int R, int N, float* __restrict__ out) {
  extern __shared__ float sh[];
  float* sW = sh;                 // [E][N]  (transposed slice)
  float* sF = sh + (size_t)E * N; // [kRbRays][E]
  for (int i = threadIdx.x; i < N * E; i += blockDim.x) {
    int n = i / E, j = i - n * E;
    sW[j * N + n] = W[(long long)n * w_stride + col_off + j];
  }
  const int r0 = blockIdx.x * kRbRays;
  for (int i = threadIdx.x; i < kRbRays * E; i += blockDim.x) {
    int rr = i / E, j = i - rr * E;
    sF[i] = (r0 + rr < R) ? feat[(long long)(r0 + rr) * feat_stride + j] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kRbRays * N; i += blockDim.x) {
    int rr = i / N, n = i - rr * N;
    if (r0 + rr >= R) break;
    float acc = bias ? bias[n] : 0.0f;
    const float* f = sF + rr * E;
    for (int j = 0; j < E; ++j) acc = fmaf(sW[j * N + n], f[j], acc);
    out[(long long)(r0 + rr) * N + n] = acc;
  }
}

// -------------------------------------------------------------------------------------------------
// alpha compositing
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_incl_prod(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float n = __shfl_up_sync(kFull, v, o);
    if (lane >= o) v *= n;
  }
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

__device__ __forceinline__ float density(float raw, int act) {
  if (act == MCF_ACT_RELU) return fmaxf(raw, 0.0f);
  return raw > 20.0f ? raw : log1pf(expf(raw));  // nn.Softplus(beta=1, threshold=20)
}
__device__ __forceinline__ float density_grad(float raw, int act) {
  if (act == MCF_ACT_RELU) return raw > 0.0f ? 1.0f : 0.0f;
  if (raw > 20.0f) return 1.0f;
  float e = expf(raw);
  return e / (e + 1.0f);
}

// sigma at sigma[m*sigma_stride], rgb (optional) at rgb[m*rgb_stride + 0..2].  kPacked: the two point into one
// [M][4] = [r,g,b,sigma] array (what NeRF.forward returns, models/nerf.py:101) -> one 16-byte load per sample.
template <bool kPacked>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
k_composite_fwd(const float* __restrict__ sigma, int sigma_stride, const float* __restrict__ rgb, int rgb_stride,
                const float* __restrict__ z, const float* __restrict__ dirs, int dir_stride,
                const float* __restrict__ noise, float noise_std, const float* __restrict__ bg, int act, int R, int S,
                float* __restrict__ weights, float* __restrict__ alphas, float* __restrict__ rgb_out,
                float* __restrict__ depth_out, float* __restrict__ opacity_out) {
  int lane = threadIdx.x & 31;
  int r = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (r >= R) return;
  const float* d = dirs + (long long)r * dir_stride;
  float dn = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);  // rendering.py:164
  long long base = (long long)r * S;
  float carry = 1.0f, sw = 0.f, sr = 0.f, sg = 0.f, sb = 0.f, sd = 0.f;
  // The chunks of a ray are a serial chain (the transmittance carry), so the loads of chunks c + 1 and c + 2 are issued
  // before the scan of chunk c: three chunks of HBM requests in flight per warp.
  struct In { float zi, zn31, raw, c0, c1, c2, ns; };
  auto fetch = [&](int c) {
    In t = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const int i = c + lane;
    if (i < S) {
      t.zi = z[base + i];
      if (lane == 31 && i + 1 < S) t.zn31 = z[base + i + 1];   // the next chunk's first depth
      if (kPacked) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(rgb) + base + i);
        t.c0 = v.x; t.c1 = v.y; t.c2 = v.z; t.raw = v.w;
      } else {
        t.raw = sigma[(base + i) * sigma_stride];
        if (rgb) {
          const float* cc = rgb + (base + i) * rgb_stride;
          t.c0 = cc[0]; t.c1 = cc[1]; t.c2 = cc[2];
        }
      }
      if (noise) t.ns = noise[base + i];
    }
    return t;
  };
  In nxt = fetch(0), nxt2 = fetch(32);   // (fetch past the end returns zeros without touching memory)
  for (int c = 0; c < S; c += 32) {
    const In cur = nxt;
    nxt = nxt2;
    if (c + 64 < S) nxt2 = fetch(c + 64);
    int i = c + lane;
    bool ok = i < S;
    float zi = cur.zi;
    // z[i+1]: the neighbour lane's value; the last lane of a chunk holds the next chunk's first element
    float zn = __shfl_down_sync(kFull, zi, 1);
    if (lane == 31) zn = cur.zn31;
    float delta = (i + 1 < S) ? (zn - zi) : 1e10f;  // rendering.py:158-160
    delta = delta * dn;
    float raw = cur.raw, c0 = cur.c0, c1 = cur.c1, c2 = cur.c2;
    if (noise && ok) raw = raw + cur.ns * noise_std;  // rendering.py:166,170
    float alpha = ok ? (1.0f - expf(-delta * density(raw, act))) : 0.f;
    float q = ok ? (1.0f - alpha + 1e-10f) : 1.0f;  // rendering.py:177
    float incl = warp_incl_prod(q, lane);
    float excl = __shfl_up_sync(kFull, incl, 1);
    if (lane == 0) excl = 1.0f;
    float T = carry * excl;  // exclusive cumprod, rendering.py:179
    float w = alpha * T;
    carry = carry * __shfl_sync(kFull, incl, 31);
    if (ok) {
      if (weights) weights[base + i] = w;
      if (alphas) alphas[base + i] = alpha;
      sw += w;
      sd += w * zi;
      if (rgb) {
        sr += w * c0;
        sg += w * c1;
        sb += w * c2;
      }
    }
  }
  sw = warp_sum(sw);
  sd = warp_sum(sd);
  if (rgb) {
    sr = warp_sum(sr);
    sg = warp_sum(sg);
    sb = warp_sum(sb);
  }
  if (lane == 0) {
    if (opacity_out) opacity_out[r] = sw;
    if (rgb && rgb_out) {
      float rem = 1.0f - sw;
      if (bg) {  // rendering.py:189-190
        sr = sr + bg[r * 3 + 0] * rem;
        sg = sg + bg[r * 3 + 1] * rem;
        sb = sb + bg[r * 3 + 2] * rem;
      }
      rgb_out[r * 3 + 0] = sr;
      rgb_out[r * 3 + 1] = sg;
      rgb_out[r * 3 + 2] = sb;
    }
    if (rgb && depth_out) depth_out[r] = sd;
  }
}

template <int kPer>
__device__ __forceinline__ void ld_vec(const float* __restrict__ p, float* dst) {
  if (kPer == 2) {
    const float2 t = *reinterpret_cast<const float2*>(p);
    dst[0] = t.x; dst[1] = t.y;
  } else {
#pragma unroll
    for (int j = 0; j < kPer; j += 4) {
      const float4 t = *reinterpret_cast<const float4*>(p + j);
      dst[j] = t.x; dst[j + 1] = t.y; dst[j + 2] = t.z; dst[j + 3] = t.w;
    }
  }
}
template <int kPer>
__device__ __forceinline__ void st_vec(float* __restrict__ p, const float* src) {
  if (kPer == 2) {
    *reinterpret_cast<float2*>(p) = make_float2(src[0], src[1]);
  } else {
#pragma unroll
    for (int j = 0; j < kPer; j += 4) *reinterpret_cast<float4*>(p + j) = make_float4(src[j], src[j + 1], src[j + 2], src[j + 3]);
  }
}

// Blocked forward for S == 32 * kPer (64 / 128 / 256 samples: every configuration of the reference's YAMLs and of
// BASELINE.json): lane l owns the kPer CONSECUTIVE samples [l*kPer, (l+1)*kPer) of the ray, so one ray needs a single
// warp scan (of the lanes' transmittance products) instead of one per 32 samples, all addresses are one base plus
// immediates, there are no bounds predicates, and depths / weights / alphas move as 16-byte vectors.  ncu on the
// chunked kernel above showed it issue-bound (85-91 % issue-active at 28-62 % of DRAM bandwidth, profiles/r02_*).
// Same formulae; the exclusive product is associated as (product of the earlier lanes) * (product inside the lane).
template <bool kPacked, int kPer>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
k_composite_fwd_blk(const float* __restrict__ sigma, const float* __restrict__ rgb, const float* __restrict__ z,
                    const float* __restrict__ dirs, int dir_stride, const float* __restrict__ noise, float noise_std,
                    const float* __restrict__ bg, int act, int R, float* __restrict__ weights,
                    float* __restrict__ alphas, float* __restrict__ rgb_out, float* __restrict__ depth_out,
                    float* __restrict__ opacity_out) {
  constexpr int S = 32 * kPer;
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (r >= R) return;
  const float* d = dirs + (long long)r * dir_stride;
  const float dn = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);  // rendering.py:164
  const long long base = (long long)r * S + lane * kPer;
  float zv[kPer + 1], raw[kPer], c0[kPer], c1[kPer], c2[kPer];
  ld_vec<kPer>(z + base, zv);
  zv[kPer] = __shfl_down_sync(kFull, zv[0], 1);   // the next lane's first depth (unused by the ray's last sample)
  if (kPacked) {
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(rgb) + base + j);
      c0[j] = v.x; c1[j] = v.y; c2[j] = v.z; raw[j] = v.w;
    }
  } else {
    ld_vec<kPer>(sigma + base, raw);
#pragma unroll
    for (int j = 0; j < kPer; ++j) c0[j] = c1[j] = c2[j] = 0.f;
  }
  if (noise) {
    float ns[kPer];
    ld_vec<kPer>(noise + base, ns);
#pragma unroll
    for (int j = 0; j < kPer; ++j) raw[j] = raw[j] + ns[j] * noise_std;   // rendering.py:166,170
  }
  float alpha[kPer], pre[kPer];   // pre[j]: product of (1 - alpha + 1e-10) over this lane's samples before j
  float run = 1.0f;
#pragma unroll
  for (int j = 0; j < kPer; ++j) {
    const bool last = (lane == 31) && (j == kPer - 1);
    float delta = last ? 1e10f : (zv[j + 1] - zv[j]);  // rendering.py:158-160
    delta = delta * dn;
    alpha[j] = 1.0f - expf(-delta * density(raw[j], act));
    pre[j] = run;
    run = run * (1.0f - alpha[j] + 1e-10f);  // rendering.py:177
  }
  const float incl = warp_incl_prod(run, lane);
  float excl = __shfl_up_sync(kFull, incl, 1);
  if (lane == 0) excl = 1.0f;
  float sw = 0.f, sr = 0.f, sg = 0.f, sb = 0.f, sd = 0.f, w[kPer];
#pragma unroll
  for (int j = 0; j < kPer; ++j) {
    const float T = excl * pre[j];  // exclusive cumprod, rendering.py:179
    w[j] = alpha[j] * T;
    sw += w[j];
    sd += w[j] * zv[j];
    if (kPacked) {
      sr += w[j] * c0[j];
      sg += w[j] * c1[j];
      sb += w[j] * c2[j];
    }
  }
  if (weights) st_vec<kPer>(weights + base, w);
  if (alphas) st_vec<kPer>(alphas + base, alpha);
  sw = warp_sum(sw);
  if (kPacked) {
    sd = warp_sum(sd);
    sr = warp_sum(sr);
    sg = warp_sum(sg);
    sb = warp_sum(sb);
  }
  if (lane == 0) {
    if (opacity_out) opacity_out[r] = sw;
    if (kPacked && rgb_out) {
      const float rem = 1.0f - sw;
      if (bg) {  // rendering.py:189-190
        sr = sr + bg[r * 3 + 0] * rem;
        sg = sg + bg[r * 3 + 1] * rem;
        sb = sb + bg[r * 3 + 2] * rem;
      }
      rgb_out[r * 3 + 0] = sr;
      rgb_out[r * 3 + 1] = sg;
      rgb_out[r * 3 + 2] = sb;
    }
    if (kPacked && depth_out) depth_out[r] = sd;
  }
}

// Analytic backward (recomputes alpha/T; per-warp smem scratch holds e=exp(-delta*dens), T, delta).
// d_sigma written at d_sigma[m*ds_stride]; d_rgb (optional) at d_rgb[m*drgb_stride+0..2].
// kPacked: inputs and gradients are [M][4] = [r,g,b,sigma] arrays -> one 16-byte load and one 16-byte store per sample.
template <bool kPacked>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
k_composite_bwd(const float* __restrict__ sigma, int sigma_stride, const float* __restrict__ rgb, int rgb_stride,
                const float* __restrict__ z, const float* __restrict__ dirs, int dir_stride,
                const float* __restrict__ noise, float noise_std, const float* __restrict__ bg, int act, int R, int S,
                const float* __restrict__ g_rgb, const float* __restrict__ g_depth, const float* __restrict__ g_opacity,
                const float* __restrict__ g_weights, float* __restrict__ d_sigma, int ds_stride,
                float* __restrict__ d_rgb, int drgb_stride) {
  extern __shared__ float scratch[];
  int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  int r = blockIdx.x * kWarpsPerBlock + wib;
  if (r >= R) return;
  float* sE = scratch + (size_t)wib * 3 * S;
  float* sT = sE + S;
  float* sD = sT + S;
  const float* d = dirs + (long long)r * dir_stride;
  float dn = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  long long base = (long long)r * S;
  float carry = 1.0f;
  for (int c = 0; c < S; c += 32) {
    int i = c + lane;
    bool ok = i < S;
    float zi = ok ? z[base + i] : 0.f;
    float zn = __shfl_down_sync(kFull, zi, 1);
    if (lane == 31) zn = (i + 1 < S) ? z[base + i + 1] : 0.f;
    float delta = ((i + 1 < S) ? (zn - zi) : 1e10f) * dn;
    float raw = ok ? sigma[(base + i) * sigma_stride] : 0.f;
    if (noise && ok) raw = raw + noise[base + i] * noise_std;
    float e = ok ? expf(-delta * density(raw, act)) : 1.0f;
    float alpha = 1.0f - e;
    float q = ok ? (1.0f - alpha + 1e-10f) : 1.0f;
    float incl = warp_incl_prod(q, lane);
    float excl = __shfl_up_sync(kFull, incl, 1);
    if (lane == 0) excl = 1.0f;
    if (ok) {
      sE[i] = e;
      sT[i] = carry * excl;
      sD[i] = delta;
    }
    carry = carry * __shfl_sync(kFull, incl, 31);
  }
  __syncwarp();
  float gr = 0.f, gg = 0.f, gb = 0.f, gd = 0.f, go = 0.f, br = 0.f, bgc = 0.f, bb = 0.f;
  if (g_rgb) {
    gr = g_rgb[r * 3 + 0];
    gg = g_rgb[r * 3 + 1];
    gb = g_rgb[r * 3 + 2];
  }
  if (g_depth) gd = g_depth[r];
  if (g_opacity) go = g_opacity[r];
  if (bg) {
    br = bg[r * 3 + 0];
    bgc = bg[r * 3 + 1];
    bb = bg[r * 3 + 2];
  }
  float suffix = 0.f;  // sum_{j>i} g_j w_j, carried from the chunks to the right
  int last_chunk = ((S - 1) / 32) * 32;
  for (int c = last_chunk; c >= 0; c -= 32) {
    int i = c + lane;
    bool ok = i < S;
    float e = ok ? sE[i] : 1.f, T = ok ? sT[i] : 0.f, delta = ok ? sD[i] : 0.f;
    float alpha = 1.0f - e;
    float w = alpha * T;
    float g = 0.f;
    float c0 = 0.f, c1 = 0.f, c2 = 0.f, raw = 0.f;
    if (ok) {
      g = go + gd * z[base + i];
      if (kPacked) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(rgb) + base + i);
        c0 = v.x; c1 = v.y; c2 = v.z; raw = v.w;
        g += gr * (c0 - br) + gg * (c1 - bgc) + gb * (c2 - bb);
      } else {
        raw = sigma[(base + i) * sigma_stride];
        if (rgb) {
          const float* cc = rgb + (base + i) * rgb_stride;
          c0 = cc[0];
          c1 = cc[1];
          c2 = cc[2];
          g += gr * (c0 - br) + gg * (c1 - bgc) + gb * (c2 - bb);
        }
      }
      if (g_weights) g += g_weights[base + i];
    }
    float term = g * w;
    // reverse inclusive scan over lanes
    float incl = term;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      float n = __shfl_down_sync(kFull, incl, o);
      if (lane + o < 32) incl += n;
    }
    float after = suffix + (incl - term);  // strictly to the right of i
    suffix = suffix + __shfl_sync(kFull, incl, 0);
    if (ok) {
      float q = 1.0f - alpha + 1e-10f;
      float dalpha = g * T - after / q;  // cumprod backward (division form, as autograd)
      if (noise) raw = raw + noise[base + i] * noise_std;
      float ddens = dalpha * (delta * e);
      const float ds = ddens * density_grad(raw, act);
      if (kPacked) {
        reinterpret_cast<float4*>(d_rgb)[base + i] = make_float4(w * gr, w * gg, w * gb, ds);
      } else {
        d_sigma[(base + i) * ds_stride] = ds;
        if (d_rgb) {
          float* o = d_rgb + (base + i) * drgb_stride;
          o[0] = w * gr;
          o[1] = w * gg;
          o[2] = w * gb;
        }
      }
    }
  }
}

// -------------------------------------------------------------------------------------------------
// sample_pdf (+ optional sort-merge with the coarse depths)
// -------------------------------------------------------------------------------------------------
// count of leading elements of the ascending array a[0..n) that satisfy (a[i] <= x) (kLE) or (a[i] < x): a branch-free
// binary search with a fixed trip count (pow2 = smallest power of two >= n + 1), no divergence between the lanes
template <bool kLE>
__device__ __forceinline__ int count_below(const float* __restrict__ a, int n, int pow2, float x) {
  int lo = 0;   // invariant: a[0..lo) satisfy the predicate
#pragma unroll 1
  for (int step = pow2 >> 1; step > 0; step >>= 1) {
    const int probe = lo + step;
    const float v = a[min(probe, n) - 1];
    const bool ok = probe <= n && (kLE ? (v <= x) : (v < x));
    lo = ok ? probe : lo;
  }
  return lo;
}

// One warp per ray.  Per-warp smem: cdf[nb+1] | bins[nb+1] | sorted samples [32*kEPL] | coarse z [n_coarse] |
// merged [n_coarse + n_imp].
// The merge with the coarse depths (rendering.py:326: sort(cat(z, samples))) does not sort 2S values: the fine samples
// are sorted in registers (bitonic network over 32*kEPL values: shuffles for partner distances < 32, register
// exchanges above), the coarse depths are already sorted, and every value's output position is its own rank plus its
// rank in the other list (binary search; ties put the coarse value first, so the positions are a permutation).
// Values are copied, never recomputed: the result equals torch.sort of the concatenation bit for bit.
template <int kEPL>   // fine samples per lane: n_imp <= 32 * kEPL
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
k_sample_pdf(const float* __restrict__ bins, int bins_stride, int bins_are_z, const float* __restrict__ wts,
             int w_stride, const float* __restrict__ cdf_in, int cdf_stride, const float* __restrict__ u, int u_stride,
             float eps, int R, int nb, int n_imp, const float* __restrict__ z_coarse, int zc_stride, int n_coarse,
             float* __restrict__ samples, int* __restrict__ inds_out, float* __restrict__ cdf_out,
             float* __restrict__ z_merged) {
  extern __shared__ float scratch[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int r = blockIdx.x * kWarpsPerBlock + wib;
  if (r >= R) return;
  const int n_tot = n_coarse + n_imp;
  auto pow2_above = [](int n) { int p = 1; while (p < n + 1) p <<= 1; return p; };
  const int p2_cdf = pow2_above(nb + 1), p2_zc = pow2_above(n_coarse), p2_imp = pow2_above(n_imp);
  const int per_warp = 2 * (nb + 1) + (z_merged ? 32 * kEPL + n_coarse + n_tot : 0);
  float* s_cdf = scratch + (size_t)wib * per_warp;
  float* s_bin = s_cdf + (nb + 1);
  float* s_srt = s_bin + (nb + 1);
  float* s_zc = s_srt + 32 * kEPL;
  float* s_out = s_zc + n_coarse;

  // bins: given, or mid-points of the coarse depths (rendering.py:321)
  const float* brow = bins + (long long)r * bins_stride;
  for (int j = lane; j <= nb; j += 32) s_bin[j] = bins_are_z ? 0.5f * (brow[j] + brow[j + 1]) : brow[j];

  if (cdf_in) {
    const float* crow = cdf_in + (long long)r * cdf_stride;
    for (int j = lane; j <= nb; j += 32) s_cdf[j] = crow[j];
  } else {
    // rendering.py:20-23.  Fixed order: weights+eps in fp32; total accumulated in fp64 and rounded
    // once; pdf = w/total (IEEE fp32 division); cdf = running fp64 sum of pdf, rounded per element.
    const float* wrow = wts + (long long)r * w_stride;
    double part = 0.0;
    for (int j = lane; j < nb; j += 32) part += static_cast<double>(wrow[j] + eps);
    float total = static_cast<float>(warp_sum_d(part));
    double run = 0.0;
    if (lane == 0) s_cdf[0] = 0.0f;
    for (int c = 0; c < nb; c += 32) {
      int j = c + lane;
      double p = (j < nb) ? static_cast<double>((wrow[j] + eps) / total) : 0.0;
      double incl = p;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        double n = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += n;
      }
      if (j < nb) s_cdf[j + 1] = static_cast<float>(run + incl);
      run += __shfl_sync(kFull, incl, 31);
    }
  }
  if (z_merged) {
    const float* zc = z_coarse + (long long)r * zc_stride;
    for (int j = lane; j < n_coarse; j += 32) s_zc[j] = zc[j];
  }
  __syncwarp();
  if (cdf_out) {
    float* co = cdf_out + (long long)r * (nb + 1);
    for (int j = lane; j <= nb; j += 32) co[j] = s_cdf[j];
  }

  const float* urow = u + (long long)r * u_stride;
  float v[kEPL];   // element e = q * 32 + lane
#pragma unroll
  for (int q = 0; q < kEPL; ++q) {
    const int k = q * 32 + lane;
    v[q] = CUDART_INF_F;
    if (k < n_imp) {
      float uk = urow[k];
      // searchsorted(cdf, u, right=True): first index with cdf[idx] > u, in [0, nb+1]   (:33)
      const int lo = count_below<true>(s_cdf, nb + 1, p2_cdf, uk);
      int below = max(lo - 1, 0), above = min(lo, nb);  // :34-35
      float c0 = s_cdf[below], c1 = s_cdf[above];
      float b0 = s_bin[below], b1 = s_bin[above];
      float denom = c1 - c0;
      if (denom < eps) denom = 1.0f;  // :41-42
      float sv = b0 + (uk - c0) / denom * (b1 - b0);  // :45
      if (samples) samples[(long long)r * n_imp + k] = sv;
      if (inds_out) inds_out[(long long)r * n_imp + k] = lo;
      v[q] = sv;
    }
  }
  if (!z_merged) return;

  // ---- bitonic sort of the 32*kEPL fine samples (padding = +inf) in registers ----
#pragma unroll
  for (int k = 2; k <= 32 * kEPL; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= 32) {
        const int dq = j >> 5;
#pragma unroll
        for (int q = 0; q < kEPL; ++q) {
          if ((q & dq) == 0) {   // element (q, lane) vs (q + dq, lane); ascending block iff (e & k) == 0
            const bool up = (((q * 32) & k) == 0);   // k >= 64 here: bit of q only
            const float a = v[q], b = v[q + dq];
            const float lo = fminf(a, b), hi = fmaxf(a, b);
            v[q] = up ? lo : hi;
            v[q + dq] = up ? hi : lo;
          }
        }
      } else {
#pragma unroll
        for (int q = 0; q < kEPL; ++q) {
          const int e = q * 32 + lane;
          const float other = __shfl_xor_sync(kFull, v[q], j);
          const bool up = (e & k) == 0;
          const bool lower = (lane & j) == 0;           // this lane holds the smaller index of the pair
          const float lo = fminf(v[q], other), hi = fmaxf(v[q], other);
          v[q] = (lower == up) ? lo : hi;
        }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < kEPL; ++q) s_srt[q * 32 + lane] = v[q];
  __syncwarp();

  // ---- rank merge ----
#pragma unroll
  for (int q = 0; q < kEPL; ++q) {
    const int e = q * 32 + lane;
    if (e < n_imp) {   // position = own rank + #coarse <= value
      const float x = v[q];
      s_out[e + count_below<true>(s_zc, n_coarse, p2_zc, x)] = x;
    }
  }
  for (int j = lane; j < n_coarse; j += 32) {   // position = own rank + #samples < value
    const float x = s_zc[j];
    s_out[j + count_below<false>(s_srt, n_imp, p2_imp, x)] = x;
  }
  __syncwarp();
  float* zo = z_merged + (long long)r * n_tot;
  for (int j = lane; j < n_tot; j += 32) zo[j] = s_out[j];
}

// -------------------------------------------------------------------------------------------------
// masked flow-consistency residual
// -------------------------------------------------------------------------------------------------
// stats: [0] masked sum, [1] masked count, [2] total sum, [3] total count (doubles)
__global__ void k_masked_l1_fwd(const float* __restrict__ a, const float* __restrict__ b,
                                const float* __restrict__ alphas, float thresh, long long M,
                                float* __restrict__ resid, double* __restrict__ stats) {
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  float res = 0.f;
  bool sel = false;
  if (idx < M) {
    float d0 = fabsf(a[idx * 3 + 0] - b[idx * 3 + 0]);
    float d1 = fabsf(a[idx * 3 + 1] - b[idx * 3 + 1]);
    float d2 = fabsf(a[idx * 3 + 2] - b[idx * 3 + 2]);
    res = ((d0 + d1) + d2) / 3.0f;  // torch.mean over the 3 coordinates
    sel = alphas[idx] >= thresh;
    if (resid) resid[idx] = res;
  }
  if (!stats) return;
  double ms = warp_sum_d(sel ? (double)res : 0.0);
  double mc = warp_sum_d(sel ? 1.0 : 0.0);
  double ts = warp_sum_d(idx < M ? (double)res : 0.0);
  __shared__ double sh[3][32];
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) {
    sh[0][w] = ms;
    sh[1][w] = mc;
    sh[2][w] = ts;
  }
  __syncthreads();
  if (w == 0) {
    int nw = blockDim.x >> 5;
    ms = warp_sum_d(lane < nw ? sh[0][lane] : 0.0);
    mc = warp_sum_d(lane < nw ? sh[1][lane] : 0.0);
    ts = warp_sum_d(lane < nw ? sh[2][lane] : 0.0);
    if (lane == 0) {
      atomicAdd(&stats[0], ms);
      atomicAdd(&stats[1], mc);
      atomicAdd(&stats[2], ts);
      long long first = blockIdx.x * (long long)blockDim.x;
      long long n_here = M - first < (long long)blockDim.x ? M - first : (long long)blockDim.x;
      atomicAdd(&stats[3], (double)(n_here > 0 ? n_here : 0));
    }
  }
}

// mean over the selected samples (all samples if none is selected): rendering.py:306-311 + torch.mean
__global__ void k_masked_l1_finalize(const double* __restrict__ stats, float* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double v = stats[1] > 0.0 ? stats[0] / stats[1] : stats[2] / stats[3];
    out[0] = static_cast<float>(v);
  }
}

// d_b[m,:] = -sign(a-b)/3 * g_m ;  g_m = g_resid[m] (compat) or g_mean * sel_m / count (fused mean)
__global__ void k_masked_l1_bwd(const float* __restrict__ a, const float* __restrict__ b,
                                const float* __restrict__ alphas, float thresh, long long M,
                                const float* __restrict__ g_resid, const float* __restrict__ g_mean,
                                const double* __restrict__ stats, float grad_mul, float* __restrict__ d_b) {
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= M) return;
  float g;
  if (g_resid) {
    g = g_resid[idx];
  } else {
    double cnt = stats[1];
    bool sel = cnt > 0.0 ? (alphas[idx] >= thresh) : true;
    double n = cnt > 0.0 ? cnt : stats[3];
    g = sel ? static_cast<float>((double)g_mean[0] * (double)grad_mul / n) : 0.f;
  }
  g = g / 3.0f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float df = a[idx * 3 + c] - b[idx * 3 + c];
    float s = df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f);
    d_b[idx * 3 + c] = -s * g;
  }
}

inline int check_launch() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : static_cast<int>(e);
}

PEArgs make_pe(int n_freqs, const float* freqs, const float* weights) {
  PEArgs a;
  a.n_freqs = n_freqs;
  for (int i = 0; i < MCF_MAX_FREQS; ++i) {
    a.freq[i] = i < n_freqs ? freqs[i] : 0.f;
    a.weight[i] = i < n_freqs ? weights[i] : 0.f;
  }
  return a;
}

}  // namespace

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

int mcf_coarse_samples(const float* rays, int ray_stride, const float* t_steps, const float* perturb_rand,
                       float perturb, int use_disp, int n_rays, int n_samples, float* z_out, float* xyz_out,
                       cudaStream_t stream) {
  if (n_rays <= 0 || n_samples <= 0) return 0;
  if (perturb > 0.f && !perturb_rand) return MCF_ERR_BAD_ARG;
  long long n = (long long)n_rays * n_samples;
  k_coarse_samples<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(rays, ray_stride, t_steps, perturb_rand, perturb,
                                                                   use_disp, n_rays, n_samples, z_out, xyz_out);
  return check_launch();
}

int mcf_ray_points(const float* rays, int ray_stride, const float* z, int n_rays, int n_samples, float* xyz_out,
                   cudaStream_t stream) {
  if (n_rays <= 0 || n_samples <= 0) return 0;
  long long n = (long long)n_rays * n_samples;
  k_points<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(rays, ray_stride, z, n_rays, n_samples, xyz_out);
  return check_launch();
}

int mcf_pe_fwd(const float* x, long long n_rows, int in_channels, int n_freqs, const float* freqs_host,
               const float* weights_host, const float* table_dev, float* out, int out_stride, cudaStream_t stream) {
  if (n_freqs > MCF_MAX_FREQS || n_freqs < 0) return MCF_ERR_BAD_ARG;
  if (n_rows <= 0) return 0;
  PEArgs pe = make_pe(n_freqs, freqs_host, weights_host);
  long long n = n_rows * in_channels * (2 * n_freqs + 1);
  k_pe_fwd<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(x, n_rows, in_channels, pe, table_dev, out_stride, out);
  return check_launch();
}

int mcf_pe_bwd(const float* x, const float* dy, long long n_rows, int in_channels, int n_freqs,
               const float* freqs_host, const float* weights_host, const float* table_dev, int dy_stride, float* dx,
               cudaStream_t stream) {
  if (n_freqs > MCF_MAX_FREQS || n_freqs < 0) return MCF_ERR_BAD_ARG;
  if (n_rows <= 0) return 0;
  PEArgs pe = make_pe(n_freqs, freqs_host, weights_host);
  long long n = n_rows * in_channels;
  k_pe_bwd<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(x, dy, n_rows, in_channels, pe, table_dev, dy_stride, dx);
  return check_launch();
}

int mcf_ray_bias(const float* W, int w_stride, int col_off, const float* bias, const float* feat, int feat_stride,
                 int n_feat, int n_rays, int n_out, float* out, cudaStream_t stream) {
  if (n_rays <= 0 || n_out <= 0) return 0;
  size_t smem = ((size_t)n_feat * n_out + (size_t)kRbRays * n_feat) * sizeof(float);
  if (smem > 160 * 1024) return MCF_ERR_UNSUPPORTED;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k_ray_bias, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  k_ray_bias<<<(n_rays + kRbRays - 1) / kRbRays, 256, smem, stream>>>(W, w_stride, col_off, bias, feat, feat_stride,
                                                                      n_feat, n_rays, n_out, out);
  return check_launch();
}

int mcf_composite_fwd(const float* sigma, int sigma_stride, const float* rgb, int rgb_stride, const float* z,
                      const float* dirs, int dir_stride, const float* noise, float noise_std, const float* background,
                      int activation, int n_rays, int n_samples, float* weights, float* alphas, float* rgb_out,
                      float* depth_out, float* opacity_out, cudaStream_t stream) {
  if (activation != MCF_ACT_RELU && activation != MCF_ACT_SOFTPLUS) return MCF_ERR_BAD_ARG;
  if (n_rays <= 0 || n_samples <= 0) return 0;
  // [r,g,b,sigma] rows: 16-byte vector path
  const bool packed = rgb != nullptr && sigma == rgb + 3 && sigma_stride == 4 && rgb_stride == 4 &&
                      (reinterpret_cast<uintptr_t>(rgb) & 15u) == 0;
  // blocked kernel: S = 64 / 128 / 256, sigma either packed with rgb or a dense [R][S] array, everything 16-byte aligned
  const bool dense_sigma = rgb == nullptr && sigma_stride == 1;
  auto al16 = [](const void* q) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
  if ((packed || dense_sigma) && (n_samples == 64 || n_samples == 128 || n_samples == 256) && al16(sigma - (packed ? 3 : 0)) &&
      al16(z) && al16(noise) && al16(weights) && al16(alphas)) {
#define MCF_CBLK(P_, K_)                                                                                       \
    k_composite_fwd_blk<P_, K_><<<grid_for_warps(n_rays), kWarpsPerBlock * 32, 0, stream>>>(                   \
        sigma, rgb, z, dirs, dir_stride, noise, noise_std, background, activation, n_rays, weights, alphas, rgb_out, \
        depth_out, opacity_out)
    if (packed) {
      if (n_samples == 64) MCF_CBLK(true, 2);
      else if (n_samples == 128) MCF_CBLK(true, 4);
      else MCF_CBLK(true, 8);
    } else {
      if (n_samples == 64) MCF_CBLK(false, 2);
      else if (n_samples == 128) MCF_CBLK(false, 4);
      else MCF_CBLK(false, 8);
    }
#undef MCF_CBLK
    return check_launch();
  }
  if (packed)
    k_composite_fwd<true><<<grid_for_warps(n_rays), kWarpsPerBlock * 32, 0, stream>>>(
        sigma, sigma_stride, rgb, rgb_stride, z, dirs, dir_stride, noise, noise_std, background, activation, n_rays,
        n_samples, weights, alphas, rgb_out, depth_out, opacity_out);
  else
    k_composite_fwd<false><<<grid_for_warps(n_rays), kWarpsPerBlock * 32, 0, stream>>>(
        sigma, sigma_stride, rgb, rgb_stride, z, dirs, dir_stride, noise, noise_std, background, activation, n_rays,
        n_samples, weights, alphas, rgb_out, depth_out, opacity_out);
  return check_launch();
}

int mcf_composite_bwd(const float* sigma, int sigma_stride, const float* rgb, int rgb_stride, const float* z,
                      const float* dirs, int dir_stride, const float* noise, float noise_std, const float* background,
                      int activation, int n_rays, int n_samples, const float* g_rgb, const float* g_depth,
                      const float* g_opacity, const float* g_weights, float* d_sigma, int d_sigma_stride, float* d_rgb,
                      int d_rgb_stride, cudaStream_t stream) {
  if (activation != MCF_ACT_RELU && activation != MCF_ACT_SOFTPLUS) return MCF_ERR_BAD_ARG;
  if (n_rays <= 0 || n_samples <= 0) return 0;
  size_t smem = (size_t)kWarpsPerBlock * 3 * n_samples * sizeof(float);
  if (smem > 200 * 1024) return MCF_ERR_UNSUPPORTED;
  const bool packed = rgb != nullptr && d_rgb != nullptr && sigma == rgb + 3 && d_sigma == d_rgb + 3 &&
                      sigma_stride == 4 && rgb_stride == 4 && d_sigma_stride == 4 && d_rgb_stride == 4 &&
                      ((reinterpret_cast<uintptr_t>(rgb) | reinterpret_cast<uintptr_t>(d_rgb)) & 15u) == 0;
  if (smem > 48 * 1024) {
    cudaError_t e = packed ? cudaFuncSetAttribute(k_composite_bwd<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                           : cudaFuncSetAttribute(k_composite_bwd<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  if (packed)
    k_composite_bwd<true><<<grid_for_warps(n_rays), kWarpsPerBlock * 32, smem, stream>>>(
        sigma, sigma_stride, rgb, rgb_stride, z, dirs, dir_stride, noise, noise_std, background, activation, n_rays,
        n_samples, g_rgb, g_depth, g_opacity, g_weights, d_sigma, d_sigma_stride, d_rgb, d_rgb_stride);
  else
    k_composite_bwd<false><<<grid_for_warps(n_rays), kWarpsPerBlock * 32, smem, stream>>>(
        sigma, sigma_stride, rgb, rgb_stride, z, dirs, dir_stride, noise, noise_std, background, activation, n_rays,
        n_samples, g_rgb, g_depth, g_opacity, g_weights, d_sigma, d_sigma_stride, d_rgb, d_rgb_stride);
  return check_launch();
}

int mcf_sample_pdf(const float* bins, int bins_stride, int bins_are_z, const float* weights, int w_stride,
                   const float* cdf_in, int cdf_stride, const float* u, int u_stride, float eps, int n_rays, int n_bins,
                   int n_importance, const float* z_coarse, int zc_stride, int n_coarse, float* samples, int* inds_out,
                   float* cdf_out, float* z_merged, cudaStream_t stream) {
  if (n_rays <= 0 || n_importance <= 0) return 0;
  if (n_bins < 1 || !u || (!weights && !cdf_in)) return MCF_ERR_BAD_ARG;
  if (z_merged && !z_coarse) return MCF_ERR_BAD_ARG;
  if (n_importance > 256) return MCF_ERR_UNSUPPORTED;
  const int epl = n_importance <= 64 ? 2 : (n_importance <= 128 ? 4 : 8);
  size_t per_warp = 2 * (size_t)(n_bins + 1) + (z_merged ? (size_t)32 * epl + n_coarse + n_coarse + n_importance : 0);
  size_t smem = (size_t)kWarpsPerBlock * per_warp * sizeof(float);
  if (smem > 200 * 1024) return MCF_ERR_UNSUPPORTED;
#define MCF_SPDF(E_)                                                                                                  \
  do {                                                                                                                \
    if (smem > 48 * 1024) {                                                                                           \
      cudaError_t e = cudaFuncSetAttribute(k_sample_pdf<E_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      if (e != cudaSuccess) return (int)e;                                                                            \
    }                                                                                                                 \
    k_sample_pdf<E_><<<grid_for_warps(n_rays), kWarpsPerBlock * 32, smem, stream>>>(                                  \
        bins, bins_stride, bins_are_z, weights, w_stride, cdf_in, cdf_stride, u, u_stride, eps, n_rays, n_bins,       \
        n_importance, z_coarse, zc_stride, n_coarse, samples, inds_out, cdf_out, z_merged);                           \
  } while (0)
  if (epl == 2) MCF_SPDF(2);
  else if (epl == 4) MCF_SPDF(4);
  else MCF_SPDF(8);
#undef MCF_SPDF
  return check_launch();
}

int mcf_masked_l1_fwd(const float* a, const float* b, const float* alphas, float thresh, long long n_points,
                      float* resid, double* stats, float* mean_out, cudaStream_t stream) {
  if (n_points <= 0) return 0;
  if (stats) {
    cudaError_t e = cudaMemsetAsync(stats, 0, 4 * sizeof(double), stream);
    if (e != cudaSuccess) return (int)e;
  }
  k_masked_l1_fwd<<<(unsigned)((n_points + 255) / 256), 256, 0, stream>>>(a, b, alphas, thresh, n_points, resid, stats);
  if (stats && mean_out) k_masked_l1_finalize<<<1, 32, 0, stream>>>(stats, mean_out);
  return check_launch();
}

int mcf_masked_l1_finalize(const double* stats, float* mean_out, cudaStream_t stream) {
  if (!stats || !mean_out) return MCF_ERR_BAD_ARG;
  k_masked_l1_finalize<<<1, 32, 0, stream>>>(stats, mean_out);
  return check_launch();
}

int mcf_masked_l1_bwd(const float* a, const float* b, const float* alphas, float thresh, long long n_points,
                      const float* g_resid, const float* g_mean, const double* stats, float grad_mul, float* d_b,
                      cudaStream_t stream) {
  if (n_points <= 0) return 0;
  if (!g_resid && !(g_mean && stats)) return MCF_ERR_BAD_ARG;
  k_masked_l1_bwd<<<(unsigned)((n_points + 255) / 256), 256, 0, stream>>>(a, b, alphas, thresh, n_points, g_resid,
                                                                         g_mean, stats, grad_mul, d_b);
  return check_launch();
}

}  // extern "C"
