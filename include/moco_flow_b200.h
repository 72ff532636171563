/* moco_flow_b200 -- C ABI of the B200-native MoCo-Flow ray-rendering hot path.
 *
 * One shared library (moco_flow_b200/csrc/libmoco_flow_b200.so), plain pointers and sizes, no C++
 * or torch types.  All pointers are DEVICE pointers unless the name ends in `_host`.  Every entry
 * is stream-ordered, allocates nothing, never synchronises, and returns 0 on success, a positive
 * cudaError_t, or a negative MCF_ERR_* code.
 *
 * The reference (wyysf-98/MoCo_Flow) has no FFI: its boundary for this path is a set of Python
 * callables.  Each entry below names the reference code it replaces (paths relative to the
 * reference root); the Python modules of `moco_flow_b200/` re-creates the Python callables on top of these entries
 * and INTEGRATION.md shows the import switch a maintainer makes.
 */
#ifndef MOCO_FLOW_B200_H_
#define MOCO_FLOW_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __CUDA_RUNTIME_H__
typedef struct CUstream_st* cudaStream_t;
#endif

#define MCF_ABI_VERSION 1
#define MCF_MAX_FREQS 24
#define MCF_TILE_ROWS 128      /* samples per tensor-core tile                      */
#define MCF_BLOCK_BYTES 16384  /* one [128 rows][64 bf16] 128B-swizzled image block */

#define MCF_ACT_RELU 0
#define MCF_ACT_SOFTPLUS 1

#define MCF_ERR_BAD_ARG (-1)
#define MCF_ERR_UNSUPPORTED (-2)
#define MCF_ERR_DEVICE_FLAG (-3)

int mcf_abi_version(void);
/* Reads (and clears) the device-side protocol error flag set by a kernel whose bounded barrier
 * wait expired.  Synchronises the device; for tests and debugging only. */
int mcf_device_error_flag(unsigned int* flag_host);

/* ---- sampling ------------------------------------------------------------------------------ */
/* models/rendering.py:245-263  z_vals (+stratified jitter) and xyz = o + d*z.
 * rays: [n_rays][ray_stride] rows [o(3) d(3) near far ...]; t_steps = linspace(0,1,n_samples). */
int mcf_coarse_samples(const float* rays, int ray_stride, const float* t_steps, const float* perturb_rand,
                       float perturb, int use_disp, int n_rays, int n_samples, float* z_out, float* xyz_out,
                       cudaStream_t stream);
/* models/rendering.py:329-330 */
int mcf_ray_points(const float* rays, int ray_stride, const float* z, int n_rays, int n_samples, float* xyz_out,
                   cudaStream_t stream);

/* ---- positional encoding (standalone Embedding.forward), models/embedding.py:42-46 ---------- */
/* table_dev (may be NULL): device [2*MCF_MAX_FREQS] floats {freq[], weight[]} that override the host arrays --
 * the form to use under CUDA-graph capture when Embedding.weights changes between replays. */
int mcf_pe_fwd(const float* x, long long n_rows, int in_channels, int n_freqs, const float* freqs_host,
               const float* weights_host, const float* table_dev, float* out, int out_stride, cudaStream_t stream);
int mcf_pe_bwd(const float* x, const float* dy, long long n_rows, int in_channels, int n_freqs,
               const float* freqs_host, const float* weights_host, const float* table_dev, int dy_stride, float* dx,
               cudaStream_t stream);

/* out[r][n] = bias[n] + sum_j W[n][col_off+j] * feat[r][j]: exact fp32 fold of the per-ray constant
 * input columns (replaces the repeat_interleave+cat of models/rendering.py:73-75,133-142). */
int mcf_ray_bias(const float* W, int w_stride, int col_off, const float* bias, const float* feat, int feat_stride,
                 int n_feat, int n_rays, int n_out, float* out, cudaStream_t stream);

/* ---- alpha compositing, models/rendering.py:158-190 ---------------------------------------- */
/* sigma at sigma[m*sigma_stride]; rgb (may be NULL: weights-only) at rgb[m*rgb_stride+0..2]. */
int mcf_composite_fwd(const float* sigma, int sigma_stride, const float* rgb, int rgb_stride, const float* z,
                      const float* dirs, int dir_stride, const float* noise, float noise_std, const float* background,
                      int activation, int n_rays, int n_samples, float* weights, float* alphas, float* rgb_out,
                      float* depth_out, float* opacity_out, cudaStream_t stream);
/* analytic backward of the above w.r.t. sigma and rgb (what autograd derives for :158-190) */
int mcf_composite_bwd(const float* sigma, int sigma_stride, const float* rgb, int rgb_stride, const float* z,
                      const float* dirs, int dir_stride, const float* noise, float noise_std, const float* background,
                      int activation, int n_rays, int n_samples, const float* g_rgb, const float* g_depth,
                      const float* g_opacity, const float* g_weights, float* d_sigma, int d_sigma_stride, float* d_rgb,
                      int d_rgb_stride, cudaStream_t stream);

/* ---- sample_pdf, models/rendering.py:5-46 (+ the sort-merge of :326) ------------------------ */
/* bins: [n_rays][n_bins+1] (or, with bins_are_z, the coarse depths whose mid-points are the bins);
 * weights: [n_rays][n_bins] (ignored when cdf_in is given: "level 1" entry that consumes a
 * caller-built cdf [n_rays][n_bins+1]); u: [n_rays][n_importance].  Optional outputs: samples,
 * inds_out (searchsorted result), cdf_out, z_merged = sort(cat(z_coarse, samples)). */
int mcf_sample_pdf(const float* bins, int bins_stride, int bins_are_z, const float* weights, int w_stride,
                   const float* cdf_in, int cdf_stride, const float* u, int u_stride, float eps, int n_rays, int n_bins,
                   int n_importance, const float* z_coarse, int zc_stride, int n_coarse, float* samples, int* inds_out,
                   float* cdf_out, float* z_merged, cudaStream_t stream);

/* ---- flow-consistency residual, models/rendering.py:304-314,363-373 ------------------------- */
/* resid[m] = mean_3 |a-b| ; stats = {masked sum, masked count, total sum, total count} (4 doubles, zeroed by the
 * call); mean_out[0] = mean of resid over alphas>=thresh (over everything if the mask is empty), or NULL to leave
 * the mean to mcf_masked_l1_finalize -- the data-parallel path all-reduces the stats of all ranks in between so that
 * the mean is the one over the whole (global) ray batch, as the reference's single-process step computes it
 * (models/rendering.py:306-311 + trainer/trainer_moco_flow.py:319-327). */
int mcf_masked_l1_fwd(const float* a, const float* b, const float* alphas, float thresh, long long n_points,
                      float* resid, double* stats, float* mean_out, cudaStream_t stream);
int mcf_masked_l1_finalize(const double* stats, float* mean_out, cudaStream_t stream);
/* g_resid: per-sample upstream gradient (dynamic-length compat path), or g_mean + stats: gradient of the masked
 * mean; grad_mul scales it (world size under data parallelism: ranks' gradients are averaged afterwards). */
int mcf_masked_l1_bwd(const float* a, const float* b, const float* alphas, float thresh, long long n_points,
                      const float* g_resid, const float* g_mean, const double* stats, float grad_mul, float* d_b,
                      cudaStream_t stream);

/* ---- fused MLP chains on tcgen05/TMEM: models/nerf.py:61-102, models/nof.py:55-85 ----------- */
/* Weights are re-packed from the fp32 nn.Linear tensors into bf16 128B-swizzled UMMA operand images
 * (and fp32 constants) by mcf_pack; a chain launch then streams those images with bulk copies.
 * The tables are built by the host shim (moco_flow_b200/plans.py) and are opaque to callers. */
typedef struct {
  uint32_t dst_off;   /* byte offset into the packed stream (kind 0) / float offset into consts (kind 1) */
  uint32_t bytes;     /* kind 0: image bytes = padded_rows*128 ; kind 1: number of floats                 */
  int32_t kind;       /* 0: bf16 swizzled image block, 1: fp32 copy                                       */
  int32_t tensor;     /* index into the pointer array                                                     */
  int32_t row0, nrows;
  int32_t col0, ncols;
  int32_t ld;         /* source leading dimension (elements)                                              */
  int32_t transposed; /* image(r,c) = src[(col0+c)*ld + row0+r] instead of src[(row0+r)*ld + col0+c]      */
} mcf_pack_t;

#define MCF_MAX_PACK_TENSORS 32
int mcf_pack(const mcf_pack_t* table_dev, int n_entries, const float* const* tensors_host, int n_tensors,
             void* wpack, float* consts, cudaStream_t stream);

typedef struct {
  uint32_t src_off; /* byte offset of the chunk image in the packed stream */
  uint32_t bytes;
  uint8_t a_buf;    /* 0: input block X0, 1: activation buffer H */
  uint8_t a_kblock; /* 64-column block of that buffer            */
  uint8_t ksteps;   /* K=16 MMAs issued from this chunk (1..4)   */
  uint8_t flags;    /* bit0: first MMA overwrites the accumulator; bit1: this chunk and the next are the 128-row halves
                     * of one [256 x 64] weight tile; bit3: one CTA's halves of this tile and of the next k-block's
                     * tile are adjacent in the packed stream (CTA pairs: one 32 KB copy)                       */
  uint16_t n;       /* MMA N                                       */
  uint16_t acc_col; /* accumulator column offset inside the slot   */
} mcf_chunk_t;

typedef struct {
  uint16_t epi;      /* MCF_EPI_*                                                   */
  uint16_t n_out;    /* accumulator columns this epilogue consumes                  */
  uint16_t acc_col;  /* first accumulator column                                    */
  uint16_t chunk_begin, chunk_end;
  int16_t raybias;   /* -1, or index of the per-ray bias matrix [n_rays][n_out]     */
  uint32_t const_off;/* float offset of the column bias in consts                   */
  uint32_t aux_off;  /* float offset of head weights (sigma / rgb) in consts        */
  uint32_t save_off; /* byte offset inside the per-tile save record, 0xFFFFFFFF none */
  uint32_t mask_off; /* word offset inside the per-tile mask record, 0xFFFFFFFF none */
  uint32_t reserved;
} mcf_round_t;

#define MCF_EPI_RELU 0         /* H = relu(acc + bias)                                  */
#define MCF_EPI_RELU_SIGMA 1   /* + sigma = h . w_sigma + b_sigma                       */
#define MCF_EPI_LINEAR 2       /* H = acc + bias                                        */
#define MCF_EPI_NERF_RGB 3     /* he = relu(acc+bias); rgb = sigmoid(W_rgb he + b_rgb)  */
#define MCF_EPI_NOF_HEAD 4     /* quaternion / residual head of NoF                     */
#define MCF_EPI_B_MASK 16      /* H = acc * mask                                        */
#define MCF_EPI_B_MASK_SIGMA 17/* H = (acc + d_sigma w_sigma) * mask                    */
#define MCF_EPI_B_LINEAR 18    /* H = acc                                               */
#define MCF_EPI_B_DPE 19       /* d_xyz += J_PE^T acc                                   */

#define MCF_PRO_PE_XYZ 0  /* X0 = PE(xyz[m]) zero-padded                      */
#define MCF_PRO_DENSE 1   /* X0 = dense[m, :dense_cols] zero-padded           */
#define MCF_PRO_B_NERF 2  /* backward of the NeRF heads                       */
#define MCF_PRO_B_NOF 3   /* backward of the NoF head                         */

typedef struct {
  /* program */
  const mcf_chunk_t* chunks;
  const mcf_round_t* rounds;
  int32_t n_chunks, n_rounds;
  int32_t width;    /* hidden width: 128 or 256 */
  int32_t prologue; /* MCF_PRO_*                */
  const void* wpack;
  const float* consts;
  const float* raybias[4];
  /* problem */
  long long n_rows;        /* M = rays*samples (or points)                   */
  int32_t rows_per_ray;    /* S: ray(m) = m / S                              */
  int32_t n_rays;
  /* prologue inputs */
  const float* xyz;        /* [M][3]                                         */
  const float* dense;      /* [M][dense_stride] (MCF_PRO_DENSE)              */
  int32_t dense_stride, dense_cols;
  int32_t pe_n_freqs, pe_pad_to;
  float pe_freq[MCF_MAX_FREQS];
  float pe_weight[MCF_MAX_FREQS];
  /* outputs */
  float* out;              /* NeRF: [M][out_stride] (rgb at +0..2, sigma at +sigma_col); NoF: [M][3] */
  int32_t out_stride, sigma_col;
  int32_t use_quat;
  float* head_save;        /* NoF: [M][12] fp32 {v,s,t,x} for the backward, or NULL */
  /* training saves (NULL = inference) */
  void* save;              /* per-tile records of bf16 images                */
  long long save_tile_bytes;
  uint32_t* masks;         /* per-tile ReLU bit masks                        */
  long long mask_tile_words;
  uint32_t x0_save_off;    /* 0xFFFFFFFF none                                */
  /* backward inputs */
  const float* g_out;      /* NeRF: [M][4] grad of rgbsigma; NoF: [M][3] grad of xyz_out */
  const float* fwd_out;    /* NeRF: forward [M][4] output                    */
  const void* fwd_save;    /* forward save records                           */
  long long fwd_save_tile_bytes;
  const uint32_t* fwd_masks;
  long long fwd_mask_tile_words;
  uint32_t fwd_x0_off, fwd_he_off;
  float* d_xyz;            /* [M][3] grad w.r.t. the input points, or NULL   */
  float* d_head;           /* fp32 head gradients: NeRF [M][4] {d_pre_rgb(3), d_sigma}; NoF [M][12] {d v,s,t} */
  /* unused since ABI rev of round 2 (the per-ray feature images are built once per ray layout by
   * mcf_rayfeat_image and shared by every evaluation, instead of being written by each chain launch) */
  const float* rayfeat;
  int32_t rayfeat_stride, rayfeat_dim;
  uint32_t extra_save_off;
  uint32_t dhead_save_off; /* backward: image block of the head gradients    */
  int32_t max_ctas;        /* 0 = one CTA per SM                             */
  /* optional instrumentation: [n_ctas][16] cycle counters (NULL = off):
   * 0-3 slot0 {prologue, wait acc_full, epilogue, save/barrier}, 4-7 slot1, 8 MMA wait act_ready,
   * 9 MMA wait weights, 10 MMA issue, 11 producer wait ring, 12 total */
  unsigned long long* timing;
  /* width 256 only: non-zero launches clusters of two CTAs that run tcgen05 cta_group::2 (one M=256 MMA stream over
   * both CTAs' tiles; every CTA fetches only its half of each weight chunk) */
  int32_t cta_pair;
  /* 0: NeRF program (sigma / rgb heads), 1: NoF program (flow head).  Selects a kernel instantiation that carries only
   * that family's prologues and epilogues; a program of the other family fails with the device error flag. */
  int32_t program_kind;
  /* optional device copy of the encoder tables, [2*MCF_MAX_FREQS] floats = {freq[], weight[]}: when non-NULL it
   * replaces pe_freq / pe_weight above.  The reference re-assigns Embedding.weights every step of its coarse-to-fine
   * schedule (trainer/trainer_moco_flow.py:280-305); values passed by pointer stay current when the launch is
   * replayed from a captured CUDA graph, values passed in this struct are frozen at capture. */
  const float* pe_table;
  /* resident != 0 (width 128, NoF programs): the whole packed weight stream (wpack_bytes <= 144 KB) is copied into
   * shared memory once per CTA instead of being streamed per tile.
   *   1: chain.cu's kernel with activations in shared memory; the first-layer operand shares the activation buffer
   *      (the plan precomputes the skip layer's x0 part in round 0);
   *   2: nof_chain.cu -- activations are the A operand in tensor memory (tcgen05.mma with A in TMEM, written by
   *      tcgen05.st from the epilogue), 16 epilogue warps per CTA; plans keep x0 available for the skip layer.
   * Plans of the three kinds (0, 1, 2) are not interchangeable. */
  uint32_t wpack_bytes;
  int32_t resident;
  /* backward, NeRF programs: gradient w.r.t. already-embedded input rows (the reference feeds NoF outputs through
   * Embedding into NeRF.forward(inputs), trainer/trainer_moco_flow.py:146-158,349-361): when non-NULL the dX rounds
   * write d_dense[m*d_dense_stride + c], c < dense_cols, instead of applying the encoder's Jacobian. */
  float* d_dense;
  int32_t d_dense_stride;
  int32_t reserved0;       /* chain.cu kernels: non-zero = signal the next layer's MMA before the training-save bulk store */
} mcf_chain_params_t;

int mcf_chain_launch(const mcf_chain_params_t* params_host, cudaStream_t stream);

/* dW[i][j] (+)= sum_rows P[row][i] * Q[row][j] over bf16 tile images (split-K over row tiles, fp32
 * atomics into a zero-initialised or accumulating output): the weight-gradient GEMM of the MLP
 * backward.  P/Q tile records: image(tile) at base + tile*tile_bytes + off; blocks of 64 columns.
 * out[i*ld_out + j]; i < n_i (<= 256), j < n_j (<= 256). */
typedef struct {
  const void* p_base; long long p_tile_bytes; uint32_t p_off; int32_t p_cols;
  const void* q_base; long long q_tile_bytes; uint32_t q_off; int32_t q_cols;
  float* out; int32_t ld_out; int32_t n_i, n_j;
  float* colsum_p;   /* optional: colsum_p[i] += sum_rows P[row][i] (bias gradient) */
  long long n_tiles;
  int32_t max_ctas;
} mcf_dw_params_t;
int mcf_dw_gemm(const mcf_dw_params_t* params_host, cudaStream_t stream);

/* All weight-gradient jobs of one MLP evaluation in ONE launch (grid = n_jobs x CTAs-per-job): removes the
 * launch gaps and lets the tail of one job overlap the head of the next.  Jobs reference the forward / backward
 * save records through a selector so that the (static) job table can stay resident on the device. */
typedef struct {
  uint32_t p_off, q_off;       /* byte offsets inside the tile records                        */
  int32_t p_src, q_src;        /* 0: forward save record, 1: backward save record             */
  int32_t p_cols, q_cols;
  uint32_t st_off;             /* float offset of the [n_i][ld] result inside staging         */
  int32_t ld, n_i, n_j;
  int32_t colsum_off;          /* float offset of colsum_p inside staging, or -1              */
  int32_t enabled;             /* 0: skip (parameter does not require grad)                   */
  int32_t q_split;             /* the first q_split 64-column blocks of Q come from q_src, the rest from the
                                * per-ray feature images (aux + tile*aux_tile_bytes + q2_off); -1: all from q_src */
  uint32_t q2_off;
} mcf_dw_job_t;
int mcf_dw_gemm_batch(const mcf_dw_job_t* jobs_dev, int n_jobs, const void* fwd_save, long long fwd_tile_bytes,
                      const void* bwd_save, long long bwd_tile_bytes, const void* aux, long long aux_tile_bytes,
                      float* staging, long long n_tiles, int ctas_per_job, cudaStream_t stream);

/* Per-ray feature columns (index / direction embedding, constant along a ray: models/rendering.py:73-75,133-142) as
 * bf16 tile images [n_tiles][128 rows][64 cols] -- the Q operand of the weight gradients of the folded input
 * columns.  It depends only on (features, rows_per_ray, n_rows), so one render call builds it once per ray layout
 * and every MLP evaluation of that layout shares it (instead of every chain launch writing its own copy). */
int mcf_rayfeat_image(const float* rayfeat, int rayfeat_stride, int rayfeat_dim, long long n_rows, int rows_per_ray,
                      void* out, cudaStream_t stream);

#define MCF_MAX_UNPACK_PTRS 40
/* Scatter the staging results of mcf_dw_gemm into the parameter-gradient buffer:
 * dst[dst_off + r*dst_ld + c] = transposed ? src[src_off + c*src_ld + r] : src[src_off + r*src_ld + c]. */
typedef struct {
  uint32_t src_off, dst_off;
  int32_t src_ld, dst_ld;
  int32_t nrows, ncols;
  int32_t transposed;
  int32_t reserved;
} mcf_unpack_t;
int mcf_unpack(const mcf_unpack_t* table_dev, int n_entries, const float* staging, float* grads,
               cudaStream_t stream);
/* Same scatter, but entry e ADDS into dst_ptrs_host[e] (+ r*dst_ld + c): accumulates straight into existing
 * parameter .grad tensors (one launch instead of one autograd add kernel per parameter). */
int mcf_unpack_accumulate(const mcf_unpack_t* table_dev, int n_entries, const float* staging,
                          float* const* dst_ptrs_host, cudaStream_t stream);

/* out[c] += sum_m src[m*stride + c], c < ncols (<= 16): bias gradients of the heads */
int mcf_colsum(const float* src, long long n_rows, int stride, int ncols, float* out, cudaStream_t stream);

/* ---- layer-program builders (host code, no CUDA calls) ---------------------------------------- */
/* The tables above for the two network families of the path, from the module shapes alone -- what a host other
 * than the Python shim needs to drive mcf_pack / mcf_chain_launch / mcf_dw_gemm_batch / mcf_unpack
 * (moco_flow_b200/plans.py builds the same tables; tests/test_host_cpu.py compares the two entry by entry).
 * Parameters are identified by canonical ids = their order in nn.Module.named_parameters():
 *   NeRF (models/nerf.py:28-59): 2i / 2i+1 = xyz_encoding_{i+1}.0.weight / .bias (i < D), then 2D.. =
 *        xyz_encoding_final.weight, .bias, extra_encoding.0.weight, .bias, sigma.weight, .bias, rgb.0.weight, .bias;
 *   NoF  (models/nof.py:40-53):  2i / 2i+1 = nof_encoding_{i+1}.0.weight / .bias, 2D / 2D+1 = nof_encoding_final.*. */
#define MCF_PLAN_MAX_PACK 192
#define MCF_PLAN_MAX_CHUNKS 128
#define MCF_PLAN_MAX_ROUNDS 24
#define MCF_PLAN_MAX_LAYERS 17
#define MCF_PLAN_MAX_JOBS 32
typedef struct {
  int32_t family;      /* 0: NeRF, 1: NoF                                                          */
  int32_t D, W, cx;    /* trunk depth, hidden width (128 | 256), in_channels_xyz (<= 64)           */
  int32_t n_skips;
  int32_t skips[8];    /* 0-based trunk layers whose input is cat([input, h])                      */
  int32_t extra_dim;   /* per-ray feature columns (index / direction embedding), 0: none           */
  int32_t use_quat;    /* NoF: 9-wide quaternion head instead of the 3-wide residual head          */
  int32_t sigma_only;  /* NeRF forward: stop after the sigma head                                  */
  int32_t training;    /* forward: also write operand images / masks for the backward pass         */
  int32_t need_dx;     /* backward: include the input-point gradient rounds                        */
  int32_t nof_kernel;  /* NoF: 0 streamed weights, 1 resident (chain.cu), 2 TMEM-resident (nof_chain.cu) */
  int32_t no_pair_merge; /* width 256: 0 = lay each layer's weight tiles out half by half, so that the CTA-pair kernel
                          * fetches one CTA's share of two k-blocks with one 32 KB copy (chunk flag bit 3); 1 = k-block major */
} mcf_plan_spec_t;

typedef struct {
  int32_t n_pack, n_chunks, n_rounds, n_tensors;
  mcf_pack_t pack[MCF_PLAN_MAX_PACK];
  mcf_chunk_t chunks[MCF_PLAN_MAX_CHUNKS];
  mcf_round_t rounds[MCF_PLAN_MAX_ROUNDS];
  int32_t tensor_ids[MCF_MAX_PACK_TENSORS];   /* canonical ids in the order of the pointer array mcf_pack takes */
  uint32_t wpack_bytes, n_consts, save_tile_bytes, mask_tile_words;
  int32_t n_raybias, kind, resident, width;   /* kind / resident / width go into mcf_chain_params_t            */
  /* offsets inside the per-tile save / mask records (0xFFFFFFFF: absent); h / dy are indexed by layer 1..D   */
  uint32_t save_x0, save_feat, save_he, mask_he;
  uint32_t save_h[MCF_PLAN_MAX_LAYERS], mask_h[MCF_PLAN_MAX_LAYERS];
  uint32_t save_dhead, save_dye, save_dyf, save_ghead;
  uint32_t save_dy[MCF_PLAN_MAX_LAYERS];
} mcf_plan_t;

typedef struct {
  int32_t n_jobs, n_unpack, n_params;
  mcf_dw_job_t jobs[MCF_PLAN_MAX_JOBS];
  int32_t job_params[MCF_PLAN_MAX_JOBS][2];   /* canonical ids a job feeds (-1: none): clear `enabled` when neither requires grad */
  mcf_unpack_t unpack[MCF_MAX_UNPACK_PTRS];   /* dst_off = offset in the flat layout below                     */
  int32_t unpack_param[MCF_MAX_UNPACK_PTRS];  /* ... or parameter id + float offset inside it (mcf_unpack_accumulate) */
  uint32_t unpack_inner[MCF_MAX_UNPACK_PTRS];
  uint32_t staging_floats;
  int32_t head_ncols, head_stride;            /* mcf_colsum of the fp32 head gradients -> staging + head_off   */
  uint32_t head_off;
  uint32_t param_offset[MCF_MAX_PACK_TENSORS];/* flat gradient layout: parameters in canonical order, 16-byte aligned */
  uint32_t total_floats;
} mcf_grad_plan_t;

int mcf_plan_forward(const mcf_plan_spec_t* spec, mcf_plan_t* out);
int mcf_plan_backward(const mcf_plan_spec_t* spec, const mcf_plan_t* forward_plan, mcf_plan_t* out);
int mcf_plan_gradients(const mcf_plan_spec_t* spec, const mcf_plan_t* forward_plan, const mcf_plan_t* backward_plan,
                       mcf_grad_plan_t* out);


/* One Adam step over n contiguous fp32 elements (parameters, gradients and both moments are flat device arrays):
 * torch.optim.Adam(lr, betas, eps, weight_decay) as built by trainer/base.py:122-133 and stepped at
 * trainer/trainer_moco_flow.py:406-420.  g = grads*grad_scale (+ weight_decay*p);
 * m += (1-b1)(g-m); v = b2 v + (1-b2) g^2; p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps), t = *step_dev + 1.
 * lr_dev / step_dev are device scalars (a captured CUDA graph keeps working while a scheduler changes the rate);
 * advance_step != 0 increments *step_dev after the update (pass it on the last segment of an optimizer). */
int mcf_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n,
                  const float* lr_dev, long long* step_dev, double beta1, double beta2, float eps, float weight_decay,
                  float grad_scale, int advance_step, cudaStream_t stream);

/* ---- ray generation / canvas scatter ------------------------------------------- */
/* utils/camera.py:29-82,134-148: rows [o(3) d(3) near far img_ind] for n_rays pixels of an H x W pinhole camera.
 * Pixel p = j*W + i has camera direction ((i-cx)/focal, -(j-cy)/focal, -1) (no half-pixel offset, one focal for both
 * axes as in the reference), rotated by c2w (host pointer to a row-major [3][4]; NULL = camera frame, origin 0) and
 * normalised.  pixel_index (device, int64) selects pixels -- the valid-ray mask gather of
 * trainer/trainer_moco_flow.py:229-232 -- or NULL for pixels 0..n_rays-1. */
int mcf_make_rays(int H, int W, float focal, float cx, float cy, const float* c2w_host, float near, float far,
                  float img_ind, const long long* pixel_index, long long n_rays, float* rays, int ray_stride,
                  cudaStream_t stream);
/* trainer/trainer_moco_flow.py:247-262: img_out[n_pixels][3] = background, depth_out = 10; the n_rays rendered
 * pixels (pixel_index, or 0..n_rays-1) get depth 8, and where opacity > 0 the rendered rgb / depth. */
int mcf_canvas_scatter(const float* background, long long n_pixels, const long long* pixel_index, long long n_rays,
                       const float* rgb, const float* depth, const float* opacity, float* img_out, float* depth_out,
                       cudaStream_t stream);

/* ---- SMPL correspondence sampling ---------------------------------------------- */
/* datasets/moco_flow_dataset.py:121-130: for every query point the nearest of n_verts vertices (what
 * knn_cuda.KNN(k=1, transpose_mode=True) returns: Euclidean distance and index; first minimum on ties),
 * inside[q] = dist < thickness, and cano[q] = (trans[ind[q]] @ [x y z 1])[:3] with trans [n_verts][4][4] row-major.
 * Any of dist / ind / cano / inside may be NULL (trans may be NULL when cano is). */
int mcf_nearest_vertex(const float* verts, int n_verts, const float* trans, const float* query, long long n_query,
                       float thickness, float* dist, long long* ind, float* cano, unsigned char* inside,
                       cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MOCO_FLOW_B200_H_ */
