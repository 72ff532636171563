/* Driving the canonical NeRF forward (models/nerf.py:61-102 with the encoder of models/embedding.py:42-46 fused in)
 * from plain C through the C ABI -- no Python, no plans.py.  Device buffers are the caller's (allocate them with the
 * CUDA runtime of your host); this file only shows the call sequence and is compiled (not run) by
 * tests/test_host_cpu.py::test_c_example_compiles.
 *
 *   gcc -c -Iinclude examples/nerf_forward.c
 */
#include <string.h>

#include "moco_flow_b200.h"

/* tensors_dev[id]: device pointers of the 24 NeRF parameters in canonical (state_dict) order. */
int nerf_forward(const float* const* tensors_dev, const float* xyz_dev, const float* ray_feat_dev, int n_rays,
                 int samples_per_ray, void* pack_dev, void* chunks_dev, void* rounds_dev, void* wpack_dev,
                 float* consts_dev, float* raybias_dev, float* rgbsigma_dev, cudaStream_t stream,
                 mcf_plan_t* plan /* out: the caller copies plan->pack / chunks / rounds to *_dev before the launches */) {
  mcf_plan_spec_t spec;
  memset(&spec, 0, sizeof(spec));
  spec.family = 0; spec.D = 8; spec.W = 256; spec.cx = 63; spec.n_skips = 1; spec.skips[0] = 4; spec.extra_dim = 5;
  int rc = mcf_plan_forward(&spec, plan);
  if (rc) return rc;

  /* pointer array in the order the pack table expects */
  const float* ordered[MCF_MAX_PACK_TENSORS];
  for (int i = 0; i < plan->n_tensors; ++i) ordered[i] = tensors_dev[plan->tensor_ids[i]];
  rc = mcf_pack((const mcf_pack_t*)pack_dev, plan->n_pack, ordered, plan->n_tensors, wpack_dev, consts_dev, stream);
  if (rc) return rc;

  /* extra_encoding sees cat([feat, index embedding]): the per-ray columns are folded into a per-ray bias */
  const float* extra_w = tensors_dev[2 * spec.D + 2];
  const float* extra_b = tensors_dev[2 * spec.D + 3];
  rc = mcf_ray_bias(extra_w, spec.W + spec.extra_dim, spec.W, extra_b, ray_feat_dev, spec.extra_dim, spec.extra_dim, n_rays,
                    spec.W / 2, raybias_dev, stream);
  if (rc) return rc;

  mcf_chain_params_t cp;
  memset(&cp, 0, sizeof(cp));
  cp.chunks = (const mcf_chunk_t*)chunks_dev; cp.rounds = (const mcf_round_t*)rounds_dev;
  cp.n_chunks = plan->n_chunks; cp.n_rounds = plan->n_rounds;
  cp.width = plan->width; cp.program_kind = plan->kind; cp.resident = plan->resident; cp.wpack_bytes = plan->wpack_bytes;
  cp.wpack = wpack_dev; cp.consts = consts_dev; cp.raybias[0] = raybias_dev; cp.cta_pair = 1;
  cp.prologue = MCF_PRO_PE_XYZ; cp.xyz = xyz_dev; cp.pe_n_freqs = 10; cp.pe_pad_to = spec.cx;
  for (int k = 0; k < 10; ++k) { cp.pe_freq[k] = (float)(1 << k); cp.pe_weight[k] = 1.0f; }
  cp.n_rows = (long long)n_rays * samples_per_ray; cp.rows_per_ray = samples_per_ray; cp.n_rays = n_rays;
  cp.out = rgbsigma_dev; cp.out_stride = 4; cp.sigma_col = 3;
  cp.x0_save_off = cp.fwd_x0_off = cp.fwd_he_off = cp.extra_save_off = cp.dhead_save_off = 0xFFFFFFFFu;
  return mcf_chain_launch(&cp, stream);
}
