"""Ray rendering with the reference's call signatures (models/rendering.py:5-375) on the CUDA library.

``sample_pdf``, ``nof_inference``, ``nerf_inference`` and ``render_rays`` take the same arguments and
return the same structures as the reference.  Differences, all opt-in keyword arguments that default
to the reference behaviour:
  * ``draws``: inject the random tensors (perturb / noise / u) instead of drawing them, for parity tests;
  * ``fused_residual_mean``: return the flow-consistency residuals as (1,) masked means without the
    host synchronisation of the reference's ``torch.any`` + boolean indexing.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch

from . import ops


class Draws:
    """Random tensors of one render_rays call, in the reference's draw order (rendering.py:259,166,30,166)."""

    def __init__(self, perturb=None, noise_coarse=None, u=None, noise_fine=None):
        self.perturb, self.noise_coarse, self.u, self.noise_fine = perturb, noise_coarse, u, noise_fine


def sample_pdf(bins, weights, N_importance, det=False, eps=1e-5, u=None):
    """Inverse-CDF sampling, rendering.py:5-46.  bins (R, n+1), weights (R, n) -> (R, N_importance)."""
    R = weights.shape[0]
    if det:
        u = torch.linspace(0, 1, N_importance, device=bins.device).expand(R, N_importance)
    elif u is None:
        u = torch.rand(R, N_importance, device=bins.device)
    samples, _, _, _ = ops.sample_pdf_raw(bins, weights.detach(), u.contiguous(), eps=eps)
    return samples


def _ray_embedding(embedding, values):
    """Per-ray embedded feature (R, C) -- evaluated once per ray, never repeated per sample, and once per rendering
    call for the same (embedding, values) pair."""
    from .mlp import memoised, _tkey
    key = ("emb", id(embedding), tuple(float(w) for w in getattr(embedding, "weights", ())), _tkey(values))
    return memoised(key, (values, embedding), lambda: embedding(values.detach().contiguous()))


def nof_inference(xyz_, ind_, nof_embeddings, nof_model):
    """rendering.py:49-83: xyz_ (R,S,3), ind_ (R,1) -> warped (R,S,3)."""
    R, S = xyz_.shape[0], xyz_.shape[1]
    ind_feat = _ray_embedding(nof_embeddings[1], ind_)
    out = nof_model.evaluate(xyz=xyz_.reshape(-1, 3), pe=nof_embeddings[0], ray_feat=ind_feat, rows_per_ray=S)
    return out.view(R, S, -1)


def _nerf_pass(xyz_, ind_, dir_, z_vals, noise_std, nerf_embeddings, nerf_model, background, weights_only,
               activate_type, noise):
    """One NeRF evaluation + compositing; returns (rgb, depth, weights, alphas, opacity) (rgb/depth None if weights_only)."""
    if activate_type not in ('relu', 'softplus'):
        raise ValueError('activation layer type: %s not support' % activate_type)
    R, S = xyz_.shape[0], xyz_.shape[1]
    dir_ = dir_.reshape(-1, 3)
    feat = None
    if not weights_only:
        if nerf_model.extra_feat_type == 'ind':
            feat = _ray_embedding(nerf_embeddings[1], ind_)
        elif nerf_model.extra_feat_type == 'dir':
            feat = _ray_embedding(nerf_embeddings[2], dir_)
        if feat is not None and feat.shape[1] > nerf_model.extra_feat_dim:
            raise RuntimeError("extra-feature embedding wider than the model's extra_feat_dim")
    raw = nerf_model.evaluate(xyz=xyz_.reshape(-1, 3), pe=nerf_embeddings[0], ray_feat=feat, rows_per_ray=S,
                              sigma_only=weights_only)
    if noise is None:  # drawn even when noise_std == 0, as the reference does (rendering.py:166)
        noise = torch.randn(R, S, device=z_vals.device)
    if weights_only:
        weights, alphas, opacity = ops.composite(raw.view(R, S), z_vals, dir_, noise, noise_std, None, activate_type)
        return None, None, weights, alphas, opacity
    return ops.composite(raw.view(R, S, 4), z_vals, dir_, noise, noise_std, background, activate_type)


def nerf_inference(xyz_, ind_, dir_, z_vals, noise_std, nerf_embeddings, nerf_model, background=None,
                   weights_only=False, activate_type='relu', noise=None):
    """rendering.py:86-192.  Returns (weights, alphas) or (rgb, depth, weights, alphas)."""
    rgb, depth, weights, alphas, _ = _nerf_pass(xyz_, ind_, dir_, z_vals, noise_std, nerf_embeddings, nerf_model,
                                                background, weights_only, activate_type, noise)
    if weights_only:
        return weights, alphas
    return rgb, depth, weights, alphas


def render_rays(rays, background, nerf_embeddings, nerf_models, nof_embeddings=None, nof_models=None,
                chain_local=False, chain_global=False, N_samples=64, N_importance=0, use_disp=False, perturb=0,
                noise_std=1, nerf_activate_type='relu', test_time=False, draws: Optional[Draws] = None,
                fused_residual_mean: bool = False) -> Dict[str, torch.Tensor]:
    """rendering.py:195-375."""
    from .mlp import call_memo
    with call_memo():
        result = _render_rays(rays, background, nerf_embeddings, nerf_models, nof_embeddings, nof_models, chain_local,
                              chain_global, N_samples, N_importance, use_disp, perturb, noise_std, nerf_activate_type,
                              test_time, draws, fused_residual_mean)
    if ops.DEBUG_SYNC and not torch.cuda.is_current_stream_capturing():
        ops.check_device()
    return result


def _render_rays(rays, background, nerf_embeddings, nerf_models, nof_embeddings, nof_models, chain_local, chain_global,
                 N_samples, N_importance, use_disp, perturb, noise_std, nerf_activate_type, test_time, draws,
                 fused_residual_mean):
    draws = draws or Draws()
    rays = rays.contiguous()
    R = rays.shape[0]
    dev = rays.device
    rays_d = rays[:, 3:6]
    img_ind = rays[:, 8:9]
    use_nof = nof_models is not None
    chained_ind = rays[:, 9:10] if (use_nof and chain_global) else None

    perturb_rand = None
    if perturb > 0:
        perturb_rand = draws.perturb if draws.perturb is not None else torch.rand(R, N_samples, device=dev)
    z_vals, xyz_coarse = ops.coarse_samples(rays, N_samples, float(perturb), perturb_rand, use_disp)

    def flow_chain(x_obs):
        bw = nof_models[0]
        x_can = nof_inference(x_obs, img_ind, nof_embeddings, bw)
        x_loc = x_glob = None
        if not test_time and (chain_local or chain_global):
            fw = nof_models[1]
            if chain_local:
                x_loc = nof_inference(x_can, img_ind, nof_embeddings, fw)
            if chain_global:
                x1 = nof_inference(x_can, chained_ind, nof_embeddings, fw)
                x2 = nof_inference(x1, chained_ind, nof_embeddings, bw)
                x_glob = nof_inference(x2, img_ind, nof_embeddings, fw)
        return x_can, x_loc, x_glob

    # fused form: statistics now, ONE all-reduce over the data-parallel ranks and the means at the end of the call
    fused = ops.ResidualMeans(dev) if fused_residual_mean else None

    def residuals(result, tag, x_obs, x_loc, x_glob, alphas):
        for on, name, x_rec in ((chain_local, 'nof_local_disp_', x_loc), (chain_global, 'nof_global_disp_', x_glob)):
            if not on:
                continue
            if fused is not None:
                result[name + tag] = None       # keeps the reference's key order; filled by fused.finish
                fused.add(name + tag, x_obs, x_rec, alphas)
            else:
                result[name + tag] = ops.flow_residual(x_obs, x_rec, alphas, False)

    if use_nof:
        nerf_in, x_loc, x_glob = flow_chain(xyz_coarse)
    else:
        nerf_in = xyz_coarse

    coarse_only_weights = N_importance > 0 and test_time
    rgb_coarse, depth_coarse, weights_coarse, alphas_coarse, opacity_coarse = _nerf_pass(
        nerf_in, img_ind, rays_d, z_vals, noise_std, nerf_embeddings, nerf_models[0], background,
        coarse_only_weights, nerf_activate_type, draws.noise_coarse)
    if coarse_only_weights:
        result = {'opacity_coarse': opacity_coarse}
    else:
        result = {'rgb_coarse': rgb_coarse, 'depth_coarse': depth_coarse, 'opacity_coarse': opacity_coarse}

    if use_nof and not test_time:
        residuals(result, 'coarse', xyz_coarse, x_loc, x_glob, alphas_coarse)

    if N_importance > 0:
        if perturb == 0:
            u = torch.linspace(0, 1, N_importance, device=dev).expand(R, N_importance).contiguous()
        else:
            u = draws.u if draws.u is not None else torch.rand(R, N_importance, device=dev)
        # bins = mid-points of the coarse depths, weights = weights_coarse[:, 1:-1]; merged and sorted in-kernel
        _, _, _, z_fine = ops.sample_pdf_raw(z_vals, weights_coarse.detach(), u, bins_are_z=True, w_offset=1,
                                             n_bins=N_samples - 2, z_coarse=z_vals, want_samples=False)
        xyz_fine = ops.ray_points(rays, z_fine)
        if use_nof:
            nerf_in_f, x_loc_f, x_glob_f = flow_chain(xyz_fine)
        else:
            nerf_in_f = xyz_fine
        rgb_fine, depth_fine, weights_fine, alphas_fine, opacity_fine = _nerf_pass(
            nerf_in_f, img_ind, rays_d, z_fine, noise_std, nerf_embeddings, nerf_models[1], background, False,
            nerf_activate_type, draws.noise_fine)
        result['rgb_fine'] = rgb_fine
        result['depth_fine'] = depth_fine
        result['opacity_fine'] = opacity_fine
        if use_nof and not test_time:
            residuals(result, 'fine', xyz_fine, x_loc_f, x_glob_f, alphas_fine)
    if fused is not None:
        fused.finish(result)
    return result
