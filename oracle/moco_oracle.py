"""CPU restatement (torch fp32, differentiable) of MoCo-Flow's ray-rendering path.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  This file restates the
algorithm of the reference's ``models/{embedding,nerf,nof,rendering}.py``; every
function cites the reference lines it follows (paths relative to the reference
root).  It is written against plain state-dicts and explicit random tensors so
that the CUDA path and the oracle can be fed *identical* weights and draws.

Parity status: PINNED.  ``tests/golden/make_golden.py`` runs the reference's own
modules (imported from ``/root/reference`` in the build container) on seeded
inputs and stores inputs + outputs under ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks this restatement against those files.
The one piece that cannot be pinned against reference-side code is kornia's
quaternion pair (kornia==0.6.5 is a pip dependency of the reference,
``docker/requirements.txt:16``, not vendored and not installed here): it is
restated from the published kornia 0.6.x algorithm in ``quat_log_to_exp`` /
``quat_to_rotmat`` and cross-checked against an independent Rodrigues formula
(rotation by 2*|v| about v/|v|) in the tests.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------
# shapes
# --------------------------------------------------------------------------
@dataclass
class PESpec:
    """models/embedding.py:5-21 (ctor state of ``Embedding``)."""
    in_channels: int
    n_freqs: int
    logscale: bool = True
    weights: Optional[Sequence[float]] = None  # per-frequency multipliers (:16, :23-28)

    @property
    def out_channels(self) -> int:  # :14
        return self.in_channels * (2 * self.n_freqs + 1)

    def freqs(self) -> List[float]:  # :18-21
        n = self.n_freqs
        if n == 0:
            return []
        if self.logscale:
            return [float(2.0 ** k) for k in torch.linspace(0, n - 1, n).tolist()]
        return torch.linspace(1, 2 ** (n - 1), n).tolist()

    def w(self) -> List[float]:
        return [1.0] * self.n_freqs if self.weights is None else [float(v) for v in self.weights]


@dataclass
class NeRFSpec:
    """models/nerf.py:6-59."""
    D: int = 8
    W: int = 256
    in_channels_xyz: int = 63
    skips: Sequence[int] = (4,)
    extra_feat_type: str = "ind"
    extra_feat_dim: int = 5


@dataclass
class NoFSpec:
    """models/nof.py:7-53."""
    D: int = 4
    W: int = 128
    in_channels_xyz: int = 33
    skips: Sequence[int] = (2,)
    extra_feat_type: str = "ind"
    extra_feat_dim: int = 33
    use_quat: bool = True


# --------------------------------------------------------------------------
# a1 positional encoding
# --------------------------------------------------------------------------
def positional_encoding(x: Tensor, spec: PESpec) -> Tensor:
    """models/embedding.py:42-46 -- [x, w0 sin(f0 x), w0 cos(f0 x), w1 sin(f1 x), ...]."""
    pieces = [x]
    for wk, fk in zip(spec.w(), spec.freqs()):
        arg = x * fk
        pieces.append(torch.sin(arg) * wk)
        pieces.append(torch.cos(arg) * wk)
    return torch.cat(pieces, dim=-1)


def _padded(feat: Tensor, width: int) -> Tensor:
    """Right zero-pad to ``width`` columns (models/rendering.py:70-72,127-129,135-136,140-141)."""
    if feat.shape[1] > width:
        raise RuntimeError("embedding wider than the model's input width")  # the reference's slice-assign would raise
    return F.pad(feat, (0, width - feat.shape[1]))


# --------------------------------------------------------------------------
# bf16 tensor-core emulation (checker for the CUDA path's arithmetic, not reference behaviour)
# --------------------------------------------------------------------------
# When EMULATE_BF16 is set, every Linear that the CUDA kernels run on tensor cores rounds its two
# operands to bf16 (round-to-nearest-even) and accumulates in fp32, exactly the arithmetic of
# tcgen05 kind::f16 with fp32 accumulators.  Columns the kernels keep in fp32 (per-ray folded
# features, the sigma / rgb heads, biases) stay fp32.  With the flag off (default) the functions
# below are the reference's fp32 algorithm.
EMULATE_BF16 = False
# Sensitivity probe: multiply every value by (1 + EMULATE_NOISE * N(0,1)) BEFORE it is rounded to bf16.  1e-6 is the
# size of an fp32 accumulation-order difference; two evaluations that differ only by that (this oracle vs a tensor
# core, or two tensor-core kernels with different K-splits) round a small fraction of activations to different bf16
# neighbours.  The tests use it to calibrate how far two faithful bf16 evaluations may be apart.
EMULATE_NOISE = 0.0
_NOISE_GEN = torch.Generator().manual_seed(20240229)


def _r16(t: Tensor) -> Tensor:
    if not EMULATE_BF16:
        return t
    src = t
    if EMULATE_NOISE:
        src = t.detach() * (1 + EMULATE_NOISE * torch.randn(t.shape, generator=_NOISE_GEN))
    return t + (src.to(torch.bfloat16).to(torch.float32) - t).detach()  # rounded value, straight-through gradient


def _tc_linear(x: Tensor, w: Tensor, b: Optional[Tensor], fp32_cols: Optional[slice] = None) -> Tensor:
    """F.linear with tensor-core emulation; ``fp32_cols`` = input columns kept in fp32 (folded per-ray features)."""
    if not EMULATE_BF16:
        return F.linear(x, w, b)
    if fp32_cols is None:
        return F.linear(_r16(x), _r16(w), b)
    keep = torch.zeros(x.shape[1], dtype=torch.bool)
    keep[fp32_cols] = True
    out = F.linear(_r16(x[:, ~keep]), _r16(w[:, ~keep]), b)
    return out + F.linear(x[:, keep], w[:, keep])


# --------------------------------------------------------------------------
# a2 NeRF MLP
# --------------------------------------------------------------------------
def nerf_mlp(p: Dict[str, Tensor], spec: NeRFSpec, inputs: Tensor, sigma_only: bool = False) -> Tensor:
    """models/nerf.py:61-102.  ``p`` uses the reference's state_dict names."""
    cx = spec.in_channels_xyz
    if sigma_only:
        pts = inputs
        extra = None
    else:
        pts, extra = inputs[:, :cx], inputs[:, cx:cx + spec.extra_feat_dim]  # :79
    h = pts
    for i in range(spec.D):  # :84-87
        if i in spec.skips:
            h = torch.cat([pts, h], dim=-1)
        h = F.relu(_tc_linear(h, p[f"xyz_encoding_{i+1}.0.weight"], p[f"xyz_encoding_{i+1}.0.bias"]))
    sigma = F.linear(h, p["sigma.weight"], p["sigma.bias"])  # :89 (fp32 head in the CUDA path too)
    if sigma_only:
        return sigma  # :90-91
    feat = _tc_linear(h, p["xyz_encoding_final.weight"], p["xyz_encoding_final.bias"])  # :93
    if spec.extra_feat_type == "latent_code":
        raise NotImplementedError("NeRF model does not support latent code yet!!!")  # :95
    W = feat.shape[1]
    if spec.extra_feat_dim > 0:
        feat = torch.cat([feat, extra], dim=-1)  # :98
    half = F.relu(_tc_linear(feat, p["extra_encoding.0.weight"], p["extra_encoding.0.bias"],
                             slice(W, W + spec.extra_feat_dim)))
    rgb = torch.sigmoid(F.linear(half, p["rgb.0.weight"], p["rgb.0.bias"]))  # :99
    return torch.cat([rgb, sigma], dim=-1)  # :101


# --------------------------------------------------------------------------
# kornia 0.6.5 pair used at models/nof.py:78-79 (restated; not on disk)
# --------------------------------------------------------------------------
def quat_log_to_exp(v: Tensor, eps: float = 1e-8) -> Tensor:
    """kornia.geometry.conversions.quaternion_log_to_exp (0.6.x, XYZW): q = [v sin|v|/|v|, cos|v|]."""
    n = torch.linalg.vector_norm(v, dim=-1, keepdim=True).clamp(min=eps)
    return torch.cat([v * torch.sin(n) / n, torch.cos(n)], dim=-1)


def quat_to_rotmat(q: Tensor) -> Tensor:
    """kornia.geometry.conversions.quaternion_to_rotation_matrix (0.6.x, XYZW, normalises first)."""
    q = F.normalize(q, p=2.0, dim=-1, eps=1e-12)
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    tx, ty, tz = 2.0 * x, 2.0 * y, 2.0 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    one = torch.ones_like(x)
    m = torch.stack([one - (tyy + tzz), txy - twz, txz + twy,
                     txy + twz, one - (txx + tzz), tyz - twx,
                     txz - twy, tyz + twx, one - (txx + tyy)], dim=-1)
    return m.view(-1, 3, 3)


def rodrigues_rotmat(v: Tensor) -> Tensor:
    """Independent cross-check: rotation by angle 2|v| about axis v/|v| (fp64 recommended)."""
    n = torch.linalg.vector_norm(v, dim=-1, keepdim=True).clamp(min=1e-30)
    k = v / n
    th = 2.0 * n[..., 0]
    K = torch.zeros(v.shape[0], 3, 3, dtype=v.dtype)
    K[:, 0, 1], K[:, 0, 2] = -k[:, 2], k[:, 1]
    K[:, 1, 0], K[:, 1, 2] = k[:, 2], -k[:, 0]
    K[:, 2, 0], K[:, 2, 1] = -k[:, 1], k[:, 0]
    eye = torch.eye(3, dtype=v.dtype).expand(v.shape[0], 3, 3)
    s, c = torch.sin(th)[:, None, None], torch.cos(th)[:, None, None]
    return eye + s * K + (1.0 - c) * (K @ K)


# --------------------------------------------------------------------------
# a3 NoF MLP
# --------------------------------------------------------------------------
def nof_mlp(p: Dict[str, Tensor], spec: NoFSpec, inputs: Tensor, xyz: Tensor) -> Tensor:
    """models/nof.py:55-85."""
    if spec.extra_feat_type == "latent_code":
        raise NotImplementedError("NoF model does not support latent code yet!!!")  # :65
    h = inputs
    cx, ce = spec.in_channels_xyz, spec.extra_feat_dim
    for i in range(spec.D):  # :69-73
        folded = None
        if i in spec.skips:
            h = torch.cat([inputs, h], dim=-1)
        if i == 0 or i in spec.skips:
            folded = slice(cx, cx + ce)
        h = F.relu(_tc_linear(h, p[f"nof_encoding_{i+1}.0.weight"], p[f"nof_encoding_{i+1}.0.bias"], folded))
    head = _tc_linear(h, p["nof_encoding_final.weight"], p["nof_encoding_final.bias"])
    if not spec.use_quat:
        return head + xyz  # :82
    v, s, t = head[:, 0:3], head[:, 3:6], head[:, 6:9]  # :77
    rot = quat_to_rotmat(quat_log_to_exp(v))  # :78-79
    rel = (xyz - s).unsqueeze(1)  # row vector times R, :80
    return torch.bmm(rel, rot).squeeze(1) + s + t


# --------------------------------------------------------------------------
# a4 sample_pdf
# --------------------------------------------------------------------------
def sample_pdf(bins: Tensor, weights: Tensor, n_importance: int, det: bool = False,
               eps: float = 1e-5, u: Optional[Tensor] = None, return_aux: bool = False):
    """models/rendering.py:5-46.  ``u`` injects the uniform draws of :30."""
    n_rays, n_bins = weights.shape
    wts = weights + eps  # :20
    pdf = wts / wts.sum(dim=-1, keepdim=True)  # :21
    cdf = torch.cumsum(pdf, dim=-1)  # :22
    cdf = torch.cat([torch.zeros_like(cdf[:, :1]), cdf], dim=-1)  # :23
    if det:
        u = torch.linspace(0, 1, n_importance, device=bins.device).expand(n_rays, n_importance)  # :27-28
    elif u is None:
        u = torch.rand(n_rays, n_importance, device=bins.device)  # :30
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, right=True)  # :33
    below = (inds - 1).clamp(min=0)  # :34
    above = inds.clamp(max=n_bins)  # :35
    cdf_lo, cdf_hi = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)  # :37-39
    bin_lo, bin_hi = torch.gather(bins, 1, below), torch.gather(bins, 1, above)
    denom = cdf_hi - cdf_lo
    denom = torch.where(denom < eps, torch.ones_like(denom), denom)  # :41-42
    samples = bin_lo + (u - cdf_lo) / denom * (bin_hi - bin_lo)  # :45
    if return_aux:
        return samples, dict(cdf=cdf, u=u, inds=inds, below=below, above=above)
    return samples


# --------------------------------------------------------------------------
# alpha compositing part of a6
# --------------------------------------------------------------------------
def composite(sigmas: Tensor, rgbs: Optional[Tensor], z_vals: Tensor, dirs: Tensor,
              noise: Optional[Tensor], background: Optional[Tensor], activate_type: str = "relu"):
    """models/rendering.py:158-190.  ``noise`` is the already scaled randn*noise_std of :166."""
    deltas = z_vals[:, 1:] - z_vals[:, :-1]  # :158
    deltas = torch.cat([deltas, torch.full_like(deltas[:, :1], 1e10)], dim=-1)  # :159-160
    deltas = deltas * torch.linalg.vector_norm(dirs.unsqueeze(1), dim=-1)  # :164
    raw = sigmas if noise is None else sigmas + noise
    if activate_type == "relu":
        dens = torch.relu(raw)  # :170
    elif activate_type == "softplus":
        dens = F.softplus(raw)  # :172 (beta=1, threshold=20)
    else:
        raise ValueError("activation layer type: %s not support" % activate_type)  # :174
    alphas = 1 - torch.exp(-deltas * dens)
    shifted = torch.cat([torch.ones_like(alphas[:, :1]), 1 - alphas + 1e-10], dim=-1)  # :176-177
    weights = alphas * torch.cumprod(shifted, dim=-1)[:, :-1]  # :178-179
    wsum = weights.sum(dim=1)  # :180
    if rgbs is None:
        return None, None, weights, alphas
    rgb = (weights.unsqueeze(-1) * rgbs).sum(dim=-2)  # :186
    depth = (weights * z_vals).sum(dim=-1)  # :187
    if background is not None:
        rgb = rgb + background * (1 - wsum.unsqueeze(-1))  # :189-190
    return rgb, depth, weights, alphas


# --------------------------------------------------------------------------
# a5 / a6 inference wrappers
# --------------------------------------------------------------------------
@dataclass
class NoFBundle:
    spec: NoFSpec
    params: Dict[str, Tensor]


@dataclass
class NeRFBundle:
    spec: NeRFSpec
    params: Dict[str, Tensor]


def nof_inference(xyz: Tensor, ind: Tensor, pe_xyz: PESpec, pe_ind: PESpec, nof: NoFBundle) -> Tensor:
    """models/rendering.py:49-83."""
    n_rays, n_samp = xyz.shape[0], xyz.shape[1]
    flat = xyz.reshape(-1, 3)
    emb = _padded(positional_encoding(flat, pe_xyz), nof.spec.in_channels_xyz)  # :70-72
    ind_emb = positional_encoding(ind, pe_ind).repeat_interleave(n_samp, dim=0)  # :73-74
    out = nof_mlp(nof.params, nof.spec, torch.cat([emb, ind_emb], dim=-1), flat)  # :75,81
    return out.view(n_rays, n_samp, -1)


def nerf_inference(xyz: Tensor, ind: Tensor, dirs: Tensor, z_vals: Tensor, noise: Optional[Tensor],
                   pes: Sequence[Optional[PESpec]], nerf: NeRFBundle, background: Optional[Tensor] = None,
                   weights_only: bool = False, activate_type: str = "relu"):
    """models/rendering.py:86-192.  ``pes`` = [xyz, ind, dir] specs (entries may be None)."""
    n_rays, n_samp = xyz.shape[0], xyz.shape[1]
    spec = nerf.spec
    flat = xyz.reshape(-1, 3)
    feats = _padded(positional_encoding(flat, pes[0]), spec.in_channels_xyz)  # :126-130
    if not weights_only:
        if spec.extra_feat_type == "ind":  # :133-137
            e = positional_encoding(ind, pes[1]).repeat_interleave(n_samp, dim=0)
            feats = torch.cat([feats, _padded(e, spec.extra_feat_dim)], dim=1)
        elif spec.extra_feat_type == "dir":  # :138-142
            e = positional_encoding(dirs.reshape(-1, 3), pes[2]).repeat_interleave(n_samp, dim=0)
            feats = torch.cat([feats, _padded(e, spec.extra_feat_dim)], dim=1)
    out = nerf_mlp(nerf.params, spec, feats, sigma_only=weights_only)  # :148
    if weights_only:
        sig, rgbs = out.view(n_rays, n_samp), None  # :151
    else:
        out = out.view(n_rays, n_samp, 4)
        rgbs, sig = out[..., :3], out[..., 3]  # :153-155
    rgb, depth, weights, alphas = composite(sig, rgbs, z_vals, dirs.reshape(-1, 3), noise, background,
                                            activate_type)
    if weights_only:
        return weights, alphas  # :182-183
    return rgb, depth, weights, alphas


# --------------------------------------------------------------------------
# a7 / a8 render_rays
# --------------------------------------------------------------------------
@dataclass
class RenderDraws:
    """The four random tensors one render_rays call consumes, in draw order."""
    perturb: Optional[Tensor] = None       # U[0,1) (R,Sc)        rendering.py:259
    noise_coarse: Optional[Tensor] = None  # N(0,1) (R,Sc)        rendering.py:166 (coarse call)
    u: Optional[Tensor] = None             # U[0,1) (R,Sf)        rendering.py:30
    noise_fine: Optional[Tensor] = None    # N(0,1) (R,Sc+Sf)     rendering.py:166 (fine call)


def _masked_residual(a: Tensor, b: Tensor, alphas: Tensor) -> Tensor:
    """models/rendering.py:306-311 / :365-370."""
    mask = alphas >= 0.01
    if not bool(mask.any()):
        mask = torch.ones_like(mask)
    return (a - b).abs()[mask].mean(dim=1)


def render_rays(rays: Tensor, background: Tensor,
                nerf_pes: Sequence[Optional[PESpec]], nerfs: Sequence[NeRFBundle],
                nof_pes: Optional[Sequence[PESpec]] = None, nofs: Optional[Sequence[NoFBundle]] = None,
                chain_local: bool = False, chain_global: bool = False,
                N_samples: int = 64, N_importance: int = 0, use_disp: bool = False,
                perturb: float = 0, noise_std: float = 1, nerf_activate_type: str = "relu",
                test_time: bool = False, draws: Optional[RenderDraws] = None,
                return_aux: bool = False) -> Dict[str, Tensor]:
    """models/rendering.py:195-375.  Missing entries of ``draws`` are drawn with torch's global RNG
    in the reference's order, so with ``draws=None`` this is a behavioural twin of the reference."""
    draws = draws or RenderDraws()
    aux: Dict[str, Tensor] = {}
    n_rays = rays.shape[0]
    o, d = rays[:, 0:3], rays[:, 3:6]  # :238
    near, far = rays[:, 6:7], rays[:, 7:8]  # :239
    ind = rays[:, 8:9]  # :240
    use_nof = nofs is not None
    ind_chain = rays[:, 9:10] if (use_nof and chain_global) else None  # :241-242

    t = torch.linspace(0, 1, N_samples)  # :245
    if not use_disp:
        z = near * (1 - t) + far * t  # :247
    else:
        z = 1 / (1 / near * (1 - t) + 1 / far * t)  # :249
    z = z.expand(n_rays, N_samples)
    if perturb > 0:  # :253-260
        mid = 0.5 * (z[:, :-1] + z[:, 1:])
        hi = torch.cat([mid, z[:, -1:]], dim=-1)
        lo = torch.cat([z[:, :1], mid], dim=-1)
        r = draws.perturb if draws.perturb is not None else torch.rand(z.shape)
        z = lo + (hi - lo) * (perturb * r)

    def pts(zv):  # :262-263, :329-330
        return o.unsqueeze(1) + d.unsqueeze(1) * zv.unsqueeze(2)

    def noise_for(given, shape):  # :166 -- drawn even when noise_std == 0
        base = given if given is not None else torch.randn(shape)
        return base * noise_std

    def flows(x_obs):  # :270-286 / :335-348
        bw, fw = nofs[0], (nofs[1] if len(nofs) > 1 else None)
        x_can = nof_inference(x_obs, ind, nof_pes[0], nof_pes[1], bw)
        x_loc = x_glob = None
        if chain_local and not test_time:
            x_loc = nof_inference(x_can, ind, nof_pes[0], nof_pes[1], fw)
        if chain_global and not test_time:
            x1 = nof_inference(x_can, ind_chain, nof_pes[0], nof_pes[1], fw)
            x2 = nof_inference(x1, ind_chain, nof_pes[0], nof_pes[1], bw)
            x_glob = nof_inference(x2, ind, nof_pes[0], nof_pes[1], fw)
        return x_can, x_loc, x_glob

    x_c = pts(z)
    if use_nof:
        xin_c, xloc_c, xglob_c = flows(x_c)
    else:
        xin_c = x_c

    result: Dict[str, Tensor] = {}
    if N_importance > 0 and test_time:  # :290-294
        w_c, a_c = nerf_inference(xin_c, ind, d, z, noise_for(draws.noise_coarse, z.shape), nerf_pes, nerfs[0],
                                  background=background, weights_only=True, activate_type=nerf_activate_type)
        result["opacity_coarse"] = w_c.sum(1)
    else:  # :296-302
        rgb_c, dep_c, w_c, a_c = nerf_inference(xin_c, ind, d, z, noise_for(draws.noise_coarse, z.shape), nerf_pes,
                                                nerfs[0], background=background, weights_only=False,
                                                activate_type=nerf_activate_type)
        result.update(rgb_coarse=rgb_c, depth_coarse=dep_c, opacity_coarse=w_c.sum(1))
    aux.update(z_coarse=z, weights_coarse=w_c, alphas_coarse=a_c)

    if use_nof and not test_time:  # :304-314
        if chain_local:
            result["nof_local_disp_coarse"] = _masked_residual(x_c, xloc_c, a_c)
        if chain_global:
            result["nof_global_disp_coarse"] = _masked_residual(x_c, xglob_c, a_c)

    if N_importance > 0:  # :317-373
        mid = 0.5 * (z[:, :-1] + z[:, 1:])
        z_new = sample_pdf(mid, w_c[:, 1:-1], N_importance, det=(perturb == 0), u=draws.u).detach()  # :322-323
        z_f, _ = torch.sort(torch.cat([z, z_new], dim=-1), dim=-1)  # :326
        x_f = pts(z_f)
        if use_nof:
            xin_f, xloc_f, xglob_f = flows(x_f)
        else:
            xin_f = x_f
        rgb_f, dep_f, w_f, a_f = nerf_inference(xin_f, ind, d, z_f, noise_for(draws.noise_fine, z_f.shape), nerf_pes,
                                                nerfs[1], background=background, weights_only=False,
                                                activate_type=nerf_activate_type)
        result.update(rgb_fine=rgb_f, depth_fine=dep_f, opacity_fine=w_f.sum(1))
        aux.update(z_new=z_new, z_fine=z_f, weights_fine=w_f, alphas_fine=a_f)
        if use_nof and not test_time:
            if chain_local:
                result["nof_local_disp_fine"] = _masked_residual(x_f, xloc_f, a_f)
            if chain_global:
                result["nof_global_disp_fine"] = _masked_residual(x_f, xglob_f, a_f)
    if return_aux:
        result["_aux"] = aux
    return result


# --------------------------------------------------------------------------
# a9 loss, and the training objective the benchmark differentiates
# --------------------------------------------------------------------------
def image_mse(result: Dict[str, Tensor], target: Tensor) -> Tensor:
    """models/losses.py:9-14."""
    loss = F.mse_loss(result["rgb_coarse"], target)
    if "rgb_fine" in result:
        loss = loss + F.mse_loss(result["rgb_fine"], target)
    return loss


def train_objective(result: Dict[str, Tensor], target: Tensor, img_w: float = 1.0,
                    local_w: float = 0.2, global_w: float = 0.2) -> Tensor:
    """trainer/trainer_moco_flow.py:317-328 (image + chain terms; aux SMPL terms are out of scope)."""
    total = image_mse(result, target) * img_w
    for key, wgt in (("nof_local_disp", local_w), ("nof_global_disp", global_w)):
        if key + "_coarse" in result:
            term = result[key + "_coarse"].mean()
            if key + "_fine" in result:
                term = term + result[key + "_fine"].mean()
            total = total + term * wgt
    return total


# --------------------------------------------------------------------------
# occupancy lattice of visualize_mesh
# --------------------------------------------------------------------------
def density_grid(nerf: NeRFBundle, pe_xyz: PESpec, n_grid: int, frame_idx: int = -1, nof: Optional[NoFBundle] = None,
                 nof_pes: Optional[Sequence[PESpec]] = None, num_frames: Optional[int] = None, chunk: int = 10000) -> Tensor:
    """trainer/trainer_moco_flow.py:484-516: sigma on the N^3 lattice over [-1.5, 1.5]^3 (numpy meshgrid, 'xy'
    indexing), optionally warped by the backward flow network of frame ``frame_idx`` (forward_nof, :160-187),
    clamped at 0, reshaped (N, N, N)."""
    import numpy as np
    t = np.linspace(-1.5, 1.5, n_grid)
    xyz = torch.FloatTensor(np.stack(np.meshgrid(t, t, t), -1).reshape(-1, 3))
    outs = []
    with torch.no_grad():
        for i in range(0, xyz.shape[0], chunk):
            x = xyz[i:i + chunk]
            if frame_idx != -1:
                ind = torch.tensor([frame_idx]).unsqueeze(0).repeat(x.shape[0], 1).float() * 2 / num_frames - 1.0  # :175
                feats = torch.cat([_padded(positional_encoding(x, nof_pes[0]), nof.spec.in_channels_xyz),
                                   _padded(positional_encoding(ind, nof_pes[1]), nof.spec.extra_feat_dim)], -1)
                x = nof_mlp(nof.params, nof.spec, feats, x).view(-1, 3)
            emb = _padded(positional_encoding(x, pe_xyz), nerf.spec.in_channels_xyz)
            outs.append(nerf_mlp(nerf.params, nerf.spec, emb, sigma_only=True))
    sigma = torch.cat(outs, 0)
    return sigma.clamp_min(0).reshape(n_grid, n_grid, n_grid)


# --------------------------------------------------------------------------
# deterministic synthetic inputs shared by tests, smoke() and bench.py
# --------------------------------------------------------------------------
def _np_rng(seed: int):
    import numpy as np
    return np.random.Generator(np.random.PCG64(seed))


def linear_init(rng, out_f: int, in_f: int):
    """nn.Linear's default U(-1/sqrt(in), 1/sqrt(in)) drawn from a numpy PCG64 stream so the
    same weights can be rebuilt anywhere without depending on torch's RNG implementation."""
    bound = 1.0 / math.sqrt(in_f)
    w = torch.from_numpy(rng.uniform(-bound, bound, size=(out_f, in_f)).astype("float32"))
    b = torch.from_numpy(rng.uniform(-bound, bound, size=(out_f,)).astype("float32"))
    return w, b


DENSE_SIGMA_STD = 5.0
DENSE_SIGMA_MEAN = 10.0


def make_nerf_params(spec: NeRFSpec, seed: int, dense: bool = False) -> Dict[str, Tensor]:
    """Random-init NeRF state-dict (reference names/shapes, models/nerf.py:28-59).  ``dense``
    re-centres and scales the sigma head so that opacities are non-degenerate."""
    rng = _np_rng(seed)
    p: Dict[str, Tensor] = {}
    for i in range(spec.D):
        in_f = spec.in_channels_xyz if i == 0 else (spec.W + spec.in_channels_xyz if i in spec.skips else spec.W)
        p[f"xyz_encoding_{i+1}.0.weight"], p[f"xyz_encoding_{i+1}.0.bias"] = linear_init(rng, spec.W, in_f)
    p["xyz_encoding_final.weight"], p["xyz_encoding_final.bias"] = linear_init(rng, spec.W, spec.W)
    ext = spec.extra_feat_dim if spec.extra_feat_type != "none" else 0
    p["extra_encoding.0.weight"], p["extra_encoding.0.bias"] = linear_init(rng, spec.W // 2, spec.W + ext)
    p["sigma.weight"], p["sigma.bias"] = linear_init(rng, 1, spec.W)
    p["rgb.0.weight"], p["rgb.0.bias"] = linear_init(rng, 3, spec.W // 2)
    if dense:
        # Re-centre and scale the density head on a fixed probe set (sigma ~ mean 10, std 5): alphas then
        # cover (0,1) and every ray saturates before the far plane, instead of the all-empty volumes the
        # default init gives.  Saturating matters for parity testing: the reference's last
        # sample has delta = 1e10, so on a ray that is still transparent there, alpha_last is the step
        # function [sigma_last > 0] and ANY perturbation of sigma (bf16 rounding included) can flip it.
        nf = max((spec.in_channels_xyz // 3 - 1) // 2, 0)
        probe = torch.from_numpy(rng.uniform([-0.5, -1.0, -0.3], [0.5, 1.0, 0.3], size=(2048, 3)).astype("float32"))
        feats = _padded(positional_encoding(probe, PESpec(3, nf)), spec.in_channels_xyz)
        with torch.no_grad():
            raw = nerf_mlp(p, spec, feats, sigma_only=True)
        gain = DENSE_SIGMA_STD / float(raw.std().clamp_min(1e-6))
        p["sigma.weight"] = p["sigma.weight"] * gain
        p["sigma.bias"] = (p["sigma.bias"] - raw.mean()) * gain + DENSE_SIGMA_MEAN
        p["rgb.0.weight"] = p["rgb.0.weight"] * 8.0
    return p


def make_nof_params(spec: NoFSpec, seed: int, scale_head: float = 1.0) -> Dict[str, Tensor]:
    """Random-init NoF state-dict (models/nof.py:40-53)."""
    rng = _np_rng(seed)
    p: Dict[str, Tensor] = {}
    cin = spec.in_channels_xyz + spec.extra_feat_dim
    for i in range(spec.D):
        in_f = cin if i == 0 else (spec.W + cin if i in spec.skips else spec.W)
        p[f"nof_encoding_{i+1}.0.weight"], p[f"nof_encoding_{i+1}.0.bias"] = linear_init(rng, spec.W, in_f)
    p["nof_encoding_final.weight"], p["nof_encoding_final.bias"] = linear_init(rng, 9 if spec.use_quat else 3, spec.W)
    if scale_head != 1.0:
        p["nof_encoding_final.weight"] = p["nof_encoding_final.weight"] * scale_head
        p["nof_encoding_final.bias"] = p["nof_encoding_final.bias"] * scale_head
    return p


def make_rays(n_rays: int, seed: int = 1, n_frames: int = 160, chained: bool = False) -> Tensor:
    """Synthetic ray bundle of SURVEY 8(d): rows [o(3), d(3), near, far, img_ind, (chained_ind)]
    (layout of utils/camera.py:142-146, trainer/trainer_moco_flow.py:309-312)."""
    rng = _np_rng(seed)
    import numpy as np
    o = np.tile(np.array([[0.0, 0.0, 2.8]], dtype=np.float32), (n_rays, 1))
    tgt = rng.uniform([-0.5, -1.0, -0.3], [0.5, 1.0, 0.3], size=(n_rays, 3)).astype(np.float32)
    d = tgt - o
    d = d / np.linalg.norm(d, axis=1, keepdims=True)
    near = np.full((n_rays, 1), 2.0, np.float32)
    far = np.full((n_rays, 1), 3.6, np.float32)
    ind = (rng.integers(0, n_frames, size=(n_rays, 1)) * 2.0 / n_frames - 1.0).astype(np.float32)
    cols = [o, d.astype(np.float32), near, far, ind]
    if chained:
        c = float(rng.integers(0, n_frames)) * 2.0 / n_frames - 1.0
        cols.append(np.full((n_rays, 1), c, np.float32))
    return torch.from_numpy(np.concatenate(cols, axis=1))


def make_draws(n_rays: int, n_coarse: int, n_fine: int, seed: int = 2, noise: bool = False) -> RenderDraws:
    """Injected random tensors; noise tensors are zeros unless ``noise``."""
    rng = _np_rng(seed)
    f32 = "float32"
    dr = RenderDraws()
    dr.perturb = torch.from_numpy(rng.random((n_rays, n_coarse), dtype=f32))
    dr.u = torch.from_numpy(rng.random((n_rays, max(n_fine, 1)), dtype=f32))[:, :n_fine]
    if noise:
        dr.noise_coarse = torch.from_numpy(rng.standard_normal((n_rays, n_coarse), dtype=f32))
        dr.noise_fine = torch.from_numpy(rng.standard_normal((n_rays, n_coarse + n_fine), dtype=f32))
    else:
        dr.noise_coarse = torch.zeros(n_rays, n_coarse)
        dr.noise_fine = torch.zeros(n_rays, n_coarse + n_fine)
    return dr


# c2f.yaml shapes (configs/people_snapshot/male-3-casual/c2f.yaml:44-102)
C2F_NERF = NeRFSpec(D=8, W=256, in_channels_xyz=63, skips=(4,), extra_feat_type="ind", extra_feat_dim=5)
C2F_NOF = NoFSpec(D=4, W=128, in_channels_xyz=33, skips=(2,), extra_feat_type="ind", extra_feat_dim=33, use_quat=True)
C2F_PE = dict(nerf_xyz=PESpec(3, 10), nerf_ind=PESpec(1, 2), nof_xyz=PESpec(3, 5), nof_ind=PESpec(1, 16))
