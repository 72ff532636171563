"""Pins oracle/moco_oracle.py against fixtures produced by the reference's own code
(tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import moco_oracle as orc

T = torch.from_numpy


def load(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, name)))


def test_pe_matches_reference(golden_dir):
    g = load(golden_dir, "pe.npz")
    specs = {"xyz10": orc.PESpec(3, 10), "ind16": orc.PESpec(1, 16),
             "xyz5_c2f": orc.PESpec(3, 5, True, [1.0, 1.0, 0.37, 0.0, 0.0]),
             "lin4": orc.PESpec(3, 4, False), "xyz0": orc.PESpec(3, 0)}
    for name, spec in specs.items():
        y = orc.positional_encoding(T(g[name + "_x"]), spec)
        assert y.shape[1] == spec.out_channels
        np.testing.assert_array_equal(y.numpy(), g[name + "_y"])


def test_modules_match_reference(golden_dir):
    g = load(golden_dir, "modules.npz")
    p = orc.make_nerf_params(orc.C2F_NERF, 11)
    out = orc.nerf_mlp(p, orc.C2F_NERF, T(g["nerf_in"]))
    np.testing.assert_allclose(out.numpy(), g["nerf_out"], rtol=0, atol=1e-6)
    sig = orc.nerf_mlp(p, orc.C2F_NERF, T(g["nerf_in"])[:, :63], sigma_only=True)
    np.testing.assert_allclose(sig.numpy(), g["nerf_sigma"], rtol=0, atol=1e-6)
    q = orc.make_nof_params(orc.C2F_NOF, 21)
    out = orc.nof_mlp(q, orc.C2F_NOF, T(g["nof_in"]), T(g["nof_xyz"]))
    np.testing.assert_allclose(out.numpy(), g["nof_out"], rtol=0, atol=1e-6)
    spec3 = orc.NoFSpec(D=4, W=128, in_channels_xyz=33, skips=(2,), extra_feat_dim=33, use_quat=False)
    q3 = orc.make_nof_params(spec3, 22)
    out = orc.nof_mlp(q3, spec3, T(g["nof_in"]), T(g["nof_xyz"]))
    np.testing.assert_allclose(out.numpy(), g["nof3_out"], rtol=0, atol=1e-6)


def test_param_counts():
    # NeRF 593 028, NoF 67 721 parameters
    assert sum(v.numel() for v in orc.make_nerf_params(orc.C2F_NERF, 0).values()) == 593028
    assert sum(v.numel() for v in orc.make_nof_params(orc.C2F_NOF, 0).values()) == 67721


def test_quaternion_restatement_vs_rodrigues():
    g = torch.Generator().manual_seed(3)
    v = torch.randn(257, 3, generator=g, dtype=torch.float64) * 0.7
    v[0] = 0.0
    v[1] = 1e-9
    r_k = orc.quat_to_rotmat(orc.quat_log_to_exp(v))
    r_r = orc.rodrigues_rotmat(v)
    assert torch.allclose(r_k, r_r, atol=1e-12)
    # proper rotations
    eye = torch.eye(3, dtype=torch.float64).expand_as(r_k)
    assert torch.allclose(r_k @ r_k.transpose(1, 2), eye, atol=1e-12)


def test_sample_pdf_matches_reference(golden_dir):
    g = load(golden_dir, "sample_pdf.npz")
    for tag, nimp in (("a", 64), ("b", 16), ("c", 128)):
        bins, w, u = T(g[f"{tag}_bins"]), T(g[f"{tag}_w"]), T(g[f"{tag}_u"])
        np.testing.assert_array_equal(orc.sample_pdf(bins, w, nimp, det=False, u=u).numpy(), g[f"{tag}_rand"])
        np.testing.assert_array_equal(orc.sample_pdf(bins, w, nimp, det=True).numpy(), g[f"{tag}_det"])


def test_composite_matches_reference(golden_dir):
    g = load(golden_dir, "composite.npz")
    for tag in ("relu", "softplus"):
        raw = T(g[f"{tag}_raw"])
        rgb, dep, w, a = orc.composite(raw[..., 3], raw[..., :3], T(g[f"{tag}_z"]), T(g[f"{tag}_dirs"]),
                                       T(g[f"{tag}_noise"]) * 0.5, T(g[f"{tag}_bg"]), tag)
        for got, key in ((rgb, "rgb"), (dep, "depth"), (w, "w"), (a, "alpha")):
            np.testing.assert_allclose(got.numpy(), g[f"{tag}_{key}"], rtol=1e-6, atol=1e-7)


RENDER_CASES = {
    # name: (Sc, Sf, use_nof, local, global, perturb, noise_std, act, test_time, dense, pes, nerf spec)
    "cfg1_nerf_only": (64, 64, False, False, False, 0.0, 0.0, "relu", False, True, None, None),
    "moco_test_time": (32, 32, True, False, False, 0.0, 0.0, "relu", True, True, None, None),
    "moco_train": (16, 16, True, True, True, 1.0, 0.0, "relu", False, True, None, None),
    "moco_train_noise": (16, 24, True, True, False, 0.5, 1.0, "relu", False, True, None, None),
    "default_init": (16, 16, True, True, True, 1.0, 0.0, "relu", False, False, None, None),
    "init_nerf_dir": (16, 16, False, False, False, 1.0, 0.0, "softplus", False, True,
                      dict(nerf_xyz=orc.PESpec(3, 0), nerf_dir=orc.PESpec(3, 4)),
                      orc.NeRFSpec(D=8, W=256, in_channels_xyz=63, skips=(4,), extra_feat_type="dir", extra_feat_dim=27)),
    "coarse_only": (24, 0, True, True, False, 0.0, 0.0, "relu", False, True, None, None),
}


def build_case(name, g):
    Sc, Sf, use_nof, loc, glob, perturb, nstd, act, tt, dense, pes, nspec = RENDER_CASES[name]
    pes = pes or orc.C2F_PE
    nspec = nspec or orc.C2F_NERF
    nerfs = [orc.NeRFBundle(nspec, orc.make_nerf_params(nspec, s, dense=dense)) for s in (101, 102)]
    nofs = [orc.NoFBundle(orc.C2F_NOF, orc.make_nof_params(orc.C2F_NOF, s, scale_head=0.25)) for s in (201, 202)] if use_nof else None
    rays, bg = T(g["rays"]), T(g["bg"])
    draws = orc.make_draws(rays.shape[0], Sc, Sf, seed=5, noise=(nstd > 0))
    nerf_pes = [pes["nerf_xyz"], pes.get("nerf_ind"), pes.get("nerf_dir")]
    nof_pes = [pes["nof_xyz"], pes["nof_ind"]] if use_nof else None
    kw = dict(chain_local=loc, chain_global=glob, N_samples=Sc, N_importance=Sf, perturb=perturb, noise_std=nstd,
              nerf_activate_type=act, test_time=tt, draws=draws)
    return rays, bg, nerf_pes, nerfs, nof_pes, nofs, kw


@pytest.mark.parametrize("name", list(RENDER_CASES))
def test_render_rays_matches_reference(golden_dir, name):
    g = load(golden_dir, f"render_{name}.npz")
    rays, bg, nerf_pes, nerfs, nof_pes, nofs, kw = build_case(name, g)
    leaves = []
    for b in nerfs + (nofs or []):
        for k in b.params:
            b.params[k] = b.params[k].clone().requires_grad_(True)
            leaves.append(b.params[k])
    res = orc.render_rays(rays, bg, nerf_pes, nerfs, nof_pes, nofs, **kw)
    keys = sorted(k[4:] for k in g if k.startswith("out_"))
    assert sorted(res.keys()) == keys
    for k in keys:
        assert tuple(res[k].shape) == g["out_" + k].shape, k
        np.testing.assert_allclose(res[k].detach().numpy(), g["out_" + k], rtol=2e-5, atol=2e-6, err_msg=k)
    if "loss" in g:
        loss = orc.train_objective(res, T(g["target"]))
        np.testing.assert_allclose(loss.item(), g["loss"], rtol=1e-5)
        names = []
        mods = {"nerf0": nerfs[0], "nerf1": nerfs[1]}
        if nofs:
            mods.update(nof0=nofs[0], nof1=nofs[1])
        plist = [(f"{mn}.{pn}", p) for mn, b in mods.items() for pn, p in b.params.items()]
        grads = torch.autograd.grad(loss, [p for _, p in plist], allow_unused=True)
        for (pn, p), gr in zip(plist, grads):
            gr = torch.zeros_like(p) if gr is None else gr
            ref_sum = g["grad_sum_" + pn]
            scale = max(ref_sum[1], 1e-12)
            assert abs(gr.double().sum().item() - ref_sum[0]) <= 2e-4 * scale + 1e-9, pn
            assert abs(gr.double().abs().sum().item() - ref_sum[1]) <= 2e-4 * scale + 1e-9, pn
            np.testing.assert_allclose(gr.reshape(-1)[:48].numpy(), g["grad_head_" + pn], rtol=2e-3,
                                       atol=2e-4 * scale / max(gr.numel(), 1) + 1e-9, err_msg=pn)
