#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REFERENCE's own modules.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):    python tests/golden/make_golden.py

What is the reference and what is not:
  * ``models/embedding.py``, ``models/nerf.py``, ``models/nof.py`` and
    ``models/rendering.py`` are imported unmodified from /root/reference.
  * ``kornia`` (a pip dependency of ``models/nof.py:4``) is not installed; the two
    functions it provides are supplied by a shim backed by the restatement in
    ``oracle/moco_oracle.py`` -- that part of the golden data is therefore NOT an
    independent pin (the tests cross-check it against a Rodrigues formula).
  * torch.rand / torch.randn inside ``models/rendering.py`` are replaced by a queue
    so that the reference consumes the same random tensors the fixtures store.
Weights are rebuilt from numpy PCG64 seeds (oracle.make_*_params) so the fixtures
only need to store inputs and outputs.
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("MOCO_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
from oracle import moco_oracle as orc  # noqa: E402


def _install_kornia_shim():
    conv = types.ModuleType("kornia.geometry.conversions")
    conv.quaternion_log_to_exp = lambda v: orc.quat_log_to_exp(v)
    conv.quaternion_to_rotation_matrix = lambda q: orc.quat_to_rotmat(q)
    geo = types.ModuleType("kornia.geometry")
    geo.conversions = conv
    top = types.ModuleType("kornia")
    top.geometry = geo
    sys.modules.update({"kornia": top, "kornia.geometry": geo, "kornia.geometry.conversions": conv})


def load_reference():
    _install_kornia_shim()
    sys.path.insert(0, REF)
    emb = importlib.import_module("models.embedding")
    nerf = importlib.import_module("models.nerf")
    nof = importlib.import_module("models.nof")
    rend = importlib.import_module("models.rendering")
    return emb, nerf, nof, rend


class _TorchProxy:
    """Stands in for the ``torch`` global of models/rendering.py: rand/randn pop a queue."""

    def __init__(self, queue):
        self._q = queue

    def __getattr__(self, name):
        return getattr(torch, name)

    def rand(self, *shape, **kw):
        t = self._q.pop(0)
        assert t[0] == "rand", t[0]
        return t[1].clone()

    def randn(self, *shape, **kw):
        t = self._q.pop(0)
        assert t[0] == "randn", t[0]
        return t[1].clone()


def ref_modules(emb_mod, nerf_mod, nof_mod, nerf_spec, nof_spec, pes, seeds, dense):
    def mk_emb(s):
        e = emb_mod.Embedding(s.in_channels, s.n_freqs, s.logscale)
        if s.weights is not None:
            e.set_weights(list(s.weights))
        return e
    nerfs, nerf_params = [], []
    for sd in seeds["nerf"]:
        m = nerf_mod.NeRF(nerf_spec.D, nerf_spec.W, nerf_spec.in_channels_xyz, list(nerf_spec.skips),
                          nerf_spec.extra_feat_type, nerf_spec.extra_feat_dim)
        p = orc.make_nerf_params(nerf_spec, sd, dense=dense)
        m.load_state_dict(p, strict=True)
        nerfs.append(m)
        nerf_params.append(p)
    nofs, nof_params = [], []
    for sd in seeds.get("nof", []):
        m = nof_mod.NoF(nof_spec.D, nof_spec.W, nof_spec.in_channels_xyz, list(nof_spec.skips),
                        nof_spec.extra_feat_type, nof_spec.extra_feat_dim, nof_spec.use_quat)
        p = orc.make_nof_params(nof_spec, sd, scale_head=seeds.get("nof_head_scale", 1.0))
        m.load_state_dict(p, strict=True)
        nofs.append(m)
        nof_params.append(p)
    nerf_embs = [mk_emb(pes["nerf_xyz"]),
                 mk_emb(pes["nerf_ind"]) if pes.get("nerf_ind") else None,
                 mk_emb(pes["nerf_dir"]) if pes.get("nerf_dir") else None]
    nof_embs = [mk_emb(pes["nof_xyz"]), mk_emb(pes["nof_ind"])] if nofs else None
    return nerfs, nofs, nerf_embs, nof_embs, nerf_params, nof_params


def np_(t):
    return t.detach().cpu().numpy()


def main():
    emb_mod, nerf_mod, nof_mod, rend = load_reference()
    torch.manual_seed(1234)
    g = np.random.Generator(np.random.PCG64(7))
    out = {}

    # ---- a1 Embedding ---------------------------------------------------
    pe_cases = {}
    for name, (cin, nf, logscale, w) in {
        "xyz10": (3, 10, True, None), "ind16": (1, 16, True, None), "xyz5_c2f": (3, 5, True, [1.0, 1.0, 0.37, 0.0, 0.0]),
        "lin4": (3, 4, False, None), "xyz0": (3, 0, True, None),
    }.items():
        x = torch.from_numpy(g.uniform(-1.5, 1.5, size=(17, cin)).astype(np.float32))
        e = emb_mod.Embedding(cin, nf, logscale)
        if w is not None:
            e.set_weights(w)
        pe_cases[name + "_x"] = np_(x)
        pe_cases[name + "_y"] = np_(e(x))
    np.savez_compressed(os.path.join(HERE, "pe.npz"), **pe_cases)

    # ---- a2/a3 module forwards ------------------------------------------
    mod = {}
    nerf_spec, nof_spec = orc.C2F_NERF, orc.C2F_NOF
    nerfs, nofs, *_ = ref_modules(emb_mod, nerf_mod, nof_mod, nerf_spec, nof_spec, orc.C2F_PE,
                                  dict(nerf=[11], nof=[21]), dense=False)
    x = torch.from_numpy(g.uniform(-1, 1, size=(40, 68)).astype(np.float32))
    mod["nerf_in"] = np_(x)
    mod["nerf_out"] = np_(nerfs[0](x))
    mod["nerf_sigma"] = np_(nerfs[0](x[:, :63], sigma_only=True))
    xi = torch.from_numpy(g.uniform(-1, 1, size=(40, 66)).astype(np.float32))
    xyz = torch.from_numpy(g.uniform(-1, 1, size=(40, 3)).astype(np.float32))
    mod["nof_in"], mod["nof_xyz"] = np_(xi), np_(xyz)
    mod["nof_out"] = np_(nofs[0](xi, xyz))
    nof3_spec = orc.NoFSpec(D=4, W=128, in_channels_xyz=33, skips=(2,), extra_feat_dim=33, use_quat=False)
    _, nofs3, *_ = ref_modules(emb_mod, nerf_mod, nof_mod, nerf_spec, nof3_spec, orc.C2F_PE,
                               dict(nerf=[], nof=[22]), dense=False)
    mod["nof3_out"] = np_(nofs3[0](xi, xyz))
    np.savez_compressed(os.path.join(HERE, "modules.npz"), **mod)

    # ---- a4 sample_pdf ---------------------------------------------------
    sp = {}
    for tag, (R, nb, nimp) in {"a": (33, 62, 64), "b": (9, 14, 16), "c": (5, 126, 128)}.items():
        bins = np.sort(g.uniform(2.0, 3.6, size=(R, nb + 1)).astype(np.float32), axis=1)
        wts = g.uniform(0, 1, size=(R, nb)).astype(np.float32) ** 4
        wts[0] = 0.0  # a ray with no mass at all
        wts[1, : nb // 2] = 0.0
        u = g.random((R, nimp), dtype=np.float32)
        u[2, 0], u[2, 1] = 0.0, np.float32(1.0) - np.float32(2 ** -24)
        bins_t, w_t, u_t = map(torch.from_numpy, (bins, wts, u))
        rend_torch = rend.torch
        rend.torch = _TorchProxy([("rand", u_t)])
        s_rand = rend.sample_pdf(bins_t, w_t, nimp, det=False)
        rend.torch = rend_torch
        s_det = rend.sample_pdf(bins_t, w_t, nimp, det=True)
        sp.update({f"{tag}_bins": bins, f"{tag}_w": wts, f"{tag}_u": u, f"{tag}_rand": np_(s_rand), f"{tag}_det": np_(s_det)})
    np.savez_compressed(os.path.join(HERE, "sample_pdf.npz"), **sp)

    # ---- a6 compositing via nerf_inference on a stub model -----------------
    comp = {}
    for tag, act in (("relu", "relu"), ("softplus", "softplus")):
        R, S = 21, 24
        z = np.sort(g.uniform(2.0, 3.6, size=(R, S)).astype(np.float32), axis=1)
        raw = (g.standard_normal((R, S, 4)) * np.array([1, 1, 1, 30.0])).astype(np.float32)
        raw[..., :3] = 1 / (1 + np.exp(-raw[..., :3]))
        dirs = g.standard_normal((R, 3)).astype(np.float32)
        bg = g.uniform(0, 1, size=(R, 3)).astype(np.float32)
        noise = g.standard_normal((R, S)).astype(np.float32)

        class Stub(torch.nn.Module):
            in_channels_xyz, extra_feat_type, extra_feat_dim = 3, "none", 0

            def forward(self, inp, sigma_only=False, img_ind=None):
                o = torch.from_numpy(raw).view(-1, 4)
                return o[:, 3:4] if sigma_only else o
        rend_torch = rend.torch
        rend.torch = _TorchProxy([("randn", torch.from_numpy(noise))])
        rgb, dep, w, a = rend.nerf_inference(torch.zeros(R, S, 3), torch.zeros(R, 1), torch.from_numpy(dirs),
                                             torch.from_numpy(z), 0.5, [emb_mod.Embedding(3, 0)], Stub(),
                                             background=torch.from_numpy(bg), weights_only=False, activate_type=act)
        rend.torch = rend_torch
        comp.update({f"{tag}_z": z, f"{tag}_raw": raw, f"{tag}_dirs": dirs, f"{tag}_bg": bg, f"{tag}_noise": noise,
                     f"{tag}_rgb": np_(rgb), f"{tag}_depth": np_(dep), f"{tag}_w": np_(w), f"{tag}_alpha": np_(a)})
    np.savez_compressed(os.path.join(HERE, "composite.npz"), **comp)

    # ---- a7 render_rays ----------------------------------------------------
    cases = {
        # name: (R, Sc, Sf, use_nof, local, global, perturb, noise_std, act, test_time, dense, pes, nerf_spec)
        "cfg1_nerf_only":   (24, 64, 64, False, False, False, 0.0, 0.0, "relu", False, True, orc.C2F_PE, orc.C2F_NERF),
        "moco_test_time":   (24, 32, 32, True, False, False, 0.0, 0.0, "relu", True, True, orc.C2F_PE, orc.C2F_NERF),
        "moco_train":       (20, 16, 16, True, True, True, 1.0, 0.0, "relu", False, True, orc.C2F_PE, orc.C2F_NERF),
        "moco_train_noise": (12, 16, 24, True, True, False, 0.5, 1.0, "relu", False, True, orc.C2F_PE, orc.C2F_NERF),
        "default_init":     (16, 16, 16, True, True, True, 1.0, 0.0, "relu", False, False, orc.C2F_PE, orc.C2F_NERF),
        "init_nerf_dir":    (16, 16, 16, False, False, False, 1.0, 0.0, "softplus", False, True,
                             dict(nerf_xyz=orc.PESpec(3, 0), nerf_dir=orc.PESpec(3, 4)),
                             orc.NeRFSpec(D=8, W=256, in_channels_xyz=63, skips=(4,), extra_feat_type="dir", extra_feat_dim=27)),
        "coarse_only":      (16, 24, 0, True, True, False, 0.0, 0.0, "relu", False, True, orc.C2F_PE, orc.C2F_NERF),
    }
    for case_no, (name, (R, Sc, Sf, use_nof, loc, glob, perturb, nstd, act, tt, dense, pes, nspec)) in enumerate(cases.items()):
        seeds = dict(nerf=[101, 102], nof=[201, 202] if use_nof else [], nof_head_scale=0.25)
        nerfs, nofs, nerf_embs, nof_embs, _, _ = ref_modules(emb_mod, nerf_mod, nof_mod, nspec, orc.C2F_NOF, pes, seeds, dense)
        rays = orc.make_rays(R, seed=300 + case_no, chained=glob)
        bg = torch.from_numpy(g.uniform(0, 1, size=(R, 3)).astype(np.float32))
        target = torch.from_numpy(g.uniform(0, 1, size=(R, 3)).astype(np.float32))
        dr = orc.make_draws(R, Sc, Sf, seed=5, noise=(nstd > 0))
        queue = []
        if perturb > 0:
            queue.append(("rand", dr.perturb))
        queue.append(("randn", dr.noise_coarse))
        if Sf > 0:
            if perturb > 0:
                queue.append(("rand", dr.u))
            queue.append(("randn", dr.noise_fine))
        rend_torch = rend.torch
        rend.torch = _TorchProxy(queue)
        res = rend.render_rays(rays, bg, nerf_embs, nerfs, nof_embeddings=nof_embs, nof_models=nofs if use_nof else None,
                               chain_local=loc, chain_global=glob, N_samples=Sc, N_importance=Sf, use_disp=False,
                               perturb=perturb, noise_std=nstd, nerf_activate_type=act, test_time=tt)
        rend.torch = rend_torch
        assert not queue, (name, [q[0] for q in queue])
        blob = dict(rays=np_(rays), bg=np_(bg), target=np_(target))
        for k, v in res.items():
            blob["out_" + k] = np_(v)
        if not tt and Sf > 0:
            # gradient pin: reference autograd of the training objective (trainer_moco_flow.py:317-328)
            loss = torch.nn.functional.mse_loss(res["rgb_coarse"], target) + torch.nn.functional.mse_loss(res["rgb_fine"], target)
            for key in ("nof_local_disp", "nof_global_disp"):
                if key + "_coarse" in res:
                    loss = loss + 0.2 * (res[key + "_coarse"].mean() + res[key + "_fine"].mean())
            blob["loss"] = np_(loss)
            mods = {"nerf0": nerfs[0], "nerf1": nerfs[1]}
            if use_nof:
                mods.update(nof0=nofs[0], nof1=nofs[1])
            params = [(f"{mn}.{pn}", p) for mn, m in mods.items() for pn, p in m.named_parameters()]
            grads = torch.autograd.grad(loss, [p for _, p in params], allow_unused=True)
            for (pn, p), gr in zip(params, grads):
                gr = torch.zeros_like(p) if gr is None else gr
                flat = gr.reshape(-1)
                blob["grad_sum_" + pn] = np.array([flat.double().sum().item(), flat.double().abs().sum().item()])
                blob["grad_head_" + pn] = np_(flat[:48])
        np.savez_compressed(os.path.join(HERE, f"render_{name}.npz"), **blob)
        print("wrote", name, {k: tuple(v.shape) for k, v in res.items()})

    print("done")


if __name__ == "__main__":
    main()
