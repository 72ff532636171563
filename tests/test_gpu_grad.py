"""GPU gradient parity: backward dX chains + weight-gradient GEMMs vs autograd of the CPU oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import moco_oracle as orc
from tests.helpers import pe_module
from tests.test_oracle_golden import build_case

pytestmark = pytest.mark.gpu
T = torch.from_numpy


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    return torch.device("cuda:0")


def rel_fro(got, ref):
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    return ((got - ref).norm() / ref.norm().clamp_min(1e-30)).item()


def compare_grads(tag, named_got, named_ref, tol):
    worst = 0.0
    for (n, g), (n2, r) in zip(named_got, named_ref):
        assert n == n2
        if r is None or r.abs().max() == 0:
            assert g is None or g.abs().max().item() <= 1e-6, n
            continue
        assert g is not None, n
        e = rel_fro(g, r)
        worst = max(worst, e)
        print(f"[grad] {tag}.{n}: rel fro err {e:.3e}  |ref| {r.norm().item():.3e}")
        assert e <= tol, (tag, n, e)
    return worst


def test_nerf_fused_gradients(dev):
    """NeRF evaluate(xyz, pe) with xyz requiring grad: d params and d xyz vs oracle autograd."""
    import moco_flow_b200 as mf
    gen = torch.Generator().manual_seed(21)
    R, S = 13, 24  # 312 rows: 3 tiles, last partial
    xyz = (torch.rand(R * S, 3, generator=gen) - 0.5) * 1.2
    ind = torch.rand(R, 1, generator=gen) * 2 - 1
    up = torch.randn(R * S, 4, generator=gen)
    p = orc.make_nerf_params(orc.C2F_NERF, 9, dense=True)
    m = mf.NeRF(8, 256, 63, [4], "ind", 5)
    m.load_state_dict(p)
    m = m.to(dev)
    pe, pe_i = mf.Embedding(3, 10), mf.Embedding(1, 2)
    xd = xyz.to(dev).requires_grad_(True)
    out = m.evaluate(xyz=xd, pe=pe, ray_feat=pe_i(ind.to(dev)), rows_per_ray=S)
    (out * up.to(dev)).sum().backward()
    torch.cuda.synchronize()
    from moco_flow_b200 import _lib as L
    assert L.device_error_flag() == 0
    po = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    xo = xyz.clone().requires_grad_(True)
    feats = torch.cat([orc.positional_encoding(xo, orc.PESpec(3, 10)),
                       orc.positional_encoding(ind, orc.PESpec(1, 2)).repeat_interleave(S, 0)], 1)
    ref = orc.nerf_mlp(po, orc.C2F_NERF, feats)
    (ref * up).sum().backward()
    assert (out.detach().cpu()[:, :3] - ref.detach()[:, :3]).abs().max().item() <= 2e-3
    got = [(n, q.grad) for n, q in m.named_parameters()]
    want = [(n, po[n].grad) for n, _ in m.named_parameters()]
    compare_grads("nerf", got, want, 3e-2)
    e = rel_fro(xd.grad, xo.grad)
    print(f"[grad] nerf.d_xyz rel fro err {e:.3e}")
    assert e <= 3e-2


@pytest.mark.parametrize("use_quat", [True, False])
def test_nof_fused_gradients(dev, use_quat):
    import moco_flow_b200 as mf
    gen = torch.Generator().manual_seed(22)
    R, S = 11, 16  # 176 rows: 2 tiles, last partial
    xyz = (torch.rand(R * S, 3, generator=gen) - 0.5) * 1.2
    ind = torch.rand(R, 1, generator=gen) * 2 - 1
    up = torch.randn(R * S, 3, generator=gen)
    spec = orc.NoFSpec(D=4, W=128, in_channels_xyz=33, skips=(2,), extra_feat_dim=33, use_quat=use_quat)
    p = orc.make_nof_params(spec, 31, scale_head=0.5)
    m = mf.NoF(4, 128, 33, [2], "ind", 33, use_quat)
    m.load_state_dict(p)
    m = m.to(dev)
    pe, pe_i = mf.Embedding(3, 5), mf.Embedding(1, 16)
    xd = xyz.to(dev).requires_grad_(True)
    out = m.evaluate(xyz=xd, pe=pe, ray_feat=pe_i(ind.to(dev)), rows_per_ray=S)
    (out * up.to(dev)).sum().backward()
    torch.cuda.synchronize()
    from moco_flow_b200 import _lib as L
    assert L.device_error_flag() == 0
    po = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    xo = xyz.clone().requires_grad_(True)
    feats = torch.cat([orc.positional_encoding(xo, orc.PESpec(3, 5)),
                       orc.positional_encoding(ind, orc.PESpec(1, 16)).repeat_interleave(S, 0)], 1)
    ref = orc.nof_mlp(po, spec, feats, xo)
    (ref * up).sum().backward()
    assert (out.detach().cpu() - ref.detach()).abs().max().item() <= 5e-3
    got = [(n, q.grad) for n, q in m.named_parameters()]
    want = [(n, po[n].grad) for n, _ in m.named_parameters()]
    compare_grads(f"nof[quat={use_quat}]", got, want, 3e-2)
    e = rel_fro(xd.grad, xo.grad)
    print(f"[grad] nof.d_xyz rel fro err {e:.3e}")
    assert e <= 3e-2


def test_nerf_module_dense_gradients(dev):
    """Reference call convention NeRF.forward(inputs) under autograd."""
    import moco_flow_b200 as mf
    gen = torch.Generator().manual_seed(23)
    x = torch.rand(200, 68, generator=gen) * 2 - 1
    up = torch.randn(200, 4, generator=gen)
    p = orc.make_nerf_params(orc.C2F_NERF, 12)
    m = mf.NeRF(8, 256, 63, [4], "ind", 5)
    m.load_state_dict(p)
    m = m.to(dev)
    (m(x.to(dev)) * up.to(dev)).sum().backward()
    po = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    (orc.nerf_mlp(po, orc.C2F_NERF, x) * up).sum().backward()
    got = [(n, q.grad) for n, q in m.named_parameters()]
    want = [(n, po[n].grad) for n, _ in m.named_parameters()]
    compare_grads("nerf-dense", got, want, 3e-2)


@pytest.mark.parametrize("name", ["moco_train", "moco_train_noise", "cfg1_nerf_only", "default_init"])
def test_render_rays_training_step(dev, golden_dir, name):
    """Full training objective (image MSE + chain terms): outputs, loss and all parameter gradients."""
    import moco_flow_b200 as mf
    g = dict(np.load(os.path.join(golden_dir, f"render_{name}.npz")))
    rays, bg, nerf_pes, nerfs, nof_pes, nofs, kw = build_case(name, g)
    spec = nerfs[0].spec
    models, nof_models = [], None
    for b in nerfs:
        m = mf.NeRF(spec.D, spec.W, spec.in_channels_xyz, list(spec.skips), spec.extra_feat_type, spec.extra_feat_dim)
        m.load_state_dict(b.params)
        models.append(m.to(dev))
    if nofs:
        nof_models = []
        for b in nofs:
            s = b.spec
            m = mf.NoF(s.D, s.W, s.in_channels_xyz, list(s.skips), s.extra_feat_type, s.extra_feat_dim, s.use_quat)
            m.load_state_dict(b.params)
            nof_models.append(m.to(dev))
    nerf_embs = [pe_module(pp, mf.Embedding) if pp is not None else None for pp in nerf_pes]
    nof_embs = [pe_module(pp, mf.Embedding) for pp in nof_pes] if nof_pes else None
    dr = kw.pop("draws")
    draws = mf.Draws(*(None if t is None else t.to(dev) for t in (dr.perturb, dr.noise_coarse, dr.u, dr.noise_fine)))
    target = T(g["target"])
    res = mf.render_rays(rays.to(dev), bg.to(dev), nerf_embs, models, nof_embeddings=nof_embs, nof_models=nof_models,
                         draws=draws, **kw)
    loss = mf.MSELoss()(res, target.to(dev))
    for key in ("nof_local_disp", "nof_global_disp"):
        if key + "_coarse" in res:
            loss = loss + 0.2 * (res[key + "_coarse"].mean() + res[key + "_fine"].mean())
    loss.backward()
    torch.cuda.synchronize()
    from moco_flow_b200 import _lib as L
    assert L.device_error_flag() == 0
    print(f"[grad] {name}: loss {loss.item():.6f} (reference {float(g['loss']):.6f})")
    assert abs(loss.item() - float(g["loss"])) <= 3e-3 * max(1.0, abs(float(g["loss"])))
    # oracle autograd on the same inputs (full gradients)
    for b in nerfs + (nofs or []):
        for k in b.params:
            b.params[k] = b.params[k].clone().requires_grad_(True)
    kw["draws"] = dr
    ref = orc.render_rays(rays, bg, nerf_pes, nerfs, nof_pes, nofs, **kw)
    orc.train_objective(ref, target).backward()
    mods = [("nerf0", models[0], nerfs[0]), ("nerf1", models[1], nerfs[1])]
    if nofs:
        mods += [("nof0", nof_models[0], nofs[0]), ("nof1", nof_models[1], nofs[1])]
    for tag, mod, bundle in mods:
        got = [(n, q.grad) for n, q in mod.named_parameters()]
        want = [(n, bundle.params[n].grad) for n, _ in mod.named_parameters()]
        # bf16 operands in every GEMM of a 10-layer fwd+bwd chain: a few percent in Frobenius norm
        compare_grads(f"{name}.{tag}", got, want, 6e-2)
