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


def cosine(a, b):
    a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    return (a @ b / (a.norm() * b.norm()).clamp_min(1e-300)).item()


def compare_grads(tag, named_got, named_ref, tol, named_fp32=None, min_cos=0.9):
    """``named_ref``: autograd of the bf16-emulating oracle (the exact gradient of what the kernels compute, up to
    bf16 rounding inside the backward GEMMs) -> Frobenius-relative tolerance ``tol``.
    ``named_fp32``: autograd of the fp32 reference algorithm.  ReLU units whose pre-activation is within bf16
    rounding of zero switch on/off between the two forwards, which changes their whole gradient contribution, so
    only the direction (cosine) is asserted against fp32."""
    bad = []
    for idx, ((n, g), (n2, r)) in enumerate(zip(named_got, named_ref)):
        assert n == n2
        if r is None or r.abs().max() == 0:
            if g is not None and g.abs().max().item() > 1e-6:
                bad.append((n, "expected zero grad"))
            continue
        if g is None:
            bad.append((n, "missing grad"))
            continue
        e = rel_fro(g, r)
        msg = f"[grad] {tag}.{n}: rel fro err vs bf16-emulated {e:.3e}  |ref| {r.norm().item():.3e}"
        if named_fp32 is not None and named_fp32[idx][1] is not None:
            c = cosine(g, named_fp32[idx][1])
            msg += f"  cos vs fp32 {c:.4f}  rel fro vs fp32 {rel_fro(g, named_fp32[idx][1]):.3e}"
            if c < min_cos:
                bad.append((n, f"cos {c}"))
        print(msg)
        if e > tol:
            bad.append((n, e))
    assert not bad, (tag, bad)


def oracle_grads(fn, params_list):
    """Runs ``fn()`` (a scalar from the oracle) twice: fp32 and bf16-emulated; returns the two grad lists."""
    out = []
    for emu in (False, True):
        for p in params_list:
            p.grad = None
        orc.EMULATE_BF16 = emu
        try:
            fn().backward()
        finally:
            orc.EMULATE_BF16 = False
        out.append([None if p.grad is None else p.grad.clone() for p in params_list])
    return out  # [fp32 grads, emulated grads]


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
    names = [n for n, _ in m.named_parameters()]

    def run():
        feats = torch.cat([orc.positional_encoding(xo, orc.PESpec(3, 10)),
                           orc.positional_encoding(ind, orc.PESpec(1, 2)).repeat_interleave(S, 0)], 1)
        return (orc.nerf_mlp(po, orc.C2F_NERF, feats) * up).sum()
    g32, gem = oracle_grads(run, [po[n] for n in names] + [xo])
    got = [(n, q.grad) for n, q in m.named_parameters()] + [("d_xyz", xd.grad)]
    compare_grads("nerf", got, list(zip(names + ["d_xyz"], gem)), 3e-2, list(zip(names + ["d_xyz"], g32)))


def test_nerf_sigma_only_gradients(dev):
    """Auxiliary density loss on a point batch (trainer/trainer_moco_flow.py:146-158,349-361): sigma-only evaluation,
    alpha = 1 - exp(-delta * softplus(sigma)), gradients to the trunk + sigma head and to the points."""
    import moco_flow_b200 as mf
    gen = torch.Generator().manual_seed(33)
    N = 700  # 6 tiles, last partial; one "ray" holding all points
    xyz = (torch.rand(N, 3, generator=gen) - 0.5) * 1.2
    up = torch.rand(N, 1, generator=gen)
    p = orc.make_nerf_params(orc.C2F_NERF, 12, dense=True)
    m = mf.NeRF(8, 256, 63, [4], "ind", 5)
    m.load_state_dict(p)
    m = m.to(dev)
    pe = mf.Embedding(3, 10)
    delta = 1.0 / 64
    xd = xyz.to(dev).requires_grad_(True)
    sig = m.evaluate(xyz=xd, pe=pe, rows_per_ray=N, sigma_only=True)
    assert sig.shape == (N, 1)
    ((1 - torch.exp(-delta * torch.nn.functional.softplus(sig))) * up.to(dev)).sum().backward()
    torch.cuda.synchronize()
    from moco_flow_b200 import _lib as L
    assert L.device_error_flag() == 0
    po = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    xo = xyz.clone().requires_grad_(True)
    trunk = [n for n, _ in m.named_parameters() if n.startswith("xyz_encoding_") and "final" not in n or n.startswith("sigma")]

    def run():
        feats = torch.cat([orc.positional_encoding(xo, orc.PESpec(3, 10)), torch.zeros(N, 5)], 1)
        s_ = orc.nerf_mlp(po, orc.C2F_NERF, feats)[:, 3:4]
        return ((1 - torch.exp(-delta * torch.nn.functional.softplus(s_))) * up).sum()
    g32, gem = oracle_grads(run, [po[n] for n in trunk] + [xo])
    named = dict(m.named_parameters())
    got = [(n, named[n].grad) for n in trunk] + [("d_xyz", xd.grad)]
    compare_grads("nerf sigma-only", got, list(zip(trunk + ["d_xyz"], gem)), 3e-2, list(zip(trunk + ["d_xyz"], g32)))
    # the colour branch is untouched by this loss
    for n, q in m.named_parameters():
        if n not in trunk:
            assert q.grad is None or float(q.grad.abs().max()) == 0.0, n


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
    names = [n for n, _ in m.named_parameters()]

    def run():
        feats = torch.cat([orc.positional_encoding(xo, orc.PESpec(3, 5)),
                           orc.positional_encoding(ind, orc.PESpec(1, 16)).repeat_interleave(S, 0)], 1)
        return (orc.nof_mlp(po, spec, feats, xo) * up).sum()
    g32, gem = oracle_grads(run, [po[n] for n in names] + [xo])
    got = [(n, q.grad) for n, q in m.named_parameters()] + [("d_xyz", xd.grad)]
    compare_grads(f"nof[quat={use_quat}]", got, list(zip(names + ["d_xyz"], gem)), 3e-2,
                  list(zip(names + ["d_xyz"], g32)))


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
    names = [n for n, _ in m.named_parameters()]
    g32, gem = oracle_grads(lambda: (orc.nerf_mlp(po, orc.C2F_NERF, x) * up).sum(), [po[n] for n in names])
    got = [(n, q.grad) for n, q in m.named_parameters()]
    compare_grads("nerf-dense", got, list(zip(names, gem)), 3e-2, list(zip(names, g32)))


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
    # oracle autograd on the same inputs (full gradients), fp32 and bf16-emulated
    for b in nerfs + (nofs or []):
        for k in b.params:
            b.params[k] = b.params[k].clone().requires_grad_(True)
    kw["draws"] = dr
    mods = [("nerf0", models[0], nerfs[0]), ("nerf1", models[1], nerfs[1])]
    if nofs:
        mods += [("nof0", nof_models[0], nofs[0]), ("nof1", nof_models[1], nofs[1])]
    plist = [bundle.params[n] for _, mod, bundle in mods for n, _ in mod.named_parameters()]
    g32, gem = oracle_grads(lambda: orc.train_objective(orc.render_rays(rays, bg, nerf_pes, nerfs, nof_pes, nofs, **kw),
                                                        target), plist)
    pos = 0
    for tag, mod, bundle in mods:
        names = [n for n, _ in mod.named_parameters()]
        got = [(n, q.grad) for n, q in mod.named_parameters()]
        sl = slice(pos, pos + len(names))
        pos += len(names)
        # end to end the emulated oracle and the kernels can still pick different fine samples / mask members
        # (discontinuous steps), hence the wider band than in the module-level tests
        compare_grads(f"{name}.{tag}", got, list(zip(names, gem[sl])), 8e-2, list(zip(names, g32[sl])), min_cos=0.8)


def test_fused_gradient_accumulation_matches_autograd(dev):
    """dp.FlatGradients(fused_accumulate=True): gradients added by the scatter kernel == autograd accumulation."""
    import moco_flow_b200 as mf
    from moco_flow_b200 import backward_mlp, dp
    gen = torch.Generator().manual_seed(5)
    R, S = 9, 32
    xyz = ((torch.rand(R * S, 3, generator=gen) - 0.5) * 1.2).to(dev)
    ind = (torch.rand(R, 1, generator=gen) * 2 - 1).to(dev)
    up = torch.randn(R * S, 3, generator=gen).to(dev)
    m = mf.NoF(4, 128, 33, [2], "ind", 33, True)
    m.load_state_dict(orc.make_nof_params(orc.C2F_NOF, 41, scale_head=0.5))
    m = m.to(dev)
    pe, pe_i = mf.Embedding(3, 5), mf.Embedding(1, 16)

    def run():
        out = m.evaluate(xyz=xyz, pe=pe, ray_feat=pe_i(ind), rows_per_ray=S)
        out2 = m.evaluate(xyz=out, pe=pe, ray_feat=pe_i(ind), rows_per_ray=S)  # same module twice: grads accumulate
        (out2 * up).sum().backward()

    run()
    ref = [p.grad.clone() for p in m.parameters()]
    for p in m.parameters():
        p.grad = None
    try:
        flat = dp.FlatGradients([m], fused_accumulate=True)
        assert backward_mlp.ACCUMULATE_INTO_GRAD
        flat.zero()
        run()
        torch.cuda.synchronize()
        off = 0
        for p, r in zip(m.parameters(), ref):
            got = flat.buffer[off:off + p.numel()].view_as(p)
            off += p.numel()
            # fp32 atomics in the GEMM epilogue make the summation order vary between runs
            assert rel_fro(got, r) <= 1e-4
    finally:
        backward_mlp.ACCUMULATE_INTO_GRAD = False


@pytest.mark.parametrize("sigma_only", [True, False])
def test_nerf_gradient_wrt_embedded_inputs(dev, sigma_only):
    """The reference's own call convention with inputs that carry grad: NoF output -> Embedding -> zero-padded rows ->
    NeRF.forward(inputs[, sigma_only=True]) (trainer/trainer_moco_flow.py:146-158,349-361).  d loss / d inputs for
    the encoded-xyz columns (and, in full mode, the extra-feature columns) against autograd of the oracle."""
    import moco_flow_b200 as mf
    gen = torch.Generator().manual_seed(29)
    N = 300
    xyz = (torch.rand(N, 3, generator=gen) - 0.5) * 1.6
    ncol = 63 if sigma_only else 68
    up = torch.randn(N, 1 if sigma_only else 4, generator=gen)
    p = orc.make_nerf_params(orc.C2F_NERF, 13, dense=True)
    m = mf.NeRF(8, 256, 63, [4], "ind", 5)
    m.load_state_dict(p)
    m = m.to(dev)
    pe = mf.Embedding(3, 10)
    extra = torch.rand(N, 5, generator=gen) * 2 - 1

    # device: xyz (leaf) -> Embedding (autograd through mcf_pe_fwd/bwd) -> padded rows -> module
    xd = xyz.to(dev).requires_grad_(True)
    rows = torch.zeros(N, ncol, device=dev)
    emb = pe(xd)
    rows[:, :emb.shape[1]] = emb
    if not sigma_only:
        ed = extra.to(dev).requires_grad_(True)
        rows = torch.cat([rows[:, :63], ed], dim=1)
    rows.retain_grad()
    out = m(rows, sigma_only=sigma_only)
    (out * up.to(dev)).sum().backward()
    torch.cuda.synchronize()
    from moco_flow_b200 import _lib as L
    assert L.device_error_flag() == 0

    po = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    xo = xyz.clone().requires_grad_(True)
    eo = extra.clone().requires_grad_(True)

    def run():
        feats = orc.positional_encoding(xo, orc.PESpec(3, 10))
        if not sigma_only:
            feats = torch.cat([feats, eo], 1)
        o = orc.nerf_mlp(po, orc.C2F_NERF, feats, sigma_only=sigma_only)
        return (o * up).sum()
    leaves = [xo] + ([eo] if not sigma_only else [])
    g32, gem = oracle_grads(run, leaves)
    got = [("d_xyz (through Embedding)", xd.grad)] + ([("d_extra", ed.grad)] if not sigma_only else [])
    names = [n for n, _ in got]
    compare_grads(f"nerf-embedded-inputs[sigma_only={sigma_only}]", got, list(zip(names, gem)), 3e-2, list(zip(names, g32)))


def test_weight_gradient_overlap_matches_serial(dev, monkeypatch):
    """backward_mlp.DW_OVERLAP_SMS: weight-gradient GEMMs on a side stream next to the following dX chains (which then
    leave SMs free) -- same gradients as the serial schedule, also when replayed from a CUDA graph."""
    import moco_flow_b200 as mf
    from moco_flow_b200 import backward_mlp, dp
    from moco_flow_b200.graph import CudaGraphStep
    gen = torch.Generator().manual_seed(77)
    R, S = 96, 64
    rays = orc.make_rays(R, seed=5, chained=True).to(dev)
    bg = torch.ones(R, 3, device=dev)
    tgt = torch.rand(R, 3, generator=gen).to(dev)
    dr = orc.make_draws(R, S, S, seed=6)
    draws = mf.Draws(*(t.to(dev) for t in (dr.perturb, dr.noise_coarse, dr.u, dr.noise_fine)))
    nerfs, nofs = [], []
    for s_ in (1, 2):
        m = mf.NeRF(8, 256, 63, [4], "ind", 5)
        m.load_state_dict(orc.make_nerf_params(orc.C2F_NERF, s_, dense=True))
        nerfs.append(m.to(dev))
    for s_ in (3, 4):
        m = mf.NoF(4, 128, 33, [2], "ind", 33, True)
        m.load_state_dict(orc.make_nof_params(orc.C2F_NOF, s_, scale_head=0.25))
        nofs.append(m.to(dev))
    nerf_embs = [mf.Embedding(3, 10), mf.Embedding(1, 2), None]
    nof_embs = [mf.Embedding(3, 5), mf.Embedding(1, 16)]
    try:
        flat = dp.FlatGradients(nerfs + nofs, fused_accumulate=True)

        def step(rays_, bg_, tgt_):
            flat.zero()
            res = mf.render_rays(rays_, bg_, nerf_embs, nerfs, nof_embeddings=nof_embs, nof_models=nofs, chain_local=True,
                                 chain_global=True, N_samples=S, N_importance=S, perturb=1.0, noise_std=0.0, draws=draws,
                                 fused_residual_mean=True)
            loss = mf.MSELoss()(res, tgt_) + 0.2 * sum(res[k].mean() for k in res if "disp" in k)
            loss.backward()
            flat.check_views()      # joins the side stream
            return loss.detach()

        step(rays, bg, tgt)
        torch.cuda.synchronize()
        serial = flat.buffer.clone()
        monkeypatch.setattr(backward_mlp, "DW_OVERLAP_SMS", 48)
        step(rays, bg, tgt)
        torch.cuda.synchronize()
        assert rel_fro(flat.buffer, serial) <= 1e-4
        g = CudaGraphStep(step, [rays, bg, tgt], warmup=1)
        flat.buffer.fill_(123.0)
        g(rays, bg, tgt)
        torch.cuda.synchronize()
        from moco_flow_b200 import _lib as L
        assert L.device_error_flag() == 0
        assert rel_fro(flat.buffer, serial) <= 1e-4
    finally:
        backward_mlp.ACCUMULATE_INTO_GRAD = False
        backward_mlp.join_weight_gradients()
