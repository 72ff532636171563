"""Parity at the BASELINE.json sizes (run on the B200 box: pytest -m gpu).

The small fixtures of test_gpu_chain / test_gpu_grad launch 2-24 tiles; the benchmark launches 2048-6144 tiles per
chain kernel on 148 persistent CTAs (7-28 work units per CTA: weight-ring stage/phase wrap across units, bulk-store
hand-over between units, CTA-pair padding tiles, the 74-CTA split-K weight-gradient jobs).  These tests run exactly
those shapes against the CPU oracle (models/rendering.py:195-375 restated), in fp32 and in bf16-emulation mode:

  * cfg2  -- 4096 rays, 64+64, bw-NoF -> NeRF -> composite, test_time            (BASELINE configs[1])
  * cfg3  -- 4096 rays, 64+64, both flow chains, loss + every parameter gradient (BASELINE configs[2])
  * cfg5  -- 512 rays, 128+128, chain_local (fw o bw), rows_per_ray 128 / 256    (BASELINE configs[4] shape)
  * sample_pdf at 64/64 and 128/128: index mismatches against the reference formula == 0

Every error figure is also written to gpurun_out/parity_r02.json (copied to profiles/r02_parity.json).
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import moco_oracle as orc
from tests.test_gpu_grad import compare_grads, cosine, rel_fro

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RECORD = {}


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    yield torch.device("cuda:0")
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    path = os.path.join(out, "parity_r02.json")
    old = {}
    if os.path.exists(path):
        try:
            old = json.load(open(path))
        except Exception:
            old = {}
    old.update(RECORD)
    json.dump(old, open(path, "w"), indent=1, sort_keys=True)


def _no_device_error():
    from moco_flow_b200 import _lib as L
    flag = L.device_error_flag()
    assert flag == 0, hex(flag)


PES = orc.C2F_PE
NERF_PES = [PES["nerf_xyz"], PES["nerf_ind"], None]
NOF_PES = [PES["nof_xyz"], PES["nof_ind"]]


def make_scene(dev, requires_grad=False):
    import moco_flow_b200 as mf
    nerf_p = [orc.make_nerf_params(orc.C2F_NERF, s, dense=True) for s in (1, 2)]
    nof_p = [orc.make_nof_params(orc.C2F_NOF, s, scale_head=0.25) for s in (3, 4)]
    nerfs, nofs = [], []
    for p in nerf_p:
        m = mf.NeRF(8, 256, 63, [4], "ind", 5)
        m.load_state_dict(p)
        nerfs.append(m.to(dev))
    for p in nof_p:
        m = mf.NoF(4, 128, 33, [2], "ind", 33, True)
        m.load_state_dict(p)
        nofs.append(m.to(dev))
    if requires_grad:
        for p in nerf_p + nof_p:
            for k in p:
                p[k] = p[k].clone().requires_grad_(True)
    o_nerfs = [orc.NeRFBundle(orc.C2F_NERF, p) for p in nerf_p]
    o_nofs = [orc.NoFBundle(orc.C2F_NOF, p) for p in nof_p]
    nerf_embs = [mf.Embedding(3, 10), mf.Embedding(1, 2), None]
    nof_embs = [mf.Embedding(3, 5), mf.Embedding(1, 16)]
    return (nerfs, nofs, nerf_embs, nof_embs), (o_nerfs, o_nofs)


def cuda_draws(dr, dev):
    import moco_flow_b200 as mf
    return mf.Draws(*(None if t is None else t.to(dev) for t in (dr.perturb, dr.noise_coarse, dr.u, dr.noise_fine)))


def slice_draws(dr, sl):
    return orc.RenderDraws(*(None if t is None else t[sl] for t in (dr.perturb, dr.noise_coarse, dr.u, dr.noise_fine)))


def oracle_render_chunked(rays, bg, o_nerfs, o_nofs, draws, kw, chunk=1024):
    """Per-ray outputs of orc.render_rays evaluated in ray chunks (bounded host memory).  The flow-residual vectors
    are returned as (sum, count) over the whole batch -- their mean is what the trainer consumes
    (trainer/trainer_moco_flow.py:319-327)."""
    per_ray, sums = {}, {}
    R = rays.shape[0]
    with torch.no_grad():
        for b in range(0, R, chunk):
            sl = slice(b, min(b + chunk, R))
            res = orc.render_rays(rays[sl], bg[sl], NERF_PES, o_nerfs, NOF_PES, o_nofs, draws=slice_draws(draws, sl),
                                  return_aux=True, **kw)
            aux = res.pop("_aux")
            for k, v in res.items():
                if "disp" in k:
                    al = aux["alphas_coarse" if k.endswith("coarse") else "alphas_fine"]
                    assert bool((al >= 0.01).any()), "chunk without a selected sample: all-true fallback would differ"
                    s, n = sums.get(k, (0.0, 0))
                    sums[k] = (s + float(v.double().sum()), n + v.numel())
                else:
                    per_ray.setdefault(k, []).append(v)
    return {k: torch.cat(v) for k, v in per_ray.items()}, sums


def emulated_pair(rays, bg, o_nerfs, o_nofs, dr, kw, chunk=1024):
    """The bf16-emulating oracle, and the same with 1e-6 relative noise before every rounding (the yardstick for how
    far two faithful bf16 evaluations may be apart, scripts/depth_sensitivity.py)."""
    orc.EMULATE_BF16 = True
    try:
        emu, sums = oracle_render_chunked(rays, bg, o_nerfs, o_nofs, dr, kw, chunk)
        orc.EMULATE_NOISE = 1e-6
        noisy, _ = oracle_render_chunked(rays, bg, o_nerfs, o_nofs, dr, kw, chunk)
    finally:
        orc.EMULATE_BF16, orc.EMULATE_NOISE = False, 0.0
    return emu, sums, noisy


def record_outputs(tag, got, emu, ref, rec, noisy=None):
    """Error figures of one output dict against the bf16-emulating and the fp32 oracle."""
    for k in sorted(ref):
        g, e, r = got[k].detach().cpu().double(), emu[k].double(), ref[k].double()
        rel32 = (g - r).abs() / r.abs().clamp_min(1e-3)
        d16 = (g - e).abs().reshape(-1)
        rec[k] = {
            "max_abs_vs_bf16_emulated": float(d16.max()),
            "p99_abs_vs_bf16_emulated": float(d16.quantile(0.99)),
            "p999_abs_vs_bf16_emulated": float(d16.quantile(0.999)),
            "emulated_oracle_max_abs_vs_fp32": float((e - r).abs().max()),
            "max_abs_vs_fp32": float((g - r).abs().max()),
            "rel_vs_fp32_median": float(rel32.median()), "rel_vs_fp32_p999": float(rel32.quantile(0.999)),
            "rel_vs_fp32_max": float(rel32.max()),
            "emulated_oracle_rel_vs_fp32_median": float(((e - r).abs() / r.abs().clamp_min(1e-3)).median()),
            "emulated_oracle_rel_vs_fp32_max": float(((e - r).abs() / r.abs().clamp_min(1e-3)).max()),
            "ref_scale": float(r.abs().max()),
        }
        if noisy is not None:
            dn = (noisy[k].double() - e).abs().reshape(-1)
            rec[k].update(emulated_self_noise_p99=float(dn.quantile(0.99)), emulated_self_noise_max=float(dn.max()))
        print(f"[scale] {tag}.{k}: " + ", ".join(f"{a}={b:.3e}" for a, b in rec[k].items()))


def assert_outputs(rec):
    for k, v in rec.items():
        if not isinstance(v, dict) or "max_abs_vs_fp32" not in v:
            continue
        if k.startswith("rgb") or k.startswith("opacity"):
            assert v["max_abs_vs_bf16_emulated"] <= 5e-4, (k, v)
            assert v["max_abs_vs_fp32"] <= 2e-3, (k, v)      # north star: <= 2e-3 on rgb for the bf16 MLP path
        elif k.startswith("depth"):
            # Depth of a random-init volume is the ill-conditioned output: an fp32 accumulation-order difference
            # (1e-6) flips a few bf16 roundings, the density field (10 octaves of encoding) turns that into ~5e-3
            # changes of single coarse weights, the inverse-cdf step moves fine samples by up to ~0.1 and depth by
            # ~1e-2 on the worst rays (profiles/r02_depth_sensitivity.json).  So the yardstick is the emulating oracle's
            # distance from its own 1e-6 perturbation: the kernel must be no further from the emulating oracle than 2x
            # that (99th percentile and max); against fp32 the north-star 2e-3 (relative) is held at the median ray and
            # the worst ray is no further from fp32 than 1.5x what the emulating oracle is.
            assert v["p99_abs_vs_bf16_emulated"] <= 2 * v["emulated_self_noise_p99"] + 1e-4, (k, v)
            assert v["max_abs_vs_bf16_emulated"] <= 2 * v["emulated_self_noise_max"] + 1e-4, (k, v)
            assert v["rel_vs_fp32_median"] <= 2e-3, (k, v)
            assert v["rel_vs_fp32_max"] <= max(2e-3, 1.5 * v["emulated_oracle_rel_vs_fp32_max"]), (k, v)


# --------------------------------------------------------------------------------------------------------------
# cfg2: 4096-ray inference render
# --------------------------------------------------------------------------------------------------------------
def test_cfg2_render_4096(dev):
    import moco_flow_b200 as mf
    R, Sc, Sf = 4096, 64, 64
    (nerfs, nofs, nerf_embs, nof_embs), (o_nerfs, o_nofs) = make_scene(dev)
    rays, bg = orc.make_rays(R, seed=1, chained=True), torch.ones(R, 3)
    dr = orc.make_draws(R, Sc, Sf, seed=2)
    kw = dict(N_samples=Sc, N_importance=Sf, perturb=1.0, noise_std=0.0, test_time=True)
    runs = []
    with torch.no_grad():
        for _ in range(2):
            runs.append(mf.render_rays(rays.to(dev), bg.to(dev), nerf_embs, nerfs, nof_embeddings=nof_embs,
                                       nof_models=nofs, draws=cuda_draws(dr, dev), **kw))
    torch.cuda.synchronize()
    _no_device_error()
    for k in runs[0]:   # the forward path has no atomics: two launches are bit-identical
        assert torch.equal(runs[0][k], runs[1][k]), k
    ref, _ = oracle_render_chunked(rays, bg, o_nerfs, o_nofs, dr, kw)
    emu, _, noisy = emulated_pair(rays, bg, o_nerfs, o_nofs, dr, kw)
    assert sorted(runs[0]) == sorted(ref)
    rec = RECORD.setdefault("cfg2_render_4096x64+64_test_time", {"deterministic": True})
    record_outputs("cfg2", runs[0], emu, ref, rec, noisy)
    assert_outputs(rec)


# --------------------------------------------------------------------------------------------------------------
# cfg3: 4096-ray training step (loss + every parameter gradient)
# --------------------------------------------------------------------------------------------------------------
def oracle_train_chunked(rays, bg, target, o_nerfs, o_nofs, draws, kw, counts, chunk=1024):
    """Loss and parameter gradients of orc.train_objective over the whole batch, accumulated chunk by chunk:
    image terms are means over R*3, chain terms are masked sums divided by the global counts of a first pass."""
    R = rays.shape[0]
    total = 0.0
    for b in range(0, R, chunk):
        sl = slice(b, min(b + chunk, R))
        res = orc.render_rays(rays[sl], bg[sl], NERF_PES, o_nerfs, NOF_PES, o_nofs, draws=slice_draws(draws, sl), **kw)
        loss = ((res["rgb_coarse"] - target[sl]) ** 2).sum() / (R * 3) + ((res["rgb_fine"] - target[sl]) ** 2).sum() / (R * 3)
        for k, n in counts.items():
            loss = loss + 0.2 * res[k].sum() / n
        loss.backward()
        total += float(loss.detach())
    return total


def test_cfg3_train_4096(dev):
    import moco_flow_b200 as mf
    R, Sc, Sf = 4096, 64, 64
    (nerfs, nofs, nerf_embs, nof_embs), (o_nerfs, o_nofs) = make_scene(dev, requires_grad=True)
    rays, bg = orc.make_rays(R, seed=1, chained=True), torch.ones(R, 3)
    target = torch.from_numpy(np.random.Generator(np.random.PCG64(7)).uniform(0, 1, (R, 3)).astype("float32"))
    dr = orc.make_draws(R, Sc, Sf, seed=2)
    kw = dict(chain_local=True, chain_global=True, N_samples=Sc, N_importance=Sf, perturb=1.0, noise_std=0.0)
    res = mf.render_rays(rays.to(dev), bg.to(dev), nerf_embs, nerfs, nof_embeddings=nof_embs, nof_models=nofs,
                         draws=cuda_draws(dr, dev), fused_residual_mean=True, **kw)
    loss = mf.MSELoss()(res, target.to(dev))
    for key in ("nof_local_disp", "nof_global_disp"):
        loss = loss + 0.2 * (res[key + "_coarse"].mean() + res[key + "_fine"].mean())
    loss.backward()
    torch.cuda.synchronize()
    _no_device_error()

    rec = RECORD.setdefault("cfg3_train_4096x64+64_both_chains", {})
    mods = [("nerf0", nerfs[0], o_nerfs[0]), ("nerf1", nerfs[1], o_nerfs[1]), ("nof0", nofs[0], o_nofs[0]),
            ("nof1", nofs[1], o_nofs[1])]
    results = {}
    for mode in ("fp32", "bf16_emulated"):
        orc.EMULATE_BF16 = mode != "fp32"
        try:
            per_ray, sums = oracle_render_chunked(rays, bg, o_nerfs, o_nofs, dr, kw)
            counts = {k: n for k, (s, n) in sums.items()}
            for _, _, b in mods:
                for p in b.params.values():
                    p.grad = None
            ol = oracle_train_chunked(rays, bg, target, o_nerfs, o_nofs, dr, kw, counts)
        finally:
            orc.EMULATE_BF16 = False
        grads = {tag: {n: (None if p.grad is None else p.grad.clone()) for n, p in b.params.items()} for tag, _, b in mods}
        results[mode] = (per_ray, sums, ol, grads)
    ref, rsums, rloss, g32 = results["fp32"]
    emu, esums, eloss, gem = results["bf16_emulated"]
    got = {k: v for k, v in res.items() if "disp" not in k}
    _, _, noisy = emulated_pair(rays, bg, o_nerfs, o_nofs, dr, kw)
    record_outputs("cfg3", got, emu, ref, rec, noisy)
    assert_outputs(rec)
    for k, (s, n) in esums.items():
        g = float(res[k].item())
        rec[k] = {"masked_mean": g, "bf16_emulated": s / n, "fp32": rsums[k][0] / rsums[k][1],
                  "count_bf16_emulated": n, "count_fp32": rsums[k][1]}
        print(f"[scale] cfg3.{k}: {rec[k]}")
        assert abs(g - s / n) <= 2e-3 * abs(s / n), k
        assert abs(g - rsums[k][0] / rsums[k][1]) <= 1e-2 * abs(rsums[k][0] / rsums[k][1]), k
    rec["loss"] = {"kernel": float(loss.item()), "bf16_emulated": eloss, "fp32": rloss}
    print(f"[scale] cfg3.loss: {rec['loss']}")
    assert abs(loss.item() - eloss) <= 5e-4 * max(1.0, abs(eloss))
    assert abs(loss.item() - rloss) <= 2e-3 * max(1.0, abs(rloss))
    grec = rec.setdefault("gradients", {})
    for tag, mod, _ in mods:
        for n, q in mod.named_parameters():
            r16, r32 = gem[tag][n], g32[tag][n]
            if r16 is None or q.grad is None:
                continue
            grec[f"{tag}.{n}"] = {"rel_fro_vs_bf16_emulated": rel_fro(q.grad, r16), "rel_fro_vs_fp32": rel_fro(q.grad, r32),
                                  "cos_vs_fp32": cosine(q.grad, r32),
                                  "emulated_oracle_rel_fro_vs_fp32": rel_fro(r16, r32)}
    worst = max(grec.items(), key=lambda kv: kv[1]["rel_fro_vs_bf16_emulated"])
    print(f"[scale] cfg3 gradients: worst rel Frobenius vs bf16-emulated {worst[0]} {worst[1]}")
    for tag, mod, _ in mods:
        names = [n for n, _ in mod.named_parameters()]
        named_got = [(n, q.grad) for n, q in mod.named_parameters()]
        compare_grads(f"cfg3.{tag}", named_got, [(n, gem[tag][n]) for n in names], 5e-2,
                      [(n, g32[tag][n]) for n in names], min_cos=0.95)


# --------------------------------------------------------------------------------------------------------------
# cfg5 shape: 128 + 128 samples, forward-o-backward flow consistency (chain_local), no grad
# --------------------------------------------------------------------------------------------------------------
def test_cfg5_shape_128_128_chain_local(dev):
    import moco_flow_b200 as mf
    R, Sc, Sf = 512, 128, 128
    (nerfs, nofs, nerf_embs, nof_embs), (o_nerfs, o_nofs) = make_scene(dev)
    rays, bg = orc.make_rays(R, seed=11, chained=True), torch.ones(R, 3)
    dr = orc.make_draws(R, Sc, Sf, seed=12)
    kw = dict(chain_local=True, N_samples=Sc, N_importance=Sf, perturb=1.0, noise_std=0.0)
    with torch.no_grad():
        res = mf.render_rays(rays.to(dev), bg.to(dev), nerf_embs, nerfs, nof_embeddings=nof_embs, nof_models=nofs,
                             draws=cuda_draws(dr, dev), **kw)
        fused = mf.render_rays(rays.to(dev), bg.to(dev), nerf_embs, nerfs, nof_embeddings=nof_embs, nof_models=nofs,
                               draws=cuda_draws(dr, dev), fused_residual_mean=True, **kw)
    torch.cuda.synchronize()
    _no_device_error()
    ref, rsums = oracle_render_chunked(rays, bg, o_nerfs, o_nofs, dr, kw, chunk=256)
    emu, esums, noisy = emulated_pair(rays, bg, o_nerfs, o_nofs, dr, kw, chunk=256)
    rec = RECORD.setdefault("cfg5shape_512x128+128_chain_local", {})
    record_outputs("cfg5", {k: v for k, v in res.items() if "disp" not in k}, emu, ref, rec, noisy)
    assert_outputs(rec)
    for k, (s, n) in esums.items():
        vec = res[k]
        rec[k] = {"selected": int(vec.numel()), "selected_bf16_emulated": n, "selected_fp32": rsums[k][1],
                  "mean": float(vec.mean()), "fused_mean": float(fused[k].item()), "mean_bf16_emulated": s / n,
                  "mean_fp32": rsums[k][0] / rsums[k][1]}
        print(f"[scale] cfg5.{k}: {rec[k]}")
        assert abs(vec.numel() - n) <= max(2, 1e-3 * n), k       # alpha >= 0.01 membership, bf16-level ties only
        assert abs(float(vec.mean()) - s / n) <= 2e-3 * abs(s / n), k
        assert abs(float(fused[k].item()) - float(vec.mean())) <= 1e-5 * abs(float(vec.mean())), k


# --------------------------------------------------------------------------------------------------------------
# sample_pdf: product path == reference formula, index for index
# --------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("Sc,Sf", [(64, 64), (128, 128)])
def test_sample_pdf_indices_exact_at_scale(dev, Sc, Sf):
    """Level 2 of the exactness contract: the product path (reference-order cdf, then the fused
    search / gather / lerp / sort-merge kernel) against the reference formula evaluated on the same device:
    every index and every sample identical, deterministic and random u.  The in-kernel fixed-order cdf (opt-in)
    is compared with the CPU oracle and its mismatch count recorded."""
    from moco_flow_b200 import ops
    gen = torch.Generator().manual_seed(40 + Sc)
    R = 8192
    z = torch.sort(torch.rand(R, Sc, generator=gen) * 1.6 + 2.0, dim=1).values
    wts = torch.rand(R, Sc, generator=gen) ** 6
    wts[::7] *= 1e-4                      # nearly empty rays: the eps term dominates
    wts[::11, Sc // 2:] = 0.0             # exact zeros: repeated cdf entries, denom < eps branch
    u = torch.rand(R, Sf, generator=gen)
    zd, wd, ud = z.to(dev), wts.to(dev), u.to(dev)
    mid_d = 0.5 * (zd[:, :-1] + zd[:, 1:])
    rec = RECORD.setdefault(f"sample_pdf_{Sc}+{Sf}_R{R}", {})
    for det in (False, True):
        u_in = torch.linspace(0, 1, Sf, device=dev).expand(R, Sf).contiguous() if det else ud
        ref, aux = orc.sample_pdf(mid_d, wd[:, 1:-1], Sf, det=det, u=None if det else ud, return_aux=True)
        s, inds, cdf, merged = ops.sample_pdf_raw(zd, wd, u_in, bins_are_z=True, w_offset=1, n_bins=Sc - 2,
                                                  z_coarse=zd, want_inds=True, want_cdf=True, fused_cdf=False)
        mism = int((inds.long() != aux["inds"]).sum())
        rec[f"reference_order_cdf_det={det}"] = {"index_mismatches": mism, "of": R * Sf,
                                                 "samples_bit_identical": bool(torch.equal(s, ref))}
        assert torch.equal(cdf, aux["cdf"])
        assert mism == 0
        assert torch.equal(s, ref)
        exp = torch.sort(torch.cat([zd, ref], dim=1), dim=1).values
        assert torch.equal(merged, exp)
        # opt-in fixed-order in-kernel cdf against the CPU oracle: report, bound loosely
        refc, auxc = orc.sample_pdf(0.5 * (z[:, :-1] + z[:, 1:]), wts[:, 1:-1], Sf, det=det, u=None if det else u,
                                    return_aux=True)
        _, inds_f, _, _ = ops.sample_pdf_raw(zd, wd, u_in, bins_are_z=True, w_offset=1, n_bins=Sc - 2,
                                             want_inds=True, fused_cdf=True)
        # the last deterministic draw u = 1.0 is decided by the last bit of cdf[-1] in the reference itself
        cols = slice(0, Sf - 1) if det else slice(0, Sf)
        mf_ = int((inds_f.cpu().long()[:, cols] != auxc["inds"][:, cols]).sum())
        rec[f"fused_cdf_vs_cpu_oracle_det={det}"] = {"index_mismatches": mf_, "of": R * (Sf - 1 if det else Sf)}
        print(f"[scale] sample_pdf {Sc}+{Sf} det={det}: reference-order mismatches {mism}, fused-cdf mismatches vs CPU oracle {mf_}")
        assert mf_ <= 1e-4 * R * Sf
    _no_device_error()
