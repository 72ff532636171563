#!/usr/bin/env python
"""How far apart may two FAITHFUL bf16 evaluations of the reference algorithm be?  (CPU only, oracle only.)

Runs the oracle (models/rendering.py:195-375 restated) on the dense synthetic scene of the parity tests and compares
  (a) fp32, coarse weights handed to sample_pdf multiplied by (1 + 1e-6 * N(0,1)): sample_pdf itself is well conditioned;
  (b) bf16 tensor-core emulation vs fp32: what any bf16 MLP path pays;
  (c) bf16 emulation vs bf16 emulation with every value multiplied by (1 + 1e-6 * N(0,1)) before rounding -- the size
      of an fp32 accumulation-order difference (tensor core vs torch's CPU GEMM).  A fraction ~1e-6 / 2^-9 of the
      activations then round to the other bf16 neighbour, the random-init density field (10 octaves of positional
      encoding: it varies on the scale of the sample spacing) turns that into per-sample weight changes of ~5e-3,
      the inverse-cdf step moves fine samples by up to ~0.1, and depth follows.  (c) is the yardstick the GPU
      parity tests hold the kernels to: a kernel is correct when it is no further from the emulating oracle than the
      emulating oracle is from its own 1e-6 perturbation.

    python scripts/depth_sensitivity.py [n_rays] > profiles/r02_depth_sensitivity.json
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import moco_oracle as orc  # noqa: E402


def main():
    R = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    Sc = Sf = 64
    pes = orc.C2F_PE
    nerfs = [orc.NeRFBundle(orc.C2F_NERF, orc.make_nerf_params(orc.C2F_NERF, s, dense=True)) for s in (1, 2)]
    nofs = [orc.NoFBundle(orc.C2F_NOF, orc.make_nof_params(orc.C2F_NOF, s, scale_head=0.25)) for s in (3, 4)]
    rays, bg = orc.make_rays(R, seed=1, chained=True), torch.ones(R, 3)
    dr = orc.make_draws(R, Sc, Sf, seed=2)
    kw = dict(N_samples=Sc, N_importance=Sf, perturb=1.0, noise_std=0.0, test_time=True)
    args = (rays, bg, [pes["nerf_xyz"], pes["nerf_ind"], None], nerfs, [pes["nof_xyz"], pes["nof_ind"]], nofs)

    def run(perturb_w=0.0, emulate=False, noise=0.0):
        orig = orc.sample_pdf
        if perturb_w:
            g = torch.Generator().manual_seed(99)

            def noisy(bins, weights, *a, **k):
                return orig(bins, weights * (1 + perturb_w * torch.randn(weights.shape, generator=g)), *a, **k)
            orc.sample_pdf = noisy
        orc.EMULATE_BF16, orc.EMULATE_NOISE = emulate, noise
        try:
            with torch.no_grad():
                return orc.render_rays(*args, draws=dr, return_aux=True, **kw)
        finally:
            orc.sample_pdf = orig
            orc.EMULATE_BF16, orc.EMULATE_NOISE = False, 0.0

    base = run()
    emu = run(emulate=True)
    out = {"rays": R, "scene": "dense NeRF (sigma ~ N(10,5)), bw-NoF, 64+64, perturb=1 with injected draws"}
    for tag, base, other in (("a_fp32_coarse_weights_x(1+1e-6*N)_vs_fp32", base, run(perturb_w=1e-6)),
                             ("b_bf16_emulation_vs_fp32", base, emu),
                             ("c_bf16_emulation_with_1e-6_noise_vs_bf16_emulation", emu, run(emulate=True, noise=1e-6))):
        rec = {}
        for k in ("opacity_coarse", "rgb_fine", "depth_fine", "opacity_fine"):
            d = (other[k] - base[k]).abs().reshape(R, -1).amax(1)
            rec[k] = {"median": float(d.median()), "p99": float(d.quantile(0.99)), "p999": float(d.quantile(0.999)),
                      "max": float(d.max())}
        dz = (other["_aux"]["z_new"] - base["_aux"]["z_new"]).abs()
        rec["coarse_weight_change_max"] = float((other["_aux"]["weights_coarse"] - base["_aux"]["weights_coarse"]).abs().max())
        rec["fine_sample_shift_max"] = float(dz.max())
        rec["fine_sample_shift_over_1e-3"] = int((dz > 1e-3).sum())
        rec["coarse_bin_width"] = 1.6 / (Sc - 1)
        out[tag] = rec
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
