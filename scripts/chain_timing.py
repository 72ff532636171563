#!/usr/bin/env python
"""In-kernel cycle breakdown of the chain kernel (GPU box): where each warp role spends its time."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _build_timing_library() -> str:
    """The cycle counters are compiled in only with -DMCF_TIMING: build that variant next to the product library."""
    import subprocess
    from moco_flow_b200 import build as B
    B.build()
    csrc = B.CSRC
    obj = os.path.join(csrc, "chain_timing.o")
    lib = os.path.join(csrc, "libmoco_flow_b200_timing.so")
    nvcc = B._nvcc()
    subprocess.run([nvcc, *B.ARCH, *B.COMMON, "-DMCF_TIMING", "-c", os.path.join(csrc, "chain.cu"), "-o", obj], check=True)
    others = [os.path.join(csrc, u.replace(".cu", ".o")) for u, _ in B.UNITS if u != "chain.cu"]
    subprocess.run([nvcc, *B.ARCH, "-shared", "-o", lib, obj, *others], check=True)
    return lib


_TIMING_LIB = _build_timing_library()
import torch  # noqa: E402
from moco_flow_b200 import _lib as _L  # noqa: E402
_L.LIB_PATH = _TIMING_LIB   # before the first call loads the library

import bench  # noqa: E402
import moco_flow_b200 as mf  # noqa: E402
from moco_flow_b200 import ops  # noqa: E402

NAMES = ["s0 prologue", "s0 wait acc", "s0 epilogue", "s0 save/bar", "s1 prologue", "s1 wait acc", "s1 epilogue",
         "s1 save/bar", "mma wait act", "mma wait w", "mma issue", "prod wait ring", "total"]


def main():
    dev = torch.device("cuda:0")
    nerfs, nofs, nerf_embs, nof_embs = bench.build_models(dev)
    rays, bg, tgt = (t.to(dev) for t in bench.synth_batch(4096, 1))

    def train():
        res = mf.render_rays(rays, bg, nerf_embs, nerfs, nof_embeddings=nof_embs, nof_models=nofs, chain_local=True,
                             chain_global=True, N_samples=64, N_importance=64, perturb=1.0, noise_std=0.0,
                             fused_residual_mean=True)
        loss = mf.MSELoss()(res, tgt) + 0.2 * sum(res[k].mean() for k in res if "disp" in k)
        loss.backward()

    def render():
        with torch.no_grad():
            mf.render_rays(rays, bg, nerf_embs, nerfs, nof_embeddings=nof_embs, nof_models=nofs, N_samples=64,
                           N_importance=64, perturb=1.0, noise_std=0.0, test_time=True)

    for name, fn in (("render", render), ("train", train)):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        ops.TIMING = {}
        fn()
        torch.cuda.synchronize()
        timing, ops.TIMING = ops.TIMING, None
        print(f"=== {name} ===")
        for tag, bufs in timing.items():
            for i, b in enumerate(bufs):
                t = b[:148, :13].double().cpu()
                active = t[:, 12] > 0
                m = t[active].mean(0)
                tot = m[12].item()
                parts = "  ".join(f"{n}={100*v/tot:4.1f}%" for n, v in zip(NAMES[:12], m[:12].tolist()))
                print(f"{tag}[{i}] ctas={int(active.sum())} total={tot/1e3:.0f}k cyc | {parts}")


if __name__ == "__main__":
    main()
