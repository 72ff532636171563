#!/usr/bin/env python
"""Benchmark of the MoCo-Flow ray-rendering hot path (contract: see the task statement / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ...]

A step of the default workload (BASELINE.json configs[2], the one the headline metric "train rays/s
(fwd+bwd, 64+64 spp)" is quoted on) is one MoCo-Flow training step on 4096 synthetic rays per GPU:
render_rays with both flow chains (5 NoF + 1 NeRF evaluations per sample, coarse 64 + fine 64 samples),
image MSE + 0.2 local + 0.2 global chain losses, backward, NCCL all-reduce of the flat gradient buffer
(N > 1) and one Adam step.  Other workloads (the remaining BASELINE.json configs, not the driver's bench line):
``cfg1`` configs[0] (canonical NeRF only, 1024 rays), ``render`` configs[1] (4096-ray MoCo-Flow render),
``frame`` configs[3] (one 540x540 frame per step, rays sharded over the GPUs, result gathered; 16 steps = the
16-frame job), ``stress`` configs[4] (1080x1080, 128+128, forward-o-backward flow consistency, sharded + gathered).

Prints ONE JSON line.  ``--impl reference`` times the CPU oracle port of the reference on the host cores
(``--device cuda [--tf32]`` runs the same eager-PyTorch port on one B200: "the reference on the same silicon").
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

RAYS_PER_GPU = 4096
N_COARSE, N_FINE = 64, 64     # rebound by set_workload() for the stress workload
N_FRAMES = 160
FRAME_HW = {"frame": (540, 540), "stress": (1080, 1080)}
CHUNK_RAYS = 1 << 17          # rays per render_rays call inside a full-frame step (the reference's inference loop
                              # chunks the same way, trainer/trainer_moco_flow.py:590-626)


def set_workload(workload: str) -> None:
    global N_COARSE, N_FINE
    N_COARSE, N_FINE = (128, 128) if workload == "stress" else (64, 64)


# --------------------------------------------------------------------------------------------------
# synthetic workload (shared by both arms)
# --------------------------------------------------------------------------------------------------
def synth_batch(n_rays: int, seed: int):
    from oracle import moco_oracle as orc
    import numpy as np
    rays = orc.make_rays(n_rays, seed=seed, n_frames=N_FRAMES, chained=True)
    g = np.random.Generator(np.random.PCG64(seed + 1000))
    target = torch.from_numpy(g.uniform(0, 1, size=(n_rays, 3)).astype("float32"))
    bg = torch.ones(n_rays, 3)
    return rays, bg, target


def render_kwargs(workload: str) -> dict:
    """render_rays keyword arguments of a workload (both arms)."""
    kw = dict(N_samples=N_COARSE, N_importance=N_FINE, perturb=1.0, noise_std=0.0)
    if workload == "train":
        kw.update(chain_local=True, chain_global=True)
    elif workload == "stress":
        kw.update(chain_local=True)        # forward-o-backward flow consistency pass, no grad
    elif workload != "cfg1":
        kw.update(test_time=True)
    return kw


def rays_per_rank(workload: str, rank: int, world: int) -> int:
    if workload in FRAME_HW:               # one frame per step, contiguous ray ranges per GPU (strong scaling)
        from moco_flow_b200 import dp
        h, w = FRAME_HW[workload]
        b, e = dp.shard_bounds(h * w, rank, world)
        return e - b
    return 1024 if workload == "cfg1" else RAYS_PER_GPU


def algorithmic_flops_per_ray(workload: str) -> float:
    """un-padded reference shapes, 2*MAC; backward counted as 2x forward."""
    nerf, nerf_sigma, nof = 1181184.0, 982528.0, 134400.0
    s_tot = N_COARSE + (N_COARSE + N_FINE)
    if workload in ("render", "frame"):  # test_time: sigma-only coarse + full fine, bw-NoF on every sample
        return nof * s_tot + nerf_sigma * N_COARSE + nerf * (N_COARSE + N_FINE)
    if workload == "cfg1":               # canonical NeRF only, full coarse + full fine
        return nerf * s_tot
    if workload == "stress":             # bw then fw flow on every sample (chain_local), full coarse + fine NeRF
        return 2 * nof * s_tot + nerf * s_tot
    return 3.0 * (5 * nof * s_tot + nerf * s_tot)


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
class Clocks:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, reasons, mx = [], set(), None
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx = float(f[2])
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                     f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        except OSError:
            pass
        if sm:
            # under load = the upper half of the samples (the sampler also sees idle gaps around the region)
            sm_sorted = sorted(sm)
            out.update(sm_mhz=statistics.median(sm_sorted[len(sm_sorted) // 2:]), sm_max_mhz=mx,
                       reasons=sorted(reasons), samples=len(sm))
        return out


def synthetic_weights():
    """Random-init weights of the c2f.yaml shapes (configs/.../c2f.yaml:44-102) for both arms, from a numpy stream so
    every rank and both arms hold the same values.  nn.Linear's default init leaves the fine-pass density head
    non-positive over the whole volume (every opacity exactly 0, the image = the background), which would make the
    self-check of the fine outputs vacuous; the density heads are therefore re-centred the way the parity tests do
    (oracle.make_nerf_params(dense=True)): same shapes, same arithmetic, semi-transparent rays."""
    from oracle import moco_oracle as orc
    nerf_p = [orc.make_nerf_params(orc.C2F_NERF, s, dense=True) for s in (1, 2)]
    nof_p = [orc.make_nof_params(orc.C2F_NOF, s, scale_head=0.25) for s in (3, 4)]
    return nerf_p, nof_p


def build_models(dev):
    import moco_flow_b200 as mf
    nerf_p, nof_p = synthetic_weights()
    nerfs, nofs = [], []
    for p in nerf_p:
        m = mf.NeRF(8, 256, 63, [4], "ind", 5)
        m.load_state_dict(p)
        nerfs.append(m.to(dev))
    for p in nof_p:
        m = mf.NoF(4, 128, 33, [2], "ind", 33, True)
        m.load_state_dict(p)
        nofs.append(m.to(dev))
    nerf_embs = [mf.Embedding(3, 10), mf.Embedding(1, 2), None]
    nof_embs = [mf.Embedding(3, 5), mf.Embedding(1, 16)]
    return nerfs, nofs, nerf_embs, nof_embs


def self_check(workload, dev, nerfs, nofs, nerf_embs, nof_embs, n_rays: int = 256):
    """The timed computation checked against the CPU oracle on a slice of the workload (same weights, same rays, same
    injected random draws): returns the error figures that go into the JSON line; raises if they are out of bounds --
    a throughput number of a wrong computation is not a result."""
    import moco_flow_b200 as mf
    from oracle import moco_oracle as orc
    rays, bg, tgt = synth_batch(n_rays, seed=1)
    dr = orc.make_draws(n_rays, N_COARSE, N_FINE, seed=2)
    draws = mf.Draws(*(t.to(dev) for t in (dr.perturb, dr.noise_coarse, dr.u, dr.noise_fine)))
    pes = orc.C2F_PE
    o_nerfs = [orc.NeRFBundle(orc.C2F_NERF, {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}) for m in nerfs]
    o_nofs = [orc.NoFBundle(orc.C2F_NOF, {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}) for m in nofs]
    train = workload == "train"
    kw = render_kwargs(workload)
    if workload == "cfg1":
        nofs = nof_embs = o_nofs = None
    from moco_flow_b200 import ops as _ops
    saved_dp, _ops.RESIDUAL_DP = _ops.RESIDUAL_DP, None   # a rank-local check: no collectives (only rank 0 runs it)
    try:
        with torch.no_grad():
            res = mf.render_rays(rays.to(dev), bg.to(dev), nerf_embs, nerfs, nof_embeddings=nof_embs, nof_models=nofs,
                                 draws=draws, fused_residual_mean=True, **kw)
    finally:
        _ops.RESIDUAL_DP = saved_dp
    with torch.no_grad():
        res = {k: v.cpu() for k, v in res.items()}
        out = {"rays": n_rays, "against": "oracle/moco_oracle.py on the host (fp32 reference algorithm, and the same with "
                                          "bf16 tensor-core emulation)"}
        refs = {}
        for mode in ("fp32", "bf16_emulated"):
            orc.EMULATE_BF16 = mode != "fp32"
            try:
                refs[mode] = orc.render_rays(rays, bg, [pes["nerf_xyz"], pes["nerf_ind"], None], o_nerfs,
                                             [pes["nof_xyz"], pes["nof_ind"]] if o_nofs else None, o_nofs, draws=dr, **kw)
            finally:
                orc.EMULATE_BF16 = False
    for key in sorted(res):
        for mode, ref in refs.items():
            if "disp" in key:   # masked residual means: the oracle returns the dynamic-length vector
                out[f"{key}_abs_vs_{mode}"] = abs(float(res[key].mean()) - float(ref[key].mean()))
            else:
                out[f"{key}_max_abs_vs_{mode}"] = float((res[key] - ref[key]).abs().max())
    if train:
        def objective(r, means):
            loss = ((r["rgb_coarse"] - tgt) ** 2).mean() + ((r["rgb_fine"] - tgt) ** 2).mean()
            for k in ("nof_local_disp", "nof_global_disp"):
                loss = loss + 0.2 * (means(r[k + "_coarse"]) + means(r[k + "_fine"]))
            return float(loss)
        out["loss"] = objective(res, lambda t: t.mean())
        for mode, ref in refs.items():
            out[f"loss_{mode}"] = objective(ref, lambda t: t.mean())
        if abs(out["loss"] - out["loss_bf16_emulated"]) > 2e-3 * max(1.0, abs(out["loss_bf16_emulated"])):
            raise SystemExit(f"bench self-check failed: loss {out}")
    # Random-init volumes are semi-transparent up to the far plane, where the reference's last sample (delta = 1e10)
    # turns alpha into the step function [sigma_last > 0]: a ray whose sigma_last is within bf16 rounding of zero may
    # legitimately flip, so the bound is on the fraction of rays, and the max is reported.
    bad = (res["rgb_fine"] - refs["bf16_emulated"]["rgb_fine"]).abs().amax(dim=1) > 1e-3
    if "rgb_coarse" in res:
        bad = bad | ((res["rgb_coarse"] - refs["bf16_emulated"]["rgb_coarse"]).abs().amax(dim=1) > 1e-3)
    out["rgb_rays_over_1e-3_vs_bf16_emulated"] = int(bad.sum())
    if float(bad.float().mean()) > 0.01:
        raise SystemExit(f"bench self-check failed: {out}")
    from moco_flow_b200 import _lib as L
    flag = L.device_error_flag()
    if flag:
        raise SystemExit(f"bench self-check: device error flag {flag:#x}")
    return out


def run_ours(args):
    import torch.distributed as dist
    import moco_flow_b200 as mf
    from moco_flow_b200 import _lib as L
    from moco_flow_b200 import dp
    from moco_flow_b200.optim import FusedAdam
    from moco_flow_b200.build import build
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    if rank == 0:
        build()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()
        dp.enable_global_residual_means()   # masked flow-residual means over the rays of all ranks
    L.lib()
    nerfs, nofs, nerf_embs, nof_embs = build_models(dev)
    dp.broadcast_parameters(nerfs + nofs)
    train = args.workload == "train"
    framed = args.workload in FRAME_HW
    R = rays_per_rank(args.workload, rank, world)
    r_max = rays_per_rank(args.workload, 0, world)   # first ranks hold the remainder: the gather pads to this
    kw = render_kwargs(args.workload)
    use_nofs = None if args.workload == "cfg1" else nofs
    use_nof_embs = None if args.workload == "cfg1" else nof_embs
    gathered = torch.empty(world * r_max, 5, device=dev) if framed else None
    # this rank's shard of the step's global batch (N * 4096 rays): weak scaling
    rays_h, bg_h, tgt_h = synth_batch(R, seed=1 + rank)
    rays_h, bg_h, tgt_h = rays_h.pin_memory(), bg_h.pin_memory(), tgt_h.pin_memory()
    rays_d, bg_d, tgt_d = rays_h.to(dev), bg_h.to(dev), tgt_h.to(dev)
    flat = dp.FlatGradients(nerfs + nofs, fused_accumulate=True, flatten_params=True) if train else None
    # the reference's optimizer (trainer/base.py:122-133: Adam, eps 1e-8) as one fused launch over the flat buffers
    opt = FusedAdam(flat.params, lr=5e-4, eps=1e-8) if train else None
    loss_fn = mf.MSELoss()

    def step(rays, bg, tgt):
        if train:
            flat.zero()
            res = mf.render_rays(rays, bg, nerf_embs, nerfs, nof_embeddings=nof_embs, nof_models=nofs,
                                 chain_local=True, chain_global=True, N_samples=N_COARSE, N_importance=N_FINE,
                                 perturb=1.0, noise_std=0.0, fused_residual_mean=True)
            loss = loss_fn(res, tgt)
            loss = loss + 0.2 * (res["nof_local_disp_coarse"].mean() + res["nof_local_disp_fine"].mean())
            loss = loss + 0.2 * (res["nof_global_disp_coarse"].mean() + res["nof_global_disp_fine"].mean())
            loss.backward()
            opt.step(grad_scale=flat.allreduce_sum())   # the 1/world of the gradient mean is folded into the update
            return loss.detach()
        with torch.no_grad():
            if not framed:
                res = mf.render_rays(rays, bg, nerf_embs, nerfs, nof_embeddings=use_nof_embs, nof_models=use_nofs,
                                     fused_residual_mean=True, **kw)
                return res["rgb_fine"].mean()
            # full-frame step: this rank's contiguous ray range in chunks, then the result ([rgb, depth, opacity] =
            # 20 B/ray) gathered on every rank -- the only communication of the inference partition
            mine = torch.zeros(r_max, 5, device=dev) if R < r_max else torch.empty(r_max, 5, device=dev)
            for b0 in range(0, R, CHUNK_RAYS):
                sl = slice(b0, min(b0 + CHUNK_RAYS, R))
                res = mf.render_rays(rays[sl], bg[sl], nerf_embs, nerfs, nof_embeddings=use_nof_embs,
                                     nof_models=use_nofs, fused_residual_mean=True, **kw)
                mine[sl, 0:3] = res["rgb_fine"]
                mine[sl, 3] = res["depth_fine"]
                mine[sl, 4] = res["opacity_fine"]
            if world > 1:
                dist.all_gather_into_tensor(gathered, mine)
                return gathered
            return mine

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step(rays_d, bg_d, tgt_d)
    sync_all()
    def dbg(msg):
        if os.environ.get("MCF_BENCH_DEBUG"):
            sys.stderr.write(f"[bench rank {rank}] {msg}\n")
            sys.stderr.flush()

    dbg("warm-up done")
    parity = None
    if rank == 0 and not args.no_self_check:
        parity = self_check(args.workload, dev, nerfs, nofs, nerf_embs, nof_embs)
        dbg(f"self-check: {parity}")
    eager_step = step
    graphed = False
    if not args.no_graph:
        gstep = None
        try:
            from moco_flow_b200.graph import CudaGraphStep
            gstep = CudaGraphStep(eager_step, [rays_d, bg_d, tgt_d], warmup=2,
                                  refresh=[e for e in nerf_embs + nof_embs if e is not None] + ([opt] if opt else []))
        except Exception as exc:  # report, then measure the eager path
            sys.stderr.write(f"[bench rank {rank}] CUDA graph capture failed ({exc!r}); timing the eager step\n")
            gstep = None
        ok = torch.tensor([1 if gstep is not None else 0], device=dev)
        if world > 1:  # every rank must take the same path, or their collectives no longer match
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 1:
            step = gstep
            graphed = True
            for _ in range(2):
                step(rays_d, bg_d, tgt_d)
        dbg(f"graph capture: {graphed}")
    sync_all()
    launches_per_step_eager = 0
    if graphed:  # the graph replays exactly the launches one eager step issues
        c0 = L.LAUNCHES
        eager_step(rays_d, bg_d, tgt_d)
        launches_per_step_eager = L.LAUNCHES - c0
        sync_all()

    # ---- device-resident timing (value) ----
    clocks = Clocks(local_rank)
    if rank == 0:
        clocks.start()
        time.sleep(0.3)
    sync_all()
    launches0 = L.LAUNCHES
    # The 4096-ray render's working set (~30 MB) would sit in the 126 MB L2 from one step to the next: write a
    # 256 MB buffer between its timed steps and bracket every step with its own pair of events.  The training step
    # and the full-frame render stream several GB per step, far more than L2, and are timed back to back.
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if args.workload in ("render", "cfg1") else None

    def timed_steps(fn):
        if flush is None:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(args.steps):
                fn()
            b.record()
            sync_all()
            return a.elapsed_time(b)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for a, b in evs:
            flush.fill_(1)
            a.record()
            fn()
            b.record()
        sync_all()
        return sum(a.elapsed_time(b) for a, b in evs)

    ms = timed_steps(lambda: step(rays_d, bg_d, tgt_d))
    launches = (launches_per_step_eager * args.steps) if graphed else (L.LAUNCHES - launches0)
    dbg("device-resident timing done")
    clk = clocks.stop() if rank == 0 else None

    # ---- end-to-end timing through the public API with host buffers ----
    sync_all()
    loss_host = torch.empty(world * r_max, 5, pin_memory=True) if framed else torch.empty((), pin_memory=True)

    def e2e_step():
        r = rays_h.to(dev, non_blocking=True)
        b = bg_h.to(dev, non_blocking=True)
        t = tgt_h.to(dev, non_blocking=True)
        out = step(r, b, t)
        loss_host.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    ms_e2e = timed_steps(e2e_step)
    dbg("e2e timing done")
    mf.check_device()   # a kernel that flagged a device-side error anywhere above fails the run

    # ---- frame workload only: pose -> device ray generation -> render -> canvas scatter -> image on the host ----
    ms_frame = None
    if args.workload == "frame":
        from moco_flow_b200 import camera
        b0, e0 = dp.shard_bounds(540 * 540, rank, world)
        pix = torch.arange(b0, e0, device=dev, dtype=torch.int64)
        pose = np.array([[1.0, 0.0, 0.0, 0.0], [0.0, 1.0, 0.0, 0.0], [0.0, 0.0, 1.0, 2.8]])
        bg_frame = torch.ones(e0 - b0, 3, device=dev)
        img_host = torch.empty(e0 - b0, 3, pin_memory=True)

        def frame_step():
            rays = camera.make_rays(540, 540, 700.0, [270.0, 270.0], pose, 2.0, 3.6, 0.25, pixel_index=pix)
            with torch.no_grad():
                res = mf.render_rays(rays, bg_frame, nerf_embs, nerfs, nof_embeddings=nof_embs, nof_models=nofs,
                                     N_samples=N_COARSE, N_importance=N_FINE, perturb=1.0, noise_std=0.0,
                                     test_time=True)
            img, _ = camera.scatter_canvas(bg_frame, None, res["rgb_fine"], res["depth_fine"], res["opacity_fine"])
            img_host.copy_(img, non_blocking=True)
            torch.cuda.current_stream().synchronize()

        frame_step()
        ms_frame = timed_steps(frame_step)
        dbg("frame pipeline timing done")

    # ---- per-kernel event timing for the roofline (extra steps, not part of the numbers above) ----
    # serial schedule here: with the weight-gradient side stream on, kernels of two streams share the GPU and a pair
    # of events around one launch would time both
    from moco_flow_b200 import backward_mlp
    overlap_sms, backward_mlp.DW_OVERLAP_SMS = backward_mlp.DW_OVERLAP_SMS, 0
    L.PROFILE = []
    for _ in range(2):
        eager_step(rays_d, bg_d, tgt_d)
    torch.cuda.synchronize()
    prof, L.PROFILE = L.PROFILE, None
    backward_mlp.DW_OVERLAP_SMS = overlap_sms
    agg = {}
    for tag, work, unit, a, b_ in prof:
        d = agg.setdefault(tag, dict(ms=0.0, work=0.0, unit=unit, n=0))
        d["ms"] += a.elapsed_time(b_)
        d["work"] += work
        d["n"] += 1

    times = torch.tensor([ms, ms_e2e, ms_frame or 0.0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms, ms_e2e, ms_frame_max = times.tolist()
    def finish():
        # Hard exit for multi-rank runs: tearing down NCCL communicators that live inside captured CUDA graphs
        # can block in ncclCommDestroy; there is nothing left to clean up at this point.
        sys.stdout.flush()
        sys.stderr.flush()
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier()
            os._exit(0)

    if rank != 0:
        finish()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    tensor_peak = peaks.get("bf16_tflops_sustained", 1400.0)
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    peak_src = "measured (MEASURED_PEAKS.json: sustained bf16, copy HBM)" if peaks else "fallback (B200_PROFILING.md)"
    total_rays = (FRAME_HW[args.workload][0] * FRAME_HW[args.workload][1] if framed else world * R) * args.steps
    kernels = {}
    for tag, d in agg.items():
        rate = d["work"] / (d["ms"] * 1e-3) if d["ms"] > 0 else 0.0
        kernels[tag] = {"launches_per_step": d["n"] // 2, "ms_per_step": round(d["ms"] / 2, 4),
                        "achieved": round(rate / (1e12 if d["unit"] == "flop" else 1e9), 2),
                        "unit": "TFLOP/s" if d["unit"] == "flop" else "GB/s"}
    dom = max(agg.items(), key=lambda kv: kv[1]["ms"])[0] if agg else None
    traffic_db = {}
    try:
        traffic_db = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        pass

    def roof(tag):
        d = agg[tag]
        tensor = d["unit"] == "flop"
        ach = d["work"] / (d["ms"] * 1e-3) / (1e12 if tensor else 1e9)
        pk = tensor_peak if tensor else hbm_peak
        out = {"kernel": tag, "bound": "tensor" if tensor else "hbm", "achieved": round(ach, 2), "peak": pk,
               "unit": "TFLOP/s" if tensor else "GB/s", "frac": round(ach / pk, 4), "traffic": None,
               "peak_source": peak_src, "share_of_step": round(d["ms"] / sum(v["ms"] for v in agg.values()), 3),
               "launches_per_step": d["n"] // 2,
               "algorithmic_per_launch": round(d["work"] / max(d["n"], 1), 1)}
        t = traffic_db.get(tag)
        if t:  # dram__bytes_read+write of one ncu --set full capture of this kernel (profiles/), per launch
            out["traffic"] = t.get("dram_bytes")
            if t.get("algorithmic_bytes"):   # algorithmic bytes of the SAME captured launch the traffic belongs to
                out["traffic_algorithmic"] = t.get("algorithmic_bytes")
            out["traffic_note"] = t.get("note")
        return out

    roofline = roof(dom) if dom else None
    # the fused MLP chains together (the tensor-core part of the step), for the tensor-pipe target of the north star
    chain_tags = [t for t in agg if t.startswith("nerf_") or t.startswith("nof_")]
    mlp = None
    if chain_tags:
        fl = sum(agg[t]["work"] for t in chain_tags)
        tm = sum(agg[t]["ms"] for t in chain_tags)
        mlp = {"kernels": sorted(chain_tags), "bound": "tensor", "achieved": round(fl / (tm * 1e-3) / 1e12, 2),
               "peak": tensor_peak, "unit": "TFLOP/s", "frac": round(fl / (tm * 1e-3) / 1e12 / tensor_peak, 4),
               "ms_per_step": round(tm / 2, 4)}
        best = max(chain_tags, key=lambda t: agg[t]["work"] / max(agg[t]["ms"], 1e-9))
        mlp["best_kernel"] = roof(best)
    flops_ray = algorithmic_flops_per_ray(args.workload)
    h2d = int(rays_h.numel() + bg_h.numel() + tgt_h.numel()) * 4
    line = {
        "metric": metric_text(args.workload),
        "value": round(total_rays / (ms * 1e-3), 1), "unit": "rays/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True,
        "scaling": "strong" if framed else "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": config_dict(args.workload, world),
        "arm": {"rays_this_rank": R, "cuda_graph": graphed,
                "dw_overlap_sms": overlap_sms if train else None,
                "gather": (f"all_gather of [rgb, depth, opacity] = 20 B/ray inside every timed step"
                           if framed and world > 1 else None)},
        "e2e": {"value": round(total_rays / (ms_e2e * 1e-3), 1), "unit": "rays/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": int(loss_host.numel()) * 4, "ms_per_step": round(ms_e2e / args.steps, 4)},
        "gpu_launches": launches,
        "parity": parity,
        "clocks": clk,
        "roofline": roofline,
        "roofline_mlp": mlp,
        "kernels": kernels,
        "model_tflops": round(total_rays * flops_ray / (ms * 1e-3) / 1e12, 2),
        "model_tensor_frac": round(total_rays * flops_ray / (ms * 1e-3) / 1e12 / (tensor_peak * world), 4),
    }
    if ms_frame is not None:
        line["frame_pipeline"] = {
            "ms_per_frame": round(ms_frame_max / args.steps, 4), "h2d_bytes_per_step": 48,
            "d2h_bytes_per_step": 540 * 540 * 12,
            "what": "camera pose (3x4, passed by value) -> mcf_make_rays -> render_rays(test_time) -> mcf_canvas_scatter "
                    "-> rgb image copied to pinned host memory; max over ranks"}
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args.workload, sample_rays=args.cpu_rays, steps=5, warmup=2)
    print(json.dumps(line))
    finish()


# --------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle port of the reference on the host cores
# --------------------------------------------------------------------------------------------------
def cpu_step_fn(workload: str, n_rays: int, device: str = "cpu"):
    """One step of the workload on ``n_rays`` rays through the oracle port (plain eager PyTorch fp32: the reference's
    algorithm op for op).  ``device='cuda'`` runs the same port on the GPU -- eager PyTorch on the same silicon."""
    from oracle import moco_oracle as orc
    torch.manual_seed(0)
    pes = orc.C2F_PE
    nerf_p, nof_p = synthetic_weights()
    train = workload == "train"
    for p in nerf_p + nof_p:
        for k in p:
            p[k] = p[k].to(device).requires_grad_(train)
    nerfs = [orc.NeRFBundle(orc.C2F_NERF, p) for p in nerf_p]
    nofs = [orc.NoFBundle(orc.C2F_NOF, p) for p in nof_p] if workload != "cfg1" else None
    nof_pes = [pes["nof_xyz"], pes["nof_ind"]] if nofs else None
    rays, bg, tgt = (t.to(device) for t in synth_batch(n_rays, seed=1))
    params = [v for p in nerf_p + nof_p for v in p.values()]
    opt = torch.optim.Adam(params, lr=5e-4, eps=1e-8) if train else None
    kw = render_kwargs(workload)

    def step():
        with torch.device(device):     # the port creates its linspace / rand / zeros tensors on the default device
            if train:
                opt.zero_grad(set_to_none=True)
                res = orc.render_rays(rays, bg, [pes["nerf_xyz"], pes["nerf_ind"], None], nerfs, nof_pes, nofs, **kw)
                loss = orc.train_objective(res, tgt)
                loss.backward()
                opt.step()
                return float(loss.detach())
            with torch.no_grad():
                res = orc.render_rays(rays, bg, [pes["nerf_xyz"], pes["nerf_ind"], None], nerfs, nof_pes, nofs, **kw)
            return float(res["rgb_fine"].mean())     # float(): also the device synchronisation of the cuda variant
    return step


def time_cpu_steps(step, steps: int, warmup: int):
    """(median s/step, all step times): warm-up steps untimed, every timed step measured on its own."""
    for _ in range(warmup):
        step()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    return statistics.median(ts), ts


def cpu_baseline(workload: str, sample_rays: int, steps: int, warmup: int):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    med, ts = time_cpu_steps(cpu_step_fn(workload, sample_rays), steps, warmup)
    return {"value": round(sample_rays / med, 2), "unit": "rays/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{sample_rays} rays of the same workload per step (cost is linear in rays), median of {steps} timed "
                      f"steps after {warmup} warm-up, {med:.2f} s/step; the reference is a Python/PyTorch code base, so "
                      f"the CPU arm is the oracle port (oracle/moco_oracle.py, torch fp32 eager, all host threads)"}


def metric_text(workload: str) -> str:
    if workload == "train":
        return "train rays/s (fwd+bwd, 64+64 spp)"
    if workload == "stress":
        return "render rays/s (128+128 spp, fw-o-bw flow consistency pass)"
    if workload == "cfg1":
        return "render rays/s (canonical NeRF only, 64+64 spp)"
    return "render rays/s (64+64 spp, test_time)"


def workload_text(workload: str) -> str:
    """The workload description both arms put into ``config`` (same text, so the two JSON lines name the same job)."""
    if workload == "train":
        return ("MoCo-Flow training step fwd+bwd (BASELINE configs[2]): 4096 rays/GPU, 64+64 samples, "
                "bw/fw NoF chains (local+global), random-init c2f.yaml shapes, Adam step, "
                "ray-sharded DP + flat-gradient NCCL all-reduce")
    if workload == "frame":
        return ("full-frame inference render (BASELINE configs[3]): one 540x540 frame per step (16 steps = the 16-frame "
                "job), rays sharded over the GPUs, result gathered, 64+64 samples, test_time; ms_per_step = ms/frame")
    if workload == "stress":
        return ("stress render (BASELINE configs[4]): one 1080x1080 frame per step, 128+128 samples, bw-NoF -> NeRF plus "
                "the forward-o-backward flow consistency pass (chain_local), no grad, rays sharded over the GPUs, "
                "result gathered")
    if workload == "cfg1":
        return ("canonical NeRF render_rays forward (BASELINE configs[0]): 1024 rays, 64 coarse + 64 fine samples, "
                "random-init 8x256 MLP, no flow networks")
    return "full MoCo-Flow ray render (BASELINE configs[1]): 4096 rays, 64+64 samples, test_time"


def config_dict(workload: str, world: int) -> dict:
    """``config`` of the JSON line -- identical in both arms (what differs between them is under ``arm``)."""
    if workload in FRAME_HW:
        rays = f"{FRAME_HW[workload][0] * FRAME_HW[workload][1]} per step over all GPUs"
    else:
        rays = f"{1024 if workload == 'cfg1' else RAYS_PER_GPU} per GPU per step"
    return {"workload": workload_text(workload), "rays": rays, "n_coarse": N_COARSE, "n_fine": N_FINE,
            "parallelism": f"dp{world} (rays sharded, weights replicated)",
            "l2": ("GPU arm: a 256 MB buffer is written between timed steps (outside the per-step events)"
                   if workload in ("render", "cfg1") else
                   "GPU arm: no flush -- one step streams several GB per GPU (saved operand images / sample tensors), "
                   "far above the 126 MB L2")}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    on_gpu = args.device == "cuda"
    if on_gpu:
        torch.backends.cuda.matmul.allow_tf32 = bool(args.tf32)
        torch.backends.cudnn.allow_tf32 = bool(args.tf32)
    full = 1024 if args.workload == "cfg1" else RAYS_PER_GPU
    n = min(args.cpu_rays, full) if not on_gpu else (args.cpu_rays if args.cpu_rays_given else full)
    if args.workload in FRAME_HW and not args.cpu_rays_given:
        n = 1024
    med, ts = time_cpu_steps(cpu_step_fn(args.workload, n, "cuda" if on_gpu else "cpu"), args.steps, args.warmup)
    dt = sum(ts)
    val = round(n * args.steps / dt, 2)
    where = (f"eager PyTorch on cuda:0 (TF32 {'on' if args.tf32 else 'off'})" if on_gpu
             else f"{torch.get_num_threads()} host threads")
    sample = (f"each step = {n} rays of the workload"
              + ("" if n == full else " (bounded sample; cost is linear in rays)")
              + f", oracle port on {where}; median step {med:.3f} s")
    line = {
        "impl": "reference",
        "metric": metric_text(args.workload),
        "value": val, "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(dt / args.steps * 1e3, 2), "higher_is_better": True,
        "scaling": "strong" if args.workload in FRAME_HW else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args.workload, args.gpus),
        "arm": {"implementation": "oracle port of the reference (oracle/moco_oracle.py, torch fp32 eager)",
                "device": "cuda" if on_gpu else "cpu", "tf32": bool(args.tf32) if on_gpu else None, "rays_per_step": n,
                "median_rays_per_s": round(n / med, 2)},
        "cpu_baseline": {"value": val, "unit": "rays/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": sample},
        "e2e": {"value": val, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="train", choices=["train", "render", "frame", "stress", "cfg1"],
                    help="train: configs[2] step; render: configs[1] 4096-ray render; frame: one 540x540 frame per step "
                         "(configs[3]; 16 steps = the 16-frame job); stress: configs[4] 1080x1080 128+128 with the "
                         "flow-consistency pass; cfg1: configs[0] canonical NeRF only, 1024 rays")
    ap.add_argument("--cpu-rays", type=int, default=None,
                    help="rays per CPU-arm step (bounded sample; default 1024 = SURVEY 8d's slice)")
    ap.add_argument("--device", default="cpu", choices=["cpu", "cuda"],
                    help="--impl reference only: run the eager-PyTorch port on the host (default) or on cuda:0")
    ap.add_argument("--tf32", type=int, default=0, help="--impl reference --device cuda: allow TF32 matmuls")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-self-check", action="store_true", help="skip the oracle comparison of the timed computation")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying a CUDA graph")
    args = ap.parse_args()
    args.cpu_rays_given = args.cpu_rays is not None
    if args.cpu_rays is None:
        args.cpu_rays = 1024
    set_workload(args.workload)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
