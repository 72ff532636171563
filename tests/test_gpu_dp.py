"""2-GPU data-parallel parity (run with: gpurun --gpus 2 -- python -m pytest tests/test_gpu_dp.py -m gpu).

SURVEY 8(e): each rank renders a contiguous shard of the step's ray batch; image terms are means over equal shards,
the masked flow-residual means are taken over the rays of ALL ranks (one all-reduce of the (sum, count) pairs), the
flat gradient buffer is summed by one NCCL all-reduce and scaled by 1/world.  The result must equal the gradients of
one process rendering the concatenated batch (models/rendering.py:306-314,365-373 +
trainer/trainer_moco_flow.py:319-327 define the global means).
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import moco_oracle as orc

pytestmark = pytest.mark.gpu
R, SC, SF = 4096, 64, 64


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _scene(dev):
    import moco_flow_b200 as mf
    nerfs, nofs = [], []
    for s in (1, 2):
        m = mf.NeRF(8, 256, 63, [4], "ind", 5)
        m.load_state_dict(orc.make_nerf_params(orc.C2F_NERF, s, dense=True))
        nerfs.append(m.to(dev))
    for s in (3, 4):
        m = mf.NoF(4, 128, 33, [2], "ind", 33, True)
        m.load_state_dict(orc.make_nof_params(orc.C2F_NOF, s, scale_head=0.25))
        nofs.append(m.to(dev))
    return nerfs, nofs, [mf.Embedding(3, 10), mf.Embedding(1, 2), None], [mf.Embedding(3, 5), mf.Embedding(1, 16)]


def _step(dev, sl, world):
    """Gradients (flat buffer, already rank-averaged) and loss of the rays ``sl`` of the global batch."""
    import moco_flow_b200 as mf
    from moco_flow_b200 import dp
    nerfs, nofs, nerf_embs, nof_embs = _scene(dev)
    flat = dp.FlatGradients(nerfs + nofs, fused_accumulate=True)
    rays, bg = orc.make_rays(R, seed=1, chained=True)[sl].to(dev), torch.ones(R, 3)[sl].to(dev)
    tgt = torch.from_numpy(np.random.Generator(np.random.PCG64(7)).uniform(0, 1, (R, 3)).astype("float32"))[sl].to(dev)
    dr = orc.make_draws(R, SC, SF, seed=2)
    draws = mf.Draws(*(t[sl].to(dev) for t in (dr.perturb, dr.noise_coarse, dr.u, dr.noise_fine)))
    flat.zero()
    res = mf.render_rays(rays, bg, nerf_embs, nerfs, nof_embeddings=nof_embs, nof_models=nofs, chain_local=True,
                         chain_global=True, N_samples=SC, N_importance=SF, perturb=1.0, noise_std=0.0, draws=draws,
                         fused_residual_mean=True)
    loss = mf.MSELoss()(res, tgt)
    for key in ("nof_local_disp", "nof_global_disp"):
        loss = loss + 0.2 * (res[key + "_coarse"].mean() + res[key + "_fine"].mean())
    loss.backward()
    scale = flat.allreduce_sum()
    torch.cuda.synchronize()
    mf.check_device()
    means = {k: float(v.item()) for k, v in res.items() if "disp" in k}
    return flat.buffer * scale, float(loss.item()), means


def _worker(rank, world, port, out):
    from moco_flow_b200 import backward_mlp, dp
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    dp.enable_global_residual_means()
    b, e = dp.shard_bounds(R, rank, world)
    grad, loss, means = _step(dev, slice(b, e), world)
    # image terms are per-rank means: the global loss is their average; residual means are already global
    t = torch.tensor([loss], device=dev, dtype=torch.float64)
    dist.all_reduce(t)
    if rank == 0:
        torch.save(dict(grad=grad.cpu(), loss_sum=float(t.item()), means=means), out)
    dist.barrier()
    dist.destroy_process_group()
    backward_mlp.ACCUMULATE_INTO_GRAD = False


def test_two_gpu_gradients_equal_single_gpu(tmp_path):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 CUDA devices")
    from moco_flow_b200 import backward_mlp, ops
    out = str(tmp_path / "dp.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    blob = torch.load(out)
    ops.RESIDUAL_DP = None
    try:
        grad1, loss1, means1 = _step(torch.device("cuda:0"), slice(0, R), 1)
    finally:
        backward_mlp.ACCUMULATE_INTO_GRAD = False
    g2, g1 = blob["grad"].double(), grad1.cpu().double()
    rel = float((g2 - g1).norm() / g1.norm())
    print(f"[dp] 2 x 2048 rays vs 1 x 4096 rays: gradient rel Frobenius {rel:.3e}, residual means {blob['means']} vs {means1}")
    for k, v in means1.items():
        assert abs(blob["means"][k] - v) <= 1e-6 * abs(v) + 1e-9, k     # global masked means, same on every rank
    # the residual terms enter every rank's loss in full (they are global), the image terms as per-rank means
    resid = 0.2 * sum(means1.values())
    assert abs((blob["loss_sum"] - 2 * resid) / 2 + resid - loss1) <= 1e-5 * max(1.0, abs(loss1))
    # same tiles, same tensor-core arithmetic; only the fp32 atomic order of the weight-gradient reduction differs
    assert rel <= 2e-4
    out_json = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(out_json, exist_ok=True)
    import json
    json.dump({"grad_rel_fro_2x2048_vs_1x4096": rel, "means_2gpu": blob["means"], "means_1gpu": means1},
              open(os.path.join(out_json, "parity_dp_r02.json"), "w"), indent=1)
