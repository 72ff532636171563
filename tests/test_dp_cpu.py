"""world_size-2 gloo tests of the ray-sharded data-parallel host logic (CPU)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from moco_flow_b200 import dp


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 4096, 4097):
        for world in (1, 2, 3, 8):
            spans = [dp.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(rank)  # different init per rank: broadcast must fix it
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    dp.broadcast_parameters([net])
    flat = dp.FlatGradients([net])
    gen = torch.Generator().manual_seed(123)
    rays = torch.randn(10, 6, generator=gen)  # the same global batch on every rank
    target = torch.randn(10, 3, generator=gen)
    b, e = dp.shard_bounds(10, rank, world)
    flat.zero()
    loss = torch.nn.functional.mse_loss(net(rays[b:e]), target[b:e])
    loss.backward()
    assert all(p.grad.data_ptr() >= flat.buffer.data_ptr() for p in net.parameters())  # still views
    flat.allreduce_mean()
    if rank == 0:
        torch.save(dict(grad=flat.buffer.clone(), state={k: v.clone() for k, v in net.state_dict().items()},
                        rays=rays, target=target), out)
    dist.destroy_process_group()


def test_two_rank_gradients_match_single_process(tmp_path):
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    blob = torch.load(out)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    net.load_state_dict(blob["state"])
    # equal shards: mean of per-shard mean losses == loss on the concatenated batch
    torch.nn.functional.mse_loss(net(blob["rays"]), blob["target"]).backward()
    ref = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    assert torch.allclose(blob["grad"], ref, atol=1e-6)


def _worker_flat(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(rank)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    # flattened parameters: one broadcast of the flat buffer makes every rank identical
    flat = dp.FlatGradients([net], flatten_params=True)
    from moco_flow_b200 import ops
    epoch0 = ops.PARAM_EPOCH
    flat.broadcast_parameters(src=0)
    assert ops.PARAM_EPOCH > epoch0          # packed bf16 weight images must be rebuilt after a p.data write
    dp.enable_global_residual_means()
    assert ops.RESIDUAL_DP is not None and ops.RESIDUAL_DP[1] == world
    gen = torch.Generator().manual_seed(123)
    rays = torch.randn(10, 6, generator=gen)
    target = torch.randn(10, 3, generator=gen)
    b, e = dp.shard_bounds(10, rank, world)
    flat.zero()
    torch.nn.functional.mse_loss(net(rays[b:e]), target[b:e]).backward()
    scale = flat.allreduce_sum()       # sum only; the 1/world goes into the optimizer step
    assert scale == 1.0 / world
    if rank == 0:
        torch.save(dict(grad=flat.buffer.clone() * scale, params=flat.param_buffer.clone(),
                        state={k: v.clone() for k, v in net.state_dict().items()}, rays=rays, target=target), out)
    dist.destroy_process_group()


def test_two_rank_flat_parameters_and_sum_allreduce(tmp_path):
    out = str(tmp_path / "r0_flat.pt")
    mp.spawn(_worker_flat, args=(2, _free_port(), out), nprocs=2, join=True)
    blob = torch.load(out)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    net.load_state_dict(blob["state"])
    assert torch.equal(torch.cat([p.detach().reshape(-1) for p in net.parameters()]), blob["params"])
    torch.nn.functional.mse_loss(net(blob["rays"]), blob["target"]).backward()
    ref = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    assert torch.allclose(blob["grad"], ref, atol=1e-6)


def test_zero_grad_keeps_flat_views():
    """FusedAdam.zero_grad() must zero in place (torch's default drops .grad, after which the backward
    would write gradients the flat all-reduce buffer never sees); a dropped view is reported, not ignored."""
    from moco_flow_b200.optim import FusedAdam
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    flat = dp.FlatGradients([net])
    opt = FusedAdam(flat.params, lr=1e-3)
    torch.nn.functional.mse_loss(net(torch.randn(4, 6)), torch.randn(4, 3)).backward()
    assert float(flat.buffer.abs().sum()) > 0
    opt.zero_grad()
    assert float(flat.buffer.abs().sum()) == 0
    flat.check_views()
    with pytest.raises(RuntimeError):
        opt.zero_grad(set_to_none=True)
    net[0].weight.grad = None
    with pytest.raises(RuntimeError):
        flat.allreduce_sum()


def test_embedding_device_table_tracks_weights():
    """the encoder weights the kernels read live in a device table that follows ``Embedding.weights``
    (CPU check of the host logic with a CPU 'device')."""
    import moco_flow_b200 as mf
    from moco_flow_b200 import _lib as L
    e = mf.Embedding(3, 4)
    t = e.device_table("cpu")
    assert t.shape == (2 * L.MAX_FREQS,) and t[:4].tolist() == [1, 2, 4, 8] and t[L.MAX_FREQS:L.MAX_FREQS + 4].tolist() == [1] * 4
    ptr = t.data_ptr()
    e.weights = [1.0, 0.5, 0.0, 0.0]         # what the coarse-to-fine schedule does every step
    e.sync_device()
    t2 = e.device_table("cpu")
    assert t2.data_ptr() == ptr and t2[L.MAX_FREQS:L.MAX_FREQS + 4].tolist() == [1.0, 0.5, 0.0, 0.0]
