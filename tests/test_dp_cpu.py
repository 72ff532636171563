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
    flat.broadcast_parameters(src=0)
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
