"""Fused Adam: host-side segment logic on CPU, kernel parity against torch.optim.Adam on the GPU."""
import pytest
import torch

from moco_flow_b200 import dp
from moco_flow_b200.optim import FusedAdam, merge_segments


def _nets():
    torch.manual_seed(0)
    a = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.Linear(5, 3))
    b = torch.nn.Linear(3, 2, bias=False)
    return [a, b]


def test_merge_segments_adjacency():
    # (param_ptr, grad_ptr, numel): 0-2 contiguous in both spaces, 3 contiguous in params only, 4 empty
    items = [(1000, 5000, 10), (1040, 5040, 6), (1064, 5064, 1), (1068, 9000, 4), (2000, 7000, 0)]
    segs = merge_segments(items)
    assert [(s[0], s[1], s[2]) for s in segs] == [(1000, 5000, 17), (1068, 9000, 4)]
    assert segs[0][3] == [0, 1, 2] and segs[1][3] == [3]
    # order of the input does not matter
    segs2 = merge_segments(list(reversed(items)))
    assert [(s[0], s[1], s[2]) for s in segs2] == [(1000, 5000, 17), (1068, 9000, 4)]


def test_flatten_params_keeps_values_and_names():
    nets = _nets()
    before = {k: v.clone() for n in nets for k, v in n.state_dict().items()}
    names = [k for n in nets for k in n.state_dict()]
    flat = dp.FlatGradients(nets, flatten_params=True)
    assert flat.param_buffer.numel() == flat.buffer.numel() == sum(p.numel() for p in flat.params)
    after = {k: v for n in nets for k, v in n.state_dict().items()}
    assert [k for n in nets for k in n.state_dict()] == names
    for k in before:
        assert torch.equal(before[k], after[k])
    # parameters and gradients are adjacent views, in the same order: one Adam segment
    items = [(p.data_ptr(), p.grad.data_ptr(), p.numel()) for p in flat.params]
    assert len(merge_segments(items)) == 1
    # the modules still compute with the flattened storage and autograd still lands in the flat gradient buffer
    x = torch.randn(4, 7)
    nets[1](nets[0](x)).sum().backward()
    assert flat.buffer.abs().sum() > 0
    flat.param_buffer.zero_()
    assert all(float(p.detach().abs().sum()) == 0.0 for p in flat.params)


def test_fused_adam_rejects_cpu_parameters():
    nets = _nets()
    flat = dp.FlatGradients(nets, flatten_params=True)
    opt = FusedAdam(flat.params, lr=1e-3)
    with pytest.raises(RuntimeError):
        opt.step()   # no CPU fallback
    with pytest.raises(ValueError):
        FusedAdam(flat.params, lr=-1.0)


@pytest.mark.gpu
@pytest.mark.parametrize("weight_decay,flatten", [(0.0, True), (0.01, True), (0.0, False)])
def test_fused_adam_matches_torch_adam(weight_decay, flatten):
    """8 steps with fresh random gradients, a MultiStepLR decay in the middle and a gradient scale; tolerance 2e-6
    relative to the parameter scale (fp32 op-order differences only)."""
    dev = torch.device("cuda:0")
    nets = [n.to(dev) for n in _nets()]
    ref_nets = [n for n in _nets()]
    flat = dp.FlatGradients(nets, flatten_params=flatten)
    opt = FusedAdam(flat.params, lr=1e-2, eps=1e-8, weight_decay=weight_decay)
    ref_params = [p for n in ref_nets for p in n.parameters()]
    ref = torch.optim.Adam(ref_params, lr=1e-2, eps=1e-8, weight_decay=weight_decay)
    sched = torch.optim.lr_scheduler.MultiStepLR(opt, milestones=[3, 6], gamma=0.5)
    ref_sched = torch.optim.lr_scheduler.MultiStepLR(ref, milestones=[3, 6], gamma=0.5)
    assert len(opt._plan(0, opt.param_groups[0])["segs"]) == (1 if flatten else len(flat.params))
    gen = torch.Generator().manual_seed(5)
    for it in range(8):
        scale = 0.5 if it % 2 else 1.0
        for p, q in zip(flat.params, ref_params):
            g = torch.randn(p.shape, generator=gen)
            p.grad.copy_(g.to(dev))
            q.grad = g * scale
        opt.step(grad_scale=scale)
        ref.step()
        sched.step()
        ref_sched.step()
    torch.cuda.synchronize()
    assert opt.param_groups[0]["lr"] == ref.param_groups[0]["lr"] == 1e-2 * 0.25
    for p, q in zip(flat.params, ref_params):
        err = (p.detach().cpu() - q.detach()).abs().max().item()
        assert err <= 2e-6 * max(1.0, q.abs().max().item()), (err, tuple(p.shape))
    st, rst = opt.state[flat.params[0]], ref.state[ref_params[0]]
    assert int(st["step"].item()) == 8 == int(rst["step"])
    assert torch.allclose(st["exp_avg"].cpu(), rst["exp_avg"], rtol=1e-5, atol=1e-7)
    assert torch.allclose(st["exp_avg_sq"].cpu(), rst["exp_avg_sq"], rtol=1e-5, atol=1e-9)


@pytest.mark.gpu
def test_fused_adam_update_reaches_the_packed_weights():
    """The update is written through raw pointers: the modules must notice and re-pack their bf16 weight images."""
    import moco_flow_b200 as mf
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    m = mf.NeRF(8, 256, 63, [4], "ind", 5).to(dev)
    flat = dp.FlatGradients([m], flatten_params=True)
    opt = FusedAdam(flat.params, lr=1e-2)
    x = torch.randn(300, 68, device=dev)
    with torch.no_grad():
        y0 = m(x).clone()
    flat.buffer.fill_(1.0)
    opt.step()
    with torch.no_grad():
        y1 = m(x)
    torch.cuda.synchronize()
    assert (y1 - y0).abs().max().item() > 1e-3
    # a CUDA-graph replay of the step keeps counting and keeps reading the device-side learning rate
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            opt.step()
    before = flat.param_buffer.clone()
    g.replay(); g.replay()
    torch.cuda.synchronize()
    assert int(opt.state[flat.params[0]]["step"].item()) == 3
    opt.param_groups[0]["lr"] = 0.0
    opt.sync_lr()
    frozen = flat.param_buffer.clone()
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(frozen, flat.param_buffer) and not torch.equal(before, frozen)


def test_reference_optimizer_and_scheduler_factories():
    """trainer/base.py:122-160: config dictionaries -> optimizer / scheduler, milestones divided by the world size."""
    from moco_flow_b200.optim import get_optimizer, get_scheduler
    nets = _nets()
    params = [p for n in nets for p in n.parameters()]
    opt = get_optimizer({'type': 'adam', 'lr': 1e-4, 'weight_decay': 0}, params)
    assert isinstance(opt, FusedAdam) and opt.defaults['eps'] == 1e-8 and opt.defaults['lr'] == 1e-4
    sgd = get_optimizer({'type': 'sgd', 'lr': 0.1, 'momentum': 0.9, 'weight_decay': 0.0}, params)
    assert isinstance(sgd, torch.optim.SGD)
    with pytest.raises(NotImplementedError):
        get_optimizer({'type': 'ranger', 'lr': 1e-3, 'weight_decay': 0}, params)
    sch = get_scheduler({'type': 'steplr', 'decay_step': [40, 80], 'decay_gamma': 0.5}, sgd, world_size=8)
    assert sorted(sch.milestones) == [5, 10]
    lrs = []
    for _ in range(12):
        lrs.append(sgd.param_groups[0]['lr'])
        sgd.step()
        sch.step()
    assert lrs[4] == 0.1 and abs(lrs[5] - 0.05) < 1e-12 and abs(lrs[10] - 0.025) < 1e-12
    for cfg in ({'type': 'explr', 'lr_decay': 0.9}, {'type': 'cosine', 'num_epochs': 10},
                {'type': 'poly', 'num_epochs': 10, 'poly_exp': 2.0}):
        assert get_scheduler(cfg, get_optimizer({'type': 'sgd', 'lr': 0.1, 'momentum': 0.0, 'weight_decay': 0.0}, params))
    with pytest.raises(NotImplementedError):
        get_scheduler({'type': 'nope'}, sgd)
