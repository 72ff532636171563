"""Checkpoint layout of the reference: trainer/base.py:279-327, trainer/trainer_moco_flow.py:46-70."""
import pytest
import torch

import moco_flow_b200 as mf
from moco_flow_b200 import checkpoint as ck
from moco_flow_b200 import dp
from oracle import moco_oracle as orc


def _nets(seed0=0):
    nets = {"coarse_NeRF": mf.NeRF(8, 256, 63, [4], "ind", 5), "fine_NeRF": mf.NeRF(8, 256, 63, [4], "ind", 5),
            "bw_NoF": mf.NoF(4, 128, 33, [2], "ind", 33, True), "fw_NoF": mf.NoF(4, 128, 33, [2], "ind", 33, True)}
    nets["coarse_NeRF"].load_state_dict(orc.make_nerf_params(orc.C2F_NERF, seed0 + 1))
    nets["fine_NeRF"].load_state_dict(orc.make_nerf_params(orc.C2F_NERF, seed0 + 2))
    nets["bw_NoF"].load_state_dict(orc.make_nof_params(orc.C2F_NOF, seed0 + 3))
    nets["fw_NoF"].load_state_dict(orc.make_nof_params(orc.C2F_NOF, seed0 + 4))
    return nets


def test_checkpoint_layout_roundtrip_and_reference_names(tmp_path):
    nets = _nets()
    path = ck.save_ckpt(str(tmp_path / "epoch1_iter10"), nets, clock={"epoch": 1, "step": 10})
    assert path.endswith(".pth")
    blob = torch.load(path)
    assert set(blob) == {"clock", "coarse_NeRF_net", "fine_NeRF_net", "bw_NoF_net", "fw_NoF_net"}
    # the parameter names a reference checkpoint holds (models/nerf.py:30-58, models/nof.py:42-53)
    assert set(blob["fine_NeRF_net"]) == set(orc.make_nerf_params(orc.C2F_NERF, 0))
    assert set(blob["bw_NoF_net"]) == set(orc.make_nof_params(orc.C2F_NOF, 0))
    other = _nets(seed0=10)
    clock = ck.load_ckpt(str(tmp_path / "epoch1_iter10"), other)
    assert clock == {"epoch": 1, "step": 10}
    for key in nets:
        for (n, a), (_, b) in zip(nets[key].state_dict().items(), other[key].state_dict().items()):
            assert torch.equal(a, b), (key, n)
    with pytest.raises(ValueError):
        ck.load_ckpt(str(tmp_path / "missing"), other)


def test_load_pretrained_nerf_takes_trunk_and_density_only(tmp_path):
    nets = _nets()
    path = ck.save_ckpt(str(tmp_path / "pre"), nets)
    target = mf.NeRF(8, 256, 63, [4], "ind", 5)
    before = {k: v.clone() for k, v in target.state_dict().items()}
    ck.load_pretrained_model(target, "fine_NeRF_net", path)    # the "only load fine NeRF" trick of the reference
    src = nets["fine_NeRF"].state_dict()
    for k, v in target.state_dict().items():
        if "xyz" in k or "sigma" in k:
            assert torch.equal(v, src[k]), k
        else:
            assert torch.equal(v, before[k]), k           # colour branch untouched
    nof = mf.NoF(4, 128, 33, [2], "ind", 33, True)
    ck.load_pretrained_model(nof, "bw_NoF_net", path)
    assert all(torch.equal(a, b) for a, b in zip(nof.state_dict().values(), nets["bw_NoF"].state_dict().values()))
    with pytest.raises(ValueError):
        ck.load_pretrained_model(nof, "no_such_net", path)


def test_flattened_parameters_survive_a_checkpoint_load(tmp_path):
    nets = _nets()
    path = ck.save_ckpt(str(tmp_path / "flat"), nets)
    other = _nets(seed0=20)
    flat = dp.FlatGradients(list(other.values()), flatten_params=True)
    ck.load_ckpt(path, other)
    # load_state_dict copies into the existing storage: the parameters are still views of the flat buffer
    off = 0
    for p in flat.params:
        assert p.data_ptr() == flat.param_buffer.data_ptr() + 4 * off
        off += p.numel()
    ref = torch.cat([p.detach().reshape(-1) for n in nets.values() for p in n.parameters()])
    assert torch.equal(flat.param_buffer, ref)


@pytest.mark.gpu
def test_fused_adam_state_roundtrip_with_torch_adam(tmp_path):
    """A torch.optim.Adam state (what a reference checkpoint holds) continues in FusedAdam, and FusedAdam's own
    state_dict is a per-parameter torch.optim.Adam-style state."""
    from moco_flow_b200.optim import FusedAdam
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.Linear(5, 3)).to(dev)
    ref_net = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.Linear(5, 3)).to(dev)
    ref_net.load_state_dict(net.state_dict())
    ref = torch.optim.Adam(ref_net.parameters(), lr=1e-2, eps=1e-8)
    gen = torch.Generator().manual_seed(1)
    grads = [[torch.randn(p.shape, generator=gen).to(dev) for p in net.parameters()] for _ in range(6)]
    for it in range(3):                      # three steps with torch's Adam, saved the reference's way
        for p, g in zip(ref_net.parameters(), grads[it]):
            p.grad = g.clone()
        ref.step()
    path = ck.save_ckpt(str(tmp_path / "opt"), {"net": ref_net}, optimizers={"moco": ref})
    flat = dp.FlatGradients([net], flatten_params=True)
    opt = FusedAdam(flat.params, lr=1e-2, eps=1e-8)
    ck.load_ckpt(path, {"net": net}, optimizers={"moco": opt}, map_location=dev)
    for it in range(3, 6):                   # both continue
        for p, q, g in zip(flat.params, ref_net.parameters(), grads[it]):
            p.grad.copy_(g)
            q.grad = g.clone()
        opt.step()
        ref.step()
    torch.cuda.synchronize()
    for p, q in zip(flat.params, ref_net.parameters()):
        assert (p - q).abs().max().item() <= 2e-6
    sd = opt.state_dict()
    assert int(float(sd["state"][0]["step"])) == 6
    assert sd["state"][0]["exp_avg"].shape == flat.params[0].shape
    assert sd["state"][0]["exp_avg"].untyped_storage().nbytes() == flat.params[0].numel() * 4   # a clone, not a view
