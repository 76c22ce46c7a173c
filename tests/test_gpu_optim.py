"""-m gpu: the fused optimizer steps (asn_sgd_step / asn_adam_step) against torch.optim run sequentially
(foreach=False) in fp64 on the CPU -- the semantics of the reference's for-loop optimizers, duplicated
parameter-group entries included.  fp32 kernels: 1e-6 relative after several steps."""
import numpy as np
import pytest
import torch
import torch.nn as nn

from conftest import rel_err
from gpu_util import gpu

pytestmark = gpu


def _nets():
    torch.manual_seed(0)
    net = nn.Sequential(nn.Conv2d(3, 5, 3), nn.Conv2d(5, 7, 1), nn.Conv2d(7, 2, 3, bias=False))
    net[1].bias.requires_grad = False          # frozen parameters stay out of the flat buffer
    ref = nn.Sequential(nn.Conv2d(3, 5, 3), nn.Conv2d(5, 7, 1), nn.Conv2d(7, 2, 3, bias=False)).double()
    ref.load_state_dict({k: v.double() for k, v in net.state_dict().items()})
    ref[1].bias.requires_grad = False
    return net.cuda(), ref


def _groups(net):
    # group 0 names conv0's parameters three times and conv1's twice (as the reference's module walk does), group 1 once
    g0 = list(net[0].parameters()) * 3 + [p for p in net[1].parameters() if p.requires_grad] * 2
    return [{"params": g0, "lr": 0.01}, {"params": list(net[2].parameters()), "lr": 0.1}]


def test_fused_sgd_matches_sequential_torch():
    from adaptsegnet_b200.optim import FlatParams, FusedSGD
    net, ref = _nets()
    flat = FlatParams(net.parameters())
    assert flat.numel % 4 == 0 and all(b % 4 == 0 for b in flat.begin)
    opt = FusedSGD(flat, _groups(net), lr=0.01, momentum=0.9, weight_decay=5e-4)
    with pytest.warns(UserWarning):
        ropt = torch.optim.SGD(_groups(ref), lr=0.01, momentum=0.9, weight_decay=5e-4, foreach=False)
    assert sorted(opt.repeat) == [1, 2, 3, 3]
    rng = np.random.default_rng(1)
    for step in range(4):
        lr0 = 0.01 * (1 - step / 10) ** 0.9
        for o in (opt, ropt):
            o.param_groups[0]["lr"], o.param_groups[1]["lr"] = lr0, lr0 * 10
        v0 = [p._version for p in flat.params]
        for p, q in zip(net.parameters(), ref.parameters()):
            if not p.requires_grad:
                continue
            g = rng.standard_normal(tuple(p.shape))
            p.grad.copy_(torch.from_numpy(g))          # the views into the flat gradient buffer
            q.grad = torch.from_numpy(g.astype(np.float32).astype(np.float64))
        opt.step()
        ropt.step()
        assert all(p._version > v for p, v in zip(flat.params, v0))
        for (k, p), q in zip(net.named_parameters(), ref.parameters()):
            assert rel_err(p.detach().cpu().numpy(), q.detach().numpy()) < 1e-6, (step, k)
    assert torch.equal(net[1].bias.cpu().double(), ref[1].bias)      # frozen: untouched


def test_fused_adam_matches_torch():
    from adaptsegnet_b200.optim import FlatParams, FusedAdam
    net, ref = _nets()
    flat = FlatParams(net.parameters())
    opt = FusedAdam(flat, lr=1e-4, betas=(0.9, 0.99))
    ropt = torch.optim.Adam([p for p in ref.parameters() if p.requires_grad], lr=1e-4, betas=(0.9, 0.99), foreach=False)
    rng = np.random.default_rng(2)
    for step in range(5):
        for o in (opt, ropt):
            o.param_groups[0]["lr"] = 1e-4 * (1 - step / 10) ** 0.9
        for p, q in zip(net.parameters(), ref.parameters()):
            if not p.requires_grad:
                continue
            g = rng.standard_normal(tuple(p.shape)) * 10.0 ** rng.integers(-4, 1)
            p.grad.copy_(torch.from_numpy(g))
            q.grad = torch.from_numpy(g.astype(np.float32).astype(np.float64))
        opt.step()
        ropt.step()
        for (k, p), q in zip(net.named_parameters(), ref.parameters()):
            assert rel_err(p.detach().cpu().numpy(), q.detach().numpy()) < 1e-6, (step, k)


def test_flat_params_keep_module_semantics():
    """parameters are views into one buffer: forward/backward, load_state_dict and channels_last layouts keep working"""
    from adaptsegnet_b200.optim import FlatParams
    torch.manual_seed(3)
    net = nn.Sequential(nn.Conv2d(4, 8, 3, padding=1), nn.ReLU(), nn.Conv2d(8, 4, 1)).cuda().to(memory_format=torch.channels_last)
    x = torch.randn(2, 4, 9, 9, device="cuda")
    y0 = net(x).detach().clone()
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    flat = FlatParams(net.parameters())
    assert torch.equal(net(x), y0)
    for p in flat.params:
        assert flat.values.data_ptr() <= p.data_ptr() < flat.values.data_ptr() + flat.values.numel() * 4
        assert p.grad.stride() == p.stride()
    net(x).square().sum().backward()
    assert flat.flat.abs().sum().item() > 0           # autograd accumulated into the flat gradient buffer
    net.load_state_dict({k: v * 2 for k, v in sd.items()})
    assert abs(flat.values.abs().sum().item() - 2 * sum(v.abs().sum().item() for v in sd.values())) < 1e-3


def test_grad_scale_equals_scaling_the_gradient_first():
    """data parallelism: `grad_scale = 1 / world` inside the fused step == `grad.mul_(1 / world)` before it, bit for bit"""
    from adaptsegnet_b200.optim import FlatParams, FusedAdam, FusedSGD
    res = []
    for deferred in (True, False):
        net, _ = _nets()
        flat = FlatParams(net.parameters())
        sgd = FusedSGD(flat, _groups(net), lr=0.01, momentum=0.9, weight_decay=5e-4)
        rng = np.random.default_rng(4)
        for step in range(3):
            flat.flat.copy_(torch.from_numpy(rng.standard_normal(flat.numel).astype(np.float32)))
            if deferred:
                flat.grad_scale = 1.0 / 3.0
            else:
                flat.flat.mul_(1.0 / 3.0)
            sgd.step()
            assert flat.grad_scale == 1.0
        res.append(flat.values.clone())
    assert torch.equal(res[0], res[1])
    res = []
    for deferred in (True, False):
        net, _ = _nets()
        flat = FlatParams(net.parameters())
        adam = FusedAdam(flat, lr=1e-3, betas=(0.9, 0.99))
        rng = np.random.default_rng(5)
        for step in range(3):
            flat.flat.copy_(torch.from_numpy(rng.standard_normal(flat.numel).astype(np.float32)))
            if deferred:
                flat.grad_scale = 0.125
            else:
                flat.flat.mul_(0.125)
            adam.step()
        res.append(flat.values.clone())
    assert torch.equal(res[0], res[1])
