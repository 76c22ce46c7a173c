"""Data-parallel host logic on CPU (gloo, world_size 2): one all-reduce per optimizer over a flat
gradient buffer == mean of the per-rank gradients; the confusion matrix all-reduce == serial sum
(SURVEY.md section 8e).  No kernels involved -- the N > 1 path's plumbing only."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from adaptsegnet_b200.train_step import FlatGrads

    torch.manual_seed(0)  # identical replicas
    net = nn.Sequential(nn.Conv2d(3, 4, 3), nn.Conv2d(4, 2, 1))
    frozen = net[0].bias
    frozen.requires_grad = False
    flat = FlatGrads(list(net.parameters()) + list(net.parameters()))  # duplicates are tolerated (Q11)
    assert flat.flat.numel() == sum(p.numel() for p in net.parameters() if p.requires_grad)
    flat.zero()
    torch.manual_seed(100 + rank)  # each rank its own batch
    x = torch.randn(2, 3, 8, 8)
    net(x).square().mean().backward()
    net(x * 0.5).sum().backward()  # a second backward accumulates into the same views
    local = [p.grad.clone() for p in net.parameters() if p.requires_grad]
    assert all(p.grad.data_ptr() >= flat.flat.data_ptr() for p in net.parameters() if p.requires_grad)
    flat.all_reduce_mean()
    torch.save({"local": local, "avg": [p.grad.clone() for p in net.parameters() if p.requires_grad]},
               os.path.join(out_dir, f"r{rank}.pt"))
    # the fused-optimizer storage (parameters AND gradients flat) with the split start / finish all-reduce the trainer
    # uses to overlap the generator's reduction with the discriminator part
    from adaptsegnet_b200.optim import FlatParams
    torch.manual_seed(0)
    net2 = nn.Sequential(nn.Conv2d(3, 4, 3), nn.Conv2d(4, 2, 1))
    fp = FlatParams(net2.parameters())
    torch.manual_seed(200 + rank)
    net2(torch.randn(2, 3, 8, 8)).square().mean().backward()
    local2 = fp.flat.clone()
    pending = fp.all_reduce_start()
    assert pending is not None
    fp.all_reduce_finish(pending)
    # the division by the world size is deferred to the fused optimizer kernel (asn_sgd_step / asn_adam_step grad_scale)
    assert fp.grad_scale == 1.0 / world
    torch.save({"local": local2, "avg": fp.flat.clone() * fp.grad_scale}, os.path.join(out_dir, f"f{rank}.pt"))
    fp.grad_scale = 1.0
    fp.flat.copy_(local2)
    fp.all_reduce_mean()          # the synchronous flavour scales in place
    assert fp.grad_scale == 1.0 and torch.allclose(fp.flat, torch.load(os.path.join(out_dir, f"f{rank}.pt"))["avg"])
    # eval: per-rank confusion matrices summed exactly
    rng = np.random.RandomState(rank)
    hist = torch.from_numpy(rng.randint(0, 1000, (19, 19)).astype(np.int64))
    total = hist.clone()
    dist.all_reduce(total, op=dist.ReduceOp.SUM)
    torch.save({"hist": hist, "total": total}, os.path.join(out_dir, f"h{rank}.pt"))
    dist.destroy_process_group()


def test_flat_grad_allreduce_and_hist_sum(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r = [torch.load(tmp_path / f"r{i}.pt") for i in range(world)]
    for k in range(len(r[0]["local"])):
        mean = (r[0]["local"][k] + r[1]["local"][k]) / 2
        assert torch.allclose(r[0]["avg"][k], mean, atol=1e-7)
        assert torch.equal(r[0]["avg"][k], r[1]["avg"][k])
    f = [torch.load(tmp_path / f"f{i}.pt") for i in range(world)]
    assert torch.allclose(f[0]["avg"], (f[0]["local"] + f[1]["local"]) / 2, atol=1e-7)
    assert torch.equal(f[0]["avg"], f[1]["avg"])
    h = [torch.load(tmp_path / f"h{i}.pt") for i in range(world)]
    assert torch.equal(h[0]["total"], h[0]["hist"] + h[1]["hist"]) and torch.equal(h[0]["total"], h[1]["total"])
