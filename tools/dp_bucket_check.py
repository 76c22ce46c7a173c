"""torchrun --nproc-per-node 2 tools/dp_bucket_check.py
Data-parallel check on real GPUs: the bucketed, event-gated all-reduce of the generator's gradient (three slices started
from inside the captured iteration's timeline) gives the SAME averaged gradients and post-step weights as one all-reduce
after the iteration -- bit for bit at 2 ranks (a two-term sum does not depend on the reduction order) -- over several
iterations with changing inputs; and both ranks end with identical weights."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from adaptsegnet_b200.train_step import AdaptSegTrainer, TrainConfig
from adaptsegnet_b200.utils.synthetic import synthetic_batch

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
res = {}
trainers = {}
for mode in ("bucketed", "single"):
    torch.manual_seed(1338)
    tr = AdaptSegTrainer(TrainConfig(lazy_upsample=True), device=dev, use_cuda_graph=True, channels_last=True, overlap=True)
    tr.bucketed_allreduce = mode == "bucketed"      # (set before the first step: the default is the single all-reduce)
    trainers[mode] = tr
hw_s, hw_t = (264, 520), (200, 392)
ok = True
def sync(dst, src):
    """same weights and optimizer state in both trainers: every iteration is an independent comparison (the classifier's
    weight gradient uses fp32 atomics, so two runs of the discriminators differ in the last bits and would drift apart)"""
    with torch.no_grad():
        dst.flat_G.values.copy_(src.flat_G.values)
        dst.flat_D.values.copy_(src.flat_D.values)
        dst.optimizer.momentum_buffer.copy_(src.optimizer.momentum_buffer)
        dst.optimizer_D.exp_avg.copy_(src.optimizer_D.exp_avg)
        dst.optimizer_D.exp_avg_sq.copy_(src.optimizer_D.exp_avg_sq)
        for a, b in zip(dst.model.buffers(), src.model.buffers()):
            a.copy_(b)
        dst.flat_G.mark_updated()
        dst.flat_D.mark_updated()


for it in range(4):
    src, lab, tgt = (t.to(dev) for t in synthetic_batch(100 + 10 * it + rank, hw_s, hw_t))
    sync(trainers["bucketed"], trainers["single"])
    grads = {}
    for mode, tr in trainers.items():
        tr.step(src, lab, tgt, i_iter=it, do_optimizer_step=False)   # leaves the AVERAGED gradients in the flat buffers
        torch.cuda.synchronize()
        grads[mode] = (tr.flat_G.flat.clone(), tr.flat_D.flat.clone())
    same_g = torch.equal(grads["bucketed"][0], grads["single"][0])
    dd = (grads["bucketed"][1] - grads["single"][1]).norm() / grads["single"][1].norm()
    same_d = bool(dd < 1e-5)      # (atomics in the classifier's weight gradient: not bit-reproducible between two runs)
    nz = float(grads["single"][0].abs().sum())
    # both ranks hold the same averaged gradient
    g0 = grads["bucketed"][0].clone()
    dist.broadcast(g0, src=0)
    same_ranks = torch.equal(g0, grads["bucketed"][0])
    ok = ok and same_g and same_d and same_ranks and nz > 0
    if rank == 0:
        print(json.dumps({"iter": it, "generator_grads_equal": same_g, "discriminator_grads_equal": same_d,
                          "ranks_equal": same_ranks, "abs_sum": nz, "used_buckets": trainers["bucketed"]._buckets is not None}))
    for tr in trainers.values():          # now really step, so that the next iteration sees new weights
        tr.optimizer.step()
        tr.optimizer_D.step()
t = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("DP_BUCKET_CHECK", "PASS" if int(t.item()) == 1 else "FAIL")
dist.destroy_process_group()
sys.exit(0 if int(t.item()) == 1 else 1)
