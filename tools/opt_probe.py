"""How long do the three optimizer steps (and zeroing the flat gradient buffers) take per iteration?"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adaptsegnet_b200.train_step import AdaptSegTrainer, TrainConfig

tr = AdaptSegTrainer(TrainConfig(lazy_upsample=True), device="cuda", channels_last=True)
src = torch.randn(1, 3, 256, 512, device="cuda") * 50
tgt = torch.randn(1, 3, 256, 512, device="cuda") * 50
lab = torch.randint(0, 19, (1, 256, 512), device="cuda")
for i in range(2):
    tr.step(src, lab, tgt, i_iter=i)
torch.cuda.synchronize()
res = {}
def timeit(name, fn, n=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    res[name] = round(e0.elapsed_time(e1) / n, 4)
timeit("sgd_G_step_ms", tr.optimizer.step)
timeit("adam_D_step_ms", tr.optimizer_D.step)
timeit("adam_D2_step_ms", tr.optimizer_D2.step)
timeit("zero_flat_ms", lambda: (tr.flat_G.zero(), tr.flat_D1.zero(), tr.flat_D2.zero()))
n_list = sum(len(g["params"]) for g in tr.optimizer.param_groups)
n_uniq = len({id(p) for g in tr.optimizer.param_groups for p in g["params"]})
res["sgd_param_entries"] = n_list
res["sgd_unique_params"] = n_uniq
res["sgd_entry_elems"] = sum(p.numel() for g in tr.optimizer.param_groups for p in g["params"])
res["sgd_unique_elems"] = sum(p.numel() for p in {id(p): p for g in tr.optimizer.param_groups for p in g["params"]}.values())
print(json.dumps(res, indent=1))
