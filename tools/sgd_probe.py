"""sgd_step / adam_step at their real sizes (the generator's 44.5 M parameters with the reference's duplicated groups;
both discriminators in one buffer): per-launch time from the library's own CUDA events, 30 launches each."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adaptsegnet_b200 import prof
from adaptsegnet_b200.model.deeplab_multi import DeeplabMulti
from adaptsegnet_b200.model.discriminator import FCDiscriminator
from adaptsegnet_b200.optim import FlatParams, FusedAdam, FusedSGD
from adaptsegnet_b200.train_step import TrainConfig

dev = "cuda"
G = DeeplabMulti(19).to(dev)
flatG = FlatParams(G.parameters())
sgd = FusedSGD(flatG, G.optim_parameters(TrainConfig()), lr=2.5e-4, momentum=0.9, weight_decay=5e-4)
flatG.flat.normal_(0, 1e-3)
D1, D2 = FCDiscriminator(19).to(dev), FCDiscriminator(19).to(dev)
flatD = FlatParams(list(D1.parameters()) + list(D2.parameters()))
adam = FusedAdam(flatD, lr=1e-4, betas=(0.9, 0.99))
flatD.flat.normal_(0, 1e-3)
for _ in range(3):
    sgd.step(); adam.step()
torch.cuda.synchronize()
prof.enable(True)
for _ in range(30):
    sgd.step(); adam.step()
torch.cuda.synchronize()
rep = prof.report()
prof.enable(False)
out = {k: {"us": round(v["ms"] / v["launches"] * 1e3, 2), "GBps": round(v["bytes"] / v["launches"] / (v["ms"] / v["launches"] * 1e-3) / 1e9, 1)}
       for k, v in rep.items()}
print(json.dumps({"env": {e: os.environ[e] for e in os.environ if e.startswith("ASN_")}, **out}))
