"""Per-kernel timing of the hot path WITHOUT the trunk (seconds, not minutes): both ASPP heads forward + backward on
channels_last features, upsample+CE, one discriminator forward + backward with parameter gradients, at the source
(720x1280) and target (512x1024) shapes of BASELINE config 2; L2 flushed between passes.  For A/B of environment
switches (e.g. ASN_GLUE=0 selects the round-1 layout kernels):
  python tools/glue_bench.py [--reps 10] [--out file.json]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from adaptsegnet_b200 import ops, prof
from adaptsegnet_b200.model.deeplab_multi import Classifier_Module
from adaptsegnet_b200.model.discriminator import FCDiscriminator

dev = "cuda"
torch.manual_seed(1338)
reps = int(sys.argv[sys.argv.index("--reps") + 1]) if "--reps" in sys.argv else 10
SHAPES = {"src": ((720, 1280), (90, 160)), "tgt": ((512, 1024), (64, 128))}
head6 = Classifier_Module(2048, [6, 12, 18, 24], [6, 12, 18, 24], 19).to(dev)
head5 = Classifier_Module(1024, [6, 12, 18, 24], [6, 12, 18, 24], 19).to(dev)
D = FCDiscriminator(19).to(dev)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
feats = {}
for k, (HW, hw) in SHAPES.items():
    f4 = (torch.randn(1, 2048, *hw, device=dev).abs() * 1.6).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    f3 = (torch.randn(1, 1024, *hw, device=dev).abs() * 4.4).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    lab = torch.randint(0, 19, (1,) + HW, device=dev)
    lab[:, :40] = 255
    feats[k] = (f3, f4, lab)


def one_pass(k):
    HW, _ = SHAPES[k]
    f3, f4, lab = feats[k]
    head6._pack.invalidate()
    head5._pack.invalidate()
    D._pack.invalidate() if hasattr(D, "_pack") else None
    z5, z6 = head5(f3), head6(f4)
    if k == "src":
        loss = ops.upsample_softmax_cross_entropy(z6, HW, lab) + 0.1 * ops.upsample_softmax_cross_entropy(z5, HW, lab)
        loss.backward()
        zd = z6.detach()                                   # D step on the source prediction: parameter gradients only
        ops.gan_loss(D(zd, from_logits=True, up_size=HW), 0.0, ops.GAN_BCE).backward()
    else:
        d = D(z6, from_logits=True, up_size=HW)            # adversarial pass: gradient to the low-res logits and the head
        (ops.gan_loss(d, 0.0, ops.GAN_BCE) + 0.0 * z5.sum()).backward()


res = {}
for k in SHAPES:
    one_pass(k)
    torch.cuda.synchronize()
    prof.enable(True)
    for _ in range(reps):
        flush.zero_()
        one_pass(k)
    torch.cuda.synchronize()
    rep = prof.report()
    prof.enable(False)
    res[k] = {n: {"us": round(v["ms"] / v["launches"] * 1e3, 2), "per_pass": v["launches"] // reps} for n, v in rep.items()}
tot = {k: round(sum(v["us"] * v["per_pass"] for v in r.values()), 1) for k, r in res.items()}
out = {"env": {e: os.environ[e] for e in os.environ if e.startswith("ASN_")}, "total_us": tot, "kernels": res}
if "--out" in sys.argv:
    json.dump(out, open(sys.argv[sys.argv.index("--out") + 1], "w"), indent=1)
names = sorted(set(res["src"]) | set(res["tgt"]))
print("total us per pass:", tot, out["env"])
for n in names:
    a, b = res["src"].get(n), res["tgt"].get(n)
    print(f"{n:26s} src {a['us'] if a else '-':>8} x{a['per_pass'] if a else 0}   tgt {b['us'] if b else '-':>8} x{b['per_pass'] if b else 0}")
