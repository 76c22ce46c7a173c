"""Runs every hot-path op once (after one warm-up) at BASELINE config-2 shapes on synthetic features --
a short command for ncu (--set full) and for back-to-back microbenchmarks of single kernels.
  python tools/hot_path_once.py            # one pass (ncu target)
  python tools/hot_path_once.py --bench    # per-op timing, 20 back-to-back launches, L2 flushed between ops
"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adaptsegnet_b200 import ops
from adaptsegnet_b200.model.deeplab_multi import Classifier_Module
from adaptsegnet_b200.model.discriminator import FCDiscriminator

dev = "cuda"
torch.manual_seed(1338)
bench = "--bench" in sys.argv
H, W, h, w = 720, 1280, 90, 160
f4 = (torch.randn(1, 2048, h, w, device=dev).abs() * 1.6).requires_grad_(True)
head = Classifier_Module(2048, [6, 12, 18, 24], [6, 12, 18, 24], 19).to(dev)
D = FCDiscriminator(19).to(dev)
lab = torch.randint(0, 19, (1, H, W), device=dev)
lab[:, :40] = 255
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def one_pass():
    logits = head(f4)
    up = ops.upsample_bilinear(logits, (H, W))
    loss = ops.softmax_cross_entropy(up, lab)
    loss.backward()
    up2 = up.detach().requires_grad_(True)
    d = D(up2, from_logits=True)
    l2 = ops.gan_loss(d, 0.0, ops.GAN_BCE)
    l2.backward()
    p = ops.softmax_channels(up.detach())
    hist, _ = ops.fast_hist(lab.reshape(-1), torch.zeros(H * W, dtype=torch.uint8, device=dev), 19)
    pred = ops.upsample_argmax(logits.detach(), (H, W))
    return loss


one_pass()
torch.cuda.synchronize()
if not bench:
    one_pass()
    torch.cuda.synchronize()
    print("ok")
    sys.exit(0)

# ---- back-to-back microbenchmarks ----
from adaptsegnet_b200 import prof
res = {}
logits = head(f4).detach()
up = ops.upsample_fwd_raw(logits, H, W)
dz = torch.randn_like(up)
cases = {
    "upsample_fwd": lambda: ops.upsample_fwd_raw(logits, H, W),
    "upsample_bwd": lambda: ops.upsample_bwd_raw(dz, h, w),
    "softmax_fwd": lambda: ops.softmax_channels(up),
    "upsample_argmax": lambda: ops.upsample_argmax(logits, (H, W)),
}
for name, fn in cases.items():
    fn()
    flush.zero_()
    torch.cuda.synchronize()
    prof.enable(True)
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    rep = prof.report()
    prof.enable(False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record()
    torch.cuda.synchronize()
    res[name] = {"batch_ms_per_call": e0.elapsed_time(e1) / 20,
                 "kernels": {k: {"us": v["ms"] / v["launches"] * 1e3,
                                 "gbs": v["bytes"] / max(v["ms"], 1e-9) / 1e6} for k, v in rep.items()}}
print(json.dumps(res, indent=1))
