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


head5 = Classifier_Module(1024, [6, 12, 18, 24], [6, 12, 18, 24], 19).to(dev)
f3 = (torch.randn(1, 1024, h, w, device=dev).abs() * 4.4).requires_grad_(True)
f4cl = f4.detach().contiguous(memory_format=torch.channels_last).requires_grad_(True)
from adaptsegnet_b200.optim import FlatParams, FusedAdam, FusedSGD
from adaptsegnet_b200.model.deeplab_multi import DeeplabMulti
from adaptsegnet_b200.train_step import TrainConfig
D_b = FCDiscriminator(19).to(dev)
flatD = FlatParams(list(D.parameters()) + list(D_b.parameters()))       # both discriminators: one buffer, one Adam launch
adam = FusedAdam(flatD, lr=1e-4, betas=(0.9, 0.99))
# the generator's optimizer at its real size (44.5 M parameters, the reference's duplicated groups): what bench.py times
G_full = DeeplabMulti(19).to(dev)
flatG = FlatParams(G_full.parameters())
sgd = FusedSGD(flatG, G_full.optim_parameters(TrainConfig()), lr=2.5e-4, momentum=0.9, weight_decay=5e-4)
flatG.flat.normal_(0, 1e-3)


tier_b_only = "--tier-b" in sys.argv
z_eval = torch.randn(1, 19, 64, 128, device=dev) * 3
lab_eval = torch.randint(0, 19, (1, 1024, 2048), device=dev, dtype=torch.uint8)
hist_eval = torch.zeros((19, 19), dtype=torch.int64, device=dev)
ovf_eval = torch.zeros(1, dtype=torch.int64, device=dev)


def one_pass():
    loss = None
    head._pack.invalidate()      # weight packing is part of every pass (in training the optimizer step invalidates it)
    head5._pack.invalidate()
    if not tier_b_only:
        # Tier-A chain (full-resolution logits materialised) on the layer4 head
        logits = head(f4)
        up = ops.upsample_bilinear(logits, (H, W))
        loss = ops.softmax_cross_entropy(up, lab)
        loss.backward()
        up2 = up.detach().requires_grad_(True)
        d = D(up2, from_logits=True)
        l2 = ops.gan_loss(d, 0.0, ops.GAN_BCE)
        l2.backward()
        p = ops.softmax_channels(up.detach())
    # Tier-B chain (what the trainer runs): channels_last features, low-res logits into the fused consumers
    z5 = head5(f3)
    z6 = head(f4cl)
    lb = ops.upsample_softmax_cross_entropy(z6, (H, W), lab) + 0.1 * ops.upsample_softmax_cross_entropy(z5, (H, W), lab)
    lb.backward()
    zd = z6.detach().requires_grad_(True)
    dd = D(zd, from_logits=True, up_size=(H, W))
    ops.gan_loss(dd, 1.0, ops.GAN_BCE).backward()
    sgd.step()
    adam.step()
    # evaluation kernels
    hist, _ = ops.fast_hist(lab.reshape(-1), torch.zeros(H * W, dtype=torch.uint8, device=dev), 19)
    pred = ops.upsample_argmax(z6.detach(), (H, W))
    # config-5 tail: both bilinear stages + argmax + label LUT + confusion matrix in one kernel, device mIoU
    ops.upsample2_argmax_hist(z_eval, (512, 1024), (1024, 2048), lab_eval, 19, hist_eval, ovf_eval)
    ops.per_class_iu_device(hist_eval)
    return lb if loss is None else loss


one_pass()
torch.cuda.synchronize()
if not bench:
    if "--sequence" in sys.argv:   # the library's own names of the second pass's kernels, in launch order (one per scope)
        from adaptsegnet_b200 import prof as _prof
        _prof.enable(True)
        one_pass()
        torch.cuda.synchronize()
        seq = _prof.sequence()
        _prof.enable(False)
        with open(sys.argv[sys.argv.index("--sequence") + 1], "w") as f:
            json.dump(seq, f)
    else:
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
