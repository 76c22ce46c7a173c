"""Kernel-for-kernel comparator for bench.py: the hot path of ONE
multi-level iteration at BASELINE config-2 shapes, run (a) through the reference's own modules in PyTorch eager on the
same B200 (ATen / cuDNN, TF32 on and off) and (b) through libasn_b200, group by group, timed with CUDA events.

The reference has no native code of its own: what its modules dispatch to on this box IS the kernel to beat.  Groups
(each starts from tensors the previous part of the iteration produced, exactly the operations the reference executes
between the trunk's forward and backward passes):

  heads         Classifier_Module forward + backward (layer5 on 1024 ch, layer6 on 2048 ch; source 90x160, target 64x128)
                model/deeplab_multi.py:117-121 and its autograd
  seg_loss      nn.Upsample -> CrossEntropyLoss(ignore_index=255), forward + backward, source, both heads
                model/deeplab_multi.py:188-189, train_gta2cityscapes_multi.py:599-605
  adversarial   everything that involves the discriminators, both levels: Upsample(target) -> softmax -> D -> BCE ->
                backward to the low-res logits (G step, D frozen), then D(softmax(pred.detach())) on source and target with
                parameter gradients (D step)  train...:617-628,642-676, model/discriminator.py:21-34
  optimizers    SGD (the reference's duplicated parameter groups, Q11) + 2 x Adam   train...:681-683

This file is BASELINE plumbing: the `ref` side executes oracle/torch_ref (the restated reference modules, allowed for
bench.py's baseline legs only); the `ours` side imports adaptsegnet_b200 only.
"""
from __future__ import annotations

import statistics

import torch
import torch.nn as nn
import torch.nn.functional as F

SRC_HW, TGT_HW = (720, 1280), (512, 1024)
SRC_F, TGT_F = (90, 160), (64, 128)


class _Timer:
    def __init__(self, dev, reps, warm):
        self.dev, self.reps, self.warm = dev, reps, warm
        self.flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def __call__(self, fn):
        for _ in range(self.warm):
            fn()
        ts = []
        for _ in range(self.reps):
            self.flush.fill_(1)                    # evict the previous repetition's data from L2 (not timed)
            torch.cuda.synchronize(self.dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize(self.dev)
            ts.append(e0.elapsed_time(e1))
        return statistics.median(ts)


def _features(dev, gen_seed=1338):
    g = torch.Generator(device="cpu").manual_seed(gen_seed)

    def feat(c, hw, rms, zeros):       # post-ReLU-like statistics of layer3 / layer4
        x = torch.randn((1, c) + hw, generator=g).abs() * rms
        x[torch.rand((1, c) + hw, generator=g) < zeros] = 0
        return x.to(dev)

    return {"f3_src": feat(1024, SRC_F, 4.4, 0.11), "f4_src": feat(2048, SRC_F, 1.6, 0.29),
            "f3_tgt": feat(1024, TGT_F, 4.4, 0.11), "f4_tgt": feat(2048, TGT_F, 1.6, 0.29)}


def _labels(dev):
    g = torch.Generator().manual_seed(1338)
    coarse = torch.randint(0, 19, (1, 1, 45, 80), generator=g).float()
    lab = F.interpolate(coarse, size=SRC_HW, mode="nearest")[0].long()
    lab[:, :36] = 255
    return lab.to(dev)


def measure_reference(dev, tf32, reps=5, warm=2):
    """the reference's modules (restated, oracle/torch_ref) in PyTorch eager on `dev`; -> {group: ms}"""
    from oracle import torch_ref as TR

    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = bool(tf32)
    torch.backends.cuda.matmul.allow_tf32 = bool(tf32)
    try:
        T = _Timer(dev, reps, warm)
        torch.manual_seed(1338)
        l5, l6 = TR.RefClassifier(1024, (6, 12, 18, 24), 19).to(dev), TR.RefClassifier(2048, (6, 12, 18, 24), 19).to(dev)
        for m in (l5, l6):
            for c in m.conv2d_list:
                c.weight.data.normal_(0, 0.01)
        D1, D2 = TR.RefFCDiscriminator(19).to(dev), TR.RefFCDiscriminator(19).to(dev)
        f = {k: v.clone().requires_grad_(True) for k, v in _features(dev).items()}   # NCHW fp32: the reference layout
        lab = _labels(dev)
        ce = nn.CrossEntropyLoss(ignore_index=255)
        bce = nn.BCEWithLogitsLoss()
        up_s = nn.Upsample(size=SRC_HW, mode="bilinear", align_corners=True)
        up_t = nn.Upsample(size=TGT_HW, mode="bilinear", align_corners=True)
        dz_s = torch.randn((1, 19) + SRC_F, device=dev) * 1e-3
        dz_t = torch.randn((1, 19) + TGT_F, device=dev) * 1e-3
        out = {}

        def heads():
            for dom, dz in (("src", dz_s), ("tgt", dz_t)):
                for head, key in ((l5, "f3_"), (l6, "f4_")):
                    x = f[key + dom]
                    y = head(x)
                    torch.autograd.grad(y, [x] + list(head.parameters()), dz)

        out["heads"] = T(heads)
        z_s = [torch.randn((1, 19) + SRC_F, device=dev).mul_(3).requires_grad_(True) for _ in range(2)]
        z_t = [torch.randn((1, 19) + TGT_F, device=dev).mul_(3).requires_grad_(True) for _ in range(2)]

        def seg_loss():
            loss = ce(up_s(z_s[1]), lab) + 0.1 * ce(up_s(z_s[0]), lab)
            torch.autograd.grad(loss, z_s)

        out["seg_loss"] = T(seg_loss)
        with torch.no_grad():
            pred_s = [up_s(z) for z in z_s]     # exists already when the reference reaches the adversarial part

        def target(d, v):                       # train...:621: built on the CPU every call (Q15)
            return torch.FloatTensor(d.data.size()).fill_(v).to(dev)

        def adversarial():
            for D in (D1, D2):
                for p in D.parameters():
                    p.requires_grad = False
            pred_t = [up_t(z) for z in z_t]
            d2 = D2(F.softmax(pred_t[1], dim=1))
            d1 = D1(F.softmax(pred_t[0], dim=1))
            loss = 0.0002 * bce(d1, target(d1, 0)) + 0.001 * bce(d2, target(d2, 0))
            torch.autograd.grad(loss, z_t)
            for D in (D1, D2):
                for p in D.parameters():
                    p.requires_grad = True
            for D, ps, pt in ((D1, pred_s[0], pred_t[0]), (D2, pred_s[1], pred_t[1])):
                d = D(F.softmax(ps.detach(), dim=1))
                torch.autograd.grad(bce(d, target(d, 0)) / 2, list(D.parameters()))
                d = D(F.softmax(pt.detach(), dim=1))
                torch.autograd.grad(bce(d, target(d, 1)) / 2, list(D.parameters()))

        out["adversarial"] = T(adversarial)
        del pred_s
        # optimizers on the full parameter set (random gradients)
        import warnings
        G = TR.RefDeeplabMulti(19).to(dev)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            sgd = torch.optim.SGD(G.optim_parameters(2.5e-4), lr=2.5e-4, momentum=0.9, weight_decay=0.0005)
        adams = [torch.optim.Adam(D.parameters(), lr=1e-4, betas=(0.9, 0.99)) for D in (D1, D2)]
        for m in (G, D1, D2):
            for p in m.parameters():
                if p.requires_grad:
                    p.grad = torch.randn_like(p) * 1e-3

        def optimizers():
            sgd.step()
            for a in adams:
                a.step()

        out["optimizers"] = T(optimizers)
        out["total"] = sum(out.values())
        return out
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def measure_ours(dev, tier="B", reps=5, warm=2):
    """the same groups through libasn_b200 (the configuration bench.py's headline runs: channels_last fp32 features,
    Tier-B consumers, fused softmax, replayed target forward, fused optimizers); -> {group: ms}"""
    from adaptsegnet_b200 import ops
    from adaptsegnet_b200.model.deeplab_multi import Classifier_Module, DeeplabMulti
    from adaptsegnet_b200.model.discriminator import FCDiscriminator
    from adaptsegnet_b200.optim import FlatParams, FusedAdam, FusedSGD
    from adaptsegnet_b200.train_step import TrainConfig

    T = _Timer(dev, reps, warm)

    def graphed(fn):
        """the same group replayed as ONE CUDA graph (how the product runs it: AdaptSegTrainer captures the iteration), i.e.
        without the Python / ctypes launch path between its 20-60 small kernels; None if the group cannot be captured"""
        try:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(2):
                    fn()
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fn()
            return T(g.replay)
        except Exception as e:   # noqa: BLE001 -- an informational extra: never let it take the bench line down
            import sys
            print(f"[bench_groups] graph replay of a group not measured: {type(e).__name__}: {e}", file=sys.stderr)
            torch.cuda.synchronize(dev)
            return None

    torch.manual_seed(1338)
    rates = [6, 12, 18, 24]
    l5, l6 = Classifier_Module(1024, rates, rates, 19).to(dev), Classifier_Module(2048, rates, rates, 19).to(dev)
    D1, D2 = FCDiscriminator(19).to(dev), FCDiscriminator(19).to(dev)
    f = {k: v.contiguous(memory_format=torch.channels_last).requires_grad_(True) for k, v in _features(dev).items()}
    lab = _labels(dev)
    dz_s = torch.randn((1, 19) + SRC_F, device=dev) * 1e-3
    dz_t = torch.randn((1, 19) + TGT_F, device=dev) * 1e-3
    out = {}

    def heads():
        for dom, dz in (("src", dz_s), ("tgt", dz_t)):
            for head, key in ((l5, "f3_"), (l6, "f4_")):
                x = f[key + dom]
                y = head(x)
                torch.autograd.grad(y, [x] + list(head.parameters()), dz)

    out["heads"] = T(heads)
    gout = {"heads": graphed(heads)}
    z_s = [torch.randn((1, 19) + SRC_F, device=dev).mul_(3).requires_grad_(True) for _ in range(2)]
    z_t = [torch.randn((1, 19) + TGT_F, device=dev).mul_(3).requires_grad_(True) for _ in range(2)]
    lazy = tier == "B"

    def seg(z):
        if lazy:
            return ops.upsample_softmax_cross_entropy(z, SRC_HW, lab, ignore_label=255)
        return ops.softmax_cross_entropy(ops.upsample_bilinear(z, SRC_HW), lab, ignore_label=255)

    def seg_loss():
        loss = seg(z_s[1]) + 0.1 * seg(z_s[0])
        torch.autograd.grad(loss, z_s)

    out["seg_loss"] = T(seg_loss)
    gout["seg_loss"] = graphed(seg_loss)
    with torch.no_grad():
        pred_s = None if lazy else [ops.upsample_bilinear(z, SRC_HW) for z in z_s]

    def adversarial():
        for D in (D1, D2):
            for p in D.parameters():
                p.requires_grad = False
        saved = {}
        if lazy:
            d2, saved[1] = D2(z_t[1], from_logits=True, return_saved=True, up_size=TGT_HW)
            d1, saved[0] = D1(z_t[0], from_logits=True, return_saved=True, up_size=TGT_HW)
        else:
            pred_t = [ops.upsample_bilinear(z, TGT_HW) for z in z_t]
            d2, saved[1] = D2(pred_t[1], from_logits=True, return_saved=True)
            d1, saved[0] = D1(pred_t[0], from_logits=True, return_saved=True)
        loss = 0.0002 * ops.gan_loss(d1, 0.0) + 0.001 * ops.gan_loss(d2, 0.0)
        torch.autograd.grad(loss, z_t)
        for D in (D1, D2):
            for p in D.parameters():
                p.requires_grad = True
        for i, D in enumerate((D1, D2)):
            if lazy:
                d = D(z_s[i].detach(), from_logits=True, up_size=SRC_HW)
            else:
                d = D(pred_s[i], from_logits=True)
            torch.autograd.grad(ops.gan_loss(d, 0.0) / 2, list(D.parameters()))
            d = D.replay(saved[i])
            torch.autograd.grad(ops.gan_loss(d, 1.0) / 2, list(D.parameters()))

    out["adversarial"] = T(adversarial)
    gout["adversarial"] = graphed(adversarial)
    G = DeeplabMulti(19).to(dev)
    cfg = TrainConfig()
    flat_g, flat_d = FlatParams(G.parameters()), [FlatParams(D.parameters()) for D in (D1, D2)]
    sgd = FusedSGD(flat_g, G.optim_parameters(cfg), lr=cfg.learning_rate, momentum=0.9, weight_decay=0.0005)
    adams = [FusedAdam(fd, lr=1e-4, betas=(0.9, 0.99)) for fd in flat_d]
    for fp in [flat_g] + flat_d:
        fp.flat.normal_(0, 1e-3)

    def optimizers():
        sgd.step()
        for a in adams:
            a.step()

    out["optimizers"] = T(optimizers)
    gout["optimizers"] = graphed(optimizers)
    out["total"] = sum(out.values())
    gout["total"] = sum(gout.values()) if all(v is not None for v in gout.values()) else None
    out["_graph"] = gout
    return out


def compare(dev, tier="B", reps=5, warm=2):
    """-> the `reference_gpu_eager.hot_path_groups` block of the bench line"""
    ours = measure_ours(dev, tier, reps, warm)
    torch.cuda.empty_cache()
    ref_tf32 = measure_reference(dev, True, reps, warm)
    torch.cuda.empty_cache()
    ref_fp32 = measure_reference(dev, False, reps, warm)
    torch.cuda.empty_cache()
    groups = {}
    ours_graph = ours.pop("_graph", {})
    for k in ours:
        groups[k] = {"ours_ms": round(ours[k], 4),
                     "ours_graph_ms": None if ours_graph.get(k) is None else round(ours_graph[k], 4),
                     "aten_tf32_ms": round(ref_tf32[k], 4),
                     "aten_fp32_ms": round(ref_fp32[k], 4),
                     "speedup_vs_tf32": round(ref_tf32[k] / ours[k], 2), "speedup_vs_fp32": round(ref_fp32[k] / ours[k], 2)}
    return groups
