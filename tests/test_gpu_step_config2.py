"""-m gpu: the whole hot path of one multi-level iteration at BASELINE config-2 shapes (720x1280 source, 512x1024
target) in the BENCHMARKED configuration -- channels_last trunk, Tier-B (lazy upsample: fused upsample+softmax+CE and
upsample+softmax inside the discriminators' input pack), fused softmax, the D step's target forward replayed from the G
step's activations, the two-stream schedule (source pipeline, target pipeline and discriminator step overlapped), everything
captured as ONE CUDA graph -- gated per backward pass at the north star's 1e-2.

Why per backward pass: a gradient compared across two DIFFERENT forwards (bf16 tensor cores here, fp32 on the host
there) measures the discontinuities of the network (LeakyReLU sign flips, softmax amplification of the logit error) and
not the kernels.  Autograd itself defines a backward pass on the activations the forward saved, so every segment is
checked against the reference's arithmetic (torch CPU fp32: F.conv2d / F.interpolate / softmax / cross_entropy and
their autograd -- the calls of model/deeplab_multi.py:117-121,188-189, model/discriminator.py:21-34,
train_gta2cityscapes_multi.py:599-676) evaluated AT THE SAME linearisation point: the features, logits and activations
this run produced.  Gates for every head and discriminator gradient: L2-relative error <= 1e-2 AND cosine >= 0.9999.

  A  source:  heads forward (logits vs fp32 heads on the same features), seg loss + its gradient at the logits,
              head backward (dW, db of both heads, dX) for that gradient
  B  target / generator step: D(softmax(interp(z))) forward for both levels, adversarial loss, gradient at the low-res
              logits (through the LeakyReLU masks of the saved activations), head backward
  C  discriminator step: parameter gradients of the source pass and of the (replayed) target pass, separately
  D  the sum of the separately checked pieces == what AdaptSegTrainer.step() leaves in its flat gradient buffers when it
     runs the same iteration as the captured two-stream CUDA graph (ties A-C to the configuration bench.py times)
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import torch_ref as TR
from gpu_util import gpu

pytestmark = gpu
SEED = 1338
SRC_HW, TGT_HW = (720, 1280), (512, 1024)
L2_TOL, COS_MIN = 1e-2, 0.9999


def _gate(name, got, ref, report):
    got = got.detach().double().cpu().reshape(-1)
    ref = ref.detach().double().cpu().reshape(-1)
    l2 = float((got - ref).norm() / ref.norm().clamp_min(1e-300))
    cos = float(torch.dot(got, ref) / (got.norm() * ref.norm()).clamp_min(1e-300))
    report[name] = {"l2_rel": l2, "cos": cos}
    assert l2 <= L2_TOL and cos >= COS_MIN, (name, l2, cos)


def _ref_heads(trainer):
    """fp32 CPU copies of the two heads (the reference's Classifier_Module arithmetic)"""
    heads = []
    for mine, cin in ((trainer.model.layer5, 1024), (trainer.model.layer6, 2048)):
        h = TR.RefClassifier(cin, (6, 12, 18, 24), 19)
        h.load_state_dict({k: v.detach().cpu() for k, v in mine.state_dict().items()})
        heads.append(h)
    return heads


def _ref_d_backward(D, a0, acts, dout):
    """model/discriminator.py:21-34 backward ON THE SAVED ACTIVATIONS (what autograd does for the reference: the in-place
    LeakyReLU keeps only its output, whose sign is the mask).  a0: the input the kernels saw; acts: A1..A4; -> (dA0, grads)"""
    convs = [D.conv1, D.conv2, D.conv3, D.conv4, D.classifier]
    inputs = [a0] + list(acts)
    g = dout
    grads = {}
    for l in range(4, -1, -1):
        w = convs[l].weight.detach().cpu()
        if l < 4:
            g = g * torch.where(acts[l] > 0, 1.0, 0.2)
        grads[l] = (torch.nn.grad.conv2d_weight(inputs[l], w.shape, g, stride=2, padding=1), g.sum((0, 2, 3)))
        g = torch.nn.grad.conv2d_input(inputs[l].shape, w, g, stride=2, padding=1)
    return g, grads


def _bf16(t):
    return t.to(torch.bfloat16).float()


def test_whole_hot_path_config2_per_backward_gates():
    import json
    import os
    from adaptsegnet_b200 import ops
    from adaptsegnet_b200.train_step import AdaptSegTrainer, TrainConfig

    torch.manual_seed(SEED)
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = TrainConfig(level="multi-level", gan="Vanilla", lazy_upsample=True)
    tr = AdaptSegTrainer(cfg, device="cuda", use_cuda_graph=True, channels_last=True, overlap=True)   # as bench.py runs it
    src, lab, tgt = (t.cuda() for t in TR.synthetic_batch(SEED, SRC_HW, TGT_HW))
    report = {}

    # ---------------- D: the benchmarked configuration itself (two captured graphs), gradients left in the flat buffers
    out_graph = tr.step(src, lab, tgt, i_iter=0, do_optimizer_step=False)
    out_graph = {k: float(v.item()) for k, v in out_graph.items()}
    flat_graph = {n: getattr(tr, n).flat.clone() for n in ("flat_G", "flat_D1", "flat_D2")}
    head_params = [p for h in (tr.model.layer5, tr.model.layer6) for p in h.parameters()]
    head_graph = [p.grad.clone() for p in head_params]

    # ---------------- the same iteration piece by piece (eager, same kernels), every piece checked against fp32
    model, D1, D2 = tr.model, tr.model_D1, tr.model_D2
    heads_ref = _ref_heads(tr)
    with torch.no_grad():
        f_src = [f.detach() for f in model.trunk(src.contiguous(memory_format=torch.channels_last))]
        f_tgt = [f.detach() for f in model.trunk(tgt.contiguous(memory_format=torch.channels_last))]
    for p in list(D1.parameters()) + list(D2.parameters()):
        p.requires_grad = False
    pieces_head = [torch.zeros_like(g) for g in head_graph]

    def head_grads():
        gs = [p.grad.clone() for p in head_params]
        for p in head_params:
            p.grad.zero_()
        return gs

    for p in head_params:
        p.grad.zero_()

    # ---- A: source ----
    xs = [f.clone().requires_grad_(True) for f in f_src]
    z = [model.layer5(xs[0]), model.layer6(xs[1])]
    for t in z:
        t.retain_grad()
    l1 = ops.upsample_softmax_cross_entropy(z[0], SRC_HW, lab, ignore_label=255)
    l2 = ops.upsample_softmax_cross_entropy(z[1], SRC_HW, lab, ignore_label=255)
    (l2 + cfg.lambda_seg * l1).backward()
    g_seg = head_grads()
    lab_c = lab.cpu()
    for i, (name, lam) in enumerate((("layer5", cfg.lambda_seg), ("layer6", 1.0))):
        x_c = xs[i].detach().cpu().contiguous().requires_grad_(True)     # fp32 features of THIS run, NCHW
        z_ref = heads_ref[i](x_c)
        _gate(f"A.{name}.logits", z[i], z_ref, report)
        # seg loss and its gradient at OUR logits (model/deeplab_multi.py:188-189 + train...:599-600)
        z_leaf = z[i].detach().cpu().requires_grad_(True)
        loss_ref = F.cross_entropy(F.interpolate(z_leaf, size=SRC_HW, mode="bilinear", align_corners=True), lab_c,
                                   ignore_index=255)
        (lam * loss_ref).backward()
        got_loss = float((l1 if i == 0 else l2).item())
        assert abs(got_loss - float(loss_ref)) <= 1e-5 * abs(float(loss_ref)), (name, got_loss, float(loss_ref))
        _gate(f"A.{name}.dlogits", z[i].grad, z_leaf.grad, report)
        # head backward for that gradient
        params = list(heads_ref[i].parameters())
        gr = torch.autograd.grad(z_ref, [x_c] + params, z[i].grad.cpu())
        _gate(f"A.{name}.dX", xs[i].grad, gr[0], report)
        mine_g = g_seg[i * 8:(i + 1) * 8]
        for (pn, _), gm, gref in zip(heads_ref[i].named_parameters(), mine_g, gr[1:]):
            _gate(f"A.{name}.{pn}", gm, gref, report)
    for a, b in zip(pieces_head, g_seg):
        a += b
    report["A.loss_seg"] = {"ours": [float(l1), float(l2)], "graph": [out_graph["loss_seg1"], out_graph["loss_seg2"]]}

    # ---- B: target, generator step (discriminators frozen) ----
    xt = [f.clone().requires_grad_(True) for f in f_tgt]
    zt = [model.layer5(xt[0]), model.layer6(xt[1])]
    for t in zt:
        t.retain_grad()
    saved, d_out, losses_adv = {}, {}, {}
    total = 0
    for i, (D, lam) in enumerate(((D1, cfg.lambda_adv_target1), (D2, cfg.lambda_adv_target2))):
        d_out[i], saved[i] = D(zt[i], from_logits=True, return_saved=True, up_size=TGT_HW)
        d_out[i].retain_grad()
        losses_adv[i] = tr.bce_loss(d_out[i], 0)
        total = total + lam * losses_adv[i]
    total.backward()
    g_adv = head_grads()
    acts_tgt, a0_tgt = {}, {}
    for i, (D, name) in enumerate(((D1, "D1"), (D2, "D2"))):
        Dc_w = D  # parameters read on the CPU inside _ref_d_backward
        z_leaf = zt[i].detach().cpu().requires_grad_(True)
        p_full = torch.softmax(F.interpolate(z_leaf, size=TGT_HW, mode="bilinear", align_corners=True), dim=1)
        Dref = TR.RefFCDiscriminator(19)
        Dref.load_state_dict({k: v.detach().cpu() for k, v in D.state_dict().items()})
        with torch.no_grad():
            out_ref = Dref(p_full.detach())
        _gate(f"B.{name}.out", d_out[i], out_ref, report)
        loss_ref = F.binary_cross_entropy_with_logits(out_ref, torch.zeros_like(out_ref))
        assert abs(float(losses_adv[i]) - float(loss_ref)) <= 1e-2 * abs(float(loss_ref))
        acts_tgt[i] = [a.cpu() for a in ops.fcd_decode_activations(saved[i].acts, saved[i].cfg)]
        a0_tgt[i] = _bf16(p_full.detach())
        dA0_ref, _ = _ref_d_backward(Dref, a0_tgt[i], acts_tgt[i], d_out[i].grad.cpu())
        (p_full * dA0_ref).sum().backward()      # softmax + interp adjoint (train...:617-618 / deeplab_multi.py:188-189)
        _gate(f"B.{name}.dlogits", zt[i].grad, z_leaf.grad, report)
        hname = ("layer5", "layer6")[i]
        x_c = xt[i].detach().cpu().contiguous().requires_grad_(True)
        z_ref = heads_ref[i](x_c)
        gr = torch.autograd.grad(z_ref, [x_c] + list(heads_ref[i].parameters()), zt[i].grad.cpu())
        _gate(f"B.{hname}.dX", xt[i].grad, gr[0], report)
        for (pn, _), gm, gref in zip(heads_ref[i].named_parameters(), g_adv[i * 8:(i + 1) * 8], gr[1:]):
            _gate(f"B.{hname}.{pn}", gm, gref, report)
    for a, b in zip(pieces_head, g_adv):
        a += b

    # ---- C: discriminator step, source pass and replayed target pass separately ----
    for p in list(D1.parameters()) + list(D2.parameters()):
        p.requires_grad = True
    tr.flat_D1.zero()
    tr.flat_D2.zero()
    pieces_D = {}
    for i, (D, name, flat) in enumerate(((D1, "D1", tr.flat_D1), (D2, "D2", tr.flat_D2))):
        Dref = TR.RefFCDiscriminator(19)
        Dref.load_state_dict({k: v.detach().cpu() for k, v in D.state_dict().items()})
        convs = ("conv1", "conv2", "conv3", "conv4", "classifier")
        # source
        d = D(z[i].detach(), from_logits=True, up_size=SRC_HW)
        d.retain_grad()
        acts_s = [a.cpu() for a in ops.fcd_saved_activations(d)]
        (tr.bce_loss(d, 0) / 2).backward()
        g_src = flat.flat.clone()
        got = {n: (getattr(D, n).weight.grad.clone(), getattr(D, n).bias.grad.clone()) for n in convs}
        p_src = torch.softmax(F.interpolate(z[i].detach().cpu(), size=SRC_HW, mode="bilinear", align_corners=True), dim=1)
        _, gref = _ref_d_backward(Dref, _bf16(p_src), acts_s, d.grad.cpu())
        for l, n in enumerate(convs):
            _gate(f"C.{name}.src.{n}.weight", got[n][0], gref[l][0], report)
            _gate(f"C.{name}.src.{n}.bias", got[n][1], gref[l][1], report)
        flat.zero()
        # target: replay of the generator step's forward (train...:665-666 repeats :617-618 with unchanged weights)
        d = D.replay(saved[i])
        d.retain_grad()
        (tr.bce_loss(d, 1) / 2).backward()
        g_tgt = flat.flat.clone()
        got = {n: (getattr(D, n).weight.grad.clone(), getattr(D, n).bias.grad.clone()) for n in convs}
        _, gref = _ref_d_backward(Dref, a0_tgt[i], acts_tgt[i], d.grad.cpu())
        for l, n in enumerate(convs):
            _gate(f"C.{name}.tgt.{n}.weight", got[n][0], gref[l][0], report)
            _gate(f"C.{name}.tgt.{n}.bias", got[n][1], gref[l][1], report)
        flat.zero()
        pieces_D[name] = g_src + g_tgt

    # ---- D: pieces == the graph-captured step ----
    for g_sum, g_graph, p in zip(pieces_head, head_graph, head_params):
        rel = float((g_sum - g_graph).norm() / g_graph.norm())
        assert rel <= 1e-5, ("head gradient, pieces vs captured step", tuple(p.shape), rel)
    for name, key in (("D1", "flat_D1"), ("D2", "flat_D2")):
        rel = float((pieces_D[name] - flat_graph[key]).norm() / flat_graph[key].norm())
        report[f"D.{name}.pieces_vs_graph"] = rel
        assert rel <= 1e-5, (name, rel)
    assert abs(float(losses_adv[0]) - out_graph["loss_adv_target1"]) <= 1e-5 and \
        abs(float(losses_adv[1]) - out_graph["loss_adv_target2"]) <= 1e-5
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/step_config2_gates.json", "w") as f:
        json.dump(report, f, indent=1)
    worst_l2 = max(v["l2_rel"] for v in report.values() if isinstance(v, dict) and "l2_rel" in v)
    worst_cos = min(v["cos"] for v in report.values() if isinstance(v, dict) and "cos" in v)
    print(f"config-2 per-backward gates: worst L2-rel {worst_l2:.2e}, worst cosine {worst_cos:.6f}")
