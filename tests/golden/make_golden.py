#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference modules.

Run in the build container only (needs /root/reference, read-only):

    python tests/golden/make_golden.py

The reference ships no golden vectors of its own, so these
fixtures -- outputs of the reference's own nn.Modules / numpy functions on seeded
inputs, torch CPU fp32 -- are what pins the oracle (oracle/np_oracle.py) and,
through it, the CUDA path.  /root/reference does not exist on the GPU box; the
.npz files travel instead.  Fixtures are kept small (tiny channel counts where
the reference's constructors allow it).
"""
import os
import sys
import warnings

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

REF = os.environ.get("ASN_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
warnings.filterwarnings("ignore")

from model.deeplab_multi import Classifier_Module as HeadMulti  # noqa: E402
from model.deeplab_vgg import Classifier_Module as HeadTwoBranch  # noqa: E402  (2-branch early return, Q9)
from model.discriminator import FCDiscriminator  # noqa: E402
from utils.loss import CrossEntropy2d  # noqa: E402
import compute_iou as ref_iou  # noqa: E402

SEED = 1338  # train_gta2cityscapes_multi.py:131
torch.set_num_threads(max(1, os.cpu_count() or 1))


def save(name, **arrs):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in arrs.items()})
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


def feature_like(gen, shape, rms, frac_zero):
    """post-ReLU-like activations."""
    x = torch.randn(shape, generator=gen).abs() * rms
    x[torch.rand(shape, generator=gen) < frac_zero] = 0
    return x


def gold_aspp():
    gen = torch.Generator().manual_seed(SEED)
    out = {}
    for tag, cls, cin, h, w in (("multi", HeadMulti, 64, 20, 28),
                                ("multi_odd", HeadMulti, 32, 45, 50),
                                ("twobranch", HeadTwoBranch, 64, 20, 28)):
        torch.manual_seed(SEED)
        head = cls(cin, [6, 12, 18, 24], [6, 12, 18, 24], 19)
        x = feature_like(gen, (1, cin, h, w), 1.5, 0.2).requires_grad_(True)
        y = head(x)
        dy = torch.randn(y.shape, generator=gen)
        y.backward(dy)
        out[tag + "_x"] = x.detach().numpy()
        out[tag + "_dy"] = dy.numpy()
        out[tag + "_y"] = y.detach().numpy()
        out[tag + "_dx"] = x.grad.numpy()
        for i, conv in enumerate(head.conv2d_list):
            out[f"{tag}_w{i}"] = conv.weight.detach().numpy()
            out[f"{tag}_b{i}"] = conv.bias.detach().numpy()
            out[f"{tag}_dw{i}"] = (conv.weight.grad if conv.weight.grad is not None
                                    else torch.zeros_like(conv.weight)).numpy()
            out[f"{tag}_db{i}"] = (conv.bias.grad if conv.bias.grad is not None
                                    else torch.zeros_like(conv.bias)).numpy()
    save("aspp", **out)


def gold_upsample():
    gen = torch.Generator().manual_seed(SEED + 1)
    out = {}
    for tag, nc, (h, w), (oh, ow) in (("a", (1, 4), (12, 20), (97, 161)), ("b", (2, 3), (9, 17), (65, 129)),
                                      ("c", (2, 19), (5, 7), (5, 7)), ("d", (1, 19), (8, 8), (1, 3)),
                                      ("e", (1, 19), (6, 10), (41, 73))):
        x = torch.randn(nc + (h, w), generator=gen).requires_grad_(True)
        y = nn.Upsample(size=(oh, ow), mode="bilinear", align_corners=True)(x)
        dy = torch.randn(y.shape, generator=gen)
        y.backward(dy)
        out[tag + "_x"] = x.detach().numpy()
        out[tag + "_y"] = y.detach().numpy()
        out[tag + "_dy"] = dy.numpy()
        out[tag + "_dx"] = x.grad.numpy()
    save("upsample", **out)


def gold_ce():
    gen = torch.Generator().manual_seed(SEED + 2)
    z = (torch.randn((2, 19, 33, 47), generator=gen) * 4).requires_grad_(True)
    y = torch.randint(0, 19, (2, 33, 47), generator=gen)
    y[torch.rand(y.shape, generator=gen) < 0.1] = 255
    loss = nn.CrossEntropyLoss(ignore_index=255)(z, y)
    loss.backward()
    out = dict(z=z.detach().numpy(), y=y.numpy(), loss=loss.item(), dz=z.grad.numpy(),
               n_valid=int((y != 255).sum()))
    # CrossEntropy2d incl. a negative label (utils/loss.py:29) and class weights
    y2 = y.clone()
    y2[0, 0, :5] = -1
    z2 = z.detach().clone().requires_grad_(True)
    l2 = CrossEntropy2d()(z2, y2)
    l2.backward()
    out.update(y_neg=y2.numpy(), loss_2d=l2.item(), dz_2d=z2.grad.numpy(),
               n_valid_2d=int(((y2 >= 0) & (y2 != 255)).sum()))
    wgt = torch.rand(19, generator=gen) + 0.5
    z3 = z.detach().clone().requires_grad_(True)
    l3 = CrossEntropy2d()(z3, y2, weight=wgt)
    l3.backward()
    out.update(weight=wgt.numpy(), loss_2d_w=l3.item(), dz_2d_w=z3.grad.numpy())
    z4 = z.detach().clone().requires_grad_(True)
    l4 = CrossEntropy2d(size_average=False)(z4, y2)
    l4.backward()
    out.update(loss_2d_sum=l4.item(), dz_2d_sum=z4.grad.numpy())
    # all ignored -> nan
    yall = torch.full_like(y, 255)
    out["loss_all_ignored"] = nn.CrossEntropyLoss(ignore_index=255)(z.detach(), yall).item()
    out["loss_2d_all_ignored"] = float(CrossEntropy2d()(z.detach(), yall))
    # the builtin and the class agree
    out["loss_2d_same_labels"] = CrossEntropy2d()(z.detach(), y).item()
    save("ce", **out)


def gold_softmax():
    gen = torch.Generator().manual_seed(SEED + 3)
    z = (torch.randn((2, 19, 17, 23), generator=gen) * 5).requires_grad_(True)
    p = F.softmax(z)  # implicit dim, as the reference calls it
    dp = torch.randn(p.shape, generator=gen)
    p.backward(dp)
    save("softmax", z=z.detach().numpy(), p=p.detach().numpy(), dp=dp.numpy(), dz=z.grad.numpy())


def fcd_params_from_seed(seed, num_classes, ndf):
    """Deterministic discriminator weights reproducible without the reference
    (torch CPU generator); shapes of model/discriminator.py:10-14."""
    gen = torch.Generator().manual_seed(seed)
    chans = [num_classes, ndf, ndf * 2, ndf * 4, ndf * 8, 1]
    sd = {}
    for i, name in enumerate(("conv1", "conv2", "conv3", "conv4", "classifier")):
        fan_in = chans[i] * 16
        bound = 1.0 / np.sqrt(fan_in)
        sd[name + ".weight"] = (torch.rand((chans[i + 1], chans[i], 4, 4), generator=gen) * 2 - 1) * bound
        sd[name + ".bias"] = (torch.rand((chans[i + 1],), generator=gen) * 2 - 1) * bound
    return sd


def gold_fcd():
    out = {}
    gen = torch.Generator().manual_seed(SEED + 4)
    for tag, ndf, (h, w) in (("small", 16, (66, 98)), ("ndf64", 64, (64, 96))):
        sd = fcd_params_from_seed(SEED + 40, 19, ndf)
        net = FCDiscriminator(19, ndf)
        net.load_state_dict(sd)
        z = torch.randn((1, 19, h, w), generator=gen) * 3
        x = F.softmax(z, dim=1).requires_grad_(True)
        o = net(x)
        do = torch.randn(o.shape, generator=gen)
        o.backward(do)
        out[tag + "_x"] = x.detach().numpy()
        out[tag + "_out"] = o.detach().numpy()
        out[tag + "_dout"] = do.numpy()
        out[tag + "_dx"] = x.grad.numpy()
        for name, p in net.named_parameters():
            g = p.grad.numpy()
            if tag == "small":
                out[f"{tag}_{name}"] = p.detach().numpy()
                out[f"{tag}_d_{name}"] = g
            else:  # weights come from the seed; keep only compact grad summaries
                out[f"{tag}_dsum_{name}"] = g.astype(np.float64).sum()
                out[f"{tag}_dl2_{name}"] = np.sqrt((g.astype(np.float64) ** 2).sum())
                out[f"{tag}_dhead_{name}"] = g.reshape(-1)[:64].copy()
    save("fcd", **out)


def gold_ganloss():
    gen = torch.Generator().manual_seed(SEED + 5)
    out = {}
    for tag, shape in (("src", (1, 1, 22, 40)), ("tgt", (1, 1, 16, 32)), ("odd", (3, 1, 5, 7))):
        x0 = torch.randn(shape, generator=gen) * 2
        out[tag + "_x"] = x0.numpy()
        for lname, fn in (("bce", nn.BCEWithLogitsLoss()), ("mse", nn.MSELoss())):
            for t in (0, 1):
                x = x0.clone().requires_grad_(True)
                # target built like train_gta2cityscapes_multi.py:621
                tgt = torch.FloatTensor(x.data.size()).fill_(t)
                loss = fn(x, tgt)
                loss.backward()
                out[f"{tag}_{lname}{t}_loss"] = loss.item()
                out[f"{tag}_{lname}{t}_dx"] = x.grad.numpy()
    save("ganloss", **out)


def gold_hist():
    rng = np.random.RandomState(SEED + 6)
    out = {}
    n = 19
    a = rng.randint(0, n, size=(64, 96)).astype(np.int64)
    a[rng.rand(*a.shape) < 0.10] = 255
    a[rng.rand(*a.shape) < 0.02] = 19
    a[rng.rand(*a.shape) < 0.02] = -1
    a[rng.rand(*a.shape) < 0.01] = 33
    b = rng.randint(0, n, size=a.shape).astype(np.uint8)
    out["a"], out["b"] = a, b
    out["hist"] = ref_iou.fast_hist(a.flatten(), b.flatten(), n)
    out["iu"] = ref_iou.per_class_iu(out["hist"].astype(np.float64))
    # blocky labels (realistic contention), uint8 labels as read from PNG
    a2 = np.repeat(np.repeat(rng.randint(0, n, size=(8, 12)), 8, 0), 8, 1).astype(np.uint8)
    a2[:8, :] = 255
    b2 = np.where(rng.rand(*a2.shape) < 0.8, np.minimum(a2, n - 1), rng.randint(0, n, size=a2.shape)).astype(np.uint8)
    out["a_u8"], out["b_u8"] = a2, b2
    out["hist_u8"] = ref_iou.fast_hist(a2.flatten(), b2.flatten(), n)
    # pred >= n under a valid label spills into the next row (flat bincount index)
    a3 = np.array([0, 0, 3, 18, 255, 5], dtype=np.int64)
    b3 = np.array([19, 37, 20, 0, 200, 5], dtype=np.uint8)
    out["a_spill"], out["b_spill"] = a3, b3
    out["hist_spill"] = ref_iou.fast_hist(a3, b3, n)
    # label_mapping with a Cityscapes-like id -> trainId table
    mapping = np.array([[i, 255] for i in range(7)] + [[7, 0], [8, 1], [11, 2], [12, 3], [13, 4],
                                                        [17, 5], [19, 6], [20, 7], [21, 8], [22, 9],
                                                        [23, 10], [24, 11], [25, 12], [26, 13],
                                                        [27, 14], [28, 15], [31, 16], [32, 17], [33, 18],
                                                        [-1, 255]], dtype=np.int64)
    raw = rng.randint(0, 34, size=(32, 48)).astype(np.uint8)
    out["map_table"], out["map_in"] = mapping, raw
    out["map_out"] = ref_iou.label_mapping(raw, mapping)
    save("hist", **out)


def gold_argmax():
    gen = torch.Generator().manual_seed(SEED + 7)
    x = torch.randn((1, 19, 16, 32), generator=gen) * 3
    # a few exact ties so the first-max rule is exercised
    x[0, 3] = x[0, 7]
    interp = nn.Upsample(size=(64, 128), mode="bilinear", align_corners=True)
    output = interp(x).cpu().data[0].numpy()          # evaluate_cityscapes.py:163
    output = output.transpose(1, 2, 0)                 # :168
    output = np.asarray(np.argmax(output, axis=2), dtype=np.uint8)  # :169
    save("argmax", x=x.numpy(), pred=output)


def gold_eval2():
    """Evaluation chain at shapes where the two bilinear stages do NOT compose into one ((out-1) % (in-1) != 0):
    ResNetMulti.forward's own upsample to the input size (model/deeplab_multi.py:188-189) followed by the script's
    interp to the label size and the CPU argmax (evaluate_cityscapes.py:153,163,168-169)."""
    import hashlib
    from model.deeplab_multi import DeeplabMulti

    out = {}
    # (a) the reference's own model end to end on a 72x136 image: features 9x17 -> 72x136 (ratio 71/8) -> 144x272
    torch.manual_seed(SEED)
    model = DeeplabMulti(19).eval()
    keep = {}
    model.layer6.register_forward_hook(lambda m, i, o: keep.__setitem__("low", o.detach().clone()))
    gen = torch.Generator().manual_seed(SEED + 11)
    image = torch.randn((1, 3, 72, 136), generator=gen) * 50
    interp = nn.Upsample(size=(144, 272), mode="bilinear", align_corners=True)      # evaluate_cityscapes.py:153
    with torch.no_grad():
        _, output2 = model(image, (136, 72))                                         # :162 (input_size = (W, H), Q1)
        output = interp(output2).cpu().data[0].numpy()                               # :163
    pred = np.asarray(np.argmax(output.transpose(1, 2, 0), axis=2), dtype=np.uint8)   # :168-169
    out["model_low"] = keep["low"].numpy()
    out["model_mid_hw"] = np.array(output2.shape[-2:])
    out["model_mid_sha"] = hashlib.sha256(output2.numpy().tobytes()).hexdigest()
    out["model_pred"] = pred
    # (b) fixed logits with exact ties, 33x65 -> 259x515 -> 518x1030
    x = torch.randn((1, 19, 33, 65), generator=gen) * 3
    x[0, 3] = x[0, 7]
    mid = nn.Upsample(size=(259, 515), mode="bilinear", align_corners=True)(x)
    up = nn.Upsample(size=(518, 1030), mode="bilinear", align_corners=True)(mid)
    out["mid_x"] = x.numpy()
    out["mid_sha"] = hashlib.sha256(mid.numpy().tobytes()).hexdigest()
    out["mid_pred"] = np.asarray(np.argmax(up[0].numpy().transpose(1, 2, 0), axis=2), dtype=np.uint8)
    # (c) BASELINE config 5 shapes: 64x128 -> 512x1024 -> 1024x2048; the input is re-created from the seed (CPU
    #     generator, platform independent), the outputs are pinned by digest and class histogram
    g5 = torch.Generator().manual_seed(SEED + 5)
    x5 = torch.randn((1, 19, 64, 128), generator=g5) * 3
    mid5 = nn.Upsample(size=(512, 1024), mode="bilinear", align_corners=True)(x5)
    up5 = nn.Upsample(size=(1024, 2048), mode="bilinear", align_corners=True)(mid5)
    pred5 = np.asarray(np.argmax(up5[0].numpy().transpose(1, 2, 0), axis=2), dtype=np.uint8)
    one = nn.Upsample(size=(1024, 2048), mode="bilinear", align_corners=True)(x5)
    pred5_one = np.asarray(np.argmax(one[0].numpy().transpose(1, 2, 0), axis=2), dtype=np.uint8)
    out["cfg5_seed"] = np.array(SEED + 5)
    out["cfg5_x_sha"] = hashlib.sha256(x5.numpy().tobytes()).hexdigest()
    out["cfg5_mid_sha"] = hashlib.sha256(mid5.numpy().tobytes()).hexdigest()
    out["cfg5_pred_sha"] = hashlib.sha256(pred5.tobytes()).hexdigest()
    out["cfg5_pred_bincount"] = np.bincount(pred5.ravel(), minlength=19)
    out["cfg5_one_stage_differs"] = np.array(int((pred5 != pred5_one).sum()))   # what a single-stage resize gets wrong
    print("cfg5: one-stage vs two-stage argmax differs on", int((pred5 != pred5_one).sum()), "pixels")
    save("eval2", **out)


def gold_preprocess():
    """dataset/gta5_dataset.py:47-71 run for real on two PNG files of exactly the crop size (so that its PIL resize is the
    identity): pins the tensor-forming tail -- BGR flip, mean subtraction, CHW transpose, id -> train id remap."""
    import tempfile
    import types
    from PIL import Image
    sys.modules.setdefault("matplotlib", types.ModuleType("matplotlib"))          # import-only dependency
    sys.modules.setdefault("matplotlib.pyplot", types.ModuleType("matplotlib.pyplot"))
    from dataset.gta5_dataset import GTA5DataSet
    rng = np.random.RandomState(SEED)
    H, W = 32, 48
    rgb = rng.randint(0, 256, size=(H, W, 3)).astype(np.uint8)
    ids = rng.randint(0, 35, size=(H, W)).astype(np.uint8)
    mean = np.array((104.00698793, 116.66876762, 122.67891434), dtype=np.float32)   # train...:30
    with tempfile.TemporaryDirectory() as d:
        os.makedirs(os.path.join(d, "images"))
        os.makedirs(os.path.join(d, "labels"))
        Image.fromarray(rgb).save(os.path.join(d, "images", "a.png"))
        Image.fromarray(ids).save(os.path.join(d, "labels", "a.png"))
        with open(os.path.join(d, "list.txt"), "w") as f:
            f.write("a.png\n")
        ds = GTA5DataSet(d, os.path.join(d, "list.txt"), crop_size=(W, H), scale=False, mirror=False, mean=mean)
        image, label, size, name = ds[0]
    save("preprocess", rgb=rgb, ids=ids, mean=mean, image=image, label=label, size=size)


def gold_step():
    """whole multi-level / single-level iteration (train_gta2cityscapes_multi.py:560-683, :373-464) driven
    through oracle/torch_ref.RefTrainer's restated loop with the REFERENCE's own modules injected."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle import torch_ref as TR
    from model.deeplab_multi import DeeplabMulti

    out = {}
    for level, gan in (("multi-level", "Vanilla"), ("single-level", "LS")):
        tag = "multi" if level == "multi-level" else "single"
        G = TR.seeded_init_(DeeplabMulti(19), SEED)
        D1 = TR.seeded_init_(FCDiscriminator(19), SEED + 1)
        D2 = TR.seeded_init_(FCDiscriminator(19), SEED + 2)
        tr = TR.RefTrainer(level=level, gan=gan, model=G, model_D1=D1, model_D2=D2)
        src, lab, tgt = TR.synthetic_batch(SEED, (129, 257), (97, 193))
        losses = tr.step(src, lab, tgt, i_iter=0, do_optimizer_step=False)
        for k, v in losses.items():
            out[f"{tag}_{k}"] = v
        for name, mod in (("G", G), ("D1", D1), ("D2", D2)):
            if name == "D1" and level != "multi-level":
                continue
            for pn, p in mod.named_parameters():
                if p.grad is None:
                    continue
                gnp = p.grad.numpy().astype(np.float64)
                if pn.startswith(("layer5", "layer6")) or name != "G" or pn == "conv1.weight" \
                        or pn.endswith("layer4.2.conv3.weight"):
                    out[f"{tag}_{name}_{pn}_l2"] = np.sqrt((gnp ** 2).sum())
                    out[f"{tag}_{name}_{pn}_head"] = gnp.reshape(-1)[:32].astype(np.float32)
        print(tag, losses)
    save("step", **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["aspp", "upsample", "ce", "softmax", "fcd", "ganloss", "hist", "argmax", "eval2", "preprocess", "step"]
    for name in which:
        globals()["gold_" + name]()
