"""Restatement of the reference's modules and training / evaluation loops in plain PyTorch
(TEST INFRASTRUCTURE: the CPU baseline of bench.py and the whole-step parity oracle).

/root/reference cannot travel to the GPU box, and its loops are not callable anyway (they live in
``main()`` and the fork's multi-level branch omits a required argument -- SURVEY.md Q1, Q2).  This
file restates, with the reference's own torch calls and in the reference's order:
  * DeeplabMulti (ResNet-101 trunk + two Classifier_Module heads + nn.Upsample)
        model/deeplab_multi.py:59-260
  * FCDiscriminator                      model/discriminator.py:5-34
  * the multi-level iteration            train_gta2cityscapes_multi.py:560-683
  * the single-level iteration           train_gta2cityscapes_multi.py:373-464
  * the eval step                        evaluate_cityscapes.py:153-169
State-dict keys equal the reference's, so weights move freely between the reference, this file and
adaptsegnet_b200.  Pinned against the real reference by tests/test_torch_ref_pin.py (runs where
/root/reference is mounted) and by tests/golden/step.npz.
"""
from __future__ import annotations

import warnings

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


# ---------------------------------------------------------------------------------------------
# model/deeplab_multi.py
# ---------------------------------------------------------------------------------------------
def _bn(c):
    m = nn.BatchNorm2d(c, affine=True)
    for p in m.parameters():
        p.requires_grad = False
    return m


class RefBottleneck(nn.Module):  # model/deeplab_multi.py:59-103
    def __init__(self, cin, planes, stride=1, dilation=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, planes, kernel_size=1, stride=stride, bias=False)
        self.bn1 = _bn(planes)
        self.conv2 = nn.Conv2d(planes, planes, kernel_size=3, stride=1, padding=dilation, bias=False, dilation=dilation)
        self.bn2 = _bn(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, kernel_size=1, bias=False)
        self.bn3 = _bn(planes * 4)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample

    def forward(self, x):
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.relu(self.bn2(self.conv2(out)))
        out = self.bn3(self.conv3(out))
        out += x if self.downsample is None else self.downsample(x)
        return self.relu(out)


class RefClassifier(nn.Module):  # model/deeplab_multi.py:106-121
    def __init__(self, cin, rates, num_classes):
        super().__init__()
        self.conv2d_list = nn.ModuleList(
            [nn.Conv2d(cin, num_classes, kernel_size=3, stride=1, padding=r, dilation=r, bias=True) for r in rates])

    def forward(self, x):
        out = self.conv2d_list[0](x)
        for i in range(len(self.conv2d_list) - 1):
            out += self.conv2d_list[i + 1](x)
        return out


class RefDeeplabMulti(nn.Module):  # model/deeplab_multi.py:124-235, DeeplabMulti :258-260
    def __init__(self, num_classes=19, layers=(3, 4, 23, 3)):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = _bn(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1, ceil_mode=False)
        cin = 64
        for i, (planes, n, stride, dil) in enumerate(zip((64, 128, 256, 512), layers, (1, 2, 1, 1), (1, 1, 2, 4))):
            ds = nn.Sequential(nn.Conv2d(cin, planes * 4, kernel_size=1, stride=stride, bias=False), _bn(planes * 4))
            blocks = [RefBottleneck(cin, planes, stride, dil, ds)]
            cin = planes * 4
            blocks += [RefBottleneck(cin, planes, dilation=dil) for _ in range(1, n)]
            setattr(self, f"layer{i + 1}", nn.Sequential(*blocks))
        self.layer5 = RefClassifier(1024, (6, 12, 18, 24), num_classes)
        self.layer6 = RefClassifier(2048, (6, 12, 18, 24), num_classes)
        for m in self.modules():  # :144-150
            if isinstance(m, nn.Conv2d):
                m.weight.data.normal_(0, 0.01)
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()

    def forward(self, x, input_size):  # :174-194 (warper=None)
        x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
        x = self.layer3(self.layer2(self.layer1(x)))
        x1 = self.layer5(x)
        x2 = self.layer6(self.layer4(x))
        up = nn.Upsample(size=(input_size[1], input_size[0]), mode="bilinear", align_corners=True)
        return up(x1), up(x2)

    def optim_parameters(self, lr):  # :196-235, duplicates included (Q11)
        def one_x():
            for root in (self.conv1, self.bn1, self.layer1, self.layer2, self.layer3, self.layer4):
                for j in root.modules():
                    for k in j.parameters():
                        if k.requires_grad:
                            yield k

        def ten_x():
            for part in (self.layer5.parameters(), self.layer6.parameters()):
                for p in part:
                    yield p

        return [{"params": one_x(), "lr": lr}, {"params": ten_x(), "lr": 10 * lr}]


class RefClassifierTwoBranch(RefClassifier):
    """model/deeplab_vgg.py:6-21 (and model/deeplab.py:101-116): the `return` sits INSIDE the loop, so only branches 0 and 1
    are summed although all four own weights"""

    def forward(self, x):
        out = self.conv2d_list[0](x)
        for i in range(len(self.conv2d_list) - 1):
            out += self.conv2d_list[i + 1](x)
            return out


class RefDeeplabVGG(nn.Module):
    """model/deeplab_vgg.py:24-54 with the Python-3 fix of its constructor (`range(23)+range(24,30)`, Q10).  forward()
    follows the single-level loop's use of a one-output model: logits are interpolated to the input size by the caller
    (evaluate_cityscapes.py:164-166); here `forward(x, input_size)` returns (None, interp(logits)) so that the restated
    training loop (RefTrainer) drives it like DeeplabMulti."""

    def __init__(self, num_classes=19):
        super().__init__()
        from torchvision import models
        features = list(models.vgg16().features.children())
        features = [features[i] for i in list(range(23)) + list(range(24, 30))]
        for i in (23, 25, 27):
            features[i].dilation = (2, 2)
            features[i].padding = (2, 2)
        fc6 = nn.Conv2d(512, 1024, kernel_size=3, padding=4, dilation=4)
        fc7 = nn.Conv2d(1024, 1024, kernel_size=3, padding=4, dilation=4)
        self.features = nn.Sequential(*(features + [fc6, nn.ReLU(inplace=True), fc7, nn.ReLU(inplace=True)]))
        self.classifier = RefClassifierTwoBranch(1024, (6, 12, 18, 24), num_classes)
        for m in self.classifier.conv2d_list:
            m.weight.data.normal_(0, 0.01)

    def forward(self, x, input_size):
        z = self.classifier(self.features(x))
        return None, nn.Upsample(size=(input_size[1], input_size[0]), mode="bilinear", align_corners=True)(z)

    def optim_parameters(self, lr):   # deeplab_vgg.py:53-54: one group
        return [{"params": list(self.parameters()), "lr": lr}]


class RefFCDiscriminator(nn.Module):  # model/discriminator.py:5-34
    def __init__(self, num_classes, ndf=64):
        super().__init__()
        self.conv1 = nn.Conv2d(num_classes, ndf, kernel_size=4, stride=2, padding=1)
        self.conv2 = nn.Conv2d(ndf, ndf * 2, kernel_size=4, stride=2, padding=1)
        self.conv3 = nn.Conv2d(ndf * 2, ndf * 4, kernel_size=4, stride=2, padding=1)
        self.conv4 = nn.Conv2d(ndf * 4, ndf * 8, kernel_size=4, stride=2, padding=1)
        self.classifier = nn.Conv2d(ndf * 8, 1, kernel_size=4, stride=2, padding=1)
        self.leaky_relu = nn.LeakyReLU(negative_slope=0.2, inplace=True)

    def forward(self, x):
        for conv in (self.conv1, self.conv2, self.conv3, self.conv4):
            x = self.leaky_relu(conv(x))
        return self.classifier(x)


# ---------------------------------------------------------------------------------------------
# train_gta2cityscapes_multi.py
# ---------------------------------------------------------------------------------------------
def lr_poly(base_lr, it, max_iter, power):  # :162-163
    return base_lr * ((1 - float(it) / max_iter) ** power)


class _Args:
    """what ResNetMulti.optim_parameters(args) reads (model/deeplab_multi.py:233-235)"""

    def __init__(self, lr):
        self.learning_rate = lr


def seeded_init_(module, seed):
    """Deterministic weights that do not depend on constructor RNG order: every parameter, in
    named_parameters() order, is redrawn from one seeded CPU generator (BatchNorm affine left at 1/0).
    Conv weights of the segmentation net ~ N(0, 0.01) as model/deeplab_multi.py:144-147; everything else
    (biases, discriminator weights) uniform in +-1/sqrt(fan_in) like nn.Conv2d's default."""
    gen = torch.Generator().manual_seed(seed)
    bn = {id(p) for m in module.modules() if isinstance(m, nn.BatchNorm2d) for p in m.parameters()}
    is_seg = hasattr(module, "layer5")
    fan_in = {}
    for m in module.modules():
        if isinstance(m, nn.Conv2d):
            f = m.weight.shape[1] * m.weight.shape[2] * m.weight.shape[3]
            fan_in[id(m.weight)] = f
            if m.bias is not None:
                fan_in[id(m.bias)] = f
    with torch.no_grad():
        for _, p in module.named_parameters():
            if id(p) in bn:
                continue
            if is_seg and p.dim() == 4:
                p.copy_(torch.randn(p.shape, generator=gen) * 0.01)
            else:
                bound = 1.0 / np.sqrt(fan_in[id(p)])
                p.copy_((torch.rand(p.shape, generator=gen) * 2 - 1) * bound)
    return module


class RefTrainer:
    """train_gta2cityscapes_multi.py:500-546 (setup) and :560-683 / :373-464 (iteration)."""

    def __init__(self, num_classes=19, level="multi-level", gan="Vanilla", device="cpu", learning_rate=2.5e-4,
                 momentum=0.9, weight_decay=0.0005, learning_rate_D=1e-4, power=0.9, num_steps=250000,
                 lambda_seg=0.1, lambda_adv_target1=0.0002, lambda_adv_target2=0.001, iter_size=1,
                 model=None, model_D1=None, model_D2=None):
        """``model`` / ``model_D*``: inject other module instances with the same interface (the golden
        generator injects the reference's own modules here)."""
        self.device = torch.device(device)
        self.multi = level == "multi-level"
        self.h = dict(lr=learning_rate, lr_D=learning_rate_D, power=power, num_steps=num_steps, lambda_seg=lambda_seg,
                      l1=lambda_adv_target1, l2=lambda_adv_target2, iter_size=iter_size)
        self.model = (model or RefDeeplabMulti(num_classes)).to(self.device).train()
        self.model_D1 = (model_D1 or RefFCDiscriminator(num_classes)).to(self.device).train() if self.multi else None
        self.model_D2 = (model_D2 or RefFCDiscriminator(num_classes)).to(self.device).train()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            groups = (self.model.optim_parameters(learning_rate) if isinstance(self.model, (RefDeeplabMulti, RefDeeplabVGG))
                      else self.model.optim_parameters(_Args(learning_rate)))
            self.optimizer = torch.optim.SGD(groups, lr=learning_rate,
                                             momentum=momentum, weight_decay=weight_decay)
        self.optimizer_D1 = (torch.optim.Adam(self.model_D1.parameters(), lr=learning_rate_D, betas=(0.9, 0.99))
                             if self.multi else None)
        self.optimizer_D2 = torch.optim.Adam(self.model_D2.parameters(), lr=learning_rate_D, betas=(0.9, 0.99))
        self.bce_loss = torch.nn.BCEWithLogitsLoss() if gan == "Vanilla" else torch.nn.MSELoss()  # :542-545
        self.seg_loss = torch.nn.CrossEntropyLoss(ignore_index=255)  # :546

    def _target(self, d_out, label):  # :621 -- built on the CPU every call (Q15)
        return torch.FloatTensor(d_out.data.size()).fill_(label).to(self.device)

    def step(self, src_images, src_labels, tgt_images, i_iter=0, do_optimizer_step=True):
        h = self.h
        it = h["iter_size"]
        dev = self.device
        Ds = [d for d in (self.model_D1, self.model_D2) if d is not None]
        opts_D = [o for o in (self.optimizer_D1, self.optimizer_D2) if o is not None]
        self.optimizer.zero_grad()
        lr = lr_poly(h["lr"], i_iter, h["num_steps"], h["power"])
        self.optimizer.param_groups[0]["lr"] = lr
        if len(self.optimizer.param_groups) > 1:
            self.optimizer.param_groups[1]["lr"] = lr * 10
        for o in opts_D:
            o.zero_grad()
            o.param_groups[0]["lr"] = lr_poly(h["lr_D"], i_iter, h["num_steps"], h["power"])
        out = {}
        for D in Ds:
            for p in D.parameters():
                p.requires_grad = False
        images = src_images.to(dev)
        labels = src_labels.long().to(dev)
        size = (images.shape[3], images.shape[2])  # input_size is (W, H): model/deeplab_multi.py:188
        pred1, pred2 = self.model(images, size)
        loss_seg2 = self.seg_loss(pred2, labels)
        if self.multi:
            loss_seg1 = self.seg_loss(pred1, labels)
            loss = loss_seg2 + h["lambda_seg"] * loss_seg1
        else:
            loss = loss_seg2
        (loss / it).backward()
        if self.multi:
            out["loss_seg1"] = loss_seg1.item() / it
        out["loss_seg2"] = loss_seg2.item() / it

        images = tgt_images.to(dev)
        size_t = (images.shape[3], images.shape[2])
        pred_target1, pred_target2 = self.model(images, size_t)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")  # F.softmax without dim, as the reference calls it (Q8)
            D_out2 = self.model_D2(F.softmax(pred_target2))
            loss_adv2 = self.bce_loss(D_out2, self._target(D_out2, 0))
            loss = h["l2"] * loss_adv2
            if self.multi:
                D_out1 = self.model_D1(F.softmax(pred_target1))
                loss_adv1 = self.bce_loss(D_out1, self._target(D_out1, 0))
                loss = h["l1"] * loss_adv1 + loss
            (loss / it).backward()
            if self.multi:
                out["loss_adv_target1"] = loss_adv1.item() / it
            out["loss_adv_target2"] = loss_adv2.item() / it

            for D in Ds:
                for p in D.parameters():
                    p.requires_grad = True
            levels = [(self.model_D2, pred2, pred_target2, "loss_D2")]
            if self.multi:
                levels.insert(0, (self.model_D1, pred1, pred_target1, "loss_D1"))
            for D, p_src, p_tgt, name in levels:
                d = D(F.softmax(p_src.detach()))
                l_src = self.bce_loss(d, self._target(d, 0)) / it / 2
                l_src.backward()
                d = D(F.softmax(p_tgt.detach()))
                l_tgt = self.bce_loss(d, self._target(d, 1)) / it / 2
                l_tgt.backward()
                out[name] = l_src.item() + l_tgt.item()
        if do_optimizer_step:
            self.optimizer.step()
            for o in opts_D:
                o.step()
        return out


# ---------------------------------------------------------------------------------------------
# evaluate_cityscapes.py:153-169 and compute_iou.py:50-61
# ---------------------------------------------------------------------------------------------
def eval_step(model, image, size=(1024, 2048)):
    """output2 -> interp -> cpu -> transpose -> argmax -> uint8"""
    interp = nn.Upsample(size=size, mode="bilinear", align_corners=True)
    with torch.no_grad():
        _, output2 = model(image, (image.shape[3], image.shape[2]))
        output = interp(output2).cpu().data[0].numpy()
    output = output.transpose(1, 2, 0)
    return np.asarray(np.argmax(output, axis=2), dtype=np.uint8)


def synthetic_batch(seed, src_hw, tgt_hw, num_classes=19, device="cpu"):
    """mean-subtracted BGR-like images, blocky labels with ~10% ignore."""
    g = torch.Generator().manual_seed(seed)
    mean = torch.tensor([104.00698793, 116.66876762, 122.67891434]).view(1, 3, 1, 1)
    src = torch.randint(0, 256, (1, 3) + tuple(src_hw), generator=g).float() - mean
    tgt = torch.randint(0, 256, (1, 3) + tuple(tgt_hw), generator=g).float() - mean
    bh, bw = max(1, src_hw[0] // 16), max(1, src_hw[1] // 16)
    coarse = torch.randint(0, num_classes, (1, 1, bh, bw), generator=g).float()
    lab = F.interpolate(coarse, size=tuple(src_hw), mode="nearest")[0].long()
    ign = F.interpolate((torch.rand((1, 1, bh, bw), generator=g) < 0.1).float(), size=tuple(src_hw), mode="nearest")[0]
    lab[ign > 0] = 255
    return src.to(device), lab.to(device), tgt.to(device)
