"""DeepLab-v2 (ResNet-101) with two ASPP classifier heads -- the reference's
``model/deeplab_multi.py`` surface (`DeeplabMulti`, `ResNetMulti`, `Classifier_Module`,
`Bottleneck`), state-dict compatible key for key.

Only the hot path is new: `Classifier_Module.forward` (reference model/deeplab_multi.py:117-121)
and the bilinear upsample at the end of `ResNetMulti.forward` (:188-189) run on libasn_b200's
sm_100a kernels.  The ResNet-101 trunk stays ordinary PyTorch modules (cuDNN), as the north star
prescribes: "timed but not rewritten".
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops

# (planes, blocks, stride, dilation) of layer1..layer4 -- reference :137-140
_TRUNK = ((64, None, 1, 1), (128, None, 2, 1), (256, None, 1, 2), (512, None, 1, 4))
_ASPP_RATES = (6, 12, 18, 24)  # reference :141-142


def _frozen_bn(channels: int) -> nn.BatchNorm2d:
    """BatchNorm with affine parameters excluded from training (reference :65-78, :130-132);
    it still normalises with batch statistics in train() mode."""
    bn = nn.BatchNorm2d(channels, affine=True)
    for p in bn.parameters():
        p.requires_grad = False
    return bn


class Bottleneck(nn.Module):
    """1x1 (strided) -> 3x3 (dilated) -> 1x1 (x4) residual unit; attribute names fix the
    state-dict keys (conv1/bn1/conv2/bn2/conv3/bn3/downsample)."""
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None):
        super().__init__()
        width_out = planes * self.expansion
        self.conv1 = nn.Conv2d(inplanes, planes, 1, stride=stride, bias=False)
        self.bn1 = _frozen_bn(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride=1, padding=dilation, dilation=dilation, bias=False)
        self.bn2 = _frozen_bn(planes)
        self.conv3 = nn.Conv2d(planes, width_out, 1, bias=False)
        self.bn3 = _frozen_bn(width_out)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        shortcut = x if self.downsample is None else self.downsample(x)
        y = self.relu(self.bn1(self.conv1(x)))
        y = self.relu(self.bn2(self.conv2(y)))
        y = self.bn3(self.conv3(y))
        y += shortcut
        return self.relu(y)


class Classifier_Module(nn.Module):
    """ASPP head: sum of dilated 3x3 convolutions to `num_classes` logits.

    Same constructor and parameters (``conv2d_list.{i}.weight/bias``, fp32 OIHW) as the reference
    (model/deeplab_multi.py:106-115); the forward is one call into the sm_100a head kernels
    (ops.aspp_head).  ``active_branches`` reproduces the early-return variants of
    model/deeplab.py:112-116 / model/deeplab_vgg.py:17-21, which sum only the first two branches
    (SURVEY.md Q9); the default sums all of them like deeplab_multi.py.
    """

    def __init__(self, inplanes, dilation_series, padding_series, num_classes, active_branches=None):
        super().__init__()
        if list(dilation_series) != list(padding_series):
            raise ValueError("the head kernels cover padding == dilation (as every reference call site uses)")
        self.conv2d_list = nn.ModuleList(
            nn.Conv2d(inplanes, num_classes, 3, stride=1, padding=p, dilation=d, bias=True)
            for d, p in zip(dilation_series, padding_series))
        for conv in self.conv2d_list:
            conv.weight.data.normal_(0, 0.01)
        self.dilations = tuple(int(d) for d in dilation_series)
        self.active_branches = len(self.dilations) if active_branches is None else int(active_branches)
        self._pack = ops.AsppWeightPack()

    def forward(self, x):
        weights = [c.weight for c in self.conv2d_list]
        biases = [c.bias for c in self.conv2d_list]
        return ops.aspp_head(x, weights, biases, self.dilations, self.active_branches, self._pack)


class ResNetMulti(nn.Module):
    def __init__(self, block, layers, num_classes):
        super().__init__()
        self.inplanes = 64
        self.conv1 = nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False)
        self.bn1 = _frozen_bn(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(3, stride=2, padding=1, ceil_mode=False)  # Q14: H/8 x W/8 exactly
        for idx, ((planes, _, stride, dilation), n_blocks) in enumerate(zip(_TRUNK, layers), start=1):
            setattr(self, f"layer{idx}", self._make_layer(block, planes, n_blocks, stride, dilation))
        self.layer5 = self._make_pred_layer(Classifier_Module, 1024, _ASPP_RATES, _ASPP_RATES, num_classes)
        self.layer6 = self._make_pred_layer(Classifier_Module, 2048, _ASPP_RATES, _ASPP_RATES, num_classes)
        for m in self.modules():  # reference :144-150
            if isinstance(m, nn.Conv2d):
                m.weight.data.normal_(0, 0.01)
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()

    def _make_layer(self, block, planes, blocks, stride=1, dilation=1):
        out_ch = planes * block.expansion
        downsample = None
        if stride != 1 or self.inplanes != out_ch or dilation in (2, 4):
            downsample = nn.Sequential(nn.Conv2d(self.inplanes, out_ch, 1, stride=stride, bias=False),
                                       _frozen_bn(out_ch))
        units = [block(self.inplanes, planes, stride, dilation=dilation, downsample=downsample)]
        self.inplanes = out_ch
        units += [block(self.inplanes, planes, dilation=dilation) for _ in range(1, blocks)]
        return nn.Sequential(*units)

    def _make_pred_layer(self, block, inplanes, dilation_series, padding_series, num_classes):
        return block(inplanes, list(dilation_series), list(padding_series), num_classes)

    # Execution mode of the (unchanged) trunk modules, SURVEY.md 8f row 1: run them under torch.autocast(bfloat16).
    # Off by default -- the reference computes the trunk in fp32 (TF32 on the GPU); the heads always receive fp32 features.
    trunk_autocast = False
    # True: forward() returns lazy upsampled-logits handles (adaptsegnet_b200/lazy.py) instead of full-resolution tensors
    lazy_outputs = False

    def trunk(self, x):
        """ResNet-101 features: (layer3 output, layer4 output), both H/8 x W/8."""
        if self.trunk_autocast and x.is_cuda:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                f3, f4 = self._trunk(x)
            return f3.float(), f4.float()
        return self._trunk(x)

    # optional callable(name, tensor) the trainer sets around ONE forward to learn when the backward has passed layer4 /
    # layer3 (it hangs gradient hooks on the two activations; used to start bucketed gradient all-reduces early)
    grad_probe = None

    def _trunk(self, x):
        x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
        l2 = self.layer2(self.layer1(x))
        f3 = self.layer3(l2)
        if self.grad_probe is not None:
            self.grad_probe("layer4.0", f3)
        return f3, self.layer4(f3)

    def forward(self, x, input_size=None, warper=None):
        """-> (x1_up, x2_up), each (N, C, input_size[1], input_size[0]).

        ``input_size`` is (W, H) as in the reference (:188-189).  It is optional here: the fork's
        own multi-level and eval call sites omit it (train...:597,615; evaluate...:162 -- SURVEY.md
        Q1); None upsamples to the input image's own size, which is what those call sites mean.
        """
        if warper is not None:
            raise NotImplementedError("the fork-only Warper path (SURVEY.md Q6) is outside the hot path")
        size = (int(x.shape[2]), int(x.shape[3])) if input_size is None else (int(input_size[1]), int(input_size[0]))
        f3, f4 = self.trunk(x)
        x1 = self.layer5(f3)
        x2 = self.layer6(f4)
        if self.lazy_outputs:
            # Tier-B behind the unchanged scripts: handles that route the script's own CrossEntropyLoss / F.softmax /
            # .detach() / nn.Upsample calls to the fused kernels and materialise on anything else (lazy.py)
            from ..lazy import UpsampledLogits
            return UpsampledLogits.make(x1, size), UpsampledLogits.make(x2, size)
        return ops.upsample_bilinear(x1, size), ops.upsample_bilinear(x2, size)

    def low_res_logits(self, x):
        """(layer5(x3), layer6(x4)) before the upsample -- for the fused eval path (K9)."""
        f3, f4 = self.trunk(x)
        return self.layer5(f3), self.layer6(f4)

    def get_1x_lr_params_NOscale(self):
        """Trainable trunk parameters.  Like the reference generator (:196-218) this walks every
        sub-module and then every parameter below it, so a parameter is yielded once per enclosing
        module; optimizers built from it behave exactly as with the reference."""
        for root in (self.conv1, self.bn1, self.layer1, self.layer2, self.layer3, self.layer4):
            for module in root.modules():
                for p in module.parameters():
                    if p.requires_grad:
                        yield p

    def get_10x_lr_params(self):
        yield from self.layer5.parameters()
        yield from self.layer6.parameters()

    def optim_parameters(self, args):
        return [{"params": self.get_1x_lr_params_NOscale(), "lr": args.learning_rate},
                {"params": self.get_10x_lr_params(), "lr": 10 * args.learning_rate}]


def DeeplabMulti(num_classes=21):
    return ResNetMulti(Bottleneck, [3, 4, 23, 3], num_classes)
