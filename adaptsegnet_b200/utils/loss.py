"""Losses of the adaptation path behind the reference's ``utils/loss.py`` surface."""
from __future__ import annotations

import torch.nn as nn

from .. import ops


class CrossEntropy2d(nn.Module):
    """2-D cross entropy with an ignore label (reference utils/loss.py:7-36): the mask
    ``(target >= 0) * (target != ignore_label)``, the NHWC gather and F.cross_entropy collapse into
    one fused softmax-CE kernel."""

    def __init__(self, size_average=True, ignore_label=255):
        super().__init__()
        self.size_average = size_average
        self.ignore_label = ignore_label

    def forward(self, predict, target, weight=None):
        assert not target.requires_grad
        assert predict.dim() == 4
        assert target.dim() == 3
        assert predict.size(0) == target.size(0), "{0} vs {1} ".format(predict.size(0), target.size(0))
        assert predict.size(2) == target.size(1), "{0} vs {1} ".format(predict.size(2), target.size(1))
        assert predict.size(3) == target.size(2), "{0} vs {1} ".format(predict.size(3), target.size(2))
        return ops.softmax_cross_entropy(predict, target.long(), ignore_label=self.ignore_label, mask_negative=True,
                                         weight=weight, size_average=self.size_average)


class SegCrossEntropy(nn.Module):
    """torch.nn.CrossEntropyLoss(ignore_index=255) as the training script uses it
    (train_gta2cityscapes_multi.py:248,359,546)."""

    def __init__(self, ignore_index=255):
        super().__init__()
        self.ignore_index = ignore_index

    def forward(self, predict, target):
        return ops.softmax_cross_entropy(predict, target, ignore_label=self.ignore_index)


class GANLoss(nn.Module):
    """BCEWithLogitsLoss ('Vanilla') or MSELoss ('LS') against a constant source/target label
    (train_gta2cityscapes_multi.py:355-358, applied :620-624).  Takes the label as a scalar: no target
    tensor is built on the CPU and copied per call."""

    def __init__(self, gan="Vanilla"):
        super().__init__()
        if gan not in ("Vanilla", "LS"):
            raise ValueError("gan must be 'Vanilla' or 'LS'")
        self.kind = ops.GAN_BCE if gan == "Vanilla" else ops.GAN_MSE

    def forward(self, d_out, label):
        return ops.gan_loss(d_out, float(label), self.kind)
