"""Fully-convolutional patch discriminator -- the reference's ``model/discriminator.py`` surface.

Parameters live in the same five ``nn.Conv2d`` holders (state-dict keys ``conv1..conv4,
classifier``), but `forward` is one call into the sm_100a implicit-GEMM kernels (ops.fcd_forward):
4x4 stride-2 convolutions with bias + LeakyReLU(0.2) fused in the epilogue
(reference model/discriminator.py:21-34)."""
from __future__ import annotations

import torch.nn as nn

from .. import ops


class FCDiscriminator(nn.Module):
    def __init__(self, num_classes, ndf=64):
        super().__init__()
        chans = (num_classes, ndf, ndf * 2, ndf * 4, ndf * 8)
        for i in range(4):
            setattr(self, f"conv{i + 1}", nn.Conv2d(chans[i], chans[i + 1], kernel_size=4, stride=2, padding=1))
        self.classifier = nn.Conv2d(chans[4], 1, kernel_size=4, stride=2, padding=1)
        self.leaky_relu = nn.LeakyReLU(negative_slope=0.2, inplace=True)  # kept for API parity; fused in-kernel
        self._pack = ops.FcdWeightPack()

    def _params(self):
        out = []
        for name in ops.FCD_LAYERS:
            conv = getattr(self, name)
            out += [conv.weight, conv.bias]
        return out

    def forward(self, x, from_logits=False, return_saved=False, up_size=None):
        """x: (N, num_classes, H, W) fp32 -> (N, 1, H/32, W/32) logits.  ``from_logits=True`` fuses the
        channel softmax the training script applies before calling D (train...:617-618).
        ``up_size=(H, W)`` (with ``from_logits``): x are the low-res logits of the segmentation heads and the
        bilinear upsample to the input resolution is fused in as well (nothing full-res is materialised).
        ``return_saved=True`` also returns a handle for :meth:`replay`."""
        return ops.fcd_forward(x, self._params(), self._pack, x_is_logits=from_logits, return_saved=return_saved,
                               up_size=up_size)

    def replay(self, saved):
        """D(x) for the same x and the same (unchanged) weights as the forward that produced ``saved``:
        re-attaches that output to the parameters so a second loss can back-propagate into them without a
        second forward (train...:665-666 repeats :617-618 bit for bit)."""
        return ops.fcd_replay(saved, self._params(), self._pack)
