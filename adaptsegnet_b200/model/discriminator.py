"""Fully-convolutional patch discriminator -- the reference's ``model/discriminator.py`` surface.

Parameters live in the same five ``nn.Conv2d`` holders (state-dict keys ``conv1..conv4,
classifier``), but `forward` is one call into the sm_100a implicit-GEMM kernels (ops.fcd_forward):
4x4 stride-2 convolutions with bias + LeakyReLU(0.2) fused in the epilogue
(reference model/discriminator.py:21-34)."""
from __future__ import annotations

import torch
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
        self._last_lazy = None   # (input key, ops.FcdSaved) of the last forward on a lazy handle that required grad

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
        if type(x) is not torch.Tensor and isinstance(x, torch.Tensor):
            # a lazy handle from the unchanged script's `model_D(F.softmax(pred))` (train...:617-618,645-646,665-666)
            from ..lazy import UpsampledSoftmax, discriminator_input
            if isinstance(x, UpsampledSoftmax) and not from_logits and not return_saved:
                return self._forward_lazy(x)
            x, lazy_logits, lazy_size = discriminator_input(x)
            from_logits, up_size = from_logits or lazy_logits, up_size or lazy_size
        return ops.fcd_forward(x, self._params(), self._pack, x_is_logits=from_logits, return_saved=return_saved,
                               up_size=up_size)

    def _forward_lazy(self, x):
        """D(softmax(interp(low))) for a lazy handle.  The script runs this forward twice per iteration on the target
        prediction with identical weights -- once with D frozen for the generator (:617-618), once on `.detach()` for D
        itself (:665-666): the second call re-attaches the first one's activations (ops.fcd_replay) instead of recomputing
        them, exactly what AdaptSegTrainer does with `reuse_target_forward`."""
        low, size = x._low, x._size
        params = self._params()
        key = (low.data_ptr(), low._version, tuple(low.shape), size, low.device)
        last = self._last_lazy
        if (last is not None and last[0] == key and ops.precision_mode() == "bf16" and torch.is_grad_enabled()
                and not low.requires_grad and all(p.requires_grad for p in params)
                and self._pack.key_on(low.device) == last[1].key == tuple((p.data_ptr(), p._version) for p in params)):
            self._last_lazy = None
            return ops.fcd_replay(last[1], params, self._pack)
        # (a forward on another input in between -- the D step on the source prediction, :645-646 -- leaves the entry alone)
        if low.requires_grad and torch.is_grad_enabled() and ops.precision_mode() == "bf16":
            out, saved = ops.fcd_forward(low, params, self._pack, x_is_logits=True, return_saved=True, up_size=size)
            if saved is not None:
                self._last_lazy = (key, saved)
            return out
        return ops.fcd_forward(low, params, self._pack, x_is_logits=True, up_size=size)

    def replay(self, saved):
        """D(x) for the same x and the same (unchanged) weights as the forward that produced ``saved``:
        re-attaches that output to the parameters so a second loss can back-propagate into them without a
        second forward (train...:665-666 repeats :617-618 bit for bit)."""
        return ops.fcd_replay(saved, self._params(), self._pack)
