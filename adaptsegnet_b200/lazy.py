"""Lazy "upsampled logits" handles (SURVEY.md section 7 / 8b last row): what lets the UNCHANGED reference scripts reach
the fused Tier-B kernels.

The fork upsamples the heads' logits to the input resolution inside ``ResNetMulti.forward``
(model/deeplab_multi.py:188-189) and the training script then applies torch builtins to the result:

    seg_loss(pred, labels)            nn.CrossEntropyLoss(ignore_index=255)   train_gta2cityscapes_multi.py:599-600
    F.softmax(pred_target)            (implicit dim = 1) -> model_D(...)      :617-618, :645-646, :665-666
    pred.detach()                                                             :642-643, :662-663
    interp(output2)                   nn.Upsample in evaluate_cityscapes.py   :153,163

With ``model.lazy_outputs = True`` the forward returns ``UpsampledLogits`` instead of 70 MB tensors: a ``torch.Tensor``
subclass without storage that remembers the low-res logits (autograd-connected) and the target size.  Through
``__torch_function__`` the calls above are routed to libasn_b200's fused kernels -- upsample+softmax+CE
(asn_upsample_ce_fwd_bwd), upsample+softmax inside the discriminator's input pack (asn_fcd_fwd_lowres) -- and ANY other
use materialises the tensor with the ordinary upsample kernel and carries on, so the scripts cannot tell the
difference except in speed.  Same mathematics either way (tests/test_gpu_lazy_handle.py).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch.utils._pytree import tree_map

from . import ops

_T = torch.Tensor
# metadata queries the storage-less wrapper answers by itself
_META = {_T.size, _T.dim, _T.numel, _T.shape.__get__, _T.dtype.__get__, _T.device.__get__, _T.ndim.__get__,
         _T.is_cuda.__get__, _T.layout.__get__, _T.is_floating_point, _T.is_complex, _T.dim_order, _T.stride,
         _T.is_contiguous, _T.element_size, _T.nelement, _T.ndimension, _T.get_device, _T.is_sparse.__get__,
         _T.is_quantized.__get__, _T.is_meta.__get__, _T.names.__get__, _T.is_leaf.__get__, _T.__len__}


class _Lazy(torch.Tensor):
    """storage-less tensor of the full-resolution shape that knows how to materialise itself"""

    @staticmethod
    def _make(cls, low: torch.Tensor, size):
        n, c = low.shape[0], low.shape[1]
        t = torch.Tensor._make_wrapper_subclass(cls, (n, c, int(size[0]), int(size[1])), dtype=low.dtype,
                                                device=low.device, requires_grad=False)
        t._low, t._size, t._dense = low, (int(size[0]), int(size[1])), None
        return t

    def materialize(self) -> torch.Tensor:
        raise NotImplementedError

    def __repr__(self):
        return f"{type(self).__name__}(low={tuple(self._low.shape)}, size={self._size}, device={self._low.device})"

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        if func in _META:
            with torch._C.DisableTorchFunctionSubclass():
                return func(*args, **kwargs)
        handler = _HANDLERS.get(func)
        if handler is not None:
            out = handler(*args, **kwargs)
            if out is not NotImplemented:
                return out
        # anything else: materialise (once per handle) and carry on with ordinary tensors
        args, kwargs = tree_map(lambda a: a.materialize() if isinstance(a, _Lazy) else a, (args, kwargs))
        return func(*args, **kwargs)

    @classmethod
    def __torch_dispatch__(cls, func, types, args=(), kwargs=None):
        # safety net below the Python API (a handle that reaches an ATen op directly): never compute on the empty wrapper
        args, kwargs = tree_map(lambda a: a.materialize() if isinstance(a, _Lazy) else a, (args, kwargs or {}))
        return func(*args, **kwargs)


class UpsampledLogits(_Lazy):
    """interp(low) for the bilinear, align_corners=True upsample of model/deeplab_multi.py:188-189"""

    @staticmethod
    def make(low, size):
        return _Lazy._make(UpsampledLogits, low, size)

    def materialize(self):
        if self._dense is None:
            self._dense = ops.upsample_bilinear(self._low, self._size)
        return self._dense


class UpsampledSoftmax(_Lazy):
    """softmax(interp(low), dim=1): what the script hands to the discriminators"""

    @staticmethod
    def make(low, size):
        return _Lazy._make(UpsampledSoftmax, low, size)

    def materialize(self):
        if self._dense is None:
            self._dense = ops.softmax_channels(ops.upsample_bilinear(self._low, self._size))
        return self._dense


class UpsampledTwice(_Lazy):
    """interp_2(interp_1(low)): evaluate_cityscapes.py:153,163 applied to the forward's output.  Materialises with two
    real resizes (the stages do not compose); adaptsegnet_b200.evaluate fuses the argmax instead."""

    @staticmethod
    def make(inner: UpsampledLogits, size):
        t = _Lazy._make(UpsampledTwice, inner._low, size)
        t._mid = inner._size
        return t

    def materialize(self):
        if self._dense is None:
            self._dense = ops.upsample_bilinear(ops.upsample_bilinear(self._low, self._mid), self._size)
        return self._dense

    def argmax_u8(self):
        """uint8 class ids, both stages + argmax in one kernel (evaluate_cityscapes.py:168-169 without the D2H of logits)"""
        return ops.upsample2_argmax(self._low, self._mid, self._size)


# ---- handlers: (args as the torch function receives them) -> result or NotImplemented --------------------------------
def _h_cross_entropy(input, target, weight=None, size_average=None, ignore_index=-100, reduce=None, reduction="mean",
                     label_smoothing=0.0):
    if not isinstance(input, UpsampledLogits) or isinstance(target, _Lazy):
        return NotImplemented
    if size_average is not None or reduce is not None:       # legacy flags, as torch resolves them
        reduction = "mean" if (size_average in (None, True) and reduce in (None, True)) else \
            ("sum" if reduce in (None, True) else "none")
    if reduction not in ("mean", "sum") or label_smoothing != 0.0 or target.dtype != torch.int64 or target.dim() != 3 \
            or not (0 <= ignore_index <= 255 or ignore_index == -100):
        return NotImplemented
    return ops.upsample_softmax_cross_entropy(input._low, input._size, target, ignore_label=ignore_index, weight=weight,
                                              size_average=reduction == "mean")


def _h_softmax(input, dim=None, _stacklevel=3, dtype=None):
    if not isinstance(input, UpsampledLogits) or dtype not in (None, torch.float32):
        return NotImplemented
    if dim is None:                      # F.softmax(pred) without dim: legacy rule, dim = 1 for 4-D
        dim = 1
    if dim not in (1, -3):
        return NotImplemented
    return UpsampledSoftmax.make(input._low, input._size)


def _h_detach(self):
    return type(self)._rewrap(self, self._low.detach())


def _rewrap(cls, old, low):
    t = _Lazy._make(cls, low, old._size)
    if isinstance(old, UpsampledTwice):
        t._mid = old._mid
    return t


_Lazy._rewrap = classmethod(_rewrap)


def _h_requires_grad(self):
    return self._low.requires_grad


def _h_grad_fn(self):
    return self._low.grad_fn


def _h_interpolate(input, size=None, scale_factor=None, mode="nearest", align_corners=None,
                   recompute_scale_factor=None, antialias=False):
    if not isinstance(input, UpsampledLogits) or mode != "bilinear" or align_corners is not True or size is None \
            or scale_factor is not None or antialias:
        return NotImplemented
    size = (size, size) if isinstance(size, int) else tuple(int(s) for s in size)
    if len(size) != 2:
        return NotImplemented
    return UpsampledTwice.make(input, size)


_HANDLERS = {
    F.cross_entropy: _h_cross_entropy,
    F.softmax: _h_softmax,
    torch.softmax: lambda input, dim, dtype=None: _h_softmax(input, dim, dtype=dtype),
    _T.softmax: lambda input, dim, dtype=None: _h_softmax(input, dim, dtype=dtype),
    _T.detach: _h_detach,
    _T.data.__get__: _h_detach,
    _T.requires_grad.__get__: _h_requires_grad,
    _T.grad_fn.__get__: _h_grad_fn,
    F.interpolate: _h_interpolate,
}


def discriminator_input(x):
    """(tensor, from_logits, up_size) for FCDiscriminator.forward: unwraps a lazy handle into the fused call"""
    if isinstance(x, UpsampledSoftmax):
        return x._low, True, x._size
    if isinstance(x, _Lazy):
        return x.materialize(), False, None
    return x, False, None
