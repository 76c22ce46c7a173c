"""Evaluation hot path (BASELINE config 5): evaluate_cityscapes.py:155-169 + compute_iou.py:50-61 without the
159 MB device->host copy, the numpy argmax and the PNG round trip."""
from __future__ import annotations

import torch

from . import ops


@torch.no_grad()
def predict_labels(model, image: torch.Tensor, size=(1024, 2048)) -> torch.Tensor:
    """uint8 (N, size[0], size[1]) class ids == np.argmax(interp(model(image)[1]), channel) of the reference
    (evaluate_cityscapes.py:162-169).  ``model(image)[1]`` is already upsampled to the image size inside the fork's
    forward (model/deeplab_multi.py:188-189), so the chain has TWO bilinear stages -- low-res logits -> image size ->
    ``size`` -- and both are evaluated (fused, with ATen's float op order) by asn_upsample2_argmax_u8."""
    _, x2 = model.low_res_logits(image)
    return ops.upsample2_argmax(x2, tuple(image.shape[-2:]), size)


class ConfusionMatrix:
    """Running 19x19 int64 confusion matrix on the device: `hist += fast_hist(label, pred, n)` of
    compute_iou.py:57 with the accumulation done by the kernel; `all_reduce` sums replicas exactly."""

    def __init__(self, n_cls=19, device="cuda"):
        self.n = n_cls
        self.hist = torch.zeros((n_cls, n_cls), dtype=torch.int64, device=device)
        self.overflow = torch.zeros(1, dtype=torch.int64, device=device)

    def update(self, label: torch.Tensor, pred: torch.Tensor, lut: torch.Tensor | None = None):
        """``lut`` (ops.mapping_lut): raw dataset ids are mapped to train ids inside the kernel
        (label_mapping + fast_hist of compute_iou.py:55-57 in one pass over the frame)."""
        _, ovf = ops.fast_hist(label.reshape(-1), pred.reshape(-1), self.n, hist=self.hist, lut=lut)
        self.overflow += ovf

    def all_reduce(self, group=None):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.hist, op=dist.ReduceOp.SUM, group=group)
            dist.all_reduce(self.overflow, op=dist.ReduceOp.SUM, group=group)

    def per_class_iu(self):
        from .compute_iou import per_class_iu
        if int(self.overflow.item()):
            raise ValueError("predictions outside [0, n_cls) under valid labels (the reference's reshape raises)")
        return per_class_iu(self.hist)
