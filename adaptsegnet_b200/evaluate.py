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


class Evaluator:
    """The evaluation hot loop on the device, per frame: evaluate_cityscapes.py:155-169 (model forward, interp, argmax) +
    compute_iou.py:50-57 (label_mapping, hist += fast_hist) without the 159 MB device->host copy, the numpy argmax and
    the PNG round trip between the two scripts.  One frame = the trunk (unchanged PyTorch
    modules), the layer6 head (the reference also computes layer5 and throws it away: `output1` is never used,
    evaluate_cityscapes.py:162-163) and ONE fused kernel for both bilinear stages, the argmax and the confusion matrix.
    The frame is captured as a CUDA graph after the first call.  ``mapping``: rows (dataset id, train id) =
    info['label2train'] when the labels are raw dataset ids (compute_iou.py:40,55)."""

    def __init__(self, model, n_cls=19, size=(1024, 2048), mapping=None, use_cuda_graph=True, channels_last=True,
                 keep_pred=False):
        self.model = model.eval()
        self.n, self.size = int(n_cls), (int(size[0]), int(size[1]))
        dev = next(model.parameters()).device
        self.hist = torch.zeros((self.n, self.n), dtype=torch.int64, device=dev)
        self.overflow = torch.zeros(1, dtype=torch.int64, device=dev)
        self.lut = ops.mapping_lut(mapping, dev) if mapping is not None else None
        self.use_cuda_graph, self.channels_last, self.keep_pred = bool(use_cuda_graph), bool(channels_last), bool(keep_pred)
        if self.channels_last:
            for name in ("conv1", "bn1", "layer1", "layer2", "layer3", "layer4"):
                getattr(self.model, name).to(memory_format=torch.channels_last)
        self._graph = None
        self._static = None
        self.pred = None
        self.frames = 0

    @torch.no_grad()
    def _frame(self, image, label):
        if self.channels_last:
            image = image.contiguous(memory_format=torch.channels_last)
        _, f4 = self.model.trunk(image)
        z = self.model.layer6(f4)
        self.pred = ops.upsample2_argmax_hist(z, tuple(image.shape[-2:]), self.size, label, self.n, self.hist,
                                              self.overflow, lut=self.lut, want_pred=self.keep_pred)

    def step(self, image: torch.Tensor, label: torch.Tensor):
        """image (N,3,h,w) fp32, label (N,H,W) uint8 / int32 / int64 at ``size``; both on the device."""
        if not self.use_cuda_graph:
            self._frame(image, label)
        else:
            if self._graph is None:
                self._static = (image.clone(), label.clone())
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                h0, o0 = self.hist.clone(), self.overflow.clone()
                with torch.cuda.stream(side):
                    for _ in range(2):                       # cuDNN autotuning, lazy kernel attributes
                        self._frame(*self._static)
                torch.cuda.current_stream().wait_stream(side)
                self.hist.copy_(h0)
                self.overflow.copy_(o0)
                self.model.layer6._pack.invalidate()
                self._graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self._graph, stream=side):
                    self._frame(*self._static)
            for dst, src in zip(self._static, (image, label)):
                if dst.data_ptr() != src.data_ptr():
                    dst.copy_(src, non_blocking=True)
            self._graph.replay()
        self.frames += image.shape[0]
        return self.pred

    def all_reduce(self, group=None):
        """frames shard across ranks; one exact int64 SUM of the 19x19 matrix at the end"""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.hist, op=dist.ReduceOp.SUM, group=group)
            dist.all_reduce(self.overflow, op=dist.ReduceOp.SUM, group=group)

    def result(self):
        """(per-class IoU float64 [n], mIoU float64 scalar) as device tensors -- compute_iou.py:61-64"""
        return ops.per_class_iu_device(self.hist)


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

    def per_class_iu_device(self):
        """(iu, mIoU) float64 device tensors (compute_iou.py:20-21,61-64); no host round trip"""
        return ops.per_class_iu_device(self.hist)

    def per_class_iu(self):
        from .compute_iou import per_class_iu
        if int(self.overflow.item()):
            raise ValueError("predictions outside [0, n_cls) under valid labels (the reference's reshape raises)")
        return per_class_iu(self.hist)
