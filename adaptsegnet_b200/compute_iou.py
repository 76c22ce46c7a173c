"""Confusion matrix / IoU behind the reference's ``compute_iou.py`` surface (fast_hist, per_class_iu)."""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def fast_hist(a, b, n, mapping=None):
    """compute_iou.py:15-17.  numpy in -> numpy int64 (n, n) out, like the reference; CUDA tensors in ->
    CUDA int64 tensor out (no host round trip).  The counting runs on the device either way.
    ``mapping`` (rows of (source id, train id), the reference's ``info['label2train']``): `a` holds raw dataset ids and
    label_mapping (compute_iou.py:24-28, applied at :55) happens inside the same kernel."""
    as_numpy = not isinstance(a, torch.Tensor)
    if as_numpy:
        a = np.ascontiguousarray(a)
        b = np.ascontiguousarray(b)
        if a.dtype not in (np.uint8, np.int32, np.int64):
            a = a.astype(np.int64)
        lab = torch.from_numpy(a.reshape(-1)).cuda(non_blocking=True)
        prd = torch.from_numpy(b.astype(np.uint8, copy=False).reshape(-1)).cuda(non_blocking=True)
    else:
        lab, prd = a.reshape(-1), b.reshape(-1)
    lut = ops.mapping_lut(mapping, lab.device) if mapping is not None else None
    hist, overflow = ops.fast_hist(lab, prd, int(n), lut=lut)
    if as_numpy:
        if int(overflow.item()):
            # np.bincount would return more than n*n bins and the reference's reshape raises
            raise ValueError(f"cannot reshape array into shape ({n},{n})")
        return hist.cpu().numpy()
    return hist


def label_mapping(input, mapping):
    """compute_iou.py:24-28 as one table lookup (the reference makes one pass over the image per mapping row).
    numpy in -> numpy int64 out; CUDA tensor in -> CUDA int64 tensor out."""
    as_numpy = not isinstance(input, torch.Tensor)
    t = torch.from_numpy(np.ascontiguousarray(input)).cuda() if as_numpy else input
    t = t.to(torch.int64)
    lut = ops.mapping_lut(mapping, t.device).to(torch.int64)
    in_range = (t >= 0) & (t < 256)
    out = torch.where(in_range, lut[t.clamp(0, 255)], t)
    for src, dst in mapping:          # ids an 8-bit label image cannot hold (e.g. Cityscapes' -1): mapped one by one
        if not 0 <= int(src) < 256:
            out = torch.where(t == int(src), torch.full_like(out, int(dst)), out)
    return out.cpu().numpy() if as_numpy else out


def per_class_iu(hist):
    """compute_iou.py:20-21 (float64, 0/0 -> nan)."""
    if isinstance(hist, torch.Tensor):
        hist = hist.cpu().numpy()
    hist = np.asarray(hist, dtype=np.float64)
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.diag(hist) / (hist.sum(1) + hist.sum(0) - np.diag(hist))
