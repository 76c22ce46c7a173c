"""Synthetic inputs of the shapes and types the reference's dataset classes produce (SURVEY.md section 8d;
dataset/gta5_dataset.py:58-71): mean-subtracted BGR-like float images and blocky int64 label maps with ~10 % ignore
label.  Used by bench.py and the tests (no datasets are available offline)."""
from __future__ import annotations

import torch
import torch.nn.functional as F

IMG_MEAN = (104.00698793, 116.66876762, 122.67891434)  # train_gta2cityscapes_multi.py:30


def synthetic_batch(seed, src_hw, tgt_hw, num_classes=19, device="cpu"):
    """-> (source image (1,3,H,W) fp32, source labels (1,H,W) int64 in {0..num_classes-1, 255}, target image)."""
    g = torch.Generator().manual_seed(seed)
    mean = torch.tensor(IMG_MEAN).view(1, 3, 1, 1)
    src = torch.randint(0, 256, (1, 3) + tuple(src_hw), generator=g).float() - mean
    tgt = torch.randint(0, 256, (1, 3) + tuple(tgt_hw), generator=g).float() - mean
    bh, bw = max(1, src_hw[0] // 16), max(1, src_hw[1] // 16)
    coarse = torch.randint(0, num_classes, (1, 1, bh, bw), generator=g).float()
    lab = F.interpolate(coarse, size=tuple(src_hw), mode="nearest")[0].long()
    ign = F.interpolate((torch.rand((1, 1, bh, bw), generator=g) < 0.1).float(), size=tuple(src_hw), mode="nearest")[0]
    lab[ign > 0] = 255
    return src.to(device), lab.to(device), tgt.to(device)
