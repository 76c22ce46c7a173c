"""`cityscapesDataSet`: ABSENT from the reference checkout; its interface is pinned by the call sites only.
Constructor keywords (root, list_path, max_iters, crop_size, scale, mirror, mean, set) as used at
train_gta2cityscapes_multi.py:333-337,516-522 and evaluate_cityscapes.py:150; items are 3-tuples (image, size, name)
(train...:418,612; evaluate...:158) with image float32 (3, H, W) BGR mean-subtracted at crop_size = (W, H).
Real files (root/leftImg8bit/<set>/<name>) are read when `list_path` exists, synthetic samples otherwise."""
import os
import os.path as osp

import numpy as np
from torch.utils import data

from . import _synthetic


class cityscapesDataSet(data.Dataset):
    def __init__(self, root, list_path, max_iters=None, crop_size=(321, 321), mean=(128, 128, 128), scale=True,
                 mirror=True, ignore_label=255, set="val"):
        self.root, self.list_path, self.set = root, list_path, set
        self.crop_size, self.mean, self.ignore_label = crop_size, mean, ignore_label
        self.scale, self.is_mirror = scale, mirror
        self.synthetic = not osp.exists(list_path)
        if self.synthetic:
            n = int(os.environ.get("ASN_SYNTHETIC_FRAMES", "8"))
            self.img_ids = ["frankfurt/synthetic_%06d_leftImg8bit.png" % i for i in range(n)]
        else:
            with open(list_path) as f:
                self.img_ids = [line.strip() for line in f if line.strip()]
        if max_iters is not None:
            self.img_ids = self.img_ids * int(np.ceil(float(max_iters) / len(self.img_ids)))

    def __len__(self):
        return len(self.img_ids)

    def __getitem__(self, index):
        name = self.img_ids[index]
        w, h = int(self.crop_size[0]), int(self.crop_size[1])
        if self.synthetic:
            return _synthetic.image(1000 + index % 8, self.crop_size, self.mean), np.array((h, w, 3)), name
        from PIL import Image
        img = Image.open(osp.join(self.root, "leftImg8bit/%s/%s" % (self.set, name))).convert("RGB")
        img = np.asarray(img.resize((w, h), Image.BICUBIC), np.float32)[:, :, ::-1] - np.asarray(self.mean, np.float32)
        return np.ascontiguousarray(img.transpose(2, 0, 1)), np.array((h, w, 3)), name
