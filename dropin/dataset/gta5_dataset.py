"""`GTA5DataSet` with the reference's interface (dataset/gta5_dataset.py:13-71): constructor keywords
(root, list_path, max_iters, crop_size, mean, scale, mirror, ignore_label), items (image, label, size, name) with
image float32 (3, H, W) BGR mean-subtracted, label float32 (H, W) of train ids / 255, size = (H, W, 3), name str.
Real files are read when `list_path` exists; otherwise samples are synthetic (see dataset/__init__.py)."""
import os.path as osp

import numpy as np
from torch.utils import data

from . import _synthetic


class GTA5DataSet(data.Dataset):
    def __init__(self, root, list_path, max_iters=None, crop_size=(321, 321), mean=(128, 128, 128), scale=True,
                 mirror=True, ignore_label=255):
        self.root, self.list_path = root, list_path
        self.crop_size, self.mean, self.ignore_label = crop_size, mean, ignore_label
        self.scale, self.is_mirror = scale, mirror
        self.synthetic = not osp.exists(list_path)
        if self.synthetic:
            self.img_ids = ["synthetic_%05d.png" % i for i in range(64)]
        else:
            with open(list_path) as f:
                self.img_ids = [line.strip() for line in f if line.strip()]
        if max_iters is not None:   # the reference repeats the list to cover max_iters items
            self.img_ids = self.img_ids * int(np.ceil(float(max_iters) / len(self.img_ids)))
        self._lut = _synthetic.trainid_lut(ignore_label)

    def __len__(self):
        return len(self.img_ids)

    def __getitem__(self, index):
        name = self.img_ids[index]
        w, h = int(self.crop_size[0]), int(self.crop_size[1])
        if self.synthetic:
            return (_synthetic.image(index % 64, self.crop_size, self.mean), _synthetic.label(index % 64, self.crop_size,
                    self.ignore_label), np.array((h, w, 3)), name)
        from PIL import Image
        img = Image.open(osp.join(self.root, "images/%s" % name)).convert("RGB").resize((w, h), Image.BICUBIC)
        lab = Image.open(osp.join(self.root, "labels/%s" % name)).resize((w, h), Image.NEAREST)
        img = np.asarray(img, np.float32)[:, :, ::-1] - np.asarray(self.mean, np.float32)     # BGR, mean-subtracted
        lab = self._lut[np.asarray(lab, np.uint8)]                                             # one table lookup
        return np.ascontiguousarray(img.transpose(2, 0, 1)), lab.copy(), np.array((h, w, 3)), name
