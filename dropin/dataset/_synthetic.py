"""Deterministic synthetic samples with the shapes / dtypes / value ranges of the real loaders' items."""
import numpy as np

ID_TO_TRAINID = {7: 0, 8: 1, 11: 2, 12: 3, 13: 4, 17: 5, 19: 6, 20: 7, 21: 8, 22: 9, 23: 10, 24: 11, 25: 12, 26: 13,
                 27: 14, 28: 15, 31: 16, 32: 17, 33: 18}   # GTA5 / Cityscapes id -> train id (dataset/gta5_dataset.py:28-30)


def trainid_lut(ignore_label=255):
    lut = np.full(256, ignore_label, dtype=np.float32)
    for k, v in ID_TO_TRAINID.items():
        lut[k] = v
    return lut


def image(index, crop_size, mean):
    """float32 (3, H, W), BGR, mean-subtracted: what `image[:, :, ::-1] - mean` transposed gives for 8-bit pixels"""
    w, h = int(crop_size[0]), int(crop_size[1])
    rng = np.random.RandomState(1338 + int(index))
    img = rng.randint(0, 256, size=(h, w, 3)).astype(np.float32)
    img -= np.asarray(mean, dtype=np.float32).reshape(1, 1, 3)
    return np.ascontiguousarray(img.transpose(2, 0, 1))


def label(index, crop_size, ignore_label=255):
    """float32 (H, W) of train ids in {0..18, ignore}: blocky raw ids pushed through the id -> train id table"""
    w, h = int(crop_size[0]), int(crop_size[1])
    rng = np.random.RandomState(7331 + int(index))
    bh, bw = max(1, h // 16), max(1, w // 16)
    raw = rng.randint(0, 34, size=(bh, bw)).astype(np.uint8)
    raw = np.repeat(np.repeat(raw, -(-h // bh), axis=0), -(-w // bw), axis=1)[:h, :w]
    return trainid_lut(ignore_label)[raw]
