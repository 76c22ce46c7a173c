"""helpers shared by the -m gpu parity tests"""
import numpy as np
import pytest
import torch

gpu = pytest.mark.gpu


def cuda(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def host(t):
    return t.detach().float().cpu().numpy() if t.is_floating_point() else t.detach().cpu().numpy()


def feature_like(rng, shape, rms=1.5, frac_zero=0.2):
    """post-ReLU-like activations"""
    x = np.abs(rng.standard_normal(shape)).astype(np.float32) * rms
    x[rng.random(shape) < frac_zero] = 0
    return x
