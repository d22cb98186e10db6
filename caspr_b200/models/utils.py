"""Gaussian helpers (mirror of reference caspr/models/utils.py:10-29)."""
from math import log, pi

import torch


def standard_normal_logprob(z):
    return -0.5 * log(2 * pi) - z.pow(2) / 2


def truncated_normal(tensor, mean=0, std=1, trunc_std=2):
    """utils.py:15-22: four candidates per value, first one inside the truncation range."""
    size = tensor.shape
    tmp = tensor.new_empty(size + (4,)).normal_()
    valid = (tmp < trunc_std) & (tmp > -trunc_std)
    ind = valid.max(-1, keepdim=True)[1]
    tensor.data.copy_(tmp.gather(-1, ind).squeeze(-1))
    tensor.data.mul_(std).add_(mean)
    return tensor


def sample_gaussian(size, truncate_std=None, device=None):
    """utils.py:24-29: drawn with the CPU generator, then moved (parity with the reference's RNG stream).

    The draw goes into pinned memory (`empty(pin_memory=True).normal_()` consumes the CPU generator exactly like
    `torch.randn`) so that the host-to-device copy is asynchronous."""
    dev = torch.device(device) if device is not None else None
    if dev is not None and dev.type == 'cuda':
        y = torch.empty(*size, dtype=torch.float32, pin_memory=True).normal_()
        y = y.to(dev, non_blocking=True)
    else:
        y = torch.randn(*size).float()
        y = y if dev is None else y.to(dev)
    if truncate_std is not None:
        truncated_normal(y, mean=0, std=1, trunc_std=truncate_std)
    return y


def sphere_surface_points(num_points, radius=0.5):
    """Random directions from a uniform cube scaled onto a sphere of `radius`
    (reference utils/transform_utils.py:80-85; reached only through decode(sample_contours=...))."""
    import numpy as np
    cube = np.random.uniform(low=-1.0, high=1.0, size=(num_points, 3))
    return cube / np.linalg.norm(cube, axis=1).reshape((-1, 1)) * radius
