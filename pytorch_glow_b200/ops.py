"""Host-side tensor helpers with the names and semantics of the reference's misc/ops.py.

These are view/shape utilities for callers (trainer, inferer, tests); the hot path does not
go through them -- splits, concatenations and per-sample reductions are folded into the
CUDA kernels' index maps.
"""
import torch


def _dims(dim):
    return sorted([dim] if isinstance(dim, int) else list(dim))


def reduce_mean(tensor, dim=None, keepdim=False, out=None):
    """misc/ops.py:4-37."""
    r = torch.mean(tensor) if dim is None else torch.mean(tensor, dim=_dims(dim), keepdim=keepdim)
    if out is not None:
        out.copy_(r)
    return r


def reduce_sum(tensor, dim=None, keepdim=False, out=None):
    """misc/ops.py:40-73."""
    r = torch.sum(tensor) if dim is None else torch.sum(tensor, dim=_dims(dim), keepdim=keepdim)
    if out is not None:
        out.copy_(r)
    return r


def tensor_equal(a, b, eps=1e-6):
    """misc/ops.py:76-92: max-abs difference within eps."""
    if a.shape != b.shape:
        return False
    return 0 <= float(torch.max(torch.abs(a - b))) <= eps


def split_channel(tensor, split_type='simple'):
    """misc/ops.py:95-113: 'simple' = first/second half, 'cross' = even/odd channels (views)."""
    assert len(tensor.shape) == 4
    assert split_type in ['simple', 'cross']
    nc = tensor.shape[1]
    if split_type == 'simple':
        return tensor[:, :nc // 2, ...], tensor[:, nc // 2:, ...]
    return tensor[:, 0::2, ...], tensor[:, 1::2, ...]


def cat_channel(a, b):
    """misc/ops.py:116-127."""
    return torch.cat((a, b), dim=1)


def count_pixels(tensor):
    """misc/ops.py:130-140."""
    assert len(tensor.shape) == 4
    return int(tensor.shape[2] * tensor.shape[3])


def onehot(y, num_classes):
    """misc/ops.py:143-160."""
    assert len(y.shape) in [1, 2], "Label y should be 1D or 2D vector"
    idx = y.unsqueeze(-1) if len(y.shape) == 1 else y
    return torch.zeros(y.shape[0], num_classes, device=y.device).scatter_(1, idx, 1)
