"""basedet/layers/common/function.py:12-54."""
import numpy as np
import torch

from .. import ops

__all__ = ["is_empty_tensor", "non_zeros", "permute_to_N_Any_K", "safelog", "meshgrid"]


def is_empty_tensor(tensor):
    return tensor.numel() == 0


def non_zeros(tensor):
    """F.cond_take(tensor != 0, tensor) -> (values, ascending int32 flat indices)."""
    if tensor.dtype == torch.bool:
        vals, idx = ops.cond_take(tensor.float(), tensor)
        return vals.bool(), idx
    return ops.cond_take(tensor)


def permute_to_N_Any_K(tensor, K):
    """(N, C, H, W) -> (N, H, W, C) -> (N, -1, K): pure layout, no arithmetic."""
    assert tensor.ndim == 4
    N = tensor.shape[0]
    return tensor.permute(0, 2, 3, 1).reshape(N, -1, K)


def safelog(tensor, eps=None):
    # loss-side helper (not on the box-op path); kept for API completeness
    if eps is None:
        eps = float(np.finfo(np.float32).tiny)
    return torch.log(torch.clamp(tensor, min=eps))


def meshgrid(x, y):
    assert len(x.shape) == 1
    assert len(y.shape) == 1
    mesh_shape = (y.shape[0], x.shape[0])
    return x.expand(mesh_shape), y.reshape(-1, 1).expand(mesh_shape)
