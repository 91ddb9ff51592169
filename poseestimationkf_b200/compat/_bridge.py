"""numpy <-> device plumbing shared by the compat modules (batch of 1 like the reference, or a
leading batch dimension)."""
from __future__ import annotations

import numpy as np
import torch

from poseestimationkf_b200 import _lib


def device():
    if not torch.cuda.is_available():
        raise _lib.PosekfError("poseestimationkf_b200 needs a CUDA device (B200); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def to_dev(a, per_item_shape):
    """array-like with shape per_item_shape or [N, *per_item_shape] -> ([k,N] float32 CUDA tensor, batched?)"""
    arr = np.asarray(a, dtype=np.float64)
    batched = arr.ndim == len(per_item_shape) + 1
    if not batched:
        arr = arr[None]
    if arr.shape[1:] != tuple(per_item_shape):
        raise ValueError(f"expected shape {per_item_shape} (optionally with a leading batch dim), got {arr.shape}")
    flat = arr.reshape(arr.shape[0], -1).T
    return torch.from_numpy(np.ascontiguousarray(flat, dtype=np.float32)).to(device()), batched


def to_host(t, per_item_shape, batched):
    arr = t.detach().cpu().numpy().astype(np.float64).T
    arr = arr.reshape(arr.shape[0], *per_item_shape)
    return arr if batched else arr[0]
