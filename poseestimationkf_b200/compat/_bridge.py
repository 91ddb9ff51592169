"""numpy <-> device plumbing shared by the compat modules (batch of 1 like the reference, or a
leading batch dimension).  Every reference-named call packs all of its inputs into ONE pinned
staging buffer (one H2D copy), launches the kernel(s) on raw pointers into that buffer, and reads
all outputs back with ONE D2H copy -- a per-call latency of a few tens of microseconds instead of
one copy per argument."""
from __future__ import annotations

import numpy as np
import torch

from poseestimationkf_b200 import _lib


def device():
    if not torch.cuda.is_available():
        raise _lib.PosekfError("poseestimationkf_b200 needs a CUDA device (B200); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def as_rows(a, per_item_shape):
    """array-like with shape per_item_shape or [N, *per_item_shape] -> (float32 [k, N] rows, batched?)"""
    arr = np.asarray(a, dtype=np.float64)
    batched = arr.ndim == len(per_item_shape) + 1
    if not batched:
        arr = arr[None]
    if arr.shape[1:] != tuple(per_item_shape):
        raise ValueError(f"expected shape {per_item_shape} (optionally with a leading batch dim), got {arr.shape}")
    return arr.reshape(arr.shape[0], -1).T.astype(np.float32), batched


def from_rows(rows, per_item_shape, batched):
    arr = rows.astype(np.float64).T
    arr = arr.reshape(arr.shape[0], *per_item_shape)
    return arr if batched else arr[0]


class _Staging:
    """Pinned + device staging buffers, grown on demand, one per process."""

    def __init__(self):
        self.cap = 0
        self.h_in = self.d_in = self.h_out = self.d_out = None

    def ensure(self, n_in, n_out):
        need = max(n_in, n_out, 256)
        if need > self.cap:
            cap = 1 << (need - 1).bit_length()
            dev = device()
            self.h_in = torch.empty(cap, dtype=torch.float32).pin_memory()
            self.h_out = torch.empty(cap, dtype=torch.float32).pin_memory()
            self.d_in = torch.empty(cap, dtype=torch.float32, device=dev)
            self.d_out = torch.empty(cap, dtype=torch.float32, device=dev)
            self.np_in = self.h_in.numpy()
            self.np_out = self.h_out.numpy()
            self.ptrs = (self.h_in.data_ptr(), self.d_in.data_ptr(), self.h_out.data_ptr(), self.d_out.data_ptr())
            self.cap = cap


_staging = _Staging()


def call(inputs, out_rows, launch):
    """inputs: list of float32 arrays [k_i, n_i] (n_i = N, or 1-D [k] for data shared by all filters);
    out_rows: list of (k_j) row counts of [k_j, N] outputs; N taken from the first 2-D input.
    launch(in_ptrs, out_ptrs, N, stream) issues the kernel(s) and returns the C status code(s)."""
    N = next(a.shape[1] for a in inputs if a.ndim == 2)
    sizes = [int(a.size) for a in inputs]
    # 4-float (16-byte) alignment of every segment keeps vector loads legal
    offs, pos = [], 0
    for s in sizes:
        offs.append(pos)
        pos += (s + 3) & ~3
    n_in = pos
    out_sizes = [k * N for k in out_rows]
    out_offs, pos = [], 0
    for s in out_sizes:
        out_offs.append(pos)
        pos += (s + 3) & ~3
    n_out = pos
    st = _staging
    st.ensure(n_in, n_out)
    for a, o, s in zip(inputs, offs, sizes):
        st.np_in[o:o + s] = a.reshape(-1)
    lib = _lib.load()
    stream = torch.cuda.current_stream().cuda_stream
    h_in, base_in, h_out, base_out = st.ptrs
    _lib.check(lib.posekf_copy_async(base_in, h_in, 4 * n_in, 1, stream), "posekf_copy_async")
    rc = launch([base_in + 4 * o for o in offs], [base_out + 4 * o for o in out_offs], N, stream)
    for code in (rc if isinstance(rc, (list, tuple)) else [rc]):
        _lib.check(code, "posekf compat call")
    _lib.check(lib.posekf_copy_async(h_out, base_out, 4 * n_out, 0, stream), "posekf_copy_async")
    _lib.check(lib.posekf_stream_sync(stream), "posekf_stream_sync")
    return [st.np_out[o:o + s].reshape(k, N).copy() for o, s, k in zip(out_offs, out_sizes, out_rows)]
