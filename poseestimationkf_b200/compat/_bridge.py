"""numpy <-> device plumbing shared by the compat modules (batch of 1 like the reference, or a
leading batch dimension).  Every reference-named call packs all of its inputs into ONE pinned
staging buffer and launches the kernel(s) on raw pointers.

Small calls (the reference's own use: one filter per call, `main_file.py:38-47`) are ZERO-COPY: pinned host memory is
mapped into the device's address space under unified addressing (same pointer value on both sides), so the kernel
reads its few dozen inputs from the pinned buffer and writes its outputs to another one directly over the host link --
one kernel launch and one stream synchronisation per call, no `cudaMemcpyAsync`.  Large batches stage through device
memory instead (one H2D copy, kernel, one D2H copy): reading megabytes through the link from inside a kernel would be
slower than a bulk copy."""
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
        if arr.shape != tuple(per_item_shape):
            raise ValueError(f"expected shape {per_item_shape} (optionally with a leading batch dim), got {arr.shape}")
        return arr.astype(np.float32).reshape(-1, 1), False      # one filter: the [k, 1] column is the flattened item
    if arr.shape[1:] != tuple(per_item_shape):
        raise ValueError(f"expected shape {per_item_shape} (optionally with a leading batch dim), got {arr.shape}")
    return arr.reshape(arr.shape[0], -1).T.astype(np.float32), batched


def from_rows(rows, per_item_shape, batched):
    if not batched:
        return rows.astype(np.float64).reshape(per_item_shape)
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


ZERO_COPY_MAX_FLOATS = 4096      # in + out floats up to which a call runs on the mapped pinned buffers


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
    zero_copy = n_in + n_out <= ZERO_COPY_MAX_FLOATS
    if zero_copy:        # the kernel works on the mapped pinned buffers themselves
        base_in, base_out = h_in, h_out
    else:
        _lib.check(lib.posekf_copy_async(base_in, h_in, 4 * n_in, 1, stream), "posekf_copy_async")
    rc = launch([base_in + 4 * o for o in offs], [base_out + 4 * o for o in out_offs], N, stream)
    for code in (rc if isinstance(rc, (list, tuple)) else [rc]):
        _lib.check(code, "posekf compat call")
    if not zero_copy:
        _lib.check(lib.posekf_copy_async(h_out, base_out, 4 * n_out, 0, stream), "posekf_copy_async")
    _lib.check(lib.posekf_stream_sync(stream), "posekf_stream_sync")
    return [st.np_out[o:o + s].reshape(k, N).copy() for o, s, k in zip(out_offs, out_sizes, out_rows)]


# ------------------------------------------------------------------------------------------------
# One-filter fast path (what the reference's own loop does, main_file.py:38-47): a call with a FIXED argument layout
# keeps its own small mapped pinned buffers with numpy views bound once, so that a call is a handful of slice
# assignments (numpy converts float64 -> float32 in C), ONE kernel launch on the mapped buffers, ONE stream
# synchronisation and a few float32 -> float64 conversions on the way out.
# ------------------------------------------------------------------------------------------------
def is_single(a, n) -> bool:
    """True for ONE item of n scalars (list / tuple / 1-D array), False for a batch with a leading dimension."""
    if isinstance(a, np.ndarray):
        return a.ndim == 1 and a.shape[0] == n
    try:
        return len(a) == n and not hasattr(a[0], "__len__")
    except TypeError:
        return False


def is_single_matrix(a, rows, cols) -> bool:
    if isinstance(a, np.ndarray):
        return a.shape == (rows, cols)
    try:
        return len(a) == rows and len(a[0]) == cols and not hasattr(a[0][0], "__len__")
    except TypeError:
        return False


class FixedCall:
    """in_shapes / out_shapes: per-argument shapes of ONE filter, e.g. [(3,), (1,), (4,), (4, 4)]."""

    def __init__(self, in_shapes, out_shapes):
        self.in_shapes, self.out_shapes = list(in_shapes), list(out_shapes)
        self.bound = False

    def _bind(self):
        def layout(shapes):
            offs, pos = [], 0
            for sh in shapes:
                offs.append(pos)
                pos += (int(np.prod(sh)) + 3) & ~3          # 16-byte aligned segments
            return offs, max(pos, 4)
        device()                                            # raises without a CUDA device
        in_offs, n_in = layout(self.in_shapes)
        out_offs, n_out = layout(self.out_shapes)
        self.h_in = torch.zeros(n_in, dtype=torch.float32).pin_memory()
        self.h_out = torch.zeros(n_out, dtype=torch.float32).pin_memory()
        np_in, np_out = self.h_in.numpy(), self.h_out.numpy()
        self.inv = [np_in[o:o + int(np.prod(sh))].reshape(sh) for o, sh in zip(in_offs, self.in_shapes)]
        self.outv = [np_out[o:o + int(np.prod(sh))].reshape(sh) for o, sh in zip(out_offs, self.out_shapes)]
        # under unified addressing pinned host memory has the same address on the device
        self.in_ptrs = [self.h_in.data_ptr() + 4 * o for o in in_offs]
        self.out_ptrs = [self.h_out.data_ptr() + 4 * o for o in out_offs]
        self.lib = _lib.load()
        self.bound = True

    def run(self, args, launch):
        """args: one array-like per input slot (None = leave the slot as it is); launch(lib, in_ptrs, out_ptrs, stream)
        returns the C status.  Returns float64 copies of the outputs."""
        if not self.bound:
            self._bind()
        for view, a in zip(self.inv, args):
            if a is not None:
                view[...] = a
        stream = torch.cuda.current_stream().cuda_stream
        rc = launch(self.lib, self.in_ptrs, self.out_ptrs, stream)
        if rc != 0:
            _lib.check(rc, "posekf compat call")
        rc = self.lib.posekf_stream_sync(stream)
        if rc != 0:
            _lib.check(rc, "posekf_stream_sync")
        return [v.astype(np.float64) for v in self.outv]
