"""``torch_scatter``-compatible front end of the native segmented reduce.

Mirrors the slice of the third-party API the reference's hot path uses
(`/root/reference/ptgnn/neuralmodels/gnn/messagepassing/abstractmessagepassing.py:4,44-50`):
``scatter(src, index, dim=0, dim_size=N, reduce=...)`` with a 2-D fp32 ``src`` and a 1-D int64 ``index``, plus
``scatter_sum/mean/max/min`` (max/min return ``(out, arg)`` with torch_scatter's conventions: empty rows are 0 with
``arg == src.size(0)``, ties go to the first occurrence).  CUDA only; no fallback.
"""
from typing import Optional, Tuple

import torch

from . import _native as N


def _run(src: torch.Tensor, index: torch.Tensor, dim: int, dim_size: Optional[int], reduce: str, want_arg: bool):
    if src.dim() != 2 or dim not in (0, -2) or index.dim() != 1:
        raise NotImplementedError("native scatter supports src [E, D] reduced along dim=0 with a 1-D index")
    src = N.require_cuda(src, "src", torch.float32)
    index = N.require_cuda(index, "index", torch.int64)
    E, D = src.shape
    if index.shape[0] != E:
        raise ValueError("index and src disagree on the number of rows")
    if dim_size is None:
        dim_size = int(index.max().item()) + 1 if E else 0
    lib = N.lib()
    out = torch.empty(dim_size, D, dtype=torch.float32, device=src.device)
    arg = torch.empty(dim_size, D, dtype=torch.int64, device=src.device) if want_arg else None
    ws_bytes = lib.ptgnn_b200_scatter_workspace_bytes(dim_size, E)
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=src.device)
    _poll_status()
    status = torch.zeros(1, dtype=torch.int32).pin_memory()     # written by the kernels, polled without synchronising
    with torch.cuda.device(src.device):
        rc = lib.ptgnn_b200_scatter_f32(N.ptr(src), N.ptr(index), E, D, dim_size, N.REDUCE[reduce], N.ptr(out), N.ptr(arg),
                                        status.data_ptr(), N.ptr(ws), ws_bytes, N.current_stream(src.device))
    N.check(rc, "ptgnn_b200_scatter_f32")
    _PENDING.append((status, dim_size))
    return out, arg


# Out-of-range indices are reported by the kernels through a pinned status word; the reference (torch_scatter) raises an
# IndexError / device assert in that case.  Checked without synchronising: at the next scatter call, or by check_scatter_status().
_PENDING: list = []


def _poll_status() -> None:
    keep = []
    try:
        for status, n in _PENDING:
            bad = int(status[0])
            if bad:
                raise IndexError(f"scatter: {bad} indices outside [0, {n}); such rows were routed to row 0, the result is not valid")
        keep = _PENDING[-8:]
    finally:
        _PENDING[:] = keep


def check_scatter_status(device=None) -> None:
    """Synchronises and raises IndexError if any scatter call issued so far saw an out-of-range index."""
    torch.cuda.synchronize(device)
    _poll_status()


def scatter(src, index, dim: int = -1, out=None, dim_size: Optional[int] = None, reduce: str = "sum") -> torch.Tensor:
    if out is not None:
        raise NotImplementedError("scatter(out=...) is not used by the reference hot path")
    if reduce not in N.REDUCE:
        raise ValueError(f"unknown reduce {reduce!r}")
    return _run(src, index, dim, dim_size, reduce, False)[0]


def scatter_sum(src, index, dim: int = -1, out=None, dim_size: Optional[int] = None) -> torch.Tensor:
    return scatter(src, index, dim, out, dim_size, "sum")


scatter_add = scatter_sum


def scatter_mean(src, index, dim: int = -1, out=None, dim_size: Optional[int] = None) -> torch.Tensor:
    return scatter(src, index, dim, out, dim_size, "mean")


def scatter_max(src, index, dim: int = -1, out=None, dim_size: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    assert out is None
    return _run(src, index, dim, dim_size, "max", True)


def scatter_min(src, index, dim: int = -1, out=None, dim_size: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    assert out is None
    return _run(src, index, dim, dim_size, "min", True)
