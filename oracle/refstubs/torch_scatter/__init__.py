"""TEST INFRASTRUCTURE ONLY -- stand-in for the third-party ``torch-scatter`` wheel.

The reference (`/root/reference/setup.py:23`, ``torch-scatter>=2.0.5``) imports
``torch_scatter`` at `ptgnn/neuralmodels/gnn/messagepassing/abstractmessagepassing.py:4`.
The wheel is not installed in this image and cannot be installed (no network), so the
reference cannot be imported as-is.  This stub restates the *published* CPU semantics of
rusty1s/pytorch_scatter 2.0.x with stock torch ops so that the
reference's own layer classes run unchanged on CPU:

* ``sum``/``add``: ``zeros(size).scatter_add_(dim, index, src)`` (edge order per target).
* ``mean``: sum divided by the per-target count clamped to >= 1 (empty -> 0).
* ``max``/``min``: strict ``>``/``<`` update starting from the dtype's lowest/highest value
  (first occurrence wins ties, NaN never wins); untouched outputs are set to 0 and their
  ``arg`` is ``src.size(dim)``.

It is put on ``sys.path`` only by ``oracle/refimport.py`` (golden-vector generation and the
reference-vs-oracle pinning tests).  Nothing in ``ptgnn_b200/`` may import it.
"""
from typing import Optional, Tuple

import torch

from . import composite  # noqa: F401  (torch_scatter.composite.*)
from .composite import scatter_log_softmax, scatter_logsumexp, scatter_softmax  # noqa: F401


def _broadcast(index: torch.Tensor, src: torch.Tensor, dim: int) -> torch.Tensor:
    if dim < 0:
        dim = src.dim() + dim
    if index.dim() == 1:
        for _ in range(dim):
            index = index.unsqueeze(0)
    while index.dim() < src.dim():
        index = index.unsqueeze(-1)
    return index.expand(src.size())


def _out_size(src, index, dim, dim_size):
    size = list(src.size())
    if dim_size is not None:
        size[dim] = int(dim_size)
    elif index.numel() == 0:
        size[dim] = 0
    else:
        size[dim] = int(index.max()) + 1
    return size


def scatter_sum(src, index, dim: int = -1, out: Optional[torch.Tensor] = None, dim_size: Optional[int] = None):
    index = _broadcast(index, src, dim)
    if out is None:
        out = torch.zeros(_out_size(src, index, dim, dim_size), dtype=src.dtype, device=src.device)
    return out.scatter_add_(dim, index, src)


scatter_add = scatter_sum


def scatter_mean(src, index, dim: int = -1, out: Optional[torch.Tensor] = None, dim_size: Optional[int] = None):
    out = scatter_sum(src, index, dim, out, dim_size)
    dim_size = out.size(dim)
    index_dim = dim
    if index_dim < 0:
        index_dim = index_dim + src.dim()
    if index.dim() <= index_dim:
        index_dim = index.dim() - 1
    ones = torch.ones(index.size(), dtype=src.dtype, device=src.device)
    count = scatter_sum(ones, index, index_dim, None, dim_size)
    count[count < 1] = 1
    count = _broadcast(count, out, dim)
    if out.is_floating_point():
        out.true_divide_(count)
    else:
        out.div_(count, rounding_mode="floor")
    return out


def _scatter_extreme(src, index, dim, dim_size, is_max: bool) -> Tuple[torch.Tensor, torch.Tensor]:
    """Sequential strict-compare loop; the algorithm of torch_scatter's scatter_cpu.cpp for MIN/MAX
    restated for a 2-D ``src`` with ``dim == 0`` (general dims are moved to that form)."""
    if dim < 0:
        dim = src.dim() + dim
    index_b = _broadcast(index, src, dim)
    size = _out_size(src, index_b, dim, dim_size)
    src_m = src.movedim(dim, 0).contiguous()
    idx_m = index_b.movedim(dim, 0).contiguous()
    n_src = src_m.shape[0]
    n_inner = 1
    for d in src_m.shape[1:]:
        n_inner *= d
    inner = src_m.reshape(n_src, n_inner)
    idx_inner = idx_m.reshape(n_src, n_inner)
    n_out = size[dim]
    if src.is_floating_point():
        init = torch.finfo(src.dtype).min if is_max else torch.finfo(src.dtype).max
    else:
        init = torch.iinfo(src.dtype).min if is_max else torch.iinfo(src.dtype).max
    out = torch.full((n_out, inner.shape[1]), init, dtype=src.dtype)
    arg = torch.full((n_out, inner.shape[1]), n_src, dtype=torch.int64)
    cols = torch.arange(inner.shape[1])
    for e in range(n_src):
        tgt = idx_inner[e]
        cur = out[tgt, cols]
        better = (inner[e] > cur) if is_max else (inner[e] < cur)
        out[tgt[better], cols[better]] = inner[e][better]
        arg[tgt[better], cols[better]] = e
    out[arg == n_src] = 0
    out_shape = [n_out] + list(src_m.shape[1:])
    return out.reshape(out_shape).movedim(0, dim), arg.reshape(out_shape).movedim(0, dim)


def _scatter_extreme_fast(src, index, dim, dim_size, is_max: bool) -> Tuple[torch.Tensor, torch.Tensor]:
    """Vectorised equivalent of ``_scatter_extreme`` for floating-point ``src`` (used for anything
    larger than a toy): same strict-compare / first-occurrence / empty->0 semantics."""
    if dim < 0:
        dim = src.dim() + dim
    index_b = _broadcast(index, src, dim)
    size = _out_size(src, index_b, dim, dim_size)
    n_src = src.shape[dim]
    init = torch.finfo(src.dtype).min if is_max else torch.finfo(src.dtype).max
    eligible = ~torch.isnan(src) & ((src > init) if is_max else (src < init))
    clean = torch.where(eligible, src, torch.full_like(src, init))
    out = torch.full(size, init, dtype=src.dtype)
    out.scatter_reduce_(dim, index_b, clean, "amax" if is_max else "amin", include_self=True)
    winners = eligible & (clean == out.gather(dim, index_b))
    shape = [1] * src.dim()
    shape[dim] = n_src
    edge_ids = torch.arange(n_src, dtype=torch.int64).reshape(shape).expand(src.size())
    cand = torch.where(winners, edge_ids, torch.full_like(edge_ids, n_src))
    arg = torch.full(size, n_src, dtype=torch.int64)
    arg.scatter_reduce_(dim, index_b, cand, "amin", include_self=True)
    out = torch.where(arg == n_src, torch.zeros_like(out), out)
    return out, arg


def scatter_max(src, index, dim: int = -1, out=None, dim_size: Optional[int] = None):
    assert out is None, "stub: `out=` is not used by the reference"
    if src.numel() <= 4096 or not src.is_floating_point():
        return _scatter_extreme(src, index, dim, dim_size, True)
    return _scatter_extreme_fast(src, index, dim, dim_size, True)


def scatter_min(src, index, dim: int = -1, out=None, dim_size: Optional[int] = None):
    assert out is None, "stub: `out=` is not used by the reference"
    if src.numel() <= 4096 or not src.is_floating_point():
        return _scatter_extreme(src, index, dim, dim_size, False)
    return _scatter_extreme_fast(src, index, dim, dim_size, False)


def scatter(src, index, dim: int = -1, out=None, dim_size: Optional[int] = None, reduce: str = "sum"):
    if reduce in ("sum", "add"):
        return scatter_sum(src, index, dim, out, dim_size)
    if reduce == "mean":
        return scatter_mean(src, index, dim, out, dim_size)
    if reduce == "max":
        return scatter_max(src, index, dim, out, dim_size)[0]
    if reduce == "min":
        return scatter_min(src, index, dim, out, dim_size)[0]
    raise ValueError(reduce)
