"""``torch_scatter.composite``: per-segment softmax family on top of the native segment max / sum kernels (same formulas
as torch_scatter 2.0.x's Python composites: subtract the per-segment max, exponentiate, normalise by the per-segment sum)."""
from typing import Optional

import torch


def _expand(per_segment: torch.Tensor, index: torch.Tensor, dim: int) -> torch.Tensor:
    return per_segment.index_select(dim, index)


def _norm_dim(src: torch.Tensor, dim: int) -> int:
    return dim + src.dim() if dim < 0 else dim


def scatter_logsumexp(src: torch.Tensor, index: torch.Tensor, dim: int = -1, out: Optional[torch.Tensor] = None,
                      dim_size: Optional[int] = None, eps: float = 1e-12) -> torch.Tensor:
    from . import scatter_max, scatter_sum, _dim_size

    assert out is None
    if not torch.is_floating_point(src):
        raise ValueError("`scatter_logsumexp` can only be computed over tensors with floating point data types.")
    dim = _norm_dim(src, dim)
    n = _dim_size(index, dim_size)
    max_per = scatter_max(src, index, dim=dim, dim_size=n)[0]
    shifted = src - _expand(max_per, index, dim)
    sums = scatter_sum(shifted.exp(), index, dim=dim, dim_size=n)
    return (sums + eps).log() + max_per


def scatter_log_softmax(src: torch.Tensor, index: torch.Tensor, dim: int = -1, eps: float = 1e-12,
                        dim_size: Optional[int] = None) -> torch.Tensor:
    from . import scatter_max, scatter_sum, _dim_size

    if not torch.is_floating_point(src):
        raise ValueError("`scatter_log_softmax` can only be computed over tensors with floating point data types.")
    dim = _norm_dim(src, dim)
    n = _dim_size(index, dim_size)
    max_per = scatter_max(src, index, dim=dim, dim_size=n)[0]
    shifted = src - _expand(max_per, index, dim)
    sums = scatter_sum(shifted.exp(), index, dim=dim, dim_size=n)
    return shifted - _expand((sums + eps).log(), index, dim)


def scatter_softmax(src: torch.Tensor, index: torch.Tensor, dim: int = -1, dim_size: Optional[int] = None) -> torch.Tensor:
    from . import scatter_max, scatter_sum, _dim_size

    if not torch.is_floating_point(src):
        raise ValueError("`scatter_softmax` can only be computed over tensors with floating point data types.")
    dim = _norm_dim(src, dim)
    n = _dim_size(index, dim_size)
    max_per = scatter_max(src, index, dim=dim, dim_size=n)[0]
    e = (src - _expand(max_per, index, dim)).exp()
    return e / _expand(scatter_sum(e, index, dim=dim, dim_size=n), index, dim)


def scatter_std(src: torch.Tensor, index: torch.Tensor, dim: int = -1, out: Optional[torch.Tensor] = None,
                dim_size: Optional[int] = None, unbiased: bool = True) -> torch.Tensor:
    from . import scatter_sum, _dim_size

    assert out is None
    dim = _norm_dim(src, dim)
    n = _dim_size(index, dim_size)
    count = scatter_sum(torch.ones_like(src), index, dim=dim, dim_size=n)
    mean = scatter_sum(src, index, dim=dim, dim_size=n) / count.clamp(min=1)
    var = scatter_sum((src - _expand(mean, index, dim)) ** 2, index, dim=dim, dim_size=n)
    denom = (count - 1 if unbiased else count).clamp(min=1)
    return (var / denom).sqrt()
