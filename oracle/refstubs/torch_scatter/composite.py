"""TEST INFRASTRUCTURE ONLY -- torch_scatter.composite stand-ins (not on the hot path)."""
import torch


def _seg_max(src, index, dim_size):
    out = torch.full((dim_size,) + tuple(src.shape[1:]), float("-inf"), dtype=src.dtype)
    idx = index.reshape((-1,) + (1,) * (src.dim() - 1)).expand(src.size())
    return out.scatter_reduce_(0, idx, src, "amax", include_self=True), idx


def scatter_logsumexp(src, index, dim: int = -1, out=None, dim_size=None, eps: float = 1e-12):
    assert dim in (0, -src.dim()) or src.dim() == 1
    n = int(dim_size) if dim_size is not None else (int(index.max()) + 1 if index.numel() else 0)
    mx, idx = _seg_max(src, index, n)
    mx = torch.where(torch.isinf(mx), torch.zeros_like(mx), mx)
    s = torch.zeros_like(mx).scatter_add_(0, idx, (src - mx.gather(0, idx)).exp())
    return (s + eps).log() + mx


def scatter_log_softmax(src, index, dim: int = -1, eps: float = 1e-12, dim_size=None):
    assert dim in (0, -src.dim()) or src.dim() == 1
    n = int(dim_size) if dim_size is not None else (int(index.max()) + 1 if index.numel() else 0)
    mx, idx = _seg_max(src, index, n)
    centred = src - mx.gather(0, idx)
    s = torch.zeros_like(mx).scatter_add_(0, idx, centred.exp())
    return centred - (s + eps).log().gather(0, idx)


def scatter_softmax(src, index, dim: int = -1, eps: float = 1e-12, dim_size=None):
    return scatter_log_softmax(src, index, dim, eps, dim_size).exp()
