"""Graph-level pooling (torch_geometric.nn.glob)."""
import torch


def _size(batch, size):
    return int(batch.max()) + 1 if size is None else int(size)


def global_add_pool(x, batch, size=None):
    return torch.zeros((_size(batch, size), x.shape[-1]), dtype=x.dtype, device=x.device).index_add_(0, batch, x)


def global_mean_pool(x, batch, size=None):
    n = _size(batch, size)
    cnt = torch.zeros(n, dtype=x.dtype, device=x.device).index_add_(0, batch, torch.ones_like(batch, dtype=x.dtype))
    return global_add_pool(x, batch, n) / cnt.clamp(min=1).unsqueeze(-1)


def global_max_pool(x, batch, size=None):
    n = _size(batch, size)
    out = torch.full((n, x.shape[-1]), -float("inf"), dtype=x.dtype, device=x.device)
    return out.scatter_reduce(0, batch.unsqueeze(-1).expand_as(x), x, "amax", include_self=True)
