"""torch_geometric.utils subset: add_self_loops, remove_self_loops, sort_edge_index, softmax, from_networkx."""
import types

import torch


def _num_nodes(edge_index, num_nodes=None):
    if num_nodes is not None:
        return int(num_nodes)
    return int(edge_index.max()) + 1 if edge_index.numel() > 0 else 0


def add_self_loops(edge_index, edge_weight=None, fill_value=1.0, num_nodes=None):
    """Appends one (i, i) pair per node AFTER the given edges; returns (edge_index, edge_weight)."""
    n = _num_nodes(edge_index, num_nodes)
    loop = torch.arange(n, dtype=edge_index.dtype, device=edge_index.device).unsqueeze(0).repeat(2, 1)
    if edge_weight is not None:
        edge_weight = torch.cat([edge_weight, edge_weight.new_full((n,), fill_value)], dim=0)
    return torch.cat([edge_index, loop], dim=1), edge_weight


def remove_self_loops(edge_index, edge_attr=None):
    keep = edge_index[0] != edge_index[1]
    return edge_index[:, keep], (None if edge_attr is None else edge_attr[keep])


def sort_edge_index(edge_index, edge_attr=None, num_nodes=None):
    """Row-major sort by (source, target)."""
    n = _num_nodes(edge_index, num_nodes)
    perm = (edge_index[0] * n + edge_index[1]).argsort()
    return edge_index[:, perm], (None if edge_attr is None else edge_attr[perm])


def softmax(src, index, ptr=None, num_nodes=None):
    """Segment softmax of `src` [E, ...] over the groups given by `index` [E]."""
    n = _num_nodes(index.view(1, -1), num_nodes)
    idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    mx = torch.full((n,) + tuple(src.shape[1:]), -float("inf"), dtype=src.dtype, device=src.device)
    mx = mx.scatter_reduce(0, idx, src, "amax", include_self=True)
    out = (src - mx.index_select(0, index)).exp()
    den = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device).index_add_(0, index, out)
    return out / (den.index_select(0, index) + 1e-16)


def from_networkx(G):
    """Edge list of a networkx graph in G.edges order (conventional_baselines...:174)."""
    edges = list(G.edges)
    ei = torch.tensor(edges, dtype=torch.long).t().contiguous().view(2, -1)
    return types.SimpleNamespace(edge_index=ei, num_nodes=G.number_of_nodes())
