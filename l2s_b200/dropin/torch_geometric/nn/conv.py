"""torch_geometric.nn.conv subset: MessagePassing (gather + scatter), GINConv, GATConv with the parameter names of
PyG 1.7.2 (so the reference's saved state dicts load: `...GIN_layers.k.eps`, `...nn.0.weight`, `...lin_l.weight`,
`...lin_r.weight`, `...att_l`, `...att_r`, `...bias`)."""
import torch
import torch.nn.functional as F
from torch import nn

from ..utils import add_self_loops, remove_self_loops, softmax

_REDUCE = {"add": "sum", "sum": "sum", "mean": "mean", "max": "amax"}


class MessagePassing(nn.Module):
    """`propagate(edge_index, x=..., size=None)` = message (x_j) -> aggregate (aggr at the target) -> update."""

    def __init__(self, aggr="add", flow="source_to_target", node_dim=-2, **kwargs):
        super().__init__()
        assert aggr in _REDUCE and flow in ("source_to_target", "target_to_source")
        self.aggr, self.flow, self.node_dim = aggr, flow, node_dim

    def message(self, x_j):
        return x_j

    def update(self, aggr_out):
        return aggr_out

    def aggregate(self, inputs, index, dim_size):
        idx = index.view(-1, *([1] * (inputs.dim() - 1))).expand_as(inputs)
        out = torch.zeros((dim_size,) + tuple(inputs.shape[1:]), dtype=inputs.dtype, device=inputs.device)
        # empty neighbourhoods keep the zero fill (torch_scatter semantics); non-empty ones ignore it
        return out.scatter_reduce(0, idx, inputs, _REDUCE[self.aggr], include_self=False)

    def propagate(self, edge_index, size=None, **kwargs):
        x = kwargs.get("x")
        xs = x[0] if isinstance(x, (tuple, list)) else x
        j, i = (0, 1) if self.flow == "source_to_target" else (1, 0)
        n_target = xs.shape[0] if size is None or size[1] is None else size[1]
        msg = self.message(xs.index_select(0, edge_index[j]))
        return self.update(self.aggregate(msg, edge_index[i], n_target))


class GINConv(MessagePassing):
    def __init__(self, nn, eps=0.0, train_eps=False, **kwargs):   # noqa: A002 (PyG's own argument name)
        kwargs.setdefault("aggr", "add")
        super().__init__(**kwargs)
        self.nn = nn
        self.initial_eps = eps
        if train_eps:
            self.eps = torch.nn.Parameter(torch.Tensor([eps]))
        else:
            self.register_buffer("eps", torch.Tensor([eps]))

    def forward(self, x, edge_index, size=None):
        out = self.propagate(edge_index, x=(x, x), size=size)
        return self.nn(out + (1 + self.eps) * x)


class GATConv(MessagePassing):
    def __init__(self, in_channels, out_channels, heads=1, concat=True, negative_slope=0.2, dropout=0.0,
                 add_self_loops=True, bias=True, **kwargs):   # noqa: A002
        kwargs.setdefault("aggr", "add")
        super().__init__(node_dim=0, **kwargs)
        self.in_channels, self.out_channels, self.heads, self.concat = in_channels, out_channels, heads, concat
        self.negative_slope, self.dropout, self.add_self_loops = negative_slope, dropout, add_self_loops
        self.lin_l = nn.Linear(in_channels, heads * out_channels, bias=False)
        self.lin_r = self.lin_l                       # one projection for sources and targets (int in_channels)
        self.att_l = nn.Parameter(torch.empty(1, heads, out_channels))
        self.att_r = nn.Parameter(torch.empty(1, heads, out_channels))
        if bias:
            self.bias = nn.Parameter(torch.zeros(heads * out_channels if concat else out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.xavier_uniform_(self.lin_l.weight)
        nn.init.xavier_uniform_(self.att_l)
        nn.init.xavier_uniform_(self.att_r)
        if self.bias is not None:
            nn.init.zeros_(self.bias)

    def forward(self, x, edge_index, size=None, return_attention_weights=None):
        H, Cc, N = self.heads, self.out_channels, x.shape[0]
        h = self.lin_l(x).view(N, H, Cc)
        a_l, a_r = (h * self.att_l).sum(-1), (h * self.att_r).sum(-1)
        if self.add_self_loops:
            edge_index, _ = remove_self_loops(edge_index)
            edge_index, _ = add_self_loops(edge_index, num_nodes=N)
        src, dst = (edge_index[0], edge_index[1]) if self.flow == "source_to_target" else (edge_index[1], edge_index[0])
        e = F.leaky_relu(a_l.index_select(0, src) + a_r.index_select(0, dst), self.negative_slope)
        alpha = F.dropout(softmax(e, dst, num_nodes=N), p=self.dropout, training=self.training)
        out = torch.zeros((N, H, Cc), dtype=x.dtype, device=x.device)
        out.index_add_(0, dst, h.index_select(0, src) * alpha.unsqueeze(-1))
        out = out.reshape(N, H * Cc) if self.concat else out.mean(dim=1)
        if self.bias is not None:
            out = out + self.bias
        if return_attention_weights:
            return out, (edge_index, alpha)
        return out
