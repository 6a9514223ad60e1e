"""Pure-torch stand-in for the sliver of torch_geometric 1.7.2 that the reference imports
(README.md:33-35 pins torch-geometric==1.7.2; this image has no torch_geometric, torch_scatter or torch_sparse).

Put `l2s_b200/dropin` on `sys.path` and the reference's own `model/actor.py` (GINConv, GATConv, global_mean_pool,
add_self_loops), `env/message_passing_evl.py` (MessagePassing with max aggregation) and
`conventional_baselines_with_long-term-mem.py` (from_networkx, Batch.from_data_list) import and run UNMODIFIED.
Only the call signatures and semantics those files use are provided; everything is plain `torch`
(index_add_ / scatter_reduce), so it runs on CPU and CUDA tensors alike.

Layer semantics (PyG 1.7.2):
  * MessagePassing(aggr, flow).propagate(edge_index, x=..., size=None): gather x_j along the edges, reduce at the
    target (flow='source_to_target': j = edge_index[0] -> i = edge_index[1]; 'target_to_source' the other way);
    'max' over an empty neighbourhood gives 0 (torch_scatter 2.0.6 scatter_max fills with 0);
  * GINConv(nn, eps, train_eps, aggr='mean'): out_i = nn(aggr_{j in N(i)} x_j + (1 + eps) x_i); `eps` is a buffer;
  * GATConv(in, out, heads, concat, dropout): shared projection lin_l (= lin_r), attention
    e_ij = leaky_relu(att_l . h_j + att_r . h_i, 0.2) over the in-edges of i after remove-then-add self loops,
    softmax per target, dropout on the coefficients, heads concatenated or averaged, + bias;
  * global_mean_pool(x, batch): per-graph mean; add_self_loops(edge_index) appends arange(N) pairs.
"""
__version__ = "1.7.2+l2s_b200.shim"
_l2s_shim = True

from . import typing, utils, nn, data  # noqa: E402,F401
