"""Pure-PyTorch policy network with the reference's architecture (GIN + DGHAN node/graph embeddings, MLP head,
pairwise action scores; paper: "Deep RL guided improvement heuristic for JSSP", reference model/actor.py:16-239)
and the reference's parameter names/shapes, so the shipped `saved_model/*.pth` state dicts load unchanged.

The policy is NOT the optimisation target of this repo (north star: "stays in PyTorch").  It exists because
torch_geometric is not available here and because the reference's `Actor.forward` forces one host round trip
per step (python lists in, `.item()` per instance out).  Two entry points:

  * `Actor.forward(batch_states, feasible_actions)` -- the reference's signature (python lists), for drop-in use;
  * `Actor.forward_device(x, edge_index_pc, edge_index_mc, batch, pairs, npairs)` -- everything stays on the GPU:
    scores are computed only for the feasible pairs (`[B, P]` dot products instead of the reference's
    `[B, n+2, n+2]` bmm), sampled on the device, and returned as an int32 `[B, 2]` action tensor that
    `JsspN5.step_device` consumes.  No `.item()`, no Python loop over instances.

Layer semantics follow torch_geometric 1.7.2 (the version the reference pins, README.md:33-35):
  GINConv(nn, eps=0, aggr='mean'):  out_i = nn( mean_{j in N_in(i)} x_j + (1+eps) x_i )   (N_in includes the
      self loop the actor appends);
  GATConv(in, out, heads, concat): h = lin_l(x) (lin_r shares the weight), e_ij = leaky_relu(a_l.h_j + a_r.h_i, 0.2)
      for edges j->i after remove-then-add self loops, alpha = softmax_j(e_ij), out_i = sum_j alpha_ij h_j (+ bias).
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.distributions.categorical import Categorical


class _AllReduceSum(torch.autograd.Function):
    """Sum over ranks, differentiable: the gradient of a sum over ranks is the sum of the gradients."""

    @staticmethod
    def forward(ctx, t, group):
        import torch.distributed as dist
        ctx.group = group
        out = t.clone()
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
        return out

    @staticmethod
    def backward(ctx, g):
        import torch.distributed as dist
        g = g.contiguous().clone()
        dist.all_reduce(g, op=dist.ReduceOp.SUM, group=ctx.group)
        return g, None


class ShardedBatchNorm1d(nn.BatchNorm1d):
    """BatchNorm1d whose TRAINING-mode statistics are taken over the rows of all ranks (one all-reduce of
    [sum, sum of squares, count] in float64 per forward, one more in backward).  The reference never calls
    `.eval()`, so its GIN layers normalise every forward with the statistics of the whole rollout
    batch (model/actor.py:74,88); with the batch sharded over ranks, per-rank statistics would change the logits.
    Same parameters / buffers / state-dict keys as BatchNorm1d; without an initialised process group (or with one
    rank) it IS BatchNorm1d.  Works with any backend (nccl on GPUs, gloo in the CPU tests)."""
    group = None

    def forward(self, x):
        import torch.distributed as dist
        if not (self.training and dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1):
            return super().forward(x)
        if x.is_cuda:
            # on the GPU: torch's fused SyncBatchNorm function (per-rank statistics kernel, one all-gather of
            # [mean, invstd, count], fused normalisation; its backward all-reduces the two gradient sums) -- a handful
            # of launches instead of ~60 small ones per layer, which matters for a launch-bound policy
            from torch.nn.modules._functions import SyncBatchNorm as _SyncBN
            group = self.group if self.group is not None else dist.group.WORLD
            if self.track_running_stats:
                self.num_batches_tracked.add_(1)
            m = self.momentum if self.momentum is not None else 1.0 / float(self.num_batches_tracked)
            return _SyncBN.apply(x.contiguous(), self.weight, self.bias, self.running_mean, self.running_var, self.eps, m,
                                 group, dist.get_world_size(group))
        xd = x.double()
        stats = torch.cat([xd.sum(0), (xd * xd).sum(0), xd.new_full((1,), float(x.shape[0]))])
        stats = _AllReduceSum.apply(stats, self.group)
        c = x.shape[1]
        cnt = stats[2 * c]
        mean = stats[:c] / cnt
        var = (stats[c:2 * c] / cnt - mean * mean).clamp(min=0.0)          # biased, like BatchNorm's normaliser
        if self.track_running_stats:
            with torch.no_grad():
                self.num_batches_tracked += 1
                m = self.momentum if self.momentum is not None else 1.0 / float(self.num_batches_tracked)
                self.running_mean.mul_(1 - m).add_(mean.to(self.running_mean.dtype), alpha=m)
                self.running_var.mul_(1 - m).add_((var * cnt / (cnt - 1).clamp(min=1)).to(self.running_var.dtype), alpha=m)
        y = (xd - mean) * torch.rsqrt(var + self.eps)
        y = y.to(x.dtype)
        if self.affine:
            y = y * self.weight + self.bias
        return y


def shard_batchnorm(module, group=None):
    """Switch every BatchNorm1d of `module` to cross-rank statistics (in place; state dict unchanged)."""
    for m in module.modules():
        if type(m) is nn.BatchNorm1d:
            m.__class__ = ShardedBatchNorm1d
            m.group = group
    return module


def _with_self_loops(edge_index, num_nodes):
    loop = torch.arange(num_nodes, dtype=edge_index.dtype, device=edge_index.device)
    return torch.cat([edge_index, loop.unsqueeze(0).repeat(2, 1)], dim=1)


class GINLayer(nn.Module):
    """Parameter layout of PyG's GINConv: `eps` buffer + `nn` Sequential(Linear, BatchNorm1d, ReLU, Linear)."""

    def __init__(self, in_dim, hidden_dim):
        super().__init__()
        self.register_buffer("eps", torch.zeros(1))
        self.nn = nn.Sequential(nn.Linear(in_dim, hidden_dim), nn.BatchNorm1d(hidden_dim), nn.ReLU(),
                                nn.Linear(hidden_dim, hidden_dim))

    def forward(self, x, src, dst, inv_deg):
        agg = torch.zeros_like(x).index_add_(0, dst, x.index_select(0, src)) * inv_deg
        return self.nn(agg + (1.0 + self.eps) * x)


class GIN(nn.Module):
    def __init__(self, in_dim, hidden_dim, layer_gin=4):
        super().__init__()
        self.layer_gin = layer_gin
        self.GIN_layers = nn.ModuleList([GINLayer(in_dim if k == 0 else hidden_dim, hidden_dim) for k in range(layer_gin)])

    def forward(self, x, edge_index, num_graphs):
        """edge_index must already contain the self loops.  Returns (sum over layers of node embeddings,
        sum over layers of per-graph mean pools) like the reference (model/actor.py:96-117)."""
        src, dst = edge_index[0], edge_index[1]
        deg = torch.zeros(x.shape[0], device=x.device, dtype=x.dtype).index_add_(0, dst, torch.ones_like(dst, dtype=x.dtype))
        inv_deg = (1.0 / deg.clamp(min=1)).unsqueeze(1)
        h, nodes, graphs = x, 0, 0
        for layer in self.GIN_layers:
            h = layer(h, src, dst, inv_deg)
            nodes = nodes + h
            graphs = graphs + h.reshape(num_graphs, -1, h.shape[-1]).mean(dim=1)   # equal-sized graphs
        return nodes, graphs


class GATLayer(nn.Module):
    """Parameter layout of PyG 1.7.2's GATConv with shared source/target projection."""

    def __init__(self, in_dim, out_dim, heads, concat, dropout):
        super().__init__()
        self.heads, self.out_dim, self.concat, self.dropout = heads, out_dim, concat, dropout
        self.lin_l = nn.Linear(in_dim, heads * out_dim, bias=False)
        self.lin_r = self.lin_l                      # same module: both keys appear in the state dict
        self.att_l = nn.Parameter(torch.empty(1, heads, out_dim))
        self.att_r = nn.Parameter(torch.empty(1, heads, out_dim))
        self.bias = nn.Parameter(torch.zeros(heads * out_dim if concat else out_dim))
        nn.init.xavier_uniform_(self.lin_l.weight)
        nn.init.xavier_uniform_(self.att_l)
        nn.init.xavier_uniform_(self.att_r)

    def forward(self, x, src, dst):
        """src/dst: edges j->i with exactly one self loop per node."""
        N = x.shape[0]
        h = self.lin_l(x).view(N, self.heads, self.out_dim)
        a_l = (h * self.att_l).sum(-1)               # [N, heads]
        a_r = (h * self.att_r).sum(-1)
        e = F.leaky_relu(a_l.index_select(0, src) + a_r.index_select(0, dst), 0.2)
        # softmax over the incoming edges of every target node
        mx = torch.full((N, self.heads), -float("inf"), device=x.device, dtype=x.dtype)
        mx = mx.scatter_reduce(0, dst.unsqueeze(1).expand_as(e), e, "amax", include_self=True)
        ex = torch.exp(e - mx.index_select(0, dst))
        den = torch.zeros((N, self.heads), device=x.device, dtype=x.dtype).index_add_(0, dst, ex)
        alpha = ex / (den.index_select(0, dst) + 1e-16)
        alpha = F.dropout(alpha, p=self.dropout, training=self.training)
        out = torch.zeros((N, self.heads, self.out_dim), device=x.device, dtype=x.dtype)
        out.index_add_(0, dst, h.index_select(0, src) * alpha.unsqueeze(-1))
        out = out.reshape(N, self.heads * self.out_dim) if self.concat else out.mean(dim=1)
        return out + self.bias


class DGHANlayer(nn.Module):
    def __init__(self, in_chnl, out_chnl, dropout, concat, heads=2):
        super().__init__()
        self.dropout = dropout
        self.opsgrp_conv = GATLayer(in_chnl, out_chnl, heads, concat, dropout)
        self.mchgrp_conv = GATLayer(in_chnl, out_chnl, heads, concat, dropout)

    def forward(self, h, pc, mc):
        h_pc = F.elu(self.opsgrp_conv(F.dropout(h, p=self.dropout, training=self.training), pc[0], pc[1]))
        h_mc = F.elu(self.mchgrp_conv(F.dropout(h, p=self.dropout, training=self.training), mc[0], mc[1]))
        return 0.5 * (h_pc + h_mc)


class DGHAN(nn.Module):
    def __init__(self, in_dim, hidden_dim, dropout, layer_dghan=4, heads=2):
        super().__init__()
        self.hidden_dim = hidden_dim
        layers = []
        if layer_dghan == 1:
            layers.append(DGHANlayer(in_dim, hidden_dim, dropout, concat=False, heads=heads))
        else:
            layers.append(DGHANlayer(in_dim, hidden_dim, dropout, concat=True, heads=heads))
            for _ in range(layer_dghan - 2):
                layers.append(DGHANlayer(heads * hidden_dim, hidden_dim, dropout, concat=True, heads=heads))
            layers.append(DGHANlayer(heads * hidden_dim, hidden_dim, dropout, concat=False, heads=1))
        self.DGHAN_layers = nn.ModuleList(layers)

    def forward(self, x, pc_loops, mc_loops, num_graphs):
        h = x
        for layer in self.DGHAN_layers:
            h = layer(h, pc_loops, mc_loops)
        return h, h.reshape(num_graphs, -1, self.hidden_dim).mean(dim=1)


class Actor(nn.Module):
    """Same constructor as the reference's Actor (model/actor.py:121-128)."""

    def __init__(self, in_dim, hidden_dim, embedding_l=4, policy_l=3, embedding_type='gin', heads=4, dropout=0.6):
        super().__init__()
        self.embedding_l, self.policy_l, self.embedding_type = embedding_l, policy_l, embedding_type
        if embedding_type == 'gin':
            self.embedding = GIN(in_dim, hidden_dim, embedding_l)
        elif embedding_type == 'dghan':
            self.embedding = DGHAN(in_dim, hidden_dim, dropout, embedding_l, heads)
        elif embedding_type == 'gin+dghan':
            self.embedding_gin = GIN(in_dim, hidden_dim, embedding_l)
            self.embedding_dghan = DGHAN(in_dim, hidden_dim, dropout, embedding_l, heads)
        else:
            raise Exception('embedding type should be either "gin", "dghan", or "gin+dghan".')
        first_in = hidden_dim * (4 if embedding_type == 'gin+dghan' else 2)
        self.policy = nn.ModuleList()
        for layer in range(max(1, policy_l)):
            self.policy.append(nn.Sequential(nn.Linear(first_in if layer == 0 else hidden_dim, hidden_dim), nn.Tanh(),
                                             nn.Linear(hidden_dim, hidden_dim)))

    # -- embeddings -> per-node action features [B, n1, hidden] ---------------------------------------------
    def node_features(self, x, edge_index_pc, edge_index_mc, num_graphs):
        N = x.shape[0]
        if self.embedding_type in ('gin', 'gin+dghan'):
            all_loops = _with_self_loops(torch.cat([edge_index_pc, edge_index_mc], dim=1), N)
        if self.embedding_type in ('dghan', 'gin+dghan'):
            pc_loops, mc_loops = _with_self_loops(edge_index_pc, N), _with_self_loops(edge_index_mc, N)
        if self.embedding_type == 'gin':
            node, graph = self.embedding(x, all_loops, num_graphs)
        elif self.embedding_type == 'dghan':
            node, graph = self.embedding(x, pc_loops, mc_loops, num_graphs)
        else:
            n1, g1 = self.embedding_gin(x, all_loops, num_graphs)
            n2, g2 = self.embedding_dghan(x, pc_loops, mc_loops, num_graphs)
            node, graph = torch.cat([n1, n2], dim=-1), torch.cat([g1, g2], dim=-1)
        n1_per = N // num_graphs
        h = torch.cat([node, graph.repeat_interleave(n1_per, dim=0)], dim=-1).reshape(num_graphs, n1_per, -1)
        for layer in self.policy:
            h = layer(h)
        return h

    # -- device-tensor path (no host round trip) ---------------------------------------------------------------
    def forward_device(self, x, edge_index_pc, edge_index_mc, batch, pairs, npairs, greedy=False):
        """pairs int32 [B, P, 2] (instance-local node ids), npairs int32 [B].  Returns
        (actions int32 [B, 2] with (0,0) where an instance has no move, log_prob float [B, 1])."""
        B, P = pairs.shape[0], pairs.shape[1]
        h = self.node_features(x, edge_index_pc, edge_index_mc, B)               # [B, n1, H]
        valid = torch.arange(P, device=x.device).unsqueeze(0) < npairs.unsqueeze(1)
        pairs = torch.where(valid.unsqueeze(-1), pairs, torch.zeros_like(pairs))   # rows are defined up to npairs[b]
        idx = pairs.long()
        ha = torch.gather(h, 1, idx[:, :, 0:1].expand(B, P, h.shape[-1]))
        hb = torch.gather(h, 1, idx[:, :, 1:2].expand(B, P, h.shape[-1]))
        score = (ha * hb).sum(-1)                                                # [B, P] = bmm(h, h^T)[b, a0, a1]
        has = npairs > 0
        score = score.masked_fill(~valid, -float("inf"))
        score = torch.where(has.unsqueeze(1), score, torch.zeros_like(score))    # no move: uniform dummy row
        dist = Categorical(logits=score, validate_args=False)   # no host-side validation: keeps the call capturable
        choice = score.argmax(dim=1) if greedy else dist.sample()
        log_prob = torch.where(has, dist.log_prob(choice), torch.zeros_like(choice, dtype=score.dtype)).unsqueeze(1)
        act = torch.gather(pairs, 1, choice.view(B, 1, 1).expand(B, 1, 2)).squeeze(1)
        act = torch.where(has.unsqueeze(1), act, torch.zeros_like(act))
        return act.to(torch.int32).contiguous(), log_prob

    # -- reference signature (python lists) ---------------------------------------------------------------------
    def forward(self, batch_states, feasible_actions):
        """Same call as the reference's Actor.forward (model/actor.py:175-239): `feasible_actions` is the list
        the environment returns; returns (list of [a0, a1], log_prob [B, 1])."""
        dev = batch_states.x.device
        B = len(feasible_actions)
        if hasattr(feasible_actions, "arrays"):          # l2s_b200.environment.MoveLists: no Python lists at all
            pairs, cnt = feasible_actions.arrays()
            pairs = np.ascontiguousarray(pairs[:, :max(1, int(cnt.max()))])
        else:
            P = max(len(f) for f in feasible_actions)
            pairs = np.zeros((B, P, 2), dtype=np.int32)
            cnt = np.zeros(B, dtype=np.int32)
            for i, f in enumerate(feasible_actions):
                if len(f) == 1 and f[0][0] == 0 and f[0][1] == 0:
                    continue
                cnt[i] = len(f)
                pairs[i, :len(f)] = np.asarray(f, dtype=np.int32)
        act, log_prob = self.forward_device(batch_states.x, batch_states.edge_index_pc, batch_states.edge_index_mc,
                                            batch_states.batch, torch.from_numpy(pairs).to(dev),
                                            torch.from_numpy(cnt).to(dev))
        return act.cpu().tolist(), log_prob
