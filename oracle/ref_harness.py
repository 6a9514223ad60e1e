"""TEST INFRASTRUCTURE ONLY -- import the UNMODIFIED reference (zcaicaros/L2S) in the build container.

This module exists to (a) validate `oracle/l2s_oracle.py` against the real reference and (b) generate
the golden vectors committed under `tests/golden/` (see `oracle/make_golden.py`).  `/root/reference`
does not exist on the GPU box, so nothing in `-m gpu` tests, `smoke()` or `bench.py` imports this file.

The reference needs packages that are absent here (torch_geometric, torch_scatter, ortools, matplotlib)
and uses a networkx API removed in 3.x.  We install tiny in-memory stand-ins for exactly the import
surface `env/environment.py`, `env/message_passing_evl.py` and
`conventional_baselines_with_long-term-mem.py` touch:

* `torch_geometric.nn.conv.MessagePassing.propagate` = gather + `scatter_reduce('amax')` into a zero
  tensor (empty segment -> 0, which is what torch_scatter 2.0.6 `scatter_max` returns), honouring `flow`;
* `torch_geometric.utils.from_networkx`, `torch_geometric.data.batch.Batch.from_data_list` for the
  greedy baseline (edge list + node-offset concatenation);
* `nx.from_numpy_matrix` -> `nx.from_numpy_array` (same row-major edge insertion order);
* `sys.argv` sanitised because `parameters.py:27` parses argv at import.

Nothing here is copied from the reference; it only makes the reference importable.
"""
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("L2S_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "env", "environment.py"))


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _install_stubs():
    import numpy as np
    import networkx as nx
    import torch

    if "torch_geometric" in sys.modules and getattr(sys.modules["torch_geometric"], "_l2s_stub", False):
        return

    class MessagePassing(torch.nn.Module):
        def __init__(self, aggr="add", flow="source_to_target", **kw):
            super().__init__()
            assert aggr == "max"
            self.aggr, self.flow = aggr, flow

        def propagate(self, edge_index, x=None, size=None):
            xs = x[0] if isinstance(x, tuple) else x
            # source_to_target: messages x_j with j = edge_index[0] are aggregated at i = edge_index[1]
            j, i = (0, 1) if self.flow == "source_to_target" else (1, 0)
            src = xs.index_select(0, edge_index[j])
            idx = edge_index[i].view(-1, *([1] * (src.dim() - 1))).expand_as(src)
            out = torch.zeros_like(xs)
            return out.scatter_reduce(0, idx, src, "amax", include_self=False)

    def from_networkx(G):
        edges = list(G.edges)
        ei = torch.tensor(edges, dtype=torch.long).t().contiguous().view(2, -1)
        return types.SimpleNamespace(edge_index=ei, num_nodes=G.number_of_nodes())

    class Batch:
        @staticmethod
        def from_data_list(datas):
            off, parts = 0, []
            for d in datas:
                parts.append(d.edge_index + off)
                off += d.num_nodes
            return types.SimpleNamespace(edge_index=torch.cat(parts, dim=-1))

    def add_self_loops(edge_index, num_nodes=None):
        n = int(edge_index.max()) + 1 if num_nodes is None else num_nodes
        loop = torch.arange(n, dtype=edge_index.dtype, device=edge_index.device).repeat(2, 1)
        return torch.cat([edge_index, loop], dim=1), None

    def _unavailable(*a, **k):
        raise RuntimeError("stubbed in oracle/ref_harness.py (policy network is out of scope)")

    tg = _mod("torch_geometric", _l2s_stub=True)
    tg.typing = _mod("torch_geometric.typing", OptPairTensor=object, Adj=object, Size=object)
    tg.nn = _mod("torch_geometric.nn", GINConv=_unavailable, GATConv=_unavailable,
                 global_mean_pool=_unavailable, MessagePassing=MessagePassing)
    tg.nn.conv = _mod("torch_geometric.nn.conv", MessagePassing=MessagePassing)
    tg.utils = _mod("torch_geometric.utils", add_self_loops=add_self_loops, sort_edge_index=_unavailable,
                    from_networkx=from_networkx)
    tg.data = _mod("torch_geometric.data")
    tg.data.batch = _mod("torch_geometric.data.batch", Batch=Batch)

    mpl = _mod("matplotlib")
    mpl.pyplot = _mod("matplotlib.pyplot")
    ort = _mod("ortools")
    ort.sat = _mod("ortools.sat")
    ort.sat.python = _mod("ortools.sat.python")
    ort.sat.python.cp_model = _mod("ortools.sat.python.cp_model", CpSolverSolutionCallback=object)

    if not hasattr(nx, "from_numpy_matrix"):
        def from_numpy_matrix(A, parallel_edges=False, create_using=None):
            return nx.from_numpy_array(np.asarray(A), parallel_edges=parallel_edges, create_using=create_using)
        nx.from_numpy_matrix = from_numpy_matrix


_loaded = {}


def load():
    """Returns a namespace with the reference's JsspN5, Evaluator, Greedy_baselines, uni_instance_gen."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    sys.dont_write_bytecode = True
    _install_stubs()
    saved_argv = sys.argv
    sys.argv = [saved_argv[0] if saved_argv else "ref_harness"]
    # our own drop-in package is also called `env`; make sure the reference's wins inside this process
    for k in [k for k in sys.modules if k == "env" or k.startswith("env.") or k in ("parameters", "model", "model.actor")]:
        del sys.modules[k]
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        import env.environment as ref_env
        import env.message_passing_evl as ref_evl
        import env.generateJSP as ref_gen
        import env.permissible_LS as ref_pls
        spec = importlib.util.spec_from_file_location(
            "ref_conventional_baselines",
            os.path.join(REFERENCE_ROOT, "conventional_baselines_with_long-term-mem.py"))
        ref_bl = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref_bl)
    finally:
        sys.argv = saved_argv
        sys.path.remove(REFERENCE_ROOT)
    _loaded.update(JsspN5=ref_env.JsspN5, BatchGraph=ref_env.BatchGraph, Evaluator=ref_evl.Evaluator,
                   CPM_batch_G=ref_evl.CPM_batch_G, uni_instance_gen=ref_gen.uni_instance_gen,
                   Greedy_baselines=ref_bl.Greedy_baselines, permissibleLeftShift=ref_pls.permissibleLeftShift,
                   env_module=ref_env, evl_module=ref_evl, baselines_module=ref_bl)
    return types.SimpleNamespace(**_loaded)
