"""TEST INFRASTRUCTURE ONLY -- import the UNMODIFIED reference (zcaicaros/L2S) in the build container.

This module exists to (a) validate `oracle/l2s_oracle.py` against the real reference, (b) generate
the golden vectors committed under `tests/golden/` (see `oracle/make_golden.py`) and (c) run the reference as the
CPU baseline / live checker on the GPU box from the vendored copy `baseline/_ref/` (`oracle/vendor_ref.py`;
git-ignored, travels with gpurun).  `/root/reference` itself does not exist on the GPU box and is never read there.

The reference needs packages that are absent here (torch_geometric, torch_scatter, ortools, matplotlib)
and uses a networkx API removed in 3.x.  We install tiny in-memory stand-ins for exactly the import
surface `env/environment.py`, `env/message_passing_evl.py` and
`conventional_baselines_with_long-term-mem.py` touch:

* `torch_geometric` = the repo's pure-torch drop-in `l2s_b200/dropin/torch_geometric` (MessagePassing.propagate =
  gather + `scatter_reduce('amax')` into a zero tensor -- empty segment -> 0, which is what torch_scatter 2.0.6
  `scatter_max` returns --, GINConv / GATConv / global_mean_pool for `model/actor.py`, `from_networkx`,
  `Batch.from_data_list` for the greedy baseline);
* `nx.from_numpy_matrix` -> `nx.from_numpy_array` (same row-major edge insertion order);
* `sys.argv` sanitised because `parameters.py:27` parses argv at import.

Nothing here is copied from the reference; it only makes the reference importable.
"""
import importlib.util
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VENDORED_ROOT = os.path.join(_REPO, "baseline", "_ref")      # written by oracle/vendor_ref.py; travels to the GPU box
SHIM_ROOT = os.path.join(_REPO, "l2s_b200", "dropin")        # pure-torch torch_geometric stand-in (product drop-in)


def _pick_root():
    env = os.environ.get("L2S_REFERENCE_ROOT")
    if env:
        return env
    for cand in ("/root/reference", VENDORED_ROOT):
        if os.path.isfile(os.path.join(cand, "env", "environment.py")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _pick_root()


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "env", "environment.py"))


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _install_stubs():
    """matplotlib / ortools are never called on this path: empty modules.  torch_geometric comes from the repo's
    pure-torch drop-in (l2s_b200/dropin/torch_geometric), loaded by path so that the drop-in's own `env` package
    (which lives next to it) never shadows the reference's."""
    import numpy as np
    import networkx as nx

    tg = sys.modules.get("torch_geometric")
    if tg is None or not getattr(tg, "_l2s_shim", False):
        for k in [k for k in sys.modules if k == "torch_geometric" or k.startswith("torch_geometric.")]:
            del sys.modules[k]
        init = os.path.join(SHIM_ROOT, "torch_geometric", "__init__.py")
        spec = importlib.util.spec_from_file_location("torch_geometric", init,
                                                      submodule_search_locations=[os.path.dirname(init)])
        mod = importlib.util.module_from_spec(spec)
        sys.modules["torch_geometric"] = mod
        spec.loader.exec_module(mod)

    if "matplotlib" not in sys.modules:
        mpl = _mod("matplotlib")
        mpl.pyplot = _mod("matplotlib.pyplot")
    if "ortools" not in sys.modules:
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
        import model.actor as ref_actor
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
                   env_module=ref_env, evl_module=ref_evl, baselines_module=ref_bl, Actor=ref_actor.Actor,
                   actor_module=ref_actor, root=REFERENCE_ROOT)
    return types.SimpleNamespace(**_loaded)
