"""CPU suite (no GPU): the oracle against the golden vectors produced by the unmodified reference, the host
logic, and the C-ABI library's exported symbols."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, golden_names
from oracle import l2s_oracle as O


def _fa_list(pairs, npairs):
    return [[list(map(int, p)) for p in pairs[b, :npairs[b]]] if npairs[b] else [[0, 0]] for b in range(len(npairs))]


@pytest.mark.parametrize("name", golden_names("traj_"))
def test_oracle_replays_reference_trajectory(name):
    g = np.load(os.path.join(GOLDEN, "traj_%s.npz" % name))
    n_j, n_m, high, fea = [int(v) for v in g["meta"]]
    K = g["tape"].shape[1]
    env = O.OracleEnv(n_j, n_m, 1, high, reward_type=str(g["reward_type"]), fea_norm_const=fea)
    state, fa, done = env.reset(g["instances"], mseq=g["mseq0"])
    assert np.array_equal(state[0], g["x_first"]) and np.array_equal(state[2], g["emc_first"])
    assert np.array_equal(state[1], g["epc"]) and np.array_equal(state[3], g["batch"])
    for k in range(K + 1):
        assert [[list(map(int, p)) for p in f] for f in fa] == _fa_list(g["pairs"][k], g["npairs"][k])
        assert np.array_equal(done.reshape(-1), g["done"][k])
        assert np.array_equal(env.current_objs.reshape(-1).astype(np.int32), g["makespan"][k])
        assert np.array_equal(env.incumbent_objs.reshape(-1), g["incumbent"][k])
        assert np.array_equal(np.stack([i.est for i in env.insts]), g["est"][k])
        assert np.array_equal(np.stack([i.lst for i in env.insts]), g["lst"][k])
        if k < K:
            state, reward, fa, done = env.step(g["tape"][:, k].tolist())
            assert np.array_equal(reward.reshape(-1), g["reward"][k])
    assert np.array_equal(state[0], g["x_last"]) and np.array_equal(state[2], g["emc_last"])


@pytest.mark.parametrize("name", golden_names("traj_"))
def test_oracle_initial_solution_equals_reference_without_ties(name):
    g = np.load(os.path.join(GOLDEN, "traj_%s.npz" % name))
    inst = g["instances"]
    checked = 0
    for b in range(inst.shape[0]):
        mseq, ties = O.rules_solver(inst[b, 0], inst[b, 1], str(g["init_type"]), 99)
        assert ties == g["init_ties"][b]
        if ties == 0:
            assert np.array_equal(mseq, g["mseq0"][b])
            checked += 1
    if name in ("10x5", "8x3_consecutive"):
        assert checked == inst.shape[0]


def test_oracle_jacobi_evaluator_equals_reference():
    g = np.load(os.path.join(GOLDEN, "eval_coo.npz"))
    for tag in ("6x4", "10x10", "20x15"):
        n_j, n_m, B = [int(v) for v in g["shape_" + tag]]
        est, lst, ms = O.evaluate_coo_jacobi(g["ei_" + tag], g["dur_" + tag], n_j, n_m)
        assert np.array_equal(est, g["est_" + tag].reshape(-1))
        assert np.array_equal(lst, g["lst_" + tag].reshape(-1))
        assert np.array_equal(ms, g["ms_" + tag].reshape(-1))


@pytest.mark.parametrize("ds,steps", [("ft6x6", 500), ("la10x5", 100), ("tai20x15", 10)])
def test_oracle_greedy_equals_reference(ds, steps):
    """Greedy policy known-answer test: incumbents logged by the reference's Greedy_baselines (and, at 500
    steps, the reference's own stored result array)."""
    g = np.load(os.path.join(GOLDEN, "greedy_%s.npz" % ds))
    inst = g["instances"]
    n_j, n_m = inst.shape[2], inst.shape[3]
    logs = [int(s) for s in g["log_steps"] if s <= steps]
    env = O.OracleEnv(n_j, n_m, 1, 99)
    _, fa, _ = env.reset(inst, mseq=g["mseq0"])
    _, got = O.run_policy(env, fa, "greedy", steps, log_steps=logs)
    assert np.array_equal(got, g["incumbents"][:len(logs)])
    if steps >= 500:
        assert np.array_equal(got[logs.index(500)], g["stored_rows"][0])


def test_n5_pair_rules_and_quirk():
    mch = np.array([1, 1, 1, 2, 2, 3, 3, 3, 3])      # machines of ops 1..9 (path = ops in order)
    path = list(range(1, 10))
    # first block (len 3): only its last pair; block of 2: its pair; last block (len 4): only its first pair
    assert O.n5_pairs(path, mch, None) == [[2, 3], [4, 5], [6, 7]]
    assert O.n5_pairs(path, mch, (4, 5)) == [[2, 3], [6, 7]]
    assert O.n5_pairs([1, 2, 4], mch, None) == [[1, 2]]
    with pytest.raises(IndexError):
        O.n5_pairs([1, 2], mch, None)


def test_random_pick_is_stable():
    assert [O.random_pick(7, i, s, 9) for i, s in ((0, 0), (1, 0), (0, 1), (12345, 499))] == \
        [O.random_pick(7, i, s, 9) for i, s in ((0, 0), (1, 0), (0, 1), (12345, 499))]
    assert all(0 <= O.random_pick(3, i, 5, 4) < 4 for i in range(200))


def test_library_exports_every_declared_symbol():
    """The C-ABI library loads and exports every function include/l2s_b200.h declares (no compute calls)."""
    from l2s_b200 import _capi, build
    path = build.build()
    lib = ctypes.CDLL(path)
    header = open(os.path.join(ROOT, "include", "l2s_b200.h")).read()
    declared = set(re.findall(r"\b(l2s_[a-z0-9_]+)\s*\(", header))
    declared -= {"l2s_handle_s"}
    assert declared, "no declarations found"
    for name in sorted(declared):
        assert hasattr(lib, name), "library does not export %s" % name
    assert declared == set(_capi.PROTOTYPES), "ctypes prototypes out of sync with the header"
    # ... and every ctypes prototype has as many arguments as the C declaration
    code = re.sub(r"/\*.*?\*/", " ", header, flags=re.S)
    for m in re.finditer(r"\b(l2s_[a-z0-9_]+)\s*\(([^()]*)\)\s*;", code):
        name, params = m.group(1), m.group(2).strip()
        n_params = 0 if params in ("", "void") else params.count(",") + 1
        assert n_params == len(_capi.PROTOTYPES[name][1]), "%s: header %d arguments, ctypes %d" % (
            name, n_params, len(_capi.PROTOTYPES[name][1]))
    assert _capi.lib().l2s_version() >= 100
    assert _capi.lib().l2s_status_string(-4).decode().startswith("illegal action")
    assert _capi.lib().l2s_eval_coo_workspace_bytes(20, 15, 8) > 0


def test_product_has_no_cpu_path_and_never_imports_the_oracle():
    import l2s_b200
    src_dir = os.path.dirname(l2s_b200.__file__)
    for dirpath, _, files in os.walk(src_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
    env = l2s_b200.JsspN5(6, 6, 1, 99)
    inst = np.load(os.path.join(GOLDEN, "traj_6x6_ties.npz"))["instances"]
    with pytest.raises(RuntimeError):
        env.reset(inst, "fdd-divide-mwkr", "cpu")


@pytest.mark.parametrize("name", golden_names("traj_"))
def test_c_oracle_replays_reference_trajectory(name):
    """The plain-C restatement (oracle/l2s_oracle.c) against the reference golden vectors."""
    from oracle import c_oracle as CO
    g = np.load(os.path.join(GOLDEN, "traj_%s.npz" % name))
    inst, tape = g["instances"], g["tape"]
    pm = g["pairs"].shape[2]
    env = CO.CEnv(inst, g["mseq0"], pmax=pm)
    for k in range(tape.shape[1] + 1):
        assert np.array_equal(env.makespan.astype(np.int32), g["makespan"][k])
        assert np.array_equal(env.npairs, g["npairs"][k]) and np.array_equal(env.pairs, g["pairs"][k])
        assert np.array_equal(env.est, g["est"][k]) and np.array_equal(env.lst, g["lst"][k])
        if k < tape.shape[1]:
            env.run("replay", 1, tape=tape[:, k:k + 1], write_state=True)
            assert env.status == 0
            assert np.array_equal(env.incumbent, g["incumbent"][k + 1])
    assert np.array_equal(env.x, g["x_last"]) and np.array_equal(env.emc, g["emc_last"])
    for b in range(inst.shape[0]):
        got, ties = CO.rules_solver(inst[b, 0], inst[b, 1], str(g["init_type"]))
        assert ties == g["init_ties"][b] or ties > 0     # C counts tie events per dispatch like the numpy oracle
        if g["init_ties"][b] == 0:
            assert np.array_equal(got, g["mseq0"][b])


@pytest.mark.parametrize("ds", ["ft6x6", "la10x5", "tai20x15"])
def test_c_oracle_greedy_reproduces_reference_stored_results(ds):
    """500 greedy steps == the reference's stored result row (test_results/conventional_results/greedy-policy)."""
    from oracle import c_oracle as CO
    g = np.load(os.path.join(GOLDEN, "greedy_%s.npz" % ds))
    env = CO.CEnv(g["instances"], g["mseq0"])
    _, logs = env.run("greedy", 500, log_steps=[int(s) for s in g["log_steps"]])
    assert np.array_equal(logs, g["incumbents"])
    assert np.array_equal(logs[list(g["log_steps"]).index(500)], g["stored_rows"][0])


@pytest.mark.parametrize("n_j,n_m", [(2, 2), (3, 1), (2, 5), (5, 2), (3, 7), (4, 32), (13, 4), (33, 3)])
def test_c_and_numpy_oracles_agree_on_odd_shapes(n_j, n_m):
    """Edge shapes (single machine, two jobs, 32 machines, more jobs than a warp): both restatements must agree,
    including the reference's IndexError corner (2-op critical path on one machine)."""
    from oracle import c_oracle as CO
    rng = np.random.RandomState(n_j * 37 + n_m)
    B, K = 6, 25
    dur = rng.randint(1, 6, size=(B, n_j, n_m))
    mch = rng.random_sample((B, n_j, n_m)).argsort(axis=2) + 1
    inst = np.stack([dur, mch], axis=1).astype(np.int32)
    mseq = np.stack([O.rules_solver(i[0], i[1])[0] for i in inst])
    for b in range(B):
        assert np.array_equal(CO.rules_solver(inst[b, 0], inst[b, 1])[0], mseq[b])
    env = O.OracleEnv(n_j, n_m, 1, 99)
    try:
        _, fa, _ = env.reset(inst, mseq=mseq)
        taken, _ = O.run_policy(env, fa, "random", K, seed=5)
        err = None
    except IndexError:
        err = -9
    cenv = CO.CEnv(inst, mseq)
    ctaken, _ = cenv.run("random", K, seed=5)
    if err is None and cenv.status == 0:
        assert np.array_equal(taken, ctaken)
        assert np.array_equal(np.array([i.makespan for i in env.insts], dtype=np.float32), cenv.makespan)
    else:
        assert cenv.status == -9 or err == -9      # the IndexError corner must be flagged by at least the C oracle


def test_dropin_module_paths_resolve_to_the_cuda_implementation():
    """`from env.environment import JsspN5, BatchGraph` (test_learned.py:6, n-step_reinforce.py:8) resolves to this
    package once l2s_b200/dropin is on sys.path; signatures mirror the reference's."""
    import importlib
    import inspect
    import sys
    import l2s_b200
    dropin = os.path.join(os.path.dirname(l2s_b200.__file__), "dropin")
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "env" or k.startswith("env.")}
    sys.path.insert(0, dropin)
    try:
        envmod = importlib.import_module("env.environment")
        evl = importlib.import_module("env.message_passing_evl")
        gen = importlib.import_module("env.generateJSP")
        assert envmod.JsspN5 is l2s_b200.JsspN5 and envmod.BatchGraph is l2s_b200.BatchGraph
        assert evl.Evaluator is l2s_b200.Evaluator
        sig = inspect.signature(envmod.JsspN5.__init__)
        assert list(sig.parameters)[1:8] == ["n_job", "n_mch", "low", "high", "reward_type", "fea_norm_const",
                                              "evaluator_type"]
        assert sig.parameters["reward_type"].default == "yaoxin" and sig.parameters["fea_norm_const"].default == 1000
        assert list(inspect.signature(envmod.JsspN5.reset).parameters)[1:5] == ["instances", "init_type", "device", "plot"]
        assert list(inspect.signature(envmod.JsspN5.step).parameters)[1:4] == ["actions", "device", "plot"]
        assert list(inspect.signature(evl.Evaluator.forward).parameters)[1:] == ["edge_index", "duration", "n_j", "n_m"]
        np.random.seed(0)
        t, m = gen.uni_instance_gen(7, 5, 1, 99)
        assert t.shape == (7, 5) and t.min() >= 1 and t.max() <= 98
        assert (np.sort(m, axis=1) == np.arange(1, 6)).all()
        bg = envmod.BatchGraph()
        bg.wrapper(1, 2, 3, 4)
        assert (bg.x, bg.edge_index_pc, bg.edge_index_mc, bg.batch) == (1, 2, 3, 4)
        bg.clean()
        assert bg.x is None
    finally:
        sys.path.remove(dropin)
        for k in [k for k in sys.modules if k == "env" or k.startswith("env.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_python_list_glue_matches_numpy_route():
    """l2s_pylists.so (CPython glue for the reference's list-shaped API) == the numpy conversions it replaces."""
    from l2s_b200 import _capi
    g = _capi.pylists()
    assert g is not None, "l2s_pylists.so should build wherever Python.h and gcc are present"
    rng = np.random.RandomState(0)
    B, pm = 257, 12
    npairs = rng.randint(0, pm + 1, size=B).astype(np.int32)
    pairs = rng.randint(1, 4000, size=(B, pm, 2)).astype(np.uint16)
    rec = np.zeros((B, 4 + pm), np.uint32)      # host records: word 0 = npairs | flags, words 4.. = a0 | a1 << 16
    rec[:, 0] = npairs.astype(np.uint32) | (np.uint32(1) << 16) * (npairs == 0) | rng.randint(0, 2, size=B).astype(np.uint32) << 24
    rec[:, 4:] = pairs[:, :, 0].astype(np.uint32) | (pairs[:, :, 1].astype(np.uint32) << 16)
    fa = g.l2s_build_move_lists(rec.ctypes.data, rec.shape[1], B)
    want = [pairs[b, :npairs[b]].tolist() if npairs[b] else [[0, 0]] for b in range(B)]
    assert fa == want and all(type(p) is list and type(p[0]) is int for f in fa for p in f)
    acts = [f[len(f) // 2] for f in fa]
    out = np.zeros((B, 2), np.int32)
    assert g.l2s_parse_actions(acts, out.ctypes.data, B) == 0 and np.array_equal(out, np.array(acts))
    mixed = [(np.int64(a[0]), np.int32(a[1])) for a in acts]           # tuples of numpy integers are fine too
    assert g.l2s_parse_actions(mixed, out.ctypes.data, B) == 0 and np.array_equal(out, np.array(acts))
    with pytest.raises(AssertionError):
        g.l2s_parse_actions(acts[:-1], out.ctypes.data, B)
    with pytest.raises((ValueError, TypeError)):
        g.l2s_parse_actions([[1, 2, 3]] * B, out.ctypes.data, B)
    with pytest.raises(TypeError):
        g.l2s_parse_actions([[1.5, 2]] * B, out.ctypes.data, B)


def test_plist_initial_orders_match_reference_golden():
    """a-14: `init_type='plist'` (env/environment.py:140-200, 391-393): with the numpy RNG seeded like the reference
    run that produced tests/golden/plist_10x5.npz, the drop-in's host-side order construction yields the reference's
    machine orders (the one RNG draw is the same `np.random.permutation` call)."""
    from l2s_b200 import JsspN5
    g = np.load(os.path.join(GOLDEN, "plist_10x5.npz"))
    n_j, n_m = int(g["meta"][0]), int(g["meta"][1])
    np.random.seed(int(g["seed"]))
    got = JsspN5(n_j, n_m, 1, 99)._plist_orders(g["instances"])
    assert np.array_equal(got, g["mseq0"])
