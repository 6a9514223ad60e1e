"""CPU parity of the oracle against the LIVE, unmodified reference (imported from /root/reference or the vendored
`baseline/_ref`) on shapes the golden fixtures do not hold: single-machine instances (n_m = 1: every operation is the
first and the last of its job, the machine chain joins consecutive node ids) and very small / very flat shapes.
The GPU suite checks the CUDA path against the same oracle on these shapes (`test_odd_shapes_against_c_oracle`,
`test_evaluator_coo_full_size[4-1-50]`, `test_evaluator_edge_orders_and_malformed_lists[3-1-9]`)."""
import numpy as np
import pytest
import torch

from oracle import l2s_oracle as O
from oracle import ref_harness

pytestmark = pytest.mark.skipif(not ref_harness.available(), reason="reference tree / baseline/_ref not present")


@pytest.mark.parametrize("n_j,n_m", [(3, 1), (4, 1), (9, 1), (2, 2), (2, 5), (7, 2)])
def test_reference_evaluator_on_odd_shapes(n_j, n_m):
    """Evaluator.forward of the reference (env/message_passing_evl.py:161-199) on random ACYCLIC machine orders --
    for n_m = 1 any permutation of the jobs -- equals the oracle's topological evaluation and its literal Jacobi
    restatement, node for node."""
    R = ref_harness.load()
    rng = np.random.RandomState(11 * n_j + n_m)
    n, n1, B = n_j * n_m, n_j * n_m + 2, 7
    eis, durs, want = [], [], []
    pc = O.edge_index_pc(n_j, n_m)
    for b in range(B):
        dur = rng.randint(1, 30, size=(n_j, n_m))
        mch = rng.random_sample((n_j, n_m)).argsort(axis=1) + 1
        mseq = O.rules_solver(dur, mch, rule="spt")[0]
        if n_m == 1:
            mseq = (rng.permutation(n_j) + 1).reshape(1, n_j)          # any job order is a valid single-machine schedule
        dur1 = O.padded_durations(dur)
        mpred, msucc = O.machine_links(mseq, n1)
        est, lst, ms = O.evaluate(dur1, mpred, msucc, n_j, n_m)
        want.append((est, lst, ms))
        eis.append(np.concatenate([pc, O.edge_index_mc(msucc, n)], axis=1) + b * n1)
        durs.append(dur1)
    ei = np.concatenate(eis, axis=1)
    ei = ei[:, rng.permutation(ei.shape[1])]
    dur = np.concatenate(durs).reshape(-1, 1)
    r_est, r_lst, r_ms = R.Evaluator().forward(torch.from_numpy(ei), torch.from_numpy(dur), n_j, n_m)
    j_est, j_lst, j_ms = O.evaluate_coo_jacobi(ei, dur, n_j, n_m)
    assert np.array_equal(r_est.numpy().reshape(-1), j_est) and np.array_equal(r_lst.numpy().reshape(-1), j_lst)
    assert np.array_equal(r_ms.numpy().reshape(-1), j_ms)
    for b, (est, lst, ms) in enumerate(want):
        assert np.array_equal(r_est.numpy().reshape(B, n1)[b].astype(np.int64), est)
        assert np.array_equal(r_lst.numpy().reshape(B, n1)[b].astype(np.int64), lst)
        assert int(r_ms[b]) == ms


@pytest.mark.parametrize("n_j", [3, 4, 9])
def test_reference_single_machine_environment(n_j):
    """JsspN5 of the reference on single-machine instances with pairwise distinct durations (no priority tie, so `spt`
    consumes no random number): initial order, features, move sets (always the dummy move: the critical path is one
    block over the only machine), rewards and incumbents equal the oracle's over a few steps."""
    R = ref_harness.load()
    rng = np.random.RandomState(5 * n_j)
    B = 6
    dur = np.stack([rng.permutation(n_j) + 1 for _ in range(B)]).reshape(B, n_j, 1)
    inst = np.stack([dur, np.ones_like(dur)], axis=1).astype(np.int64)
    env = R.JsspN5(n_j, 1, 1, 99)
    state, fa, done = env.reset(inst, "spt", "cpu")
    oenv = O.OracleEnv(n_j, 1, 1, 99)
    mseq = np.stack([O.rules_solver(i[0], i[1], rule="spt")[0] for i in inst])
    _, ofa, _ = oenv.reset(inst, mseq=mseq)
    for k in range(4):
        x = state[0].numpy().reshape(B, -1, 3)
        for b, ins in enumerate(oenv.insts):
            assert np.array_equal(x[b], O.features(ins.dur1, ins.est, ins.lst, 99, 1000)), (k, b)
            assert [list(p) for p in fa[b]] == [list(p) for p in ofa[b]], (k, b)
        assert np.array_equal(env.current_objs.numpy().reshape(-1), np.array([i.makespan for i in oenv.insts], np.float32))
        acts = [list(f[0]) for f in fa]
        state, rew, fa, done = env.step(acts, "cpu")
        _, orew, ofa, _ = oenv.step(acts)
        assert np.array_equal(rew.numpy().reshape(-1), np.asarray(orew, dtype=np.float32).reshape(-1))


@pytest.mark.parametrize("n_j,n_m,steps", [(100, 20, 12), (150, 25, 8)])
def test_reference_large_instances(n_j, n_m, steps):
    """BASELINE configs[3] shapes, which no golden fixture holds (the reference needs seconds per instance there): one
    instance stepped by the live reference, by the numpy oracle and by the C oracle from the reference's own initial
    machine orders -- features, machine edges, move sets in order, makespan and reward at every step.  (The CUDA path
    is checked against the C oracle at these shapes in tests/test_gpu_fullsize.py: this closes the chain.)"""
    from oracle import c_oracle as CO
    from oracle.make_golden import ref_initial_orders
    R = ref_harness.load()
    rng = np.random.RandomState(n_j)
    dur = rng.randint(1, 99, size=(1, n_j, n_m))
    mch = rng.random_sample((1, n_j, n_m)).argsort(axis=2) + 1
    inst = np.stack([dur, mch], axis=1).astype(np.int64)
    env, state, fa, done, mseq0 = ref_initial_orders(R, inst, "fdd-divide-mwkr", 1)
    oenv = O.OracleEnv(n_j, n_m, 1, 99)
    _, ofa, _ = oenv.reset(inst, mseq=mseq0)
    cenv = CO.CEnv(inst, mseq0, pmax=64)
    cenv.run("replay", 0, tape=np.zeros((1, 0, 2), np.int32), write_state=True)
    for k in range(steps + 1):
        ins = oenv.insts[0]
        assert np.array_equal(state[0].numpy(), O.features(ins.dur1, ins.est, ins.lst, 99, 1000)), k
        assert [list(p) for p in fa[0]] == [list(p) for p in ofa[0]], k
        assert float(env.current_objs[0]) == ins.makespan
        assert np.array_equal(state[2].numpy(), O.edge_index_mc(ins.msucc, n_j * n_m)), k
        assert cenv.status == 0 and np.array_equal(cenv.est[0], ins.est) and np.array_equal(cenv.lst[0], ins.lst), k
        if k:       # (the C oracle writes x / edge_index_mc from its step loop)
            assert np.array_equal(cenv.x, state[0].numpy()) and np.array_equal(cenv.emc, state[2].numpy()), k
        assert cenv.pairs[0, :cenv.npairs[0]].tolist() == [list(p) for p in fa[0]], k
        assert float(cenv.makespan[0]) == ins.makespan and float(cenv.incumbent[0]) == float(env.incumbent_objs[0])
        if k == steps:
            break
        acts = [list(fa[0][rng.randint(len(fa[0]))])]
        state, rew, fa, done = env.step(acts, "cpu")
        _, orew, ofa, _ = oenv.step(acts)
        cenv.run("replay", 1, tape=np.asarray(acts, np.int32).reshape(1, 1, 2), write_state=True)
        assert float(rew[0]) == float(np.asarray(orew).reshape(-1)[0])


@pytest.mark.parametrize("case", range(24))
def test_reference_differential_on_random_small_shapes(case):
    """Differential run, live reference vs numpy oracle, on random small shapes (2-9 jobs, 1-6 machines, short
    durations so that equal-length paths and tabu interplay are frequent), both reward types, random feasible moves
    mixed with dummy moves: every returned quantity at every step; the reference's IndexError corner
    (env/environment.py:89-90) must be raised by both or by neither."""
    from oracle.make_golden import ref_initial_orders
    R = ref_harness.load()
    rng = np.random.RandomState(1000 + case)
    n_j, n_m = int(rng.randint(2, 10)), int(rng.randint(1, 7))
    B, K = 4, 15
    high = 99          # (the reference's left shift uses the global `args.h` = 99 as its "empty slot" sentinel,
    #                     env/permissible_LS.py:31: any other `high` makes its reset raise IndexError)
    reward_type = "yaoxin" if case % 2 == 0 else "consecutive"
    init = "spt" if case % 3 == 0 else "fdd-divide-mwkr"
    dur = rng.randint(1, int(rng.choice([4, 9, 99])), size=(B, n_j, n_m))
    mch = rng.random_sample((B, n_j, n_m)).argsort(axis=2) + 1
    inst = np.stack([dur, mch], axis=1).astype(np.int64)
    n1 = n_j * n_m + 2

    def both(f_ref, f_orc):
        out = []
        for f in (f_ref, f_orc):
            try:
                out.append(("ok", f()))
            except IndexError:
                out.append(("IndexError", None))
        assert out[0][0] == out[1][0], (out[0][0], out[1][0])
        return out[0][0] == "ok", out[0][1], out[1][1]

    state_env = {}

    def ref_reset():
        np.random.seed(case)
        env = R.JsspN5(n_j, n_m, 1, high, reward_type=reward_type, fea_norm_const=1000)
        state, fa, done = env.reset(inst, init, "cpu")
        emc = state[2].numpy()
        E = emc.shape[1] // B
        from oracle.make_golden import mseq_from_emc
        state_env["mseq0"] = np.stack([mseq_from_emc(emc[:, b * E:(b + 1) * E] - b * n1, inst[b, 1], n_j, n_m)
                                       for b in range(B)]) if E else None
        return env, state, fa

    try:
        env, state, fa = ref_reset()
        ref_err = None
    except IndexError:
        ref_err = "IndexError"
    if state_env.get("mseq0") is None and n_j > 1:
        # (reset raised before the state existed: take the oracle's own initial orders; only valid without ties)
        state_env["mseq0"] = np.stack([O.rules_solver(i[0], i[1], rule=init, high=high)[0] for i in inst])
    oenv = O.OracleEnv(n_j, n_m, 1, high, reward_type=reward_type, fea_norm_const=1000)
    try:
        _, ofa, _ = oenv.reset(inst, mseq=state_env["mseq0"])
        orc_err = None
    except IndexError:
        orc_err = "IndexError"
    if ref_err or orc_err:
        # with the oracle's own initial orders a tie may have led the two to different schedules: only when the
        # reference produced its state can the corner be compared
        if ref_err is None:
            assert orc_err is None, "oracle raised IndexError on a state the reference accepts"
        return
    from oracle import c_oracle as CO
    cenv = CO.CEnv(inst, state_env["mseq0"], pmax=64, high=float(high))
    for k in range(K):
        x = state[0].numpy().reshape(B, n1, 3)
        for b, ins in enumerate(oenv.insts):
            assert np.array_equal(x[b], O.features(ins.dur1, ins.est, ins.lst, high, 1000)), (k, b)
            assert [list(p) for p in fa[b]] == [list(p) for p in ofa[b]], (k, b, fa[b], ofa[b])
            # ... and the C oracle (what the full-size GPU tests compare the CUDA path with)
            assert np.array_equal(cenv.est[b], ins.est) and np.array_equal(cenv.lst[b], ins.lst), (k, b)
            assert (cenv.pairs[b, :cenv.npairs[b]].tolist() or [[0, 0]]) == [list(p) for p in fa[b]], (k, b)
        assert cenv.status == 0
        assert np.array_equal(env.current_objs.numpy().reshape(-1), np.array([i.makespan for i in oenv.insts], np.float32))
        assert np.array_equal(env.current_objs.numpy().reshape(-1), cenv.makespan)
        assert np.array_equal(env.incumbent_objs.numpy().reshape(-1), cenv.incumbent)
        acts = [[0, 0] if rng.rand() < 0.15 else list(f[rng.randint(len(f))]) for f in fa]
        ok, r_out, o_out = both(lambda: env.step(acts, "cpu"), lambda: oenv.step(acts))
        cenv.run("replay", 1, tape=np.asarray(acts, np.int32).reshape(B, 1, 2))
        if not ok:
            assert cenv.status == -9
            return
        state, rew, fa, done = r_out
        _, orew, ofa, odone = o_out
        assert np.array_equal(rew.numpy().reshape(-1), np.asarray(orew, dtype=np.float32).reshape(-1)), k
        assert np.array_equal(env.incumbent_objs.numpy().reshape(-1), np.asarray(oenv.incumbent_objs, np.float32).reshape(-1)), k


def test_reference_index_error_corner():
    """Two jobs on one machine: the critical path is two operations of the same machine and `_get_pairs`
    (env/environment.py:89-90) indexes past its end -- the reference raises IndexError from `reset`; so does the oracle
    (the CUDA path reports status -9 -> IndexError, tests/test_gpu_fullsize.py::test_odd_shapes_against_c_oracle)."""
    R = ref_harness.load()
    inst = np.array([[[[3], [5]], [[1], [1]]]], dtype=np.int64)
    with pytest.raises(IndexError):
        R.JsspN5(2, 1, 1, 99).reset(inst, "spt", "cpu")
    with pytest.raises(IndexError):
        O.OracleEnv(2, 1, 1, 99).reset(inst, "spt")


@pytest.mark.parametrize("n_j,n_m", [(6, 4), (8, 3), (5, 5), (10, 2), (4, 6)])
def test_reference_greedy_baseline_on_random_instances(n_j, n_m):
    """`Greedy_baselines` of the reference (conventional_baselines_with_long-term-mem.py:154-221: best neighbour by
    full re-evaluation, first minimum) on random instances == the C oracle's GREEDY policy (the one the fused CUDA
    policy is checked against), incumbents at every logged step."""
    import random
    from oracle import c_oracle as CO
    from oracle.make_golden import ref_initial_orders
    R = ref_harness.load()
    rng = np.random.RandomState(3 * n_j + n_m)
    B, steps = 5, [1, 2, 5, 10, 20, 40]
    dur = rng.randint(1, 99, size=(B, n_j, n_m))
    mch = rng.random_sample((B, n_j, n_m)).argsort(axis=2) + 1
    inst = np.stack([dur, mch], axis=1).astype(np.int64)
    _, _, _, _, mseq0 = ref_initial_orders(R, inst, "fdd-divide-mwkr", 1)
    random.seed(1)
    np.random.seed(1)
    res, _ = R.Greedy_baselines(inst, max(steps), list(steps), "cpu")
    cenv = CO.CEnv(inst, mseq0)
    _, logs = cenv.run("greedy", max(steps), log_steps=steps)
    assert cenv.status == 0
    assert np.array_equal(np.asarray(res, dtype=np.float32).reshape(len(steps), B), logs)


@pytest.mark.parametrize("n_j,n_m", [(3, 1), (2, 4), (6, 6), (10, 5), (15, 10), (20, 15)])
def test_plist_orders_of_the_dropin_equal_the_live_reference(n_j, n_m):
    """a-14: `reset(..., init_type='plist')` (env/environment.py:140-200, 391-393).  The drop-in builds the machine
    orders on the host (`JsspN5._plist_orders`, one `np.random.permutation` like the reference) and installs them with
    `l2s_set_solution`; with the numpy RNG seeded alike, the orders equal those of the live reference's reset (read
    back from its machine edges) for every instance of the batch -- and the RNG is left in the same state."""
    from l2s_b200 import JsspN5
    from oracle.make_golden import mseq_from_emc
    R = ref_harness.load()
    rng = np.random.RandomState(17 * n_j + n_m)
    B = 4
    dur = rng.randint(1, 99, size=(B, n_j, n_m))
    mch = rng.random_sample((B, n_j, n_m)).argsort(axis=2) + 1
    inst = np.stack([dur, mch], axis=1).astype(np.int64)
    n1 = n_j * n_m + 2
    np.random.seed(77)
    try:
        state, fa, done = R.JsspN5(n_j, n_m, 1, 99).reset(inst, "plist", "cpu")
    except IndexError:
        pytest.skip("the reference's own IndexError corner (two-operation critical path on one machine)")
    after_ref = np.random.randint(1 << 30)
    emc = state[2].numpy()
    E = emc.shape[1] // B
    want = np.stack([mseq_from_emc(emc[:, b * E:(b + 1) * E] - b * n1, inst[b, 1], n_j, n_m) for b in range(B)])
    np.random.seed(77)
    got = JsspN5(n_j, n_m, 1, 99)._plist_orders(inst)
    assert np.random.randint(1 << 30) == after_ref
    assert np.array_equal(np.asarray(got).reshape(B, n_m, n_j), want)


def test_reinforce_loss_equals_the_live_reference_learn():
    """f-4: `training.reinforce_loss` (vectorised) against the reference's own `RL2S4JSSP.learn`
    (n-step_reinforce.py:40-61, loaded from the unmodified script): same loss gradient with respect to every
    log-probability, for random rewards and done patterns (instances that finish part-way included)."""
    import importlib.util
    import os
    import sys
    import types
    from l2s_b200.training import reinforce_loss
    R = ref_harness.load()
    sys.path.insert(0, R.root)
    try:
        spec = importlib.util.spec_from_file_location("ref_nstep_reinforce", os.path.join(R.root, "n-step_reinforce.py"))
        mod = importlib.util.module_from_spec(spec)
        saved_argv, sys.argv = sys.argv, [sys.argv[0]]
        try:
            spec.loader.exec_module(mod)
        finally:
            sys.argv = saved_argv
    finally:
        sys.path.remove(R.root)
    gen = torch.Generator().manual_seed(5)
    B, n = 7, 10
    for trial in range(4):
        rewards = [torch.randint(0, 30, (B, 1), generator=gen).float() for _ in range(n)]
        first_done = torch.randint(3, n + 4, (B, 1), generator=gen)              # some instances never finish
        dones = [torch.full((B, 1), t, dtype=torch.long) >= first_done for t in range(n)]
        base = [torch.randn(B, 1, generator=gen) for _ in range(n)]
        lp_ref = [b.clone().requires_grad_(True) for b in base]
        lp_own = [b.clone().requires_grad_(True) for b in base]
        me = types.SimpleNamespace(eps=np.finfo(np.float32).eps.item())
        opt = types.SimpleNamespace(zero_grad=lambda: None, step=lambda: None)
        mod.RL2S4JSSP.learn(me, rewards, lp_ref, dones, opt)
        reinforce_loss(rewards, lp_own, dones, gamma=float(mod.args.gamma)).backward()
        for a, b in zip(lp_ref, lp_own):
            ga = a.grad if a.grad is not None else torch.zeros_like(a)
            gb = b.grad if b.grad is not None else torch.zeros_like(b)
            assert torch.allclose(ga, gb, rtol=1e-5, atol=1e-6), trial


def test_checkpoint_names_follow_the_reference_scheme():
    """f-4: `training.model_name` built from the reference's OWN default arguments (parameters.py) is the name of the
    model the reference ships for that configuration (saved_model/, loaded by test_learned.py:141-148)."""
    import os
    import sys
    from l2s_b200.training import model_name
    R = ref_harness.load()
    sys.path.insert(0, R.root)
    try:
        saved_argv, sys.argv = sys.argv, [sys.argv[0]]
        try:
            from parameters import args
        finally:
            sys.argv = saved_argv
    finally:
        sys.path.remove(R.root)
    name = model_name('incumbent', args.j, args.m, args.l, args.h, args.init_type, args.reward_type, args.gamma,
                      args.hidden_dim, args.embedding_layer, args.policy_layer, args.embedding_type, args.heads,
                      args.drop_out, args.lr, args.steps_learn, args.transit, args.batch_size, args.episodes,
                      args.step_validation)
    assert os.path.isfile(os.path.join(R.root, "saved_model", name)), name


@pytest.mark.parametrize("rule", ["spt", "fdd-divide-mwkr"])
def test_initial_solutions_against_the_live_reference(rule):
    """a-12 / a-13: `_rules_solver` + `permissibleLeftShift` (env/environment.py:203-289, env/permissible_LS.py) of the
    live reference on random shapes with tie-heavy and tie-free durations.  The oracle drawing its tie-breaks from an
    identically seeded RandomState reproduces the reference's machine orders ALWAYS; its first-minimum rule (what the
    device builder implements, DESIGN.md 4 'known deviations') reproduces them whenever no tie occurred; the C oracle's
    builder equals the numpy one."""
    from oracle import c_oracle as CO
    from oracle.make_golden import mseq_from_emc
    R = ref_harness.load()
    ran = with_ties = 0
    for case in range(60):
        rng = np.random.RandomState(case)
        n_j, n_m = int(rng.randint(3, 13)), int(rng.randint(2, 9))
        high = 99      # (must equal the reference's global args.h, env/permissible_LS.py:31)
        dur = rng.randint(1, int(rng.choice([3, 6, 99])), size=(1, n_j, n_m))
        mch = rng.random_sample((1, n_j, n_m)).argsort(axis=2) + 1
        inst = np.stack([dur, mch], axis=1).astype(np.int64)
        np.random.seed(case)
        try:
            state, fa, done = R.JsspN5(n_j, n_m, 1, high).reset(inst, rule, "cpu")
        except IndexError:
            continue                                    # (the reference's two-operation critical path corner)
        want = mseq_from_emc(state[2].numpy(), inst[0, 1], n_j, n_m)
        first, ties = O.rules_solver(inst[0, 0], inst[0, 1], rule, high)
        drawn, _ = O.rules_solver(inst[0, 0], inst[0, 1], rule, high, rng=np.random.RandomState(case))
        assert np.array_equal(drawn, want), case
        if ties == 0:
            assert np.array_equal(first, want), case
        assert np.array_equal(CO.rules_solver(inst[0, 0], inst[0, 1], rule, high)[0], first), case
        ran += 1
        with_ties += ties > 0
    assert ran >= 30 and with_ties >= 5 and ran - with_ties >= 5, (ran, with_ties)
