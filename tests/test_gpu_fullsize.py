"""GPU parity at BASELINE.json's full sizes against the plain-C oracle (same seeded inputs), plus
size-independent properties: swap involution, shard invariance, greedy known answers."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def gen_instances(n_j, n_m, B, seed):
    rng = np.random.RandomState(seed)
    dur = rng.randint(1, 99, size=(B, n_j, n_m))
    mch = rng.random_sample((B, n_j, n_m)).argsort(axis=2) + 1
    return np.stack([dur, mch], axis=1).astype(np.int32)


@pytest.mark.parametrize("n_j,n_m,B,K", [(20, 15, 8192, 30), (10, 10, 512, 60), (100, 20, 64, 12), (150, 25, 32, 8),
                                          (30, 20, 128, 20), (50, 15, 64, 20), (6, 6, 300, 50), (17, 3, 100, 40)])
def test_full_size_random_search_equals_c_oracle(n_j, n_m, B, K):
    from l2s_b200 import JsspN5
    from oracle import c_oracle as CO
    inst = gen_instances(n_j, n_m, B, seed=n_j * 100 + n_m)
    env = JsspN5(n_j, n_m, 1, 99)
    env.instance_offset = 77
    state, fa, done = env.reset(inst, "fdd-divide-mwkr", "cuda")
    init = env.solutions().cpu().numpy()
    # device initial-solution builder == C restatement of _rules_solver (first-min tie rule) on every instance
    for b in range(0, B, max(1, B // 256)):
        want, _ = CO.rules_solver(inst[b, 0], inst[b, 1], "fdd-divide-mwkr")
        assert np.array_equal(init[b], want), "initial solution of instance %d" % b
    cenv = CO.CEnv(inst, init, inst_offset=77)
    assert np.array_equal(env.current_objs.cpu().numpy().reshape(-1), cenv.makespan)
    logs, taken = env.run("random", K, seed=2024, log_steps=[1, K], return_actions=True)
    ctaken, clogs = cenv.run("random", K, seed=2024, log_steps=[1, K])
    assert cenv.status == 0
    assert np.array_equal(taken.cpu().numpy(), ctaken)
    assert np.array_equal(logs.cpu().numpy(), clogs)
    # one more step through the step API with every output compared
    state, fa, done = env.observe()
    acts = [f[0] for f in fa]
    state, reward, fa, done = env.step(acts, "cuda")
    cenv.run("replay", 1, tape=np.array(acts, dtype=np.int32).reshape(B, 1, 2), write_state=True)
    assert np.array_equal(env.current_objs.cpu().numpy().reshape(-1), cenv.makespan)
    assert np.array_equal(env.incumbent_objs.cpu().numpy().reshape(-1), cenv.incumbent)
    assert np.array_equal(state[0].cpu().numpy(), cenv.x)
    assert np.array_equal(state[2].cpu().numpy(), cenv.emc)
    ex = env.export_state()
    assert np.array_equal(ex["est"].cpu().numpy(), cenv.est) and np.array_equal(ex["lst"].cpu().numpy(), cenv.lst)
    assert np.array_equal(ex["jobfirst"].cpu().numpy()[:, 1:-1], cenv.jobfirst[:, 1:-1])
    got_np = np.array([0 if f == [[0, 0]] else len(f) for f in fa])
    assert np.array_equal(got_np, cenv.npairs)
    for b in range(0, B, max(1, B // 64)):
        if cenv.npairs[b]:
            assert fa[b] == cenv.pairs[b, :cenv.npairs[b]].tolist()
    # size-independent invariants: est+dur <= lst+dur <= makespan, est[T] = lst[T] = makespan, lst[S] = 0
    est, lst = ex["est"].cpu().numpy(), ex["lst"].cpu().numpy()
    assert (est <= lst).all() and (lst[:, 0] == 0).all() and (est[:, -1] == lst[:, -1]).all()


@pytest.mark.parametrize("ds", ["ft6x6", "la10x5", "tai20x15"])
def test_greedy_policy_reproduces_reference_stored_results(ds):
    """l2s_run(GREEDY) for 500 steps == the reference's stored greedy result row
    (test_results/conventional_results/greedy-policy/*_result.npy, row 0)."""
    from l2s_b200 import JsspN5
    g = np.load(os.path.join(GOLDEN, "greedy_%s.npz" % ds))
    inst = g["instances"]
    env = JsspN5(inst.shape[2], inst.shape[3], 1, 99)
    env.reset(inst, "fdd-divide-mwkr", "cuda", init_solutions=g["mseq0"])
    steps = [int(s) for s in g["log_steps"]]
    logs, _ = env.run("greedy", 500, log_steps=steps)
    assert np.array_equal(logs.cpu().numpy(), g["incumbents"])
    assert np.array_equal(logs.cpu().numpy()[steps.index(500)], g["stored_rows"][0])


def test_swap_is_an_involution_and_shards_are_invariant():
    from l2s_b200 import JsspN5
    n_j, n_m, B = 20, 15, 600
    inst = gen_instances(n_j, n_m, B, seed=5)
    env = JsspN5(n_j, n_m, 1, 99)
    s0, fa, _ = env.reset(inst, "fdd-divide-mwkr", "cuda")
    ms0 = env.current_objs.clone()
    acts = [f[len(f) // 2] for f in fa]
    env.step(acts, "cuda")
    s2, _, _, _ = env.step([[a[1], a[0]] for a in acts], "cuda")      # swap back (tabu only filters the move LIST)
    assert torch.equal(env.current_objs, ms0) and torch.equal(s2[0], s0[0]) and torch.equal(s2[2], s0[2])
    # two slices stepped by separate handles == one handle (RANDOM policy keyed by the global instance index)
    full = JsspN5(n_j, n_m, 1, 99)
    full.reset(inst, "fdd-divide-mwkr", "cuda")
    lf, tf = full.run("random", 25, seed=3, log_steps=[25], return_actions=True)
    parts = []
    for lo, hi in ((0, 250), (250, 600)):
        e = JsspN5(n_j, n_m, 1, 99)
        e.instance_offset = lo
        e.reset(inst[lo:hi], "fdd-divide-mwkr", "cuda")
        parts.append(e.run("random", 25, seed=3, log_steps=[25], return_actions=True))
    assert torch.equal(torch.cat([p[0] for p in parts], dim=1), lf)
    assert torch.equal(torch.cat([p[1] for p in parts], dim=0), tf)


def test_input_validation_errors():
    from l2s_b200 import JsspN5, L2SError
    inst = gen_instances(6, 6, 4, seed=1)
    bad = inst.copy()
    bad[1, 0, 2, 3] = 0                      # zero duration (reference quirk H4b) -> explicit error
    with pytest.raises(L2SError):
        JsspN5(6, 6, 1, 99).reset(bad, "fdd-divide-mwkr", "cuda")
    bad = inst.copy()
    bad[2, 1, 0, :] = 1                      # a job visiting machine 1 six times
    with pytest.raises(L2SError):
        JsspN5(6, 6, 1, 99).reset(bad, "fdd-divide-mwkr", "cuda")
    sol = np.zeros((4, 6, 6), dtype=np.int32)
    with pytest.raises(L2SError):
        JsspN5(6, 6, 1, 99).reset(inst, "fdd-divide-mwkr", "cuda", init_solutions=sol)
    with pytest.raises(L2SError):
        JsspN5(40, 40, 1, 99).reset(gen_instances(40, 40, 1, 1), "spt", "cuda")      # n_m > 32 unsupported


@pytest.mark.parametrize("n_j,n_m,B", [(20, 15, 4096), (10, 10, 300), (16, 16, 64), (100, 20, 32), (150, 25, 16), (7, 32, 40), (4, 1, 50)])
def test_evaluator_coo_full_size(n_j, n_m, B):
    """Evaluator.forward on shuffled int64 COO == the integer est/lst of the environment (itself checked against
    the oracle above) for the same schedules; malformed inputs are rejected."""
    from l2s_b200 import Evaluator, JsspN5, L2SError
    inst = gen_instances(n_j, n_m, B, seed=3 * n_j + n_m)
    env = JsspN5(n_j, n_m, 1, 99)
    state, fa, done = env.reset(inst, "spt", "cuda")
    env.run("random", 15, seed=1)
    state, fa, done = env.observe()
    ex = env.export_state()
    n1 = n_j * n_m + 2
    ei = torch.cat([state[1], state[2]], dim=1)
    ei = ei[:, torch.randperm(ei.shape[1], device=ei.device)]
    dur = torch.zeros((B, n1), dtype=torch.int64, device="cuda")
    dur[:, 1:-1] = torch.from_numpy(inst[:, 0].reshape(B, -1).astype(np.int64)).cuda()
    est, lst, ms = Evaluator().forward(ei, dur.reshape(-1, 1), n_j, n_m)
    assert torch.equal(est.reshape(B, n1), ex["est"].float()) and torch.equal(lst.reshape(B, n1), ex["lst"].float())
    assert torch.equal(ms, env.current_objs)
    # a duplicated machine arc / a missing edge / a reversed arc are rejected (conjunctive arcs are implicit in
    # the node numbering: they are counted, not verified one by one)
    bad = torch.cat([state[1], state[2]], dim=1)
    bad[:, -1] = bad[:, -2]
    with pytest.raises(L2SError):
        Evaluator().forward(bad, dur.reshape(-1, 1), n_j, n_m)
    with pytest.raises(L2SError):
        Evaluator().forward(ei[:, :-1], dur.reshape(-1, 1), n_j, n_m)
    mc = state[2].clone()
    mc[:, 0] = torch.flip(mc[:, 0], dims=[0])           # reverse one machine arc -> 2-cycle or broken chain
    with pytest.raises(L2SError):
        Evaluator().forward(torch.cat([state[1], mc], dim=1), dur.reshape(-1, 1), n_j, n_m)


@pytest.mark.parametrize("n_j,n_m", [(2, 2), (3, 1), (2, 5), (5, 2), (3, 7), (4, 32), (13, 4), (33, 3), (40, 16), (40, 17)])
def test_odd_shapes_against_c_oracle(n_j, n_m):
    """Edge shapes: single machine, two jobs, 32 machines (all lanes), more jobs than lanes, the dual/single
    sweep boundary (16 vs 17 machines).  Tie-heavy durations.  The 2-op-path IndexError corner is mirrored."""
    from l2s_b200 import JsspN5
    from oracle import c_oracle as CO
    rng = np.random.RandomState(n_j * 37 + n_m)
    B, K = 40, 25
    dur = rng.randint(1, 6, size=(B, n_j, n_m))
    mch = rng.random_sample((B, n_j, n_m)).argsort(axis=2) + 1
    inst = np.stack([dur, mch], axis=1).astype(np.int32)
    env = JsspN5(n_j, n_m, 1, 99)
    mseq = np.stack([CO.rules_solver(i[0], i[1])[0] for i in inst])
    cenv = CO.CEnv(inst, mseq)
    ctaken, _ = cenv.run("random", K, seed=5)
    if cenv.status == -9:
        with pytest.raises(IndexError):
            env.reset(inst, "fdd-divide-mwkr", "cuda")
            env.run("random", K, seed=5)
        return
    assert cenv.status == 0
    env.reset(inst, "fdd-divide-mwkr", "cuda")
    assert np.array_equal(env.solutions().cpu().numpy(), mseq)
    _, taken = env.run("random", K, seed=5, return_actions=True)
    assert np.array_equal(taken.cpu().numpy(), ctaken)
    state, fa, done = env.observe()
    cenv.run("replay", 1, tape=np.zeros((B, 1, 2), np.int32), write_state=True)   # dummy step: refresh outputs
    ex = env.export_state()
    assert np.array_equal(ex["est"].cpu().numpy(), cenv.est) and np.array_equal(ex["lst"].cpu().numpy(), cenv.lst)
    assert np.array_equal(np.array([0 if f == [[0, 0]] else len(f) for f in fa]), cenv.npairs)


def test_randomised_shapes_and_durations_against_c_oracle():
    """Fuzz: 24 random (n_j, n_m, duration range, init rule) combinations, each stepped through the per-step API with
    every output compared to the plain-C oracle at every step (bit-exact x, edges, pairs, rewards, incumbents)."""
    from l2s_b200 import JsspN5
    from oracle import c_oracle as CO
    # L2S_FUZZ_SEED / L2S_FUZZ_CASES: longer one-off campaigns (e.g. 300 cases, also with L2S_CTA_WARPS forced)
    rng = np.random.RandomState(int(os.environ.get("L2S_FUZZ_SEED", "2024")))
    for case in range(int(os.environ.get("L2S_FUZZ_CASES", "24"))):
        n_j, n_m = int(rng.randint(3, 28)), int(rng.randint(2, 24))
        hi = int(rng.choice([3, 5, 20, 99, 400]))
        B, K = int(rng.randint(3, 40)), int(rng.randint(5, 30))
        rule = str(rng.choice(["fdd-divide-mwkr", "spt"]))
        dur = rng.randint(1, hi + 1, size=(B, n_j, n_m))
        mch = rng.random_sample((B, n_j, n_m)).argsort(axis=2) + 1
        inst = np.stack([dur, mch], axis=1).astype(np.int32)
        fea = float(rng.choice([1000, 999.5, 37]))
        env = JsspN5(n_j, n_m, 1, max(hi, 99), fea_norm_const=fea, reward_type=str(rng.choice(["yaoxin", "consecutive"])))
        try:
            state, fa, done = env.reset(inst, rule, "cuda")
        except IndexError:
            continue          # the reference's 2-op-path corner; covered elsewhere
        mseq = np.stack([CO.rules_solver(i[0], i[1], rule, max(hi, 99))[0] for i in inst])
        assert np.array_equal(env.solutions().cpu().numpy(), mseq), "case %d initial solution" % case
        cenv = CO.CEnv(inst, mseq, high=float(max(hi, 99)), fea=fea, pmax=env.pmax)
        if cenv.status != 0:
            continue
        prev_cur = cenv.makespan.copy()
        for k in range(K):
            acts = [f[int(rng.randint(len(f)))] for f in fa]
            prev_inc = cenv.incumbent.copy()
            cenv.run("replay", 1, tape=np.array(acts, dtype=np.int32).reshape(B, 1, 2), write_state=True)
            if cenv.status != 0:
                break
            state, reward, fa, done = env.step(acts, "cuda")
            assert np.array_equal(state[0].cpu().numpy(), cenv.x), "case %d step %d x" % (case, k)
            assert np.array_equal(state[2].cpu().numpy(), cenv.emc), "case %d step %d edges" % (case, k)
            assert np.array_equal(env.current_objs.cpu().numpy().reshape(-1), cenv.makespan)
            assert np.array_equal(env.incumbent_objs.cpu().numpy().reshape(-1), cenv.incumbent)
            want_r = (prev_cur - cenv.makespan) if env.reward_type == "consecutive" else np.maximum(prev_inc - cenv.makespan, 0)
            assert np.array_equal(reward.cpu().numpy().reshape(-1), want_r.astype(np.float32))
            prev_cur = cenv.makespan.copy()
            for b in range(B):
                want = cenv.pairs[b, :cenv.npairs[b]].tolist() if cenv.npairs[b] else [[0, 0]]
                assert fa[b] == want, "case %d step %d instance %d pairs" % (case, k, b)


def test_evaluator_workspace_is_handed_back_clean():
    """One Evaluator (one workspace) across shapes, int32 / int64 durations and rejected inputs: every call leaves
    the workspace zero-filled, so the next one -- which skips its memset -- is still exact."""
    from l2s_b200 import Evaluator, JsspN5, L2SError
    ev = Evaluator()
    for rnd, (n_j, n_m, B) in enumerate([(20, 15, 64), (10, 10, 200), (15, 15, 33), (20, 15, 64), (6, 6, 500)]):
        inst = gen_instances(n_j, n_m, B, seed=11 + rnd)
        env = JsspN5(n_j, n_m, 1, 99)
        env.reset(inst, "fdd-divide-mwkr", "cuda")
        env.run("random", 7, seed=rnd)
        state, fa, done = env.observe()
        ex = env.export_state()
        n1 = n_j * n_m + 2
        ei = torch.cat([state[1], state[2]], dim=1)
        dur = torch.zeros((B, n1), dtype=torch.int64 if rnd % 2 == 0 else torch.int32, device="cuda")
        dur[:, 1:-1] = torch.from_numpy(inst[:, 0].reshape(B, -1)).cuda().to(dur.dtype)
        if rnd in (1, 3):        # a rejected call in between: duplicated machine arc
            bad = ei.clone()
            bad[:, -1] = bad[:, -2]
            with pytest.raises(L2SError):
                ev.forward(bad, dur.reshape(-1, 1), n_j, n_m)
        est, lst, ms = ev.forward(ei, dur.reshape(-1, 1), n_j, n_m)
        assert torch.equal(est.reshape(B, n1), ex["est"].float()) and torch.equal(lst.reshape(B, n1), ex["lst"].float())
        assert torch.equal(ms, env.current_objs)
        assert int(ev._ws.count_nonzero()) == 0


@pytest.mark.parametrize("n_j,n_m,B,K,cta", [(100, 20, 24, 40, None), (150, 25, 8, 25, None), (30, 20, 40, 60, "4"),
                                              (20, 15, 50, 60, "2"), (12, 7, 30, 80, "1")])
def test_incremental_step_equals_full_sweep_and_c_oracle(n_j, n_m, B, K, cta, monkeypatch):
    """The CTA step kernel resumes its forward / backward sweeps from the evaluation it kept in HBM.  The same long
    trajectory (random feasible moves with dummy moves, evaluate-only launches, a fused run and an illegal action in
    between) must give bit-identical outputs at every step as a handle created with L2S_INCREMENTAL=0 (always sweeps
    everything) and as the plain-C oracle."""
    from l2s_b200 import JsspN5
    from oracle import c_oracle as CO
    import networkx as nx
    if cta:
        monkeypatch.setenv("L2S_CTA_WARPS", cta)           # force the CTA kernel for shapes that default to the warp kernel
    inst = gen_instances(n_j, n_m, B, seed=31 + n_j)
    monkeypatch.setenv("L2S_INCREMENTAL", "1")
    inc = JsspN5(n_j, n_m, 1, 99)
    s_i, fa_i, _ = inc.reset(inst, "fdd-divide-mwkr", "cuda")
    monkeypatch.setenv("L2S_INCREMENTAL", "0")
    full = JsspN5(n_j, n_m, 1, 99)
    s_f, fa_f, _ = full.reset(inst, "fdd-divide-mwkr", "cuda")
    monkeypatch.delenv("L2S_INCREMENTAL")
    cenv = CO.CEnv(inst, inc.solutions().cpu().numpy())
    rng = np.random.RandomState(3)

    def same(tag):
        assert torch.equal(s_i[0], s_f[0]) and torch.equal(s_i[2], s_f[2]), tag
        assert fa_i == fa_f, tag
        assert torch.equal(inc.current_objs, full.current_objs) and torch.equal(inc.incumbent_objs, full.incumbent_objs), tag

    same("reset")
    for k in range(K):
        acts = []
        for f in fa_i:
            p = f[rng.randint(len(f))]
            acts.append([0, 0] if rng.rand() < 0.1 else [int(p[0]), int(p[1])])     # 10 % dummy moves
        s_i, r_i, fa_i, d_i = inc.step(acts, "cuda")
        s_f, r_f, fa_f, d_f = full.step(acts, "cuda")
        cenv.run("replay", 1, tape=np.array(acts, dtype=np.int32).reshape(B, 1, 2), write_state=True)
        same("step %d" % k)
        assert torch.equal(r_i, r_f) and torch.equal(d_i, d_f)
        assert np.array_equal(s_i[0].cpu().numpy(), cenv.x) and np.array_equal(s_i[2].cpu().numpy(), cenv.emc), k
        if k % 9 == 4:          # evaluate-only launches use the cached evaluation unchanged
            s_i, fa_i, _ = inc.observe()
            s_f, fa_f, _ = full.observe()
            same("observe %d" % k)
            ex_i, ex_f = inc.export_state(), full.export_state()
            assert torch.equal(ex_i["est"], ex_f["est"]) and torch.equal(ex_i["lst"], ex_f["lst"])
            assert np.array_equal(ex_i["est"].cpu().numpy(), cenv.est) and np.array_equal(ex_i["lst"].cpu().numpy(), cenv.lst)
        if k == K // 2:         # a fused run moves the machine orders without the step kernel: the cache must be dropped
            inc.run("random", 7, seed=5)
            full.run("random", 7, seed=5)
            cenv.run("random", 7, seed=5)
            s_i, fa_i, _ = inc.observe()
            s_f, fa_f, _ = full.observe()
            same("after run")
        if k == K // 3:         # an illegal action is reported and leaves state and cache untouched
            bad = [list(map(int, f[0])) for f in fa_i]
            bad[0] = [1, 1 + n_m]           # first operations of jobs 0 and 1: never an adjacent machine pair? (only if same machine)
            if inst[0, 1, 0, 0] != inst[0, 1, 1, 0]:
                with pytest.raises(nx.NetworkXError):
                    inc.step(bad, "cuda")
                with pytest.raises(nx.NetworkXError):
                    full.step(bad, "cuda")
                cenv.run("replay", 1, tape=np.array([[0, 0]] + bad[1:], dtype=np.int32).reshape(B, 1, 2), write_state=True)
                inc.itr += 1
                full.itr += 1
                s_i, fa_i, _ = inc.observe()
                s_f, fa_f, _ = full.observe()
                same("after illegal action")
                assert np.array_equal(s_i[0].cpu().numpy(), cenv.x)


@pytest.mark.parametrize("n_j,n_m,B,cta", [(20, 15, 40, "2"), (12, 7, 30, "1"), (30, 20, 24, "4"), (100, 20, 10, None),
                                           (9, 32, 12, "8")])
@pytest.mark.parametrize("fast", ["0", "1", "2"])
def test_cta_sweep_variants_equal_c_oracle(n_j, n_m, B, cta, fast, monkeypatch):
    """The three sweep variants of the CTA step kernel (lean, latency, latency + two-operation look-ahead) against the
    plain-C oracle on a 25-step random trajectory: features, machine edges, est / lst, move sets."""
    from l2s_b200 import JsspN5
    from oracle import c_oracle as CO
    if cta:
        monkeypatch.setenv("L2S_CTA_WARPS", cta)
    monkeypatch.setenv("L2S_FAST_SWEEP", fast)
    inst = gen_instances(n_j, n_m, B, seed=11 * n_j + n_m)
    env = JsspN5(n_j, n_m, 1, 99)
    state, fa, done = env.reset(inst, "fdd-divide-mwkr", "cuda")
    cenv = CO.CEnv(inst, env.solutions().cpu().numpy())
    rng = np.random.RandomState(int(fast) + 5)
    for k in range(25):
        acts = [[int(v) for v in f[rng.randint(len(f))]] for f in fa]
        state, reward, fa, done = env.step(acts, "cuda")
        cenv.run("replay", 1, tape=np.array(acts, dtype=np.int32).reshape(B, 1, 2), write_state=True)
        assert np.array_equal(state[0].cpu().numpy(), cenv.x), k
        assert np.array_equal(state[2].cpu().numpy(), cenv.emc), k
        got_np = np.array([0 if f == [[0, 0]] else len(f) for f in fa])
        assert np.array_equal(got_np, cenv.npairs), k
    ex = env.export_state()
    assert np.array_equal(ex["est"].cpu().numpy(), cenv.est) and np.array_equal(ex["lst"].cpu().numpy(), cenv.lst)


@pytest.mark.parametrize("n_j,n_m,B", [(20, 15, 700), (10, 10, 257), (6, 6, 33), (30, 20, 70), (5, 32, 40), (3, 1, 9)])
def test_evaluator_edge_orders_and_malformed_lists(n_j, n_m, B):
    """The reference hands Evaluator.forward cat([edge_index_pc, edge_index_mc]) (env/environment.py:308); the kernels
    accept any column order.  The environment's own order, block-wise rearrangements and a full shuffle give identical
    outputs; lists that are not a batch of complete selections are rejected, and the evaluator stays usable."""
    from l2s_b200 import Evaluator, JsspN5, L2SError
    inst = gen_instances(n_j, n_m, B, seed=5 * n_j + n_m)
    env = JsspN5(n_j, n_m, 1, 99)
    env.reset(inst, "fdd-divide-mwkr", "cuda")
    env.run("random", 9, seed=2)
    state, fa, done = env.observe()
    ex = env.export_state()
    n1 = n_j * n_m + 2
    pc, mc = state[1], state[2]
    E_pc, E_mc = pc.shape[1] // B, mc.shape[1] // B
    ei = torch.cat([pc, mc], dim=1)
    dur = torch.zeros((B, n1), dtype=torch.int64, device="cuda")
    dur[:, 1:-1] = torch.from_numpy(inst[:, 0].reshape(B, -1).astype(np.int64)).cuda()
    dur = dur.reshape(-1, 1)

    def same(out):
        est, lst, ms = out
        return (torch.equal(est.reshape(B, n1), ex["est"].float()) and torch.equal(lst.reshape(B, n1), ex["lst"].float())
                and torch.equal(ms, env.current_objs))

    ev = Evaluator()
    for dt in (torch.int64, torch.int32):
        assert same(ev.forward(ei, dur.to(dt), n_j, n_m))
    g = B // 2
    variants = {}
    v = ei.clone()                                   # one graph's machine block reversed
    v[:, B * E_pc + g * E_mc: B * E_pc + (g + 1) * E_mc] = torch.flip(mc[:, g * E_mc:(g + 1) * E_mc], dims=[1])
    variants["descending"] = v
    v = ei.clone()                                   # two conjunctive columns of one graph exchanged
    v[:, [g * E_pc, g * E_pc + 1]] = v[:, [g * E_pc + 1, g * E_pc]]
    variants["pc-swapped"] = v
    v = ei.clone()                                   # the machine blocks of two graphs exchanged
    a, b = B * E_pc, B * E_pc + (B - 1) * E_mc
    v[:, a:a + E_mc], v[:, b:b + E_mc] = ei[:, b:b + E_mc], ei[:, a:a + E_mc]
    variants["blocks-exchanged"] = v
    variants["mc-first"] = torch.cat([mc, pc], dim=1)
    variants["shuffled"] = ei[:, torch.randperm(ei.shape[1], device="cuda")]
    for name, v in variants.items():
        assert same(ev.forward(v, dur, n_j, n_m)), name
    bad = {}
    v = ei.clone(); v[:, -1] = v[:, -2]; bad["duplicate arc"] = v
    v = ei.clone(); v[1, B * E_pc + g * E_mc] = v[0, B * E_pc + g * E_mc]; bad["self loop"] = v
    v = ei.clone(); v[1, B * E_pc + g * E_mc] = g * n1; bad["arc into S"] = v
    v = ei.clone(); v[0, B * E_pc] = n1 + 1; bad["source in another graph"] = v
    for name, v in bad.items():
        with pytest.raises(L2SError):
            ev.forward(v, dur, n_j, n_m)
        assert same(ev.forward(ei, dur, n_j, n_m)), name                           # and the workspace is usable again
    ev = Evaluator(strict=False)
    out = ev.forward(variants["shuffled"], dur, n_j, n_m)
    ev.check()
    assert same(out)
