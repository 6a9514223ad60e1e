"""GPU-box parity against the LIVE, unmodified reference: the reference's `JsspN5` / `Evaluator` / `CPM_batch_G`
(imported from the vendored `baseline/_ref`, written by oracle/vendor_ref.py; CPU) and the CUDA drop-in step the
same instances with the same actions, and everything the API returns is compared bit for bit at every step.
(/root/reference does not exist on the GPU box; nothing here reads it.)"""
import os
import random

import numpy as np
import pytest
import torch

from oracle import ref_harness

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_harness.available(), reason="baseline/_ref not present (oracle/vendor_ref.py)")]


def _mseq_from_state(state, instances, n_j, n_m):
    from oracle.make_golden import mseq_from_emc
    n1 = n_j * n_m + 2
    emc = state[2].cpu().numpy()
    B = instances.shape[0]
    E = emc.shape[1] // B
    return np.stack([mseq_from_emc(emc[:, b * E:(b + 1) * E] - b * n1, instances[b, 1], n_j, n_m) for b in range(B)])


def _norm(fa):
    return [[[int(p[0]), int(p[1])] for p in f] for f in fa]


def _gen(n_j, n_m, B, seed):
    rng = np.random.RandomState(seed)
    dur = rng.randint(1, 99, size=(B, n_j, n_m))
    mch = rng.random_sample((B, n_j, n_m)).argsort(axis=2) + 1
    return np.stack([dur, mch], axis=1).astype(np.int32)


@pytest.mark.timeout(900)
@pytest.mark.parametrize("n_j,n_m,B,K,evaluator_type", [(20, 15, 16, 20, "message-passing"), (10, 10, 12, 30, "CPM"),
                                                        (30, 20, 2, 6, "message-passing")])
def test_cuda_step_equals_live_reference(n_j, n_m, B, K, evaluator_type):
    from l2s_b200 import JsspN5
    R = ref_harness.load()
    inst = _gen(n_j, n_m, B, seed=100 + n_j)
    np.random.seed(5)
    random.seed(5)
    renv = R.JsspN5(n_j, n_m, 1, 99, evaluator_type=evaluator_type)
    rstate, rfa, rdone = renv.reset(inst, "fdd-divide-mwkr", "cpu")
    env = JsspN5(n_j, n_m, 1, 99, evaluator_type=evaluator_type)
    state, fa, done = env.reset(inst, "fdd-divide-mwkr", "cuda", init_solutions=_mseq_from_state(rstate, inst, n_j, n_m))
    rnd = random.Random(11)

    def same(k):
        for i in range(4):
            assert state[i].dtype == rstate[i].dtype
            assert np.array_equal(state[i].cpu().numpy(), rstate[i].numpy()), "state[%d] differs at step %d" % (i, k)
        assert _norm(fa) == _norm(rfa), "move sets differ at step %d" % k
        assert np.array_equal(done.cpu().numpy(), rdone.numpy())
        assert np.array_equal(env.current_objs.cpu().numpy(), renv.current_objs.numpy())
        assert np.array_equal(env.incumbent_objs.cpu().numpy(), renv.incumbent_objs.numpy())

    same(0)
    for k in range(K):
        acts = [list(map(int, rnd.choice(f))) for f in rfa]
        rstate, rrew, rfa, rdone = renv.step(acts, "cpu")
        state, rew, fa, done = env.step(acts, "cuda")
        assert np.array_equal(rew.cpu().numpy(), rrew.numpy()), "reward differs at step %d" % (k + 1)
        same(k + 1)
        assert env.itr == renv.itr
    assert env.tabu_lists == [[[int(a), int(b)] for a, b in t] for t in renv.tabu_lists]
    # lazily built networkx views (current_graphs / sub_graphs_mc): same arcs, same weights, same critical paths
    import networkx as nx
    for G, Gm, rG, rGm in zip(env.current_graphs, env.sub_graphs_mc, renv.current_graphs, renv.sub_graphs_mc):
        assert sorted(G.edges(data='weight')) == sorted((u, v, int(w)) for u, v, w in rG.edges(data='weight'))
        assert sorted(Gm.edges()) == sorted(rGm.edges())
        assert nx.dag_longest_path(G) == nx.dag_longest_path(rG)


def test_cuda_evaluator_equals_live_reference_evaluator_and_cpm():
    """Evaluator.forward on shuffled COO (int64 and int32 durations) and CPM_batch_G give the same est/lst/makespan."""
    from l2s_b200 import Evaluator
    from l2s_b200.message_passing_evl import CPM_batch_G
    R = ref_harness.load()
    n_j, n_m, B = 15, 10, 6
    inst = _gen(n_j, n_m, B, seed=9)
    np.random.seed(2)
    renv = R.JsspN5(n_j, n_m, 1, 99)
    rstate, rfa, _ = renv.reset(inst, "spt", "cpu")
    for _ in range(5):
        rstate, _, rfa, _ = renv.step([f[0] for f in rfa], "cpu")
    ei = torch.cat([rstate[1], rstate[2]], dim=1)
    ei = ei[:, torch.from_numpy(np.random.RandomState(0).permutation(ei.shape[1]))]
    dur = np.zeros((B, n_j * n_m + 2), dtype=np.int64)
    dur[:, 1:-1] = inst[:, 0].reshape(B, -1)
    dur = torch.from_numpy(dur.reshape(-1, 1))
    r_est, r_lst, r_ms = R.Evaluator().forward(ei, dur, n_j, n_m)
    c_est, c_lst, c_ms = R.CPM_batch_G(renv.current_graphs, dev="cpu")
    for dt in (torch.int64, torch.int32):
        est, lst, ms = Evaluator().forward(ei.cuda(), dur.to(dt).cuda(), n_j, n_m)
        assert np.array_equal(est.cpu().numpy(), r_est.numpy()) and np.array_equal(lst.cpu().numpy(), r_lst.numpy())
        assert np.array_equal(ms.cpu().numpy(), r_ms.numpy())
    # the networkx route of the drop-in: same graphs in, same numbers out (env/message_passing_evl.py:275-287)
    est, lst, ms = CPM_batch_G(renv.current_graphs, dev="cuda")
    assert np.array_equal(est.cpu().numpy(), c_est.numpy()) and np.array_equal(lst.cpu().numpy(), c_lst.numpy())
    assert np.array_equal(ms.cpu().numpy(), c_ms.numpy())
    assert np.array_equal(est.cpu().numpy(), r_est.numpy())


def test_unmodified_test_learned_loop_runs_on_the_dropin():
    """The rollout loop of the reference's test_learned.py:185-189 (BatchGraph.wrapper -> reference Actor ->
    env.step) with the reference's OWN Actor, and `JsspN5`/`BatchGraph` resolved from the drop-in, for 50 steps on
    syn10x10 -- and the same loop on the reference's CPU environment with identically seeded sampling: the two
    runs must produce the same actions and the same incumbents (logits agree, so do the samples)."""
    from l2s_b200 import BatchGraph, JsspN5
    R = ref_harness.load()
    inst = np.load(os.path.join(R.root, "test_data", "syn10x10.npy"))[:8]
    sd = torch.load(os.path.join(R.root, "saved_model", "incumbent_10x10[1,99]_fdd-divide-mwkr_yaoxin_1_128_4_4_"
                                 "gin+dghan_1_0.0_5e-05_10_500_64_128000_10.pth"), map_location="cpu")
    kw = dict(in_dim=3, hidden_dim=128, embedding_l=4, policy_l=4, embedding_type='gin+dghan', heads=1, dropout=0.)

    # (a) reference env + reference actor, CPU
    np.random.seed(1)
    renv = R.JsspN5(n_job=10, n_mch=10, low=1, high=99, reward_type='yaoxin', fea_norm_const=1000)
    pol_cpu = R.Actor(**kw)
    pol_cpu.load_state_dict(sd)
    rstates, rfa, _ = renv.reset(instances=inst, init_type='fdd-divide-mwkr', device='cpu', plot=False)
    init = _mseq_from_state(rstates, inst, 10, 10)

    # (b) the drop-in: same lines as test_learned.py:182-189, device = cuda
    dev = 'cuda'
    env = JsspN5(n_job=10, n_mch=10, low=1, high=99, reward_type='yaoxin', fea_norm_const=1000)
    policy = R.Actor(**kw).to(dev)
    policy.load_state_dict(sd)
    batch_data = BatchGraph()
    states, feasible_actions, _ = env.reset(instances=inst, init_type='fdd-divide-mwkr', device=dev, plot=False,
                                            init_solutions=init)
    rbatch = R.BatchGraph()
    greedy_saved = R.actor_module.Categorical

    class _Greedy(torch.distributions.Categorical):
        def sample(self, sample_shape=torch.Size()):
            return self.probs.argmax(dim=-1)

    R.actor_module.Categorical = _Greedy        # deterministic sampler on both sides (CPU and CUDA RNG streams differ)
    try:
        while env.itr < 50:
            batch_data.wrapper(*states)
            with torch.no_grad():
                actions, _ = policy(batch_data, feasible_actions)
            states, _, feasible_actions, _ = env.step(actions, dev, plot=False)
            rbatch.wrapper(*rstates)
            with torch.no_grad():
                ractions, _ = pol_cpu(rbatch, rfa)
            rstates, _, rfa, _ = renv.step(ractions, 'cpu')
            assert [list(map(int, a)) for a in actions] == [list(map(int, a)) for a in ractions], \
                "greedy actions of the CUDA and CPU rollouts differ at step %d" % env.itr
    finally:
        R.actor_module.Categorical = greedy_saved
    assert np.array_equal(env.incumbent_objs.cpu().numpy(), renv.incumbent_objs.numpy())
    assert np.array_equal(states[0].cpu().numpy(), rstates[0].numpy())
