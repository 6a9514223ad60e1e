"""TEST INFRASTRUCTURE ONLY -- generate `tests/golden/*.npz` from the UNMODIFIED reference.

Run in the build container (needs /root/reference):   python -m oracle.make_golden [--greedy]
Every fixture is produced by the reference's own `JsspN5.reset/step`, `Evaluator.forward` and
`Greedy_baselines` (imported through `oracle/ref_harness.py`); while generating, the numpy oracle
(`oracle/l2s_oracle.py`) is asserted equal to the reference on everything that is stored.
"""
import argparse
import os
import random
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import l2s_oracle as O  # noqa: E402
from oracle import ref_harness  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
REF_ROOT = ref_harness.REFERENCE_ROOT


def mseq_from_emc(emc_local, mch, n_j, n_m):
    """Machine order [n_m,n_j] from one instance's machine edges (chains followed from their heads)."""
    n = n_j * n_m
    msucc = np.zeros(n + 2, dtype=np.int64)
    mpred = np.zeros(n + 2, dtype=np.int64)
    for u, v in emc_local.T:
        msucc[u], mpred[v] = v, u
    mseq = np.zeros((n_m, n_j), dtype=np.int64)
    flat = np.asarray(mch).reshape(-1)
    for v in range(1, n + 1):
        if mpred[v] == 0:
            m, k, u = flat[v - 1] - 1, 0, v
            while u:
                mseq[m, k] = u
                k += 1
                u = msucc[u]
            assert k == n_j
    return mseq


def ref_initial_orders(R, instances, init_type, seed):
    n_j, n_m = instances.shape[2], instances.shape[3]
    np.random.seed(seed)
    random.seed(seed)
    env = R.JsspN5(n_j, n_m, 1, 99)
    state, fa, done = env.reset(instances, init_type, "cpu")
    n1 = n_j * n_m + 2
    emc = state[2].numpy()
    B = instances.shape[0]
    E = emc.shape[1] // B
    mseq0 = np.stack([mseq_from_emc(emc[:, b * E:(b + 1) * E] - b * n1, instances[b, 1], n_j, n_m) for b in range(B)])
    return env, state, fa, done, mseq0


def pad_pairs(fa, pmax):
    B = len(fa)
    cnt = np.zeros(B, dtype=np.int32)
    out = np.zeros((B, pmax, 2), dtype=np.int32)
    for b, f in enumerate(fa):
        f = [[int(p[0]), int(p[1])] for p in f]
        if f == [[0, 0]]:
            continue
        cnt[b] = len(f)
        out[b, :len(f)] = np.array(f)
    return out, cnt


def trajectory(R, name, n_j, n_m, B, K, lo, hi, seed, reward_type="yaoxin", init_type="fdd-divide-mwkr", pmax=32):
    np.random.seed(seed)
    instances = np.array([R.uni_instance_gen(n_j, n_m, lo, hi) for _ in range(B)]).astype(np.int32)
    env, state, fa, done, mseq0 = ref_initial_orders(R, instances, init_type, seed)
    env.reward_type = reward_type
    oenv = O.OracleEnv(n_j, n_m, 1, 99, reward_type=reward_type)
    ostate, ofa, odone = oenv.reset(instances, mseq=mseq0)
    n1 = n_j * n_m + 2
    rec = dict(makespan=[], incumbent=[], done=[], pairs=[], npairs=[], est=[], lst=[], reward=[], tape=[])

    def record(state, fa, done):
        x = state[0].numpy()
        assert np.array_equal(x, ostate[0]) and np.array_equal(state[2].numpy(), ostate[2])
        assert [[list(map(int, p)) for p in f] for f in fa] == [[list(map(int, p)) for p in f] for f in ofa]
        assert np.array_equal(done.numpy(), odone)
        rec["makespan"].append(env.current_objs.numpy().reshape(-1).astype(np.int32))
        rec["incumbent"].append(env.incumbent_objs.numpy().reshape(-1).copy())
        rec["done"].append(done.numpy().reshape(-1).copy())
        p, c = pad_pairs(fa, pmax)
        rec["pairs"].append(p)
        rec["npairs"].append(c)
        rec["est"].append(np.stack([i.est for i in oenv.insts]).astype(np.int32))
        rec["lst"].append(np.stack([i.lst for i in oenv.insts]).astype(np.int32))
        # the oracle's integer est/lst reproduce the reference's features bit for bit (asserted above)

    record(state, fa, done)
    x_first, emc_first = state[0].numpy().copy(), state[2].numpy().copy()
    rnd = random.Random(seed)
    for k in range(K):
        acts = [list(map(int, rnd.choice(f))) for f in fa]
        state, r, fa, done = env.step(acts, "cpu")
        ostate, orr, ofa, odone = oenv.step(acts)
        assert np.array_equal(r.numpy(), orr)
        rec["tape"].append(np.array(acts, dtype=np.int32))
        rec["reward"].append(r.numpy().reshape(-1).copy())
        record(state, fa, done)
    own = [O.rules_solver(instances[b, 0], instances[b, 1], init_type, 99) for b in range(B)]
    ties = np.array([t for _, t in own], dtype=np.int32)
    for b in range(B):
        if ties[b] == 0:
            assert np.array_equal(own[b][0], mseq0[b]), "oracle initial solution differs from the reference without a tie"
    np.savez_compressed(
        os.path.join(OUT, "traj_%s.npz" % name),
        instances=instances, mseq0=mseq0.astype(np.int32), init_ties=ties,
        tape=np.stack(rec["tape"], axis=1), makespan=np.stack(rec["makespan"]),
        incumbent=np.stack(rec["incumbent"]), done=np.stack(rec["done"]), pairs=np.stack(rec["pairs"]),
        npairs=np.stack(rec["npairs"]), est=np.stack(rec["est"]), lst=np.stack(rec["lst"]),
        reward=np.stack(rec["reward"]), x_first=x_first, x_last=state[0].numpy(), emc_first=emc_first,
        emc_last=state[2].numpy(), epc=state[1].numpy(), batch=state[3].numpy(),
        meta=np.array([n_j, n_m, 99, 1000], dtype=np.int64), reward_type=np.array(reward_type),
        init_type=np.array(init_type))
    print("traj", name, "B", B, "K", K, "init ties", int(ties.sum()), "max pairs", int(np.stack(rec["npairs"]).max()))


def evaluator_coo(R):
    """Reference Evaluator.forward on shuffled COO edges of random schedules (int32 and int64 durations)."""
    import torch
    rng = np.random.RandomState(7)
    out = {}
    for tag, (n_j, n_m, B) in {"6x4": (6, 4, 5), "10x10": (10, 10, 4), "20x15": (20, 15, 3), "30x20": (30, 20, 2)}.items():
        np.random.seed(11)
        instances = np.array([R.uni_instance_gen(n_j, n_m, 1, 99) for _ in range(B)]).astype(np.int32)
        _, state, _, _, _ = ref_initial_orders(R, instances, "spt", 3)
        ei = np.concatenate([state[1].numpy(), state[2].numpy()], axis=1)
        ei = ei[:, rng.permutation(ei.shape[1])]
        dur = np.concatenate([O.padded_durations(i[0]) for i in instances]).reshape(-1, 1)
        est, lst, ms = R.Evaluator().forward(torch.from_numpy(ei), torch.from_numpy(dur), n_j, n_m)
        e2, l2, m2 = O.evaluate_coo_jacobi(ei, dur, n_j, n_m)
        assert np.array_equal(est.numpy().reshape(-1), e2) and np.array_equal(lst.numpy().reshape(-1), l2)
        assert np.array_equal(ms.numpy().reshape(-1), m2)
        out.update({"ei_" + tag: ei, "dur_" + tag: dur.astype(np.int32), "est_" + tag: est.numpy(),
                    "lst_" + tag: lst.numpy(), "ms_" + tag: ms.numpy(), "shape_" + tag: np.array([n_j, n_m, B])})
    np.savez_compressed(os.path.join(OUT, "eval_coo.npz"), **out)
    print("eval_coo ok")


def greedy(R, ds, log_steps, seed=1):
    """Reference greedy policy (conventional_baselines_with_long-term-mem.py:154-221) run here, plus the
    reference's own stored result row for the same data set."""
    instances = np.load(os.path.join(REF_ROOT, "test_data", ds + ".npy"))
    n_j, n_m = instances.shape[2], instances.shape[3]
    _, _, _, _, mseq0 = ref_initial_orders(R, instances, "fdd-divide-mwkr", seed)
    random.seed(seed)
    np.random.seed(seed)
    res, _ = R.Greedy_baselines(instances, max(log_steps), list(log_steps), "cpu")
    stored = np.load(os.path.join(REF_ROOT, "test_results", "conventional_results", "greedy-policy",
                                  ds + "_fdd-divide-mwkr_result.npy"))
    if 500 in log_steps:
        same = np.array_equal(res[list(log_steps).index(500)], stored[0])
        print("greedy", ds, "500-step row equals the reference's stored array:", same)
    np.savez_compressed(os.path.join(OUT, "greedy_%s.npz" % ds), instances=instances.astype(np.int32),
                        mseq0=mseq0.astype(np.int32), log_steps=np.array(log_steps), incumbents=res,
                        stored_rows=stored, stored_steps=np.array([500, 1000, 2000, 5000]))


def plist(R, name="plist_10x5", n_j=10, n_m=5, B=6, K=40, seed=7):
    """`init_type='plist'` (JsspN5._p_list_solver, env/environment.py:140-200, :391-393): one random priority list
    (drawn from np.random at reset) shared by the batch.  Stores the reference's initial machine orders, first state
    and a K-step trajectory."""
    np.random.seed(21)
    instances = np.array([R.uni_instance_gen(n_j, n_m, 1, 99) for _ in range(B)]).astype(np.int32)
    np.random.seed(seed)
    env = R.JsspN5(n_j, n_m, 1, 99)
    state, fa, done = env.reset(instances, "plist", "cpu")
    n1 = n_j * n_m + 2
    emc = state[2].numpy()
    E = emc.shape[1] // B
    mseq0 = np.stack([mseq_from_emc(emc[:, b * E:(b + 1) * E] - b * n1, instances[b, 1], n_j, n_m) for b in range(B)])
    oenv = O.OracleEnv(n_j, n_m, 1, 99)
    ostate, ofa, odone = oenv.reset(instances, mseq=mseq0)
    assert np.array_equal(state[0].numpy(), ostate[0]) and np.array_equal(emc, ostate[2])
    x_first, ms_first = state[0].numpy().copy(), env.current_objs.numpy().reshape(-1).copy()
    p0, c0 = pad_pairs(fa, 32)
    rnd = random.Random(seed)
    tape, ms, pairs, npairs = [], [], [], []
    for k in range(K):
        acts = [list(map(int, rnd.choice(f))) for f in fa]
        state, r, fa, done = env.step(acts, "cpu")
        ostate, orr, ofa, odone = oenv.step(acts)
        assert np.array_equal(state[0].numpy(), ostate[0]) and np.array_equal(r.numpy(), orr)
        tape.append(np.array(acts, dtype=np.int32))
        ms.append(env.current_objs.numpy().reshape(-1).astype(np.int32))
        p, c = pad_pairs(fa, 32)
        pairs.append(p)
        npairs.append(c)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), instances=instances, mseq0=mseq0.astype(np.int32),
                        seed=np.array(seed), x_first=x_first, emc_first=emc, makespan_first=ms_first, pairs_first=p0,
                        npairs_first=c0, tape=np.stack(tape, axis=1), makespan=np.stack(ms), pairs=np.stack(pairs),
                        npairs=np.stack(npairs), x_last=state[0].numpy(), meta=np.array([n_j, n_m, 99, 1000]))
    print("plist", name, "makespans", ms_first)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--greedy", action="store_true", help="also run the (slow) reference greedy baseline")
    ap.add_argument("--only-greedy", action="store_true")
    ap.add_argument("--only-plist", action="store_true")
    args = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    R = ref_harness.load()
    if args.only_plist:
        plist(R)
        return
    if not args.only_greedy:
        trajectory(R, "6x6_ties", 6, 6, 8, 120, 1, 4, 1)
        trajectory(R, "10x10_ties", 10, 10, 6, 100, 1, 3, 2)
        trajectory(R, "10x5", 10, 5, 8, 100, 1, 99, 3)
        trajectory(R, "8x3_consecutive", 8, 3, 6, 100, 1, 99, 6, reward_type="consecutive")
        trajectory(R, "10x10", 10, 10, 8, 100, 1, 99, 8)
        trajectory(R, "15x10_spt", 15, 10, 4, 60, 1, 99, 9, init_type="spt")
        trajectory(R, "20x15_ties", 20, 15, 3, 60, 1, 5, 4)
        trajectory(R, "20x15", 20, 15, 4, 60, 1, 99, 5)
        trajectory(R, "30x20", 30, 20, 2, 30, 1, 99, 10)
        trajectory(R, "50x20", 50, 20, 1, 12, 1, 99, 12)
        evaluator_coo(R)
        plist(R)
    if args.greedy or args.only_greedy:
        greedy(R, "ft6x6", [10, 50, 100, 500])
        greedy(R, "la10x5", [10, 50, 100, 500])
        greedy(R, "tai20x15", [10, 50, 100, 500])


if __name__ == "__main__":
    main()
