"""GPU parity: the CUDA environment (through the C-ABI via the Python drop-in) against the golden vectors
generated from the unmodified reference (tests/golden/traj_*.npz, made by oracle/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, golden_names

pytestmark = pytest.mark.gpu


def _pairs_from_fa(fa, pmax):
    B = len(fa)
    cnt = np.zeros(B, dtype=np.int32)
    out = np.zeros((B, pmax, 2), dtype=np.int32)
    for b, f in enumerate(fa):
        if f == [[0, 0]]:
            continue
        cnt[b] = len(f)
        out[b, :len(f)] = np.array(f)
    return out, cnt


@pytest.mark.parametrize("name", golden_names("traj_"))
def test_trajectory_matches_reference(name):
    from l2s_b200 import JsspN5
    g = np.load(os.path.join(GOLDEN, "traj_%s.npz" % name))
    n_j, n_m, high, fea = [int(v) for v in g["meta"]]
    inst = g["instances"]
    B, K = inst.shape[0], g["tape"].shape[1]
    pmax = g["pairs"].shape[2]
    env = JsspN5(n_j, n_m, 1, high, reward_type=str(g["reward_type"]), fea_norm_const=fea)
    state, fa, done = env.reset(inst, str(g["init_type"]), "cuda", init_solutions=g["mseq0"])

    def check(k, state, fa, done):
        p, c = _pairs_from_fa(fa, pmax)
        assert np.array_equal(c, g["npairs"][k]), "npairs step %d" % k
        assert np.array_equal(p, g["pairs"][k]), "pairs step %d" % k
        assert np.array_equal(done.cpu().numpy().reshape(-1), g["done"][k])
        assert np.array_equal(env.current_objs.cpu().numpy().reshape(-1).astype(np.int32), g["makespan"][k])
        assert np.array_equal(env.incumbent_objs.cpu().numpy().reshape(-1), g["incumbent"][k])
        ex = env.export_state()
        assert np.array_equal(ex["est"].cpu().numpy(), g["est"][k]), "est step %d" % k
        assert np.array_equal(ex["lst"].cpu().numpy(), g["lst"][k]), "lst step %d" % k

    check(0, state, fa, done)
    assert state[0].dtype == torch.float32 and state[1].dtype == torch.int64 and state[2].dtype == torch.int64
    assert np.array_equal(state[0].cpu().numpy(), g["x_first"])          # bit-exact float32 features
    assert np.array_equal(state[1].cpu().numpy(), g["epc"])
    assert np.array_equal(state[2].cpu().numpy(), g["emc_first"])
    assert np.array_equal(state[3].cpu().numpy(), g["batch"])
    for k in range(K):
        acts = g["tape"][:, k].tolist()
        state, reward, fa, done = env.step(acts, "cuda")
        assert np.array_equal(reward.cpu().numpy().reshape(-1), g["reward"][k]), "reward step %d" % k
        check(k + 1, state, fa, done)
        assert env.itr == k + 1
    assert np.array_equal(state[0].cpu().numpy(), g["x_last"])
    assert np.array_equal(state[2].cpu().numpy(), g["emc_last"])
    # tabu list = last applied move reversed
    last = g["tape"][:, K - 1]
    tl = env.tabu_lists
    for b in range(B):
        if last[b].any():
            assert tl[b] == [[int(last[b, 1]), int(last[b, 0])]]


def test_evaluator_coo_matches_reference():
    from l2s_b200 import Evaluator
    g = np.load(os.path.join(GOLDEN, "eval_coo.npz"))
    ev = Evaluator()
    for tag in ("6x4", "10x10", "20x15", "30x20"):
        n_j, n_m, B = [int(v) for v in g["shape_" + tag]]
        ei = torch.from_numpy(g["ei_" + tag]).cuda()
        for dt in (torch.int32, torch.int64):
            dur = torch.from_numpy(g["dur_" + tag]).to(dt).cuda()
            est, lst, ms = ev.forward(ei, dur, n_j, n_m)
            assert est.dtype == torch.float32 and est.shape == (B * (n_j * n_m + 2), 1) and ms.shape == (B, 1)
            assert np.array_equal(est.cpu().numpy(), g["est_" + tag])
            assert np.array_equal(lst.cpu().numpy(), g["lst_" + tag])
            assert np.array_equal(ms.cpu().numpy(), g["ms_" + tag])


@pytest.mark.parametrize("name", golden_names("traj_"))
def test_initial_solution_builder(name):
    """l2s_build_initial == the reference's _rules_solver whenever no priority tie occurred (the reference
    breaks ties with np.random.choice), and == the oracle's first-min rule always."""
    from l2s_b200 import JsspN5
    from oracle import l2s_oracle as O
    g = np.load(os.path.join(GOLDEN, "traj_%s.npz" % name))
    n_j, n_m, high, fea = [int(v) for v in g["meta"]]
    inst = g["instances"]
    env = JsspN5(n_j, n_m, 1, high)
    env.reset(inst, str(g["init_type"]), "cuda")
    got = env.solutions().cpu().numpy()
    for b in range(inst.shape[0]):
        want, ties = O.rules_solver(inst[b, 0], inst[b, 1], str(g["init_type"]), high)
        assert np.array_equal(got[b], want), "instance %d differs from the oracle builder" % b
        if g["init_ties"][b] == 0:
            assert np.array_equal(got[b], g["mseq0"][b]), "instance %d differs from the reference reset" % b


@pytest.mark.parametrize("name", ["6x6_ties", "10x10", "20x15", "20x15_ties", "30x20"])
def test_fused_run_replay_and_random(name):
    """l2s_run (K steps, one launch) reproduces the per-step path and the reference trajectory."""
    from l2s_b200 import JsspN5
    from oracle import l2s_oracle as O
    g = np.load(os.path.join(GOLDEN, "traj_%s.npz" % name))
    n_j, n_m, high, fea = [int(v) for v in g["meta"]]
    inst, tape = g["instances"], g["tape"]
    B, K = tape.shape[0], tape.shape[1]
    env = JsspN5(n_j, n_m, 1, high)
    env.reset(inst, "fdd-divide-mwkr", "cuda", init_solutions=g["mseq0"])
    logs, taken = env.run("replay", K, tape=tape, log_steps=[1, K // 2, K], return_actions=True)
    assert np.array_equal(taken.cpu().numpy(), tape)
    inc = g["incumbent"]
    assert np.array_equal(logs.cpu().numpy(), np.stack([inc[1], inc[K // 2], inc[K]]))
    state, fa, done = env.observe()
    assert np.array_equal(state[0].cpu().numpy(), g["x_last"])
    assert np.array_equal(env.current_objs.cpu().numpy().reshape(-1).astype(np.int32), g["makespan"][K])
    assert env.itr == K
    # RANDOM policy against the oracle driven by the same counter-based RNG
    env.instance_offset = 1000
    env.reset(inst, "fdd-divide-mwkr", "cuda", init_solutions=g["mseq0"])
    K2 = min(K, 40)
    logs, taken = env.run("random", K2, seed=12345, log_steps=[K2], return_actions=True)
    oenv = O.OracleEnv(n_j, n_m, 1, high)
    _, ofa, _ = oenv.reset(inst, mseq=g["mseq0"])
    otaken, ologs = O.run_policy(oenv, ofa, "random", K2, seed=12345, inst_offset=1000, log_steps=[K2])
    assert np.array_equal(taken.cpu().numpy(), otaken)
    assert np.array_equal(logs.cpu().numpy(), ologs)


def test_illegal_action_and_errors():
    import networkx as nx
    from l2s_b200 import JsspN5
    g = np.load(os.path.join(GOLDEN, "traj_10x5.npz"))
    inst = g["instances"]
    env = JsspN5(10, 5, 1, 99)
    state, fa, done = env.reset(inst, "fdd-divide-mwkr", "cuda", init_solutions=g["mseq0"])
    acts = [f[0] for f in fa]
    acts[2] = [1, 2]      # same job: never an adjacent machine pair
    with pytest.raises(nx.NetworkXError):
        env.step(acts, "cuda")
    with pytest.raises(AssertionError):
        env.reset(inst, "no-such-rule", "cuda")
    bad = JsspN5(10, 5, 1, 99, reward_type="nope")
    bad.reset(inst, "fdd-divide-mwkr", "cuda")
    with pytest.raises(ValueError):
        bad.step([[0, 0]] * inst.shape[0], "cuda")
    with pytest.raises(RuntimeError):
        JsspN5(10, 5, 1, 99).reset(inst, "fdd-divide-mwkr", "cpu")


def test_dummy_actions_freeze_state():
    from l2s_b200 import JsspN5
    g = np.load(os.path.join(GOLDEN, "traj_10x5.npz"))
    inst = g["instances"]
    env = JsspN5(10, 5, 1, 99)
    s0, fa0, _ = env.reset(inst, "fdd-divide-mwkr", "cuda", init_solutions=g["mseq0"])
    s1, r, fa1, _ = env.step([[0, 0]] * inst.shape[0], "cuda")
    assert torch.equal(s0[0], s1[0]) and torch.equal(s0[2], s1[2]) and fa0 == fa1
    assert float(r.abs().sum()) == 0 and env.itr == 1


def test_plist_initial_solutions_and_nonstrict_evaluator():
    """'plist' initial solutions (env/environment.py:140-177,391-393): one shared random priority list; every
    machine processes its operations in list order.  Checked through the evaluator against the numpy oracle."""
    from l2s_b200 import Evaluator, JsspN5, L2SError
    from oracle import l2s_oracle as O
    g = np.load(os.path.join(GOLDEN, "traj_10x10.npz"))
    inst = g["instances"]
    np.random.seed(3)
    env = JsspN5(10, 10, 1, 99)
    state, fa, done = env.reset(inst, "plist", "cuda")
    sol = env.solutions().cpu().numpy()
    oenv = O.OracleEnv(10, 10, 1, 99)
    ostate, ofa, odone = oenv.reset(inst, mseq=sol)
    assert np.array_equal(state[0].cpu().numpy(), ostate[0]) and np.array_equal(state[2].cpu().numpy(), ostate[2])
    assert [[list(map(int, p)) for p in f] for f in fa] == [[list(map(int, p)) for p in f] for f in ofa]
    # every machine row holds exactly that machine's operations
    mch = inst[:, 1].reshape(len(inst), -1)
    for b in range(len(inst)):
        for m in range(10):
            assert (mch[b][sol[b, m] - 1] == m + 1).all()
    # asynchronous evaluator: errors surface at check()
    ev = Evaluator(strict=False)
    ei = torch.cat([state[1], state[2]], dim=1)
    dur = torch.from_numpy(np.concatenate([O.padded_durations(i[0]) for i in inst]).reshape(-1, 1)).cuda()
    est, lst, ms = ev.forward(ei, dur, 10, 10)
    ev.check()
    assert torch.equal(ms, env.current_objs)
    bad = ei.clone()
    bad[:, -1] = bad[:, -2]
    ev.forward(bad, dur, 10, 10)
    with pytest.raises(L2SError):
        ev.check()
    ev.forward(ei, dur, 10, 10)
    ev.check()          # the status word was cleared by the failed check


def test_device_api_errors_are_reported_asynchronously():
    """step_device never synchronises; an illegal action surfaces at check_errors() / the next host-side step, and
    a replayed tape with an illegal action raises from run()."""
    import networkx as nx
    from l2s_b200 import JsspN5
    g = np.load(os.path.join(GOLDEN, "traj_10x5.npz"))
    inst = g["instances"]
    env = JsspN5(10, 5, 1, 99)
    state, pairs, npairs, done = env.reset_device(inst, "fdd-divide-mwkr", "cuda", init_solutions=g["mseq0"])
    acts = pairs[:, 0, :].clone()
    acts[1] = torch.tensor([1, 2], dtype=torch.int32)          # same job: never an adjacent machine pair
    before = env.solutions().clone()
    env.step_device(acts)
    with pytest.raises(nx.NetworkXError):
        env.check_errors()
    after = env.solutions()
    assert torch.equal(before[1], after[1])                     # the offending instance was left untouched
    env.check_errors()                                          # error word was drained
    env2 = JsspN5(10, 5, 1, 99)
    env2.reset(inst, "fdd-divide-mwkr", "cuda", init_solutions=g["mseq0"])
    tape = g["tape"][:, :5].copy()
    tape[3, 2] = [7, 7]
    with pytest.raises(nx.NetworkXError):
        env2.run("replay", 5, tape=tape)


def test_plain_c_host_through_the_c_abi(tmp_path):
    """examples/capi_demo.c: a C program that links libl2s_b200.so and drives reset + steps with host buffers gives
    the same makespans / move counts / features as the Python drop-in on the same instances."""
    import re
    import subprocess
    from conftest import ROOT
    from l2s_b200 import JsspN5, build
    exe = str(tmp_path / "capi_demo")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cmd = ["gcc", "-O2", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(cuda, "include"),
           os.path.join(ROOT, "examples", "capi_demo.c"), "-o", exe, "-L", build.LIB_DIR, "-ll2s_b200",
           "-L", os.path.join(cuda, "lib64"), "-lcudart", "-Wl,-rpath," + build.LIB_DIR]
    subprocess.run(cmd, check=True, capture_output=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    assert out.strip().endswith("ok")
    dur = np.array([[5, 3, 7, 2, 8, 4, 6, 1, 9, 3, 3, 5], [4, 4, 2, 9, 1, 6, 3, 7, 5, 2, 8, 6]]).reshape(2, 4, 3)
    mch = np.array([[1, 2, 3, 2, 1, 3, 3, 1, 2, 1, 3, 2], [2, 3, 1, 1, 2, 3, 3, 2, 1, 2, 1, 3]]).reshape(2, 4, 3)
    env = JsspN5(4, 3, 1, 99)
    state, fa, done = env.reset(np.stack([dur, mch], axis=1), "fdd-divide-mwkr", "cuda")
    lines = out.strip().split("\n")
    m = re.match(r"reset: makespans (\d+) (\d+), moves (\d+) (\d+)", lines[0])
    assert [int(v) for v in m.groups()[:2]] == env.current_objs.reshape(-1).int().tolist()
    for k in range(6):
        acts = [f[0] for f in fa]
        state, r, fa, done = env.step(acts, "cuda")
        m = re.match(r"step \d+: actions \((\d+),(\d+)\) \((\d+),(\d+)\) -> makespans (\d+) (\d+), rewards (\d+) (\d+), moves (\d+) (\d+)",
                     lines[1 + k])
        g = [int(v) for v in m.groups()]
        assert [g[0:2], g[2:4]] == acts
        assert g[4:6] == env.current_objs.reshape(-1).int().tolist() and g[6:8] == r.reshape(-1).int().tolist()
        assert g[8:10] == [0 if f == [[0, 0]] else len(f) for f in fa]
    x1 = [float(v) for v in lines[-2].split("=")[1].split()]
    assert np.allclose(x1, state[0][1].cpu().numpy(), atol=1e-6)


@pytest.mark.parametrize("records,zerocopy", [("0", "1"), ("1", "0"), ("0", "0")])
def test_staged_host_paths_match_the_zero_copy_default(records, zerocopy, monkeypatch):
    """L2S_HOST_RECORDS=0 (records via device memory + one bulk D2H) and L2S_HOST_ZEROCOPY=0 (actions via an H2D
    copy) give the same trajectory as the default path where the kernel reads / writes pinned host memory itself."""
    from l2s_b200 import JsspN5
    g = np.load(os.path.join(GOLDEN, "traj_20x15_ties.npz"))
    n_j, n_m, high, fea = [int(v) for v in g["meta"]]
    pmax = g["pairs"].shape[2]
    monkeypatch.setenv("L2S_HOST_RECORDS", records)
    monkeypatch.setenv("L2S_HOST_ZEROCOPY", zerocopy)
    env = JsspN5(n_j, n_m, 1, high, reward_type=str(g["reward_type"]), fea_norm_const=fea)   # reads the env at create
    state, fa, done = env.reset(g["instances"], str(g["init_type"]), "cuda", init_solutions=g["mseq0"])
    for k in range(g["tape"].shape[1]):
        state, reward, fa, done = env.step(g["tape"][:, k].tolist(), "cuda")
        p, c = _pairs_from_fa(fa, pmax)
        assert np.array_equal(c, g["npairs"][k + 1]) and np.array_equal(p, g["pairs"][k + 1])
        assert np.array_equal(reward.cpu().numpy().reshape(-1), g["reward"][k])
        assert np.array_equal(done.cpu().numpy().reshape(-1), g["done"][k + 1])
        assert np.array_equal(env.incumbent_objs.cpu().numpy().reshape(-1), g["incumbent"][k + 1])


def test_step_host_reads_pinned_caller_memory_in_place():
    """l2s_step_host with the actions in the caller's own pinned host memory (read in place by the kernel) gives the
    same records as with pageable memory (staged through the handle's pinned buffer)."""
    import ctypes as C
    from l2s_b200 import JsspN5, _capi
    g = np.load(os.path.join(GOLDEN, "traj_10x10.npz"))
    n_j, n_m, high, fea = [int(v) for v in g["meta"]]
    lib = _capi.lib()
    recs = []
    for pinned in (True, False):
        env = JsspN5(n_j, n_m, 1, high, reward_type=str(g["reward_type"]), fea_norm_const=fea)
        state, fa, done = env.reset(g["instances"], str(g["init_type"]), "cuda", init_solutions=g["mseq0"])
        B = g["instances"].shape[0]
        x = torch.empty_like(state[0])
        emc = torch.empty_like(state[2])
        io = _capi.StepIO(actions=None, high=float(high), fea_norm_const=float(fea),
                          reward_type=_capi.REWARD[str(g["reward_type"])], is_reset=0, x=x.data_ptr(),
                          edge_index_mc=emc.data_ptr(), makespan=None, incumbent=None, reward=None, done=None,
                          pairs=None, npairs=None, status=None)
        sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        out = []
        for k in range(g["tape"].shape[1]):
            acts = np.ascontiguousarray(g["tape"][:, k]).astype(np.int32)
            keep = torch.from_numpy(acts).pin_memory() if pinned else None
            ptr = keep.data_ptr() if pinned else acts.ctypes.data
            _capi.check(lib.l2s_step_host(env._h, C.byref(io), ptr, None, None, None, None, None, None, sp), "step")
            rec = env._h_rec.copy()
            npairs = rec[:, 0] & 0xffff
            rec[:, 4:][np.arange(rec.shape[1] - 4)[None, :] >= npairs[:, None]] = 0     # stale words beyond npairs
            out.append((rec, x.cpu().numpy().copy()))
            assert np.array_equal(npairs, g["npairs"][k + 1])
            assert np.array_equal(rec[:, 1].view(np.float32).astype(np.int32), g["makespan"][k + 1])
        recs.append(out)
    for (r0, x0), (r1, x1) in zip(*recs):
        assert np.array_equal(r0, r1) and np.array_equal(x0, x1)


def test_step_host_begin_wait_pipelines_independent_environments():
    """Two environments driven by one host thread through l2s_step_host_begin / l2s_step_host_wait with both steps in
    flight on the same stream: every step's records and device outputs equal the golden trajectory (= the blocking
    call); a second begin before wait, and a wait without begin, are rejected."""
    import ctypes as C
    from l2s_b200 import JsspN5, _capi
    g = np.load(os.path.join(GOLDEN, "traj_10x10.npz"))
    n_j, n_m, high, fea = [int(v) for v in g["meta"]]
    lib = _capi.lib()
    B = g["instances"].shape[0]
    half = B // 2
    sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    envs, ios, xs, pins = [], [], [], []
    for sl in (slice(0, half), slice(half, B)):
        env = JsspN5(n_j, n_m, 1, high, reward_type=str(g["reward_type"]), fea_norm_const=fea)
        state, fa, done = env.reset(g["instances"][sl], str(g["init_type"]), "cuda", init_solutions=g["mseq0"][sl])
        x = torch.empty_like(state[0])
        emc = torch.empty_like(state[2])
        nb = len(range(*sl.indices(B)))
        ms = torch.empty((nb, 1), device="cuda")
        ios.append(_capi.StepIO(actions=None, high=float(high), fea_norm_const=float(fea),
                                reward_type=_capi.REWARD[str(g["reward_type"])], is_reset=0, x=x.data_ptr(),
                                edge_index_mc=emc.data_ptr(), makespan=ms.data_ptr(), incumbent=None, reward=None,
                                done=None, pairs=None, npairs=None, status=None))
        envs.append(env)
        xs.append((x, emc, ms, sl))
    assert lib.l2s_step_host_wait(envs[0]._h, None, None, None, None, None, None) < 0          # nothing begun
    for k in range(g["tape"].shape[1]):
        keep = [torch.from_numpy(np.ascontiguousarray(g["tape"][sl, k]).astype(np.int32)).pin_memory() for *_, sl in xs]
        for e in (0, 1):
            _capi.check(lib.l2s_step_host_begin(envs[e]._h, C.byref(ios[e]), keep[e].data_ptr(), sp), "begin")
        assert lib.l2s_step_host_begin(envs[0]._h, C.byref(ios[0]), keep[0].data_ptr(), sp) < 0   # one in flight per handle
        for e in (1, 0):                                                          # (waited for out of order)
            nb = xs[e][2].shape[0]
            npairs = np.zeros(nb, dtype=np.int32)
            mk = np.zeros(nb, dtype=np.float32)
            _capi.check(lib.l2s_step_host_wait(envs[e]._h, None, npairs.ctypes.data, None, None, mk.ctypes.data, None),
                        "wait")
            x, emc, ms, sl = xs[e]
            assert np.array_equal(npairs, g["npairs"][k + 1][sl])
            assert np.array_equal(mk.astype(np.int32), g["makespan"][k + 1][sl])
        torch.cuda.synchronize()
        for x, emc, ms, sl in xs:
            assert np.array_equal(ms.cpu().numpy().reshape(-1).astype(np.int32), g["makespan"][k + 1][sl])


def test_plist_reset_matches_reference_golden():
    """a-14: reset(init_type='plist') under the reference run's numpy seed reproduces its first state, move sets and a
    40-step trajectory (tests/golden/plist_10x5.npz, generated from the unmodified reference)."""
    from l2s_b200 import JsspN5
    g = np.load(os.path.join(GOLDEN, "plist_10x5.npz"))
    n_j, n_m = int(g["meta"][0]), int(g["meta"][1])
    env = JsspN5(n_j, n_m, 1, 99)
    np.random.seed(int(g["seed"]))
    state, fa, done = env.reset(g["instances"], "plist", "cuda")
    assert np.array_equal(env.solutions().cpu().numpy(), g["mseq0"])
    assert np.array_equal(state[0].cpu().numpy(), g["x_first"]) and np.array_equal(state[2].cpu().numpy(), g["emc_first"])
    assert np.array_equal(env.current_objs.cpu().numpy().reshape(-1), g["makespan_first"])
    p, c = _pairs_from_fa(fa, 32)
    assert np.array_equal(p, g["pairs_first"]) and np.array_equal(c, g["npairs_first"])
    for k in range(g["tape"].shape[1]):
        state, reward, fa, done = env.step(g["tape"][:, k].tolist(), "cuda")
        assert np.array_equal(env.current_objs.cpu().numpy().reshape(-1).astype(np.int32), g["makespan"][k])
        p, c = _pairs_from_fa(fa, 32)
        assert np.array_equal(p, g["pairs"][k]) and np.array_equal(c, g["npairs"][k])
    assert np.array_equal(state[0].cpu().numpy(), g["x_last"])


@pytest.mark.parametrize("cta_warps", ["2", "4", "8"])
def test_evaluator_cta_kernel_matches_reference(cta_warps, monkeypatch):
    """The CTA-per-graph evaluator kernel (the default for graphs of which fewer than 16 fit an SM) forced onto the
    small golden graphs with 2 / 4 / 8 warps per graph, lean and latency sweep: same est / lst / makespan as the
    reference's Evaluator.forward."""
    from l2s_b200 import Evaluator
    g = np.load(os.path.join(GOLDEN, "eval_coo.npz"))
    monkeypatch.setenv("L2S_EVAL_CTA", cta_warps)
    for fast in ("0", "1", "2"):
        monkeypatch.setenv("L2S_FAST_SWEEP", fast)
        ev = Evaluator()
        for tag in ("6x4", "10x10", "20x15", "30x20"):
            n_j, n_m, B = [int(v) for v in g["shape_" + tag]]
            ei = torch.from_numpy(g["ei_" + tag]).cuda()
            for dt in (torch.int32, torch.int64):
                dur = torch.from_numpy(g["dur_" + tag]).to(dt).cuda()
                est, lst, ms = ev.forward(ei, dur, n_j, n_m)
                assert np.array_equal(est.cpu().numpy(), g["est_" + tag]), (tag, fast)
                assert np.array_equal(lst.cpu().numpy(), g["lst_" + tag]), (tag, fast)
                assert np.array_equal(ms.cpu().numpy(), g["ms_" + tag]), (tag, fast)
