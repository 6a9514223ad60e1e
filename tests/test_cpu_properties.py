"""CPU-only domain properties of the oracle (the GPU parity suite compares the CUDA path against it) and of the
host-record format of the C ABI.  Seeded random instances; no reference needed at run time."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

sys.path.insert(0, ROOT)
from oracle import l2s_oracle as O   # noqa: E402


def _instance(rng, n_j, n_m, hi):
    dur = rng.randint(1, hi + 1, size=(n_j, n_m))
    mch = rng.random_sample((n_j, n_m)).argsort(axis=1) + 1
    mseq = O.rules_solver(dur, mch, "fdd-divide-mwkr", max(hi, 99))[0]
    return O.OracleInstance(dur, mch, mseq)


@pytest.mark.parametrize("seed", range(6))
def test_schedule_invariants_along_a_random_search(seed):
    """At every step of a random N5 search: the critical path is a chain of tight arcs from time 0 to the makespan,
    est <= lst with equality exactly... on the path, every offered pair is adjacent on one machine AND consecutive on
    the path, the tabu move is never offered, and swapping a pair back restores the machine order."""
    rng = np.random.RandomState(100 + seed)
    n_j, n_m = int(rng.randint(3, 13)), int(rng.randint(2, 9))
    inst = _instance(rng, n_j, n_m, int(rng.choice([3, 9, 99])))
    n = n_j * n_m
    for step in range(25):
        est, lst, ms, dur1 = inst.est, inst.lst, inst.makespan, inst.dur1
        assert est[n + 1] == ms and lst[n + 1] == ms and est[0] == 0
        assert (lst[1:n + 1] >= est[1:n + 1]).all()
        path = [int(v) for v in inst.path()]
        assert est[path[0]] == 0 and est[path[-1]] + dur1[path[-1]] == ms
        for u, v in zip(path[:-1], path[1:]):
            assert est[u] + dur1[u] == est[v]                                   # tight arc
            assert v == u + 1 or inst.msucc[u] == v                             # job arc or machine arc
        assert all(est[v] == lst[v] for v in path)                              # zero slack on the critical path
        pairs = [tuple(int(x) for x in p) for p in inst.moves()]
        pos = {v: i for i, v in enumerate(path)}
        for a0, a1 in pairs:
            assert inst.msucc[a0] == a1 and inst.mch_flat[a0 - 1] == inst.mch_flat[a1 - 1]
            assert pos[a1] == pos[a0] + 1
            assert inst.tabu is None or (a0, a1) != tuple(inst.tabu)
        if not pairs:
            break
        a0, a1 = pairs[int(rng.randint(len(pairs)))]
        before = inst.mseq.copy()
        mseq, mp, msu, jf = inst.mseq.copy(), inst.mpred.copy(), inst.msucc.copy(), inst.jobfirst.copy()
        O.apply_move(mseq, mp, msu, jf, inst.mch_flat, a0, a1)
        O.apply_move(mseq, mp, msu, jf, inst.mch_flat, a1, a0)
        assert np.array_equal(mseq, before)                                     # the swap is an involution
        assert inst.neighbour_makespan(a0, a1) >= max(dur1[1:n + 1].reshape(n_j, n_m).sum(axis=1).max(), 1)
        inst.apply(a0, a1)
        inst.evaluate()
        assert sorted(inst.mseq.reshape(-1).tolist()) == list(range(1, n + 1))  # still a permutation of the operations


def test_host_record_macros_match_the_python_decoding(tmp_path):
    """include/l2s_b200.h: L2S_REC_NPAIRS / DONE / STATUS decode word 0 of a host record exactly like the Python
    drop-in does (l2s_b200/environment.py)."""
    src = tmp_path / "rec.c"
    src.write_text(
        '#include <stdio.h>\n#include "l2s_b200.h"\n'
        "int main(void) {\n"
        "  const uint32_t w[] = {0u, 7u, 0x10000u, 0xfb010000u, 0xfc000012u, 0xf600ffffu, 0xff00002au};\n"
        "  for (unsigned i = 0; i < sizeof(w) / sizeof(w[0]); ++i)\n"
        '    printf("%u %u %d\\n", (unsigned)L2S_REC_NPAIRS(w[i]), (unsigned)L2S_REC_DONE(w[i]), L2S_REC_STATUS(w[i]));\n'
        '  printf("%d\\n", L2S_REC_HEADER_WORDS);\n  return 0;\n}\n')
    exe = tmp_path / "rec"
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    words = np.array([0, 7, 0x10000, 0xfb010000, 0xfc000012, 0xf600ffff, 0xff00002a], dtype=np.uint32)
    npairs = words & 0xffff
    done = (words >> 16) & 1
    status = (words >> 24).astype(np.uint8).view(np.int8)
    for i in range(len(words)):
        assert [int(v) for v in out[i].split()] == [int(npairs[i]), int(done[i]), int(status[i])]
    from l2s_b200 import _capi
    assert int(out[len(words)]) == _capi.REC_HEADER_WORDS


def test_move_lists_behave_like_the_reference_nested_lists():
    """`MoveLists` (the lazy `feasible_actions` of JsspN5.step) against plain nested lists built from the same host
    records: len, indexing (incl. negative / slices), iteration, equality both ways, [[0, 0]] for instances without
    moves, and the numpy view `arrays()`; the C glue and the pure-Python route must agree."""
    import numpy as np
    from l2s_b200 import _capi
    from l2s_b200.environment import MoveLists
    rng = np.random.RandomState(0)
    B, pmax = 37, 12
    stride = 16
    rec = np.zeros((B, stride), dtype=np.uint32)
    want = []
    for b in range(B):
        c = 0 if b % 5 == 0 else int(rng.randint(1, pmax + 1))
        pairs = rng.randint(1, 400, size=(c, 2))
        rec[b, 0] = c | ((1 << 16) if c == 0 else 0)
        rec[b, 4:4 + c] = pairs[:, 0] | (pairs[:, 1] << 16)
        rec[b, 4 + c:] = 0xdeadbeef                     # stale words behind the valid prefix must be ignored
        want.append(pairs.tolist() if c else [[0, 0]])
    fa = MoveLists(rec.copy())
    assert len(fa) == B and fa[3] == want[3] and fa[-1] == want[-1] and fa[0] == [[0, 0]]
    assert len(fa[7]) == len(want[7]) and fa[7][0][1] == want[7][0][1]
    assert fa == want and want == fa.tolist() and not (fa != want)
    assert [f for f in fa] == want and fa[2:5] == want[2:5]
    assert MoveLists(rec.copy()) == fa
    with pytest.raises(IndexError):
        fa[B]
    pairs, cnt = fa.arrays()
    assert pairs.shape == (B, stride - 4, 2) and cnt.tolist() == [0 if w == [[0, 0]] else len(w) for w in want]
    for b in range(B):
        assert pairs[b, :cnt[b]].tolist() == (want[b] if cnt[b] else [])
        assert not pairs[b, cnt[b]:].any()
    lazy = MoveLists(rec.copy())
    assert [lazy._row(b) for b in range(B)] == want       # pure-Python route == C glue route
    if _capi.pylists() is not None:
        assert lazy.tolist() == want


def test_n5_bit_logic_equals_reference_case_analysis():
    """The kernels enumerate N5 moves with `take_i = s_i & ~(s_(i-1) & s_(i+1))` (s_i: path operations i, i+1 share a
    machine; sentinels s_-1 = s_rg = 1).  Exhaustive check over all machine-label sequences of length 2..9 over 3
    labels against the reference's case analysis (JsspN5._get_pairs, env/environment.py:84-105, restated in the oracle),
    including the IndexError corner (two operations on one machine)."""
    import itertools
    from oracle import l2s_oracle as O

    def formula(cb):
        rg = len(cb) - 1
        s = [1] + [1 if cb[i] == cb[i + 1] else 0 for i in range(rg)] + [1]
        return [i for i in range(rg) if s[i + 1] and not (s[i] and s[i + 2])]

    def reference(cb):
        out, rg = [], len(cb) - 1
        for i in range(rg):
            if cb[i] == cb[i + 1]:
                if i == 0:
                    if cb[i + 1] != cb[i + 2]:
                        out.append(i)
                elif cb[i] != cb[i - 1]:
                    out.append(i)
                elif i + 1 == rg:
                    pass
                elif cb[i + 1] != cb[i + 2]:
                    out.append(i)
        return out

    checked = 0
    for n in range(2, 10):
        for cb in itertools.product(range(3), repeat=n):
            try:
                want = reference(cb)
            except IndexError:
                assert n == 2 and cb[0] == cb[1]
                continue
            assert formula(cb) == want, cb
            checked += 1
    assert checked > 25000
    # and the oracle's own pair enumeration agrees with the formula on a path with real node ids
    path, mch = [5, 9, 2, 7, 11, 3], [1, 1, 1, 2, 2, 1]
    pairs = O.get_pairs(path, mch, None) if hasattr(O, "get_pairs") else None
    if pairs is not None:
        assert [[int(a), int(b)] for a, b in pairs] == [[path[i], path[i + 1]] for i in formula(mch)]


def test_vendored_reference_is_importable_on_its_own(tmp_path):
    """`baseline/_ref` (oracle/vendor_ref.py) is what the GPU box has instead of /root/reference: a fresh interpreter
    pointed at it alone must import the reference (on the repo's torch_geometric drop-in), reset, step and run the
    reference Actor with the shipped weights."""
    from oracle import vendor_ref
    if vendor_ref.vendor() is None:
        pytest.skip("no reference tree and no vendored copy")
    code = r'''
import os, sys
sys.path.insert(0, os.environ["L2S_ROOT"])
import numpy as np, torch
from oracle import ref_harness as H
assert H.REFERENCE_ROOT == os.environ["L2S_REFERENCE_ROOT"], H.REFERENCE_ROOT
R = H.load()
inst = np.load(os.path.join(R.root, "test_data", "syn10x10.npy"))[:3]
np.random.seed(1)
env = R.JsspN5(10, 10, 1, 99)
state, fa, done = env.reset(inst, "fdd-divide-mwkr", "cpu")
pol = R.Actor(3, 128, embedding_l=4, policy_l=4, embedding_type="gin+dghan", heads=1, dropout=0.0)
pol.load_state_dict(torch.load(os.path.join(R.root, "saved_model", "incumbent_10x10[1,99]_fdd-divide-mwkr_yaoxin_1_128_4_4_gin+dghan_1_0.0_5e-05_10_500_64_128000_10.pth"), map_location="cpu"))
bg = R.BatchGraph(); bg.wrapper(*state)
acts, lp = pol(bg, fa)
state, r, fa, done = env.step(acts, "cpu")
res, _ = R.Greedy_baselines(inst[:1], 3, [3], "cpu")
print("OK", env.itr, float(env.current_objs.sum()), np.asarray(res).shape)
'''
    env = dict(os.environ, L2S_ROOT=ROOT, L2S_REFERENCE_ROOT=vendor_ref.DST)
    res = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "OK 1" in res.stdout, res.stderr[-2000:]


def test_step_io_layout_and_enums_match_the_header(tmp_path):
    """The ctypes mirror of `l2s_step_io` (l2s_b200/_capi.py) has the C struct's size and field offsets, and the
    Python-side status / reward / rule / policy codes are the header's enumerators (compiled with gcc from
    include/l2s_b200.h)."""
    import ctypes as C
    from l2s_b200 import _capi
    fields = [f[0] for f in _capi.StepIO._fields_]
    enums = (["L2S_ERR_" + k if k != "OK" else "L2S_OK" for k in _capi.STATUS]
             + ["L2S_REWARD_" + k.upper() for k in _capi.REWARD]
             + ["L2S_RULE_" + k.upper().replace("-", "_") for k in _capi.RULE]
             + ["L2S_POLICY_" + k.upper() for k in _capi.POLICY])
    src = tmp_path / "io.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "l2s_b200.h"\nint main(void) {\n'
        '  printf("%zu\\n", sizeof(l2s_step_io));\n'
        + "".join('  printf("%%zu\\n", offsetof(l2s_step_io, %s));\n' % f for f in fields)
        + "".join('  printf("%%d\\n", (int)%s);\n' % e for e in enums)
        + "  return 0;\n}\n")
    exe = tmp_path / "io"
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = [int(v) for v in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    assert out[0] == C.sizeof(_capi.StepIO)
    assert out[1:1 + len(fields)] == [getattr(_capi.StepIO, f).offset for f in fields]
    want = list(_capi.STATUS.values()) + list(_capi.REWARD.values()) + list(_capi.RULE.values()) + list(_capi.POLICY.values())
    assert out[1 + len(fields):] == want
