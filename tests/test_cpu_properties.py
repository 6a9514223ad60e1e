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
