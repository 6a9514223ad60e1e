"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper of the C oracle (oracle/l2s_oracle.c)."""
import ctypes as C

import numpy as np

from . import build_c

_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build_c.build())
        L.orc_run_batch.restype = C.c_int
        L.orc_run_batch.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                    C.c_void_p, C.c_uint64, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_void_p,
                                    C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_int]
        L.orc_rules_solver.restype = C.c_int
        L.orc_rules_solver.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data if a is not None else None


POLICY = {"replay": 0, "random": 1, "greedy": 2}


class CEnv:
    """Batch of instances stepped by the C oracle.  State (machine orders, tie-break history, tabu, incumbent)
    is kept in numpy arrays between calls, so `run` can be called repeatedly."""

    def __init__(self, instances, mseq, pmax=32, high=99.0, fea=1000.0, inst_offset=0, threads=0):
        inst = np.asarray(instances)
        self.B, _, self.n_j, self.n_m = inst.shape
        self.n = self.n_j * self.n_m
        self.dur = np.ascontiguousarray(inst[:, 0].reshape(self.B, -1), dtype=np.int32)
        self.mch = np.ascontiguousarray(inst[:, 1].reshape(self.B, -1), dtype=np.int32)
        self.mseq = np.array(np.asarray(mseq).reshape(self.B, -1), dtype=np.int32, order='C', copy=True)   # mutated in place by C
        self.pmax, self.high, self.fea, self.inst_offset, self.threads = pmax, high, fea, inst_offset, threads
        self.jobfirst = np.zeros((self.B, self.n + 2), dtype=np.uint8)
        self.tabu = np.zeros((self.B, 2), dtype=np.int32)
        self.incumbent = np.zeros(self.B, dtype=np.float32)
        self.makespan = np.zeros(self.B, dtype=np.float32)
        self.itr = 0
        self.run("replay", 0, tape=np.zeros((self.B, 0, 2), np.int32))

    def run(self, policy, K, tape=None, seed=0, log_steps=(), write_state=False, return_actions=True):
        B, n1 = self.B, self.n + 2
        taken = np.zeros((B, K, 2), dtype=np.int32) if return_actions else None
        logs = np.zeros((len(log_steps), B), dtype=np.float32)
        ls = np.asarray(list(log_steps), dtype=np.int32)
        tape = None if tape is None else np.ascontiguousarray(tape, dtype=np.int32)
        self.pairs = np.zeros((B, self.pmax, 2), dtype=np.int32)
        self.npairs = np.zeros(B, dtype=np.int32)
        self.est = np.zeros((B, n1), dtype=np.int32)
        self.lst = np.zeros((B, n1), dtype=np.int32)
        self.x = np.zeros((B * n1, 3), dtype=np.float32) if write_state else None
        self.emc = np.zeros((2, B * self.n_m * (self.n_j - 1)), dtype=np.int64) if write_state else None
        st = lib().orc_run_batch(B, self.n_j, self.n_m, _p(self.dur), _p(self.mch), _p(self.mseq), POLICY[policy], K,
                                 _p(tape), seed, self.inst_offset, self.itr, self.pmax, _p(taken), _p(ls), len(ls),
                                 _p(logs), _p(self.makespan), _p(self.incumbent), 1 if write_state else 0,
                                 self.high, self.fea, _p(self.x), _p(self.emc), _p(self.pairs), _p(self.npairs),
                                 _p(self.est), _p(self.lst), _p(self.jobfirst), _p(self.tabu), self.threads)
        self.itr += K
        self.status = st
        return taken, logs


def rules_solver(dur, mch, rule="fdd-divide-mwkr", high=99):
    dur = np.ascontiguousarray(dur, dtype=np.int32)
    mch = np.ascontiguousarray(mch, dtype=np.int32)
    n_j, n_m = dur.shape
    out = np.zeros((n_m, n_j), dtype=np.int32)
    ties = C.c_int32(0)
    lib().orc_rules_solver(n_j, n_m, _p(dur), _p(mch), {"spt": 0, "fdd-divide-mwkr": 1}[rule], high, _p(out), C.byref(ties))
    return out, ties.value
