"""Drop-in `Evaluator` (reference: env/message_passing_evl.py:156-199) backed by the sm_100a kernels.

`Evaluator().forward(edge_index, duration, n_j, n_m)` takes the same COO edge list (any edge order,
conjunctive + machine arcs of a batch of complete selections, graphs offset by n_j*n_m+2) and durations
([N,1] int32/int64, 0 for S and T) and returns `(est [N,1], lst [N,1], make_span [B,1])` as float32 tensors
holding the same integer values the reference's iterated max-aggregation converges to.
"""
import ctypes as C

import torch

from . import _capi
from ._capi import L2SError


class Evaluator:
    """`strict=True` (default) reads the device status word back after every call and raises on malformed input
    (one 4-byte D2H + sync; the reference synchronises 2*depth times per call).  `strict=False` keeps the call
    fully asynchronous like any CUDA op; call `check()` when convenient to surface a recorded error."""

    def __init__(self, strict=True):
        self._lib = _capi.lib()
        self._ws = None
        self._flag = None
        self._sticky = None
        self.strict = strict

    def check(self):
        if self._sticky is not None:
            flag = int(self._sticky.item())
            self._sticky.zero_()
            if flag < 0:
                raise L2SError(flag, "Evaluator.forward")

    def forward(self, edge_index, duration, n_j, n_m):
        if not edge_index.is_cuda or not duration.is_cuda:
            raise RuntimeError("l2s_b200.Evaluator runs on CUDA tensors only (no CPU fallback)")
        device = edge_index.device
        n1 = n_j * n_m + 2
        N = duration.shape[0]
        if N % n1 != 0:
            raise ValueError("duration has %d rows, not a multiple of n_j*n_m+2 = %d" % (N, n1))
        B = N // n1
        ei = edge_index.contiguous()
        if ei.dtype != torch.int64:
            ei = ei.to(torch.int64)
        dur = duration.reshape(-1).contiguous()
        if dur.dtype not in (torch.int32, torch.int64):
            dur = dur.to(torch.int64)
        need = self._lib.l2s_eval_coo_workspace_bytes(n_j, n_m, B)
        _capi.check(need, "l2s_eval_coo_workspace_bytes")
        if self._ws is None or self._ws.numel() < need or self._ws.device != device:
            self._ws = torch.zeros(need, dtype=torch.uint8, device=device)   # every call hands it back zero-filled
            self._flag = torch.zeros(1, dtype=torch.int32, device=device)
            self._sticky = torch.zeros(1, dtype=torch.int32, device=device)
        if self.strict:
            self._flag.zero_()
        est = torch.empty((N, 1), dtype=torch.float32, device=device)
        lst = torch.empty((N, 1), dtype=torch.float32, device=device)
        ms = torch.empty((B, 1), dtype=torch.float32, device=device)
        stream = C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
        _capi.check(self._lib.l2s_eval_coo(
            C.c_void_p(ei.data_ptr()), ei.shape[1], C.c_void_p(dur.data_ptr()), 1 if dur.dtype == torch.int64 else 0,
            n_j, n_m, B, C.c_void_p(est.data_ptr()), C.c_void_p(lst.data_ptr()), C.c_void_p(ms.data_ptr()),
            C.c_void_p(self._ws.data_ptr()), self._ws.numel(), 1,
            C.c_void_p((self._flag if self.strict else self._sticky).data_ptr()),
            device.index if device.index is not None else torch.cuda.current_device(), stream), "l2s_eval_coo")
        if self.strict:
            flag = int(self._flag.item())   # the reference synchronises 2*depth times per call; we do it once
            if flag < 0:
                raise L2SError(flag, "Evaluator.forward")
        return est, lst, ms


_cpm_evaluator = None


def CPM_batch_G(Gs, dev):
    """Drop-in for the reference's networkx critical-path evaluator (env/message_passing_evl.py:275-287):
    `Gs` is a list of `nx.DiGraph` disjunctive graphs (nodes 0..n+1, S = 0, T = n+1, every arc weighted with the
    duration of its source node, as `JsspN5._init_solver` builds them, env/environment.py:262-265); returns
    `(est [N,1], lst [N,1], makespan [B,1])` float32 on `dev`.  The reference walks each graph in topological
    order in Python; here the arcs are flattened into one COO list and evaluated by the same CUDA kernels as
    `Evaluator.forward` (the two evaluators of the reference agree, `:367-371`, so the numbers are identical).
    `n_j` is read off the graph (out-degree of S), `n_m = n / n_j`; all graphs of a batch must share the shape."""
    global _cpm_evaluator
    import numpy as np
    device = torch.device(dev)
    if device.type != 'cuda':
        raise RuntimeError("l2s_b200.CPM_batch_G runs on CUDA devices only (no CPU fallback); got dev=%r" % (dev,))
    n1 = Gs[0].number_of_nodes()
    n_j = Gs[0].out_degree(0)
    n_m = (n1 - 2) // n_j
    eis, durs = [], []
    for k, G in enumerate(Gs):
        if G.number_of_nodes() != n1 or G.out_degree(0) != n_j:
            raise ValueError("all graphs of a batch must have the same n_j x n_m shape")
        e = np.fromiter((v for u, w, d in G.edges(data='weight') for v in (u, w, d)), dtype=np.int64).reshape(-1, 3)
        dur = np.zeros(n1, dtype=np.int64)
        dur[e[:, 0]] = e[:, 2]                      # every arc carries its source's duration
        eis.append(e[:, :2].T + k * n1)
        durs.append(dur)
    ei = torch.from_numpy(np.ascontiguousarray(np.concatenate(eis, axis=1))).to(device)
    dur = torch.from_numpy(np.concatenate(durs)).reshape(-1, 1).to(device)
    if _cpm_evaluator is None:
        _cpm_evaluator = Evaluator()
    return _cpm_evaluator.forward(ei, dur, n_j, n_m)
