"""Drop-in `JsspN5` / `BatchGraph` backed by the sm_100a kernels (mirror of the reference's
env/environment.py:19-423 public surface: same constructor, `reset`, `step`, attributes, tensor layouts).

The whole search state lives on the GPU inside an `l2s_handle`; one `step` is one kernel launch that
applies the N5 swaps, re-evaluates every schedule, writes the policy inputs (x, edge_index_mc) and
enumerates the next move sets.  Only the feasible pairs travel back to the host, because the reference API
hands them to the policy as Python lists (model/actor.py:216-226).

There is no CPU path: `device` must be a CUDA device and the CUDA library must be built.
"""
import ctypes as C
from collections.abc import Sequence

import numpy as np
import torch

from . import _capi
from ._capi import L2SError, StepIO


class BatchGraph:
    """Same 4-field holder as the reference (env/environment.py:19-36)."""

    def __init__(self):
        self.x = None
        self.edge_index_pc = None
        self.edge_index_mc = None
        self.batch = None

    def wrapper(self, x, edge_index_pc, edge_index_mc, batch):
        self.x, self.edge_index_pc, self.edge_index_mc, self.batch = x, edge_index_pc, edge_index_mc, batch

    def clean(self):
        self.x = self.edge_index_pc = self.edge_index_mc = self.batch = None


class MoveLists(Sequence):
    """`feasible_actions` of one step: behaves like the reference's `list[B] of list[P_i] of [a0, a1]`
    (env/environment.py:411-423; an instance without moves holds `[[0, 0]]`), but is a lazy view of a snapshot of
    the per-instance host records the step kernel wrote: nothing is allocated per instance until someone asks.
    The reference policy needs `len(fa)`, `len(fa[i])`, `fa[i][j][0|1]` (model/actor.py:216-222); tests compare it with
    plain lists (`==`); `arrays()` hands the same data out as numpy for consumers that want no Python lists at all
    (l2s_b200.policy.Actor.forward).  Building ~10 small lists per instance cost ~4 ms per 8192-instance step."""
    __slots__ = ("_rec", "_lists")

    def __init__(self, records):
        self._rec = records              # [B, 4 + pmax] uint32, private copy
        self._lists = None

    def __len__(self):
        return self._rec.shape[0]

    def _row(self, b):
        row = self._rec[b]
        c = int(row[0]) & 0xffff
        if c == 0:
            return [[0, 0]]
        w = row[_capi.REC_HEADER_WORDS:_capi.REC_HEADER_WORDS + c]
        return np.stack([w & 0xffff, w >> 16], axis=1).tolist()

    def tolist(self):
        """The fully materialised nested list (built once, at C speed when the CPython glue is available)."""
        if self._lists is None:
            glue = _capi.pylists()
            if glue is not None:
                self._lists = glue.l2s_build_move_lists(self._rec.ctypes.data, self._rec.shape[1], self._rec.shape[0])
            else:
                self._lists = [self._row(b) for b in range(len(self))]
        return self._lists

    def __getitem__(self, b):
        if self._lists is not None or isinstance(b, slice):
            return self.tolist()[b]
        n = self._rec.shape[0]
        if b < -n or b >= n:
            raise IndexError("instance index out of range")
        return self._row(b)

    def __iter__(self):
        return iter(self.tolist())

    def __eq__(self, other):
        if isinstance(other, MoveLists):
            other = other.tolist()
        return self.tolist() == other

    def __ne__(self, other):
        return not self.__eq__(other)

    __hash__ = None

    def __repr__(self):
        return "MoveLists(%r)" % (self.tolist(),)

    def counts(self):
        """int32 [B]: number of feasible pairs per instance (0 where the reference returns [[0, 0]])."""
        return (self._rec[:, 0] & 0xffff).astype(np.int32)

    def arrays(self):
        """(pairs int32 [B, pmax, 2], npairs int32 [B]); rows are valid up to npairs[b], zero beyond."""
        cnt = self.counts()
        w = self._rec[:, _capi.REC_HEADER_WORDS:]
        valid = np.arange(w.shape[1])[None, :] < cnt[:, None]
        w = np.where(valid, w, 0)
        return np.stack([w & 0xffff, w >> 16], axis=2).astype(np.int32), cnt


def _stream_ptr(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _np_view(ptr, shape, dtype):
    n = int(np.prod(shape))
    buf = (C.c_byte * (n * np.dtype(dtype).itemsize)).from_address(C.addressof(ptr.contents))
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


def _raise_status(status, instance, where):
    """Map device status codes to the exceptions the reference raises at the same place."""
    if status == _capi.STATUS["SHORT_PATH"]:
        raise IndexError("index 2 is out of bounds: critical path of two operations on one machine "
                         "(instance %d; reference env/environment.py:89-90)" % instance)
    if status == _capi.STATUS["ILLEGAL_ACTION"]:
        try:
            import networkx as nx
            exc = nx.NetworkXError
        except Exception:  # pragma: no cover
            exc = L2SError
        if exc is L2SError:
            raise L2SError(status, where, instance)
        raise exc("The edge of the action given for instance %d is not in the graph "
                  "(not an adjacent pair on one machine)." % instance)
    raise L2SError(status, where, instance)


class JsspN5:
    """N5 local-search environment for batches of JSSP instances (reference: env/environment.py:39-423)."""

    def __init__(self, n_job, n_mch, low, high, reward_type='yaoxin', fea_norm_const=1000,
                 evaluator_type='message-passing', pmax=None):
        self.n_job = n_job
        self.n_mch = n_mch
        self.n_oprs = n_job * n_mch
        self.low = low
        self.high = high
        self.itr = 0
        self.instances = None
        self.current_objs = None
        self.incumbent_objs = None
        self.tabu_size = 1
        self.reward_type = reward_type
        self.fea_norm_const = fea_norm_const
        # 'message-passing' and 'CPM' give identical numbers in the reference (env/message_passing_evl.py
        # :367-371); both are served by the same kernel here.
        self.evaluator_type = evaluator_type
        # capacity of the per-instance move list; grown automatically if a reset finds more N5 pairs
        self.pmax = pmax if pmax else (32 if n_job <= 30 else (64 if n_job <= 60 else 128))
        self.instance_offset = 0      # global index of the first instance (multi-GPU sharding)
        self._h = None
        self._key = None
        self._device = None
        self._const = None
        self._h_rec = self._h_actions = self._h_actions_ptr = None
        self._fa = self._flag = None
        self._lib = _capi.lib()

    # -- handle management -----------------------------------------------------------------------------
    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def close(self):
        """Destroy the handle.  The numpy views of its pinned host buffers die with it."""
        if self._h is not None:
            self._lib.l2s_destroy(self._h)
            self._h = None
        self._key = None
        self._h_rec = self._h_actions = self._h_actions_ptr = None
        self._fa = self._flag = None

    def _live(self, what):
        if self._h is None:
            raise RuntimeError("JsspN5.%s: no live environment -- call reset() first (or the handle was closed)" % what)

    def _ensure_handle(self, B, device):
        device = torch.device(device)
        if device.type != 'cuda':
            raise RuntimeError("l2s_b200 runs on CUDA devices only (no CPU fallback); got device=%r" % (device,))
        if device.index is None:
            device = torch.device('cuda', torch.cuda.current_device())
        key = (B, device.index)
        if self._key != key:
            self.close()
            h = C.c_void_p()
            _capi.check(self._lib.l2s_create(self.n_job, self.n_mch, B, self.pmax, device.index, C.byref(h)),
                        "l2s_create")
            self._h, self._key, self._device = h, key, device
            rec, stride, act = _capi.c_u32p(), C.c_int(0), _capi.c_i32p()
            _capi.check(self._lib.l2s_host_records(h, C.byref(rec), C.byref(stride), C.byref(act)),
                        "l2s_host_records")
            # pinned host records the step kernel writes (include/l2s_b200.h): word 0 = npairs | done << 16 |
            # status << 24, words 4.. = a0 | a1 << 16
            self._h_rec = _np_view(rec, (B, stride.value), np.uint32)
            self._h_actions = _np_view(act, (B, 2), np.int32)
            self._h_actions_ptr = C.cast(act, C.c_void_p)
            n1 = self.n_oprs + 2
            # static tensors, cached: conjunctive edges in the reference's order (env/environment.py:62-69,
            # 301) and the node->graph map (:314)
            first = torch.arange(self.n_job, dtype=torch.int64) * self.n_mch + 1
            u = torch.arange(1, self.n_oprs + 1, dtype=torch.int64)
            v = torch.where(u % self.n_mch != 0, u + 1, torch.full_like(u, self.n_oprs + 1))
            pc = torch.stack([torch.cat([torch.zeros(self.n_job, dtype=torch.int64), u]), torch.cat([first, v])])
            off = (torch.arange(B, dtype=torch.int64) * n1).view(B, 1, 1)
            epc = (pc.unsqueeze(0) + off).permute(1, 0, 2).reshape(2, -1).contiguous().to(device)
            batch = torch.arange(B, dtype=torch.int64).repeat_interleave(n1).to(device)
            self._const = (epc, batch)
        return device

    # -- reference API ---------------------------------------------------------------------------------
    def reset(self, instances, init_type, device, plot=False, init_solutions=None):
        """instances: np.ndarray [B, 2, n_j, n_m] (durations, 1-based machines).  Returns
        ((x, edge_index_pc, edge_index_mc, batch), feasible_actions, done) like env/environment.py:389-409.
        `init_solutions` [B, n_m, n_j] (extension) installs given machine orders instead of a rule."""
        instances = np.asarray(instances)
        B = instances.shape[0]
        assert instances.shape[1:] == (2, self.n_job, self.n_mch), "instances must be [B,2,n_j,n_m]"
        device = self._ensure_handle(B, device)
        stream = _stream_ptr(device)
        self.instances = instances
        di = torch.from_numpy(np.ascontiguousarray(instances.reshape(B, 2, -1)).astype(np.int32)).to(device)
        dur, mch = di[:, 0].contiguous(), di[:, 1].contiguous()
        _capi.check(self._lib.l2s_load_instances(self._h, _ptr(dur), _ptr(mch), self.instance_offset, stream),
                    "l2s_load_instances")
        self._poll("reset")      # malformed instances are reported before anything is built from them
        if init_solutions is not None:
            sol = torch.from_numpy(np.ascontiguousarray(np.asarray(init_solutions)).astype(np.int32)).to(device)
            assert sol.shape == (B, self.n_mch, self.n_job)
            _capi.check(self._lib.l2s_set_solution(self._h, _ptr(sol), stream), "l2s_set_solution")
        elif init_type == 'plist':
            sol = torch.from_numpy(self._plist_orders(instances)).to(device)
            _capi.check(self._lib.l2s_set_solution(self._h, _ptr(sol), stream), "l2s_set_solution")
        elif init_type in _capi.RULE:
            _capi.check(self._lib.l2s_build_initial(self._h, _capi.RULE[init_type], int(self.high), stream),
                        "l2s_build_initial")
        else:
            assert False, 'Initial solution type = "plist", "spt", "fdd-divide-mwkr".'
        self._poll("reset")
        self.itr = 0
        try:
            state, _, fa, done = self._launch(None, device, is_reset=True)
        except L2SError as e:
            if e.status != _capi.STATUS["PAIR_OVERFLOW"] or self.pmax >= 4096:
                raise
            # more N5 pairs than the move list holds: start over with a larger list (nothing to migrate at reset)
            sol = self.solutions().cpu().numpy()
            self.pmax *= 2
            self.close()
            return self.reset(instances, init_type, device, plot=plot, init_solutions=sol)
        return state, fa, done

    def step(self, actions, device, plot=False):
        """actions: list of B [a0, a1] ([0, 0] = dummy).  Returns (state, reward, feasible_actions, done)
        like env/environment.py:356-387."""
        if self.reward_type not in _capi.REWARD:
            raise ValueError('reward type must be "yaoxin" or "consecutive".')
        self._live("step")
        glue = _capi.pylists()
        if glue is not None and not isinstance(actions, np.ndarray):
            if glue.l2s_parse_actions(actions, self._h_actions.ctypes.data, self._key[0]) != 0:
                raise RuntimeError("could not parse actions")      # (a Python exception is normally already set)
            acts = self._h_actions
        else:
            acts = np.asarray(actions, dtype=np.int32).reshape(-1, 2)
            assert acts.shape[0] == self._key[0], "one action per instance"
        out = self._launch(acts, self._device, is_reset=False)
        self.itr = self.itr + 1
        return out

    def feasible_actions(self, device=None):
        """Move sets of the current solutions (env/environment.py:411-423) from the last launch."""
        return self._fa, self._flag

    @property
    def tabu_lists(self):
        """list[B] of [[a1, a0]] (last move reversed) or [] -- env/environment.py:370-380."""
        if self._h is None:
            return None
        B = self._key[0]
        tabu = torch.empty((B, 2), dtype=torch.int32, device=self._device)
        _capi.check(self._lib.l2s_export_state(self._h, None, None, None, None, _ptr(tabu), None, None, None,
                                               _stream_ptr(self._device)), "l2s_export_state")
        return [[] if (a == 0 and b == 0) else [[a, b]] for a, b in tabu.cpu().tolist()]

    def _nx_graphs(self):
        """networkx views of the current solutions, built on demand (the reference keeps them as its primary state,
        env/environment.py:262-268; here the state lives on the GPU and only the baselines scripts read these,
        conventional_baselines_with_long-term-mem.py:100).  Arc weights = duration of the source node; the in-arcs
        of every node are inserted in the order networkx's `G.pred[v]` has in the reference after the same moves
        (job predecessor first iff the tie-break history bit is set or its id is smaller), so `nx.dag_longest_path`
        on these graphs returns the reference's critical paths."""
        import networkx as nx
        B, n, n_m, n_j = self._key[0], self.n_oprs, self.n_mch, self.n_job
        mseq = self.solutions().cpu().numpy()
        jf = self.export_state()["jobfirst"].cpu().numpy()
        dur = np.asarray(self.instances)[:, 0].reshape(B, -1)
        full, mc = [], []
        for b in range(B):
            mpred = np.zeros(n + 2, dtype=np.int64)
            mpred[mseq[b][:, 1:].reshape(-1)] = mseq[b][:, :-1].reshape(-1)
            d = np.concatenate([[0], dur[b], [0]])
            G = nx.DiGraph()
            G.add_nodes_from(range(n + 2))
            Gm = nx.DiGraph()
            Gm.add_nodes_from(range(n + 2))
            for v in range(1, n + 1):
                jp = v - 1 if (v - 1) % n_m != 0 else 0
                mp = int(mpred[v])
                preds = [u for u in ([jp, mp] if (jf[b, v] or (jp and jp < mp)) else [mp, jp]) if u]
                if not jp:
                    preds.append(0)                       # S -> first operation of a job is (re-)added last (:264)
                for u in preds:
                    G.add_edge(u, v, weight=int(d[u]))
                if mp:
                    Gm.add_edge(mp, v, weight=1)
            for j in range(n_j):
                G.add_edge(j * n_m + n_m, n + 1, weight=int(d[j * n_m + n_m]))
            full.append(G)
            mc.append(Gm)
        return full, mc

    @property
    def current_graphs(self):
        """list[B] of nx.DiGraph: full disjunctive graphs (lazily built, see `_nx_graphs`)."""
        return self._nx_graphs()[0] if self._h is not None else None

    @property
    def sub_graphs_mc(self):
        """list[B] of nx.DiGraph: machine arcs only (lazily built)."""
        return self._nx_graphs()[1] if self._h is not None else None

    def solutions(self):
        """Current machine orders [B, n_m, n_j] int32 (device tensor) -- extension."""
        B = self._key[0]
        out = torch.empty((B, self.n_mch, self.n_job), dtype=torch.int32, device=self._device)
        _capi.check(self._lib.l2s_get_solution(self._h, _ptr(out), _stream_ptr(self._device)), "l2s_get_solution")
        return out

    def export_state(self):
        """Test introspection: integer est/lst [B, n+2], critical paths, tabu, tie-break history."""
        B, dev, n1 = self._key[0], self._device, self.n_oprs + 2
        est = torch.empty((B, n1), dtype=torch.int32, device=dev)
        lst = torch.empty((B, n1), dtype=torch.int32, device=dev)
        path = torch.zeros((B, self.n_oprs), dtype=torch.int32, device=dev)
        plen = torch.zeros((B,), dtype=torch.int32, device=dev)
        tabu = torch.empty((B, 2), dtype=torch.int32, device=dev)
        jf = torch.empty((B, n1), dtype=torch.uint8, device=dev)
        ms = torch.empty((B,), dtype=torch.float32, device=dev)
        inc = torch.empty((B,), dtype=torch.float32, device=dev)
        _capi.check(self._lib.l2s_export_state(self._h, _ptr(est), _ptr(lst), _ptr(path), _ptr(plen), _ptr(tabu),
                                               _ptr(jf), _ptr(ms), _ptr(inc), _stream_ptr(dev)), "l2s_export_state")
        self._poll("export_state")
        return dict(est=est, lst=lst, path=path, path_len=plen, tabu=tabu, jobfirst=jf, makespan=ms, incumbent=inc)

    # -- device-tensor interface: no host round trip per step ---------------------------------
    def reset_device(self, instances, init_type, device, init_solutions=None):
        """Like `reset`, but the move sets stay on the GPU: returns (state, pairs int32 [B,pmax,2],
        npairs int32 [B], done bool [B,1]).  Pair rows are valid up to npairs[b]."""
        self.reset(instances, init_type, device, init_solutions=init_solutions)
        state, _, pairs, npairs, done = self._launch_device(None, is_reset=True)
        return state, pairs, npairs, done

    def step_device(self, actions, out=None):
        """actions: int32 device tensor [B,2] ((0,0) = dummy).  One kernel launch, nothing is copied to the
        host and nothing synchronises; call `check_errors()` whenever convenient (e.g. every few hundred steps).
        Returns (state, reward [B,1], pairs int32 [B,pmax,2], npairs int32 [B], done bool [B,1])."""
        if self.reward_type not in _capi.REWARD:
            raise ValueError('reward type must be "yaoxin" or "consecutive".')
        self._live("step_device")
        assert actions.dtype == torch.int32 and actions.is_cuda and actions.shape == (self._key[0], 2)
        res = self._launch_device(actions.contiguous(), is_reset=False, out=out)
        self.itr = self.itr + 1
        return res

    def check_errors(self):
        """Synchronise and raise if any instance reported a device-side error since the last check."""
        self._poll("step_device")

    def device_buffers(self):
        """Freshly allocated output tensors of one step (pass them as `out=` to `step_device` to write in place,
        e.g. when the policy + step pair is captured in a CUDA graph)."""
        B, device = self._key[0], self._device
        n1 = self.n_oprs + 2
        e_mc = self.n_mch * (self.n_job - 1)
        pm = self._lib.l2s_pmax(self._h)       # (the host records are padded to whole 64-byte lines: not their width)
        return dict(x=torch.empty((B * n1, 3), dtype=torch.float32, device=device),
                    emc=torch.empty((2, B * e_mc), dtype=torch.int64, device=device),
                    ms=torch.empty((B, 1), dtype=torch.float32, device=device),
                    inc=torch.empty((B, 1), dtype=torch.float32, device=device),
                    reward=torch.empty((B, 1), dtype=torch.float32, device=device),
                    done=torch.empty((B, 1), dtype=torch.bool, device=device),
                    pairs=torch.empty((B, pm, 2), dtype=torch.int32, device=device),
                    npairs=torch.empty((B,), dtype=torch.int32, device=device))

    def _launch_device(self, actions, is_reset, out=None):
        device = self._device
        o = out if out is not None else self.device_buffers()
        x, emc, ms, inc, reward, done, pairs, npairs = (o[k] for k in ("x", "emc", "ms", "inc", "reward", "done",
                                                                       "pairs", "npairs"))
        io = StepIO(actions=actions.data_ptr() if actions is not None else None, high=float(self.high),
                    fea_norm_const=float(self.fea_norm_const), reward_type=_capi.REWARD.get(self.reward_type, 0),
                    is_reset=1 if is_reset else 0, x=x.data_ptr(), edge_index_mc=emc.data_ptr(), makespan=ms.data_ptr(),
                    incumbent=inc.data_ptr(), reward=reward.data_ptr(), done=done.data_ptr(), pairs=pairs.data_ptr(),
                    npairs=npairs.data_ptr(), status=None)
        _capi.check(self._lib.l2s_step(self._h, C.byref(io), _stream_ptr(device)), "l2s_step")
        self.current_objs, self.incumbent_objs = ms, inc
        state = (x, self._const[0], emc, self._const[1])
        self._last_device = (state, pairs, npairs, done)
        return state, reward, pairs, npairs, done

    # -- fused rollouts (no host round trip inside the loop) -------------------------------------------
    def run(self, policy, steps, tape=None, seed=0, log_steps=(), return_actions=False):
        """K = `steps` fused steps in ONE launch with an on-device policy ('replay' with tape [B,K,2],
        'random', 'greedy').  Returns (incumbents at log_steps [len(log_steps), B] float32 device tensor,
        actions taken [B,K,2] int32 or None).  Call `observe()` afterwards for the policy inputs."""
        self._live("run")
        B, dev = self._key[0], self._device
        tape_t = None
        if policy == 'replay':
            tape_t = torch.as_tensor(np.asarray(tape), dtype=torch.int32).reshape(B, steps, 2).contiguous().to(dev)
        taken = torch.zeros((B, steps, 2), dtype=torch.int32, device=dev) if return_actions else None
        log_steps = [int(s) for s in log_steps]
        logs = torch.zeros((len(log_steps), B), dtype=torch.float32, device=dev)
        arr = (C.c_int32 * max(1, len(log_steps)))(*log_steps)
        _capi.check(self._lib.l2s_run(self._h, _capi.POLICY[policy], int(steps), _ptr(tape_t), int(seed),
                                      _ptr(taken), arr, len(log_steps), _ptr(logs) if log_steps else None,
                                      _stream_ptr(dev)), "l2s_run")
        self._poll("run")
        self.itr += int(steps)
        return logs, taken

    def observe(self):
        """Re-evaluate without moving (a launch without actions changes neither the incumbent nor the step
        counter): refreshes current/incumbent tensors and returns (state, feasible_actions, done)."""
        self._live("observe")
        state, _, fa, done = self._launch(None, self._device, is_reset=False)
        return state, fa, done

    # -- internals -------------------------------------------------------------------------------------
    def _plist_orders(self, instances):
        """'plist' initial solutions (env/environment.py:140-177,391-393): one random priority list of job
        ids shared by the batch; every machine processes its operations in list order."""
        B = instances.shape[0]
        plist = np.random.permutation(np.arange(self.n_job).repeat(self.n_mch))
        nxt = np.zeros(self.n_job, dtype=np.int64)
        fill = np.zeros((B, self.n_mch), dtype=np.int64)
        out = np.zeros((B, self.n_mch, self.n_job), dtype=np.int32)
        rows = np.arange(B)
        for j in plist:
            m = instances[:, 1, j, nxt[j]].astype(np.int64) - 1
            out[rows, m, fill[rows, m]] = j * self.n_mch + nxt[j] + 1
            fill[rows, m] += 1
            nxt[j] += 1
        return out

    def _poll(self, where):
        inst = C.c_int32(0)
        st = self._lib.l2s_poll_error(self._h, C.byref(inst), _stream_ptr(self._device))
        if st < 0:
            _raise_status(st, inst.value, where)

    def _launch(self, acts, device, is_reset):
        B = self._key[0]
        n1 = self.n_oprs + 2
        e_mc = self.n_mch * (self.n_job - 1)
        x = torch.empty((B * n1, 3), dtype=torch.float32, device=device)
        emc = torch.empty((2, B * e_mc), dtype=torch.int64, device=device)
        ms = torch.empty((B, 1), dtype=torch.float32, device=device)
        inc = torch.empty((B, 1), dtype=torch.float32, device=device)
        reward = torch.empty((B, 1), dtype=torch.float32, device=device)
        done = torch.empty((B, 1), dtype=torch.bool, device=device)
        io = StepIO(actions=None, high=float(self.high), fea_norm_const=float(self.fea_norm_const),
                    reward_type=_capi.REWARD.get(self.reward_type, 0), is_reset=1 if is_reset else 0,
                    x=x.data_ptr(), edge_index_mc=emc.data_ptr(), makespan=ms.data_ptr(), incumbent=inc.data_ptr(),
                    reward=reward.data_ptr(), done=done.data_ptr(), pairs=None, npairs=None, status=None)
        a_ptr = None
        if acts is not None:
            if acts is not self._h_actions:
                self._h_actions[:] = acts
            a_ptr = self._h_actions_ptr
        _capi.check(self._lib.l2s_step_host(self._h, C.byref(io), a_ptr, None, None, None, None, None, None,
                                            _stream_ptr(device)), "l2s_step")
        rec = self._h_rec
        head = rec[:, 0]
        if (head >> 24).any():
            status = (head >> 24).astype(np.uint8).view(np.int8)
            b = int(np.nonzero(status)[0][0])
            # drain the sticky device error word so that the next call starts clean
            self._lib.l2s_poll_error(self._h, None, _stream_ptr(device))
            _raise_status(int(status[b]), b, "step")
        self.current_objs = ms
        self.incumbent_objs = inc
        # the pinned records are overwritten by the next step: snapshot the header + the longest valid pair list
        width = _capi.REC_HEADER_WORDS + max(1, int((head & 0xffff).max()))
        fa = MoveLists(rec[:, :width].copy())
        self._fa = fa
        self._flag = ~done
        return (x, self._const[0], emc, self._const[1]), reward, fa, done
