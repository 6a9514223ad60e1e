"""TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the L2S evaluator + N5 local-search step.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this file, and only as the checker / timed CPU baseline.  The product path (`l2s_b200`) never
imports it and has no CPU fallback.

Parity status: the reference (zcaicaros/L2S) has no unit tests for this path ("parity unpinned" in the
formal sense).  This restatement is pinned instead by (i) outputs of the unmodified
reference run in the build container and committed as fixtures under `tests/golden/` by
`oracle/make_golden.py` (est/lst/makespan/x/critical-path move sets/rewards along replayed
trajectories, initial machine orders of `reset`), and (ii) the reference's stored deterministic
greedy-policy result arrays (`test_results/conventional_results/greedy-policy/*_result.npy`).

Conventions (same as the reference): an instance is `[2, n_j, n_m]` (durations, 1-based machine ids);
node ids are S=0, operations 1..n row-major by (job, position), T=n+1; a solution is the processing
order on every machine, `mseq[m, k]` = node id of the k-th operation on machine m (0-based m).

Every function cites the reference lines it restates (paths relative to the reference root).
"""
import numpy as np

MASK64 = (1 << 64) - 1


# ----------------------------------------------------------------------------------------------
# graph helpers
# ----------------------------------------------------------------------------------------------
def machine_links(mseq, n_nodes):
    """mpred/msucc arrays (0 = none) from a machine order.  Machine arcs ops[k] -> ops[k+1]:
    env/environment.py:254-259."""
    mpred = np.zeros(n_nodes, dtype=np.int64)
    msucc = np.zeros(n_nodes, dtype=np.int64)
    for row in np.asarray(mseq):
        for a, b in zip(row[:-1], row[1:]):
            msucc[a] = b
            mpred[b] = a
    return mpred, msucc


def padded_durations(dur_mat):
    """[n_j, n_m] -> [n+2] with 0 for S and T: env/environment.py:299."""
    return np.concatenate([[0], np.asarray(dur_mat).reshape(-1), [0]]).astype(np.int64)


def evaluate(dur1, mpred, msucc, n_j, n_m):
    """Earliest starts, latest starts, makespan of one disjunctive graph.

    Same quantities as Evaluator.forward (env/message_passing_evl.py:161-199), which iterates
    max-aggregation message passing until every node is 'ready'; the fixed point is the longest-path
    recurrence below, evaluated here once in topological order (Kahn).
    Returns integer arrays est[n+2], lst[n+2] and int makespan; raises on a cyclic selection.
    """
    n = n_j * n_m
    T = n + 1
    est = np.zeros(n + 2, dtype=np.int64)
    indeg = np.zeros(n + 2, dtype=np.int64)
    for v in range(1, n + 1):
        indeg[v] = (1 if (v - 1) % n_m != 0 else 0) + (1 if mpred[v] else 0)
    stack = [v for v in range(1, n + 1) if indeg[v] == 0]
    order = []
    while stack:
        v = stack.pop()
        order.append(v)
        fin = est[v] + dur1[v]
        for s in ((v + 1) if v % n_m != 0 else 0, msucc[v]):
            if s:
                if fin > est[s]:
                    est[s] = fin
                indeg[s] -= 1
                if indeg[s] == 0:
                    stack.append(s)
    if len(order) != n:
        raise ValueError("machine order induces a cycle")
    ms = max(est[j * n_m + n_m] + dur1[j * n_m + n_m] for j in range(n_j))
    est[T] = ms
    lst = np.zeros(n + 2, dtype=np.int64)
    lst[T] = ms
    for v in reversed(order):
        best = lst[v + 1] if v % n_m != 0 else ms
        if msucc[v] and lst[msucc[v]] < best:
            best = lst[msucc[v]]
        lst[v] = best - dur1[v]
    lst[0] = min(lst[j * n_m + 1] for j in range(n_j))
    return est, lst, int(ms)


def evaluate_coo_jacobi(edge_index, duration, n_j, n_m):
    """Literal numpy restatement of Evaluator.forward (env/message_passing_evl.py:161-199) for a batch
    given as COO edges [2,E] (any order) and durations [N] -- used to pin the drop-in `Evaluator`.

    forward:  est <- max over in-edges of (dur[src] + (est[src] if ready[src] else 0)), 0 if no in-edge
              notready <- max over in-edges of notready[src]            (lines 176-181)
    backward: same on reversed edges with negated latest starts, T pinned to -makespan (187-197).
    Returns float32 est[N], lst[N], makespan[B] like the reference.
    """
    ei = np.asarray(edge_index).astype(np.int64)
    dur = np.asarray(duration).reshape(-1).astype(np.float32)
    N = dur.shape[0]
    n1 = n_j * n_m + 2
    B = N // n1
    src, dst = ei[0], ei[1]
    S = np.arange(B) * n1
    T = S + n1 - 1

    def agg(values, at, frm):
        out = np.full(N, -np.inf, dtype=np.float32)
        np.maximum.at(out, at, values[frm])
        out[np.isinf(out)] = 0
        return out

    est = np.zeros(N, dtype=np.float32)
    notready = np.ones(N, dtype=np.float32)
    notready[S] = 0
    for _ in range(N):
        if notready.sum() == 0:
            break
        x = dur + np.where(notready > 0, 0, est).astype(np.float32)
        est = agg(x, dst, src)
        notready = agg(notready, dst, src)
    ms = est[T].copy()
    lst = -np.ones(N, dtype=np.float32)
    lst[T] = -ms
    notready = np.ones(N, dtype=np.float32)
    notready[T] = 0
    for _ in range(N):
        if notready.sum() == 0:
            break
        x = np.where(notready > 0, 0, lst).astype(np.float32)
        lst = agg(x, src, dst) + dur
        lst[T] = -ms
        notready = agg(notready, src, dst)
    return est, np.abs(lst), ms


def critical_path(est, dur1, mpred, jobfirst, n_j, n_m):
    """Operations of the longest S->T path == nx.dag_longest_path(G)[1:-1] (env/environment.py:77).

    networkx keeps, per node, the FIRST predecessor (in G.pred[v] insertion order) that attains the
    longest distance.  Insertion order restated: T lists job-last ops by ascending
    job; an operation lists its predecessors ascending by node id (from_numpy_matrix is row-major,
    env/environment.py:263) until a swap re-adds its machine in-edge, after which the job predecessor
    comes first (`jobfirst`, set in `apply_move`); the S edge of a job-first op is added last (:264).
    """
    n = n_j * n_m
    ms = int(est[n + 1])
    v = 0
    for j in range(n_j):
        last = j * n_m + n_m
        if est[last] + dur1[last] == ms:
            v = last
            break
    path = [v]
    while True:
        jp = v - 1 if (v - 1) % n_m != 0 else 0
        mp = int(mpred[v])
        if jp and mp:
            cand = [jp, mp] if (jobfirst[v] or jp < mp) else [mp, jp]
        else:
            cand = [c for c in (jp, mp) if c]
        # a job-first op also has the weight-0 S edge, inserted last: a predecessor must beat length 0
        # (always true for a machine predecessor when durations are >= 1, the supported precondition)
        best_len, best = (0, 0) if jp == 0 else (-1, 0)
        for u in cand:
            if est[u] + dur1[u] > best_len:
                best_len, best = est[u] + dur1[u], u
        if best == 0:
            break
        path.append(int(best))
        v = int(best)
    path.reverse()
    return path


def n5_pairs(path, mch_flat, tabu):
    """N5 block-boundary pairs in path order, tabu-filtered: JsspN5._get_pairs
    (env/environment.py:84-105).  `mch_flat[v-1]` = machine of op v; `tabu` = (a, b) or None.
    Raises IndexError exactly where the reference does (2-op path on one machine, :89-90)."""
    cb = [int(mch_flat[v - 1]) for v in path]
    rg = len(path) - 1
    out = []
    for i in range(rg):
        if cb[i] != cb[i + 1]:
            continue
        if i == 0:
            if i + 2 >= len(cb):
                raise IndexError("critical path of two operations on one machine (reference quirk)")
            take = cb[1] != cb[2]
        elif cb[i] != cb[i - 1]:
            take = True
        elif i + 1 == rg:
            take = False
        else:
            take = cb[i + 1] != cb[i + 2]
        if take and (tabu is None or (path[i], path[i + 1]) != (int(tabu[0]), int(tabu[1]))):
            out.append([path[i], path[i + 1]])
    return out


def apply_move(mseq, mpred, msucc, jobfirst, mch_flat, a0, a1):
    """N5 swap of the adjacent machine pair (a0 -> a1): JsspN5.change_nxgraph_topology
    (env/environment.py:318-354).  Mutates all arrays in place; raises ValueError on a non-adjacent
    pair (the reference fails inside networkx remove_edge, :348)."""
    if msucc[a0] != a1 or a0 == 0:
        raise ValueError("illegal N5 action (%d,%d): not an adjacent machine pair" % (a0, a1))
    s, t = int(mpred[a0]), int(msucc[a1])
    if s:
        msucc[s] = a1
    mpred[a1] = s
    msucc[a1] = a0
    mpred[a0] = a1
    msucc[a0] = t
    if t:
        mpred[t] = a0
        jobfirst[t] = True
    jobfirst[a0] = True
    jobfirst[a1] = True
    m = int(mch_flat[a0 - 1]) - 1
    row = mseq[m]
    p = int(np.nonzero(row == a0)[0][0])
    row[p], row[p + 1] = a1, a0


def features(dur1, est, lst, high, fea_norm_const):
    """x = [dur/high, est/c, lst/c] as IEEE float32 divisions: env/environment.py:312."""
    d = dur1.astype(np.float32) / np.float32(high)
    e = est.astype(np.float32) / np.float32(fea_norm_const)
    l = lst.astype(np.float32) / np.float32(fea_norm_const)
    return np.stack([d, e, l], axis=-1).astype(np.float32)


def edge_index_pc(n_j, n_m):
    """Static conjunctive edges of one instance in the reference's order (nonzero of the row-major
    adjacency, env/environment.py:62-69,301): S->first ops, then u->u+1 / last->T ascending."""
    n = n_j * n_m
    src = [0] * n_j
    dst = [j * n_m + 1 for j in range(n_j)]
    for u in range(1, n + 1):
        src.append(u)
        dst.append(u + 1 if u % n_m != 0 else n + 1)
    return np.array([src, dst], dtype=np.int64)


def edge_index_mc(msucc, n):
    """Machine edges ascending by source node: env/environment.py:300-302."""
    src = [u for u in range(1, n + 1) if msucc[u]]
    return np.array([src, [int(msucc[u]) for u in src]], dtype=np.int64).reshape(2, -1)


# ----------------------------------------------------------------------------------------------
# initial solutions: JsspN5._rules_solver + permissibleLeftShift
# ----------------------------------------------------------------------------------------------
def _left_shift_insert(a, dur, mch, starts, ops, n_m, high):
    """Place op `a` (0-based id) on its machine with permissible left shift.
    Restates env/permissible_LS.py:5-78 on one machine row (`starts`, `ops` mutated in place)."""
    flat_dur = dur.reshape(-1)
    flat_mch = mch.reshape(-1)
    m = flat_mch[a] - 1
    row_s, row_o = starts[m], ops[m]
    # job-ready / machine-ready times (permissible_LS.py:60-78)
    job_rdy = 0
    if a % n_m != 0:
        jp = a - 1
        mj = flat_mch[jp] - 1
        job_rdy = int(starts[mj][np.nonzero(ops[mj] == jp)[0][0]] + flat_dur[jp])
    placed = np.nonzero(row_o >= 0)[0]
    mch_rdy = int(row_s[placed[-1]] + flat_dur[row_o[placed[-1]]]) if len(placed) else 0
    d = int(flat_dur[a])
    later = np.nonzero(job_rdy < row_s)[0]                     # ops already starting after job_rdy (:13)
    if len(later):
        p0 = later[0]
        # end of the op before the first candidate slot; index -1 wraps to the (unfilled) row tail
        # exactly as in the reference (permissible_LS.py:41)
        before = int(row_s[p0 - 1] + flat_dur[row_o[p0 - 1]])
        gap_open = [max(job_rdy, before)] + [int(row_s[p] + flat_dur[row_o[p]]) for p in later[:-1]]
        for i, p in enumerate(later):
            if d <= int(row_s[p]) - gap_open[i]:                # first idle gap that fits (:44-45,50-53)
                row_s[p + 1:] = row_s[p:-1].copy()
                row_o[p + 1:] = row_o[p:-1].copy()
                row_s[p], row_o[p] = gap_open[i], a
                return gap_open[i], True
    k = np.nonzero(row_s == -high)[0][0]                       # append (:28-35)
    row_s[k], row_o[k] = max(job_rdy, mch_rdy), a
    return int(row_s[k]), False


def rules_solver(dur, mch, rule="fdd-divide-mwkr", high=99, rng=None):
    """Initial machine order by priority dispatching: JsspN5._rules_solver (env/environment.py:203-255).

    rule 'spt' (min duration, :227-231) or 'fdd-divide-mwkr' (min cumulative-duration / remaining-work,
    float64, :232-238).  Priority ties: the reference draws `np.random.choice` among the tied
    candidates; here `rng=None` takes the FIRST tied candidate and the number of tie events is
    returned so callers know whether the reference's RNG could have mattered; pass a
    `np.random.RandomState` to draw like the reference.
    Returns (mseq [n_m, n_j] of 1-based node ids, n_ties).
    """
    dur = np.asarray(dur).astype(np.int64)
    mch = np.asarray(mch).astype(np.int64)
    n_j, n_m = dur.shape
    n = n_j * n_m
    starts = -high * np.ones((n_m, n_j), dtype=np.int64)
    ops = -n_j * np.ones((n_m, n_j), dtype=np.int64)
    nxt = np.zeros(n_j, dtype=np.int64)            # position of each job's next op
    cum = np.cumsum(dur, axis=1)
    total = cum[:, -1]
    done_work = np.zeros(n_j, dtype=np.int64)
    ties = 0
    for _ in range(n):
        jobs = np.nonzero(nxt < n_m)[0]
        if rule == "spt":
            pri = dur[jobs, nxt[jobs]].astype(np.float64)
        elif rule == "fdd-divide-mwkr":
            pri = cum[jobs, nxt[jobs]] / (total[jobs] - done_work[jobs])
        else:
            raise AssertionError('rule must be "spt" or "fdd-divide-mwkr"')
        tied = np.nonzero(pri == pri.min())[0]
        if len(tied) > 1:
            ties += 1
        pick = tied[0] if rng is None else rng.choice(tied)
        j = jobs[pick]
        a = j * n_m + nxt[j]
        _left_shift_insert(a, dur, mch, starts, ops, n_m, high)
        done_work[j] += dur[j, nxt[j]]
        nxt[j] += 1
    return (ops + 1).astype(np.int64), ties


# ----------------------------------------------------------------------------------------------
# counter-based RNG shared with the CUDA RANDOM policy (not part of the reference)
# ----------------------------------------------------------------------------------------------
def mix64(seed, inst, step):
    z = (seed ^ ((inst * 0x9E3779B97F4A7C15) & MASK64) ^ ((step * 0xD1B54A32D192ED03) & MASK64)) & MASK64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
    return z ^ (z >> 31)


def random_pick(seed, inst, step, count):
    return ((mix64(seed, inst, step) >> 32) * count) >> 32


# ----------------------------------------------------------------------------------------------
# the environment
# ----------------------------------------------------------------------------------------------
class OracleInstance:
    """Mutable search state of one instance."""

    def __init__(self, dur, mch, mseq):
        self.dur = np.asarray(dur).astype(np.int64)
        self.mch = np.asarray(mch).astype(np.int64)
        self.n_j, self.n_m = self.dur.shape
        self.n = self.n_j * self.n_m
        self.dur1 = padded_durations(self.dur)
        self.mch_flat = self.mch.reshape(-1)
        self.mseq = np.array(mseq, dtype=np.int64)
        self.mpred, self.msucc = machine_links(self.mseq, self.n + 2)
        self.jobfirst = np.zeros(self.n + 2, dtype=bool)
        self.tabu = None
        self.evaluate()

    def evaluate(self):
        self.est, self.lst, self.makespan = evaluate(self.dur1, self.mpred, self.msucc, self.n_j, self.n_m)

    def path(self):
        return critical_path(self.est, self.dur1, self.mpred, self.jobfirst, self.n_j, self.n_m)

    def moves(self):
        return n5_pairs(self.path(), self.mch_flat, self.tabu)

    def apply(self, a0, a1):
        """swap + tabu update (env/environment.py:357,370-380); dummy [0,0] is a no-op."""
        if a0 == 0 and a1 == 0:
            return
        apply_move(self.mseq, self.mpred, self.msucc, self.jobfirst, self.mch_flat, a0, a1)
        self.tabu = (a1, a0)

    def neighbour_makespan(self, a0, a1):
        mseq, mp, ms, jf = self.mseq.copy(), self.mpred.copy(), self.msucc.copy(), self.jobfirst.copy()
        apply_move(mseq, mp, ms, jf, self.mch_flat, a0, a1)
        return evaluate(self.dur1, mp, ms, self.n_j, self.n_m)[2]


class OracleEnv:
    """numpy counterpart of JsspN5 (env/environment.py:39-423): same reset/step semantics, numpy outputs."""

    def __init__(self, n_job, n_mch, low=1, high=99, reward_type="yaoxin", fea_norm_const=1000):
        self.n_job, self.n_mch, self.low, self.high = n_job, n_mch, low, high
        if reward_type not in ("yaoxin", "consecutive"):
            raise ValueError('reward type must be "yaoxin" or "consecutive".')
        self.reward_type, self.fea_norm_const = reward_type, fea_norm_const
        self.itr = 0

    def reset(self, instances, init_type="fdd-divide-mwkr", mseq=None):
        """env/environment.py:389-409.  `mseq` [B, n_m, n_j] injects a given initial solution
        (e.g. the reference's own, to stay clear of its RNG tie-breaks)."""
        instances = np.asarray(instances)
        self.insts = []
        for b, ins in enumerate(instances):
            if mseq is not None:
                order = mseq[b]
            else:
                assert init_type in ("spt", "fdd-divide-mwkr"), 'Initial solution type = "plist", "spt", "fdd-divide-mwkr".'
                order, _ = rules_solver(ins[0], ins[1], init_type, self.high)
            self.insts.append(OracleInstance(ins[0], ins[1], order))
        self.current_objs = np.array([[i.makespan] for i in self.insts], dtype=np.float32)
        self.incumbent_objs = self.current_objs.copy()
        self.itr = 0
        fa, flag = self.feasible_actions()
        return self.state(), fa, ~flag

    def feasible_actions(self):
        """env/environment.py:411-423."""
        fa, flag = [], []
        for ins in self.insts:
            pairs = ins.moves()
            fa.append(pairs if pairs else [[0, 0]])
            flag.append(bool(pairs))
        return fa, np.array(flag, dtype=bool).reshape(-1, 1)

    def state(self):
        """(x, edge_index_pc, edge_index_mc, batch): env/environment.py:291-316."""
        n1 = self.n_job * self.n_mch + 2
        xs, pcs, mcs = [], [], []
        pc = edge_index_pc(self.n_job, self.n_mch)
        for b, ins in enumerate(self.insts):
            xs.append(features(ins.dur1, ins.est, ins.lst, self.high, self.fea_norm_const))
            pcs.append(pc + b * n1)
            mcs.append(edge_index_mc(ins.msucc, ins.n) + b * n1)
        batch = np.repeat(np.arange(len(self.insts), dtype=np.int64), n1)
        return np.concatenate(xs), np.concatenate(pcs, axis=1), np.concatenate(mcs, axis=1), batch

    def step(self, actions):
        """env/environment.py:356-387."""
        for ins, a in zip(self.insts, actions):
            ins.apply(int(a[0]), int(a[1]))
            ins.evaluate()
        ms = np.array([[i.makespan] for i in self.insts], dtype=np.float32)
        if self.reward_type == "consecutive":
            reward = self.current_objs - ms
        else:
            reward = np.maximum(self.incumbent_objs - ms, 0).astype(np.float32)
        self.incumbent_objs = np.minimum(self.incumbent_objs, ms)
        self.current_objs = ms
        self.itr += 1
        fa, flag = self.feasible_actions()
        return self.state(), reward, fa, ~flag


def greedy_action(ins):
    """Best neighbour by makespan, first minimum in move-list order, always moves:
    Greedy_baselines (conventional_baselines_with_long-term-mem.py:186-197)."""
    pairs = ins.moves()
    if not pairs:
        return [0, 0]
    vals = [ins.neighbour_makespan(a0, a1) for a0, a1 in pairs]
    return pairs[int(np.argmin(vals))]


def run_policy(env, fa, policy, steps, tape=None, seed=0, inst_offset=0, log_steps=()):
    """Drive `env` for `steps` steps with a device-style policy: 'replay' (tape [B,steps,2]),
    'random' (counter-based `random_pick` among the feasible list) or 'greedy'.
    Returns (actions taken [B,steps,2], incumbents at log_steps [len(log_steps),B])."""
    B = len(env.insts)
    taken = np.zeros((B, steps, 2), dtype=np.int64)
    logs = []
    for k in range(steps):
        acts = []
        for b in range(B):
            if policy == "replay":
                a = [int(tape[b, k, 0]), int(tape[b, k, 1])]
            elif policy == "random":
                if fa[b] == [[0, 0]]:
                    a = [0, 0]
                else:
                    a = fa[b][random_pick(seed, inst_offset + b, env.itr, len(fa[b]))]
            elif policy == "greedy":
                a = greedy_action(env.insts[b])
            else:
                raise ValueError(policy)
            acts.append([int(a[0]), int(a[1])])
        taken[:, k] = np.array(acts)
        _, _, fa, _ = env.step(acts)
        if (k + 1) in log_steps:
            logs.append(env.incumbent_objs.reshape(-1).copy())
    return taken, (np.stack(logs) if logs else np.zeros((0, B), dtype=np.float32))
