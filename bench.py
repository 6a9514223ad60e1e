#!/usr/bin/env python
"""Benchmark of the L2S hot path: N5 local-search steps per second (instances x steps / s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--shape 20x15] [--batch 8192]

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): batched synthetic 20x15
instances, durations U{1..98} (`uni_instance_gen(20,15,1,99)` semantics), batch 8192 PER GPU (weak scaling:
instances are independent, one batch slice per rank, no data-path collective), fdd-divide-mwkr initial
solutions built on the device, actions = a recorded tape of uniformly random feasible N5 moves.

A "step" is ONE `l2s_step` launch over one batch: swap + re-evaluation (est, lst, makespan) + features x +
machine edges + reward/incumbent/tabu + critical path + N5 move set for every instance of the batch.

  value : device-resident throughput (state and action tape already in HBM), CUDA events, max over ranks.  The
          batch is stepped as 4 independent environments of `--batch` instances (they also make the per-step working
          set exceed the L2), round-robin, each on a CUDA stream of its own: a step of one environment waits for that
          environment's previous step, while the ramp-down of one launch overlaps the start-up of the next.  The
          timed region is `reps` back-to-back windows of exactly K steps (reps = ceil(1600 / K), so that the region
          lasts >= ~40 ms whatever K is: a 20-step window is 0.5 ms, far too short next to rank skew and clock
          sampling); `steps_timed` = K * reps, `ms_per_step` = region / steps_timed.  No collective inside; the one
          collective of a search -- the final all-gather of makespans and incumbents -- is warmed up and then timed
          on its own (`final_gather_ms`).  `single_stream` holds the same launches serialised on ONE stream (one
          launch timed alone; the `roofline` object is computed from it, `roofline.steady_state` from the headline).
  e2e   : the same steps through the C-ABI with HOST buffers: the kernel reads the host actions over PCIe from pinned
          memory and writes one result record per instance (move set, reward, makespan, incumbent, done, status) over
          PCIe into pinned host memory.  One host thread drives the 4 environments through `l2s_step_host_begin` /
          `l2s_step_host_wait` with 3 of them in flight; every step's records are in host memory and read before that
          environment's next step is issued.  `e2e.blocking`: one blocking `l2s_step_host` call per step.
  e2e_python_api : the same through the drop-in `JsspN5.step(actions, device)` (Python lists in and out).
  evaluator : `Evaluator.forward` on the int64 COO edge lists of the same schedules (graphs/s, GB/s).
  large_instances : BASELINE configs[3]: 100x20 and 150x25, 1024 instances sharded over the ranks.
  training : BASELINE configs[4]: n-step REINFORCE on 10x10, 64 rollouts sharded over the ranks, 500-step episode.
  learned_policy : BASELINE configs[0] / [1] (N = 1): shipped models on syn10x10 x 500 steps and tai20x15 x 5000 steps.
  roofline : algorithmic bytes of one step (SURVEY.md 8d formula) / mean launch duration vs measured HBM peak.
  cpu_baseline : the UNMODIFIED reference `JsspN5.step` (vendored copy in baseline/_ref, oracle/vendor_ref.py) on the
          box's host cores, one process per core on disjoint instance slices, bounded sample.

`--impl reference` times that same unmodified reference path and prints the same JSON line with "impl": "reference".
"""
import argparse
import collections
import ctypes as C
import json
import os
import subprocess
import sys
import threading  # noqa: F401
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "N5 LS steps/sec (instances x steps)"
UNIT = "inst-steps/s"
PIPE_DEPTH = 3     # independent environments in flight in the pipelined e2e loop (< REPLICAS)
REPLICAS = 4        # rotating independent batches so that the per-step working set exceeds the 126 MB L2
MIN_TIMED_STEPS = 1600


def workload_name(shape, batch):
    return ("syn%s U{1..98}, batch %d per GPU, fdd-divide-mwkr init, replayed random feasible N5 moves; one step = one "
            "environment step (l2s_step launch) over the batch" % (shape, batch))


def config_of(args, world):
    """Identical in both arms."""
    return {"workload": workload_name(args.shape, args.batch), "shape": args.shape, "batch_per_gpu": args.batch,
            "global_batch": world * args.batch, "parallelism": "instance-sharded x%d" % world}


def gen_instances(n_j, n_m, B, seed):
    """uni_instance_gen semantics (env/generateJSP.py:14-18), vectorised: durations randint(1, 99) ->
    U{1..98}; machine rows = independent permutations of 1..n_m."""
    rng = np.random.RandomState(seed)
    dur = rng.randint(1, 99, size=(B, n_j, n_m))
    mch = rng.random_sample((B, n_j, n_m)).argsort(axis=2) + 1
    return np.stack([dur, mch], axis=1).astype(np.int32)


def bytes_per_step(n_j, n_m, pbar):
    """Algorithmic bytes of one instance-step behind the step API (SURVEY.md 8d / BASELINE.md 4)."""
    n = n_j * n_m
    n1, e_mc = n + 2, n_m * (n_j - 1)
    return (17 * n + 20) + (12 * n1 + 16 * e_mc + 27 + 4 + 8 * pbar + 13)


def bytes_per_graph(n_j, n_m):
    n1, e_mc = n_j * n_m + 2, n_m * (n_j - 1)
    return 8 * n1 + 16 * (n_j * (n_m + 1) + e_mc) + 8 * n1 + 4


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def host_cores():
    try:
        return sorted(os.sched_getaffinity(0))
    except Exception:
        return list(range(os.cpu_count() or 1))


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        # one long-lived nvidia-smi in loop mode (20 ms period); its stdout is parsed at stop()
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is not None:
            try:
                self.proc.terminate()
                out, _ = self.proc.communicate(timeout=10)
                self.rows = [[c.strip() for c in ln.split(",")] for ln in out.strip().split("\n") if ln.strip()]
            except Exception:
                self.rows = []
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def pin_to_gpu_numa(local_rank, world, force=False):
    """Bind this rank to its own share of the cores local to its GPU's NUMA node (the e2e loop is one host thread
    per rank that launches, waits and reads pinned memory: migrations and remote-node memory cost microseconds per
    step).  Only when the box has more than one NUMA node (`force` binds anyway: diagnostics).  Returns a description,
    or None when the topology cannot be read."""
    try:
        import torch
        def cpulist(i):
            p = torch.cuda.get_device_properties(i)
            path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/local_cpulist" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
            cores = set()
            for part in open(path).read().strip().split(","):
                if "-" in part:
                    a, b = part.split("-")
                    cores.update(range(int(a), int(b) + 1))
                elif part:
                    cores.add(int(part))
            return tuple(sorted(cores))
        mine = cpulist(local_rank)
        allowed = set(host_cores())
        local = [c for c in mine if c in allowed]
        if not local:
            return None
        if len(local) == len(allowed) and not force:
            # one NUMA node: every core is local to every GPU, nothing to gain from a binding (measured on this pool's
            # 8-GPU box, 32 cores, one node: pipelined e2e 26.5 us per step unbound vs 28-35 us with 4 cores per rank;
            # profiles/e2e_pipeline_ranks_r2.json)
            return {"numa_cores": len(local), "ranks_on_node": min(world, torch.cuda.device_count()), "bound_to": None}
        n_local = min(world, torch.cuda.device_count())
        sharing = [i for i in range(n_local) if cpulist(i) == mine]
        k, idx = len(sharing), sharing.index(local_rank)
        per = max(1, len(local) // k)
        share = local[idx * per:(idx + 1) * per] or local
        os.sched_setaffinity(0, set(share))
        return {"numa_cores": len(local), "ranks_on_node": k, "bound_to": "%d-%d (%d cores)" % (share[0], share[-1], len(share))}
    except Exception as e:   # no sysfs / not permitted: run unpinned
        return {"error": str(e)[:120]}


# ------------------------------------------------------------------------------------------------------
# CPU baseline: the UNMODIFIED reference (baseline/_ref or /root/reference through oracle/ref_harness.py)
# ------------------------------------------------------------------------------------------------------
def _ref_worker(args):
    """One process = one slice: reset with fdd-divide-mwkr, then time env.step only (perf_counter), actions drawn
    uniformly among the feasible moves like the recorded tape of our arm."""
    n_j, n_m, lo, hi, steps, warm, seed, threads, batch = args
    import random
    import torch
    from oracle import ref_harness
    torch.set_num_threads(threads)
    R = ref_harness.load()
    inst = gen_instances(n_j, n_m, batch, seed=1)[lo:hi]    # the first instances of our arm's first batch
    np.random.seed(seed)
    random.seed(seed)
    env = R.JsspN5(n_j, n_m, 1, 99)
    t0 = time.perf_counter()
    _, fa, _ = env.reset(inst, "fdd-divide-mwkr", "cpu")
    t_reset = time.perf_counter() - t0
    rnd = random.Random(seed)
    total = 0.0
    for k in range(warm + steps):
        acts = [rnd.choice(f) for f in fa]
        t0 = time.perf_counter()
        _, _, fa, _ = env.step(acts, "cpu")
        if k >= warm:
            total += time.perf_counter() - t0
    return (hi - lo) * steps, total, t_reset


def reference_rate(n_j, n_m, steps, warm, procs, per, batch=8192):
    """inst-steps/s of the reference's JsspN5.step (env/environment.py:356-387): `procs` processes x `per` instances,
    one torch thread each (procs == 1: a single process with all torch threads -- how the reference itself runs)."""
    from oracle import ref_harness
    if not ref_harness.available():
        return None
    if procs <= 1:
        units, secs, t_reset = _ref_worker((n_j, n_m, 0, per, steps, warm, 1, len(host_cores()), batch))
        return {"value": units / secs, "reset_s_per_instance": t_reset / per}
    import multiprocessing as mp
    jobs = [(n_j, n_m, i * per, (i + 1) * per, steps, warm, 1 + i, 1, batch) for i in range(procs)]
    with mp.get_context("spawn").Pool(procs) as pool:
        res = pool.map(_ref_worker, jobs)
    # every process times only its env.step loop; the job rate is the sum of the per-process rates
    return {"value": sum(u / s for u, s, _ in res), "reset_s_per_instance": float(np.mean([r for _, _, r in res])) / per}


def cpu_baseline_reference(n_j, n_m, steps, warm, batch=8192):
    cores = host_cores()
    procs = max(1, min(len(cores), 64))
    per = 4 if n_j * n_m <= 400 else 1
    t0 = time.perf_counter()
    allc = reference_rate(n_j, n_m, steps, warm, procs, per, batch)
    if allc is None:
        return None
    wall = time.perf_counter() - t0
    one = reference_rate(n_j, n_m, max(2, min(steps, 6)), min(warm, 2), 1, 16 if n_j * n_m <= 400 else 2, batch)
    return {"value": allc["value"], "unit": UNIT, "cores": procs, "kind": "reference",
            "sample": "unmodified reference JsspN5.step (env/environment.py:356-387; vendored baseline/_ref, networkx + "
                      "torch CPU): %d processes x %d instances x %d timed env.step calls after %d warm-up steps, one "
                      "torch thread per process, same instances as our first batch, actions uniform among the feasible "
                      "moves; %.1f s wall incl. reset (%.0f ms per instance)"
                      % (procs, per, steps, warm, wall, 1e3 * allc["reset_s_per_instance"]),
            "single_process": {"value": one["value"], "unit": UNIT, "threads": len(cores),
                               "what": "one process, torch.set_num_threads(all cores): how the reference itself runs"}}


def c_port_rate(n_j, n_m, procs):
    """The plain-C restatement (oracle/l2s_oracle.c, OpenMP) on all cores: the strongest CPU code in the repo."""
    try:
        from oracle import c_oracle as CO
        cb = 64 * procs if n_j * n_m <= 400 else 4 * procs
        inst = gen_instances(n_j, n_m, cb, 5)
        mseq = np.stack([CO.rules_solver(i[0], i[1])[0] for i in inst])
        cenv = CO.CEnv(inst, mseq, threads=procs)
        cenv.run("random", 5, seed=1, write_state=True, return_actions=False)
        ksteps = 50 if n_j * n_m <= 400 else 10
        t1 = time.perf_counter()
        cenv.run("random", ksteps, seed=2, write_state=True, return_actions=False)
        return {"value": cb * ksteps / (time.perf_counter() - t1), "unit": UNIT, "threads": procs, "kind": "port",
                "what": "oracle/l2s_oracle.c orc_run_batch (OpenMP): per step swap + O(n) topological evaluation + "
                        "features + machine edges + critical path + N5 pairs, %d instances x %d steps" % (cb, ksteps)}
    except Exception as e:   # optional extra
        return {"error": str(e)[:200]}


def run_reference(args, rank, world):
    if rank != 0:
        return
    n_j, n_m = [int(v) for v in args.shape.split("x")]
    steps, warm = max(1, args.steps), max(0, args.warmup)
    # bounded: the reference needs ~4-5 ms per instance-step at 20x15, the whole run must end within minutes
    steps_t, warm_t = min(steps, 30), min(warm, 5)
    t0 = time.perf_counter()
    cpu = cpu_baseline_reference(n_j, n_m, steps_t, warm_t, args.batch)
    wall = time.perf_counter() - t0
    if cpu is None:
        print(json.dumps({"impl": "reference", "unavailable": "baseline/_ref is missing: run `python -m oracle.vendor_ref` "
                          "where /root/reference exists (build() does it)"}), flush=True)
        return
    rate = cpu["value"]
    per = 4 if n_j * n_m <= 400 else 1
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "steps_timed": steps_t,
        "ms_per_step": 1e3 * per * cpu["cores"] / rate,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32/int64 (torch CPU + networkx)",
        "data": "synthetic", "config": config_of(args, world),
        "sample": "each timed step advances %d instances of the workload (%d processes x %d); throughput is per "
                  "instance-step and the reference's per-instance cost does not fall with the batch size"
                  % (per * cpu["cores"], cpu["cores"], per),
        "cpu_baseline": cpu,
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": wall, "c_port": c_port_rate(n_j, n_m, cpu["cores"]),
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------
class Replicas:
    """REPLICAS independent environments of one shape with recorded action tapes and preallocated outputs."""

    def __init__(self, n_j, n_m, B, T, rank, dev, seed0=1):
        import torch
        from l2s_b200 import JsspN5
        self.envs, self.tapes, self.outs = [], [], []
        n1, e_mc = n_j * n_m + 2, n_m * (n_j - 1)
        for r in range(REPLICAS):
            env = JsspN5(n_j, n_m, 1, 99)
            env.instance_offset = (rank * REPLICAS + r) * B
            inst = gen_instances(n_j, n_m, B, seed=seed0 + rank * REPLICAS + r)
            env.reset(inst, "fdd-divide-mwkr", dev)
            init = env.solutions().clone()
            _, taken = env.run("random", T, seed=7 + r, return_actions=True)   # record the action tape
            tape = taken.permute(1, 0, 2).contiguous()                          # [T, B, 2] device
            env.reset(inst, "fdd-divide-mwkr", dev, init_solutions=init.cpu().numpy())
            self.envs.append(env)
            self.tapes.append(tape)
            self.outs.append(dict(x=torch.empty((B * n1, 3), dtype=torch.float32, device=dev),
                                  emc=torch.empty((2, B * e_mc), dtype=torch.int64, device=dev),
                                  ms=torch.empty(B, dtype=torch.float32, device=dev),
                                  inc=torch.empty(B, dtype=torch.float32, device=dev),
                                  rew=torch.empty(B, dtype=torch.float32, device=dev),
                                  done=torch.empty(B, dtype=torch.uint8, device=dev),
                                  pairs=torch.empty((B, env.pmax, 2), dtype=torch.int32, device=dev),
                                  npairs=torch.empty(B, dtype=torch.int32, device=dev),
                                  status=torch.empty(B, dtype=torch.int32, device=dev)))
        self.pos = [0] * REPLICAS

    def io_for(self, r, t):
        from l2s_b200 import _capi
        o = self.outs[r]
        return _capi.StepIO(actions=self.tapes[r][t].data_ptr(), high=99.0, fea_norm_const=1000.0, reward_type=0,
                            is_reset=0, x=o["x"].data_ptr(), edge_index_mc=o["emc"].data_ptr(),
                            makespan=o["ms"].data_ptr(), incumbent=o["inc"].data_ptr(), reward=o["rew"].data_ptr(),
                            done=o["done"].data_ptr(), pairs=o["pairs"].data_ptr(), npairs=o["npairs"].data_ptr(),
                            status=o["status"].data_ptr())

    def next_ios(self, count):
        ios = []
        for k in range(count):
            r = k % REPLICAS
            ios.append((r, self.io_for(r, self.pos[r])))
            self.pos[r] += 1
        return ios


def timed_device_steps(rep, W, K_total, lib, sp, stream, barrier, dev):
    """W warm-up launches, then K_total launches between two CUDA events on the launching stream."""
    import torch
    from l2s_b200 import _capi
    ios = rep.next_ios(W + K_total)
    for r, io in ios[:W]:
        _capi.check(lib.l2s_step(rep.envs[r]._h, C.byref(io), sp), "l2s_step")
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for r, io in ios[W:]:
        lib.l2s_step(rep.envs[r]._h, C.byref(io), sp)
    e1.record(stream)
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    st = torch.stack([o["status"] for o in rep.outs])
    assert int(st.abs().sum()) == 0, "device-side error during the timed region"
    return ms


def timed_device_steps_streams(rep, W, K_total, lib, dev):
    """The same launches with every replica (an independent environment) on a stream of its own: the tail of one
    launch overlaps the head of the next.  Aggregate time between a fork event and the join of all streams."""
    import torch
    from l2s_b200 import _capi
    main = torch.cuda.current_stream(dev)
    streams = [torch.cuda.Stream(device=dev) for _ in range(REPLICAS)]
    sps = [C.c_void_p(st.cuda_stream) for st in streams]
    ios = rep.next_ios(W + K_total)
    for st in streams:
        st.wait_stream(main)
    for r, io in ios[:W]:
        _capi.check(lib.l2s_step(rep.envs[r]._h, C.byref(io), sps[r]), "l2s_step")
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for st in streams:
        st.wait_event(e0)
    for r, io in ios[W:]:
        lib.l2s_step(rep.envs[r]._h, C.byref(io), sps[r])
    for st in streams:
        main.wait_stream(st)
    e1.record(main)
    torch.cuda.synchronize(dev)
    st_ = torch.stack([o["status"] for o in rep.outs])
    assert int(st_.abs().sum()) == 0, "device-side error during the multi-stream region"
    return e0.elapsed_time(e1)


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from l2s_b200 import JsspN5, _capi  # noqa: F401

    n_j, n_m = [int(v) for v in args.shape.split("x")]
    B, K, W = args.batch, args.steps, max(3, args.warmup)
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (l2s_b200 has no CPU path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    affinity = pin_to_gpu_numa(local_rank, world) if not args.no_pin else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _capi.lib()
    n, n1, e_mc = n_j * n_m, n_j * n_m + 2, n_m * (n_j - 1)

    reps = max(1, -(-MIN_TIMED_STEPS // K))
    K1 = K * reps                                   # device-resident steps actually timed
    k_e2e = max(1, args.e2e_steps)                  # python-API steps (~0.5 ms each)
    w_e2e = 3
    K3 = K * max(1, -(-1000 // K))                  # C-ABI e2e steps actually timed
    T = 2 * (-(-(W + K1) // REPLICAS)) + -(-(w_e2e + k_e2e) // REPLICAS) + 2 * (-(-(w_e2e + K3) // REPLICAS)) + 2
    rep = Replicas(n_j, n_m, B, T, rank, dev)
    envs, tapes, outs = rep.envs, rep.tapes, rep.outs
    stream = torch.cuda.current_stream(dev)
    sp = C.c_void_p(stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # the one collective of a search: final gather of makespans and incumbents -- preallocated, warmed up here so that
    # NCCL's lazy channel / connection set-up never lands in a timed region, timed separately below
    gather_in = torch.zeros((2, B), dtype=torch.float32, device=dev)
    gather_out = torch.zeros((world, 2, B), dtype=torch.float32, device=dev)
    if world > 1:
        for _ in range(3):
            dist.all_gather_into_tensor(gather_out, gather_in)
    barrier()

    # ---- phase 1: device-resident ------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_dev = timed_device_steps(rep, W, K1, lib, sp, stream, barrier, dev)
    last = outs[(W + K1 - 1) % REPLICAS]
    gather_ms = 0.0
    if world > 1:
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        gather_in[0].copy_(last["ms"])
        gather_in[1].copy_(last["inc"])
        dist.all_gather_into_tensor(gather_out, gather_in)
        e1.record(stream)
        torch.cuda.synchronize(dev)
        gather_ms = e0.elapsed_time(e1)
    barrier()
    ms_streams = timed_device_steps_streams(rep, W, K1, lib, dev)
    barrier()
    pbar = float(torch.stack([o["npairs"] for o in outs]).float().mean())
    if args.device_only:
        if rank == 0:
            sampler.stop()
            print(json.dumps({"shape": args.shape, "batch": B, "launch_us": 1e3 * ms_dev / K1,
                              "value": world * B * K1 / (ms_dev / 1e3),
                              "streams_us_per_step": 1e3 * ms_streams / K1}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return
    pos = rep.pos

    # ---- phase 2: end to end through the Python drop-in API ------------------------------------------
    host_tapes = [t.cpu().numpy() for t in tapes]
    lists = []
    p2 = list(pos)
    for k in range(w_e2e + k_e2e):
        r = k % REPLICAS
        lists.append((r, host_tapes[r][p2[r]].tolist()))
        p2[r] += 1
    for r, acts in lists[:w_e2e]:
        envs[r].step(acts, dev)
    barrier()
    t0 = time.perf_counter()
    sink = 0
    for r, acts in lists[w_e2e:]:
        state, reward, fa, done = envs[r].step(acts, dev)
        sink += len(fa) + len(fa[0])
    torch.cuda.synchronize(dev)
    t_py = time.perf_counter() - t0
    pos = p2

    # ---- phase 3: end to end through the C-ABI with host buffers -------------------------------------
    # the step's inputs start in pinned host memory (one [B,2] int32 action block per step)
    pinned_keep = [torch.from_numpy(np.ascontiguousarray(t)).pin_memory() for t in host_tapes]
    pinned_tapes = [t.numpy() for t in pinned_keep]
    hb = []
    p3 = list(pos)
    for k in range(2 * (w_e2e + K3)):
        r = k % REPLICAS
        hb.append((r, pinned_tapes[r][p3[r]]))
        p3[r] += 1
    # one l2s_step_io per environment, filled once (the device output pointers do not change between steps)
    host_ios = []
    for r in range(REPLICAS):
        o = outs[r]
        host_ios.append(_capi.StepIO(actions=None, high=99.0, fea_norm_const=1000.0, reward_type=0, is_reset=0,
                                     x=o["x"].data_ptr(), edge_index_mc=o["emc"].data_ptr(), makespan=o["ms"].data_ptr(),
                                     incumbent=o["inc"].data_ptr(), reward=o["rew"].data_ptr(), done=o["done"].data_ptr(),
                                     pairs=None, npairs=None, status=None))
    step_host = lib.l2s_step_host
    io_refs = [C.byref(io) for io in host_ios]
    handles = [e._h for e in envs]
    recs = [e._h_rec for e in envs]

    def capi_step(r, acts_ptr):
        st = step_host(handles[r], io_refs[r], acts_ptr, None, None, None, None, None, None, sp)
        if st < 0:
            _capi.check(st, "l2s_step_host")
        rec = recs[r]
        return int(rec[0, 0]) + int(rec[-1, 0])     # touch the host results (first and last record)

    calls = [(r, acts.ctypes.data) for r, acts in hb]     # hb keeps the action arrays alive
    for r, ptr in calls[:w_e2e]:
        capi_step(r, ptr)
    barrier()
    t0 = time.perf_counter()
    for r, ptr in calls[w_e2e:w_e2e + K3]:
        capi_step(r, ptr)
    torch.cuda.synchronize(dev)
    t_capi = time.perf_counter() - t0
    assert not any((e._h_rec[:, 0] >> 24).any() for e in envs)

    # the same steps with the call split in two (l2s_step_host_begin / l2s_step_host_wait): the host thread keeps
    # PIPE_DEPTH of the REPLICAS independent environments in flight, so the launch + completion latency of one step
    # hides behind the kernels of the others; every step still takes its actions from host memory and its records are
    # in host memory (and touched) before that environment's next step is issued
    begin, wait = lib.l2s_step_host_begin, lib.l2s_step_host_wait
    own_streams = [torch.cuda.Stream(device=dev) for _ in range(REPLICAS)]
    for st in own_streams:
        st.wait_stream(stream)
    sps_own = [C.c_void_p(st.cuda_stream) for st in own_streams]

    def finish(r):
        st = wait(handles[r], None, None, None, None, None, None)
        if st < 0:
            _capi.check(st, "l2s_step_host_wait")
        rec = recs[r]
        return int(rec[0, 0]) + int(rec[-1, 0])

    def pipelined(seq):
        inflight = collections.deque()
        for r, ptr in seq:
            if len(inflight) == PIPE_DEPTH:
                finish(inflight.popleft())
            st = begin(handles[r], io_refs[r], ptr, sps_own[r])
            if st < 0:
                _capi.check(st, "l2s_step_host_begin")
            inflight.append(r)
        while inflight:
            finish(inflight.popleft())

    rest = calls[w_e2e + K3:]
    pipelined(rest[:w_e2e])
    barrier()
    t0 = time.perf_counter()
    pipelined(rest[w_e2e:])
    torch.cuda.synchronize(dev)
    t_pipe = time.perf_counter() - t0
    assert not any((e._h_rec[:, 0] >> 24).any() for e in envs)

    # ---- phase 4: stateless evaluator (Evaluator.forward drop-in) on the current schedules ----------------
    ms_ev, k_ev = time_evaluator(envs, n_j, n_m, B, K, dev, stream, barrier)
    ms_ev_strict = time_evaluator_strict(envs, n_j, n_m, B, dev)
    ms_ev_st, k_ev_st = time_evaluator_streams(envs, n_j, n_m, B, K, dev, barrier)

    # ---- phase 5: fused K-step search (l2s_run): on-device policy, state resident in shared memory ------------
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k_f = max(10, min(K, 100))
    envs[0].run("random", 5, seed=11)
    barrier()
    e0.record(stream)
    for r in range(REPLICAS):
        envs[r].run("random", k_f, seed=12 + r)
    e1.record(stream)
    barrier()
    ms_fused = e0.elapsed_time(e1)
    k_g = max(2, min(K, 10))
    barrier()
    e0.record(stream)
    envs[0].run("greedy", k_g)
    e1.record(stream)
    barrier()
    ms_greedy = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    del rep, envs, tapes, outs, host_tapes, pinned_keep, pinned_tapes, hb, calls, recs, handles, io_refs, host_ios
    torch.cuda.empty_cache()

    # ---- BASELINE configs[3] and configs[4] ------------------------------------------------------------------
    large = None if args.no_large else large_instances(rank, world, dev, lib, sp, stream, barrier)
    training = None if args.no_training else training_episode(rank, world, dev, barrier)
    learned = learned_policy_rollouts(dev) if (world == 1 and rank == 0 and not args.no_learned) else None

    # ---- reduce over ranks ---------------------------------------------------------------------------
    times = torch.tensor([ms_dev / 1e3, t_py, t_capi, ms_ev / 1e3, ms_fused / 1e3, ms_greedy / 1e3, gather_ms / 1e3,
                          ms_streams / 1e3, t_pipe, ms_ev_st / 1e3], dtype=torch.float64, device=dev)
    per_rank = None
    if world > 1:
        allt = [torch.zeros_like(times) for _ in range(world)]
        dist.all_gather(allt, times)
        per_rank = torch.stack(allt).cpu().numpy()
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    t_dev, t_py, t_capi, t_ev, t_fused, t_greedy, t_gather, t_streams, t_pipe, t_ev_st = [float(v) for v in times.cpu()]
    value = world * B * K1 / t_streams              # headline: every independent environment on a stream of its own
    e2e_py = world * B * k_e2e / t_py
    e2e_capi = world * B * K3 / t_capi

    if rank == 0:
        peak, peak_src = measured_peak()
        bps = bytes_per_step(n_j, n_m, pbar)
        launch_s = t_dev / K1
        achieved = bps * B / launch_s / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "step_kernel_traffic.json")
        if os.path.isfile(tp):
            try:
                traffic = json.load(open(tp)).get("%s_b%d" % (args.shape, B))
            except Exception:
                traffic = None
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_baseline_reference(n_j, n_m, 20, 3, B)
        bpg = bytes_per_graph(n_j, n_m)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "steps_timed": K1, "ms_per_step": 1e3 * t_streams / K1, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": config_of(args, world),
            "run": {"timed_region": "%d windows of %d steps back to back = %d l2s_step launches, the %d independent environments "
                                    "round-robin, each on a CUDA stream of its own (a step of one environment still waits "
                                    "for that environment's previous step), between a fork event and the join of the "
                                    "streams (%.1f ms); no collective inside; the same launches on ONE stream: see "
                                    "single_stream" % (reps, K, K1, REPLICAS, 1e3 * t_streams),
                    "l2": "%d independent batches stepped round-robin: per-step working set x%d > 126 MB L2"
                          % (REPLICAS, REPLICAS),
                    "mean_pairs": pbar, "final_gather_ms": 1e3 * t_gather,
                    "final_gather": "all_gather_into_tensor of (makespan, incumbent) [2,%d] f32 per rank, once per search "
                                    "(500 steps in the named config), warmed up, timed on its own" % B,
                    "affinity": affinity,
                    "per_rank_ms": None if per_rank is None else
                    {"device_steps": [round(1e3 * v, 3) for v in per_rank[:, 7]],
                     "device_steps_single_stream": [round(1e3 * v, 3) for v in per_rank[:, 0]],
                     "e2e": [round(1e3 * v, 3) for v in per_rank[:, 8]],
                     "e2e_blocking": [round(1e3 * v, 3) for v in per_rank[:, 2]]}},
            "e2e": {"value": world * B * K3 / t_pipe, "unit": UNIT,
                    "api": "C-ABI l2s_step_host_begin / l2s_step_host_wait (the two halves of l2s_step_host): host int32 "
                           "actions in (read in place from pinned memory by the kernel), host records out "
                           "(pairs/npairs/done/reward/makespan/incumbent/status, written into pinned memory by the kernel); "
                           "one host thread keeps %d of the %d independent environments in flight, each on its own stream; "
                           "a step's records are in host memory and read before that environment's next step is issued"
                           % (PIPE_DEPTH, REPLICAS),
                    "h2d_bytes_per_step": B * 8, "d2h_bytes_per_step": int(B * (16 + 4 * pbar)),
                    "steps": K3, "us_per_step": 1e6 * t_pipe / K3, "depth": PIPE_DEPTH,
                    "blocking": {"value": e2e_capi, "unit": UNIT, "us_per_step": 1e6 * t_capi / K3, "steps": K3,
                                 "api": "l2s_step_host (one call per step, returns when the records are in host memory; "
                                        "what the Python drop-in's JsspN5.step uses): launch + completion latency exposed "
                                        "on every step"}},
            "e2e_python_api": {"value": e2e_py, "unit": UNIT,
                               "api": "drop-in JsspN5.step(actions: list[B][2], device) -> (state, reward, feasible_actions "
                                      "(lazy list[B][P][2] over the pinned records), done)",
                               "steps": k_e2e},
            "evaluator": {"graphs_per_s": world * B * k_ev / t_ev, "ms_per_call": 1e3 * t_ev / k_ev,
                          "api": "Evaluator.forward(edge_index int64 COO, duration int64, n_j, n_m), strict=False (status checked "
                                 "once after the loop)",
                          "strict_ms_per_call": ms_ev_strict,
                          "strict": "default mode: one 4-byte status read-back + sync per call (the reference synchronises "
                                    "2 x depth times per call), wall clock, rank 0",
                          "algorithmic_bytes_per_graph": bpg, "achieved_gbs": bpg * B * k_ev / t_ev / 1e9,
                          "frac_of_peak": bpg * B * k_ev / t_ev / 1e9 / peak,
                          "streams": {"graphs_per_s": world * B * k_ev_st / t_ev_st, "ms_per_call": 1e3 * t_ev_st / k_ev_st,
                                      "frac_of_peak": bpg * B * k_ev_st / t_ev_st / 1e9 / peak,
                                      "what": "the same calls with one Evaluator (workspace) and one CUDA stream per "
                                              "independent input set (%d sets): calls on different streams overlap "
                                              "head to tail" % REPLICAS}},
            "fused_run": {"random_policy_inst_steps_per_s": world * B * k_f * REPLICAS / t_fused,
                          "greedy_policy_inst_steps_per_s": world * B * k_g / t_greedy,
                          "api": "l2s_run: K steps per launch, on-device policy, no host round trip, no x/edge outputs",
                          "steps_per_launch": k_f},
            "single_stream": {"value": world * B * K1 / t_dev, "unit": UNIT, "us_per_step": 1e6 * t_dev / K1,
                              "what": "the same %d launches back to back on ONE stream (round-1 / early round-2 headline): "
                                      "launches serialise, so every launch pays its own start-up burst and ramp-down "
                                      "(9 us + 17.2 us per wave of 5920 instances); the roofline object is computed from "
                                      "this launch time" % K1},
            "large_instances": large, "training": training, "learned_policy": learned,
            "gpu_launches": K1,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": "l2s::step_kernel",
                         "algorithmic_bytes_per_launch": bps * B, "launch_ms": 1e3 * launch_s,
                         "launch": "one launch timed alone (single_stream: launches serialised on one stream)",
                         "steady_state": {"achieved": bps * B * K1 / t_streams / 1e9,
                                          "frac": bps * B * K1 / t_streams / 1e9 / peak,
                                          "how": "algorithmic bytes of all timed launches / timed region of the headline "
                                                 "schedule (launches of independent environments overlap head to tail)"}},
            "cpu_baseline": cpu, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def time_evaluator(envs, n_j, n_m, B, K, dev, stream, barrier):
    import torch
    from l2s_b200 import Evaluator
    n1 = n_j * n_m + 2
    ev = Evaluator(strict=False)      # asynchronous calls; the status word is checked once after the loop
    ev_in = []
    for r in range(len(envs)):
        state, fa, done = envs[r].observe()
        ei = torch.cat([state[1], state[2]], dim=1).contiguous()
        dur = torch.zeros((B, n1), dtype=torch.int64, device=dev)
        dur[:, 1:-1] = torch.from_numpy(envs[r].instances[:, 0].reshape(B, -1).astype(np.int64)).to(dev)
        ev_in.append((ei, dur.reshape(-1, 1)))
    k_ev = max(24, min(K, 40))
    # warm-up: every input twice, lazy kernel loading settled -- and the outputs BOUND like in the timed loop, so that two
    # sets of output tensors are alive at a time and the caching allocator has grown to that before the window (with
    # the results dropped, the first timed call paid a cudaMalloc: 100x20 at 128 graphs read 79-510 us instead of 33)
    for r in range(2 * len(envs)):
        est, lst, mk = ev.forward(ev_in[r % len(envs)][0], ev_in[r % len(envs)][1], n_j, n_m)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for k in range(k_ev):
        ei, dur = ev_in[k % len(envs)]
        est, lst, mk = ev.forward(ei, dur, n_j, n_m)
    e1.record(stream)
    barrier()
    ms_ev = e0.elapsed_time(e1)
    ev.check()
    assert torch.equal(mk.reshape(-1), envs[(k_ev - 1) % len(envs)].current_objs.reshape(-1))
    return ms_ev, k_ev


def time_evaluator_streams(envs, n_j, n_m, B, K, dev, barrier):
    """The same calls with one Evaluator (= one workspace) and one CUDA stream per independent input set: the two
    launches of a call stay ordered on their stream, calls on different streams overlap head to tail."""
    import torch
    from l2s_b200 import Evaluator
    n1 = n_j * n_m + 2
    main = torch.cuda.current_stream(dev)
    evs = [Evaluator(strict=False) for _ in envs]
    streams = [torch.cuda.Stream(device=dev) for _ in envs]
    ev_in = []
    for r in range(len(envs)):
        state, fa, done = envs[r].observe()
        ei = torch.cat([state[1], state[2]], dim=1).contiguous()
        dur = torch.zeros((B, n1), dtype=torch.int64, device=dev)
        dur[:, 1:-1] = torch.from_numpy(envs[r].instances[:, 0].reshape(B, -1).astype(np.int64)).to(dev)
        ev_in.append((ei, dur.reshape(-1, 1)))
    k_ev = max(24, min(K, 40))
    k_ev -= k_ev % len(envs)
    torch.cuda.synchronize(dev)
    outs = [None] * len(envs)
    for k in range(2 * len(envs)):
        r = k % len(envs)
        with torch.cuda.stream(streams[r]):
            outs[r] = evs[r].forward(ev_in[r][0], ev_in[r][1], n_j, n_m)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for st in streams:
        st.wait_event(e0)
    for k in range(k_ev):
        r = k % len(envs)
        with torch.cuda.stream(streams[r]):
            outs[r] = evs[r].forward(ev_in[r][0], ev_in[r][1], n_j, n_m)
    for st in streams:
        main.wait_stream(st)
    e1.record(main)
    barrier()
    for r, ev in enumerate(evs):
        ev.check()
        assert torch.equal(outs[r][2].reshape(-1), envs[r].current_objs.reshape(-1))
    return e0.elapsed_time(e1), k_ev


def time_evaluator_strict(envs, n_j, n_m, B, dev):
    """Default (strict) mode of the drop-in: one status read-back + sync per call, wall clock."""
    import torch
    from l2s_b200 import Evaluator
    n1 = n_j * n_m + 2
    ev = Evaluator()
    state, fa, done = envs[0].observe()
    ei = torch.cat([state[1], state[2]], dim=1).contiguous()
    dur = torch.zeros((B, n1), dtype=torch.int64, device=dev)
    dur[:, 1:-1] = torch.from_numpy(envs[0].instances[:, 0].reshape(B, -1).astype(np.int64)).to(dev)
    dur = dur.reshape(-1, 1)
    for _ in range(3):
        ev.forward(ei, dur, n_j, n_m)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(20):
        ev.forward(ei, dur, n_j, n_m)
    torch.cuda.synchronize(dev)
    return 1e3 * (time.perf_counter() - t0) / 20


def large_instances(rank, world, dev, lib, sp, stream, barrier):
    """BASELINE configs[3]: synthetic 100x20 and 150x25, batch 1024 sharded over the ranks (1024 / N per GPU)."""
    import torch
    import torch.distributed as dist
    peak, _ = measured_peak()
    out = {}
    for shape, (n_j, n_m) in (("100x20", (100, 20)), ("150x25", (150, 25))):
        Bg = int(os.environ.get("L2S_BENCH_LARGE_BATCH", "1024"))      # (override: reproduce an N-rank slice on one GPU)
        B = Bg // world
        Kt, Wt = 240, 8
        T = 2 * (-(-(Wt + Kt) // REPLICAS)) + 6
        rep = Replicas(n_j, n_m, B, T, rank, dev, seed0=1000 + n_j)
        ms = timed_device_steps(rep, Wt, Kt, lib, sp, stream, barrier, dev)
        barrier()
        ms_st = timed_device_steps_streams(rep, Wt, Kt, lib, dev)      # the headline schedule: a stream per environment
        barrier()
        pbar = float(torch.stack([o["npairs"] for o in rep.outs]).float().mean())
        ms_ev, k_ev = time_evaluator(rep.envs, n_j, n_m, B, 12, dev, stream, barrier)
        t = torch.tensor([ms, ms_ev, ms_st], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_ev, ms_st = [float(v) for v in t.cpu()]
        bps, bpg = bytes_per_step(n_j, n_m, pbar), bytes_per_graph(n_j, n_m)
        out[shape] = {"global_batch": Bg, "batch_per_gpu": B, "steps_timed": Kt, "us_per_step": 1e3 * ms / Kt,
                      "inst_steps_per_s": Bg * Kt / (ms / 1e3),
                      "roofline_frac": bps * B * Kt / (ms / 1e3) / 1e9 / peak,
                      "schedule": "us_per_step / inst_steps_per_s / roofline_frac: launches serialised on one stream (what one "
                                  "environment's sequential search sees); streams_*: the 4 independent environments on a "
                                  "stream each (the headline schedule of `value`)",
                      "streams_us_per_step": 1e3 * ms_st / Kt, "streams_inst_steps_per_s": Bg * Kt / (ms_st / 1e3),
                      "streams_roofline_frac": bps * B * Kt / (ms_st / 1e3) / 1e9 / peak,
                      "evaluator_graphs_per_s": Bg * k_ev / (ms_ev / 1e3), "evaluator_us_per_call": 1e3 * ms_ev / k_ev,
                      "evaluator_frac_of_peak": bpg * B * k_ev / (ms_ev / 1e3) / 1e9 / peak}
        del rep
        torch.cuda.empty_cache()
    return out


def learned_policy_rollouts(dev):
    """BASELINE configs[0] / configs[1]: the shipped GIN+DGHAN incumbent models rolled out on syn10x10 (100 instances,
    500 steps) and tai20x15 (10 instances, 5000 steps) -- the loop of test_learned.py:185-189 with the policy + step pair
    captured in one CUDA graph (l2s_b200.rollout.GraphedRollout).  Seconds per instance and optimality gaps beside the
    reference's own stored numbers (unstated GPU)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "examples"))
    from rollout_learned import rollout
    pol = os.path.join(ROOT, "tests", "golden", "policy")
    out = {}
    for name, data, weights, steps, ref in (
            ("syn10x10", "syn10x10", "incumbent_10x10_gin+dghan.pth", 500,
             {"s_per_instance": {"500": 1.25}, "gap": {"500": 0.0444}}),
            ("tai20x15", "tai20x15", "incumbent_20x15_gin+dghan.pth", 5000,
             {"s_per_instance": {"500": 2.93, "5000": 29.45}, "gap": {"500": 0.1163, "1000": 0.1042, "2000": 0.0945, "5000": 0.0831}})):
        inst = np.load(os.path.join(pol, data + ".npy"))
        opt = np.load(os.path.join(pol, data + "_result.npy"))
        rollout(inst, os.path.join(pol, weights), 20, mode="graph")          # warm-up (allocator, capture path)
        torch.cuda.synchronize(dev)
        res, secs = rollout(inst, os.path.join(pol, weights), steps, mode="graph")
        out[name] = {"instances": int(len(inst)), "steps": steps, "seconds": secs, "s_per_instance": secs / len(inst),
                     "inst_steps_per_s": len(inst) * steps / secs,
                     "gap": {str(k): float(((v - opt) / opt).mean()) for k, v in sorted(res.items())},
                     "reference_stored": ref}
    out["what"] = ("shipped incumbent models, sampling policy (seed 1), policy forward + l2s_step in one CUDA graph per "
                   "step; the reference's stored s_per_instance include its policy forward on an unstated GPU")
    return out


def training_episode(rank, world, dev, barrier):
    """BASELINE configs[4]: n-step REINFORCE on 10x10, 64 rollouts (sharded over the ranks), 500-step episode,
    gradient step every 10 environment steps, policy-gradient all-reduce."""
    import torch
    from l2s_b200 import JsspN5
    from l2s_b200.policy import Actor, shard_batchnorm
    from l2s_b200.training import train_batch
    torch.manual_seed(1)
    policy = Actor(3, 128, embedding_l=4, policy_l=4, embedding_type='gin+dghan', heads=1, dropout=0.0).to(dev)
    if world > 1:
        shard_batchnorm(policy)
    optimizer = torch.optim.Adam(policy.parameters(), lr=5e-5)
    torch.manual_seed(100 + rank)
    env = JsspN5(10, 10, 1, 99)
    inst = gen_instances(10, 10, 64, seed=77)
    train_batch(policy, optimizer, env, inst, transit=30, steps_learn=10, device=dev, rank=rank, world=world)   # warm-up
    barrier()
    timing = {}
    t0 = time.perf_counter()
    inc, cur, n_upd = train_batch(policy, optimizer, env, inst, transit=500, steps_learn=10, device=dev, rank=rank,
                                  world=world, timing=timing)
    torch.cuda.synchronize(dev)
    secs = time.perf_counter() - t0
    t = torch.tensor([secs, timing.get("allreduce_s", 0.0)], dtype=torch.float64, device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    secs, ar = [float(v) for v in t.cpu()]
    return {"shape": "10x10", "global_batch": 64, "batch_per_gpu": 64 // world, "episode_steps": 500,
            "gradient_steps": n_upd, "inst_steps_per_s": 64 * 500 / secs, "episode_s": secs,
            "allreduce_s": ar, "allreduce_share": ar / secs,
            "what": "policy forward (torch) + l2s_step per env step, loss/backward/Adam every 10 steps, one flattened "
                    "gradient all-reduce per update; BatchNorm statistics all-reduced per forward (sync_bn)"
                    if world > 1 else "policy forward (torch) + l2s_step per env step, loss/backward/Adam every 10 steps"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--shape", default="20x15")
    ap.add_argument("--batch", type=int, default=8192)
    ap.add_argument("--e2e-steps", type=int, default=48)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-large", action="store_true", help="skip the 100x20 / 150x25 section")
    ap.add_argument("--no-training", action="store_true", help="skip the REINFORCE episode")
    ap.add_argument("--no-learned", action="store_true", help="skip the learned-policy rollouts (configs 0 / 1)")
    ap.add_argument("--no-pin", action="store_true", help="do not bind ranks to their GPU's NUMA-local cores")
    ap.add_argument("--device-only", action="store_true", help="kernel-only timing (skips the e2e phases)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29531"] + sys.argv
        raise SystemExit(subprocess.call(cmd))
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
