#!/usr/bin/env python
"""Benchmark of the L2S hot path: N5 local-search steps per second (instances x steps / s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--shape 20x15] [--batch 8192]

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): batched synthetic 20x15
instances, durations U{1..98} (`uni_instance_gen(20,15,1,99)` semantics), batch 8192 PER GPU (weak scaling:
instances are independent, one batch slice per rank, no data-path collective), fdd-divide-mwkr initial
solutions built on the device, actions = a recorded tape of uniformly random feasible N5 moves.

A "step" is ONE `l2s_step` launch over one batch: swap + re-evaluation (est, lst, makespan) + features x +
machine edges + reward/incumbent/tabu + critical path + N5 move set for every instance of the batch.

  value : device-resident throughput (state and action tape already in HBM), CUDA events, max over ranks.
  e2e   : the same steps through the C-ABI `l2s_step_host` with HOST buffers: the kernel reads the host actions
          over PCIe from pinned memory and writes one result record per instance (move set, reward, makespan,
          incumbent, done, status) over PCIe into pinned host memory; stream synchronised every step.
  e2e_python_api : the same through the drop-in `JsspN5.step(actions, device)` (Python lists in and out).
  evaluator : `Evaluator.forward` on the int64 COO edge lists of the same schedules (graphs/s, GB/s).
  roofline : algorithmic bytes of one step (SURVEY.md 8d formula) / mean launch duration vs measured HBM peak.
  cpu_baseline : the CPU oracle port of the reference path (oracle/l2s_oracle.py) on a bounded sample.

`--impl reference` times the CPU port of the reference's own path (the reference is pure Python and cannot
travel to the GPU box) on all host cores and prints the same JSON line with "impl": "reference".
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "N5 LS steps/sec (instances x steps)"
UNIT = "inst-steps/s"
REPLICAS = 4   # rotating independent batches so that the per-step working set exceeds the 126 MB L2


def workload_name(shape, batch):
    return ("syn%s U{1..98}, batch %d per GPU, fdd-divide-mwkr init, replayed random feasible N5 moves; one step = one "
            "environment step (l2s_step launch) over the batch" % (shape, batch))


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


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


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


# ------------------------------------------------------------------------------------------------------
# CPU port of the reference path (oracle) -- timed baseline
# ------------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    n_j, n_m, B, steps, warm, seed = args
    from oracle import l2s_oracle as O
    inst = gen_instances(n_j, n_m, B, seed)
    env = O.OracleEnv(n_j, n_m, 1, 99)
    _, fa, _ = env.reset(inst, "fdd-divide-mwkr")
    rnd = np.random.RandomState(seed)
    total = 0.0
    for k in range(warm + steps):
        acts = [f[rnd.randint(len(f))] for f in fa]
        t0 = time.perf_counter()
        _, _, fa, _ = env.step(acts)
        if k >= warm:
            total += time.perf_counter() - t0
    return B * steps, total


def cpu_port_rate(n_j, n_m, B, steps, warm, procs):
    """inst-steps/s of OracleEnv.step (env/environment.py:356-387 restated) with `procs` processes."""
    if procs <= 1:
        units, secs = _cpu_worker((n_j, n_m, B, steps, warm, 1))
        return units / secs
    import multiprocessing as mp
    per = max(1, B // procs)
    with mp.get_context("spawn").Pool(procs) as pool:
        t0 = time.perf_counter()
        res = pool.map(_cpu_worker, [(n_j, n_m, per, steps, warm, 1 + i) for i in range(procs)])
        wall = time.perf_counter() - t0
    # every process times only its env.step loop; the job rate is the sum of the per-process rates
    return sum(u / s for u, s in res)


def run_reference(args, rank, world):
    if rank != 0:
        return
    n_j, n_m = [int(v) for v in args.shape.split("x")]
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 64))
    per = 4 if n_j * n_m <= 400 else 1
    t0 = time.perf_counter()
    rate = cpu_port_rate(n_j, n_m, per * procs, max(1, args.steps), max(0, min(args.warmup, 3)), procs)
    wall = time.perf_counter() - t0
    units = per * procs * args.steps
    # additionally: the plain-C restatement (oracle/l2s_oracle.c) on all cores -- the strongest CPU number we have
    c_port = None
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
        c_port = {"value": cb * ksteps / (time.perf_counter() - t1), "unit": UNIT, "threads": procs,
                  "what": "oracle/l2s_oracle.c orc_run_batch (OpenMP): per step swap + O(n) topological evaluation + "
                          "features + machine edges + critical path + N5 pairs, %d instances x %d steps" % (cb, ksteps)}
    except Exception as e:   # the C oracle is optional for this line
        c_port = {"error": str(e)[:200]}
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * units / rate / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": workload_name(args.shape, args.batch), "shape": args.shape, "batch_per_gpu": args.batch,
                   "sample": "each timed step advances %d instances of that workload (bounded sample; throughput "
                             "is per instance-step, so it does not depend on the batch size)" % (per * procs)},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": procs, "kind": "port",
                         "sample": "%d processes x %d instances x %d env.step calls of the numpy/Python oracle port "
                                   "(oracle/l2s_oracle.py OracleEnv.step); the reference itself is pure Python "
                                   "(networkx + torch CPU) and cannot travel to this box" % (procs, per, args.steps)},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": wall, "c_port": c_port,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from l2s_b200 import JsspN5, _capi

    n_j, n_m = [int(v) for v in args.shape.split("x")]
    B, K, W = args.batch, args.steps, args.warmup
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (l2s_b200 has no CPU path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _capi.lib()
    n, n1, e_mc = n_j * n_m, n_j * n_m + 2, n_m * (n_j - 1)

    # phases per replica: [device-resident: W+K steps spread over replicas][e2e python][e2e capi]
    per_rep = -(-(W + K) // REPLICAS)
    T = 3 * per_rep
    envs, tapes, outs = [], [], []
    for r in range(REPLICAS):
        env = JsspN5(n_j, n_m, 1, 99)
        env.instance_offset = (rank * REPLICAS + r) * B
        inst = gen_instances(n_j, n_m, B, seed=1 + rank * REPLICAS + r)
        env.reset(inst, "fdd-divide-mwkr", dev)
        init = env.solutions().clone()
        _, taken = env.run("random", T, seed=7 + r, return_actions=True)   # record the action tape
        tape = taken.permute(1, 0, 2).contiguous()                          # [T, B, 2] device
        env.reset(inst, "fdd-divide-mwkr", dev, init_solutions=init.cpu().numpy())
        envs.append(env)
        tapes.append(tape)
        outs.append(dict(x=torch.empty((B * n1, 3), dtype=torch.float32, device=dev),
                         emc=torch.empty((2, B * e_mc), dtype=torch.int64, device=dev),
                         ms=torch.empty(B, dtype=torch.float32, device=dev),
                         inc=torch.empty(B, dtype=torch.float32, device=dev),
                         rew=torch.empty(B, dtype=torch.float32, device=dev),
                         done=torch.empty(B, dtype=torch.uint8, device=dev),
                         pairs=torch.empty((B, env.pmax, 2), dtype=torch.int32, device=dev),
                         npairs=torch.empty(B, dtype=torch.int32, device=dev),
                         status=torch.empty(B, dtype=torch.int32, device=dev)))
    stream = torch.cuda.current_stream(dev)
    sp = C.c_void_p(stream.cuda_stream)

    def io_for(r, t):
        o = outs[r]
        return _capi.StepIO(actions=tapes[r][t].data_ptr(), high=99.0, fea_norm_const=1000.0, reward_type=0,
                            is_reset=0, x=o["x"].data_ptr(), edge_index_mc=o["emc"].data_ptr(),
                            makespan=o["ms"].data_ptr(), incumbent=o["inc"].data_ptr(), reward=o["rew"].data_ptr(),
                            done=o["done"].data_ptr(), pairs=o["pairs"].data_ptr(), npairs=o["npairs"].data_ptr(),
                            status=o["status"].data_ptr())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- phase 1: device-resident ------------------------------------------------------------------
    pos = [0] * REPLICAS
    ios = []
    for k in range(W + K):
        r = k % REPLICAS
        ios.append((r, io_for(r, pos[r])))
        pos[r] += 1
    for r, io in ios[:W]:
        _capi.check(lib.l2s_step(envs[r]._h, C.byref(io), sp), "l2s_step")
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for r, io in ios[W:]:
        lib.l2s_step(envs[r]._h, C.byref(io), sp)
    if world > 1:   # the only collective of a search: final gather of makespans and incumbents
        res = torch.stack([outs[(W + K - 1) % REPLICAS]["ms"], outs[(W + K - 1) % REPLICAS]["inc"]])
        gathered = [torch.empty_like(res) for _ in range(world)]
        dist.all_gather(gathered, res)
    e1.record(stream)
    barrier()
    ms_dev = e0.elapsed_time(e1)
    st = torch.stack([o["status"] for o in outs])
    assert int(st.abs().sum()) == 0, "device-side error during the timed region"
    pbar = float(torch.stack([o["npairs"] for o in outs]).float().mean())
    if args.device_only:
        if rank == 0:
            sampler.stop()
            print(json.dumps({"shape": args.shape, "batch": B, "launch_us": 1e3 * ms_dev / K,
                              "value": world * B * K / (ms_dev / 1e3)}), flush=True)
        return

    # ---- phase 2: end to end through the Python drop-in API ------------------------------------------
    # sync the Python-side step counters with the C handle, then time JsspN5.step with host lists
    host_tapes = [t.cpu().numpy() for t in tapes]
    k_e2e = max(1, min(K, args.e2e_steps))
    w_e2e = min(W, 3)
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
        sink += len(fa)
    torch.cuda.synchronize(dev)
    t_py = time.perf_counter() - t0
    pos = p2

    # ---- phase 3: end to end through the C-ABI with host buffers -------------------------------------
    # the step's inputs start in pinned host memory (one [B,2] int32 action block per step)
    pinned_keep = [torch.from_numpy(np.ascontiguousarray(t)).pin_memory() for t in host_tapes]
    pinned_tapes = [t.numpy() for t in pinned_keep]
    hb = []
    p3 = list(pos)
    for k in range(w_e2e + K):
        r = k % REPLICAS
        hb.append((r, pinned_tapes[r][p3[r]]))
        p3[r] += 1
    # host results are read in place from the handle's pinned records (l2s_host_records): zero extra copies
    sink = 0

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
    for r, ptr in calls[w_e2e:]:
        capi_step(r, ptr)
    torch.cuda.synchronize(dev)
    t_capi = time.perf_counter() - t0
    assert not any((e._h_rec[:, 0] >> 24).any() for e in envs)

    # ---- phase 4: stateless evaluator (Evaluator.forward drop-in) on the current schedules ----------------
    from l2s_b200 import Evaluator
    ev = Evaluator(strict=False)      # asynchronous calls; the status word is checked once after the loop
    ev_in = []
    for r in range(REPLICAS):
        state, fa, done = envs[r].observe()
        ei = torch.cat([state[1], state[2]], dim=1).contiguous()
        dur = torch.zeros((B, n1), dtype=torch.int64, device=dev)
        dur[:, 1:-1] = torch.from_numpy(envs[r].instances[:, 0].reshape(B, -1).astype(np.int64)).to(dev)
        ev_in.append((ei, dur.reshape(-1, 1)))
    k_ev = max(4, min(K, 40))
    for r in range(min(REPLICAS, 3)):
        ev.forward(ev_in[r][0], ev_in[r][1], n_j, n_m)
    barrier()
    e0.record(stream)
    for k in range(k_ev):
        ei, dur = ev_in[k % REPLICAS]
        est, lst, mk = ev.forward(ei, dur, n_j, n_m)
    e1.record(stream)
    barrier()
    ms_ev = e0.elapsed_time(e1)
    ev.check()
    assert torch.equal(mk.reshape(-1), envs[(k_ev - 1) % REPLICAS].current_objs.reshape(-1))

    # ---- phase 5: fused K-step search (l2s_run): on-device policy, state resident in shared memory ------------
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

    # ---- reduce over ranks ---------------------------------------------------------------------------
    times = torch.tensor([ms_dev / 1e3, t_py, t_capi, ms_ev / 1e3, ms_fused / 1e3, ms_greedy / 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    t_dev, t_py, t_capi, t_ev, t_fused, t_greedy = [float(v) for v in times.cpu()]
    value = world * B * K / t_dev
    e2e_py = world * B * k_e2e / t_py
    e2e_capi = world * B * K / t_capi

    if rank == 0:
        peak, peak_src = measured_peak()
        bps = bytes_per_step(n_j, n_m, pbar)
        launch_s = t_dev / K
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
            cb, cs = (64, 150) if n <= 400 else (4, 10)
            t0 = time.perf_counter()
            rate = cpu_port_rate(n_j, n_m, cb, cs, 2, 1)
            cpu = {"value": rate, "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": "%d instances x %d env.step calls of the numpy/Python oracle port of the reference "
                             "path (oracle/l2s_oracle.py), %.1f s" % (cb, cs, time.perf_counter() - t0)}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": 1e3 * t_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic",
            "config": {"workload": workload_name(args.shape, B),
                       "shape": args.shape, "batch_per_gpu": B, "global_batch": world * B,
                       "l2": "%d independent batches stepped round-robin: per-step working set x%d > 126 MB L2"
                             % (REPLICAS, REPLICAS),
                       "mean_pairs": pbar, "parallelism": "instance-sharded x%d" % world},
            "e2e": {"value": e2e_capi, "unit": UNIT,
                    "api": "C-ABI l2s_step_host: host int32 actions in (read in place from pinned memory by the kernel), "
                           "host records out (pairs/npairs/done/reward/makespan/incumbent/status, written into pinned "
                           "memory by the kernel), sync per step",
                    "h2d_bytes_per_step": B * 8, "d2h_bytes_per_step": int(B * (16 + 4 * pbar)),
                    "steps": K},
            "e2e_python_api": {"value": e2e_py, "unit": UNIT,
                               "api": "drop-in JsspN5.step(actions: list[B][2], device) -> (state, reward, list[B][P][2], done); "
                                      "dominated by building ~%d Python lists per step, as the reference API requires" % int(B * (1 + pbar)),
                               "steps": k_e2e},
            "evaluator": {"graphs_per_s": world * B * k_ev / t_ev, "ms_per_call": 1e3 * t_ev / k_ev,
                          "api": "Evaluator.forward(edge_index int64 COO, duration int64, n_j, n_m)",
                          "algorithmic_bytes_per_graph": 8 * n1 + 16 * (n_j * (n_m + 1) + e_mc) + 8 * n1 + 4,
                          "achieved_gbs": (8 * n1 + 16 * (n_j * (n_m + 1) + e_mc) + 8 * n1 + 4) * B * k_ev / t_ev / 1e9 / 1.0,
                          "frac_of_peak": (8 * n1 + 16 * (n_j * (n_m + 1) + e_mc) + 8 * n1 + 4) * B * k_ev / t_ev / 1e9 / peak},
            "fused_run": {"random_policy_inst_steps_per_s": world * B * k_f * REPLICAS / t_fused,
                          "greedy_policy_inst_steps_per_s": world * B * k_g / t_greedy,
                          "api": "l2s_run: K steps per launch, on-device policy, no host round trip, no x/edge outputs",
                          "steps_per_launch": k_f},
            "gpu_launches": K,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": "l2s::step_kernel",
                         "algorithmic_bytes_per_launch": bps * B, "launch_ms": 1e3 * launch_s},
            "cpu_baseline": cpu, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--shape", default="20x15")
    ap.add_argument("--batch", type=int, default=8192)
    ap.add_argument("--e2e-steps", type=int, default=12)
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
