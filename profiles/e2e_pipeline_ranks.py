#!/usr/bin/env python
"""Diagnostic (run under torchrun): per-rank microseconds per step of the pipelined host-driven step
(`l2s_step_host_begin` / `l2s_step_host_wait`, 3 of 4 environments in flight, one stream each) when only some ranks
are active, with the core binding normal / reversed / off, and with whole-line records.  Separates "these cores are
slow" from "these GPUs share a PCIe uplink".

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 profiles/e2e_pipeline_ranks.py"""
import argparse
import collections
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

SCENARIOS = [  # (name, active ranks (None = all), affinity, record_lines, depth)
    ("all", None, "normal", 0, 3),
    ("all+lines", None, "normal", 1, 3),
    ("all+reversed-cores", None, "reversed", 0, 3),
    ("all+unpinned", None, "none", 0, 3),
    ("low-half", "low", "normal", 0, 3),
    ("high-half", "high", "normal", 0, 3),
    ("even", "even", "normal", 0, 3),
    ("all+depth2", None, "normal", 0, 2),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=800)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from l2s_b200 import _capi
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _capi.lib()
    n_j, n_m, B, K = 20, 15, 8192, args.steps
    free_affinity = os.sched_getaffinity(0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    T = (len(SCENARIOS) + 1) * (-(-(8 + K) // bench.REPLICAS)) + 4
    rep = bench.Replicas(n_j, n_m, B, T, rank, dev)
    main_stream = torch.cuda.current_stream(dev)
    keep = [torch.from_numpy(np.ascontiguousarray(t.cpu().numpy())).pin_memory() for t in rep.tapes]
    tapes = [t.numpy() for t in keep]
    ios = []
    for r in range(bench.REPLICAS):
        o = rep.outs[r]
        ios.append(_capi.StepIO(actions=None, high=99.0, fea_norm_const=1000.0, reward_type=0, is_reset=0,
                                x=o["x"].data_ptr(), edge_index_mc=o["emc"].data_ptr(), makespan=o["ms"].data_ptr(),
                                incumbent=o["inc"].data_ptr(), reward=o["rew"].data_ptr(), done=o["done"].data_ptr(),
                                pairs=None, npairs=None, status=None))
    refs = [C.byref(io) for io in ios]
    handles = [e._h for e in rep.envs]
    recs = [e._h_rec for e in rep.envs]
    streams = [torch.cuda.Stream(dev) for _ in range(bench.REPLICAS)]
    for st in streams:
        st.wait_stream(main_stream)
    sps = [C.c_void_p(s.cuda_stream) for s in streams]
    pos = list(rep.pos)
    rows = []
    for name, active, aff, lines, depth in SCENARIOS:
        on = (active is None or (active == "low" and rank < world // 2) or (active == "high" and rank >= world // 2)
              or (active == "even" and rank % 2 == 0))
        os.sched_setaffinity(0, free_affinity)
        bound = None
        if aff == "normal":
            bound = bench.pin_to_gpu_numa(local, world, force=True)
        elif aff == "reversed":
            bound = bench.pin_to_gpu_numa(world - 1 - local, world, force=True)    # (all GPUs report the same local cores here)
        for h in handles:
            _capi.check(lib.l2s_set_option(h, b"host_record_lines", lines), "opt")
        calls = []
        for k in range(8 + K):
            r = k % bench.REPLICAS
            calls.append((r, tapes[r][pos[r]].ctypes.data))
            if on:
                pos[r] += 1

        def run(seq):
            inflight = collections.deque()
            for r, ptr in seq:
                if len(inflight) == depth:
                    q = inflight.popleft()
                    lib.l2s_step_host_wait(handles[q], None, None, None, None, None, None)
                    _ = int(recs[q][0, 0]) + int(recs[q][-1, 0])
                _capi.check(lib.l2s_step_host_begin(handles[r], refs[r], ptr, sps[r]), "begin")
                inflight.append(r)
            while inflight:
                lib.l2s_step_host_wait(handles[inflight.popleft()], None, None, None, None, None, None)

        t = 0.0
        if on:
            run(calls[:8])
        barrier()
        if on:
            t0 = time.perf_counter()
            run(calls[8:])
            torch.cuda.synchronize(dev)
            t = time.perf_counter() - t0
            assert not any((e._h_rec[:, 0] >> 24).any() for e in rep.envs)
        barrier()
        tt = torch.tensor([1e6 * t / K], dtype=torch.float64, device=dev)
        allt = [torch.zeros_like(tt) for _ in range(world)]
        if world > 1:
            dist.all_gather(allt, tt)
        else:
            allt = [tt]
        rows.append({"scenario": name, "us_per_step": [round(float(v), 1) for v in torch.cat(allt).cpu()],
                     "bound_to_rank0": bound})
    if rank == 0:
        print(json.dumps({"world": world, "cores": len(free_affinity), "steps": K, "rows": rows}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
