#!/usr/bin/env python
"""Diagnostic: where does the host-driven step (C-ABI `l2s_step_host`) lose time when N ranks run side by side?
Run under torchrun (or alone).  For every mode it times the device-resident `l2s_step` loop and the `l2s_step_host`
loop on every rank and prints the per-rank microseconds per step.

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 profiles/e2e_scaling.py [--steps 600]

Modes: completion by cudaStreamSynchronize or by a polled stream-ordered flag (L2S_HOST_WAIT=flag), ranks unpinned
or bound to their GPU's NUMA-local cores, records written by the kernel into host memory or staged through device
memory + one bulk D2H (L2S_HOST_RECORDS=0), actions read in place or copied H2D first (L2S_HOST_ZEROCOPY=0)."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=600)
    ap.add_argument("--shape", default="20x15")
    ap.add_argument("--batch", type=int, default=8192)
    ap.add_argument("--modes", default="pin,pin+lines,pin+prefetch,pin+lines+prefetch,pin+devrec,pin+h2d,pin+norec,pin+noact,pin+norec+noact")
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
    n_j, n_m = [int(v) for v in args.shape.split("x")]
    B, K = args.batch, args.steps
    stream = torch.cuda.current_stream(dev)
    sp = C.c_void_p(stream.cuda_stream)
    free_affinity = os.sched_getaffinity(0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    rows = []
    T = (len(args.modes.split(",")) + 1) * (-(-(8 + K) // bench.REPLICAS)) + 4
    rep = bench.Replicas(n_j, n_m, B, T, rank, dev)
    ms_dev = bench.timed_device_steps(rep, 8, K, lib, sp, stream, barrier, dev)
    host_tapes = [t.cpu().numpy() for t in rep.tapes]
    keep = [torch.from_numpy(np.ascontiguousarray(t)).pin_memory() for t in host_tapes]
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
    pos = list(rep.pos)
    desynced = False
    for mode in args.modes.split(","):
        parts = mode.split("+")
        opts = {"host_wait_flag": int("flag" in parts), "host_records_direct": int("devrec" not in parts),
                "host_zero_copy": int("h2d" not in parts), "host_record_lines": int("lines" in parts),
                "host_action_prefetch": int("prefetch" in parts),
                "host_debug": (1 if "norec" in parts else 0) | (2 if "noact" in parts else 0)}
        for h in handles:
            for k, v in opts.items():
                _capi.check(lib.l2s_set_option(h, k.encode(), v), "l2s_set_option")
        os.sched_setaffinity(0, free_affinity)
        aff = bench.pin_to_gpu_numa(local, world, force=True) if "pin" in parts else None
        calls = []
        for k in range(8 + K):
            r = k % bench.REPLICAS
            calls.append((r, tapes[r][pos[r]].ctypes.data))
            pos[r] += 1
        for r, ptr in calls[:8]:
            _capi.check(lib.l2s_step_host(handles[r], refs[r], ptr, None, None, None, None, None, None, sp), "host")
        barrier()
        t0 = time.perf_counter()
        for r, ptr in calls[8:]:
            lib.l2s_step_host(handles[r], refs[r], ptr, None, None, None, None, None, None, sp)
        torch.cuda.synchronize(dev)
        t_host = time.perf_counter() - t0
        # ("noact" replaces the taped moves by dummy moves: the tape is out of step from then on, so those modes go last)
        desynced = desynced or "noact" in parts
        assert "norec" in parts or desynced or not any((e._h_rec[:, 0] >> 24).any() for e in rep.envs)
        t = torch.tensor([1e3 * ms_dev / K, 1e6 * t_host / K], dtype=torch.float64, device=dev)
        allt = [torch.zeros_like(t) for _ in range(world)]
        if world > 1:
            dist.all_gather(allt, t)
        else:
            allt = [t]
        per = torch.stack(allt).cpu().numpy()
        rows.append({"mode": mode, "device_us": [round(float(v), 1) for v in per[:, 0]],
                     "host_us": [round(float(v), 1) for v in per[:, 1]],
                     "e2e_inst_steps_per_s": world * B / (per[:, 1].max() * 1e-6), "affinity": aff})
    if rank == 0:
        print(json.dumps({"world": world, "cores": len(free_affinity), "shape": args.shape, "batch": B, "steps": K,
                          "rows": rows}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
