#!/usr/bin/env python
"""Diagnostic: the host-driven step with the call split in two (`l2s_step_host_begin` / `l2s_step_host_wait`) and
1-3 of the 4 independent environments in flight, on one stream or on a stream per environment, with the action
prefetch kernel and whole-line records on or off.  Prints microseconds per step for every combination.

    python profiles/e2e_pipeline.py [--steps 1200] [--shape 20x15] [--batch 8192]"""
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=1200)
    ap.add_argument("--shape", default="20x15")
    ap.add_argument("--batch", type=int, default=8192)
    args = ap.parse_args()
    import torch
    from l2s_b200 import _capi
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    lib = _capi.lib()
    n_j, n_m = [int(v) for v in args.shape.split("x")]
    B, K = args.batch, args.steps
    combos = [(d, own, pf, ln) for own in (0, 1) for d in (1, 2, 3) for pf in (0, 1) for ln in (0, 1)]
    T = (len(combos) + 1) * (-(-(8 + K) // bench.REPLICAS)) + 4
    rep = bench.Replicas(n_j, n_m, B, T, 0, dev)
    main_stream = torch.cuda.current_stream(dev)
    sp = C.c_void_p(main_stream.cuda_stream)
    ms_dev = bench.timed_device_steps(rep, 8, K, lib, sp, main_stream, lambda: torch.cuda.synchronize(dev), dev)
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
    own_streams = [torch.cuda.Stream(dev) for _ in range(bench.REPLICAS)]
    pos = list(rep.pos)
    bench.pin_to_gpu_numa(0, 1, force=True)
    rows = []
    for depth, own, pf, ln in combos:
        for h in handles:
            _capi.check(lib.l2s_set_option(h, b"host_action_prefetch", pf), "opt")
            _capi.check(lib.l2s_set_option(h, b"host_record_lines", ln), "opt")
        sps = [C.c_void_p(s.cuda_stream) for s in own_streams] if own else [sp] * bench.REPLICAS
        calls = []
        for k in range(8 + K):
            r = k % bench.REPLICAS
            calls.append((r, tapes[r][pos[r]].ctypes.data))
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
                q = inflight.popleft()
                lib.l2s_step_host_wait(handles[q], None, None, None, None, None, None)

        torch.cuda.synchronize(dev)
        run(calls[:8])
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        run(calls[8:])
        torch.cuda.synchronize(dev)
        t = time.perf_counter() - t0
        assert not any((e._h_rec[:, 0] >> 24).any() for e in rep.envs)
        rows.append({"depth": depth, "own_streams": own, "action_prefetch": pf, "record_lines": ln,
                     "us_per_step": round(1e6 * t / K, 2), "inst_steps_per_s": B * K / t})
    print(json.dumps({"shape": args.shape, "batch": B, "steps": K, "device_us_per_step": 1e3 * ms_dev / K, "rows": rows}))


if __name__ == "__main__":
    main()
