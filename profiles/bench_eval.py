#!/usr/bin/env python
"""Evaluator.forward timing for one shape / batch (CUDA events):  python profiles/bench_eval.py 100x20 1024"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import torch  # noqa: E402

shape, B = sys.argv[1], int(sys.argv[2])
n_j, n_m = [int(v) for v in shape.split("x")]
dev = torch.device("cuda", 0)
rep = bench.Replicas(n_j, n_m, B, 8, 0, dev)
stream = torch.cuda.current_stream(dev)
ms, k = bench.time_evaluator(rep.envs, n_j, n_m, B, 40, dev, stream, lambda: torch.cuda.synchronize(dev))
peak, _ = bench.measured_peak()
print(json.dumps({"shape": shape, "batch": B, "us_per_call": 1e3 * ms / k, "graphs_per_s": B * k / (ms / 1e3),
                  "frac": bench.bytes_per_graph(n_j, n_m) * B * k / (ms / 1e3) / 1e9 / peak,
                  "env": {e: os.environ[e] for e in ("L2S_EVAL_CTA", "L2S_FAST_SWEEP") if e in os.environ}}))
