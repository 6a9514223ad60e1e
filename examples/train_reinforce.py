#!/usr/bin/env python
"""n-step REINFORCE training with the GPU environment (config 5 of BASELINE.json: 10x10, batch 64, 500-step
episodes, gradient step every 10 env steps).  Single GPU:  python examples/train_reinforce.py --batches 3
Multi GPU:  torchrun --nproc-per-node N --master-addr 127.0.0.1 examples/train_reinforce.py --batches 3"""
import argparse
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from l2s_b200 import JsspN5  # noqa: E402
from l2s_b200.policy import Actor  # noqa: E402
from l2s_b200.training import train_batch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--j", type=int, default=10)
    ap.add_argument("--m", type=int, default=10)
    ap.add_argument("--batch_size", type=int, default=64)
    ap.add_argument("--transit", type=int, default=500)
    ap.add_argument("--steps_learn", type=int, default=10)
    ap.add_argument("--batches", type=int, default=3)
    ap.add_argument("--lr", type=float, default=5e-5)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(1)                       # identical initial weights on every rank
    policy = Actor(3, 128, embedding_l=4, policy_l=4, embedding_type='gin+dghan', heads=1, dropout=0.0).to(dev)
    optimizer = torch.optim.Adam(policy.parameters(), lr=args.lr)
    torch.manual_seed(100 + rank)              # different action samples per rank
    env = JsspN5(args.j, args.m, 1, 99)
    rng = np.random.RandomState(1)
    for b in range(args.batches):
        dur = rng.randint(1, 99, size=(args.batch_size, args.j, args.m))
        mch = rng.random_sample((args.batch_size, args.j, args.m)).argsort(axis=2) + 1
        inst = np.stack([dur, mch], axis=1).astype(np.int32)
        t0 = time.time()
        inc, cur, n_upd = train_batch(policy, optimizer, env, inst, args.transit, args.steps_learn, device=dev,
                                      rank=rank, world=world)
        torch.cuda.synchronize()
        if rank == 0:
            print("batch %d: %d gradient steps, mean incumbent %.1f, mean final %.1f, %.2f s (%.0f inst-steps/s per rank)"
                  % (b, n_upd, inc.mean().item(), cur.mean().item(), time.time() - t0,
                     len(inst) / world * args.transit / (time.time() - t0)))
    if world > 1:
        # all ranks must hold identical weights after the all-reduced updates
        flat = torch.cat([p.detach().reshape(-1) for p in policy.parameters()])
        ref = flat.clone()
        dist.broadcast(ref, 0)
        assert torch.equal(flat, ref), "ranks diverged"
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
