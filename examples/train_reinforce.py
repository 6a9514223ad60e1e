#!/usr/bin/env python
"""n-step REINFORCE training with the GPU environment (config 5 of BASELINE.json: 10x10, batch 64, 500-step
episodes, gradient step every 10 env steps; reference: n-step_reinforce.py).
Single GPU:  python examples/train_reinforce.py --batches 3
Multi GPU:   python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 examples/train_reinforce.py --batches 3
Every rank rolls batch_size / N instances; the collectives are the BatchNorm-statistics all-reduce inside the policy
forward (so the logits are those of the unsharded batch), one gradient all-reduce per update, and the all-gather of
the validation makespans; rank 0 validates / checkpoints under the reference's file names."""
import argparse
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from l2s_b200 import JsspN5  # noqa: E402
from l2s_b200.policy import Actor, shard_batchnorm  # noqa: E402
from l2s_b200.training import Validator, train_batch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--j", type=int, default=10)
    ap.add_argument("--m", type=int, default=10)
    ap.add_argument("--l", type=int, default=1)
    ap.add_argument("--h", type=int, default=99)
    ap.add_argument("--init_type", default="fdd-divide-mwkr")
    ap.add_argument("--reward_type", default="yaoxin")
    ap.add_argument("--gamma", type=float, default=1)
    ap.add_argument("--hidden_dim", type=int, default=128)
    ap.add_argument("--embedding_layer", type=int, default=4)
    ap.add_argument("--policy_layer", type=int, default=4)
    ap.add_argument("--embedding_type", default="gin+dghan")
    ap.add_argument("--heads", type=int, default=1)
    ap.add_argument("--drop_out", type=float, default=0.)
    ap.add_argument("--lr", type=float, default=5e-5)
    ap.add_argument("--steps_learn", type=int, default=10)
    ap.add_argument("--transit", type=int, default=500)
    ap.add_argument("--batch_size", type=int, default=64)
    ap.add_argument("--episodes", type=int, default=128000)
    ap.add_argument("--step_validation", type=int, default=10)
    ap.add_argument("--batches", type=int, default=3, help="stop after this many batches (0: episodes / batch_size)")
    ap.add_argument("--validation_data", default=None, help=".npy [V,2,j,m]; default: 100 fresh instances (seed 3)")
    ap.add_argument("--out", default=None, help="directory for the checkpoints (reference file names)")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(1)                       # identical initial weights on every rank
    policy = Actor(3, args.hidden_dim, embedding_l=args.embedding_layer, policy_l=args.policy_layer,
                   embedding_type=args.embedding_type, heads=args.heads, dropout=args.drop_out).to(dev)
    if world > 1:
        shard_batchnorm(policy)                # batch statistics over all ranks: the logits of the unsharded batch
    optimizer = torch.optim.Adam(policy.parameters(), lr=args.lr)
    torch.manual_seed(100 + rank)              # different action samples per rank
    env = JsspN5(args.j, args.m, args.l, args.h, reward_type=args.reward_type)
    rng = np.random.RandomState(1)             # every rank draws the same global batch and takes its slice

    def gen(n, r):
        dur = r.randint(args.l, args.h, size=(n, args.j, args.m))
        mch = r.random_sample((n, args.j, args.m)).argsort(axis=2) + 1
        return np.stack([dur, mch], axis=1).astype(np.int32)

    val_data = np.load(args.validation_data) if args.validation_data else gen(100, np.random.RandomState(3))
    name_args = (args.j, args.m, args.l, args.h, args.init_type, args.reward_type, args.gamma, args.hidden_dim,
                 args.embedding_layer, args.policy_layer, args.embedding_type, args.heads, args.drop_out, args.lr,
                 args.steps_learn, args.transit, args.batch_size, args.episodes, args.step_validation)
    validate = Validator(args.j, args.m, args.l, args.h, val_data, args.out, name_args, reward_type=args.reward_type,
                         init_type=args.init_type, rank=rank, world=world)
    n_batches = args.batches or args.episodes // args.batch_size
    log, vlog = [], []
    for b in range(1, n_batches + 1):
        inst = gen(args.batch_size, rng)
        t0 = time.time()
        inc, cur, n_upd = train_batch(policy, optimizer, env, inst, args.transit, args.steps_learn, gamma=args.gamma,
                                      init_type=args.init_type, device=dev, rank=rank, world=world)
        torch.cuda.synchronize()
        dt = time.time() - t0
        log.append(cur.mean().item())
        if rank == 0:
            print("Batch %d training takes: %.2f  Mean Performance: %.2f  (%d gradient steps, %.0f inst-steps/s over %d rank(s))"
                  % (b, dt, log[-1], n_upd, len(inst) * args.transit / dt, world), flush=True)
        if b % args.step_validation == 0 or b == n_batches:
            t0 = time.time()
            v1, v2, saved = validate(policy, args.transit, dev)
            vlog.append([v1, v2])
            if rank == 0:
                print("Incumbent objs and final step objs for validation are: %.2f  %.2f  validation takes:%.2f  saved: %s"
                      % (v1, v2, time.time() - t0, saved or "-"), flush=True)
    if world > 1:
        # all ranks must hold identical weights after the all-reduced updates
        flat = torch.cat([p.detach().reshape(-1) for p in policy.parameters()])
        ref = flat.clone()
        dist.broadcast(ref, 0)
        assert torch.equal(flat, ref), "ranks diverged"
        if rank == 0:
            print("all %d ranks hold identical weights" % world)
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
