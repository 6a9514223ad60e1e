#!/usr/bin/env python
"""Learned-policy rollout on the GPU environment, the loop of the reference's test_learned.py:185-189:

    while env.itr < horizon:  actions = policy(state, feasible_actions);  state = env.step(actions)

with the shipped GIN+DGHAN weights.  Two modes:
  --mode lists   : the reference's own API (python lists between env and policy, one host round trip per step)
  --mode device  : move sets and actions stay on the GPU (`step_device` + `Actor.forward_device`), no host sync.
  --mode graph   : device mode with the policy+step pair captured in one CUDA graph (`GraphedRollout`).

    python examples/rollout_learned.py --data tests/golden/policy/syn10x10.npy \
        --optimum tests/golden/policy/syn10x10_result.npy --weights tests/golden/policy/incumbent_10x10_gin+dghan.pth --steps 500
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from l2s_b200 import BatchGraph, JsspN5  # noqa: E402
from l2s_b200.policy import Actor  # noqa: E402


def rollout(instances, weights, steps, mode="device", seed=1, device="cuda", log_steps=(500, 1000, 2000, 5000)):
    torch.manual_seed(seed)
    n_j, n_m = instances.shape[2], instances.shape[3]
    env = JsspN5(n_j, n_m, 1, 99, reward_type='yaoxin', fea_norm_const=1000)
    policy = Actor(3, 128, embedding_l=4, policy_l=4, embedding_type='gin+dghan', heads=1, dropout=0.0).to(device)
    if weights:
        policy.load_state_dict(torch.load(weights, map_location=device))
    results, t0 = {}, time.time()
    with torch.no_grad():      # the env is not differentiated through; the reference only keeps log_prob for training
        if mode == "graph":
            from l2s_b200.rollout import GraphedRollout
            ro = GraphedRollout(env, policy).reset(instances, 'fdd-divide-mwkr', device)
            logs = ro.run(steps, log_steps=log_steps)
            for k, v in logs.items():
                results[k] = v.cpu().numpy()
        elif mode == "device":
            state, pairs, npairs, done = env.reset_device(instances, 'fdd-divide-mwkr', device)
            while env.itr < steps:
                actions, _ = policy.forward_device(*state, pairs, npairs)
                state, reward, pairs, npairs, done = env.step_device(actions)
                if env.itr in log_steps:
                    results[env.itr] = env.incumbent_objs.cpu().numpy().reshape(-1)
            env.check_errors()
        else:
            batch_data = BatchGraph()
            state, feasible_actions, done = env.reset(instances, 'fdd-divide-mwkr', device)
            while env.itr < steps:
                batch_data.wrapper(*state)
                actions, _ = policy(batch_data, feasible_actions)
                state, reward, feasible_actions, done = env.step(actions, device)
                if env.itr in log_steps:
                    results[env.itr] = env.incumbent_objs.cpu().numpy().reshape(-1)
    torch.cuda.synchronize()
    results[env.itr] = env.incumbent_objs.cpu().numpy().reshape(-1)
    return results, time.time() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--data", default="tests/golden/policy/syn10x10.npy")
    ap.add_argument("--optimum", default="tests/golden/policy/syn10x10_result.npy")
    ap.add_argument("--weights", default="tests/golden/policy/incumbent_10x10_gin+dghan.pth")
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--mode", default="device", choices=["device", "lists", "graph"])
    args = ap.parse_args()
    inst = np.load(args.data)
    res, secs = rollout(inst, args.weights, args.steps, args.mode)
    opt = np.load(args.optimum) if args.optimum else None
    for k in sorted(res):
        gap = "" if opt is None else "  mean gap to optimum %.2f %%" % (100 * ((res[k] - opt) / opt).mean())
        print("after %5d steps: mean incumbent makespan %.1f%s" % (k, res[k].mean(), gap))
    print("%d instances x %d steps in %.2f s (%s mode): %.0f inst-steps/s incl. the policy network"
          % (len(inst), args.steps, secs, args.mode, len(inst) * args.steps / secs))


if __name__ == "__main__":
    main()
