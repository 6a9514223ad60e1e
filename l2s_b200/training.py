"""Data-parallel n-step REINFORCE on the GPU environment (the learner of the reference's
n-step_reinforce.py:40-61,118-175, re-expressed with device tensors and instance sharding).

Per rank: `batch_size / world` rollouts in one `JsspN5` handle, policy forward on the same GPU, no host round trip
inside the rollout (`step_device` + `Actor.forward_device`).  Every `steps_learn` environment steps the n-step
returns are normalised per instance over its not-yet-done steps and the REINFORCE loss is back-propagated; the
ONLY collective is one flattened all-reduce of the policy gradients (mean over ranks) before `optimizer.step()`
(`sharding.allreduce_mean_gradients`).  With equal slice sizes this equals the reference's batch-mean loss.
"""
import numpy as np
import torch

from .environment import JsspN5
from .sharding import allreduce_mean_gradients, shard_range


def reinforce_loss(rewards, log_probs, dones, gamma=1.0, eps=float(np.finfo(np.float32).eps)):
    """n-step REINFORCE loss of the reference's `learn` (n-step_reinforce.py:40-59), vectorised.
    rewards / log_probs / dones: lists (length n) of [B,1] tensors; `dones[t]` is the done flag observed BEFORE
    action t.  Returns the batch mean of sum_t( -log_prob * normalised return ) over not-done steps."""
    R = torch.zeros_like(rewards[0])
    returns = []
    for r in reversed(rewards):
        R = r + gamma * R
        returns.insert(0, R)
    returns = torch.cat(returns, dim=-1)                 # [B, n]
    alive = ~torch.cat(dones, dim=-1)                    # [B, n]
    logp = torch.cat(log_probs, dim=-1)
    cnt = alive.sum(dim=1, keepdim=True).clamp(min=1)
    mean = (returns * alive).sum(dim=1, keepdim=True) / cnt
    var = (((returns - mean) ** 2) * alive).sum(dim=1, keepdim=True) / cnt      # unbiased=False
    norm = (returns - mean) / (var.sqrt() + eps)
    return (-(logp * norm) * alive).sum(dim=1).mean()


def train_batch(policy, optimizer, env, instances, transit=500, steps_learn=10, gamma=1.0, init_type='fdd-divide-mwkr',
                device='cuda', rank=0, world=1):
    """One training batch (n-step_reinforce.py:147-175) on this rank's slice of `instances`."""
    lo, hi = shard_range(len(instances), rank, world)
    env.instance_offset = lo
    state, pairs, npairs, done = env.reset_device(instances[lo:hi], init_type, device)
    rewards, logps, dones = [], [], [done]
    n_updates = 0
    while env.itr < transit:
        actions, logp = policy.forward_device(*state, pairs, npairs)
        state, reward, pairs, npairs, done = env.step_device(actions)
        rewards.append(reward)
        logps.append(logp)
        dones.append(done)
        if env.itr % steps_learn == 0:
            loss = reinforce_loss(rewards, logps, dones[:-1], gamma)
            optimizer.zero_grad()
            loss.backward()
            allreduce_mean_gradients(list(policy.parameters()))      # the one collective of training
            optimizer.step()
            n_updates += 1
            rewards, logps, dones = [], [], [done]
    env.check_errors()
    return env.incumbent_objs, env.current_objs, n_updates
