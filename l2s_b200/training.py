"""Data-parallel n-step REINFORCE on the GPU environment (the learner of the reference's
n-step_reinforce.py:40-61,118-175 and its validation / checkpoint logic :63-116, re-expressed with device tensors
and instance sharding).

Per rank: `batch_size / world` rollouts in one `JsspN5` handle, policy forward on the same GPU, no host round trip
inside the rollout (`step_device` + `Actor.forward_device`).  Every `steps_learn` environment steps the n-step
returns are normalised per instance over its not-yet-done steps and the REINFORCE loss is back-propagated.
Collectives (never inside the environment step):
  * one flattened all-reduce of the policy gradients (mean over ranks) before `optimizer.step()`
    (`sharding.allreduce_mean_gradients`);
  * with `sync_bn` (default when world > 1): the GIN BatchNorm layers take their batch statistics over ALL ranks'
    rows (`policy.ShardedBatchNorm1d`: 2 x 128 + 1 doubles per layer and forward).  The reference normalises with the
    statistics of the whole rollout batch (it never calls `.eval()`), so only with this are the sharded logits -- and,
    with equal slice sizes, the all-reduced gradient of the batch-mean loss -- those of an unsharded run (up to
    float rounding; tests/test_sharding_gloo.py compares them).  Without it every rank normalises with its own
    slice's statistics: still a valid learner, but not the reference's numbers.
Sampling differs by construction: every rank draws its actions from its own RNG stream.
"""
import os
import time

import numpy as np
import torch

from .environment import JsspN5
from .sharding import allreduce_mean_gradients, gather_concat, shard_range


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
                device='cuda', rank=0, world=1, timing=None):
    """One training batch (n-step_reinforce.py:147-175) on this rank's slice of `instances`.
    `timing` (optional dict) receives `allreduce_s`: host-observed seconds spent in the gradient all-reduce
    (device-synchronised around the call, so only pass it when measuring)."""
    lo, hi = shard_range(len(instances), rank, world)
    env.instance_offset = lo
    state, pairs, npairs, done = env.reset_device(instances[lo:hi], init_type, device)
    rewards, logps, dones = [], [], [done]
    n_updates = 0
    ar = 0.0
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
            if timing is not None and world > 1:
                torch.cuda.synchronize()
                t0 = time.perf_counter()
            allreduce_mean_gradients(list(policy.parameters()))      # the one collective of the learner
            if timing is not None and world > 1:
                torch.cuda.synchronize()
                ar += time.perf_counter() - t0
            optimizer.step()
            n_updates += 1
            rewards, logps, dones = [], [], [done]
    env.check_errors()
    if timing is not None:
        timing["allreduce_s"] = ar
    return env.incumbent_objs, env.current_objs, n_updates


def model_name(kind, j, m, l, h, init_type, reward_type, gamma, hidden_dim, embedding_layer, policy_layer,
               embedding_type, heads, drop_out, lr, steps_learn, transit, batch_size, episodes, step_validation):
    """File-name scheme of the reference's checkpoints (n-step_reinforce.py:85-107; the key `test_learned.py:141-148`
    rebuilds to load them)."""
    dghan = 'NAN' if embedding_type == 'gin' else '{}_{}'.format(heads, drop_out)
    return ('{}_{}x{}[{},{}]_{}_{}_{}_{}_{}_{}_{}_{}_{}_{}_{}_{}_{}_{}.pth'
            .format(kind, j, m, l, h, init_type, reward_type, gamma, hidden_dim, embedding_layer, policy_layer,
                    embedding_type, dghan, lr, steps_learn, transit, batch_size, episodes, step_validation))


class Validator:
    """Validation + checkpointing of the reference (`RL2S4JSSP.validation`, n-step_reinforce.py:63-116): roll the
    policy for `transit` steps on the fixed validation set, average incumbent and last-step makespans, and save the
    state dict under the reference's file names whenever either average improves.  Sharded: every rank rolls its
    slice of the validation set, the makespans are all-gathered, RANK 0 alone compares and writes."""

    def __init__(self, n_j, n_m, low, high, instances, out_dir, name_args, reward_type='yaoxin',
                 init_type='fdd-divide-mwkr', rank=0, world=1):
        self.env = JsspN5(n_j, n_m, low, high, reward_type=reward_type)
        self.instances, self.out_dir, self.name_args = instances, out_dir, name_args
        self.init_type, self.rank, self.world = init_type, rank, world
        self.incumbent_validation_result = np.inf
        self.current_validation_result = np.inf

    @torch.no_grad()
    def __call__(self, policy, transit, device):
        lo, hi = shard_range(len(self.instances), self.rank, self.world)
        env = self.env
        env.instance_offset = lo
        state, pairs, npairs, done = env.reset_device(self.instances[lo:hi], self.init_type, device)
        while env.itr < transit:
            actions, _ = policy.forward_device(*state, pairs, npairs)
            state, _, pairs, npairs, done = env.step_device(actions)
        env.check_errors()
        inc = gather_concat(env.incumbent_objs.reshape(-1)).mean().item()
        cur = gather_concat(env.current_objs.reshape(-1)).mean().item()
        saved = []
        if inc < self.incumbent_validation_result:
            self.incumbent_validation_result = inc
            saved.append('incumbent')
        if cur < self.current_validation_result:
            self.current_validation_result = cur
            saved.append('last-step')
        if self.rank == 0 and self.out_dir:
            os.makedirs(self.out_dir, exist_ok=True)
            for kind in saved:
                torch.save(policy.state_dict(), os.path.join(self.out_dir, model_name(kind, *self.name_args)))
        return inc, cur, saved
