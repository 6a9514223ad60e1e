"""Instance sharding across the GPUs of one node (one process per GPU, torch.distributed).

Instances are independent (no cross-instance term anywhere in the reference's env/), so a batch is cut into
contiguous slices, every rank steps its slice with its own handle and the identical kernels, and the only
collectives are (i) the final gather of makespans / incumbents and (ii) in training, the all-reduce of the
policy gradients before `optimizer.step()` (n-step_reinforce.py:58-61).  Nothing here runs inside the search
loop.  All helpers work with any backend (`nccl` on GPUs; `gloo` in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """Contiguous slice [lo, hi) of `total` instances for `rank`; the remainder goes to the LAST ranks."""
    base, rem = divmod(total, world)
    first_big = world - rem                       # ranks >= first_big hold base+1 instances
    lo = rank * base + max(0, rank - first_big)
    hi = lo + base + (1 if rank >= first_big else 0)
    return lo, hi


def gather_concat(t, group=None):
    """all_gather of per-rank tensors whose first dimension may differ; returns the rank-order concatenation
    (i.e. the tensor an unsharded run would hold)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return t
    world = dist.get_world_size(group)
    sizes = [torch.zeros(1, dtype=torch.int64, device=t.device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device), group=group)
    sizes = [int(s.item()) for s in sizes]
    mx = max(sizes)
    pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[:t.shape[0]] = t
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)], dim=0)


def allreduce_mean_gradients(parameters, group=None):
    """Average gradients over ranks with ONE flattened all-reduce (policy-gradient step of a data-parallel
    REINFORCE learner: every rank holds a slice of the rollout batch; n-step_reinforce.py:58-61)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    grads = [p.grad for p in parameters if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat /= dist.get_world_size(group)
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()


class ShardedJsspN5:
    """`JsspN5` over this rank's slice of a global batch.  `reset` takes the GLOBAL instance array; `step`
    takes this rank's actions.  `gather_objs()` returns the global (current, incumbent) makespans."""

    def __init__(self, n_job, n_mch, low, high, rank=None, world=None, **kw):
        from .environment import JsspN5
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        self.env = JsspN5(n_job, n_mch, low, high, **kw)
        self.lo = self.hi = 0

    def reset(self, instances, init_type, device, init_solutions=None):
        self.lo, self.hi = shard_range(len(instances), self.rank, self.world)
        self.env.instance_offset = self.lo
        sol = None if init_solutions is None else init_solutions[self.lo:self.hi]
        return self.env.reset(instances[self.lo:self.hi], init_type, device, init_solutions=sol)

    def step(self, actions, device):
        return self.env.step(actions, device)

    def run(self, *a, **kw):
        return self.env.run(*a, **kw)

    def gather_objs(self):
        return gather_concat(self.env.current_objs), gather_concat(self.env.incumbent_objs)
