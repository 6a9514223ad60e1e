"""Learned-policy rollouts with the (policy forward + environment step) pair captured in ONE CUDA graph
(SURVEY.md 8 f-1): no Python lists, no `.item()`, no per-kernel launch overhead.  One `graph.replay()` advances
every instance of the batch by one local-search step; the step kernel writes the next state into the same
static buffers the policy reads."""
import torch

from .environment import JsspN5  # noqa: F401


class GraphedRollout:
    def __init__(self, env, policy, greedy=False):
        self.env, self.policy, self.greedy = env, policy, greedy
        self.graph = None
        self.buf = None
        self._ident = None        # (handle address, pmax) the buffers and the captured graph belong to

    def reset(self, instances, init_type, device, init_solutions=None):
        env = self.env
        state, pairs, npairs, done = env.reset_device(instances, init_type, device, init_solutions=init_solutions)
        # the captured launch holds the handle's device pointers and its pmax: a new handle (different batch, a
        # regrown pair list after PAIR_OVERFLOW, close + reset) invalidates both the graph and the buffers
        ident = (env._h.value, env.pmax, tuple(state[0].shape))
        if self.buf is None or self._ident != ident:
            self.buf = env.device_buffers()
            self.graph = None
            self._ident = ident
        b = self.buf
        b["x"].copy_(state[0]); b["emc"].copy_(state[2]); b["pairs"].copy_(pairs); b["npairs"].copy_(npairs)
        b["done"].copy_(done); b["ms"].copy_(env.current_objs); b["inc"].copy_(env.incumbent_objs)
        env.current_objs, env.incumbent_objs = b["ms"], b["inc"]
        self.epc, self.batch = state[1], state[3]
        return self

    def _one_step(self):
        b = self.buf
        actions, _ = self.policy.forward_device(b["x"], self.epc, b["emc"], self.batch, b["pairs"], b["npairs"],
                                                greedy=self.greedy)
        self.env.step_device(actions, out=b)       # in place: the next replay reads what this one wrote

    def capture(self, warmup=3):
        """Warm up on a side stream (as CUDA graphs require), then capture one policy+step iteration.
        The warm-up iterations ARE search steps (env.itr advances)."""
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s), torch.no_grad():
            for _ in range(warmup):
                self._one_step()
        torch.cuda.current_stream().wait_stream(s)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self._one_step()
        self.env.itr -= 1        # the captured call only recorded the work; it did not advance the search
        self._sync_itr()
        return self

    def _sync_itr(self):
        """Graph replays do not pass through l2s_step, so the C handle's step counter (the RANDOM policy's RNG key)
        is set from the Python-side one."""
        from . import _capi
        _capi.check(_capi.lib().l2s_set_itr(self.env._h, int(self.env.itr)), "l2s_set_itr")

    def run(self, until_itr, log_steps=()):
        """Replay until env.itr == until_itr.  Returns {step: incumbents [B] (device tensor clone)}."""
        if self.graph is None:
            self.capture()
        env, logs = self.env, {}
        while env.itr < until_itr:
            self.graph.replay()
            env.itr += 1
            if env.itr in log_steps:
                logs[env.itr] = env.incumbent_objs.reshape(-1).clone()
        self._sync_itr()
        env.check_errors()
        return logs
