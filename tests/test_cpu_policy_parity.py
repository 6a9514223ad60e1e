"""Learned-policy parity at LOGIT level (north star: "same final makespans within the stated float tolerance on
policy logits"): the reference's own, unmodified `model/actor.py:Actor` (imported from /root/reference or the
vendored `baseline/_ref`, running on the repo's pure-torch torch_geometric drop-in) against `l2s_b200.policy.Actor`
with the same shipped weights on the same states.

Tolerances (float32; the two implementations order the reductions differently: bmm over all (n+2)^2 pairs + masked
softmax vs. gathered dot products over the feasible pairs):
  * per-node action embeddings (the operands of the score bmm, |h| <= ~3): max abs difference <= 1e-5;
  * log pi of the chosen action: <= 2e-6 * max|score| (the scores are dot products of magnitude ~130-150, whose
    float32 spacing is 1.5e-5, so an absolute 1e-5 would be below one ulp of the logits).
Greedy (argmax) trajectories must be IDENTICAL.  The environment here is the reference's own CPU `JsspN5`, so the
whole test runs without a GPU; tests/test_gpu_policy.py repeats it with the CUDA environment.
"""
import os

import numpy as np
import pytest
import torch

from oracle import ref_harness

pytestmark = pytest.mark.skipif(not ref_harness.available(), reason="reference tree / baseline/_ref not present")

TOL_H = 1e-5          # action embeddings
TOL_REL = 2e-6        # log-probabilities, relative to the largest |score|
MODEL = "saved_model/incumbent_{s}[1,99]_fdd-divide-mwkr_yaoxin_1_128_4_4_gin+dghan_1_0.0_5e-05_10_500_64_128000_10.pth"


def _actors(R, shape, device="cpu"):
    from l2s_b200.policy import Actor
    kw = dict(embedding_l=4, policy_l=4, embedding_type='gin+dghan', heads=1, dropout=0.0)
    sd = torch.load(os.path.join(R.root, MODEL.format(s=shape)), map_location=device)
    ref = R.Actor(3, 128, **kw).to(device)
    ours = Actor(3, 128, **kw).to(device)
    ref.load_state_dict(sd, strict=True)
    ours.load_state_dict(sd, strict=True)
    return ref, ours


def pairs_tensors(fa, device):
    B, P = len(fa), max(len(f) for f in fa)
    pairs = np.zeros((B, P, 2), dtype=np.int32)
    cnt = np.zeros(B, dtype=np.int32)
    for i, f in enumerate(fa):
        if len(f) == 1 and int(f[0][0]) == 0 and int(f[0][1]) == 0:
            continue
        cnt[i] = len(f)
        pairs[i, :len(f)] = np.asarray(f, dtype=np.int32)
    return torch.from_numpy(pairs).to(device), torch.from_numpy(cnt).to(device)


def ours_log_probs(ours, state, fa):
    """log pi over the feasible pairs [B, P] from l2s_b200.policy (same maths as forward_device)."""
    dev = state[0].device
    pairs, npairs = pairs_tensors(fa, dev)
    B, P = pairs.shape[:2]
    h = ours.node_features(state[0], state[1], state[2], B)
    idx = pairs.long()
    ha = torch.gather(h, 1, idx[:, :, 0:1].expand(B, P, h.shape[-1]))
    hb = torch.gather(h, 1, idx[:, :, 1:2].expand(B, P, h.shape[-1]))
    score = (ha * hb).sum(-1)
    valid = torch.arange(P, device=dev).unsqueeze(0) < npairs.unsqueeze(1)
    return torch.log_softmax(score.masked_fill(~valid, -float("inf")), dim=1), pairs, npairs


class _Greedy(torch.distributions.Categorical):
    """Test-side stand-in for the sampler only (the reference keeps its argmax line commented out,
    model/actor.py:235): `sample()` returns the mode."""

    def sample(self, sample_shape=torch.Size()):
        return self.probs.argmax(dim=-1)


def run_parity(R, ref, ours, env_reset, env_step, steps, greedy, seed=0):
    """Drive the env with the REFERENCE actor's actions; at every step compare its log-probabilities with ours."""
    saved = R.actor_module.Categorical
    if greedy:
        R.actor_module.Categorical = _Greedy
    try:
        torch.manual_seed(seed)
        state, fa, _ = env_reset()
        bg = R.BatchGraph()
        worst = 0.0
        grab = {}
        hook = ref.policy[-1].register_forward_hook(lambda m, i, o: grab.__setitem__("h", o))
        for _ in range(steps):
            bg.wrapper(*state)
            with torch.no_grad():
                acts, lp_ref = ref(bg, fa)
                lsm, pairs, npairs = ours_log_probs(ours, state, fa)
                h_ours = ours.node_features(state[0], state[1], state[2], len(fa))
                assert float((grab["h"] - h_ours).abs().max()) <= TOL_H, "action embeddings differ"
                scale = float(torch.bmm(h_ours, h_ours.transpose(1, 2)).abs().max())
                if greedy:
                    a_ours, lp_ours = ours.forward_device(*state, pairs, npairs, greedy=True)
                    assert a_ours.cpu().tolist() == [[int(a[0]), int(a[1])] for a in acts], "greedy actions differ"
            for b, a in enumerate(acts):
                if int(npairs[b]) == 0:
                    assert [int(a[0]), int(a[1])] == [0, 0]
                    continue
                j = [[int(p[0]), int(p[1])] for p in fa[b]].index([int(a[0]), int(a[1])])
                worst = max(worst, abs(float(lsm[b, j]) - float(lp_ref.reshape(-1)[b])) / scale)
            state, _, fa, _ = env_step(acts)
        hook.remove()
        return worst
    finally:
        R.actor_module.Categorical = saved


@pytest.mark.parametrize("shape,data,B,steps", [("10x10", "syn10x10", 6, 40), ("20x15", "tai20x15", 3, 12)])
@pytest.mark.parametrize("greedy", [False, True])
def test_reference_actor_logits_equal_ours_on_reference_env(shape, data, B, steps, greedy):
    if greedy:
        steps = 100          # identical greedy trajectories over 100 steps on syn10x10 and tai20x15
    R = ref_harness.load()
    n_j, n_m = [int(v) for v in shape.split("x")]
    inst = np.load(os.path.join(R.root, "test_data", data + ".npy"))[:B]
    ref, ours = _actors(R, shape)
    np.random.seed(1)
    env = R.JsspN5(n_j, n_m, 1, 99)
    worst = run_parity(R, ref, ours, lambda: env.reset(inst, "fdd-divide-mwkr", "cpu"),
                       lambda acts: env.step(acts, "cpu"), steps, greedy)
    assert worst <= TOL_REL, "relative log-prob difference %.3g exceeds %.0e" % (worst, TOL_REL)


def test_pyg_shim_layers_against_dense_formulas():
    """The drop-in's GINConv / GATConv / MessagePassing(max) against dense matrix formulas on a small graph."""
    ref_harness.load()
    import torch_geometric
    from torch_geometric.nn import GATConv, GINConv, global_mean_pool
    from torch_geometric.nn.conv import MessagePassing
    from torch_geometric.utils import add_self_loops
    assert getattr(torch_geometric, "_l2s_shim", False)
    torch.manual_seed(3)
    N, F_in = 7, 5
    ei = torch.tensor([[0, 1, 2, 3, 4, 5, 1, 6], [1, 2, 3, 4, 5, 6, 4, 0]])
    x = torch.randn(N, F_in)
    A = torch.zeros(N, N)
    A[ei[1], ei[0]] = 1.0                                        # A[i, j] = 1 for an edge j -> i
    # GIN, mean aggregation with self loops appended (model/actor.py:179-182)
    lin = torch.nn.Linear(F_in, 4)
    gin = GINConv(lin, eps=0, train_eps=False, aggr='mean', flow="source_to_target")
    ei_l = add_self_loops(ei)[0]
    Al = A + torch.eye(N)
    want = lin((Al @ x) / Al.sum(1, keepdim=True) + x)
    assert torch.allclose(gin(x, ei_l), want, atol=1e-6)
    # GAT, 2 heads, concat
    gat = GATConv(F_in, 3, heads=2, dropout=0.0, concat=True)
    h = gat.lin_l(x).view(N, 2, 3)
    al, ar = (h * gat.att_l).sum(-1), (h * gat.att_r).sum(-1)
    e = torch.nn.functional.leaky_relu(al.unsqueeze(0) + ar.unsqueeze(1), 0.2)        # e[i, j, head]
    e = e.masked_fill((Al == 0).unsqueeze(-1), -float("inf"))
    alpha = torch.softmax(e, dim=1)
    want = torch.einsum("ijh,jhc->ihc", alpha, h).reshape(N, 6) + gat.bias
    assert torch.allclose(gat(x, ei_l), want, atol=1e-6)
    assert torch.allclose(gat(x, ei), want, atol=1e-6)           # remove-then-add self loops: same result
    # max aggregation with zero fill, both flows (env/message_passing_evl.py:123-153)
    v = torch.tensor([[3.], [1.], [4.], [1.], [5.], [9.], [2.]])
    fwd = MessagePassing(aggr='max', flow='source_to_target').propagate(ei, x=v)
    bwd = MessagePassing(aggr='max', flow='target_to_source').propagate(ei, x=v)
    want_f = torch.stack([(v[ei[0][ei[1] == i]].max() if (ei[1] == i).any() else torch.tensor(0.)) for i in range(N)])
    want_b = torch.stack([(v[ei[1][ei[0] == i]].max() if (ei[0] == i).any() else torch.tensor(0.)) for i in range(N)])
    assert torch.equal(fwd.view(-1), want_f) and torch.equal(bwd.view(-1), want_b)
    batch = torch.tensor([0, 0, 0, 1, 1, 1, 1])
    assert torch.allclose(global_mean_pool(x, batch), torch.stack([x[:3].mean(0), x[3:].mean(0)]), atol=1e-6)
