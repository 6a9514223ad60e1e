"""Real multi-GPU parity (skipped unless the box has >= 2 GPUs): 2 NCCL ranks drive `ShardedJsspN5` over a global
batch and the gathered per-instance results must equal the 1-rank run bit for bit -- under a REPLAYed action tape
(step API), the on-device RANDOM policy (keyed by the GLOBAL instance index) and the GREEDY policy."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT

pytestmark = pytest.mark.gpu

WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["L2S_ROOT"])
from l2s_b200 import JsspN5
from l2s_b200.sharding import ShardedJsspN5, gather_concat, shard_range
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
g = np.load(os.path.join(os.environ["L2S_GOLDEN"], "traj_10x10.npz"))
inst, mseq0, tape = g["instances"][:7], g["mseq0"][:7], g["tape"][:7]          # 7 instances: ranks hold 3 + 4
K = tape.shape[1]
lo, hi = shard_range(len(inst), rank, world)
out = {}
# (1) step API with the recorded tape
env = ShardedJsspN5(10, 10, 1, 99)
env.reset(inst, "fdd-divide-mwkr", dev, init_solutions=mseq0)
for k in range(K):
    state, reward, fa, done = env.step(tape[lo:hi, k].tolist(), dev)
cur, inc = env.gather_objs()
out["step_cur"], out["step_inc"] = cur.cpu().numpy().reshape(-1), inc.cpu().numpy().reshape(-1)
out["x"] = gather_concat(state[0].reshape(hi - lo, -1)).cpu().numpy()
# (2) fused policies
for pol in ("random", "greedy"):
    env = ShardedJsspN5(10, 10, 1, 99)
    env.reset(inst, "fdd-divide-mwkr", dev, init_solutions=mseq0)
    logs, taken = env.run(pol, 40, seed=5, log_steps=(10, 40), return_actions=True)
    out[pol + "_inc"] = gather_concat(logs.t().contiguous()).cpu().numpy()
    out[pol + "_taken"] = gather_concat(taken).cpu().numpy()
# (3) sharded BatchNorm on the GPU (torch's fused SyncBatchNorm function underneath): embeddings and all-reduced
#     gradients of the 2-rank run == the unsharded run, which every rank recomputes locally with plain BatchNorm
sys.path.insert(0, os.path.join(os.environ["L2S_ROOT"], "tests"))
from test_sharding_gloo import _policy_state, _slice_state
from l2s_b200.policy import Actor, shard_batchnorm
from l2s_b200.sharding import allreduce_mean_gradients
n_j, n_m, B = 5, 4, 6
n1, e_pc, e_mc = n_j * n_m + 2, n_j * (n_m + 1), n_m * (n_j - 1)
st = tuple(t.to(dev) for t in _policy_state(B, n_j, n_m, seed=4))
torch.manual_seed(7)
kw = dict(embedding_l=2, policy_l=2, embedding_type='gin+dghan', heads=1, dropout=0.0)
ref, sharded = Actor(3, 32, **kw).to(dev), Actor(3, 32, **kw).to(dev)
sharded.load_state_dict(ref.state_dict())
shard_batchnorm(sharded)
def loss_of(model, s):
    h = model.node_features(s[0], s[1], s[2], s[4].shape[0])
    idx = s[4].long()
    ha = torch.gather(h, 1, idx[:, :, 0:1].expand(-1, -1, h.shape[-1]))
    hb = torch.gather(h, 1, idx[:, :, 1:2].expand(-1, -1, h.shape[-1]))
    return h, -torch.log_softmax((ha * hb).sum(-1), dim=1)[:, 0].mean()
h_full, loss_full = loss_of(ref, st)
loss_full.backward()
blo, bhi = rank * B // world, (rank + 1) * B // world
h_loc, loss_loc = loss_of(sharded, _slice_state(st, blo, bhi, n1, e_pc, e_mc))
loss_loc.backward()
allreduce_mean_gradients(list(sharded.parameters()))
gmax = max(float(p.grad.abs().max()) for p in ref.parameters() if p.grad is not None)
errs = torch.tensor([float((h_loc - h_full[blo:bhi]).detach().abs().max()),
                     max(float((a.grad - b.grad).abs().max()) / gmax for a, b in zip(sharded.parameters(), ref.parameters())
                         if b.grad is not None),
                     max(float((dict(sharded.named_buffers())[k] - v).abs().max()) for k, v in ref.named_buffers()
                         if v.dtype.is_floating_point)], device=dev)
dist.all_reduce(errs, op=dist.ReduceOp.MAX)
out["bn_errs"] = errs.cpu().numpy()
if rank == 0:
    np.savez(os.environ["L2S_OUT"], **out)
dist.barrier()
dist.destroy_process_group()
'''


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.timeout(600)
@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_two_nccl_ranks_equal_one_rank(tmp_path):
    from l2s_b200 import JsspN5
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    outp = str(tmp_path / "out.npz")
    envv = dict(os.environ, L2S_ROOT=ROOT, L2S_GOLDEN=GOLDEN, L2S_OUT=outp)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), str(script)]
    res = subprocess.run(cmd, env=envv, capture_output=True, text=True, timeout=500)
    assert res.returncode == 0, res.stderr[-3000:]
    got = np.load(outp)
    g = np.load(os.path.join(GOLDEN, "traj_10x10.npz"))
    inst, mseq0, tape = g["instances"][:7], g["mseq0"][:7], g["tape"][:7]
    K = tape.shape[1]
    env = JsspN5(10, 10, 1, 99)
    env.reset(inst, "fdd-divide-mwkr", "cuda:0", init_solutions=mseq0)
    for k in range(K):
        state, reward, fa, done = env.step(tape[:, k].tolist(), "cuda:0")
    assert np.array_equal(got["step_cur"], env.current_objs.cpu().numpy().reshape(-1))
    assert np.array_equal(got["step_inc"], env.incumbent_objs.cpu().numpy().reshape(-1))
    assert np.array_equal(got["step_cur"].astype(np.int32), g["makespan"][K][:7])        # and the reference's numbers
    assert np.array_equal(got["x"], state[0].reshape(7, -1).cpu().numpy())
    for pol in ("random", "greedy"):
        env = JsspN5(10, 10, 1, 99)
        env.reset(inst, "fdd-divide-mwkr", "cuda:0", init_solutions=mseq0)
        logs, taken = env.run(pol, 40, seed=5, log_steps=(10, 40), return_actions=True)
        assert np.array_equal(got[pol + "_inc"], logs.t().cpu().numpy()), pol
        assert np.array_equal(got[pol + "_taken"], taken.cpu().numpy()), pol
    err_h, err_g, err_buf = [float(v) for v in got["bn_errs"]]
    assert err_h < 1e-5 and err_g < 1e-4 and err_buf < 1e-5, (err_h, err_g, err_buf)
