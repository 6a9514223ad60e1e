"""N>1 host logic on CPU: world_size-2 gloo processes exercise the sharding helpers (slice arithmetic, the
final gather, the gradient all-reduce) and check that a sharded search gives the unsharded per-instance
results (the search itself is stood in for by the CPU oracle here; the GPU version is tests/test_gpu_*.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN, ROOT


def test_shard_range_partitions_everything():
    from l2s_b200.sharding import shard_range
    for total in (1, 7, 8, 10, 8192, 1000):
        for world in (1, 2, 3, 4, 8):
            pieces = [shard_range(total, r, world) for r in range(world)]
            assert pieces[0][0] == 0 and pieces[-1][1] == total
            assert all(pieces[i][1] == pieces[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in pieces]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes)      # remainder on the last ranks


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from l2s_b200.sharding import allreduce_mean_gradients, gather_concat, shard_range
    from oracle import c_oracle as CO
    g = np.load(os.path.join(GOLDEN, "traj_10x5.npz"))
    inst, mseq0 = g["instances"][:7], g["mseq0"][:7]            # 7 instances over 2 ranks: 3 + 4
    lo, hi = shard_range(len(inst), rank, world)
    env = CO.CEnv(inst[lo:hi], mseq0[lo:hi], inst_offset=lo)
    env.run("random", 30, seed=99)
    cur = gather_concat(torch.from_numpy(env.makespan.copy()).view(-1, 1))
    inc = gather_concat(torch.from_numpy(env.incumbent.copy()).view(-1, 1))
    # gradient all-reduce: rank r holds grad = r+1 -> mean over ranks
    lin = torch.nn.Linear(3, 2)
    for p in lin.parameters():
        p.grad = torch.full_like(p, float(rank + 1))
    allreduce_mean_gradients(list(lin.parameters()))
    ok_grad = all(torch.allclose(p.grad, torch.full_like(p, (world + 1) / 2)) for p in lin.parameters())
    if rank == 0:
        full = CO.CEnv(inst, mseq0, inst_offset=0)
        full.run("random", 30, seed=99)
        out["same_cur"] = bool(np.array_equal(cur.numpy().reshape(-1), full.makespan))
        out["same_inc"] = bool(np.array_equal(inc.numpy().reshape(-1), full.incumbent))
        out["ok_grad"] = ok_grad
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sharded_search_equals_unsharded():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert out.get("same_cur") and out.get("same_inc") and out.get("ok_grad"), dict(out)
