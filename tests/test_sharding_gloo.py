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


def _policy_state(B, n_j, n_m, seed):
    """Synthetic policy inputs of B equal-sized graphs (the policy is pure torch, so this runs on CPU)."""
    rng = np.random.RandomState(seed)
    n1 = n_j * n_m + 2
    x = torch.from_numpy(rng.rand(B * n1, 3).astype(np.float32))
    pcs, mcs, pairs = [], [], []
    for b in range(B):
        u = np.arange(1, n_j * n_m + 1)
        v = np.where(u % n_m != 0, u + 1, n1 - 1)
        first = np.arange(n_j) * n_m + 1
        pcs.append(np.stack([np.concatenate([np.zeros(n_j, np.int64), u]), np.concatenate([first, v])]) + b * n1)
        perm = rng.permutation(u)[:n_m * (n_j - 1) + 1]
        mcs.append(np.stack([perm[:-1], perm[1:]]) + b * n1)
        pairs.append(rng.randint(1, n_j * n_m, size=(4, 2)))
    pc = torch.from_numpy(np.concatenate(pcs, axis=1)).long()
    mc = torch.from_numpy(np.concatenate(mcs, axis=1)).long()
    batch = torch.arange(B).repeat_interleave(n1)
    return x, pc, mc, batch, torch.from_numpy(np.stack(pairs)).int(), torch.full((B,), 4, dtype=torch.int32)


def _slice_state(st, lo, hi, n1, e_pc, e_mc):
    x, pc, mc, batch, pairs, npairs = st
    return (x[lo * n1:hi * n1], pc[:, lo * e_pc:hi * e_pc] - lo * n1, mc[:, lo * e_mc:hi * e_mc] - lo * n1,
            batch[lo * n1:hi * n1] - lo, pairs[lo:hi], npairs[lo:hi])


def _bn_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from l2s_b200.policy import Actor, shard_batchnorm
    from l2s_b200.sharding import allreduce_mean_gradients
    n_j, n_m, B = 5, 4, 6
    n1, e_pc, e_mc = n_j * n_m + 2, n_j * (n_m + 1), n_m * (n_j - 1)
    st = _policy_state(B, n_j, n_m, seed=4)
    torch.manual_seed(7)
    ref = Actor(3, 32, embedding_l=2, policy_l=2, embedding_type='gin+dghan', heads=1, dropout=0.0)
    sharded = Actor(3, 32, embedding_l=2, policy_l=2, embedding_type='gin+dghan', heads=1, dropout=0.0)
    sharded.load_state_dict(ref.state_dict())
    shard_batchnorm(sharded)

    def loss_of(model, s):
        h = model.node_features(s[0], s[1], s[2], s[4].shape[0])
        idx = s[4].long()
        ha = torch.gather(h, 1, idx[:, :, 0:1].expand(-1, -1, h.shape[-1]))
        hb = torch.gather(h, 1, idx[:, :, 1:2].expand(-1, -1, h.shape[-1]))
        logp = torch.log_softmax((ha * hb).sum(-1), dim=1)
        return h, -logp[:, 0].mean()                  # batch-mean loss, like the REINFORCE learner

    # unsharded run (every rank computes it locally: plain BatchNorm over all B graphs)
    h_full, loss_full = loss_of(ref, st)
    loss_full.backward()
    # sharded run: rank r holds B/world graphs, BatchNorm statistics all-reduced, gradients averaged
    lo, hi = rank * B // world, (rank + 1) * B // world
    h_loc, loss_loc = loss_of(sharded, _slice_state(st, lo, hi, n1, e_pc, e_mc))
    loss_loc.backward()
    allreduce_mean_gradients(list(sharded.parameters()))
    err_h = float((h_loc - h_full[lo:hi]).detach().abs().max())
    # relative to the largest gradient entry of the model (a Linear bias in front of a BatchNorm has a gradient that
    # is zero up to rounding noise, so a per-tensor relative error would compare noise with noise)
    gmax = max(float(b.grad.abs().max()) for b in ref.parameters() if b.grad is not None)
    err_g = max(float((a.grad - b.grad).abs().max()) / gmax
                for a, b in zip(sharded.parameters(), ref.parameters()) if b.grad is not None)
    bufs = max(float((dict(sharded.named_buffers())[k] - v).abs().max()) for k, v in ref.named_buffers()
               if v.dtype.is_floating_point)
    # control: per-rank statistics (plain BatchNorm on the slice) do NOT reproduce the unsharded embeddings
    plain = Actor(3, 32, embedding_l=2, policy_l=2, embedding_type='gin+dghan', heads=1, dropout=0.0)
    plain.load_state_dict(ref.state_dict())
    h_plain, _ = loss_of(plain, _slice_state(st, lo, hi, n1, e_pc, e_mc))
    out["r%d" % rank] = (err_h, err_g, bufs, float((h_plain - h_full[lo:hi]).detach().abs().max()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sharded_batchnorm_reproduces_unsharded_embeddings_and_gradients():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_bn_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    for r in range(2):
        err_h, err_g, bufs, err_plain = out["r%d" % r]
        assert err_h < 1e-5, "embeddings differ from the unsharded run: %g" % err_h
        assert err_g < 1e-4, "all-reduced gradients differ from the unsharded run: %g (relative)" % err_g
        assert bufs < 1e-5, "BatchNorm running statistics differ: %g" % bufs
        assert err_plain > 1e-3, "control failed: per-rank statistics should change the embeddings"
