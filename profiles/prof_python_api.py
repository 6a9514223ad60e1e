import sys, time, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from l2s_b200 import JsspN5
dev = torch.device("cuda", 0)
rep = bench.Replicas(20, 15, 8192, 40, 0, dev)
tapes = [t.cpu().numpy() for t in rep.tapes]
lists = [(k % 4, tapes[k % 4][k // 4].tolist()) for k in range(100)]
for r, a in lists[:8]:
    rep.envs[r].step(a, dev)
torch.cuda.synchronize()
t0 = time.perf_counter()
for r, a in lists[8:60]:
    s, rw, fa, d = rep.envs[r].step(a, dev)
    x = len(fa) + len(fa[0])
torch.cuda.synchronize()
print("us per step", (time.perf_counter() - t0) / 52 * 1e6)
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for r, a in lists[60:100]:
    s, rw, fa, d = rep.envs[r].step(a, dev)
    x = len(fa) + len(fa[0])
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
