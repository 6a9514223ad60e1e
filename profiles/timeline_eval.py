"""Per-graph timeline of one `l2s::coo_eval_kernel` launch (Evaluator.forward), diagnostic build (-DL2S_TIMELINE).
usage: python profiles/timeline_eval.py [batch] [shape]"""
import os, subprocess, sys, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TL_LIB = os.path.join(ROOT, "l2s_b200", "lib", "libl2s_b200_timeline.so")
if not os.path.isfile(TL_LIB):
    subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--shared",
                    "-Xcompiler", "-fPIC", "-DL2S_TIMELINE", "-o", TL_LIB, "l2s_capi.cu"],
                   cwd=os.path.join(ROOT, "l2s_b200", "csrc"), check=True)
os.environ["L2S_B200_LIB"] = TL_LIB
import torch
sys.path.insert(0, ROOT)
from l2s_b200 import JsspN5, Evaluator, _capi
from bench import gen_instances
raw = C.CDLL(TL_LIB)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
n_j, n_m = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "20x15").split("x")]
dev = torch.device("cuda", 0)
env = JsspN5(n_j, n_m, 1, 99)
state, fa, done = env.reset(gen_instances(n_j, n_m, B, 1), "fdd-divide-mwkr", dev)
n1 = n_j * n_m + 2
ei = torch.cat([state[1], state[2]], dim=1).contiguous()
dur = torch.zeros((B, n1), dtype=torch.int64, device=dev)
dur[:, 1:-1] = torch.from_numpy(env.instances[:, 0].reshape(B, -1).astype(np.int64)).to(dev)
dur = dur.reshape(-1, 1)
ev = Evaluator(strict=False)
tl = torch.zeros((B, 12), dtype=torch.int64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
raw.l2s_debug_timeline.argtypes = [C.c_void_p]
for k in range(6):
    if k == 5:
        assert raw.l2s_debug_timeline(C.c_void_p(tl.data_ptr())) == 0
    flush.zero_()
    est, lst, mk = ev.forward(ei, dur, n_j, n_m)
torch.cuda.synchronize()
ev.check()
t = tl.cpu().numpy().astype(np.int64)
g0, g1 = t[:, 0], t[:, 1]
base = g0.min()
start, end = (g0 - base) / 1e3, (g1 - base) / 1e3
print("eval kernel span us %.1f" % end.max())
print("start pct 0/25/50/75/100:", np.percentile(start, [0, 25, 50, 75, 100]).round(1))
print("end   pct 0/25/50/75/100:", np.percentile(end, [0, 25, 50, 75, 100]).round(1))
names = ["durations -> node words", "wait for TMA", "heads", "chains -> walk rows", "sweeps", "outputs"]
ph = np.diff(t[:, 2:9].astype(np.float64), axis=1) / 1.965e3
first = start < 2.0
for grp, m in (("first wave", first), ("later", ~first)):
    if m.any():
        print(grp, "(%d graphs) latency us mean %.2f" % (m.sum(), (end - start)[m].mean()), " phases:",
              " | ".join("%s %.2f" % (n, v) for n, v in zip(names, ph[m].mean(axis=0))))
