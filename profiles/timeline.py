"""Per-instance timeline of one `l2s::step_kernel` launch (20x15): when every instance starts / ends (globaltimer)
and how long each phase takes (clock64), from a build with -DL2S_TIMELINE (lane 0 of every warp stores 11 stamps).
Diagnostic build only -- never used for bench numbers.   usage: python profiles/timeline.py [batch] [shape]"""
import os, subprocess, sys, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TL_LIB = os.path.join(ROOT, "l2s_b200", "lib", "libl2s_b200_timeline.so")
_src = os.path.join(ROOT, "l2s_b200", "csrc")
if not os.path.isfile(TL_LIB) or any(os.path.getmtime(os.path.join(_src, f)) > os.path.getmtime(TL_LIB)
                                     for f in ("l2s_capi.cu", "l2s_kernels.cuh", "l2s_device.cuh")):
    subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--shared",
                    "-Xcompiler", "-fPIC", "-DL2S_TIMELINE", "-o", TL_LIB, "l2s_capi.cu"],
                   cwd=os.path.join(ROOT, "l2s_b200", "csrc"), check=True)
os.environ["L2S_B200_LIB"] = TL_LIB
import torch
sys.path.insert(0, ROOT)
from l2s_b200 import JsspN5, _capi
from bench import gen_instances
lib = _capi.lib()
raw = C.CDLL(TL_LIB)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
n_j, n_m = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "20x15").split("x")]
dev = torch.device("cuda", 0)
env = JsspN5(n_j, n_m, 1, 99)
env.reset(gen_instances(n_j, n_m, B, 1), "fdd-divide-mwkr", dev)
init = env.solutions().clone()
_, taken = env.run("random", 40, seed=7, return_actions=True)
tape = taken.permute(1, 0, 2).contiguous()
env.reset(gen_instances(n_j, n_m, B, 1), "fdd-divide-mwkr", dev, init_solutions=init.cpu().numpy())
out = env.device_buffers()
tl = torch.zeros((B, 12), dtype=torch.int64, device=dev)
raw.l2s_debug_timeline.argtypes = [C.c_void_p]
# big buffer to flush L2 between steps
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for k in range(30):
    if k == 29:
        assert raw.l2s_debug_timeline(C.c_void_p(tl.data_ptr())) == 0
        torch.cuda.synchronize()
        raw.l2s_debug_levels(None, 1)
    flush.zero_()
    env.step_device(tape[k], out=out)
torch.cuda.synchronize()
t = tl.cpu().numpy().astype(np.int64)
lv = (C.c_ulonglong * 2)()
raw.l2s_debug_levels(lv, 0)
print("sweep levels per wavefront call: mean %.1f over %d calls" % (lv[0] / max(1, lv[1]), lv[1]))
g0, g1 = t[:, 0], t[:, 1]
base = g0.min()
print("launch span us", (g1.max() - base) / 1e3)
start = (g0 - base) / 1e3; end = (g1 - base) / 1e3
print("start  pct 0/25/50/75/100:", np.percentile(start, [0, 25, 50, 75, 100]).round(1))
print("end    pct 0/25/50/75/100:", np.percentile(end, [0, 25, 50, 75, 100]).round(1))
dur = end - start
first = start < 2.0
print("instances starting in first 2us:", first.sum(), " later:", (~first).sum())
names = ["stage+search", "swap", "evaluate", "x out", "link (+emc, warp kernel)", "trace (CTA kernel: overlaps emc)", "moves", "scalars+rec"]
staged = (t[:, 11] - t[:, 2]) / 1.965e3   # warp kernel only: TMA staging complete (part of stage+search)
ck = t[:, 2:11].astype(np.float64)
ph = np.diff(ck, axis=1) / 1.965e3   # us at 1965 MHz
for grp, m in (("first wave", first), ("later", ~first)):
    if m.any() and (t[m, 11] > 0).all():
        print(grp, "of which staging wait (TMA complete) %.2f us" % staged[m].mean())
    print(grp, "latency us mean %.2f" % dur[m].mean(), " phases:", " | ".join("%s %.2f" % (n, v) for n, v in zip(names, ph[m].mean(axis=0))))
# histogram of active instances over time
ts = np.arange(0, end.max(), 2.0)
print("active instances at t:", [(float(x), int(((start <= x) & (end > x)).sum())) for x in ts])
