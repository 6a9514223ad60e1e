"""Share of executed warp instructions / stall samples per phase of l2s::step_kernel.
Input: `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` export.  Phase boundaries are found by
searching the CURRENT sources for their marker lines, so run it against the sources the profile was built from.
usage: by_phase.py file.csv [warp_instructions_per_launch instances_per_launch]"""
import csv
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MARKS = {
    "l2s_device.cuh": [("smem helpers", r"smem_addr\(const void"), ("fill_pending", r"void fill_pending"),
                       ("stage", r"void stage_instance"), ("wavefront", r"bool wavefront\("),
                       ("evaluate glue", r"int evaluate_instance"), ("link_pass", r"void link_pass"),
                       ("apply_swap", r"int apply_swap"), ("trace_path", r"int trace_path"),
                       ("n5_moves", r"int n5_moves"), ("rng", r"uint64_t mix64")],
    "l2s_kernels.cuh": [("setup kernels", r"__global__ void load_kernel"), ("div_rn", r"float div_rn"),
                        ("prologue+search", r"step_kernel\(const __grid_constant__"), ("swap glue", r"// 1\. N5 swap"),
                        ("eval glue / S,T rows", r"// 2\. evaluate"), ("x output", r"// 3\. node features"),
                        ("emc output", r"// 4\. machine edges"), ("path+moves glue", r"// 5\. critical path"),
                        ("scalars", r"// 6\. reward"), ("other kernels", r"// K fused steps")],
}
ranges = {}
for f, marks in MARKS.items():
    lines = open(os.path.join(ROOT, "l2s_b200", "csrc", f)).read().split("\n")
    starts = []
    for name, pat in marks:
        for i, ln in enumerate(lines):
            if re.search(pat, ln):
                starts.append((i + 1, name))
                break
    starts.sort()
    ranges[f] = [(s, (starts[k + 1][0] - 1 if k + 1 < len(starts) else 10 ** 9), nm) for k, (s, nm) in enumerate(starts)]

rows = list(csv.reader(open(sys.argv[1])))
cur, hdr, agg = None, None, {}
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split('/')[-1]
        continue
    if len(r) > 5 and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) > 10 and r[0].isdigit():
        ln = int(r[0])
        name = cur + ":other"
        for lo, hi, nm in ranges.get(cur, []):
            if lo <= ln <= hi:
                name = nm
        try:
            a = agg.setdefault(name, [0, 0])
            a[0] += int(r[hdr.index("Instructions Executed")])
            a[1] += int(r[hdr.index("# Samples")])
        except ValueError:
            pass
ti = sum(v[0] for v in agg.values()) or 1
ts = sum(v[1] for v in agg.values()) or 1
per = float(sys.argv[2]) / float(sys.argv[3]) if len(sys.argv) > 3 else 0
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("%-28s inst=%5.1f%% (%5.0f/instance)  stall samples=%5.1f%%" % (k, 100 * v[0] / ti, per * v[0] / ti, 100 * v[1] / ts))
