"""Share of executed warp instructions / stall samples per phase of l2s::step_kernel
(input: `ncu --page source --csv --print-source cuda,sass` export).  usage: by_phase.py file.csv insts_per_launch B"""
import csv
import sys

PH = {"l2s_device.cuh": [(75, 125, "smem helpers (ld/st)"), (126, 132, "fill_pending"), (133, 150, "stage"), (151, 193, "wavefront"),
                          (194, 218, "evaluate glue"), (219, 241, "link_pass"), (242, 262, "apply_swap"), (263, 286, "trace_path"),
                          (287, 322, "n5_moves")],
      "l2s_kernels.cuh": [(262, 270, "div_rn (x)"), (303, 331, "prologue+search"), (332, 352, "swap glue"), (353, 380, "eval glue/S,T rows"),
                          (381, 400, "x output"), (401, 420, "emc output"), (421, 438, "path+moves glue"), (439, 475, "scalars")]}
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
        for lo, hi, nm in PH.get(cur, []):
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
