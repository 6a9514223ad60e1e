"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export by CUDA source line:
share of executed warp instructions and of stall samples.  usage: by_line.py file.csv [top]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
agg = collections.OrderedDict()
cur_file, hdr = None, None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split('/')[-1]
        continue
    if len(r) > 5 and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) > 10 and r[0] != "" and r[0].isdigit():
        i_inst, i_samp = hdr.index("Instructions Executed"), hdr.index("# Samples")
        try:
            a = agg.setdefault((cur_file, int(r[0]), r[1][:100]), [0, 0])
            a[0] += int(r[i_inst])
            a[1] += int(r[i_samp])
        except ValueError:
            pass
tot = sum(v[0] for v in agg.values()) or 1
tots = sum(v[1] for v in agg.values()) or 1
print("total warp instructions (all captured launches)", tot, "stall samples", tots)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-16s %4d inst=%5.1f%% samp=%5.1f%%  %s" % (k[0][:16], k[1], 100 * v[0] / tot, 100 * v[1] / tots, k[2]))
