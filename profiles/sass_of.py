"""Print the SASS (with per-instruction executed counts) attributed to a source line range of one file.
usage: sass_of.py export.csv file.cuh 'start regex' 'end regex'"""
import csv
import re
import sys
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = list(csv.reader(open(sys.argv[1])))
fname = sys.argv[2]
src = open(os.path.join(ROOT, "l2s_b200", "csrc", fname)).read().split("\n")
lo = [i + 1 for i, l in enumerate(src) if re.search(sys.argv[3], l)][0]
hi = [i + 1 for i, l in enumerate(src) if re.search(sys.argv[4], l) and i + 1 > lo][0]
cur = hdr = None
on = False
seen = set()
tot = 0
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split('/')[-1]
        continue
    if len(r) > 5 and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) > 10:
        if r[0].isdigit():
            on = cur == fname and lo <= int(r[0]) < hi
            if on:
                print("--- L%s: %s" % (r[0], r[1][:100]))
        elif on and r[2].startswith('0x') and r[2] not in seen:
            seen.add(r[2])
            e = int(r[hdr.index('Instructions Executed')])
            tot += e
            print("      %-64s exec=%d" % (r[3][:64], e))
print("total executed in range:", tot)
