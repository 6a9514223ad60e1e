#!/usr/bin/env python
"""Pretty-print (parts of) a bench.py JSON line:  python profiles/show_bench.py file.json [key ...]"""
import json
import sys


def show(o, ind=0):
    if isinstance(o, dict):
        for k, v in o.items():
            if isinstance(v, dict):
                print(" " * ind + str(k) + ":")
                show(v, ind + 2)
            else:
                print(" " * ind + "%s: %s" % (k, str(v)[:160]))
    else:
        print(" " * ind + str(o)[:300])


d = json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
if len(sys.argv) > 2:
    for k in sys.argv[2:]:
        print(k + ":")
        show(d.get(k), 2)
else:
    show(d)
