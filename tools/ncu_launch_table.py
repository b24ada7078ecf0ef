#!/usr/bin/env python
"""tools/ncu_launch_table.py launches.csv [--second-half] — one line per launch from an `ncu --metrics … --csv --log-file` capture"""
import collections
import csv
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if not l.startswith("==")))
h = rows[0]
k, m, v, i, g = (h.index(x) for x in ("Kernel Name", "Metric Name", "Metric Value", "ID", "Grid Size"))
d = collections.OrderedDict()
for r in rows[1:]:
    if len(r) > v:
        d.setdefault(r[i], {"name": r[k][:120], "grid": r[g]})[r[m]] = r[v]
items = list(d.values())
if "--second-half" in sys.argv:
    items = items[len(items) // 2:]
for e in items:
    print(e["name"], e["grid"], " ".join(f"{kk.split('__')[-1]}={vv}" for kk, vv in e.items() if kk not in ("name", "grid")))
