#!/usr/bin/env python
"""print gpurun_out/sweep.jsonl (or argv[1]) as a table"""
import collections, json, sys
rows = [json.loads(l) for l in open(sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/sweep.jsonl')]
by = collections.OrderedDict()
for r in rows:
    by.setdefault(r['case'], []).append(r)
for case, rs in by.items():
    print(case)
    for r in rs:
        print(f"   {r['impl']:6s} {str(r['knobs']):24s} {r['us']:9.2f} us {r['gbs']:8.1f} GB/s {r['frac']:.3f}")
