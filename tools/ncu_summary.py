#!/usr/bin/env python
"""tools/ncu_summary.py raw.csv… — one line per profiled launch from `ncu --page raw --csv` exports"""
import csv, sys
COLS = [('gpu__time_duration.sum', 't'), ('dram__bytes_read.sum', 'rd'), ('dram__bytes_write.sum', 'wr'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram%'), ('l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1%'),
        ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l2%'), ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue%'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ%'), ('launch__registers_per_thread', 'regs'),
        ('smsp__inst_executed.sum', 'inst')]
for f in sys.argv[1:]:
    rows = list(csv.reader(open(f)))
    if len(rows) < 3:
        continue
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    print('==', f)
    for r in data:
        name = r[idx['Kernel Name']]
        name = name.replace('void ', '').replace('hptb::', '')[:58]
        out = [f"{name:58s}", f"grid {r[idx['Grid Size']]:>12s}"]
        for k, lab in COLS:
            if k in idx:
                v = r[idx[k]].replace(',', '')
                try:
                    v = f"{float(v):.1f}"
                except ValueError:
                    pass
                u = units[idx[k]]
                u = {'usecond': 'us', 'us': 'us', 'Mbyte': 'MB', 'Gbyte': 'GB', 'Kbyte': 'KB', '%': '', 'register/thread': '', 'inst': ''}.get(u, u)
                out.append(f"{lab}={v}{u}")
        print(' '.join(out))
