#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: stall-reason totals and the instructions holding most samples.
usage: ncu -i rep.ncu-rep --page source --csv [--launch-skip N --launch-count 1] > src.csv; tools/ncu_source_top.py src.csv [n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
print(rows[0][:2])
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr) and r[ix['# Samples']].isdigit()]
tot = sum(int(r[ix['# Samples']]) for r in body)
print('total samples', tot, 'instructions', len(body))
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {s: sum(int(r[ix[s]]) for r in body) for s in stalls}
print(sorted(agg.items(), key=lambda kv: -kv[1])[:8])
for i, r in enumerate(body):
    r.append(i)
top = sorted(body, key=lambda r: -int(r[ix['# Samples']]))[:n]
for r in top:
    ss = sorted(((s, int(r[ix[s]])) for s in stalls if int(r[ix[s]]) > 0), key=lambda kv: -kv[1])
    print('%5d %6s  #%-5d %-72s %s' % (int(r[ix['# Samples']]), r[ix['Instructions Executed']], r[-1], r[ix['Source']].strip()[:72], ss[:2]))
