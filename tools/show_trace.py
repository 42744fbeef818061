#!/usr/bin/env python
"""Merge a (role, tag, clock) kernel trace (COCR_GEMM_TRACE / COCR_ATTN_TRACE) into one timeline.
usage: tools/show_trace.py file [start_event_index] [count]"""
import sys
rows = [tuple(int(x) for x in l.split()) for l in open(sys.argv[1])]
rows.sort(key=lambda e: e[2])
t0 = rows[0][2]
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
n = int(sys.argv[3]) if len(sys.argv) > 3 else 80
names = {0: 'EPI', 1: 'MMA', 2: 'TMA'}
print('events', len(rows), 'span cycles', rows[-1][2] - t0)
for r, t, c in rows[lo:lo + n]:
    print('%9d %-4s %d' % (c - t0, names.get(r, str(r)), t))
