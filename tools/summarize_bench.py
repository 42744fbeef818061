#!/usr/bin/env python
"""Print the per-kernel-family breakdown of a bench.py JSON line (developer tool)."""
import json
import sys

d = json.load(open(sys.argv[1]))
print('lines/s %.1f  ms/step %.3f  e2e %s  launches %s' % (d['value'], d['ms_per_step'], d['e2e']['value'] if d.get('e2e') else None, d.get('gpu_launches')))
fam = d.get('kernel_families') or {}
tot = sum(v['ms_per_step'] for v in fam.values())
for k, v in sorted(fam.items(), key=lambda kv: -kv[1]['ms_per_step']):
    ms = max(v['ms_per_step'], 1e-9)
    print('%-28s n=%4d %8.3f ms (%4.1f%%) %8.1f us/launch %9.1f TFLOP/s %9.1f GB/s' % (
        k, v['launches'], v['ms_per_step'], 100 * v['ms_per_step'] / max(tot, 1e-9), 1e3 * ms / max(v['launches'], 1),
        v['flops_per_step'] / ms / 1e9, v['bytes_per_step'] / ms / 1e6))
print('sum of families %.3f ms' % tot)
if d.get('roofline'):
    print('roofline', json.dumps(d['roofline']))
