#!/bin/bash
mkdir -p gpurun_out
for mb in 0 48 80 96; do
echo "== COCR_L2_PERSIST=$mb"
COCR_L2_PERSIST=$mb timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-e2e > gpurun_out/l2_$mb.json 2> gpurun_out/l2_$mb.err; echo "exit $?"; grep cocr: gpurun_out/l2_$mb.err | head -2
python - <<PY
import json
d=json.load(open('gpurun_out/l2_$mb.json'))
print('value',round(d['value']),'ms/pass',round(d['ms_per_pass'],3))
print(' '.join(f"{k}={v['avg_launch_ms']*1e3:.1f}" for k,v in sorted(d['kernel_families'].items(), key=lambda kv:-kv[1]['ms_per_pass'])[:9]))
PY
done
