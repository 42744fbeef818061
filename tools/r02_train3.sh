#!/bin/bash
mkdir -p gpurun_out
COCR_TRAIN_PRECISION=bf16 timeout 600 python tools/train_check.py 2>&1 | tee gpurun_out/train_check_bf16.log | grep -v "^   tap" | tail -40
grep "<<<" gpurun_out/train_check_bf16.log | head -20
timeout 600 python bench.py --config 4 --precision bf16 --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/train_bench_bf16.json 2> gpurun_out/train_bench_bf16.err; echo "bench exit $?"; tail -5 gpurun_out/train_bench_bf16.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/train_bench_bf16.json'))
print('lines/s',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e'] and d['e2e']['value'],'launches',d['gpu_launches'], 'loss', d['config']['final_loss'])
for k,v in sorted(d['kernel_families'].items(), key=lambda kv:-kv[1]['ms_per_step'])[:40]:
    print(f"  {k:28s} n={v['launches']:4d} ms/step={v['ms_per_step']:8.3f} avg_us={v['avg_launch_ms']*1e3:8.1f} {v['achieved']:8.1f} {v['unit']} frac={v['frac']:.3f}")
PY
