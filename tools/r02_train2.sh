#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -x -q -s 2>&1 | tail -25 | tee gpurun_out/train_pytest.log
timeout 600 python bench.py --config 4 --steps 5 --warmup 3 > gpurun_out/train_bench.json 2> gpurun_out/train_bench.err; echo "bench exit $?"; tail -5 gpurun_out/train_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/train_bench.json'))
print('lines/s',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e'] and d['e2e']['value'],'launches',d['gpu_launches'])
print('roofline_step',d['roofline_step']['frac'],'eager',d['gpu_eager_baseline'],'cpu',d['cpu_baseline'] and d['cpu_baseline']['value'])
for k,v in sorted(d['kernel_families'].items(), key=lambda kv:-kv[1]['ms_per_step'])[:25]:
    print(f"  {k:28s} n={v['launches']:4d} ms/step={v['ms_per_step']:8.3f} avg_us={v['avg_launch_ms']*1e3:8.1f} {v['achieved']:8.1f} {v['unit']} frac={v['frac']:.3f}")
PY
