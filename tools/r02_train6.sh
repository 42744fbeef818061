#!/bin/bash
mkdir -p gpurun_out
timeout 120 python - <<'PY' 2>&1 | tail -8
import ctypes
from conformer_ocr_b200 import _lib
lib=_lib.load(); e=ctypes.c_double()
for M,N,K,fl in [(384,767,36,4),(300,144,576,0),(300,576,144,2),(26,26,16,4),(250,200,100,7),(1000,288,144,0),(129,48,64,0),(77,16,40,1)]:
    rc=lib.cocr_selftest(0, 30, M, N, K+100000*fl, ctypes.byref(e)); print(f'M={M} N={N} K={K} flags={fl} rc={rc} err={e.value:.3e}', flush=True)
PY
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --config 4 --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/train_bench_bf16.json 2> gpurun_out/train_bench_bf16.err; echo "bench exit $?"; tail -3 gpurun_out/train_bench_bf16.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/train_bench_bf16.json'))
print('lines/s',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e'] and d['e2e']['value'],'launches',d['gpu_launches'], 'loss', d['config']['final_loss'])
tot=sum(v['ms_per_step'] for v in d['kernel_families'].values()); print('sum of kernel ms', tot)
for k,v in sorted(d['kernel_families'].items(), key=lambda kv:-kv[1]['ms_per_step'])[:22]:
    print(f"  {k:28s} n={v['launches']:4d} ms/step={v['ms_per_step']:8.3f} avg_us={v['avg_launch_ms']*1e3:8.1f} {v['achieved']:8.1f} {v['unit']} frac={v['frac']:.3f}")
PY
