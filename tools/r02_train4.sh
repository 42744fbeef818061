#!/bin/bash
mkdir -p gpurun_out
for prec in fp32 bf16; do
COCR_TRAIN_PRECISION=$prec timeout 600 python tools/train_check.py 2>&1 | tee gpurun_out/train_check_$prec.log | grep -v "^   tap\|^   grad" | tail -20
done
grep "grad " gpurun_out/train_check_fp32.log | head
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -x -q 2>&1 | tail -5
for prec in bf16 fp32; do
timeout 600 python bench.py --config 4 --precision $prec --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/train_bench_$prec.json 2> gpurun_out/train_bench_$prec.err; echo "bench $prec exit $?"; tail -3 gpurun_out/train_bench_$prec.err
python - $prec <<'PY'
import json,sys
d=json.load(open(f'gpurun_out/train_bench_{sys.argv[1]}.json'))
print('lines/s',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e'] and d['e2e']['value'],'launches',d['gpu_launches'], 'loss', d['config']['final_loss'])
tot=sum(v['ms_per_step'] for v in d['kernel_families'].values()); print('sum of kernel ms', tot)
for k,v in sorted(d['kernel_families'].items(), key=lambda kv:-kv[1]['ms_per_step'])[:18]:
    print(f"  {k:28s} n={v['launches']:4d} ms/step={v['ms_per_step']:8.3f} avg_us={v['avg_launch_ms']*1e3:8.1f} {v['achieved']:8.1f} {v['unit']} frac={v['frac']:.3f}")
PY
done
