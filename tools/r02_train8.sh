#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_bf16.py -m gpu -x -q -k "train or bgemm" 2>&1 | tail -4
timeout 300 python bench.py --config 4 --steps 20 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/t8.json 2> gpurun_out/t8.err; echo "bench exit $?"; tail -2 gpurun_out/t8.err
python - <<'PY'
import json
for n in ('t8',):
    d=json.load(open(f'gpurun_out/{n}.json'))
    print(n,'lines/s',round(d['value'],1),'ms/step',round(d['ms_per_step'],3),'e2e',d['e2e'] and round(d['e2e']['value'],1),'launches/step',d['gpu_launches']//d['steps'],'loss',d['config']['final_loss'], 'rf', d['roofline'] and round(d['roofline']['frac'],4))
PY
