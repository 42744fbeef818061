#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ctc.py tests/test_gpu_train.py -m gpu -x -q 2>&1 | tail -4
timeout 300 python bench.py --config 4 --steps 20 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/t9.json 2> gpurun_out/t9.err; echo "bench exit $?"; tail -2 gpurun_out/t9.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/t9.json'))
print('lines/s',round(d['value'],1),'ms/step',round(d['ms_per_step'],3),'e2e',d['e2e'] and round(d['e2e']['value'],1),'loss',d['config']['final_loss'])
f=d['kernel_families']
for k in ('ctc_alpha','ctc_beta_grad','row_logsumexp','ctc_grad_softmax'):
    if k in f: print('  ',k, round(f[k]['avg_launch_ms']*1e3,1),'us')
PY
