#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_bf16.py -m gpu -x -q -k "train or bgemm" 2>&1 | tail -4
for v in "" "COCR_TRAIN_NO_PDL=1"; do
env $v timeout 300 python bench.py --config 4 --steps 20 --warmup 3 --no-cpu-baseline --no-eager-baseline --profile-steps 0 > gpurun_out/t7.json 2> gpurun_out/t7.err; echo "[$v] bench exit $?"; tail -2 gpurun_out/t7.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/t7.json'))
print('   lines/s',round(d['value'],1),'ms/step',round(d['ms_per_step'],3),'e2e',d['e2e'] and round(d['e2e']['value'],1),'loss',d['config']['final_loss'])
PY
done
