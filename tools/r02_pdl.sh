#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_configs.py -m gpu -x -q > gpurun_out/pdl_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pdl_pytest.log
for v in 1 0 1 0; do
if [ $v = 1 ]; then export COCR_NO_PDL=1; else unset COCR_NO_PDL; fi
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-e2e --profile-steps 0 > gpurun_out/pdl_$v.json 2> gpurun_out/pdl_$v.err; echo "NO_PDL=$v exit $?"
python - <<PY
import json
d=json.load(open('gpurun_out/pdl_$v.json'))
print('   value',round(d['value']),'ms/pass',round(d['ms_per_pass'],4))
PY
done
