#!/bin/bash
# A/B of two environment settings on the same box: usage tools/r02_ab.sh "ENV_A=1" "ENV_B=1" (empty string = default)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_configs.py -m gpu -x -q -k "not selftest and not soak and not long_run" > gpurun_out/ab_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/ab_pytest.log
for rep in 1 2; do
for v in "$1" "$2"; do
env $v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-e2e > gpurun_out/ab.json 2> gpurun_out/ab.err; echo "[$v] exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/ab.json'))
f=d['kernel_families']
print('   ms/pass',round(d['ms_per_pass'],4),' '.join(f"{k}={v['avg_launch_ms']*1e3:.1f}" for k,v in sorted(f.items(), key=lambda kv:-kv[1]['ms_per_pass'])[:4]))
PY
done; done
