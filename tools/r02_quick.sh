#!/bin/bash
# quick check of a kernel change: bf16 parity tests + short bench with per-family times
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_configs.py -m gpu -x -q -k "not selftest" > gpurun_out/quick_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/quick_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/quick_bench.json 2> gpurun_out/quick_bench.err; echo "bench exit $?"; tail -3 gpurun_out/quick_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/quick_bench.json'))
print('value',round(d['value']),'ms/pass',round(d['ms_per_pass'],3),'e2e',round(d['e2e']['value']),'api',round(d['e2e_api']['value']),'step frac',round(d['roofline_step']['frac'],4))
for k,v in sorted(d['kernel_families'].items(), key=lambda kv:-kv[1]['ms_per_pass']): print(f"  {k:22s} {v['avg_launch_ms']*1e3:7.1f} us x{v['launches']}")
PY
