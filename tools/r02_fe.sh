#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_bf16.py -m gpu -x -q -k "golden or uint8 or odd_shapes" > gpurun_out/fe_pytest.log 2>&1; echo "pytest exit $?"; tail -12 gpurun_out/fe_pytest.log
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/fe_bench.json 2> gpurun_out/fe_bench.err; echo "bench exit $?"; tail -3 gpurun_out/fe_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/fe_bench.json'))
print('value',round(d['value']),'ms/pass',round(d['ms_per_pass'],3),'e2e',round(d['e2e']['value']),'u8',round(d['e2e']['uint8_input']['value']))
for k,v in sorted(d['kernel_families'].items(), key=lambda kv:-kv[1]['ms_per_pass']): print(f"  {k:22s} {v['avg_launch_ms']*1e3:7.1f} us x{v['launches']}")
PY
