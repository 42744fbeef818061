#!/bin/bash
# FFN kernel check: self tests, timing of the stand-alone kernel, short bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bf16.py -m gpu -x -q -k "ffn or golden or long_run" > gpurun_out/ffn_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/ffn_pytest.log
timeout 120 python - <<'PY' 2>&1 | tee gpurun_out/ffn_time.log
import ctypes
from conformer_ocr_b200 import _lib
lib=_lib.load(); e=ctypes.c_double()
for i in range(2):
    print('rc', lib.cocr_selftest(0, 21, 98304, 6, 4+8, ctypes.byref(e)), 'err', e.value)
PY
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/ffn_bench.json 2> gpurun_out/ffn_bench.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/ffn_bench.json'))
print('value',d['value'],'ms/pass',d['ms_per_pass'],'e2e',d['e2e']['value'])
for k,v in sorted(d['kernel_families'].items(), key=lambda kv:-kv[1]['ms_per_pass']): print(k, round(v['avg_launch_ms']*1e3,1))
PY
