#!/bin/bash
mkdir -p gpurun_out
for bn in 96 144; do
COCR_PW1_BN=$bn timeout 300 python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-eager-baseline --no-e2e > gpurun_out/pw1_$bn.json 2> gpurun_out/pw1_$bn.err; echo "bn=$bn exit $?"; tail -2 gpurun_out/pw1_$bn.err
python - <<PY
import json
d=json.load(open('gpurun_out/pw1_$bn.json'))
print('  ms/pass',round(d['ms_per_pass'],3),'pw1',round(d['kernel_families']['gemm_pw1_glu']['avg_launch_ms']*1e3,1))
PY
done
timeout 600 python -m pytest tests/test_gpu_bf16.py -m gpu -x -q -k "golden or odd_shapes or full_width" 2>&1 | tail -2
