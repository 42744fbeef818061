#!/bin/bash
# final measurement set of round 2 (one B200): GPU tests, smoke, bench lines of every config and both arms
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/final2_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/final2_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/final2_bench.json 2> gpurun_out/final2_bench.err; echo "bench exit $?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final2_ref.json 2> gpurun_out/final2_ref.err; echo "ref exit $?"
timeout 300 python bench.py --config 1 --steps 20 --warmup 3 > gpurun_out/final2_cfg1.json 2> gpurun_out/final2_cfg1.err; echo "cfg1 exit $?"
timeout 300 python bench.py --config 5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/final2_cfg5.json 2> gpurun_out/final2_cfg5.err; echo "cfg5 exit $?"
timeout 600 python bench.py --config 4 --steps 20 --warmup 3 > gpurun_out/final2_cfg4.json 2> gpurun_out/final2_cfg4.err; echo "cfg4 exit $?"
timeout 300 python bench.py --config 4 --impl reference --steps 2 --warmup 1 > gpurun_out/final2_cfg4_ref.json 2> gpurun_out/final2_cfg4_ref.err; echo "cfg4 ref exit $?"
timeout 300 python bench.py --config 4 --precision fp32 --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/final2_cfg4_fp32.json 2> gpurun_out/final2_cfg4_fp32.err; echo "cfg4 fp32 exit $?"
python - <<'PY'
import json
for n in ('bench','ref','cfg1','cfg5','cfg4','cfg4_ref','cfg4_fp32'):
    try:
        d=json.load(open(f'gpurun_out/final2_{n}.json'))
        print(n, d.get('impl','ours'), round(d['value'],1), d['unit'], 'ms/step', round(d['ms_per_step'],3), 'e2e', d.get('e2e') and round(d['e2e']['value'],1), 'rf_step', d.get('roofline_step') and round(d['roofline_step']['frac'],4), 'eager', d.get('gpu_eager_baseline') and {k:round(v['value'],1) for k,v in d['gpu_eager_baseline'].items() if isinstance(v,dict)}, 'cpu', d.get('cpu_baseline') and round(d['cpu_baseline']['value'],2))
    except Exception as e:
        print(n, 'ERR', e)
PY
