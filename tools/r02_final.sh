#!/bin/bash
# the round's bench lines on one GPU: default (configs[1]), --impl reference, configs[0] and configs[4]
mkdir -p gpurun_out
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err; echo "ref exit $?"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench exit $?"; tail -3 gpurun_out/final_bench.err
timeout 900 python bench.py --config 1 --steps 20 --warmup 5 > gpurun_out/final_cfg1.json 2> gpurun_out/final_cfg1.err; echo "cfg1 exit $?"; tail -3 gpurun_out/final_cfg1.err
timeout 900 python bench.py --config 5 --steps 10 --warmup 3 > gpurun_out/final_cfg5.json 2> gpurun_out/final_cfg5.err; echo "cfg5 exit $?"; tail -3 gpurun_out/final_cfg5.err
python - <<'PY'
import json
for n in ('final_bench','final_cfg1','final_cfg5','final_ref'):
    try:
        d=json.loads(open(f'gpurun_out/{n}.json').read().strip().splitlines()[-1])
        e=d.get('e2e') or {}
        print(n,'value',round(d['value'],1),'ms/step',round(d['ms_per_step'],2),'ms/pass',round(d.get('ms_per_pass',0),3),'e2e',round(e.get('value',0),1),'api',round((d.get('e2e_api') or {}).get('value',0),1),'frac',round((d.get('roofline_step') or {}).get('frac',0),4), 'eager',(d.get('gpu_eager_baseline') or {}))
    except Exception as ex: print(n,'ERR',ex)
PY
