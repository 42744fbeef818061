#!/bin/bash
# usage: tools/r02_scale.sh N   -- bench.py on N GPUs exactly as the driver launches it, NCCL INFO kept
N=$1
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err
else
  NCCL_DEBUG=INFO timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err
fi
echo "N=$N exit $?"
python - <<PY
import json
d=json.loads(open('gpurun_out/scale_$N.json').read().strip().splitlines()[-1])
print('N=$N value',round(d['value']),'ms/step',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value']),'u8',round(d['e2e']['uint8_input']['value']),'lines',d['config']['lines_in_timed_region'])
PY
grep -c "NCCL INFO" gpurun_out/scale_$N.err; grep "nranks" gpurun_out/scale_$N.err | head -3
