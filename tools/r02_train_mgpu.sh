#!/bin/bash
# usage (GPU box with N GPUs): tools/r02_train_mgpu.sh N
N=$1
mkdir -p gpurun_out
if [ "$2" != "nocheck" ]; then
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/train_allreduce_check.py 2> gpurun_out/train_allreduce_$N.err | tee gpurun_out/train_allreduce_check_${N}gpu.json
echo "check exit $?"; tail -3 gpurun_out/train_allreduce_$N.err
fi
NCCL_DEBUG=INFO timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --config 4 --gpus $N --steps 20 --warmup 3 > gpurun_out/train_bench_${N}gpu.json 2> gpurun_out/train_bench_${N}gpu.err
echo "bench exit $?"; grep -c "NCCL INFO" gpurun_out/train_bench_${N}gpu.err; grep "nranks\|Init COMPLETE" gpurun_out/train_bench_${N}gpu.err | head -4
python - $N <<'PY'
import json,sys
d=json.load(open(f'gpurun_out/train_bench_{sys.argv[1]}gpu.json'))
print('N',d['n_gpus'],'lines/s',round(d['value'],1),'ms/step',round(d['ms_per_step'],2),'e2e',d['e2e'] and round(d['e2e']['value'],1))
PY
