#!/bin/bash
# CTC tests on 1 GPU of the box + 2-GPU bench (gather inside the timed loops, NCCL INFO kept)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ctc.py -m gpu -x -q -s > gpurun_out/ctc_pytest.log 2>&1; echo "ctc pytest exit $?"; tail -15 gpurun_out/ctc_pytest.log
NCCL_DEBUG=INFO timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "2gpu exit $?"
tail -c 600 gpurun_out/bench_2gpu.json; grep -c "NCCL INFO" gpurun_out/bench_2gpu.err; grep -m3 "nranks\|Init COMPLETE\|NVLS" gpurun_out/bench_2gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_2gpu.json').read().strip().splitlines()[-1])
print('N=2 value',round(d['value']),'ms/step',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value']), 'e2e u8', round(d['e2e']['uint8_input']['value']))
PY
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2>&1; tail -c 400 gpurun_out/bench_ref.json
