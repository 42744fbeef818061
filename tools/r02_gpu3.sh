#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ctc.py tests/test_gpu_configs.py -m gpu -x -q -s > gpurun_out/gpu3_pytest.log 2>&1; echo "pytest exit $?"; grep -v "^$" gpurun_out/gpu3_pytest.log | tail -25
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/train_step_partial.py > gpurun_out/train_partial_2gpu.json 2> gpurun_out/train_partial_2gpu.err; echo "train2 exit $?"; tail -2 gpurun_out/train_partial_2gpu.json
timeout 300 python tools/train_step_partial.py > gpurun_out/train_partial_1gpu.json 2> gpurun_out/train_partial_1gpu.err; echo "train1 exit $?"; tail -1 gpurun_out/train_partial_1gpu.json
