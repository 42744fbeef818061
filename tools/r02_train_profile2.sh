#!/bin/bash
# ncu launch list of one training step of the final tree (bench.py --config 4, graph off so that every kernel is listed)
mkdir -p gpurun_out
CMD="python bench.py --config 4 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-eager-baseline --profile-steps 0"
timeout 300 $CMD > gpurun_out/r02_train_plain.log 2>&1 || { echo "plain failed"; tail -5 gpurun_out/r02_train_plain.log; exit 1; }
COCR_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 1450 --csv --log-file gpurun_out/r02_train_launches.csv $CMD > gpurun_out/r02_train_ncu_list.log 2>&1; echo "ncu list exit $?"
ls -la gpurun_out/
