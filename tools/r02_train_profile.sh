#!/bin/bash
# ncu evidence of the training step (bench.py --config 4): launch list of one short run, then --set full of its main kernels
mkdir -p gpurun_out
CMD="python bench.py --config 4 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-eager-baseline --profile-steps 0"
timeout 300 $CMD > gpurun_out/r02_train_plain.log 2>&1 || { echo "plain failed"; tail -5 gpurun_out/r02_train_plain.log; exit 1; }
COCR_NO_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1700 -c 1700 --csv --log-file gpurun_out/r02_train_launches.csv $CMD > gpurun_out/r02_train_ncu_list.log 2>&1; echo "ncu list exit $?"
COCR_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_bgemm_tc|k_att_softmax_tc|k_att_ds_tc|k_ln_bwd_fused|k_dwconv1d_wgrad|k_resid_drop|k_silu_drop" -s 700 -c 60 -f -o gpurun_out/r02_train_prof $CMD > gpurun_out/r02_train_ncu_full.log 2>&1; echo "ncu full exit $?"
ls -la gpurun_out/r02_train_prof.ncu-rep gpurun_out/r02_train_launches.csv
