#!/bin/bash
# ncu evidence of the training step (bench.py --config 4): launch list of one step, then --set full of its main kernels.
# The .ncu-rep stays on the box (too large for gpurun_out); its raw page comes back as CSV.
mkdir -p gpurun_out
CMD="python bench.py --config 4 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-eager-baseline --profile-steps 0"
timeout 300 $CMD > gpurun_out/r02_train_plain.log 2>&1 || { echo "plain failed"; tail -5 gpurun_out/r02_train_plain.log; exit 1; }
COCR_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1700 -c 1700 --csv --log-file gpurun_out/r02_train_launches.csv $CMD > gpurun_out/r02_train_ncu_list.log 2>&1; echo "ncu list exit $?"
COCR_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none -k regex:"k_bgemm_tc|k_att_softmax_tc|k_att_ds_tc|k_ln_bwd_fused|k_dwconv1d_wgrad" -s 760 -c 24 -f -o /tmp/r02_train_prof $CMD > gpurun_out/r02_train_ncu_full.log 2>&1; echo "ncu full exit $?"
ncu -i /tmp/r02_train_prof.ncu-rep --page raw --csv > gpurun_out/r02_train_ncu_full_raw.csv 2>/dev/null
ls -la /tmp/r02_train_prof.ncu-rep gpurun_out/
