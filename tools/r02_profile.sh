#!/bin/bash
# round-2 profile set: bench line, ncu launch list (same command, short), ncu --set full of the top kernels
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --passes-per-step 2 --no-e2e --no-cpu-baseline --no-eager-baseline --profile-steps 0"
timeout 300 $CMD > gpurun_out/r02_plain.log 2>&1 || { echo "plain failed"; tail -5 gpurun_out/r02_plain.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/r02_ncu_list.log 2>&1; echo "ncu list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_attn_tc|k_gemm_tc|k_ffn_fused|k_dw_v2|k_frontend_tc" -s 131 -c 14 -f -o gpurun_out/r02_prof $CMD > gpurun_out/r02_ncu_full.log 2>&1; echo "ncu full exit $?"
ls -la gpurun_out/r02_prof.ncu-rep gpurun_out/r02_launches.csv
