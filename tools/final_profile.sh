#!/bin/bash
# usage (GPU box): tools/final_profile.sh  -- the round's bench line, then the two ncu passes of the profiling recipe
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 600 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench exit $?"
tail -c 600 gpurun_out/final_bench.json
$CMD > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 720 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1; echo "ncu list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_attn_tc|k_gemm_tc|k_ffn_fused|k_dw_fast|k_frontend_convs2" -s 136 -c 12 -o gpurun_out/prof $CMD > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"
ls -la gpurun_out/prof.ncu-rep gpurun_out/launches.csv
