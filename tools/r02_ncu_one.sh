#!/bin/bash
# usage: tools/r02_ncu_one.sh <kernel regex> <out name> [skip] [count]   -- one ncu --set full capture of a kernel of bench.py
mkdir -p gpurun_out
K=$1; OUT=$2; SKIP=${3:-5}; CNT=${4:-1}
CMD="python bench.py --steps 1 --warmup 1 --passes-per-step 2 --no-e2e --no-cpu-baseline --no-eager-baseline --profile-steps 0"
COCR_NO_GRAPH=1 timeout 300 $CMD > gpurun_out/${OUT}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${OUT}_plain.log; exit 1; }
COCR_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$K" -s $SKIP -c $CNT -f -o gpurun_out/$OUT $CMD > gpurun_out/${OUT}_ncu.log 2>&1; echo "ncu exit $?"
tail -3 gpurun_out/${OUT}_ncu.log; ls -la gpurun_out/$OUT.ncu-rep
