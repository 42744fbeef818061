#!/bin/bash
# usage (GPU box): tools/stress_versions.sh <runs> <steps> <dir>...  -- failure count of a long single-GPU bench per checkout
runs=$1; steps=$2; shift; shift
for d in "$@"; do
  fails=0
  for i in $(seq 1 $runs); do
    (cd $d && timeout 300 python bench.py --steps $steps --warmup 3 --no-cpu-baseline --no-e2e --profile-steps 0 > /tmp/sv.json 2> /tmp/sv.err) || fails=$((fails+1))
  done
  echo "[$d] $fails failures of $runs"; grep -m1 "Error" /tmp/sv.err
done
