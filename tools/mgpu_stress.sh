#!/bin/bash
# usage (2-GPU box): tools/mgpu_stress.sh <runs> <steps>  -- failure counts of the 2-GPU bench under a few switches
runs=${1:-4}; steps=${2:-200}
mkdir -p gpurun_out
run2() {  # env-string
  fails=0
  for i in $(seq 1 $runs); do
    env $1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps $steps --warmup 3 --no-cpu-baseline --no-e2e --profile-steps 0 > gpurun_out/ms.json 2> gpurun_out/ms.err || fails=$((fails+1))
  done
  echo "[$1] $fails failures of $runs"
}
run2 "X=1"
run2 "COCR_FFN_SPLIT=1"
run2 "COCR_ATTN_STREAM_KV=1"
# two independent single-GPU processes at the same time (no NCCL)
fails=0
for i in $(seq 1 $runs); do
  CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps $steps --warmup 3 --no-cpu-baseline --no-e2e --profile-steps 0 > gpurun_out/ms0.json 2> gpurun_out/ms0.err &
  p0=$!
  CUDA_VISIBLE_DEVICES=1 timeout 300 python bench.py --steps $steps --warmup 3 --no-cpu-baseline --no-e2e --profile-steps 0 > gpurun_out/ms1.json 2> gpurun_out/ms1.err &
  p1=$!
  wait $p0 || fails=$((fails+1)); wait $p1 || fails=$((fails+1))
done
echo "[two independent processes] $fails failures of $((2*runs))"
