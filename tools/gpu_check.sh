#!/bin/bash
# usage (on the GPU box, from the repo root): tools/gpu_check.sh <tag> [pytest -k expression] [extra bench args]
# runs the bf16 GPU tests and a short bf16 bench, leaving logs under gpurun_out/
tag=$1; kexpr=${2:-"not selftest"}; shift; shift
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_bf16.py -x -q -s -k "$kexpr" > gpurun_out/${tag}_pytest.log 2>&1; echo "exit $?" >> gpurun_out/${tag}_pytest.log
tail -14 gpurun_out/${tag}_pytest.log
timeout 300 python bench.py --precision bf16 --steps 5 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python tools/summarize_bench.py gpurun_out/${tag}_bench.json; tail -5 gpurun_out/${tag}_bench.err
