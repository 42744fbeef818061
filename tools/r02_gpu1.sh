#!/bin/bash
# round-2 first GPU pass: tests, bench line, FFN trace, sanitizer on smoke
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
tail -c 1500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
COCR_FFN_TRACE=gpurun_out/ffn_trace.txt timeout 300 python bench.py --steps 1 --warmup 1 --passes-per-step 2 --no-e2e --no-cpu-baseline --no-eager-baseline --profile-steps 0 > gpurun_out/trace_run.log 2>&1; echo "trace exit $?"
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_memcheck_smoke.log 2>&1; echo "memcheck exit $?"
tail -5 gpurun_out/sanitizer_memcheck_smoke.log
