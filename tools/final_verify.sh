#!/bin/bash
# usage (GPU box): tools/final_verify.sh  -- the whole GPU test suite, smoke(), then tools/final_profile.sh
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/final_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/final_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/final_smoke.log
tools/final_profile.sh
