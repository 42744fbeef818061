#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q -s -k "bf16" 2>&1 | grep -v "^$" | grep "bf16\|Error\|error\|passed\|failed" | cut -c1-300 | tail -30
