#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/train_check.py 2>&1 | tee gpurun_out/train_check.log | tail -80
