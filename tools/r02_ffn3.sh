#!/bin/bash
# which resource bounds k_ffn_fused: remove one at a time (results are garbage, only the time counts)
mkdir -p gpurun_out
for dbg in 0 1 4 8 16 5 17 21 9 12 13 29; do
echo "COCR_FFN_DBG=$dbg (1 no weight loads, 4 no final-epilogue memory traffic, 8 no MMAs, 16 no xn tile loads)"
COCR_FFN_DBG=$dbg timeout 60 python - <<'PY' 2>&1 | grep selftest_ffn
import ctypes, os
from conformer_ocr_b200 import _lib
lib=_lib.load(); e=ctypes.c_double()
rc=lib.cocr_selftest(0, 21, 98304, 6, 4+8, ctypes.byref(e))
PY
done 2>&1 | tee gpurun_out/ffn_knockout.log
