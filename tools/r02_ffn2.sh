#!/bin/bash
mkdir -p gpurun_out
for dbg in 0 $((11*256+7*4096)) $((11*256+7*4096+64)) $((9*256+6*4096)) $((3*256+2*4096)) $((1*256+1*4096)) $((11*256+4*4096)) $((6*256+7*4096)) $((64)); do
echo "dbg=$dbg G1_SPLIT=$(( (dbg>>8)&15 )) G2_SPLIT=$(( (dbg>>12)&15 )) (value-1; 0 = default 5/3) seq=$(( (dbg>>6)&1 ))"
COCR_FFN_DBG=$dbg timeout 60 python - <<'PY' 2>&1 | tail -3
import ctypes, os
from conformer_ocr_b200 import _lib
lib=_lib.load(); e=ctypes.c_double()
rc=lib.cocr_selftest(0, 21, 98304, 6, 4+8, ctypes.byref(e))
print('rc',rc,'err',e.value)
PY
done 2>&1 | tee gpurun_out/ffn_variants.log
