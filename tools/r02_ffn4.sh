#!/bin/bash
mkdir -p gpurun_out
timeout 120 python - <<'PY' 2>&1 | tee gpurun_out/ffn2_selftest.log
import ctypes
from conformer_ocr_b200 import _lib
lib=_lib.load(); e=ctypes.c_double()
for M,NC,fl in ((128,6,4),(300,6,6),(777,2,2),(2500,1,6),(4096,6,4),(20481,8,4),(98304,6,12)):
    rc=lib.cocr_selftest(0, 21, M, NC, fl, ctypes.byref(e)); print('M',M,'NC',NC,'flags',fl,'rc',rc,'err',e.value, flush=True)
PY
echo "v1:"; COCR_FFN_V1=1 timeout 120 python - <<'PY' 2>&1 | tail -2
import ctypes
from conformer_ocr_b200 import _lib
lib=_lib.load(); e=ctypes.c_double()
rc=lib.cocr_selftest(0, 21, 98304, 6, 12, ctypes.byref(e)); print('rc',rc,'err',e.value)
PY
