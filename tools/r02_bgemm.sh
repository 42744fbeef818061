#!/bin/bash
mkdir -p gpurun_out
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/bgemm_selftest.log
import ctypes
from conformer_ocr_b200 import _lib
lib=_lib.load(); e=ctypes.c_double()
cases=[(128,64,64,0),(128,64,64,1),(128,64,64,2),(128,64,64,3),
       (384,384,36,4),(384,767,36,4),(384,36,384,4|2),(384,36,384,4|3),(767,36,384,4|3),
       (300,144,576,0),(300,576,144,2),(576,144,1000,3|8),(144,576,1000,3|8),(432,144,5000,3|8),
       (26,26,16,4),(26,16,26,4|2),(51,16,26,4|3),(10,10,16,4),(250,200,100,7),(1000,288,144,0)]
for M,N,K,fl in cases:
    rc=lib.cocr_selftest(0, 30, M, N, K+100000*fl, ctypes.byref(e))
    msg = '' if rc==0 else lib.cocr_last_error(None).decode()
    print(f'M={M} N={N} K={K} flags={fl:2d} rc={rc} err={e.value:.3e} {msg}', flush=True)
PY
