"""Fused FFN kernel (cocr_selftest 21): parity against the host reference on a small case and time at the benchmark size."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from conformer_ocr_b200 import _lib
lib = _lib.load()
for M, NC, flags in ((256, 6, 6), (1000, 6, 4), (98304, 6, 4 + 8)) if len(sys.argv) < 2 else ((int(sys.argv[1]), 6, 4 + 8),):
    err = ctypes.c_double(-1)
    rc = lib.cocr_selftest(0, 21, M, NC, flags, ctypes.byref(err))
    print(f'M={M} NC={NC} flags={flags}: rc={rc} max err {err.value:.3e}', flush=True)
