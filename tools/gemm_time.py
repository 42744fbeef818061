"""time k_gemm_tc shapes through cocr_selftest (COCR_SELFTEST_TIME=1; COCR_GEMM_DBG=256 = main loop only)"""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from conformer_ocr_b200 import _lib
lib = _lib.load()
M = 98304
for (N, K, bn) in [(96, 144, 0), (288, 144, 0), (576, 144, 0), (1280, 144, 0), (192, 144, 0), (256, 144, 0), (96, 128, 0), (288, 128, 0), (288, 192, 0), (96, 64, 0), (288, 64, 0)]:
    e = ctypes.c_double(-1)
    rc = lib.cocr_selftest(0, 10, M, N, K + 100000 * bn, ctypes.byref(e))
