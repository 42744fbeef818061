"""cycles of a GROUP of n tcgen05.mma + commit + mbarrier wait (completion latency included)"""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.microbench.load import load
lib = load()
for a_tmem in (0, 1):
    for N in (64, 128, 256):
        for n in (1, 2, 4, 8, 12, 24, 48):
            c = ctypes.c_double(-1)
            rc = lib.cocr_microbench(0, 40, N, 1, a_tmem + 10 * n, ctypes.byref(c))
            print(f'A={"tmem" if a_tmem else "smem"} N={N:3d} group={n:2d}: {c.value * n:8.1f} cycles per group, {c.value:7.1f} per MMA  rc={rc}')
