import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.microbench.load import load
lib = load()
for a_tmem in (0, 1):
    for mode in (0, 1, 2, 10, 12):
        c = ctypes.c_double(-1)
        rc = lib.cocr_microbench(0, 42, mode, a_tmem, 0, ctypes.byref(c))
        print(f'12 x N=128 MMAs, A={"tmem" if a_tmem else "smem"}, B {"3 distinct tiles" if mode >= 10 else "one tile"}, other warps: {["idle", "tcgen05.ld loop", "tcgen05.ld+st loop"][mode % 10]}: {c.value:7.1f} cycles per group rc={rc}')
