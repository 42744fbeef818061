"""Tensor-pipe micro-benchmark (cocr_microbench 40): cycles per tcgen05.mma for N / accumulator chains / A source."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.microbench.load import load
lib = load()
for a_tmem in (0, 1):
    for N in (64, 128, 256):
        for chains in (1, 2, 3):
            if chains * N > 448:
                continue
            c = ctypes.c_double(-1)
            rc = lib.cocr_microbench(0, 40, N, chains, a_tmem, ctypes.byref(c))
            print(f'A={"tmem" if a_tmem else "smem"} N={N:3d} chains={chains}: {c.value:7.1f} cycles/MMA (floor {N/2:.0f}) rc={rc}')
for shape in (0, 1):
    for depth in (1, 2):
        for with_mma in (0, 1):
            c = ctypes.c_double(-1)
            rc = lib.cocr_microbench(0, 41, depth, with_mma, shape, ctypes.byref(c))
            print(f'tcgen05.ld {"32x32b.x32" if shape == 0 else "16x256b.x8"} depth={depth} mma={with_mma}: {c.value:7.1f} B/clk/SM rc={rc}')
