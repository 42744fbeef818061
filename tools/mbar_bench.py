"""mbarrier hand-shake latency (cocr_microbench 45): cycles per hop of a two-warp ping-pong."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.microbench.load import load
lib = load()
wn = {0: 'try_wait, all lanes', 1: 'try_wait, lane 0 + syncwarp', 2: 'test_wait spin, all lanes', 3: 'test_wait spin, lane 0 + syncwarp'}
for extra in (0, 16, 32, 48):
    for wm in range(4):
        c = ctypes.c_double(-1)
        rc = lib.cocr_microbench(0, 45, wm + extra, 0, 0, ctypes.byref(c))
        print(f'{wn[wm]:36s} {"+12 suspended warps" if extra & 16 else "":20s} {"pong = tcgen05.commit" if extra & 32 else "":22s}: {c.value:7.1f} cycles/hop rc={rc}')
