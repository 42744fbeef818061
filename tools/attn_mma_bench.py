"""Tensor-pipe cost of the attention kernel's MMA sequence (cocr_microbench 44): cycles per key tile, total and issue-side,
alone and with eight warps imitating the softmax warps (TMEM traffic / MUFU+FMA work / named barriers)."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.microbench.load import load
lib = load()
names = {1: 'PV MN-major', 2: 'PV N=48', 4: 'no PV', 8: 'no S', 16: 'commit+wait every group', 64: 'S as N=256',
         128: 'sm:tmem', 256: 'sm:ex2', 512: 'sm:barrier', 2048: 'bulk copy stream', 4096: 'pollers all lanes', 8192: 'pollers lane 0'}
grid = int(os.environ.get('GRID', '1'))
modes = [int(a) for a in sys.argv[1:]] or [1, 1 + 128, 1 + 256, 1 + 512, 1 + 128 + 256, 1 + 128 + 256 + 512, 17 + 128 + 256 + 512, 4 + 128 + 256 + 512, 9 + 128 + 256 + 512]
for mode in modes:
    tot, iss, sm = ctypes.c_double(-1), ctypes.c_double(-1), ctypes.c_double(-1)
    rc = lib.cocr_microbench(0, 44, mode, 0, grid, ctypes.byref(tot))
    rc |= lib.cocr_microbench(0, 44, mode, 1, grid, ctypes.byref(iss))
    rc |= lib.cocr_microbench(0, 44, mode, 2, grid, ctypes.byref(sm))
    cp = ctypes.c_double(-1)
    rc |= lib.cocr_microbench(0, 44, mode, 3, grid, ctypes.byref(cp))
    desc = ', '.join(v for k, v in names.items() if mode & k) or 'S(12xN128) + PV(8xN64 K-major)'
    print(f'grid {grid} mode {mode:4d} [{desc}]: {tot.value:7.1f} cycles/tile, issue thread busy {iss.value:7.1f}, softmax-like loop {sm.value:7.1f} cycles/iter, copy {cp.value:6.1f} B/clk rc={rc}')
