import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.microbench.load import load
lib = load()
for kind, name in enumerate(['ex2.approx.ftz.f32', 'ex2.approx.f16x2', 'tanh.approx.f32', 'tanh.approx.f16x2', 'ex2.approx.ftz.bf16x2', 'rcp.approx.ftz.f32']):
    c = ctypes.c_double(-1)
    rc = lib.cocr_microbench(0, 43, kind, 0, 0, ctypes.byref(c))
    print(f'{name:24s}: {c.value:6.2f} cycles per warp instruction per scheduler  rc={rc}')
