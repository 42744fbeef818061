"""ctypes loader of tools/microbench/libcocr_microbench.so (build: make -C tools/microbench)."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))


def load():
    path = os.path.join(_HERE, 'libcocr_microbench.so')
    if not os.path.exists(path):
        raise RuntimeError(f'{path} is missing: run `make -C tools/microbench`')
    lib = ctypes.CDLL(path)
    lib.cocr_microbench.restype = ctypes.c_int
    lib.cocr_microbench.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong,
                                    ctypes.POINTER(ctypes.c_double)]
    lib.cocr_microbench_error.restype = ctypes.c_char_p
    return lib
