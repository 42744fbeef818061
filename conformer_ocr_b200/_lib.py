"""ctypes binding of libcocr_b200.so (include/cocr_b200.h).  This is the binding a maintainer of the
reference would add to call the C ABI from its Python modules."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libcocr_b200.so')

COCR_OK = 0
COCR_ERR_BAD_ARG = -1
COCR_ERR_UNSUPPORTED = -2
COCR_ERR_OOM = -3
COCR_ERR_CUDA = -4
COCR_ERR_STATE = -5

COCR_FP32 = 0
COCR_BF16 = 1
COCR_MASK_REFERENCE_UNMASKED = 0
COCR_MASK_RAGGED = 1


class cocr_hparams(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        'height', 'encoder_dim', 'num_layers', 'num_heads', 'ff_expansion', 'conv_expansion',
        'conv_kernel_size', 'half_step_residual', 'subsampling_conv_channels', 'subsampling_factor',
        'num_classes', 'precision', 'mask_mode')]


class cocr_train_config(C.Structure):
    _fields_ = [('input_dropout_p', C.c_float), ('feed_forward_dropout_p', C.c_float), ('attention_dropout_p', C.c_float),
                ('conv_dropout_p', C.c_float), ('bn_momentum', C.c_float), ('compute', C.c_uint32), ('seed', C.c_uint64)]


_P = C.c_void_p
_I32 = C.c_int32
_I64 = C.c_int64

# name -> (restype, argtypes); every symbol include/cocr_b200.h declares
SIGNATURES = {
    'cocr_version': (C.c_char_p, []),
    'cocr_last_error': (C.c_char_p, [_P]),
    'cocr_create': (C.c_int, [C.POINTER(cocr_hparams), C.c_int, C.POINTER(_P)]),
    'cocr_destroy': (C.c_int, [_P]),
    'cocr_set_weight': (C.c_int, [_P, C.c_char_p, _P, C.c_int, C.POINTER(_I64)]),
    'cocr_finalize_weights': (C.c_int, [_P]),
    'cocr_out_frames': (_I32, [_I32]),
    'cocr_encode': (C.c_int, [_P, _P, _I64, _I64, _I64, _P, _I32, _I32, _P, _P, _P]),
    'cocr_logits': (C.c_int, [_P, _P, _I64, _I64, _I64, _P, _I32, _I32, _P, _P, _P]),
    'cocr_greedy_decode': (C.c_int, [_P, _P, _P, _I32, _I32, _I32, _P, _P, _P, _P, _P, _P]),
    'cocr_forward_decode': (C.c_int, [_P, _P, _I64, _I64, _I64, _P, _I32, _I32, _P, _P, _P, _P, _P, _P, _P]),
    'cocr_recognize_host': (C.c_int, [_P, _P, _P, _I32, _I32, _P, _P, _P, _P, _P, _P, _P]),
    'cocr_recognize_host_begin': (C.c_int, [_P, _I32, _P, _P, _I32, _I32, _P]),
    'cocr_forward_decode_u8': (C.c_int, [_P, _P, _I64, _I64, _I64, _P, _I32, _I32, _P, _P, _P, _P, _P, _P, _P]),
    'cocr_recognize_host_u8_begin': (C.c_int, [_P, _I32, _P, _P, _I32, _I32, _P]),
    'cocr_recognize_host_end': (C.c_int, [_P, _I32, _P, _P, _P, _P, _P, _P]),
    'cocr_log_softmax': (C.c_int, [_P, _P, _I64, _I32, _P, _P]),
    'cocr_ctc_loss': (C.c_int, [_P, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _P, _P, _P, _P]),
    'cocr_project': (C.c_int, [_P, _P, _I64, _P, _P]),
    'cocr_decoder_backward': (C.c_int, [_P, _P, _P, _I64, _P, _P, _P, _P]),
    'cocr_train_layout': (C.c_int, [_P, C.POINTER(_I64), C.POINTER(_I64), C.POINTER(_I32)]),
    'cocr_train_entry': (C.c_int, [_P, _I32, C.c_char_p, C.c_size_t, C.POINTER(_I64), C.POINTER(_I64), C.POINTER(_I32)]),
    'cocr_train_workspace_bytes': (C.c_int, [_P, _I32, _I32, C.POINTER(C.c_size_t)]),
    'cocr_train_step': (C.c_int, [_P, _P, _P, _P, _I64, _I64, _I64, _P, _I32, _I32, _P, _P, _I32, C.POINTER(cocr_train_config), _P, _P, _P, _P]),
    'cocr_adamw_step': (C.c_int, [_P, _P, _P, _P, _P, _I64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, _I32, _P]),
    'cocr_dropout_mask': (C.c_int, [_P, C.c_uint64, C.c_uint32, C.c_float, _I64, _P, _P]),
    'cocr_slot_records_dev': (C.c_int, [_P, _I32, C.POINTER(_P), C.POINTER(C.c_size_t), C.POINTER(_I32), C.POINTER(_I32)]),
    'cocr_workspace_bytes': (C.c_int, [_P, _I32, _I32, C.POINTER(C.c_size_t)]),
    'cocr_kernel_plan': (C.c_int, [_P, C.c_char_p, C.c_size_t, C.POINTER(_I32)]),
    'cocr_launch_count': (_I64, [_P]),
    'cocr_set_profiling': (C.c_int, [_P, C.c_int]),
    'cocr_profile_report': (C.c_int, [_P, C.c_char_p, C.c_size_t]),
    'cocr_selftest': (C.c_int, [C.c_int, C.c_int, _I64, _I64, _I64, C.POINTER(C.c_double)]),
}

_lib = None


def load():
    """Load the shared library; raise (never fall back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f'{LIB_PATH} is missing: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                           '(or `make -C conformer_ocr_b200/csrc`). There is no CPU fallback.')
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def raise_for_status(rc: int, handle=None):
    """Map a cocr_status to the exception type the reference would raise."""
    if rc == COCR_OK:
        return
    msg = load().cocr_last_error(handle).decode('utf-8', 'replace')
    if rc == COCR_ERR_OOM:
        # lightning's is_oom_error looks for this prefix (model.py:146-151)
        raise RuntimeError(msg if msg.startswith('CUDA out of memory') else 'CUDA out of memory. ' + msg)
    if rc == COCR_ERR_BAD_ARG:
        raise ValueError(msg)
    if rc == COCR_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise RuntimeError(f'libcocr_b200 error {rc}: {msg}')
