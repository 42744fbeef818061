"""Thin owner of one `cocr_handle`: weight upload and the forward / decode calls.

PyTorch is used for plumbing only (device memory of inputs/outputs, the current CUDA stream);
all arithmetic happens inside libcocr_b200.so.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Mapping, Optional, Tuple

import numpy as np
import torch

from . import _lib

_PRECISIONS = {'fp32': _lib.COCR_FP32, 'float32': _lib.COCR_FP32, '32': _lib.COCR_FP32,
               'bf16': _lib.COCR_BF16, 'bfloat16': _lib.COCR_BF16}
_MASKS = {'reference_unmasked': _lib.COCR_MASK_REFERENCE_UNMASKED, 'ragged': _lib.COCR_MASK_RAGGED}


def out_frames(widths):
    """calc_length (conformer/convolution.py:240-247) for subsampling factor 4, on the host, with
    the reference's float arithmetic; returns an int32 tensor like the reference does."""
    t = torch.as_tensor(widths)
    t = t.to(dtype=torch.float)
    for _ in range(2):
        t = torch.floor(torch.div(t + (2.0 - 3.0), 2) + 1.0)
    return t.to(dtype=torch.int)


class Engine:
    """One device handle (encoder [+ decoder]) on one GPU."""

    def __init__(self, *, height: int, encoder_dim: int, num_layers: int, num_heads: int,
                 ff_expansion: int, conv_expansion: int, conv_kernel_size: int, half_step_residual: bool,
                 subsampling_conv_channels: int, subsampling_factor: int, num_classes: int,
                 device: torch.device, precision: str = 'fp32', mask_mode: str = 'reference_unmasked'):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        if precision not in _PRECISIONS:
            raise ValueError(f'unknown precision {precision!r}; expected one of {sorted(_PRECISIONS)}')
        if mask_mode not in _MASKS:
            raise ValueError(f'unknown mask_mode {mask_mode!r}; expected one of {sorted(_MASKS)}')
        device = torch.device(device)
        if device.type != 'cuda':
            raise RuntimeError('conformer_ocr_b200 computes on a CUDA device only (no CPU fallback); '
                               f'got device {device}')
        if not torch.cuda.is_available():
            raise RuntimeError('no CUDA device available; conformer_ocr_b200 has no CPU fallback')
        self.device = torch.device('cuda', device.index if device.index is not None else torch.cuda.current_device())
        self.hp = _lib.cocr_hparams(height, encoder_dim, num_layers, num_heads, ff_expansion, conv_expansion,
                                    conv_kernel_size, int(bool(half_step_residual)), subsampling_conv_channels,
                                    subsampling_factor, num_classes, _PRECISIONS[precision], _MASKS[mask_mode])
        self.precision = precision
        self.mask_mode = mask_mode
        rc = self._lib.cocr_create(C.byref(self.hp), self.device.index, C.byref(self._h))
        _lib.raise_for_status(rc, None)
        self.num_classes = num_classes
        self.encoder_dim = encoder_dim
        self.height = height
        self._loaded = False

    # -- life cycle -------------------------------------------------------------------------
    def close(self):
        if getattr(self, '_h', None) is not None and self._h.value:
            self._lib.cocr_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        _lib.raise_for_status(rc, self._h)

    # -- weights ----------------------------------------------------------------------------
    def load_state_dict(self, state: Mapping[str, object]):
        """`state`: keys as in `PytorchRecognitionModel.nn.state_dict()` ("encoder.*", "decoder.*");
        values torch tensors or numpy arrays."""
        for key, val in state.items():
            if key.endswith('num_batches_tracked'):
                continue
            if isinstance(val, torch.Tensor):
                arr = val.detach().to(device='cpu', dtype=torch.float32).contiguous().numpy()
            else:
                arr = np.ascontiguousarray(np.asarray(val, dtype=np.float32))
            shape = (C.c_int64 * max(arr.ndim, 1))(*arr.shape)
            self._check(self._lib.cocr_set_weight(self._h, key.encode(), arr.ctypes.data_as(C.c_void_p), arr.ndim, shape))
        self._check(self._lib.cocr_finalize_weights(self._h))
        self._loaded = True
        plan, nf = self.kernel_plan()
        if nf and not Engine._warned_fallback:
            import warnings
            Engine._warned_fallback = True
            warnings.warn('conformer_ocr_b200: these hyper-parameters are outside the geometry the fastest kernels are built for; '
                          f'{nf} component(s) run on general kernels:\n' + plan, RuntimeWarning, stacklevel=2)

    _warned_fallback = False

    def kernel_plan(self) -> Tuple[str, int]:
        """(description of the kernels selected for these hyper-parameters, number of components on a general fallback kernel)."""
        buf = C.create_string_buffer(2048)
        nf = C.c_int32()
        self._check(self._lib.cocr_kernel_plan(self._h, buf, len(buf), C.byref(nf)))
        return buf.value.decode(), nf.value

    # -- helpers ----------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _prep(self, inputs: torch.Tensor, lens, allow_u8: bool = False) -> Tuple[torch.Tensor, Optional[torch.Tensor], int, int]:
        """inputs: (B, W, H) view (any strides) on this engine's device."""
        if inputs.dim() != 3:
            raise ValueError(f'expected (batch, width, height) input, got shape {tuple(inputs.shape)}')
        if inputs.device != self.device:
            if inputs.device.type != 'cuda':
                raise RuntimeError('input tensor is on the CPU; move it to the GPU (no CPU fallback)')
            raise RuntimeError(f'input on {inputs.device}, engine on {self.device}')
        if inputs.dtype == torch.uint8 and not (allow_u8 and self.precision == 'bf16'):
            # raw pixels -> ToTensor scaling.  A tensor divisor: dividing by a Python scalar multiplies by the rounded
            # reciprocal on CUDA and differs from p / 255 in the last bit for some pixel values
            inputs = inputs.float() / torch.full((), 255.0, device=inputs.device)
        elif inputs.dtype not in (torch.float32, torch.uint8):   # uint8 kept: scaled by 1/255 inside the first kernel
            inputs = inputs.float()
        B, W, H = inputs.shape
        if H != self.height:
            raise ValueError(f'input height {H} does not match model height {self.height}')
        wd = None
        if lens is not None:
            wd = torch.as_tensor(lens).to(device=self.device, dtype=torch.int32, non_blocking=True).contiguous()
            if wd.numel() != B:
                raise ValueError(f'{wd.numel()} sequence lengths for a batch of {B}')
        return inputs, wd, B, W

    @staticmethod
    def _ptr(t: Optional[torch.Tensor]):
        return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p()

    # -- forward calls ----------------------------------------------------------------------
    def encode(self, inputs: torch.Tensor, lens=None) -> Tuple[torch.Tensor, torch.Tensor]:
        """ConformerEncoder.forward: (B,W,H) -> ((B,T,d) fp32 on device, (B,) int32 device lens)."""
        x, wd, B, W = self._prep(inputs, lens)
        T = int(self._lib.cocr_out_frames(W))
        out = torch.empty((B, T, self.encoder_dim), dtype=torch.float32, device=self.device)
        olens = torch.empty((B,), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self._lib.cocr_encode(self._h, self._ptr(x), x.stride(0), x.stride(1), x.stride(2), self._ptr(wd),
                                              B, W, self._stream(), self._ptr(out), self._ptr(olens)))
        return out, olens

    def logits(self, inputs: torch.Tensor, lens=None) -> Tuple[torch.Tensor, torch.Tensor]:
        x, wd, B, W = self._prep(inputs, lens)
        T = int(self._lib.cocr_out_frames(W))
        out = torch.empty((B, T, self.num_classes), dtype=torch.float32, device=self.device)
        olens = torch.empty((B,), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self._lib.cocr_logits(self._h, self._ptr(x), x.stride(0), x.stride(1), x.stride(2), self._ptr(wd),
                                              B, W, self._stream(), self._ptr(out), self._ptr(olens)))
        return out, olens

    def forward_decode(self, inputs: torch.Tensor, lens=None) -> Dict[str, torch.Tensor]:
        """Fused encoder -> projection -> argmax -> CTC collapse; results stay on the device."""
        x, wd, B, W = self._prep(inputs, lens, allow_u8=True)
        T = int(self._lib.cocr_out_frames(W))
        i32 = dict(dtype=torch.int32, device=self.device)
        res = dict(labels=torch.empty((B, T), **i32), starts=torch.empty((B, T), **i32), ends=torch.empty((B, T), **i32),
                   conf=torch.empty((B, T), dtype=torch.float32, device=self.device),
                   counts=torch.empty((B,), **i32), out_lens=torch.empty((B,), **i32))
        fn = self._lib.cocr_forward_decode_u8 if x.dtype == torch.uint8 else self._lib.cocr_forward_decode
        with torch.cuda.device(self.device):
            self._check(fn(
                self._h, self._ptr(x), x.stride(0), x.stride(1), x.stride(2), self._ptr(wd), B, W, self._stream(),
                self._ptr(res['labels']), self._ptr(res['starts']), self._ptr(res['ends']), self._ptr(res['conf']),
                self._ptr(res['counts']), self._ptr(res['out_lens'])))
        return res

    def greedy_decode(self, logits: torch.Tensor, lens=None) -> Dict[str, torch.Tensor]:
        """greedy_decoder over device logits (B,T,C)."""
        if logits.device != self.device or logits.dtype != torch.float32:
            logits = logits.to(device=self.device, dtype=torch.float32)
        logits = logits.contiguous()
        B, T, Cn = logits.shape
        ld = None
        if lens is not None:
            ld = torch.as_tensor(lens).to(device=self.device, dtype=torch.int32).contiguous()
        i32 = dict(dtype=torch.int32, device=self.device)
        res = dict(labels=torch.empty((B, T), **i32), starts=torch.empty((B, T), **i32), ends=torch.empty((B, T), **i32),
                   conf=torch.empty((B, T), dtype=torch.float32, device=self.device), counts=torch.empty((B,), **i32))
        with torch.cuda.device(self.device):
            self._check(self._lib.cocr_greedy_decode(self._h, self._ptr(logits), self._ptr(ld), B, T, Cn, self._stream(),
                                                     self._ptr(res['labels']), self._ptr(res['starts']),
                                                     self._ptr(res['ends']), self._ptr(res['conf']), self._ptr(res['counts'])))
        return res

    def recognize_host(self, img: torch.Tensor, widths: Optional[torch.Tensor] = None) -> Dict[str, np.ndarray]:
        """Host-buffer entry point (cli/test.py:185-191 equivalent): `img` (N,1,H,W) fp32 host tensor
        (pinned for full speed), `widths` int32 host tensor.  H2D, forward, decode and D2H all happen
        inside the call; returns numpy arrays."""
        if img.device.type != 'cpu' or img.dtype != torch.float32 or not img.is_contiguous():
            raise ValueError('recognize_host expects a contiguous fp32 host tensor (N,1,H,W)')
        B, ch, H, W = img.shape
        if ch != 1 or H != self.height:
            raise ValueError(f'expected (N,1,{self.height},W) input, got {tuple(img.shape)}')
        T = int(self._lib.cocr_out_frames(W))
        wh = None
        if widths is not None:
            wh = torch.as_tensor(widths).to(dtype=torch.int32).contiguous()
        res = dict(labels=np.empty((B, T), np.int32), starts=np.empty((B, T), np.int32), ends=np.empty((B, T), np.int32),
                   conf=np.empty((B, T), np.float32), counts=np.empty((B,), np.int32), out_lens=np.empty((B,), np.int32))
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        with torch.cuda.device(self.device):
            self._check(self._lib.cocr_recognize_host(
                self._h, C.c_void_p(img.data_ptr()), self._ptr(wh), B, W, self._stream(), p(res['labels']),
                p(res['starts']), p(res['ends']), p(res['conf']), p(res['counts']), p(res['out_lens'])))
        return res

    def recognize_host_begin(self, slot: int, img: torch.Tensor, widths: Optional[torch.Tensor] = None) -> None:
        """Enqueue one host batch in `slot` (0 or 1) without waiting: H2D on the handle's copy stream, forward + decode
        + D2H on the current stream.  Keep `img` / `widths` alive and unchanged until `recognize_host_end(slot)`.
        Two slots let the copy of batch i+1 overlap the forward pass of batch i."""
        if img.device.type != 'cpu' or img.dtype not in (torch.float32, torch.uint8) or not img.is_contiguous():
            raise ValueError('recognize_host_begin expects a contiguous fp32 (or uint8: raw pixels) host tensor (N,1,H,W)')
        B, ch, H, W = img.shape
        if ch != 1 or H != self.height:
            raise ValueError(f'expected (N,1,{self.height},W) input, got {tuple(img.shape)}')
        wh = None
        if widths is not None:
            wh = torch.as_tensor(widths).to(dtype=torch.int32).contiguous()
        fn = self._lib.cocr_recognize_host_u8_begin if img.dtype == torch.uint8 else self._lib.cocr_recognize_host_begin
        with torch.cuda.device(self.device):
            self._check(fn(self._h, slot, C.c_void_p(img.data_ptr()), self._ptr(wh), B, W, self._stream()))
        self._pending = getattr(self, '_pending', {})
        self._pending[slot] = (img, wh, B, int(self._lib.cocr_out_frames(W)))      # keeps the host buffers alive

    def recognize_host_end(self, slot: int) -> Dict[str, np.ndarray]:
        pending = getattr(self, '_pending', {})
        if slot not in pending:
            dummy = np.empty((1,), np.int32)
            q = dummy.ctypes.data_as(C.c_void_p)
            self._check(self._lib.cocr_recognize_host_end(self._h, slot, q, q, q, q, q, q))     # raises: nothing pending
            raise RuntimeError(f'slot {slot} holds no batch')
        _, _, B, T = pending.pop(slot)
        res = dict(labels=np.empty((B, T), np.int32), starts=np.empty((B, T), np.int32), ends=np.empty((B, T), np.int32),
                   conf=np.empty((B, T), np.float32), counts=np.empty((B,), np.int32), out_lens=np.empty((B,), np.int32))
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        with torch.cuda.device(self.device):
            self._check(self._lib.cocr_recognize_host_end(self._h, slot, p(res['labels']), p(res['starts']), p(res['ends']),
                                                          p(res['conf']), p(res['counts']), p(res['out_lens'])))
        return res

    # -- training-side tail (model.py:136-142) -------------------------------------------------
    def log_softmax(self, logits: torch.Tensor) -> torch.Tensor:
        """nn.functional.log_softmax(logits, dim=-1) (model.py:136) on device fp32 logits."""
        if logits.device != self.device or logits.dtype != torch.float32:
            logits = logits.to(device=self.device, dtype=torch.float32)
        logits = logits.contiguous()
        out = torch.empty_like(logits)
        Cn = logits.shape[-1]
        with torch.cuda.device(self.device):
            self._check(self._lib.cocr_log_softmax(self._h, self._ptr(logits), logits.numel() // Cn, Cn, self._stream(), self._ptr(out)))
        return out

    def ctc_loss(self, logits: torch.Tensor, targets, input_lens, target_lens, want_grad: bool = False):
        """nn.CTCLoss(reduction='sum', zero_infinity=True)(log_softmax(logits).transpose(0, 1), targets, input_lens,
        target_lens) (model.py:119,136-142) from the RAW logits (B,T,C).  Returns (loss (0-d device tensor), per-line losses
        (B,), grad wrt logits (B,T,C) or None)."""
        if logits.device != self.device or logits.dtype != torch.float32:
            logits = logits.to(device=self.device, dtype=torch.float32)
        logits = logits.contiguous()
        B, T, Cn = logits.shape
        tl_host = torch.as_tensor(target_lens).to('cpu', torch.int64)
        if tl_host.numel() != B:
            raise ValueError(f'{tl_host.numel()} target lengths for a batch of {B}')
        tg = torch.as_tensor(targets).to(device=self.device, dtype=torch.int32).contiguous().view(-1)
        if int(tl_host.sum()) != tg.numel():
            raise ValueError('targets must be the 1-D concatenation of the label sequences (model.py:131-133)')
        if tg.numel() and (int(tg.min()) < 1 or int(tg.max()) >= Cn):
            raise ValueError('target labels must lie in 1 .. num_classes-1 (0 is the CTC blank)')
        tl = tl_host.to(device=self.device, dtype=torch.int32)
        il = None if input_lens is None else torch.as_tensor(input_lens).to(device=self.device, dtype=torch.int32).contiguous()
        loss = torch.zeros((), dtype=torch.float32, device=self.device)
        per = torch.empty((B,), dtype=torch.float32, device=self.device)
        grad = torch.empty_like(logits) if want_grad else None
        with torch.cuda.device(self.device):
            self._check(self._lib.cocr_ctc_loss(self._h, self._ptr(logits), self._ptr(tg), self._ptr(tl), self._ptr(il), B, T, Cn,
                                                int(tl_host.max()) if B else 0, self._stream(), self._ptr(loss), self._ptr(per),
                                                self._ptr(grad)))
        return loss, per, grad

    def project(self, enc: torch.Tensor) -> torch.Tensor:
        """decoder nn.Linear on encoder outputs (..., d) -> logits (..., C), fp32 (pred.py:121, model.py:135)."""
        enc = enc.to(device=self.device, dtype=torch.float32).contiguous()
        out = torch.empty(enc.shape[:-1] + (self.num_classes,), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self._lib.cocr_project(self._h, self._ptr(enc), enc.numel() // enc.shape[-1], self._stream(), self._ptr(out)))
        return out

    def decoder_backward(self, grad_logits: torch.Tensor, enc: torch.Tensor, want_grad_enc: bool = True):
        """Backward pass of the decoder nn.Linear: (grad_weight (C,d), grad_bias (C,), grad_enc like enc or None)."""
        g = grad_logits.to(device=self.device, dtype=torch.float32).contiguous()
        e = enc.to(device=self.device, dtype=torch.float32).contiguous()
        rows = e.numel() // e.shape[-1]
        if g.numel() != rows * self.num_classes:
            raise ValueError('grad_logits and enc disagree on the number of frames')
        gw = torch.empty((self.num_classes, self.encoder_dim), dtype=torch.float32, device=self.device)
        gb = torch.empty((self.num_classes,), dtype=torch.float32, device=self.device)
        ge = torch.empty_like(e) if want_grad_enc else None
        with torch.cuda.device(self.device):
            self._check(self._lib.cocr_decoder_backward(self._h, self._ptr(g), self._ptr(e), rows, self._stream(), self._ptr(gw),
                                                        self._ptr(gb), self._ptr(ge)))
        return gw, gb, ge

    def slot_records(self, slot: int) -> Dict[str, torch.Tensor]:
        """Device views (no copy) of the records of the batch last begun in `slot`: labels/starts/ends (B,T) int32,
        conf (B,T) fp32, counts/out_lens (B,) int32.  Valid until the slot is begun again; complete after
        `recognize_host_end(slot)`.  Used to feed the gather of decoded records across GPUs without a host round trip."""
        ptr, nbytes, B, T = C.c_void_p(), C.c_size_t(), C.c_int32(), C.c_int32()
        self._check(self._lib.cocr_slot_records_dev(self._h, slot, C.byref(ptr), C.byref(nbytes), C.byref(B), C.byref(T)))

        class _Block:      # __cuda_array_interface__ v2 over library-owned memory
            pass
        blk = _Block()
        blk.__cuda_array_interface__ = dict(shape=(nbytes.value // 4,), typestr='<i4', data=(ptr.value, False), version=2)
        flat = torch.as_tensor(blk, device=self.device)
        b, t = B.value, T.value
        bt = b * t
        out = dict(labels=flat[:bt].view(b, t), starts=flat[bt:2 * bt].view(b, t), ends=flat[2 * bt:3 * bt].view(b, t),
                   conf=flat[3 * bt:4 * bt].view(torch.float32).view(b, t), counts=flat[4 * bt:4 * bt + b],
                   out_lens=flat[4 * bt + b:4 * bt + 2 * b])
        return out

    def workspace_bytes(self, B: int, W: int) -> int:
        n = C.c_size_t()
        self._check(self._lib.cocr_workspace_bytes(self._h, B, W, C.byref(n)))
        return int(n.value)

    @property
    def launch_count(self) -> int:
        return int(self._lib.cocr_launch_count(self._h))


def unpack_decoded(res: Mapping[str, object]) -> List[List[Tuple[int, int, int, float]]]:
    """Device/host decode block -> the list-of-tuples format of kraken's greedy_decoder."""
    def host(v):
        return v.cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
    labels, starts, ends, conf, counts = (host(res[k]) for k in ('labels', 'starts', 'ends', 'conf', 'counts'))
    out = []
    for b in range(labels.shape[0]):
        n = int(counts[b])
        out.append(list(zip(labels[b, :n].tolist(), starts[b, :n].tolist(), ends[b, :n].tolist(), conf[b, :n].tolist())))
    return out
