"""Host-side mirror of `conformer_ocr.pred.PytorchRecognitionModel` (pred.py:50-210).

`forward`, `predict`, `predict_string`, `predict_labels`, `load_checkpoint` and `load_safetensors`
keep the reference's signatures, attributes (`nn`, `codec`, `ctc_decoder`, `height`, `channels`,
`width`) and return formats.  Differences, all on the device side:
  * encoder + decoder projection run in libcocr_b200.so (one engine holding both);
  * with the default `ctc_decoder` the argmax / run-collapse also runs on the GPU and only the
    collapsed (label, start, end, confidence) records cross PCIe instead of the (N, C, T) logits
    (pred.py:138,158,173).  A user-supplied `ctc_decoder` callable is honoured exactly like the
    reference does: logits are copied to the host and the callable sees `(C, T_i)` arrays.
"""
from __future__ import annotations

import json
import tarfile
from typing import TYPE_CHECKING, Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn

from .codec import LabelCodec
from .conformer.encoder import ConformerEncoder, EngineOwner
from .engine import out_frames, unpack_decoded

if TYPE_CHECKING:
    from os import PathLike


def greedy_decoder(outputs: np.ndarray) -> List[Tuple[int, int, int, float]]:
    """Default value of the `ctc_decoder` argument.  When `PytorchRecognitionModel` sees this
    function it decodes on the GPU (csrc/fp32_path.cu: k_argmax + k_ctc_collapse); it is never
    evaluated on the host by this package.  Calling it directly decodes one (C, T) score matrix on
    the current CUDA device with the same kernels."""
    from .engine import Engine  # noqa: F401  (kept local: pred -> engine only)
    t = torch.as_tensor(np.ascontiguousarray(outputs.T), dtype=torch.float32, device='cuda')[None]
    eng = _decode_engine(t.device)
    return unpack_decoded(eng.greedy_decode(t))[0]


_DECODE_ENGINES: Dict[int, object] = {}


def _decode_engine(device: torch.device):
    """A weight-less handle used only for `cocr_greedy_decode` on stand-alone logits."""
    from .engine import Engine
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _DECODE_ENGINES:
        _DECODE_ENGINES[idx] = Engine(height=8, encoder_dim=8, num_layers=0, num_heads=1, ff_expansion=1, conv_expansion=2,
                                      conv_kernel_size=3, half_step_residual=True, subsampling_conv_channels=1,
                                      subsampling_factor=4, num_classes=0, device=torch.device('cuda', idx))
    return _DECODE_ENGINES[idx]


def _make_codec(spec):
    """kraken's PytorchCodec when kraken is installed (pred.py:189,206), else the built-in table."""
    if spec is None or hasattr(spec, 'decode'):
        return spec
    try:
        from kraken.lib.codec import PytorchCodec  # type: ignore
        return PytorchCodec(spec)
    except ImportError:
        return LabelCodec(spec)


class PytorchRecognitionModel(EngineOwner):
    def __init__(self,
                 num_classes: int,
                 height: int,
                 encoder_dim: int,
                 num_encoder_layers: int,
                 num_attention_heads: int,
                 feed_forward_expansion_factor: int,
                 conv_expansion_factor: int,
                 input_dropout_p: float,
                 feed_forward_dropout_p: float,
                 attention_dropout_p: float,
                 conv_dropout_p: float,
                 conv_kernel_size: int,
                 half_step_residual: bool,
                 subsampling_conv_channels: int,
                 subsampling_factor: int,
                 codec=None,
                 ctc_decoder: Callable = greedy_decoder,
                 precision: str = 'fp32',
                 mask_mode: str = 'reference_unmasked',
                 **kwargs):
        """An nn.Module version of conformer_ocr.model.RecognitionModel for inference
        (pred.py:51-98); unknown hyper-parameters in **kwargs are ignored like in the reference."""
        super().__init__()
        encoder = ConformerEncoder(in_channels=1,
                                   input_dim=height,
                                   encoder_dim=encoder_dim,
                                   num_layers=num_encoder_layers,
                                   num_attention_heads=num_attention_heads,
                                   feed_forward_expansion_factor=feed_forward_expansion_factor,
                                   conv_expansion_factor=conv_expansion_factor,
                                   input_dropout_p=input_dropout_p,
                                   feed_forward_dropout_p=feed_forward_dropout_p,
                                   attention_dropout_p=attention_dropout_p,
                                   conv_dropout_p=conv_dropout_p,
                                   conv_kernel_size=conv_kernel_size,
                                   half_step_residual=half_step_residual,
                                   subsampling_conv_channels=subsampling_conv_channels,
                                   subsampling_factor=subsampling_factor,
                                   precision=precision, mask_mode=mask_mode)
        decoder = nn.Linear(encoder_dim, num_classes, bias=True)
        for p in decoder.parameters():
            p.requires_grad_(False)
        self.nn = nn.ModuleDict({'encoder': encoder,
                                 'decoder': decoder})
        self.codec = _make_codec(codec)
        self.ctc_decoder = ctc_decoder
        self.height = height
        self.channels = 1
        self.width = 0
        self.num_classes = num_classes
        self._init_engine_state(precision, mask_mode)
        self.eval()

    # -- engine plumbing ---------------------------------------------------------------------
    def _engine_kwargs(self):
        return dict(num_classes=self.num_classes, **self.nn['encoder'].hparams)

    def _engine_state(self):
        return dict(self.nn.state_dict())

    def train(self, mode: bool = True):
        if mode:
            raise NotImplementedError('conformer_ocr_b200 implements the inference path only')
        return super().train(False)

    @staticmethod
    def _as_bwh(line: torch.Tensor) -> torch.Tensor:
        if line.dim() != 4 or line.shape[1] != 1:
            raise ValueError(f'expected an (N, 1, H, W) line tensor, got {tuple(line.shape)}')
        return line.squeeze(1).transpose(1, 2)                       # pred.py:119 (a view, no copy)

    # -- reference API -----------------------------------------------------------------------
    def forward(self, line: torch.Tensor, lens: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """(N, 1, H, W) lines -> (probits (N, T, C) fp32 on the device, output lengths (N,) int32)
        (pred.py:101-122; the docstring there says numpy, the code returns tensors)."""
        x = self._as_bwh(line)
        eng = self.engine_for(x.device)
        probits, _ = eng.logits(x, lens)
        olens = out_frames(lens if lens is not None else torch.full((x.shape[0],), x.shape[1]))
        return probits, olens

    def _decode(self, line: torch.Tensor, lens) -> List[List[Tuple[int, int, int, float]]]:
        if self.ctc_decoder is greedy_decoder:
            return unpack_decoded(self._decode_arrays(line, lens))
        o, olens = self.forward(line, lens)
        o = o.transpose(1, 2).cpu().float().numpy()                   # pred.py:138
        return [self.ctc_decoder(seq[:, :int(seq_len)]) for seq, seq_len in zip(o, olens)]

    def _model_device(self) -> torch.device:
        return self.nn['decoder'].weight.device

    def _decode_arrays(self, line: torch.Tensor, lens) -> Dict[str, np.ndarray]:
        """Fused device decode of one batch -> host arrays (labels/starts/ends/conf (B,T), counts/out_lens (B,)).
        A host batch (the DataLoader hands out CPU tensors, cli/test.py:185-189) goes through
        `cocr_recognize_host_begin/_end`: H2D, forward, decode and the D2H of the records all inside the library, with no
        (N, C, T) logits and no per-label Python objects.  A device batch goes through `cocr_forward_decode`."""
        if line.device.type == 'cpu':
            eng = self.engine_for(self._model_device())
            if line.dim() != 4 or line.shape[1] != 1:
                raise ValueError(f'expected an (N, 1, H, W) line tensor, got {tuple(line.shape)}')
            if line.dtype not in (torch.float32, torch.uint8):
                line = line.float()
            slot = 0 if 0 not in getattr(eng, '_pending', {}) else 1
            eng.recognize_host_begin(slot, line.contiguous(), lens)
            return eng.recognize_host_end(slot)
        x = self._as_bwh(line)
        res = self.engine_for(x.device).forward_decode(x, lens)
        # one D2H per array, all queued before the first wait
        host = {k: v.to('cpu', non_blocking=True) for k, v in res.items()}
        torch.cuda.current_stream(x.device).synchronize()
        return {k: v.numpy() for k, v in host.items()}

    def _strings(self, line: torch.Tensor, lens) -> List[str]:
        if self.ctc_decoder is greedy_decoder and isinstance(self.codec, LabelCodec):
            # collapsed labels stay a (B, T) block: one table look-up for the batch instead of a tuple per label
            res = self._decode_arrays(line, lens)
            return self.codec.decode_strings(res['labels'], res['counts'])
        return [''.join(x[0] for x in self.codec.decode(locs)) for locs in self._decode(line, lens)]

    def predict_labels_arrays(self, line: torch.Tensor, lens: Optional[torch.Tensor] = None) -> Dict[str, np.ndarray]:
        """`predict_labels` without the list-of-tuples packaging: the (label, start, end, max) columns of pred.py:165-178
        as (B, T) arrays plus `counts` (records per line).  `unpack_decoded` turns them into the reference's format."""
        if self.ctc_decoder is not greedy_decoder:
            raise NotImplementedError('array output needs the built-in greedy decoder')
        return self._decode_arrays(line, lens)

    def predict_string_stream(self, batches, arrays: bool = False):
        """Pipelined `predict_string` over an iterable of host batches `(line (N,1,H,W) fp32/uint8 CPU tensor, lens)`, e.g.
        a DataLoader (cli/test.py:185-191).  Two batches are in flight: while batch i computes, batch i+1 is copied to
        the device on the handle's copy stream and the records of batch i-1 travel back.  Yields one result per batch, in
        order: the strings of `predict_string`, or the arrays of `predict_labels_arrays` when `arrays`.  Pinned batches
        (DataLoader(pin_memory=True)) overlap fully; pageable ones still work."""
        if self.ctc_decoder is not greedy_decoder or not (arrays or isinstance(self.codec, LabelCodec)):
            for line, lens in batches:                      # no fused path: plain per-batch calls
                yield self._strings(line, lens)
            return
        eng = self.engine_for(self._model_device())

        def finish(slot):
            res = eng.recognize_host_end(slot)
            return res if arrays else self.codec.decode_strings(res['labels'], res['counts'])
        pending: List[int] = []
        slot = 0
        try:
            for line, lens in batches:
                if line.device.type != 'cpu':
                    raise ValueError('predict_string_stream takes host batches; call predict_string for device tensors')
                if line.dtype not in (torch.float32, torch.uint8):
                    line = line.float()
                if len(pending) == 2:
                    yield finish(pending.pop(0))
                eng.recognize_host_begin(slot, line.contiguous(), lens)
                pending.append(slot)
                slot ^= 1
            while pending:
                yield finish(pending.pop(0))
        finally:
            # a consumer that stops early (or an exception) must not leave batches in the handle's slots
            while pending:
                try:
                    eng.recognize_host_end(pending.pop(0))
                except Exception:
                    pass

    def predict(self, line: torch.Tensor, lens: Optional[torch.Tensor] = None) -> List[str]:
        """pred.py:124-145 (returns the decoded strings, as the reference's code does)."""
        return self._strings(line, lens)

    def predict_string(self, line: torch.Tensor, lens: Optional[torch.Tensor] = None) -> List[str]:
        """pred.py:148-163"""
        return self._strings(line, lens)

    def predict_labels(self, line: torch.Tensor, lens: Optional[torch.Tensor] = None) -> List[List[Tuple[int, int, int, float]]]:
        """pred.py:165-178: list of (class, start, end, max) per line."""
        return self._decode(line, lens)

    # -- loaders (formats unchanged: SURVEY.md §5 "Checkpoint / resume") ------------------------
    @classmethod
    def load_safetensors(cls, path: 'PathLike', **engine_kwargs):
        """tar{model.safetensors (keys without the `nn.` prefix), metadata.json} (pred.py:180-195).  The container is
        parsed with numpy (convert.read_safetensors): the module's parameters become views of the archive's buffer (no random
        initialisation, no second copy) and the device engine is fed from the same views on first use."""
        from .convert import read_archive
        from .conformer import encoder as _enc
        metadata, weights = read_archive(path)
        _enc._SKIP_INIT = True                       # parameters are overwritten below: skip xavier / kaiming draws
        try:
            net = cls(**metadata['hyper_params'], codec=metadata['codec'], **engine_kwargs)
        finally:
            _enc._SKIP_INIT = False
        state = {k: torch.from_numpy(np.array(v, copy=True) if not v.flags.writeable else v) for k, v in weights.items()}
        net.nn.load_state_dict(state)
        return net.eval()

    @classmethod
    def load_checkpoint(cls, path: 'PathLike', **engine_kwargs):
        """Lightning checkpoint with `state_dict` (keys `nn.*`), `hyper_parameters` and
        `TextLineDataModule.codec` (pred.py:197-210)."""
        state_dict = torch.load(path, map_location='cpu', weights_only=False)
        if 'TextLineDataModule' not in state_dict:
            raise ValueError('Checkpoint does not contain data module state.')
        if 'hyper_parameters' not in state_dict:
            raise ValueError('No hyperparameters in state_dict')
        net = cls(**state_dict['hyper_parameters'], codec=state_dict['TextLineDataModule']['codec'], **engine_kwargs)
        net.load_state_dict(state_dict['state_dict'], strict=False)
        return net.eval()
