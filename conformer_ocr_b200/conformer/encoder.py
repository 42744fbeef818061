"""Host-side mirror of `conformer_ocr.conformer.encoder.ConformerEncoder` (encoder.py:107-204).

Same constructor signature, same `forward(inputs, input_lengths) -> (outputs, output_lengths)`
contract and the same state-dict key names / shapes, so checkpoints of the reference load with
`load_state_dict` unchanged.  The module owns parameters only; the arithmetic runs in
libcocr_b200.so (conformer_ocr_b200/csrc) through `Engine`.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn
from torch import Tensor

from ..engine import Engine, out_frames


def encoder_param_spec(input_dim: int, encoder_dim: int, num_layers: int, num_attention_heads: int,
                       feed_forward_expansion_factor: int, conv_kernel_size: int,
                       subsampling_conv_channels: int) -> List[Tuple[str, Tuple[int, ...], str]]:
    """(state-dict key, shape, kind) of every tensor of the reference encoder.
    kind: 'xavier' | 'default' | 'zeros' | 'ones' | 'buffer0' | 'buffer1' selects the init the
    reference would apply (modules.py:50-55, attention.py:65-68, torch defaults)."""
    d, ff, cc, K, h = encoder_dim, encoder_dim * feed_forward_expansion_factor, subsampling_conv_channels, \
        conv_kernel_size, num_attention_heads
    f = int(out_frames(input_dim))
    spec = [('conv_subsample.conv.0.weight', (cc, 1, 3, 3), 'default'), ('conv_subsample.conv.0.bias', (cc,), 'default_b:9'),
            ('conv_subsample.conv.2.weight', (cc, 1, 3, 3), 'default'), ('conv_subsample.conv.2.bias', (cc,), 'default_b:9'),
            ('conv_subsample.conv.3.weight', (cc, cc, 1, 1), 'default'), ('conv_subsample.conv.3.bias', (cc,), f'default_b:{cc}'),
            ('conv_subsample.out.0.weight', (d, cc * f), 'default'), ('conv_subsample.out.0.bias', (d,), f'default_b:{cc * f}')]
    def ffn(q):
        return [(q + '0.weight', (d,), 'ones'), (q + '0.bias', (d,), 'zeros'),
                (q + '1.linear.weight', (ff, d), 'xavier'), (q + '1.linear.bias', (ff,), 'zeros'),
                (q + '4.linear.weight', (d, ff), 'xavier'), (q + '4.linear.bias', (d,), 'zeros')]
    # Order = the reference's module registration order (encoder.py:67-103), i.e. the order of `parameters()` and
    # `state_dict()`: optimizer states of a Lightning checkpoint are indexed by it.
    for i in range(num_layers):
        p = f'layers.{i}.sequential.'
        spec += ffn(f'{p}0.module.sequential.')
        q = f'{p}1.module.'
        spec += [(q + 'layer_norm.weight', (d,), 'ones'), (q + 'layer_norm.bias', (d,), 'zeros'),
                 (q + 'attention.u_bias', (h, d // h), 'xavier'), (q + 'attention.v_bias', (h, d // h), 'xavier')]
        for nm in ('query', 'key', 'value'):
            spec += [(q + f'attention.{nm}_proj.linear.weight', (d, d), 'xavier'),
                     (q + f'attention.{nm}_proj.linear.bias', (d,), 'zeros')]
        spec += [(q + 'attention.pos_proj.linear.weight', (d, d), 'xavier'),
                 (q + 'attention.out_proj.linear.weight', (d, d), 'xavier'),
                 (q + 'attention.out_proj.linear.bias', (d,), 'zeros')]
        q = f'{p}2.module.sequential.'
        spec += [(q + '0.weight', (d,), 'ones'), (q + '0.bias', (d,), 'zeros'),
                 (q + '2.conv.weight', (2 * d, d, 1), 'default'), (q + '2.conv.bias', (2 * d,), f'default_b:{d}'),
                 (q + '4.conv.weight', (d, 1, K), 'default'),
                 (q + '5.weight', (d,), 'ones'), (q + '5.bias', (d,), 'zeros'),
                 (q + '5.running_mean', (d,), 'buffer0'), (q + '5.running_var', (d,), 'buffer1'),
                 (q + '5.num_batches_tracked', (), 'counter'),
                 (q + '7.conv.weight', (d, d, 1), 'default'), (q + '7.conv.bias', (d,), f'default_b:{d}')]
        spec += ffn(f'{p}3.module.sequential.')
        spec += [(p + '4.weight', (d,), 'ones'), (p + '4.bias', (d,), 'zeros')]
    return spec


class _Holder(nn.Module):
    """Name-space node of the parameter tree (no forward of its own)."""


def _attach(root: nn.Module, dotted: str, tensor: Tensor, buffer: bool):
    node = root
    parts = dotted.split('.')
    for part in parts[:-1]:
        if part not in node._modules:
            node.add_module(part, _Holder())
        node = node._modules[part]
    if buffer:
        node.register_buffer(parts[-1], tensor)
    else:
        node.register_parameter(parts[-1], nn.Parameter(tensor, requires_grad=False))


_SKIP_INIT = False      # set by the safetensors fast path: every tensor is about to be overwritten by load_state_dict


def _init_tensor(shape, kind: str) -> Tensor:
    t = torch.empty(shape, dtype=torch.float32)
    if _SKIP_INIT:
        return t.zero_()
    if kind == 'ones' or kind == 'buffer1':
        return t.fill_(1.0)
    if kind == 'zeros' or kind == 'buffer0':
        return t.zero_()
    if kind == 'xavier':
        return nn.init.xavier_uniform_(t)
    if kind == 'default':
        return nn.init.kaiming_uniform_(t, a=math.sqrt(5))
    if kind.startswith('default_b:'):
        bound = 1.0 / math.sqrt(int(kind.split(':')[1]))
        return t.uniform_(-bound, bound)
    raise ValueError(kind)


class EngineOwner(nn.Module):
    """Shared logic: lazily (re)build the device engine from the module's own state dict."""

    def _init_engine_state(self, precision: str, mask_mode: str):
        self.precision = precision
        self.mask_mode = mask_mode
        self._engine: Optional[Engine] = None
        self._engine_key = None
        # `net.nn.load_state_dict(weights)` is the reference idiom (pred.py:193): torch runs the post hooks of the module
        # load_state_dict was called on and of its descendants only, so the hook sits on every sub-module, not just on
        # the owner.  Called after the parameter tree is complete.
        import weakref
        owner = weakref.ref(self)

        def _invalidate(module, incompatible):
            o = owner()
            if o is not None:
                o.invalidate_engine()
        for m in self.modules():
            m.register_load_state_dict_post_hook(_invalidate)

    def invalidate_engine(self):
        """Call after editing parameters in place; the next forward re-uploads the weights."""
        self._engine_key = None

    def _apply(self, fn, *a, **kw):
        self.invalidate_engine()
        return super()._apply(fn, *a, **kw)

    def _engine_kwargs(self) -> Dict:
        raise NotImplementedError

    def _engine_state(self) -> Dict[str, Tensor]:
        raise NotImplementedError

    def engine_for(self, device: torch.device) -> Engine:
        device = torch.device(device)
        if device.type != 'cuda':
            raise RuntimeError('conformer_ocr_b200 runs on CUDA devices only; the input is on ' + str(device) +
                               ' (there is deliberately no CPU fallback)')
        idx = device.index if device.index is not None else torch.cuda.current_device()
        key = (idx, self.precision, self.mask_mode)
        if self._engine is None or self._engine_key != key:
            if self._engine is not None:
                self._engine.close()
            eng = Engine(device=torch.device('cuda', idx), precision=self.precision, mask_mode=self.mask_mode,
                         **self._engine_kwargs())
            eng.load_state_dict(self._engine_state())
            self._engine, self._engine_key = eng, key
        return self._engine


class ConformerEncoder(EngineOwner):
    """Drop-in for `conformer_ocr.conformer.encoder.ConformerEncoder` (inference).

    Extra keyword arguments (not in the reference): `precision` ('fp32' | 'bf16') and `mask_mode`
    ('reference_unmasked' — what the reference computes — or 'ragged')."""

    def __init__(self, in_channels: int = 1, input_dim: int = 80, encoder_dim: int = 512, num_layers: int = 17,
                 num_attention_heads: int = 8, feed_forward_expansion_factor: int = 4, conv_expansion_factor: int = 2,
                 input_dropout_p: float = 0.1, feed_forward_dropout_p: float = 0.1, attention_dropout_p: float = 0.1,
                 conv_dropout_p: float = 0.1, conv_kernel_size: int = 31, half_step_residual: bool = True,
                 subsampling_conv_channels: int = 256, subsampling_factor: int = 4, *, precision: str = 'fp32',
                 mask_mode: str = 'reference_unmasked'):
        super().__init__()
        # the reference's constructor checks, same messages
        assert encoder_dim % num_attention_heads == 0, "d_model % num_heads should be zero."              # attention.py:53
        assert (conv_kernel_size - 1) % 2 == 0, "kernel_size should be a odd number for 'SAME' padding"   # convolution.py:132
        assert conv_expansion_factor == 2, "Currently, Only Supports expansion_factor 2"                  # convolution.py:133
        if not math.log(subsampling_factor, 2).is_integer():
            raise ValueError('Sampling factor should be a power of 2.')                                   # convolution.py:180-181
        if subsampling_factor != 4:
            raise NotImplementedError('conformer_ocr_b200 implements subsampling_factor 4 only')
        self.hparams = dict(height=input_dim, encoder_dim=encoder_dim, num_layers=num_layers,
                            num_heads=num_attention_heads, ff_expansion=feed_forward_expansion_factor,
                            conv_expansion=conv_expansion_factor, conv_kernel_size=conv_kernel_size,
                            half_step_residual=half_step_residual, subsampling_conv_channels=subsampling_conv_channels,
                            subsampling_factor=subsampling_factor)
        # dropout probabilities: the identity in inference, used by the training step (conformer_ocr_b200.train.Trainer)
        self.dropout_ps = dict(input_dropout_p=input_dropout_p, feed_forward_dropout_p=feed_forward_dropout_p,
                               attention_dropout_p=attention_dropout_p, conv_dropout_p=conv_dropout_p)
        for key, shape, kind in encoder_param_spec(input_dim, encoder_dim, num_layers, num_attention_heads,
                                                   feed_forward_expansion_factor, conv_kernel_size,
                                                   subsampling_conv_channels):
            if kind == 'counter':
                _attach(self, key, torch.tensor(0, dtype=torch.long), buffer=True)
            else:
                _attach(self, key, _init_tensor(shape, kind), buffer=kind.startswith('buffer'))
        self._init_engine_state(precision, mask_mode)
        self.eval()

    # -- reference API ----------------------------------------------------------------------
    def count_parameters(self) -> int:
        """ Count parameters of encoder (encoder.py:173-175) """
        return sum(p.numel() for p in self.parameters())

    def update_dropout(self, dropout_p: float) -> None:
        """Dropout is the identity in inference; kept for interface parity (encoder.py:177-181)."""

    def train(self, mode: bool = True):
        """The encoder on its own has no training step (the loss lives in RecognitionModel, model.py:129-145): as a child of a
        RecognitionModel mirror it follows the parent's mode; stand-alone `forward` always computes in eval-mode arithmetic."""
        return super().train(mode)

    def _engine_kwargs(self):
        return dict(num_classes=0, **self.hparams)

    def _engine_state(self):
        return {'encoder.' + k: v for k, v in self.state_dict().items()}

    def forward(self, inputs: Tensor, input_lengths: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
        """inputs (batch, time=width, dim=height), input_lengths (batch) -> outputs (batch, T, d),
        output_lengths (batch) int32 on the device of `input_lengths` (encoder.py:183-204)."""
        eng = self.engine_for(inputs.device)
        out, _ = eng.encode(inputs, input_lengths)
        if input_lengths is None:
            lens = out_frames(torch.full((inputs.shape[0],), inputs.shape[1]))
        else:
            lens = out_frames(input_lengths)
        return out, lens
