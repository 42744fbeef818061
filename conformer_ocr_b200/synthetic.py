"""Deterministic synthetic weights / line images for benchmarks and smoke runs (there is no network for
checkpoints or datasets).  Same generator as the test suite's, so bench inputs and parity inputs agree;
tests/test_host_api.py asserts the two produce identical arrays."""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np



def weight_spec(hp: Dict, num_classes: int) -> List[Tuple[str, Tuple[int, ...]]]:
    """(state-dict key, shape) of `PytorchRecognitionModel.nn` (pred.py:75-92)."""
    return _draw_order(hp, num_classes)


def make_weights(hp: Dict, num_classes: int, seed: int = 0) -> Dict[str, np.ndarray]:
    """Seeded fp32 tensors for every state-dict key; biases, BatchNorm statistics and norm scales are
    non-trivial so the folded paths are exercised.  Draw order is fixed by
    `_DRAW_ORDER` so that the arrays equal the test suite's generator bit for bit."""
    rng = np.random.RandomState(seed)
    out: Dict[str, np.ndarray] = {}
    for key, shape in _draw_order(hp, num_classes):
        leaf = key.rsplit('.', 1)[-1]
        if leaf == 'running_var':
            v = rng.uniform(0.5, 1.5, size=shape)
        elif leaf == 'running_mean':
            v = rng.normal(0.0, 0.1, size=shape)
        elif leaf in ('u_bias', 'v_bias'):
            a = math.sqrt(6.0 / (shape[0] + shape[1]))
            v = rng.uniform(-a, a, size=shape)
        elif len(shape) == 1:
            v = rng.normal(1.0 if leaf == 'weight' else 0.0, 0.05, size=shape)
        else:
            fan_out = shape[0] * int(np.prod(shape[2:])) if len(shape) > 2 else shape[0]
            fan_in = int(np.prod(shape[1:]))
            a = math.sqrt(6.0 / (fan_in + fan_out))
            v = rng.uniform(-a, a, size=shape)
        out[key] = np.ascontiguousarray(v.astype(np.float32))
    return out


def _draw_order(hp: Dict, num_classes: int) -> List[Tuple[str, Tuple[int, ...]]]:
    d = hp['encoder_dim']
    ff = d * hp['feed_forward_expansion_factor']
    cc = hp['subsampling_conv_channels']
    K = hp['conv_kernel_size']
    h = hp['num_attention_heads']
    f = hp['height']
    for _ in range(2):
        f = (f + 2 - 3) // 2 + 1
    spec: List[Tuple[str, Tuple[int, ...]]] = [
        ('encoder.conv_subsample.conv.0.weight', (cc, 1, 3, 3)), ('encoder.conv_subsample.conv.0.bias', (cc,)),
        ('encoder.conv_subsample.conv.2.weight', (cc, 1, 3, 3)), ('encoder.conv_subsample.conv.2.bias', (cc,)),
        ('encoder.conv_subsample.conv.3.weight', (cc, cc, 1, 1)), ('encoder.conv_subsample.conv.3.bias', (cc,)),
        ('encoder.conv_subsample.out.0.weight', (d, cc * f)), ('encoder.conv_subsample.out.0.bias', (d,)),
    ]
    for i in range(hp['num_encoder_layers']):
        p = f'encoder.layers.{i}.sequential.'
        for ffn in ('0', '3'):
            q = f'{p}{ffn}.module.sequential.'
            spec += [(q + '0.weight', (d,)), (q + '0.bias', (d,)),
                     (q + '1.linear.weight', (ff, d)), (q + '1.linear.bias', (ff,)),
                     (q + '4.linear.weight', (d, ff)), (q + '4.linear.bias', (d,))]
        q = f'{p}1.module.'
        spec += [(q + 'layer_norm.weight', (d,)), (q + 'layer_norm.bias', (d,)),
                 (q + 'attention.u_bias', (h, d // h)), (q + 'attention.v_bias', (h, d // h))]
        for nm in ('query', 'key', 'value', 'out'):
            spec += [(q + f'attention.{nm}_proj.linear.weight', (d, d)), (q + f'attention.{nm}_proj.linear.bias', (d,))]
        spec += [(q + 'attention.pos_proj.linear.weight', (d, d))]
        q = f'{p}2.module.sequential.'
        spec += [(q + '0.weight', (d,)), (q + '0.bias', (d,)),
                 (q + '2.conv.weight', (2 * d, d, 1)), (q + '2.conv.bias', (2 * d,)),
                 (q + '4.conv.weight', (d, 1, K)),
                 (q + '5.weight', (d,)), (q + '5.bias', (d,)),
                 (q + '5.running_mean', (d,)), (q + '5.running_var', (d,)),
                 (q + '7.conv.weight', (d, d, 1)), (q + '7.conv.bias', (d,))]
        spec += [(p + '4.weight', (d,)), (p + '4.bias', (d,))]
    spec += [('decoder.weight', (num_classes, d)), ('decoder.bias', (num_classes,))]
    return spec


def make_lines(batch: int, height: int, width: int, seed: int = 1234,
               widths: Optional[Sequence[int]] = None) -> Tuple[np.ndarray, np.ndarray]:
    """(N,1,H,W) fp32 in [0,1), right-zero-padded, plus the true widths (int64): the layout kraken's
    collate_sequences hands to the model."""
    rng = np.random.RandomState(seed)
    img = rng.uniform(0.0, 1.0, size=(batch, 1, height, width)).astype(np.float32)
    if widths is None:
        lens = np.full((batch,), width, dtype=np.int64)
    else:
        lens = np.asarray(widths, dtype=np.int64)
        for b, wv in enumerate(lens):
            img[b, :, :, int(wv):] = 0.0
    return img, lens


def flops_per_line(width: int, num_classes: int, hp: Dict) -> float:
    """Algorithmic FLOPs (2*MAC) of one line: SURVEY.md §8(d) formula (positional scores counted once
    per (i,j); pos_proj excluded as batch-level; elementwise / softmax / LayerNorm excluded)."""
    d = hp['encoder_dim']
    ff = d * hp['feed_forward_expansion_factor']
    cc = hp['subsampling_conv_channels']
    K = hp['conv_kernel_size']
    L = hp['num_encoder_layers']
    f2 = (hp['height'] + 2 - 3) // 2 + 1
    f = (f2 + 2 - 3) // 2 + 1
    t2 = (width - 1) // 2 + 1
    t = (t2 - 1) // 2 + 1
    f_sub = 2 * 9 * cc * f2 * t2 + 2 * 9 * cc * f * t + 2 * cc * cc * f * t + 2 * (cc * f) * d * t
    f_ffn = 2 * (2 * d * ff) * 2 * t
    f_proj = 4 * (2 * d * d) * t
    f_att = 3 * (2 * t * t * d)
    f_conv = (2 * d * 2 * d + 2 * K * d + 2 * d * d) * t
    f_dec = 2 * d * num_classes * t
    return float(f_sub + L * (f_ffn + f_proj + f_att + f_conv) + f_dec)
