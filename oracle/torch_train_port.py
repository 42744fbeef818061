"""TEST INFRASTRUCTURE (checker only; never imported by conformer_ocr_b200/).

Training-mode restatement of `RecognitionModel._step` (/root/reference/conformer_ocr/model.py:129-145) on plain torch
ops with autograd: encoder in train() mode (BatchNorm1d batch statistics, convolution.py:141; dropout at the places
the reference has it, feed_forward.py:48-53, attention.py:99,134, convolution.py:146,225) -> decoder -> log_softmax ->
CTCLoss(reduction='sum', zero_infinity=True) (model.py:119,136-142).  Pinned against the UNMODIFIED reference modules
by tests/golden/make_golden_train.py (dropout 0: the reference's own masks come from torch's generator and cannot be
reproduced by another implementation).  Works on CPU or CUDA tensors, any floating dtype."""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch
import torch.nn.functional as F

from .torch_port import calc_length, rel_pos

BN_MOMENTUM = 0.1          # nn.BatchNorm1d default (convolution.py:141)


_MASK_FN = None      # (site, shape) -> tensor of dropout factors (0 or 1/(1-p)); set by train_forward(mask_fn=...)


def _drop(x, p, site=None):
    """nn.Dropout in train mode.  With a mask function installed the factors are REPLAYED (site numbering of csrc/train.cu:
    0 = input; layer l: 16 l + 1 + {0,1: first FFN hidden/out, 2,3: attention probabilities/out, 4: conv out, 6,7: second FFN})
    instead of drawn from torch's generator."""
    if p <= 0:
        return x
    if _MASK_FN is not None:
        return x * _MASK_FN(site, x.shape).to(x.dtype)
    return F.dropout(x, p, training=True)


def _ffn(x, w, p, pd, s0=None):
    y = F.layer_norm(x, x.shape[-1:], w[p + 'sequential.0.weight'], w[p + 'sequential.0.bias'])
    y = _drop(F.silu(F.linear(y, w[p + 'sequential.1.linear.weight'], w[p + 'sequential.1.linear.bias'])), pd, s0)
    return _drop(F.linear(y, w[p + 'sequential.4.linear.weight'], w[p + 'sequential.4.linear.bias']), pd, None if s0 is None else s0 + 1)


def _mhsa(x, w, p, h, pe, pd, s0=None):
    B, T, d = x.shape
    dh = d // h
    a = p + 'attention.'
    xn = F.layer_norm(x, (d,), w[p + 'layer_norm.weight'], w[p + 'layer_norm.bias'])
    q = F.linear(xn, w[a + 'query_proj.linear.weight'], w[a + 'query_proj.linear.bias']).view(B, T, h, dh)
    k = F.linear(xn, w[a + 'key_proj.linear.weight'], w[a + 'key_proj.linear.bias']).view(B, T, h, dh).permute(0, 2, 1, 3)
    v = F.linear(xn, w[a + 'value_proj.linear.weight'], w[a + 'value_proj.linear.bias']).view(B, T, h, dh).permute(0, 2, 1, 3)
    pos = F.linear(pe, w[a + 'pos_proj.linear.weight']).view(1, 2 * T - 1, h, dh)
    content = torch.matmul((q + w[a + 'u_bias']).transpose(1, 2), k.transpose(2, 3))
    ps = torch.matmul((q + w[a + 'v_bias']).transpose(1, 2), pos.permute(0, 2, 3, 1))
    ps = F.pad(ps, (1, 0)).reshape(B, h, 2 * T, T)[:, :, 1:].reshape(B, h, T, 2 * T - 1)[..., :T]     # attention.py:105-113
    attn = _drop(F.softmax((content + ps) / math.sqrt(dh), -1), pd, s0)
    ctx = torch.matmul(attn, v).transpose(1, 2).contiguous().view(B, T, d)
    return _drop(F.linear(ctx, w[a + 'out_proj.linear.weight'], w[a + 'out_proj.linear.bias']), pd, None if s0 is None else s0 + 1)


def _conv(x, w, p, K, pd, stats, s0=None):
    d = x.shape[-1]
    y = F.layer_norm(x, (d,), w[p + 'sequential.0.weight'], w[p + 'sequential.0.bias']).transpose(1, 2)
    y = F.glu(F.conv1d(y, w[p + 'sequential.2.conv.weight'], w[p + 'sequential.2.conv.bias']), dim=1)
    y = F.conv1d(y, w[p + 'sequential.4.conv.weight'], None, padding=(K - 1) // 2, groups=d)
    rm = w[p + 'sequential.5.running_mean'].detach().clone()
    rv = w[p + 'sequential.5.running_var'].detach().clone()
    y = F.batch_norm(y, rm, rv, w[p + 'sequential.5.weight'], w[p + 'sequential.5.bias'], training=True,
                     momentum=BN_MOMENTUM, eps=1e-5)
    stats[p + 'sequential.5.running_mean'] = rm
    stats[p + 'sequential.5.running_var'] = rv
    y = F.conv1d(F.silu(y), w[p + 'sequential.7.conv.weight'], w[p + 'sequential.7.conv.bias'])
    return _drop(y.transpose(1, 2).contiguous(), pd, s0)


def train_forward(line: torch.Tensor, lens: torch.Tensor, w: Dict[str, torch.Tensor], hp: Dict, dropout: float = 0.0,
                  taps: Dict | None = None, mask_fn=None) -> Tuple[torch.Tensor, torch.Tensor, Dict[str, torch.Tensor]]:
    """(N,1,H,W) -> (probits (N,T,C), encoder_lens, new BatchNorm running statistics); model.py:131-135 in train mode.
    `taps`, when given, receives the residual stream after every module (retain_grad'ed) under 'x<layer>.<module>'."""
    global _MASK_FN
    _MASK_FN = mask_fn
    x = line.squeeze(1).transpose(1, 2)
    s = 'encoder.conv_subsample.'
    olens = calc_length(lens)
    y = x.unsqueeze(1)
    y = F.relu(F.conv2d(y, w[s + 'conv.0.weight'], w[s + 'conv.0.bias'], stride=2, padding=1))
    y = F.conv2d(y, w[s + 'conv.2.weight'], w[s + 'conv.2.bias'], stride=2, padding=1, groups=y.shape[1])
    y = F.relu(F.conv2d(y, w[s + 'conv.3.weight'], w[s + 'conv.3.bias']))
    b, c, t, f = y.shape
    y = _drop(F.linear(y.transpose(1, 2).reshape(b, t, -1), w[s + 'out.0.weight'], w[s + 'out.0.bias']), dropout, 0)
    fac = 0.5 if hp.get('half_step_residual', True) else 1.0
    pe = rel_pos(t, hp['encoder_dim']).to(device=y.device, dtype=y.dtype)
    stats: Dict[str, torch.Tensor] = {}

    def tap(name, v):
        if taps is not None:
            if v.requires_grad:
                v.retain_grad()
            taps[name] = v
        return v

    y = tap('x0', y)
    for i in range(hp['num_encoder_layers']):
        p = f'encoder.layers.{i}.sequential.'
        y = tap(f'x{i}.0', _ffn(y, w, p + '0.module.', dropout, 16 * i + 1) * fac + y)
        y = tap(f'x{i}.1', _mhsa(y, w, p + '1.module.', hp['num_attention_heads'], pe, dropout, 16 * i + 3) + y)
        y = tap(f'x{i}.2', _conv(y, w, p + '2.module.', hp['conv_kernel_size'], dropout, stats, 16 * i + 5) + y)
        y = tap(f'x{i}.3', _ffn(y, w, p + '3.module.', dropout, 16 * i + 7) * fac + y)
        y = tap(f'x{i}.4', F.layer_norm(y, y.shape[-1:], w[p + '4.weight'], w[p + '4.bias']))
    return F.linear(y, w['decoder.weight'], w['decoder.bias']), olens, stats


def train_step_loss(line, lens, targets, target_lens, w, hp, dropout: float = 0.0, taps=None, mask_fn=None):
    """model.py:136-142: log_softmax, (N,T,C)->(T,N,C), CTCLoss(sum, zero_infinity).  Returns (loss, probits, lens, stats)."""
    probits, olens, stats = train_forward(line, lens, w, hp, dropout, taps, mask_fn)
    logp = F.log_softmax(probits, dim=-1)
    loss = F.ctc_loss(logp.transpose(0, 1), targets, olens, target_lens, blank=0, reduction='sum', zero_infinity=True)
    return loss, probits, olens, stats


def trainable_keys(w):
    return [k for k in w if not k.endswith(('running_mean', 'running_var', 'num_batches_tracked'))]


def philox_dropout_factors(seed: int, site: int, p: float, n: int):
    """numpy restatement of the kernels' dropout masks (csrc/dropout.cuh): Philox-4x32-10 with key = (seed low, seed high) and
    counter = (e >> 2 low, e >> 2 high, site, 0x5eed5eed); element e takes word e & 3 and is kept (factor 1 / (1 - p)) when the
    word >= floor(p * 2**32).  Returns float32 factors, shape (n,)."""
    import numpy as np
    if not p > 0:
        return np.ones(n, dtype=np.float32)
    thresh = np.uint64(min(max(int(p * 4294967296.0), 1), 4294967295))
    nc = (n + 3) // 4
    ctr = np.arange(nc, dtype=np.uint64)
    M32 = np.uint64(0xFFFFFFFF)
    c0, c1 = ctr & M32, ctr >> np.uint64(32)
    c2 = np.full(nc, site, dtype=np.uint64)
    c3 = np.full(nc, 0x5eed5eed, dtype=np.uint64)
    k0, k1 = np.uint64(seed & 0xFFFFFFFF), np.uint64((seed >> 32) & 0xFFFFFFFF)
    A, B = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    for _ in range(10):
        p0, p1 = A * c0, B * c2                      # 32 x 32 -> 64 bit products
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & M32, p1 >> np.uint64(32), p1 & M32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & M32, lo1, (hi0 ^ c3 ^ k1) & M32, lo0
        k0, k1 = (k0 + np.uint64(0x9E3779B9)) & M32, (k1 + np.uint64(0xBB67AE85)) & M32
    words = np.stack([c0, c1, c2, c3], axis=1).reshape(-1)[:n]
    return np.where(words >= thresh, np.float32(1.0 / (1.0 - p)), np.float32(0.0)).astype(np.float32)
