"""CPU port of the hot path on torch's CPU operators (TEST / BASELINE INFRASTRUCTURE).

The reference executes this path as ATen CPU operators (conv2d, addmm, bmm, softmax, layer_norm,
glu, batch_norm ...).  `/root/reference` does not exist on the GPU box and a Python
reference cannot travel, so this file restates the same operator sequence functionally, reading the
weights from the flat state-dict the rest of the repo uses.  It is used (a) as a second, multi-threaded
checker next to the numpy oracle and (b) as `bench.py`'s `cpu_baseline` / `--impl reference` arm
(kind "port"), timed on the box's host cores.  Pinned against the golden vectors of the real
reference in tests/test_oracle_golden.py.  Never imported by conformer_ocr_b200/.
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from .conformer_oracle import greedy_decoder


def to_torch(weights: Dict[str, np.ndarray]) -> Dict[str, torch.Tensor]:
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in weights.items()}


def calc_length(lengths: torch.Tensor, repeat_num: int = 2) -> torch.Tensor:
    """convolution.py:240-247"""
    lengths = lengths.to(dtype=torch.float)
    for _ in range(repeat_num):
        lengths = torch.floor(torch.div(lengths + (2.0 - 3.0), 2) + 1.0)
    return lengths.to(dtype=torch.int)


def rel_pos(T: int, d: int) -> torch.Tensor:
    """embedding.py:42-56,66 for exactly the rows a length-T input uses: offsets T-1 .. -(T-1)."""
    position = torch.arange(T - 1, -T, -1, dtype=torch.float32).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d, 2, dtype=torch.float32) * -(math.log(10000.0) / d))
    pe = torch.zeros(2 * T - 1, d)
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe


def _ffn(x, w, p):
    y = F.layer_norm(x, x.shape[-1:], w[p + 'sequential.0.weight'], w[p + 'sequential.0.bias'])
    y = F.silu(F.linear(y, w[p + 'sequential.1.linear.weight'], w[p + 'sequential.1.linear.bias']))
    return F.linear(y, w[p + 'sequential.4.linear.weight'], w[p + 'sequential.4.linear.bias'])


def _mhsa(x, w, p, h, pe):
    B, T, d = x.shape
    dh = d // h
    a = p + 'attention.'
    xn = F.layer_norm(x, (d,), w[p + 'layer_norm.weight'], w[p + 'layer_norm.bias'])
    q = F.linear(xn, w[a + 'query_proj.linear.weight'], w[a + 'query_proj.linear.bias']).view(B, T, h, dh)
    k = F.linear(xn, w[a + 'key_proj.linear.weight'], w[a + 'key_proj.linear.bias']).view(B, T, h, dh).permute(0, 2, 1, 3)
    v = F.linear(xn, w[a + 'value_proj.linear.weight'], w[a + 'value_proj.linear.bias']).view(B, T, h, dh).permute(0, 2, 1, 3)
    pos = F.linear(pe, w[a + 'pos_proj.linear.weight']).view(1, 2 * T - 1, h, dh)
    content = torch.matmul((q + w[a + 'u_bias']).transpose(1, 2), k.transpose(2, 3))
    ps = torch.matmul((q + w[a + 'v_bias']).transpose(1, 2), pos.permute(0, 2, 3, 1))        # (B,h,T,2T-1)
    # relative shift (attention.py:105-113) as a strided gather: out[i, j] = ps[i, T-1-i+j]
    ps = F.pad(ps, (1, 0)).reshape(B, h, 2 * T, T)[:, :, 1:].reshape(B, h, T, 2 * T - 1)[..., :T]
    score = (content + ps) / math.sqrt(dh)
    attn = F.softmax(score, -1)
    ctx = torch.matmul(attn, v).transpose(1, 2).contiguous().view(B, T, d)
    return F.linear(ctx, w[a + 'out_proj.linear.weight'], w[a + 'out_proj.linear.bias'])


def _conv(x, w, p, K):
    d = x.shape[-1]
    y = F.layer_norm(x, (d,), w[p + 'sequential.0.weight'], w[p + 'sequential.0.bias']).transpose(1, 2)
    y = F.glu(F.conv1d(y, w[p + 'sequential.2.conv.weight'], w[p + 'sequential.2.conv.bias']), dim=1)
    y = F.conv1d(y, w[p + 'sequential.4.conv.weight'], None, padding=(K - 1) // 2, groups=d)
    y = F.batch_norm(y, w[p + 'sequential.5.running_mean'], w[p + 'sequential.5.running_var'],
                     w[p + 'sequential.5.weight'], w[p + 'sequential.5.bias'], training=False, eps=1e-5)
    y = F.conv1d(F.silu(y), w[p + 'sequential.7.conv.weight'], w[p + 'sequential.7.conv.bias'])
    return y.transpose(1, 2)


@torch.no_grad()
def recognition_forward(line: torch.Tensor, lens: torch.Tensor, w: Dict[str, torch.Tensor], hp: Dict
                        ) -> Tuple[torch.Tensor, torch.Tensor]:
    """pred.py:118-122 on CPU tensors: (N,1,H,W) -> (logits (N,T,C), out lens)."""
    x = line.squeeze(1).transpose(1, 2)
    s = 'encoder.conv_subsample.'
    olens = calc_length(lens)
    y = x.unsqueeze(1)
    y = F.relu(F.conv2d(y, w[s + 'conv.0.weight'], w[s + 'conv.0.bias'], stride=2, padding=1))
    y = F.conv2d(y, w[s + 'conv.2.weight'], w[s + 'conv.2.bias'], stride=2, padding=1, groups=y.shape[1])
    y = F.relu(F.conv2d(y, w[s + 'conv.3.weight'], w[s + 'conv.3.bias']))
    b, c, t, f = y.shape
    y = F.linear(y.transpose(1, 2).reshape(b, t, -1), w[s + 'out.0.weight'], w[s + 'out.0.bias'])
    fac = 0.5 if hp.get('half_step_residual', True) else 1.0
    pe = rel_pos(t, hp['encoder_dim']).to(y.device)
    for i in range(hp['num_encoder_layers']):
        p = f'encoder.layers.{i}.sequential.'
        y = _ffn(y, w, p + '0.module.') * fac + y
        y = _mhsa(y, w, p + '1.module.', hp['num_attention_heads'], pe) + y
        y = _conv(y, w, p + '2.module.', hp['conv_kernel_size']) + y
        y = _ffn(y, w, p + '3.module.') * fac + y
        y = F.layer_norm(y, y.shape[-1:], w[p + '4.weight'], w[p + '4.bias'])
    return F.linear(y, w['decoder.weight'], w['decoder.bias']), olens


def predict_labels(line: torch.Tensor, lens: torch.Tensor, w, hp) -> List[List[Tuple[int, int, int, float]]]:
    """pred.py:172-178 with the restated greedy decoder."""
    o, olens = recognition_forward(line, lens, w, hp)
    o = o.transpose(1, 2).cpu().float().numpy()
    return [greedy_decoder(seq[:, :int(n)]) for seq, n in zip(o, olens)]
