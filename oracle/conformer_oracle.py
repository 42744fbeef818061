"""CPU oracle for the conformer_ocr hot path (TEST INFRASTRUCTURE, not product code).

This file is a numpy restatement of the algorithm the reference implements with
torch modules.  It exists only so that `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` legs can check (and time) the
CUDA path against an independent CPU implementation.  Nothing under
`conformer_ocr_b200/` imports it; the product path fails loudly when the CUDA
library is missing.

Parity pinning
--------------
* encoder / decoder arithmetic: PINNED.  `tests/golden/make_golden.py` imports
  the reference's own `conformer_ocr.conformer.*` modules from /root/reference,
  runs them on seeded inputs/weights and commits the outputs under
  `tests/golden/*.npz`; `tests/test_oracle_golden.py` checks this restatement
  against every one of those vectors.
* greedy CTC decode: "parity unpinned" for the 4th (confidence) column.  The
  arithmetic lives in third-party `kraken.lib.ctc_decoder.greedy_decoder`
  (setup.cfg:36 `kraken>=4.3.13`, not vendored, not installable here) and the
  reference has no test or fixture for it.  label/start/end follow from the
  call-site contract (pred.py:143,162,177; model.py:167): argmax over classes,
  merge runs, drop blank 0.

Every function cites the reference file:line it follows (paths relative to
/root/reference/conformer_ocr/).
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import numpy as np

Array = np.ndarray


# --------------------------------------------------------------------------- #
# hyper-parameters (default_specs.py:48-61)
# --------------------------------------------------------------------------- #
DEFAULT_HPARAMS = dict(
    height=96,
    encoder_dim=144,
    num_encoder_layers=16,
    num_attention_heads=4,
    feed_forward_expansion_factor=4,
    conv_expansion_factor=2,
    conv_kernel_size=31,
    half_step_residual=True,
    subsampling_conv_channels=32,
    subsampling_factor=4,
)


# --------------------------------------------------------------------------- #
# small helpers
# --------------------------------------------------------------------------- #
def _sigmoid(x: Array) -> Array:
    return 1.0 / (1.0 + np.exp(-x))


def silu(x: Array) -> Array:
    """nn.SiLU: x * sigmoid(x) (feed_forward.py:48, convolution.py:142)."""
    return x * _sigmoid(x)


def layer_norm(x: Array, weight: Array, bias: Array, eps: float = 1e-5) -> Array:
    """nn.LayerNorm over the last axis, biased variance, eps 1e-5
    (feed_forward.py:46, attention.py:138,148, convolution.py:136, encoder.py:99)."""
    mu = x.mean(axis=-1, keepdims=True)
    xc = x - mu
    var = (xc * xc).mean(axis=-1, keepdims=True)
    return xc / np.sqrt(var + eps) * weight + bias


def linear(x: Array, weight: Array, bias: Array | None = None) -> Array:
    """nn.Linear: x @ W^T + b (modules.py:47-58)."""
    y = x @ weight.T
    if bias is not None:
        y = y + bias
    return y


def calc_length(lengths: Array, all_paddings: int = 2, kernel_size: int = 3, stride: int = 2,
                repeat_num: int = 2) -> Array:
    """convolution.py:240-247: floor((L + pad - k)/stride + 1), `repeat_num` times, in
    float32, result int32."""
    add_pad = np.float32(all_paddings - kernel_size)
    out = np.asarray(lengths).astype(np.float32)
    for _ in range(repeat_num):
        out = np.floor((out + add_pad) / np.float32(stride) + np.float32(1.0))
    return out.astype(np.int32)


# --------------------------------------------------------------------------- #
# convolutional front-end (convolution.py:151-237)
# --------------------------------------------------------------------------- #
def _conv2d_3x3_s2_p1(x: Array, weight: Array, bias: Array, depthwise: bool) -> Array:
    """3x3 / stride 2 / padding 1 convolution in NCHW.

    `x` (B, Cin, Hh, Ww); dense: weight (Cout, Cin, 3, 3); depthwise: (C, 1, 3, 3)
    (convolution.py:192-205)."""
    B, Cin, Hh, Ww = x.shape
    Ho = (Hh + 2 - 3) // 2 + 1
    Wo = (Ww + 2 - 3) // 2 + 1
    xp = np.zeros((B, Cin, Hh + 2, Ww + 2), dtype=x.dtype)
    xp[:, :, 1:Hh + 1, 1:Ww + 1] = x
    Cout = weight.shape[0]
    out = np.zeros((B, Cout, Ho, Wo), dtype=x.dtype)
    for kh in range(3):
        for kw in range(3):
            patch = xp[:, :, kh:kh + 2 * Ho - 1:2, kw:kw + 2 * Wo - 1:2]  # (B, Cin, Ho, Wo)
            if depthwise:
                out += patch * weight[None, :, 0, kh, kw, None, None]
            else:
                out += np.einsum('bchw,oc->bohw', patch, weight[:, :, kh, kw])
    return out + bias[None, :, None, None]


def conv2d_subsampling(inputs: Array, input_lengths: Array, w: Dict[str, Array],
                       prefix: str = 'conv_subsample.') -> Tuple[Array, Array]:
    """Conv2dSubsampling.forward for subsampling_factor 4 (convolution.py:227-237).

    inputs (B, W, H): the conv "height" axis is time (W px), the conv "width"
    axis is the image height.  Returns ((B, T, d), (B,) int32)."""
    out_lens = calc_length(input_lengths)
    x = inputs[:, None, :, :]
    x = _conv2d_3x3_s2_p1(x, w[prefix + 'conv.0.weight'], w[prefix + 'conv.0.bias'], depthwise=False)
    x = np.maximum(x, 0)                                                   # ReLU  :198
    x = _conv2d_3x3_s2_p1(x, w[prefix + 'conv.2.weight'], w[prefix + 'conv.2.bias'], depthwise=True)
    pw = w[prefix + 'conv.3.weight'][:, :, 0, 0]                          # 1x1   :207-212
    x = np.einsum('bcth,oc->both', x, pw) + w[prefix + 'conv.3.bias'][None, :, None, None]
    x = np.maximum(x, 0)                                                   # ReLU  :213
    b, c, t, f = x.shape
    x = x.transpose(0, 2, 1, 3).reshape(b, t, c * f)                       # :236 feature = c*f_dim + f
    x = linear(x, w[prefix + 'out.0.weight'], w[prefix + 'out.0.bias'])    # :224
    return x, out_lens


# --------------------------------------------------------------------------- #
# relative positional encoding (embedding.py:35-67)
# --------------------------------------------------------------------------- #
def rel_pos_table(T: int, d_model: int, dtype=np.float32) -> Array:
    """Rows k = 0..2T-2 hold the sinusoid of relative offset r = T-1-k
    (positive offsets first, flipped; then negative) (embedding.py:42-56,66)."""
    r = (T - 1 - np.arange(2 * T - 1)).astype(np.float32)[:, None]
    div = np.exp(np.arange(0, d_model, 2, dtype=np.float32) * np.float32(-(math.log(10000.0) / d_model)))
    pe = np.zeros((2 * T - 1, d_model), dtype=np.float32)
    pe[:, 0::2] = np.sin(r * div)
    pe[:, 1::2] = np.cos(r * div)
    return pe.astype(dtype)


def relative_shift_reference_form(pos_score: Array) -> Array:
    """The pad-one-column / reinterpret / slice idiom of attention.py:105-113,
    restated with numpy reshape.  Kept to prove `relative_shift` below equal."""
    b, h, t1, t2 = pos_score.shape
    padded = np.concatenate([np.zeros((b, h, t1, 1), dtype=pos_score.dtype), pos_score], axis=-1)
    padded = padded.reshape(b, h, t2 + 1, t1)
    return padded[:, :, 1:].reshape(b, h, t1, t2)[:, :, :, : t2 // 2 + 1]


def relative_shift(pos_score: Array) -> Array:
    """shift(M)[i, j] = M[i, T-1-i+j] for j < T (closed form of attention.py:105-113)."""
    b, h, T, t2 = pos_score.shape
    assert t2 == 2 * T - 1
    i = np.arange(T)[:, None]
    j = np.arange(T)[None, :]
    return pos_score[:, :, i, T - 1 - i + j]


# --------------------------------------------------------------------------- #
# modules of one conformer block
# --------------------------------------------------------------------------- #
def feed_forward(x: Array, w: Dict[str, Array], p: str) -> Array:
    """FeedForwardModule (feed_forward.py:45-55); dropout is identity in eval."""
    y = layer_norm(x, w[p + 'sequential.0.weight'], w[p + 'sequential.0.bias'])
    y = linear(y, w[p + 'sequential.1.linear.weight'], w[p + 'sequential.1.linear.bias'])
    y = silu(y)
    return linear(y, w[p + 'sequential.4.linear.weight'], w[p + 'sequential.4.linear.bias'])


def rel_mhsa(x: Array, w: Dict[str, Array], p: str, num_heads: int, mask: Array | None = None) -> Array:
    """MultiHeadedSelfAttentionModule + RelativeMultiHeadAttention
    (attention.py:143-151, 72-103).  `mask` (B, 1, T) bool, True = masked; the
    reference never passes one (encoder.py:67-103 is an nn.Sequential)."""
    B, T, d = x.shape
    dh = d // num_heads
    pe = rel_pos_table(T, d, x.dtype)                                               # :145-146
    xn = layer_norm(x, w[p + 'layer_norm.weight'], w[p + 'layer_norm.bias'])        # :148
    a = p + 'attention.'
    q = linear(xn, w[a + 'query_proj.linear.weight'], w[a + 'query_proj.linear.bias']).reshape(B, T, num_heads, dh)
    k = linear(xn, w[a + 'key_proj.linear.weight'], w[a + 'key_proj.linear.bias']).reshape(B, T, num_heads, dh)
    v = linear(xn, w[a + 'value_proj.linear.weight'], w[a + 'value_proj.linear.bias']).reshape(B, T, num_heads, dh)
    pos = linear(pe, w[a + 'pos_proj.linear.weight']).reshape(2 * T - 1, num_heads, dh)  # :85 (no bias)
    qu = (q + w[a + 'u_bias']).transpose(0, 2, 1, 3)                                # (B,h,T,dh)  :87
    qv = (q + w[a + 'v_bias']).transpose(0, 2, 1, 3)                                #             :88
    content = qu @ k.transpose(0, 2, 3, 1)                                          # (B,h,T,T)
    pos_score = qv @ pos.transpose(1, 2, 0)[None]                                   # (B,h,T,2T-1)
    pos_score = relative_shift(pos_score)                                           # :89
    score = (content + pos_score) / np.asarray(math.sqrt(dh), dtype=x.dtype)        # :91
    if mask is not None:
        score = np.where(mask[:, None], np.asarray(-1e9, dtype=x.dtype), score)     # :93-95
    score = score - score.max(axis=-1, keepdims=True)
    e = np.exp(score)
    attn = e / e.sum(axis=-1, keepdims=True)                                        # :97
    ctx = (attn @ v.transpose(0, 2, 1, 3)).transpose(0, 2, 1, 3).reshape(B, T, d)   # :100-101 head-major
    return linear(ctx, w[a + 'out_proj.linear.weight'], w[a + 'out_proj.linear.bias'])


def conv_module(x: Array, w: Dict[str, Array], p: str, kernel_size: int) -> Array:
    """ConformerConvModule in eval mode (convolution.py:135-148)."""
    B, T, d = x.shape
    y = layer_norm(x, w[p + 'sequential.0.weight'], w[p + 'sequential.0.bias'])
    pw1 = w[p + 'sequential.2.conv.weight'][:, :, 0]                                # (2d, d)
    y = linear(y, pw1, w[p + 'sequential.2.conv.bias'])                             # (B,T,2d)
    y = y[..., :d] * _sigmoid(y[..., d:])                                           # GLU(dim=channels) :139
    pad = (kernel_size - 1) // 2
    yp = np.zeros((B, T + 2 * pad, d), dtype=x.dtype)
    yp[:, pad:pad + T] = y
    dw = w[p + 'sequential.4.conv.weight'][:, 0, :]                                 # (d, K), no bias :140
    z = np.zeros_like(y)
    for kk in range(kernel_size):
        z += yp[:, kk:kk + T] * dw[None, None, :, kk]
    bn = p + 'sequential.5.'
    z = (z - w[bn + 'running_mean']) / np.sqrt(w[bn + 'running_var'] + np.asarray(1e-5, dtype=x.dtype))
    z = z * w[bn + 'weight'] + w[bn + 'bias']                                       # BatchNorm1d eval :141
    z = silu(z)                                                                     # :142
    pw2 = w[p + 'sequential.7.conv.weight'][:, :, 0]
    return linear(z, pw2, w[p + 'sequential.7.conv.bias'])                          # :143


def conformer_block(x: Array, w: Dict[str, Array], p: str, num_heads: int, kernel_size: int,
                    half_step_residual: bool = True, mask: Array | None = None) -> Array:
    """ConformerBlock (encoder.py:67-103) with ResidualConnectionModule
    (modules.py:31-32): out = module(x) * module_factor + x."""
    f = 0.5 if half_step_residual else 1.0
    f = np.asarray(f, dtype=x.dtype)
    x = feed_forward(x, w, p + 'sequential.0.module.') * f + x
    x = rel_mhsa(x, w, p + 'sequential.1.module.', num_heads, mask) + x
    x = conv_module(x, w, p + 'sequential.2.module.', kernel_size) + x
    x = feed_forward(x, w, p + 'sequential.3.module.') * f + x
    return layer_norm(x, w[p + 'sequential.4.weight'], w[p + 'sequential.4.bias'])


# --------------------------------------------------------------------------- #
# encoder / decoder / full forward
# --------------------------------------------------------------------------- #
def encoder_forward(inputs: Array, input_lengths: Array, w: Dict[str, Array], hp: Dict,
                    prefix: str = '', masked: bool = False) -> Tuple[Array, Array]:
    """ConformerEncoder.forward (encoder.py:183-204).  inputs (B, W, H).

    masked=False is the reference behaviour (no attention mask is ever applied);
    masked=True masks keys beyond each line's output length (ragged mode)."""
    x, out_lens = conv2d_subsampling(inputs, input_lengths, w, prefix + 'conv_subsample.')
    mask = None
    if masked:
        T = x.shape[1]
        mask = (np.arange(T)[None, :] >= out_lens[:, None])[:, None, :]
    for i in range(hp['num_encoder_layers']):
        x = conformer_block(x, w, f'{prefix}layers.{i}.', hp['num_attention_heads'],
                            hp['conv_kernel_size'], hp.get('half_step_residual', True), mask)
    return x, out_lens


def recognition_forward(line: Array, lens: Array, w: Dict[str, Array], hp: Dict,
                        masked: bool = False) -> Tuple[Array, Array]:
    """PytorchRecognitionModel.forward (pred.py:118-122): line (N,1,H,W) ->
    squeeze(1).transpose(1,2) -> encoder -> decoder Linear.  Keys carry the
    ModuleDict names `encoder.` / `decoder.` (pred.py:91)."""
    x = np.ascontiguousarray(line[:, 0].transpose(0, 2, 1))
    enc, enc_lens = encoder_forward(x, lens, w, hp, prefix='encoder.', masked=masked)
    logits = linear(enc, w['decoder.weight'], w['decoder.bias'])
    return logits, enc_lens


def log_softmax(x: Array) -> Array:
    """model.py:136 (training only)."""
    m = x.max(axis=-1, keepdims=True)
    return x - m - np.log(np.exp(x - m).sum(axis=-1, keepdims=True))


# --------------------------------------------------------------------------- #
# greedy CTC decode (kraken.lib.ctc_decoder.greedy_decoder; call sites pred.py:143,162,177)
# --------------------------------------------------------------------------- #
def greedy_decoder(outputs: Array) -> List[Tuple[int, int, int, float]]:
    """outputs (C, T).  argmax over classes (first maximum wins ties), merge
    consecutive equal labels, drop label 0 (blank), emit
    (label, first_frame, last_frame_inclusive, max winning score inside the run).
    The 4th column is "parity unpinned" (see module docstring)."""
    C, T = outputs.shape
    if T == 0:
        return []
    labels = outputs.argmax(axis=0)
    best = outputs.max(axis=0)
    res: List[Tuple[int, int, int, float]] = []
    start = 0
    for t in range(1, T + 1):
        if t == T or labels[t] != labels[start]:
            lab = int(labels[start])
            if lab != 0:
                res.append((lab, start, t - 1, float(best[start:t].max())))
            start = t
    return res


def predict_labels(line: Array, lens: Array, w: Dict[str, Array], hp: Dict,
                   masked: bool = False) -> List[List[Tuple[int, int, int, float]]]:
    """PytorchRecognitionModel.predict_labels (pred.py:172-178)."""
    logits, olens = recognition_forward(line, lens, w, hp, masked)
    o = logits.transpose(0, 2, 1).astype(np.float32)
    return [greedy_decoder(seq[:, :int(n)]) for seq, n in zip(o, olens)]


# --------------------------------------------------------------------------- #
# CTC loss (model.py:119,139-142: nn.CTCLoss(reduction='sum', zero_infinity=True), blank 0)
# --------------------------------------------------------------------------- #
def ctc_loss_sum(log_probs: Array, targets: Array, input_lengths: Sequence[int],
                 target_lengths: Sequence[int]) -> float:
    """log_probs (T, B, C) log-softmaxed; targets 1-D concatenation.  Standard
    alpha recursion in float64; infinite losses are zeroed (zero_infinity)."""
    total = 0.0
    off = 0
    lp = log_probs.astype(np.float64)
    for b, (Tn, Ln) in enumerate(zip(input_lengths, target_lengths)):
        Tn, Ln = int(Tn), int(Ln)
        tgt = targets[off:off + Ln]
        off += Ln
        ext = np.zeros(2 * Ln + 1, dtype=np.int64)
        ext[1::2] = tgt
        S = ext.shape[0]
        alpha = np.full(S, -np.inf)
        if Tn == 0:
            loss = 0.0 if Ln == 0 else np.inf
        else:
            alpha[0] = lp[0, b, 0]
            if S > 1:
                alpha[1] = lp[0, b, ext[1]]
            for t in range(1, Tn):
                prev = alpha
                a1 = np.concatenate([[-np.inf], prev[:-1]])[:S]
                a2 = np.concatenate([[-np.inf, -np.inf], prev[:-2]])[:S]
                skip_ok = np.zeros(S, dtype=bool)
                skip_ok[2:] = (ext[2:] != 0) & (ext[2:] != ext[:-2])
                a2 = np.where(skip_ok, a2, -np.inf)
                m = np.maximum(np.maximum(prev, a1), a2)
                msafe = np.where(np.isfinite(m), m, 0.0)
                s = np.exp(prev - msafe) + np.exp(a1 - msafe) + np.exp(a2 - msafe)
                with np.errstate(divide='ignore'):
                    alpha = np.where(np.isfinite(m), msafe + np.log(s), -np.inf) + lp[t, b, ext]
            tail = alpha[-1] if S == 1 else np.logaddexp(alpha[-1], alpha[-2])
            loss = -tail
        if not np.isfinite(loss):
            loss = 0.0
        total += float(loss)
    return total


# --------------------------------------------------------------------------- #
# deterministic synthetic weights / inputs shared by tests, bench and smoke
# --------------------------------------------------------------------------- #
def weight_spec(hp: Dict, num_classes: int) -> List[Tuple[str, Tuple[int, ...]]]:
    """(state-dict key, shape) for every tensor of `PytorchRecognitionModel.nn`
    (pred.py:75-92): `encoder.*` + `decoder.{weight,bias}`."""
    d = hp['encoder_dim']
    ff = d * hp['feed_forward_expansion_factor']
    cc = hp['subsampling_conv_channels']
    K = hp['conv_kernel_size']
    h = hp['num_attention_heads']
    feat = cc * int(calc_length(np.array(hp['height'])))
    spec: List[Tuple[str, Tuple[int, ...]]] = [
        ('encoder.conv_subsample.conv.0.weight', (cc, 1, 3, 3)), ('encoder.conv_subsample.conv.0.bias', (cc,)),
        ('encoder.conv_subsample.conv.2.weight', (cc, 1, 3, 3)), ('encoder.conv_subsample.conv.2.bias', (cc,)),
        ('encoder.conv_subsample.conv.3.weight', (cc, cc, 1, 1)), ('encoder.conv_subsample.conv.3.bias', (cc,)),
        ('encoder.conv_subsample.out.0.weight', (d, feat)), ('encoder.conv_subsample.out.0.bias', (d,)),
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


def make_weights(hp: Dict, num_classes: int, seed: int = 0) -> Dict[str, Array]:
    """Seeded fp32 weights with non-trivial biases / BN statistics so that the
    folded paths are exercised.  Matrix weights are scaled
    uniform(-a, a) with the xavier/kaiming magnitude torch would give them; the
    exact distribution is irrelevant to parity, only determinism matters."""
    rng = np.random.RandomState(seed)
    out: Dict[str, Array] = {}
    for key, shape in weight_spec(hp, num_classes):
        leaf = key.rsplit('.', 1)[-1]
        if leaf == 'running_var':
            v = rng.uniform(0.5, 1.5, size=shape)
        elif leaf == 'running_mean':
            v = rng.normal(0.0, 0.1, size=shape)
        elif leaf in ('u_bias', 'v_bias'):
            a = math.sqrt(6.0 / (shape[0] + shape[1]))
            v = rng.uniform(-a, a, size=shape)
        elif len(shape) == 1:
            # 1-D 'weight' = LayerNorm/BatchNorm scale ~ N(1, .05); every 1-D 'bias' ~ N(0, .05)
            v = rng.normal(1.0 if leaf == 'weight' else 0.0, 0.05, size=shape)
        else:
            fan_out = shape[0] * int(np.prod(shape[2:])) if len(shape) > 2 else shape[0]
            fan_in = int(np.prod(shape[1:]))
            a = math.sqrt(6.0 / (fan_in + fan_out))
            v = rng.uniform(-a, a, size=shape)
        out[key] = np.ascontiguousarray(v.astype(np.float32))
    return out


def make_lines(batch: int, height: int, width: int, seed: int = 1234,
               widths: Sequence[int] | None = None) -> Tuple[Array, Array]:
    """Synthetic text-line batch in the collate layout:
    (N,1,H,W) fp32 in [0,1), right-zero-padded, plus the true widths (int64)."""
    rng = np.random.RandomState(seed)
    img = rng.uniform(0.0, 1.0, size=(batch, 1, height, width)).astype(np.float32)
    if widths is None:
        lens = np.full((batch,), width, dtype=np.int64)
    else:
        lens = np.asarray(widths, dtype=np.int64)
        for b, wv in enumerate(lens):
            img[b, :, :, int(wv):] = 0.0
    return img, lens
