"""Golden vectors of one TRAINING step from the UNMODIFIED reference modules (build container only):

    python tests/golden/make_golden_train.py

`conformer_ocr.conformer.encoder.ConformerEncoder` (from /root/reference) in train() mode with every dropout probability 0
(the reference's masks come from torch's generator: not reproducible elsewhere), `nn.Linear` decoder, then the three lines of
`RecognitionModel._step` (model.py:136-142; the module itself needs lightning/kraken and cannot be imported):
log_softmax -> transpose -> CTCLoss(reduction='sum', zero_infinity=True); `loss.backward()`.
Stored per case: loss, probits, every parameter gradient, the updated BatchNorm running statistics, and the parameters after
one `torch.optim.AdamW(lr, weight_decay)` step (model.py:272-274).  Inputs/weights/targets are regenerated from the seeds."""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference')

from oracle import conformer_oracle as O  # noqa: E402
from conformer_ocr.conformer.encoder import ConformerEncoder  # noqa: E402  (the reference)

TINY = dict(num_encoder_layers=2, encoder_dim=32, num_attention_heads=2, height=16, subsampling_conv_channels=4,
            conv_kernel_size=7)
CASES = {
    # name: (hparams overrides, num_classes, batch, width, widths or None, weight seed, input seed, target seed)
    'train_tiny_equal': (TINY, 11, 3, 40, None, 21, 22, 23),
    'train_tiny_ragged': (TINY, 11, 4, 53, [53, 41, 22, 9], 24, 25, 26),
    'train_default_1layer': (dict(num_encoder_layers=1), 24, 2, 101, [101, 77], 27, 28, 29),
}
LR, WD = 1e-3, 1e-3       # default_specs.py lr / model.py:50 weight_decay


def make_targets(batch, out_lens, num_classes, seed):
    """Targets a CTC path exists for: length U{1..max(1, T_i // 3)}, labels U{1..C-1}."""
    rng = np.random.RandomState(seed)
    tl = np.array([rng.randint(1, max(2, int(n) // 3 + 1)) for n in out_lens], dtype=np.int64)
    tg = rng.randint(1, num_classes, size=int(tl.sum())).astype(np.int64)
    return tg, tl


def run_case(name):
    ov, C, B, W, widths, ws, is_, ts = CASES[name]
    hp = dict(O.DEFAULT_HPARAMS)
    hp.update(ov)
    w = O.make_weights(hp, C, seed=ws)
    img, lens = O.make_lines(B, hp['height'], W, seed=is_, widths=widths)
    enc = ConformerEncoder(in_channels=1, input_dim=hp['height'], encoder_dim=hp['encoder_dim'],
                           num_layers=hp['num_encoder_layers'], num_attention_heads=hp['num_attention_heads'],
                           feed_forward_expansion_factor=hp['feed_forward_expansion_factor'],
                           conv_expansion_factor=hp['conv_expansion_factor'], input_dropout_p=0.0,
                           feed_forward_dropout_p=0.0, attention_dropout_p=0.0, conv_dropout_p=0.0,
                           conv_kernel_size=hp['conv_kernel_size'], half_step_residual=hp['half_step_residual'],
                           subsampling_conv_channels=hp['subsampling_conv_channels'],
                           subsampling_factor=hp['subsampling_factor'])
    net = torch.nn.ModuleDict({'encoder': enc, 'decoder': torch.nn.Linear(hp['encoder_dim'], C, bias=True)})
    missing, unexpected = net.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in w.items()}, strict=False)
    assert not unexpected and all(m.endswith('num_batches_tracked') for m in missing), (missing, unexpected)
    net.train()
    opt = torch.optim.AdamW(net.parameters(), lr=LR, weight_decay=WD)
    x = torch.from_numpy(img).squeeze(1).transpose(1, 2)                     # model.py:132
    out, olens = net['encoder'](x, torch.from_numpy(lens))
    probits = net['decoder'](out)
    tg, tl = make_targets(B, olens.tolist(), C, ts)
    logp = torch.nn.functional.log_softmax(probits, dim=-1)                   # model.py:136
    crit = torch.nn.CTCLoss(reduction='sum', zero_infinity=True)              # model.py:119
    loss = crit(logp.transpose(0, 1), torch.from_numpy(tg), olens, torch.from_numpy(tl))
    loss.backward()
    rec = {'loss': np.float64(loss.item()), 'probits': probits.detach().numpy(), 'out_lens': olens.numpy(),
           'targets': tg, 'target_lens': tl}
    for k, p in net.named_parameters():
        rec['grad/' + k] = p.grad.detach().numpy().copy()
    opt.step()
    sd = net.state_dict()
    for k, v in sd.items():
        if k.endswith(('running_mean', 'running_var')):
            rec['stat/' + k] = v.numpy().copy()
    for k, p in net.named_parameters():
        rec['adamw/' + k] = p.detach().numpy().copy()
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **rec)
    meta = dict(hparams=hp, num_classes=C, batch=B, width=W, widths=widths, weight_seed=ws, input_seed=is_, target_seed=ts,
                lr=LR, weight_decay=WD, loss=float(loss.item()))
    print(name, 'loss', float(loss.item()), 'T', olens.tolist(), 'target lens', tl.tolist())
    return meta


if __name__ == '__main__':
    torch.manual_seed(0)
    torch.set_num_threads(4)
    idx = {n: run_case(n) for n in CASES}
    with open(os.path.join(HERE, 'index_train.json'), 'w') as f:
        json.dump(idx, f, indent=1, sort_keys=True)
