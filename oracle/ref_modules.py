"""TEST INFRASTRUCTURE (checker / reported baseline only; never imported by conformer_ocr_b200/).

Loader of `oracle/_ref/`: the reference's own `conformer_ocr.conformer` modules, compiled from /root/reference by
`oracle/build_ref.py` into marshalled code objects (compiled outputs only; nothing here reads /root/reference at run time).  Used
  * by `tests/test_ref_modules.py` to pin the numpy oracle and the torch ports against the UNMODIFIED reference modules
    on fresh seeded inputs (the committed goldens under tests/golden/ were made from the same modules), and
  * by `bench.py --impl reference` / `cpu_baseline` as the reference's CPU implementation of the path (`kind: "reference"`);
    without `oracle/_ref` both fall back to the ports (`kind: "port"`).

What is the reference's and what is restated: encoder, sub-sampling, attention, convolution and feed-forward modules, their
initialisation and their train()/eval() behaviour are the reference's compiled code.  `conformer_ocr/pred.py` and `model.py`
cannot be imported (kraken / lightning absent, SyntaxError at pred.py:213-217), so the lines that drive the modules are
restated below, each citing its source; `greedy_decoder` is kraken's (absent): `oracle.conformer_oracle.greedy_decoder`
(PARITY UNPINNED for the confidence column).
"""
from __future__ import annotations

import importlib
import importlib.abc
import importlib.machinery
import importlib.util
import marshal
import os
import sys
from typing import Dict, List, Mapping, Optional, Tuple

import numpy as np
import torch

from . import conformer_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, '_ref')
_ENC = None
_WHY = ''


class _RefFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """Serves `conformer_ocr`, `conformer_ocr.conformer` and `conformer_ocr.conformer.*` from oracle/_ref/<name>.mcode."""
    PACKAGES = ('conformer_ocr', 'conformer_ocr.conformer')

    def find_spec(self, fullname, path=None, target=None):
        if fullname != 'conformer_ocr' and not fullname.startswith('conformer_ocr.'):
            return None
        origin = os.path.join(REF_DIR, fullname + '.mcode')
        if not os.path.isfile(origin):
            return None
        return importlib.machinery.ModuleSpec(fullname, self, origin=origin, is_package=fullname in self.PACKAGES)

    def create_module(self, spec):
        return None

    def exec_module(self, module):
        origin = module.__spec__.origin
        with open(origin, 'rb') as fp:
            if fp.read(len(importlib.util.MAGIC_NUMBER)) != importlib.util.MAGIC_NUMBER:
                raise ImportError(f'{origin} was compiled by another interpreter version')
            code = marshal.loads(fp.read())
        module.__file__ = origin
        exec(code, module.__dict__)


def _load():
    global _ENC, _WHY
    if _ENC is not None or _WHY:
        return _ENC
    if not os.path.isfile(os.path.join(REF_DIR, 'conformer_ocr.conformer.encoder.mcode')):
        _WHY = 'oracle/_ref is not built (python oracle/build_ref.py in the build container)'
        return None
    finder = _RefFinder()
    sys.meta_path.insert(0, finder)
    try:
        mod = importlib.import_module('conformer_ocr.conformer.encoder')
        where = os.path.realpath(getattr(mod, '__file__', '') or '')
        if not where.startswith(os.path.realpath(REF_DIR) + os.sep) and not where.startswith('/root/reference/'):
            raise ImportError(f'conformer_ocr resolved to {where}, not to oracle/_ref')
        _ENC = mod
    except Exception as e:                       # stale code objects (other interpreter) or a shadowing package: use the ports
        _WHY = f'oracle/_ref cannot be imported: {type(e).__name__}: {e}'
    finally:
        sys.meta_path.remove(finder)
    return _ENC


def available() -> bool:
    return _load() is not None


def why_unavailable() -> str:
    _load()
    return _WHY


def build_model(hp: Mapping, num_classes: int, weights: Optional[Mapping[str, np.ndarray]] = None,
                dropout: float = 0.1) -> torch.nn.ModuleDict:
    """`nn.ModuleDict({'encoder': ConformerEncoder(...), 'decoder': nn.Linear(encoder_dim, num_classes)})`
    as pred.py:75-90 / model.py:99-115 build it; `weights` (state-dict keys 'encoder.*', 'decoder.*') are loaded by key."""
    enc_mod = _load()
    if enc_mod is None:
        raise RuntimeError(_WHY)
    enc = enc_mod.ConformerEncoder(in_channels=1, input_dim=hp['height'], encoder_dim=hp['encoder_dim'],
                                   num_layers=hp['num_encoder_layers'], num_attention_heads=hp['num_attention_heads'],
                                   feed_forward_expansion_factor=hp['feed_forward_expansion_factor'],
                                   conv_expansion_factor=hp['conv_expansion_factor'], input_dropout_p=dropout,
                                   feed_forward_dropout_p=dropout, attention_dropout_p=dropout, conv_dropout_p=dropout,
                                   conv_kernel_size=hp['conv_kernel_size'], half_step_residual=hp['half_step_residual'],
                                   subsampling_conv_channels=hp['subsampling_conv_channels'],
                                   subsampling_factor=hp['subsampling_factor'])
    dec = torch.nn.Linear(hp['encoder_dim'], num_classes, bias=True)
    net = torch.nn.ModuleDict({'encoder': enc, 'decoder': dec})
    if weights is not None:
        sd = {k: torch.as_tensor(np.array(v, copy=True)) for k, v in weights.items()}
        missing, unexpected = net.load_state_dict(sd, strict=False)
        missing = [m for m in missing if not m.endswith('num_batches_tracked')]
        if missing or unexpected:
            raise ValueError(f'state dict does not fit the reference modules: missing {missing[:3]}, unexpected {unexpected[:3]}')
    return net


def forward(net: torch.nn.ModuleDict, line: torch.Tensor, lens: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """PytorchRecognitionModel.forward, pred.py:116-122."""
    with torch.no_grad():
        line = line.squeeze(1).transpose(1, 2)
        encoder_outputs, encoder_lens = net['encoder'](line, lens)
        probits = net['decoder'](encoder_outputs)
    return probits, encoder_lens


def predict_labels(net: torch.nn.ModuleDict, line: torch.Tensor, lens: Optional[torch.Tensor] = None) -> List[List[Tuple[int, int, int, float]]]:
    """PytorchRecognitionModel.predict_labels, pred.py:171-178 (ctc_decoder = kraken's greedy_decoder, restated)."""
    o, olens = forward(net, line, lens)
    o = o.transpose(1, 2).cpu().float().numpy()
    oseqs = []
    for seq, seq_len in zip(o, olens):
        oseqs.append(O.greedy_decoder(seq[:, :int(seq_len)]))
    return oseqs


def step_loss(net: torch.nn.ModuleDict, batch: Mapping[str, torch.Tensor], criterion=None) -> Dict[str, torch.Tensor]:
    """RecognitionModel._step, model.py:129-145 (criterion: model.py:119)."""
    if criterion is None:
        criterion = torch.nn.CTCLoss(reduction='sum', zero_infinity=True)
    input, target = batch['image'], batch['target']
    input = input.squeeze(1).transpose(1, 2)
    seq_lens, label_lens = batch['seq_lens'], batch['target_lens']
    encoder_outputs, encoder_lens = net['encoder'](input, seq_lens)
    probits = net['decoder'](encoder_outputs)
    logits = torch.nn.functional.log_softmax(probits, dim=-1)
    loss = criterion(logits.transpose(0, 1), target, encoder_lens, label_lens)
    return {'loss': loss, 'probits': probits, 'output_lens': encoder_lens}


def train_step_factory(net: torch.nn.ModuleDict, batch: Mapping[str, torch.Tensor], lr: float = 1e-3, weight_decay: float = 1e-3):
    """One optimiser step the way Lightning drives the reference: `training_step` (model.py:153-158) -> backward ->
    `optimizer.step` of `torch.optim.AdamW(lr, weight_decay)` (model.py:272-274 with the default_specs optimiser)."""
    net.train()
    opt = torch.optim.AdamW(net.parameters(), lr=lr, weight_decay=weight_decay)
    criterion = torch.nn.CTCLoss(reduction='sum', zero_infinity=True)

    def one_step() -> float:
        opt.zero_grad(set_to_none=True)
        loss = step_loss(net, batch, criterion)['loss']
        loss.backward()
        opt.step()
        return float(loss.item())
    return one_step
