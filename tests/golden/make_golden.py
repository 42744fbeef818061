"""Generate golden vectors by running the UNMODIFIED reference modules.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

It imports `conformer_ocr.conformer.encoder.ConformerEncoder` from /root/reference
(torch-only, importable), loads the deterministic numpy weights from
`oracle.conformer_oracle.make_weights` into it by state-dict key, runs the forward of
`PytorchRecognitionModel.forward` (pred.py:118-122: squeeze/transpose -> encoder ->
nn.Linear decoder) on CPU fp32 and stores logits + output lengths (+ per-module
intermediates for the small case).  Inputs and weights are NOT stored: they are
re-generated from the seeds recorded in each file.  `conformer_ocr.pred` itself cannot be
imported (kraken/lightning missing and a SyntaxError at pred.py:213-217), so the three
wrapper lines are restated here.
"""
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

CASES = {
    # name: (hparams overrides, num_classes, batch, width, widths(ragged) or None, weight seed, input seed)
    'tiny_equal': (dict(num_encoder_layers=2, encoder_dim=32, num_attention_heads=2, height=16,
                        subsampling_conv_channels=4, conv_kernel_size=7), 11, 3, 40, None, 1, 2),
    'tiny_ragged': (dict(num_encoder_layers=2, encoder_dim=32, num_attention_heads=2, height=16,
                         subsampling_conv_channels=4, conv_kernel_size=7), 11, 4, 53, [53, 41, 22, 5], 3, 4),
    'default_equal': (dict(), 64, 2, 128, None, 5, 6),
    'default_ragged': (dict(), 40, 3, 101, [101, 77, 30], 7, 8),
    'default_2layer_long': (dict(num_encoder_layers=2), 300, 1, 1025, None, 9, 10),
}


def build_reference(hp, num_classes, weights):
    enc = ConformerEncoder(in_channels=1, input_dim=hp['height'], encoder_dim=hp['encoder_dim'],
                           num_layers=hp['num_encoder_layers'], num_attention_heads=hp['num_attention_heads'],
                           feed_forward_expansion_factor=hp['feed_forward_expansion_factor'],
                           conv_expansion_factor=hp['conv_expansion_factor'], input_dropout_p=0.1,
                           feed_forward_dropout_p=0.1, attention_dropout_p=0.1, conv_dropout_p=0.1,
                           conv_kernel_size=hp['conv_kernel_size'], half_step_residual=hp['half_step_residual'],
                           subsampling_conv_channels=hp['subsampling_conv_channels'],
                           subsampling_factor=hp['subsampling_factor'])
    dec = torch.nn.Linear(hp['encoder_dim'], num_classes, bias=True)
    net = torch.nn.ModuleDict({'encoder': enc, 'decoder': dec})
    sd = {k: torch.from_numpy(v.copy()) for k, v in weights.items()}
    missing, unexpected = net.load_state_dict(sd, strict=False)
    missing = [m for m in missing if not m.endswith('num_batches_tracked')]
    assert not missing and not unexpected, (missing, unexpected)
    return net.eval()


def main():
    torch.manual_seed(0)
    torch.set_num_threads(4)
    index = {}
    for name, (over, C, B, W, widths, wseed, iseed) in CASES.items():
        hp = dict(O.DEFAULT_HPARAMS)
        hp.update(over)
        weights = O.make_weights(hp, C, seed=wseed)
        net = build_reference(hp, C, weights)
        img, lens = O.make_lines(B, hp['height'], W, seed=iseed, widths=widths)
        inter = {}
        if name.startswith('tiny'):
            blk = net['encoder'].layers[0].sequential
            hooks = [net['encoder'].conv_subsample.register_forward_hook(lambda m, i, o: inter.__setitem__('subsample', o[0].numpy().copy()))]
            for idx, nm in ((0, 'l0_ffn1'), (1, 'l0_mhsa'), (2, 'l0_conv'), (3, 'l0_ffn2'), (4, 'l0_out')):
                hooks.append(blk[idx].register_forward_hook(lambda m, i, o, nm=nm: inter.__setitem__(nm, o.numpy().copy())))
        with torch.no_grad():
            line = torch.from_numpy(img).squeeze(1).transpose(1, 2)               # pred.py:119
            enc_out, enc_lens = net['encoder'](line, torch.from_numpy(lens))      # pred.py:120
            logits = net['decoder'](enc_out)                                      # pred.py:121
        out = dict(logits=logits.numpy().astype(np.float32), enc_lens=enc_lens.numpy().astype(np.int32),
                   meta=np.frombuffer(json.dumps(dict(hparams=hp, num_classes=C, batch=B, width=W, widths=widths,
                                                      weight_seed=wseed, input_seed=iseed,
                                                      torch=torch.__version__)).encode(), dtype=np.uint8))
        out.update({'inter_' + k: v for k, v in inter.items()})
        np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
        index[name] = dict(logits_shape=list(logits.shape), enc_lens=enc_lens.tolist())
        print(name, logits.shape, enc_lens.tolist(), float(logits.abs().max()))
    with open(os.path.join(HERE, 'index.json'), 'w') as fp:
        json.dump(index, fp, indent=1)


if __name__ == '__main__':
    main()
