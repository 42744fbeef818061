"""CPU: the checkers against the reference's OWN modules, live (oracle/_ref = conformer_ocr.conformer.* byte-compiled from the
reference tree by oracle/build_ref.py; loader oracle/ref_modules.py).

The committed goldens (tests/golden/*.npz) pin the oracle at five + three fixed cases; with the compiled modules at hand the
same comparison runs on FRESH seeded inputs, on shapes the goldens do not hold (odd widths, a one-frame line, the default
spec at full depth), and in train() mode.  Skipped where oracle/_ref is not built (it is git-ignored; `__graft_entry__.build()`
makes it wherever /root/reference exists, and it travels with the snapshot)."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN_CASES, golden_inputs, load_golden, rel_err
from oracle import conformer_oracle as O
from oracle import ref_modules as RM
from oracle import torch_port as TP
from oracle import torch_train_port as TTP

pytestmark = pytest.mark.skipif(not RM.available(), reason=RM.why_unavailable() or 'oracle/_ref unavailable')

TINY = dict(num_encoder_layers=2, encoder_dim=32, num_attention_heads=2, height=16, subsampling_conv_channels=4, conv_kernel_size=7)
CASES = {
    # name: (hparams overrides, classes, batch, width, widths, weight seed, input seed)
    'tiny_odd_widths': (TINY, 13, 5, 61, [61, 60, 33, 7, 1], 101, 102),
    'tiny_one_line': (TINY, 5, 1, 9, None, 103, 104),
    'default_3layer_ragged': (dict(num_encoder_layers=3), 97, 3, 257, [257, 130, 64], 105, 106),
    'default_full_depth': (dict(), 256, 2, 192, [192, 150], 107, 108),
}


def _case(name):
    over, C, B, W, widths, ws, iseed = CASES[name]
    hp = dict(O.DEFAULT_HPARAMS)
    hp.update(over)
    w = O.make_weights(hp, C, seed=ws)
    img, lens = O.make_lines(B, hp['height'], W, seed=iseed, widths=widths)
    return hp, C, w, img, lens


@pytest.mark.parametrize('name', sorted(CASES))
def test_numpy_oracle_and_torch_port_match_the_reference_modules(name):
    hp, C, w, img, lens = _case(name)
    net = RM.build_model(hp, C, w).eval()
    ref, ref_lens = RM.forward(net, torch.from_numpy(img), torch.from_numpy(lens))
    ref = ref.numpy()
    got, got_lens = O.recognition_forward(img, lens, w, hp)
    assert ref_lens.dtype == torch.int32 and got_lens.tolist() == ref_lens.tolist()
    assert got.shape == ref.shape and rel_err(got, ref) < 2e-5
    np.testing.assert_array_equal(got.argmax(-1), ref.argmax(-1))
    port, port_lens = TP.recognition_forward(torch.from_numpy(img), torch.from_numpy(lens), TP.to_torch(w), hp)
    assert port_lens.tolist() == ref_lens.tolist()
    assert rel_err(port.numpy(), ref) < 2e-5
    # pred.py:171-178 over the reference modules against the two restatements: identical (label, start, end) records
    recs = RM.predict_labels(net, torch.from_numpy(img), torch.from_numpy(lens))
    for other in (O.predict_labels(img, lens, w, hp), TP.predict_labels(torch.from_numpy(img), torch.from_numpy(lens), TP.to_torch(w), hp)):
        assert [[r[:3] for r in line] for line in other] == [[r[:3] for r in line] for line in recs]


@pytest.mark.parametrize('name', GOLDEN_CASES)
def test_committed_goldens_reproduce_from_the_reference_modules(name):
    """The fixtures under tests/golden/ are what the reference modules compute today (same seeds, same weights)."""
    meta, arrs = load_golden(name)
    hp, w, img, lens = golden_inputs(meta)
    net = RM.build_model(hp, meta['num_classes'], w).eval()
    ref, ref_lens = RM.forward(net, torch.from_numpy(img), torch.from_numpy(lens))
    np.testing.assert_array_equal(ref_lens.numpy(), arrs['enc_lens'])
    assert rel_err(ref.numpy(), arrs['logits']) < 1e-6          # 0.0 with the generating thread count; BLAS blocking otherwise


@pytest.mark.parametrize('name', ['tiny_odd_widths', 'default_3layer_ragged'])
def test_train_port_matches_the_reference_modules_in_train_mode(name):
    """model.py:129-145 over the reference modules in train() mode (dropout 0, BatchNorm batch statistics) against the
    float64 training port: loss, every parameter gradient, updated running statistics."""
    hp, C, w, img, lens = _case(name)
    B = img.shape[0]
    rng = np.random.RandomState(7)
    olens = O.calc_length(lens)
    tl = np.minimum(rng.randint(1, 6, size=B), np.maximum(olens // 2, 1)).astype(np.int64)
    tg = rng.randint(1, C, size=int(tl.sum())).astype(np.int64)
    batch = {'image': torch.from_numpy(img), 'seq_lens': torch.from_numpy(lens), 'target': torch.from_numpy(tg), 'target_lens': torch.from_numpy(tl)}
    net = RM.build_model(hp, C, w, dropout=0.0).train()
    out = RM.step_loss(net, batch)
    out['loss'].backward()
    ref_grads = {k: p.grad.numpy() for k, p in net.named_parameters()}
    wt = {k: torch.from_numpy(v.copy()).double().requires_grad_(not k.endswith(('running_mean', 'running_var'))) for k, v in w.items()}
    loss, probits, polens, stats = TTP.train_step_loss(torch.from_numpy(img).double(), batch['seq_lens'], batch['target'], batch['target_lens'], wt, hp)
    loss.backward()
    assert polens.tolist() == out['output_lens'].tolist()
    assert abs(loss.item() - out['loss'].item()) <= 1e-5 * abs(out['loss'].item())
    keys = TTP.trainable_keys(wt)
    assert sorted(keys) == sorted(ref_grads)
    gmax = max(np.abs(ref_grads[k]).max() for k in keys)
    for k in keys:
        assert np.abs(wt[k].grad.numpy() - ref_grads[k]).max() <= 1e-4 * max(np.abs(ref_grads[k]).max(), 1e-3 * gmax), k
    sd = net.state_dict()
    for k, v in stats.items():
        assert np.abs(v.numpy() - sd[k].numpy()).max() <= 1e-6, k


def test_every_module_of_every_layer_matches_the_reference_modules():
    """Rows A2-A8 one by one, fed with the REFERENCE's own input of each module (so an error cannot hide behind, or grow
    from, an earlier module): sub-sampling, then per layer the half-step FFN, rel-pos MHSA, conv module, second FFN and
    block-final LayerNorm of a 3-layer default-spec encoder on a ragged batch."""
    hp, C, w, img, lens = _case('default_3layer_ragged')
    net = RM.build_model(hp, C, w).eval()
    enc = net['encoder']
    taps = {}
    hooks = [enc.conv_subsample.register_forward_hook(lambda m, i, o: taps.__setitem__('sub', (i[0].numpy().copy(), o[0].numpy().copy())))]
    for l, layer in enumerate(enc.layers):
        for j in range(5):
            hooks.append(layer.sequential[j].register_forward_hook(
                lambda m, i, o, key=(l, j): taps.__setitem__(key, (i[0].numpy().copy(), o.numpy().copy()))))
    RM.forward(net, torch.from_numpy(img), torch.from_numpy(lens))
    for h in hooks:
        h.remove()
    x_in, sub_out = taps['sub']
    got, _ = O.conv2d_subsampling(x_in, lens, w, 'encoder.conv_subsample.')
    assert rel_err(got, sub_out) < 2e-5
    half = np.float32(0.5)
    for l in range(hp['num_encoder_layers']):
        p = f'encoder.layers.{l}.sequential.'
        x, y = taps[(l, 0)]
        assert rel_err(O.feed_forward(x, w, p + '0.module.') * half + x, y) < 2e-5, (l, 'ffn1')
        x, y = taps[(l, 1)]
        assert rel_err(O.rel_mhsa(x, w, p + '1.module.', hp['num_attention_heads']) + x, y) < 2e-5, (l, 'mhsa')
        x, y = taps[(l, 2)]
        assert rel_err(O.conv_module(x, w, p + '2.module.', hp['conv_kernel_size']) + x, y) < 2e-5, (l, 'conv')
        x, y = taps[(l, 3)]
        assert rel_err(O.feed_forward(x, w, p + '3.module.') * half + x, y) < 2e-5, (l, 'ffn2')
        x, y = taps[(l, 4)]
        assert rel_err(O.layer_norm(x, w[p + '4.weight'], w[p + '4.bias']), y) < 2e-5, (l, 'final LayerNorm')
