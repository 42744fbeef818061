"""Parity of the fp32 CUDA path (through the C ABI) with the reference's golden vectors and the oracle.

Tolerance (BASELINE.json north_star): fp32 logits max-norm relative error <= 1e-4; collapsed label
sequences bit-exact given identical argmax."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN_CASES, golden_inputs, load_golden, rel_err
from oracle import conformer_oracle as O

pytestmark = pytest.mark.gpu
FP32_TOL = 1e-4
DROP = dict(input_dropout_p=0.1, feed_forward_dropout_p=0.1, attention_dropout_p=0.1, conv_dropout_p=0.1)


def build_net(hp, C, w, device, **kw):
    from conformer_ocr_b200 import PytorchRecognitionModel
    net = PytorchRecognitionModel(num_classes=C, codec=None, **hp, **DROP, **kw)
    net.nn.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()}, strict=False)
    return net.to(device)


@pytest.mark.parametrize('name', GOLDEN_CASES)
def test_logits_match_reference_golden(name, cuda_device):
    meta, arrs = load_golden(name)
    hp, w, img, lens = golden_inputs(meta)
    net = build_net(hp, meta['num_classes'], w, cuda_device)
    logits, olens = net(torch.from_numpy(img).to(cuda_device), torch.from_numpy(lens))
    assert logits.dtype == torch.float32 and logits.is_cuda
    assert olens.dtype == torch.int32
    np.testing.assert_array_equal(olens.cpu().numpy(), arrs['enc_lens'])
    got = logits.cpu().numpy()
    assert got.shape == arrs['logits'].shape
    assert rel_err(got, arrs['logits']) < FP32_TOL
    # top-1 agreement on every frame of every line
    np.testing.assert_array_equal(got.argmax(-1), arrs['logits'].argmax(-1))


@pytest.mark.parametrize('name', GOLDEN_CASES)
def test_decoded_labels_bit_exact(name, cuda_device):
    meta, arrs = load_golden(name)
    hp, w, img, lens = golden_inputs(meta)
    net = build_net(hp, meta['num_classes'], w, cuda_device)
    got = net.predict_labels(torch.from_numpy(img).to(cuda_device), torch.from_numpy(lens))
    ref = [O.greedy_decoder(arrs['logits'][b, :arrs['enc_lens'][b]].T) for b in range(meta['batch'])]
    assert [[t[:3] for t in l] for l in got] == [[t[:3] for t in l] for l in ref]
    # confidence column: parity unpinned upstream (kraken absent); compared with tolerance, not gated bit-exact
    for g, r in zip(got, ref):
        np.testing.assert_allclose([t[3] for t in g], [t[3] for t in r], rtol=1e-3, atol=1e-3)


def test_encoder_module_and_input_layouts(cuda_device):
    from conformer_ocr_b200 import ConformerEncoder
    meta, _ = load_golden('default_ragged')
    hp, w, img, lens = golden_inputs(meta)
    enc = ConformerEncoder(in_channels=1, input_dim=hp['height'], encoder_dim=hp['encoder_dim'],
                           num_layers=hp['num_encoder_layers'], num_attention_heads=hp['num_attention_heads'],
                           feed_forward_expansion_factor=hp['feed_forward_expansion_factor'],
                           conv_expansion_factor=2, conv_kernel_size=hp['conv_kernel_size'],
                           half_step_residual=True, subsampling_conv_channels=hp['subsampling_conv_channels'],
                           subsampling_factor=4)
    enc.load_state_dict({k[len('encoder.'):]: torch.from_numpy(v) for k, v in w.items() if k.startswith('encoder.')},
                        strict=False)
    x_view = torch.from_numpy(img).to(cuda_device).squeeze(1).transpose(1, 2)       # (B,W,H) view of NCHW
    out_a, lens_a = enc(x_view, torch.from_numpy(lens))
    out_b, lens_b = enc(x_view.contiguous(), torch.from_numpy(lens))                 # (B,W,H) contiguous
    ref, ref_lens = O.encoder_forward(np.ascontiguousarray(img[:, 0].transpose(0, 2, 1)), lens, w, hp, prefix='encoder.')
    assert torch.equal(out_a, out_b)
    assert rel_err(out_a.cpu().numpy(), ref) < FP32_TOL
    np.testing.assert_array_equal(lens_a.numpy(), ref_lens)


def test_ragged_mask_mode_matches_masked_oracle(cuda_device):
    meta, _ = load_golden('tiny_ragged')
    hp, w, img, lens = golden_inputs(meta)
    net = build_net(hp, meta['num_classes'], w, cuda_device, mask_mode='ragged')
    logits, olens = net(torch.from_numpy(img).to(cuda_device), torch.from_numpy(lens))
    ref, ref_lens = O.recognition_forward(img, lens, w, hp, masked=True)
    got = logits.cpu().numpy()
    for b, n in enumerate(ref_lens):
        # frames inside the line: attention ignores the padding keys exactly like the masked oracle
        assert rel_err(got[b, :n], ref[b, :n]) < FP32_TOL


@pytest.mark.parametrize('B,T,C', [(1, 1, 2), (3, 17, 5), (5, 700, 300), (2, 64, 10000)])
def test_greedy_decode_kernel_vs_oracle(B, T, C, cuda_device):
    from conformer_ocr_b200.engine import unpack_decoded
    from conformer_ocr_b200.pred import _decode_engine
    rng = np.random.RandomState(B * 1000 + T)
    # few distinct values: long runs, many exact ties (lowest index must win), blanks
    logits = rng.randint(0, 3, size=(B, T, C)).astype(np.float32)
    logits[:, :, 0] += (rng.uniform(size=(B, T)) < 0.3) * 2.0
    lens = rng.randint(0, T + 1, size=(B,)).astype(np.int32)
    lens[0] = T
    eng = _decode_engine(cuda_device)
    got = unpack_decoded(eng.greedy_decode(torch.from_numpy(logits).to(cuda_device), torch.from_numpy(lens)))
    ref = [O.greedy_decoder(logits[b, :lens[b]].T) for b in range(B)]
    assert got == ref            # integer columns AND the max-score column are exact here (small integers)


def test_host_buffer_entry_point_equals_device_path(cuda_device):
    meta, _ = load_golden('default_ragged')
    hp, w, img, lens = golden_inputs(meta)
    net = build_net(hp, meta['num_classes'], w, cuda_device)
    eng = net.engine_for(cuda_device)
    from conformer_ocr_b200.engine import unpack_decoded
    host = eng.recognize_host(torch.from_numpy(img).pin_memory(), torch.from_numpy(lens).to(torch.int32))
    dev = net.predict_labels(torch.from_numpy(img).to(cuda_device), torch.from_numpy(lens))
    assert unpack_decoded(host) == dev
    np.testing.assert_array_equal(host['out_lens'], O.calc_length(lens))


def test_pipelined_host_entry_points(cuda_device):
    """cocr_recognize_host_begin/_end with two batches in flight return what the one-shot call returns, and the slot
    protocol is enforced."""
    from conformer_ocr_b200.engine import Engine
    meta, _ = load_golden('default_ragged')
    hp, w, img, lens = golden_inputs(meta)
    net = build_net(hp, meta['num_classes'], w, cuda_device)
    eng = net.engine_for(cuda_device)
    a = torch.from_numpy(img).contiguous().pin_memory()
    b = torch.from_numpy(img[::-1].copy()).contiguous().pin_memory()
    la, lb = torch.from_numpy(lens).int(), torch.from_numpy(lens[::-1].copy()).int()
    ref_a, ref_b = eng.recognize_host(a, la), eng.recognize_host(b, lb)
    eng.recognize_host_begin(0, a, la)
    eng.recognize_host_begin(1, b, lb)
    with pytest.raises(RuntimeError):
        eng.recognize_host_begin(1, b, lb)          # slot still busy
    got_a, got_b = eng.recognize_host_end(0), eng.recognize_host_end(1)
    for got, ref in ((got_a, ref_a), (got_b, ref_b)):
        for k in ('labels', 'starts', 'ends', 'counts', 'out_lens'):
            n = ref['counts']
            if k in ('labels', 'starts', 'ends'):
                for i in range(len(n)):
                    np.testing.assert_array_equal(got[k][i, :n[i]], ref[k][i, :n[i]])
            else:
                np.testing.assert_array_equal(got[k], ref[k])
    with pytest.raises(RuntimeError):
        eng.recognize_host_end(0)                    # nothing pending


def test_errors_cross_the_abi_as_python_exceptions(cuda_device):
    meta, _ = load_golden('tiny_equal')
    hp, w, img, lens = golden_inputs(meta)
    net = build_net(hp, meta['num_classes'], w, cuda_device)
    with pytest.raises(ValueError):
        net(torch.zeros(2, 1, hp['height'] + 1, 40, device=cuda_device), torch.tensor([40, 40]))
    with pytest.raises(ValueError):
        net(torch.from_numpy(img).to(cuda_device), torch.tensor([40]))          # wrong number of lengths
    bad = dict(w)
    bad['decoder.bias'] = np.zeros(3, dtype=np.float32)
    from conformer_ocr_b200.engine import Engine
    eng = Engine(device=cuda_device, num_classes=meta['num_classes'], **net.nn['encoder'].hparams)
    with pytest.raises(ValueError, match='size mismatch'):
        eng.load_state_dict(bad)


def test_full_width_batch_properties(cuda_device):
    """BASELINE config 1 size (batch 8, 1024 px, default spec): line independence (a property the
    domain offers at sizes the numpy oracle would need minutes for) + oracle parity on 2 of the lines."""
    hp = dict(O.DEFAULT_HPARAMS)
    C = 256
    w = O.make_weights(hp, C, seed=21)
    img, lens = O.make_lines(8, 96, 1024, seed=22)
    net = build_net(hp, C, w, cuda_device)
    x = torch.from_numpy(img).to(cuda_device)
    logits, _ = net(x, torch.from_numpy(lens))
    # equal-width lines are independent: any sub-batch gives the same rows
    sub, _ = net(x[2:5], torch.from_numpy(lens[2:5]))
    assert torch.equal(logits[2:5], sub)
    perm = torch.tensor([7, 0, 3])
    sub2, _ = net(x[perm], torch.from_numpy(lens[perm.numpy()]))
    assert torch.equal(logits[perm], sub2)
    ref, _ = O.recognition_forward(img[:2], lens[:2], w, hp)
    got = logits[:2].cpu().numpy()
    assert rel_err(got, ref) < FP32_TOL
    assert (got.argmax(-1) == ref.argmax(-1)).mean() == 1.0
