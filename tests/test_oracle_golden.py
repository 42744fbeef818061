"""The numpy oracle against the golden vectors produced by the reference modules
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from conftest import GOLDEN_CASES, golden_inputs, load_golden, rel_err
from oracle import conformer_oracle as O

# fp32 numpy vs fp32 torch: both round differently (BLAS blocking, exp), nothing more.
ORACLE_TOL = 2e-5


@pytest.mark.parametrize('name', GOLDEN_CASES)
def test_oracle_matches_reference_logits(name):
    meta, arrs = load_golden(name)
    hp, w, img, lens = golden_inputs(meta)
    logits, olens = O.recognition_forward(img, lens, w, hp)
    assert logits.dtype == np.float32
    assert olens.dtype == np.int32
    np.testing.assert_array_equal(olens, arrs['enc_lens'])
    assert logits.shape == arrs['logits'].shape
    assert rel_err(logits, arrs['logits']) < ORACLE_TOL
    # argmax identical on every frame -> identical collapsed label sequences
    np.testing.assert_array_equal(logits.argmax(-1), arrs['logits'].argmax(-1))


@pytest.mark.parametrize('name', ['tiny_equal', 'tiny_ragged'])
def test_oracle_matches_reference_intermediates(name):
    meta, arrs = load_golden(name)
    hp, w, img, lens = golden_inputs(meta)
    x = np.ascontiguousarray(img[:, 0].transpose(0, 2, 1))
    sub, _ = O.conv2d_subsampling(x, lens, w, 'encoder.conv_subsample.')
    assert rel_err(sub, arrs['inter_subsample']) < ORACLE_TOL
    p = 'encoder.layers.0.sequential.'
    h = hp['num_attention_heads']
    y = O.feed_forward(sub, w, p + '0.module.') * np.float32(0.5) + sub
    assert rel_err(y, arrs['inter_l0_ffn1']) < ORACLE_TOL
    y2 = O.rel_mhsa(y, w, p + '1.module.', h) + y
    assert rel_err(y2, arrs['inter_l0_mhsa']) < ORACLE_TOL
    y3 = O.conv_module(y2, w, p + '2.module.', hp['conv_kernel_size']) + y2
    assert rel_err(y3, arrs['inter_l0_conv']) < ORACLE_TOL
    y4 = O.feed_forward(y3, w, p + '3.module.') * np.float32(0.5) + y3
    assert rel_err(y4, arrs['inter_l0_ffn2']) < ORACLE_TOL
    y5 = O.layer_norm(y4, w[p + '4.weight'], w[p + '4.bias'])
    assert rel_err(y5, arrs['inter_l0_out']) < ORACLE_TOL


def test_float64_oracle_close_to_float32_reference():
    meta, arrs = load_golden('default_equal')
    hp, w, img, lens = golden_inputs(meta)
    w64 = {k: v.astype(np.float64) for k, v in w.items()}
    logits, _ = O.recognition_forward(img.astype(np.float64), lens, w64, hp)
    assert rel_err(logits, arrs['logits']) < ORACLE_TOL


def test_calc_length_values():
    # SURVEY.md §8 notation, probed on the reference
    got = O.calc_length(np.array([1024, 1025, 1536, 2048, 4096, 100, 5, 4, 3, 2, 1]))
    np.testing.assert_array_equal(got, [256, 257, 384, 512, 1024, 25, 2, 1, 1, 1, 1])
    assert int(O.calc_length(np.array(96))) == 24


def test_relative_shift_closed_form_equals_reference_idiom():
    rng = np.random.RandomState(0)
    for T in (1, 2, 5, 16, 33):
        m = rng.normal(size=(2, 3, T, 2 * T - 1)).astype(np.float32)
        np.testing.assert_array_equal(O.relative_shift(m), O.relative_shift_reference_form(m))


def test_rel_pos_table_offsets():
    T, d = 7, 12
    pe = O.rel_pos_table(T, d)
    assert pe.shape == (2 * T - 1, d)
    # row T-1 is offset 0: sin 0 = 0, cos 0 = 1
    np.testing.assert_allclose(pe[T - 1, 0::2], 0.0, atol=1e-7)
    np.testing.assert_allclose(pe[T - 1, 1::2], 1.0, atol=1e-7)
    # rows are antisymmetric in sin, symmetric in cos
    np.testing.assert_allclose(pe[::-1, 0::2], -pe[:, 0::2], atol=1e-6)
    np.testing.assert_allclose(pe[::-1, 1::2], pe[:, 1::2], atol=1e-6)


def test_masked_mode_equals_unmasked_for_equal_widths_and_differs_for_ragged():
    meta, _ = load_golden('tiny_ragged')
    hp, w, img, lens = golden_inputs(meta)
    a, _ = O.recognition_forward(img, lens, w, hp, masked=False)
    b, olens = O.recognition_forward(img, lens, w, hp, masked=True)
    np.testing.assert_allclose(a[0], b[0], rtol=0, atol=1e-5)   # widest line: nothing masked
    assert np.abs(a[2, :olens[2]] - b[2, :olens[2]]).max() > 1e-3


@pytest.mark.parametrize('name', GOLDEN_CASES)
def test_torch_port_matches_reference_logits(name):
    import torch
    from oracle import torch_port as TP
    meta, arrs = load_golden(name)
    hp, w, img, lens = golden_inputs(meta)
    logits, olens = TP.recognition_forward(torch.from_numpy(img), torch.from_numpy(lens), TP.to_torch(w), hp)
    np.testing.assert_array_equal(olens.numpy(), arrs['enc_lens'])
    assert rel_err(logits.numpy(), arrs['logits']) < ORACLE_TOL
