"""bf16 tensor-core path: kernel self checks through the C ABI + end-to-end parity with the reference goldens.

Tolerance (BASELINE.json north_star): bf16 logits max-norm relative error <= 2e-2; top-1 agreement is
reported per case and gated on frames whose fp32 top-2 margin exceeds the bf16 noise floor."""
import ctypes

import numpy as np
import pytest
import torch

from conftest import GOLDEN_CASES, golden_inputs, load_golden, rel_err
from oracle import conformer_oracle as O

pytestmark = pytest.mark.gpu
BF16_TOL = 2e-2
DROP = dict(input_dropout_p=0.1, feed_forward_dropout_p=0.1, attention_dropout_p=0.1, conv_dropout_p=0.1)

EPI_BF16, EPI_SILU, EPI_F32, EPI_ARGMAX, EPI_GLU = 0, 1, 2, 3, 6
GEMM_CASES = [
    # (epilogue, M, N, K, block_n override)
    (EPI_BF16, 128, 64, 64, 0), (EPI_BF16, 128, 144, 144, 0), (EPI_BF16, 300, 144, 144, 0),
    (EPI_SILU, 4096, 576, 144, 0), (EPI_BF16, 4096, 144, 576, 0), (EPI_BF16, 5000, 432, 144, 0),
    (EPI_BF16, 3000, 288, 144, 0), (EPI_BF16, 2500, 144, 768, 0), (EPI_BF16, 777, 32, 32, 0),
    (EPI_F32, 1000, 256, 144, 0), (EPI_F32, 515, 11, 32, 0), (EPI_F32, 2048, 10000, 144, 0),
    (EPI_ARGMAX, 1000, 256, 144, 0), (EPI_ARGMAX, 3000, 10000, 144, 0), (EPI_ARGMAX, 130, 11, 32, 0),
    (EPI_BF16, 40000, 576, 144, 0), (EPI_BF16, 40000, 144, 576, 0), (EPI_BF16, 1000, 144, 144, 32), (EPI_BF16, 1000, 144, 144, 96),
    (EPI_BF16, 5000, 1280, 144, 0), (EPI_SILU, 777, 576, 144, 0), (EPI_BF16, 333, 288, 144, 0), (EPI_BF16, 1000, 600, 144, 0),
    (EPI_SILU, 60000, 576, 144, 0), (EPI_BF16, 129, 96, 48, 0),
    # first pointwise conv + GLU of the conv module (convolution.py:139-141): the production tiling (two N tiles of 144, whose
    # last 32-column piece is half a piece: it must not spill into the next tile's outputs), three tiles of 96, a tiny model
    (EPI_GLU, 3000, 288, 144, 144), (EPI_GLU, 40000, 288, 144, 144), (EPI_GLU, 777, 288, 144, 144), (EPI_GLU, 3000, 288, 144, 0),
    (EPI_GLU, 1000, 64, 32, 0),
]


@pytest.mark.parametrize('epi,M,N,K,bn', GEMM_CASES)
def test_tcgen05_gemm_selftest(epi, M, N, K, bn, cuda_device):
    from conformer_ocr_b200 import _lib
    lib = _lib.load()
    err = ctypes.c_double(-1.0)
    rc = lib.cocr_selftest(0, 10 + epi, M, N, K + 100000 * bn, ctypes.byref(err))
    _lib.raise_for_status(rc, None)
    # bf16 outputs: error relative to 1+|ref| bounded by bf16 rounding (2^-8) (+ tanh.approx for SiLU);
    # fp32 / arg-max outputs: accumulation-order differences only
    tol = {EPI_BF16: 6e-3, EPI_SILU: 8e-3, EPI_F32: 2e-3, EPI_ARGMAX: 2e-3, EPI_GLU: 8e-3}[epi]
    assert 0.0 <= err.value < tol, f'max err {err.value}'


@pytest.mark.parametrize('M,N,K,flags', [(1000, 144, 144, 4), (4096, 144, 576, 4), (777, 144, 576, 6), (5000, 144, 576, 2),
                                          (3000, 144, 768, 5), (300, 144, 144, 0), (100, 144, 64, 6), (2000, 144, 144, 7),
                                          (40000, 144, 576, 6), (40001, 144, 144, 4)])
def test_residual_layernorm_epilogue_selftest(M, N, K, flags, cuda_device):
    from conformer_ocr_b200 import _lib
    lib = _lib.load()
    err = ctypes.c_double(-1.0)
    rc = lib.cocr_selftest(0, 20, M, N, K + 100000 * flags, ctypes.byref(err))
    _lib.raise_for_status(rc, None)
    # xn is bf16 (2^-8 rounding); the fp32 stream error is weighted 8x inside the self test
    assert 0.0 <= err.value < 6e-3, f'max err {err.value}'


@pytest.mark.parametrize('M,N,K,flags', [
    (128, 64, 64, 0), (128, 64, 64, 1), (128, 64, 64, 2), (128, 64, 64, 3),                      # every operand-major combination
    (384, 384, 36, 4), (384, 767, 36, 4), (384, 36, 384, 6), (384, 36, 384, 7), (767, 36, 384, 7),   # attention shapes, batched 3 x 2
    (300, 144, 576, 0), (300, 576, 144, 2), (1000, 288, 144, 0),                                   # linears and data gradients
    (576, 144, 1000, 11), (144, 576, 1000, 11), (432, 144, 5000, 11),                              # weight gradients: split contraction
    (26, 26, 16, 4), (26, 16, 26, 6), (51, 16, 26, 7), (10, 10, 16, 4), (250, 200, 100, 7), (77, 16, 40, 1),   # ragged edges
])
def test_bgemm_selftest(M, N, K, flags, cuda_device):
    """Batched tcgen05 GEMM of the training step (bgemm_tc.cu) against a scalar device reference on the same bf16 operands.
    flags: 1 A stored [k][m], 2 B stored [k][n], 4 batched 3 x 2 with strided blocks, 8 split contraction with atomic output.
    Both sides accumulate in fp32; the difference is the summation order (<= 1e-3 absolute at K = 5000 with |a|,|b| <= 1)."""
    from conformer_ocr_b200 import _lib
    lib = _lib.load()
    err = ctypes.c_double()
    rc = lib.cocr_selftest(0, 30, M, N, K + 100000 * flags, ctypes.byref(err))
    assert rc == 0, lib.cocr_last_error(None).decode()
    assert 0 <= err.value < 1e-3, err.value


@pytest.mark.parametrize('M,NC,flags', [(128, 6, 4), (300, 6, 6), (1000, 6, 0), (777, 2, 2), (4096, 6, 4), (2500, 1, 6), (19000, 6, 6), (20481, 8, 4), (5000, 3, 2)])
def test_fused_ffn_selftest(M, NC, flags, cuda_device):
    """ffn_tc.cuh against a host reference: x + 0.5 * (W2 silu(W1 xn + b1) + b2), LayerNorm(s), bf16 hidden operand."""
    from conformer_ocr_b200 import _lib
    lib = _lib.load()
    err = ctypes.c_double(-1.0)
    rc = lib.cocr_selftest(0, 21, M, NC, flags, ctypes.byref(err))
    _lib.raise_for_status(rc, None)
    # the fp32 stream differs by accumulation order and tanh.approx (both ~1e-3 of the module output); xn adds bf16 rounding
    assert 0.0 <= err.value < 6e-3, f'max err {err.value}'


def build_net(hp, C, w, device, **kw):
    from conformer_ocr_b200 import PytorchRecognitionModel
    net = PytorchRecognitionModel(num_classes=C, codec=None, precision='bf16', **hp, **DROP, **kw)
    net.nn.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()}, strict=False)
    return net.to(device)


def top1_report(got, ref):
    """agreement overall and on frames whose reference top-2 margin is above 4x the observed max error."""
    agree = got.argmax(-1) == ref.argmax(-1)
    srt = np.sort(ref, axis=-1)
    margin = srt[..., -1] - srt[..., -2]
    noise = np.abs(got - ref).max()
    clear = margin > 4 * noise
    return float(agree.mean()), float(agree[clear].mean()) if clear.any() else 1.0, float(clear.mean())


@pytest.mark.parametrize('name', GOLDEN_CASES)
def test_bf16_logits_match_reference_golden(name, cuda_device):
    meta, arrs = load_golden(name)
    hp, w, img, lens = golden_inputs(meta)
    net = build_net(hp, meta['num_classes'], w, cuda_device)
    logits, olens = net(torch.from_numpy(img).to(cuda_device), torch.from_numpy(lens))
    got = logits.cpu().numpy()
    np.testing.assert_array_equal(olens.cpu().numpy(), arrs['enc_lens'])
    err = rel_err(got, arrs['logits'])
    overall, clear, frac_clear = top1_report(got, arrs['logits'])
    print(f'{name}: bf16 rel err {err:.3e}; top-1 agreement {overall:.4f} overall, {clear:.4f} on the '
          f'{frac_clear:.2%} frames with a clear margin')
    assert err < BF16_TOL
    assert clear == 1.0
    assert overall >= 0.999          # north_star: top-1 agreement on >= 99.9 % of ALL frames, not only the clear ones


def test_bf16_decode_consistent_with_own_logits(cuda_device):
    """the fused projection+arg-max kernel must pick exactly the arg-max of the logits the same path emits,
    and the collapse must be bit-exact given that arg-max."""
    meta, _ = load_golden('default_ragged')
    hp, w, img, lens = golden_inputs(meta)
    net = build_net(hp, meta['num_classes'], w, cuda_device)
    x = torch.from_numpy(img).to(cuda_device)
    logits, olens = net(x, torch.from_numpy(lens))
    got = net.predict_labels(x, torch.from_numpy(lens))
    lg = logits.cpu().numpy()
    ref = [O.greedy_decoder(lg[b, :int(olens[b])].T) for b in range(lg.shape[0])]
    assert [[t[:3] for t in l] for l in got] == [[t[:3] for t in l] for l in ref]


def test_bf16_full_width_batch(cuda_device):
    """BASELINE config 1 shape (batch 8 x 1024 px, default spec) against the fp32 CUDA path (itself pinned to
    the reference within 1e-4): tolerance + line independence."""
    from conformer_ocr_b200 import PytorchRecognitionModel
    hp = dict(O.DEFAULT_HPARAMS)
    C = 256
    w = O.make_weights(hp, C, seed=21)
    img, lens = O.make_lines(8, 96, 1024, seed=22)
    x = torch.from_numpy(img).to(cuda_device)
    net16 = build_net(hp, C, w, cuda_device)
    net32 = PytorchRecognitionModel(num_classes=C, codec=None, precision='fp32', **hp, **DROP)
    net32.nn.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()}, strict=False)
    l16, _ = net16(x, torch.from_numpy(lens))
    l32, _ = net32(x, torch.from_numpy(lens))
    a, b = l16.cpu().numpy(), l32.cpu().numpy()
    err = rel_err(a, b)
    overall, clear, frac_clear = top1_report(a, b)
    print(f'b8 w1024: bf16 vs fp32 rel err {err:.3e}; top-1 {overall:.4f} overall, {clear:.4f} on {frac_clear:.2%} clear frames')
    assert err < BF16_TOL
    assert clear == 1.0
    sub, _ = net16(x[3:6], torch.from_numpy(lens[3:6]))
    assert torch.equal(l16[3:6], sub)


def test_bf16_long_lines_large_alphabet(cuda_device):
    """BASELINE config 5 shape at reduced depth: 4096-px lines (T = 1024 frames: 8 query tiles per line, the E table
    streams with K instead of staying resident) and a 2 500-class alphabet (10 N tiles in the fused arg-max projection),
    against the fp32 CUDA path.  Also checks that the fused decode equals the arg-max of the path's own logits."""
    from conformer_ocr_b200 import PytorchRecognitionModel
    hp = dict(O.DEFAULT_HPARAMS)
    hp['num_encoder_layers'] = 3
    C = 2500
    w = O.make_weights(hp, C, seed=31)
    img, lens = O.make_lines(3, 96, 4096, seed=32)
    x = torch.from_numpy(img).to(cuda_device)
    tl = torch.from_numpy(lens)
    net16 = build_net(hp, C, w, cuda_device)
    net32 = PytorchRecognitionModel(num_classes=C, codec=None, precision='fp32', **hp, **DROP)
    net32.nn.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()}, strict=False)
    l16, ol = net16(x, tl)
    l32, _ = net32(x, tl)
    assert l16.shape == (3, 1024, C)
    a, b = l16.cpu().numpy(), l32.cpu().numpy()
    err = rel_err(a, b)
    overall, clear, frac_clear = top1_report(a, b)
    print(f'b3 w4096 C2500: bf16 vs fp32 rel err {err:.3e}; top-1 {overall:.4f} overall, {clear:.4f} on {frac_clear:.2%} clear frames')
    assert err < BF16_TOL
    assert clear == 1.0
    got = net16.predict_labels(x, tl)
    ref = [O.greedy_decoder(a[i, :int(ol[i])].T) for i in range(3)]
    assert [[t[:3] for t in l] for l in got] == [[t[:3] for t in l] for l in ref]


def test_bf16_ragged_mask_mode(cuda_device):
    """mask_mode='ragged' on the tensor-core path: keys beyond a line's own length are excluded.  Checked against the
    fp32 CUDA path in the same mode (itself pinned to the masked oracle), on lines whose lengths end inside a key tile,
    exactly on a tile boundary and below one tile."""
    from conformer_ocr_b200 import PytorchRecognitionModel
    hp = dict(O.DEFAULT_HPARAMS)
    hp['num_encoder_layers'] = 2
    C = 64
    w = O.make_weights(hp, C, seed=41)
    widths = [1536, 1100, 1024, 512, 300, 40]
    img, lens = O.make_lines(len(widths), 96, 1536, seed=42, widths=widths)
    x = torch.from_numpy(img).to(cuda_device)
    tl = torch.from_numpy(lens)
    net16 = build_net(hp, C, w, cuda_device, mask_mode='ragged')
    net32 = PytorchRecognitionModel(num_classes=C, codec=None, precision='fp32', mask_mode='ragged', **hp, **DROP)
    net32.nn.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()}, strict=False)
    l16, ol = net16(x, tl)
    l32, _ = net32(x, tl)
    a, b = l16.cpu().numpy(), l32.cpu().numpy()
    for i in range(len(widths)):
        n = int(ol[i])
        err = rel_err(a[i, :n], b[i, :n])
        assert err < BF16_TOL, f'line {i} (width {widths[i]}): rel err {err}'


@pytest.mark.parametrize('B,W', [(1, 5), (3, 100), (5, 1025), (2, 3001), (7, 513), (130, 64)])
def test_bf16_odd_shapes_against_fp32(B, W, cuda_device):
    """Degenerate and ragged tile shapes (one frame per line, partial M / query / key tiles, more lines than one CTA wave
    of work items, streamed E table) on the tensor-core path against the fp32 CUDA path."""
    from conformer_ocr_b200 import PytorchRecognitionModel
    hp = dict(O.DEFAULT_HPARAMS)
    hp['num_encoder_layers'] = 2
    C = 97
    w = O.make_weights(hp, C, seed=51)
    img, lens = O.make_lines(B, 96, W, seed=52 + B)
    x = torch.from_numpy(img).to(cuda_device)
    tl = torch.from_numpy(lens)
    net16 = build_net(hp, C, w, cuda_device)
    net32 = PytorchRecognitionModel(num_classes=C, codec=None, precision='fp32', **hp, **DROP)
    net32.nn.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()}, strict=False)
    l16, ol16 = net16(x, tl)
    l32, ol32 = net32(x, tl)
    assert torch.equal(ol16.cpu(), ol32.cpu())
    a, b = l16.cpu().numpy(), l32.cpu().numpy()
    assert np.isfinite(a).all()
    assert rel_err(a, b) < BF16_TOL
    got = net16.predict_labels(x, tl)
    ref = [O.greedy_decoder(a[i, :int(ol16[i])].T) for i in range(B)]
    assert [[t[:3] for t in l] for l in got] == [[t[:3] for t in l] for l in ref]


def test_uint8_input_is_bit_identical_to_scaled_float(cuda_device):
    """raw 8-bit pixels through cocr_forward_decode_u8 / cocr_recognize_host_u8_begin give exactly what
    `img.float() / 255` gives through the float entry points (the kernel divides by 255 with correct rounding)."""
    from conformer_ocr_b200 import default_specs as D
    from conformer_ocr_b200 import synthetic as S
    from conformer_ocr_b200.engine import Engine
    hp = dict(D.RECOGNITION_HYPER_PARAMS)
    C_, B, W = 60, 5, 700
    eng = Engine(height=hp['height'], encoder_dim=hp['encoder_dim'], num_layers=2, num_heads=hp['num_attention_heads'],
                 ff_expansion=hp['feed_forward_expansion_factor'], conv_expansion=hp['conv_expansion_factor'],
                 conv_kernel_size=hp['conv_kernel_size'], half_step_residual=hp['half_step_residual'],
                 subsampling_conv_channels=hp['subsampling_conv_channels'], subsampling_factor=hp['subsampling_factor'],
                 num_classes=C_, device=cuda_device, precision='bf16')
    hp2 = dict(hp); hp2['num_encoder_layers'] = 2
    eng.load_state_dict(S.make_weights(hp2, C_, seed=3))
    g = torch.Generator().manual_seed(5)
    u8 = torch.randint(0, 256, (B, 1, hp['height'], W), generator=g, dtype=torch.uint8)
    lens = torch.tensor([700, 650, 700, 123, 400], dtype=torch.int32)
    f32 = u8.float() / 255
    a = eng.forward_decode(f32.to(cuda_device).squeeze(1).transpose(1, 2), lens)
    b = eng.forward_decode(u8.to(cuda_device).squeeze(1).transpose(1, 2), lens)
    for k in ('labels', 'starts', 'ends', 'counts', 'out_lens', 'conf'):
        n = a['counts'].cpu().numpy()
        x, y = a[k].cpu().numpy(), b[k].cpu().numpy()
        if x.ndim == 2:
            for i in range(B):
                np.testing.assert_array_equal(x[i, :n[i]], y[i, :n[i]])
        else:
            np.testing.assert_array_equal(x, y)
    # host entry points: uint8 pinned batch in, same records out
    eng.recognize_host_begin(0, f32.contiguous().pin_memory(), lens)
    eng.recognize_host_begin(1, u8.contiguous().pin_memory(), lens)
    ra, rb = eng.recognize_host_end(0), eng.recognize_host_end(1)
    np.testing.assert_array_equal(ra['counts'], rb['counts'])
    for i in range(B):
        np.testing.assert_array_equal(ra['labels'][i, :ra['counts'][i]], rb['labels'][i, :rb['counts'][i]])
        np.testing.assert_array_equal(ra['conf'][i, :ra['counts'][i]], rb['conf'][i, :rb['counts'][i]])
    # the encoder / logits calls accept uint8 too (converted on the device)
    e1, _ = eng.encode(f32.to(cuda_device).squeeze(1).transpose(1, 2), lens)
    e2, _ = eng.encode(u8.to(cuda_device).squeeze(1).transpose(1, 2), lens)
    assert torch.equal(e1, e2)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs in one process')
def test_two_devices_in_one_process(cuda_device):
    """Handles on two GPUs of the same process: per-device function attributes (dynamic shared memory opt-in), per-handle
    workspaces and graphs; both devices give the same records for the same batch."""
    from conformer_ocr_b200.pred import PytorchRecognitionModel
    meta, _ = load_golden('default_equal')
    hp, w, img, lens = golden_inputs(meta)
    outs = []
    for dev in (torch.device('cuda', 1), torch.device('cuda', 0)):          # second device first: nothing cached for it yet
        net = build_net(hp, meta['num_classes'], w, dev)
        x = torch.from_numpy(img).to(dev)
        for _ in range(3):                                                  # eager, capture, replay
            rec = net.predict_labels(x, torch.from_numpy(lens))
        outs.append(rec)
    assert [[r[0] for r in line] for line in outs[0]] == [[r[0] for r in line] for line in outs[1]]


def test_bf16_long_run_is_stable(cuda_device):
    """300 back-to-back steps of the benchmark workload: every step must succeed and reproduce the first step's records.
    (A variant of the fused FFN kernel with two MMA-issuing warps passed every parity test and died with an unspecified
    launch failure about once in a hundred steps; short tests never saw it.)"""
    from conformer_ocr_b200 import default_specs as D
    from conformer_ocr_b200 import synthetic as S
    from conformer_ocr_b200.engine import Engine
    hp = dict(D.RECOGNITION_HYPER_PARAMS)
    B, W, C_ = 256, 1536, 256
    eng = Engine(height=hp['height'], encoder_dim=hp['encoder_dim'], num_layers=hp['num_encoder_layers'],
                 num_heads=hp['num_attention_heads'], ff_expansion=hp['feed_forward_expansion_factor'],
                 conv_expansion=hp['conv_expansion_factor'], conv_kernel_size=hp['conv_kernel_size'],
                 half_step_residual=hp['half_step_residual'], subsampling_conv_channels=hp['subsampling_conv_channels'],
                 subsampling_factor=hp['subsampling_factor'], num_classes=C_, device=cuda_device, precision='bf16')
    eng.load_state_dict(S.make_weights(hp, C_, seed=0))
    g = torch.Generator(device=cuda_device).manual_seed(99)
    x = torch.rand((B, 1, hp['height'], W), generator=g, device=cuda_device).squeeze(1).transpose(1, 2)
    lens = torch.full((B,), W, dtype=torch.int32)
    first = None
    for step in range(300):
        res = eng.forward_decode(x, lens)
        if step % 50 == 0 or step == 299:
            torch.cuda.synchronize(cuda_device)
            counts, labels = res['counts'].cpu(), res['labels'].cpu()
            valid = torch.arange(labels.shape[1])[None, :] < counts[:, None]          # entries behind a row's count are unset
            cur = (counts, torch.where(valid, labels, torch.zeros_like(labels)))
            if first is None:
                first = cur
            else:
                assert torch.equal(cur[0], first[0]) and torch.equal(cur[1], first[1]), f'step {step} differs from step 0'
    torch.cuda.synchronize(cuda_device)


def test_bf16_soak_many_shapes_both_mask_modes(cuda_device):
    """2 400 forward passes cycling over six (batch, width) shapes in both mask modes on long-lived engines (graph capture,
    replay, LRU cache, growing E table, ragged lengths): every pass must succeed and reproduce the first result of its
    shape bit for bit.  (the 300-step test covers one shape; compute-sanitizer is not available on this pool.)"""
    from conformer_ocr_b200 import default_specs as D
    from conformer_ocr_b200 import synthetic as S
    from conformer_ocr_b200.engine import Engine
    hp = dict(D.RECOGNITION_HYPER_PARAMS)
    hp2 = dict(hp); hp2['num_encoder_layers'] = 4
    C_ = 120
    w = S.make_weights(hp2, C_, seed=5)
    shapes = [(32, 1536), (7, 700), (64, 512), (3, 4096), (16, 1534), (1, 37)]
    for mask_mode in ('reference_unmasked', 'ragged'):
        eng = Engine(height=hp['height'], encoder_dim=hp['encoder_dim'], num_layers=4, num_heads=hp['num_attention_heads'],
                     ff_expansion=hp['feed_forward_expansion_factor'], conv_expansion=hp['conv_expansion_factor'],
                     conv_kernel_size=hp['conv_kernel_size'], half_step_residual=hp['half_step_residual'],
                     subsampling_conv_channels=hp['subsampling_conv_channels'], subsampling_factor=hp['subsampling_factor'],
                     num_classes=C_, device=cuda_device, precision='bf16', mask_mode=mask_mode)
        eng.load_state_dict(w)
        g = torch.Generator(device=cuda_device).manual_seed(7)
        inputs, first = [], {}
        for B, W in shapes:
            x = torch.rand((B, 1, hp['height'], W), generator=g, device=cuda_device).squeeze(1).transpose(1, 2)
            lens = torch.randint(max(1, W // 2), W + 1, (B,), generator=torch.Generator().manual_seed(B + W), dtype=torch.int32)
            lens[0] = W
            inputs.append((x, lens))
        for step in range(1200):
            i = step % len(shapes) if step % 7 else (step // 7) % len(shapes)          # irregular order
            x, lens = inputs[i]
            res = eng.forward_decode(x, lens)
            if step % 97 == 0 or step < len(shapes) * 7:
                counts, labels = res['counts'].cpu(), res['labels'].cpu()
                valid = torch.arange(labels.shape[1])[None, :] < counts[:, None]
                cur = (counts, torch.where(valid, labels, torch.zeros_like(labels)))
                if i not in first:
                    first[i] = cur
                else:
                    assert torch.equal(cur[0], first[i][0]) and torch.equal(cur[1], first[i][1]), f'{mask_mode} step {step} shape {shapes[i]}'
        torch.cuda.synchronize(cuda_device)
        assert len(first) == len(shapes)
        eng.close()


def test_kernel_plan_reports_fallbacks(cuda_device):
    """The default spec runs every component on its fast kernel; other hyper-parameters are served by general kernels and say so
    (cocr_kernel_plan + one RuntimeWarning from the shim) instead of dropping to them silently."""
    import warnings
    from conformer_ocr_b200.engine import Engine
    from conformer_ocr_b200 import synthetic as S
    hp = dict(O.DEFAULT_HPARAMS, num_encoder_layers=1)
    kw = dict(height=hp['height'], encoder_dim=hp['encoder_dim'], num_layers=1, num_heads=hp['num_attention_heads'],
              ff_expansion=4, conv_expansion=2, conv_kernel_size=hp['conv_kernel_size'], half_step_residual=True,
              subsampling_conv_channels=hp['subsampling_conv_channels'], subsampling_factor=4, num_classes=16, device=cuda_device, precision='bf16')
    eng = Engine(**kw)
    eng.load_state_dict(O.make_weights(hp, 16, seed=1))
    plan, nf = eng.kernel_plan()
    assert nf == 0 and 'k_attn_tc' in plan and 'k_frontend_tc' in plan, plan
    eng.close()
    tiny = dict(O.DEFAULT_HPARAMS, num_encoder_layers=1, encoder_dim=32, num_attention_heads=2, height=16, subsampling_conv_channels=4, conv_kernel_size=7)
    kw.update(height=16, encoder_dim=32, num_heads=2, conv_kernel_size=7, subsampling_conv_channels=4)
    Engine._warned_fallback = False
    eng = Engine(**kw)
    with warnings.catch_warnings(record=True) as rec:
        warnings.simplefilter('always')
        eng.load_state_dict(O.make_weights(tiny, 16, seed=1))
    plan, nf = eng.kernel_plan()
    assert nf >= 3 and 'CUDA-core attention' in plan, plan
    assert any(issubclass(r.category, RuntimeWarning) and 'general kernels' in str(r.message) for r in rec)
    eng.close()
