"""Training-side tail of the path (model.py:136-142): log_softmax and CTCLoss(reduction='sum', zero_infinity=True) on
the GPU through the C ABI, against the numpy oracle (oracle/conformer_oracle.py: ctc_loss_sum, log_softmax) and against
torch's own CTCLoss + autograd for the gradient with respect to the decoder outputs.

Tolerance: the checker is torch's CTCLoss + autograd in float64 on the host (exact to ~1e-12).  The kernels work in fp32
log space like torch's own CUDA implementation (the arithmetic the reference actually runs): alpha and beta reach minus a few
hundred to a few thousand, every log-add rounds at 6e-8 of that, so a posterior carries an absolute error of order 1e-6 x the
line's loss.  The gradient (entries in [-1, 1]) must stay within max(3e-5, 2e-6 x largest per-line loss) of the float64 result
AND within twice the distance torch's fp32 CUDA CTCLoss has from it (measured: 2.5e-3 against torch's 3.0e-3 at B=32, T=384);
the loss within 1e-5 relative."""
import numpy as np
import pytest
import torch

from oracle import conformer_oracle as O

pytestmark = pytest.mark.gpu


def engine(device, C):
    from conformer_ocr_b200.engine import Engine
    return Engine(height=8, encoder_dim=8, num_layers=0, num_heads=1, ff_expansion=1, conv_expansion=2, conv_kernel_size=3,
                  half_step_residual=True, subsampling_conv_channels=1, subsampling_factor=4, num_classes=0, device=device)


def make_case(B, T, C, lmin, lmax, seed, ragged=True, repeats=False):
    rng = np.random.RandomState(seed)
    logits = rng.normal(0, 2.0, size=(B, T, C)).astype(np.float32)
    tl = rng.randint(lmin, lmax + 1, size=B)
    il = rng.randint(max(1, T // 2), T + 1, size=B) if ragged else np.full(B, T)
    tg = []
    for n in tl:
        row = rng.randint(1, C, size=n)
        if repeats and n > 3:
            row[1] = row[0]
            row[3] = row[2]
        tg.append(row)
    return logits, np.concatenate(tg) if len(tg) and sum(tl) else np.zeros(0, np.int64), il.astype(np.int64), tl.astype(np.int64)


def torch_reference(logits, targets, il, tl, device=None, dtype=torch.float64):
    """model.py:136-142 with torch itself: float64 on the host by default (the checker), fp32 on `device` for comparison."""
    z = torch.from_numpy(logits).to(dtype)
    if device is not None:
        z = z.to(device)
    z.requires_grad_(True)
    lp = torch.nn.functional.log_softmax(z, dim=-1)
    crit = torch.nn.CTCLoss(reduction='sum', zero_infinity=True)
    tg = torch.from_numpy(targets)
    loss = crit(lp.transpose(0, 1), tg.to(z.device), torch.from_numpy(il), torch.from_numpy(tl))
    loss.backward()
    return float(loss.detach()), z.grad.detach().cpu().numpy()


@pytest.mark.parametrize('B,T,C,lmin,lmax,ragged,repeats', [
    (4, 50, 20, 3, 12, True, False),
    (32, 384, 256, 20, 80, False, False),          # BASELINE configs[3]: batch 32, W = 1536 -> T = 384, targets U{20..80}
    (8, 100, 30, 0, 5, True, True),                # empty targets, repeated labels
    (6, 40, 12, 25, 45, True, True),               # some targets longer than the line can emit: infinite loss -> 0
    (3, 1024, 10000, 50, 300, True, False),        # configs[4] geometry (T = 1024, 10 000 classes)
    (2, 7, 5, 1, 3, False, True),
])
def test_ctc_loss_and_gradient_match_torch_and_oracle(B, T, C, lmin, lmax, ragged, repeats, cuda_device):
    logits, targets, il, tl = make_case(B, T, C, lmin, lmax, seed=B * 1000 + T, ragged=ragged, repeats=repeats)
    eng = engine(cuda_device, C)
    loss, per_line, grad = eng.ctc_loss(torch.from_numpy(logits).to(cuda_device), targets, il, tl, want_grad=True)
    want, want_grad = torch_reference(logits, targets, il, tl)
    got = float(loss)
    assert abs(got - want) <= 1e-5 * max(1.0, abs(want)), (got, want)
    assert abs(float(per_line.sum()) - got) <= 1e-5 * max(1.0, abs(got))
    g = grad.cpu().numpy()
    assert np.isfinite(g).all()
    atol = max(3e-5, 2e-6 * float(per_line.max()))
    err = float(np.abs(g - want_grad).max())
    _, cuda_grad = torch_reference(logits, targets, il, tl, device=cuda_device, dtype=torch.float32)
    print(f'B={B} T={T} C={C}: loss {got:.4f} (float64 {want:.4f}); grad max err {err:.2e} (torch fp32 CUDA: '
          f'{float(np.abs(cuda_grad - want_grad).max()):.2e}; bound {atol:.1e})')
    assert err <= atol
    assert err <= 2.0 * float(np.abs(cuda_grad - want_grad).max()) + 3e-5
    if B * T * C <= 4_000_000:                      # numpy oracle (float64 alpha recursion) on the small cases
        lp = O.log_softmax(logits.astype(np.float64))
        ref = O.ctc_loss_sum(lp.transpose(1, 0, 2), targets, il, tl)
        assert abs(got - ref) <= 1e-5 * max(1.0, abs(ref)), (got, ref)
    # bit-reproducible: fixed summation order, no atomics
    loss2, _, grad2 = eng.ctc_loss(torch.from_numpy(logits).to(cuda_device), targets, il, tl, want_grad=True)
    assert torch.equal(loss, loss2) and torch.equal(grad, grad2)


def test_log_softmax_matches_oracle(cuda_device):
    rng = np.random.RandomState(3)
    for shape in ((5, 7, 11), (2, 384, 256), (3, 33, 10000), (1, 1, 2)):
        x = (rng.normal(0, 3, size=shape) + rng.normal(0, 20, size=shape[:-1] + (1,))).astype(np.float32)
        eng = engine(cuda_device, shape[-1])
        got = eng.log_softmax(torch.from_numpy(x).to(cuda_device)).cpu().numpy()
        ref = O.log_softmax(x.astype(np.float64))
        np.testing.assert_allclose(got, ref, rtol=1e-6, atol=2e-5)
        np.testing.assert_allclose(np.exp(got.astype(np.float64)).sum(-1), 1.0, atol=1e-4)


def test_recognition_model_step_matches_reference_semantics(cuda_device):
    """RecognitionModel._step (model.py:129-145): loss from the module's own logits equals torch's CTCLoss on
    log_softmax of the same logits; output_lens = calc_length(seq_lens)."""
    from conformer_ocr_b200.model import RecognitionModel
    from conformer_ocr_b200 import synthetic as S
    hp = dict(O.DEFAULT_HPARAMS)
    hp['num_encoder_layers'] = 2
    C, B, W = 40, 4, 400
    w = S.make_weights(hp, C, seed=6)
    net = RecognitionModel(num_classes=C, **hp)
    net.nn.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()}, strict=False)
    net = net.to(cuda_device)
    img, lens = S.make_lines(B, hp['height'], W, seed=7, widths=[400, 390, 300, 120])
    rng = np.random.RandomState(8)
    tl = np.array([20, 10, 5, 0])
    tg = rng.randint(1, C, size=int(tl.sum()))
    batch = dict(image=torch.from_numpy(img).to(cuda_device), target=torch.from_numpy(tg), seq_lens=torch.from_numpy(lens),
                 target_lens=torch.from_numpy(tl))
    out = net._step(batch, want_grad=True)
    assert out['output_lens'].tolist() == [100, 98, 75, 30]
    want, want_grad = torch_reference(out['probits'].cpu().numpy(), tg.astype(np.int64), out['output_lens'].numpy().astype(np.int64),
                                      tl.astype(np.int64))
    assert abs(float(out['loss']) - want) <= 1e-5 * abs(want)
    np.testing.assert_allclose(out['grad_probits'].cpu().numpy(), want_grad, atol=max(3e-5, 2e-6 * float(out['loss_per_line'].max())), rtol=0)
    lp, _ = net.validation_logprobs(batch['image'].squeeze(1).transpose(1, 2), batch['seq_lens'])
    np.testing.assert_allclose(lp.cpu().numpy(), O.log_softmax(out['probits'].cpu().numpy().astype(np.float64)), atol=2e-5, rtol=1e-6)


def test_decoder_backward_and_training_step_match_autograd(cuda_device):
    """decoder_training_step: loss, d loss / d decoder.{weight,bias} and d loss / d encoder outputs equal torch autograd
    through nn.Linear + log_softmax + CTCLoss on the same encoder outputs (fp32, TF32 off)."""
    from conformer_ocr_b200.model import RecognitionModel
    from conformer_ocr_b200 import synthetic as S
    hp = dict(O.DEFAULT_HPARAMS)
    hp['num_encoder_layers'] = 2
    C, B, W = 70, 6, 600
    w = S.make_weights(hp, C, seed=16)
    net = RecognitionModel(num_classes=C, **hp)
    net.nn.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()}, strict=False)
    net = net.to(cuda_device)
    img, lens = S.make_lines(B, hp['height'], W, seed=17, widths=[600, 600, 580, 400, 333, 90])
    rng = np.random.RandomState(18)
    tl = np.array([30, 25, 12, 8, 3, 1])
    tg = rng.randint(1, C, size=int(tl.sum()))
    batch = dict(image=torch.from_numpy(img).to(cuda_device), target=torch.from_numpy(tg), seq_lens=torch.from_numpy(lens),
                 target_lens=torch.from_numpy(tl))
    out = net.decoder_training_step(batch)
    enc, _ = net.engine_for(cuda_device).encode(batch['image'].squeeze(1).transpose(1, 2), batch['seq_lens'])
    tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        e = enc.double().requires_grad_(True)
        lin = torch.nn.Linear(hp['encoder_dim'], C).to(cuda_device).double()
        with torch.no_grad():
            lin.weight.copy_(torch.from_numpy(w['decoder.weight']))
            lin.bias.copy_(torch.from_numpy(w['decoder.bias']))
        z = lin(e)
        loss = torch.nn.CTCLoss(reduction='sum', zero_infinity=True)(
            torch.nn.functional.log_softmax(z, -1).transpose(0, 1), batch['target'].to(cuda_device), out['output_lens'], batch['target_lens'])
        loss.backward()
    finally:
        torch.backends.cuda.matmul.allow_tf32 = tf32
    np.testing.assert_allclose(out['probits'].cpu().numpy(), z.detach().cpu().numpy(), rtol=1e-4, atol=1e-4)
    assert abs(float(out['loss']) - float(loss.detach())) <= 2e-5 * abs(float(loss.detach()))
    scale = float(lin.weight.grad.abs().max())
    np.testing.assert_allclose(out['grad_decoder_weight'].cpu().numpy(), lin.weight.grad.cpu().numpy(), rtol=0, atol=2e-4 * scale)
    np.testing.assert_allclose(out['grad_decoder_bias'].cpu().numpy(), lin.bias.grad.cpu().numpy(), rtol=0, atol=2e-4 * float(lin.bias.grad.abs().max()))
    np.testing.assert_allclose(out['grad_encoder_outputs'].cpu().numpy(), e.grad.cpu().numpy(), rtol=0, atol=2e-4 * float(e.grad.abs().max()))
    # single process: the all-reduced values equal the local ones
    assert torch.equal(out['loss_allreduced'], out['loss']) and torch.equal(out['grad_decoder_weight_allreduced'], out['grad_decoder_weight'])
    again = net.decoder_training_step(batch)
    assert torch.equal(again['grad_decoder_weight'], out['grad_decoder_weight'])       # bit-reproducible
