"""One training step on the GPU through the C ABI (cocr_train_step / cocr_adamw_step; csrc/train.cu) against
 (1) the golden vectors of the UNMODIFIED reference modules in train() mode (model.py:129-145; dropout 0),
 (2) the torch port's autograd in float64 at the sequence length of BASELINE configs[3] (16 layers, 1536 px, C = 256),
 (3) the same with dropout 0.1 everywhere, the port replaying the kernels' own masks (cocr_dropout_mask),
 (4) torch.optim.AdamW.
Tolerance: fp32 kernels against a float64 checker; a gradient tensor must be within 2e-4 of max(its own largest entry,
1e-3 x the largest gradient entry of the model) - the floor covers tensors whose exact gradient is zero (key-projection
bias: softmax is invariant to it)."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR
from oracle import conformer_oracle as O
from oracle import torch_train_port as TP

pytestmark = pytest.mark.gpu

INDEX = json.load(open(os.path.join(GOLDEN_DIR, 'index_train.json')))
HP_KEYS = ('height', 'encoder_dim', 'num_encoder_layers', 'num_attention_heads', 'feed_forward_expansion_factor',
           'conv_expansion_factor', 'conv_kernel_size', 'half_step_residual', 'subsampling_conv_channels', 'subsampling_factor')


BF16_NOTE = """bf16 mode (tcgen05 operands rounded to 8 significant bits, fp32 accumulation and master copy): the tolerance is the
usual one of mixed-precision training - loss within 5e-3, every gradient tensor within 0.15 of max(its largest entry, 1e-3 x the
model's largest gradient entry ... 5e-2 x for the
key-projection bias, whose exact gradient is zero and whose computed one is the rounding noise of the attention backward pass),
and the whole gradient vector within 2 degrees of the reference's (cosine >= 0.9994) on the shallow golden cases; the 16-layer case is
rated against torch's own bf16 autocast (see test_train_step_bf16_default_spec_against_port)."""


def make_trainer(hp, C, device, dropout=0.0, **kw):
    from conformer_ocr_b200.train import Trainer
    return Trainer(num_classes=C, input_dropout_p=dropout, feed_forward_dropout_p=dropout, attention_dropout_p=dropout,
                   conv_dropout_p=dropout, device=device, **{k: hp[k] for k in HP_KEYS}, **kw)


def check_grads(tr, ref_grads, tol=2e-4, zero_floor=1e-3):
    """Every gradient tensor within `tol` of max(its own largest reference entry, zero_floor x the model's largest)."""
    keys = tr.trainable_keys()
    gmax = max(float(np.abs(ref_grads[k]).max()) for k in keys)
    worst = 0.0
    for k in keys:
        ref = np.asarray(ref_grads[k], dtype=np.float64)
        got = tr.view(k, 'grads').cpu().double().numpy()
        assert got.shape == ref.shape, k
        e = np.abs(got - ref).max() / max(np.abs(ref).max(), zero_floor * gmax)
        worst = max(worst, e)
        assert e <= tol, f'{k}: gradient rel err {e:.3e}'
    return worst


@pytest.mark.parametrize('name', sorted(INDEX))
def test_train_step_matches_reference_golden(cuda_device, name):
    m = INDEX[name]
    g = np.load(os.path.join(GOLDEN_DIR, name + '.npz'))
    hp = m['hparams']
    w = O.make_weights(hp, m['num_classes'], seed=m['weight_seed'])
    img, lens = O.make_lines(m['batch'], hp['height'], m['width'], seed=m['input_seed'], widths=m['widths'])
    tr = make_trainer(hp, m['num_classes'], cuda_device)
    assert sorted(tr.layout) == sorted(k for k in w)
    tr.load_state_dict(w)
    batch = {'image': torch.from_numpy(img), 'seq_lens': torch.from_numpy(lens), 'target': torch.from_numpy(g['targets']),
             'target_lens': torch.from_numpy(g['target_lens'])}
    n0 = tr.launch_count
    loss = tr.forward_backward(batch)
    assert tr.launch_count > n0
    assert abs(loss.item() - float(g['loss'])) <= 1e-5 * abs(float(g['loss']))
    worst = check_grads(tr, {k: g['grad/' + k] for k in tr.trainable_keys()})
    for k in tr.layout:
        if k.endswith(('running_mean', 'running_var')):
            assert np.abs(tr.view(k).cpu().numpy() - g['stat/' + k]).max() <= 1e-6, k
    # a second pass over the same batch gives the same gradients (atomics reorder sums: 1e-5), statistics move again
    g1 = tr.grads[:tr.n_trainable].clone()
    tr.forward_backward(batch)
    assert (tr.grads[:tr.n_trainable] - g1).abs().max().item() <= 1e-5 * g1.abs().max().item()
    print(f'{name}: loss {loss.item():.6f}, worst gradient rel err {worst:.2e}')
    tr.close()


def synthetic_batch(B, H, W, C, seed, widths=None):
    img, lens = O.make_lines(B, H, W, seed=seed, widths=widths)
    rng = np.random.RandomState(seed + 1)
    tl = rng.randint(20, 81, size=B).astype(np.int64)                     # SURVEY 8d cfg 4: target length U{20..80}
    tg = rng.randint(1, C, size=int(tl.sum())).astype(np.int64)
    return {'image': torch.from_numpy(img), 'seq_lens': torch.from_numpy(lens), 'target': torch.from_numpy(tg), 'target_lens': torch.from_numpy(tl)}


def port_grads(batch, w, hp, device, mask_fn=None, dropout=0.0):
    wt = {k: torch.from_numpy(v.copy()).to(device=device, dtype=torch.float64).requires_grad_(not k.endswith(('running_mean', 'running_var')))
          for k, v in w.items()}
    loss, _, _, stats = TP.train_step_loss(batch['image'].to(device).double(), batch['seq_lens'], batch['target'].to(device),
                                           batch['target_lens'], wt, hp, dropout=dropout, mask_fn=mask_fn)
    loss.backward()
    return loss.item(), {k: wt[k].grad.cpu().numpy() for k in TP.trainable_keys(wt)}, {k: v.cpu().numpy() for k, v in stats.items()}


@pytest.mark.parametrize('dropout', [0.0, 0.1])
def test_train_step_default_spec_against_port(cuda_device, dropout):
    """16 layers, d = 144, 1536 px (T = 384), C = 256, ragged widths; B = 3 keeps the float64 checker to seconds."""
    hp = dict(O.DEFAULT_HPARAMS)
    C, B, W = 256, 3, 1536
    w = O.make_weights(hp, C, seed=31)
    batch = synthetic_batch(B, hp['height'], W, C, seed=32, widths=[1536, 1400, 1111])
    tr = make_trainer(hp, C, cuda_device, dropout=dropout, seed=5)
    tr.load_state_dict(w)
    loss = tr.forward_backward(batch)
    seed = tr.cfg.seed
    prob = dropout

    def mask_fn(site, shape):
        return tr.dropout_mask(site, prob, int(np.prod(shape)), seed=seed).view(shape)

    ref_loss, ref_g, ref_stats = port_grads(batch, w, hp, cuda_device, mask_fn=mask_fn if dropout > 0 else None, dropout=dropout)
    assert abs(loss.item() - ref_loss) <= 2e-5 * abs(ref_loss), (loss.item(), ref_loss)
    worst = check_grads(tr, ref_g, tol=2e-3)     # 16 layers of fp32 accumulation against float64
    for k, v in ref_stats.items():
        assert np.abs(tr.view(k).cpu().numpy() - v).max() <= 1e-5, k
    if dropout > 0:
        mk = tr.dropout_mask(3, prob, 1 << 20, seed=seed)
        keep = (mk > 0).float().mean().item()
        assert abs(keep - (1 - prob)) < 2e-3 and abs(mk.max().item() - 1 / (1 - prob)) < 1e-6
    print(f'default spec, dropout {dropout}: loss {loss.item():.4f} (port {ref_loss:.4f}), worst gradient rel err {worst:.2e}')
    tr.close()


def flat_cosine(tr, ref_grads):
    a = np.concatenate([tr.view(k, 'grads').cpu().double().numpy().ravel() for k in tr.trainable_keys()])
    b = np.concatenate([np.asarray(ref_grads[k], dtype=np.float64).ravel() for k in tr.trainable_keys()])
    return float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b)))


@pytest.mark.parametrize('name', sorted(INDEX))
def test_train_step_bf16_matches_reference_golden(cuda_device, name):
    """The tensor-core path against the goldens of the unmodified reference modules (tolerance: BF16_NOTE)."""
    m = INDEX[name]
    g = np.load(os.path.join(GOLDEN_DIR, name + '.npz'))
    hp = m['hparams']
    w = O.make_weights(hp, m['num_classes'], seed=m['weight_seed'])
    img, lens = O.make_lines(m['batch'], hp['height'], m['width'], seed=m['input_seed'], widths=m['widths'])
    tr = make_trainer(hp, m['num_classes'], cuda_device, precision='bf16')
    tr.load_state_dict(w)
    batch = {'image': torch.from_numpy(img), 'seq_lens': torch.from_numpy(lens), 'target': torch.from_numpy(g['targets']),
             'target_lens': torch.from_numpy(g['target_lens'])}
    ref = {k: g['grad/' + k] for k in tr.trainable_keys()}
    for rep in range(3):            # plain launches, capture, replay of the captured step: same result every time
        loss = tr.forward_backward(batch)
        assert abs(loss.item() - float(g['loss'])) <= 5e-3 * abs(float(g['loss'])), (rep, loss.item())
        worst = check_grads(tr, ref, tol=0.15, zero_floor=0.05)
        cos = flat_cosine(tr, ref)
        assert cos >= 0.9994, (rep, cos)
    print(f'{name} bf16: loss {loss.item():.4f} (reference {float(g["loss"]):.4f}), worst gradient rel err {worst:.3f}, cosine {cos:.6f}')
    tr.close()


def test_train_step_bf16_default_spec_against_port(cuda_device):
    """16 layers at 1536 px with dropout 0.1 everywhere, the float64 port replaying the kernels' masks.  Rounding noise of bf16
    operands grows with depth (measured on B200, profiles/r02_train_bf16_noise.log: cosine 0.99998 / 0.9998 / 0.990 at 1 / 4 / 16
    layers), so the yardstick here is the reference's own mixed-precision mode: torch's bf16 autocast over the same operator
    sequence, masks and batch.  Measured: without dropout 0.990 against autocast's 0.958; with dropout 0.1 (this test) 0.961 - 0.984
    against 0.989 - at this depth the figure is one realisation of the rounding noise: it moves by a few hundredths with the batch,
    the masks, and even with the summation order of an fp32 reduction (a change of 1e-7 that flips bf16 roundings downstream), for
    either implementation.  Required: cosine >= 0.95, not more than 0.04 below autocast's, and every tensor with a non-negligible
    gradient within cosine 0.8.  The shallow cases (goldens, the long-line case) hold 0.9994."""
    hp = dict(O.DEFAULT_HPARAMS)
    C, B, W, prob = 256, 3, 1536, 0.1
    w = O.make_weights(hp, C, seed=31)
    batch = synthetic_batch(B, hp['height'], W, C, seed=32, widths=[1536, 1400, 1111])
    tr = make_trainer(hp, C, cuda_device, dropout=prob, seed=5, precision='bf16')
    tr.load_state_dict(w)
    loss = tr.forward_backward(batch)
    seed = tr.cfg.seed

    def mask_fn(site, shape):
        return tr.dropout_mask(site, prob, int(np.prod(shape)), seed=seed).view(shape)

    ref_loss, ref_g, _ = port_grads(batch, w, hp, cuda_device, mask_fn=mask_fn, dropout=prob)
    assert abs(loss.item() - ref_loss) <= 5e-3 * abs(ref_loss), (loss.item(), ref_loss)
    # per tensor: direction (cosine) rather than the largest single entry, which at layer 0 of 16 carries the noise of the whole chain
    gnorm = np.sqrt(sum(float((np.asarray(v, np.float64) ** 2).sum()) for v in ref_g.values()))
    worst = 1.0
    for k in tr.trainable_keys():
        r = np.asarray(ref_g[k], np.float64).ravel()
        if np.linalg.norm(r) < 1e-3 * gnorm:
            continue                                              # tensors with (near-)zero gradient: covered by the global cosine
        gk = tr.view(k, 'grads').cpu().double().numpy().ravel()
        ck = float(gk @ r / (np.linalg.norm(gk) * np.linalg.norm(r)))
        worst = min(worst, ck)
        assert ck >= 0.8, f'{k}: gradient cosine {ck:.4f}'        # measured minimum 0.92 (a bias gradient of layer 8)
    cos = flat_cosine(tr, ref_g)
    # torch autocast (bf16 matmuls, fp32 master weights) on the same batch and masks
    wt = {k: torch.from_numpy(v.copy()).to(cuda_device).requires_grad_(not k.endswith(('running_mean', 'running_var'))) for k, v in w.items()}
    with torch.autocast('cuda', dtype=torch.bfloat16):
        probits, olens, _ = TP.train_forward(batch['image'].to(cuda_device), batch['seq_lens'], wt, hp, dropout=prob, mask_fn=mask_fn)
    logp = torch.nn.functional.log_softmax(probits.float(), -1)
    tl = torch.nn.functional.ctc_loss(logp.transpose(0, 1), batch['target'].to(cuda_device), olens, batch['target_lens'], blank=0,
                                      reduction='sum', zero_infinity=True)
    tl.backward()
    keys = tr.trainable_keys()
    a = np.concatenate([wt[k].grad.double().cpu().numpy().ravel() for k in keys])
    b = np.concatenate([np.asarray(ref_g[k], dtype=np.float64).ravel() for k in keys])
    cos_torch = float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b)))
    print(f'default spec bf16, dropout {prob}: loss {loss.item():.3f} (port {ref_loss:.3f}), worst per-tensor gradient cosine {worst:.4f}, '
          f'cosine {cos:.6f} (torch bf16 autocast: {cos_torch:.6f})')
    assert cos >= 0.95 and cos >= cos_torch - 0.04, (cos, cos_torch)
    tr.close()


def test_training_bf16_tracks_fp32(cuda_device):
    """Six optimiser steps on one batch in both precisions from the same weights and masks: the loss curves stay within 2 %."""
    hp = dict(O.DEFAULT_HPARAMS, num_encoder_layers=4)
    C = 64
    batch = synthetic_batch(4, hp['height'], 768, C, seed=52)
    batch['target_lens'] = torch.clamp(batch['target_lens'], max=48)
    batch['target'] = batch['target'][:int(batch['target_lens'].sum())]
    curves = {}
    for prec in ('fp32', 'bf16'):
        tr = make_trainer(hp, C, cuda_device, dropout=0.1, lr=1e-3, weight_decay=1e-3, seed=3, precision=prec)
        tr.load_state_dict(O.make_weights(hp, C, seed=51))
        curves[prec] = [tr.step(batch).item() for _ in range(6)]
        tr.close()
    a, b = np.array(curves['fp32']), np.array(curves['bf16'])
    assert np.all(np.isfinite(b)) and b[-1] < 0.9 * b[0], curves
    assert np.abs(a - b).max() <= 0.02 * a.max(), curves


def test_adamw_matches_torch(cuda_device):
    hp = dict(O.DEFAULT_HPARAMS, num_encoder_layers=1)
    tr = make_trainer(hp, 32, cuda_device, lr=3e-3, weight_decay=1e-2)
    gen = torch.Generator(device='cpu').manual_seed(3)
    n = tr.n_trainable
    p0 = torch.randn(n, generator=gen)
    tr.params[:n] = p0.to(cuda_device)
    ref = torch.nn.Parameter(p0.clone().to(cuda_device))
    opt = torch.optim.AdamW([ref], lr=3e-3, weight_decay=1e-2)
    for step in range(3):
        gr = torch.randn(n, generator=gen) * (10.0 ** (step - 1))
        tr.grads[:n] = gr.to(cuda_device)
        ref.grad = gr.to(cuda_device)
        tr.optimizer_step()
        opt.step()
        assert (tr.params[:n] - ref.detach()).abs().max().item() <= 2e-6
    tr.close()


def test_training_reduces_the_loss(cuda_device):
    """A few steps on one batch with the reference's optimiser settings: the loss goes down, parameters stay finite."""
    hp = dict(O.DEFAULT_HPARAMS, num_encoder_layers=2)
    C = 48
    tr = make_trainer(hp, C, cuda_device, dropout=0.1, lr=1e-3, weight_decay=1e-3, seed=9)
    tr.load_state_dict(O.make_weights(hp, C, seed=41))
    batch = synthetic_batch(4, hp['height'], 512, C, seed=42)
    batch['target_lens'] = torch.clamp(batch['target_lens'], max=40)
    batch['target'] = batch['target'][:int(batch['target_lens'].sum())]
    losses = [tr.step(batch).item() for _ in range(8)]
    assert all(np.isfinite(losses)) and losses[-1] < 0.9 * losses[0], losses
    assert torch.isfinite(tr.params).all()
    tr.close()


def test_train_rejects_bad_arguments(cuda_device):
    hp = dict(O.DEFAULT_HPARAMS, num_encoder_layers=1)
    tr = make_trainer(hp, 16, cuda_device)
    with pytest.raises(ValueError):
        tr.load_state_dict({'decoder.weight': np.zeros((16, 144), np.float32)})
    batch = synthetic_batch(2, 64, 128, 16, seed=1)
    with pytest.raises(ValueError):
        tr.forward_backward(batch)                      # wrong line height
    tr.cfg.conv_dropout_p = 1.5
    with pytest.raises(ValueError):
        tr.forward_backward(synthetic_batch(2, hp['height'], 128, 16, seed=1))
    tr.close()


def test_recognition_model_training_step_interface(cuda_device):
    """The drop-in interface: RecognitionModel.train() -> training_step(batch, idx) (model.py:153-158) -> eval() -> forward sees
    the trained weights (parameters and BatchNorm statistics copied back into the module's state dict)."""
    from conformer_ocr_b200 import RecognitionModel
    hp = dict(O.DEFAULT_HPARAMS, num_encoder_layers=2)
    C = 40
    kw = dict(input_dropout_p=0.1, feed_forward_dropout_p=0.1, attention_dropout_p=0.1, conv_dropout_p=0.1)
    net = RecognitionModel(num_classes=C, precision='fp32', **hp, **kw)
    w = O.make_weights(hp, C, seed=61)
    net.nn.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()}, strict=False)
    batch = synthetic_batch(4, hp['height'], 512, C, seed=62)
    batch['target_lens'] = torch.clamp(batch['target_lens'], max=40)
    batch['target'] = batch['target'][:int(batch['target_lens'].sum())]
    x = batch['image'].to(cuda_device).squeeze(1).transpose(1, 2)
    with pytest.raises(RuntimeError):
        net.training_step(batch, 0)                                   # not in train mode
    before, _ = net(x, batch['seq_lens'])
    net.configure_training(lr=1e-3, weight_decay=1e-3, precision='bf16', seed=4)
    net.train()
    losses = [net.training_step(batch, i).item() for i in range(6)]
    assert losses[-1] < 0.9 * losses[0], losses
    net.eval()
    after, _ = net(x, batch['seq_lens'])
    assert (after - before).abs().max().item() > 1e-3                 # the inference engine was rebuilt from the trained weights
    sd = net.nn.state_dict()
    tsd = net._trainer.state_dict()
    for k, v in tsd.items():
        assert torch.equal(sd[k].cpu(), v.cpu()), k
    # BatchNorm running statistics moved (train mode) and are what eval mode now uses
    k = 'encoder.layers.0.sequential.2.module.sequential.5.running_mean'
    assert (sd[k].cpu() - torch.from_numpy(w[k])).abs().max().item() > 1e-4


def test_training_soak_over_changing_shapes(cuda_device):
    """45 optimiser steps cycling through three batch shapes (each shape: plain launches, capture, then graph replays; the
    staging buffers grow with the widest batch): losses stay finite and fall for every shape, parameters stay finite, and a
    fresh trainer replaying the same sequence lands on the same loss curve (within 2 %: the atomics of the bias / LayerNorm /
    depthwise sums reorder additions, and 45 Adam steps amplify that to 1e-2 in individual parameters - measured)."""
    hp = dict(O.DEFAULT_HPARAMS, num_encoder_layers=2)
    C = 48
    shapes = [(4, 512), (3, 768), (5, 384)]
    batches = []
    for i, (B, W) in enumerate(shapes):
        b = synthetic_batch(B, hp['height'], W, C, seed=70 + i)
        b['target_lens'] = torch.clamp(b['target_lens'], max=W // 16)
        b['target'] = b['target'][:int(b['target_lens'].sum())]
        batches.append(b)
    finals = []
    for run in range(2):
        tr = make_trainer(hp, C, cuda_device, dropout=0.0, lr=5e-4, weight_decay=1e-3, seed=11, precision='bf16')
        tr.load_state_dict(O.make_weights(hp, C, seed=71))
        losses = {i: [] for i in range(len(shapes))}
        for step in range(45):
            i = step % len(shapes)
            losses[i].append(tr.step(batches[i]).item())
        for i, ls in losses.items():
            assert all(np.isfinite(ls)) and ls[-1] < 0.9 * ls[0], (i, ls)
        assert torch.isfinite(tr.params).all()
        finals.append(np.array([ls[-1] for ls in losses.values()]))
        tr.close()
    assert np.abs(finals[0] - finals[1]).max() <= 0.02 * finals[0].max(), finals


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_step_graphs_survive_alternating_widths(cuda_device, precision):
    """One trainer fed batches of two widths in turns (plain launches, capture, replays of both step graphs, back and forth):
    every visit reproduces the loss and the gradient buffer of a FRESH trainer's first (plain-launch) step on that batch.
    Dropout 0 and no optimiser step, so nothing but the BatchNorm running statistics (unused in train mode) changes between
    visits.  Guards the addresses a captured step bakes in - the relative positional table of its frame count above all:
    one table per frame count lives as long as the handle (csrc/train.cu ensure_pe), so a replay never reads a table that
    was freed or rewritten for another width.  Tolerance: the fp32 atomics of the bias / LayerNorm sums reorder additions."""
    hp = dict(O.DEFAULT_HPARAMS, num_encoder_layers=2)
    C = 40
    w = O.make_weights(hp, C, seed=91)
    batches = {}
    for name, (B, W, seed) in {'A': (3, 768, 92), 'B': (3, 512, 93)}.items():
        b = synthetic_batch(B, hp['height'], W, C, seed=seed)
        b['target_lens'] = torch.clamp(b['target_lens'], max=W // 16)
        b['target'] = b['target'][:int(b['target_lens'].sum())]
        batches[name] = b
    fresh = {}
    for name, b in batches.items():
        tr = make_trainer(hp, C, cuda_device, precision=precision)
        tr.load_state_dict(w)
        fresh[name] = (tr.forward_backward(b).item(), tr.grads[:tr.n_trainable].clone())
        tr.close()
    tr = make_trainer(hp, C, cuda_device, precision=precision)
    tr.load_state_dict(w)
    for visit, name in enumerate('AAABBBAABBAB'):
        loss = tr.forward_backward(batches[name]).item()
        ref_loss, ref_g = fresh[name]
        assert abs(loss - ref_loss) <= 1e-4 * abs(ref_loss), (visit, name, loss, ref_loss)
        err = ((tr.grads[:tr.n_trainable] - ref_g).abs().max() / ref_g.abs().max()).item()
        assert err <= 1e-3, (visit, name, err)
    tr.close()


def test_train_step_long_lines_bf16_against_fp32_path(cuda_device):
    """Lines of 2304 px (T = 576 > 512: the softmax kernel's streaming path, several key tiles and contraction chunks per GEMM, a
    ragged batch) and an odd class count: the tensor-core mode against the fp32 CUDA-core mode on the same batch, one layer."""
    hp = dict(O.DEFAULT_HPARAMS, num_encoder_layers=1)
    C, W = 51, 2304
    w = O.make_weights(hp, C, seed=81)
    batch = synthetic_batch(2, hp['height'], W, C, seed=82, widths=[2304, 1999])
    grads = {}
    for prec in ('fp32', 'bf16'):
        tr = make_trainer(hp, C, cuda_device, precision=prec)
        tr.load_state_dict(w)
        loss = tr.forward_backward(batch).item()
        grads[prec] = (loss, tr.grads[:tr.n_trainable].double().cpu().numpy(), {k: tr.view(k, 'grads').cpu().double().numpy() for k in tr.trainable_keys()})
        tr.close()
    (l32, g32, k32), (l16, g16, k16) = grads['fp32'], grads['bf16']
    assert abs(l16 - l32) <= 5e-3 * abs(l32), (l16, l32)
    cos = float(g16 @ g32 / (np.linalg.norm(g16) * np.linalg.norm(g32)))
    assert cos >= 0.9994, cos
    gmax = max(np.abs(v).max() for v in k32.values())
    for k, ref in k32.items():
        e = np.abs(k16[k] - ref).max() / max(np.abs(ref).max(), 0.05 * gmax)
        assert e <= 0.15, f'{k}: {e:.3f}'


@pytest.mark.parametrize('seed,site,p,n', [(0, 0, 0.1, 1000), (12345678901234567, 17, 0.1, 4099), (2 ** 63 + 5, 255, 0.5, 70001), (7, 3, 0.0, 64)])
def test_dropout_masks_are_philox(cuda_device, seed, site, p, n):
    """The masks of the training step are a pure function of (seed, site, element): cocr_dropout_mask against the numpy
    restatement of Philox-4x32-10 in the oracle, bit for bit."""
    hp = dict(O.DEFAULT_HPARAMS, num_encoder_layers=1)
    tr = make_trainer(hp, 16, cuda_device)
    got = tr.dropout_mask(site, p, n, seed=seed).cpu().numpy()
    want = TP.philox_dropout_factors(seed, site, p, n)
    assert got.dtype == want.dtype and np.array_equal(got, want)
    tr.close()


def test_train_step_benchmark_shape_bf16_against_fp32_path(cuda_device):
    """BASELINE configs[3] exactly as `bench.py --config 4` runs it (32 lines of 96 x 1536 px, 256 classes, 16 layers; the
    weight-gradient contractions are split over 28 slices at this size, 18 at the 3-line shape of the other default-spec tests):
    the tensor-core mode against the fp32 CUDA-core mode (pinned to the reference's gradients to 1e-5 at the small shapes) on the
    same batch, dropout off.  Loss within 5e-3; the whole gradient vector within 14 degrees (cosine >= 0.97; 0.990 was measured
    against a float64 port at 16 layers and 3 lines, torch's own bf16 autocast reaches 0.958 there - profiles/r02_train_bf16_noise.log).
    Measured on B200: loss 52 459.8 (fp32) / 52 511.9 (bf16), cosine 0.9958."""
    hp = dict(O.DEFAULT_HPARAMS)
    B, W, C = 32, 1536, 256
    w = O.make_weights(hp, C, seed=0)
    batch = synthetic_batch(B, hp['height'], W, C, seed=1234)
    out = {}
    for prec in ('fp32', 'bf16'):
        tr = make_trainer(hp, C, cuda_device, precision=prec)
        tr.load_state_dict(w)
        loss = tr.forward_backward(batch).item()
        out[prec] = (loss, tr.grads[:tr.n_trainable].double().cpu())
        tr.close()
    (l32, g32), (l16, g16) = out['fp32'], out['bf16']
    assert np.isfinite(l32) and np.isfinite(l16) and torch.isfinite(g16).all()
    cos = float((g32 * g16).sum() / (g32.norm() * g16.norm()))
    print(f'training step at the benchmark shape: loss fp32 {l32:.3f} bf16 {l16:.3f}, gradient cosine {cos:.6f}')
    assert abs(l16 - l32) <= 5e-3 * abs(l32), (l16, l32)
    assert cos >= 0.97, cos
