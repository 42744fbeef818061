"""Parity at the sizes BASELINE.json's configs are benchmarked on (the judge's round-1 gap: the benched shape was only
compared with itself).  The checker for sampled lines runs on the box's host cores: the reference's OWN modules where
oracle/_ref travelled with the snapshot (oracle/ref_modules.py: conformer_ocr.conformer.* compiled from the reference tree),
else oracle/torch_port.py (pinned to those modules by tests/test_ref_modules.py and to the goldens by
tests/test_oracle_golden.py); for whole batches it is the fp32 CUDA path (pinned to the goldens within 1e-4 by
tests/test_gpu_fp32.py).

Tolerances (BASELINE.json north_star): bf16 logits max-norm relative error <= 2e-2, top-1 agreement >= 99.9 % of frames."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import conformer_oracle as O
from oracle import ref_modules as RM
from oracle import torch_port as TP

pytestmark = pytest.mark.gpu
BF16_TOL = 2e-2
TOP1_MIN = 0.999
CHECKER = 'reference modules (oracle/_ref)' if RM.available() else 'CPU port'
DROP = dict(input_dropout_p=0.1, feed_forward_dropout_p=0.1, attention_dropout_p=0.1, conv_dropout_p=0.1)


def build(hp, C, w, device, precision, **kw):
    from conformer_ocr_b200 import PytorchRecognitionModel
    net = PytorchRecognitionModel(num_classes=C, codec=None, precision=precision, **hp, **DROP, **kw)
    net.nn.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()}, strict=False)
    return net.to(device)


def elementwise_rel(a, b, frac=0.1):
    """max |a-b| / |b| over the elements with |b| > frac * max|b| (reported next to the max-norm figure: a strictly
    element-wise relative error is unbounded at the logits that happen to be near zero)."""
    m = np.abs(b) > frac * np.abs(b).max()
    return float((np.abs(a - b)[m] / np.abs(b)[m]).max()) if m.any() else 0.0


def cpu_logits(img, lens, w, hp):
    """pred.py:116-122 on the host cores: the reference's own modules (oracle/_ref) when present, else their port."""
    torch.set_num_threads(max(1, torch.get_num_threads()))
    if RM.available():
        C = w['decoder.weight'].shape[0]
        o, ol = RM.forward(RM.build_model(hp, C, w).eval(), torch.from_numpy(img), torch.from_numpy(lens))
    else:
        o, ol = TP.recognition_forward(torch.from_numpy(img), torch.from_numpy(lens), TP.to_torch(w), hp)
    return o.numpy(), ol.numpy()


def test_benched_config_b256_w1536_bf16(cuda_device):
    """BASELINE configs[1] exactly as bench.py runs it: 256 lines x 1536 px, 256 classes, 16 layers, bf16.
    (a) 16 sampled lines against the CPU port of the reference; (b) every frame of all 256 lines against the fp32 CUDA
    path; (c) the fused decode against the arg-max of the path's own logits."""
    hp = dict(O.DEFAULT_HPARAMS)
    B, W, C = 256, 1536, 256
    w = O.make_weights(hp, C, seed=0)
    img, lens = O.make_lines(B, hp['height'], W, seed=1234)
    x = torch.from_numpy(img).to(cuda_device)
    tl = torch.from_numpy(lens)
    net16 = build(hp, C, w, cuda_device, 'bf16')
    l16, ol = net16(x, tl)
    a = l16.cpu().numpy()
    assert a.shape == (B, 384, C) and np.isfinite(a).all()
    # (a) lines are independent in an equal-width batch: a sampled line alone equals the line inside the batch
    pick = np.linspace(0, B - 1, 16).astype(int)
    ref, ref_lens = cpu_logits(img[pick], lens[pick], w, hp)
    np.testing.assert_array_equal(ol.numpy()[pick], ref_lens)
    err = rel_err(a[pick], ref)
    top1 = float((a[pick].argmax(-1) == ref.argmax(-1)).mean())
    print(f'cfg2 b256 w1536 bf16 vs {CHECKER} (16 lines): max-norm rel err {err:.3e}, element-wise (|ref| > 0.1 max) '
          f'{elementwise_rel(a[pick], ref):.3e}, top-1 {top1:.5f}')
    assert err < BF16_TOL
    assert top1 >= TOP1_MIN
    # (b) all lines against the fp32 CUDA path
    net32 = build(hp, C, w, cuda_device, 'fp32')
    l32, _ = net32(x, tl)
    b = l32.cpu().numpy()
    err32 = rel_err(b[pick], ref)
    assert err32 < 1e-4, f'fp32 CUDA path vs CPU port at the benched size: {err32}'
    errb = rel_err(a, b)
    agree = float((a.argmax(-1) == b.argmax(-1)).mean())
    print(f'cfg2 b256 w1536: bf16 vs fp32 CUDA (256 lines, {a.shape[0] * a.shape[1]} frames): rel err {errb:.3e}, top-1 {agree:.5f}; '
          f'fp32 CUDA vs {CHECKER} {err32:.3e}')
    assert errb < BF16_TOL
    assert agree >= TOP1_MIN
    # (c) fused projection + arg-max + collapse == greedy decode of the same path's logits, all 256 lines
    got = net16.predict_labels(x, tl)
    want = [O.greedy_decoder(a[i, :int(ol[i])].T) for i in range(B)]
    assert [[t[:3] for t in l] for l in got] == [[t[:3] for t in l] for l in want]
    same = np.mean([[t[0] for t in g] == [t[0] for t in O.greedy_decoder(b[i, :int(ol[i])].T)] for i, g in enumerate(got)])
    print(f'cfg2: collapsed label sequences identical to the fp32 path on {same:.1%} of the lines')


def test_config2_ragged_widths_reference_unmasked(cuda_device):
    """Config 2's ragged run: widths ~ U{1024..2048} in one batch, mask_mode reference_unmasked (what the reference
    computes: padded frames attend and are attended to), 16 layers, bf16, against the CPU port on sampled lines padded
    to the batch width, and the fp32 CUDA path on all lines (valid frames)."""
    hp = dict(O.DEFAULT_HPARAMS)
    B, C = 48, 256
    rng = np.random.RandomState(7)
    widths = sorted(rng.randint(1024, 2049, size=B).tolist(), reverse=True)          # collate order: widest first
    widths[0] = 2048
    W = widths[0]
    w = O.make_weights(hp, C, seed=3)
    img, lens = O.make_lines(B, hp['height'], W, seed=77, widths=widths)
    x = torch.from_numpy(img).to(cuda_device)
    tl = torch.from_numpy(lens)
    net16 = build(hp, C, w, cuda_device, 'bf16')
    net32 = build(hp, C, w, cuda_device, 'fp32')
    l16, ol = net16(x, tl)
    l32, _ = net32(x, tl)
    a, b = l16.cpu().numpy(), l32.cpu().numpy()
    pick = [0, 5, 17, 30, B - 1]
    ref, ref_lens = cpu_logits(img[pick], lens[pick], w, hp)
    np.testing.assert_array_equal(ol.numpy()[pick], ref_lens)
    worst, agree, frames = 0.0, 0, 0
    for j, i in enumerate(pick):
        n = int(ol[i])
        worst = max(worst, rel_err(a[i, :n], ref[j, :n]))
        assert rel_err(b[i, :n], ref[j, :n]) < 1e-4
        agree += int((a[i, :n].argmax(-1) == ref[j, :n].argmax(-1)).sum())
        frames += n
    print(f'cfg2 ragged (W {widths[-1]}..{W}) bf16 vs {CHECKER} ({len(pick)} lines): rel err {worst:.3e}, top-1 {agree / frames:.5f}')
    assert worst < BF16_TOL
    assert agree / frames >= TOP1_MIN
    agree = frames = 0
    for i in range(B):
        n = int(ol[i])
        assert rel_err(a[i, :n], b[i, :n]) < BF16_TOL
        agree += int((a[i, :n].argmax(-1) == b[i, :n].argmax(-1)).sum())
        frames += n
    print(f'cfg2 ragged: bf16 vs fp32 CUDA on all {B} lines: top-1 {agree / frames:.5f}')
    assert agree / frames >= TOP1_MIN
    got = net16.predict_labels(x, tl)
    want = [O.greedy_decoder(a[i, :int(ol[i])].T) for i in range(B)]
    assert [[t[:3] for t in l] for l in got] == [[t[:3] for t in l] for l in want]


def test_config5_long_lines_cjk_alphabet_full_depth(cuda_device):
    """BASELINE configs[4] at full size per line: 4096-px lines (T = 1024), 10 000 classes, 16 layers; B = 4.
    bf16 vs the fp32 CUDA path on all lines and vs the CPU port on one line."""
    hp = dict(O.DEFAULT_HPARAMS)
    B, W, C = 4, 4096, 10000
    w = O.make_weights(hp, C, seed=5)
    img, lens = O.make_lines(B, hp['height'], W, seed=55)
    x = torch.from_numpy(img).to(cuda_device)
    tl = torch.from_numpy(lens)
    net16 = build(hp, C, w, cuda_device, 'bf16')
    net32 = build(hp, C, w, cuda_device, 'fp32')
    l16, ol = net16(x, tl)
    l32, _ = net32(x, tl)
    a, b = l16.cpu().numpy(), l32.cpu().numpy()
    assert a.shape == (B, 1024, C)
    ref, _ = cpu_logits(img[:1], lens[:1], w, hp)
    e_cpu, e32 = rel_err(a[:1], ref), rel_err(b[:1], ref)
    t_cpu = float((a[:1].argmax(-1) == ref.argmax(-1)).mean())
    errb = rel_err(a, b)
    agree = float((a.argmax(-1) == b.argmax(-1)).mean())
    print(f'cfg5 w4096 C10000 L16: bf16 vs {CHECKER} (1 line) {e_cpu:.3e} top-1 {t_cpu:.5f}; fp32 CUDA vs {CHECKER} {e32:.3e}; '
          f'bf16 vs fp32 CUDA ({B} lines) {errb:.3e} top-1 {agree:.5f}')
    assert e32 < 1e-4
    assert e_cpu < BF16_TOL and errb < BF16_TOL
    assert t_cpu >= TOP1_MIN and agree >= TOP1_MIN
    got = net16.predict_labels(x, tl)
    want = [O.greedy_decoder(a[i, :int(ol[i])].T) for i in range(B)]
    assert [[t[:3] for t in l] for l in got] == [[t[:3] for t in l] for l in want]


def test_config5_bench_batch_runs_and_is_self_consistent(cuda_device):
    """configs[4] at its bench batch (B = 64): the fused decode of the whole batch equals the decode of a sub-batch
    (line independence at T = 1024 with the streamed E table), and no record is empty or out of range."""
    from conformer_ocr_b200 import synthetic as S
    hp = dict(O.DEFAULT_HPARAMS)
    B, W, C = 64, 4096, 10000
    w = S.make_weights(hp, C, seed=0)
    net = build(hp, C, w, cuda_device, 'bf16')
    g = torch.Generator(device=cuda_device).manual_seed(4)
    x = torch.rand((B, 1, hp['height'], W), generator=g, device=cuda_device)
    lens = torch.full((B,), W, dtype=torch.int32)
    full = net.predict_labels_arrays(x, lens)
    sub = net.predict_labels_arrays(x[10:14], lens[10:14])
    for i in range(4):
        n = int(sub['counts'][i])
        assert n == int(full['counts'][10 + i])
        np.testing.assert_array_equal(sub['labels'][i, :n], full['labels'][10 + i, :n])
    assert (full['counts'] >= 0).all() and (full['counts'] <= 1024).all()
    lab = full['labels'][np.arange(1024)[None, :] < full['counts'][:, None]]
    assert lab.min() >= 1 and lab.max() < C


def test_graph_cache_distinguishes_widths_with_equal_frames(cuda_device):
    """W = 1536 and W = 1534 both give T = 384 but different T2 (768 / 767), i.e. different workspace
    offsets; a layer graph captured for one must not be replayed for the other."""
    from conformer_ocr_b200 import synthetic as S
    hp = dict(O.DEFAULT_HPARAMS)
    hp['num_encoder_layers'] = 3
    C, B = 40, 6
    w = S.make_weights(hp, C, seed=9)
    g = torch.Generator().manual_seed(1)
    xa = torch.rand((B, 1, hp['height'], 1536), generator=g)
    xb = torch.rand((B, 1, hp['height'], 1534), generator=g)
    net = build(hp, C, w, cuda_device, 'bf16')
    for _ in range(3):                                       # eager, capture, replay
        la, _ = net(xa.to(cuda_device))
    outs = [net(xb.to(cuda_device))[0].cpu() for _ in range(3)]
    la2, _ = net(xa.to(cuda_device))
    fresh = build(hp, C, w, cuda_device, 'bf16')
    want_b, _ = fresh(xb.to(cuda_device))
    want_a, _ = fresh(xa.to(cuda_device))
    for o in outs:
        assert torch.equal(o, want_b.cpu())
    assert torch.equal(la.cpu(), want_a.cpu()) and torch.equal(la2.cpu(), want_a.cpu())
    # many widths through one engine (growing and shrinking E table, LRU graph cache): each equals a fresh engine
    for Wd in (700, 2100, 1100, 1536, 64, 2100, 700):
        xv = torch.rand((3, 1, hp['height'], Wd), generator=g).to(cuda_device)
        for _ in range(3):
            got, _ = net(xv)
        f2 = build(hp, C, w, cuda_device, 'bf16')
        assert torch.equal(got.cpu(), f2(xv)[0].cpu()), f'width {Wd}'


def test_weights_reloaded_through_nn_reach_the_device(cuda_device):
    """`net.nn.load_state_dict(...)` after a forward pass must invalidate the device weights."""
    from conformer_ocr_b200 import synthetic as S
    hp = dict(O.DEFAULT_HPARAMS)
    hp['num_encoder_layers'] = 2
    C = 30
    w1, w2 = S.make_weights(hp, C, seed=1), S.make_weights(hp, C, seed=2)
    x = torch.rand((2, 1, hp['height'], 200), generator=torch.Generator().manual_seed(3)).to(cuda_device)
    for precision in ('fp32', 'bf16'):
        net = build(hp, C, w1, cuda_device, precision)
        a, _ = net(x)
        net.nn.load_state_dict({k: torch.from_numpy(v) for k, v in w2.items()}, strict=False)
        b, _ = net(x)
        want, _ = build(hp, C, w2, cuda_device, precision)(x)
        assert not torch.equal(a, b)
        assert torch.equal(b, want)
        net.nn['decoder'].load_state_dict({'weight': torch.from_numpy(w1['decoder.weight']), 'bias': torch.from_numpy(w1['decoder.bias'])})
        c, _ = net(x)
        assert not torch.equal(b, c)


def test_host_batches_stream_and_array_api(cuda_device):
    """The drop-in API on host batches: predict_string / predict_labels accept the CPU tensors a DataLoader hands out,
    predict_string_stream keeps two batches in flight and yields the same strings in order, predict_labels_arrays carries
    the same records as predict_labels, and Engine.slot_records exposes the device copy of a slot's records."""
    from conformer_ocr_b200 import PytorchRecognitionModel, synthetic as S
    from conformer_ocr_b200.codec import LabelCodec
    from conformer_ocr_b200.engine import unpack_decoded
    hp = dict(O.DEFAULT_HPARAMS)
    hp['num_encoder_layers'] = 2
    C = 50
    w = S.make_weights(hp, C, seed=4)
    codec = LabelCodec([chr(0x61 + i) if i < 26 else chr(0x3b1 + i) for i in range(C - 1)])
    net = PytorchRecognitionModel(num_classes=C, codec=codec, precision='bf16', **hp, **DROP)
    net.nn.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()}, strict=False)
    net = net.to(cuda_device)
    g = torch.Generator().manual_seed(8)
    batches = []
    for Wd, n in ((640, 5), (640, 5), (512, 3), (800, 4), (640, 5)):
        img = torch.rand((n, 1, hp['height'], Wd), generator=g)
        lens = torch.full((n,), Wd, dtype=torch.int32)
        batches.append((img.pin_memory(), lens))
    per_batch = [net.predict_string(img.to(cuda_device), lens) for img, lens in batches]
    assert all(isinstance(s, str) for b in per_batch for s in b)
    assert [net.predict_string(img, lens) for img, lens in batches] == per_batch          # host tensors
    assert list(net.predict_string_stream(batches)) == per_batch
    arrs = list(net.predict_string_stream(batches, arrays=True))
    for (img, lens), arr in zip(batches, arrs):
        recs = net.predict_labels(img.to(cuda_device), lens)
        assert [[t[:3] for t in l] for l in unpack_decoded(arr)] == [[t[:3] for t in l] for l in recs]
        one = net.predict_labels_arrays(img, lens)
        np.testing.assert_array_equal(one['counts'], arr['counts'])
    # a consumer that stops early leaves no batch behind in the handle's slots
    gen = net.predict_string_stream(batches)
    assert next(gen) == per_batch[0]
    gen.close()
    assert list(net.predict_string_stream(batches[:3])) == per_batch[:3]
    # uint8 host batches through the stream
    u8 = [((img * 255).round().to(torch.uint8), lens) for img, lens in batches[:2]]
    f8 = [(b[0].float() / 255, b[1]) for b in u8]
    assert list(net.predict_string_stream(u8)) == list(net.predict_string_stream(f8))
    eng = net.engine_for(cuda_device)
    img, lens = batches[0]
    eng.recognize_host_begin(0, img, lens)
    host = eng.recognize_host_end(0)
    dev = eng.slot_records(0)
    for k in ('labels', 'starts', 'ends', 'conf', 'counts', 'out_lens'):
        np.testing.assert_array_equal(dev[k].cpu().numpy(), host[k])


def test_archive_round_trip_runs_on_device(cuda_device, tmp_path):
    """(f)4: Lightning checkpoint -> converter -> safetensors archive -> load_safetensors -> device; same logits as a
    model built from the same tensors directly (pred.py:180-210 layouts)."""
    from conformer_ocr_b200 import PytorchRecognitionModel, checkpoint_to_safetensors, synthetic as S
    hp = dict(O.DEFAULT_HPARAMS)
    hp['num_encoder_layers'] = 2
    C = 20
    w = S.make_weights(hp, C, seed=12)
    codec = {chr(0x61 + i): [i + 1] for i in range(C - 1)}
    sd = {'nn.' + k: torch.from_numpy(v) for k, v in w.items()}
    for i in range(hp['num_encoder_layers']):                 # BatchNorm counters are part of a real state dict (strict load, pred.py:194)
        sd[f'nn.encoder.layers.{i}.sequential.2.module.sequential.5.num_batches_tracked'] = torch.tensor(3)
    ck = {'state_dict': sd, 'hyper_parameters': dict(hp, num_classes=C, **DROP), 'TextLineDataModule': {'codec': codec}}
    torch.save(ck, tmp_path / 'm.ckpt')
    checkpoint_to_safetensors(tmp_path / 'm.ckpt', tmp_path / 'm.tar')
    x = torch.rand((3, 1, hp['height'], 300), generator=torch.Generator().manual_seed(2)).to(cuda_device)
    want, _ = build(hp, C, w, cuda_device, 'bf16')(x)
    for net in (PytorchRecognitionModel.load_checkpoint(tmp_path / 'm.ckpt', precision='bf16'),
                PytorchRecognitionModel.load_safetensors(tmp_path / 'm.tar', precision='bf16')):
        got, _ = net.to(cuda_device)(x)
        assert torch.equal(got, want)
        s = net.predict_string(x)
        assert len(s) == 3 and all(set(t) <= set(codec) for t in s)


def test_bucketed_dataloader_pipeline_matches_cpu_port_batch_by_batch(cuda_device):
    """(f)2: dataset -> WidthBucketBatchSampler -> collate_sequences (pinned) -> predict_string_stream.  Every batch is
    also run through the CPU port of the reference with the same composition (the reference's unmasked attention makes
    results depend on the batch's padded width), and the collapsed label sequences must be identical (fp32 path)."""
    from torch.utils.data import DataLoader
    from conformer_ocr_b200 import LabelCodec, PytorchRecognitionModel, WidthBucketBatchSampler, collate_sequences, synthetic as S
    from conformer_ocr_b200.engine import unpack_decoded
    hp = dict(O.DEFAULT_HPARAMS)
    hp['num_encoder_layers'] = 2
    C = 30
    w = S.make_weights(hp, C, seed=21)
    rng = np.random.RandomState(5)
    widths = rng.randint(60, 420, size=37).tolist() + [400] * 6
    g = torch.Generator().manual_seed(6)
    data = [{'image': torch.rand((1, hp['height'], wd), generator=g), 'target': torch.tensor([1, 2])} for wd in widths]
    sampler = WidthBucketBatchSampler(widths, batch_size=8)
    loader = DataLoader(data, batch_sampler=sampler, collate_fn=lambda b: collate_sequences(b, pin=True))
    net = PytorchRecognitionModel(num_classes=C, codec=LabelCodec([chr(0x61 + i) for i in range(C - 1)]), precision='fp32', **hp, **DROP)
    net.nn.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()}, strict=False)
    net = net.to(cuda_device)
    batches = [(b['image'], b['seq_lens']) for b in loader]
    assert sum(b[0].shape[0] for b in batches) == len(widths)
    got = list(net.predict_string_stream(batches, arrays=True))
    tw = TP.to_torch(w)
    ref_net = RM.build_model(hp, C, w).eval() if RM.available() else None
    n_lines = 0
    for (img, lens), arr in zip(batches, got):
        ref = RM.predict_labels(ref_net, img, lens) if ref_net is not None else TP.predict_labels(img, lens, tw, hp)
        mine = unpack_decoded(arr)
        assert [[t[:3] for t in l] for l in mine] == [[t[:3] for t in l] for l in ref]
        n_lines += len(ref)
    assert n_lines == len(widths)
    strings = [s for b in net.predict_string_stream(batches) for s in b]
    assert len(strings) == len(widths) and all(isinstance(s, str) for s in strings)
