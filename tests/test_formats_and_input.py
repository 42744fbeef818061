"""Checkpoint formats, converter, codec and the input side (collate / width bucketing): host logic, no GPU.
SURVEY.md 8(f) ranks 1, 2 and 4."""
import io
import json
import tarfile

import numpy as np
import pytest
import torch

from conformer_ocr_b200 import (LabelCodec, PytorchRecognitionModel, WidthBucketBatchSampler, checkpoint_to_safetensors,
                                collate_sequences)
from conformer_ocr_b200 import synthetic as S
from conformer_ocr_b200.convert import read_archive, read_safetensors, write_safetensors
from conformer_ocr_b200.default_specs import RECOGNITION_HYPER_PARAMS as HP
from conformer_ocr_b200.sharding import out_frames_int


def small_hp(layers=2):
    hp = dict(HP)
    hp['num_encoder_layers'] = layers
    return hp


def lightning_checkpoint(tmp_path, hp, C, codec, seed=0):
    """A file with the layout of a Lightning checkpoint of the reference's RecognitionModel + TextLineDataModule."""
    w = S.make_weights(hp, C, seed=seed)
    sd = {'nn.' + k: torch.from_numpy(v) for k, v in w.items()}
    for i in range(hp['num_encoder_layers']):
        sd[f'nn.encoder.layers.{i}.sequential.2.module.sequential.5.num_batches_tracked'] = torch.tensor(7)
    sd['val_loss.mean_value'] = torch.tensor(0.0)             # a torchmetrics buffer of the LightningModule (model.py:121-123)
    ck = {'state_dict': sd, 'hyper_parameters': dict(hp, num_classes=C, lr=1e-3, optimizer='AdamW', batch_size=32),
          'TextLineDataModule': {'codec': codec}, 'epoch': 3, 'global_step': 1234}
    path = tmp_path / 'checkpoint_03-0.9000.ckpt'
    torch.save(ck, path)
    return path, w


def test_safetensors_container_round_trip_matches_the_library():
    import safetensors.torch
    rng = np.random.RandomState(0)
    t = {'a.weight': rng.normal(size=(3, 5)).astype(np.float32), 'b': rng.normal(size=(7,)).astype(np.float16),
         'c.n': np.asarray(3, dtype=np.int64), 'z': np.zeros((0, 4), np.float32)}
    blob = write_safetensors(t)
    ref = safetensors.torch.load(blob)                         # the library reads what we write
    for k, v in t.items():
        np.testing.assert_array_equal(ref[k].numpy(), v)
    blob2 = safetensors.torch.save({k: torch.from_numpy(v) for k, v in t.items()} | {'h': torch.ones(2, 2, dtype=torch.bfloat16) * 1.5})
    got = read_safetensors(blob2)                              # and we read what the library writes (bf16 widened to fp32)
    for k, v in t.items():
        np.testing.assert_array_equal(got[k], v)
    np.testing.assert_array_equal(got['h'], np.full((2, 2), 1.5, np.float32))
    with pytest.raises(ValueError):
        read_safetensors(b'\x01\x02')


def test_checkpoint_loader_and_converter_round_trip(tmp_path):
    hp, C = small_hp(), 12
    codec = {chr(0x61 + i): [i + 1] for i in range(C - 1)}
    ck_path, w = lightning_checkpoint(tmp_path, hp, C, codec)
    net = PytorchRecognitionModel.load_checkpoint(ck_path)      # pred.py:197-210
    assert not net.training and net.num_classes == C and net.height == hp['height']
    assert isinstance(net.codec, LabelCodec) and net.codec.c2l == codec
    sd = net.nn.state_dict()
    for k, v in w.items():
        np.testing.assert_array_equal(sd[k].numpy(), v)
    # converter (working form of pred.py:213-254) -> archive -> load_safetensors (pred.py:180-195)
    user_meta = tmp_path / 'meta.json'
    user_meta.write_text(json.dumps({'authors': ['x'], 'license': 'Apache-2.0'}))
    out = tmp_path / 'model.tar'
    checkpoint_to_safetensors(ck_path, out, metadata=user_meta)
    with tarfile.open(out) as tf:
        assert sorted(tf.getnames()) == ['_kraken_metadata.json', 'metadata.json', 'model.safetensors']
        meta = json.load(tf.extractfile('metadata.json'))
        assert json.load(tf.extractfile('_kraken_metadata.json'))['license'] == 'Apache-2.0'
        import safetensors.torch
        st = safetensors.torch.load(tf.extractfile('model.safetensors').read())     # readable by the reference's own loader code
    assert meta['codec'] == codec and meta['model_type'] == 'recognition' and meta['kraken_min_version'] == 5.3
    assert meta['hyper_params']['encoder_dim'] == hp['encoder_dim'] and meta['hyper_params']['num_classes'] == C
    assert set(st) == {k for k in sd}                          # keys WITHOUT the nn. prefix, counters included, metrics dropped
    net2 = PytorchRecognitionModel.load_safetensors(out, precision='bf16')
    assert net2.precision == 'bf16' and net2.codec.c2l == codec
    for k, v in w.items():
        np.testing.assert_array_equal(net2.nn.state_dict()[k].numpy(), v)
    metadata, arrays = read_archive(out)
    np.testing.assert_array_equal(arrays['decoder.weight'], w['decoder.weight'])


def test_loaders_reject_malformed_files_like_the_reference(tmp_path):
    hp, C = small_hp(1), 5
    ck_path, _ = lightning_checkpoint(tmp_path, hp, C, {'a': [1]})
    ck = torch.load(ck_path, weights_only=False)
    for missing, msg in (('TextLineDataModule', 'data module state'), ('hyper_parameters', 'hyperparameters')):
        bad = {k: v for k, v in ck.items() if k != missing}
        p = tmp_path / f'no_{missing}.ckpt'
        torch.save(bad, p)
        with pytest.raises(ValueError, match=msg):             # pred.py:204,207
            PytorchRecognitionModel.load_checkpoint(p)
        with pytest.raises(ValueError):
            checkpoint_to_safetensors(p, tmp_path / 'x.tar')
    for drop, msg in (('codec', 'No codec'), ('hyper_params', 'No hyperparameters')):       # pred.py:188,191
        meta = {'codec': {'a': [1]}, 'hyper_params': dict(hp, num_classes=C)}
        del meta[drop]
        p = tmp_path / f'no_{drop}.tar'
        with tarfile.open(p, 'w') as tar:
            for name, blob in (('metadata.json', json.dumps(meta).encode()), ('model.safetensors', write_safetensors({}))):
                ti = tarfile.TarInfo(name)
                ti.size = len(blob)
                tar.addfile(ti, io.BytesIO(blob))
        with pytest.raises(ValueError, match=msg):
            PytorchRecognitionModel.load_safetensors(p)


def test_codec_one_to_one_and_n_to_m():
    c = LabelCodec({'a': [1], 'b': [2], 'ch': [3, 4], 'c': [3], 'é': [5]})
    assert not c.one_to_one and c.max_label == 5
    recs = [(1, 0, 1, .9), (3, 2, 2, .5), (4, 3, 5, .7), (3, 6, 6, .4), (9, 7, 7, .1), (2, 8, 9, .8)]
    assert c.decode(recs) == [('a', 0, 1, .9), ('ch', 2, 5, .7), ('c', 6, 6, .4), ('b', 8, 9, .8)]      # longest match, unknown skipped
    assert c.encode('achcbé') == [1, 3, 4, 3, 2, 5]
    with pytest.raises(KeyError):
        c.encode('x')
    labels = np.array([[1, 3, 4, 3, 2, 0], [5, 5, 0, 0, 0, 0]])
    assert c.decode_strings(labels, np.array([5, 2])) == ['achcb', 'éé']
    with pytest.raises(ValueError):
        LabelCodec({'a': [0]})
    one = LabelCodec('xyz')
    assert one.one_to_one and one.encode('zxy') == [3, 1, 2]
    assert ''.join(g for g, *_ in one.decode([(3, 0, 0, 1.), (1, 1, 1, 1.)])) == 'zx'


def test_collate_sequences_layout():
    g = torch.Generator().manual_seed(0)
    widths = [50, 120, 75, 120]
    batch = [{'image': torch.rand((1, 96, w), generator=g), 'target': torch.arange(1, 1 + i + 2)} for i, w in enumerate(widths)]
    out = collate_sequences(batch)
    assert out['image'].shape == (4, 1, 96, 120) and out['image'].dtype == torch.float32
    assert out['seq_lens'].tolist() == [120, 120, 75, 50] and out['seq_lens'].dtype == torch.int64
    assert out['target_lens'].tolist() == [3, 5, 4, 2]          # stable sort: sample 1 before sample 3
    assert out['target'].tolist() == [1, 2, 3, 1, 2, 3, 4, 5, 1, 2, 3, 4, 1, 2]
    assert torch.equal(out['image'][2, :, :, :75], batch[2]['image']) and float(out['image'][2, :, :, 75:].abs().sum()) == 0.0
    u8 = collate_sequences(batch, uint8=True)
    assert u8['image'].dtype == torch.uint8
    np.testing.assert_array_equal(u8['image'][3, 0, :, :50].numpy(), (batch[0]['image'][0] * 255).round().numpy().astype(np.uint8))
    strs = collate_sequences([{'image': torch.zeros(1, 96, 10), 'target': 'ab'}, {'image': torch.zeros(1, 96, 20), 'target': 'c'}])
    assert strs['target'] == ['c', 'ab'] and strs['target_lens'].tolist() == [1, 2]


def test_width_bucket_sampler_gives_equal_frame_batches():
    rng = np.random.RandomState(1)
    widths = rng.randint(30, 2049, size=500).tolist() + [1536] * 40
    s = WidthBucketBatchSampler(widths, batch_size=16)
    seen = []
    for b in s:
        assert 1 <= len(b) <= 16
        assert len({out_frames_int(widths[i]) for i in b}) == 1
        assert max(widths[i] for i in b) - min(widths[i] for i in b) <= 3
        seen += b
    assert sorted(seen) == list(range(len(widths)))
    assert len(s) == len(list(s))
    sh = WidthBucketBatchSampler(widths, batch_size=16, shuffle=True, seed=3)
    a, b2 = [tuple(x) for x in sh], [tuple(x) for x in sh]
    assert a == b2 and sorted(a) == sorted(tuple(x) for x in s) and a != [tuple(x) for x in s]
    sh.set_epoch(1)
    assert [tuple(x) for x in sh] != a
    full = WidthBucketBatchSampler(widths, batch_size=16, drop_last=True)
    assert all(len(b) == 16 for b in full)
    with pytest.raises(ValueError):
        WidthBucketBatchSampler([10, 0], 4)


def test_mirror_parameter_order_is_the_references():
    """`nn.parameters()` / `state_dict()` order of the mirrors = the reference's module registration order (encoder.py:67-103):
    Lightning indexes `optimizer_states` by it.  First layer spelled out; against the reference's own modules where oracle/_ref
    is built."""
    from conformer_ocr_b200 import RecognitionModel
    hp = small_hp(2)
    m = RecognitionModel(num_classes=9, **hp)
    names = [k for k, _ in m.nn.named_parameters()]
    p = 'encoder.layers.0.sequential.'
    first = ['0.module.sequential.0.weight', '0.module.sequential.0.bias', '0.module.sequential.1.linear.weight', '0.module.sequential.1.linear.bias',
             '0.module.sequential.4.linear.weight', '0.module.sequential.4.linear.bias', '1.module.layer_norm.weight', '1.module.layer_norm.bias',
             '1.module.attention.u_bias', '1.module.attention.v_bias', '1.module.attention.query_proj.linear.weight',
             '1.module.attention.query_proj.linear.bias', '1.module.attention.key_proj.linear.weight', '1.module.attention.key_proj.linear.bias',
             '1.module.attention.value_proj.linear.weight', '1.module.attention.value_proj.linear.bias', '1.module.attention.pos_proj.linear.weight',
             '1.module.attention.out_proj.linear.weight', '1.module.attention.out_proj.linear.bias', '2.module.sequential.0.weight',
             '2.module.sequential.0.bias', '2.module.sequential.2.conv.weight', '2.module.sequential.2.conv.bias', '2.module.sequential.4.conv.weight',
             '2.module.sequential.5.weight', '2.module.sequential.5.bias', '2.module.sequential.7.conv.weight', '2.module.sequential.7.conv.bias',
             '3.module.sequential.0.weight', '3.module.sequential.0.bias', '3.module.sequential.1.linear.weight', '3.module.sequential.1.linear.bias',
             '3.module.sequential.4.linear.weight', '3.module.sequential.4.linear.bias', '4.weight', '4.bias']
    assert names[:8] == ['encoder.conv_subsample.conv.0.weight', 'encoder.conv_subsample.conv.0.bias', 'encoder.conv_subsample.conv.2.weight',
                         'encoder.conv_subsample.conv.2.bias', 'encoder.conv_subsample.conv.3.weight', 'encoder.conv_subsample.conv.3.bias',
                         'encoder.conv_subsample.out.0.weight', 'encoder.conv_subsample.out.0.bias']
    assert names[8:8 + len(first)] == [p + k for k in first]
    assert names[-2:] == ['decoder.weight', 'decoder.bias']
    from oracle import ref_modules as RM
    if RM.available():
        ref = RM.build_model(hp, 9)
        assert [k for k, _ in ref.named_parameters()] == names
        assert [(k, tuple(v.shape), v.dtype) for k, v in ref.state_dict().items()] == [(k, tuple(v.shape), v.dtype) for k, v in m.nn.state_dict().items()]


def test_adamw_state_mapping_between_flat_buffers_and_torch():
    """The trainer keeps AdamW's moments in two flat buffers; checkpoints carry them as `torch.optim.AdamW.state_dict()` indexed in
    `nn.parameters()` order.  A real torch optimiser's state -> flat -> state dict is the identity, and torch loads the result."""
    from conformer_ocr_b200 import RecognitionModel
    from conformer_ocr_b200.train import adamw_state_to_flat, flat_to_adamw_state
    hp = small_hp(2)
    m = RecognitionModel(num_classes=9, **hp)
    names = [k for k, _ in m.nn.named_parameters()]
    params = [torch.nn.Parameter(p.detach().clone()) for _, p in m.nn.named_parameters()]
    # a layout like cocr_train_entry's: trainable tensors back to back, 8-element aligned, in an order of its own
    layout, off = {}, 0
    for k, p in sorted(zip(names, params), key=lambda kp: kp[0]):
        off = (off + 7) // 8 * 8
        layout[k] = (off, p.numel(), True, tuple(p.shape))
        off += p.numel()
    opt = torch.optim.AdamW(params, lr=3e-4, weight_decay=1e-5)
    g = torch.Generator().manual_seed(0)
    for _ in range(3):
        for p in params:
            p.grad = torch.randn(p.shape, generator=g)
        opt.step()
    want = opt.state_dict()
    exp_avg, exp_avg_sq = torch.full((off,), 7.0), torch.full((off,), 7.0)
    step = adamw_state_to_flat(layout, want, names, exp_avg, exp_avg_sq)
    assert step == 3
    for i, k in enumerate(names):
        o, n, _, shape = layout[k]
        assert torch.equal(exp_avg[o:o + n].view(shape), want['state'][i]['exp_avg'])
        assert torch.equal(exp_avg_sq[o:o + n].view(shape), want['state'][i]['exp_avg_sq'])
    got = flat_to_adamw_state(layout, exp_avg, exp_avg_sq, step, names, 3e-4, (0.9, 0.999), 1e-8, 1e-5)
    assert got['param_groups'][0]['params'] == want['param_groups'][0]['params']
    for i in range(len(names)):
        assert float(got['state'][i]['step']) == 3.0
        assert torch.equal(got['state'][i]['exp_avg'], want['state'][i]['exp_avg'])
        assert torch.equal(got['state'][i]['exp_avg_sq'], want['state'][i]['exp_avg_sq'])
    opt2 = torch.optim.AdamW([torch.nn.Parameter(p.detach().clone()) for p in params], lr=1.0)
    opt2.load_state_dict(got)                                   # torch accepts it as its own
    assert opt2.param_groups[0]['lr'] == 3e-4 and opt2.param_groups[0]['weight_decay'] == 1e-5
    assert flat_to_adamw_state(layout, exp_avg, exp_avg_sq, 0, names, 1e-3, (0.9, 0.999), 1e-8, 0.0)['state'] == {}
    with pytest.raises(ValueError):
        adamw_state_to_flat(layout, want, names[:-1], exp_avg, exp_avg_sq)


def test_recognition_model_checkpoint_round_trip(tmp_path):
    """`RecognitionModel.save_checkpoint` (model.py:219-220) writes the layout the reference's loaders read:
    `PytorchRecognitionModel.load_checkpoint` (pred.py:197-210), the converter and `load_from_checkpoint` (cli/train.py:328)."""
    from conformer_ocr_b200 import RecognitionModel
    hp, C = small_hp(2), 12
    codec = {chr(0x61 + i): [i + 1] for i in range(C - 1)}
    w = S.make_weights(hp, C, seed=3)
    m = RecognitionModel(num_classes=C, lr=3e-4, optimizer='AdamW', **hp)
    m.nn.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()}, strict=False)
    path = tmp_path / 'checkpoint_00.ckpt'
    m.save_checkpoint(path, codec=codec)
    ck = torch.load(path, weights_only=False)
    assert list(ck['state_dict']) == ['nn.' + k for k in m.nn.state_dict()]
    assert ck['hyper_parameters']['num_classes'] == C and ck['hyper_parameters']['lr'] == 3e-4 and ck['hyper_parameters']['optimizer'] == 'AdamW'
    assert ck['TextLineDataModule'] == {'codec': codec} and 'optimizer_states' not in ck and ck['global_step'] == 0
    net = PytorchRecognitionModel.load_checkpoint(path)
    assert net.codec.c2l == codec
    for k, v in w.items():
        np.testing.assert_array_equal(net.nn.state_dict()[k].numpy(), v)
    checkpoint_to_safetensors(path, tmp_path / 'm.tar')
    net2 = PytorchRecognitionModel.load_safetensors(tmp_path / 'm.tar')
    np.testing.assert_array_equal(net2.nn.state_dict()['decoder.weight'].numpy(), w['decoder.weight'])
    back = RecognitionModel.load_from_checkpoint(path, precision='bf16')
    assert back.precision == 'bf16' and back.codec == codec and back.hyper_parameters['lr'] == 3e-4 and back._resume is None
    for k, v in w.items():
        np.testing.assert_array_equal(back.nn.state_dict()[k].numpy(), v)
    # a checkpoint WITH optimiser state (as Lightning writes it) keeps it through load -> save until a trainer takes it over
    ck['optimizer_states'] = [{'state': {}, 'param_groups': [{'lr': 3e-4, 'params': list(range(len(list(m.nn.parameters()))))}]}]
    ck['global_step'] = 41
    torch.save(ck, tmp_path / 'resume.ckpt')
    res = RecognitionModel.load_from_checkpoint(tmp_path / 'resume.ckpt')
    assert res._resume['global_step'] == 41
    res.save_checkpoint(tmp_path / 'again.ckpt')
    again = torch.load(tmp_path / 'again.ckpt', weights_only=False)
    assert again['global_step'] == 41 and again['optimizer_states'] == ck['optimizer_states'] and again['TextLineDataModule'] == {'codec': codec}
    bad = {k: v for k, v in ck.items() if k != 'hyper_parameters'}
    torch.save(bad, tmp_path / 'bad.ckpt')
    with pytest.raises(ValueError, match='hyperparameters'):
        RecognitionModel.load_from_checkpoint(tmp_path / 'bad.ckpt')
