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
