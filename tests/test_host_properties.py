"""CPU: property tests (hypothesis, bounded) of the host logic either side of the device path - the record wire format, the
shard planner, the width-bucket sampler, the safetensors container, the label codec and the five statements of
`calc_length` (convolution.py:240-247) that have to agree: numpy oracle, host shim, integer form, C ABI, and - where
oracle/_ref is built - the reference's own function."""
import numpy as np
import pytest
import torch

pytest.importorskip('hypothesis')           # a missing optional package must not break the collection of the other files
from hypothesis import given, settings  # noqa: E402
from hypothesis import strategies as st  # noqa: E402

from conformer_ocr_b200 import sharding as S
from conformer_ocr_b200.codec import LabelCodec
from conformer_ocr_b200.collate import WidthBucketBatchSampler, collate_sequences
from conformer_ocr_b200.convert import read_safetensors, write_safetensors
from conformer_ocr_b200.engine import out_frames
from oracle import conformer_oracle as O
from oracle import ref_modules as RM

FAST = settings(max_examples=60, deadline=None)

f32 = st.floats(width=32, allow_nan=False, allow_infinity=True)
record = st.tuples(st.integers(1, 2 ** 31 - 1), st.integers(0, 10 ** 6), st.integers(0, 10 ** 6), f32)


@FAST
@given(st.dictionaries(st.integers(0, 10 ** 6), st.lists(record, max_size=12), max_size=10))
def test_record_wire_format_round_trips(recs):
    out = S.unpack_records(S.pack_records(recs))
    assert out.keys() == recs.keys()
    for k, rows in recs.items():
        assert [r[:3] for r in out[k]] == [r[:3] for r in rows]
        # confidences travel as their fp32 bit patterns
        assert [np.float32(r[3]).tobytes() for r in out[k]] == [np.float32(r[3]).tobytes() for r in rows]


@FAST
@given(st.lists(st.integers(1, 5000), max_size=300), st.integers(1, 8), st.integers(1, 64), st.booleans())
def test_shard_plan_invariants(widths, world, batch_lines, equal_frames):
    shards = S.plan_shards(widths, world, batch_lines, equal_frames=equal_frames)
    assert len(shards) == world
    seen = sorted(i for sh in shards for b in sh for i in b.lines)
    assert seen == list(range(len(widths)))                                   # every line exactly once
    for sh in shards:
        for b in sh:
            assert 1 <= len(b.lines) <= batch_lines
            ws = [widths[i] for i in b.lines]
            assert ws == sorted(ws, reverse=True)                             # collate order: widest first
            assert b.width == ws[0] and b.frames == S.out_frames_int(ws[0])
            if equal_frames:
                assert len({S.out_frames_int(w) for w in ws}) == 1
    # longest-processing-time-first: no rank carries more than the mean plus one batch
    loads = [sum(b.cost for b in sh) for sh in shards]
    biggest = max((b.cost for sh in shards for b in sh), default=0.0)
    assert max(loads, default=0.0) <= sum(loads) / world + biggest + 1e-9


@FAST
@given(st.lists(st.integers(1, 3000), min_size=1, max_size=200), st.integers(1, 32), st.booleans(), st.integers(0, 5))
def test_width_bucket_sampler_invariants(widths, batch_size, shuffle, epoch):
    smp = WidthBucketBatchSampler(widths, batch_size, shuffle=shuffle, seed=3)
    smp.set_epoch(epoch)
    batches = list(smp)
    assert len(batches) == len(smp)
    assert sorted(i for b in batches for i in b) == list(range(len(widths)))
    for b in batches:
        assert 1 <= len(b) <= batch_size
        assert len({S.out_frames_int(widths[i]) for i in b}) == 1
        assert max(widths[i] for i in b) - min(widths[i] for i in b) <= 3      # one frame count spans at most 4 widths
    assert sorted(map(tuple, batches)) == sorted(map(tuple, WidthBucketBatchSampler(widths, batch_size)))   # shuffling permutes whole batches


@FAST
@given(st.lists(st.integers(1, 40), min_size=1, max_size=6), st.booleans())
def test_collate_pads_right_with_zeros_and_sorts_by_width(widths, uint8):
    rng = np.random.RandomState(len(widths))
    samples = [{'image': torch.from_numpy(rng.rand(1, 4, w).astype(np.float32)), 'target': torch.arange(1, 1 + (i % 3) + 1)} for i, w in enumerate(widths)]
    out = collate_sequences(samples, uint8=uint8)
    lens = out['seq_lens'].tolist()
    assert lens == sorted(widths, reverse=True) and out['image'].shape == (len(widths), 1, 4, max(widths))
    assert out['image'].dtype == (torch.uint8 if uint8 else torch.float32)
    for row, n in zip(out['image'], lens):
        assert not row[..., n:].any()                                          # right padding is zero
    assert int(out['target_lens'].sum()) == out['target'].numel()
    if not uint8:                                                              # pixels survive unchanged, in sorted order
        order = sorted(range(len(widths)), key=lambda i: widths[i], reverse=True)
        for row, i in zip(out['image'], order):
            if sum(1 for w in widths if w == widths[i]) == 1:
                assert torch.equal(row[..., :widths[i]], samples[i]['image'])


_DT = [np.float32, np.float16, np.int64, np.int32, np.uint8, np.float64]


@FAST
@given(st.dictionaries(st.text(alphabet='abcdefgh._0123', min_size=1, max_size=12),
                       st.tuples(st.sampled_from(range(len(_DT))), st.lists(st.integers(0, 5), max_size=3), st.integers(0, 2 ** 31 - 1)), max_size=6))
def test_safetensors_container_round_trips(spec):
    tensors = {}
    for name, (dt, shape, seed) in spec.items():
        rng = np.random.RandomState(seed)
        tensors[name] = np.asarray(rng.rand(*shape) * 200 - 100).astype(_DT[dt])
    blob = write_safetensors(tensors)
    back = read_safetensors(blob)
    assert back.keys() == tensors.keys()
    for k, v in tensors.items():
        assert back[k].dtype == v.dtype and back[k].shape == v.shape
        assert back[k].tobytes() == v.tobytes()
    (n,) = np.frombuffer(blob[:8], '<u8')
    assert (8 + int(n)) % 8 == 0                                               # the data section starts 8-byte aligned


@pytest.mark.parametrize('bad', [b'', b'\x00' * 7, (10 ** 6).to_bytes(8, 'little') + b'{}'])
def test_safetensors_rejects_truncated_blobs(bad):
    with pytest.raises(ValueError):
        read_safetensors(bad)


_ALPHABET = [chr(c) for c in range(0x61, 0x7b)] + ['ä', 'ß', '漢', '字', ' ', 'é']


@FAST
@given(st.text(alphabet=_ALPHABET, max_size=40))
def test_codec_one_to_one_round_trip_and_batched_strings(s):
    codec = LabelCodec(_ALPHABET)
    labels = codec.encode(s)
    assert all(1 <= lab <= len(_ALPHABET) for lab in labels)
    recs = [(lab, i, i, 0.5) for i, lab in enumerate(labels)]
    assert ''.join(g for g, *_ in codec.decode(recs)) == s
    block = np.zeros((2, max(len(labels), 1) + 3), np.int32)
    block[0, :len(labels)] = labels
    block[1, :len(labels)] = labels[::-1]
    got = codec.decode_strings(block, np.array([len(labels), len(labels)]))
    assert got == [s, s[::-1]]


@FAST
@given(st.lists(st.sampled_from(['a', 'b', 'ab', 'abc', 'c', 'ca']), max_size=15))
def test_codec_n_to_m_longest_match(graphemes):
    """kraken-style n:m codec: graphemes of several code points, label sequences of several labels; encode picks the longest
    grapheme, decode the longest label sequence, and decoding an encoding gives the string back."""
    c2l = {'a': [1], 'b': [2], 'ab': [3, 4], 'abc': [5], 'c': [6, 7], 'ca': [8]}
    codec = LabelCodec(c2l)
    s = ''.join(graphemes)
    labels = codec.encode(s)
    recs = [(lab, 2 * i, 2 * i + 1, float(i)) for i, lab in enumerate(labels)]
    dec = codec.decode(recs)
    assert ''.join(g for g, *_ in dec) == s
    for g, start, end, conf in dec:                                            # spans and confidences come from the covered records
        assert start <= end and start % 2 == 0 and end % 2 == 1
    assert [d[1] for d in dec] == sorted(d[1] for d in dec)
    block = np.asarray([labels + [0, 0]], np.int32) if labels else np.zeros((1, 2), np.int32)
    assert codec.decode_strings(block, np.array([len(labels)])) == [s]


def test_calc_length_statements_agree_on_every_width():
    widths = np.arange(1, 6001)
    want = O.calc_length(widths)
    np.testing.assert_array_equal(out_frames(torch.from_numpy(widths)).numpy(), want)
    np.testing.assert_array_equal([S.out_frames_int(int(w)) for w in widths], want)
    from conformer_ocr_b200 import _lib
    lib = _lib.load()
    np.testing.assert_array_equal([lib.cocr_out_frames(int(w)) for w in widths], want)
    if RM.available():
        import conformer_ocr.conformer.convolution as ref_conv                 # the reference's own function (oracle/_ref)
        got = ref_conv.calc_length(torch.from_numpy(widths), all_paddings=2, kernel_size=3, stride=2, repeat_num=2)
        assert got.dtype == torch.int32
        np.testing.assert_array_equal(got.numpy(), want)
