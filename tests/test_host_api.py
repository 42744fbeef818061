"""Host-side logic and the C-ABI surface, without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from oracle import conformer_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    from conformer_ocr_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    from conformer_ocr_b200 import _lib
    header = open(os.path.join(ROOT, 'include', 'cocr_b200.h')).read()
    declared = set(re.findall(r'\b(cocr_[a-z_0-9]+)\s*\(', header))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name
    assert b'sm_100a' in lib.cocr_version()


def test_out_frames_matches_calc_length(lib):
    from conformer_ocr_b200.engine import out_frames
    widths = np.array([1, 2, 3, 4, 5, 100, 1024, 1025, 1536, 2048, 4096])
    want = O.calc_length(widths)
    np.testing.assert_array_equal(out_frames(torch.from_numpy(widths)).numpy(), want)
    np.testing.assert_array_equal([lib.cocr_out_frames(int(w)) for w in widths], want)


def test_create_rejects_bad_hparams_like_the_reference(lib):
    from conformer_ocr_b200 import _lib
    hp = _lib.cocr_hparams(96, 144, 16, 5, 4, 2, 31, 1, 32, 4, 10, 0, 0)   # 144 % 5 != 0
    h = ctypes.c_void_p()
    rc = lib.cocr_create(ctypes.byref(hp), 0, ctypes.byref(h))
    assert rc == _lib.COCR_ERR_BAD_ARG
    assert b'num_heads' in lib.cocr_last_error(None)
    with pytest.raises(ValueError):
        _lib.raise_for_status(rc)
    hp = _lib.cocr_hparams(96, 144, 16, 4, 4, 2, 30, 1, 32, 4, 10, 0, 0)   # even kernel size
    assert lib.cocr_create(ctypes.byref(hp), 0, ctypes.byref(h)) == _lib.COCR_ERR_BAD_ARG


def test_module_mirrors_reference_state_dict():
    from conformer_ocr_b200 import PytorchRecognitionModel, ConformerEncoder, RECOGNITION_HYPER_PARAMS
    hp = dict(RECOGNITION_HYPER_PARAMS)
    net = PytorchRecognitionModel(num_classes=37, codec=['a', 'b'], **hp)
    want = dict(O.weight_spec(O.DEFAULT_HPARAMS, 37))
    got = {k: tuple(v.shape) for k, v in net.nn.state_dict().items() if not k.endswith('num_batches_tracked')}
    assert got == want
    assert net.nn['encoder'].count_parameters() == 8217904        # SURVEY.md §0 [probed on the reference]
    assert (net.height, net.channels, net.width) == (96, 1, 0)
    # constructor checks of the reference
    with pytest.raises(AssertionError):
        ConformerEncoder(input_dim=96, encoder_dim=144, num_attention_heads=5)
    with pytest.raises(AssertionError):
        ConformerEncoder(input_dim=96, encoder_dim=144, num_attention_heads=4, conv_kernel_size=30)
    with pytest.raises(ValueError):
        ConformerEncoder(input_dim=96, encoder_dim=144, num_attention_heads=4, subsampling_factor=6)


def test_no_cpu_fallback():
    from conformer_ocr_b200 import PytorchRecognitionModel, RECOGNITION_HYPER_PARAMS
    net = PytorchRecognitionModel(num_classes=5, codec=None, **RECOGNITION_HYPER_PARAMS)
    with pytest.raises(RuntimeError, match='no CPU fallback|CUDA'):
        net(torch.zeros(1, 1, 96, 32), torch.tensor([32]))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'conformer_ocr_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.h', '.cuh')):
                src = open(os.path.join(dirpath, f)).read()
                assert 'oracle' not in src.replace('test / baseline', ''), f'{f} mentions the oracle'


def test_label_codec_decode_contract():
    from conformer_ocr_b200.codec import LabelCodec
    c = LabelCodec({'a': [1], 'b': [2]})
    assert c.decode([(2, 0, 1, .5), (1, 3, 3, .25), (9, 4, 4, .1)]) == [('b', 0, 1, .5), ('a', 3, 3, .25)]
    assert LabelCodec('xy').decode([(1, 0, 0, 1.0)]) == [('x', 0, 0, 1.0)]


def test_product_synthetic_generator_equals_test_generator():
    from conformer_ocr_b200 import synthetic as S
    hp = dict(O.DEFAULT_HPARAMS, num_encoder_layers=2)
    a, b = S.make_weights(hp, 17, seed=3), O.make_weights(hp, 17, seed=3)
    assert list(a) == list(b)
    for k in a:
        np.testing.assert_array_equal(a[k], b[k])
    ia, la = S.make_lines(3, 96, 50, seed=4, widths=[50, 20, 7])
    ib, lb = O.make_lines(3, 96, 50, seed=4, widths=[50, 20, 7])
    np.testing.assert_array_equal(ia, ib)
    np.testing.assert_array_equal(la, lb)
    # 4.95 GF/line at W=1024,C=200 ; 8.11 at 1536 ; 11.71 at 2048 ; 33.57 at 4096/10k
    for (W, C), want in {(1024, 200): 4.95, (1536, 200): 8.11, (2048, 200): 11.71, (4096, 10000): 33.57}.items():
        assert abs(S.flops_per_line(W, C, O.DEFAULT_HPARAMS) / 1e9 - want) < 0.01


def test_label_codec_batched_strings_match_tuple_decode():
    """the vectorised label -> string path gives the strings of `codec.decode` + join (pred.py:163)."""
    import time
    import numpy as np
    from conformer_ocr_b200.codec import LabelCodec
    rng = np.random.default_rng(7)
    charset = [chr(c) for c in list(range(0x61, 0x7b)) + [0x20, 0xe9, 0x3b1, 0x4e2d, 0x1f600]]      # ASCII, Latin-1, Greek, CJK, astral
    codec = LabelCodec(charset)
    B, T = 64, 200
    labels = rng.integers(-1, len(charset) + 4, size=(B, T)).astype(np.int32)                       # includes unknown labels
    counts = rng.integers(0, T + 1, size=B).astype(np.int32)
    counts[0], counts[1] = 0, T                                                                     # empty and full rows
    want = [''.join(g for g, *_ in codec.decode([(int(l), 0, 0, 1.0) for l in row[:n]])) for row, n in zip(labels, counts)]
    assert codec.decode_strings(labels, counts) == want
    # multi-code-point graphemes fall back to the per-label join
    c2 = LabelCodec({'ab': [1], 'c': [2]})
    assert c2.decode_strings(np.array([[1, 2, 1, 9]]), np.array([4])) == ['abcab']
    with pytest.raises(ValueError):
        codec.decode_strings(labels, counts[:-1])
    # and it is the faster of the two (what the row is for); generous factor so the test is not a benchmark
    big_l = rng.integers(1, len(charset) + 1, size=(2048, 384)).astype(np.int32)
    big_n = np.full(2048, 300, np.int32)
    t0 = time.perf_counter(); s_fast = codec.decode_strings(big_l, big_n); t1 = time.perf_counter()
    s_slow = [''.join(g for g, *_ in codec.decode(list(zip(r[:n].tolist(), [0] * n, [0] * n, [1.0] * n)))) for r, n in zip(big_l[:256], big_n[:256])]
    t2 = time.perf_counter()
    assert s_fast[:256] == s_slow
    assert (t1 - t0) / 2048 < (t2 - t1) / 256


def test_load_state_dict_through_submodules_invalidates_engine():
    """`net.nn.load_state_dict(weights)` is the reference idiom (pred.py:193); torch runs only the post hooks
    of the module it was called on and its descendants, so every sub-module carries the invalidation hook."""
    import torch
    from conformer_ocr_b200 import PytorchRecognitionModel
    from conformer_ocr_b200.default_specs import RECOGNITION_HYPER_PARAMS as HP
    hp = {k: HP[k] for k in ('height', 'encoder_dim', 'num_attention_heads', 'feed_forward_expansion_factor', 'conv_expansion_factor',
                             'conv_kernel_size', 'half_step_residual', 'subsampling_conv_channels', 'subsampling_factor')}
    net = PytorchRecognitionModel(num_classes=7, num_encoder_layers=1, input_dropout_p=0.1, feed_forward_dropout_p=0.1,
                                  attention_dropout_p=0.1, conv_dropout_p=0.1, **hp)
    for target in (net, net.nn, net.nn['encoder'], net.nn['decoder']):
        net._engine_key = ('stale',)
        target.load_state_dict(target.state_dict())
        assert net._engine_key is None, type(target).__name__
    enc = net.nn['encoder']
    enc._engine_key = ('stale',)
    net.nn.load_state_dict(net.nn.state_dict())
    assert enc._engine_key is None


def test_recognition_model_signature_defaults_match_reference():
    """model.py:58-72: RecognitionModel(num_classes) builds the 512 / 18 / 8 / 9 / 256 architecture, not the default spec."""
    import inspect
    from conformer_ocr_b200.model import RecognitionModel
    sig = inspect.signature(RecognitionModel.__init__).parameters
    want = dict(height=96, encoder_dim=512, num_encoder_layers=18, num_attention_heads=8, feed_forward_expansion_factor=4,
                conv_expansion_factor=2, conv_kernel_size=9, half_step_residual=True, subsampling_conv_channels=256, subsampling_factor=4)
    for k, v in want.items():
        assert sig[k].default == v, k


@pytest.mark.parametrize('force_port', [False, True])
def test_bench_reference_arm_prints_one_contract_line(force_port):
    """`bench.py --impl reference` runs on the host cores (no GPU needed) and prints ONE JSON line with the contract's keys:
    over the reference's own modules where oracle/_ref is built (kind 'reference'), over the port otherwise (kind 'port')."""
    import json
    import subprocess
    import sys
    from oracle import ref_modules as RM
    env = dict(os.environ)
    env.pop('COCR_BENCH_CPU_PORT', None)
    if force_port:
        env['COCR_BENCH_CPU_PORT'] = '1'
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '1',
                          '--cpu-sample-lines', '1', '--width', '256'], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline',
              'dtype', 'data', 'config', 'cpu_baseline', 'e2e', 'gpu_launches'):
        assert k in d, k
    assert d['impl'] == 'reference' and d['unit'] == 'lines/s' and d['value'] > 0 and d['gpu_launches'] == 0
    assert d['cpu_baseline']['kind'] == ('reference' if RM.available() and not force_port else 'port') and d['e2e']['h2d_bytes_per_step'] == 0


def test_bench_algorithmic_work_sums_to_survey_8d():
    """bench.algorithmic_work (per kernel family) adds up to SURVEY 8(d)'s FLOPs per line, for configs[1] and configs[4]."""
    import importlib.util
    spec = importlib.util.spec_from_file_location('bench_mod', os.path.join(ROOT, 'bench.py'))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    from conformer_ocr_b200 import synthetic as S
    hp = bench.arch_hp(bench.hparams())
    for B, W, C in ((256, 1536, 256), (64, 4096, 10000)):
        a = bench.algorithmic_work(hp, B, W, C)
        once = ('frontend_convs', 'gemm_frontend_pw', 'gemm_sub_out', 'gemm_ctc_argmax', 'ctc_collapse')
        per_layer = ('ffn_fused', 'ffn_fused_final', 'gemm_qkvx_pos', 'attention_tc', 'gemm_attn_out', 'gemm_pw1_glu', 'dwconv_bn_silu_bf16', 'gemm_pw2')
        total = sum(a[k]['flops'] for k in once) + hp['num_encoder_layers'] * sum(a[k]['flops'] for k in per_layer)
        assert total == pytest.approx(B * S.flops_per_line(W, C, hp), rel=1e-12)
        assert a['frontend_tc']['flops'] == a['frontend_convs']['flops'] + a['gemm_frontend_pw']['flops']
        assert a['attention_tc']['bound'] == 'tensor' and a['dwconv_bn_silu_bf16']['bound'] == 'hbm'


def test_tensor_core_kernels_hold_tcgen05_and_tma_instructions(lib):
    """The hot kernels of the shipped library are tcgen05 / TMEM / TMA code, not a recompiled SIMT or mma.sync path: per kernel,
    `cuobjdump -sass` shows UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld) and - where operands arrive by TMA - UTMALDG
    (cp.async.bulk.tensor); mnemonics as in /opt/skills/guides/B200_PROFILING.md.  No GPU needed."""
    import shutil
    import subprocess
    from conformer_ocr_b200 import _lib
    tool = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
    if not os.path.exists(tool):
        pytest.skip('cuobjdump not installed')
    sass = subprocess.run([tool, '-sass', _lib.LIB_PATH], capture_output=True, text=True, timeout=300, check=True).stdout
    counts = {}
    for part in re.split(r'\n\s*Function : ', sass)[1:]:
        name = part.split('\n', 1)[0].strip()
        counts[name] = {m: len(re.findall(r'\b' + m, part)) for m in ('UTCHMMA', 'LDTM', 'STTM', 'UTMALDG', 'HMMA')}

    def kernels(fragment):
        hit = {k: v for k, v in counts.items() if fragment in k}
        assert hit, f'no kernel named *{fragment}* in the library'
        return hit
    for frag, tma in (('k_gemm_tc', True), ('k_ffn_fused2', True), ('k_ffn_fused', True), ('k_attn_tc', True), ('k_bgemm_tc', True),
                      ('k_frontend_tc', False)):          # the front-end gathers its image rows with plain loads (strided caller memory)
        for name, c in kernels(frag).items():
            assert c['UTCHMMA'] > 0 and c['LDTM'] > 0, (name, c)
            assert not tma or c['UTMALDG'] > 0, (name, c)
    assert len(kernels('k_gemm_tc')) == 7                   # one instantiation per epilogue
    assert all(c['STTM'] > 0 for c in kernels('k_attn_tc').values())      # Q and P are written to TMEM as A operands
    assert all(c['STTM'] > 0 for c in kernels('k_ffn_fused2').values())   # so are the LayerNorm'd rows of the fused FFN
    assert not any(c['HMMA'] for c in counts.values())      # no mma.sync / wmma kernel anywhere in the library


def test_no_function_local_static_frozen_by_the_first_model():
    """A function-local `static` whose initialiser reads the handle or its dimensions keeps the value of the FIRST model the
    process ran (that is how the GPU suite, which starts with a tiny golden, ran another pw1 tiling than the benchmark did;
    DESIGN.md 5).  Statics in the host code of the library may depend on the environment only."""
    csrc = os.path.join(ROOT, 'conformer_ocr_b200', 'csrc')
    bad = []
    for fn in sorted(os.listdir(csrc)):
        if not fn.endswith(('.cu', '.cuh', '.h')):
            continue
        for no, line in enumerate(open(os.path.join(csrc, fn)), 1):
            m = re.match(r'\s*(?:\{\s*)?static\s+(?:const\s+)?(?:bool|int|size_t|float|unsigned|uint32_t|int64_t)\s+\w+\s*=\s*([^;]*)', line)
            if m and 'constexpr' not in line and re.search(r'\bh->|\bdm\.|\bhp\.|\bd\s*==|\bp\.[A-Za-z]|\bB\b|\bT\b', m.group(1)):
                bad.append(f'{fn}:{no}: {line.strip()}')
    assert not bad, '\n'.join(bad)
