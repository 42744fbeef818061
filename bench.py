#!/usr/bin/env python
"""Benchmark of the conformer_ocr hot path (encoder forward -> CTC projection -> greedy decode).

    python bench.py --gpus N --steps K --warmup W            # our arm (libcocr_b200.so on N B200s)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (port), host cores
    python bench.py --config 1|2|5 ...                       # BASELINE.json configs[0|1|4] (default 2 = configs[1])
    python bench.py --config 4 ...                           # configs[3]: training step, 32 lines per GPU, CTC loss + all-reduce + AdamW

Workload (default, BASELINE.json configs[1]): default-spec conformer, batches of 256 lines, 96 x 1536 px, 256 classes,
bf16.  One "step" = `--passes-per-step` (default 16) passes of the hot path over 256-line batches = 4096 lines per GPU:
a slice of config 3's shard (100 000 lines over 8 GPUs), long enough that K = 20 steps time > 2 s.  At N > 1 every rank
runs its own shard (no data-path collective) and the decoded records of the whole timed region are gathered over NCCL
inside the timed region (one collective per shard).  Prints ONE JSON line (rank 0).  DESIGN.md 6 explains
every key.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

CONFIGS = {
    # BASELINE.json configs[0]: the reference's own CPU-runnable case
    1: dict(batch=8, width=1024, classes=256, precision='fp32', passes=1, name='configs[0]'),
    # configs[1]: the configuration the metric is quoted on
    2: dict(batch=256, width=1536, classes=256, precision='bf16', passes=16, name='configs[1]'),
    # configs[3]: one data-parallel TRAINING step (forward in train() mode, CTC loss, backward, gradient all-reduce, AdamW);
    # 32 lines per GPU per step (the weak-scaling variant of SURVEY.md 8d cfg 4)
    4: dict(batch=32, width=1536, classes=256, precision='bf16', passes=1, name='configs[3]'),
    # configs[4]: long lines, CJK-sized alphabet
    5: dict(batch=64, width=4096, classes=10000, precision='bf16', passes=4, name='configs[4]'),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', choices=['ours', 'reference'], default='ours')
    ap.add_argument('--config', type=int, choices=sorted(CONFIGS), default=2)
    ap.add_argument('--precision', choices=['fp32', 'bf16'], default=None)
    ap.add_argument('--batch', type=int, default=None, help='lines per pass per GPU')
    ap.add_argument('--width', type=int, default=None)
    ap.add_argument('--classes', type=int, default=None)
    ap.add_argument('--passes-per-step', type=int, default=None, help='hot-path passes (batches) per step')
    ap.add_argument('--cpu-sample-lines', type=int, default=16, help='lines per CPU-baseline pass')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-eager-baseline', action='store_true')
    ap.add_argument('--profile-steps', type=int, default=2, help='extra passes with per-kernel CUDA events (roofline)')
    a = ap.parse_args()
    c = CONFIGS[a.config]
    a.batch = a.batch or c['batch']
    a.width = a.width or c['width']
    a.classes = a.classes or c['classes']
    a.precision = a.precision or c['precision']
    a.passes = a.passes_per_step or c['passes']
    return a


def load_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as fp:
            p = json.load(fp)
        return dict(hbm_gbs=p['hbm_gbs'], tflops_burst=p['bf16_tflops'], tflops_sustained=p.get('bf16_tflops_sustained', p['bf16_tflops']),
                    source='measured')
    # fallback stated in /opt/skills/guides/B200_PROFILING.md
    return dict(hbm_gbs=6650.0, tflops_burst=1590.0, tflops_sustained=1400.0, source='fallback')


class ClockSampler(threading.Thread):
    """Samples nvidia-smi SM clocks / throttle reasons while the timed region runs."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu_index = gpu_index
        self.rows = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '100', '-i', str(self.gpu_index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                parts = [p.strip() for p in line.split(',')]
                if len(parts) >= 8:
                    self.rows.append(parts)
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except ValueError:
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[4:8]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        return dict(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))


def hparams():
    from conformer_ocr_b200.default_specs import RECOGNITION_HYPER_PARAMS
    return dict(RECOGNITION_HYPER_PARAMS)


def arch_hp(hp):
    return {k: hp[k] for k in ('height', 'encoder_dim', 'num_encoder_layers', 'num_attention_heads',
                               'feed_forward_expansion_factor', 'conv_expansion_factor', 'conv_kernel_size',
                               'half_step_residual', 'subsampling_conv_channels', 'subsampling_factor')}


def workload_name(args):
    return (f'BASELINE {CONFIGS[args.config]["name"]}: default_specs conformer (d144 L16 h4 ff576 k31), batch {args.batch} lines/GPU, '
            f'96x{args.width} px, {args.classes} classes, {args.precision} inference + greedy CTC decode')


def out_frames(W):
    return (((W - 1) // 2 + 1) - 1) // 2 + 1


# ----------------------------------------------------------------------------------------------
# Algorithmic work of every kernel family for ONE pass over a (B, W) batch: SURVEY.md 8(d).
# FLOPs are 2*MAC of the reference's contractions (positional scores once per (i, j); the positional projection is
# batch-level and excluded); bytes are each activation once in / once out at the path's storage types (bf16
# activations, fp32 residual stream).  `bound` is 8(d)'s table: tensor pipe for GEMMs and attention, HBM for the
# LayerNorm / depthwise / GLU / front-end-conv / arg-max / collapse kernels.  Nothing here depends on how a kernel is
# launched (padding, fused columns, tile sizes): a kernel that moves or multiplies more than this is charged for it.
# ----------------------------------------------------------------------------------------------
def algorithmic_work(hp, B, W, C, in_bytes_per_px=4):
    d = hp['encoder_dim']
    ff = d * hp['feed_forward_expansion_factor']
    cc = hp['subsampling_conv_channels']
    K = hp['conv_kernel_size']
    H = hp['height']
    F2 = (H + 2 - 3) // 2 + 1
    F = (F2 + 2 - 3) // 2 + 1
    T2 = (W - 1) // 2 + 1
    T = (T2 - 1) // 2 + 1
    M = B * T
    act, xs = 2 * d, 4 * d                                  # bytes of one bf16 activation row / one fp32 residual row
    fam = {
        'frontend_convs': dict(flops=B * (2 * 9 * cc * F2 * T2 + 2 * 9 * cc * F * T), bytes=B * H * W * in_bytes_per_px + M * F * cc * 2, bound='hbm'),
        'gemm_frontend_pw': dict(flops=M * 2 * cc * cc * F, bytes=2 * M * F * cc * 2, bound='tensor'),
        'gemm_sub_out': dict(flops=M * 2 * (cc * F) * d, bytes=M * (F * cc * 2 + xs + act), bound='tensor'),
        'ffn_fused': dict(flops=M * 2 * (2 * d * ff), bytes=M * (act + xs + xs + act), bound='tensor'),
        'gemm_qkvx_pos': dict(flops=M * 3 * (2 * d * d), bytes=M * (act + 3 * act), bound='tensor'),
        'attention_tc': dict(flops=B * 3 * (2 * T * T * d), bytes=M * (3 * act + act), bound='tensor'),
        'gemm_attn_out': dict(flops=M * 2 * d * d, bytes=M * (act + xs + xs + act), bound='tensor'),
        'gemm_pw1_glu': dict(flops=M * 2 * d * 2 * d, bytes=M * (act + act), bound='tensor'),
        'dwconv_bn_silu_bf16': dict(flops=M * 2 * K * d, bytes=M * (act + act), bound='hbm'),
        'gemm_pw2': dict(flops=M * 2 * d * d, bytes=M * (act + xs + xs + act), bound='tensor'),
        'gemm_ctc_argmax': dict(flops=M * 2 * d * C, bytes=M * (act + 8), bound='tensor'),
        'ctc_collapse': dict(flops=0, bytes=M * (8 + 16), bound='hbm'),
    }
    fam['ffn_fused_final'] = dict(fam['ffn_fused'])
    # the one-kernel front-end (frontend_tc.cu): conv1 + depthwise + 1x1 conv; image in, feature rows out
    fam['frontend_tc'] = dict(flops=fam['frontend_convs']['flops'] + fam['gemm_frontend_pw']['flops'],
                              bytes=B * H * W * in_bytes_per_px + M * F * cc * 2, bound='hbm')
    return fam


# ----------------------------------------------------------------------------------------------
# CPU baseline (reference arm): the port of the reference's ATen-CPU operator sequence (oracle/torch_port.py)
# ----------------------------------------------------------------------------------------------
def cpu_pass_factory(args):
    """One pass of the reference's CPU implementation of the path over a bounded sample of the workload.  With `oracle/_ref`
    (the reference's own encoder modules, compiled by oracle/build_ref.py) the pass is pred.py:171-178 over those modules
    (kind 'reference'); without it, the port of their operator sequence (kind 'port')."""
    from oracle import torch_port as TP          # allowed here: cpu_baseline / --impl reference leg only
    from oracle import ref_modules as RM
    from conformer_ocr_b200 import synthetic as S
    hp = arch_hp(hparams())
    wnp = S.make_weights(hp, args.classes, seed=0)
    n = max(1, min(args.cpu_sample_lines, args.batch))
    img, lens = S.make_lines(n, hp['height'], args.width, seed=1234)
    ti, tl = torch.from_numpy(img), torch.from_numpy(lens)
    torch.set_num_threads(os.cpu_count() or 1)
    if RM.available() and not os.environ.get('COCR_BENCH_CPU_PORT'):
        net = RM.build_model(hp, args.classes, wnp).eval()

        def one_pass():
            return RM.predict_labels(net, ti, tl)
        return one_pass, n, torch.get_num_threads(), 'reference'
    w = TP.to_torch(wnp)

    def one_pass():
        return TP.predict_labels(ti, tl, w, hp)
    return one_pass, n, torch.get_num_threads(), 'port'


def time_cpu(args, reps: int, warm: int):
    one_pass, n, cores, kind = cpu_pass_factory(args)
    for _ in range(warm):
        one_pass()
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        one_pass()
        times.append(time.perf_counter() - t0)
    return n, cores, times, kind


CPU_NOTES = {
    'reference': ('the reference\'s own conformer modules (oracle/_ref: conformer_ocr.conformer.* byte-compiled from the reference tree by '
                  'oracle/build_ref.py) driven as pred.py:171-178 does: encoder -> nn.Linear -> transpose / D2H-style numpy copy -> greedy decode '
                  'per line (kraken\'s greedy_decoder is absent: restated in oracle/conformer_oracle.py)'),
    'port': ('port of the reference\'s ATen CPU operator sequence (oracle/torch_port.py; oracle/_ref is not built on this machine); the '
             'relative-position table is built once per pass, not repeated per batch row as attention.py:145 does, which flatters the '
             'reference slightly'),
}
CPU_WHAT = {'reference': 'reference modules (oracle/_ref)', 'port': 'oracle/torch_port.py'}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    n, cores, times, kind = time_cpu(args, reps=max(1, args.steps), warm=max(1, min(args.warmup, 2)))
    total = sum(times)
    lps = n * len(times) / total
    T = out_frames(args.width)
    sample = f'{n} lines of 96x{args.width} px per step ({len(times)} steps), fp32, {CPU_WHAT[kind]}, torch CPU ops, {cores} threads'
    out = dict(impl='reference', metric='text lines/sec', value=lps, unit='lines/s', n_gpus=args.gpus, steps=len(times),
               warmup=args.warmup, ms_per_step=1e3 * total / len(times), higher_is_better=True, scaling='weak',
               vs_baseline=None, dtype='f32', data='synthetic', frames_per_s=lps * T,
               config=dict(workload=workload_name(args).replace(args.precision + ' inference', 'fp32 CPU inference'),
                           sample=sample, note=CPU_NOTES[kind]),
               cpu_baseline=dict(value=lps, unit='lines/s', cores=cores, kind=kind, sample=sample),
               e2e=dict(value=lps, unit='lines/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0),
               gpu_launches=0)
    print(json.dumps(out), flush=True)


# ----------------------------------------------------------------------------------------------
# the practical bar: the reference's operator sequence run by torch on the SAME GPU, eager fp32 and bf16 autocast
# (SURVEY.md 2.3 / 8d).  oracle/torch_port.py is the port of those modules; library kernels (cuBLAS / cuDNN / ATen).
# ----------------------------------------------------------------------------------------------
def gpu_eager_baseline(args, hp, dev, x_dev, reps=3):
    from oracle import torch_port as TP          # baseline leg, same status as cpu_baseline
    from conformer_ocr_b200 import synthetic as S
    out = {}
    try:
        w = {k: v.to(dev) for k, v in TP.to_torch(S.make_weights(hp, args.classes, seed=0)).items()}
        lens = torch.full((args.batch,), args.width, dtype=torch.int64)
        tf32 = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False

        def run(autocast):
            with torch.autocast('cuda', dtype=torch.bfloat16, enabled=autocast):
                o, olens = TP.recognition_forward(x_dev, lens, w, hp)
            o = o.transpose(1, 2).cpu().float().numpy()                     # pred.py:138: full-logit D2H
            return [TP.greedy_decoder(seq[:, :int(n)]) for seq, n in zip(o, olens)]
        for name, ac in (('fp32', False), ('bf16_autocast', True)):
            run(ac)
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            for _ in range(reps):
                run(ac)
            torch.cuda.synchronize(dev)
            dt = (time.perf_counter() - t0) / reps
            out[name] = dict(value=args.batch / dt, unit='lines/s', ms_per_pass=1e3 * dt)
        torch.backends.cuda.matmul.allow_tf32 = tf32
        out['what'] = (f'oracle/torch_port.py (the reference\'s ATen operator sequence) on this GPU, eager, batch {args.batch}, full-logit '
                       f'D2H + host greedy decode as pred.py:138-143 does; {reps} passes after 1 warm-up')
    except Exception as e:                                                   # an OOM of the O(T^2) fp32 scores is a result too
        out['error'] = f'{type(e).__name__}: {str(e)[:200]}'
        torch.cuda.empty_cache()
    return out


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    from conformer_ocr_b200 import synthetic as S
    from conformer_ocr_b200.engine import Engine

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    json_fd = None
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        # NCCL prints its banner / INFO lines on the process's stdout from C; this script's stdout carries ONE JSON line, so
        # file descriptor 1 points at stderr until that line is written (NCCL_DEBUG is left as the caller set it)
        sys.stdout.flush()
        json_fd = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    assert torch.cuda.is_available(), 'bench.py (our arm) needs a CUDA device; there is no CPU fallback'
    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)

    hp = arch_hp(hparams())
    B, W, C, P = args.batch, args.width, args.classes, args.passes
    T = out_frames(W)
    eng = Engine(height=hp['height'], encoder_dim=hp['encoder_dim'], num_layers=hp['num_encoder_layers'],
                 num_heads=hp['num_attention_heads'], ff_expansion=hp['feed_forward_expansion_factor'],
                 conv_expansion=hp['conv_expansion_factor'], conv_kernel_size=hp['conv_kernel_size'],
                 half_step_residual=hp['half_step_residual'], subsampling_conv_channels=hp['subsampling_conv_channels'],
                 subsampling_factor=hp['subsampling_factor'], num_classes=C, device=dev, precision=args.precision)
    weights = S.make_weights(hp, C, seed=0)
    eng.load_state_dict(weights)

    # synthetic lines: every rank its own shard (seeded by rank); inputs resident in HBM for `value`
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    x_dev = torch.rand((B, 1, hp['height'], W), generator=g, device=dev, dtype=torch.float32)
    widths = torch.full((B,), W, dtype=torch.int32)
    x_view = x_dev.squeeze(1).transpose(1, 2)
    in_bytes = x_dev.numel() * 4
    n_pass = args.steps * P                                            # passes inside a timed region
    lines_total = world * B * n_pass

    # the shard's decoded records [count | labels] per line, kept on the device until the exchange step
    shard = torch.zeros((n_pass, B, T + 1), dtype=torch.int32, device=dev)
    gathered = torch.empty((world,) + tuple(shard.shape), dtype=torch.int32, device=dev) if world > 1 else None
    # every rank receives every shard on its device (all-gather); rank 0, the consumer of the job's output, also brings the
    # gathered block to its host
    gathered_host = torch.empty((world,) + tuple(shard.shape), dtype=torch.int32).pin_memory() if (world > 1 and rank == 0) else None

    def keep(i, res):
        shard[i, :, 0].copy_(res['counts'])
        shard[i, :, 1:].copy_(res['labels'])

    def exchange():
        """the path's one exchange step: all ranks receive the records of every shard, once per shard"""
        if world > 1:
            dist.all_gather_into_tensor(gathered, shard)
            if gathered_host is not None:
                gathered_host.copy_(gathered, non_blocking=True)

    def fence():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    for i in range(args.warmup * P):
        keep(i % n_pass, eng.forward_decode(x_view, widths))
    exchange()
    fence()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    launches0 = eng.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fence()
    e0.record()
    for i in range(n_pass):
        keep(i, eng.forward_decode(x_view, widths))
    exchange()
    e1.record()
    fence()
    ms_total = e0.elapsed_time(e1)
    launches = eng.launch_count - launches0
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    lines_per_s = lines_total / (ms_total * 1e-3)

    # ---- end to end through the host-buffer entry point (H2D + D2H inside the timed region) ----
    e2e = None
    if not args.no_e2e:
        # two pinned host batches alternate through the handle's two slots: the H2D copy of pass i+1 runs on the copy
        # stream while pass i computes; every pass's input crosses PCIe and every pass's records come back to the host.
        # At N > 1 the device copy of the records also goes into the shard buffer and is gathered once, inside the region.
        def host_loop(hosts):
            for sl in (0, 1):                               # warm both slots (their buffers are allocated on first use)
                eng.recognize_host_begin(sl, hosts[sl], widths)
            eng.recognize_host_end(0); eng.recognize_host_end(1)
            fence()
            t0 = time.perf_counter()
            pending = []
            r = None

            def finish(j, sl):
                rr = eng.recognize_host_end(sl)
                if world > 1:
                    keep(j, eng.slot_records(sl))          # D2D on the compute stream, ahead of the slot's next pass
                return rr
            for i in range(n_pass):
                slot = i & 1
                if len(pending) == 2:
                    r = finish(*pending.pop(0))
                eng.recognize_host_begin(slot, hosts[slot], widths)
                pending.append((i, slot))
            while pending:
                r = finish(*pending.pop(0))
            exchange()
            torch.cuda.synchronize(dev)
            dt = time.perf_counter() - t0
            tt = torch.tensor([dt], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            assert int(r['counts'].sum()) >= 0
            return float(tt.item())
        x_hosts = [torch.empty((B, 1, hp['height'], W), dtype=torch.float32).pin_memory() for _ in range(2)]
        for xh in x_hosts:
            xh.copy_(x_dev)
        dt = host_loop(x_hosts)
        d2h = 4 * (4 * B * T + 2 * B)
        gather_bytes = (world * shard.numel() * 4) if world > 1 else 0
        e2e = dict(value=lines_total / dt, unit='lines/s', h2d_bytes_per_step=P * (in_bytes + 4 * B),
                   d2h_bytes_per_step=P * d2h + gather_bytes // args.steps, ms_per_step=1e3 * dt / args.steps,
                   api='Engine.recognize_host_begin/_end -> cocr_recognize_host_begin/_end (pinned fp32 (N,1,H,W) in, decoded '
                       'records out; two slots, H2D of pass i+1 overlaps the forward pass of pass i)'
                       + ('; records of the whole shard all-gathered over NCCL (every rank) and copied to rank 0\'s host inside the timed region'
                          if world > 1 else ''))
        del x_hosts
        # the same loop on raw 8-bit pixels (an extension of the reference interface, reported beside
        # the like-for-like fp32 number, never instead of it): a quarter of the bytes cross PCIe
        if args.precision == 'bf16':
            u_hosts = [torch.empty((B, 1, hp['height'], W), dtype=torch.uint8).pin_memory() for _ in range(2)]
            for uh in u_hosts:
                uh.copy_((x_dev * 255).round().clamp_(0, 255).to(torch.uint8))
            dtu = host_loop(u_hosts)
            e2e['uint8_input'] = dict(value=lines_total / dtu, unit='lines/s', h2d_bytes_per_step=P * (in_bytes // 4 + 4 * B),
                                      d2h_bytes_per_step=e2e['d2h_bytes_per_step'])
            del u_hosts

    if sampler:
        time.sleep(0.2)
        sampler.stop()

    # ---- the drop-in API: PytorchRecognitionModel.predict_string_stream, pinned host batches -> Python strings ----
    e2e_api = None
    if not args.no_e2e and rank == 0:
        from conformer_ocr_b200 import PytorchRecognitionModel
        from conformer_ocr_b200.codec import LabelCodec
        kw = dict(input_dropout_p=0.1, feed_forward_dropout_p=0.1, attention_dropout_p=0.1, conv_dropout_p=0.1)
        codec = LabelCodec([chr(0x4E00 + i) for i in range(C - 1)])
        net = PytorchRecognitionModel(num_classes=C, codec=codec, precision=args.precision, **hp, **kw)
        net.nn.load_state_dict({k: torch.from_numpy(v) for k, v in weights.items()}, strict=False)
        net = net.to(dev)
        hosts = [torch.empty((B, 1, hp['height'], W), dtype=torch.float32).pin_memory() for _ in range(2)]
        for xh in hosts:
            xh.copy_(x_dev)

        def batches(n):
            for i in range(n):
                yield hosts[i & 1], widths
        for _ in net.predict_string_stream(batches(4)):
            pass
        torch.cuda.synchronize(dev)
        n_api = max(2, min(n_pass, 64))
        t0 = time.perf_counter()
        chars = 0
        for strings in net.predict_string_stream(batches(n_api)):
            chars += sum(len(s) for s in strings)
        dt = time.perf_counter() - t0
        e2e_api = dict(value=B * n_api / dt, unit='lines/s', passes=n_api, chars_decoded=chars,
                       api='PytorchRecognitionModel.predict_string_stream (pinned host batches -> Python strings; LabelCodec LUT)',
                       note='rank 0 only')
        del net, hosts

    # ---- per-kernel CUDA-event timing (rank 0): the roofline of every family against SURVEY 8(d)'s algorithmic work ----
    roofline = None
    families = {}
    if rank == 0 and args.profile_steps > 0:
        import ctypes as Ct
        lib, h = eng._lib, eng._h
        lib.cocr_set_profiling(h, 1)
        for _ in range(args.profile_steps):
            eng.forward_decode(x_view, widths)
        buf = Ct.create_string_buffer(1 << 16)
        lib.cocr_profile_report(h, buf, len(buf))
        lib.cocr_set_profiling(h, 0)
        peaks = load_peaks()
        algo = algorithmic_work(hp, B, W, C)
        traffic = {}
        tpath = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
        if os.path.exists(tpath) and (B, W) == (256, 1536):                  # the captures are of this workload
            with open(tpath) as fp:
                traffic = json.load(fp)
        step_ms = 0.0
        for line in buf.value.decode().splitlines():
            name, n, ms, fl, by = line.split()
            n_l = int(n) // args.profile_steps
            ms_p = float(ms) / args.profile_steps
            step_ms += ms_p
            f = dict(launches=n_l, ms_per_pass=ms_p, avg_launch_ms=ms_p / max(n_l, 1))
            a = algo.get(name)
            if a:
                # `a` is one launch's algorithmic work over the whole (B, W) batch
                f['bound'] = a['bound']
                f['algorithmic_flops_per_launch'] = a['flops']
                f['algorithmic_bytes_per_launch'] = a['bytes']
                sec = f['avg_launch_ms'] * 1e-3
                if a['bound'] == 'tensor':
                    f['achieved'] = a['flops'] / sec / 1e12
                    f['unit'] = 'TFLOP/s'
                    f['peak'] = peaks['tflops_sustained']
                else:
                    f['achieved'] = a['bytes'] / sec / 1e9
                    f['unit'] = 'GB/s'
                    f['peak'] = peaks['hbm_gbs']
                f['frac'] = f['achieved'] / f['peak']
                f['hbm_frac_of_algorithmic_bytes'] = a['bytes'] / sec / 1e9 / peaks['hbm_gbs']
                tr = traffic.get(name)
                if tr:
                    f['traffic'] = tr['dram_bytes_per_launch']
                    f['traffic_over_algorithmic'] = tr['dram_bytes_per_launch'] / a['bytes']
                    f['traffic_source'] = tr['source']
            f['as_launched_flops_per_launch'] = float(fl) / max(int(n), 1)
            f['as_launched_bytes_per_launch'] = float(by) / max(int(n), 1)
            families[name] = f
        for f in families.values():
            f['share_of_pass'] = f['ms_per_pass'] / max(step_ms, 1e-9)
        rated = {k: v for k, v in families.items() if 'frac' in v}
        if rated:
            top = max(rated, key=lambda k: rated[k]['ms_per_pass'])
            f = rated[top]
            roofline = dict(kernel=top, bound=f['bound'], achieved=f['achieved'], peak=f['peak'], unit=f['unit'], frac=f['frac'],
                            traffic=f.get('traffic'), traffic_over_algorithmic=f.get('traffic_over_algorithmic'),
                            traffic_source=f.get('traffic_source'),
                            algorithmic_flops_per_launch=f['algorithmic_flops_per_launch'],
                            algorithmic_bytes_per_launch=f['algorithmic_bytes_per_launch'],
                            peak_source=peaks['source'] + (' (sustained bf16 cuBLAS: the kernel is timed inside a long step)'
                                                           if f['bound'] == 'tensor' else ' (copy)'),
                            launches_per_pass=f['launches'], avg_launch_ms=f['avg_launch_ms'], share_of_pass=f['share_of_pass'],
                            how=f'dominant kernel family by CUDA-event time (events around every launch over {args.profile_steps} extra '
                                'passes); algorithmic FLOPs / bytes per SURVEY.md 8(d), not the launch\'s padded operands')

    eager = None
    if rank == 0 and world == 1 and not args.no_eager_baseline:
        eager = gpu_eager_baseline(args, hp, dev, x_dev)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # whole-step roofline against the tensor-core peak (SURVEY.md 8d algorithmic FLOPs per line)
    peaks = load_peaks()
    fl_line = S.flops_per_line(W, C, hp)
    step_tflops = fl_line * B * P / (ms_step * 1e-3) / 1e12

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:         # rank 0 at N = 1 only (the other ranks' barrier spins would share the cores)
        n, cores, times, kind = time_cpu(args, reps=3, warm=1)
        v = n * len(times) / sum(times)
        cpu_baseline = dict(value=v, unit='lines/s', cores=cores, kind=kind,
                            sample=f'{n} lines of 96x{W} px x {len(times)} passes, fp32, {CPU_WHAT[kind]}, {cores} threads',
                            note=CPU_NOTES[kind])

    clocks = sampler.summary() if sampler else dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
    out = dict(metric='text lines/sec', value=lines_per_s, unit='lines/s', n_gpus=world, steps=args.steps,
               warmup=args.warmup, ms_per_step=ms_step, ms_per_pass=ms_step / P, higher_is_better=True, scaling='weak',
               vs_baseline=None, dtype='bf16' if args.precision == 'bf16' else 'f32', data='synthetic',
               frames_per_s=lines_per_s * T,
               config=dict(workload=workload_name(args), lines_per_pass_per_gpu=B, passes_per_step=P, lines_per_step_per_gpu=B * P,
                           lines_in_timed_region=lines_total, width_px=W, frames_per_line=T, classes=C,
                           parallelism=(f'dp{world}: lines sharded, no data-path collective; the [count | labels] records of the whole timed '
                                        f'region ({lines_total} lines) are all-gathered over NCCL once, inside the timed region')
                           if world > 1 else 'single GPU',
                           l2='input batch is %.0f MB > 126 MB L2; no explicit flush' % (in_bytes / 1e6) if in_bytes > 126e6 else
                              'workspace of a pass (activations of 16 layers) exceeds the 126 MB L2; no explicit flush',
                           mask_mode='reference_unmasked'),
               clocks=clocks, e2e=e2e, e2e_api=e2e_api, gpu_launches=launches,
               roofline=roofline,
               roofline_step=dict(bound='tensor', achieved=step_tflops, peak=peaks['tflops_sustained'], unit='TFLOP/s',
                                  frac=step_tflops / peaks['tflops_sustained'], frac_of_burst=step_tflops / peaks['tflops_burst'],
                                  gflop_per_line=fl_line / 1e9, peak_source=peaks['source'] + ' (sustained bf16 cuBLAS)'),
               kernel_families=families, gpu_eager_baseline=eager, cpu_baseline=cpu_baseline)
    if json_fd is not None:
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(out) + '\n').encode())
        os.close(json_fd)
    else:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------
# configs[3]: the training step
# ----------------------------------------------------------------------------------------------
def train_batch(args, hp, seed, device=None):
    """SURVEY.md 8d cfg 4: B lines of 96 x W px, targets of length U{20..80}, labels U{1..C-1}."""
    rng = np.random.RandomState(seed)
    B, W, C = args.batch, args.width, args.classes
    tl = rng.randint(20, 81, size=B).astype(np.int64)
    tg = rng.randint(1, C, size=int(tl.sum())).astype(np.int64)
    g = torch.Generator().manual_seed(seed)
    img = torch.rand((B, 1, hp['height'], W), generator=g, dtype=torch.float32)
    return {'image': img, 'seq_lens': torch.full((B,), W, dtype=torch.int64), 'target': torch.from_numpy(tg), 'target_lens': torch.from_numpy(tl)}


def train_flops_per_line(W, C, hp):
    """forward + data gradients + weight gradients = 3 x the forward contractions (attention: 3 forward, 6 backward)."""
    from conformer_ocr_b200 import synthetic as S
    return 3.0 * S.flops_per_line(W, C, hp)


def port_train_step_factory(args, hp, device, n_lines, autocast=False):
    """The reference's training step as torch runs it (oracle/torch_train_port.py: train()-mode modules, autograd,
    torch.optim.AdamW) - baseline leg only."""
    from oracle import torch_train_port as TTP
    from conformer_ocr_b200 import synthetic as S
    w = {k: torch.from_numpy(v.copy()).to(device) for k, v in S.make_weights(hp, args.classes, seed=0).items()}
    for k, v in w.items():
        v.requires_grad_(not k.endswith(('running_mean', 'running_var')))
    opt = torch.optim.AdamW([v for v in w.values() if v.requires_grad], lr=1e-3, weight_decay=1e-3)
    a = argparse.Namespace(**vars(args))
    a.batch = n_lines
    b = train_batch(a, hp, seed=77)
    img, tg = b['image'].to(device), b['target'].to(device)

    def one_step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast(device.type, dtype=torch.bfloat16, enabled=autocast):
            probits, olens, _ = TTP.train_forward(img, b['seq_lens'], w, hp, dropout=0.1)
        logp = torch.nn.functional.log_softmax(probits.float(), dim=-1)
        loss = torch.nn.functional.ctc_loss(logp.transpose(0, 1), tg, olens, b['target_lens'], blank=0, reduction='sum', zero_infinity=True)
        loss.backward()
        opt.step()
        return float(loss.item())
    return one_step


def ref_train_step_factory(args, hp, n_lines):
    """The reference's training step over its own modules (oracle/_ref) on the host cores: model.py:129-145 + backward +
    torch.optim.AdamW, train() mode with the default dropout 0.1 - baseline leg only."""
    from oracle import ref_modules as RM
    from conformer_ocr_b200 import synthetic as S
    net = RM.build_model(hp, args.classes, S.make_weights(hp, args.classes, seed=0), dropout=0.1)
    a = argparse.Namespace(**vars(args))
    a.batch = n_lines
    return RM.train_step_factory(net, train_batch(a, hp, seed=77), lr=1e-3, weight_decay=1e-3)


def time_cpu_train(args, hp, reps, warm):
    from oracle import ref_modules as RM
    torch.set_num_threads(os.cpu_count() or 1)
    n = max(1, min(4, args.batch))
    kind = 'reference' if RM.available() and not os.environ.get('COCR_BENCH_CPU_PORT') else 'port'
    step = ref_train_step_factory(args, hp, n) if kind == 'reference' else port_train_step_factory(args, hp, torch.device('cpu'), n)
    for _ in range(warm):
        step()
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    return n, torch.get_num_threads(), times, kind


TRAIN_CPU_NOTES = {
    'reference': ('the reference\'s training step on the host cores over its own modules in train() mode (oracle/_ref, byte-compiled from the '
                  'reference tree): encoder -> decoder -> log_softmax -> CTCLoss(sum, zero_infinity) as model.py:129-145, autograd, '
                  'torch.optim.AdamW; dropout 0.1'),
    'port': ('the reference\'s training step as torch runs it on the host cores: train()-mode port of its modules '
             '(oracle/torch_train_port.py, pinned to the unmodified reference modules by tests/golden/train_*.npz; oracle/_ref is not built '
             'on this machine), autograd, torch.optim.AdamW; dropout 0.1'),
}


def run_reference_train(args):
    if int(os.environ.get('RANK', '0')) != 0:
        return
    hp = arch_hp(hparams())
    n, cores, times, kind = time_cpu_train(args, hp, reps=max(1, args.steps), warm=1)
    total = sum(times)
    lps = n * len(times) / total
    sample = f'{n} lines of 96x{args.width} px per step ({len(times)} steps), fp32, {CPU_WHAT[kind]}, torch CPU ops + autograd + AdamW, {cores} threads'
    out = dict(impl='reference', metric='training lines/sec', value=lps, unit='lines/s', n_gpus=args.gpus, steps=len(times), warmup=args.warmup,
               ms_per_step=1e3 * total / len(times), higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32', data='synthetic',
               config=dict(workload=train_workload_name(args), sample=sample, note=TRAIN_CPU_NOTES[kind]),
               cpu_baseline=dict(value=lps, unit='lines/s', cores=cores, kind=kind, sample=sample),
               e2e=dict(value=lps, unit='lines/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(out), flush=True)


def train_workload_name(args):
    return (f'BASELINE {CONFIGS[4]["name"]}: training step of the default_specs conformer (d144 L16 h4 ff576 k31), {args.batch} lines/GPU of '
            f'96x{args.width} px, {args.classes} classes, targets U{{20..80}}, dropout 0.1, CTC loss (sum) + backward + gradient all-reduce + AdamW, ' + ('bf16 tcgen05 operands, fp32 accumulation / master weights' if args.precision == 'bf16' else 'fp32 CUDA-core kernels'))


def run_train(args):
    import torch.distributed as dist
    from conformer_ocr_b200 import synthetic as S
    from conformer_ocr_b200.train import Trainer

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    json_fd = None
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        sys.stdout.flush()
        json_fd = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    assert torch.cuda.is_available(), 'bench.py (our arm) needs a CUDA device; there is no CPU fallback'
    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    hp = arch_hp(hparams())
    B, W, C = args.batch, args.width, args.classes
    T = out_frames(W)
    tr = Trainer(num_classes=C, device=dev, lr=1e-3, weight_decay=1e-3, warmup=0, seed=1, input_dropout_p=0.1, feed_forward_dropout_p=0.1,
                 attention_dropout_p=0.1, conv_dropout_p=0.1, precision=args.precision, **hp)
    tr.load_state_dict(S.make_weights(hp, C, seed=0))                    # every rank the same initial weights (DDP broadcast)
    host = train_batch(args, hp, seed=1234 + rank)                       # every rank its own shard of the global batch
    dev_batch = {k: v.to(dev) for k, v in host.items()}
    in_bytes = host['image'].numel() * 4 + host['target'].numel() * 4 + 2 * B * 4

    def fence():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    first_loss = None
    for i in range(args.warmup):
        lw = tr.step(dev_batch)
        if i == 0:
            first_loss = float(lw.item())        # loss of the initial weights: with `final_loss` it shows the timed steps train
    fence()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    launches0 = tr.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fence()
    e0.record()
    for _ in range(args.steps):
        loss = tr.step(dev_batch)
    e1.record()
    fence()
    ms_total = e0.elapsed_time(e1)
    launches = tr.launch_count - launches0
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    lines_per_s = world * B * args.steps / (ms_total * 1e-3)
    last_loss = float(loss.item())

    # ---- end to end: the batch comes from pinned host memory every step, the loss goes back to the host every step ----
    e2e = None
    if not args.no_e2e:
        pinned = {k: v.pin_memory() for k, v in host.items()}
        for _ in range(2):
            tr.step(pinned)                      # the staging buffers are allocated (and the step re-captured for them) here
        fence()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            lv = tr.step(pinned).item()          # H2D of the pinned host batch into the trainer's staging buffers; D2H of the loss
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        assert np.isfinite(lv)
        e2e = dict(value=world * B * args.steps / dt, unit='lines/s', h2d_bytes_per_step=in_bytes, d2h_bytes_per_step=4, ms_per_step=1e3 * dt / args.steps,
                   api='Trainer.step(batch): pinned host batch (image, seq_lens, target, target_lens as the reference\'s collate hands them to '
                       '_step) -> cocr_train_step -> all-reduce -> cocr_adamw_step -> loss.item()')
    if sampler:
        time.sleep(0.2)
        sampler.stop()

    # ---- per-kernel CUDA-event timing (rank 0) ----
    families, roofline = {}, None
    if rank == 0 and args.profile_steps > 0:
        import ctypes as Ct
        lib, h = tr._lib, tr._h
        lib.cocr_set_profiling(h, 1)
        for _ in range(args.profile_steps):
            tr.forward_backward(dev_batch)       # rank 0 alone: no collective in the profiling passes
            tr.optimizer_step()
        buf = Ct.create_string_buffer(1 << 16)
        lib.cocr_profile_report(h, buf, len(buf))
        lib.cocr_set_profiling(h, 0)
        peaks = load_peaks()
        tot = 0.0
        for line in buf.value.decode().splitlines():
            name, n, ms, fl, by = line.split()
            n_l = int(n) // args.profile_steps
            ms_p = float(ms) / args.profile_steps
            tot += ms_p
            f = dict(launches=n_l, ms_per_step=ms_p, avg_launch_ms=ms_p / max(n_l, 1), flops_per_launch=float(fl) / max(int(n), 1),
                     bytes_per_launch=float(by) / max(int(n), 1))
            sec = f['avg_launch_ms'] * 1e-3
            dense = f['flops_per_launch'] > 0 and (name.startswith('train_tc_') or 'gemm' in name or name.startswith('train_att_') or name in (
                'train_ffn_up', 'train_ffn_down', 'train_qkv', 'train_att_out', 'train_pw1', 'train_pw2', 'train_decoder', 'train_fe_pw', 'train_fe_out',
                'train_pos_proj')) and name not in ('train_tc_att_softmax', 'train_tc_att_ds', 'train_tc_att_prep', 'train_tc_weight_cast', 'train_att_softmax')
            if dense:       # tensor pipe for the dense contractions
                f.update(bound='tensor', achieved=f['flops_per_launch'] / sec / 1e12, unit='TFLOP/s', peak=peaks['tflops_sustained'], kernel='k_bgemm_tc' if name.startswith('train_tc_') else 'cuda-core')
            else:
                f.update(bound='hbm', achieved=f['bytes_per_launch'] / sec / 1e9, unit='GB/s', peak=peaks['hbm_gbs'])
            f['frac'] = f['achieved'] / f['peak']
            families[name] = f
        for f in families.values():
            f['share_of_step'] = f['ms_per_step'] / max(tot, 1e-9)
        # the dominant KERNEL of the tensor-core mode is the batched GEMM (every dense contraction is a launch of it): rate it as a whole
        bg = [f for f in families.values() if f.get('kernel') == 'k_bgemm_tc']
        if bg:
            ms = sum(f['ms_per_step'] for f in bg)
            fl = sum(f['flops_per_launch'] * f['launches'] for f in bg)
            n = sum(f['launches'] for f in bg)
            ach = fl / (ms * 1e-3) / 1e12
            roofline = dict(kernel='k_bgemm_tc (all dense contractions of the step: linears, data / weight gradients, attention)', bound='tensor',
                            achieved=ach, peak=peaks['tflops_sustained'], unit='TFLOP/s', frac=ach / peaks['tflops_sustained'], traffic=None,
                            launches_per_step=n, avg_launch_ms=ms / max(n, 1), share_of_step=ms / max(tot, 1e-9), flops_per_step=fl,
                            peak_source=peaks['source'] + ' (sustained bf16 cuBLAS)',
                            how='all launches of the batched tcgen05 GEMM by CUDA-event time; exact 2*MAC FLOPs of each launch (extents are not padded). '
                                'At 32 lines per GPU the contractions are small (M = 12 288 frames, K or N = 144): a launch lasts 15-45 us and is bound by '
                                'writing its fp32 result (a 12 288 x 576 output is 28 MB), not by the tensor pipe')
        else:
            top = max(families, key=lambda k: families[k]['ms_per_step'])
            f = families[top]
            roofline = dict(kernel=top, bound=f['bound'], achieved=f['achieved'], peak=f['peak'], unit=f['unit'], frac=f['frac'], traffic=None,
                            launches_per_step=f['launches'], avg_launch_ms=f['avg_launch_ms'], share_of_step=f['share_of_step'],
                            peak_source=peaks['source'], how='dominant kernel family of the training step by CUDA-event time; exact 2*MAC FLOPs / '
                            'operand bytes of each launch')

    eager = None
    if rank == 0 and world == 1 and not args.no_eager_baseline:
        eager = {}
        try:
            tf32 = torch.backends.cuda.matmul.allow_tf32
            torch.backends.cuda.matmul.allow_tf32 = False
            for name, ac in (('fp32', False), ('bf16_autocast', True)):
                step = port_train_step_factory(args, hp, dev, B, autocast=ac)
                step()
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                for _ in range(3):
                    step()
                torch.cuda.synchronize(dev)
                dt = (time.perf_counter() - t0) / 3
                eager[name] = dict(value=B / dt, unit='lines/s', ms_per_step=1e3 * dt)
                del step
                torch.cuda.empty_cache()
            torch.backends.cuda.matmul.allow_tf32 = tf32
            eager['what'] = (f'oracle/torch_train_port.py (the reference\'s modules in train() mode) + autograd + torch.optim.AdamW on this GPU, '
                             f'eager, batch {B}; 3 steps after 1 warm-up')
        except Exception as e:
            eager['error'] = f'{type(e).__name__}: {str(e)[:200]}'
            torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    peaks = load_peaks()
    fl_line = train_flops_per_line(W, C, hp)
    step_tflops = fl_line * B / (ms_step * 1e-3) / 1e12
    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        n, cores, times, kind = time_cpu_train(args, hp, reps=2, warm=1)
        cpu_baseline = dict(value=n * len(times) / sum(times), unit='lines/s', cores=cores, kind=kind,
                            sample=f'{n} lines of 96x{W} px x {len(times)} steps, fp32 forward + backward + AdamW, {CPU_WHAT[kind]}, {cores} threads',
                            note=TRAIN_CPU_NOTES[kind])
    clocks = sampler.summary() if sampler else dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
    out = dict(metric='training lines/sec', value=lines_per_s, unit='lines/s', n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=ms_step,
               higher_is_better=True, scaling='weak', vs_baseline=None, dtype='bf16' if args.precision == 'bf16' else 'f32', data='synthetic', frames_per_s=lines_per_s * T,
               config=dict(workload=train_workload_name(args), lines_per_step_per_gpu=B, global_batch=world * B, width_px=W, frames_per_line=T,
                           classes=C, parameters=tr.n_trainable, workspace_bytes=tr.workspace_bytes(B, W), first_loss=first_loss, final_loss=last_loss,
                           parallelism=(f'dp{world}: every rank its own {B} lines; ONE NCCL all-reduce(sum) of the flat gradient buffer '
                                        f'({tr.n_trainable * 4 / 1e6:.1f} MB) per step; BatchNorm statistics per rank') if world > 1 else 'single GPU',
                           l2='activations saved for the backward pass (%.1f GB) exceed the 126 MB L2; no explicit flush' % (tr.workspace_bytes(B, W) / 1e9)),
               clocks=clocks, e2e=e2e, gpu_launches=launches, roofline=roofline,
               roofline_step=dict(bound='tensor', achieved=step_tflops, peak=peaks['tflops_sustained'], unit='TFLOP/s',
                                  frac=step_tflops / peaks['tflops_sustained'], gflop_per_line=fl_line / 1e9,
                                  note='3 x SURVEY 8(d) forward FLOPs per line (forward + data gradients + weight gradients)'),
               kernel_families=families, gpu_eager_baseline=eager, cpu_baseline=cpu_baseline)
    if json_fd is not None:
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(out) + '\n').encode())
        os.close(json_fd)
    else:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.config == 4:
        (run_reference_train if args.impl == 'reference' else run_train)(args)
    elif args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
