#!/usr/bin/env python
"""Benchmark of the conformer_ocr hot path (encoder forward -> CTC projection -> greedy decode).

    python bench.py --gpus N --steps K --warmup W            # our arm (libcocr_b200.so on N B200s)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (port), host cores

One "step" = one pass of the hot path over one batch of synthetic text lines.  Workload at N=1 is
BASELINE.json configs[1]: default-spec conformer, batch 256 lines, 96 x 1536 px, 256 classes, bf16.
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the meaning of every key.
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

DEFAULT_PRECISION = 'bf16'     # BASELINE.json configs[1]: bf16 inference (fp32 is the parity-grade path)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', choices=['ours', 'reference'], default='ours')
    ap.add_argument('--precision', choices=['fp32', 'bf16'], default=DEFAULT_PRECISION)
    ap.add_argument('--batch', type=int, default=256, help='lines per step per GPU')
    ap.add_argument('--width', type=int, default=1536)
    ap.add_argument('--classes', type=int, default=256)
    ap.add_argument('--cpu-sample-lines', type=int, default=16, help='lines per CPU-baseline pass')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--profile-steps', type=int, default=2, help='extra steps with per-kernel CUDA events (roofline)')
    return ap.parse_args()


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
    return (f'default_specs conformer (d144 L16 h4 ff576 k31), batch {args.batch} lines/GPU, 96x{args.width} px, '
            f'{args.classes} classes, {args.precision} inference + greedy CTC decode')


# ----------------------------------------------------------------------------------------------
# CPU baseline (reference arm): the port of the reference's ATen-CPU operator sequence (oracle/torch_port.py)
# ----------------------------------------------------------------------------------------------
def cpu_pass_factory(args):
    from oracle import torch_port as TP          # allowed here: cpu_baseline / --impl reference leg only
    from conformer_ocr_b200 import synthetic as S
    hp = arch_hp(hparams())
    w = TP.to_torch(S.make_weights(hp, args.classes, seed=0))
    n = max(1, min(args.cpu_sample_lines, args.batch))
    img, lens = S.make_lines(n, hp['height'], args.width, seed=1234)
    ti, tl = torch.from_numpy(img), torch.from_numpy(lens)
    torch.set_num_threads(os.cpu_count() or 1)

    def one_pass():
        return TP.predict_labels(ti, tl, w, hp)
    return one_pass, n, torch.get_num_threads()


def time_cpu(args, reps: int, warm: int):
    one_pass, n, cores = cpu_pass_factory(args)
    for _ in range(warm):
        one_pass()
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        one_pass()
        times.append(time.perf_counter() - t0)
    return n, cores, times


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    n, cores, times = time_cpu(args, reps=max(1, args.steps), warm=max(1, min(args.warmup, 2)))
    total = sum(times)
    lps = n * len(times) / total
    T = (((args.width - 1) // 2 + 1) - 1) // 2 + 1
    sample = f'{n} lines of 96x{args.width} px per step ({len(times)} steps), fp32, torch CPU ops, {cores} threads'
    out = dict(impl='reference', metric='text lines/sec', value=lps, unit='lines/s', n_gpus=args.gpus, steps=len(times),
               warmup=args.warmup, ms_per_step=1e3 * total / len(times), higher_is_better=True, scaling='weak',
               vs_baseline=None, dtype='f32', data='synthetic', frames_per_s=lps * T,
               config=dict(workload=workload_name(args).replace(args.precision + ' inference', 'fp32 CPU inference'),
                           sample=sample),
               cpu_baseline=dict(value=lps, unit='lines/s', cores=cores, kind='port', sample=sample),
               e2e=dict(value=lps, unit='lines/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0),
               gpu_launches=0)
    print(json.dumps(out), flush=True)


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
        if os.environ.get('NCCL_DEBUG', '').upper() in ('', 'VERSION', 'INFO'):
            os.environ['NCCL_DEBUG'] = 'WARN'
        # NCCL prints its version banner on the process's stdout from C (whatever NCCL_DEBUG says on some boxes); this
        # script's stdout carries ONE JSON line, so file descriptor 1 points at stderr until that line is written
        sys.stdout.flush()
        json_fd = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    assert torch.cuda.is_available(), 'bench.py (our arm) needs a CUDA device; there is no CPU fallback'
    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)

    hp = arch_hp(hparams())
    B, W, C = args.batch, args.width, args.classes
    T = (((W - 1) // 2 + 1) - 1) // 2 + 1
    eng = Engine(height=hp['height'], encoder_dim=hp['encoder_dim'], num_layers=hp['num_encoder_layers'],
                 num_heads=hp['num_attention_heads'], ff_expansion=hp['feed_forward_expansion_factor'],
                 conv_expansion=hp['conv_expansion_factor'], conv_kernel_size=hp['conv_kernel_size'],
                 half_step_residual=hp['half_step_residual'], subsampling_conv_channels=hp['subsampling_conv_channels'],
                 subsampling_factor=hp['subsampling_factor'], num_classes=C, device=dev, precision=args.precision)
    eng.load_state_dict(S.make_weights(hp, C, seed=0))

    # synthetic lines: every rank its own shard (seeded by rank); inputs resident in HBM for `value`
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    x_dev = torch.rand((B, 1, hp['height'], W), generator=g, device=dev, dtype=torch.float32)
    widths = torch.full((B,), W, dtype=torch.int32)
    x_view = x_dev.squeeze(1).transpose(1, 2)
    in_bytes = x_dev.numel() * 4

    def step():
        res = eng.forward_decode(x_view, widths)
        if world > 1:
            # the path's one exchange step: gather the collapsed label records of every shard (NCCL/NVLink)
            packed = torch.cat([res['counts'].view(B, 1), res['labels']], dim=1)
            out = torch.empty((world,) + tuple(packed.shape), dtype=packed.dtype, device=dev)
            dist.all_gather_into_tensor(out, packed)
            return out
        return res

    def fence():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        step()
    fence()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    launches0 = eng.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fence()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    fence()
    ms_total = e0.elapsed_time(e1)
    launches = eng.launch_count - launches0
    if sampler:
        time.sleep(0.2)
        sampler.stop()
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    lines_per_s = world * B * args.steps / (ms_total * 1e-3)

    # ---- end to end through the host-buffer entry point (H2D + D2H inside the timed region) ----
    e2e = None
    if not args.no_e2e:
        # two pinned host batches alternate through the handle's two slots: the H2D copy of step i+1 runs on the copy
        # stream while step i computes; every step's input crosses PCIe and every step's records come back to the host
        x_hosts = [torch.empty((B, 1, hp['height'], W), dtype=torch.float32).pin_memory() for _ in range(2)]
        for xh in x_hosts:
            xh.copy_(x_dev)
        for _ in range(2):                                  # warm both slots (their buffers are allocated on first use)
            eng.recognize_host_begin(0, x_hosts[0], widths)
            eng.recognize_host_begin(1, x_hosts[1], widths)
            eng.recognize_host_end(0)
            eng.recognize_host_end(1)
        fence()
        t0 = time.perf_counter()
        pending = None
        for i in range(args.steps):
            slot = i & 1
            eng.recognize_host_begin(slot, x_hosts[slot], widths)
            if pending is not None:
                r = eng.recognize_host_end(pending)
            pending = slot
        r = eng.recognize_host_end(pending)
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        d2h = 4 * (4 * B * T + 2 * B)
        e2e = dict(value=world * B * args.steps / dt, unit='lines/s', h2d_bytes_per_step=in_bytes + 4 * B,
                   d2h_bytes_per_step=d2h, ms_per_step=1e3 * dt / args.steps,
                   api='Engine.recognize_host_begin/_end -> cocr_recognize_host_begin/_end (pinned fp32 (N,1,H,W) in, decoded '
                       'records out; two slots, H2D of step i+1 overlaps the forward pass of step i)')
        assert int(r['counts'].sum()) >= 0
        # the same loop on raw 8-bit pixels (an extension of the reference interface, reported beside
        # the like-for-like fp32 number, never instead of it): a quarter of the bytes cross PCIe
        if args.precision == 'bf16':
            u_hosts = [torch.empty((B, 1, hp['height'], W), dtype=torch.uint8).pin_memory() for _ in range(2)]
            for uh in u_hosts:
                uh.copy_((x_dev * 255).round().clamp_(0, 255).to(torch.uint8))
            for sl in (0, 1):
                eng.recognize_host_begin(sl, u_hosts[sl], widths)
            eng.recognize_host_end(0); eng.recognize_host_end(1)
            fence()
            t0 = time.perf_counter()
            pending = None
            for i in range(args.steps):
                slot = i & 1
                eng.recognize_host_begin(slot, u_hosts[slot], widths)
                if pending is not None:
                    eng.recognize_host_end(pending)
                pending = slot
            eng.recognize_host_end(pending)
            torch.cuda.synchronize(dev)
            dtu = time.perf_counter() - t0
            tt = torch.tensor([dtu], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            e2e['uint8_input'] = dict(value=world * B * args.steps / float(tt.item()), unit='lines/s',
                                      h2d_bytes_per_step=in_bytes // 4 + 4 * B, d2h_bytes_per_step=d2h)

    # ---- per-kernel CUDA-event timing for the roofline of the dominant kernel (rank 0) ----
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
        for line in buf.value.decode().splitlines():
            name, n, ms, fl, by = line.split()
            families[name] = dict(launches=int(n) // args.profile_steps, ms_per_step=float(ms) / args.profile_steps,
                                  flops_per_step=float(fl) / args.profile_steps, bytes_per_step=float(by) / args.profile_steps)
        peaks = load_peaks()
        if families:
            top = max(families, key=lambda k: families[k]['ms_per_step'])
            f = families[top]
            per_launch_ms = f['ms_per_step'] / max(f['launches'], 1)
            tensor_bound = f['flops_per_step'] > 0 and (f['flops_per_step'] / max(f['bytes_per_step'], 1.0)) > (
                peaks['tflops_sustained'] * 1e12 / (peaks['hbm_gbs'] * 1e9))
            if tensor_bound:
                ach = f['flops_per_step'] / (f['ms_per_step'] * 1e-3) / 1e12
                roofline = dict(kernel=top, bound='tensor', achieved=ach, peak=peaks['tflops_sustained'], unit='TFLOP/s',
                                frac=ach / peaks['tflops_sustained'], traffic=None)
            else:
                ach = f['bytes_per_step'] / (f['ms_per_step'] * 1e-3) / 1e9
                roofline = dict(kernel=top, bound='hbm', achieved=ach, peak=peaks['hbm_gbs'], unit='GB/s',
                                frac=ach / peaks['hbm_gbs'], traffic=None)
            tpath = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
            if os.path.exists(tpath):
                with open(tpath) as fp:
                    tr = json.load(fp).get(top)
                if tr and (args.batch, args.width) == (256, 1536):      # the capture is of this workload
                    roofline['traffic'] = tr['dram_bytes_per_launch']
                    roofline['traffic_source'] = tr['source']
                    roofline['algorithmic_bytes_per_launch'] = f['bytes_per_step'] / max(f['launches'], 1)
            roofline.update(peak_source=peaks['source'] + (' (sustained bf16 cuBLAS)' if tensor_bound else ' (copy)'),
                            launches_per_step=f['launches'], avg_launch_ms=per_launch_ms,
                            share_of_step=f['ms_per_step'] / max(sum(v['ms_per_step'] for v in families.values()), 1e-9),
                            how=f'CUDA events around every launch of the family over {args.profile_steps} extra steps')

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # whole-step roofline against the tensor-core peak (SURVEY.md §8d algorithmic FLOPs per line)
    peaks = load_peaks()
    fl_line = S.flops_per_line(W, C, hp)
    step_tflops = fl_line * B / (ms_step * 1e-3) / 1e12

    cpu_baseline = None
    if not args.no_cpu_baseline:
        n, cores, times = time_cpu(args, reps=3, warm=1)
        v = n * len(times) / sum(times)
        cpu_baseline = dict(value=v, unit='lines/s', cores=cores, kind='port',
                            sample=f'{n} lines of 96x{W} px x {len(times)} passes, fp32, oracle/torch_port.py (the reference\'s '
                                   f'ATen CPU operator sequence), {cores} threads')

    clocks = sampler.summary() if sampler else dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
    out = dict(metric='text lines/sec', value=lines_per_s, unit='lines/s', n_gpus=world, steps=args.steps,
               warmup=args.warmup, ms_per_step=ms_step, higher_is_better=True, scaling='weak', vs_baseline=None,
               dtype='bf16' if args.precision == 'bf16' else 'f32', data='synthetic',
               frames_per_s=lines_per_s * T,
               config=dict(workload=workload_name(args), lines_per_gpu=B, width_px=W, frames_per_line=T, classes=C,
                           parallelism=f'dp{world} (lines sharded, decoded records all-gathered over NCCL)' if world > 1 else 'single GPU',
                           l2='input batch is %.0f MB > 126 MB L2; no explicit flush' % (in_bytes / 1e6),
                           mask_mode='reference_unmasked'),
               clocks=clocks, e2e=e2e, gpu_launches=launches,
               roofline=roofline,
               roofline_step=dict(bound='tensor', achieved=step_tflops, peak=peaks['tflops_sustained'], unit='TFLOP/s',
                                  frac=step_tflops / peaks['tflops_sustained'], gflop_per_line=fl_line / 1e9,
                                  peak_source=peaks['source'] + ' (sustained bf16 cuBLAS)'),
               kernel_families=families, cpu_baseline=cpu_baseline)
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
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
