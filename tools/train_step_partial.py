#!/usr/bin/env python
"""BASELINE configs[3] (training step, batch 32, CTC loss + gradient all-reduce), the part that is built:
encoder forward -> decoder -> fused log-softmax + CTC loss -> d loss / d logits -> decoder backward -> ONE NCCL all-reduce of
[loss | grad_weight | grad_bias].  The backward pass through the encoder is NOT built, so this is not a training-step
benchmark; it checks the exchange step on real GPUs (all-reduced gradients == sum of the per-rank gradients) and reports
the time of what exists.  Run: python -m torch.distributed.run --nproc-per-node N tools/train_step_partial.py"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from conformer_ocr_b200 import synthetic as S
from conformer_ocr_b200.default_specs import RECOGNITION_HYPER_PARAMS as HP
from conformer_ocr_b200.model import RecognitionModel

world = int(os.environ.get('WORLD_SIZE', '1'))
rank = int(os.environ.get('RANK', '0'))
local = int(os.environ.get('LOCAL_RANK', '0'))
dev = torch.device('cuda', local)
torch.cuda.set_device(dev)
if world > 1:
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    dist.init_process_group('nccl', device_id=dev)
hp = {k: HP[k] for k in ('height', 'encoder_dim', 'num_encoder_layers', 'num_attention_heads', 'feed_forward_expansion_factor',
                         'conv_expansion_factor', 'conv_kernel_size', 'half_step_residual', 'subsampling_conv_channels', 'subsampling_factor')}
C, W, B_global = 256, 1536, 32
B = B_global // world
net = RecognitionModel(num_classes=C, precision='bf16', **hp)
net.nn.load_state_dict({k: torch.from_numpy(v) for k, v in S.make_weights(hp, C, seed=0).items()}, strict=False)
net = net.to(dev)
rng = np.random.RandomState(1000 + rank)
img, lens = S.make_lines(B, hp['height'], W, seed=2000 + rank)
tl = rng.randint(20, 81, size=B)
batch = dict(image=torch.from_numpy(img).to(dev), seq_lens=torch.from_numpy(lens), target=torch.from_numpy(rng.randint(1, C, size=int(tl.sum()))),
             target_lens=torch.from_numpy(tl))
for _ in range(3):
    out = net.decoder_training_step(batch)
torch.cuda.synchronize(dev)
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
steps = 20
for _ in range(steps):
    out = net.decoder_training_step(batch)
torch.cuda.synchronize(dev)
dt = (time.perf_counter() - t0) / steps
ok = True
if world > 1:
    parts = [torch.empty_like(out['grad_decoder_weight']) for _ in range(world)]
    dist.all_gather(parts, out['grad_decoder_weight'])
    want = torch.stack(parts).sum(0)
    err = float((want - out['grad_decoder_weight_allreduced']).abs().max() / want.abs().max())
    losses = [torch.empty_like(out['loss']) for _ in range(world)]
    dist.all_gather(losses, out['loss'])
    lerr = abs(float(torch.stack(losses).sum()) - float(out['loss_allreduced'])) / float(out['loss_allreduced'])
    ok = err < 1e-6 and lerr < 1e-6
else:
    err = lerr = 0.0
if rank == 0:
    print(json.dumps(dict(what='partial training step (no encoder backward): forward + CTC loss + decoder backward + all-reduce',
                          n_gpus=world, global_batch=B_global, per_gpu_batch=B, width_px=W, classes=C, ms_per_step=1e3 * dt,
                          loss_sum_allreduced=float(out['loss_allreduced']), allreduce_rel_err_vs_sum_of_ranks=err, loss_rel_err=lerr,
                          allreduce_matches_sum=ok)))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
sys.exit(0 if ok else 1)
