#!/usr/bin/env python
"""N GPUs of one box (torchrun): every rank runs the training step on its own 32/N lines, the flat gradient buffers are all-reduced
(NCCL) and compared with the sum of the per-rank buffers gathered separately; then three optimiser steps, after which the
parameters of all ranks must be identical (same initial weights, same summed gradients, BatchNorm statistics aside).
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/train_allreduce_check.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from conformer_ocr_b200 import synthetic as S
from conformer_ocr_b200.default_specs import RECOGNITION_HYPER_PARAMS as HP
from conformer_ocr_b200.train import Trainer

world, rank, local = int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('RANK', '0')), int(os.environ.get('LOCAL_RANK', '0'))
dev = torch.device('cuda', local)
torch.cuda.set_device(dev)
os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
dist.init_process_group('nccl', device_id=dev)
hp = {k: HP[k] for k in ('height', 'encoder_dim', 'num_encoder_layers', 'num_attention_heads', 'feed_forward_expansion_factor',
                         'conv_expansion_factor', 'conv_kernel_size', 'half_step_residual', 'subsampling_conv_channels', 'subsampling_factor')}
C, W, B = 256, 1536, 32 // world
out = {}
for prec in ('fp32', 'bf16'):
    tr = Trainer(num_classes=C, device=dev, precision=prec, seed=1, **hp)
    tr.load_state_dict(S.make_weights(hp, C, seed=0))
    rng = np.random.RandomState(1000 + rank)
    img, lens = S.make_lines(B, hp['height'], W, seed=2000 + rank)
    tl = rng.randint(20, 81, size=B)
    batch = dict(image=torch.from_numpy(img), seq_lens=torch.from_numpy(lens), target=torch.from_numpy(rng.randint(1, C, size=int(tl.sum()))),
                 target_lens=torch.from_numpy(tl))
    tr.forward_backward(batch)
    own = tr.grads[:tr.n_trainable].clone()
    gathered = [torch.empty_like(own) for _ in range(world)]
    dist.all_gather(gathered, own)
    tr.allreduce_gradients()
    want = torch.stack(gathered).double().sum(0)
    err = float((tr.grads[:tr.n_trainable].double() - want).abs().max() / want.abs().max())
    tr.optimizer_step()
    for _ in range(2):
        tr.step(batch)
    p = tr.params[:tr.n_trainable].clone()
    ps = [torch.empty_like(p) for _ in range(world)]
    dist.all_gather(ps, p)
    spread = float(max((q - ps[0]).abs().max() for q in ps))
    out[prec] = dict(allreduce_rel_err=err, parameter_spread_after_3_steps=spread, finite=bool(torch.isfinite(p).all()))
    tr.close()
if rank == 0:
    print(json.dumps(dict(world=world, lines_per_rank=B, **out)), flush=True)
    assert all(v['allreduce_rel_err'] < 1e-6 and v['parameter_spread_after_3_steps'] == 0.0 and v['finite'] for v in out.values()), out
dist.barrier()
dist.destroy_process_group()
