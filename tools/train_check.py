"""GPU box: compare the CUDA training step with the golden vectors of the reference (tests/golden/train_*.npz) and with the
torch port's autograd (oracle/torch_train_port.py, float64 on the host), module by module.  usage: python tools/train_check.py [case ...]"""
import json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import conformer_oracle as O
from oracle import torch_train_port as TP
from conformer_ocr_b200.train import Trainer

idx = json.load(open(os.path.join(ROOT, 'tests/golden/index_train.json')))
PREC = os.environ.get('COCR_TRAIN_PRECISION', 'fp32')
cases = sys.argv[1:] or list(idx)
for name in cases:
    m = idx[name]
    g = np.load(os.path.join(ROOT, f'tests/golden/{name}.npz'))
    hp = m['hparams']
    w = O.make_weights(hp, m['num_classes'], seed=m['weight_seed'])
    img, lens = O.make_lines(m['batch'], hp['height'], m['width'], seed=m['input_seed'], widths=m['widths'])
    kw = {k: hp[k] for k in ('height', 'encoder_dim', 'num_encoder_layers', 'num_attention_heads', 'feed_forward_expansion_factor',
                             'conv_expansion_factor', 'conv_kernel_size', 'half_step_residual', 'subsampling_conv_channels',
                             'subsampling_factor')}
    tr = Trainer(num_classes=m['num_classes'], input_dropout_p=0.0, feed_forward_dropout_p=0.0, attention_dropout_p=0.0,
                 conv_dropout_p=0.0, lr=m['lr'], weight_decay=m['weight_decay'], precision=PREC, **kw)
    tr.load_state_dict(w)
    batch = {'image': torch.from_numpy(img), 'seq_lens': torch.from_numpy(lens), 'target': torch.from_numpy(g['targets']),
             'target_lens': torch.from_numpy(g['target_lens'])}
    loss, taps = tr.forward_backward(batch, debug_taps=True)
    torch.cuda.synchronize()
    print(f'== {name}: loss {loss.item():.6f} (reference {float(g["loss"]):.6f})')
    # module-level taps against the float64 port
    wt = {k: torch.from_numpy(v.copy()).double().requires_grad_(not k.endswith(('running_mean', 'running_var'))) for k, v in w.items()}
    tp = {}
    l64, _, _, stats = TP.train_step_loss(torch.from_numpy(img).double(), torch.from_numpy(lens), torch.from_numpy(g['targets']),
                                          torch.from_numpy(g['target_lens']), wt, hp, taps=tp)
    l64.backward()
    names = ['x0'] + [f'x{l}.{k}' for l in range(hp['num_encoder_layers']) for k in range(5)]
    tc = taps.cpu().double().numpy()
    for i, n in enumerate(names):
        ref, refg = tp[n].detach().numpy(), tp[n].grad.numpy()
        ef = np.abs(tc[0, i] - ref).max() / max(np.abs(ref).max(), 1e-30)
        eg = np.abs(tc[1, i] - refg).max() / max(np.abs(refg).max(), 1e-30)
        flag = ' <<<' if max(ef, eg) > (1e-4 if PREC == 'fp32' else 3e-2) else ''
        print(f'   tap {n:6s} fwd {ef:.2e}  grad {eg:.2e}{flag}')
    gmax = max(np.abs(g['grad/' + k]).max() for k in tr.trainable_keys())
    worst = 0
    for k in tr.trainable_keys():
        ref = g['grad/' + k]
        got = tr.view(k, 'grads').cpu().numpy()
        e = np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-6 * gmax)
        worst = max(worst, e)
        if e > (1e-4 if PREC == 'fp32' else 3e-2):
            print(f'   grad {k}: rel err {e:.3e} (|ref| {np.abs(ref).max():.3e}, |got| {np.abs(got).max():.3e})')
    print(f'   worst parameter-gradient rel err {worst:.3e}')
    se = max(np.abs(tr.view(k).cpu().numpy() - g['stat/' + k]).max() for k in tr.layout if k.endswith(('running_mean', 'running_var')))
    print(f'   running statistics max abs err {se:.3e}')
    tr.optimizer_step()
    torch.cuda.synchronize()
    pe = max(np.abs(tr.view(k).cpu().numpy() - g['adamw/' + k]).max() for k in tr.trainable_keys())
    print(f'   parameters after AdamW: max abs err {pe:.3e}')
    tr.close()
