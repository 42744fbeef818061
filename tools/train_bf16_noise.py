"""GPU box: bf16 training-step gradient error against the float64 port at T = 384, for 1 / 4 / 16 layers, next to the error of
torch's own bf16 autocast on the same batch (the yardstick for mixed-precision noise)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from oracle import conformer_oracle as O
from oracle import torch_train_port as TP
from test_gpu_train import make_trainer, synthetic_batch, port_grads
dev = torch.device('cuda:0')
for L in (1, 4, 16):
    hp = dict(O.DEFAULT_HPARAMS, num_encoder_layers=L)
    C, B, W = 256, 3, 1536
    w = O.make_weights(hp, C, seed=31)
    batch = synthetic_batch(B, hp['height'], W, C, seed=32, widths=[1536, 1400, 1111])
    ref_loss, ref_g, _ = port_grads(batch, w, hp, dev)
    keys = list(ref_g)
    gmax = max(np.abs(ref_g[k]).max() for k in keys)
    def report(name, loss, grads):
        a = np.concatenate([np.asarray(grads[k], np.float64).ravel() for k in keys]); b = np.concatenate([ref_g[k].ravel() for k in keys])
        cos = a @ b / np.linalg.norm(a) / np.linalg.norm(b)
        errs = {k: np.abs(np.asarray(grads[k], np.float64) - ref_g[k]).max() / max(np.abs(ref_g[k]).max(), 0.05 * gmax) for k in keys}
        wk = max(errs, key=errs.get)
        print(f'L={L:2d} {name:14s} loss {loss:.3f} (ref {ref_loss:.3f}) cosine {cos:.6f} worst {errs[wk]:.3f} at {wk}', flush=True)
    for prec in ('fp32', 'bf16'):
        tr = make_trainer(hp, C, dev, precision=prec)
        tr.load_state_dict(w)
        loss = tr.forward_backward(batch).item()
        report('cuda ' + prec, loss, {k: tr.view(k, 'grads').cpu().numpy() for k in keys})
        tr.close()
    # torch autocast bf16 on the port
    wt = {k: torch.from_numpy(v.copy()).to(dev).requires_grad_(not k.endswith(('running_mean', 'running_var'))) for k, v in w.items()}
    with torch.autocast('cuda', dtype=torch.bfloat16):
        probits, olens, _ = TP.train_forward(batch['image'].to(dev), batch['seq_lens'], wt, hp)
    logp = torch.nn.functional.log_softmax(probits.float(), -1)
    loss = torch.nn.functional.ctc_loss(logp.transpose(0, 1), batch['target'].to(dev), olens, batch['target_lens'], blank=0, reduction='sum', zero_infinity=True)
    loss.backward()
    report('torch autocast', loss.item(), {k: wt[k].grad.float().cpu().numpy() for k in keys})
