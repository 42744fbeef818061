"""Host side of one data-parallel training step (BASELINE configs[3]).

`Trainer` owns, per rank, the flat fp32 parameter / gradient / AdamW-moment buffers (torch tensors) and a `cocr_handle`;
`step(batch)` is what `RecognitionModel.training_step` + Lightning's backward / DDP all-reduce / `optimizer.step` do in the
reference (model.py:129-158, 272-274, 226-235): forward in train() mode, CTC loss (sum), backward through every module in
hand-written CUDA kernels (csrc/train.cu), ONE NCCL all-reduce(sum) of the flat gradient buffer (the loss is a sum, model.py:119,
so gradients are summed, not averaged), AdamW with the reference's linear warm-up.  BatchNorm statistics are per rank (no
SyncBN), as under Lightning's default DDP strategy.  No autograd, no torch ops on the data path."""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Mapping, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from . import _lib
from .conformer.encoder import encoder_param_spec
from .engine import out_frames
from .sharding import allreduce_flat_prefix


def flat_to_adamw_state(layout: Mapping[str, Tuple[int, int, bool, Tuple[int, ...]]], exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor,
                        step: int, param_names: Sequence[str], lr: float, betas: Tuple[float, float], eps: float, weight_decay: float) -> Dict:
    """The flat AdamW moment buffers as `torch.optim.AdamW(...).state_dict()` (what Lightning stores under `optimizer_states`):
    state[i] belongs to the i-th name of `param_names` (= `nn.parameters()` order).  Before the first step the state is empty,
    like torch's."""
    state = {}
    if step > 0:
        for i, k in enumerate(param_names):
            off, num, trainable, shape = layout[k]
            if not trainable:
                raise ValueError(f'{k} is not a trainable parameter')
            state[i] = {'step': torch.tensor(float(step)), 'exp_avg': exp_avg[off:off + num].detach().cpu().clone().view(shape),
                        'exp_avg_sq': exp_avg_sq[off:off + num].detach().cpu().clone().view(shape)}
    group = {'lr': lr, 'betas': tuple(betas), 'eps': eps, 'weight_decay': weight_decay, 'amsgrad': False, 'maximize': False, 'foreach': None,
             'capturable': False, 'differentiable': False, 'fused': None, 'params': list(range(len(param_names)))}
    return {'state': state, 'param_groups': [group]}


def adamw_state_to_flat(layout: Mapping[str, Tuple[int, int, bool, Tuple[int, ...]]], opt_state: Mapping, param_names: Sequence[str],
                        exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor) -> int:
    """Inverse of `flat_to_adamw_state`: copies the moments of a `torch.optim.AdamW` state dict into the flat buffers (in place)
    and returns the step count (0 for an empty state).  Parameters are matched by position in `param_names` and checked by shape."""
    groups = opt_state.get('param_groups', [])
    ids = [i for g in groups for i in g['params']]
    if len(ids) != len(param_names):
        raise ValueError(f'optimizer state covers {len(ids)} parameters, the model has {len(param_names)}')
    steps = set()
    exp_avg.zero_()
    exp_avg_sq.zero_()
    for pid, k in zip(ids, param_names):
        st = opt_state['state'].get(pid)
        if st is None:
            continue
        off, num, trainable, shape = layout[k]
        if tuple(st['exp_avg'].shape) != tuple(shape):
            raise ValueError(f'optimizer state {pid} has shape {tuple(st["exp_avg"].shape)}, parameter {k} has {tuple(shape)}')
        exp_avg[off:off + num].copy_(st['exp_avg'].reshape(-1).to(torch.float32))
        exp_avg_sq[off:off + num].copy_(st['exp_avg_sq'].reshape(-1).to(torch.float32))
        steps.add(int(st['step']))
    if len(steps) > 1:
        raise ValueError(f'parameters disagree on the step count: {sorted(steps)}')
    return steps.pop() if steps else 0


class Trainer:
    def __init__(self, *, num_classes: int, height: int = 96, encoder_dim: int = 144, num_encoder_layers: int = 16,
                 num_attention_heads: int = 4, feed_forward_expansion_factor: int = 4, conv_expansion_factor: int = 2,
                 input_dropout_p: float = 0.1, feed_forward_dropout_p: float = 0.1, attention_dropout_p: float = 0.1,
                 conv_dropout_p: float = 0.1, conv_kernel_size: int = 31, half_step_residual: bool = True,
                 subsampling_conv_channels: int = 32, subsampling_factor: int = 4, device='cuda', lr: float = 1e-3,
                 weight_decay: float = 1e-3, betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8, warmup: int = 0,
                 seed: int = 0, group: Optional[dist.ProcessGroup] = None, precision: str = 'fp32', **kwargs):
        """Hyper-parameter names are RecognitionModel's (model.py:58-72); the remaining training-loop arguments arrive in
        **kwargs and are ignored."""
        self._lib = _lib.load()
        device = torch.device(device)
        if device.type != 'cuda' or not torch.cuda.is_available():
            raise RuntimeError('conformer_ocr_b200 trains on a CUDA device only (no CPU fallback)')
        self.device = torch.device('cuda', device.index if device.index is not None else torch.cuda.current_device())
        self.hp = _lib.cocr_hparams(height, encoder_dim, num_encoder_layers, num_attention_heads, feed_forward_expansion_factor,
                                    conv_expansion_factor, conv_kernel_size, int(bool(half_step_residual)),
                                    subsampling_conv_channels, subsampling_factor, num_classes, _lib.COCR_FP32,
                                    _lib.COCR_MASK_REFERENCE_UNMASKED)
        self._h = C.c_void_p()
        _lib.raise_for_status(self._lib.cocr_create(C.byref(self.hp), self.device.index, C.byref(self._h)), None)
        if precision not in ('fp32', 'bf16'):
            raise ValueError(f"unknown precision {precision!r}; expected 'fp32' (CUDA-core parity path) or 'bf16' (tcgen05, fp32 master)")
        self.precision = precision
        self.cfg = _lib.cocr_train_config(input_dropout_p, feed_forward_dropout_p, attention_dropout_p, conv_dropout_p, 0.1,
                                          1 if precision == 'bf16' else 0, seed)
        self.seed = int(seed)
        self.num_classes, self.encoder_dim, self.num_layers = num_classes, encoder_dim, num_encoder_layers
        self.lr, self.weight_decay, self.betas, self.eps, self.warmup = lr, weight_decay, betas, eps, warmup
        self.group = group
        self.global_step = 0
        # layout
        nt, ntr, ne = C.c_int64(), C.c_int64(), C.c_int32()
        self._check(self._lib.cocr_train_layout(self._h, C.byref(nt), C.byref(ntr), C.byref(ne)))
        self.n_total, self.n_trainable = nt.value, ntr.value
        shapes = {'encoder.' + k: s for k, s, _ in encoder_param_spec(height, encoder_dim, num_encoder_layers, num_attention_heads,
                                                                        feed_forward_expansion_factor, conv_kernel_size,
                                                                        subsampling_conv_channels)}
        shapes['decoder.weight'] = (num_classes, encoder_dim)
        shapes['decoder.bias'] = (num_classes,)
        self.layout: Dict[str, Tuple[int, int, bool, Tuple[int, ...]]] = {}
        buf = C.create_string_buffer(256)
        for i in range(ne.value):
            off, num, tr = C.c_int64(), C.c_int64(), C.c_int32()
            self._check(self._lib.cocr_train_entry(self._h, i, buf, 256, C.byref(off), C.byref(num), C.byref(tr)))
            key = buf.value.decode()
            self.layout[key] = (off.value, num.value, bool(tr.value), tuple(shapes[key]))
        self.params = torch.zeros(self.n_total, dtype=torch.float32, device=self.device)
        self.grads = torch.zeros(self.n_total, dtype=torch.float32, device=self.device)
        self.exp_avg = torch.zeros(self.n_trainable, dtype=torch.float32, device=self.device)
        self.exp_avg_sq = torch.zeros(self.n_trainable, dtype=torch.float32, device=self.device)
        self._loss = torch.zeros(1, dtype=torch.float32, device=self.device)
        self._stage: Dict[str, torch.Tensor] = {}

    # -- life cycle -------------------------------------------------------------------------
    def close(self):
        if getattr(self, '_h', None) is not None and self._h.value:
            self._lib.cocr_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        _lib.raise_for_status(rc, self._h)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # -- parameters -------------------------------------------------------------------------
    def view(self, key: str, which: str = 'params') -> torch.Tensor:
        """The tensor of state-dict key `key` inside the flat buffer `which` ('params' | 'grads' | 'exp_avg' | 'exp_avg_sq')."""
        off, num, _, shape = self.layout[key]
        return getattr(self, which)[off:off + num].view(shape)

    def load_state_dict(self, state: Mapping[str, object]):
        """Keys as in `RecognitionModel.nn.state_dict()`; `num_batches_tracked` is ignored (model.py:115)."""
        missing = [k for k in self.layout if k not in state]
        if missing:
            raise ValueError(f'missing key in state_dict: {missing[0]}')
        for k, v in state.items():
            if k.endswith('num_batches_tracked'):
                continue
            if k not in self.layout:
                raise ValueError(f'unexpected key in state_dict: {k}')
            t = torch.as_tensor(v, dtype=torch.float32)
            if tuple(t.shape) != self.layout[k][3]:
                raise ValueError(f'size mismatch for {k}: got {tuple(t.shape)} expected {self.layout[k][3]}')
            self.view(k).copy_(t.to(self.device))

    def state_dict(self) -> Dict[str, torch.Tensor]:
        return {k: self.view(k).detach().clone() for k in self.layout}

    def trainable_keys(self) -> List[str]:
        return [k for k, v in self.layout.items() if v[2]]

    # -- one step ---------------------------------------------------------------------------
    def forward_backward(self, batch: Mapping[str, torch.Tensor], debug_taps: bool = False):
        """`batch` as the reference's collate function hands it to `_step` (model.py:131-133): 'image' (N,1,H,W) float,
        'seq_lens' (N), 'target' 1-D labels, 'target_lens' (N).  Leaves d loss / d param in `self.grads`; returns the loss
        tensor (1 element, on the device) [and the debug taps]."""
        img = batch['image']
        if img.dim() != 4 or img.shape[1] != 1:
            raise ValueError(f'expected an (N,1,H,W) image batch, got {tuple(img.shape)}')
        if img.shape[2] != self.hp.height:
            raise ValueError(f'line height {img.shape[2]} does not match the model ({self.hp.height})')
        # Persistent device staging: the library replays the step's launch sequence as a CUDA graph as long as the buffer
        # addresses repeat, so host batches are copied into buffers that live as long as the trainer.
        if not (img.is_cuda and img.dtype == torch.float32 and img.device == self.device):
            img = self._staged('image', img, torch.float32)
        x = img.squeeze(1).transpose(1, 2)                                  # model.py:132: (N, W, H) view
        B, W, H = x.shape
        widths = self._staged('seq_lens', torch.as_tensor(batch['seq_lens']), torch.int32)
        tl_host = torch.as_tensor(batch['target_lens']).to('cpu', dtype=torch.int64)
        tlens = self._staged('target_lens', tl_host, torch.int32)
        tg = torch.as_tensor(batch['target'])
        if tl_host.numel() != B:
            raise ValueError(f'{tl_host.numel()} target lengths for a batch of {B}')
        if tg.numel() != int(tl_host.sum()):
            raise ValueError('target has to be the 1-D concatenation of the label sequences')
        if not tg.is_cuda and tg.numel() and (int(tg.min()) < 1 or int(tg.max()) >= self.num_classes):
            # the loss kernels index the logits with the labels; a device-resident target is taken on trust (no host sync per step)
            raise ValueError('target labels must lie in 1 .. num_classes-1 (0 is the CTC blank)')
        targets = self._staged('target', tg, torch.int32, capacity=max(int(tg.numel()), B * 128))
        taps = None
        if debug_taps:
            T = int(out_frames(torch.tensor([W]))[0])
            taps = torch.zeros(2, 1 + 5 * self.num_layers, B, T, self.encoder_dim, device=self.device)
        self.cfg.seed = (self.seed * 1000003 + self.global_step) & 0xFFFFFFFFFFFFFFFF
        sb, sw, sh = x.stride()
        self._check(self._lib.cocr_train_step(self._h, self.params.data_ptr(), self.grads.data_ptr(), x.data_ptr(), sb, sw, sh,
                                              widths.data_ptr(), B, W, targets.data_ptr(), tlens.data_ptr(),
                                              int(tl_host.max()) if B else 0, C.byref(self.cfg), self._stream(),
                                              self._loss.data_ptr(), None, taps.data_ptr() if taps is not None else None))
        return (self._loss, taps) if debug_taps else self._loss

    def _staged(self, name: str, src: torch.Tensor, dtype, capacity: Optional[int] = None) -> torch.Tensor:
        """Copy `src` into a persistent device buffer (re-allocated only when it has to grow) and return the view of it."""
        n = src.numel()
        buf = self._stage.get(name)
        if buf is None or buf.numel() < n or buf.dtype != dtype:
            buf = torch.empty(max(n, capacity or n), dtype=dtype, device=self.device)
            self._stage[name] = buf
        view = buf[:n].view(src.shape)
        view.copy_(src, non_blocking=True)
        return view

    def allreduce_gradients(self):
        """ONE all-reduce(sum) of the trainable prefix of the flat gradient buffer (NCCL over NVLink).  The loss stays the
        rank's own sum, as the value `training_step` logs per rank under Lightning's DDP (model.py:155-157)."""
        allreduce_flat_prefix(self.grads, self.n_trainable, self.group)

    def optimizer_step(self):
        """torch.optim.AdamW(lr, weight_decay) (model.py:272-274) with the linear warm-up of model.py:230-235."""
        self.global_step += 1
        lr = self.lr
        if self.warmup and self.global_step <= self.warmup:
            lr = min(1.0, float(self.global_step) / self.warmup) * self.lr
        self._check(self._lib.cocr_adamw_step(self._h, self.params.data_ptr(), self.grads.data_ptr(), self.exp_avg.data_ptr(),
                                              self.exp_avg_sq.data_ptr(), self.n_trainable, lr, self.betas[0], self.betas[1], self.eps,
                                              self.weight_decay, self.global_step, self._stream()))

    def optimizer_state_dict(self, param_names: Optional[Sequence[str]] = None) -> Dict:
        """`torch.optim.AdamW.state_dict()` form of the flat moment buffers (for checkpoints); `param_names` = `nn.parameters()`
        order of the module the checkpoint belongs to (default: the trainable keys in layout order)."""
        return flat_to_adamw_state(self.layout, self.exp_avg, self.exp_avg_sq, self.global_step, list(param_names or self.trainable_keys()),
                                   self.lr, self.betas, self.eps, self.weight_decay)

    def load_optimizer_state_dict(self, opt_state: Mapping, param_names: Optional[Sequence[str]] = None) -> None:
        """Resume: moments and step count of a `torch.optim.AdamW` state dict (Lightning's `optimizer_states[0]`)."""
        self.global_step = adamw_state_to_flat(self.layout, opt_state, list(param_names or self.trainable_keys()), self.exp_avg, self.exp_avg_sq)

    def step(self, batch: Mapping[str, torch.Tensor]) -> torch.Tensor:
        loss = self.forward_backward(batch)
        self.allreduce_gradients()
        self.optimizer_step()
        return loss

    def dropout_mask(self, site: int, p: float, n: int, seed: Optional[int] = None) -> torch.Tensor:
        out = torch.empty(n, dtype=torch.float32, device=self.device)
        self._check(self._lib.cocr_dropout_mask(self._h, self.cfg.seed if seed is None else seed, site, p, n, self._stream(), out.data_ptr()))
        return out

    @property
    def launch_count(self) -> int:
        return int(self._lib.cocr_launch_count(self._h))

    def workspace_bytes(self, B: int, W: int) -> int:
        n = C.c_size_t()
        self._check(self._lib.cocr_train_workspace_bytes(self._h, B, W, C.byref(n)))
        return n.value
