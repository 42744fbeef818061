"""Host-side mirror of the hot-path slice of `conformer_ocr.model.RecognitionModel`
(model.py:39-158): construction (`:99-116`), `forward(x, seq_lens)` (`:125-127`), `_step` (`:129-145`: encoder, decoder,
log_softmax, CTC loss) and `training_step` (`:153-158`) together with what Lightning does around it - backward pass, gradient
all-reduce, `optimizer_step` with the reference's warm-up (`:226-235`) - through `conformer_ocr_b200.train.Trainer`
(csrc/train.cu).  The Lightning loop itself, metrics, schedulers other than the warm-up and the data module are not built
(SURVEY.md 2 C9/C10)."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn as nn

from .conformer.encoder import ConformerEncoder, EngineOwner
from .engine import out_frames


class RecognitionModel(EngineOwner):
    def __init__(self, num_classes: int, height: int = 96, encoder_dim: int = 512, num_encoder_layers: int = 18,
                 num_attention_heads: int = 8, feed_forward_expansion_factor: int = 4, conv_expansion_factor: int = 2,
                 input_dropout_p: float = 0.1, feed_forward_dropout_p: float = 0.1, attention_dropout_p: float = 0.1,
                 conv_dropout_p: float = 0.1, conv_kernel_size: int = 9, half_step_residual: bool = True,
                 subsampling_conv_channels: int = 256, subsampling_factor: int = 4, precision: str = 'fp32',
                 mask_mode: str = 'reference_unmasked', **kwargs):
        """Signature defaults are the reference's (model.py:58-72: 512 / 18 / 8 / 9 / 256); the 144 / 16 / 4 / 31 / 32
        default spec comes from `default_specs.RECOGNITION_HYPER_PARAMS` at the call site, as in the reference CLI
        (cli/train.py:254-279).  Training-loop arguments (optimizer, schedule, ...) arrive in **kwargs and are ignored."""
        super().__init__()
        encoder = ConformerEncoder(in_channels=1, input_dim=height, encoder_dim=encoder_dim, num_layers=num_encoder_layers,
                                   num_attention_heads=num_attention_heads,
                                   feed_forward_expansion_factor=feed_forward_expansion_factor,
                                   conv_expansion_factor=conv_expansion_factor, input_dropout_p=input_dropout_p,
                                   feed_forward_dropout_p=feed_forward_dropout_p, attention_dropout_p=attention_dropout_p,
                                   conv_dropout_p=conv_dropout_p, conv_kernel_size=conv_kernel_size,
                                   half_step_residual=half_step_residual,
                                   subsampling_conv_channels=subsampling_conv_channels,
                                   subsampling_factor=subsampling_factor, precision=precision, mask_mode=mask_mode)
        decoder = nn.Linear(encoder_dim, num_classes, bias=True)
        for p in decoder.parameters():
            p.requires_grad_(False)
        self.nn = nn.ModuleDict({'encoder': encoder, 'decoder': decoder})    # model.py:115
        self.num_classes = num_classes
        # `self.save_hyperparameters()` (model.py:91): what a checkpoint stores under 'hyper_parameters'
        self.hyper_parameters = dict(num_classes=num_classes, height=height, encoder_dim=encoder_dim, num_encoder_layers=num_encoder_layers,
                                     num_attention_heads=num_attention_heads, feed_forward_expansion_factor=feed_forward_expansion_factor,
                                     conv_expansion_factor=conv_expansion_factor, input_dropout_p=input_dropout_p,
                                     feed_forward_dropout_p=feed_forward_dropout_p, attention_dropout_p=attention_dropout_p,
                                     conv_dropout_p=conv_dropout_p, conv_kernel_size=conv_kernel_size, half_step_residual=half_step_residual,
                                     subsampling_conv_channels=subsampling_conv_channels, subsampling_factor=subsampling_factor, **kwargs)
        self.codec = None                 # c2l dict of the data module (dataset.py:169-175), carried through checkpoints
        self._resume = None               # optimiser state of a checkpoint, handed to the trainer when it is created
        self._init_engine_state(precision, mask_mode)
        self.eval()

    def _engine_kwargs(self):
        return dict(num_classes=self.num_classes, **self.nn['encoder'].hparams)

    def _engine_state(self):
        return dict(self.nn.state_dict())

    def train(self, mode: bool = True):
        """train(True) selects the training step (`training_step`); leaving it copies the trained parameters and BatchNorm
        statistics back into this module's tensors, so `forward` / `predict*` see them (the inference engine re-uploads)."""
        if not mode and getattr(self, '_trainer', None) is not None and self._trainer_dirty:
            self.sync_from_trainer()
        return super().train(mode)

    # -- training (model.py:153-158 + Lightning's backward / DDP / optimizer_step) ---------------------------------------
    def configure_training(self, lr: float = 1e-3, weight_decay: float = 1e-3, warmup: int = 0, precision: str = 'bf16', seed: int = 0,
                           group=None, **kwargs):
        """The optimiser settings `_configure_optimizer_and_lr_scheduler` reads from hparams (model.py:255-274; AdamW only) and the
        arithmetic of the step: 'bf16' = tcgen05 operands with fp32 accumulation and master weights, 'fp32' = CUDA-core parity path."""
        self._train_cfg = dict(lr=lr, weight_decay=weight_decay, warmup=warmup, precision=precision, seed=seed, group=group)
        self._trainer = None

    def trainer_for(self, device):
        from .train import Trainer
        device = torch.device(device)
        if getattr(self, '_trainer', None) is None or self._trainer.device != torch.device('cuda', device.index if device.index is not None else torch.cuda.current_device()):
            cfg = dict(getattr(self, '_train_cfg', None) or dict(lr=1e-3, weight_decay=1e-3, warmup=0, precision='bf16', seed=0, group=None))
            hp = dict(self.nn['encoder'].hparams)
            dp = self.nn['encoder'].dropout_ps
            self._trainer = Trainer(num_classes=self.num_classes, height=hp['height'], encoder_dim=hp['encoder_dim'],
                                    num_encoder_layers=hp['num_layers'], num_attention_heads=hp['num_heads'],
                                    feed_forward_expansion_factor=hp['ff_expansion'], conv_expansion_factor=hp['conv_expansion'],
                                    conv_kernel_size=hp['conv_kernel_size'], half_step_residual=hp['half_step_residual'],
                                    subsampling_conv_channels=hp['subsampling_conv_channels'], subsampling_factor=hp['subsampling_factor'],
                                    device=device, **dp, **cfg)
            self._trainer.load_state_dict(self.nn.state_dict())
            self._trainer_dirty = False
            if self._resume is not None:                                     # resuming: AdamW moments and the step count of the checkpoint
                self._trainer.load_optimizer_state_dict(self._resume['optimizer'], [k for k, _ in self.nn.named_parameters()])
                self._trainer.global_step = int(self._resume['global_step'])
                self._resume = None
        return self._trainer

    def training_step(self, batch, batch_idx: int = 0):
        """model.py:153-158 plus the backward pass, the all-reduce of the gradients over the process group and the optimiser step.
        Returns the batch's summed CTC loss (a 1-element device tensor).  Out-of-memory errors skip the batch like `_step`."""
        if not self.training:
            raise RuntimeError('call .train() before training_step')
        try:
            tr = self.trainer_for(batch['image'].device if batch['image'].is_cuda else torch.device('cuda', torch.cuda.current_device()))
            loss = tr.step(batch)
            self._trainer_dirty = True
            return loss
        except RuntimeError as e:
            if 'out of memory' in str(e).lower():
                torch.cuda.empty_cache()
                return None
            raise

    def sync_from_trainer(self):
        """Copy parameters and BatchNorm running statistics of the trainer's flat buffer into this module (state-dict keys)."""
        sd = self._trainer.state_dict()
        own = self.nn.state_dict()
        with torch.no_grad():
            for k, v in sd.items():
                own[k].copy_(v.to(own[k].device))
        self._trainer_dirty = False
        self.invalidate_engine()

    # -- checkpoints (model.py:219-220 `save_checkpoint`; `load_from_checkpoint` as cli/train.py:328 and pred.py:229 call it) ----------
    def save_checkpoint(self, filename, codec=None):
        """Write a checkpoint with the layout of the reference's Lightning checkpoints: `state_dict` (keys `nn.*`, in the
        reference's order), `hyper_parameters`, `TextLineDataModule` = {'codec': c2l} (what `PytorchRecognitionModel.load_checkpoint`,
        pred.py:197-210, and the converter read), `global_step`, and - when a training step has run - `optimizer_states` in
        `torch.optim.AdamW.state_dict()` form, indexed in `nn.parameters()` order like Lightning's."""
        if getattr(self, '_trainer', None) is not None and self._trainer_dirty:
            self.sync_from_trainer()
        codec = codec if codec is not None else self.codec
        if hasattr(codec, 'c2l'):
            codec = codec.c2l
        ck = {'state_dict': {'nn.' + k: v.detach().cpu().clone() for k, v in self.nn.state_dict().items()},
              'hyper_parameters': dict(self.hyper_parameters), 'epoch': 0, 'global_step': 0}
        if codec is not None:
            ck['TextLineDataModule'] = {'codec': {k: list(v) for k, v in dict(codec).items()}}
        tr = getattr(self, '_trainer', None)
        if tr is not None:
            ck['global_step'] = int(tr.global_step)
            ck['optimizer_states'] = [tr.optimizer_state_dict([k for k, _ in self.nn.named_parameters()])]
        elif self._resume is not None:
            ck['global_step'] = int(self._resume['global_step'])
            ck['optimizer_states'] = [self._resume['optimizer']]
        torch.save(ck, filename)

    @classmethod
    def load_from_checkpoint(cls, checkpoint_path, map_location='cpu', **kwargs):
        """Lightning's classmethod as the reference uses it (cli/train.py:328-336, pred.py:229): the module is rebuilt from the
        checkpoint's `hyper_parameters` (keyword arguments override them), `nn.*` tensors are loaded strictly, metric buffers of the
        LightningModule are ignored; optimiser moments and the step count are kept for the next `training_step`."""
        ck = torch.load(checkpoint_path, map_location=map_location, weights_only=False)
        if 'hyper_parameters' not in ck:
            raise ValueError('No hyperparameters in state_dict')
        if 'state_dict' not in ck:
            raise ValueError('Checkpoint does not contain a state_dict.')
        hp = dict(ck['hyper_parameters'])
        hp.update(kwargs)
        model = cls(**hp)
        own = {k[len('nn.'):]: v for k, v in ck['state_dict'].items() if k.startswith('nn.')}
        model.nn.load_state_dict(own, strict=True)
        if 'TextLineDataModule' in ck:
            model.codec = ck['TextLineDataModule'].get('codec')
        if ck.get('optimizer_states'):
            model._resume = {'optimizer': ck['optimizer_states'][0], 'global_step': int(ck.get('global_step', 0))}
        return model

    def forward(self, x: torch.Tensor, seq_lens: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """x (N, W, H) as `_step` passes it (model.py:132-135) -> (logits (N,T,C), encoder_lens)."""
        eng = self.engine_for(x.device)
        logits, _ = eng.logits(x, seq_lens)
        lens = out_frames(seq_lens if seq_lens is not None else torch.full((x.shape[0],), x.shape[1]))
        return logits, lens

    def _step(self, batch, want_grad: bool = False):
        """model.py:129-151 in eval-mode arithmetic (dropout off, BatchNorm running statistics): encoder -> decoder ->
        log_softmax -> CTCLoss(reduction='sum', zero_infinity=True).  The log-softmax is fused into the loss kernels
        (only its per-frame normaliser is computed).  Returns the reference's dict {'loss', 'probits', 'output_lens'} plus
        'loss_per_line' and, when `want_grad`, 'grad_probits' = d loss / d probits (what autograd would hand to the
        decoder's backward).  Like the reference, an out-of-memory error skips the batch (returns None)."""
        try:
            input, target = batch['image'], batch['target']
            input = input.squeeze(1).transpose(1, 2)
            seq_lens, label_lens = batch['seq_lens'], batch['target_lens']
            eng = self.engine_for(input.device)
            probits, _ = eng.logits(input, seq_lens)
            encoder_lens = out_frames(seq_lens)
            loss, per_line, grad = eng.ctc_loss(probits, target, encoder_lens, label_lens, want_grad=want_grad)
            out = {'loss': loss, 'probits': probits, 'output_lens': encoder_lens, 'loss_per_line': per_line}
            if want_grad:
                out['grad_probits'] = grad
            return out
        except RuntimeError as e:
            if 'out of memory' in str(e).lower():                      # lightning's is_oom_error (model.py:146-149)
                torch.cuda.empty_cache()
                return None
            raise

    def decoder_training_step(self, batch, group=None):
        """Training step with a FROZEN encoder (eval-mode arithmetic): encoder forward, decoder projection, fused log-softmax +
        CTC loss (sum), its gradient with respect to the decoder outputs, the decoder's weight / bias gradients and the gradient
        flowing back into the encoder outputs; then ONE all-reduce (NCCL) of [loss | grad_weight | grad_bias] across the ranks of
        `group`.  Returns a dict with the local and the all-reduced values.  The full training step (encoder backward included,
        BASELINE configs[3]) is `training_step`."""
        from .sharding import allreduce_bucket
        input = batch['image'].squeeze(1).transpose(1, 2)
        seq_lens, label_lens = batch['seq_lens'], batch['target_lens']
        eng = self.engine_for(input.device)
        enc, _ = eng.encode(input, seq_lens)
        probits = eng.project(enc)
        encoder_lens = out_frames(seq_lens)
        loss, per_line, grad = eng.ctc_loss(probits, batch['target'], encoder_lens, label_lens, want_grad=True)
        gw, gb, ge = eng.decoder_backward(grad, enc)
        loss_all, gw_all, gb_all = allreduce_bucket([loss, gw, gb], group=group)
        return {'loss': loss, 'probits': probits, 'output_lens': encoder_lens, 'loss_per_line': per_line, 'grad_probits': grad,
                'grad_decoder_weight': gw, 'grad_decoder_bias': gb, 'grad_encoder_outputs': ge,
                'loss_allreduced': loss_all, 'grad_decoder_weight_allreduced': gw_all, 'grad_decoder_bias_allreduced': gb_all}

    def validation_logprobs(self, x: torch.Tensor, seq_lens: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """log_softmax(decoder(encoder(x)), dim=-1) (model.py:134-136) as a materialised tensor."""
        logits, lens = self.forward(x, seq_lens)
        return self.engine_for(x.device).log_softmax(logits), lens
