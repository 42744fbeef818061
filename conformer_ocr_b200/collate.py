"""Input side of the path: what sits between the DataLoader's samples and `predict_string`.

* `collate_sequences` restates kraken's collate function (dataset.py:158,166 pass `kraken.lib.dataset.collate_sequences`
  to the DataLoader; kraken is absent here, so PARITY UNPINNED): samples `{'image': (1, H, W_i), 'target': labels}` are
  sorted by width, widest first, right-zero-padded to the widest and stacked to `(N, 1, H, W_max)`; `seq_lens` are the true
  widths (int64), `target` the 1-D concatenation of the label sequences, `target_lens` their lengths: exactly the
  batch layout `PytorchRecognitionModel.forward` / `RecognitionModel._step` consume (pred.py:119, model.py:131-133).
  With `uint8=True` the stack is an 8-bit tensor (raw pixels), which the bf16 engine converts inside its first kernel
  (cocr_forward_decode_u8): a quarter of the bytes cross PCIe; `pin=True` allocates the stack in pinned memory.
* `WidthBucketBatchSampler` yields index batches whose lines share one output frame count T (and differ by at most 3 px in
  width).  The reference never masks attention, so a line's output depends on how wide its batch is
  padded; equal-T batches bound that dependence to the last frame, keep the batch free of padding work, and let the
  engine replay one captured graph per (batch size, T) instead of re-planning on every batch.
"""
from __future__ import annotations

from typing import Dict, Iterator, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

from .sharding import out_frames_int


def collate_sequences(batch: Sequence[Dict], uint8: bool = False, pin: bool = False) -> Dict[str, object]:
    srt = sorted(batch, key=lambda x: x['image'].shape[2], reverse=True)
    seqs = [x['image'] for x in srt]
    seq_lens = torch.LongTensor([s.shape[2] for s in seqs])
    max_len = int(seqs[0].shape[2])
    first = seqs[0]
    dtype = torch.uint8 if uint8 else first.dtype
    out = torch.zeros((len(seqs),) + tuple(first.shape[:2]) + (max_len,), dtype=dtype, pin_memory=pin)
    for i, s in enumerate(seqs):
        if uint8 and s.dtype != torch.uint8:
            s = (s * 255.0).round().clamp_(0, 255).to(torch.uint8)        # ToTensor's inverse for 8-bit sources
        out[i, :, :, :s.shape[2]] = s
    res: Dict[str, object] = {'image': out, 'seq_lens': seq_lens}
    if 'target' in srt[0]:
        if isinstance(srt[0]['target'], str):
            res['target'] = [x['target'] for x in srt]
        else:
            res['target'] = torch.cat([torch.as_tensor(x['target']) for x in srt]).long()
        res['target_lens'] = torch.LongTensor([len(x['target']) for x in srt])
    return res


class WidthBucketBatchSampler:
    """`batch_sampler` for a DataLoader: batches of at most `batch_size` lines that share one output frame count.

    widths: the pixel width of every sample (after the dataset's own transforms).  Lines are sorted by width (widest
    first, like the collate function), cut into equal-T runs, and the batches are optionally shuffled as whole units."""

    def __init__(self, widths: Sequence[int], batch_size: int, shuffle: bool = False, seed: int = 0, drop_last: bool = False):
        if batch_size < 1:
            raise ValueError('batch_size must be positive')
        self.widths = [int(w) for w in widths]
        if any(w < 1 for w in self.widths):
            raise ValueError('widths must be positive')
        self.batch_size = batch_size
        self.shuffle = shuffle
        self.seed = seed
        self.drop_last = drop_last
        self.epoch = 0
        order = sorted(range(len(self.widths)), key=lambda i: (-self.widths[i], i))
        self._batches: List[List[int]] = []
        cur: List[int] = []
        cur_t = -1
        for i in order:
            t = out_frames_int(self.widths[i])
            if cur and (len(cur) == batch_size or t != cur_t):
                self._batches.append(cur)
                cur = []
            if not cur:
                cur_t = t
            cur.append(i)
        if cur:
            self._batches.append(cur)
        if drop_last:
            self._batches = [b for b in self._batches if len(b) == batch_size]

    def set_epoch(self, epoch: int) -> None:
        self.epoch = epoch

    def __len__(self) -> int:
        return len(self._batches)

    def __iter__(self) -> Iterator[List[int]]:
        idx = list(range(len(self._batches)))
        if self.shuffle:
            rng = np.random.RandomState(self.seed + self.epoch)
            rng.shuffle(idx)
        for i in idx:
            yield list(self._batches[i])
