"""Minimal label codec with the `decode` contract of kraken's `PytorchCodec` (used by
pred.py:163 / model.py:168) for installations without kraken: a one-to-one label -> string table
applied to (label, start, end, confidence) tuples.  Label 0 is the CTC blank and never decoded."""
from __future__ import annotations

from typing import Dict, List, Mapping, Optional, Sequence, Tuple, Union

import numpy as np


class LabelCodec:
    def __init__(self, charset: Union[Mapping[str, Sequence[int]], Sequence[str], str]):
        """charset: kraken-style c2l mapping {grapheme: [label]} (dataset.py:169-175) or a sequence of
        graphemes which are assigned labels 1..N in order."""
        if isinstance(charset, Mapping):
            self.c2l: Dict[str, List[int]] = {k: list(v) for k, v in charset.items()}
        else:
            self.c2l = {c: [i + 1] for i, c in enumerate(charset)}
        self.l2c: Dict[int, str] = {}
        for c, labs in self.c2l.items():
            if len(labs) != 1:
                raise ValueError('LabelCodec supports one label per grapheme; install kraken for n:m codecs')
            self.l2c[int(labs[0])] = c

    @property
    def max_label(self) -> int:
        return max(self.l2c) if self.l2c else 0

    def decode(self, labels: Sequence[Tuple[int, int, int, float]]) -> List[Tuple[str, int, int, float]]:
        return [(self.l2c[lab], s, e, c) for lab, s, e, c in labels if lab in self.l2c]

    __call__ = decode

    # -- batched string assembly -------------------------------------------
    # After the GPU collapse the reference still walks every line in Python: `codec.decode` builds one tuple per label
    # and `''.join` picks the graphemes (pred.py:163).  For a whole batch of collapsed label rows the same strings come
    # from one table look-up: labels -> UTF-32 code points (numpy fancy indexing over the (B, T) block), one
    # `bytes.decode` per line.  Labels the codec does not know are skipped, exactly like `decode`.
    def _codepoint_lut(self) -> Optional[np.ndarray]:
        lut = getattr(self, '_lut', None)
        if lut is None:
            if any(len(c) != 1 for c in self.l2c.values()):
                lut = False                       # multi-code-point graphemes: no fixed-width table
            else:
                lut = np.zeros(self.max_label + 1, dtype='<u4')
                for lab, c in self.l2c.items():
                    if lab >= 0:
                        lut[lab] = ord(c)
            self._lut = lut
        return None if lut is False else lut

    def decode_strings(self, labels: np.ndarray, counts: np.ndarray) -> List[str]:
        """labels: (B, T) collapsed label rows, counts: (B,) labels per row -> the B strings
        `''.join(g for g, *_ in self.decode(row))`."""
        labels = np.asarray(labels)
        counts = np.asarray(counts).astype(np.int64)
        if labels.ndim != 2 or counts.shape != (labels.shape[0],):
            raise ValueError(f'expected (B, T) labels and (B,) counts, got {labels.shape} and {counts.shape}')
        lut = self._codepoint_lut()
        if lut is None:
            return [''.join(self.l2c[int(lab)] for lab in row[:n] if int(lab) in self.l2c) for row, n in zip(labels, counts)]
        idx = np.clip(labels, 0, lut.shape[0] - 1)
        cps = np.where((labels >= 0) & (labels < lut.shape[0]), lut[idx], 0).astype('<u4')
        out = []
        for row, n in zip(cps, counts):
            r = row[:n]
            if not r.all():
                r = r[r != 0]                     # unknown labels (and a NUL grapheme, which kraken's codec never holds)
            out.append(r.tobytes().decode('utf-32-le'))
        return out
