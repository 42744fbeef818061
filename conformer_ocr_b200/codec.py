"""Minimal label codec with the `decode` contract of kraken's `PytorchCodec` (used by
pred.py:163 / model.py:168) for installations without kraken: a one-to-one label -> string table
applied to (label, start, end, confidence) tuples.  Label 0 is the CTC blank and never decoded."""
from __future__ import annotations

from typing import Dict, List, Mapping, Sequence, Tuple, Union


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
