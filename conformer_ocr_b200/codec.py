"""Label codec with the `decode` / `encode` contract of kraken's `PytorchCodec` (used by pred.py:163 / model.py:168;
state = the `c2l` dict of dataset.py:169-175) for installations without kraken.  kraken is a third-party dependency that is
neither vendored nor installable here (setup.cfg:36 `kraken>=4.3.13`), so this is a restatement of its documented
behaviour, PARITY UNPINNED: grapheme -> label-sequence table (one label per grapheme in the common case, n labels for
n:m codecs); `decode` walks the (label, start, end, confidence) records and, at every position, emits the grapheme of the
LONGEST label sequence of the table that matches there (start of its first record, end of its last, maximum confidence),
skipping labels no sequence starts with.  Label 0 is the CTC blank and never decoded."""
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
        self.l2c: Dict[int, str] = {}                      # single-label graphemes (the table the batched path uses)
        self.l2c_multi: Dict[Tuple[int, ...], str] = {}    # every label sequence, single ones included
        for c, labs in self.c2l.items():
            seq = tuple(int(x) for x in labs)
            if not seq or 0 in seq:
                raise ValueError(f'grapheme {c!r}: label sequences must be non-empty and must not contain the blank 0')
            self.l2c_multi[seq] = c
            if len(seq) == 1:
                self.l2c[seq[0]] = c
        self._max_seq = max((len(k) for k in self.l2c_multi), default=1)
        self.one_to_one = self._max_seq == 1

    @property
    def max_label(self) -> int:
        return max((max(k) for k in self.l2c_multi), default=0)

    def encode(self, s: str) -> List[int]:
        """Greedy longest-grapheme-first encoding of a string into labels (the training target of model.py:131);
        characters no grapheme starts with raise KeyError."""
        out: List[int] = []
        graphemes = sorted(self.c2l, key=len, reverse=True)
        i = 0
        while i < len(s):
            for g in graphemes:
                if g and s.startswith(g, i):
                    out.extend(self.c2l[g])
                    i += len(g)
                    break
            else:
                raise KeyError(f'no grapheme of the codec matches {s[i]!r}')
        return out

    def decode(self, labels: Sequence[Tuple[int, int, int, float]]) -> List[Tuple[str, int, int, float]]:
        if self.one_to_one:
            return [(self.l2c[lab], s, e, c) for lab, s, e, c in labels if lab in self.l2c]
        out: List[Tuple[str, int, int, float]] = []
        labs = [int(r[0]) for r in labels]
        i = 0
        while i < len(labs):
            for n in range(min(self._max_seq, len(labs) - i), 0, -1):
                g = self.l2c_multi.get(tuple(labs[i:i + n]))
                if g is not None:
                    out.append((g, labels[i][1], labels[i + n - 1][2], max(r[3] for r in labels[i:i + n])))
                    i += n
                    break
            else:
                i += 1                                          # undecodable label: skipped
        return out

    __call__ = decode

    # -- batched string assembly -------------------------------------------
    # After the GPU collapse the reference still walks every line in Python: `codec.decode` builds one tuple per label
    # and `''.join` picks the graphemes (pred.py:163).  For a whole batch of collapsed label rows the same strings come
    # from one table look-up: labels -> UTF-32 code points (numpy fancy indexing over the (B, T) block), one
    # `bytes.decode` per line.  Labels the codec does not know are skipped, exactly like `decode`.
    def _codepoint_lut(self) -> Optional[np.ndarray]:
        lut = getattr(self, '_lut', None)
        if lut is None:
            if not self.one_to_one or any(len(c) != 1 for c in self.l2c.values()):
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
            if self.one_to_one:
                return [''.join(self.l2c[int(lab)] for lab in row[:n] if int(lab) in self.l2c) for row, n in zip(labels, counts)]
            return [''.join(g for g, *_ in self.decode([(int(lab), 0, 0, 0.0) for lab in row[:n]])) for row, n in zip(labels, counts)]
        idx = np.clip(labels, 0, lut.shape[0] - 1)
        cps = np.where((labels >= 0) & (labels < lut.shape[0]), lut[idx], 0).astype('<u4')
        out = []
        for row, n in zip(cps, counts):
            r = row[:n]
            if not r.all():
                r = r[r != 0]                     # unknown labels (and a NUL grapheme, which kraken's codec never holds)
            out.append(r.tobytes().decode('utf-32-le', 'surrogatepass'))       # a codec may hold any code point a Python str can
        return out
