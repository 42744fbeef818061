"""TEST INFRASTRUCTURE (checker only; never imported by conformer_ocr_b200/).

Recipe for `oracle/_ref/`: the reference's own encoder modules, compiled where they lie.

The reference is a Python package whose build backend (`pbr`, pyproject.toml) is absent from this image, so
`pip install --target ... /root/reference` fails (`ModuleNotFoundError: No module named 'pbr'`; recorded in DESIGN.md 5).
Its `conformer_ocr.conformer` sub-package - encoder.py, convolution.py, attention.py, embedding.py, feed_forward.py,
modules.py: every module of the hot path, SURVEY.md 8a rows A1-A9 - depends on torch alone.  This recipe compiles those
files FROM /root/reference (CPython's compiler, like gcc on a C reference) into `oracle/_ref/<dotted module name>.mcode`
(interpreter magic + marshalled code object): compiled outputs only, no reference source enters the repo.  `oracle/_ref/` is
git-ignored and travels to the GPU box with the snapshot, where the same interpreter executes the code objects through the
small finder in `oracle/ref_modules.py`.  (Plain sourceless `.pyc` files were tried first: the snapshot leaves `*.pyc` behind.)

`conformer_ocr/model.py` and `pred.py` are NOT compiled: they import lightning / kraken (absent) and pred.py has a
SyntaxError (pred.py:213-217); the ten wrapper lines of `pred.py:118-122,172-178` and `model.py:129-145` that drive the
modules are restated in `oracle/ref_modules.py`, each citing its line.

    python oracle/build_ref.py            # needs /root/reference; a no-op message without it
"""
from __future__ import annotations

import hashlib
import importlib.util
import json
import os
import marshal
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = '/root/reference'
OUT = os.path.join(HERE, '_ref')
FILES = ['conformer_ocr/__init__.py', 'conformer_ocr/conformer/__init__.py', 'conformer_ocr/conformer/encoder.py',
         'conformer_ocr/conformer/convolution.py', 'conformer_ocr/conformer/attention.py', 'conformer_ocr/conformer/embedding.py',
         'conformer_ocr/conformer/feed_forward.py', 'conformer_ocr/conformer/modules.py']


def build(ref_root: str = REF_ROOT, out: str = OUT, quiet: bool = False) -> bool:
    """Returns True when oracle/_ref was (re)built, False when the reference is not on this machine."""
    if not os.path.isdir(os.path.join(ref_root, 'conformer_ocr', 'conformer')):
        if not quiet:
            print(f'oracle/_ref: {ref_root} is absent; keeping what is there')
        return False
    manifest = {'python': sys.version.split()[0], 'magic': importlib.util.MAGIC_NUMBER.hex(), 'files': {}}
    os.makedirs(out, exist_ok=True)
    for rel in FILES:
        src = os.path.join(ref_root, rel)
        name = rel[:-3].replace('/', '.')
        if name.endswith('.__init__'):
            name = name[:-len('.__init__')]
        with open(src, 'rb') as fp:
            text = fp.read()
        code = compile(text, src, 'exec', dont_inherit=True)       # `src` is only the file name tracebacks show
        with open(os.path.join(out, name + '.mcode'), 'wb') as fp:
            fp.write(importlib.util.MAGIC_NUMBER)
            fp.write(marshal.dumps(code))
        manifest['files'][rel] = hashlib.sha256(text).hexdigest()
    with open(os.path.join(out, 'MANIFEST.json'), 'w') as fp:
        json.dump(manifest, fp, indent=1, sort_keys=True)
    if not quiet:
        print(f'oracle/_ref: compiled {len(FILES)} reference modules from {ref_root}')
    return True


if __name__ == '__main__':
    build()
