"""Checkpoint formats of the reference (SURVEY.md 5 "Checkpoint / resume"), unchanged:

* Lightning checkpoint (`torch.save` dict): `state_dict` with keys `nn.encoder.*` / `nn.decoder.{weight,bias}`
  (ModuleDict `self.nn`, model.py:115), `hyper_parameters` (model.py:91 `save_hyperparameters()`), and
  `TextLineDataModule` = {'codec': c2l dict} (dataset.py:169-175).
* safetensors archive: a tar with `model.safetensors` (the same tensors WITHOUT the `nn.` prefix, pred.py:194 <-> :247),
  `metadata.json` {'codec', 'hyper_params', ...} (read by pred.py:185-192) and `_kraken_metadata.json`.

`checkpoint_to_safetensors` is the working form of pred.py:213-254.  The reference's version does not compile (positional
parameters after defaulted ones, pred.py:213-217), writes the user's metadata file under `_kraken_metadata.json` and tars a
`metadata.json` it never writes (pred.py:249-254).  Here `metadata.json` receives the record `load_safetensors` needs (the
reference's `kraken_metadata` dict, pred.py:230-237) and `_kraken_metadata.json` the optional user metadata, which is the only
reading under which the reference's own loader (pred.py:180-195) can open the result.  No Lightning / kraken import is needed:
the checkpoint is a plain dict.

`read_safetensors` / `load_archive_weights` are the fast path of SURVEY.md 8(f) rank 4: the safetensors container is parsed
with numpy (header JSON + one contiguous little-endian buffer) and every tensor is handed to `cocr_set_weight` as a view of
that buffer: no torch module, no random initialisation, no intermediate state-dict copy.
"""
from __future__ import annotations

import io
import json
import struct
import tarfile
from typing import TYPE_CHECKING, Dict, Mapping, Optional, Tuple

import numpy as np

if TYPE_CHECKING:
    from os import PathLike

_ST_DTYPES = {'F64': np.float64, 'F32': np.float32, 'F16': np.float16, 'I64': np.int64, 'I32': np.int32, 'I16': np.int16,
              'I8': np.int8, 'U8': np.uint8, 'BOOL': np.bool_}


def _bf16_to_f32(raw: np.ndarray) -> np.ndarray:
    return (raw.astype(np.uint32) << 16).view(np.float32)


def read_safetensors(buf: bytes) -> Dict[str, np.ndarray]:
    """Parse a safetensors blob into {name: array}.  Arrays are read-only views of `buf` (BF16 is widened to fp32)."""
    if len(buf) < 8:
        raise ValueError('not a safetensors file: too short')
    (n,) = struct.unpack('<Q', buf[:8])
    if n > len(buf) - 8:
        raise ValueError('not a safetensors file: header length exceeds the file')
    header = json.loads(buf[8:8 + n].decode('utf-8'))
    base = 8 + n
    out: Dict[str, np.ndarray] = {}
    for name, info in header.items():
        if name == '__metadata__':
            continue
        lo, hi = info['data_offsets']
        shape = tuple(info['shape'])
        dt = info['dtype']
        if dt == 'BF16':
            arr = _bf16_to_f32(np.frombuffer(buf, dtype='<u2', count=(hi - lo) // 2, offset=base + lo))
        elif dt in _ST_DTYPES:
            arr = np.frombuffer(buf, dtype=np.dtype(_ST_DTYPES[dt]).newbyteorder('<'), count=(hi - lo) // np.dtype(_ST_DTYPES[dt]).itemsize,
                                offset=base + lo)
        else:
            raise ValueError(f'unsupported safetensors dtype {dt} for {name}')
        out[name] = arr.reshape(shape)
    return out


def write_safetensors(tensors: Mapping[str, np.ndarray]) -> bytes:
    """Serialise {name: array} (fp32 / fp16 / integer arrays) as a safetensors blob; keys sorted like safetensors does."""
    rev = {np.dtype(v): k for k, v in _ST_DTYPES.items()}
    header, chunks, off = {}, [], 0
    for name in sorted(tensors):
        a = np.require(tensors[name], requirements='C')          # (ascontiguousarray would turn a 0-d tensor, e.g. num_batches_tracked, into shape (1,))
        if a.dtype not in rev:
            raise ValueError(f'unsupported dtype {a.dtype} for {name}')
        raw = a.astype(a.dtype.newbyteorder('<'), copy=False).tobytes()
        header[name] = {'dtype': rev[a.dtype], 'shape': list(a.shape), 'data_offsets': [off, off + len(raw)]}
        chunks.append(raw)
        off += len(raw)
    hj = json.dumps(header, separators=(',', ':')).encode('utf-8')
    hj += b' ' * ((8 - len(hj) % 8) % 8)
    return struct.pack('<Q', len(hj)) + hj + b''.join(chunks)


def read_archive(path: 'PathLike') -> Tuple[dict, Dict[str, np.ndarray]]:
    """tar{model.safetensors, metadata.json} -> (metadata, {state-dict key below `nn.`: array}); pred.py:184-193."""
    with tarfile.open(path, 'r') as tf:
        metadata = json.load(tf.extractfile('metadata.json'))
        if 'codec' not in metadata:
            raise ValueError('No codec in metadata record')
        if 'hyper_params' not in metadata:
            raise ValueError('No hyperparameters in metadata record')
        weights = read_safetensors(tf.extractfile('model.safetensors').read())
    return metadata, weights


def _plain(v):
    """hyper-parameters as JSON-serialisable values"""
    if isinstance(v, (bool, int, float, str)) or v is None:
        return v
    if isinstance(v, (list, tuple)):
        return [_plain(x) for x in v]
    if isinstance(v, Mapping):
        return {str(k): _plain(x) for k, x in v.items()}
    if hasattr(v, 'item'):
        return v.item()
    return str(v)


def checkpoint_to_safetensors(checkpoint_path: 'PathLike', output_path: 'PathLike', metadata: Optional['PathLike'] = None) -> None:
    """Lightning checkpoint of a RecognitionModel + TextLineDataModule -> safetensors archive for inference
    (pred.py:213-254; see the module docstring for the two defects of the reference's version that are not reproduced)."""
    import torch
    state = torch.load(checkpoint_path, map_location='cpu', weights_only=False)
    if 'TextLineDataModule' not in state:
        raise ValueError('Checkpoint does not contain data module state.')
    if 'hyper_parameters' not in state:
        raise ValueError('No hyperparameters in state_dict')
    codec = state['TextLineDataModule']['codec']
    tensors = {}
    for k, v in state['state_dict'].items():
        if not k.startswith('nn.'):
            continue                                            # metrics / criterion buffers of the LightningModule
        a = v.detach().cpu()
        tensors[k[len('nn.'):]] = (a.float() if a.is_floating_point() else a).numpy()
    kraken_metadata = {'codec': _plain(codec), 'seg_type': 'baselines', 'one_channel_mode': 'L', 'model_type': 'recognition',
                       'legacy_polygons': False, 'hyper_params': _plain(dict(state['hyper_parameters'])),
                       'kraken_min_version': 5.3}             # pred.py:230-237
    user = {}
    if metadata:
        with open(metadata, 'r') as fp:
            user = json.load(fp)
    with tarfile.open(output_path, 'w') as tar:
        for name, blob in (('model.safetensors', write_safetensors(tensors)),
                           ('_kraken_metadata.json', json.dumps(user).encode()),
                           ('metadata.json', json.dumps(kraken_metadata).encode())):
            ti = tarfile.TarInfo(name)
            ti.size = len(blob)
            tar.addfile(ti, io.BytesIO(blob))
