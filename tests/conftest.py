import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')
GOLDEN_CASES = ['tiny_equal', 'tiny_ragged', 'default_equal', 'default_ragged', 'default_2layer_long']


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def load_golden(name):
    """Return (meta dict, arrays dict) of one committed golden vector."""
    z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'))
    meta = json.loads(bytes(z['meta']).decode())
    return meta, {k: z[k] for k in z.files if k != 'meta'}


def golden_inputs(meta):
    """Re-generate the weights / inputs a golden vector was produced from."""
    from oracle import conformer_oracle as O
    hp = meta['hparams']
    w = O.make_weights(hp, meta['num_classes'], seed=meta['weight_seed'])
    img, lens = O.make_lines(meta['batch'], hp['height'], meta['width'], seed=meta['input_seed'],
                             widths=meta['widths'])
    return hp, w, img, lens


def rel_err(a, b):
    """max-norm relative error used by every parity test: max|a-b| / max|b|."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.fixture(scope='session')
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    return torch.device('cuda:0')
