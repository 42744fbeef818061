"""Greedy CTC decode / CTC loss restatements of the oracle (CPU only)."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from oracle import conformer_oracle as O


def _brute_force(outputs):
    """itertools.groupby formulation of the kraken greedy decoder contract."""
    from itertools import groupby
    labels = np.argmax(outputs, 0)
    res, pos = [], 0
    for lab, grp in groupby(labels):
        n = len(list(grp))
        if lab != 0:
            res.append((int(lab), pos, pos + n - 1, float(outputs[lab, pos:pos + n].max())))
        pos += n
    return res


def test_greedy_known_answer():
    #            t: 0    1    2    3    4    5    6
    lab = [0, 2, 2, 0, 2, 3, 3]
    C, T = 4, len(lab)
    o = np.full((C, T), -1.0, dtype=np.float32)
    for t, l in enumerate(lab):
        o[l, t] = 1.0 + 0.1 * t
    got = O.greedy_decoder(o)
    assert [(g[0], g[1], g[2]) for g in got] == [(2, 1, 2), (2, 4, 4), (3, 5, 6)]
    np.testing.assert_allclose([g[3] for g in got], [1.2, 1.4, 1.6], rtol=1e-6)


def test_greedy_edge_cases():
    assert O.greedy_decoder(np.zeros((5, 0), dtype=np.float32)) == []
    # all blank
    o = np.zeros((3, 4), dtype=np.float32)
    o[0] = 1.0
    assert O.greedy_decoder(o) == []
    # ties: lowest class index wins (numpy argmax) -> blank beats label 2 on a tie
    o = np.zeros((3, 2), dtype=np.float32)
    assert O.greedy_decoder(o) == []
    o[1, 1] = o[2, 1] = 5.0
    assert [(g[0], g[1], g[2]) for g in O.greedy_decoder(o)] == [(1, 1, 1)]


@settings(max_examples=60, deadline=None)
@given(st.integers(2, 9), st.integers(0, 40), st.integers(0, 2 ** 31 - 1))
def test_greedy_matches_groupby(C, T, seed):
    rng = np.random.RandomState(seed)
    # few distinct values -> plenty of runs and ties
    o = rng.randint(0, 3, size=(C, T)).astype(np.float32)
    assert O.greedy_decoder(o) == _brute_force(o)


def test_ctc_loss_matches_torch():
    rng = np.random.RandomState(0)
    T, B, C = 30, 4, 7
    logits = rng.normal(size=(T, B, C)).astype(np.float32)
    lp = O.log_softmax(logits)
    tl = [5, 0, 12, 20]          # the last one is infeasible for its short input -> zero_infinity
    il = [30, 10, 25, 15]
    targets = np.concatenate([rng.randint(1, C, size=n) for n in tl]).astype(np.int64)
    crit = torch.nn.CTCLoss(reduction='sum', zero_infinity=True)      # model.py:119
    ref = crit(torch.from_numpy(lp), torch.from_numpy(targets), torch.tensor(il), torch.tensor(tl)).item()
    got = O.ctc_loss_sum(lp, targets, il, tl)
    assert abs(got - ref) / abs(ref) < 1e-5
