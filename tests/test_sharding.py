"""Host-side sharding logic: batch planning, the wire format, and the N>1 gather on a
world-size-2 gloo group (CPU)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conformer_ocr_b200 import sharding as S


def fake_records(idx):
    rng = np.random.default_rng(1000 + idx)
    n = int(rng.integers(0, 9))
    starts = np.sort(rng.choice(200, size=n, replace=False))
    return [(int(rng.integers(1, 255)), int(s), int(s + rng.integers(0, 3)), float(np.float32(rng.normal()))) for s in starts]


def test_out_frames_matches_reference_values():
    # SURVEY.md §8 notation block (probed on the reference's calc_length)
    for w, t in ((1024, 256), (1025, 257), (1536, 384), (2048, 512), (4096, 1024), (100, 25), (5, 2), (1, 1), (4, 1)):
        assert S.out_frames_int(w) == t


@pytest.mark.parametrize('world', [1, 2, 8])
def test_plan_covers_every_line_once_and_balances(world):
    rng = np.random.default_rng(3)
    widths = rng.integers(900, 2100, size=1000).tolist()
    shards = S.plan_shards(widths, world, batch_lines=32)
    seen = sorted(i for sh in shards for b in sh for i in b.lines)
    assert seen == list(range(1000))
    for sh in shards:
        for b in sh:
            assert len(b.lines) <= 32
            assert len({S.out_frames_int(widths[i]) for i in b.lines}) == 1          # equal-T batches
            assert b.width == max(widths[i] for i in b.lines) and b.frames == S.out_frames_int(b.width)
    loads = [sum(b.cost for b in sh) for sh in shards]
    assert max(loads) <= 1.15 * (sum(loads) / world)


def test_plan_rejects_bad_input():
    with pytest.raises(ValueError):
        S.plan_shards([10, 0, 5], 2, 4)
    with pytest.raises(ValueError):
        S.plan_shards([10], 0, 4)
    assert S.plan_shards([], 2, 4) == [[], []]


def test_pack_unpack_round_trip():
    recs = {i: fake_records(i) for i in (5, 0, 17, 3)}
    recs[9] = []
    assert S.unpack_records(S.pack_records(recs)) == recs
    assert S.unpack_records(S.pack_records({})) == {}


def _worker(rank, world, port, widths, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        shards = S.plan_shards(widths, world, batch_lines=4)
        local = {i: fake_records(i) for b in shards[rank] for i in b.lines}
        full = S.gather_records(local)
        q.put((rank, full))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_gather_records_world_size_2_gloo():
    widths = [1024, 1500, 1500, 990, 2048, 1024, 1024, 700, 1300]
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, widths, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = {i: fake_records(i) for i in range(len(widths))}
    assert got[0] == want and got[1] == want


def _ar_worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(100 + rank)
        ts = [torch.rand((), generator=g), torch.rand((5, 3), generator=g), torch.rand((7,), generator=g).to(torch.float64)]
        out = S.allreduce_bucket(ts)
        q.put((rank, [t.clone() for t in ts], out))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_allreduce_bucket_world_size_2_gloo():
    """training exchange step: loss + gradients summed over ranks with one flat all-reduce (sum, not mean: model.py:119)"""
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_ar_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = {r: (a, b) for r, a, b in (q.get(timeout=120) for _ in range(2))}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for i in range(3):
        want = got[0][0][i].double() + got[1][0][i].double()
        for r in (0, 1):
            assert got[r][1][i].shape == got[r][0][i].shape and got[r][1][i].dtype == got[r][0][i].dtype
            torch.testing.assert_close(got[r][1][i].double(), want, rtol=1e-6, atol=1e-6)
    solo = S.allreduce_bucket([torch.ones(3)])                 # no process group: identity
    assert torch.equal(solo[0], torch.ones(3)) and S.allreduce_bucket([]) == []


def _flat_worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(200 + rank)
        flat = torch.rand(1000, generator=g)
        before = flat.clone()
        S.allreduce_flat_prefix(flat, 900)
        q.put((rank, before, flat.clone()))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_allreduce_flat_prefix_world_size_2_gloo():
    """the Trainer's exchange step: the trainable prefix of the flat gradient buffer is summed over the ranks in place, the
    per-rank tail (BatchNorm running statistics) is not touched"""
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_flat_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = {r: (a, b) for r, a, b in (q.get(timeout=120) for _ in range(2))}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = got[0][0][:900] + got[1][0][:900]
    for r in (0, 1):
        torch.testing.assert_close(got[r][1][:900], want, rtol=1e-6, atol=1e-6)
        assert torch.equal(got[r][1][900:], got[r][0][900:])
    import pytest
    with pytest.raises(ValueError):
        S.allreduce_flat_prefix(torch.zeros(4), 5)
    assert torch.equal(S.allreduce_flat_prefix(torch.ones(4), 4), torch.ones(4))        # no process group: identity
