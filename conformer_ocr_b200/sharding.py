"""Data-parallel sharding of text lines over the GPUs of one box.

Lines are independent in eval mode (LayerNorm per frame, BatchNorm with running statistics, attention per
line), so the path shards with no data-path collective: every rank runs the whole model on its own lines.
The one exchange step is the gather of the decoded records.  This module is host-side plumbing only:

* `plan_shards`   — bucket lines into equal-frame-count batches and deal the batches to ranks by cost.
  Equal-width batches also make the reference's unmasked attention independent of batch
  composition, so sharded results equal single-GPU results.
* `pack_records` / `unpack_records` — the wire format: one int32 vector per rank.
* `gather_records` — size exchange + one padded all-gather (`torch.distributed`; NCCL on GPUs, gloo in tests).
* `allreduce_bucket` — the training configuration's exchange step (BASELINE configs[3]): the loss and the gradient tensors a
  rank produced are flattened into ONE buffer and summed across ranks with one all-reduce (the loss is `reduction='sum'`,
  model.py:119, so gradients are summed, not averaged).  With 8.2 M parameters (33 MB fp32) the wire time on NVLink 5 is
  well under 100 us, so a single bucket is latency-optimal; call it once per group of layers to overlap with a backward pass.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

Record = Tuple[int, int, int, float]          # (label, start frame, end frame inclusive, confidence)


def out_frames_int(width: int) -> int:
    """calc_length (conformer/convolution.py:240-247) for subsampling factor 4, integer form."""
    t2 = (width - 1) // 2 + 1
    return (t2 - 1) // 2 + 1


@dataclass
class Batch:
    lines: List[int]          # indices into the caller's line list
    width: int                # padded width of the batch (max of its lines)
    frames: int               # output frames of the padded width

    @property
    def cost(self) -> float:
        # per line: linear part (projections, FFN, conv) + quadratic part (attention), SURVEY.md §8d
        t = float(self.frames)
        return len(self.lines) * (t * 1.0 + t * t / 840.0)


def plan_shards(widths: Sequence[int], world_size: int, batch_lines: int, equal_frames: bool = True) -> List[List[Batch]]:
    """Returns, per rank, the batches it processes.  Lines are sorted by width (descending, like kraken's
    collate order), cut into batches of at most `batch_lines` lines that share one output frame count when
    `equal_frames`, and the batches are dealt longest-processing-time-first to the least loaded rank."""
    if world_size < 1 or batch_lines < 1:
        raise ValueError('world_size and batch_lines must be positive')
    order = sorted(range(len(widths)), key=lambda i: (-int(widths[i]), i))
    batches: List[Batch] = []
    cur: List[int] = []
    cur_frames = -1
    for i in order:
        w = int(widths[i])
        if w < 1:
            raise ValueError(f'line {i} has non-positive width {w}')
        t = out_frames_int(w)
        if cur and (len(cur) == batch_lines or (equal_frames and t != cur_frames)):
            batches.append(Batch(cur, int(widths[cur[0]]), out_frames_int(int(widths[cur[0]]))))
            cur = []
        if not cur:
            cur_frames = t
        cur.append(i)
    if cur:
        batches.append(Batch(cur, int(widths[cur[0]]), out_frames_int(int(widths[cur[0]]))))
    shards: List[List[Batch]] = [[] for _ in range(world_size)]
    load = [0.0] * world_size
    for b in sorted(batches, key=lambda b: -b.cost):
        r = min(range(world_size), key=lambda k: (load[k], k))
        shards[r].append(b)
        load[r] += b.cost
    return shards


def pack_records(records: Dict[int, List[Record]]) -> torch.Tensor:
    """{line index: records} -> int32 vector [n_lines, (line, count, labels.., starts.., ends.., conf bits..)*]."""
    out: List[np.ndarray] = [np.asarray([len(records)], np.int32)]
    for idx in sorted(records):
        rec = records[idx]
        n = len(rec)
        out.append(np.asarray([idx, n], np.int32))
        if n:
            lab, st, en, cf = zip(*rec)
            out.append(np.asarray(lab, np.int32))
            out.append(np.asarray(st, np.int32))
            out.append(np.asarray(en, np.int32))
            out.append(np.asarray(cf, np.float32).view(np.int32))
    return torch.from_numpy(np.concatenate(out))


def unpack_records(buf: torch.Tensor) -> Dict[int, List[Record]]:
    a = buf.detach().cpu().numpy().astype(np.int32, copy=False)
    n_lines = int(a[0])
    pos = 1
    out: Dict[int, List[Record]] = {}
    for _ in range(n_lines):
        idx, n = int(a[pos]), int(a[pos + 1])
        pos += 2
        lab = a[pos:pos + n]; st = a[pos + n:pos + 2 * n]; en = a[pos + 2 * n:pos + 3 * n]
        cf = a[pos + 3 * n:pos + 4 * n].view(np.float32)
        pos += 4 * n
        out[idx] = list(zip(lab.tolist(), st.tolist(), en.tolist(), cf.tolist()))
    return out


def gather_records(local: Dict[int, List[Record]], device: Optional[torch.device] = None,
                   group: Optional[dist.ProcessGroup] = None) -> Dict[int, List[Record]]:
    """All ranks contribute the records of their own lines and receive everybody's.  Two collectives: the
    packed sizes, then one all-gather of buffers padded to the largest."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return dict(local)
    world = dist.get_world_size(group)
    dev = torch.device(device) if device is not None else torch.device('cpu')
    mine = pack_records(local).to(dev)
    sizes = torch.zeros(world, dtype=torch.int64, device=dev)
    sizes[dist.get_rank(group)] = mine.numel()
    dist.all_reduce(sizes, op=dist.ReduceOp.SUM, group=group)
    cap = int(sizes.max().item())
    send = torch.zeros(cap, dtype=torch.int32, device=dev)
    send[:mine.numel()] = mine
    recv = torch.empty(world * cap, dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(recv, send, group=group)
    out: Dict[int, List[Record]] = {}
    for r in range(world):
        part = unpack_records(recv[r * cap:r * cap + int(sizes[r].item())])
        dup = set(part) & set(out)
        if dup:
            raise RuntimeError(f'lines {sorted(dup)[:5]} were decoded by more than one rank')
        out.update(part)
    return out


def allreduce_bucket(tensors: Sequence[torch.Tensor], group: Optional[dist.ProcessGroup] = None) -> List[torch.Tensor]:
    """Sum `tensors` (same device, floating point) over all ranks with ONE all-reduce of a flat fp32 bucket; returns new
    tensors with the shapes / dtypes of the inputs.  A no-op copy when no process group is initialised or world size is 1."""
    ts = [t.detach() for t in tensors]
    if not ts:
        return []
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return [t.clone() for t in ts]
    flat = torch.cat([t.reshape(-1).to(torch.float32) for t in ts])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    out, off = [], 0
    for t in ts:
        n = t.numel()
        out.append(flat[off:off + n].view(t.shape).to(t.dtype))
        off += n
    return out


def allreduce_flat_prefix(flat: torch.Tensor, n: int, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """The exchange step of the data-parallel training step: sum the first `n` elements of the flat gradient
    buffer over all ranks IN PLACE with one all-reduce (the loss is a sum, model.py:119: gradients are summed, not averaged).
    Elements beyond `n` (the BatchNorm running statistics of the rank) are left alone.  No-op without a process group."""
    if n < 0 or n > flat.numel() or flat.dim() != 1:
        raise ValueError(f'bad prefix {n} of a buffer with shape {tuple(flat.shape)}')
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1 and n > 0:
        dist.all_reduce(flat[:n], op=dist.ReduceOp.SUM, group=group)
    return flat
