"""Wireframe-level data parallelism: one process per GPU, whole batches per rank.

The reference has no distributed code.  Wireframes are independent
units of ``forward_eval`` -- the only cross-sample couplings are per BATCH: F = max(num_input)
(model_para.py:187) and the global stop predicate (model_para.py:232).  To stay tensor-identical
to a single-GPU run the batch composition is therefore fixed FIRST and whole batches are dealt to
ranks; there is no collective on the data path.  Collectives (NCCL over NVLink on GPUs, gloo in
the CPU tests):

  * ``broadcast_weights``  one broadcast of the packed fp32 weight blob from rank 0
  * ``gather_predictions`` one all_gather of fixed-shape int32 prediction shards at the end
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, List, Sequence

import numpy as np


@dataclass(frozen=True)
class Batch:
    index: int                 # position in the global batch order
    items: np.ndarray          # dataset indices of the wireframes in this batch (in order)
    cost: float                # relative decode cost estimate


def batch_cost(num_edges: np.ndarray, seq_len: int) -> float:
    """Relative cost of decoding a batch in parallel mode: every real anchor is one sequence
    re-running its whole prefix each step (sum_P P token passes), each token pass dominated by the
    fixed-size linear layers plus cross-attention over n_i+4 memory rows."""
    ne = np.asarray(num_edges, dtype=np.float64)
    passes = seq_len * (seq_len - 1) / 2.0
    return float(np.sum((ne + 1.0) * passes * (1.0 + (ne + 4.0) / 2560.0)))


def plan_batches(num_edges: Sequence[int], batch_size: int, seq_len: int, sort: bool = False) -> List[Batch]:
    """Fix the global batch composition.  ``sort=False`` keeps dataset order (what a DataLoader
    with shuffle=False yields, trainer.py:52-54); ``sort=True`` buckets wireframes of similar edge
    count together so that little work is spent on padded anchors."""
    ne = np.asarray(num_edges, dtype=np.int64)
    order = np.argsort(-ne, kind="stable") if sort else np.arange(len(ne))
    out = []
    for b, s in enumerate(range(0, len(ne), batch_size)):
        items = order[s:s + batch_size]
        out.append(Batch(index=b, items=items, cost=batch_cost(ne[items], seq_len)))
    return out


def assign_batches(batches: Sequence[Batch], world_size: int) -> List[List[int]]:
    """Deal whole batches to ranks, longest-processing-time first.  Deterministic: every rank
    computes the same assignment.  Returns per-rank lists of batch indices (ascending)."""
    load = [0.0] * world_size
    mine: List[List[int]] = [[] for _ in range(world_size)]
    for b in sorted(batches, key=lambda x: (-x.cost, x.index)):
        r = min(range(world_size), key=lambda i: (load[i], i))
        load[r] += b.cost
        mine[r].append(b.index)
    return [sorted(m) for m in mine]


def broadcast_weights(blob, src: int = 0):
    """Broadcast the packed weight blob (torch tensor, in place) from ``src`` to all ranks."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(blob, src=src)
    return blob


def gather_predictions(local: Dict[int, np.ndarray], batches: Sequence[Batch], assignment: List[List[int]],
                       num_lines: int, seq_len: int, device=None) -> Dict[int, np.ndarray]:
    """All-gather the per-batch ``predict`` tensors of every rank.

    local: batch index -> int64 [N_b, F_b, T] decoded on this rank.  Every rank contributes one
    fixed-shape int32 buffer [max_batches_per_rank, max_N, num_lines, T] (padded with -1) plus the
    (N_b, F_b) sizes, so a single all_gather moves everything.  Returns batch index -> predict
    for ALL batches, identical on every rank."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    if world == 1:
        return dict(local)
    max_b = max(len(a) for a in assignment)
    max_n = max(len(b.items) for b in batches)
    dev = device if device is not None else "cpu"
    buf = torch.full((max_b, max_n, num_lines, seq_len), -1, dtype=torch.int32)
    dims = torch.zeros((max_b, 2), dtype=torch.int32)
    for slot, bi in enumerate(assignment[rank]):
        p = np.asarray(local[bi])
        buf[slot, :p.shape[0], :p.shape[1]] = torch.from_numpy(p.astype(np.int32))
        dims[slot, 0], dims[slot, 1] = p.shape[0], p.shape[1]
    buf, dims = buf.to(dev), dims.to(dev)
    all_buf = [torch.empty_like(buf) for _ in range(world)]
    all_dims = [torch.empty_like(dims) for _ in range(world)]
    dist.all_gather(all_buf, buf)
    dist.all_gather(all_dims, dims)
    out: Dict[int, np.ndarray] = {}
    for r in range(world):
        rb, rd = all_buf[r].cpu().numpy(), all_dims[r].cpu().numpy()
        for slot, bi in enumerate(assignment[r]):
            n, f = int(rd[slot, 0]), int(rd[slot, 1])
            out[bi] = rb[slot, :n, :f].astype(np.int64)
    return out


def run_sharded(decode_batch: Callable[[Batch], np.ndarray], batches: Sequence[Batch],
                num_lines: int, seq_len: int, device=None, gather: bool = True):
    """Decode this rank's batches with ``decode_batch`` and all-gather the results."""
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    assignment = assign_batches(batches, world)
    local = {bi: decode_batch(batches[bi]) for bi in assignment[rank]}
    if not gather:
        return local, assignment
    return gather_predictions(local, batches, assignment, num_lines, seq_len, device), assignment
