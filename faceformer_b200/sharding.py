"""Wireframe-level data parallelism: one process per GPU.

Two partitionings, both tensor-identical to a single-GPU run:
  * whole batches per rank (``plan_batches`` / ``assign_batches`` / ``run_sharded``): no coupling between ranks at all;
  * ONE batch split over all ranks (``split_batch`` / ``SplitDecoder``; BASELINE.json configs[2], batch=128 over 8 GPUs): the two
    per-batch couplings are carried across ranks -- F is fixed on the host (FFB_OPT_FORCE_F) and the per-step stop predicate crosses
    NVLink inside the step-end kernel (ffb_stop_exchange_*), so a test split of 507 wireframes at batch 128 keeps all 8 GPUs busy.

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


# ---- one batch split over all ranks -------------------------------------------------------------------------------------------------
def split_batch(num_input: Sequence[int], world_size: int) -> List[np.ndarray]:
    """Deal the wireframes of ONE batch to ranks, longest-processing-time first on the number of distinct sequences (n_i real anchors
    + the shared padded-anchor sequence), ties to the lower rank.  Deterministic; returns per-rank ascending index arrays (a rank may
    get none when the batch has fewer wireframes than ranks)."""
    ne = np.asarray(num_input, dtype=np.int64)
    load = [0] * world_size
    mine: List[List[int]] = [[] for _ in range(world_size)]
    for i in sorted(range(len(ne)), key=lambda j: (-int(ne[j]), j)):
        r = min(range(world_size), key=lambda k: (load[k], k))
        load[r] += int(ne[i]) + 1
        mine[r].append(i)
    return [np.asarray(sorted(m), dtype=np.int64) for m in mine]


def merge_split(shares: Sequence[np.ndarray], parts: Sequence[np.ndarray], n: int) -> np.ndarray:
    """Inverse of split_batch on the outputs: shares[r] = predict [len(parts[r]), F, T] of rank r -> predict [n, F, T]."""
    ref = next(s for s in shares if s.shape[0] > 0)
    out = np.zeros((n,) + tuple(ref.shape[1:]), dtype=ref.dtype)
    for idx, sh in zip(parts, shares):
        if len(idx):
            out[idx] = sh[:len(idx)]
    return out


class SplitDecoder:
    """model(batch) for one global batch spread over all ranks of the default process group (one Engine per rank)."""

    def __init__(self, engine, group=None):
        import torch.distributed as dist
        self.eng = engine
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.group = group
        self.connected = False

    def connect(self):
        """Exchange the CUDA IPC handles of the ranks' stop-flag buffers (the only set-up collective)."""
        import torch.distributed as dist
        if self.world == 1 or self.connected:
            return
        mine = self.eng.stop_exchange_export()
        handles = [None] * self.world
        dist.all_gather_object(handles, mine, group=self.group)
        self.eng.stop_exchange_connect(self.rank, self.world, handles)
        self.connected = True

    def forward_eval(self, coords, pad_mask, num_input, gather: bool = True):
        """coords / pad_mask / num_input: the WHOLE batch (CUDA tensors, identical on every rank).  Returns (predict [N,F,T] int64 CUDA
        tensor -- complete on every rank when gather, else only this rank's rows filled --, executed steps)."""
        import torch
        import torch.distributed as dist
        from .lib import FFB_OPT_FORCE_F
        ni = num_input.cpu().numpy() if hasattr(num_input, "cpu") else np.asarray(num_input)
        n, F = len(ni), int(ni.max())
        T = self.eng.cfg.seq_len(self.eng.mode)
        parts = split_batch(ni, self.world)
        if min(len(p) for p in parts) == 0:
            raise ValueError("batch has fewer wireframes than ranks")
        idx = torch.as_tensor(parts[self.rank], device=coords.device)
        self.eng.set_option(FFB_OPT_FORCE_F, F if self.world > 1 else 0)
        local, steps = self.eng.forward_eval(coords.index_select(0, idx), pad_mask.index_select(0, idx), num_input.index_select(0, idx))
        out = torch.zeros((n, F, T), dtype=torch.int64, device=coords.device)
        if self.world == 1 or not gather:
            out[idx] = local
            return out, steps
        cap = max(len(p) for p in parts)
        send = torch.zeros((cap, F, T), dtype=torch.int32, device=coords.device)
        send[:len(idx)] = local.to(torch.int32)
        recv = torch.empty((self.world, cap, F, T), dtype=torch.int32, device=coords.device)
        dist.all_gather_into_tensor(recv, send, group=self.group)          # predicted face loops of every rank (NCCL over NVLink)
        for r, p in enumerate(parts):
            out[torch.as_tensor(p, device=coords.device)] = recv[r, :len(p)].to(torch.int64)
        return out, steps
