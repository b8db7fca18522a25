"""Multi-rank plumbing on CPU: world_size-2 gloo, a deterministic stand-in for the per-batch decode.
(The GPU path itself is covered by the -m gpu tests; this covers partitioning, broadcast and all-gather.)"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from faceformer_b200 import sharding
from faceformer_b200.config import OURS


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def fake_decode(batch: sharding.Batch, num_edges, T):
    """Deterministic function of the batch content, shaped like `predict` [N, F, T]."""
    ne = num_edges[batch.items]
    F = int(ne.max())
    out = np.zeros((len(ne), F, T), np.int64)
    for i, it in enumerate(batch.items):
        out[i] = (np.arange(F)[:, None] * 7 + np.arange(T)[None, :] * 3 + int(it)) % 97
    return out


def _worker(rank, world, port, num_edges, batch_size, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        T = OURS.max_face_length
        blob = torch.arange(1000, dtype=torch.float32) if rank == 0 else torch.zeros(1000)
        sharding.broadcast_weights(blob, 0)
        batches = sharding.plan_batches(num_edges, batch_size, T)
        res, assignment = sharding.run_sharded(lambda b: fake_decode(b, num_edges, T), batches, OURS.num_lines, T)
        q.put((rank, float(blob.sum()), {k: v.tolist() for k, v in res.items()}, assignment))
    finally:
        dist.destroy_process_group()


def test_plan_is_deterministic_and_covers_everything():
    rng = np.random.default_rng(0)
    ne = rng.integers(24, 217, size=507)                       # the synthetic "ours" test split
    batches = sharding.plan_batches(ne, 128, OURS.max_face_length)
    assert [len(b.items) for b in batches] == [128, 128, 128, 123]
    assert np.array_equal(np.concatenate([b.items for b in batches]), np.arange(507))
    for world in (1, 2, 4, 8):
        a = sharding.assign_batches(batches, world)
        assert sorted(sum(a, [])) == list(range(len(batches)))
        assert a == sharding.assign_batches(batches, world)
    srt = sharding.plan_batches(ne, 128, OURS.max_face_length, sort=True)
    assert sorted(np.concatenate([b.items for b in srt]).tolist()) == list(range(507))
    assert sum(b.cost for b in srt) <= sum(b.cost for b in batches) * 1.0001     # cost is per wireframe: the sum is unchanged


def test_assignment_balances_cost():
    rng = np.random.default_rng(1)
    ne = rng.integers(24, 217, size=2048)
    batches = sharding.plan_batches(ne, 32, OURS.max_face_length, sort=True)
    a = sharding.assign_batches(batches, 8)
    load = [sum(batches[i].cost for i in r) for r in a]
    assert max(load) / (sum(load) / 8) < 1.1


@pytest.mark.timeout(120)
def test_two_ranks_gather_equals_single_process():
    rng = np.random.default_rng(2)
    num_edges = rng.integers(3, 40, size=23).astype(np.int64)
    T = OURS.max_face_length
    batch_size = 4
    batches = sharding.plan_batches(num_edges, batch_size, T)
    want = {b.index: fake_decode(b, num_edges, T) for b in batches}
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, num_edges, batch_size, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=100) for _ in procs]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    for rank, blob_sum, res, assignment in got:
        assert blob_sum == float(sum(range(1000)))                                   # broadcast reached every rank
        assert sorted(res) == sorted(want)
        for k, v in res.items():
            assert np.array_equal(np.asarray(v), want[k])                            # gathered == single-process
        assert len(assignment) == 2 and all(len(a) > 0 for a in assignment)


def test_split_batch_is_deterministic_balanced_and_invertible():
    rng = np.random.default_rng(3)
    ne = rng.integers(24, 217, size=128)
    for world in (1, 2, 4, 8):
        parts = sharding.split_batch(ne, world)
        assert sorted(np.concatenate(parts).tolist()) == list(range(128))
        assert all(np.array_equal(a, b) for a, b in zip(parts, sharding.split_batch(ne, world)))
        load = [int((ne[p] + 1).sum()) for p in parts]
        assert max(load) / (sum(load) / world) < 1.03                       # 16 wireframes per rank at world 8: LPT is near-perfect
        T, F = 5, int(ne.max())
        full = rng.integers(0, 50, size=(128, F, T))
        shares = [full[p] for p in parts]
        assert np.array_equal(sharding.merge_split(shares, parts, 128), full)


def _split_worker(rank, world, port, ne, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # the host-side protocol of SplitDecoder with a stand-in for the engine: every rank decodes its share with F fixed globally,
        # the shares are all-gathered (gloo here, NCCL on GPUs) and merged
        parts = sharding.split_batch(ne, world)
        F, T = int(ne.max()), 6
        full = (np.arange(len(ne))[:, None, None] * 1000 + np.arange(F)[None, :, None] * 10 + np.arange(T)[None, None, :]).astype(np.int64)
        cap = max(len(p) for p in parts)
        send = torch.zeros((cap, F, T), dtype=torch.int32)
        send[:len(parts[rank])] = torch.from_numpy(full[parts[rank]].astype(np.int32))
        recv = [torch.empty_like(send) for _ in range(world)]
        dist.all_gather(recv, send)
        merged = sharding.merge_split([r.numpy().astype(np.int64) for r in recv], parts, len(ne))
        handles = [None] * world
        dist.all_gather_object(handles, bytes([rank]) * 64)                 # the IPC-handle exchange of SplitDecoder.connect
        q.put((rank, bool(np.array_equal(merged, full)), [h[0] for h in handles]))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_ranks_split_one_batch_and_merge():
    ne = np.random.default_rng(4).integers(3, 30, size=9).astype(np.int64)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_split_worker, args=(r, 2, port, ne, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=100) for _ in procs]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    for rank, same, hs in got:
        assert same and hs == [0, 1]
