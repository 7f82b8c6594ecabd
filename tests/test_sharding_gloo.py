"""The N>1 host-side logic on CPU: world_size-2 gloo process group, no GPU."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gnn_pe_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)          # same stream on every rank: the full truth
        n_slots = 7
        full = [np.unique(rng.integers(0, 500, size=rng.integers(0, 60))) for _ in range(n_slots)]
        full[3] = np.zeros(0, dtype=np.int64)   # an empty candidate set
        # shard r sees the candidates v with v % world == r, plus some overlap with the other shard
        mine = [np.unique(np.concatenate([c[c % world == rank], c[: len(c) // 5]])) for c in full]
        counts = torch.tensor([len(c) for c in mine], dtype=torch.int32)
        cand = torch.from_numpy(np.concatenate(mine + [np.zeros(0, np.int64)]).astype(np.int32))
        all_counts, all_cand, stride = sharding.allgather_candidates(counts, cand)
        assert all_counts.shape == (world, n_slots) and all_cand.shape == (world, stride)
        merged = sharding.union_reference(all_counts.numpy(), all_cand.numpy())
        ok = all(np.array_equal(m, f) for m, f in zip(merged, full))
        # C1, bitmap form: same sets as bitmaps of 512 bits per slot
        bm = np.zeros((n_slots, 64), dtype=np.uint8)
        for sl, c in enumerate(mine):
            np.bitwise_or.at(bm[sl], c >> 3, (1 << (c & 7)).astype(np.uint8))
        allb = sharding.allgather_bitmaps(torch.from_numpy(bm.reshape(-1)))
        assert allb.shape == (world, n_slots * 64)
        un = sharding.union_bitmaps_reference(allb.numpy()).reshape(n_slots, 64)
        for sl, f in enumerate(full):
            ok = ok and np.array_equal(np.flatnonzero(np.unpackbits(un[sl], bitorder="little")), f)
        # C2: match counts summed over shards, start candidates dealt round-robin
        total = 1001
        mine_n = sharding.start_candidates_of_rank(total, rank, world)
        summed = sharding.allreduce_counts(np.array([mine_n, 2**40 + rank], dtype=np.uint64), "cpu")
        ok = ok and int(summed[0]) == total and int(summed[1]) == world * 2**40 + sum(range(world))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_candidate_allgather_and_count_allreduce_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(results) == [(0, True), (1, True)]


def test_partition_assignment():
    for p, world in [(8, 1), (8, 2), (8, 4), (8, 8), (5, 2), (13, 8)]:
        masks = [sharding.partitions_of_rank(p, r, world) for r in range(world)]
        assert np.array_equal(np.sum(masks, axis=0), np.ones(p, dtype=np.uint8))   # every partition exactly once
    for total in (0, 1, 7, 8, 1001):
        for world in (1, 2, 4, 8):
            assert sum(sharding.start_candidates_of_rank(total, r, world) for r in range(world)) == total
