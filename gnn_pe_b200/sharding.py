"""Multi-GPU plumbing: one process per GPU, torch.distributed for the exchange step.

The reference is single-process; its only parallelism is one OpenMP thread per graph partition whose
std::set results are merged serially (src/main.cpp:160-172).  Here GPU `r` of `world` owns the partitions
`i % world == r` (a path belongs to the partition of its FIRST vertex, custom.h:74), scans only its own
path table, and the serial merge becomes
  C1: an all-gather of the shards' candidate bitmaps (fixed size: one bit per vertex of the slot's label class,
      identical layout on every shard, so there is no count exchange and no host sync), whose union is fused
      into the compaction's popcount pass (gpe_batch_bitmap_merge).  The list form of the same exchange
      (allgather_candidates + gpe_batch_cand_merge: counts, then lists padded to the longest shard) is kept
      for callers that hold sorted lists;
  C2: an all-reduce(sum) of the per-query match counts after the join has been split by start candidate.
The CSR, labels and vertex embeddings are replicated, so the join needs no other exchange.
These helpers are backend-agnostic (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def partitions_of_rank(p: int, rank: int, world: int) -> np.ndarray:
    """uint8 selection mask over the reference's partitions for one GPU."""
    return np.array([1 if i % world == rank else 0 for i in range(p)], dtype=np.uint8)


def start_candidates_of_rank(total: int, rank: int, world: int) -> int:
    """How many of `total` start candidates (dealt round-robin by index) a rank enumerates."""
    return (total - rank + world - 1) // world if total > rank else 0


def allgather_candidates(counts: torch.Tensor, cand: torch.Tensor, group=None):
    """C1.  counts: int32 [n_slots] sizes of this shard's lists; cand: int32 [>= counts.sum()] the lists
    concatenated.  Returns (all_counts [world, n_slots], all_cand [world, stride], stride)."""
    world = dist.get_world_size(group)
    total = torch.tensor([int(counts.sum().item()) if counts.numel() else 0], dtype=torch.int64, device=counts.device)
    totals = torch.empty(world, dtype=torch.int64, device=counts.device)
    dist.all_gather_into_tensor(totals, total, group=group)
    stride = max(int(totals.max().item()), 1)
    padded = torch.zeros(stride, dtype=torch.int32, device=cand.device)
    n = int(total.item())
    if n:
        padded[:n] = cand[:n]
    all_counts = torch.empty(world * counts.numel(), dtype=torch.int32, device=counts.device)
    all_cand = torch.empty(world * stride, dtype=torch.int32, device=cand.device)
    dist.all_gather_into_tensor(all_counts, counts.contiguous(), group=group)
    dist.all_gather_into_tensor(all_cand, padded, group=group)
    return all_counts.view(world, -1), all_cand.view(world, stride), stride


class _DevMem:
    """A span of device memory owned by libgpe, exposed through __cuda_array_interface__ (zero-copy torch view)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = dict(shape=(nbytes,), typestr="|u1", data=(ptr, False), version=3, strides=None)


def allgather_bitmaps(local: torch.Tensor, group=None) -> torch.Tensor:
    """C1, bitmap form.  local: uint8 [n_bytes] (this shard's candidate bitmaps).  Returns uint8 [world, n_bytes]."""
    world = dist.get_world_size(group)
    out = torch.empty(world * local.numel(), dtype=torch.uint8, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out.view(world, local.numel())


def union_bitmaps_reference(all_bitmaps: np.ndarray) -> np.ndarray:
    """Host statement of what gpe_batch_bitmap_merge computes: the OR of the shards' bitmaps."""
    return np.bitwise_or.reduce(all_bitmaps, axis=0)


def union_reference(all_counts: np.ndarray, all_cand: np.ndarray):
    """Host statement of what gpe_batch_cand_merge computes: per slot, the sorted union of the shards' lists."""
    world, n_slots = all_counts.shape
    out = []
    offs = np.zeros((world, n_slots + 1), dtype=np.int64)
    offs[:, 1:] = np.cumsum(all_counts, axis=1)
    for s in range(n_slots):
        parts = [all_cand[r, offs[r, s]:offs[r, s + 1]] for r in range(world)]
        out.append(np.unique(np.concatenate(parts)) if parts else np.zeros(0, np.int32))
    return out


def allreduce_counts(raw: np.ndarray, device, group=None) -> np.ndarray:
    """C2.  Sum of per-query match counts over the shards (u64 carried as int64)."""
    t = torch.from_numpy(raw.astype(np.int64)).to(device)
    dist.all_reduce(t, group=group)
    return t.cpu().numpy().astype(np.uint64)


class ShardedEngine:
    """The online stage over a path table sharded across the ranks of the default process group."""

    def __init__(self, ctx, rank: int, world: int):
        self.ctx, self.rank, self.world = ctx, rank, world

    def build(self, g, l, e, p, sorted_nodes, membership, vde):
        self.ctx.set_graph(g.offsets, g.nbrs, g.labels)
        self.ctx.set_embeddings(vde)
        n_rows, rows_pp = self.ctx.enumerate(l + 1, sorted_nodes, membership, p)
        sel = partitions_of_rank(p, self.rank, self.world)
        table_rows = self.ctx.build_table(sel if self.world > 1 else None)
        return n_rows, rows_pp, table_rows

    def exchange(self):
        """C1 on the context's stream: NCCL orders the all-gather after the scan and the merge after the all-gather,
        no host sync in between."""
        ptr, nbytes = self.ctx.batch_bitmap()
        stream = torch.cuda.ExternalStream(self.ctx.stream)
        with torch.cuda.stream(stream):
            if nbytes == 0:
                return self.ctx.batch_bitmap_merge(1, ptr)
            local = torch.as_tensor(_DevMem(ptr, nbytes), device="cuda")
            self._all = allgather_bitmaps(local)  # kept alive until the next exchange
            self.ctx.batch_bitmap_merge(self.world, self._all.data_ptr())

    def exchange_lists(self):
        """The list form of C1 (sorted candidate lists, variable length)."""
        n_slots, total = self.ctx.batch_cand_info()
        counts = torch.empty(max(n_slots, 1), dtype=torch.int32, device="cuda")
        cand = torch.zeros(max(total, 1), dtype=torch.int32, device="cuda")
        self.ctx.batch_cand_export(counts.data_ptr(), cand.data_ptr())
        all_counts, all_cand, stride = allgather_candidates(counts[:n_slots], cand[:total] if total else cand[:0])
        torch.cuda.current_stream().synchronize()
        self.ctx.batch_cand_merge(self.world, all_counts.data_ptr(), all_cand.data_ptr(), stride)

    def step(self, lists: bool = False):
        """scan (local shard) -> C1 -> join (this rank's start candidates).  Batch must be uploaded."""
        if self.world == 1:
            self.ctx.batch_filter()
        elif lists:
            self.ctx.batch_filter()
            self.exchange_lists()
        else:
            self.ctx.batch_scan()
            self.exchange()
        self.ctx.batch_join(self.rank, self.world)

    def finish(self, limits):
        raw = self.ctx.batch_download()
        if self.world > 1:
            raw = allreduce_counts(raw, "cuda")
        return np.array([self.ctx.clamp(r, l) for r, l in zip(raw, limits)], dtype=np.uint64)
