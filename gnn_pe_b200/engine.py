"""The reference's offline/online pipeline (src/main.cpp:75-182) on one GPU, through the C ABI."""
from __future__ import annotations

import numpy as np

from . import gpe
from .graph_io import CSRGraph, block_membership, degree_order


class Engine:
    """offline(): main.cpp:77-119 (enumeration) + the table the online stage scans.
    online(): main.cpp:122-179 for one query or a batch.  Host steps (embeddings, query plans) run in
    libgpe's host mirror; everything data-parallel runs in the CUDA kernels."""

    def __init__(self, device: int = 0):
        self.ctx = gpe.GpeContext(device)
        self.g = None

    def close(self):
        self.ctx.close()

    def offline(self, g: CSRGraph, l: int = 2, e: int = 2, p: int = 5, sorted_nodes=None, membership=None,
                part_select=None):
        self.g, self.l, self.e, self.p = g, l, e, p
        self.sorted_nodes = degree_order(g) if sorted_nodes is None else np.asarray(sorted_nodes, dtype=np.uint32)
        self.membership = block_membership(g.V, p) if membership is None else np.asarray(membership, dtype=np.uint32)
        self.ctx.set_graph(g.offsets, g.nbrs, g.labels)
        self.x, self.vde = gpe.host_gen_vde(g.offsets, g.nbrs, g.labels, e)
        self.ctx.set_embeddings(self.vde)
        self.n_rows, self.rows_per_partition = self.ctx.enumerate(l + 1, self.sorted_nodes, self.membership, p)
        self.table_rows = self.ctx.build_table(part_select)
        return self.n_rows

    def online(self, q: CSRGraph, limit: int = gpe.LIMIT_MAX) -> int:
        return int(self.ctx.query_batch([q], [limit])[0])

    def online_batch(self, queries, limits=None) -> np.ndarray:
        return self.ctx.query_batch(queries, limits)

    def write_offline_files(self, dataset_dir: str, chunk_rows: int = 1 << 22):
        """all_paths.txt + partition_paths.txt exactly as main.cpp:98-119 writes them."""
        import os
        if self.n_rows >= 1 << 32:
            raise ValueError("the reference's text format holds 32-bit path counts (custom.h:548)")
        base = os.path.join(dataset_dir, "gnn-pe")
        with open(os.path.join(base, "all_paths.txt"), "w") as f:
            f.write(f"{self.n_rows}\n")
            for first in range(0, self.n_rows, chunk_rows):
                rows = self.ctx.dump_paths(first, min(chunk_rows, self.n_rows - first))
                f.write("".join(" ".join(map(str, r)) + " \n" for r in rows.tolist()))
        start = self.ctx.start_rows()
        member_by_rank = self.membership[self.sorted_nodes]
        for i in range(self.p):
            ranks = np.nonzero(member_by_rank == i)[0]
            ids = np.concatenate([np.arange(start[r], start[r + 1], dtype=np.uint64) for r in ranks] +
                                 [np.zeros(0, np.uint64)])
            with open(os.path.join(base, "partitions", f"partition-{i}", "partition_paths.txt"), "w") as f:
                f.write(f"{len(ids)}\n")
                f.write("".join(f"{int(x)}\n" for x in ids))
