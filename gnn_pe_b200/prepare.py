"""The preparation step the reference does in gnnpe.py (GNN-PE/gnnpe.py:44-76, GNN-PGE/gnnpge.py): create
``<dataset>/gnn-pe/partitions/partition-i/`` and write ``gnn-pe/membership.txt`` -- one ``vertex partition`` line per
vertex, in ascending-degree order (stable, :71-76), which is the order the offline stage enumerates in.

    python -m gnn_pe_b200.prepare --f DIR/ --d data.graph --p 5 [--variant pe|pge] [--partitioner auto|metis|rcm|block]

Same flags as the reference's script (``--l`` is accepted and unused there too).  The data graph is the ``.graph`` text file
the C++ side reads (the reference's script reads a networkx pickle of the same graph; given one, and networkx installed, it
is read as well).  Partitioners: ``metis`` = pymetis.part_graph(p, recursive=True) as the reference calls it (:66); ``rcm`` =
equal cuts of a reverse Cuthill-McKee order (locality without METIS); ``block`` = contiguous id blocks; ``auto`` = metis
when pymetis imports, else rcm.  Path tables, candidate sets and answers do not depend on the assignment:
it only decides which rows share a partition, i.e. a GPU in the sharded mode.
"""
from __future__ import annotations

import argparse
import os
import shutil

import numpy as np

from . import graph_io


def _read_any(path: str) -> graph_io.CSRGraph:
    with open(path, "rb") as f:
        magic = f.read(2)
    if magic == b"\x1f\x8b" or magic[:1] == b"\x80":   # gzip or a bare pickle (the reference's Test/*.gpickle.gz is the latter)
        import gzip
        import pickle
        with (gzip.open(path, "rb") if magic == b"\x1f\x8b" else open(path, "rb")) as f:
            G = pickle.load(f)                          # a networkx graph with integer nodes 0..V-1 (gnnpe.py:52-53)
        nodes = sorted(G.nodes())
        if nodes != list(range(len(nodes))):
            raise ValueError("the pickled graph's nodes are not 0..V-1")
        labels = np.array([G.nodes[v].get("label", 0) for v in nodes], dtype=np.uint32)   # only the topology is used here
        edges = np.array([(u, v) for u, v in G.edges() if u != v], dtype=np.int64).reshape(-1, 2)
        return graph_io.csr_from_edges(len(nodes), edges, labels)
    return graph_io.read_graph(path)


def partition(g: graph_io.CSRGraph, p: int, how: str = "auto") -> np.ndarray:
    """membership[v] in [0, p) for every vertex."""
    V = g.V
    if p <= 0:
        raise ValueError("partition number must be positive")
    if how == "auto":
        try:
            import pymetis  # noqa: F401
            how = "metis"
        except ImportError:
            how = "rcm"
    if how == "metis":
        import pymetis
        off = g.offsets.astype(np.int64)
        adjacency = [g.nbrs[off[v]:off[v + 1]] for v in range(V)]
        _, membership = pymetis.part_graph(p, adjacency=adjacency, recursive=True)   # gnnpe.py:66
        return np.asarray(membership, dtype=np.uint32)
    if how == "block" or V == 0:
        return graph_io.block_membership(V, p)
    if how == "rcm":
        from scipy.sparse import csr_matrix
        from scipy.sparse.csgraph import reverse_cuthill_mckee
        a = csr_matrix((np.ones(len(g.nbrs), dtype=np.int8), g.nbrs.astype(np.int32), g.offsets.astype(np.int32)), shape=(V, V))
        order = reverse_cuthill_mckee(a, symmetric_mode=True)
        membership = np.zeros(V, dtype=np.uint32)
        membership[order] = graph_io.block_membership(V, p)
        return membership
    raise ValueError(f"unknown partitioner {how!r}")


def edge_cut(g: graph_io.CSRGraph, membership: np.ndarray) -> int:
    src = np.repeat(np.arange(g.V), np.diff(g.offsets.astype(np.int64)))
    return int((membership[src] != membership[g.nbrs]).sum() // 2)


def prepare(dataset_dir: str, g: graph_io.CSRGraph, p: int, variant: str = "pe", how: str = "auto") -> np.ndarray:
    if not dataset_dir.endswith("/"):
        dataset_dir += "/"
    base = dataset_dir + ("gnn-pe" if variant == "pe" else "gnn-pge")
    shutil.rmtree(base, ignore_errors=True)           # gnnpe.py:57: a fresh tree every time
    for i in range(p):
        os.makedirs(f"{base}/partitions/partition-{i}")
    membership = partition(g, p, how)
    graph_io.write_membership(f"{base}/membership.txt", graph_io.degree_order(g), membership)
    return membership


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--f", type=str, default="../Test/", help="file path")
    ap.add_argument("--d", type=str, default="../Test/data_graph.graph", help="data graph")
    ap.add_argument("--p", type=int, default=5, help="partition number")
    ap.add_argument("--l", type=int, default=2, help="path length (unused, as in the reference)")
    ap.add_argument("--variant", choices=["pe", "pge"], default="pe")
    ap.add_argument("--partitioner", choices=["auto", "metis", "rcm", "block"], default="auto")
    a = ap.parse_args(argv)
    g = _read_any(a.d)
    membership = prepare(a.f, g, a.p, a.variant, a.partitioner)
    sizes = np.bincount(membership, minlength=a.p)
    print(f"|V|: {g.V}, |E|: {g.E}, partitions: {a.p}, sizes {sizes.min()}..{sizes.max()}, edge cut {edge_cut(g, membership)}")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
